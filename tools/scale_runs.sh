#!/bin/bash
# bench.py at N = 1, 2, 4, 8 on one multi-GPU box (EXTRA=--no-config4 for configs[3] only) (gpurun --gpus 8 -- "bash tools/scale_runs.sh"), plus the rp_multi_* shim test
cd $GRAFT_REPO_ROOT
O=gpurun_out; mkdir -p $O; T=${TAG:-r02}
timeout 600 python bench.py --no-cpu-baseline $EXTRA > $O/${T}_bench_1gpu_samebox.json 2> $O/${T}_scale_1.err
for N in 2 4 8; do
  timeout 900 python -m torch.distributed.run --nnodes=1 --nproc-per-node $N --master-addr 127.0.0.1 --master-port $((29500 + N)) bench.py --gpus $N --no-cpu-baseline $EXTRA > $O/${T}_bench_${N}gpu.json 2> $O/${T}_scale_$N.err
done
for N in 1 2 4 8; do f=$O/${T}_bench_${N}gpu.json; [ $N = 1 ] && f=$O/${T}_bench_1gpu_samebox.json; python - $f <<'PY'
import json,sys
try:
    j=json.loads(open(sys.argv[1]).read().strip().splitlines()[-1])
    print(j["n_gpus"], round(j["value"]), j["ms_per_step"], round(j["e2e"]["value"]), j.get("config4",{}).get("value"))
except Exception as e: print(sys.argv[1], "failed", e)
PY
done
timeout 600 python -m pytest tests/test_cpp_shim.py -x -q -m gpu 2>&1 | tail -n 3
