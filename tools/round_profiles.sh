#!/bin/bash
# One GPU-box pass for a round: bench (both arms), the ncu launch list of the bench command, and --set full captures of
# the dominant kernels.  Run as: gpurun -- "bash tools/round_profiles.sh"; outputs land in gpurun_out/.
set -x
cd $GRAFT_REPO_ROOT
O=gpurun_out; mkdir -p $O
timeout 900 python bench.py --impl reference > $O/${TAG:-r02j}_bench_reference.json 2> $O/${TAG:-r02j}_bench_reference.err
timeout 900 python bench.py > $O/${TAG:-r02j}_bench.json 2> $O/${TAG:-r02j}_bench.err
timeout 600 ncu --metrics gpu__time_duration.sum --clock-control none -c 400 --csv --log-file $O/${TAG:-r02j}_launches.csv python bench.py --steps 2 --warmup 1 --no-cpu-baseline > $O/${TAG:-r02j}_launches.log 2>&1
timeout 600 ncu --set full --clock-control none --import-source on -k regex:"mcc_band_kernel|unstru_kernel" -c 3 -o $O/${TAG:-r02j}_band python tools/perf_probe.py --workload mica_ompa --num 1000 --reps 1 > $O/${TAG:-r02j}_ncu_band.log 2>&1
timeout 800 ncu --set full --clock-control none --import-source on -k regex:mcc_persistent -c 1 -o $O/${TAG:-r02j}_general python tools/perf_probe.py --workload synthetic --num 148 --reps 1 > $O/${TAG:-r02j}_ncu_general.log 2>&1
for f in $O/${TAG:-r02j}_bench.json $O/${TAG:-r02j}_bench_reference.json; do tail -n 2 $f; done
if [ -f ractip_b200/libractip_prob_tune.so ]; then
  RP_PROFILE=1 RP_LIB=ractip_b200/libractip_prob_tune.so timeout 300 python tools/perf_probe.py --workload mica_ompa --num 1000 --reps 1 > $O/${TAG:-r02j}_phase_profile.txt 2>&1
  RP_PROFILE=1 RP_LIB=ractip_b200/libractip_prob_tune.so timeout 300 python tools/perf_probe.py --workload synthetic --num 148 --reps 1 > $O/${TAG:-r02j}_phase_profile_synthetic.txt 2>&1
fi
timeout 300 python tests/bundled_latency.py > $O/${TAG:-r02j}_bundled_latency.jsonl 2> $O/${TAG:-r02j}_bundled_latency.err
timeout 900 python -m pytest tests -x -q -m gpu 2>&1 | tail -n 3
timeout 300 python -c "import __graft_entry__ as g; g.smoke(); print(\"smoke ok\")" 2>&1 | tail -n 2
