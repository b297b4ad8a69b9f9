#!/usr/bin/env python3
"""Kernel-time probe: runs the device-resident batch a few times and prints the CUDA-event
time of each launch.  Used under ncu (short batches) and for A/B timing of kernel changes.

    python tools/perf_probe.py [--workload mica_ompa|synthetic|dis|copa] [--num N] [--reps R]
"""
import argparse
import json
import sys
from pathlib import Path

ROOT = Path(__file__).resolve().parent.parent
sys.path.insert(0, str(ROOT))

import numpy as np  # noqa: E402

from ractip_b200 import ProbabilityStage, default_opts, zscore_shuffles  # noqa: E402

ap = argparse.ArgumentParser()
ap.add_argument("--workload", default="mica_ompa")
ap.add_argument("--num", type=int, default=1000)
ap.add_argument("--reps", type=int, default=3)
ap.add_argument("--only", default="", help="l1,l2 lengths override: synthetic random pair of these lengths")
args = ap.parse_args()
seqs = json.loads((ROOT / "tests" / "golden" / "bundled_pairs.json").read_text())["sequences"]
if args.only:
    l1, l2 = [int(x) for x in args.only.split(",")]
    rng = np.random.default_rng(1)
    s1 = "".join("ACGU"[x] for x in rng.integers(0, 4, l1))
    s2 = "".join("ACGU"[x] for x in rng.integers(0, 4, l2))
elif args.workload == "mica_ompa":
    s1, s2 = seqs["MicA"], seqs["ompA"]
elif args.workload == "dis":
    s1, s2 = seqs["DIS"], seqs["DIS"]
elif args.workload == "copa":
    s1, s2 = seqs["CopA"], seqs["CopT"]
else:
    rng = np.random.default_rng(20261018)
    s1 = "".join("ACGU"[x] for x in rng.integers(0, 4, 1000))
    s2 = "".join("ACGU"[x] for x in rng.integers(0, 4, 500))
r1, r2 = zscore_shuffles(s1, s2, args.num, 1)
st = ProbabilityStage()
b = st.batch(list(zip(r1, r2)), default_opts())
for r in range(args.reps):
    b.run()
    b.sync()
    t = st.last_timing()
    print(f"{args.workload} n1={len(s1)} n2={len(s2)} num={args.num} rep={r} kernel_ms={t.ms_total:.3f} "
          f"TFLOPs={t.alg_flops / t.ms_total / 1e9:.3f} pairs/s={args.num / t.ms_total * 1e3:.1f}", flush=True)
b.close()
st.close()
