#!/usr/bin/env python3
"""BASELINE configs[1] / configs[2]: the bundled pairs one at a time and as one ragged batch.

Prints, per pair, the CUDA-event time of the kernels (device-resident batch), the end-to-end time of
rp_run_dense (host buffers in and out) and the oracle port's single-thread time for the same pair
(the reference is single-threaded, src/ractip.cpp:1494).  Measurement aid that lives under tests/ because it
times the oracle next to the GPU path (not collected by pytest); the product path never touches the oracle.

    python tests/bundled_latency.py [--reps 20]
"""
import argparse
import json
import sys
import time
from pathlib import Path

ROOT = Path(__file__).resolve().parent.parent
sys.path.insert(0, str(ROOT))

from oracle.oracle import Oracle  # noqa: E402
from ractip_b200 import ProbabilityStage, default_model, default_opts  # noqa: E402

ap = argparse.ArgumentParser()
ap.add_argument("--reps", type=int, default=20)
args = ap.parse_args()
g = json.loads((ROOT / "tests" / "golden" / "bundled_pairs.json").read_text())
seqs, names = g["sequences"], g["pairs"]
model, opts = default_model(), default_opts()
st = ProbabilityStage(model)
orc = Oracle(model)


def gpu_times(pairs):
    b = st.batch(pairs, opts)
    for _ in range(3):
        b.run(); b.sync()
    ks = []
    for _ in range(args.reps):
        b.run(); b.sync()
        ks.append(st.last_timing().ms_total)
    b.close()
    for _ in range(3):
        st.run_dense(pairs, opts)
    t0 = time.perf_counter()
    for _ in range(args.reps):
        st.run_dense(pairs, opts)
    e2e = (time.perf_counter() - t0) / args.reps * 1e3
    return min(ks), sorted(ks)[len(ks) // 2], e2e


def cpu_time(pairs):
    t0 = time.perf_counter()
    for s1, s2 in pairs:
        orc.rnafold(s1, opts.max_w); orc.rnafold(s2, opts.max_w); orc.rnaduplex(s1, s2, opts.th_hy)
    return (time.perf_counter() - t0) * 1e3


rows = []
allp = []
for a, b in names:
    p = [(seqs[a], seqs[b])]
    allp += p
    kmin, kmed, e2e = gpu_times(p)
    rows.append({"pair": f"{a} x {b}", "n1": len(seqs[a]), "n2": len(seqs[b]), "kernel_ms_min": kmin, "kernel_ms_median": kmed,
                 "e2e_ms": e2e, "cpu_port_1thread_ms": cpu_time(p)})
kmin, kmed, e2e = gpu_times(allp)
rows.append({"pair": "all 8 bundled pairs, one ragged batch", "kernel_ms_min": kmin, "kernel_ms_median": kmed, "e2e_ms": e2e,
             "cpu_port_1thread_ms": cpu_time(allp)})
for r in rows:
    print(json.dumps(r))
st.close()
