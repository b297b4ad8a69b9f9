#!/usr/bin/env python3
"""Wall-clock breakdown of the one-shot host call (rp_run_dense = create + run + fetch + destroy)
on the 1000-shuffle MicA x ompA batch.  Tuning aid for the e2e number of bench.py."""
import ctypes as C
import json
import sys
import time
from pathlib import Path

ROOT = Path(__file__).resolve().parent.parent
sys.path.insert(0, str(ROOT))

from ractip_b200 import ProbabilityStage, default_opts, zscore_shuffles  # noqa: E402
from ractip_b200.stage import _make_pairs  # noqa: E402

seqs = json.loads((ROOT / "tests" / "golden" / "bundled_pairs.json").read_text())["sequences"]
r1, r2 = zscore_shuffles(seqs["MicA"], seqs["ompA"], 1000, 1)
pairs = list(zip(r1, r2))
st = ProbabilityStage()
lib = st.lib
opts = default_opts()
arr, keep = _make_pairs(pairs)
n = len(pairs)
b0 = st.batch(pairs, opts)
tot = b0.total_floats
b0.close()
lib.rp_host_alloc.restype = C.c_void_p
pin = lib.rp_host_alloc(C.c_size_t(tot * 4))
for rep in range(4):
    t0 = time.perf_counter()
    h = C.c_void_p()
    st._check(lib.rp_batch_create(st.ctx, arr, n, C.byref(opts), C.byref(h)))
    t1 = time.perf_counter()
    st._check(lib.rp_batch_run(h))
    st._check(lib.rp_batch_sync(h))
    t2 = time.perf_counter()
    st._check(lib.rp_batch_fetch_dense(h, C.c_void_p(pin), C.c_size_t(tot)))
    t3 = time.perf_counter()
    lib.rp_batch_destroy(h)
    t4 = time.perf_counter()
    print(f"rep {rep}: create {1e3*(t1-t0):.2f} ms  run+sync {1e3*(t2-t1):.2f} ms  fetch({tot*4/1e6:.0f} MB) {1e3*(t3-t2):.2f} ms  "
          f"destroy {1e3*(t4-t3):.2f} ms  total {1e3*(t4-t0):.2f} ms", flush=True)
st.close()
