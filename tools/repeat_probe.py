#!/usr/bin/env python3
"""Bitwise repeatability of the multi-CTA (cluster) route: the same two long pairs four times."""
import sys, numpy as np
from pathlib import Path
sys.path.insert(0, str(Path(__file__).resolve().parent.parent))
from ractip_b200 import ProbabilityStage, default_opts
rng = np.random.default_rng(5)
rs = lambda n: "".join("ACGU"[x] for x in rng.integers(0, 4, n))
pairs = [(rs(700), rs(350)), (rs(300), rs(260))]
st = ProbabilityStage()
ref = None
for k in range(4):
    r = st.run_dense(pairs, default_opts())
    cat = np.concatenate([np.concatenate([x.bp1, x.bp2, x.up1.ravel(), x.up2.ravel(), x.hp.ravel()]) for x in r])
    if ref is None: ref = cat
    print(k, "identical" if np.array_equal(ref, cat) else "DIFFERENT", float(cat.sum()))
st.close()
