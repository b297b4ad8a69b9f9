#!/usr/bin/env python3
"""Small mixed batch (both band launch shapes + the general kernel + --duplex) for compute-sanitizer runs."""
import json
import sys
from pathlib import Path

ROOT = Path(__file__).resolve().parent.parent
sys.path.insert(0, str(ROOT))
import numpy as np  # noqa: E402

from ractip_b200 import ProbabilityStage, default_opts  # noqa: E402

seqs = json.loads((ROOT / "tests" / "golden" / "bundled_pairs.json").read_text())["sequences"]
rng = np.random.default_rng(5)
rnd = lambda n: "".join("ACGU"[x] for x in rng.integers(0, 4, n))
pairs = [(seqs["DIS"], seqs["DIS"]), (seqs["MicA"], seqs["ompA"]), (rnd(9), rnd(4)), (rnd(120), rnd(101))]
st = ProbabilityStage()
r = st.run_dense(pairs, default_opts())
r2 = st.run_dense(pairs[:2], default_opts(use_pf_duplex=1))
print("ok", [float(x.hp.max()) for x in r], float(r2[0].hp.max()))
st.close()
