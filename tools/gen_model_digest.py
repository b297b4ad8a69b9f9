#!/usr/bin/env python3
"""Write tests/golden/model_digest.txt = rp_model_digest(rp_model_default(use_bl=1))."""
import ctypes as C
import sys
from pathlib import Path
ROOT = Path(__file__).resolve().parent.parent
sys.path.insert(0, str(ROOT))
from ractip_b200 import _lib, default_model  # noqa: E402
lib = _lib.load()
d = lib.rp_model_digest(C.byref(default_model()))
(ROOT / "tests" / "golden" / "model_digest.txt").write_text(str(d) + "\n")
print(d)
