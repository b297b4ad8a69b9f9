import ctypes as C
import json
import subprocess
import sys
from pathlib import Path

import numpy as np
import pytest

ROOT = Path(__file__).resolve().parent.parent
sys.path.insert(0, str(ROOT))


def pytest_configure(config):
    config.addinivalue_line("markers", "gpu: needs a CUDA device (run on the B200 box)")


def _ensure_built():
    from ractip_b200 import build as rb
    rb.build()
    from oracle import oracle as orc
    if not orc.LIB_PATH.exists():
        orc.build()


@pytest.fixture(scope="session")
def lib():
    _ensure_built()
    from ractip_b200 import _lib
    return _lib.load()


@pytest.fixture(scope="session")
def model(lib):
    from ractip_b200 import default_model
    return default_model()


@pytest.fixture(scope="session")
def oracle(model):
    from oracle.oracle import Oracle
    return Oracle(model)


@pytest.fixture(scope="session")
def bundled():
    return json.loads((ROOT / "tests" / "golden" / "bundled_pairs.json").read_text())


@pytest.fixture(scope="session")
def emul(model):
    """Host emulation of the kernel phase functions (tests/emul)."""
    d = ROOT / "tests" / "emul"
    subprocess.run(["make", "-C", str(d)], check=True, stdout=subprocess.PIPE, stderr=subprocess.STDOUT)
    em = C.CDLL(str(d / "libemul.so"))
    em.emul_problem.argtypes = [C.c_void_p, C.c_char_p, C.c_int, C.c_int, C.c_int, C.c_int, C.c_int, C.c_int,
                                C.c_float, C.c_int, C.c_void_p, C.c_void_p, C.c_void_p, C.c_void_p]
    em.emul_problem.restype = C.c_int
    em.emul_band_problem.argtypes = em.emul_problem.argtypes
    em.emul_band_problem.restype = C.c_int
    em.emul_wide_problem.argtypes = [C.c_int] + em.emul_problem.argtypes
    em.emul_wide_problem.restype = C.c_int

    class Emul:
        @staticmethod
        def _fn(band, wide):
            """band: shared-memory band schedule; wide=W: general schedule with W-diagonal split-sum bands"""
            if wide:
                return lambda *a: em.emul_wide_problem(wide, *a)
            return em.emul_band_problem if band else em.emul_problem

        def linear(self, s, w, T=64, band=False, wide=0):
            n = len(s)
            bp = np.zeros((n + 1) * (n + 2) // 2, dtype=np.float32)
            up = np.zeros((n, w), dtype=np.float32)
            lz = C.c_double()
            fn = self._fn(band, wide)
            rc = fn(C.addressof(model), s.encode(), n, 0, 0, w, 0, 0, 0.0, T, bp.ctypes.data,
                                 up.ctypes.data, None, C.byref(lz))
            assert rc == 0
            return bp, up, lz.value

        def cofold(self, s1, s2, th=0.1, T=64, band=False, wide=0):
            n1, n2 = len(s1), len(s2)
            hp = np.zeros((n1 + 1, n2 + 1), dtype=np.float32)
            lz = C.c_double()
            fn = self._fn(band, wide)
            rc = fn(C.addressof(model), (s1 + s2).encode(), n1 + n2, n1 + 1, 1, 0, n1, n2, th, T,
                                 None, None, hp.ctypes.data, C.byref(lz))
            assert rc == 0
            return hp, lz.value

    return Emul()


def has_gpu():
    try:
        import torch
        return torch.cuda.is_available()
    except Exception:
        return False


@pytest.fixture(scope="session")
def stage(lib):
    from ractip_b200 import ProbabilityStage
    st = ProbabilityStage()
    yield st
    st.close()


def rand_seq(rng, n, alphabet="ACGU"):
    return "".join(alphabet[x] for x in rng.integers(0, len(alphabet), n))
