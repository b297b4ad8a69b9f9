"""ctypes wrapper of oracle/liboracle.so (the CPU fp64 restatement).

TEST INFRASTRUCTURE ONLY: imported by tests/, __graft_entry__.smoke() and
bench.py's cpu_baseline / --impl reference legs.  Never imported by ractip_b200.
"""
from __future__ import annotations

import ctypes as C
import subprocess
from pathlib import Path

import numpy as np

HERE = Path(__file__).resolve().parent
LIB_PATH = HERE / "liboracle.so"
REF_USHUFFLE = HERE / "_ref" / "libushuffle_ref.so"

_lib = None


def build() -> None:
    """Compile the oracle (and oracle/_ref when /root/reference is present)."""
    subprocess.run(["make", "-C", str(HERE)], check=True, stdout=subprocess.PIPE, stderr=subprocess.STDOUT)


def load() -> C.CDLL:
    global _lib
    if _lib is not None:
        return _lib
    if not LIB_PATH.exists():
        build()
    lib = C.CDLL(str(LIB_PATH))
    vp, i, d = C.c_void_p, C.c_int, C.c_double
    lib.orc_params_new.restype = vp
    lib.orc_params_new.argtypes = [vp]
    lib.orc_params_free.argtypes = [vp]
    lib.orc_pf_scale.restype = d
    lib.orc_pf_scale.argtypes = [vp]
    lib.orc_fold.restype = i
    lib.orc_fold.argtypes = [vp, C.c_char_p, i, i, vp, vp, i, vp]
    lib.orc_enumerate.restype = i
    lib.orc_enumerate.argtypes = [vp, C.c_char_p, i, i, vp, vp, i, vp]
    lib.orc_pf_duplex.restype = i
    lib.orc_pf_duplex.argtypes = [vp, C.c_char_p, i, C.c_char_p, i, vp, vp]
    lib.orc_enum_duplex.restype = i
    lib.orc_enum_duplex.argtypes = [vp, C.c_char_p, i, C.c_char_p, i, vp, vp]
    lib.orc_rnafold.restype = i
    lib.orc_rnafold.argtypes = [vp, C.c_char_p, i, i, vp, vp]
    lib.orc_rnaduplex.restype = i
    lib.orc_rnaduplex.argtypes = [vp, C.c_char_p, i, C.c_char_p, i, C.c_float, i, vp]
    _lib = lib
    return lib


def _ptr(a):
    return a.ctypes.data_as(C.c_void_p) if a is not None else None


class Oracle:
    """CPU oracle bound to one integer energy model (an RpModel ctypes struct)."""

    def __init__(self, model):
        self.lib = load()
        self._model = model  # keep alive
        self.P = self.lib.orc_params_new(C.addressof(model))
        if not self.P:
            raise ValueError("oracle: bad model (temperature must be 37)")

    def __del__(self):
        try:
            if getattr(self, "P", None):
                self.lib.orc_params_free(self.P)
                self.P = None
        except Exception:
            pass

    @property
    def pf_scale(self) -> float:
        return self.lib.orc_pf_scale(self.P)

    def fold(self, seq: str, cp: int = 0, max_w: int = 0):
        """(pr[(n+1),(n+1)], up[n,max_w] or None, logZ) in fp64."""
        n = len(seq)
        pr = np.zeros((n + 1, n + 1))
        up = np.zeros((n, max_w)) if max_w > 0 else None
        lz = C.c_double()
        rc = self.lib.orc_fold(self.P, seq.encode(), n, cp, _ptr(pr), _ptr(up), max_w, C.byref(lz))
        if rc:
            raise RuntimeError("orc_fold rc=%d" % rc)
        return pr, up, lz.value

    def enumerate(self, seq: str, cp: int = 0, max_w: int = 0):
        n = len(seq)
        pr = np.zeros((n + 1, n + 1))
        up = np.zeros((n, max(max_w, 1)))
        lz = C.c_double()
        rc = self.lib.orc_enumerate(self.P, seq.encode(), n, cp, _ptr(pr), _ptr(up), max_w, C.byref(lz))
        if rc:
            raise RuntimeError("orc_enumerate rc=%d" % rc)
        return pr, (up if max_w > 0 else None), lz.value

    def pf_duplex(self, s1: str, s2: str):
        pr = np.zeros((len(s1) + 1, len(s2) + 1))
        es = C.c_double()
        rc = self.lib.orc_pf_duplex(self.P, s1.encode(), len(s1), s2.encode(), len(s2), _ptr(pr), C.byref(es))
        if rc:
            raise RuntimeError("orc_pf_duplex rc=%d" % rc)
        return pr, es.value

    def enum_duplex(self, s1: str, s2: str):
        pr = np.zeros((len(s1) + 1, len(s2) + 1))
        lz = C.c_double()
        rc = self.lib.orc_enum_duplex(self.P, s1.encode(), len(s1), s2.encode(), len(s2), _ptr(pr), C.byref(lz))
        if rc:
            raise RuntimeError("orc_enum_duplex rc=%d" % rc)
        return pr, lz.value

    def rnafold(self, seq: str, max_w: int):
        """Reference layouts: bp float32[(L+1)(L+2)/2], up float32[L,max_w]."""
        L = len(seq)
        bp = np.zeros((L + 1) * (L + 2) // 2, dtype=np.float32)
        up = np.zeros((L, max_w), dtype=np.float32)
        rc = self.lib.orc_rnafold(self.P, seq.encode(), L, max_w, _ptr(bp), _ptr(up))
        if rc:
            raise RuntimeError("orc_rnafold rc=%d" % rc)
        return bp, up

    def rnaduplex(self, s1: str, s2: str, th_hy: float = 0.1, use_pf_duplex: bool = False):
        hp = np.zeros((len(s1) + 1, len(s2) + 1), dtype=np.float32)
        rc = self.lib.orc_rnaduplex(self.P, s1.encode(), len(s1), s2.encode(), len(s2),
                                    C.c_float(th_hy), int(use_pf_duplex), _ptr(hp))
        if rc:
            raise RuntimeError("orc_rnaduplex rc=%d" % rc)
        return hp


def ref_ushuffle(seq: str, k: int, seed: int, num: int):
    """Shuffles from the reference's own src/ushuffle.c (oracle/_ref), seeded via libc srandom."""
    if not REF_USHUFFLE.exists():
        return None
    ref = C.CDLL(str(REF_USHUFFLE))
    libc = C.CDLL(None)
    libc.srandom(C.c_uint(seed))
    ref.set_randfunc(C.cast(libc.random, C.c_void_p))
    out = []
    buf = C.create_string_buffer(len(seq) + 1)
    for _ in range(num):
        ref.shuffle(seq.encode(), buf, len(seq), k)
        out.append(buf.raw[:len(seq)].decode())
    return out
