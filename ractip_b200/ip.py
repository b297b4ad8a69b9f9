"""Host-side consumer of the probability stage: RactIP's integer programme and energy evaluation.

Mirrors what RactIP::solve does after the three probability calls (reference src/ractip.cpp:552-1352),
RactIP::solve_ss (:1366-1465), RactIP::energy_of_duplex (:1528-1559) and the z-score statistic of
RactIP::run (:1626-1669).  The model itself is built in C++ (ractip_b200/csrc/ipmodel.cpp, C ABI in
include/ractip_ip.h) column for column and row for row as the reference builds it; the MIP solver is
scipy.optimize.milp (HiGHS) because none of the solver libraries the reference links behind
src/ip.cpp (GLPK, Gurobi, CPLEX, SCIP, HiGHS C++) exists in this image.  Host code, no GPU.
"""
from __future__ import annotations

import ctypes as C
import math
from dataclasses import dataclass
from typing import List, Optional, Sequence, Tuple

import numpy as np

from . import _lib
from ._lib import RpIpOpts, RpModel
from .stage import PairProbabilities, PairRecords, RpError

FR, LO, UP, DB, FX = range(5)


def default_ip_opts(**kw) -> RpIpOpts:
    lib = _lib.load()
    o = RpIpOpts()
    lib.rp_ip_opts_default(C.byref(o))
    for k, v in kw.items():
        if not hasattr(o, k):
            raise TypeError(f"unknown option {k}")
        setattr(o, k, v)
    return o


def _fp(a: Optional[np.ndarray]):
    if a is None:
        return None
    a = np.ascontiguousarray(a, dtype=np.float32)
    return a


class IPModel:
    """A model built by rp_ip_build / rp_ip_build_sparse / rp_ip_build_ss."""

    def __init__(self, handle: C.c_void_p, n1: int, n2: int):
        self.lib = _lib.load()
        self.handle, self.n1, self.n2 = handle, n1, n2
        nc, nr, nz = C.c_int(), C.c_int(), C.c_int()
        self._check(self.lib.rp_ip_dims(handle, C.byref(nc), C.byref(nr), C.byref(nz)))
        self.n_cols, self.n_rows, self.nnz = nc.value, nr.value, nz.value
        self.obj = np.zeros(self.n_cols)
        self.row_kind = np.zeros(max(self.n_rows, 1), dtype=np.int32)
        self.row_lo = np.zeros(max(self.n_rows, 1))
        self.row_hi = np.zeros(max(self.n_rows, 1))
        self.ia = np.zeros(max(self.nnz, 1), dtype=np.int32)
        self.ja = np.zeros(max(self.nnz, 1), dtype=np.int32)
        self.ar = np.zeros(max(self.nnz, 1))
        self._check(self.lib.rp_ip_export(handle, self.obj.ctypes.data, self.row_kind.ctypes.data,
                                          self.row_lo.ctypes.data, self.row_hi.ctypes.data,
                                          self.ia.ctypes.data, self.ja.ctypes.data, self.ar.ctypes.data))

    def _check(self, rc: int):
        if rc:
            raise RpError(rc, self.lib.rp_strerror(rc).decode())

    # ---------------------------------------------------------------- construction
    @classmethod
    def joint(cls, probs: PairProbabilities, opts: Optional[RpIpOpts] = None) -> "IPModel":
        """RactIP::solve's model (src/ractip.cpp:552-1225) from the dense matrices."""
        lib = _lib.load()
        opts = opts if opts is not None else default_ip_opts()
        n1, n2 = probs.hp.shape[0] - 1, probs.hp.shape[1] - 1
        bp1, bp2, up1, up2, hp = (_fp(probs.bp1), _fp(probs.bp2), _fp(probs.up1), _fp(probs.up2), _fp(probs.hp))
        h = C.c_void_p()
        rc = lib.rp_ip_build(C.byref(opts), n1, n2, bp1.ctypes.data, bp2.ctypes.data, up1.ctypes.data,
                             up2.ctypes.data, hp.ctypes.data, C.byref(h))
        if rc:
            raise RpError(rc, lib.rp_strerror(rc).decode())
        return cls(h, n1, n2)

    @classmethod
    def joint_sparse(cls, n1: int, n2: int, recs: PairRecords, opts: Optional[RpIpOpts] = None) -> "IPModel":
        """The same model from the thresholded lists of the sparse path (x, y, z, v, w)."""
        lib = _lib.load()
        opts = opts if opts is not None else default_ip_opts()
        arrs = [np.ascontiguousarray(a) for a in (recs.x, recs.y, recs.z, recs.v, recs.w)]
        h = C.c_void_p()
        args = []
        for a in arrs:
            args += [C.c_void_p(a.ctypes.data if a.size else 0), int(a.size)]
        rc = lib.rp_ip_build_sparse(C.byref(opts), n1, n2, *args, C.byref(h))
        if rc:
            raise RpError(rc, lib.rp_strerror(rc).decode())
        return cls(h, n1, n2)

    @classmethod
    def single(cls, bp: np.ndarray, n: int, opts: Optional[RpIpOpts] = None,
               usable: Optional[np.ndarray] = None) -> "IPModel":
        """RactIP::solve_ss's model (src/ractip.cpp:1366-1436)."""
        lib = _lib.load()
        opts = opts if opts is not None else default_ip_opts()
        bpf = _fp(bp)
        u = None if usable is None else np.ascontiguousarray(usable, dtype=np.uint8)
        h = C.c_void_p()
        rc = lib.rp_ip_build_ss(C.byref(opts), n, bpf.ctypes.data, C.c_void_p(u.ctypes.data if u is not None else 0),
                                C.byref(h))
        if rc:
            raise RpError(rc, lib.rp_strerror(rc).decode())
        return cls(h, n, 0)

    # --------------------------------------------------------------------- solving
    def solve(self) -> Tuple[float, np.ndarray]:
        """IP::solve (src/ip.cpp:110-122): maximise; every column is binary.  Returns (objective, values)."""
        from scipy.optimize import Bounds, LinearConstraint, milp
        from scipy.sparse import coo_matrix
        if self.n_cols == 0:
            return 0.0, np.zeros(0)
        cons = []
        if self.n_rows:
            A = coo_matrix((self.ar[:self.nnz], (self.ia[:self.nnz], self.ja[:self.nnz])),
                           shape=(self.n_rows, self.n_cols)).tocsr()
            lb = np.full(self.n_rows, -np.inf)
            ub = np.full(self.n_rows, np.inf)
            k = self.row_kind[:self.n_rows]
            lo, hi = self.row_lo[:self.n_rows], self.row_hi[:self.n_rows]
            lb[(k == LO) | (k == DB)] = lo[(k == LO) | (k == DB)]
            ub[(k == UP) | (k == DB)] = hi[(k == UP) | (k == DB)]
            lb[k == FX] = lo[k == FX]
            ub[k == FX] = lo[k == FX]
            cons = [LinearConstraint(A, lb, ub)]
        res = milp(c=-self.obj, constraints=cons, integrality=np.ones(self.n_cols),
                   bounds=Bounds(np.zeros(self.n_cols), np.ones(self.n_cols)),
                   options={"mip_rel_gap": 0.0})
        if res.x is None:
            raise RuntimeError(f"MIP solver failed: {res.message}")
        return float(-res.fun), np.asarray(res.x)

    def decode(self, values: np.ndarray):
        """Column values -> (r1, r2, used1, used2) as src/ractip.cpp:1227-1250,1286-1295."""
        v = np.ascontiguousarray(values, dtype=np.float64)
        r1 = C.create_string_buffer(self.n1 + 1)
        r2 = C.create_string_buffer(self.n2 + 1)
        u1 = np.zeros(max(self.n1, 1), dtype=np.uint8)
        u2 = np.zeros(max(self.n2, 1), dtype=np.uint8)
        self._check(self.lib.rp_ip_decode(self.handle, v.ctypes.data, r1, r2, u1.ctypes.data, u2.ctypes.data))
        return r1.value.decode(), r2.value.decode(), u1[:self.n1], u2[:self.n2]

    def close(self):
        if self.handle:
            self.lib.rp_ip_free(self.handle)
            self.handle = None

    def __del__(self):
        try:
            self.close()
        except Exception:
            pass


# ------------------------------------------------------------------------- energies
def energy_of_structure(model: RpModel, seq: str, structure: str, cut_point: int = -1) -> float:
    """Vienna::energy_of_structure as RactIP calls it (src/ractip.cpp:1254,1299,1457,1551)."""
    lib = _lib.load()
    e = C.c_float()
    rc = lib.rp_energy_of_structure(C.byref(model), seq.encode(), structure.encode(), len(seq), cut_point, C.byref(e))
    if rc:
        raise RpError(rc, lib.rp_strerror(rc).decode())
    return e.value


def energy_of_duplex(model: RpModel, s1: str, s2: str, r1: str, r2: str) -> float:
    """RactIP::energy_of_duplex (src/ractip.cpp:1528-1559)."""
    lib = _lib.load()
    e = C.c_float()
    rc = lib.rp_energy_of_duplex(C.byref(model), s1.encode(), len(s1), s2.encode(), len(s2), r1.encode(), r2.encode(),
                                 C.byref(e))
    if rc:
        raise RpError(rc, lib.rp_strerror(rc).decode())
    return e.value


# --------------------------------------------------------------- RactIP::solve / run
@dataclass
class JointPrediction:
    r1: str
    r2: str
    objective: float
    e1: Optional[float] = None   # energy of s1's internal structure (or of its accessible regions with --acc-max)
    e2: Optional[float] = None
    e3: Optional[float] = None   # energy_of_duplex


def solve_ss(model: RpModel, seq: str, bp: np.ndarray, opts: Optional[RpIpOpts] = None,
             usable: Optional[np.ndarray] = None, r0: Optional[str] = None, energy: bool = False):
    """RactIP::solve_ss (src/ractip.cpp:1354-1465): (objective, structure[, energy])."""
    ip = IPModel.single(bp, len(seq), opts, usable)
    ea, vals = ip.solve()
    r, _, _, _ = ip.decode(vals)
    if r0 is not None:   # the --acc-max-ss branch draws into the string that already holds the brackets
        r = "".join(c if c != "." else k for c, k in zip(r0, r))
    ip.close()
    if energy:
        return ea, r, energy_of_structure(model, seq, r, -1)
    return ea, r


def solve_joint(model: RpModel, s1: str, s2: str, probs: PairProbabilities, opts: Optional[RpIpOpts] = None,
                energies: bool = False, recs: Optional[PairRecords] = None) -> JointPrediction:
    """RactIP::solve after the probability calls (src/ractip.cpp:552-1352).  With `recs` the model is built
    from the thresholded lists of the sparse path instead of the dense matrices."""
    opts = opts if opts is not None else default_ip_opts()
    ip = IPModel.joint_sparse(len(s1), len(s2), recs, opts) if recs is not None else IPModel.joint(probs, opts)
    ea, vals = ip.solve()
    r1, r2, used1, used2 = ip.decode(vals)
    out = JointPrediction(r1, r2, ea)
    if energies:
        out.e3 = energy_of_duplex(model, s1, s2, r1, r2)
    kT = (model.temperature + 273.15) * 1.98717 / 1000.0
    for which, (s, bp, up, used) in enumerate(((s1, probs.bp1, probs.up1, used1), (s2, probs.bp2, probs.up2, used2))):
        r = out.r1 if which == 0 else out.r2
        e = None
        if not opts.acc_max:
            if energies:
                e = energy_of_structure(model, s, r, -1)
        else:
            if opts.acc_max_ss:   # :1263-1271
                usable = (1 - used).astype(np.uint8)
                add, r = solve_ss(model, s, bp, opts, usable, r0=r)
                out.objective += add
            if energies:          # :1272-1283: -kT ln(up) of the chosen regions
                e = 0.0
                n = len(s)
                i = 0
                while i < n:
                    if used[i]:
                        j = i
                        while j + 1 < n and used[j + 1]:
                            j += 1
                        e += -math.log(float(up[i][j - i])) * kT
                        i = j + 1
                    else:
                        i += 1
        if which == 0:
            out.r1, out.e1 = r, e
        else:
            out.r2, out.e2 = r, e
    ip.close()
    return out


def zscore_statistic(e_joint: float, e_ss: float, shuffled: Sequence[Tuple[float, float]]) -> Tuple[float, float]:
    """The two numbers `ractip --zscore` prints (src/ractip.cpp:1626-1669), in float as the reference:
    e_joint = e1+e2+e3 of the real pair, e_ss = e1s+e2s; shuffled = [(ee, ee1s+ee2s)] per shuffle."""
    f = np.float32
    s = s2 = ss = ss2 = f(0)
    for ee, ees_ss in shuffled:
        ee = f(ee)
        ees = f(ee - f(ees_ss))
        s = f(s + ee); s2 = f(s2 + f(ee * ee))
        ss = f(ss + ees); ss2 = f(ss2 + f(ees * ees))
    n = f(len(shuffled))
    m = f(s / n); v = f(f(s2 / n) - f(m * m))
    v = max(v, f(0))
    ms = f(ss / n); vs = f(f(ss2 / n) - f(ms * ms))
    vs = max(vs, f(0))
    with np.errstate(divide="ignore", invalid="ignore"):
        return float(f(f(e_joint) - m) / np.sqrt(v)), float(f(f(e_joint) - f(e_ss) - ms) / np.sqrt(vs))
