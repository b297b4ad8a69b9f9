"""The drop-in boundary: the shared library loads, exports every symbol the header declares,
plans layouts on the host, and refuses to compute without a GPU (no CPU fallback)."""
import ctypes as C
import re
from pathlib import Path

import numpy as np
import pytest

from conftest import has_gpu

ROOT = Path(__file__).resolve().parent.parent


def test_library_exports_every_declared_symbol(lib):
    declared = set()
    for h in sorted((ROOT / "include").glob("*.h")):   # ractip_prob.h (probability stage), ractip_ip.h (consumer side)
        hdr = re.sub(r"/\*.*?\*/", " ", h.read_text(), flags=re.S)
        declared |= set(re.findall(r"\b(rp_[a-z0-9_]+)\s*\(", hdr))
    from ractip_b200 import _lib
    assert declared == set(_lib.EXPORTS), declared ^ set(_lib.EXPORTS)
    for name in declared:
        assert hasattr(lib, name), name
    assert b"sm_100a" in lib.rp_version()


def test_struct_sizes_match_the_header(lib):
    from ractip_b200._lib import RpModel, RpDenseLayout, RpRec, RpSparseCounts
    # rp_model: 4-byte ints and 8-byte doubles with natural alignment
    ints = 64 + 31 * 3 + 200 * 6 + 40 * 2 + 1600 + 8000 + 40000 + 7 + 200 + 40 + 200
    assert C.sizeof(RpModel) >= ints * 4 + 1401 + 241 + 1801
    assert C.sizeof(RpRec) == 12 and C.sizeof(RpSparseCounts) == 24 and C.sizeof(RpDenseLayout) == 80


def test_dense_and_sparse_plans(lib):
    from ractip_b200._lib import RpPair, RpDenseLayout, RpSparseLayout
    from ractip_b200 import default_opts
    pairs = (RpPair * 2)()
    pairs[0].s1, pairs[0].n1, pairs[0].s2, pairs[0].n2 = b"ACGUACGUAC", 10, b"GGGAAACCC", 9
    pairs[1].s1, pairs[1].n1, pairs[1].s2, pairs[1].n2 = b"A" * 72, 72, b"C" * 137, 137
    o = default_opts()
    lay = (RpDenseLayout * 2)()
    tot = C.c_size_t()
    assert lib.rp_dense_plan(pairs, 2, C.byref(o), lay, C.byref(tot)) == 0
    # sizes of src/ractip.cpp:314 (bp), :370 (up), :393/:405 (hp)
    assert (lay[0].n_bp1, lay[0].n_bp2, lay[0].n_up1, lay[0].n_up2, lay[0].n_hp) == (66, 55, 150, 135, 110)
    assert (lay[1].n_bp1, lay[1].n_bp2, lay[1].n_hp) == (2701, 9591, 73 * 138)
    assert lay[1].bp1 == 66 + 55 + 150 + 135 + 110
    assert tot.value == lay[1].hp + lay[1].n_hp
    sl = (RpSparseLayout * 2)()
    tr, tf = C.c_size_t(), C.c_size_t()
    assert lib.rp_sparse_plan(pairs, 2, C.byref(o), sl, C.byref(tr), C.byref(tf)) == 0
    assert sl[1].cap_x == 73 and sl[1].cap_y == 138 and sl[1].cap_z == 72 * 10
    assert sl[1].cap_v == 72 * 11 and sl[1].cap_w == 137 * 11 and sl[1].w == sl[1].v + sl[1].cap_v   # lengths 5..15 (:622)
    assert tf.value == (10 + 9 + 72 + 137) * 15
    # bad arguments are reported, not crashed on
    assert lib.rp_dense_plan(None, 1, C.byref(o), lay, C.byref(tot)) == 1
    pairs[0].n1 = 0
    assert lib.rp_dense_plan(pairs, 2, C.byref(o), lay, C.byref(tot)) == 1


def test_alg_flops_match_survey_table(lib):
    # F_mcc in GFLOP
    for n, g in [(35, 0.0003), (70, 0.0035), (72, 0.0039), (137, 0.0226), (209, 0.0648), (1000, 3.067), (1500, 8.831)]:
        f = lib.rp_alg_flops_mcc(n) / 1e9
        assert abs(f - g) <= max(0.00006, 0.002 * g), (n, f, g)


@pytest.mark.skipif(has_gpu(), reason="only meaningful on a box without a GPU")
def test_compute_fails_loudly_without_gpu(lib, model):
    from ractip_b200 import ProbabilityStage, RpError
    ctx = C.c_void_p()
    rc = lib.rp_create(C.byref(ctx), C.byref(model), 0)
    assert rc == 2 and not ctx.value  # RP_ERR_NO_DEVICE
    assert b"no CPU fallback" in lib.rp_last_error(None)
    with pytest.raises(RpError):
        ProbabilityStage(model)


def test_product_does_not_reference_the_oracle():
    """The product path must never import, link or call anything under oracle/."""
    pkg = ROOT / "ractip_b200"
    for p in list(pkg.rglob("*.py")) + list(pkg.rglob("*.cpp")) + list(pkg.rglob("*.cu")) + list(pkg.rglob("*.h")):
        text = p.read_text()
        assert "liboracle" not in text and "rp_oracle" not in text and "from oracle" not in text \
            and "import oracle" not in text and "orc_" not in text, p


def test_kernel_plan_routes_lengths_by_shared_memory(lib):
    """Host-side routing: short problems run two band CTAs per SM, the MicA x ompA two-strand
    problem (209 nt) must fit the one-CTA band kernel on a B200, long ones fall to the general kernel."""
    smem = C.c_size_t()
    assert lib.rp_kernel_plan(72, 0, C.byref(smem)) == 1 and smem.value <= (232448 + 1024) // 2 - 1024
    assert lib.rp_kernel_plan(137, 0, C.byref(smem)) == 0
    assert lib.rp_kernel_plan(209, 0, C.byref(smem)) == 0 and smem.value <= 232448
    assert lib.rp_kernel_plan(215, 0, C.byref(smem)) == 0 and smem.value <= 232448
    assert lib.rp_kernel_plan(223, 0, C.byref(smem)) == 0 and smem.value <= 232448 - 64   # OxyS x fhlA (222) still fits
    assert lib.rp_kernel_plan(224, 0, None) == 2
    assert lib.rp_kernel_plan(1500, 0, None) == 3 and lib.rp_kernel_plan(699, 0, None) == 2 and lib.rp_kernel_plan(700, 0, None) == 3
    # monotone in n, and every band answer respects the limit it was given
    prev = 1
    for n in range(1, 400):
        k = lib.rp_kernel_plan(n, 0, C.byref(smem))
        assert k in (0, 1, 2) and (k >= prev or (prev == 1 and k == 0)), n
        if k != 2:
            assert smem.value <= 232448
        prev = k
    assert lib.rp_kernel_plan(72, 48 * 1024, None) == 2   # a device without the opt-in carve-out
