"""Pins the CPU oracle: its DP (inside, outside, unpaired windows, two-strand variant,
pf_duplex) must agree with exhaustive enumeration of every structure under the same loop
model (the reference has no golden vectors for this path)."""
import ctypes as C

import numpy as np
import pytest

from conftest import rand_seq

TOL = 1e-12


@pytest.mark.parametrize("alphabet", ["ACGU", "GCGCAU", "GGGCCCAU"])
def test_linear_dp_equals_enumeration(oracle, alphabet):
    rng = np.random.default_rng(hash(alphabet) % 1000)
    for _ in range(25):
        n = int(rng.integers(5, 18))
        s = rand_seq(rng, n, alphabet)
        pr, up, lz = oracle.fold(s, 0, 6)
        pe, ue, le = oracle.enumerate(s, 0, 6)
        assert abs(lz - le) < TOL, s
        assert np.abs(pr - pe).max() < TOL, s
        assert np.abs(up - ue).max() < TOL, s


def test_linear_multiloops_present(oracle):
    # long enough, GC-rich: multiloop states carry real weight, windows up to 15
    rng = np.random.default_rng(3)
    for _ in range(4):
        s = rand_seq(rng, 20, "GCGCGA")
        pr, up, lz = oracle.fold(s, 0, 15)
        pe, ue, le = oracle.enumerate(s, 0, 15)
        assert abs(lz - le) < TOL and np.abs(pr - pe).max() < TOL and np.abs(up - ue).max() < TOL


def test_two_strand_dp_equals_enumeration(oracle):
    rng = np.random.default_rng(11)
    for _ in range(60):
        n = int(rng.integers(6, 18))
        s = rand_seq(rng, n, "GCGCAU")
        cp = int(rng.integers(2, n + 1))
        pr, _, lz = oracle.fold(s, cp, 0)
        pe, _, le = oracle.enumerate(s, cp, 0)
        assert abs(lz - le) < TOL, (s, cp)
        assert np.abs(pr - pe).max() < TOL, (s, cp)


def test_small_maxloop_bound_is_exercised(model):
    """With MAXLOOP lowered to 3 the interior-loop size bound actually cuts structures at n<=16."""
    from oracle.oracle import Oracle
    o = Oracle(model)
    o.lib.orc_set_maxloop.argtypes = [C.c_void_p, C.c_int]
    o.lib.orc_set_maxloop(o.P, 3)
    rng = np.random.default_rng(5)
    for _ in range(30):
        n = int(rng.integers(10, 17))
        s = rand_seq(rng, n, "GCGCAU")
        cp = int(rng.integers(0, n + 1))
        cp = 0 if cp < 2 else cp
        pr, up, lz = o.fold(s, cp, 5 if cp == 0 else 0)
        pe, ue, le = o.enumerate(s, cp, 5 if cp == 0 else 0)
        assert abs(lz - le) < TOL and np.abs(pr - pe).max() < TOL
        if cp == 0:
            assert np.abs(up - ue).max() < TOL


def test_pf_duplex_equals_enumeration(oracle):
    """orc_pf_duplex restates src/pf_duplex.c:128-206 in log space; brute force over all duplexes."""
    rng = np.random.default_rng(2)
    for _ in range(30):
        s1 = rand_seq(rng, int(rng.integers(2, 9)), "GCGCAU")
        s2 = rand_seq(rng, int(rng.integers(2, 9)), "GCGCAU")
        pr, es = oracle.pf_duplex(s1, s2)
        pe, le = oracle.enum_duplex(s1, s2)
        if np.isfinite(le):
            assert abs(es - le) < 1e-10 and np.abs(pr - pe).max() < 1e-10


def test_model_invariants_on_bundled_sequences(oracle, bundled):
    for name in ["DIS", "MicA", "CopA", "Tar"]:
        s = bundled["sequences"][name]
        n = len(s)
        pr, up, lz = oracle.fold(s, 0, 15)
        assert np.isfinite(lz)
        prs = pr + pr.T
        # every base is paired or unpaired: sum_j p(i,j) + P(i unpaired) = 1
        assert np.abs(prs[1:, 1:].sum(axis=1) + up[:, 0] - 1.0).max() < 1e-10
        # longer windows are never more likely; windows past the 3' end are 0
        assert (np.diff(up, axis=1) <= 1e-12).all()
        for i in range(n):
            for d in range(15):
                if i + d >= n:
                    assert up[i, d] == 0.0
        assert (pr >= -1e-15).all() and (pr <= 1 + 1e-12).all()


def test_readme_dis_known_answer(oracle, bundled):
    """The only known-answer in the reference (README.md:91-97, `ractip DIS.fa DIS.fa`):
    ((((.(((((((..[[[[[[.)))))))...))))  on both strands.  The IP maximises the sum of
    (p - threshold) over compatible pairs, so every pair of that answer must be a variable
    (p > th_ss = 0.5 inside, p > th_hy = 0.1 between) -- and here the thresholded sets are
    exactly the answer's helices, which pins the energy model end to end."""
    s = bundled["sequences"]["DIS"]
    n = len(s)
    ans = bundled["readme_dis"]["s1"]
    assert ans == bundled["readme_dis"]["s2"].replace("]", "[")
    stack, want_in = [], set()
    for k, ch in enumerate(ans, start=1):
        if ch == "(":
            stack.append(k)
        elif ch == ")":
            want_in.add((stack.pop(), k))
    kiss = [k for k, ch in enumerate(ans, start=1) if ch == "["]
    want_ext = {(a, b) for a, b in zip(kiss, reversed(kiss))}  # antiparallel, non-crossing
    pr, _, _ = oracle.fold(s, 0, 0)
    got_in = {(int(i), int(j)) for i, j in np.argwhere(pr > 0.5)}
    assert got_in == want_in
    pc, _, _ = oracle.fold(s + s, n + 1, 0)
    inter = pc[1:n + 1, n + 1:]
    for a, b in want_ext:
        assert inter[a - 1, b - 1] > 0.9


def test_rnaduplex_layout_and_threshold(oracle, bundled):
    s1, s2 = bundled["sequences"]["DIS"], bundled["sequences"]["DIS"]
    hp = oracle.rnaduplex(s1, s2, 0.1)
    assert hp.shape == (36, 36) and (hp[0] == 0).all() and (hp[:, 0] == 0).all()
    nz = hp[hp != 0]
    assert nz.size > 0 and (nz > np.float32(0.1)).all()
    dense = oracle.rnaduplex(s1, s2, 0.0)
    assert np.array_equal(np.where(dense > np.float32(0.1), dense, 0), hp)
