// mcc_driver.h -- the phase schedule of one problem, written once against an
// "executor" that runs a per-thread phase function on every thread of the CTA
// and then synchronises them.  kernels.cu instantiates it with a CTA executor
// (threadIdx.x + __syncthreads); tests/emul with a serial one.
#ifndef RP_MCC_DRIVER_H
#define RP_MCC_DRIVER_H

#include "mcc_core.h"

namespace rp {

// sh: CTA-shared scratch (see Shared).  dense: base of the dense
// float output.  logz: 3 doubles per pair (s1, s2, s1&s2) or nullptr.
template <class Exec>
RP_HD void solve_mcc(Exec& ex, Ctx& c, const Problem& p, float* dense, double* logz, const Shared& sh) {
  const int T = sh.T;
  const int n = c.n;
  if (n + 2 <= RP_SMEM_SEQ) {
    ex.phase(0, [&](int tid) { stage_sequence(c, sh, tid); });
    c.S = sh.S;
  }
  ex.phase(1, [&](int tid) { prologue(c, sh, tid); });
  ex.phase(2, [&](int tid) { prologue2(c, sh, tid); });

  // ---- inside: anti-diagonal wavefront, shortest spans first
  for (int d = TURN + 1; d <= n - 1; d++) {
    const int cells = n - d;
    const int chunk = make_split(cells, T).Cp;
    for (int i0 = 1; i0 <= cells; i0 += chunk) {
      const int C = cells - i0 + 1 < chunk ? cells - i0 + 1 : chunk;
      ex.phase(3, [&](int tid) { inside_A(c, sh, d, i0, C, tid); });
      ex.phase(4, [&](int tid) { inside_B(c, sh, d, i0, C, tid); });
    }
  }
  inside_end(c);
  if (logz) {
    ex.phase(18, [&](int tid) {
      if (tid == 0) logz[(size_t)p.pair * 3 + p.which] = log(TB(c, T_Q, n - 1, 1)) + n * log(c.M->pf_scale);
    });
  }

  // ---- outside: longest spans first
  for (int d = n - 1; d >= TURN + 1; d--) {
    if (c.cp > 0) {
      ex.phase(5, [&](int tid) { outside_nick1(c, sh, d, tid); });
      ex.phase(6, [&](int tid) { outside_nick2(c, sh, d, tid); });
    }
    const int cells = n - d;
    const int chunk = make_split(cells, T).Cp;
    for (int i0 = 1; i0 <= cells; i0 += chunk) {
      const int C = cells - i0 + 1 < chunk ? cells - i0 + 1 : chunk;
      ex.phase(7, [&](int tid) { outside_A(c, sh, d, i0, C, tid); });
      ex.phase(8, [&](int tid) { outside_B(c, sh, d, i0, C, tid); });
    }
  }

  // ---- outputs
  if (p.kind == KIND_LINEAR) {
    if (p.out_bp >= 0) {
      float* bp = dense + p.out_bp;
      ex.phase(9, [&](int tid) { write_bp(c, bp, tid, T); });
      ex.phase(9, [&](int tid) { write_bp2(c, bp, tid, T); });
    }
    if (p.out_up >= 0 && p.max_w > 0) {
      float* up = dense + p.out_up;
      ex.phase(10, [&](int tid) { unstru_hairpin(c, tid, T); });
      ex.phase(11, [&](int tid) { unstru_gaps(c, 0, tid, T); });
      ex.phase(12, [&](int tid) { unstru_gaps(c, 1, tid, T); });
      ex.phase(13, [&](int tid) { unstru_dom_rows(c, tid, T); });
      ex.phase(14, [&](int tid) { unstru_dom_cols(c, tid, T); });
      ex.phase(15, [&](int tid) { unstru_ml_tables(c, tid, T); });
      ex.phase(16, [&](int tid) { unstru_windows(c, up, tid, T); });
    }
  } else if (p.kind == KIND_COFOLD) {
    if (p.out_hp >= 0) {
      float* hp = dense + p.out_hp;
      ex.phase(17, [&](int tid) { write_hp(c, hp, p.n1, p.n2, p.th_hy, tid, T); });
    }
  }
}

}  // namespace rp
#endif
