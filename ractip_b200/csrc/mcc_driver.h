// mcc_driver.h -- the phase schedules, written once against an "executor" that
// runs a per-thread phase function on every thread of the CTA and then
// synchronises them.  kernels.cu instantiates them with a CTA executor
// (threadIdx.x + __syncthreads); tests/emul with a serial one.
//
//   solve_mcc / solve_mcc_wide / solve_mcc_cluster   one problem per CTA (cluster), tables in HBM (any length)
//   solve_band                                       one problem per CTA, interior-loop operands in shared memory
#ifndef RP_MCC_DRIVER_H
#define RP_MCC_DRIVER_H

#include "mcc_band.h"
#include "mcc_core.h"
#ifdef __CUDACC__
#include "mcc_band_shfl.cuh"
#include "mcc_wide_shfl.cuh"
#endif

namespace rp {

enum {
  PH_STAGE = 0, PH_PROLOGUE, PH_PROLOGUE2, PH_INSIDE_A, PH_INSIDE_B, PH_GENERIC_IN, PH_GENERIC_OUT, PH_OUTSIDE_A, PH_OUTSIDE_B,
  PH_WRITE_BP, PH_UN_HAIRPIN, PH_UN_GAPS0, PH_UN_GAPS1, PH_UN_DOMROWS, PH_UN_DOMCOLS, PH_UN_MLTAB, PH_UN_WINDOWS,
  PH_WRITE_HP, PH_LOGZ, PH_BAND_A, PH_BAND_B, PH_CFAC, PH_COUNT
};

template <class Exec, class MT>
RP_HD void emit_unpaired(Exec& ex, Ctx& c, const Problem& p, float* dense, const MT& SM, const double* gfull);

// outputs of a finished problem
// `SM`: the model the table-driven gap loops read (DevModel, or the band kernel's shared-memory copy);
// `gfull`: DevModel::gfull or a shared-memory copy of it
template <class Exec, class MT>
RP_HD void emit_outputs(Exec& ex, Ctx& c, const Problem& p, float* dense, const MT& SM, const double* gfull) {
  const int nct = ex.nthreads();
  if (p.kind == KIND_LINEAR) {
    if (p.out_bp >= 0) {
      ex.phase(PH_WRITE_BP, [&](int tid) { write_bp(c, dense + p.out_bp, tid, nct); });
      ex.phase(PH_WRITE_BP, [&](int tid) { write_bp2(c, dense + p.out_bp, tid, nct); });
    }
    if (p.max_w > 0 && !p.defer_up) emit_unpaired(ex, c, p, dense, SM, gfull);
  } else if (p.kind == KIND_COFOLD) {
    if (p.out_hp >= 0)
      ex.phase(PH_WRITE_HP, [&](int tid) { write_hp(c, dense + p.out_hp, p.n1, p.n2, p.th_hy, tid, nct); });
  }
}

// the unpaired-window pass of a finished single-strand problem (pf_unstru, src/ractip.cpp:371-375): reads the tables
// the two wavefronts left in the problem's workspace.  Run right after them (emit_outputs) or, for the band-kernel
// classes, later by unstru_kernel with its own launch shape.
template <class Exec, class MT>
RP_HD void emit_unpaired(Exec& ex, Ctx& c, const Problem& p, float* dense, const MT& SM, const double* gfull) {
  const int nct = ex.nthreads();
  {
    {
      ex.phase(PH_UN_HAIRPIN, [&](int tid) {
#ifdef __CUDA_ARCH__
        long long* prof = RP_PROF(c);
        const long long t0 = (prof && tid == 0) ? clock64() : 0;
#endif
        unstru_hairpin(c, tid, nct);
#ifdef __CUDA_ARCH__
        const long long t1 = (prof && tid == 0) ? clock64() : 0;
#endif
        unstru_gap_specials(c, SM, tid, nct);
#ifdef __CUDA_ARCH__
        if (prof && tid == 0) {   // RP_PROFILE probes of thread 0: the two halves of the phase
          atomicAdd(reinterpret_cast<unsigned long long*>(prof + 22), (unsigned long long)(t1 - t0));
          atomicAdd(reinterpret_cast<unsigned long long*>(prof + 32 + 22), 1ull);
          atomicAdd(reinterpret_cast<unsigned long long*>(prof + 23), (unsigned long long)(clock64() - t1));
          atomicAdd(reinterpret_cast<unsigned long long*>(prof + 32 + 23), 1ull);
        }
#endif
      });
      ex.phase(PH_UN_GAPS0, [&](int tid) { unstru_gaps(c, gfull, 0, tid, nct); });
      ex.phase(PH_UN_GAPS1, [&](int tid) { unstru_gaps(c, gfull, 1, tid, nct); });
      ex.phase(PH_UN_DOMROWS, [&](int tid) { unstru_dom_rows(c, tid, nct); });
      ex.phase(PH_UN_DOMCOLS, [&](int tid) { unstru_dom_cols(c, tid, nct); });
      ex.phase(PH_UN_MLTAB, [&](int tid) { unstru_ml_tables(c, tid, nct); });
      if (p.out_up >= 0) ex.phase(PH_UN_WINDOWS, [&](int tid) { unstru_windows(c, dense + p.out_up, tid, nct); });
    }
  }
}

// ---------------------------------------------------------------------------
// general kernel: one problem, sliced cells.  dense: base of the dense float
// output.  logz: 3 doubles per pair (s1, s2, s1&s2) or nullptr.
// ---------------------------------------------------------------------------
template <class Exec>
RP_HD void solve_mcc(Exec& ex, Ctx& c, const Problem& p, float* dense, double* logz, const Shared& sh) {
  const int T = sh.T;
  const int n = c.n;
  if (n + 2 <= RP_SMEM_SEQ) {
    const uint8_t* gS = c.S;
    ex.phase(PH_STAGE, [&](int tid) {
      for (int x = tid; x <= n + 1; x += T) sh.S[x] = gS[x];
    });
    c.S = sh.S;
  }
  ex.phase(PH_PROLOGUE, [&](int tid) {
    load_shared_model(*c.M, sh, tid);
    prologue_vectors(c, tid, T);
    prologue_lists(c, tid, T);
  });
  ex.phase(PH_PROLOGUE2, [&](int tid) { prologue2(c, tid, T); });

  // ---- inside: anti-diagonal wavefront, shortest spans first; split sums one band ahead
  for (int d = TURN + 1; d <= n - 1; d++) {
    if (d == band_start_inside(d)) {
      const int rows = n - d;
      int chunk = make_split(rows, T).Cp;
#ifdef __CUDA_ARCH__
      if (chunk > HW * (T / 32)) chunk = HW * (T / 32);   // shuffle variant: 28 rows per warp
#endif
      for (int i0 = 1; i0 <= rows; i0 += chunk) {
        const int C = rows - i0 + 1 < chunk ? rows - i0 + 1 : chunk;
        ex.phase(PH_BAND_A, [&](int tid) {
#ifdef __CUDA_ARCH__
          inside_band_A_shfl<Exec::kBatch>(c, sh, d, i0, C, tid);   // same sums, operands passed along the warp
#else
          inside_band_A(c, sh, d, i0, C, tid);
#endif
        });
        ex.phase(PH_BAND_B, [&](int tid) {
#ifdef __CUDA_ARCH__
          inside_band_B_shfl(c, sh, d, i0, C, tid);
#else
          inside_band_B(c, sh, d, i0, C, tid);
#endif
        });
      }
    }
    const int cells = n - d;
    const int chunk = cells < T ? cells : T;
    for (int i0 = 1; i0 <= cells; i0 += chunk) {
      const int C = cells - i0 + 1 < chunk ? cells - i0 + 1 : chunk;
      ex.phase(PH_INSIDE_A, [&](int tid) { inside_A(c, sh, d, i0, C, tid); });
      ex.phase(PH_INSIDE_B, [&](int tid) { inside_B(c, sh, d, i0, C, tid); });
    }
  }
  inside_end(c);
  if (logz) {
    ex.phase(PH_LOGZ, [&](int tid) {
      if (tid == 0) logz[(size_t)p.pair * 3 + p.which] = log(TB(c, T_Q, n - 1, 1)) + n * log(c.M->pf_scale);
    });
  }

  // ---- outside: longest spans first
  // (two strands: only the inter-strand cells of a diagonal, no nick sums -- cross_lo / cross_hi in mcc_core.h)
  for (int d = n - 1; d >= TURN + 1; d--) {
    if ((n - 1 - d) % BAND == 0) {  // a band of diagonals d, d-1, ..., d-BAND+1 starts here
      const int rows = n - d + BAND - 1;
      int chunk = make_split(rows, T).Cp;
#ifdef __CUDA_ARCH__
      if (chunk > HW * (T / 32)) chunk = HW * (T / 32);
#endif
      for (int r0 = 0; r0 < rows; r0 += chunk) {
        const int C = rows - r0 < chunk ? rows - r0 : chunk;
        ex.phase(PH_BAND_A, [&](int tid) {
#ifdef __CUDA_ARCH__
          outside_band_A_shfl<Exec::kBatch>(c, sh, d, r0, C, tid);
#else
          outside_band_A(c, sh, d, r0, C, tid);
#endif
        });
        ex.phase(PH_BAND_B, [&](int tid) {
#ifdef __CUDA_ARCH__
          outside_band_B_shfl(c, sh, d, r0, C, tid);
#else
          outside_band_B(c, sh, d, r0, C, tid);
#endif
        });
      }
    }
    const int lo = cross_lo(c, d), hi = cross_hi(c, d), cells = hi - lo + 1;
    const int chunk = cells < T ? cells : T;
    for (int i0 = lo; i0 <= hi; i0 += chunk) {
      const int C = hi - i0 + 1 < chunk ? hi - i0 + 1 : chunk;
      ex.phase(PH_OUTSIDE_A, [&](int tid) { outside_A(c, sh, d, i0, C, tid); });
      ex.phase(PH_OUTSIDE_B, [&](int tid) { outside_B(c, sh, d, i0, C, tid); });
    }
  }
  emit_outputs(ex, c, p, dense, *c.M, &c.M->gfull[0][0]);
}

// ---------------------------------------------------------------------------
// general kernel, wide-band schedule (long problems): as solve_mcc, but the split sums are computed
// W = Exec::kWide diagonals at a time (far pass, "Wide bands" in mcc_core.h) and the finishing phase
// of a diagonal adds the few terms the far pass could not see.  `sh` must be carved with W.
// ---------------------------------------------------------------------------
template <class Exec>
RP_HD void solve_mcc_wide(Exec& ex, Ctx& c, const Problem& p, float* dense, double* logz, const Shared& sh) {
  constexpr int W = Exec::kWide;
  const int T = sh.T;
  const int n = c.n;
  if (n + 2 <= RP_SMEM_SEQ) {
    const uint8_t* gS = c.S;
    ex.phase(PH_STAGE, [&](int tid) {
      for (int x = tid; x <= n + 1; x += T) sh.S[x] = gS[x];
    });
    c.S = sh.S;
  }
  ex.phase(PH_PROLOGUE, [&](int tid) {
    load_shared_model(*c.M, sh, tid);
    prologue_vectors(c, tid, T);
    prologue_lists(c, tid, T);
  });
  ex.phase(PH_PROLOGUE2, [&](int tid) {
    prologue2(c, tid, T);
    prologue_rowmajor(c, tid, T);
  });

  for (int d = TURN + 1; d <= n - 1; d++) {
    if (d == wide_start_inside<W>(d)) {
      const int rows = n - d;
      int chunk = make_split(rows, T).Cp;
#ifdef __CUDA_ARCH__
      if (chunk > wide_chunk<W>(T)) chunk = wide_chunk<W>(T);
#endif
      for (int i0 = 1; i0 <= rows; i0 += chunk) {
        const int C = rows - i0 + 1 < chunk ? rows - i0 + 1 : chunk;
        ex.phase(PH_BAND_A, [&](int tid) {
#ifdef __CUDA_ARCH__
          wide_inside_A_shfl<W, Exec::kWideBatch>(c, sh, d, i0, C, tid);
#else
          wide_inside_A<W>(c, sh, d, i0, C, tid);
#endif
        });
        ex.phase(PH_BAND_B, [&](int tid) {
#ifdef __CUDA_ARCH__
          wide_inside_B_shfl<W>(c, sh, d, i0, C, tid);
#else
          wide_inside_B<W>(c, sh, d, i0, C, tid);
#endif
        });
      }
    }
    // Chunks of up to T cells that do not straddle a strand segment (both ends on strand 1 / joining the strands /
    // both on strand 2): the generic-class taps of a chunk are summed densely out of a shared-memory tile
    // (stage_generic_tile + generic_items), the ends and the table-driven shapes per cell (inside_A).
    const int cells = n - d;
    int sb[4] = {1, cells + 1, cells + 1, cells + 1};
    if (c.cp > 0) {
      int b1 = c.cp - d, b2 = c.cp;
      if (b1 < 1) b1 = 1;
      if (b1 > cells + 1) b1 = cells + 1;
      if (b2 > cells + 1) b2 = cells + 1;
      if (b2 < b1) b2 = b1;
      sb[1] = b1; sb[2] = b2;
    }
    const bool staged = W > BAND && gs_smax(n, d, 1) >= GS_ROW0;   // (the tile lives in the wide builds' partial-sum buffer)
    for (int sgm = 0; sgm < 3; sgm++) {
      for (int i0 = sb[sgm]; i0 < sb[sgm + 1]; i0 += T) {
        const int C = sb[sgm + 1] - i0 < T ? sb[sgm + 1] - i0 : T;
        const bool crossing = c.cp > 0 && sgm == 1;
        ex.phase(PH_INSIDE_A, [&](int tid) {
          if (staged) {
            stage_generic_tile<1>(c, sh, d, i0, C, crossing, tid);
            ends_items<1>(c, sh, d, i0, C, tid);
          } else {
            inside_A(c, sh, d, i0, C, tid);
          }
        });
        if (staged) ex.phase(PH_GENERIC_IN, [&](int tid) { generic_items<1>(c, sh, d, i0, C, tid); });
        ex.phase(PH_INSIDE_B, [&](int tid) { wide_inside_finish<W>(c, sh, d, i0, C, tid, staged); });
      }
    }
  }
  inside_end(c);
  if (logz) {
    ex.phase(PH_LOGZ, [&](int tid) {
      if (tid == 0) logz[(size_t)p.pair * 3 + p.which] = log(TB(c, T_Q, n - 1, 1)) + n * log(c.M->pf_scale);
    });
  }

  for (int d = n - 1; d >= TURN + 1; d--) {
    if (d == wide_start_outside<W>(n, d)) {
      const int rows = n - d + W - 1;
      int chunk = make_split(rows, T).Cp;
#ifdef __CUDA_ARCH__
      if (chunk > wide_chunk<W>(T)) chunk = wide_chunk<W>(T);
#endif
      for (int r0 = 0; r0 < rows; r0 += chunk) {
        const int C = rows - r0 < chunk ? rows - r0 : chunk;
        ex.phase(PH_BAND_A, [&](int tid) {
#ifdef __CUDA_ARCH__
          wide_outside_A_shfl<W, Exec::kWideBatch>(c, sh, d, r0, C, tid);
#else
          wide_outside_A<W>(c, sh, d, r0, C, tid);
#endif
        });
        ex.phase(PH_BAND_B, [&](int tid) {
#ifdef __CUDA_ARCH__
          wide_outside_B_shfl<W>(c, sh, d, r0, C, tid);
#else
          wide_outside_B<W>(c, sh, d, r0, C, tid);
#endif
        });
      }
    }
    const int lo = cross_lo(c, d), hi = cross_hi(c, d), cells = hi - lo + 1;
    const int chunk = cells < T ? cells : T;
    const bool staged = W > BAND && gs_smax(n, d, -1) >= GS_ROW0;
    for (int i0 = lo; i0 <= hi; i0 += chunk) {
      const int C = hi - i0 + 1 < chunk ? hi - i0 + 1 : chunk;
      ex.phase(PH_OUTSIDE_A, [&](int tid) {
        if (staged) {
          stage_generic_tile<-1>(c, sh, d, i0, C, false, tid);   // (enclosing pairs of inter-strand cells join the strands anyway)
          ends_items<-1>(c, sh, d, i0, C, tid);
        } else {
          outside_A(c, sh, d, i0, C, tid);
        }
      });
      if (staged) ex.phase(PH_GENERIC_OUT, [&](int tid) { generic_items<-1>(c, sh, d, i0, C, tid); });
      ex.phase(PH_OUTSIDE_B, [&](int tid) { wide_outside_finish<W>(c, sh, d, i0, C, tid, staged); });
    }
  }
  emit_outputs(ex, c, p, dense, *c.M, &c.M->gfull[0][0]);
}

#ifdef __CUDACC__
// ---------------------------------------------------------------------------
// general kernel, ONE problem on a thread-block CLUSTER (long problems when there are too few of them
// to fill the GPU: a single 1000 x 500 pair instead of a shuffle batch).  Same phases as solve_mcc_wide;
// the chunks of a diagonal (cells, far-pass rows) and the items of the flat passes (prologue, unpaired
// windows, outputs) are dealt out over the G CTAs of the cluster.  A chunk's two phases (partial sums,
// then the reduction / the finish of its cells) stay inside one CTA and need only __syncthreads; what
// crosses CTAs -- every table in the problem's HBM workspace -- is ordered by one cluster barrier
// (release/acquire at cluster scope, which also drops the L1) per dependency level: after the far pass
// of a band and after the finishing phase of every diagonal.
// Exec: phase(id, f) = f(threadIdx.x) + __syncthreads; csync() = cluster barrier; rank(), nranks().
// ---------------------------------------------------------------------------
template <class Exec>
__device__ void solve_mcc_cluster(Exec& ex, Ctx& c, const Problem& p, float* dense, double* logz, const Shared& sh) {
  constexpr int W = Exec::kWide;
  const int T = sh.T, n = c.n;
  const int R = ex.rank(), G = ex.nranks(), GT = G * T;
  auto share = [&](int total, int cap) {   // chunk size that spreads `total` items over the G CTAs, at most `cap`
    int ch = (total + G - 1) / G;
    if (ch > cap) ch = cap;
    return ch < 1 ? 1 : ch;
  };
  if (n + 2 <= RP_SMEM_SEQ) {
    const uint8_t* gS = c.S;
    ex.phase(PH_STAGE, [&](int tid) {
      for (int x = tid; x <= n + 1; x += T) sh.S[x] = gS[x];
    });
    c.S = sh.S;
  }
  ex.phase(PH_PROLOGUE, [&](int tid) {
    load_shared_model(*c.M, sh, tid);
    prologue_vectors(c, R * T + tid, GT);
    prologue_lists(c, R * T + tid, GT);
  });
  ex.csync();
  ex.phase(PH_PROLOGUE2, [&](int tid) {
    prologue2(c, R * T + tid, GT);
    prologue_rowmajor(c, R * T + tid, GT);
  });
  ex.csync();

  for (int d = TURN + 1; d <= n - 1; d++) {
    if (d == wide_start_inside<W>(d)) {
      const int rows = n - d;
      const int chunk = share(rows, wide_chunk<W>(T));
      for (int i0 = 1 + R * chunk; i0 <= rows; i0 += G * chunk) {
        const int C = rows - i0 + 1 < chunk ? rows - i0 + 1 : chunk;
        ex.phase(PH_BAND_A, [&](int tid) { wide_inside_A_shfl<W, Exec::kWideBatch>(c, sh, d, i0, C, tid); });
        ex.phase(PH_BAND_B, [&](int tid) { wide_inside_B_shfl<W>(c, sh, d, i0, C, tid); });
      }
      ex.csync();
    }
    const int cells = n - d;
    const int chunk = share(cells, T);
    for (int i0 = 1 + R * chunk; i0 <= cells; i0 += G * chunk) {
      const int C = cells - i0 + 1 < chunk ? cells - i0 + 1 : chunk;
      ex.phase(PH_INSIDE_A, [&](int tid) { inside_A(c, sh, d, i0, C, tid); });
      ex.phase(PH_INSIDE_B, [&](int tid) { wide_inside_finish<W>(c, sh, d, i0, C, tid); });
    }
    ex.csync();
  }
  inside_end(c);
  if (logz && R == 0) {
    ex.phase(PH_LOGZ, [&](int tid) {
      if (tid == 0) logz[(size_t)p.pair * 3 + p.which] = log(TB(c, T_Q, n - 1, 1)) + n * log(c.M->pf_scale);
    });
  }

  for (int d = n - 1; d >= TURN + 1; d--) {
    if (d == wide_start_outside<W>(n, d)) {
      const int rows = n - d + W - 1;
      const int chunk = share(rows, wide_chunk<W>(T));
      for (int r0 = R * chunk; r0 < rows; r0 += G * chunk) {
        const int C = rows - r0 < chunk ? rows - r0 : chunk;
        ex.phase(PH_BAND_A, [&](int tid) { wide_outside_A_shfl<W, Exec::kWideBatch>(c, sh, d, r0, C, tid); });
        ex.phase(PH_BAND_B, [&](int tid) { wide_outside_B_shfl<W>(c, sh, d, r0, C, tid); });
      }
      ex.csync();
    }
    const int lo = cross_lo(c, d), hi = cross_hi(c, d), cells = hi - lo + 1;
    const int chunk = share(cells, T);
    for (int i0 = lo + R * chunk; i0 <= hi; i0 += G * chunk) {
      const int C = hi - i0 + 1 < chunk ? hi - i0 + 1 : chunk;
      ex.phase(PH_OUTSIDE_A, [&](int tid) { outside_A(c, sh, d, i0, C, tid); });
      ex.phase(PH_OUTSIDE_B, [&](int tid) { wide_outside_finish<W>(c, sh, d, i0, C, tid); });
    }
    ex.csync();
  }

  // outputs and the unpaired-window pass: flat loops over all threads of the cluster, a cluster barrier
  // wherever emit_outputs has a phase boundary
  if (p.kind == KIND_LINEAR) {
    if (p.out_bp >= 0) {
      ex.phase(PH_WRITE_BP, [&](int tid) { write_bp(c, dense + p.out_bp, R * T + tid, GT); });
      ex.csync();
      ex.phase(PH_WRITE_BP, [&](int tid) { write_bp2(c, dense + p.out_bp, R * T + tid, GT); });
    }
    if (p.max_w > 0) {
      ex.phase(PH_UN_HAIRPIN, [&](int tid) {
        unstru_hairpin(c, R * T + tid, GT);
        unstru_gap_specials(c, *c.M, R * T + tid, GT);
      });
      ex.csync();
      ex.phase(PH_UN_GAPS0, [&](int tid) { unstru_gaps(c, &c.M->gfull[0][0], 0, R * T + tid, GT); });
      ex.csync();
      ex.phase(PH_UN_GAPS1, [&](int tid) { unstru_gaps(c, &c.M->gfull[0][0], 1, R * T + tid, GT); });
      ex.csync();
      ex.phase(PH_UN_DOMROWS, [&](int tid) { unstru_dom_rows(c, R * T + tid, GT); });
      ex.csync();
      ex.phase(PH_UN_DOMCOLS, [&](int tid) { unstru_dom_cols(c, R * T + tid, GT); });
      ex.csync();
      ex.phase(PH_UN_MLTAB, [&](int tid) { unstru_ml_tables(c, R * T + tid, GT); });
      ex.csync();
      if (p.out_up >= 0) ex.phase(PH_UN_WINDOWS, [&](int tid) { unstru_windows(c, dense + p.out_up, R * T + tid, GT); });
    }
  } else if (p.kind == KIND_COFOLD) {
    if (p.out_hp >= 0)
      ex.phase(PH_WRITE_HP, [&](int tid) { write_hp(c, dense + p.out_hp, p.n1, p.n2, p.th_hy, R * T + tid, GT); });
  }
  ex.csync();   // the slot is reused by the cluster's next problem
}
#endif

// ---------------------------------------------------------------------------
// band kernel: one problem per CTA, interior-loop operands in a shared-memory
// ring of the last 32 diagonals (mcc_band.h); split sums, nick sums, unpaired
// windows and outputs as in solve_mcc.  `smem` is the CTA's dynamic shared
// memory (band_shared_bytes(n, T) bytes).
// ---------------------------------------------------------------------------
template <class Exec>
RP_HD void solve_band(Exec& ex, Ctx& c, const Problem& p, float* dense, double* logz, void* smem) {
  const int T = ex.nthreads();
  const int n = c.n;
  Shared sh;
  BandShared bs;
  carve_band(sh, bs, smem, n, T);
  const bool wide = c.kind == KIND_LINEAR && c.max_w > 0;  // the unpaired-window pass reads the class tables' full history
  {
    const uint8_t* gS = c.S;
    ex.phase(PH_STAGE, [&](int tid) {
      for (int x = tid; x <= n + 1; x += T) sh.S[x] = gS[x];
      load_band_weights(*c.M, bs, tid, T);
      prologue_vectors(c, tid, T);
    });
    c.S = sh.S;
  }
  ex.phase(PH_PROLOGUE2, [&](int tid) {
    prologue2(c, tid, T);
    if (tid == T - 1 && TURN + 1 <= n - 1) band_make_desc(bs.desc[(TURN + 1) & (NDESC - 1)], n, c.cp, TURN + 1, T);
  });

  // Per-diagonal descriptors (strand segments, item schedule, thread roles: DiagDesc) are worked out by ONE
  // thread during the long phase of the step before their first use.
  // Schedule: the completion ("finish") of diagonal d runs one step late, in the same long phase
  // as the interior items of the next diagonal, so that its HBM/L2 round trips and the nick sums
  // hide behind the items' arithmetic.  Per diagonal: one long phase, then one quick phase that
  // collects the items' partial sums (freeing the partial buffer for the split-sum band phases,
  // which follow every BAND diagonals) and sets up the next diagonal's closing factors.
  // ---- inside: diagonals TURN+1 .. n-1
  ex.phase(PH_CFAC, [&](int tid) { band_collect<1>(c, bs, -1, TURN + 1 <= n - 1 ? TURN + 1 : -1, tid, T); });
  for (int d = TURN + 1; d <= n; d++) {
    const int dfin = d - 1 >= TURN + 1 ? d - 1 : -1;   // diagonal to complete in this step
    const int dnew = d <= n - 1 ? d : -1;              // diagonal whose items run in this step
    ex.phase(PH_INSIDE_A, [&](int tid) {
      if (dfin >= 0) band_inside_B(c, sh, bs, dfin, wide, tid);
      if (dnew >= 0) band_interior_A<1>(c, bs, dnew, tid, T);
      if (tid == T - 1 && d + 1 <= n - 1) band_make_desc(bs.desc[(d + 1) & (NDESC - 1)], n, c.cp, d + 1, T);
    });
    if (dnew < 0) break;
    ex.phase(PH_CFAC, [&](int tid) { band_collect<1>(c, bs, dnew, d + 1 <= n - 1 ? d + 1 : -1, tid, T); });
    // split sums of diagonals d .. d+BAND-1: every operand lies on a diagonal < d, complete now
    if (d == band_start_inside(d)) {
      const int rows = n - d;
      int chunk = make_split(rows, T).Cp;
#ifdef __CUDA_ARCH__
      if (chunk > HW * (T / 32)) chunk = HW * (T / 32);   // shuffle variant: 28 rows per warp
#endif
      for (int i0 = 1; i0 <= rows; i0 += chunk) {
        const int C = rows - i0 + 1 < chunk ? rows - i0 + 1 : chunk;
        ex.phase(PH_BAND_A, [&](int tid) {
#ifdef __CUDA_ARCH__
          inside_band_A_shfl(c, sh, d, i0, C, tid);   // same sums, operands passed along the warp
#else
          inside_band_A(c, sh, d, i0, C, tid);
#endif
        });
        ex.phase(PH_BAND_B, [&](int tid) {
#ifdef __CUDA_ARCH__
          inside_band_B_shfl(c, sh, d, i0, C, tid);
#else
          inside_band_B(c, sh, d, i0, C, tid);
#endif
        });
      }
    }
  }
  inside_end(c);
  if (logz) {
    ex.phase(PH_LOGZ, [&](int tid) {
      if (tid == 0) logz[(size_t)p.pair * 3 + p.which] = log(TB(c, T_Q, n - 1, 1)) + n * log(c.M->pf_scale);
    });
  }

  // ---- outside: diagonals n-1 .. TURN+1; step d completes diagonal d+1 and runs the items of d.
  if (n - 1 >= TURN + 1)
    ex.phase(PH_CFAC, [&](int tid) { if (tid == 0) band_make_desc(bs.desc[(n - 1) & (NDESC - 1)], n, c.cp, n - 1, T, true); });
  ex.phase(PH_CFAC, [&](int tid) { band_collect<-1>(c, bs, -1, n - 1 >= TURN + 1 ? n - 1 : -1, tid, T); });
  for (int d = n - 1; d >= TURN; d--) {
    const int dfin = d + 1 <= n - 1 ? d + 1 : -1;
    const int dnew = d >= TURN + 1 ? d : -1;
    // split sums of diagonals d .. d-BAND+1: operands on diagonals >= d+2, complete after the previous step
    if (dnew >= 0 && (n - 1 - d) % BAND == 0) {
      const int rows = n - d + BAND - 1;
      int chunk = make_split(rows, T).Cp;
#ifdef __CUDA_ARCH__
      if (chunk > HW * (T / 32)) chunk = HW * (T / 32);   // shuffle variant: 28 rows per warp
#endif
      for (int r0 = 0; r0 < rows; r0 += chunk) {
        const int C = rows - r0 < chunk ? rows - r0 : chunk;
        ex.phase(PH_BAND_A, [&](int tid) {
#ifdef __CUDA_ARCH__
          outside_band_A_shfl(c, sh, d, r0, C, tid);
#else
          outside_band_A(c, sh, d, r0, C, tid);
#endif
        });
        ex.phase(PH_BAND_B, [&](int tid) {
#ifdef __CUDA_ARCH__
          outside_band_B_shfl(c, sh, d, r0, C, tid);
#else
          outside_band_B(c, sh, d, r0, C, tid);
#endif
        });
      }
    }
    ex.phase(PH_OUTSIDE_A, [&](int tid) {
      if (dfin >= 0) band_outside_B(c, sh, bs, dfin, wide, tid);
      if (dnew >= 0) band_interior_A<-1>(c, bs, dnew, tid, T);
      if (tid == T - 1 && d - 1 >= TURN + 1) band_make_desc(bs.desc[(d - 1) & (NDESC - 1)], n, c.cp, d - 1, T, true);
    });
    if (dnew < 0) break;
    ex.phase(PH_CFAC, [&](int tid) {
      band_collect<-1>(c, bs, dnew, d - 1 >= TURN + 1 ? d - 1 : -1, tid, T);
    });
  }
  // the ring is free now: it takes a copy of the gap-sum weights
  const double* gfull = &c.M->gfull[0][0];
  if (wide && !p.defer_up && (size_t)3 * BSLOTS * bs.LDB >= (size_t)(MAXLOOP + 1) * GROW_LD) {
    ex.phase(PH_STAGE, [&](int tid) {
      for (int x = tid; x < (MAXLOOP + 1) * GROW_LD; x += T) bs.TI[x] = (&c.M->gfull[0][0])[x];
    });
    gfull = bs.TI;
  }
  emit_outputs(ex, c, p, dense, *bs.sm, gfull);
}

// the unpaired-window pass of a problem the band kernel finished earlier, run as a job of its own inside that kernel
template <class Exec>
RP_HD void solve_band_unpaired(Exec& ex, Ctx& c, const Problem& p, float* dense, void* smem) {
  const int T = ex.nthreads(), n = c.n;
  Shared sh;
  BandShared bs;
  carve_band(sh, bs, smem, n, T);
  const uint8_t* gS = c.S;
  ex.phase(PH_STAGE, [&](int tid) {
    for (int x = tid; x <= n + 1; x += T) sh.S[x] = gS[x];
    load_band_weights(*c.M, bs, tid, T);
  });
  c.S = sh.S;
  inside_end(c);
  const double* gfull = &c.M->gfull[0][0];
  if ((size_t)3 * BSLOTS * bs.LDB >= (size_t)(MAXLOOP + 1) * GROW_LD) {
    ex.phase(PH_STAGE, [&](int tid) {
      for (int x = tid; x < (MAXLOOP + 1) * GROW_LD; x += T) bs.TI[x] = (&c.M->gfull[0][0])[x];
    });
    gfull = bs.TI;
  }
  emit_unpaired(ex, c, p, dense, *bs.sm, gfull);
}

}  // namespace rp
#endif