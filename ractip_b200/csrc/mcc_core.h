// mcc_core.h -- the McCaskill wavefront, written once as per-thread phase
// functions.  The CUDA kernel (kernels.cu) runs them with tid = threadIdx.x and
// __syncthreads() between phases; tests/emul compiles the SAME functions for
// the host and runs the threads of a CTA one after another, so kernel logic is
// debugged on a CPU-only box.  (The emulator is test infrastructure; the
// product only ever runs the CUDA build.)
//
// What is computed (reference call sites):
//   linear   : Vienna::pf_fold + export_bppm  (src/ractip.cpp:356-367)
//              Vienna::pf_unstru, sum H+I+M+E (src/ractip.cpp:371-375)
//   two-strand: Vienna::co_pf_fold + export_co_bppm (src/ractip.cpp:442-447)
// with dangles=2, TURN=3, MAXLOOP=30, pf_scale from DevModel.
//
// Layout: every O(n^2) table is stored DIAGONAL-MAJOR: cell (i,j), 1<=i<=j<=n,
// lives at [d*ld + i] with d=j-i, ld=n+1.  Cells of one anti-diagonal wavefront
// are contiguous, so "thread t handles cell i0+t" makes every operand stream of
// every recurrence a unit-stride (coalesced) access:
//   sum_k A(i,k-1)*B(k,j)  ->  sum_a A[a][i] * B[d-1-a][i+1+a]
//   interior window        ->  sum_taps g * B[d-dd][i+po]
#ifndef RP_MCC_CORE_H
#define RP_MCC_CORE_H

#include <math.h>
#include <stdint.h>

#include "dev_model.h"

#ifdef __CUDACC__
#define RP_HD __host__ __device__ __forceinline__
#define RP_D __device__ __forceinline__
#else
#define RP_HD inline
#define RP_D inline
#endif

namespace rp {

// ---------------------------------------------------------------------------
// problem descriptor and per-slot workspace
// ---------------------------------------------------------------------------
enum { KIND_LINEAR = 0, KIND_COFOLD = 1, KIND_DUPLEX = 2 };

struct Problem {
  int seq_off;        // offset of S[1] in the batch's encoded-sequence buffer
  int n;              // length (n1+n2 for two-strand problems)
  int cp;             // first index of strand 2 (0: single strand)
  int kind;
  int pair;           // index of the rp_pair this problem belongs to
  int which;          // 0: s1, 1: s2, 2: s1&s2
  int max_w;
  int n1, n2;
  long long out_bp;   // float offsets into the dense output, -1 if absent
  long long out_up;
  long long out_hp;
  float th_hy;
  int pad;
};

enum {
  T_Q = 0, T_QQ, T_QM, T_QM1, T_QM2, T_QB, T_QBI, T_QB1N, T_QBAU,
  T_OUT, T_OUTI, T_OUT1N, T_OUTAU, T_MC, T_PR, T_PRML, T_PMLB, T_PL,
  T_DG, T_RR, T_LL, T_XX,
  T_LIST,   // not doubles: per-diagonal lists of pairable cells (uint16), see listp()/posp()
  T_COUNT
};
enum {
  V_SCALE = 0, V_MLB, V_HPW, V_SP3, V_SP4, V_SP6, V_U0, V_U1,
  V_QR, V_QROUT, V_QL, V_QLOUT,
  V_COUNT
};

struct Ctx {
  const DevModel* M;
  const uint8_t* S;   // S[1..n]; low 3 bits base code 0..4, bit 3 = "letter is not A/C/G/U"
  int n, cp, ld, kind, max_w;
  double* ws;         // slot workspace: T_COUNT tables then V_COUNT vectors
  size_t te, ve;      // elements per table / per vector
  double invZ;        // set after the inside pass
  int dbg;            // tuning aid (RP_DEBUG_SKIP): 1 skip interior rows, 2 skip split sums, 4 skip gap sums
};

RP_HD size_t table_elems(int n) { return (size_t)n * (size_t)(n + 1) + 8; }
RP_HD size_t vector_elems(int n) { return (size_t)n + 8; }
RP_HD size_t slot_doubles(int n) { return T_COUNT * table_elems(n) + V_COUNT * vector_elems(n); }

RP_HD void bind_ctx(Ctx& c, const DevModel* M, const uint8_t* S, const Problem& p, double* ws) {
  c.M = M; c.S = S; c.n = p.n; c.cp = p.cp; c.ld = p.n + 1; c.kind = p.kind; c.max_w = p.max_w;
  c.ws = ws; c.te = table_elems(p.n); c.ve = vector_elems(p.n);
  c.invZ = 0;
  c.dbg = 0;
}
RP_HD double* tabp(const Ctx& c, int t) { return c.ws + (size_t)t * c.te; }
RP_HD double* vecp(const Ctx& c, int v) { return c.ws + (size_t)T_COUNT * c.te + (size_t)v * c.ve; }

#define TB(c, t, d, i) (tabp(c, t)[(size_t)(d) * (c).ld + (i)])

// Per-diagonal compaction of the cells that can pair (static per sequence):
//   LIST[d*ld + r] = i of the r-th pairable cell (i,i+d);  POS[d*ld + i] = #pairable cells (i',i'+d), i' < i.
// Interior-loop work is dealt out over these lists, so no thread idles on a cell that cannot pair.
RP_HD uint16_t* listp(const Ctx& c) { return reinterpret_cast<uint16_t*>(tabp(c, T_LIST)); }
RP_HD uint16_t* posp(const Ctx& c) { return reinterpret_cast<uint16_t*>(tabp(c, T_LIST)) + c.te * 2; }

// CTA-shared scratch (CUDA shared memory; a heap block in the host emulation)
constexpr int RP_SMEM_SEQ = 4096;  // sequences up to this length are staged in shared memory
struct Shared {
  int T;
  double* part;     // [3][T] partial sums of the current phase
  double* grow;     // [MAXLOOP+1][GROW_LD] run weights of the factorised interior loops (DevModel::grow)
  double* ghead_b;  // [GROW_LD]
  double* ghead_1;  // [GROW_LD]
  double* red;      // [128] small reductions (nick sums)
  uint8_t* S;       // [RP_SMEM_SEQ + 8] staged sequence
};
RP_HD size_t shared_bytes(int T) {
  return sizeof(double) * (3 * (size_t)T + (MAXLOOP + 1) * GROW_LD + 2 * GROW_LD + 128) + RP_SMEM_SEQ + 16;
}
RP_HD void carve_shared(Shared& sh, void* base, int T) {
  sh.T = T;
  double* p = static_cast<double*>(base);
  sh.part = p; p += 3 * (size_t)T;
  sh.grow = p; p += (MAXLOOP + 1) * GROW_LD;
  sh.ghead_b = p; p += GROW_LD;
  sh.ghead_1 = p; p += GROW_LD;
  sh.red = p; p += 128;
  sh.S = reinterpret_cast<uint8_t*>(p);
}

// ---------------------------------------------------------------------------
// small helpers
// ---------------------------------------------------------------------------
RP_HD int base(const Ctx& c, int i) { return c.S[i] & 7; }

RP_HD int pair_type(int a, int b) {
  // CG=1 GC=2 GU=3 UG=4 AU=5 UA=6 ; a,b in 0..4 (N,A,C,G,U).  One nibble per (a-1,b-1).
  const uint64_t LUT = (5ull << 12) | (1ull << 24) | (2ull << 36) | (3ull << 44) | (6ull << 48) | (4ull << 56);
  const int idx = ((a - 1) & 3) * 4 + ((b - 1) & 3);
  const int t = (int)((LUT >> (4 * idx)) & 15);
  return (a && b) ? t : 0;
}
RP_HD int rtype(int t) { return t == 0 ? 0 : (t == 7 ? 7 : ((t - 1) ^ 1) + 1); }

// ViennaRNA SAME_STRAND(a,b) for a<b
RP_HD bool ss(const Ctx& c, int a, int b) { return c.cp <= 0 || a >= c.cp || b < c.cp; }

RP_HD double ext_stem(const DevModel& M, int type, int s5, int s3) {
  double e = 1.0;
  if (s5 >= 0 && s3 >= 0) e = M.mmExt[type][s5][s3];
  else if (s5 >= 0) e = M.dangle5[type][s5];
  else if (s3 >= 0) e = M.dangle3[type][s3];
  if (type > 2) e *= M.expTermAU;
  return e;
}
RP_HD double ml_stem(const DevModel& M, int type, int s5, int s3) {
  double e = 1.0;
  if (s5 >= 0 && s3 >= 0) e = M.mmM[type][s5][s3];
  else if (s5 >= 0) e = M.dangle5[type][s5];
  else if (s3 >= 0) e = M.dangle3[type][s3];
  if (type > 2) e *= M.expTermAU;
  return e * M.expMLintern;
}

// Interior-loop weight (unscaled) for the table-driven small cases and, for
// completeness, every other case.  type2 is rtype of the inner pair.
RP_HD double int_loop(const DevModel& M, int u1, int u2, int type, int type2, int si1, int sj1, int sp1, int sq1) {
  const int ul = u1 > u2 ? u1 : u2, us = u1 > u2 ? u2 : u1;
  if (ul == 0) return M.expstack[type][type2];
  if (us == 0) {
    double z = M.expbulge[ul];
    if (ul == 1) z *= M.expstack[type][type2];
    else {
      if (type > 2) z *= M.expTermAU;
      if (type2 > 2) z *= M.expTermAU;
    }
    return z;
  }
  if (us == 1) {
    if (ul == 1) return M.int11[type][type2][si1][sj1];
    if (ul == 2) return u1 == 1 ? M.int21[type][type2][si1][sq1][sj1] : M.int21[type2][type][sq1][si1][sp1];
    return M.expinternal[ul + us] * M.mm1n[type][si1][sj1] * M.mm1n[type2][sq1][sp1] * M.expninio[ul - us];
  }
  if (us == 2) {
    if (ul == 2) return M.int22[type][type2][si1][sp1][sq1][sj1];
    if (ul == 3) return M.expinternal[5] * M.mm23[type][si1][sj1] * M.mm23[type2][sq1][sp1] * M.expninio[1];
  }
  return M.expinternal[ul + us] * M.mmI[type][si1][sj1] * M.mmI[type2][sq1][sp1] * M.expninio[ul - us];
}

// the nine (u1,u2) combinations that do not factorise
#define RP_N_SPECIAL 9
RP_HD void special_uv(int s, int& u1, int& u2) {
  const int U1[RP_N_SPECIAL] = {0, 1, 0, 1, 1, 2, 2, 2, 3};
  const int U2[RP_N_SPECIAL] = {0, 0, 1, 1, 2, 1, 2, 3, 2};
  u1 = U1[s]; u2 = U2[s];
}

// hairpin weight of pair (i,j), including scale[u+2]
RP_HD double hairpin(const Ctx& c, int i, int j, int type) {
  const int u = j - i - 1;
  if (c.M->special_hp) {
    if (u == 4 && vecp(c, V_SP4)[i] >= 0.) return vecp(c, V_SP4)[i];  // type==7 never occurs for ACGU pairs
    if (u == 6 && vecp(c, V_SP6)[i] >= 0.) return vecp(c, V_SP6)[i];
    if (u == 3) {
      if (vecp(c, V_SP3)[i] >= 0.) return vecp(c, V_SP3)[i];
      return type > 2 ? vecp(c, V_HPW)[3] * c.M->expTermAU : vecp(c, V_HPW)[3];
    }
  }
  return vecp(c, V_HPW)[u] * c.M->mmH[type][base(c, i + 1)][base(c, j - 1)];
}

// partition of a chunk of `C` cells over T threads: Cp cells x S slices
struct Split {
  int Cp, S;
};
RP_HD Split make_split(int C, int T) {
  Split s;
  int Cp = (C + 31) & ~31;
  if (Cp > T) Cp = T;
  s.Cp = Cp;
  s.S = T / Cp;
  if (s.S < 1) s.S = 1;
  return s;
}

// ---------------------------------------------------------------------------
// prologue: shared tap tables, per-problem vectors, pair lists, d<=TURN diagonals
// ---------------------------------------------------------------------------
// copy S[0..n+1] into shared memory (the caller then points c.S at sh.S)
RP_HD void stage_sequence(const Ctx& c, const Shared& sh, int tid) {
  for (int x = tid; x <= c.n + 1; x += sh.T) sh.S[x] = c.S[x];
}

RP_HD void prologue(Ctx& c, const Shared& sh, int tid) {
  const DevModel& M = *c.M;
  const int n = c.n, T = sh.T;
  // interior-loop run weights into shared memory
  for (int x = tid; x < (MAXLOOP + 1) * GROW_LD; x += T) sh.grow[x] = M.grow[x / GROW_LD][x % GROW_LD];
  for (int x = tid; x < GROW_LD; x += T) {
    sh.ghead_b[x] = M.ghead_b[x];
    sh.ghead_1[x] = M.ghead_1[x];
  }
  // scale[k] = pf_scale^-k, mlb[k] = (expMLbase/pf_scale)^k: built by repeated
  // multiplication by one thread so that every consumer sees the same values
  if (tid == 0) {
    double s = 1.0, b = 1.0;
    double* sc = vecp(c, V_SCALE);
    double* ml = vecp(c, V_MLB);
    for (int k = 0; k <= n + 2; k++) {
      sc[k] = s;
      ml[k] = b;
      s *= M.scale1;
      b *= M.mlb1;
    }
  }
  // pair lists: one thread per diagonal
  uint16_t* LIST = listp(c);
  uint16_t* POS = posp(c);
  for (int d = tid; d < n; d += T) {
    uint16_t* L = LIST + (size_t)d * c.ld;
    uint16_t* P = POS + (size_t)d * c.ld;
    int cnt = 0;
    for (int i = 1; i <= n - d; i++) {
      P[i] = (uint16_t)cnt;
      if (d > TURN && pair_type(base(c, i), base(c, i + d))) L[cnt++] = (uint16_t)i;
    }
    P[n - d + 1 <= n ? n - d + 1 : n] = (uint16_t)cnt;  // d = 0: i = n+1 does not exist and is never asked for
    if (d == 0) P[n] = 0;
  }
}
RP_HD void prologue2(Ctx& c, const Shared& sh, int tid) {
  const DevModel& M = *c.M;
  const int n = c.n, T = sh.T;
  for (int u = tid; u <= n; u += T) {
    double q;
    if (u <= 30) q = M.exphairpin[u];
    else q = M.exphairpin[30] * exp(-(M.lxc * log(u / 30.)) * 10. / M.kT);
    vecp(c, V_HPW)[u] = q * vecp(c, V_SCALE)[u + 2];
  }
  for (int i = tid; i <= n + 1; i += T) {
    double s3 = -1., s4 = -1., s6 = -1.;
    if (i >= 1) {
      // window codes: base-8 digits, 7 for letters that cannot match a list entry
      int code = 0;
      bool in = true;
      for (int k = 0; k < 8; k++) {
        int p = i + k;
        int dgt = 0;
        if (p <= n) dgt = (c.S[p] & 8) ? 7 : (c.S[p] & 7);
        else in = false;
        code = code * 8 + dgt;
        // a hairpin window never spans the nick (hairpins need ss(i,j))
        if (k == 4 && in) {
          for (int e = 0; e < M.n_tri; e++)
            if (M.tri_code[e] == code) { s3 = M.exptri[e] * vecp(c, V_SCALE)[5]; break; }
        } else if (k == 5 && in) {
          for (int e = 0; e < M.n_tetra; e++)
            if (M.tetra_code[e] == code) { s4 = M.exptetra[e] * vecp(c, V_SCALE)[6]; break; }
        } else if (k == 7 && in) {
          for (int e = 0; e < M.n_hex; e++)
            if (M.hex_code[e] == code) { s6 = M.exphex[e] * vecp(c, V_SCALE)[8]; break; }
        }
      }
    }
    vecp(c, V_SP3)[i] = s3; vecp(c, V_SP4)[i] = s4; vecp(c, V_SP6)[i] = s6;
    vecp(c, V_U0)[i] = 0.; vecp(c, V_U1)[i] = 0.;
    vecp(c, V_QR)[i] = 0.; vecp(c, V_QROUT)[i] = 0.; vecp(c, V_QL)[i] = 0.; vecp(c, V_QLOUT)[i] = 0.;
  }
  // diagonals 0..TURN: q = scale[d+1], everything else 0
  const int dmax = TURN < n - 1 ? TURN : n - 1;
  const int cells = (dmax + 1) * c.ld;
  for (int x = tid; x < cells; x += T) {
    int d = x / c.ld, i = x % c.ld;
    bool valid = i >= 1 && i + d <= n;
    TB(c, T_Q, d, i) = valid ? vecp(c, V_SCALE)[d + 1] : 0.;
    TB(c, T_QQ, d, i) = 0.; TB(c, T_QM, d, i) = 0.; TB(c, T_QM1, d, i) = 0.; TB(c, T_QM2, d, i) = 0.;
    TB(c, T_QB, d, i) = 0.; TB(c, T_QBI, d, i) = 0.; TB(c, T_QB1N, d, i) = 0.; TB(c, T_QBAU, d, i) = 0.;
    TB(c, T_OUT, d, i) = 0.; TB(c, T_OUTI, d, i) = 0.; TB(c, T_OUT1N, d, i) = 0.; TB(c, T_OUTAU, d, i) = 0.;
    TB(c, T_MC, d, i) = 0.; TB(c, T_PR, d, i) = 0.; TB(c, T_PRML, d, i) = 0.; TB(c, T_PMLB, d, i) = 0.;
    TB(c, T_PL, d, i) = 0.; TB(c, T_DG, d, i) = 0.;
  }
}

// ---------------------------------------------------------------------------
// shared pieces of the inside and outside phases
// ---------------------------------------------------------------------------
// weighted sum along one row of the factorised interior loops:
//   sum_{u2=lo..hi} g[u2] * p[u2*step]
RP_HD double row_sum(const double* g, const double* p, int step, int lo, int hi) {
  double a0 = 0., a1 = 0., a2 = 0., a3 = 0.;
  const double* q = p + (long)lo * step;
  g += lo;
  int cnt = hi - lo + 1;
  for (; cnt >= 4; cnt -= 4) {
    a0 += g[0] * q[0];
    a1 += g[1] * q[step];
    a2 += g[2] * q[2 * step];
    a3 += g[3] * q[3 * step];
    g += 4;
    q += 4 * step;
  }
  for (; cnt > 0; cnt--) {
    a0 += g[0] * q[0];
    g++;
    q += step;
  }
  return (a0 + a1) + (a2 + a3);
}

// The factorised part of the interior-loop sum of one cell, restricted to the
// row pairs (q, 30-q), q = sl, sl+SI, ... (each pair holds 32 terms, so slices
// are balanced).  SIGN=+1: inside, inner pair (i+1+u1, j-1-u2) lies u1+u2+2
// diagonals below the cell; SIGN=-1: outside, enclosing pair (k-1-u1, l+1+u2)
// lies above.  cell0 = d*ld + i.  Bounds: u1 <= u1max, u2 <= u2cap,
// u1+u2+2 <= ddmax; every element touched is a valid cell, so no guards.
template <int SIGN>
RP_HD void interior_rows(const Shared& sh, const double* TI, const double* T1, const double* TA, int cell0, int ld,
                         int u1max, int u2cap, int ddmax, int sl, int SI, double& sI, double& s1, double& sA) {
  const int step = -SIGN * ld;
  for (int q = sl; q <= MAXLOOP / 2; q += SI) {
    for (int h = 0; h < 2; h++) {
      const int u1 = h == 0 ? q : MAXLOOP - q;
      if (h == 1 && u1 == q) break;
      if (u1 > u1max) continue;
      int u2hi = MAXLOOP - u1;
      if (u2cap < u2hi) u2hi = u2cap;
      if (ddmax - 2 - u1 < u2hi) u2hi = ddmax - 2 - u1;
      if (u2hi < 0) continue;
      const int o0 = cell0 - SIGN * ((u1 + 2) * ld - (1 + u1));  // element (u1, u2=0)
      const double* g = sh.grow + u1 * GROW_LD;
      if (u1 == 0) {
        if (u2hi >= 2) sA += row_sum(g, TA + o0, step, 2, u2hi);
      } else if (u1 == 1) {
        if (u2hi >= 3) s1 += row_sum(g, T1 + o0, step, 3, u2hi);
      } else {
        sA += sh.ghead_b[u1] * TA[o0];
        if (u2hi >= 1) s1 += sh.ghead_1[u1] * T1[o0 + step];
        if (u2hi >= 2) sI += row_sum(g, TI + o0, step, 2, u2hi);
      }
    }
  }
}

// the nine loop shapes that do not factorise, closing pair `type` with
// neighbours (si1,sj1), inner pair of reversed type t2r with neighbours (sp1,sq1)
RP_HD double special_loop(const DevModel& M, int s, int type, int t2r, int si1, int sj1, int sp1, int sq1) {
  switch (s) {
    case 0: return M.expstack[type][t2r] * M.scale_small[2];
    case 1:
    case 2: return M.expbulge[1] * M.expstack[type][t2r] * M.scale_small[3];
    case 3: return M.int11[type][t2r][si1][sj1] * M.scale_small[4];
    case 4: return M.int21[type][t2r][si1][sq1][sj1] * M.scale_small[5];   // u1=1,u2=2
    case 5: return M.int21[t2r][type][sq1][si1][sp1] * M.scale_small[5];   // u1=2,u2=1
    case 6: return M.int22[type][t2r][si1][sp1][sq1][sj1] * M.scale_small[6];
    default: return M.expinternal[5] * M.expninio[1] * M.mm23[type][si1][sj1] * M.mm23[t2r][sq1][sp1] * M.scale_small[7];
  }
}

// sum of A[x*sa] * B[x*sb] over x in [x0,x1) with x = s (mod S)
RP_HD double dot_range(const double* A, int sa, const double* B, int sb, int x0, int x1, int s, int S) {
  int x = x0 + ((s - x0) % S + S) % S;
  if (x >= x1) return 0.;
  int cnt = (x1 - 1 - x) / S + 1;
  const double* a = A + (long)x * sa;
  const double* b = B + (long)x * sb;
  const int da = S * sa, db = S * sb;
  double a0 = 0., a1 = 0., a2 = 0., a3 = 0.;
  for (; cnt >= 4; cnt -= 4) {
    a0 += a[0] * b[0];
    a1 += a[da] * b[db];
    a2 += a[2 * da] * b[2 * db];
    a3 += a[3 * da] * b[3 * db];
    a += 4 * da;
    b += 4 * db;
  }
  for (; cnt > 0; cnt--) {
    a0 += a[0] * b[0];
    a += da;
    b += db;
  }
  return (a0 + a1) + (a2 + a3);
}
// same over x in [0,cnt) skipping x == skip (the split that would fall on the nick)
RP_HD double strided_dot(const double* A, int sa, const double* B, int sb, int cnt, int s, int S, int skip) {
  if (skip < 0 || skip >= cnt) return dot_range(A, sa, B, sb, 0, cnt, s, S);
  return dot_range(A, sa, B, sb, 0, skip, s, S) + dot_range(A, sa, B, sb, skip + 1, cnt, s, S);
}

// work split of the interior-loop items of a chunk: cnt pairable cells x SI slices
struct ISplit {
  int lo, cnt, cntp, SI;
};
RP_HD ISplit make_isplit(const Ctx& c, int d, int i0, int C, int T) {
  const uint16_t* P = posp(c) + (size_t)d * c.ld;
  ISplit s;
  s.lo = P[i0];
  s.cnt = (int)P[i0 + C] - s.lo;
  int cp = (s.cnt + 31) & ~31;
  if (cp > T) cp = T;
  if (cp < 32) cp = 32;
  s.cntp = cp;
  int si = T / cp;  // slices per cell: a power of two <= 16 (there are 16 row pairs)
  s.SI = si >= 16 ? 16 : si >= 8 ? 8 : si >= 4 ? 4 : si >= 2 ? 2 : 1;
  return s;
}

// ---------------------------------------------------------------------------
// inside pass, diagonal d >= TURN+1; cells i0 .. i0+C-1 handled as a chunk
// sh.part: [3][T] doubles (interior, QM2, q-split)
// ---------------------------------------------------------------------------
RP_HD void inside_A(const Ctx& c, const Shared& sh, int d, int i0, int C, int tid) {
  const DevModel& M = *c.M;
  const int T = sh.T;
  const long ld = c.ld;
  // --- (1) interior loops: items = (pairable cell, slice) ---------------------
  const int ddmax = d - (TURN + 1) < MAXLOOP + 2 ? d - (TURN + 1) : MAXLOOP + 2;
  if (ddmax >= 2) {
    const ISplit is = make_isplit(c, d, i0, C, T);
    const int r = tid % is.cntp, sl = tid / is.cntp;
    if (sl < is.SI && r < is.cnt) {
      const int i = listp(c)[(size_t)d * ld + is.lo + r], j = i + d;
      const int type = pair_type(base(c, i), base(c, j));
      // strand guards: inner 5' end must stay on i's strand, inner 3' end on j's
      const int maxpo = (c.cp > 0 && i < c.cp) ? c.cp - 1 - i : 1000;
      const int maxu2 = (c.cp > 0 && j >= c.cp) ? j - 1 - c.cp : 1000;
      const int si1 = base(c, i + 1), sj1 = base(c, j - 1);
      int u1max = ddmax - 2 < MAXLOOP ? ddmax - 2 : MAXLOOP;
      if (maxpo - 1 < u1max) u1max = maxpo - 1;
      double sI = 0., s1 = 0., sA = 0.;
      if (!(c.dbg & 1)) interior_rows<1>(sh, tabp(c, T_QBI), tabp(c, T_QB1N), tabp(c, T_QBAU), d * (int)ld + i, (int)ld, u1max, maxu2,
                       ddmax, sl, is.SI, sI, s1, sA);
      double accI = M.mmI[type][si1][sj1] * sI + M.mm1n[type][si1][sj1] * s1 + (type > 2 ? M.expTermAU : 1.0) * sA;
      // table-driven small loops
      for (int s = sl; s < RP_N_SPECIAL; s += is.SI) {
        int u1, u2;
        special_uv(s, u1, u2);
        const int dd = u1 + u2 + 2;
        if (dd > ddmax || u1 + 1 > maxpo || u2 > maxu2) continue;
        const int k = i + 1 + u1, l = j - 1 - u2;
        const int t2 = pair_type(base(c, k), base(c, l));
        if (!t2) continue;
        accI += TB(c, T_QB, d - dd, k) * special_loop(M, s, type, rtype(t2), si1, sj1, base(c, k - 1), base(c, l + 1));
      }
      sh.part[tid] = accI;
    }
  }
  // --- (2) split sums: items = (cell, slice) ---------------------------------
  const Split sp = make_split(C, T);
  const int cell = tid % sp.Cp, slice = tid / sp.Cp;
  if (slice < sp.S && cell < C) {
    const int i = i0 + cell;
    // QM2(i,j) = sum_a qm[a][i] * qm1[d-1-a][i+1+a], a = TURN+1 .. d-2-TURN; the split k=i+1+a may not be the nick
    const int cntM = d - 2 * TURN - 2;   // number of terms
    double accM = 0., accQ = 0.;
    if (cntM > 0 && !(c.dbg & 2)) {
      const int skip = c.cp > 0 ? c.cp - 1 - i - (TURN + 1) : -1;  // a = cp-1-i  <=> k = cp
      accM = strided_dot(tabp(c, T_QM) + (size_t)(TURN + 1) * ld + i, ld,
                         tabp(c, T_QM1) + (size_t)(d - 2 - TURN) * ld + i + TURN + 2, 1 - ld, cntM, slice, sp.S, skip);
    }
    // sum_a q[a][i] * qq[d-1-a][i+1+a], a = 0 .. d-2-TURN
    const int cntQ = d - 1 - TURN;
    if (cntQ > 0 && !(c.dbg & 2))
      accQ = strided_dot(tabp(c, T_Q) + i, ld, tabp(c, T_QQ) + (size_t)(d - 1) * ld + i + 1, 1 - ld, cntQ, slice, sp.S, -1);
    sh.part[T + tid] = accM;
    sh.part[2 * T + tid] = accQ;
  }
}

RP_HD void inside_B(Ctx& c, const Shared& sh, int d, int i0, int C, int tid) {
  const DevModel& M = *c.M;
  const int T = sh.T;
  const Split sp = make_split(C, T);
  if (tid >= sp.Cp || tid >= C) return;
  const int i = i0 + tid, j = i + d, n = c.n;
  double sI = 0., sM = 0., sQ = 0.;
  for (int s = 0; s < sp.S; s++) {
    sM += sh.part[T + s * sp.Cp + tid];
    sQ += sh.part[2 * T + s * sp.Cp + tid];
  }
  const int type = pair_type(base(c, i), base(c, j));
  const double* scale = vecp(c, V_SCALE);
  double qb = 0.;
  if (type) {
    if (d - (TURN + 1) >= 2) {
      const ISplit is = make_isplit(c, d, i0, C, T);
      const int r = (int)posp(c)[(size_t)d * c.ld + i] - is.lo;
      for (int s = 0; s < is.SI; s++) sI += sh.part[s * is.cntp + r];
    }
    if (ss(c, i, j)) qb += hairpin(c, i, j, type);
    qb += sI;
    if (ss(c, i, i + 1) && ss(c, j - 1, j))
      qb += TB(c, T_QM2, d - 2, i + 1) * M.expMLclosing * ml_stem(M, rtype(type), base(c, j - 1), base(c, i + 1)) * scale[2];
    if (!ss(c, i, j)) {
      // the loop that contains the nick is an exterior loop
      double t = scale[2];
      if (i + 1 <= c.cp - 1) t *= TB(c, T_Q, c.cp - 2 - i, i + 1);
      if (c.cp <= j - 1) t *= TB(c, T_Q, j - 1 - c.cp, c.cp);
      t *= ext_stem(M, rtype(type), ss(c, j - 1, j) ? base(c, j - 1) : -1, ss(c, i, i + 1) ? base(c, i + 1) : -1);
      qb += t;
    }
  }
  TB(c, T_QB, d, i) = qb;
  TB(c, T_QM2, d, i) = sM;
  double fI = 0., f1 = 0., fA = 0.;
  if (type && qb != 0.) {
    const int t2 = rtype(type), sq1 = j < n ? base(c, j + 1) : 0, sp1 = i > 1 ? base(c, i - 1) : 0;
    fI = qb * M.mmI[t2][sq1][sp1];
    f1 = qb * M.mm1n[t2][sq1][sp1];
    fA = type > 2 ? qb * M.expTermAU : qb;
  }
  TB(c, T_QBI, d, i) = fI;
  TB(c, T_QB1N, d, i) = f1;
  TB(c, T_QBAU, d, i) = fA;
  // qm1: one stem starting at i, unpaired to its right
  double qm1 = ss(c, j - 1, j) ? TB(c, T_QM1, d - 1, i) * M.mlb1 : 0.;
  if (type && ss(c, i - 1, i) && ss(c, j, j + 1))
    qm1 += qb * ml_stem(M, type, i > 1 ? base(c, i - 1) : -1, j < n ? base(c, j + 1) : -1);
  TB(c, T_QM1, d, i) = qm1;
  // U(i,j) = sum_k mlb[k-i] qm1(k,j) = mlb1*(qm1(i+1,j) + U(i+1,j)), cut at the nick
  const double* Uprev = vecp(c, (d & 1) ? V_U0 : V_U1);
  double* Ucur = vecp(c, (d & 1) ? V_U1 : V_U0);
  const double U = ss(c, i, i + 1) ? M.mlb1 * (TB(c, T_QM1, d - 1, i + 1) + Uprev[i + 1]) : 0.;
  Ucur[i] = U;
  TB(c, T_QM, d, i) = qm1 + sM + U;
  // exterior
  double qq = TB(c, T_QQ, d - 1, i) * M.scale1;
  if (type)
    qq += qb * ext_stem(M, type, (i > 1 && ss(c, i - 1, i)) ? base(c, i - 1) : -1,
                        (j < n && ss(c, j, j + 1)) ? base(c, j + 1) : -1);
  TB(c, T_QQ, d, i) = qq;
  TB(c, T_Q, d, i) = scale[d + 1] + qq + sQ;
}

RP_HD void inside_end(Ctx& c) { c.invZ = 1.0 / TB(c, T_Q, c.n - 1, 1); }

// ---------------------------------------------------------------------------
// outside pass, diagonal d from n-1 down to TURN+1.
// out(k,l) = Z_outside(k,l)/Z  (ViennaRNA's probs[] before the final *qb).
// ---------------------------------------------------------------------------
// two-strand only, twice per diagonal before outside_A: the closing pairs that
// straddle the nick feed the stems sitting directly in the nicked loop.
//   Qr(r)    = sum_{p<cp} out(p,r) ExtClose(p,r) scale[2] q(p+1,cp-1)        complete after diag r-cp+1
//   Qrout(l) = sum_{r>l} Qr(r) q(l+1,r-1)
//   Ql(p)    = sum_{r>=cp} out(p,r) ExtClose(p,r) scale[2] q(cp,r-1)         complete after diag cp-p
//   Qlout(k) = sum_{p<k} Ql(p) q(p+1,k-1)
// Step d finalises Qr(d+cp), Qrout(d+cp-1), Ql(cp-1-d), Qlout(cp-d).  Each sum
// is split over 32 threads (fixed partition => deterministic), partials in sh.red.
RP_HD double nick_close(const Ctx& c, int p, int r) {
  const int tp = pair_type(base(c, p), base(c, r));
  if (!tp || r - p <= TURN) return 0.;
  const double o = TB(c, T_OUT, r - p, p);
  if (o == 0.) return 0.;
  return o * vecp(c, V_SCALE)[2] *
         ext_stem(*c.M, rtype(tp), ss(c, r - 1, r) ? base(c, r - 1) : -1, ss(c, p, p + 1) ? base(c, p + 1) : -1);
}
RP_HD void outside_nick1(Ctx& c, const Shared& sh, int d, int tid) {
  if (c.cp <= 0) return;
  const int n = c.n, cp = c.cp;
  for (int w = tid; w < 128; w += sh.T) {
    const int lane = w & 31, job = w >> 5;
    double s = 0.;
    if (job == 0) {          // Qr(r), r = d+cp: closing pairs (p,r), all of diag >= d+1
      const int r = d + cp;
      if (r >= cp && r <= n)
        for (int p = 1 + lane; p < cp; p += 32) {
          const double v = nick_close(c, p, r);
          if (v != 0.) s += v * (p + 1 <= cp - 1 ? TB(c, T_Q, cp - 2 - p, p + 1) : 1.0);
        }
    } else if (job == 1) {   // Ql(p), p = cp-1-d
      const int p = cp - 1 - d;
      if (p >= 1 && p < cp)
        for (int rr = cp + lane; rr <= n; rr += 32) {
          const double v = nick_close(c, p, rr);
          if (v != 0.) s += v * (cp <= rr - 1 ? TB(c, T_Q, rr - 1 - cp, cp) : 1.0);
        }
    } else if (job == 2) {   // part of Qrout(l), l = d+cp-1, that uses Qr(r), r >= l+2 (earlier steps)
      const int l = d + cp - 1;
      if (l >= cp && l < n)
        for (int r = l + 2 + lane; r <= n; r += 32) s += vecp(c, V_QR)[r] * TB(c, T_Q, r - 2 - l, l + 1);
    } else {                 // part of Qlout(k), k = cp-d, that uses Ql(p), p <= k-2 (earlier steps)
      const int k = cp - d;
      if (k >= 2 && k < cp)
        for (int p = 1 + lane; p <= k - 2; p += 32) s += vecp(c, V_QL)[p] * TB(c, T_Q, k - 2 - p, p + 1);
    }
    sh.red[w] = s;
  }
}
RP_HD void outside_nick2(Ctx& c, const Shared& sh, int d, int tid) {
  if (c.cp <= 0) return;
  const int n = c.n, cp = c.cp;
  if (tid == 0) {
    double qr = 0., rest = 0.;
    for (int t = 0; t < 32; t++) { qr += sh.red[t]; rest += sh.red[64 + t]; }
    const int r = d + cp, l = d + cp - 1;
    if (r >= cp && r <= n) vecp(c, V_QR)[r] = qr;
    // Qrout(l) = Qr(l+1)*q(l+1,l) + rest, q of the empty segment is 1
    if (l >= cp && l < n) vecp(c, V_QROUT)[l] = qr + rest;
  } else if (tid == 32 || (sh.T <= 32 && tid == 1)) {
    double ql = 0., rest = 0.;
    for (int t = 0; t < 32; t++) { ql += sh.red[32 + t]; rest += sh.red[96 + t]; }
    const int p = cp - 1 - d, k = cp - d;
    if (p >= 1 && p < cp) vecp(c, V_QL)[p] = ql;
    if (k >= 2 && k < cp) vecp(c, V_QLOUT)[k] = ql + rest;
  }
}

// sh.part: [3][T] doubles (interior, PR, ML-left)
RP_HD void outside_A(const Ctx& c, const Shared& sh, int d, int i0, int C, int tid) {
  const DevModel& M = *c.M;
  const int T = sh.T, n = c.n;
  const long ld = c.ld;
  // --- (1) interior loops seen from the inner pair (k,l): items = (pairable cell, slice)
  const int ddmax = n - 1 - d < MAXLOOP + 2 ? n - 1 - d : MAXLOOP + 2;
  if (ddmax >= 2) {
    const ISplit is = make_isplit(c, d, i0, C, T);
    const int r = tid % is.cntp, sl = tid / is.cntp;
    if (sl < is.SI && r < is.cnt) {
      const int k = listp(c)[(size_t)d * ld + is.lo + r], l = k + d;
      double accI = 0.;
      // enclosing pair (i,j) = (k-po, l+1+u2)
      int maxpo = k - 1, maxu2 = n - l - 1;
      if (c.cp > 0) {
        if (k >= c.cp && k - c.cp < maxpo) maxpo = k - c.cp;        // i must stay on k's strand
        if (l < c.cp && c.cp - 2 - l < maxu2) maxu2 = c.cp - 2 - l;  // j must stay on l's strand
      }
      if (maxpo >= 1 && maxu2 >= 0 && TB(c, T_QB, d, k) != 0.) {
        const int type = pair_type(base(c, k), base(c, l));
        const int t2 = rtype(type), sp1 = base(c, k - 1), sq1 = base(c, l + 1);
        int u1max = ddmax - 2 < MAXLOOP ? ddmax - 2 : MAXLOOP;
        if (maxpo - 1 < u1max) u1max = maxpo - 1;
        double sI = 0., s1 = 0., sA = 0.;
        if (!(c.dbg & 1)) interior_rows<-1>(sh, tabp(c, T_OUTI), tabp(c, T_OUT1N), tabp(c, T_OUTAU), d * (int)ld + k, (int)ld, u1max,
                          maxu2, ddmax, sl, is.SI, sI, s1, sA);
        accI = M.mmI[t2][sq1][sp1] * sI + M.mm1n[t2][sq1][sp1] * s1 + (type > 2 ? M.expTermAU : 1.0) * sA;
        for (int s = sl; s < RP_N_SPECIAL; s += is.SI) {
          int u1, u2;
          special_uv(s, u1, u2);
          const int dd = u1 + u2 + 2;
          if (dd > ddmax || u1 + 1 > maxpo || u2 > maxu2) continue;
          const int i = k - 1 - u1, j = l + 1 + u2;
          const int t1 = pair_type(base(c, i), base(c, j));
          if (!t1) continue;
          const double o = TB(c, T_OUT, d + dd, i);
          if (o == 0.) continue;
          accI += o * special_loop(M, s, t1, t2, base(c, i + 1), base(c, j - 1), sp1, sq1);
        }
      }
      sh.part[tid] = accI;
    }
  }
  // --- (2) multiloop sums: items = (cell, slice) --------------------------------
  const Split sp = make_split(C, T);
  const int cell = tid % sp.Cp, slice = tid / sp.Cp;
  if (slice < sp.S && cell < C) {
    const int k = i0 + cell, l = k + d;
    double accP = 0., accL = 0.;
    // PR(k,l) = sum_b Mc[d+2+b][k] * qm[b][l+1], b = TURN+1 .. n-l-2   [k is the closing 5' end]
    if (l + 2 <= n && ss(c, l, l + 1)) {
      const int cnt = n - l - 2 - TURN;
      if (cnt > 0 && !(c.dbg & 2))
        accP = strided_dot(tabp(c, T_MC) + (size_t)(d + 3 + TURN) * ld + k, ld,
                           tabp(c, T_QM) + (size_t)(TURN + 1) * ld + l + 1, ld, cnt, slice, sp.S, -1);
    }
    // ML-left(k,l) = sum_cc PRML[d+2+cc][k-2-cc] * qm[cc][k-1-cc], cc = TURN+1 .. k-3
    if (l < n && k > 2 && ss(c, k - 1, k) && ss(c, l, l + 1) && posp(c)[(size_t)d * ld + k + 1] != posp(c)[(size_t)d * ld + k]) {
      const int cnt = k - 3 - TURN;
      if (cnt > 0 && !(c.dbg & 2) && TB(c, T_QB, d, k) != 0.)
        accL = strided_dot(tabp(c, T_PRML) + (size_t)(d + 3 + TURN) * ld + k - 3 - TURN, ld - 1,
                           tabp(c, T_QM) + (size_t)(TURN + 1) * ld + k - 2 - TURN, ld - 1, cnt, slice, sp.S, -1);
    }
    sh.part[T + tid] = accP;
    sh.part[2 * T + tid] = accL;
  }
}

RP_HD void outside_B(Ctx& c, const Shared& sh, int d, int i0, int C, int tid) {
  const DevModel& M = *c.M;
  const int T = sh.T;
  const Split sp = make_split(C, T);
  if (tid >= sp.Cp || tid >= C) return;
  const int k = i0 + tid, l = k + d, n = c.n;
  double sI = 0., sP = 0., sL = 0.;
  for (int s = 0; s < sp.S; s++) {
    sP += sh.part[T + s * sp.Cp + tid];
    sL += sh.part[2 * T + s * sp.Cp + tid];
  }
  const double* scale = vecp(c, V_SCALE);
  const bool mlr = l < n && ss(c, l, l + 1);  // something may follow l inside a multiloop
  // right side all unpaired: PL(k,l) = sum_{j>l} Mc(k,j) mlb^(j-l-1)
  const double PL = mlr ? TB(c, T_PL, d + 1, k) * M.mlb1 + TB(c, T_MC, d + 1, k) : 0.;
  const double PR = mlr ? sP : 0.;
  TB(c, T_PL, d, k) = PL;
  TB(c, T_PR, d, k) = PR;
  TB(c, T_PRML, d, k) = PR + PL;
  // left side all unpaired: PMLB(k,l) = sum_{i<k} PR(i,l) mlb^(k-1-i)
  double PMLB = 0.;
  if (k > 1 && ss(c, k - 1, k)) PMLB = TB(c, T_PMLB, d + 1, k - 1) * M.mlb1 + TB(c, T_PR, d + 1, k - 1);
  TB(c, T_PMLB, d, k) = PMLB;

  const int type = pair_type(base(c, k), base(c, l));
  double out = 0.;
  if (type && TB(c, T_QB, d, k) != 0.) {
    if ((n - 1 - d) >= 2) {
      const ISplit is = make_isplit(c, d, i0, C, T);
      const int r = (int)posp(c)[(size_t)d * c.ld + k] - is.lo;
      for (int s = 0; s < is.SI; s++) sI += sh.part[s * is.cntp + r];
    }
    const double q5 = k > 1 ? TB(c, T_Q, k - 2, 1) : 1.0;
    const double q3 = l < n ? TB(c, T_Q, n - l - 1, l + 1) : 1.0;
    out = q5 * q3 * c.invZ *
          ext_stem(M, type, (k > 1 && ss(c, k - 1, k)) ? base(c, k - 1) : -1, (l < n && ss(c, l, l + 1)) ? base(c, l + 1) : -1);
    out += sI;
    if (mlr && k > 1 && ss(c, k - 1, k)) out += (PMLB + sL) * ml_stem(M, type, base(c, k - 1), base(c, l + 1)) * scale[2];
    if (c.cp > 0) {
      if (k >= c.cp) {
        const double qo = vecp(c, V_QROUT)[l];
        if (qo != 0.)
          out += qo * (k > c.cp ? TB(c, T_Q, k - 1 - c.cp, c.cp) : 1.0) *
                 ext_stem(M, type, k > c.cp ? base(c, k - 1) : -1, base(c, l + 1));
      } else if (l < c.cp) {
        const double qo = vecp(c, V_QLOUT)[k];
        if (qo != 0.)
          out += qo * (l + 1 <= c.cp - 1 ? TB(c, T_Q, c.cp - 2 - l, l + 1) : 1.0) *
                 ext_stem(M, type, base(c, k - 1), l + 1 < c.cp ? base(c, l + 1) : -1);
      }
    }
  }
  TB(c, T_OUT, d, k) = out;
  double fI = 0., f1 = 0., fA = 0., mc = 0.;
  if (out != 0.) {
    const int si1 = base(c, k + 1), sj1 = base(c, l - 1);
    fI = out * M.mmI[type][si1][sj1];
    f1 = out * M.mm1n[type][si1][sj1];
    fA = type > 2 ? out * M.expTermAU : out;
    if (ss(c, k, k + 1) && ss(c, l - 1, l)) mc = out * M.expMLclosing * ml_stem(M, rtype(type), sj1, si1);
  }
  TB(c, T_OUTI, d, k) = fI;
  TB(c, T_OUT1N, d, k) = f1;
  TB(c, T_OUTAU, d, k) = fA;
  TB(c, T_MC, d, k) = mc;
}

// ---------------------------------------------------------------------------
// unpaired windows (single strand): up(i,d) = P(i..i+d unpaired), d < max_w
// ---------------------------------------------------------------------------
// U1: DG(p,o) = out(p,o)*hairpin(p,o)   (loop whose unpaired run is (p,o))
RP_HD void unstru_hairpin(Ctx& c, int tid, int T) {
  const int n = c.n;
  const size_t total = (size_t)n * c.ld;
  for (size_t x = tid; x < total; x += T) {
    const int d = (int)(x / c.ld), i = (int)(x % c.ld);
    if (i < 1 || i + d > n) continue;
    double v = 0.;
    if (d > TURN) {
      const double o = TB(c, T_OUT, d, i);
      if (o != 0.) v = o * hairpin(c, i, i + d, pair_type(base(c, i), base(c, i + d)));
    }
    TB(c, T_DG, d, i) = v;
  }
}
// U2 (side=0): DG(p,k) += weight of all interior loops closed by some (p,o) with inner pair (k,l): 5' gap (p,k)
// U3 (side=1): DG(l,o) += the same loops seen from their 3' gap (l,o)
// One item per gap.  For the factorised classes the sum over the free pair end is a dot
// product of two table rows that already carry the pair factors:
//   side 0:  sum_l  outX(p, l+1+u2) * qbX(k, l)          (both advance one diagonal per l)
//   side 1:  sum_p  outX(p, o)      * qbX(p+1+u1, l)     (both step one diagonal down, one cell right)
// Up to 8 dot products that share the streamed operand A:
//   acc[t] += sum_{x<cnt} A[x*sa] * B[t*tb + x*sb],  t < nu
// The B streams of neighbouring t overlap (a sliding window), so all but one of the B loads of
// an iteration hit L1: per 8 FMAs only two new values travel from L2.
RP_HD void multi_dot(const double* A, int sa, const double* B, int sb, int tb, int cnt, int nu, double* acc) {
  for (int x = 0; x < cnt; x++, A += sa, B += sb) {
    const double a = *A;
#pragma unroll
    for (int t = 0; t < 8; t++)
      if (t < nu) acc[t] += a * B[t * tb];
  }
}

RP_HD int special_index(int u1, int u2) {
  // inverse of special_uv: (0,0)->0 (1,0)->1 (0,1)->2 (1,1)->3 (1,2)->4 (2,1)->5 (2,2)->6 (2,3)->7 (3,2)->8
  if (u1 == 0) return u2 == 0 ? 0 : 2;
  if (u1 == 1) return u2 == 0 ? 1 : (u2 == 1 ? 3 : 4);
  if (u1 == 2) return u2 == 1 ? 5 : (u2 == 2 ? 6 : 7);
  return 8;
}
RP_HD void unstru_gaps(Ctx& c, int side, int tid, int T) {
  const DevModel& M = *c.M;
  const int n = c.n, ld = c.ld;
  const int items = (c.dbg & 4) ? 0 : n * (MAXLOOP + 1);
  const int tabO[3] = {T_OUTI, T_OUT1N, T_OUTAU};
  const int tabQ[3] = {T_QBI, T_QB1N, T_QBAU};
  for (int x = tid; x < items; x += T) {
    const int ug = x / n;      // size of the gap this item owns
    const int a = x % n + 1;   // gap is the open interval (a, a+ug+1)
    const int b = a + ug + 1;
    if (b > n || ug < 1) continue;  // an empty gap cannot contain a window
    double acc = 0.;
    if (side == 0) {
      const int p = a, k = b, u1 = ug;
      const int lmin = k + TURN + 1;
      for (int u2 = 0; u1 + u2 <= MAXLOOP;) {
        if (n - 1 - u2 < lmin) break;
        const int cls = M.gcls[u1][u2];
        if (cls != CLS_SPECIAL) {
          // run of up to 8 consecutive u2 of the same class
          int nu = 1;
          while (nu < 8 && u1 + u2 + nu <= MAXLOOP && M.gcls[u1][u2 + nu] == cls && n - 1 - (u2 + nu) >= lmin) nu++;
          double av[8] = {0., 0., 0., 0., 0., 0., 0., 0.};
          const double* Q = tabp(c, tabQ[cls]) + (size_t)(lmin - k) * ld + k;           // qbX(k,l), one diagonal per l
          const double* O = tabp(c, tabO[cls]) + (size_t)(lmin + 1 + u2 - p) * ld + p;  // outX(p,l+1+u2), +ld per u2
          const int cmain = n - 1 - (u2 + nu - 1) - lmin + 1;  // l range valid for every u2 of the run
          multi_dot(Q, ld, O, ld, ld, cmain, nu, av);
          for (int t = 0; t < nu; t++) {
            // the shorter shifts reach further: l up to n-1-(u2+t)
            const int cnt = n - 1 - (u2 + t) - lmin + 1;
            if (cnt > cmain) av[t] += dot_range(Q, ld, O + (size_t)t * ld, ld, cmain, cnt, 0, 1);
            acc += M.gfull[u1][u2 + t] * av[t];
          }
          u2 += nu;
        } else {
          const int lmax = n - 1 - u2;
          const int sidx = special_index(u1, u2);
          const int sp1 = base(c, k - 1), si1 = base(c, p + 1);
          for (int l = lmin; l <= lmax; l++) {
            const double qb = TB(c, T_QB, l - k, k);
            if (qb == 0.) continue;
            const int o = l + 1 + u2;
            const double ou = TB(c, T_OUT, o - p, p);
            if (ou == 0.) continue;
            acc += ou * qb * special_loop(M, sidx, pair_type(base(c, p), base(c, o)), rtype(pair_type(base(c, k), base(c, l))),
                                          si1, base(c, o - 1), sp1, base(c, l + 1));
          }
          u2++;
        }
      }
    } else {
      const int l = a, o = b, u2 = ug;
      for (int u1 = 0; u1 + u2 <= MAXLOOP;) {
        if (l - TURN - 2 - u1 < 1) break;  // k = p+1+u1 <= l-TURN-1 needs p >= 1
        const int cls = M.gcls[u1][u2];
        if (cls != CLS_SPECIAL) {
          int nu = 1;
          while (nu < 8 && u1 + nu + u2 <= MAXLOOP && M.gcls[u1 + nu][u2] == cls && l - TURN - 2 - (u1 + nu) >= 1) nu++;
          double av[8] = {0., 0., 0., 0., 0., 0., 0., 0.};
          const double* O = tabp(c, tabO[cls]) + (size_t)(o - 1) * ld + 1;               // outX(p,o): one diagonal down, one cell right per p
          const double* Q = tabp(c, tabQ[cls]) + (size_t)(l - 2 - u1) * ld + 2 + u1;     // qbX(p+1+u1,l); per u1: 1-ld
          const int cmain = l - TURN - 2 - (u1 + nu - 1);  // p = 1..cmain valid for every u1 of the run
          multi_dot(O, 1 - ld, Q, 1 - ld, 1 - ld, cmain, nu, av);
          for (int t = 0; t < nu; t++) {
            const int cnt = l - TURN - 2 - (u1 + t);
            if (cnt > cmain) av[t] += dot_range(O, 1 - ld, Q + (long)t * (1 - ld), 1 - ld, cmain, cnt, 0, 1);
            acc += M.gfull[u1 + t][u2] * av[t];
          }
          u1 += nu;
        } else {
          const int pmax = l - TURN - 2 - u1;
          const int sidx = special_index(u1, u2);
          const int sq1 = base(c, l + 1), sj1 = base(c, o - 1);
          for (int p = 1; p <= pmax; p++) {
            const double ou = TB(c, T_OUT, o - p, p);
            if (ou == 0.) continue;
            const int k = p + 1 + u1;
            const double qb = TB(c, T_QB, l - k, k);
            if (qb == 0.) continue;
            acc += ou * qb * special_loop(M, sidx, pair_type(base(c, p), base(c, o)), rtype(pair_type(base(c, k), base(c, l))),
                                          base(c, p + 1), sj1, base(c, k - 1), sq1);
          }
          u1++;
        }
      }
    }
    TB(c, T_DG, b - a, a) += acc;
  }
}
// U4a: suffix sums over b for each a ; U4b: prefix sums over a for each b
RP_HD void unstru_dom_rows(Ctx& c, int tid, int T) {
  const int n = c.n;
  for (int a = 1 + tid; a <= n; a += T) {
    double s = 0.;
    for (int b = n; b > a; b--) {
      s += TB(c, T_DG, b - a, a);
      TB(c, T_DG, b - a, a) = s;
    }
  }
}
RP_HD void unstru_dom_cols(Ctx& c, int tid, int T) {
  const int n = c.n;
  for (int b = 2 + tid; b <= n; b += T) {
    double s = 0.;
    for (int a = 1; a < b; a++) {
      s += TB(c, T_DG, b - a, a);
      TB(c, T_DG, b - a, a) = s;
    }
  }
}
// U5: RR(p,j) = sum_{o>=j+2} Mc(p,o) QM2(j+1,o-1)
//     LL(i,o) = sum_{p<=i-2} Mc(p,o) QM2(p+1,i-1)
//     XX(i,o) = sum_{p<=i-2} Mc(p,o) qm (p+1,i-1)
RP_HD void unstru_ml_tables(Ctx& c, int tid, int T) {
  const int n = c.n;
  const size_t total = (size_t)n * c.ld;
  const double* MC = tabp(c, T_MC);
  const double* QM2 = tabp(c, T_QM2);
  const double* QM = tabp(c, T_QM);
  for (size_t x = tid; x < total; x += T) {
    const int e = (int)(x / c.ld), i = (int)(x % c.ld);
    if (i < 1 || i + e > n) continue;
    const int o = i + e;  // cell (i,o); also (p,j) for RR
    double r = 0., l2 = 0., l1 = 0.;
    for (int b = 2 * TURN + 3; b <= n - o - 2; b++) r += MC[(size_t)(e + 2 + b) * c.ld + i] * QM2[(size_t)b * c.ld + o + 1];
    for (int cc = TURN + 1; cc <= i - 3; cc++) {
      const double m = MC[(size_t)(e + 2 + cc) * c.ld + i - 2 - cc];
      l2 += m * QM2[(size_t)cc * c.ld + i - 1 - cc];
      l1 += m * QM[(size_t)cc * c.ld + i - 1 - cc];
    }
    TB(c, T_RR, e, i) = r;
    TB(c, T_LL, e, i) = l2;
    TB(c, T_XX, e, i) = l1;
  }
}
// U6: assemble; writes fp32 in the reference layout up[(i-1)*max_w + d]
RP_HD void unstru_windows(Ctx& c, float* up, int tid, int T) {
  const int n = c.n, w = c.max_w;
  const double* scale = vecp(c, V_SCALE);
  const double* mlb = vecp(c, V_MLB);
  for (int x = tid; x < n * w; x += T) {
    const int i = x / w + 1, dd = x % w, j = i + dd;
    double v = 0.;
    if (j <= n) {
      const double q5 = i > 1 ? TB(c, T_Q, i - 2, 1) : 1.0;
      const double q3 = j < n ? TB(c, T_Q, n - j - 1, j + 1) : 1.0;
      v = q5 * scale[dd + 1] * q3 * c.invZ;
      if (i > 1 && j < n) {
        v += TB(c, T_DG, j - i + 2, i - 1);
        double m1 = 0., m2 = 0., m3 = 0.;
        for (int p = 1; p < i; p++) m1 += mlb[j - p] * TB(c, T_RR, j - p, p);
        for (int o = j + 1; o <= n; o++) m2 += mlb[o - i] * TB(c, T_LL, o - i, i);
        for (int o = j + 2 + TURN + 1; o <= n; o++) m3 += TB(c, T_QM, o - j - 2, j + 1) * TB(c, T_XX, o - i, i);
        // Mc carries no scale factor for the closing pair's two bases: apply it here
        v += (m1 + m2 + m3 * mlb[dd + 1]) * scale[2];
      }
    }
    up[x] = (float)v;
  }
}

// ---------------------------------------------------------------------------
// outputs in the reference's layouts
// ---------------------------------------------------------------------------
// bp[offset[i]+j] = (float) pr(i,j), offset[i] = i*(2L+1-i)/2   (src/ractip.cpp:314-317,365-367)
RP_HD void write_bp(const Ctx& c, float* bp, int tid, int T) {
  const int L = c.n;
  const size_t total = (size_t)(L + 1) * (L + 2) / 2;
  for (size_t x = tid; x < total; x += T) bp[x] = 0.f;
}
RP_HD void write_bp2(const Ctx& c, float* bp, int tid, int T) {
  const int L = c.n;
  const size_t total = (size_t)L * c.ld;
  for (size_t x = tid; x < total; x += T) {
    const int d = (int)(x / c.ld), i = (int)(x % c.ld);
    if (i < 1 || i + d > L || d < 1) continue;
    const double p = d > TURN ? TB(c, T_OUT, d, i) * TB(c, T_QB, d, i) : 0.;
    bp[(size_t)i * (2 * L + 1 - i) / 2 + (i + d)] = (float)p;
  }
}
// hp[i][j-cp+1] = p if i<cp<=j and p>th_hy (float compare)   (src/ractip.cpp:404-405,447-453)
RP_HD void write_hp(const Ctx& c, float* hp, int n1, int n2, float th_hy, int tid, int T) {
  const int cp = c.cp;
  const int total = (n1 + 1) * (n2 + 1);
  for (int x = tid; x < total; x += T) {
    const int i = x / (n2 + 1), jj = x % (n2 + 1);
    float v = 0.f;
    if (i >= 1 && jj >= 1) {
      const int j = jj + cp - 1, d = j - i;
      if (d > TURN) {
        const double p = TB(c, T_OUT, d, i) * TB(c, T_QB, d, i);
        const float pf = (float)p;
        if (p >= (double)th_hy && pf > th_hy) v = pf;
      }
    }
    hp[x] = v;
  }
}

}  // namespace rp
#endif
