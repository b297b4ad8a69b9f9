/*
 * rp_oracle.c -- CPU fp64 restatement of RactIP's probability stage.
 *
 * TEST INFRASTRUCTURE ONLY.  Nothing under ractip_b200/ may include, link or
 * call this file; only tests/, __graft_entry__.smoke() and bench.py's
 * cpu_baseline / --impl reference legs use it, as the checker and the timed
 * CPU baseline, never as the product.
 *
 * PARITY STATUS: **parity unpinned**.  All arithmetic of the reference's hot
 * path lives in ViennaRNA (RNAlib2, ">= 2.2.0", version unpinned in the
 * reference: Dockerfile:15; the only concrete version named is 2.4.17 in a
 * comment, Dockerfile:17-22), which is neither vendored in /root/reference nor
 * installed in the build container.  The reference has no tests or golden
 * vectors for this path.  This file therefore restates the
 * published ViennaRNA 2.4.x algorithms (McCaskill inside/outside with dangles=2,
 * MAXLOOP=30, TURN=3; the two-strand co_pf_fold variant; pf_unstru's unpaired
 * window probabilities) from their model definition, anchored on the
 * reference's call sites:
 *     rnafold   src/ractip.cpp:308-382  (pf_fold + export_bppm + pf_unstru)
 *     rnaduplex src/ractip.cpp:384-459  (co_pf_fold + assign_plist_from_pr,
 *                                        or pf_duplex under --duplex)
 *     pf_duplex src/pf_duplex.c:34-40,67-206 (fully in-tree; restated 1:1)
 * What pins it instead: exhaustive structure enumeration (orc_enumerate below,
 * an independent evaluation of the same loop-energy model) must reproduce Z,
 * every pair marginal and every unpaired-window marginal of the DP to 1e-12
 * (tests/test_oracle_enum.py), plus the model invariants listed in SURVEY 8c.
 *
 * Conventions: sequences 1-based internally, S[i] in {0:N,1:A,2:C,3:G,4:U};
 * pair types CG=1 GC=2 GU=3 UG=4 AU=5 UA=6; rtype={0,2,1,4,3,6,5,7}.
 * cp = index of the first base of the second strand (0 = single strand);
 * ss(a,b) = "a and b on the same strand" == ViennaRNA's SAME_STRAND(a,b).
 */
#include <math.h>
#include <stdio.h>
#include <stdlib.h>
#include <string.h>
#include <ctype.h>

#include "../include/ractip_prob.h"

#define MAXLOOP RP_MAXLOOP
#define TURN RP_TURN
#define NBPAIRS RP_NBPAIRS
#define MIN2(a, b) ((a) < (b) ? (a) : (b))
#define MAX2(a, b) ((a) > (b) ? (a) : (b))

static const int RTYPE[8] = {0, 2, 1, 4, 3, 6, 5, 7};
static const int PAIR[5][5] = {
    /*      N  A  C  G  U */
    /*N*/ {0, 0, 0, 0, 0},
    /*A*/ {0, 0, 0, 0, 5},
    /*C*/ {0, 0, 0, 1, 0},
    /*G*/ {0, 0, 2, 0, 3},
    /*U*/ {0, 6, 0, 4, 0}};

/* ------------------------------------------------------------------------ */
/* parameters: integer (scale_parameters) and Boltzmann (pf) tables at 37 C  */
/* ------------------------------------------------------------------------ */
typedef struct orc_params {
  rp_model m;
  double kT; /* cal/mol */
  double pf_scale;
  /* Boltzmann factors, unscaled */
  double expstack[8][8], exphairpin[31], expbulge[31], expinternal[31];
  double expmismatchI[8][5][5], expmismatchH[8][5][5], expmismatchM[8][5][5];
  double expmismatchExt[8][5][5], expmismatch1nI[8][5][5], expmismatch23I[8][5][5];
  double expdangle5[8][5], expdangle3[8][5];
  double expint11[8][8][5][5], expint21[8][8][5][5][5], expint22[8][8][5][5][5][5];
  double expninio[MAXLOOP + 1];
  double expMLclosing, expMLintern[8], expMLbase, expTermAU, expDuplexInit, lxc;
  double exptetra[200], exptri[40], exphex[200];
  /* integer tables as scale_parameters() leaves them (pf_duplex path) */
  int dangle5[8][5], dangle3[8][5], mismatchM[8][5][5], mismatchExt[8][5][5];
  int maxloop; /* MAXLOOP (30); tests lower it so that small enumerations exercise the bound */
} orc_params;

/* ViennaRNA params.c SMOOTH(): dangles/exterior+multi mismatches are made
 * non-positive smoothly.  X is -energy in 0.01 kcal/mol. */
static double smooth(const rp_model* m, double X) {
  if (!m->pf_smooth) return X < 0 ? 0 : X; /* CLIP_NEGATIVE */
  if (X / 10. < -1.2283697) return 0;
  if (X / 10. > 0.8660254) return X;
  double s = sin(X / 10. - 0.34242663) + 1;
  return 10. * 0.38490018 * s * s;
}

orc_params* orc_params_new(const rp_model* m) {
  if (!m || m->temperature != 37.0) return NULL;
  orc_params* P = (orc_params*)calloc(1, sizeof(orc_params));
  P->m = *m;
  P->maxloop = MAXLOOP;
  double kT = (m->temperature + RP_K0) * RP_GASCONST;
  P->kT = kT;
  /* pf_scale = -1 (src/ractip.cpp:325,392,442) => ViennaRNA's estimate from
   * the mean energy of random sequences, -185 cal/nt at 37 C, times sfact. */
  P->pf_scale = exp(-(m->sfact * (-185.0 + (m->temperature - 37.) * 7.27)) / kT);
  if (P->pf_scale < 1.) P->pf_scale = 1.;
#define BF(E) exp(-(double)(E) * 10. / kT)
  for (int i = 0; i < 31; i++) {
    P->exphairpin[i] = BF(m->hairpin37[i]);
    P->expbulge[i] = BF(m->bulge37[i]);
    P->expinternal[i] = BF(m->internal_loop37[i]);
  }
  P->lxc = m->lxc37;
  P->expDuplexInit = BF(m->DuplexInit37);
  P->expMLclosing = BF(m->ML_closing37);
  for (int i = 0; i <= NBPAIRS; i++) P->expMLintern[i] = BF(m->ML_intern37);
  P->expTermAU = BF(m->TerminalAU37);
  P->expMLbase = BF(m->ML_BASE37);
  for (int j = 0; j <= MAXLOOP; j++) P->expninio[j] = BF(MIN2(m->MAX_NINIO, j * m->ninio37));
  for (int i = 0; (size_t)(i * 7) < strlen(m->Tetraloops) && i < 200; i++) P->exptetra[i] = BF(m->Tetraloop37[i]);
  for (int i = 0; (size_t)(i * 6) < strlen(m->Triloops) && i < 40; i++) P->exptri[i] = BF(m->Triloop37[i]);
  for (int i = 0; (size_t)(i * 9) < strlen(m->Hexaloops) && i < 200; i++) P->exphex[i] = BF(m->Hexaloop37[i]);
  for (int i = 0; i <= NBPAIRS; i++)
    for (int j = 0; j <= NBPAIRS; j++) P->expstack[i][j] = BF(m->stack37[i][j]);
  for (int i = 0; i <= NBPAIRS; i++)
    for (int j = 0; j < 5; j++) {
      if (m->dangles) {
        P->expdangle5[i][j] = exp(smooth(m, -(double)m->dangle5_37[i][j]) * 10. / kT);
        P->expdangle3[i][j] = exp(smooth(m, -(double)m->dangle3_37[i][j]) * 10. / kT);
      } else {
        P->expdangle5[i][j] = P->expdangle3[i][j] = 1.;
      }
      P->dangle5[i][j] = m->dangle5_37[i][j] > 0 ? 0 : m->dangle5_37[i][j];
      P->dangle3[i][j] = m->dangle3_37[i][j] > 0 ? 0 : m->dangle3_37[i][j];
      for (int k = 0; k < 5; k++) {
        P->expmismatchI[i][j][k] = BF(m->mismatchI37[i][j][k]);
        P->expmismatch1nI[i][j][k] = BF(m->mismatch1nI37[i][j][k]);
        P->expmismatchH[i][j][k] = BF(m->mismatchH37[i][j][k]);
        P->expmismatch23I[i][j][k] = BF(m->mismatch23I37[i][j][k]);
        if (m->dangles) {
          P->expmismatchM[i][j][k] = exp(smooth(m, -(double)m->mismatchM37[i][j][k]) * 10. / kT);
          P->expmismatchExt[i][j][k] = exp(smooth(m, -(double)m->mismatchExt37[i][j][k]) * 10. / kT);
          P->mismatchM[i][j][k] = m->mismatchM37[i][j][k] > 0 ? 0 : m->mismatchM37[i][j][k];
          P->mismatchExt[i][j][k] = m->mismatchExt37[i][j][k] > 0 ? 0 : m->mismatchExt37[i][j][k];
        } else {
          P->expmismatchM[i][j][k] = P->expmismatchExt[i][j][k] = 1.;
        }
      }
    }
  for (int i = 0; i <= NBPAIRS; i++)
    for (int j = 0; j <= NBPAIRS; j++)
      for (int k = 0; k < 5; k++)
        for (int l = 0; l < 5; l++) {
          P->expint11[i][j][k][l] = BF(m->int11_37[i][j][k][l]);
          for (int a = 0; a < 5; a++) {
            P->expint21[i][j][k][l][a] = BF(m->int21_37[i][j][k][l][a]);
            for (int b = 0; b < 5; b++) P->expint22[i][j][k][l][a][b] = BF(m->int22_37[i][j][k][l][a][b]);
          }
        }
#undef BF
  return P;
}

void orc_params_free(orc_params* P) { free(P); }
void orc_set_maxloop(orc_params* P, int maxloop) { if (P && maxloop >= 0 && maxloop <= MAXLOOP) P->maxloop = maxloop; }
double orc_pf_scale(const orc_params* P) { return P->pf_scale; }

/* ------------------------------------------------------------------------ */
/* loop weights (ViennaRNA loop_energies.h, 2.4.x semantics)                 */
/* ------------------------------------------------------------------------ */
/* hairpin closed by (i,j); str points at the character of position i */
static double exp_hairpin(const orc_params* P, int u, int type, int si1, int sj1, const char* str) {
  double q;
  if (u <= 30) q = P->exphairpin[u];
  else q = P->exphairpin[30] * exp(-(P->lxc * log(u / 30.)) * 10. / P->kT);
  if (u < 3) return q;
  if (P->m.special_hp) {
    if (u == 4) {
      char tl[7] = {0};
      const char* ts;
      strncpy(tl, str, 6);
      if ((ts = strstr(P->m.Tetraloops, tl))) {
        if (type != 7) return P->exptetra[(ts - P->m.Tetraloops) / 7];
        q *= P->exptetra[(ts - P->m.Tetraloops) / 7];
      }
    } else if (u == 6) {
      char tl[9] = {0};
      const char* ts;
      strncpy(tl, str, 8);
      if ((ts = strstr(P->m.Hexaloops, tl))) return P->exphex[(ts - P->m.Hexaloops) / 9];
    } else if (u == 3) {
      char tl[6] = {0};
      const char* ts;
      strncpy(tl, str, 5);
      if ((ts = strstr(P->m.Triloops, tl))) return P->exptri[(ts - P->m.Triloops) / 6];
      return type > 2 ? q * P->expTermAU : q;
    }
  }
  return q * P->expmismatchH[type][si1][sj1];
}

/* interior loop: closing (i,j) of `type`, inner pair given as type2 =
 * rtype[type(k,l)]; si1=S[i+1], sj1=S[j-1], sp1=S[k-1], sq1=S[l+1] */
static double exp_intloop(const orc_params* P, int u1, int u2, int type, int type2, int si1, int sj1, int sp1, int sq1) {
  int ul = u1 > u2 ? u1 : u2, us = u1 > u2 ? u2 : u1;
  if (ul == 0) return P->expstack[type][type2];
  if (us == 0) {
    double z = P->expbulge[ul];
    if (ul == 1) z *= P->expstack[type][type2];
    else {
      if (type > 2) z *= P->expTermAU;
      if (type2 > 2) z *= P->expTermAU;
    }
    return z;
  }
  if (us == 1) {
    if (ul == 1) return P->expint11[type][type2][si1][sj1];
    if (ul == 2) {
      if (u1 == 1) return P->expint21[type][type2][si1][sq1][sj1];
      return P->expint21[type2][type][sq1][si1][sp1];
    }
    return P->expinternal[ul + us] * P->expmismatch1nI[type][si1][sj1] * P->expmismatch1nI[type2][sq1][sp1] * P->expninio[ul - us];
  }
  if (us == 2) {
    if (ul == 2) return P->expint22[type][type2][si1][sp1][sq1][sj1];
    if (ul == 3) return P->expinternal[5] * P->expmismatch23I[type][si1][sj1] * P->expmismatch23I[type2][sq1][sp1] * P->expninio[1];
  }
  return P->expinternal[ul + us] * P->expmismatchI[type][si1][sj1] * P->expmismatchI[type2][sq1][sp1] * P->expninio[ul - us];
}

static double exp_extstem(const orc_params* P, int type, int s5, int s3) {
  double e = 1.0;
  if (s5 >= 0 && s3 >= 0) e = P->expmismatchExt[type][s5][s3];
  else if (s5 >= 0) e = P->expdangle5[type][s5];
  else if (s3 >= 0) e = P->expdangle3[type][s3];
  if (type > 2) e *= P->expTermAU;
  return e;
}

static double exp_mlstem(const orc_params* P, int type, int s5, int s3) {
  double e = 1.0;
  if (s5 >= 0 && s3 >= 0) e = P->expmismatchM[type][s5][s3];
  else if (s5 >= 0) e = P->expdangle5[type][s5];
  else if (s3 >= 0) e = P->expdangle3[type][s3];
  if (type > 2) e *= P->expTermAU;
  return e * P->expMLintern[type];
}

/* integer twins (scale_parameters view), used by pf_duplex only */
static int E_intloop(const orc_params* P, int n1, int n2, int type, int type_2, int si1, int sj1, int sp1, int sq1) {
  const rp_model* m = &P->m;
  int nl = n1 > n2 ? n1 : n2, ns = n1 > n2 ? n2 : n1;
  if (nl == 0) return m->stack37[type][type_2];
  if (ns == 0) {
    int e = nl <= MAXLOOP ? m->bulge37[nl] : m->bulge37[30] + (int)(m->lxc37 * log(nl / 30.));
    if (nl == 1) e += m->stack37[type][type_2];
    else {
      if (type > 2) e += m->TerminalAU37;
      if (type_2 > 2) e += m->TerminalAU37;
    }
    return e;
  }
  if (ns == 1) {
    if (nl == 1) return m->int11_37[type][type_2][si1][sj1];
    if (nl == 2) {
      if (n1 == 1) return m->int21_37[type][type_2][si1][sq1][sj1];
      return m->int21_37[type_2][type][sq1][si1][sp1];
    }
    int e = nl + 1 <= MAXLOOP ? m->internal_loop37[nl + 1] : m->internal_loop37[30] + (int)(m->lxc37 * log((nl + 1) / 30.));
    e += MIN2(m->MAX_NINIO, (nl - ns) * m->ninio37);
    e += m->mismatch1nI37[type][si1][sj1] + m->mismatch1nI37[type_2][sq1][sp1];
    return e;
  }
  if (ns == 2) {
    if (nl == 2) return m->int22_37[type][type_2][si1][sp1][sq1][sj1];
    if (nl == 3) return m->internal_loop37[5] + m->ninio37 + m->mismatch23I37[type][si1][sj1] + m->mismatch23I37[type_2][sq1][sp1];
  }
  int u = nl + ns;
  int e = u <= MAXLOOP ? m->internal_loop37[u] : m->internal_loop37[30] + (int)(m->lxc37 * log(u / 30.));
  e += MIN2(m->MAX_NINIO, (nl - ns) * m->ninio37);
  e += m->mismatchI37[type][si1][sj1] + m->mismatchI37[type_2][sq1][sp1];
  return e;
}

static int E_extloop(const orc_params* P, int type, int si1, int sj1) {
  int e = 0;
  if (si1 >= 0 && sj1 >= 0) e += P->mismatchExt[type][si1][sj1];
  else if (si1 >= 0) e += P->dangle5[type][si1];
  else if (sj1 >= 0) e += P->dangle3[type][sj1];
  if (type > 2) e += P->m.TerminalAU37;
  return e;
}

/* ------------------------------------------------------------------------ */
/* sequence helpers                                                          */
/* ------------------------------------------------------------------------ */
static int enc(char c) {
  switch (toupper((unsigned char)c)) {
    case 'A': return 1;
    case 'C': return 2;
    case 'G': return 3;
    case 'U': case 'T': return 4;
    default: return 0;
  }
}

typedef struct {
  int n, cp;
  int* S;     /* 1..n, S[0]=S[n+1]=0 */
  char* str;  /* upper-cased copy, str[i-1] = char of position i */
} seq_t;

static void seq_init(seq_t* s, const char* seq, int n, int cp) {
  s->n = n; s->cp = cp;
  s->S = (int*)calloc(n + 2, sizeof(int));
  s->str = (char*)calloc(n + 16, 1);
  for (int i = 1; i <= n; i++) { s->S[i] = enc(seq[i - 1]); s->str[i - 1] = (char)toupper((unsigned char)seq[i - 1]); }
}
static void seq_free(seq_t* s) { free(s->S); free(s->str); }
static inline int same_strand(int cp, int a, int b) { return cp <= 0 || a >= cp || b < cp; }

/* ------------------------------------------------------------------------ */
/* McCaskill inside/outside (+ two-strand variant, + unpaired windows)       */
/* Follows the call sites src/ractip.cpp:356,359-367 (cp=0) and :444-447     */
/* (cp=|s1|+1); recurrences as in SURVEY.md Appendix A.4-A.6.                */
/* ------------------------------------------------------------------------ */
typedef struct {
  int n, cp, ld;
  double *q, *qb, *qm, *qm1, *qm2, *out; /* (n+2)x(n+2), [i][j] */
  double *scale, *mlb;                   /* scale[k]=pf_scale^-k ; mlb[k]=expMLbase^k*scale[k] */
  double Z;
} tables_t;

#define T(a, i, j) ((a)[(size_t)(i) * ld + (j)])

static void tables_alloc(tables_t* t, int n, int cp, double pf_scale, double expMLbase) {
  t->n = n; t->cp = cp; t->ld = n + 2;
  size_t sz = (size_t)(n + 2) * (n + 2);
  t->q = (double*)calloc(sz, sizeof(double));
  t->qb = (double*)calloc(sz, sizeof(double));
  t->qm = (double*)calloc(sz, sizeof(double));
  t->qm1 = (double*)calloc(sz, sizeof(double));
  t->qm2 = (double*)calloc(sz, sizeof(double));
  t->out = (double*)calloc(sz, sizeof(double));
  t->scale = (double*)calloc(n + 3, sizeof(double));
  t->mlb = (double*)calloc(n + 3, sizeof(double));
  t->scale[0] = 1.; t->mlb[0] = 1.;
  for (int k = 1; k <= n + 2; k++) { t->scale[k] = t->scale[k - 1] / pf_scale; t->mlb[k] = t->mlb[k - 1] * expMLbase / pf_scale; }
}
static void tables_free(tables_t* t) {
  free(t->q); free(t->qb); free(t->qm); free(t->qm1); free(t->qm2); free(t->out); free(t->scale); free(t->mlb);
}

static void inside(const orc_params* P, const seq_t* sq, tables_t* t) {
  const int n = t->n, cp = t->cp, ld = t->ld;
  const int* S = sq->S;
  double *q = t->q, *qb = t->qb, *qm = t->qm, *qm1 = t->qm1, *qm2 = t->qm2;
  const double* scale = t->scale;
  double* qq = (double*)calloc((size_t)(n + 2) * (n + 2), sizeof(double));
#define SS(a, b) same_strand(cp, a, b)
  for (int i = 1; i <= n + 1; i++) T(q, i, i - 1) = 1.0; /* empty segment */
  for (int d = 0; d <= TURN && d < n; d++)
    for (int i = 1; i + d <= n; i++) T(q, i, i + d) = scale[d + 1];
  for (int j = TURN + 2; j <= n; j++) {
    for (int i = j - TURN - 1; i >= 1; i--) {
      int type = PAIR[S[i]][S[j]];
      int u = j - i - 1;
      double qbt = 0.;
      /* qm2(i+1,j-1): >= 2 stems; also the multiloop-closing split sum */
      double m2 = 0.;
      for (int k = i + 2; k <= j - 1; k++)
        if (SS(k - 1, k)) m2 += T(qm, i + 1, k - 1) * T(qm1, k, j - 1);
      T(qm2, i + 1, j - 1) = m2;
      if (type) {
        if (SS(i, j)) qbt += exp_hairpin(P, u, type, S[i + 1], S[j - 1], sq->str + i - 1) * scale[u + 2];
        for (int k = i + 1; k <= MIN2(i + P->maxloop + 1, j - TURN - 2); k++) {
          int u1 = k - i - 1;
          if (!SS(i, k)) break;
          for (int l = MAX2(k + TURN + 1, j - 1 - P->maxloop + u1); l < j; l++) {
            if (!SS(l, j)) continue;
            int t2 = PAIR[S[k]][S[l]];
            if (!t2) continue;
            qbt += T(qb, k, l) * exp_intloop(P, u1, j - l - 1, type, RTYPE[t2], S[i + 1], S[j - 1], S[k - 1], S[l + 1]) * scale[u1 + j - l + 1];
          }
        }
        if (SS(i, i + 1) && SS(j - 1, j))
          qbt += m2 * P->expMLclosing * exp_mlstem(P, RTYPE[type], S[j - 1], S[i + 1]) * scale[2];
        if (!SS(i, j)) {
          /* loop containing the nick: scored as an exterior loop */
          double tmp = T(q, i + 1, cp - 1) * T(q, cp, j - 1) * scale[2];
          tmp *= exp_extstem(P, RTYPE[type], SS(j - 1, j) ? S[j - 1] : -1, SS(i, i + 1) ? S[i + 1] : -1);
          qbt += tmp;
        }
      }
      T(qb, i, j) = qbt;
      /* qm1: exactly one stem starting at i, unpaired to its right */
      double v = SS(j - 1, j) ? T(qm1, i, j - 1) * t->mlb[1] : 0.;
      if (type && SS(i - 1, i) && SS(j, j + 1))
        v += qbt * exp_mlstem(P, type, i > 1 ? S[i - 1] : -1, j < n ? S[j + 1] : -1);
      T(qm1, i, j) = v;
      /* qm: >= 1 stem */
      double tm = 0.;
      for (int k = i + 1; k <= j; k++) {
        double a = 0.;
        if (SS(k - 1, k)) a += T(qm, i, k - 1);
        if (SS(i, k)) a += t->mlb[k - i];
        tm += a * T(qm1, k, j);
      }
      T(qm, i, j) = tm + v;
      /* qq / q: exterior */
      double e = T(qq, i, j - 1) * scale[1];
      if (type) e += qbt * exp_extstem(P, type, (i > 1 && SS(i - 1, i)) ? S[i - 1] : -1, (j < n && SS(j, j + 1)) ? S[j + 1] : -1);
      T(qq, i, j) = e;
      double tq = scale[j - i + 1] + e;
      for (int k = i + 1; k <= j; k++) tq += T(q, i, k - 1) * T(qq, k, j);
      T(q, i, j) = tq;
    }
  }
  t->Z = T(q, 1, n);
  free(qq);
#undef SS
}

/* out(k,l) = (partition function of everything outside the pair (k,l)) / Z,
 * i.e. ViennaRNA's probs[] before the final multiplication by qb. */
static void outside(const orc_params* P, const seq_t* sq, tables_t* t) {
  const int n = t->n, cp = t->cp, ld = t->ld;
  const int* S = sq->S;
  double *q = t->q, *qb = t->qb, *qm = t->qm, *out = t->out;
  const double* scale = t->scale;
  const double Z = t->Z;
#define SS(a, b) same_strand(cp, a, b)
  /* Mc(i,j) = out(i,j) * expMLclosing * MLstem_close * scale[2] (0 unless the
   * closing pair may close a multiloop: boundaries i|i+1 and j-1|j intact) */
  double* prml = (double*)calloc(n + 2, sizeof(double));
  double* prm_l = (double*)calloc(n + 2, sizeof(double));
  double* prm_l1 = (double*)calloc(n + 2, sizeof(double));
  double* Ql = (double*)calloc(n + 2, sizeof(double));    /* per p  */
  double* Qr = (double*)calloc(n + 2, sizeof(double));    /* per r  */
  double* Qlout = (double*)calloc(n + 2, sizeof(double)); /* per k  */
  int ql_ready = 0;

  /* 1. exterior context */
  for (int i = 1; i <= n; i++)
    for (int j = i + TURN + 1; j <= n; j++) {
      int type = PAIR[S[i]][S[j]];
      if (type && T(qb, i, j) > 0.)
        T(out, i, j) = T(q, 1, i - 1) * T(q, j + 1, n) / Z *
                       exp_extstem(P, type, (i > 1 && SS(i - 1, i)) ? S[i - 1] : -1, (j < n && SS(j, j + 1)) ? S[j + 1] : -1);
    }

  for (int l = n; l > TURN + 1; l--) {
    /* 2. (k,l) enclosed by (i,j) in an interior loop */
    for (int k = 1; k < l - TURN; k++) {
      int t2 = PAIR[S[k]][S[l]];
      if (!t2 || T(qb, k, l) == 0.) continue;
      t2 = RTYPE[t2];
      double acc = 0.;
      for (int i = MAX2(1, k - P->maxloop - 1); i <= k - 1; i++) {
        if (!SS(i, k)) continue;
        int u1 = k - i - 1;
        for (int j = l + 1; j <= MIN2(l + P->maxloop - u1 + 1, n); j++) {
          if (!SS(l, j)) break;
          int type = PAIR[S[i]][S[j]];
          if (!type || T(out, i, j) == 0.) continue;
          acc += T(out, i, j) * scale[u1 + j - l + 1] *
                 exp_intloop(P, u1, j - l - 1, type, t2, S[i + 1], S[j - 1], S[k - 1], S[l + 1]);
        }
      }
      T(out, k, l) += acc;
    }
    /* 3. (k,l) as a stem of a multiloop closed by (i,j) */
    if (l < n && SS(l, l + 1)) {
      double prm_MLb = 0.;
      for (int k = 2; k < l - TURN; k++) {
        int i = k - 1;
        double prmt = 0., prmt1 = 0.;
        if (SS(i, i + 1)) {
          /* j = l+1: nothing right of the stem */
          int tt = PAIR[S[i]][S[l + 1]];
          if (tt && l + 1 - i > TURN)
            prmt1 = T(out, i, l + 1) * P->expMLclosing * exp_mlstem(P, RTYPE[tt], S[l], S[i + 1]);
          for (int j = l + 2; j <= n; j++) {
            if (!SS(j - 1, j)) continue;
            tt = PAIR[S[i]][S[j]];
            if (!tt || T(out, i, j) == 0.) continue;
            prmt += T(out, i, j) * exp_mlstem(P, RTYPE[tt], S[j - 1], S[i + 1]) * T(qm, l + 1, j - 1);
          }
          prmt *= P->expMLclosing;
        }
        /* prm_l[i] = sum_{j>l} out(i,j) close(i,j) expMLbase^(j-l-1): right side unpaired */
        prm_l[i] = prm_l1[i] * t->mlb[1] + prmt1;
        /* prm_MLb = sum_{i<k} prmt(i) expMLbase^(k-i-1): left side unpaired */
        if (SS(k - 1, k)) prm_MLb = prm_MLb * t->mlb[1] + prmt;
        else prm_MLb = 0.;
        /* the unpaired stretch i+1..k-1 is empty for i=k-1: boundary i|k is i|i+1, checked above */
        prml[i] = prmt + prm_l[i];
        int tk = PAIR[S[k]][S[l]];
        if (!tk || T(qb, k, l) == 0.) continue;
        if (!SS(k - 1, k)) continue; /* stem may not sit in a multiloop at the nick */
        double temp = prm_MLb;
        for (int ii = 1; ii <= k - 2; ii++)
          if (SS(ii, ii + 1)) temp += prml[ii] * T(qm, ii + 1, k - 1);
        temp *= exp_mlstem(P, tk, S[k - 1], S[l + 1]) * scale[2];
        T(out, k, l) += temp;
      }
    } else {
      for (int i = 0; i <= n; i++) prm_l[i] = 0.;
    }
    { double* tmp = prm_l1; prm_l1 = prm_l; prm_l = tmp; }

    /* 4. (k,l) directly inside the loop that contains the nick */
    if (cp > 0) {
      if (l >= cp) {
        /* Qr(l) needs out(p,l) for all p<cp: complete once steps 2-3 at this l are done;
         * it is used by stems with smaller l, so update after use. */
        /* stems (k,l) on strand 2: cp <= k < l < r */
        double Qrout = 0.;
        for (int r = l + 1; r <= n; r++) Qrout += Qr[r] * T(q, l + 1, r - 1);
        if (Qrout != 0.)
          for (int k = cp; k < l - TURN; k++) {
            int tk = PAIR[S[k]][S[l]];
            if (!tk || T(qb, k, l) == 0.) continue;
            T(out, k, l) += Qrout * T(q, cp, k - 1) * exp_extstem(P, tk, k > cp ? S[k - 1] : -1, S[l + 1]);
          }
        double s = 0.;
        for (int p = 1; p < cp; p++) {
          int tp = PAIR[S[p]][S[l]];
          if (!tp || l - p <= TURN || T(out, p, l) == 0.) continue;
          s += T(out, p, l) * scale[2] * T(q, p + 1, cp - 1) *
               exp_extstem(P, RTYPE[tp], SS(l - 1, l) ? S[l - 1] : -1, SS(p, p + 1) ? S[p + 1] : -1);
        }
        Qr[l] = s;
      } else {
        if (!ql_ready) {
          /* all out(p,r), r>=cp are final now */
          for (int p = 1; p < cp; p++) {
            double s = 0.;
            for (int r = cp; r <= n; r++) {
              int tp = PAIR[S[p]][S[r]];
              if (!tp || r - p <= TURN || T(out, p, r) == 0.) continue;
              s += T(out, p, r) * scale[2] * T(q, cp, r - 1) *
                   exp_extstem(P, RTYPE[tp], SS(r - 1, r) ? S[r - 1] : -1, SS(p, p + 1) ? S[p + 1] : -1);
            }
            Ql[p] = s;
          }
          for (int k = 2; k < cp; k++) {
            double s = 0.;
            for (int p = 1; p < k; p++) s += Ql[p] * T(q, p + 1, k - 1);
            Qlout[k] = s;
          }
          ql_ready = 1;
        }
        for (int k = 2; k < l - TURN; k++) {
          int tk = PAIR[S[k]][S[l]];
          if (!tk || T(qb, k, l) == 0. || Qlout[k] == 0.) continue;
          T(out, k, l) += Qlout[k] * T(q, l + 1, cp - 1) * exp_extstem(P, tk, S[k - 1], l + 1 < cp ? S[l + 1] : -1);
        }
      }
    }
  }
  free(prml); free(prm_l); free(prm_l1); free(Ql); free(Qr); free(Qlout);
#undef SS
}

/* unpaired-window probabilities: up[(i-1)*max_w + d] = P(i..i+d unpaired),
 * 1<=i<=n, 0<=d<max_w; the sum H+I+M+E the reference takes at
 * src/ractip.cpp:373-375 (pf_unstru).  Single strand. */
static void unpaired_windows(const orc_params* P, const seq_t* sq, const tables_t* t, int w, double* up) {
  const int n = t->n, ld = t->ld;
  const int* S = sq->S;
  const double *q = t->q, *qb = t->qb, *qm = t->qm, *qm2 = t->qm2, *out = t->out;
  const double* scale = t->scale;
  size_t sz = (size_t)(n + 2) * (n + 2);
  double* D = (double*)calloc(sz, sizeof(double));  /* hairpin + interior gap weights then dominance sums */
  double* Mc = (double*)calloc(sz, sizeof(double));
  double* R = (double*)calloc(sz, sizeof(double));  /* R(p,j) = sum_{o>j} Mc(p,o) qm2(j+1,o-1) */
  double* L = (double*)calloc(sz, sizeof(double));  /* L(i,o) = sum_{p<i} Mc(p,o) qm2(p+1,i-1) */
  double* X = (double*)calloc(sz, sizeof(double));  /* X(i,o) = sum_{p<i} Mc(p,o) qm(p+1,i-1)  */
  /* G(a,b): total weight (already /Z via out) of loops closed by some (p,o)
   * whose unpaired run is exactly the open interval (a,b): a<i, j<b. */
  for (int p = 1; p <= n; p++)
    for (int o = p + TURN + 1; o <= n; o++) {
      int type = PAIR[S[p]][S[o]];
      double po = T(out, p, o);
      if (!type || po == 0.) continue;
      int u = o - p - 1;
      T(D, p, o) += po * exp_hairpin(P, u, type, S[p + 1], S[o - 1], sq->str + p - 1) * scale[u + 2];
      for (int k = p + 1; k <= MIN2(p + P->maxloop + 1, o - TURN - 2); k++) {
        int u1 = k - p - 1;
        for (int l = MAX2(k + TURN + 1, o - 1 - P->maxloop + u1); l < o; l++) {
          int t2 = PAIR[S[k]][S[l]];
          if (!t2) continue;
          double wgt = po * T(qb, k, l) * scale[u1 + o - l + 1] *
                       exp_intloop(P, u1, o - l - 1, type, RTYPE[t2], S[p + 1], S[o - 1], S[k - 1], S[l + 1]);
          T(D, p, k) += wgt; /* 5' gap (p,k) */
          T(D, l, o) += wgt; /* 3' gap (l,o) */
        }
      }
      T(Mc, p, o) = po * P->expMLclosing * exp_mlstem(P, RTYPE[type], S[o - 1], S[p + 1]) * scale[2];
    }
  /* dominance sums: D(a,b) <- sum_{a'<=a, b'>=b} D(a',b') */
  for (int a = 1; a <= n; a++)
    for (int b = n; b >= 1; b--)
      T(D, a, b) += T(D, a - 1, b) + (b < n ? T(D, a, b + 1) : 0.) - (b < n ? T(D, a - 1, b + 1) : 0.);
  for (int p = 1; p <= n; p++)
    for (int j = p + 1; j <= n; j++) {
      double s = 0.;
      for (int o = j + 2; o <= n; o++) s += T(Mc, p, o) * T(qm2, j + 1, o - 1);
      T(R, p, j) = s;
    }
  for (int o = 1; o <= n; o++)
    for (int i = 2; i < o; i++) {
      double s2 = 0., s1 = 0.;
      for (int p = 1; p <= i - 2; p++) {
        s2 += T(Mc, p, o) * T(qm2, p + 1, i - 1);
        s1 += T(Mc, p, o) * T(qm, p + 1, i - 1);
      }
      T(L, i, o) = s2;
      T(X, i, o) = s1;
    }
  for (int i = 1; i <= n; i++)
    for (int d = 0; d < w; d++) {
      int j = i + d;
      double v = 0.;
      if (j <= n) {
        v = T(q, 1, i - 1) * scale[j - i + 1] * T(q, j + 1, n) / t->Z; /* exterior */
        if (i > 1 && j < n) {
          v += T(D, i - 1, j + 1); /* hairpin + interior */
          double m1 = 0., m2 = 0., m3 = 0.;
          for (int p = 1; p < i; p++) m1 += t->mlb[j - p] * T(R, p, j);
          for (int o = j + 1; o <= n; o++) m2 += t->mlb[o - i] * T(L, i, o);
          for (int o = j + 2; o <= n; o++) m3 += T(qm, j + 1, o - 1) * T(X, i, o);
          v += m1 + m2 + m3 * t->mlb[j - i + 1];
        }
      }
      up[(size_t)(i - 1) * w + d] = v;
    }
  free(D); free(Mc); free(R); free(L); free(X);
}

/* pr: (n+1)*(n+1) row-major doubles, pr[i*(n+1)+j] for 1<=i<j<=n.
 * up: n*max_w doubles or NULL.  Returns 0 on success. */
int orc_fold(const orc_params* P, const char* seq, int n, int cp, double* pr, double* up, int max_w, double* logZ) {
  if (!P || n < 1) return 1;
  seq_t sq; tables_t t;
  seq_init(&sq, seq, n, cp);
  tables_alloc(&t, n, cp, P->pf_scale, P->expMLbase);
  inside(P, &sq, &t);
  if (logZ) *logZ = log(t.Z) + n * log(P->pf_scale);
  if (pr || up) outside(P, &sq, &t);
  if (pr) {
    memset(pr, 0, sizeof(double) * (size_t)(n + 1) * (n + 1));
    const int ld = t.ld;
    for (int i = 1; i <= n; i++)
      for (int j = i + TURN + 1; j <= n; j++) pr[(size_t)i * (n + 1) + j] = T(t.out, i, j) * T(t.qb, i, j);
  }
  if (up && max_w > 0) unpaired_windows(P, &sq, &t, max_w, up);
  tables_free(&t);
  seq_free(&sq);
  return 0;
}

/* ------------------------------------------------------------------------ */
/* pf_duplex: log-space forward/backward over pure duplexes                  */
/* 1:1 restatement of src/pf_duplex.c:34-40 (LogAdd), :67-117, :128-164 (fw), */
/* :166-206 (bk).  pr: (n1+1)*(n2+1) row-major.                              */
/* ------------------------------------------------------------------------ */
static double logadd(double x, double y) {
  if (x <= -INFINITY) return y;
  if (y <= -INFINITY) return x;
  return x > y ? log1p(exp(y - x)) + x : log1p(exp(x - y)) + y;
}

int orc_pf_duplex(const orc_params* P, const char* s1, int n1, const char* s2, int n2, double* pr, double* Esum_out) {
  if (!P || n1 < 1 || n2 < 1) return 1;
  const double kT = P->kT;
  const double NEG = -INFINITY;
  int* A = (int*)calloc(n1 + 2, sizeof(int));
  int* B = (int*)calloc(n2 + 2, sizeof(int));
  for (int i = 1; i <= n1; i++) A[i] = enc(s1[i - 1]);
  for (int j = 1; j <= n2; j++) B[j] = enc(s2[j - 1]);
  size_t ld = n2 + 2;
  double* fw = (double*)malloc(sizeof(double) * (n1 + 2) * ld);
  double* bk = (double*)malloc(sizeof(double) * (n1 + 2) * ld);
  for (size_t x = 0; x < (size_t)(n1 + 2) * ld; x++) fw[x] = bk[x] = NEG;
  double Esum = NEG;
  for (int i = 1; i <= n1; i++)
    for (int j = n2; j > 0; j--) {
      int type = PAIR[A[i]][B[j]];
      if (!type) continue;
      int E = P->m.DuplexInit37 + E_extloop(P, type, i > 1 ? A[i - 1] : -1, j < n2 ? B[j + 1] : -1);
      double f = -E * 10. / kT;
      for (int k = i - 1; k > 0 && k > i - P->maxloop - 2; k--)
        for (int l = j + 1; l <= n2; l++) {
          if (i - k + l - j - 2 > P->maxloop) break;
          int type2 = PAIR[A[k]][B[l]];
          if (!type2) continue;
          E = E_intloop(P, i - k - 1, l - j - 1, type2, RTYPE[type], A[k + 1], B[l - 1], A[i - 1], B[j + 1]);
          f = logadd(f, fw[k * ld + l] - E * 10. / kT);
        }
      fw[i * ld + j] = f;
      E = E_extloop(P, RTYPE[type], j > 1 ? B[j - 1] : -1, i < n1 ? A[i + 1] : -1);
      Esum = logadd(Esum, f - E * 10. / kT);
    }
  for (int i = n1; i > 0; i--)
    for (int j = 1; j <= n2; j++) {
      int type = PAIR[A[i]][B[j]];
      if (!type) continue;
      int E = E_extloop(P, RTYPE[type], j > 1 ? B[j - 1] : -1, i < n1 ? A[i + 1] : -1);
      bk[i * ld + j] = logadd(bk[i * ld + j], -E * 10. / kT);
      for (int k = i - 1; k > 0 && k > i - P->maxloop - 2; k--)
        for (int l = j + 1; l <= n2; l++) {
          if (i - k + l - j - 2 > P->maxloop) break;
          int type2 = PAIR[A[k]][B[l]];
          if (!type2) continue;
          E = E_intloop(P, i - k - 1, l - j - 1, type2, RTYPE[type], A[k + 1], B[l - 1], A[i - 1], B[j + 1]);
          bk[k * ld + l] = logadd(bk[k * ld + l], bk[i * ld + j] - E * 10. / kT);
        }
    }
  if (pr) {
    memset(pr, 0, sizeof(double) * (size_t)(n1 + 1) * (n2 + 1));
    for (int i = 1; i <= n1; i++)
      for (int j = 1; j <= n2; j++) pr[(size_t)i * (n2 + 1) + j] = exp(fw[i * ld + j] + bk[i * ld + j] - Esum);
  }
  if (Esum_out) *Esum_out = Esum;
  free(A); free(B); free(fw); free(bk);
  return 0;
}

/* ------------------------------------------------------------------------ */
/* reference-layout wrappers (what RactIP::rnafold / rnaduplex leave behind) */
/* ------------------------------------------------------------------------ */
/* bp: (L+1)(L+2)/2 floats, bp[offset[i]+j]; up: L*max_w floats.
 * src/ractip.cpp:314-317,365-375 */
int orc_rnafold(const orc_params* P, const char* seq, int L, int max_w, float* bp, float* up) {
  double* pr = (double*)malloc(sizeof(double) * (size_t)(L + 1) * (L + 1));
  double* u = (double*)malloc(sizeof(double) * (size_t)L * (max_w > 0 ? max_w : 1));
  int rc = orc_fold(P, seq, L, 0, pr, max_w > 0 ? u : NULL, max_w, NULL);
  if (!rc) {
    memset(bp, 0, sizeof(float) * (size_t)(L + 1) * (L + 2) / 2);
    for (int i = 1; i < L; i++) {
      size_t off = (size_t)i * ((L + 1) + (L + 1) - i - 1) / 2;
      for (int j = i + 1; j <= L; j++) bp[off + j] = (float)pr[(size_t)i * (L + 1) + j];
    }
    for (size_t x = 0; x < (size_t)L * max_w; x++) up[x] = (float)u[x];
  }
  free(pr); free(u);
  return rc;
}

/* hp: (L1+1)*(L2+1) floats row-major.  src/ractip.cpp:390-458 */
int orc_rnaduplex(const orc_params* P, const char* s1, int L1, const char* s2, int L2, float th_hy, int use_pf_duplex, float* hp) {
  memset(hp, 0, sizeof(float) * (size_t)(L1 + 1) * (L2 + 1));
  if (use_pf_duplex) {
    double* pr = (double*)malloc(sizeof(double) * (size_t)(L1 + 1) * (L2 + 1));
    int rc = orc_pf_duplex(P, s1, L1, s2, L2, pr, NULL);
    if (!rc)
      for (int i = 1; i <= L1; i++)
        for (int j = 1; j <= L2; j++) hp[(size_t)i * (L2 + 1) + j] = (float)pr[(size_t)i * (L2 + 1) + j];
    free(pr);
    return rc;
  }
  int n = L1 + L2, cp = L1 + 1;
  char* s = (char*)malloc(n + 1);
  memcpy(s, s1, L1); memcpy(s + L1, s2, L2); s[n] = 0;
  double* pr = (double*)malloc(sizeof(double) * (size_t)(n + 1) * (n + 1));
  int rc = orc_fold(P, s, n, cp, pr, NULL, 0, NULL);
  if (!rc)
    for (int i = 1; i < cp; i++)
      for (int j = cp; j <= n; j++) {
        /* plist entry p is a float; kept iff p >= cutoff and then p > th_hy (:447,452) */
        float p = (float)pr[(size_t)i * (n + 1) + j];
        if (pr[(size_t)i * (n + 1) + j] >= (double)th_hy && p > th_hy) hp[(size_t)i * (L2 + 1) + (j - cp + 1)] = p;
      }
  free(pr); free(s);
  return rc;
}

/* ------------------------------------------------------------------------ */
/* exhaustive enumeration: independent evaluation of the same loop model.    */
/* Unscaled weights.  n <= ~18.                                              */
/* ------------------------------------------------------------------------ */
typedef struct {
  const orc_params* P;
  const seq_t* sq;
  int n, cp, w;
  int* pt;
  double Z;
  double *pr, *up;
} enum_t;

static double loop_weight_closed(const enum_t* E, int i, int j) {
  const orc_params* P = E->P;
  const int* S = E->sq->S;
  const int* pt = E->pt;
  const int cp = E->cp;
  int type = PAIR[S[i]][S[j]];
  /* children */
  int nchild = 0, unp = 0, k1 = 0, l1 = 0, nick_in_child = 0;
  for (int x = i + 1; x < j;) {
    if (pt[x] > x) {
      if (!nchild) { k1 = x; l1 = pt[x]; }
      nchild++;
      if (cp > 0 && x < cp && pt[x] >= cp) nick_in_child = 1;
      x = pt[x] + 1;
    } else { unp++; x++; }
  }
  int nicked = cp > 0 && i < cp && j >= cp && !nick_in_child;
#define SSE(a, b) same_strand(cp, a, b)
  if (nicked) {
    double wgt = exp_extstem(P, RTYPE[type], SSE(j - 1, j) ? S[j - 1] : -1, SSE(i, i + 1) ? S[i + 1] : -1);
    for (int x = i + 1; x < j;) {
      if (pt[x] > x) {
        int k = x, l = pt[x];
        wgt *= exp_extstem(P, PAIR[S[k]][S[l]], SSE(k - 1, k) ? S[k - 1] : -1, SSE(l, l + 1) ? S[l + 1] : -1);
        x = l + 1;
      } else x++;
    }
    return wgt;
  }
  if (nchild == 0) return exp_hairpin(P, j - i - 1, type, S[i + 1], S[j - 1], E->sq->str + i - 1);
  if (nchild == 1) {
    int u1 = k1 - i - 1, u2 = j - l1 - 1;
    if (u1 + u2 > P->maxloop) return 0.;
    return exp_intloop(P, u1, u2, type, RTYPE[PAIR[S[k1]][S[l1]]], S[i + 1], S[j - 1], S[k1 - 1], S[l1 + 1]);
  }
  double wgt = P->expMLclosing * exp_mlstem(P, RTYPE[type], S[j - 1], S[i + 1]);
  for (int x = i + 1; x < j;) {
    if (pt[x] > x) {
      int k = x, l = pt[x];
      wgt *= exp_mlstem(P, PAIR[S[k]][S[l]], S[k - 1], S[l + 1]);
      x = l + 1;
    } else { wgt *= P->expMLbase; x++; }
  }
  (void)unp;
  return wgt;
}

static void enum_eval(enum_t* E) {
  const orc_params* P = E->P;
  const int* S = E->sq->S;
  const int n = E->n, cp = E->cp;
  const int* pt = E->pt;
  double wgt = 1.;
  for (int x = 1; x <= n;) { /* exterior loop */
    if (pt[x] > x) {
      int k = x, l = pt[x];
      wgt *= exp_extstem(P, PAIR[S[k]][S[l]], (k > 1 && SSE(k - 1, k)) ? S[k - 1] : -1, (l < n && SSE(l, l + 1)) ? S[l + 1] : -1);
      x = l + 1;
    } else x++;
  }
  for (int i = 1; i <= n && wgt != 0.; i++)
    if (pt[i] > i) wgt *= loop_weight_closed(E, i, pt[i]);
  if (wgt == 0.) return;
  E->Z += wgt;
  for (int i = 1; i <= n; i++)
    if (pt[i] > i) E->pr[(size_t)i * (n + 1) + pt[i]] += wgt;
  if (E->up)
    for (int i = 1; i <= n; i++)
      for (int d = 0; d < E->w && i + d <= n; d++) {
        if (pt[i + d]) break;
        E->up[(size_t)(i - 1) * E->w + d] += wgt;
      }
#undef SSE
}

static void enum_rec(enum_t* E, int pos, int* stack, int sp) {
  const int n = E->n;
  if (pos > n) { enum_eval(E); return; }
  if (E->pt[pos]) { /* closes an open pair */
    enum_rec(E, pos + 1, stack, sp - 1);
    return;
  }
  enum_rec(E, pos + 1, stack, sp); /* unpaired */
  int limit = sp > 0 ? stack[sp - 1] - 1 : n;
  for (int j = pos + TURN + 1; j <= limit; j++) {
    if (E->pt[j] || !PAIR[E->sq->S[pos]][E->sq->S[j]]) continue;
    E->pt[pos] = j; E->pt[j] = pos;
    int saved = stack[sp]; /* deeper frames reuse this slot after popping: restore on return */
    stack[sp] = j;
    enum_rec(E, pos + 1, stack, sp + 1);
    stack[sp] = saved;
    E->pt[pos] = 0; E->pt[j] = 0;
  }
}

int orc_enumerate(const orc_params* P, const char* seq, int n, int cp, double* pr, double* up, int max_w, double* logZ) {
  if (!P || n < 1 || n > 24) return 1;
  seq_t sq;
  seq_init(&sq, seq, n, cp);
  enum_t E;
  E.P = P; E.sq = &sq; E.n = n; E.cp = cp; E.w = max_w;
  E.pt = (int*)calloc(n + 2, sizeof(int));
  E.Z = 0.;
  E.pr = pr; E.up = (max_w > 0) ? up : NULL;
  memset(pr, 0, sizeof(double) * (size_t)(n + 1) * (n + 1));
  if (E.up) memset(up, 0, sizeof(double) * (size_t)n * max_w);
  int* stack = (int*)calloc(n + 2, sizeof(int));
  enum_rec(&E, 1, stack, 0);
  for (size_t x = 0; x < (size_t)(n + 1) * (n + 1); x++) pr[x] /= E.Z;
  if (E.up) for (size_t x = 0; x < (size_t)n * max_w; x++) up[x] /= E.Z;
  if (logZ) *logZ = log(E.Z);
  free(stack); free(E.pt);
  seq_free(&sq);
  return 0;
}

/* brute-force duplex ensemble for pf_duplex: every chain of pairs
 * (i1<i2<...; j1>j2>...) with consecutive gaps forming a <=P->maxloop loop */
int orc_enum_duplex(const orc_params* P, const char* s1, int n1, const char* s2, int n2, double* pr, double* logZ) {
  if (n1 > 10 || n2 > 10) return 1;
  const double kT = P->kT;
  int A[16] = {0}, B[16] = {0};
  for (int i = 1; i <= n1; i++) A[i] = enc(s1[i - 1]);
  for (int j = 1; j <= n2; j++) B[j] = enc(s2[j - 1]);
  memset(pr, 0, sizeof(double) * (size_t)(n1 + 1) * (n2 + 1));
  double Z = 0.;
  /* choose subsets of positions: mask1 over s1, mask2 over s2 with equal popcount */
  for (unsigned m1 = 1; m1 < (1u << n1); m1++)
    for (unsigned m2 = 1; m2 < (1u << n2); m2++) {
      if (__builtin_popcount(m1) != __builtin_popcount(m2)) continue;
      int is[16], js[16], c = 0, d = 0;
      for (int i = 1; i <= n1; i++) if (m1 >> (i - 1) & 1) is[c++] = i;
      for (int j = n2; j >= 1; j--) if (m2 >> (j - 1) & 1) js[d++] = j;
      int ok = 1, E = P->m.DuplexInit37;
      for (int x = 0; x < c && ok; x++) {
        int type = PAIR[A[is[x]]][B[js[x]]];
        if (!type) { ok = 0; break; }
        if (x == 0) E += E_extloop(P, type, is[0] > 1 ? A[is[0] - 1] : -1, js[0] < n2 ? B[js[0] + 1] : -1);
        else {
          int k = is[x - 1], l = js[x - 1], i = is[x], j = js[x];
          if (i - k + l - j - 2 > P->maxloop) { ok = 0; break; }
          E += E_intloop(P, i - k - 1, l - j - 1, PAIR[A[k]][B[l]], RTYPE[type], A[k + 1], B[l - 1], A[i - 1], B[j + 1]);
        }
        if (x == c - 1) E += E_extloop(P, RTYPE[type], js[x] > 1 ? B[js[x] - 1] : -1, is[x] < n1 ? A[is[x] + 1] : -1);
      }
      if (!ok) continue;
      double wgt = exp(-E * 10. / kT);
      Z += wgt;
      for (int x = 0; x < c; x++) pr[(size_t)is[x] * (n2 + 1) + js[x]] += wgt;
    }
  for (size_t x = 0; x < (size_t)(n1 + 1) * (n2 + 1); x++) pr[x] /= Z;
  if (logZ) *logZ = log(Z);
  return 0;
}
