// ipmodel.cpp -- RactIP's integer programme, restated from the reference (host code, no GPU).
//
// Builds the model of RactIP::solve (reference src/ractip.cpp:552-1225) and RactIP::solve_ss
// (:1366-1465): the same binary variables in the same creation order (= column order), the same
// constraint rows in the same order, the same float arithmetic for the objective weights, so that a
// solver sees the very problem the reference hands to GLPK / HiGHS (src/ip.cpp:49-135,489-622).
// Decoding restates :1227-1250,1286-1295.  Not modelled: structure constraints (-c,
// --force-constraint; :655-713,1170-1222).
//
// The model is kept as plain arrays (objective, row bounds, coefficient triplets in
// IP::add_constraint call order); include/ractip_ip.h documents the C ABI.
#include <cstring>
#include <utility>
#include <vector>

#include "ractip_ip.h"

namespace {

typedef std::vector<int> VI;
typedef std::vector<VI> VVI;

struct Rec { int i, j; float p; };

}  // namespace

struct rp_ip_model {
  int n1 = 0, n2 = 0;
  bool ss_model = false;
  rp_ip_opts o{};
  std::vector<double> obj;
  std::vector<int> kind;
  std::vector<double> lo, hi;
  std::vector<int> ia, ja;
  std::vector<double> ar;
  // what decoding needs: (i, j, column) of the pair variables, (first, last, column) of the regions
  std::vector<Rec> xs, ys, zs, vs, ws;   // Rec::p unused; i, j and the column in a parallel vector
  VI xcol, ycol, zcol, vcol, wcol;

  // IP::make_variable(coef): a binary column (src/ip.cpp:69-76)
  int make_variable(double coef) { obj.push_back(coef); return (int)obj.size() - 1; }
  // IP::make_constraint(bnd, l, u) (src/ip.cpp:87-99)
  int make_constraint(int k, double l, double u) { kind.push_back(k); lo.push_back(l); hi.push_back(u); return (int)kind.size() - 1; }
  // IP::add_constraint(row, col, val) (src/ip.cpp:101-108)
  void add_constraint(int row, int col, double val) { ia.push_back(row); ja.push_back(col); ar.push_back(val); }
};

namespace {

// bp[offset[i]+j], offset[i] = i*((L+1)+(L+1)-i-1)/2 (src/ractip.cpp:316-317)
inline size_t bp_index(int L, int i, int j) { return (size_t)i * ((L + 1) + (L + 1) - i - 1) / 2 + j; }

// thresholded lists in the creation order of src/ractip.cpp:557-567 (x, y), :598-609 (z), :619-628 (v, w)
void lists_from_dense(const rp_ip_opts& o, int n1, int n2, const float* bp1, const float* bp2, const float* up1,
                      const float* up2, const float* hp, std::vector<Rec>& X, std::vector<Rec>& Y, std::vector<Rec>& Z,
                      std::vector<Rec>& V, std::vector<Rec>& W) {
  auto pairs = [&](int L, const float* bp, std::vector<Rec>& out) {
    if (!bp) return;
    for (int j = 1; j < L; ++j)
      for (int i = j - 1; i >= 0; --i) {
        const float p = bp[bp_index(L, i + 1, j + 1)];
        if (p > o.th_ss) out.push_back({i, j, p});
      }
  };
  pairs(n1, bp1, X);
  pairs(n2, bp2, Y);
  if (hp)
    for (int i = 0; i < n1; ++i)
      for (int j = 0; j < n2; ++j) {
        const float p = hp[(size_t)(i + 1) * (n2 + 1) + (j + 1)];
        if (p > o.th_hy) Z.push_back({i, j, p});
      }
  const int width = o.max_w > 1 ? o.max_w : 1;   // rnafold is called with std::max(1, max_w_) (:546-547)
  auto regions = [&](int L, const float* up, std::vector<Rec>& out) {
    if (!up) return;
    for (int i = 0; i < L; ++i)
      for (int j = o.min_w - 1; j < width; ++j) {
        if (j < 0) continue;
        const float p = up[(size_t)i * width + j];
        if (p > o.th_ac) out.push_back({i, j, p});
      }
  };
  regions(n1, up1, V);
  regions(n2, up2, W);
}

// the stacking rows of one sequence (src/ractip.cpp:1059-1112 for s1, :1114-1146 for s2, :1402-1436 solve_ss)
void stacking_rows(rp_ip_model& M, const VVI& x, int n) {
  // upstream
  for (int i = 0; i < n; ++i) {
    const int row = M.make_constraint(RP_IP_LO, 0, 0);
    for (int j = 0; j < i; ++j)
      if (x[j][i] >= 0) M.add_constraint(row, x[j][i], -1);
    if (i > 0)
      for (int j = 0; j < i - 1; ++j)
        if (x[j][i - 1] >= 0) M.add_constraint(row, x[j][i - 1], 1);
    if (i + 1 < n)
      for (int j = 0; j < i + 1; ++j)
        if (x[j][i + 1] >= 0) M.add_constraint(row, x[j][i + 1], 1);
  }
  // downstream
  for (int i = 0; i < n; ++i) {
    const int row = M.make_constraint(RP_IP_LO, 0, 0);
    for (int j = i + 1; j < n; ++j)
      if (x[i][j] >= 0) M.add_constraint(row, x[i][j], -1);
    if (i > 0)
      for (int j = i; j < n; ++j)
        if (x[i - 1][j] >= 0) M.add_constraint(row, x[i - 1][j], 1);
    if (i + 1 < n)
      for (int j = i + 2; j < n; ++j)
        if (x[i + 1][j] >= 0) M.add_constraint(row, x[i + 1][j], 1);
  }
}

int build_joint(const rp_ip_opts& o, int n1, int n2, const std::vector<Rec>& X, const std::vector<Rec>& Y,
                const std::vector<Rec>& Z, const std::vector<Rec>& V, const std::vector<Rec>& W, rp_ip_model** out) {
  rp_ip_model* Mp = new rp_ip_model;
  rp_ip_model& M = *Mp;
  M.n1 = n1; M.n2 = n2; M.o = o;
  const bool enable_accessibility = o.min_w > 1 && o.max_w >= o.min_w;   // :526
  const bool enable_structure_s1 = !o.acc_max, enable_structure_s2 = !o.acc_max;   // :527-528

  // ---- variables, in creation order (:552-653)
  VVI x(n1, VI(n1, -1)), xx(n1);
  VI x_un(n1, -1);
  if (enable_structure_s1) {
    for (const Rec& r : X) {
      x[r.i][r.j] = x[r.j][r.i] = M.make_variable(r.p - o.th_ss);   // float arithmetic, as `p-th_ss_` (:562)
      xx[r.i].push_back(r.j);
      M.xs.push_back(r); M.xcol.push_back(x[r.i][r.j]);
    }
    for (int i = 0; i < n1; ++i) x_un[i] = M.make_variable(0.0);
  }
  VVI y(n2, VI(n2, -1)), yy(n2);
  VI y_un(n2, -1);
  if (enable_structure_s2) {
    for (const Rec& r : Y) {
      y[r.i][r.j] = y[r.j][r.i] = M.make_variable(r.p - o.th_ss);
      yy[r.i].push_back(r.j);
      M.ys.push_back(r); M.ycol.push_back(y[r.i][r.j]);
    }
    for (int i = 0; i < n2; ++i) y_un[i] = M.make_variable(0.0);
  }
  VVI z(n1, VI(n2, -1)), zz(n1);
  VI z_un1(n1, -1), z_un2(n2, -1);
  for (const Rec& r : Z) {
    z[r.i][r.j] = M.make_variable(o.alpha * (r.p - o.th_hy));   // float: alpha_*(p-th_hy_) (:605)
    zz[r.i].push_back(r.j);
    M.zs.push_back(r); M.zcol.push_back(z[r.i][r.j]);
  }
  for (int i = 0; i < n1; ++i) z_un1[i] = M.make_variable(0.0);
  for (int i = 0; i < n2; ++i) z_un2[i] = M.make_variable(0.0);

  VI v, w;
  std::vector<std::pair<int, int> > vv, ww;
  VI v_st(n1, -1), v_en(n1, -1), w_st(n2, -1), w_en(n2, -1);
  if (enable_accessibility)
    for (const Rec& r : V) {
      v.push_back(M.make_variable(o.beta * (r.p - o.th_ac)));   // float: beta_*(up-th_ac_) (:625)
      vv.push_back(std::make_pair(r.i, r.i + r.j));
      M.vs.push_back({r.i, r.i + r.j, r.p}); M.vcol.push_back(v.back());
    }
  for (int i = 0; i < n1; ++i) { v_st[i] = M.make_variable(0.0); v_en[i] = M.make_variable(0.0); }
  if (enable_accessibility)
    for (const Rec& r : W) {
      w.push_back(M.make_variable(o.beta * (r.p - o.th_ac)));
      ww.push_back(std::make_pair(r.i, r.i + r.j));
      M.ws.push_back({r.i, r.i + r.j, r.p}); M.wcol.push_back(w.back());
    }
  for (int i = 0; i < n2; ++i) { w_st[i] = M.make_variable(0.0); w_en[i] = M.make_variable(0.0); }
  for (size_t k = 0; k < vv.size(); k++)
    if (vv[k].second >= n1) { delete Mp; return RP_ERR_ARG; }   // a window past the 3' end has probability 0
  for (size_t k = 0; k < ww.size(); k++)
    if (ww[k].second >= n2) { delete Mp; return RP_ERR_ARG; }

  // ---- constraints for helper variables (:717-762)
  if (enable_structure_s1)
    for (int i = 0; i < n1; ++i) {   // sum_j x[i][j] + x_un[i] = 1
      const int row = M.make_constraint(RP_IP_FX, 1, 1);
      M.add_constraint(row, x_un[i], 1);
      for (int j = 0; j < n1; ++j)
        if (x[i][j] >= 0) M.add_constraint(row, x[i][j], 1);
    }
  for (int i = 0; i < n1; ++i) {     // sum_j z[i][j] + z_un1[i] = 1
    const int row = M.make_constraint(RP_IP_FX, 1, 1);
    M.add_constraint(row, z_un1[i], 1);
    for (int j = 0; j < n2; ++j)
      if (z[i][j] >= 0) M.add_constraint(row, z[i][j], 1);
  }
  if (enable_structure_s2)
    for (int i = 0; i < n2; ++i) {
      const int row = M.make_constraint(RP_IP_FX, 1, 1);
      M.add_constraint(row, y_un[i], 1);
      for (int j = 0; j < n2; ++j)
        if (y[i][j] >= 0) M.add_constraint(row, y[i][j], 1);
    }
  for (int i = 0; i < n2; ++i) {     // sum_j z[j][i] + z_un2[i] = 1
    const int row = M.make_constraint(RP_IP_FX, 1, 1);
    M.add_constraint(row, z_un2[i], 1);
    for (int j = 0; j < n1; ++j)
      if (z[j][i] >= 0) M.add_constraint(row, z[j][i], 1);
  }

  if (enable_accessibility) {        // region start / end counters (:764-799)
    VI row_v_st(n1, -1), row_v_en(n1, -1);
    for (int i = 0; i < n1; ++i) {
      row_v_st[i] = M.make_constraint(RP_IP_FX, 0, 0);
      M.add_constraint(row_v_st[i], v_st[i], -1);
      row_v_en[i] = M.make_constraint(RP_IP_FX, 0, 0);
      M.add_constraint(row_v_en[i], v_en[i], -1);
    }
    for (size_t i = 0; i < v.size(); ++i) {
      M.add_constraint(row_v_st[vv[i].first], v[i], 1);
      M.add_constraint(row_v_en[vv[i].second], v[i], 1);
    }
    VI row_w_st(n2, -1), row_w_en(n2, -1);
    for (int i = 0; i < n2; ++i) {
      row_w_st[i] = M.make_constraint(RP_IP_FX, 0, 0);
      M.add_constraint(row_w_st[i], w_st[i], -1);
      row_w_en[i] = M.make_constraint(RP_IP_FX, 0, 0);
      M.add_constraint(row_w_en[i], w_en[i], -1);
    }
    for (size_t i = 0; i < w.size(); ++i) {
      M.add_constraint(row_w_st[ww[i].first], w[i], 1);
      M.add_constraint(row_w_en[ww[i].second], w[i], 1);
    }
  }

  if (!enable_accessibility) {       // each base pairs at most once (:802-829)
    if (enable_structure_s1)
      for (int i = 0; i < n1; ++i) {
        const int row = M.make_constraint(RP_IP_LO, 1, 0);
        M.add_constraint(row, x_un[i], 1);
        M.add_constraint(row, z_un1[i], 1);
      }
    if (enable_structure_s2)
      for (int i = 0; i < n2; ++i) {
        const int row = M.make_constraint(RP_IP_LO, 1, 0);
        M.add_constraint(row, y_un[i], 1);
        M.add_constraint(row, z_un2[i], 1);
      }
  } else {                           // accessibility rows (:830-982)
    auto cover = [&](const VI& row, const VI& var, const std::vector<std::pair<int, int> >& span) {
      for (size_t j = 0; j < var.size(); ++j)
        for (int i = span[j].first; i <= span[j].second; ++i) M.add_constraint(row[i], var[j], 1);
    };
    if (enable_structure_s1) {       // internal pairs of s1 are not accessible
      VI row(n1);
      for (int i = 0; i < n1; ++i) {
        row[i] = M.make_constraint(RP_IP_UP, 0, 0);
        M.add_constraint(row[i], x_un[i], -1);
      }
      cover(row, v, vv);
    }
    {                                // external pairs of s1 lie in accessible regions
      VI row(n1, 0);
      for (int i = 0; i < n1; ++i) {
        row[i] = M.make_constraint(RP_IP_LO, 1, 0);
        M.add_constraint(row[i], z_un1[i], 1);
      }
      cover(row, v, vv);
    }
    if (enable_structure_s2) {
      VI row(n2);
      for (int i = 0; i < n2; ++i) {
        row[i] = M.make_constraint(RP_IP_UP, 0, 0);
        M.add_constraint(row[i], y_un[i], -1);
      }
      cover(row, w, ww);
    }
    {
      VI row(n2);
      for (int i = 0; i < n2; ++i) {
        row[i] = M.make_constraint(RP_IP_LO, 1, 0);
        M.add_constraint(row[i], z_un2[i], 1);
      }
      cover(row, w, ww);
    }
    {                                // each position of s1 lies in at most one region
      VI row(n1, -1);
      for (int i = 0; i < n1; ++i) row[i] = M.make_constraint(RP_IP_UP, 0, 1);
      cover(row, v, vv);
    }
    for (int i = 1; i < n1; ++i) {   // regions do not adjoin
      const int row = M.make_constraint(RP_IP_UP, 0, 1);
      M.add_constraint(row, v_en[i - 1], 1);
      M.add_constraint(row, v_st[i], 1);
    }
    {
      VI row(n2, -1);
      for (int i = 0; i < n2; ++i) row[i] = M.make_constraint(RP_IP_UP, 0, 1);
      cover(row, w, ww);
    }
    for (int i = 1; i < n2; ++i) {
      const int row = M.make_constraint(RP_IP_UP, 0, 1);
      M.add_constraint(row, w_en[i - 1], 1);
      M.add_constraint(row, w_st[i], 1);
    }
    if (o.beta > 0.0) {              // every region holds an external pair (:927-950)
      for (size_t j = 0; j < v.size(); ++j) {
        const int row = M.make_constraint(RP_IP_UP, 0, vv[j].second - vv[j].first + 1);
        M.add_constraint(row, v[j], 1);
        for (int i = vv[j].first; i <= vv[j].second; ++i) M.add_constraint(row, z_un1[i], 1);
      }
      for (size_t j = 0; j < w.size(); ++j) {
        const int row = M.make_constraint(RP_IP_UP, 0, ww[j].second - ww[j].first + 1);
        M.add_constraint(row, w[j], 1);
        for (int i = ww[j].first; i <= ww[j].second; ++i) M.add_constraint(row, z_un2[i], 1);
      }
    }
    if (o.acc_num > 0) {             // at most acc_num regions (:969-982)
      int row = M.make_constraint(RP_IP_UP, 0, o.acc_num);
      for (size_t i = 0; i < v.size(); ++i) M.add_constraint(row, v[i], 1);
      row = M.make_constraint(RP_IP_UP, 0, o.acc_num);
      for (size_t i = 0; i < w.size(); ++i) M.add_constraint(row, w[i], 1);
    }
  }
  if (enable_accessibility && o.acc_num > 0) {   // the reference emits these two rows a second time (:985-994)
    int row = M.make_constraint(RP_IP_UP, 0, o.acc_num);
    for (size_t i = 0; i < v.size(); ++i) M.add_constraint(row, v[i], 1);
    row = M.make_constraint(RP_IP_UP, 0, o.acc_num);
    for (size_t i = 0; i < w.size(); ++i) M.add_constraint(row, w[i], 1);
  }

  // ---- no crossing external pairs (:996-1012)
  for (size_t i = 0; i < zz.size(); ++i)
    for (size_t k = i + 1; k < zz.size(); ++k)
      for (size_t p = 0; p < zz[i].size(); ++p) {
        const int j = zz[i][p];
        for (size_t q = 0; q < zz[k].size(); ++q) {
          const int l = zz[k][q];
          if (j < l) {
            const int row = M.make_constraint(RP_IP_UP, 0, 1);
            M.add_constraint(row, z[i][j], 1);
            M.add_constraint(row, z[k][l], 1);
          }
        }
      }

  // ---- no internal pseudoknots (:1014-1057)
  if (o.in_pk) {
    auto no_pk = [&](const VVI& xp, const VVI& xxp) {
      for (size_t i = 0; i < xxp.size(); ++i)
        for (size_t p = 0; p < xxp[i].size(); ++p) {
          const int j = xxp[i][p];
          for (int k = (int)i + 1; k < j; ++k)
            for (size_t q = 0; q < xxp[k].size(); ++q) {
              const int l = xxp[k][q];
              if (j < l) {
                const int row = M.make_constraint(RP_IP_UP, 0, 1);
                M.add_constraint(row, xp[i][j], 1);
                M.add_constraint(row, xp[k][l], 1);
              }
            }
        }
    };
    if (enable_structure_s1) no_pk(x, xx);
    if (enable_structure_s2) no_pk(y, yy);
  }

  // ---- stacking (no isolated pairs) (:1059-1168)
  if (o.stacking) {
    if (enable_structure_s1) stacking_rows(M, x, n1);
    if (enable_structure_s2) stacking_rows(M, y, n2);
    for (int i = 0; i < n2; ++i) {   // external pairs, seen from s2
      const int row = M.make_constraint(RP_IP_LO, 0, 0);
      for (int j = 0; j < n1; ++j)
        if (z[j][i] >= 0) M.add_constraint(row, z[j][i], -1);
      if (i > 0)
        for (int j = 0; j < n1; ++j)
          if (z[j][i - 1] >= 0) M.add_constraint(row, z[j][i - 1], 1);
      if (i + 1 < n2)
        for (int j = 0; j < n1; ++j)
          if (z[j][i + 1] >= 0) M.add_constraint(row, z[j][i + 1], 1);
    }
    for (int i = 0; i < n1; ++i) {   // ... and from s1
      const int row = M.make_constraint(RP_IP_LO, 0, 0);
      for (int j = 0; j < n2; ++j)
        if (z[i][j] >= 0) M.add_constraint(row, z[i][j], -1);
      if (i > 0)
        for (int j = 0; j < n2; ++j)
          if (z[i - 1][j] >= 0) M.add_constraint(row, z[i - 1][j], 1);
      if (i + 1 < n1)
        for (int j = 0; j < n2; ++j)
          if (z[i + 1][j] >= 0) M.add_constraint(row, z[i + 1][j], 1);
    }
  }
  *out = Mp;
  return RP_OK;
}

bool check_recs(const rp_rec* r, int n, int lim_i, int lim_j) {
  if (n < 0 || (n > 0 && !r)) return false;
  for (int k = 0; k < n; k++)
    if (r[k].i < 0 || r[k].i >= lim_i || r[k].j < 0 || r[k].j >= lim_j) return false;
  return true;
}

}  // namespace

extern "C" {

void rp_ip_opts_default(rp_ip_opts* o) {
  if (!o) return;
  // src/cmdline.c:151-186 as mapped at src/ractip.cpp:1474-1498
  o->alpha = 0.7f; o->beta = 0.0f; o->th_ss = 0.5f; o->th_hy = 0.1f; o->th_ac = 0.003f;
  o->max_w = 15; o->min_w = 5; o->acc_max = 0; o->acc_max_ss = 0; o->acc_num = 1; o->in_pk = 1; o->stacking = 1;
}

int rp_ip_build(const rp_ip_opts* o, int n1, int n2, const float* bp1, const float* bp2, const float* up1,
                const float* up2, const float* hp, rp_ip_model** out) {
  if (!o || !out || n1 < 1 || n2 < 1 || !hp) return RP_ERR_ARG;
  if (!o->acc_max && (!bp1 || !bp2)) return RP_ERR_ARG;
  const bool acc = o->min_w > 1 && o->max_w >= o->min_w;
  if (acc && (!up1 || !up2)) return RP_ERR_ARG;
  std::vector<Rec> X, Y, Z, V, W;
  lists_from_dense(*o, n1, n2, bp1, bp2, acc ? up1 : nullptr, acc ? up2 : nullptr, hp, X, Y, Z, V, W);
  return build_joint(*o, n1, n2, X, Y, Z, V, W, out);
}

int rp_ip_build_sparse(const rp_ip_opts* o, int n1, int n2, const rp_rec* x, int nx, const rp_rec* y, int ny,
                       const rp_rec* z, int nz, const rp_rec* v, int nv, const rp_rec* w, int nw, rp_ip_model** out) {
  if (!o || !out || n1 < 1 || n2 < 1) return RP_ERR_ARG;
  const int width = o->max_w > 1 ? o->max_w : 1;
  if (!(o->min_w > 1 && o->max_w >= o->min_w)) nv = nw = 0;   // accessibility off: the reference creates no v/w (:526)
  if (o->acc_max) nx = ny = 0;                                // --acc-max: no internal-pair variables (:527-528)
  if (!check_recs(x, nx, n1, n1) || !check_recs(y, ny, n2, n2) || !check_recs(z, nz, n1, n2) ||
      !check_recs(v, nv, n1, width) || !check_recs(w, nw, n2, width))
    return RP_ERR_ARG;
  auto conv = [](const rp_rec* r, int n) {
    std::vector<Rec> o2((size_t)n);
    for (int k = 0; k < n; k++) o2[k] = {r[k].i, r[k].j, r[k].p};
    return o2;
  };
  return build_joint(*o, n1, n2, conv(x, nx), conv(y, ny), conv(z, nz), conv(v, nv), conv(w, nw), out);
}

int rp_ip_build_ss(const rp_ip_opts* o, int n, const float* bp, const unsigned char* usable, rp_ip_model** out) {
  if (!o || !out || n < 1 || !bp) return RP_ERR_ARG;
  rp_ip_model* Mp = new rp_ip_model;
  rp_ip_model& M = *Mp;
  M.n1 = n; M.n2 = 0; M.ss_model = true; M.o = *o;
  VVI x(n, VI(n, -1));
  for (int j = 1; j < n; ++j) {        // :1376-1389
    if (usable && !usable[j]) continue;
    for (int i = j - 1; i >= 0; --i) {
      if (usable && !usable[i]) continue;
      const float p = bp[bp_index(n, i + 1, j + 1)];
      if (p > o->th_ss) {
        x[i][j] = x[j][i] = M.make_variable(p - o->th_ss);
        M.xs.push_back({i, j, p}); M.xcol.push_back(x[i][j]);
      }
    }
  }
  for (int i = 0; i < n; ++i) {        // each base pairs at most once (:1393-1400)
    const int row = M.make_constraint(RP_IP_UP, 0, 1);
    for (int j = 0; j < n; ++j)
      if (x[i][j] >= 0) M.add_constraint(row, x[i][j], 1);
  }
  if (o->stacking) stacking_rows(M, x, n);   // :1402-1436
  *out = Mp;
  return RP_OK;
}

int rp_ip_dims(const rp_ip_model* m, int* n_cols, int* n_rows, int* n_nonzeros) {
  if (!m) return RP_ERR_ARG;
  if (n_cols) *n_cols = (int)m->obj.size();
  if (n_rows) *n_rows = (int)m->kind.size();
  if (n_nonzeros) *n_nonzeros = (int)m->ia.size();
  return RP_OK;
}

int rp_ip_export(const rp_ip_model* m, double* obj, int* row_kind, double* row_lo, double* row_hi, int* ia, int* ja,
                 double* ar) {
  if (!m) return RP_ERR_ARG;
  if (obj) std::memcpy(obj, m->obj.data(), m->obj.size() * sizeof(double));
  if (row_kind) std::memcpy(row_kind, m->kind.data(), m->kind.size() * sizeof(int));
  if (row_lo) std::memcpy(row_lo, m->lo.data(), m->lo.size() * sizeof(double));
  if (row_hi) std::memcpy(row_hi, m->hi.data(), m->hi.size() * sizeof(double));
  if (ia) std::memcpy(ia, m->ia.data(), m->ia.size() * sizeof(int));
  if (ja) std::memcpy(ja, m->ja.data(), m->ja.size() * sizeof(int));
  if (ar) std::memcpy(ar, m->ar.data(), m->ar.size() * sizeof(double));
  return RP_OK;
}

int rp_ip_decode(const rp_ip_model* m, const double* val, char* r1, char* r2, unsigned char* used1, unsigned char* used2) {
  if (!m || !val || !r1) return RP_ERR_ARG;
  std::memset(r1, '.', (size_t)m->n1);
  r1[m->n1] = 0;
  if (m->ss_model) {                   // :1440-1450
    for (size_t k = 0; k < m->xs.size(); k++)
      if (val[m->xcol[k]] > 0.5) { r1[m->xs[k].i] = '('; r1[m->xs[k].j] = ')'; }
    return RP_OK;
  }
  if (!r2) return RP_ERR_ARG;
  std::memset(r2, '.', (size_t)m->n2);
  r2[m->n2] = 0;
  for (size_t k = 0; k < m->zs.size(); k++)   // :1232-1237
    if (val[m->zcol[k]] > 0.5) { r1[m->zs[k].i] = '['; r2[m->zs[k].j] = ']'; }
  if (!m->o.acc_max && m->o.in_pk) {          // :1241-1250, :1286-1295
    for (size_t k = 0; k < m->xs.size(); k++)
      if (val[m->xcol[k]] > 0.5) { r1[m->xs[k].i] = '('; r1[m->xs[k].j] = ')'; }
    for (size_t k = 0; k < m->ys.size(); k++)
      if (val[m->ycol[k]] > 0.5) { r2[m->ys[k].i] = '('; r2[m->ys[k].j] = ')'; }
  }
  if (used1) {                                // chosen accessible regions (:1263-1268)
    std::memset(used1, 0, (size_t)m->n1);
    for (size_t k = 0; k < m->vs.size(); k++)
      if (val[m->vcol[k]] > 0.5)
        for (int i = m->vs[k].i; i <= m->vs[k].j; i++) used1[i] = 1;
  }
  if (used2) {
    std::memset(used2, 0, (size_t)m->n2);
    for (size_t k = 0; k < m->ws.size(); k++)
      if (val[m->wcol[k]] > 0.5)
        for (int i = m->ws[k].i; i <= m->ws[k].j; i++) used2[i] = 1;
  }
  return RP_OK;
}

void rp_ip_free(rp_ip_model* m) { delete m; }

}  // extern "C"
