// eval.cpp -- free energy of a given secondary structure under the integer energy tables (host code).
//
// RactIP evaluates the predicted joint structure with ViennaRNA's energy_of_structure(seq, structure, -1)
// per strand (reference src/ractip.cpp:1254,1299,1457) and, for the interaction, on the concatenation
// with cut_point = |s1|+1 (energy_of_duplex, :1528-1559).  ViennaRNA is not vendored in the reference;
// this restates its evaluation at the settings RactIP runs with (temperature 37, dangles = 2,
// tetra_loop on, logML off) [VRNA-recalled, SURVEY.md Appendix A.3 / B]:
//   * integer tables as scale_parameters() leaves them at 37 C: dangles, mismatchM and mismatchExt are
//     clipped to <= 0, everything else is taken as is;
//   * loop decomposition: hairpin, interior (stack / bulge / 1x1 / 1x2 / 2x2 / 1xn / 2x3 / generic with
//     the ninio asymmetry term), multiloop (closing + one ML stem term per branch + ML_BASE per unpaired
//     base), exterior loop; with dangles = 2 every stem of an exterior or multi-loop takes both
//     neighbouring bases when they exist (and lie on the stem's strand);
//   * two strands: a loop that contains the nick is scored as an exterior loop (its closing pair seen
//     from inside counts as one more exterior stem), and DuplexInit is added once when at least one
//     pair joins the strands.
// The same loop model, exponentiated, is what the partition-function kernels sum over; tests check
// sum_S exp(-E(S)/kT) against the oracle's Z by exhaustive enumeration (pf_smooth = 0).
#include <cmath>
#include <cstring>
#include <string>
#include <vector>

#include "ractip_ip.h"
#include "seq_encode.h"

namespace {

struct Eval {
  const rp_model& P;
  int n, cp;                 // cp: first position of strand 2 (1-based), 0 = single strand
  std::vector<int> S;        // S[1..n] in 0..4, S[0] = S[n+1] = 0
  std::vector<int> pt;       // pair table, 1-based, 0 = unpaired

  int pair_type(int i, int j) const {
    static const int T[5][5] = {{0, 0, 0, 0, 0}, {0, 0, 0, 0, 5}, {0, 0, 0, 1, 0}, {0, 0, 2, 0, 3}, {0, 6, 0, 4, 0}};
    const int t = T[S[i]][S[j]];
    return t ? t : 7;        // a pair the structure demands although it is not canonical
  }
  bool same_strand(int a, int b) const { return cp <= 0 || a >= cp || b < cp; }   // a < b
  int d5(int t, int x) const { return P.dangle5_37[t][x] < 0 ? P.dangle5_37[t][x] : 0; }
  int d3(int t, int x) const { return P.dangle3_37[t][x] < 0 ? P.dangle3_37[t][x] : 0; }

  int ext_stem(int type, int s5, int s3) const {
    int e = 0;
    if (s5 >= 0 && s3 >= 0) e += P.mismatchExt37[type][s5][s3] < 0 ? P.mismatchExt37[type][s5][s3] : 0;
    else if (s5 >= 0) e += d5(type, s5);
    else if (s3 >= 0) e += d3(type, s3);
    if (type > 2) e += P.TerminalAU37;
    return e;
  }
  int ml_stem(int type, int s5, int s3) const {
    int e = 0;
    if (s5 >= 0 && s3 >= 0) e += P.mismatchM37[type][s5][s3] < 0 ? P.mismatchM37[type][s5][s3] : 0;
    else if (s5 >= 0) e += d5(type, s5);
    else if (s3 >= 0) e += d3(type, s3);
    if (type > 2) e += P.TerminalAU37;
    return e + P.ML_intern37;
  }
  int hairpin(int i, int j, const std::string& seq) const {
    const int u = j - i - 1, type = pair_type(i, j);
    int e = u <= 30 ? P.hairpin37[u] : P.hairpin37[30] + (int)(P.lxc37 * std::log(u / 30.));
    if (u < 3) return e;
    if (P.special_hp && (u == 3 || u == 4 || u == 6)) {
      const std::string loop = seq.substr((size_t)i - 1, (size_t)u + 2);
      const char* list = u == 4 ? P.Tetraloops : (u == 6 ? P.Hexaloops : P.Triloops);
      const int* en = u == 4 ? P.Tetraloop37 : (u == 6 ? P.Hexaloop37 : P.Triloop37);
      const char* hit = std::strstr(list, loop.c_str());
      if (hit) return en[(hit - list) / (u + 3)];
      if (u == 3) return e + (type > 2 ? P.TerminalAU37 : 0);
    } else if (u == 3) {
      return e + (type > 2 ? P.TerminalAU37 : 0);
    }
    return e + P.mismatchH37[type][S[i + 1]][S[j - 1]];
  }
  // E_IntLoop(n1, n2, type, type_2, si1, sj1, sp1, sq1) as src/pf_duplex.c:153-154 calls it
  int int_loop(int n1, int n2, int type, int type2, int si1, int sj1, int sp1, int sq1) const {
    const int nl = n1 > n2 ? n1 : n2, ns = n1 > n2 ? n2 : n1;
    if (nl == 0) return P.stack37[type][type2];
    if (ns == 0) {
      int e = nl <= 30 ? P.bulge37[nl] : P.bulge37[30] + (int)(P.lxc37 * std::log(nl / 30.));
      if (nl == 1) e += P.stack37[type][type2];
      else {
        if (type > 2) e += P.TerminalAU37;
        if (type2 > 2) e += P.TerminalAU37;
      }
      return e;
    }
    auto ninio = [&](int a, int b) { const int x = (a > b ? a - b : b - a) * P.ninio37; return x < P.MAX_NINIO ? x : P.MAX_NINIO; };
    if (ns == 1) {
      if (nl == 1) return P.int11_37[type][type2][si1][sj1];
      if (nl == 2) return n1 == 1 ? P.int21_37[type][type2][si1][sq1][sj1] : P.int21_37[type2][type][sq1][si1][sp1];
      int e = nl + 1 <= 30 ? P.internal_loop37[nl + 1] : P.internal_loop37[30] + (int)(P.lxc37 * std::log((nl + 1) / 30.));
      return e + ninio(nl, ns) + P.mismatch1nI37[type][si1][sj1] + P.mismatch1nI37[type2][sq1][sp1];
    }
    if (ns == 2) {
      if (nl == 2) return P.int22_37[type][type2][si1][sp1][sq1][sj1];
      if (nl == 3) return P.internal_loop37[5] + P.ninio37 + P.mismatch23I37[type][si1][sj1] + P.mismatch23I37[type2][sq1][sp1];
    }
    const int u = nl + ns;
    int e = u <= 30 ? P.internal_loop37[u] : P.internal_loop37[30] + (int)(P.lxc37 * std::log(u / 30.));
    return e + ninio(nl, ns) + P.mismatchI37[type][si1][sj1] + P.mismatchI37[type2][sq1][sp1];
  }

  // a loop scored as an exterior loop: the exterior loop itself (i = 0) or the loop closed by
  // (i, pt[i]) that contains the nick
  int ext_loop(int i) const {
    int e = 0;
    const int j = i > 0 ? pt[i] : n + 1;
    if (i > 0) {   // the closing pair, seen from inside the loop
      const int tt = pair_type(j, i);
      e += ext_stem(tt, same_strand(j - 1, j) ? S[j - 1] : -1, same_strand(i, i + 1) ? S[i + 1] : -1);
    }
    for (int p = i + 1; p < j; p++) {
      if (!pt[p]) continue;
      const int q = pt[p];
      e += ext_stem(pair_type(p, q), (p > 1 && same_strand(p - 1, p)) ? S[p - 1] : -1,
                    (q < n && same_strand(q, q + 1)) ? S[q + 1] : -1);
      p = q;
    }
    return e;
  }
  int ml_loop(int i) const {
    const int j = pt[i];
    int e = P.ML_closing37 + ml_stem(pair_type(j, i), S[j - 1], S[i + 1]);
    int unpaired = 0;
    for (int p = i + 1; p < j; p++) {
      if (!pt[p]) { unpaired++; continue; }
      const int q = pt[p];
      e += ml_stem(pair_type(p, q), S[p - 1], S[q + 1]);
      p = q;
    }
    return e + unpaired * P.ML_BASE37;
  }
  bool nick_in_loop(int i) const {   // does the loop closed by (i, pt[i]) contain the nick directly?
    if (cp <= 0) return false;
    const int j = pt[i];
    if (!(i < cp && j >= cp)) return false;
    for (int p = i + 1; p < j; p++) {
      if (!pt[p]) continue;
      const int q = pt[p];
      if (p < cp && q >= cp) return false;   // the nick is inside this branch
      p = q;
    }
    return true;
  }
  int loop_energy(int i, const std::string& seq) const {   // the loop closed by (i, pt[i])
    const int j = pt[i];
    int branches = 0, p1 = 0, q1 = 0;
    for (int p = i + 1; p < j; p++) {
      if (!pt[p]) continue;
      if (!branches) { p1 = p; q1 = pt[p]; }
      branches++;
      p = pt[p];
    }
    if (nick_in_loop(i)) return ext_loop(i);
    if (branches == 0) return hairpin(i, j, seq);
    if (branches == 1)
      return int_loop(p1 - i - 1, j - q1 - 1, pair_type(i, j), pair_type(q1, p1), S[i + 1], S[j - 1], S[p1 - 1], S[q1 + 1]);
    return ml_loop(i);
  }
};

}  // namespace

extern "C" int rp_energy_of_structure(const rp_model* model, const char* seq, const char* structure, int n, int cut_point,
                                      float* energy) {
  if (!model || !seq || !structure || n < 1 || !energy) return RP_ERR_ARG;
  if (model->temperature != 37.0 || model->dangles != 2) return RP_ERR_UNSUPPORTED;
  Eval E{*model, n, cut_point > 1 && cut_point <= n ? cut_point : 0, std::vector<int>((size_t)n + 2, 0),
         std::vector<int>((size_t)n + 2, 0)};
  std::string s((size_t)n, 'N');
  for (int i = 1; i <= n; i++) {
    const int code = rp::encode_base(seq[i - 1]);
    E.S[i] = code & 7;
    s[(size_t)i - 1] = (code & 8) ? 'X' : "NACGU"[E.S[i]];   // a window with a letter outside ACGU never matches a special hairpin (string compare)
  }
  std::vector<int> stack;
  for (int i = 1; i <= n; i++) {
    if (structure[i - 1] == '(') stack.push_back(i);
    else if (structure[i - 1] == ')') {
      if (stack.empty()) return RP_ERR_FORMAT;
      E.pt[i] = stack.back(); E.pt[stack.back()] = i;
      stack.pop_back();
    }
  }
  if (!stack.empty()) return RP_ERR_FORMAT;
  int e = E.ext_loop(0);
  bool joined = false;
  for (int i = 1; i <= n; i++)
    if (E.pt[i] > i) {
      e += E.loop_energy(i, s);
      if (E.cp > 0 && i < E.cp && E.pt[i] >= E.cp) joined = true;
    }
  if (joined) e += model->DuplexInit37;
  *energy = (float)e / 100.f;
  return RP_OK;
}

extern "C" int rp_energy_of_duplex(const rp_model* model, const char* s1, int n1, const char* s2, int n2, const char* r1,
                                   const char* r2, float* energy) {
  if (!s1 || !s2 || !r1 || !r2 || n1 < 1 || n2 < 1) return RP_ERR_ARG;
  // src/ractip.cpp:1533-1546: concatenate; '(' ')' -> '.', '[' -> '(', ']' -> ')'
  std::string ss = std::string(s1, (size_t)n1) + std::string(s2, (size_t)n2);
  std::string rr = std::string(r1, (size_t)n1) + std::string(r2, (size_t)n2);
  for (char& x : rr) {
    switch (x) {
      case '(': case ')': x = '.'; break;
      case '[': x = '('; break;
      case ']': x = ')'; break;
      default: break;
    }
  }
  return rp_energy_of_structure(model, ss.c_str(), rr.c_str(), n1 + n2, n1 + 1, energy);
}
