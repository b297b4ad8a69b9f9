// Compiles and exercises the C++ shim (include/ractip_prob.hpp) the way a RactIP build would.
//   shim_check host          : host-only entry points; the stage must refuse to exist without a GPU
//   shim_check gpu S1 S2     : fills RactIP's members for one pair and prints a few numbers
//   shim_check multi S1 S2 N : the z-score shuffle batch (N shuffles) on every visible GPU through rp_multi_*, compared
//                              with the same batch on one GPU
#include <cstdio>
#include <cstdlib>
#include <cstring>
#include <string>

#include "ractip_prob.hpp"

int main(int argc, char** argv) {
  const std::string mode = argc > 1 ? argv[1] : "host";
  std::vector<std::pair<std::string, std::string> > sh;
  rp::zscore_shuffles("GAAAGACGCGCAUUUGUUAUCAUCAUCCCUGAAUUCAGAGAUGAAAUUUUGGCCACUCACGAGUGGCCUUUU", "GCCAGGGGUGCUCGGCAUAAGCCGAAGAUAUCGG", 12, 1, 3, sh);
  if (sh.size() != 3 || sh[0].first.size() != 72) return 2;
  std::printf("shuffle0 %s\n", sh[0].first.c_str());
  if (mode == "host") {
    try {
      rp::ProbabilityStage st;
      std::printf("stage created (a GPU is present)\n");
    } catch (const std::exception& e) {
      std::printf("no stage: %s\n", e.what());
      if (!std::strstr(e.what(), "no CPU fallback")) return 3;
    }
    return 0;
  }
  if (argc < 4) return 4;
  if (mode == "multi") {
    const int num = argc > 4 ? std::atoi(argv[4]) : 16;
    std::vector<std::pair<std::string, std::string> > batch;
    rp::zscore_shuffles(argv[2], argv[3], 12, 1, num, batch);
    std::vector<rp::PairProbabilities> one, all;
    {
      rp::ProbabilityStage single;
      single.solve_batch(batch, one);
    }
    rp::ProbabilityStage many(nullptr, 0, 0);
    many.solve_batch(batch, all);
    std::printf("devices %d pairs %d\n", many.devices(), num);
    if (one.size() != all.size()) return 7;
    for (size_t k = 0; k < one.size(); k++)
      if (one[k].bp1 != all[k].bp1 || one[k].bp2 != all[k].bp2 || one[k].up1 != all[k].up1 || one[k].up2 != all[k].up2 ||
          one[k].hp != all[k].hp || one[k].offset2 != all[k].offset2)
        return 8;
    std::printf("multi == single\n");
    return 0;
  }
  rp::ProbabilityStage st;
  rp::PairProbabilities r;
  st.solve_probabilities(argv[2], argv[3], r);
  const int L1 = static_cast<int>(std::strlen(argv[2]));
  double sbp = 0, sup = 0, shp = 0;
  for (float v : r.bp1) sbp += v;
  for (auto& row : r.up1) for (float v : row) sup += v;
  for (auto& row : r.hp) for (float v : row) shp += v;
  std::printf("L1 %d offset1[1] %d bp1[offset1[1]+%d] %.9g sum_bp1 %.9g sum_up1 %.9g sum_hp %.9g\n", L1, r.offset1[1], L1,
              r.bp1[r.offset1[1] + L1], sbp, sup, shp);
  rp::VF bp; rp::VI off; rp::VVF up;
  st.rnafold(argv[2], bp, off, up, 15);
  if (bp != r.bp1 || off != r.offset1 || up != r.up1) return 5;
  rp::VVF hp;
  st.rnaduplex(argv[2], argv[3], hp);
  if (hp != r.hp) return 6;
  return 0;
}
