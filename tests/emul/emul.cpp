// emul.cpp -- TEST INFRASTRUCTURE.  Runs the kernel's per-thread phase
// functions (ractip_b200/csrc/mcc_core.h, mcc_driver.h) on the host, one
// emulated thread after another, so that the CUDA kernel's logic can be checked
// against the oracle on a CPU-only box.  Never loaded by the product.
#include <cstdlib>
#include <cstring>
#include <vector>

#include "dev_model.h"
#include "mcc_driver.h"
#include "seq_encode.h"

namespace {
struct SerialExec {
  int T;
  int nthreads() const { return T; }
  template <class F>
  void phase(int, F f) {
    for (int t = 0; t < T; t++) f(t);
  }
};
}  // namespace

extern "C" int emul_problem(const rp_model* m, const char* seq, int n, int cp, int kind, int max_w, int n1, int n2,
                            float th_hy, int T, float* bp, float* up, float* hp, double* logz) {
  static rp::DevModel M;  // large
  int rc = rp::build_dev_model(*m, &M);
  if (rc) return rc;
  std::vector<uint8_t> S(n + 16, 0);
  for (int i = 1; i <= n; i++) S[i] = rp::encode_base(seq[i - 1]);
  rp::Problem p;
  std::memset(&p, 0, sizeof p);
  p.seq_off = 0; p.n = n; p.cp = cp; p.kind = kind; p.pair = 0; p.which = 0; p.max_w = max_w;
  p.n1 = n1; p.n2 = n2; p.th_hy = th_hy;
  // lay the three outputs out in one float buffer
  size_t nbp = (size_t)(n + 1) * (n + 2) / 2, nup = (size_t)n * (max_w > 0 ? max_w : 0), nhp = (size_t)(n1 + 1) * (n2 + 1);
  std::vector<float> dense(nbp + nup + nhp + 8, 0.f);
  p.out_bp = (kind == rp::KIND_LINEAR && bp) ? 0 : -1;
  p.out_up = (kind == rp::KIND_LINEAR && up && max_w > 0) ? (long long)nbp : -1;
  p.out_hp = (kind == rp::KIND_COFOLD && hp) ? (long long)(nbp + nup) : -1;
  std::vector<double> ws(rp::slot_doubles(n), 1e300);  // poison: stale data must never be read
  std::vector<double> smem(rp::shared_bytes(T) / sizeof(double) + 2, 0.0);
  rp::Shared sh;
  rp::carve_shared(sh, smem.data(), T);
  double lz[3] = {0, 0, 0};
  rp::Ctx c;
  rp::bind_ctx(c, &M, S.data(), p, ws.data());
  SerialExec ex{T};
  rp::solve_mcc(ex, c, p, dense.data(), lz, sh);
  if (p.out_bp >= 0) std::memcpy(bp, dense.data(), nbp * sizeof(float));
  if (p.out_up >= 0) std::memcpy(up, dense.data() + nbp, nup * sizeof(float));
  if (p.out_hp >= 0) std::memcpy(hp, dense.data() + nbp + nup, nhp * sizeof(float));
  if (logz) *logz = lz[0];
  return 0;
}
