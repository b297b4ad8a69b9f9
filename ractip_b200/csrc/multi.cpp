// multi.cpp -- all GPUs of one box behind the C ABI (host code over the single-device entry points).
//
// The z-score loop of the reference (src/ractip.cpp:1638-1657) is a sequence of independent solve() calls;
// a RactIP build that hands the whole shuffle batch to the probability stage can spread it
// over every visible GPU without a process per device: one rp_ctx per device, one host thread each, the
// batch cut into contiguous blocks (pair k of n goes to device k*G/n), every device writing its block
// straight into the caller's one host buffer -- the dense sections and the record lists of a block are
// contiguous in the layouts of rp_dense_plan / rp_sparse_plan, so no result is copied twice and no
// collective is needed (results end in host memory, where the integer programmes run).
#include <algorithm>
#include <string>
#include <thread>
#include <vector>

#include <cuda_runtime.h>

#include "ractip_prob.h"

struct rp_multi {
  std::vector<rp_ctx*> ctx;
  std::vector<int> dev;
  std::string err;
};

namespace {
thread_local std::string g_multi_err;
int fail(rp_multi* m, int code, const std::string& msg) {
  (m ? m->err : g_multi_err) = msg;
  return code;
}
// contiguous block of device r
void block(int n, int G, int r, int& lo, int& hi) {
  lo = (int)((long long)n * r / G);
  hi = (int)((long long)n * (r + 1) / G);
}
}  // namespace

extern "C" {

int rp_multi_create(rp_multi** out, const rp_model* model, const int* devices, int n_devices) {
  if (!out || !model) return fail(nullptr, RP_ERR_ARG, "rp_multi_create: null argument");
  *out = nullptr;
  int visible = 0;
  if (cudaGetDeviceCount(&visible) != cudaSuccess || visible < 1)
    return fail(nullptr, RP_ERR_NO_DEVICE, "rp_multi_create: no CUDA device; the probability stage has no CPU fallback");
  std::vector<int> dev;
  if (devices && n_devices > 0) dev.assign(devices, devices + n_devices);
  else {
    const int g = n_devices > 0 ? std::min(n_devices, visible) : visible;
    for (int d = 0; d < g; d++) dev.push_back(d);
  }
  rp_multi* m = new rp_multi;
  for (int d : dev) {
    rp_ctx* c = nullptr;
    const int rc = rp_create(&c, model, d);
    if (rc) {
      const std::string msg = std::string("rp_multi_create: device ") + std::to_string(d) + ": " + rp_last_error(nullptr);
      for (rp_ctx* x : m->ctx) rp_destroy(x);
      delete m;
      return fail(nullptr, rc, msg);
    }
    m->ctx.push_back(c);
    m->dev.push_back(d);
  }
  *out = m;
  return RP_OK;
}

int rp_multi_destroy(rp_multi* m) {
  if (!m) return RP_OK;
  for (rp_ctx* c : m->ctx) rp_destroy(c);
  delete m;
  return RP_OK;
}

int rp_multi_devices(const rp_multi* m) { return m ? (int)m->ctx.size() : 0; }
const char* rp_multi_last_error(const rp_multi* m) { return m ? m->err.c_str() : g_multi_err.c_str(); }

int rp_multi_run_dense(rp_multi* m, const rp_pair* pairs, int n_pairs, const rp_opts* opts, float* out, size_t out_floats) {
  if (!m || !out) return fail(m, RP_ERR_ARG, "rp_multi_run_dense: null argument");
  std::vector<rp_dense_layout> lay((size_t)std::max(n_pairs, 1));
  size_t total = 0;
  int rc = rp_dense_plan(pairs, n_pairs, opts, lay.data(), &total);
  if (rc) return fail(m, rc, "rp_multi_run_dense: bad pairs or options");
  if (out_floats < total) return fail(m, RP_ERR_CAPACITY, "rp_multi_run_dense: buffer too small");
  const int G = (int)m->ctx.size();
  std::vector<int> rcs((size_t)G, RP_OK);
  std::vector<std::thread> th;
  for (int r = 0; r < G; r++) {
    int lo, hi;
    block(n_pairs, G, r, lo, hi);
    if (hi <= lo) continue;
    const size_t first = lay[lo].bp1, end = hi < n_pairs ? lay[hi].bp1 : total;
    th.emplace_back([=, &rcs]() { rcs[r] = rp_run_dense(m->ctx[r], pairs + lo, hi - lo, opts, out + first, end - first); });
  }
  for (auto& t : th) t.join();
  for (int r = 0; r < G; r++)
    if (rcs[r]) return fail(m, rcs[r], std::string("rp_multi_run_dense: device ") + std::to_string(m->dev[r]) + ": " + rp_last_error(m->ctx[r]));
  return RP_OK;
}

int rp_multi_run_sparse(rp_multi* m, const rp_pair* pairs, int n_pairs, const rp_opts* opts, rp_rec* recs, size_t n_recs,
                        rp_sparse_counts* counts) {
  if (!m || !recs || !counts) return fail(m, RP_ERR_ARG, "rp_multi_run_sparse: null argument");
  std::vector<rp_sparse_layout> lay((size_t)std::max(n_pairs, 1));
  size_t total = 0, total_f = 0;
  int rc = rp_sparse_plan(pairs, n_pairs, opts, lay.data(), &total, &total_f);
  if (rc) return fail(m, rc, "rp_multi_run_sparse: bad pairs or options");
  if (n_recs < total) return fail(m, RP_ERR_CAPACITY, "rp_multi_run_sparse: buffer too small");
  const int G = (int)m->ctx.size();
  std::vector<int> rcs((size_t)G, RP_OK);
  std::vector<std::thread> th;
  for (int r = 0; r < G; r++) {
    int lo, hi;
    block(n_pairs, G, r, lo, hi);
    if (hi <= lo) continue;
    const size_t first = lay[lo].x, end = hi < n_pairs ? lay[hi].x : total;
    th.emplace_back([=, &rcs]() {
      rcs[r] = rp_run_sparse(m->ctx[r], pairs + lo, hi - lo, opts, recs + first, end - first, nullptr, 0, counts + lo);
    });
  }
  for (auto& t : th) t.join();
  for (int r = 0; r < G; r++)
    if (rcs[r]) return fail(m, rcs[r], std::string("rp_multi_run_sparse: device ") + std::to_string(m->dev[r]) + ": " + rp_last_error(m->ctx[r]));
  return RP_OK;
}

}  // extern "C"
