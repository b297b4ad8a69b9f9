// api.cu -- the C ABI of include/ractip_prob.h over the CUDA kernels.
//
// Host-side mirror of what RactIP::solve does before it builds the IP
// (reference src/ractip.cpp:546-548): for every pair, two single-strand
// problems (rnafold: pair probabilities + unpaired windows) and one two-strand
// problem (rnaduplex), batched over all pairs of a call -- the --zscore shuffle
// loop (src/ractip.cpp:1638-1657) hands its whole batch over at once.
//
// There is no CPU fallback: every compute entry point needs a CUDA device.
#include <cuda_runtime.h>

#include <algorithm>
#include <cmath>
#include <cstdio>
#include <cstdlib>
#include <cstring>
#include <numeric>
#include <string>
#include <utility>
#include <vector>

#include "dev_model.h"
#include "kernels.h"
#include "ractip_prob.h"
#include "seq_encode.h"

using rp::Problem;

struct rp_ctx {
  int device = 0;
  int sm_count = 0;
  int ctas_per_sm = 0;       // general kernel, 64-register build (two CTAs per SM)
  int ctas_per_sm1 = 0;      // general kernel, 128-register build (one CTA per SM: long problems)
  int mcc_long_n = 700;      // problems at least this long run the 128-register build (RP_MCC_LONG_N)
  int mcc_wide = 10;         // ... with split-sum bands of this many diagonals (RP_MCC_WIDE: 5, 10, 15)
  // Band kernel (mcc_band.h): problems whose 32-diagonal ring fits in shared memory.  Class L: 512
  // threads, one CTA per SM; class S: 256 threads, two CTAs per SM.  RP_BAND=0 disables it (A/B aid).
  bool band = true;
  size_t smem_optin = 0;       // max dynamic shared memory per CTA
  int threads = RP_MCC_THREADS;  // CTA width of the wavefront kernel (RP_MCC_THREADS env var narrows it: tuning aid)
  rp::DevModel* d_model = nullptr;
  cudaStream_t own_stream = nullptr;
  cudaStream_t side_stream = nullptr;   // the second band launch shape runs here, so that its CTAs fill the SMs the first one leaves
  cudaEvent_t ev_fork = nullptr, ev_join = nullptr;
  cudaStream_t copy_stream = nullptr;   // rp_batch_fetch_dense: outputs of the long band class leave while the short class still runs
  cudaEvent_t ev_main = nullptr, ev_copy = nullptr;
  cudaStream_t stream = nullptr;
  cudaEvent_t ev[6] = {};
  cudaEvent_t ev_dom[2] = {};   // around the launch that carries most of the algorithmic flops (rp_timing::ms_dominant)
  bool dom_timed = false;
  double* ws = nullptr;       // workspace slots (grow-only)
  size_t ws_bytes = 0;
  rp_timing timing = {};
  bool timing_pending = false;
  bool timed_copies = false;
  std::string err;
  // device buffers of destroyed batches, reused by the next batch (cudaMalloc/cudaFree cost
  // milliseconds and cudaFree synchronises the device: fatal for the one-shot host calls)
  std::vector<std::pair<void*, size_t>> pool_free;
  std::vector<std::pair<void*, size_t>> pool_live;
  // lifetime: batches hold a reference; rp_destroy with batches alive only marks the context closed, the last
  // rp_batch_destroy tears it down (a batch handle never outlives the memory it points into)
  int up_ctas_per_sm = 0;             // unstru_kernel occupancy
  int live_batches = 0;
  bool closing = false;
  cudaStream_t ws_stream = nullptr;   // stream of the last launch that used the workspace / a pooled buffer
  long long* d_prof = nullptr;        // RP_PROFILE counters (per context: per device)
};

struct rp_batch {
  rp_ctx* ctx = nullptr;
  int n_pairs = 0;
  rp_opts opts = {};
  std::vector<Problem> probs;
  std::vector<int> order;
  std::vector<rp_dense_layout> layout;
  size_t total_floats = 0;
  int maxn = 0;
  int n_mcc = 0, n_duplex = 0;
  size_t slot_doubles = 0;
  double alg_flops = 0;
  // device
  uint8_t* d_seq = nullptr;
  Problem* d_probs = nullptr;
  int* d_order = nullptr;
  int mcc_minb = 2;           // register budget of the general kernel for this batch (kernels.h)
  int grid_cached = -1;       // launch shape, fixed at the first run (cudaMemGetInfo is slow)
  int n_general = 0;          // problems left to the general kernel + duplex (entries of order after the band classes)
  // band classes: order = [class L | class S | general]
  int n_band[2] = {0, 0};
  int band_maxn[2] = {0, 0};
  int band_grid[2] = {-1, -1};
  int* d_counter = nullptr;
  float* d_dense = nullptr;
  // Split fetch (uniform batches whose problems fall into both band classes, nothing else launched): sections of a
  // pair's dense block written by the long class / by the short class, as (offset, length) in floats within pair 0's
  // block; every pair has the same block layout, `pair_stride` floats apart.
  bool split_fetch = false;
  size_t pair_stride = 0;
  std::vector<std::pair<size_t, size_t>> sect_long, sect_short;
  double* d_logz = nullptr;
  // sparse
  std::vector<rp_sparse_layout> slayout;
  size_t total_recs = 0, total_upf = 0;
  // deferred unpaired-window passes (unstru_kernel): queue entries appended to `order`, private workspaces
  int n_defer = 0;
  size_t defer_slot = 0;      // doubles per private workspace
  int defer_maxn = 0;
  int* d_done = nullptr;      // completion flags of the deferred problems (indexed by problem)
  int up_mode = 0;            // 0: pass fused with the wavefronts, 1: unstru_kernel after them, 2: jobs of their own inside the band kernel
  int up_grid = -1;
  int dom_kind = -1;          // launch with the largest share of the algorithmic flops: 0 / 1 band class, 2 general
  double dom_flops = 0;
  bool has_single = false;              // some pair has n2 == 0: its unused output sections are zero-filled once
  cudaStream_t last_stream = nullptr;   // stream the batch last ran on (rp_set_stream may have moved the context on)
  rp::SparsePair* d_spairs = nullptr;
  rp_rec* d_recs = nullptr;
  float* d_ups = nullptr;
  rp_sparse_counts* d_counts = nullptr;
};

namespace {

thread_local std::string g_err;

int fail(rp_ctx* ctx, int code, const std::string& msg) {
  if (ctx) ctx->err = msg;
  g_err = msg;
  return code;
}

#define CU(call)                                                                              \
  do {                                                                                        \
    cudaError_t e__ = (call);                                                                 \
    if (e__ != cudaSuccess)                                                                   \
      return fail(ctx, RP_ERR_CUDA, std::string(#call) + ": " + cudaGetErrorString(e__));     \
  } while (0)

size_t bp_len(int L) { return (size_t)(L + 1) * (size_t)(L + 2) / 2; }

int check_pairs(const rp_pair* pairs, int n_pairs, const rp_opts* opts) {
  if (!pairs || !opts || n_pairs < 0) return RP_ERR_ARG;
  for (int p = 0; p < n_pairs; p++)
    if (!pairs[p].s1 || pairs[p].n1 < 1 || pairs[p].n2 < 0 || (pairs[p].n2 > 0 && !pairs[p].s2)) return RP_ERR_ARG;   // n2 == 0: s1 alone
  return RP_OK;
}

int ensure_workspace(rp_ctx* ctx, size_t bytes) {
  if (bytes <= ctx->ws_bytes) return RP_OK;
  if (ctx->ws) {
    if (ctx->ws_stream && ctx->ws_stream != ctx->stream) CU(cudaStreamSynchronize(ctx->ws_stream));
    CU(cudaStreamSynchronize(ctx->stream));
    CU(cudaFree(ctx->ws));
    ctx->ws = nullptr;
    ctx->ws_bytes = 0;
  }
  cudaError_t e = cudaMalloc(&ctx->ws, bytes);
  if (e != cudaSuccess) {   // give the cached buffers of destroyed batches back to the driver and retry once
    cudaGetLastError();
    for (auto& f : ctx->pool_free) cudaFree(f.first);
    ctx->pool_free.clear();
    e = cudaMalloc(&ctx->ws, bytes);
  }
  if (e != cudaSuccess) { ctx->ws = nullptr; return fail(ctx, RP_ERR_CUDA, std::string("cudaMalloc workspace: ") + cudaGetErrorString(e)); }
  ctx->ws_bytes = bytes;
  return RP_OK;
}

size_t pool_cached_bytes(const rp_ctx* ctx) {
  size_t t = 0;
  for (const auto& f : ctx->pool_free) t += f.second;
  return t;
}

size_t pool_cached_bytes(const rp_ctx* ctx);
cudaError_t pool_alloc(rp_ctx* ctx, void** out, size_t bytes) {
  bytes = std::max<size_t>(bytes, 256);
  int best = -1;
  for (size_t k = 0; k < ctx->pool_free.size(); k++)
    if (ctx->pool_free[k].second >= bytes && (best < 0 || ctx->pool_free[k].second < ctx->pool_free[best].second)) best = (int)k;
  if (best >= 0 && ctx->pool_free[best].second <= 4 * bytes + (1 << 20)) {
    *out = ctx->pool_free[best].first;
    ctx->pool_live.push_back(ctx->pool_free[best]);
    ctx->pool_free.erase(ctx->pool_free.begin() + best);
    return cudaSuccess;
  }
  cudaError_t e = cudaMalloc(out, bytes);
  if (e != cudaSuccess) {  // give cached buffers back to the driver and retry once
    for (auto& f : ctx->pool_free) cudaFree(f.first);
    ctx->pool_free.clear();
    e = cudaMalloc(out, bytes);
  }
  if (e == cudaSuccess) ctx->pool_live.emplace_back(*out, bytes);
  return e;
}
template <class T>
cudaError_t pool_alloc(rp_ctx* ctx, T** out, size_t bytes) {
  return pool_alloc(ctx, reinterpret_cast<void**>(out), bytes);
}
void pool_release(rp_ctx* ctx, void* p) {
  if (!p) return;
  for (size_t k = 0; k < ctx->pool_live.size(); k++)
    if (ctx->pool_live[k].first == p) {
      ctx->pool_free.push_back(ctx->pool_live[k]);
      ctx->pool_live.erase(ctx->pool_live.begin() + k);
      // keep the cache bounded: beyond 64 buffers or 8 GB the oldest entries go back to the driver
      while (ctx->pool_free.size() > 64 || pool_cached_bytes(ctx) > ((size_t)8 << 30)) {
        cudaFree(ctx->pool_free.front().first);
        ctx->pool_free.erase(ctx->pool_free.begin());
      }
      return;
    }
  cudaFree(p);
}

// number of CTAs (= workspace slots) for a batch
int grid_for(const rp_ctx* ctx, int nprob, size_t slot_bytes, int minb) {
  int g = ctx->sm_count * std::max(1, minb >= 2 ? ctx->ctas_per_sm : ctx->ctas_per_sm1);
  g = std::min(g, std::max(1, nprob));
  if (const char* e = std::getenv("RP_GRID")) {  // tuning aid
    int v = std::atoi(e);
    if (v >= 1 && v < g) g = v;
  }
  // keep the workspace within a budget (long sequences: fewer, fatter slots).  cudaMemGetInfo costs
  // milliseconds: not asked when the workspace the context already holds is large enough.
  size_t free_b = 0, total_b = 0;
  if ((size_t)g * slot_bytes > ctx->ws_bytes && cudaMemGetInfo(&free_b, &total_b) == cudaSuccess) {
    size_t budget = (free_b + ctx->ws_bytes + pool_cached_bytes(ctx)) / 10 * 7;   // cached buffers are freed on demand
    while (g > 1 && (size_t)g * slot_bytes > budget) g--;
  }
  return g;
}

}  // namespace

extern "C" {

const char* rp_version(void) { return "ractip_b200 0.2 (sm_100a)"; }

int rp_kernel_plan(int n, size_t smem_limit, size_t* smem_bytes) {
  if (smem_limit == 0) smem_limit = 232448;
  const size_t half_sm = (smem_limit + 1024) / 2 - 1024;   // two CTAs per SM, 1 KB reserved each
  if (n < 1) n = 1;
  const size_t s256 = rp::band_shared_bytes(n, 256), s512 = rp::band_shared_bytes(n, 512);
  if (s256 <= half_sm) { if (smem_bytes) *smem_bytes = s256; return RP_KERNEL_BAND_2CTA; }
  if (smem_bytes) *smem_bytes = s512;
  if (s512 <= smem_limit) return RP_KERNEL_BAND_1CTA;
  return n >= 700 ? RP_KERNEL_GENERAL_WIDE : RP_KERNEL_GENERAL;   // rp_ctx::mcc_long_n default
}

const char* rp_strerror(int code) {
  switch (code) {
    case RP_OK: return "ok";
    case RP_ERR_ARG: return "bad argument";
    case RP_ERR_NO_DEVICE: return "no usable CUDA device (the probability stage has no CPU fallback)";
    case RP_ERR_CUDA: return "CUDA error";
    case RP_ERR_IO: return "cannot read file";
    case RP_ERR_FORMAT: return "malformed parameter data";
    case RP_ERR_NO_DEFAULTS: return "requested energy tables are not embedded; load a parameter file";
    case RP_ERR_SEQ: return "bad sequence";
    case RP_ERR_TOO_LONG: return "sequence too long";
    case RP_ERR_CAPACITY: return "output buffer too small";
    case RP_ERR_UNSUPPORTED: return "unsupported model setting";
    default: return "unknown error";
  }
}

const char* rp_last_error(const rp_ctx* ctx) { return ctx ? ctx->err.c_str() : g_err.c_str(); }

void rp_opts_default(rp_opts* o) {
  if (!o) return;
  // defaults of src/cmdline.c:151-186 as mapped at src/ractip.cpp:1474-1498
  o->max_w = 15; o->min_w = 5; o->th_ss = 0.5f; o->th_hy = 0.1f; o->th_ac = 0.003f; o->use_pf_duplex = 0;
}

int rp_dense_plan(const rp_pair* pairs, int n_pairs, const rp_opts* opts, rp_dense_layout* layout,
                  size_t* total_floats) {
  int rc = check_pairs(pairs, n_pairs, opts);
  if (rc) return rc;
  const size_t w = opts->max_w > 0 ? (size_t)opts->max_w : 0;
  size_t off = 0;
  for (int p = 0; p < n_pairs; p++) {
    rp_dense_layout l;
    l.n_bp1 = bp_len(pairs[p].n1); l.n_bp2 = bp_len(pairs[p].n2);
    l.n_up1 = (size_t)pairs[p].n1 * w; l.n_up2 = (size_t)pairs[p].n2 * w;
    l.n_hp = (size_t)(pairs[p].n1 + 1) * (size_t)(pairs[p].n2 + 1);
    l.bp1 = off; off += l.n_bp1;
    l.bp2 = off; off += l.n_bp2;
    l.up1 = off; off += l.n_up1;
    l.up2 = off; off += l.n_up2;
    l.hp = off; off += l.n_hp;
    if (layout) layout[p] = l;
  }
  if (total_floats) *total_floats = off;
  return RP_OK;
}

int rp_sparse_plan(const rp_pair* pairs, int n_pairs, const rp_opts* opts, rp_sparse_layout* layout, size_t* total_recs,
                   size_t* total_floats) {
  int rc = check_pairs(pairs, n_pairs, opts);
  if (rc) return rc;
  const size_t w = opts->max_w > 0 ? (size_t)opts->max_w : 0;
  // Each base pairs with total probability <= 1, so fewer than 1/th partners
  // exceed th: at most n/(2 th) pairs inside one RNA, min(n1,n2)/th across.
  auto cap_in = [&](int n) -> size_t {
    size_t all = (size_t)n * (size_t)(n - 1) / 2;
    if (!(opts->th_ss > 0.f)) return all;
    return std::min(all, (size_t)std::floor(n / (2.0 * opts->th_ss)) + 1);
  };
  size_t roff = 0, foff = 0;
  for (int p = 0; p < n_pairs; p++) {
    rp_sparse_layout l;
    l.cap_x = cap_in(pairs[p].n1);
    l.cap_y = cap_in(pairs[p].n2);
    size_t all = (size_t)pairs[p].n1 * (size_t)pairs[p].n2;
    l.cap_z = !(opts->th_hy > 0.f) ? all
                                   : std::min(all, (size_t)std::min(pairs[p].n1, pairs[p].n2) *
                                                       ((size_t)std::floor(1.0 / opts->th_hy) + 1));
    // accessible regions: one variable per (start, length) with length in min_w..max_w, when the
    // reference creates them at all (enable_accessibility, src/ractip.cpp:526)
    const bool acc = opts->min_w > 1 && opts->max_w >= opts->min_w;
    const size_t nlen = acc ? (size_t)(opts->max_w - opts->min_w + 1) : 0;
    l.cap_v = (size_t)pairs[p].n1 * nlen;
    l.cap_w = (size_t)pairs[p].n2 * nlen;
    l.x = roff; roff += l.cap_x;
    l.y = roff; roff += l.cap_y;
    l.z = roff; roff += l.cap_z;
    l.v = roff; roff += l.cap_v;
    l.w = roff; roff += l.cap_w;
    l.n_up1 = (size_t)pairs[p].n1 * w; l.n_up2 = (size_t)pairs[p].n2 * w;
    l.up1 = foff; foff += l.n_up1;
    l.up2 = foff; foff += l.n_up2;
    if (layout) layout[p] = l;
  }
  if (total_recs) *total_recs = roff;
  if (total_floats) *total_floats = foff;
  return RP_OK;
}

static double alg_flops_mcc_uncached(int n);
double rp_alg_flops_mcc(int n) {
  // memoised: the O(n^2) count below used to be recomputed for all 3 problems of every pair of
  // every one-shot call (tens of milliseconds per 1000-pair batch, more than the copies)
  static thread_local std::vector<std::pair<int, double>> memo;
  for (const auto& kv : memo)
    if (kv.first == n) return kv.second;
  const double v = alg_flops_mcc_uncached(n);
  if (memo.size() < 4096) memo.emplace_back(n, v);
  return v;
}
static double alg_flops_mcc_uncached(int n) {
  // F_mcc(n) = 6 I(n) + 2 S_in(n) + 2 S_out(n), TURN=3, MAXLOOP=30
  if (n < 5) return 0.;
  double I = 0, Sin = 0, Sout = 0;
  for (int d = 4; d <= n - 1; d++) {
    const int m = std::min(30, d - 6);
    if (m >= 0) I += (double)(n - d) * (m + 1) * (m + 2) / 2.0;
    Sin += (double)(n - d) * ((d - 2) + d + d);
  }
  for (int l = 5; l <= n; l++)
    for (int k = 2; k <= l - 4; k++) Sout += std::max(0, n - l - 1) + (k - 2);
  return 6. * I + 2. * Sin + 2. * Sout;
}

// ------------------------------------------------------------------ context
int rp_create(rp_ctx** out, const rp_model* m, int device) {
  rp_ctx* ctx = nullptr;
  if (!out || !m) return fail(nullptr, RP_ERR_ARG, "rp_create: null argument");
  *out = nullptr;
  int ndev = 0;
  cudaError_t e = cudaGetDeviceCount(&ndev);
  if (e != cudaSuccess || ndev < 1)
    return fail(nullptr, RP_ERR_NO_DEVICE,
                std::string("rp_create: no CUDA device (") + (e == cudaSuccess ? "count 0" : cudaGetErrorString(e)) +
                    "); the probability stage has no CPU fallback");
  if (device < 0 || device >= ndev) return fail(nullptr, RP_ERR_ARG, "rp_create: bad device index");
  std::vector<rp::DevModel> host(1);
  int rc = rp::build_dev_model(*m, host.data());
  if (rc) return fail(nullptr, rc, "rp_create: unsupported model (temperature must be 37, dangles 2)");
  ctx = new rp_ctx;
  ctx->device = device;
  auto bail = [&](int code, const std::string& msg) {
    rp_destroy(ctx);
    return fail(nullptr, code, msg);
  };
  if ((e = cudaSetDevice(device)) != cudaSuccess) return bail(RP_ERR_CUDA, cudaGetErrorString(e));
  cudaDeviceProp prop;
  if ((e = cudaGetDeviceProperties(&prop, device)) != cudaSuccess) return bail(RP_ERR_CUDA, cudaGetErrorString(e));
  ctx->sm_count = prop.multiProcessorCount;
  ctx->smem_optin = prop.sharedMemPerBlockOptin;
  if (const char* e = std::getenv("RP_BAND")) ctx->band = std::atoi(e) != 0;
  if ((e = cudaStreamCreateWithFlags(&ctx->own_stream, cudaStreamNonBlocking)) != cudaSuccess)
    return bail(RP_ERR_CUDA, cudaGetErrorString(e));
  ctx->stream = ctx->own_stream;
  if ((e = cudaStreamCreateWithFlags(&ctx->side_stream, cudaStreamNonBlocking)) != cudaSuccess)
    return bail(RP_ERR_CUDA, cudaGetErrorString(e));
  if ((e = cudaEventCreateWithFlags(&ctx->ev_fork, cudaEventDisableTiming)) != cudaSuccess) return bail(RP_ERR_CUDA, cudaGetErrorString(e));
  if ((e = cudaEventCreateWithFlags(&ctx->ev_join, cudaEventDisableTiming)) != cudaSuccess) return bail(RP_ERR_CUDA, cudaGetErrorString(e));
  if ((e = cudaStreamCreateWithFlags(&ctx->copy_stream, cudaStreamNonBlocking)) != cudaSuccess) return bail(RP_ERR_CUDA, cudaGetErrorString(e));
  if ((e = cudaEventCreateWithFlags(&ctx->ev_main, cudaEventDisableTiming)) != cudaSuccess) return bail(RP_ERR_CUDA, cudaGetErrorString(e));
  if ((e = cudaEventCreateWithFlags(&ctx->ev_copy, cudaEventDisableTiming)) != cudaSuccess) return bail(RP_ERR_CUDA, cudaGetErrorString(e));
  for (auto& ev : ctx->ev)
    if ((e = cudaEventCreate(&ev)) != cudaSuccess) return bail(RP_ERR_CUDA, cudaGetErrorString(e));
  for (auto& ev : ctx->ev_dom)
    if ((e = cudaEventCreate(&ev)) != cudaSuccess) return bail(RP_ERR_CUDA, cudaGetErrorString(e));
  if ((e = cudaMalloc(&ctx->d_model, sizeof(rp::DevModel))) != cudaSuccess) return bail(RP_ERR_CUDA, cudaGetErrorString(e));
  if ((e = cudaMemcpy(ctx->d_model, host.data(), sizeof(rp::DevModel), cudaMemcpyHostToDevice)) != cudaSuccess)
    return bail(RP_ERR_CUDA, cudaGetErrorString(e));
  if (const char* e = std::getenv("RP_MCC_THREADS")) {
    int t = std::atoi(e) / 32 * 32;
    if (t >= 64 && t <= RP_MCC_THREADS) ctx->threads = t;
  }
  if (const char* e = std::getenv("RP_MCC_LONG_N")) ctx->mcc_long_n = std::atoi(e);
  if (const char* e = std::getenv("RP_MCC_WIDE")) ctx->mcc_wide = std::atoi(e);
  ctx->ctas_per_sm = rp::mcc_max_ctas_per_sm(ctx->threads, 2, ctx->mcc_wide);
  ctx->ctas_per_sm1 = rp::mcc_max_ctas_per_sm(ctx->threads, 1, ctx->mcc_wide);
  ctx->up_ctas_per_sm = rp::unstru_max_ctas_per_sm();
  if (ctx->ctas_per_sm < 1)
    return bail(RP_ERR_CUDA, "rp_create: kernel image not loadable on this device (built for sm_100a)");
  *out = ctx;
  return RP_OK;
}

int rp_destroy(rp_ctx* ctx) {
  if (!ctx) return RP_OK;
  if (ctx->live_batches > 0) {   // batches still point into this context: the last rp_batch_destroy finishes the job
    ctx->closing = true;
    return RP_OK;
  }
  cudaSetDevice(ctx->device);
  if (ctx->own_stream) cudaStreamSynchronize(ctx->own_stream);
  if (ctx->ws) cudaFree(ctx->ws);
  for (auto& f : ctx->pool_free) cudaFree(f.first);
  for (auto& f : ctx->pool_live) cudaFree(f.first);
  if (ctx->d_model) cudaFree(ctx->d_model);
  if (ctx->d_prof) cudaFree(ctx->d_prof);
  for (auto& ev : ctx->ev)
    if (ev) cudaEventDestroy(ev);
  for (auto& ev : ctx->ev_dom)
    if (ev) cudaEventDestroy(ev);
  if (ctx->side_stream) { cudaStreamSynchronize(ctx->side_stream); cudaStreamDestroy(ctx->side_stream); }
  if (ctx->copy_stream) { cudaStreamSynchronize(ctx->copy_stream); cudaStreamDestroy(ctx->copy_stream); }
  if (ctx->ev_main) cudaEventDestroy(ctx->ev_main);
  if (ctx->ev_copy) cudaEventDestroy(ctx->ev_copy);
  if (ctx->ev_fork) cudaEventDestroy(ctx->ev_fork);
  if (ctx->ev_join) cudaEventDestroy(ctx->ev_join);
  if (ctx->own_stream) cudaStreamDestroy(ctx->own_stream);
  delete ctx;
  return RP_OK;
}

int rp_set_stream(rp_ctx* ctx, void* cuda_stream) {
  if (!ctx) return RP_ERR_ARG;
  ctx->stream = cuda_stream ? static_cast<cudaStream_t>(cuda_stream) : ctx->own_stream;
  return RP_OK;
}

void* rp_host_alloc(size_t bytes) {
  void* p = nullptr;
  if (cudaMallocHost(&p, bytes) != cudaSuccess) return nullptr;
  return p;
}
void rp_host_free(void* p) {
  if (p) cudaFreeHost(p);
}

// -------------------------------------------------------------------- batch
int rp_batch_destroy(rp_batch* b) {
  if (!b) return RP_OK;
  if (b->ctx) {
    cudaSetDevice(b->ctx->device);
    if (b->last_stream && b->last_stream != b->ctx->stream) cudaStreamSynchronize(b->last_stream);
    cudaStreamSynchronize(b->ctx->stream);
  }
  if (b->ctx) {
    rp_ctx* ctx = b->ctx;
    for (void* p : {(void*)b->d_seq, (void*)b->d_probs, (void*)b->d_order, (void*)b->d_counter, (void*)b->d_dense,
                    (void*)b->d_logz, (void*)b->d_done, (void*)b->d_spairs, (void*)b->d_recs, (void*)b->d_ups, (void*)b->d_counts})
      pool_release(ctx, p);
    ctx->live_batches--;
    if (ctx->closing && ctx->live_batches == 0) {
      delete b;
      return rp_destroy(ctx);
    }
  }
  delete b;
  return RP_OK;
}

int rp_batch_create(rp_ctx* ctx, const rp_pair* pairs, int n_pairs, const rp_opts* opts, rp_batch** out) {
  if (!ctx || !out) return fail(ctx, RP_ERR_ARG, "rp_batch_create: null argument");
  *out = nullptr;
  int rc = check_pairs(pairs, n_pairs, opts);
  if (rc) return fail(ctx, rc, "rp_batch_create: bad pairs/opts");
  for (int p = 0; p < n_pairs; p++)
    if ((long long)pairs[p].n1 + pairs[p].n2 > rp::RP_MAX_N)
      return fail(ctx, RP_ERR_TOO_LONG, "rp_batch_create: n1+n2 exceeds the 32-bit workspace indexing limit (12000 nt)");
  CU(cudaSetDevice(ctx->device));
  rp_batch* b = new rp_batch;
  b->ctx = ctx;
  ctx->live_batches++;
  b->n_pairs = n_pairs;
  b->opts = *opts;
  b->layout.resize(n_pairs);
  rp_dense_plan(pairs, n_pairs, opts, b->layout.data(), &b->total_floats);
  b->slayout.resize(n_pairs);
  rp_sparse_plan(pairs, n_pairs, opts, b->slayout.data(), &b->total_recs, &b->total_upf);

  // encoded sequences: per pair  0 s1 0 s2 0 s1s2 0
  std::vector<uint8_t> seq;
  size_t bytes = 0;
  for (int p = 0; p < n_pairs; p++) bytes += 2 * ((size_t)pairs[p].n1 + pairs[p].n2) + 4;
  seq.reserve(bytes + 16);
  const int w = opts->max_w > 0 ? opts->max_w : 0;
  size_t ws_need = 0;
  for (int p = 0; p < n_pairs; p++) {
    const rp_pair& pr = pairs[p];
    const rp_dense_layout& L = b->layout[p];
    seq.push_back(0);
    const int off1 = (int)seq.size();
    for (int i = 0; i < pr.n1; i++) seq.push_back(rp::encode_base(pr.s1[i]));
    seq.push_back(0);
    const int off2 = (int)seq.size();
    for (int i = 0; i < pr.n2; i++) seq.push_back(rp::encode_base(pr.s2[i]));
    seq.push_back(0);
    const int off12 = (int)seq.size();
    for (int i = 0; i < pr.n1; i++) seq.push_back(rp::encode_base(pr.s1[i]));
    for (int i = 0; i < pr.n2; i++) seq.push_back(rp::encode_base(pr.s2[i]));
    seq.push_back(0);
    Problem q;
    std::memset(&q, 0, sizeof q);
    q.pair = p; q.max_w = w; q.n1 = pr.n1; q.n2 = pr.n2; q.th_hy = opts->th_hy;
    q.out_bp = q.out_up = q.out_hp = -1;
    q.ws_off = -1;
    // rnafold(fa1, ...), rnafold(fa2, ...)  src/ractip.cpp:546-547
    Problem a = q;
    a.kind = rp::KIND_LINEAR; a.which = 0; a.seq_off = off1; a.n = pr.n1; a.cp = 0;
    a.out_bp = (long long)L.bp1; a.out_up = w > 0 ? (long long)L.up1 : -1;
    b->probs.push_back(a);
    a.which = 1; a.seq_off = off2; a.n = pr.n2;
    a.out_bp = (long long)L.bp2; a.out_up = w > 0 ? (long long)L.up2 : -1;
    b->probs.push_back(a);
    if (pr.n2 == 0) {   // a lone sequence (RactIP::rnafold on its own, src/ractip.cpp:1599-1601): no second fold, no hybridization
      Problem none = q;
      none.kind = rp::KIND_LINEAR; none.which = 2; none.n = 0;
      b->probs.push_back(none);
      b->maxn = std::max(b->maxn, pr.n1);
      b->alg_flops += rp_alg_flops_mcc(pr.n1);
      b->has_single = true;
      continue;
    }
    // rnaduplex(fa1, fa2, hp_)  src/ractip.cpp:548
    Problem c = q;
    c.which = 2; c.seq_off = off12; c.out_hp = (long long)L.hp;
    if (opts->use_pf_duplex) {
      c.kind = rp::KIND_DUPLEX; c.n = pr.n1 + pr.n2; c.cp = pr.n1 + 1;
      b->n_duplex++;
      ws_need = std::max(ws_need, 2 * (size_t)(pr.n1 + 2) * (size_t)(pr.n2 + 2));
    } else {
      c.kind = rp::KIND_COFOLD; c.n = pr.n1 + pr.n2; c.cp = pr.n1 + 1;
      b->maxn = std::max(b->maxn, c.n);
      b->alg_flops += rp_alg_flops_mcc(c.n);
    }
    b->probs.push_back(c);
    b->maxn = std::max(b->maxn, std::max(pr.n1, pr.n2));
    b->alg_flops += rp_alg_flops_mcc(pr.n1) + rp_alg_flops_mcc(pr.n2);
  }
  // LPT cost.  Band-kernel lengths (measured, profiles/r01b_phase_ablation.txt): ~30 k cycles per unit of
  // length for the two wavefronts (fixed cost per diagonal) + ~160 n^2 for the unpaired-window pass.
  auto cost = [&](const Problem& q) {
    if (q.kind == rp::KIND_DUPLEX) return 0.0;
    if (ctx->band && rp_kernel_plan(q.n, ctx->smem_optin, nullptr) < RP_KERNEL_GENERAL)
      return 3.0e4 * q.n + ((q.kind == rp::KIND_LINEAR && q.max_w > 0) ? 160.0 * q.n * q.n : 0.0);
    return (double)q.n * q.n * (q.n + 1500.0);
  };
  // general queue: most expensive first (LPT), ties by index for determinism
  for (size_t k = 0; k < b->probs.size(); k++)
    if (b->probs[k].n > 0) b->order.push_back((int)k);
  std::stable_sort(b->order.begin(), b->order.end(), [&](int x, int y) { return cost(b->probs[x]) > cost(b->probs[y]); });
  if (ctx->band) {
    // class of a problem: 0 = L (512 threads, ring needs more than half an SM), 1 = S (256 threads, 2 CTAs/SM), -1 = general
    auto cls = [&](const Problem& q) {
      if (q.kind == rp::KIND_DUPLEX) return -1;
      const int k = rp_kernel_plan(q.n, ctx->smem_optin, nullptr);
      return k >= RP_KERNEL_GENERAL ? -1 : k;   // RP_KERNEL_GENERAL / RP_KERNEL_GENERAL_WIDE: the general queue
    };
    std::vector<int> part[3];
    for (int k : b->order) {
      const int c = cls(b->probs[k]);
      part[c < 0 ? 2 : c].push_back(k);
      if (c >= 0) b->band_maxn[c] = std::max(b->band_maxn[c], b->probs[k].n);
    }
    b->n_band[0] = (int)part[0].size();
    b->n_band[1] = (int)part[1].size();
    b->order.clear();
    for (auto& v : part) b->order.insert(b->order.end(), v.begin(), v.end());
  }
  b->n_general = (int)b->order.size() - b->n_band[0] - b->n_band[1];
  // Unpaired-window passes of the band classes, three ways:
  //   mode 1 (large batches, >= 4.5 single-strand problems per SM): a launch of their own after the wavefront kernels
  //           (unstru_kernel: one 768-thread CTA per SM -- the pass needs no shared-memory ring, and the tables of 148
  //           resident problems stay in the L2);
  //   mode 2 (batches with at least three problems per SM): JOBS of their own in the band kernel's queue, after all
  //           wavefront jobs, each waiting for its problem's completion flag -- finer jobs shorten the tail;
  //   mode 0 (a handful of problems, or no room for private workspaces): fused with the wavefronts.
  // Modes 1 and 2 keep a deferred problem's tables in a private workspace.  RP_DEFER_UP=0/1/2 forces a mode.
  {
    const char* e = std::getenv("RP_DEFER_UP");
    int maxn = 0;
    std::vector<int> up;
    for (int x = 0; ctx->band && x < b->n_band[0] + b->n_band[1]; x++) {
      const Problem& q = b->probs[b->order[x]];
      if (q.kind == rp::KIND_LINEAR && q.max_w > 0 && q.out_up >= 0) { up.push_back(b->order[x]); maxn = std::max(maxn, q.n); }
    }
    const size_t slot = up.empty() ? 0 : rp::slot_doubles(maxn);
    int mode = 0;
    if (!up.empty() && (double)slot * sizeof(double) * up.size() <= 24e9) {
      // (measured, MicA x ompA shuffles, kernel ms fused / own launch / jobs of their own: 125 pairs 5.09 / 5.13 / 5.29,
      //  180 pairs 7.70 / 8.05 / 7.15, 250 pairs 10.40 / 9.89 / 9.83, 350 pairs 13.94 / 13.34 / 13.45, 500 pairs - / 18.3 / 19.1,
      //  1000 pairs 39.0 / 36.2 / -)
      if (2 * (int)up.size() >= 9 * ctx->sm_count && ctx->up_ctas_per_sm > 0) mode = 1;
      else if (b->n_band[0] + b->n_band[1] >= 3 * ctx->sm_count) mode = 2;
      if (e) mode = std::atoi(e);
    }
    b->up_mode = mode;
    if (mode) {
      b->n_defer = (int)up.size();
      b->defer_slot = slot;
      b->defer_maxn = maxn;
      for (size_t k = 0; k < up.size(); k++) {
        b->probs[up[k]].defer_up = 1;
        b->probs[up[k]].ws_off = (long long)(k * slot);
      }
      // the deferred passes, costliest (longest) first
      std::stable_sort(up.begin(), up.end(), [&](int x, int y) { return b->probs[x].n > b->probs[y].n; });
      if (mode == 1) {
        b->order.insert(b->order.end(), up.begin(), up.end());   // queue of unstru_kernel, after the wavefront queues
      } else {
        // per band class: wavefront jobs with a dependent job first (their successor should start early), then the
        // others, then the unpaired-window jobs
        std::vector<int> q2;
        int off = 0;
        for (int cl = 0; cl < 2; cl++) {
          std::vector<int> dep, oth, ups;
          for (int x = off; x < off + b->n_band[cl]; x++) (b->probs[b->order[x]].defer_up ? dep : oth).push_back(b->order[x]);
          const int maxn_cl = b->band_maxn[cl];
          for (int k : up)
            if (rp_kernel_plan(b->probs[k].n, ctx->smem_optin, nullptr) == cl) ups.push_back(k | rp::RP_JOB_UNPAIRED);
          (void)maxn_cl;
          off += b->n_band[cl];
          q2.insert(q2.end(), dep.begin(), dep.end());
          q2.insert(q2.end(), oth.begin(), oth.end());
          q2.insert(q2.end(), ups.begin(), ups.end());
          b->n_band[cl] = (int)(dep.size() + oth.size() + ups.size());
        }
        q2.insert(q2.end(), b->order.begin() + off, b->order.end());   // the general queue
        b->order.swap(q2);
      }
    }
  }
  int gen_maxn = 0, gen_mcc = 0;
  for (size_t x = (size_t)b->n_band[0] + b->n_band[1]; x < (size_t)(b->n_band[0] + b->n_band[1] + b->n_general); x++) {
    const int k = b->order[x];
    if (b->probs[k].kind != rp::KIND_DUPLEX) { gen_maxn = std::max(gen_maxn, b->probs[k].n); gen_mcc++; }
  }
  b->n_mcc = gen_mcc;
  b->slot_doubles = std::max(gen_mcc ? rp::slot_doubles(gen_maxn) : (size_t)0, ws_need);
  b->mcc_minb = (gen_maxn >= ctx->mcc_long_n && ctx->ctas_per_sm1 > 0) ? 1 : 2;
  // Split fetch plan.  The short band class runs last (side stream, on the SMs the long class leaves), so
  // everything the long class wrote can cross PCIe meanwhile.  Needs a uniform batch (the z-score shuffle batch:
  // every pair has the lengths of the original pair) so that the sections are strided 2-D copies.
  b->split_fetch = false;
  if (ctx->band && n_pairs >= 2 && b->n_band[0] > 0 && b->n_band[1] > 0 && b->n_general == 0 &&
      !std::getenv("RP_NO_SPLIT_FETCH")) {
    bool uniform = true;
    for (int p = 1; p < n_pairs && uniform; p++) uniform = pairs[p].n1 == pairs[0].n1 && pairs[p].n2 == pairs[0].n2;
    if (uniform) {
      const rp_dense_layout& L = b->layout[0];
      b->pair_stride = b->layout[1].bp1 - L.bp1;
      auto is_short = [&](int n) { return rp_kernel_plan(n, ctx->smem_optin, nullptr) == RP_KERNEL_BAND_2CTA; };
      const bool s1 = is_short(pairs[0].n1), s2 = is_short(pairs[0].n2), s12 = is_short(pairs[0].n1 + pairs[0].n2);
      struct Sec { size_t off, len; bool sh; };
      const bool late = b->up_mode == 1;   // the deferred unpaired-window launch writes the up sections last
      const Sec secs[5] = {{L.bp1, L.n_bp1, s1}, {L.bp2, L.n_bp2, s2}, {L.up1, L.n_up1, s1 || late}, {L.up2, L.n_up2, s2 || late}, {L.hp, L.n_hp, s12}};
      b->sect_long.clear(); b->sect_short.clear();
      for (const Sec& x : secs) {   // sections are in layout order: merge neighbours of the same class
        if (!x.len) continue;
        auto& v = x.sh ? b->sect_short : b->sect_long;
        if (!v.empty() && v.back().first + v.back().second == x.off) v.back().second += x.len;
        else v.push_back({x.off - L.bp1, x.len});
      }
      b->split_fetch = !b->sect_long.empty() && !b->sect_short.empty();
    }
  }

  auto bail = [&](cudaError_t e, const char* what) {
    rp_batch_destroy(b);
    return fail(ctx, RP_ERR_CUDA, std::string(what) + ": " + cudaGetErrorString(e));
  };
  cudaError_t e;
  const size_t np = b->probs.size();
  if ((e = pool_alloc(ctx, &b->d_seq, seq.size() + 16)) != cudaSuccess) return bail(e, "cudaMalloc seq");
  if ((e = pool_alloc(ctx, &b->d_probs, std::max<size_t>(1, np) * sizeof(Problem))) != cudaSuccess) return bail(e, "cudaMalloc probs");
  if ((e = pool_alloc(ctx, &b->d_order, std::max<size_t>(1, b->order.size()) * sizeof(int))) != cudaSuccess) return bail(e, "cudaMalloc order");
  if ((e = pool_alloc(ctx, &b->d_done, std::max<size_t>(1, np) * sizeof(int))) != cudaSuccess) return bail(e, "cudaMalloc done");
  // The private workspaces of the deferred passes are part of the context's workspace (grown on demand by rp_batch_run,
  // kept between calls).  If the device cannot hold them, the pass stays fused with the wavefronts.
  bool up_fits = true;
  if (b->n_defer) {
    const size_t need = b->defer_slot * sizeof(double) * (size_t)b->n_defer;
    size_t free_b = 0, total_b = 0;
    if (need > ctx->ws_bytes && cudaMemGetInfo(&free_b, &total_b) == cudaSuccess)
      up_fits = need <= (free_b + ctx->ws_bytes + pool_cached_bytes(ctx)) / 10 * 6;
  }
  if (b->n_defer && !up_fits) {
    for (auto& q : b->probs) { q.defer_up = 0; q.ws_off = -1; }
    if (b->up_mode == 1) b->order.resize(b->order.size() - (size_t)b->n_defer);
    else {   // drop the unpaired-window jobs from the band queues
      std::vector<int> q2;
      int off = 0;
      for (int cl = 0; cl < 2; cl++) {
        int kept = 0;
        for (int x = off; x < off + b->n_band[cl]; x++)
          if (!(b->order[x] & rp::RP_JOB_UNPAIRED)) { q2.push_back(b->order[x]); kept++; }
        off += b->n_band[cl];
        b->n_band[cl] = kept;
      }
      q2.insert(q2.end(), b->order.begin() + off, b->order.end());
      b->order.swap(q2);
    }
    b->n_defer = 0;
    b->up_mode = 0;
  }
  if ((e = pool_alloc(ctx, &b->d_counter, 4 * sizeof(int))) != cudaSuccess) return bail(e, "cudaMalloc counter");
  if ((e = pool_alloc(ctx, &b->d_dense, std::max<size_t>(1, b->total_floats) * sizeof(float))) != cudaSuccess) return bail(e, "cudaMalloc dense");
  if ((e = pool_alloc(ctx, &b->d_logz, std::max<size_t>(1, (size_t)n_pairs * 3) * sizeof(double))) != cudaSuccess) return bail(e, "cudaMalloc logz");
  cudaStream_t st = ctx->stream;
  if ((e = cudaMemcpyAsync(b->d_seq, seq.data(), seq.size(), cudaMemcpyHostToDevice, st)) != cudaSuccess) return bail(e, "H2D seq");
  if (np) {
    if ((e = cudaMemcpyAsync(b->d_probs, b->probs.data(), np * sizeof(Problem), cudaMemcpyHostToDevice, st)) != cudaSuccess) return bail(e, "H2D probs");
    if (!b->order.empty() && (e = cudaMemcpyAsync(b->d_order, b->order.data(), b->order.size() * sizeof(int), cudaMemcpyHostToDevice, st)) != cudaSuccess) return bail(e, "H2D order");
  }
  if (b->has_single && (e = cudaMemsetAsync(b->d_dense, 0, std::max<size_t>(1, b->total_floats) * sizeof(float), st)) != cudaSuccess) return bail(e, "memset dense");
  if ((e = cudaMemsetAsync(b->d_logz, 0, std::max<size_t>(1, (size_t)n_pairs * 3) * sizeof(double), st)) != cudaSuccess) return bail(e, "memset logz");
  if ((e = cudaStreamSynchronize(st)) != cudaSuccess) return bail(e, "sync");  // host vectors go out of scope
  *out = b;
  return RP_OK;
}

int rp_batch_run(rp_batch* b) {
  if (!b) return RP_ERR_ARG;
  rp_ctx* ctx = b->ctx;
  CU(cudaSetDevice(ctx->device));
  const int nprob = (int)b->probs.size();
  ctx->timing = rp_timing{};
  ctx->timing.alg_flops = b->alg_flops;
  if (nprob == 0) { ctx->timing_pending = false; return RP_OK; }
  const size_t slot_bytes = b->slot_doubles * sizeof(double);
  if (b->grid_cached < 0) b->grid_cached = b->n_general > 0 ? grid_for(ctx, b->n_general, std::max<size_t>(slot_bytes, 8), b->mcc_minb) : 0;
  const int grid = b->grid_cached;
  // the two kernels run one after the other on the stream and share the workspace
  // band classes: grid = resident CTAs, one workspace slot each, placed after the general kernel's slots
  const int band_threads[2] = {512, 256};
  size_t band_smem[2] = {0, 0}, band_slot[2] = {0, 0};
  for (int k = 0; k < 2; k++) {
    if (!b->n_band[k]) { b->band_grid[k] = 0; continue; }
    band_smem[k] = rp::band_shared_bytes(b->band_maxn[k], band_threads[k]);
    band_slot[k] = rp::slot_doubles(b->band_maxn[k]);
    if (b->band_grid[k] < 0) {
      const int occ = rp::band_max_ctas_per_sm(band_threads[k], band_smem[k]);
      if (occ < 1) return fail(ctx, RP_ERR_CUDA, "rp_batch_run: band kernel does not fit on this device");
      b->band_grid[k] = std::min(b->n_band[k], ctx->sm_count * occ);
      if (const char* e = std::getenv("RP_GRID")) {
        int v = std::atoi(e);
        if (v >= 1 && v < b->band_grid[k]) b->band_grid[k] = v;
      }
    }
  }
  const size_t gen_bytes = (size_t)grid * slot_bytes;
  const size_t bandL_bytes = (size_t)b->band_grid[0] * band_slot[0] * sizeof(double);
  const size_t bandS_bytes = (size_t)b->band_grid[1] * band_slot[1] * sizeof(double);
  const size_t up_bytes = b->n_defer ? b->defer_slot * sizeof(double) * (size_t)b->n_defer : 0;
  int rc = ensure_workspace(ctx, gen_bytes + bandL_bytes + bandS_bytes + up_bytes);
  if (rc) return rc;
  const int n_bandall = b->n_band[0] + b->n_band[1];
  rp::BatchDev d;
  d.model = ctx->d_model; d.seq = b->d_seq; d.probs = b->d_probs; d.order = b->d_order + n_bandall; d.nprob = b->n_general;
  d.counter = b->d_counter; d.ws = ctx->ws; d.slot_stride = b->slot_doubles; d.nslots = std::max(grid, 1);
  d.dense = b->d_dense; d.logz = b->d_logz;
  d.ws_up = ctx->ws + (gen_bytes + bandL_bytes + bandS_bytes) / sizeof(double); d.done = b->d_done;
  d.prof = nullptr;
  d.dbg = std::getenv("RP_DEBUG_SKIP") ? std::atoi(std::getenv("RP_DEBUG_SKIP")) : 0;
  long long*& d_prof = ctx->d_prof;
  const bool profile = std::getenv("RP_PROFILE") != nullptr;
  if (profile) {
    if (!d_prof) CU(cudaMalloc(&d_prof, 64 * sizeof(long long)));
    CU(cudaMemsetAsync(d_prof, 0, 64 * sizeof(long long), ctx->stream));
    d.prof = d_prof;
  }
  if (b->dom_kind < 0) {   // which launch carries most of the algorithmic flops (timed by itself for the roofline)
    double f[3] = {0, 0, 0};
    int off = 0;
    for (int k = 0; k < 3; k++) {
      const int cnt = k < 2 ? b->n_band[k] : b->n_general;
      for (int x = off; x < off + cnt; x++) {
        if (b->order[x] & rp::RP_JOB_UNPAIRED) continue;   // (an unpaired-window job carries no credited flops)
        const Problem& q = b->probs[b->order[x]];
        if (q.kind != rp::KIND_DUPLEX) f[k] += rp_alg_flops_mcc(q.n);
      }
      off += cnt;
    }
    b->dom_kind = f[0] >= f[1] && f[0] >= f[2] ? 0 : (f[1] >= f[2] ? 1 : 2);
    b->dom_flops = f[b->dom_kind];
    if (b->dom_flops <= 0) b->dom_kind = 3;   // nothing to time
  }
  ctx->dom_timed = false;
  ctx->timing.dominant_kind = b->dom_kind < 3 ? b->dom_kind : -1;
  ctx->timing.alg_flops_dominant = b->dom_flops;
  ctx->timing.ms_dominant = 0;
  cudaStream_t st = ctx->stream;
  b->last_stream = st;
  ctx->ws_stream = st;
  CU(cudaMemsetAsync(b->d_counter, 0, 4 * sizeof(int), st));
  if (b->n_defer) CU(cudaMemsetAsync(b->d_done, 0, b->probs.size() * sizeof(int), st));
  CU(cudaEventRecord(ctx->ev[0], st));
  int launches = 0;
  {
    size_t ws_off = gen_bytes / sizeof(double);
    int ord_off = 0;
    if (b->n_band[0] && b->n_band[1]) CU(cudaEventRecord(ctx->ev_fork, st));   // inputs and counters are ready here
    for (int k = 0; k < 2; k++) {
      if (b->n_band[k]) {
        rp::BatchDev db = d;
        db.order = b->d_order + ord_off; db.nprob = b->n_band[k];
        db.counter = b->d_counter + 1 + k;
        db.ws = ctx->ws + ws_off; db.slot_stride = band_slot[k]; db.nslots = b->band_grid[k];
        cudaStream_t ls = st;
        if (k == 1 && b->n_band[0]) {   // the short class runs on the side stream (forked before the first launch)
          CU(cudaStreamWaitEvent(ctx->side_stream, ctx->ev_fork, 0));
          ls = ctx->side_stream;
        }
        if (b->dom_kind == k) CU(cudaEventRecord(ctx->ev_dom[0], ls));
        CU(rp::launch_band(db, b->band_grid[k], band_threads[k], band_smem[k], ls));
        if (b->dom_kind == k) { CU(cudaEventRecord(ctx->ev_dom[1], ls)); ctx->dom_timed = true; }
        if (k == 0 && b->split_fetch) CU(cudaEventRecord(ctx->ev_main, st));   // the long class's outputs are final here
        if (ls != st) {                 // join before anything else of the batch runs
          CU(cudaEventRecord(ctx->ev_join, ls));
          CU(cudaStreamWaitEvent(st, ctx->ev_join, 0));
        }
        launches++;
      }
      ws_off += (size_t)b->band_grid[k] * band_slot[k];
      ord_off += b->n_band[k];
    }
  }
  if (b->up_mode == 1) {   // the unpaired-window passes of the band classes, after both of their launches (the streams have joined)
    rp::BatchDev du = d;
    du.order = b->d_order + n_bandall + b->n_general; du.nprob = b->n_defer;
    du.counter = b->d_counter + 3;
    if (b->up_grid < 0) b->up_grid = std::min(b->n_defer, ctx->sm_count * std::max(1, ctx->up_ctas_per_sm));
    CU(rp::launch_unstru(du, b->up_grid, st));
    launches++;
  }
  if (b->n_mcc > 0) {
    // Few long problems (a single long pair, not a shuffle batch): one problem per thread-block cluster, so
    // that its wavefront runs on several SMs -- 16 CTAs per problem while that leaves no cluster waiting,
    // 8 up to ~5 rounds of clusters.  Beyond that one CTA per problem has the higher throughput (a 1500-nt
    // problem: 57 ms on 16 SMs, 76 ms on 8, ~400 ms on one).
    // RP_CLUSTER=0 switches the multi-CTA path off, 8 / 16 force it with that cluster size.
    const char* ce = std::getenv("RP_CLUSTER");
    int G = 0;
    if (ce) G = std::atoi(ce) >= 16 ? 16 : std::atoi(ce) > 0 ? 8 : 0;
    else if (b->n_mcc * 16 <= ctx->sm_count) G = 16;                      // any general-kernel length (measured: one 250 x 100 pair 17.2 -> 9.3 ms)
    else if (b->n_mcc * 8 <= ctx->sm_count * (b->mcc_minb == 1 ? 5 : 4)) G = 8;   // up to 4-5 rounds of clusters (20 pairs of 400 x 300: 87 -> 56 ms; 40 pairs: 93 vs 106 ms)
    // a cluster shape the device (or its current partitioning) cannot schedule fails at launch, before
    // anything ran: fall back to the smaller cluster, then to one CTA per problem
    if (b->dom_kind == 2) CU(cudaEventRecord(ctx->ev_dom[0], st));
    bool launched = false;
    for (; G && !launched; G = G == 16 ? 8 : 0) {
      const int ncl = std::max(1, std::min({b->n_mcc, grid, ctx->sm_count / G}));
      if (rp::launch_mcc_cluster(d, ncl, G, ctx->threads, st) == cudaSuccess) launched = true;
      else cudaGetLastError();
    }
    if (!launched) CU(rp::launch_mcc(d, grid, ctx->threads, b->mcc_minb, ctx->mcc_wide, st));
    if (b->dom_kind == 2) { CU(cudaEventRecord(ctx->ev_dom[1], st)); ctx->dom_timed = true; }
    launches++;
  }
  if (b->n_duplex > 0) {
    CU(rp::launch_duplex(d, std::max(1, std::min(grid, b->n_general)), st));
    launches++;
  }
  CU(cudaEventRecord(ctx->ev[1], st));
  if (profile) {
    long long h[64];
    CU(cudaMemcpyAsync(h, d_prof, sizeof h, cudaMemcpyDeviceToHost, st));
    CU(cudaStreamSynchronize(st));
    static const char* names[32] = {"stage", "prologue", "prologue2", "inside_A", "inside_B", "generic_in", "generic_out", "outside_A", "outside_B", "write_bp", "un_hairpin", "un_gaps0", "un_gaps1", "un_domrows", "un_domcols", "un_mltab", "un_windows", "write_hp", "logz", "band_A", "band_B", "collect", "un_hairpin.cells(t0)", "un_hairpin.specials(t0)", "finish_in.loads(t0)", "finish_in.rest(t0)", "bandA_in.head(t0)", "bandA_in.main(t0)", "bandA_in.tail(t0)", "bandA_out.PR(t0)", "bandA_out.need(t0)", "bandA_out.ML(t0)"};
    long long tot = 0;
    for (int k = 0; k < 22; k++) tot += h[k];
    for (int k = 0; k < 32; k++)
      if (h[32 + k]) std::fprintf(stderr, "[rp_profile] %-11s %6.2f%%  calls %9lld  cycles/call %9.0f\n", names[k], 100.0 * h[k] / tot, h[32 + k], (double)h[k] / h[32 + k]);
  }
  ctx->timing.kernel_launches = launches;
  ctx->timing_pending = true;
  ctx->timed_copies = false;
  return RP_OK;
}

int rp_batch_sync(rp_batch* b) {
  if (!b) return RP_ERR_ARG;
  rp_ctx* ctx = b->ctx;
  CU(cudaStreamSynchronize(ctx->stream));
  return RP_OK;
}

int rp_batch_fetch_dense(rp_batch* b, float* out, size_t out_floats) {
  if (!b || !out) return RP_ERR_ARG;
  rp_ctx* ctx = b->ctx;
  if (out_floats < b->total_floats) return fail(ctx, RP_ERR_CAPACITY, "rp_batch_fetch_dense: buffer too small");
  CU(cudaSetDevice(ctx->device));
  CU(cudaEventRecord(ctx->ev[2], ctx->stream));
  if (b->total_floats && b->split_fetch) {
    // long-class sections: on the copy stream, as soon as the long class is done (ev_main), while the short class
    // still computes; short-class sections: on the main stream, which has joined the side stream
    const size_t pitch = b->pair_stride * sizeof(float);
    CU(cudaStreamWaitEvent(ctx->copy_stream, ctx->ev_main, 0));
    for (const auto& sc : b->sect_long)
      CU(cudaMemcpy2DAsync(out + sc.first, pitch, b->d_dense + sc.first, pitch, sc.second * sizeof(float), (size_t)b->n_pairs,
                           cudaMemcpyDeviceToHost, ctx->copy_stream));
    CU(cudaEventRecord(ctx->ev_copy, ctx->copy_stream));
    for (const auto& sc : b->sect_short)
      CU(cudaMemcpy2DAsync(out + sc.first, pitch, b->d_dense + sc.first, pitch, sc.second * sizeof(float), (size_t)b->n_pairs,
                           cudaMemcpyDeviceToHost, ctx->stream));
    CU(cudaStreamWaitEvent(ctx->stream, ctx->ev_copy, 0));
  } else if (b->total_floats) {
    CU(cudaMemcpyAsync(out, b->d_dense, b->total_floats * sizeof(float), cudaMemcpyDeviceToHost, ctx->stream));
  }
  CU(cudaEventRecord(ctx->ev[3], ctx->stream));
  CU(cudaStreamSynchronize(ctx->stream));
  ctx->timed_copies = true;
  return RP_OK;
}

}  // extern "C" (reopened below)

namespace {
int sparse_prepare(rp_batch* b) {
  rp_ctx* ctx = b->ctx;
  const int np = b->n_pairs;
  if (b->d_spairs || np == 0) return RP_OK;
  std::vector<rp::SparsePair> sp(np);
  for (int p = 0; p < np; p++) {
    const rp_dense_layout& L = b->layout[p];
    const rp_sparse_layout& S = b->slayout[p];
    rp::SparsePair& q = sp[p];
    q.n1 = b->probs[3 * p].n; q.n2 = b->probs[3 * p + 1].n;
    q.bp1 = (long long)L.bp1; q.bp2 = (long long)L.bp2; q.hp = (long long)L.hp;
    q.x = (long long)S.x; q.y = (long long)S.y; q.z = (long long)S.z;
    q.cap_x = (int)S.cap_x; q.cap_y = (int)S.cap_y; q.cap_z = (int)S.cap_z;
    q.up1_src = (long long)L.up1; q.up2_src = (long long)L.up2;
    q.up1_dst = (long long)S.up1; q.up2_dst = (long long)S.up2;
    q.n_up1 = (int)S.n_up1; q.n_up2 = (int)S.n_up2;
    q.v = (long long)S.v; q.w = (long long)S.w;
    q.cap_v = (int)S.cap_v; q.cap_w = (int)S.cap_w;
  }
  CU(pool_alloc(ctx, &b->d_spairs, np * sizeof(rp::SparsePair)));
  CU(cudaMemcpy(b->d_spairs, sp.data(), np * sizeof(rp::SparsePair), cudaMemcpyHostToDevice));
  return RP_OK;
}

int sparse_launch(rp_batch* b, rp_rec* d_recs, float* d_ups, rp_sparse_counts* d_counts) {
  rp_ctx* ctx = b->ctx;
  rp::SparseDev s;
  s.pairs = b->d_spairs; s.dense = b->d_dense; s.recs = d_recs; s.ups = d_ups; s.counts = d_counts;
  s.th_ss = b->opts.th_ss; s.th_hy = b->opts.th_hy; s.th_ac = b->opts.th_ac;
  s.min_w = b->opts.min_w; s.max_w = b->opts.max_w;
  CU(cudaMemsetAsync(d_counts, 0, b->n_pairs * sizeof(rp_sparse_counts), ctx->stream));
  CU(rp::launch_sparse(s, b->n_pairs, d_ups != nullptr, ctx->stream));
  ctx->timing.kernel_launches += d_ups ? 2 : 1;
  return RP_OK;
}
}  // namespace

extern "C" int rp_batch_sparse_device(rp_batch* b, void* recs_dev, size_t n_recs, void* ups_dev, size_t n_floats,
                                      void* counts_dev) {
  if (!b || !recs_dev || !counts_dev) return RP_ERR_ARG;
  rp_ctx* ctx = b->ctx;
  if (n_recs < b->total_recs || (ups_dev && n_floats < b->total_upf))
    return fail(ctx, RP_ERR_CAPACITY, "rp_batch_sparse_device: buffer too small");
  CU(cudaSetDevice(ctx->device));
  if (b->n_pairs == 0) return RP_OK;
  int rc = sparse_prepare(b);
  if (rc) return rc;
  return sparse_launch(b, static_cast<rp_rec*>(recs_dev), static_cast<float*>(ups_dev),
                       static_cast<rp_sparse_counts*>(counts_dev));
}

extern "C" int rp_batch_fetch_sparse(rp_batch* b, rp_rec* recs, size_t n_recs, float* ups, size_t n_floats,
                                     rp_sparse_counts* counts) {
  if (!b || !recs || !counts) return RP_ERR_ARG;
  rp_ctx* ctx = b->ctx;
  if (n_recs < b->total_recs || (ups && n_floats < b->total_upf))
    return fail(ctx, RP_ERR_CAPACITY, "rp_batch_fetch_sparse: buffer too small");
  CU(cudaSetDevice(ctx->device));
  cudaStream_t st = ctx->stream;
  const int np = b->n_pairs;
  if (np == 0) return RP_OK;
  int rc = sparse_prepare(b);
  if (rc) return rc;
  if (!b->d_recs) {
    CU(pool_alloc(ctx, &b->d_recs, std::max<size_t>(1, b->total_recs) * sizeof(rp_rec)));
    CU(pool_alloc(ctx, &b->d_ups, std::max<size_t>(1, b->total_upf) * sizeof(float)));
    CU(pool_alloc(ctx, &b->d_counts, np * sizeof(rp_sparse_counts)));
  }
  rc = sparse_launch(b, b->d_recs, ups ? b->d_ups : nullptr, b->d_counts);
  if (rc) return rc;
  CU(cudaEventRecord(ctx->ev[2], st));
  if (b->total_recs) CU(cudaMemcpyAsync(recs, b->d_recs, b->total_recs * sizeof(rp_rec), cudaMemcpyDeviceToHost, st));
  if (ups && b->total_upf) CU(cudaMemcpyAsync(ups, b->d_ups, b->total_upf * sizeof(float), cudaMemcpyDeviceToHost, st));
  CU(cudaMemcpyAsync(counts, b->d_counts, np * sizeof(rp_sparse_counts), cudaMemcpyDeviceToHost, st));
  CU(cudaEventRecord(ctx->ev[3], st));
  CU(cudaStreamSynchronize(st));
  ctx->timed_copies = true;
  for (int p = 0; p < np; p++)
    if (counts[p].overflow) return fail(ctx, RP_ERR_CAPACITY, "rp_batch_fetch_sparse: record capacity exceeded");
  return RP_OK;
}

extern "C" {

int rp_batch_fetch_logz(rp_batch* b, double* logz, size_t n) {
  if (!b || !logz) return RP_ERR_ARG;
  rp_ctx* ctx = b->ctx;
  if (n < (size_t)b->n_pairs * 3) return fail(ctx, RP_ERR_CAPACITY, "rp_batch_fetch_logz: buffer too small");
  CU(cudaSetDevice(ctx->device));
  if (b->n_pairs)
    CU(cudaMemcpyAsync(logz, b->d_logz, (size_t)b->n_pairs * 3 * sizeof(double), cudaMemcpyDeviceToHost, ctx->stream));
  CU(cudaStreamSynchronize(ctx->stream));
  return RP_OK;
}

int rp_last_timing(const rp_ctx* cctx, rp_timing* t) {
  if (!cctx || !t) return RP_ERR_ARG;
  rp_ctx* ctx = const_cast<rp_ctx*>(cctx);
  if (ctx->timing_pending) {
    CU(cudaSetDevice(ctx->device));
    CU(cudaStreamSynchronize(ctx->stream));
    float ms = 0;
    CU(cudaEventElapsedTime(&ms, ctx->ev[0], ctx->ev[1]));
    ctx->timing.ms_total = ms;
    if (ctx->dom_timed) {
      CU(cudaEventElapsedTime(&ms, ctx->ev_dom[0], ctx->ev_dom[1]));
      ctx->timing.ms_dominant = ms;
    }
    if (ctx->timed_copies) {
      CU(cudaEventElapsedTime(&ms, ctx->ev[2], ctx->ev[3]));
      ctx->timing.ms_d2h = ms;
    }
    ctx->timing_pending = false;
  }
  *t = ctx->timing;
  return RP_OK;
}

// ------------------------------------------------------- one-shot host calls
int rp_run_dense(rp_ctx* ctx, const rp_pair* pairs, int n_pairs, const rp_opts* opts, float* out, size_t out_floats) {
  if (!ctx || !out) return fail(ctx, RP_ERR_ARG, "rp_run_dense: null argument");
  rp_batch* b = nullptr;
  int rc = rp_batch_create(ctx, pairs, n_pairs, opts, &b);
  if (rc) return rc;
  if (out_floats < b->total_floats) {
    rp_batch_destroy(b);
    return fail(ctx, RP_ERR_CAPACITY, "rp_run_dense: buffer too small");
  }
  rc = rp_batch_run(b);
  if (!rc) rc = rp_batch_fetch_dense(b, out, out_floats);
  rp_timing t;
  if (!rc) rp_last_timing(ctx, &t);
  rp_batch_destroy(b);
  return rc;
}

int rp_run_sparse(rp_ctx* ctx, const rp_pair* pairs, int n_pairs, const rp_opts* opts, rp_rec* recs, size_t n_recs,
                  float* ups, size_t n_floats, rp_sparse_counts* counts) {
  if (!ctx || !recs || !counts) return fail(ctx, RP_ERR_ARG, "rp_run_sparse: null argument");
  rp_batch* b = nullptr;
  int rc = rp_batch_create(ctx, pairs, n_pairs, opts, &b);
  if (rc) return rc;
  rc = rp_batch_run(b);
  if (!rc) rc = rp_batch_fetch_sparse(b, recs, n_recs, ups, n_floats, counts);
  rp_timing t;
  if (!rc) rp_last_timing(ctx, &t);
  rp_batch_destroy(b);
  return rc;
}

// ------------------------------------------------------------------- peaks
int rp_measure_peaks(rp_ctx* ctx, double* fp64_tflops, double* smem_gbs) {
  if (!ctx) return RP_ERR_ARG;
  CU(cudaSetDevice(ctx->device));
  const int grid = ctx->sm_count * 8, threads = 256;
  double* d_out = nullptr;
  CU(cudaMalloc(&d_out, (size_t)grid * threads * sizeof(double)));
  cudaStream_t st = ctx->stream;
  float best_f = 1e30f, best_s = 1e30f;
  const int it_f = 1 << 15, it_s = 1 << 12;
  for (int rep = 0; rep < 4; rep++) {
    CU(cudaEventRecord(ctx->ev[4], st));
    CU(rp::launch_peak_fp64(d_out, grid, it_f, st));
    CU(cudaEventRecord(ctx->ev[5], st));
    CU(cudaStreamSynchronize(st));
    float ms;
    CU(cudaEventElapsedTime(&ms, ctx->ev[4], ctx->ev[5]));
    if (rep > 0) best_f = std::min(best_f, ms);
    CU(cudaEventRecord(ctx->ev[4], st));
    CU(rp::launch_peak_smem(d_out, grid, it_s, st));
    CU(cudaEventRecord(ctx->ev[5], st));
    CU(cudaStreamSynchronize(st));
    CU(cudaEventElapsedTime(&ms, ctx->ev[4], ctx->ev[5]));
    if (rep > 0) best_s = std::min(best_s, ms);
  }
  CU(cudaFree(d_out));
  if (fp64_tflops) *fp64_tflops = 2.0 * 8 * (double)it_f * grid * threads / (best_f * 1e-3) / 1e12;
  if (smem_gbs) *smem_gbs = 8.0 * 16 * (double)it_s * grid * threads / (best_s * 1e-3) / 1e9;
  return RP_OK;
}

}  // extern "C"
