// C ABI of libpnb200.so (see include/pyneapple_b200.h).
#include <cuda_runtime.h>

#include <atomic>
#include <condition_variable>
#include <cstdio>
#include <cstdlib>
#include <cstring>
#include <mutex>
#include <string>
#include <thread>
#include <vector>

#include "../../include/pyneapple_b200.h"
#include "pnb_internal.h"
#include "pnb_trf_kernel.cuh"

namespace {

thread_local std::string g_err;
std::atomic<long long> g_launches{0};

int fail(int code, const std::string &msg) {
  g_err = msg;
  return code;
}
int cuda_fail(cudaError_t e, const char *what) {
  g_err = std::string(what) + ": " + cudaGetErrorString(e);
  return (int)e;
}
#define PNB_CUDA(call)                                  \
  do {                                                  \
    cudaError_t e_ = (call);                            \
    if (e_ != cudaSuccess) return cuda_fail(e_, #call); \
  } while (0)

using LaunchFn = cudaError_t (*)(const pnb::TrfDeviceArgs *, cudaStream_t);

}  // namespace

namespace pnb {
cudaError_t trf_cov_launch(int n_free, long long n_vox, int m, const int *status, double *cov, cudaStream_t stream,
                           int absolute_sigma) {
#define PNB_COV_CASE(nf)                                                                           \
  case nf: {                                                                                       \
    constexpr int TV = pnb::CovTile<nf>::TV;                                                       \
    cov_kernel<nf><<<(unsigned)((n_vox + TV - 1) / TV), TV, 0, stream>>>(n_vox, m, status, cov, absolute_sigma);   \
  } break;
  switch (n_free) {
    PNB_COV_CASE(2) PNB_COV_CASE(3) PNB_COV_CASE(4) PNB_COV_CASE(5) PNB_COV_CASE(6) PNB_COV_CASE(7)
    default: return cudaErrorInvalidValue;
  }
  g_launches.fetch_add(1);
  return cudaGetLastError();
}
}  // namespace pnb

namespace pnbi {
int fail(int code, const std::string &msg) { return ::fail(code, msg); }
int cuda_fail(cudaError_t e, const char *what) { return ::cuda_fail(e, what); }
void count_launch() { g_launches.fetch_add(1); }

bool is_pageable(const void *ptr) {
  cudaPointerAttributes attr;
  if (cudaPointerGetAttributes(&attr, ptr) != cudaSuccess) { cudaGetLastError(); return true; }
  return attr.type == cudaMemoryTypeUnregistered;
}

namespace {
// Persistent copy workers: a staged host pipeline issues ~100 copies of 2-35 MB per call, and
// creating 15 threads for each of them cost more than the copies of the small ones.
class CopyPool {
 public:
  explicit CopyPool(unsigned n) : n_(n) {
    for (unsigned t = 0; t < n_; t++) workers_.emplace_back([this, t] { run(t); });
  }
  ~CopyPool() {
    {
      std::lock_guard<std::mutex> lk(mu_);
      stop_ = true;
      gen_++;
    }
    cv_.notify_all();
    for (auto &w : workers_) w.join();
  }
  // copies [part * (t + 1), ...) pieces on the workers while the caller copies piece 0
  void copy(char *dst, const char *src, size_t bytes, size_t part, unsigned pieces) {
    std::unique_lock<std::mutex> lk(mu_);
    dst_ = dst; src_ = src; bytes_ = bytes; part_ = part; pieces_ = pieces;
    pending_ = pieces - 1;
    gen_++;
    lk.unlock();
    cv_.notify_all();
    std::memcpy(dst, src, part < bytes ? part : bytes);
    lk.lock();
    done_.wait(lk, [this] { return pending_ == 0; });
  }

 private:
  void run(unsigned t) {
    unsigned long long seen = 0;
    for (;;) {
      std::unique_lock<std::mutex> lk(mu_);
      cv_.wait(lk, [&] { return gen_ != seen; });
      seen = gen_;
      if (stop_) return;
      const unsigned piece = t + 1;
      if (piece >= pieces_) continue;
      char *d = dst_; const char *s = src_;
      const size_t off = part_ * piece, total = bytes_, part = part_;
      lk.unlock();
      if (off < total) std::memcpy(d + off, s + off, off + part > total ? total - off : part);
      lk.lock();
      if (--pending_ == 0) done_.notify_one();
    }
  }
  unsigned n_;
  std::vector<std::thread> workers_;
  std::mutex mu_;
  std::condition_variable cv_, done_;
  unsigned long long gen_ = 0;
  bool stop_ = false;
  char *dst_ = nullptr; const char *src_ = nullptr;
  size_t bytes_ = 0, part_ = 0;
  unsigned pieces_ = 0, pending_ = 0;
};
std::mutex g_copy_mu;  // one parallel copy at a time (the pool has one job slot)
}  // namespace

void parallel_memcpy(void *dst, const void *src, size_t bytes) {
  static const unsigned hw = std::thread::hardware_concurrency();
  // PNB_COPY_THREADS overrides the default of half the cores (4 .. 12) per staging copy: on the 16-core
  // boxes 8 threads already move 35 GB/s and leave cores to the thread that drives the pipeline
  // (end to end from pageable memory: 198 Mvoxel/s with 8, 155 with 16; profiles/r2_pageable_probe.log)
  static const unsigned want = [] { const char *e = std::getenv("PNB_COPY_THREADS"); return e ? (unsigned)std::atoi(e) : 0u; }();
  static const unsigned nt = want ? want : (hw >= 24 ? 12 : (hw >= 8 ? hw / 2 : (hw ? hw : 1)));
  if (bytes < (4u << 20) || nt < 2) { std::memcpy(dst, src, bytes); return; }
  static CopyPool *pool = new CopyPool(nt - 1);  // lives for the process (workers sleep on a condition variable)
  const size_t part = ((bytes / nt) + 4095) & ~(size_t)4095;
  const unsigned pieces = (unsigned)((bytes + part - 1) / part);
  std::lock_guard<std::mutex> lk(g_copy_mu);
  pool->copy((char *)dst, (const char *)src, bytes, part, pieces);
}
}  // namespace pnbi

#define PNB_DECL(id, t1)                                                                            \
  extern "C" cudaError_t pnb_trf_launch_##id##_##t1(const pnb::TrfDeviceArgs *, cudaStream_t);     \
  extern "C" cudaError_t pnb_dogbox_launch_##id##_##t1(const pnb::TrfDeviceArgs *, cudaStream_t);  \
  extern "C" cudaError_t pnb_lm_launch_##id##_##t1(const pnb::TrfDeviceArgs *, cudaStream_t);
PNB_DECL(0, 0) PNB_DECL(1, 0) PNB_DECL(2, 0) PNB_DECL(3, 0) PNB_DECL(4, 0) PNB_DECL(5, 0) PNB_DECL(6, 0)
// the EXTRAS instantiations (curve_fit sigma / robust loss): method trf, no T1 parameter
#define PNB_DECLX(id) extern "C" cudaError_t pnb_trfx_launch_##id##_0(const pnb::TrfDeviceArgs *, cudaStream_t);
PNB_DECLX(0) PNB_DECLX(1) PNB_DECLX(2) PNB_DECLX(3) PNB_DECLX(4) PNB_DECLX(5) PNB_DECLX(6)
#ifdef PNB_WITH_T1
PNB_DECL(0, 1) PNB_DECL(1, 1) PNB_DECL(2, 1) PNB_DECL(3, 1) PNB_DECL(4, 1) PNB_DECL(5, 1) PNB_DECL(6, 1)
PNB_DECL(0, 2) PNB_DECL(1, 2) PNB_DECL(2, 2) PNB_DECL(3, 2) PNB_DECL(4, 2) PNB_DECL(5, 2) PNB_DECL(6, 2)
#endif

namespace {

#define PNB_ROW(name, t1) {name##_0_##t1, name##_1_##t1, name##_2_##t1, name##_3_##t1, name##_4_##t1, name##_5_##t1, name##_6_##t1}
#define PNB_NOROW {nullptr, nullptr, nullptr, nullptr, nullptr, nullptr, nullptr}
LaunchFn trf_launcher(int model_id, int t1_mode, int method = 0, bool extras = false) {
  if (extras) {
    static const LaunchFn xtable[7] = PNB_ROW(pnb_trfx_launch, 0);
    if (model_id < 0 || model_id > 6 || t1_mode != 0 || method != 0) return nullptr;
    return xtable[model_id];
  }
  static const LaunchFn table[3][3][7] = {
#ifdef PNB_WITH_T1
      {PNB_ROW(pnb_trf_launch, 0), PNB_ROW(pnb_trf_launch, 1), PNB_ROW(pnb_trf_launch, 2)},
      {PNB_ROW(pnb_dogbox_launch, 0), PNB_ROW(pnb_dogbox_launch, 1), PNB_ROW(pnb_dogbox_launch, 2)},
      {PNB_ROW(pnb_lm_launch, 0), PNB_ROW(pnb_lm_launch, 1), PNB_ROW(pnb_lm_launch, 2)},
#else
      {PNB_ROW(pnb_trf_launch, 0), PNB_NOROW, PNB_NOROW},
      {PNB_ROW(pnb_dogbox_launch, 0), PNB_NOROW, PNB_NOROW},
      {PNB_ROW(pnb_lm_launch, 0), PNB_NOROW, PNB_NOROW},
#endif
  };
  if (model_id < 0 || model_id > 6 || t1_mode < 0 || t1_mode > 2 || method < 0 || method > 2) return nullptr;
  return table[method][t1_mode][model_id];
}

// weights or a robust loss need the EXTRAS kernel
bool wants_extras(const pnb_trf_problem *p) {
  bool steps = false;
  for (int i = 0; i < 8; i++) steps = steps || p->diff_step[i] > 0.0;
  return p->weights != nullptr || p->loss != PNB_LOSS_LINEAR || steps;
}

int model_n_params(int model_id, int t1_mode) {
  static const int base[7] = {2, 3, 4, 4, 5, 6, 6};
  return base[model_id] + (t1_mode ? 1 : 0);
}

int check_problem(const pnb_trf_problem *p) {
  if (!p) return fail(PNB_E_BADARG, "null problem");
  if (p->model_id < 0 || p->model_id > 6 || p->t1_mode < 0 || p->t1_mode > 2)
    return fail(PNB_E_UNSUPPORTED, "unknown model_id / t1_mode");
  if (p->method != PNB_METHOD_TRF && p->method != PNB_METHOD_DOGBOX && p->method != PNB_METHOD_LM)
    return fail(PNB_E_UNSUPPORTED, "method must be PNB_METHOD_TRF, PNB_METHOD_DOGBOX or PNB_METHOD_LM");
  if (!trf_launcher(p->model_id, p->t1_mode, p->method))
    return fail(PNB_E_UNSUPPORTED, "this build has no kernel for the requested model / T1 mode");
  if (p->loss < PNB_LOSS_LINEAR || p->loss > PNB_LOSS_ARCTAN) return fail(PNB_E_BADARG, "loss must be one of PNB_LOSS_*");
  if (wants_extras(p) && !trf_launcher(p->model_id, p->t1_mode, p->method, true))
    return fail(PNB_E_UNSUPPORTED, "sigma / loss / diff_step are built for method trf without a T1 parameter");
  if (p->loss != PNB_LOSS_LINEAR && p->method == PNB_METHOD_LM)
    return fail(PNB_E_BADARG, "method='lm' supports only 'linear' loss function.");
  if (p->n_params != model_n_params(p->model_id, p->t1_mode))
    return fail(PNB_E_BADARG, "n_params does not match the model");
  if (p->n_b < 1 || p->n_b > 512) return fail(PNB_E_BADARG, "n_b must be in [1, 512]");
  if (p->n_vox < 0) return fail(PNB_E_BADARG, "n_vox < 0");
  if (p->max_nfev < 1) return fail(PNB_E_BADARG, "max_nfev must be positive");
  if (p->jac_mode < 0 || p->jac_mode > 2) return fail(PNB_E_BADARG, "jac_mode must be 0, 1 or 2");
  if ((p->jac_mode == 2) != (p->method == PNB_METHOD_LM) && p->jac_mode != 0)
    return fail(PNB_E_BADARG, "jac_mode 2 (MINPACK forward differences) goes with PNB_METHOD_LM, jac_mode 1 with trf / dogbox");
  if (p->n_vox > 0 &&
      (!p->xdata || !p->ydata || !p->p0 || !p->lb || !p->ub || !p->params || !p->status || !p->nfev))
    return fail(PNB_E_BADARG, "null array pointer");
  const unsigned all = (1u << p->n_params) - 1u;
  if ((p->frozen_mask & all) == all) return fail(PNB_E_BADARG, "all parameters are fixed");
  return 0;
}

pnb::TrfOptions make_options(const pnb_trf_problem *p) {
  pnb::TrfOptions o;
  o.ftol = p->ftol; o.xtol = p->xtol; o.gtol = p->gtol;
  o.max_nfev = p->max_nfev; o.jac_mode = p->jac_mode; o.x_scale_jac = p->x_scale_jac;
  o.method = p->method;
  o.frozen = p->frozen_mask & ((1u << p->n_params) - 1u);
  for (int i = 0; i < 8; i++) o.x_scale[i] = (p->x_scale[i] > 0.0) ? p->x_scale[i] : 1.0;
  o.tr = p->repetition_time; o.tm = p->mixing_time;
  o.finish_wait = (p->finish_wait > 0 && p->finish_wait <= 64) ? p->finish_wait : pnb::kTrfFinishWait;
  for (int i = 0; i < 8; i++) o.diff_step[i] = (p->diff_step[i] > 0.0) ? p->diff_step[i] : 0.0;
  o.loss = p->loss;
  o.f_scale = (p->f_scale > 0.0) ? p->f_scale : 1.0;
  o.absolute_sigma = p->absolute_sigma ? 1 : 0;
  return o;
}

int popcount(unsigned v) { int c = 0; while (v) { c += v & 1u; v >>= 1; } return c; }

// A small ring of work counters per device (one per in-flight launch).
struct CounterRing {
  unsigned long long *buf = nullptr;
  int next = 0;
  static constexpr int kSlots = 256;
};
std::mutex g_mu;
CounterRing g_rings[16];

int next_counter(unsigned long long **out) {
  int dev = 0;
  PNB_CUDA(cudaGetDevice(&dev));
  std::lock_guard<std::mutex> lk(g_mu);
  CounterRing &r = g_rings[dev & 15];
  if (!r.buf) PNB_CUDA(cudaMalloc(&r.buf, sizeof(unsigned long long) * CounterRing::kSlots));
  *out = r.buf + r.next;
  r.next = (r.next + 1) % CounterRing::kSlots;
  return 0;
}

}  // namespace

extern "C" int pnb_abi_version(void) { return PNB_ABI_VERSION; }
extern "C" int pnb_sizeof_trf_problem(void) { return (int)sizeof(pnb_trf_problem); }
extern "C" const char *pnb_last_error(void) { return g_err.c_str(); }
extern "C" int64_t pnb_launch_count(void) { return g_launches.load(); }

extern "C" int pnb_device_count(void) {
  int n = 0;
  if (cudaGetDeviceCount(&n) != cudaSuccess) { cudaGetLastError(); return 0; }
  return n;
}

extern "C" int pnb_host_alloc(void **ptr, int64_t bytes) {
  PNB_CUDA(cudaHostAlloc(ptr, (size_t)bytes, cudaHostAllocDefault));
  return 0;
}
extern "C" int pnb_host_free(void *ptr) {
  PNB_CUDA(cudaFreeHost(ptr));
  return 0;
}

extern "C" int pnb_trf_fit_device(const pnb_trf_problem *p, void *cuda_stream) {
  if (int rc = check_problem(p)) return rc;
  if (p->n_vox == 0) return 0;
  cudaStream_t stream = (cudaStream_t)cuda_stream;
  pnb::TrfDeviceArgs a;
  a.n_b = p->n_b; a.n_vox = p->n_vox; a.b = p->xdata; a.y = p->ydata;
  a.p0 = p->p0; a.lb = p->lb; a.ub = p->ub;
  a.p0_row_stride = p->p0_per_voxel ? p->n_vox : 1;
  a.p0_vox_stride = p->p0_per_voxel ? 1 : 0;
  a.bd_row_stride = p->bounds_per_voxel ? p->n_vox : 1;
  a.bd_vox_stride = p->bounds_per_voxel ? 1 : 0;
  a.opt = make_options(p);
  a.params = p->params; a.cov = p->cov; a.status = p->status; a.nfev = p->nfev;
  a.njev = p->njev; a.cost = p->cost; a.r2 = p->r_squared;
  if (int rc = next_counter(&a.counter)) return rc;
  a.n_failed = nullptr;
  a.w = p->weights;
  cudaError_t e = trf_launcher(p->model_id, p->t1_mode, p->method, wants_extras(p))(&a, stream);
  if (e != cudaSuccess) return cuda_fail(e, "trf kernel launch");
  g_launches.fetch_add(1);
  return 0;
}

// ---------------------------------------------------------------------
// host pipeline: chunks of voxels flow H2D -> kernel -> D2H on kSlots streams
// ---------------------------------------------------------------------
namespace {

struct Slot {
  cudaStream_t stream = nullptr;
  double *y = nullptr, *p0 = nullptr, *lb = nullptr, *ub = nullptr;
  double *params = nullptr, *cov = nullptr, *cost = nullptr, *r2 = nullptr;
  int *status = nullptr, *nfev = nullptr, *njev = nullptr;
  unsigned long long *counter = nullptr;
  cudaEvent_t kernel_done = nullptr;  // after the slot's most recent solver kernel
  size_t cap_y = 0, cap_p0 = 0, cap_lb = 0, cap_ub = 0, cap_par = 0, cap_cov = 0;
  size_t cap_cost = 0, cap_r2 = 0, cap_st = 0, cap_nf = 0, cap_nj = 0;
  // page-locked staging block for pageable caller memory
  char *pin = nullptr;
  size_t cap_pin = 0;
  size_t pend_start = 0, pend_n = 0;  // chunk whose results still sit in the staging block
};

struct Pipeline {
  static constexpr int kSlots = 3;
  Slot slots[kSlots];
  double *b = nullptr, *vec = nullptr;  // xdata, broadcast p0|lb|ub
  double *w = nullptr;                  // weights (1 / sigma), when the problem has them
  size_t cap_b = 0, cap_w = 0;
  unsigned long long *n_failed = nullptr;  // voxels of the current call that ended with status <= 0
  cudaStream_t cov_stream = nullptr;       // the one covariance pass over a device-resident covariance range
  bool init = false;
};
std::atomic<long long> g_last_failed{0};  // failures of the most recent host call (all devices of a multi call)
Pipeline g_pipes[16];
std::mutex g_pipe_mu[16];  // one host pipeline per device; different devices run concurrently

template <class T> int grow(T **ptr, size_t *cap, size_t need) {
  if (need <= *cap) return 0;
  if (*ptr) PNB_CUDA(cudaFree(*ptr));
  *ptr = nullptr; *cap = 0;
  PNB_CUDA(cudaMalloc(ptr, need * sizeof(T)));
  *cap = need;
  return 0;
}

}  // namespace

namespace {
// Voxels [v0, v1) of the problem through the host pipeline of `device`, results written into the
// same positions of the caller's arrays.  `cov_dev`: when not null, a DEVICE buffer on `device` that
// receives the covariances of voxels v0 .. v1-1 (row 0 = voxel v0) instead of p->cov: they stay on
// the GPU and cost no D2H traffic.
int trf_host_range(const pnb_trf_problem *p, int device, int64_t chunk_vox, size_t v0, size_t v1, double *cov_dev) {
  if (v1 <= v0) return 0;
  if (pnb_device_count() <= device || device < 0) return fail(PNB_E_NODEVICE, "no such CUDA device");
  pnbi::DeviceScope dev_scope(device);
  PNB_CUDA(dev_scope.error());
  std::lock_guard<std::mutex> lk(g_pipe_mu[device & 15]);
  Pipeline &P = g_pipes[device & 15];
  const int np = p->n_params, nb = p->n_b;
  const int nfree = np - popcount(p->frozen_mask & ((1u << np) - 1u));
  if (chunk_vox <= 0) chunk_vox = 1 << 18;
  if ((size_t)chunk_vox > v1 - v0) chunk_vox = (int64_t)(v1 - v0);
  const size_t C = (size_t)chunk_vox;
  double *const host_cov = cov_dev ? nullptr : p->cov;  // covariance that travels to the host
  const bool any_cov = cov_dev || p->cov;

  if (!P.init) {
    for (auto &s : P.slots) {
      PNB_CUDA(cudaStreamCreateWithFlags(&s.stream, cudaStreamNonBlocking));
      PNB_CUDA(cudaMalloc(&s.counter, sizeof(unsigned long long)));
      PNB_CUDA(cudaEventCreateWithFlags(&s.kernel_done, cudaEventDisableTiming));
    }
    PNB_CUDA(cudaStreamCreateWithFlags(&P.cov_stream, cudaStreamNonBlocking));
    PNB_CUDA(cudaMalloc(&P.vec, sizeof(double) * 3 * 8));
    PNB_CUDA(cudaMalloc(&P.n_failed, sizeof(unsigned long long)));
    P.init = true;
  }
  if (int rc = grow(&P.b, &P.cap_b, (size_t)nb)) return rc;
  if (p->weights)
    if (int rc = grow(&P.w, &P.cap_w, (size_t)nb)) return rc;
  for (auto &s : P.slots) {
    if (int rc = grow(&s.y, &s.cap_y, C * nb)) return rc;
    const size_t need_p = C * np;
    if (int rc = grow(&s.p0, &s.cap_p0, need_p)) return rc;
    if (int rc = grow(&s.lb, &s.cap_lb, need_p)) return rc;
    if (int rc = grow(&s.ub, &s.cap_ub, need_p)) return rc;
    if (int rc = grow(&s.params, &s.cap_par, need_p)) return rc;
    if (host_cov)
      if (int rc = grow(&s.cov, &s.cap_cov, C * nfree * nfree)) return rc;
    if (int rc = grow(&s.cost, &s.cap_cost, C)) return rc;
    if (int rc = grow(&s.r2, &s.cap_r2, C)) return rc;
    if (int rc = grow(&s.status, &s.cap_st, C)) return rc;
    if (int rc = grow(&s.nfev, &s.cap_nf, C)) return rc;
    if (int rc = grow(&s.njev, &s.cap_nj, C)) return rc;
  }
  // shared small inputs (synchronous, tiny)
  cudaStream_t s0 = P.slots[0].stream;
  PNB_CUDA(cudaMemsetAsync(P.n_failed, 0, sizeof(unsigned long long), s0));
  PNB_CUDA(cudaMemcpyAsync(P.b, p->xdata, sizeof(double) * nb, cudaMemcpyHostToDevice, s0));
  if (p->weights) PNB_CUDA(cudaMemcpyAsync(P.w, p->weights, sizeof(double) * nb, cudaMemcpyHostToDevice, s0));
  if (!p->p0_per_voxel)
    PNB_CUDA(cudaMemcpyAsync(P.vec, p->p0, sizeof(double) * np, cudaMemcpyHostToDevice, s0));
  if (!p->bounds_per_voxel) {
    PNB_CUDA(cudaMemcpyAsync(P.vec + 8, p->lb, sizeof(double) * np, cudaMemcpyHostToDevice, s0));
    PNB_CUDA(cudaMemcpyAsync(P.vec + 16, p->ub, sizeof(double) * np, cudaMemcpyHostToDevice, s0));
  }
  PNB_CUDA(cudaStreamSynchronize(s0));

  const pnb::TrfOptions opt = make_options(p);
  LaunchFn launch = trf_launcher(p->model_id, p->t1_mode, p->method, wants_extras(p));
  const size_t NV = (size_t)p->n_vox;
  // pageable caller memory is staged through page-locked blocks with multi-threaded host copies;
  // inputs and outputs independently (a caller may hold pageable images and page-locked result arrays)
  const bool stage_in = pnbi::is_pageable(p->ydata);
  const bool stage_out = pnbi::is_pageable(p->params);
  const bool staged = stage_in || stage_out;
  const size_t D = sizeof(double), I = sizeof(int);
  const size_t o_y = 0, o_p0 = o_y + C * nb * D, o_lb = o_p0 + C * np * D, o_ub = o_lb + C * np * D;
  const size_t o_par = o_ub + C * np * D, o_cov = o_par + C * np * D;
  const size_t o_cost = o_cov + (host_cov ? C * nfree * nfree * D : 0), o_r2 = o_cost + C * D, o_st = o_r2 + C * D;
  const size_t o_nf = o_st + C * I, o_nj = o_nf + C * I, pin_bytes = o_nj + C * I;
  if (staged) {
    for (auto &s : P.slots) {
      s.pend_n = 0;
      if (pin_bytes > s.cap_pin) {
        if (s.pin) PNB_CUDA(cudaFreeHost(s.pin));
        s.pin = nullptr; s.cap_pin = 0;
        PNB_CUDA(cudaHostAlloc((void **)&s.pin, pin_bytes, cudaHostAllocDefault));
        s.cap_pin = pin_bytes;
      }
    }
  }
  // copy a finished chunk from the slot's staging block to the caller's arrays
  auto drain = [&](Slot &s) -> int {
    if (!s.pend_n) return 0;
    PNB_CUDA(cudaStreamSynchronize(s.stream));
    if (!stage_out) { s.pend_n = 0; return 0; }  // only the staging block of the inputs had to be free
    const size_t st = s.pend_start, n = s.pend_n;
    for (int k = 0; k < np; k++)
      pnbi::parallel_memcpy(p->params + (size_t)k * NV + st, s.pin + o_par + (size_t)k * n * D, n * D);
    if (host_cov) pnbi::parallel_memcpy(host_cov + st * nfree * nfree, s.pin + o_cov, n * nfree * nfree * D);
    std::memcpy(p->status + st, s.pin + o_st, n * I);
    std::memcpy(p->nfev + st, s.pin + o_nf, n * I);
    if (p->njev) std::memcpy(p->njev + st, s.pin + o_nj, n * I);
    if (p->cost) std::memcpy(p->cost + st, s.pin + o_cost, n * D);
    if (p->r_squared) std::memcpy(p->r_squared + st, s.pin + o_r2, n * D);
    s.pend_n = 0;
    return 0;
  };
  int slot = 0;
  // Chunk schedule: C-sized chunks with a short ramp at both ends (C/4, C/2 ... C/2, C/4), so the first
  // kernel starts after a quarter of a chunk's upload and only a quarter chunk's results are still on
  // their way when the last kernel ends.
  std::vector<size_t> sizes;
  {
    size_t rem = v1 - v0;
    const size_t q = C / 4, h = C / 2;
    const bool ramp = q >= 4096 && rem > 3 * C;
    if (ramp) { sizes.push_back(q); sizes.push_back(h); rem -= q + h; }
    const size_t tail = ramp ? q + h : 0;
    while (rem > tail) {
      const size_t n = (rem - tail < C) ? rem - tail : C;
      sizes.push_back(n);
      rem -= n;
    }
    if (ramp) { sizes.push_back(h); sizes.push_back(q); }
  }
  size_t start = v0;
  for (size_t ci = 0; ci < sizes.size(); start += sizes[ci], ci++, slot = (slot + 1) % Pipeline::kSlots) {
    Slot &s = P.slots[slot];
    const size_t n = sizes[ci];
    const double *src_y = p->ydata + start * nb, *src_p0 = p->p0 + start, *src_lb = p->lb + start,
                 *src_ub = p->ub + start;
    size_t pitch = NV * D;  // row pitch of the caller's (n_params, n_vox) arrays
    if (staged)
      if (int rc = drain(s)) return rc;  // also waits until the slot's previous chunk is done
    if (stage_in) {
      pnbi::parallel_memcpy(s.pin + o_y, src_y, n * nb * D);
      src_y = reinterpret_cast<const double *>(s.pin + o_y);
      if (p->p0_per_voxel) {
        for (int k = 0; k < np; k++) std::memcpy(s.pin + o_p0 + (size_t)k * n * D, p->p0 + (size_t)k * NV + start, n * D);
        src_p0 = reinterpret_cast<const double *>(s.pin + o_p0);
      }
      if (p->bounds_per_voxel) {
        for (int k = 0; k < np; k++) {
          std::memcpy(s.pin + o_lb + (size_t)k * n * D, p->lb + (size_t)k * NV + start, n * D);
          std::memcpy(s.pin + o_ub + (size_t)k * n * D, p->ub + (size_t)k * NV + start, n * D);
        }
        src_lb = reinterpret_cast<const double *>(s.pin + o_lb);
        src_ub = reinterpret_cast<const double *>(s.pin + o_ub);
      }
      pitch = n * D;
    }
    // the slot's previous chunk has been fully enqueued on the same stream, so
    // stream order already protects the device buffers; no host sync needed here.
    PNB_CUDA(cudaMemcpyAsync(s.y, src_y, D * n * nb, cudaMemcpyHostToDevice, s.stream));
    if (p->p0_per_voxel)
      PNB_CUDA(cudaMemcpy2DAsync(s.p0, n * D, src_p0, pitch, n * D, np, cudaMemcpyHostToDevice, s.stream));
    if (p->bounds_per_voxel) {
      PNB_CUDA(cudaMemcpy2DAsync(s.lb, n * D, src_lb, pitch, n * D, np, cudaMemcpyHostToDevice, s.stream));
      PNB_CUDA(cudaMemcpy2DAsync(s.ub, n * D, src_ub, pitch, n * D, np, cudaMemcpyHostToDevice, s.stream));
    }
    pnb::TrfDeviceArgs a;
    a.n_b = nb; a.n_vox = (long long)n; a.b = P.b; a.y = s.y;
    a.w = p->weights ? P.w : nullptr;
    a.p0 = p->p0_per_voxel ? s.p0 : P.vec;
    a.lb = p->bounds_per_voxel ? s.lb : P.vec + 8;
    a.ub = p->bounds_per_voxel ? s.ub : P.vec + 16;
    a.p0_row_stride = p->p0_per_voxel ? (long long)n : 1;
    a.p0_vox_stride = p->p0_per_voxel ? 1 : 0;
    a.bd_row_stride = p->bounds_per_voxel ? (long long)n : 1;
    a.bd_vox_stride = p->bounds_per_voxel ? 1 : 0;
    a.opt = opt;
    a.params = s.params; a.status = s.status; a.nfev = s.nfev;
    a.cov = cov_dev ? cov_dev + (start - v0) * nfree * nfree : (host_cov ? s.cov : nullptr);
    a.njev = p->njev ? s.njev : nullptr; a.cost = p->cost ? s.cost : nullptr;
    a.r2 = p->r_squared ? s.r2 : nullptr;
    a.counter = s.counter;
    a.n_failed = P.n_failed;
    // covariances that stay on the GPU: the chunks only park J^T J and the cost, ONE pass over the
    // whole range follows the last chunk instead of a small kernel queued behind every chunk (which
    // held each chunk's downloads back while the next chunk's persistent grid owned the SMs)
    a.defer_cov = cov_dev ? 1 : 0;
    cudaError_t e = launch(&a, s.stream);
    if (e != cudaSuccess) return cuda_fail(e, "trf kernel launch");
    if (cov_dev) PNB_CUDA(cudaEventRecord(s.kernel_done, s.stream));
    g_launches.fetch_add(1);
    if (staged) { s.pend_start = start; s.pend_n = n; }
    if (stage_out) {
      PNB_CUDA(cudaMemcpyAsync(s.pin + o_par, s.params, D * n * np, cudaMemcpyDeviceToHost, s.stream));
      if (host_cov)
        PNB_CUDA(cudaMemcpyAsync(s.pin + o_cov, s.cov, D * n * nfree * nfree, cudaMemcpyDeviceToHost, s.stream));
      PNB_CUDA(cudaMemcpyAsync(s.pin + o_st, s.status, I * n, cudaMemcpyDeviceToHost, s.stream));
      PNB_CUDA(cudaMemcpyAsync(s.pin + o_nf, s.nfev, I * n, cudaMemcpyDeviceToHost, s.stream));
      if (p->njev) PNB_CUDA(cudaMemcpyAsync(s.pin + o_nj, s.njev, I * n, cudaMemcpyDeviceToHost, s.stream));
      if (p->cost) PNB_CUDA(cudaMemcpyAsync(s.pin + o_cost, s.cost, D * n, cudaMemcpyDeviceToHost, s.stream));
      if (p->r_squared) PNB_CUDA(cudaMemcpyAsync(s.pin + o_r2, s.r2, D * n, cudaMemcpyDeviceToHost, s.stream));
      continue;
    }
    PNB_CUDA(cudaMemcpy2DAsync(p->params + start, NV * D, s.params, n * D, n * D, np, cudaMemcpyDeviceToHost, s.stream));
    if (host_cov)
      PNB_CUDA(cudaMemcpyAsync(host_cov + start * nfree * nfree, s.cov, D * n * nfree * nfree,
                               cudaMemcpyDeviceToHost, s.stream));
    PNB_CUDA(cudaMemcpyAsync(p->status + start, s.status, I * n, cudaMemcpyDeviceToHost, s.stream));
    PNB_CUDA(cudaMemcpyAsync(p->nfev + start, s.nfev, I * n, cudaMemcpyDeviceToHost, s.stream));
    if (p->njev)
      PNB_CUDA(cudaMemcpyAsync(p->njev + start, s.njev, I * n, cudaMemcpyDeviceToHost, s.stream));
    if (p->cost)
      PNB_CUDA(cudaMemcpyAsync(p->cost + start, s.cost, D * n, cudaMemcpyDeviceToHost, s.stream));
    if (p->r_squared)
      PNB_CUDA(cudaMemcpyAsync(p->r_squared + start, s.r2, D * n, cudaMemcpyDeviceToHost, s.stream));
  }
  if (cov_dev && nfree >= 2) {
    const size_t n_chunks = sizes.size();
    for (int k = 0; k < Pipeline::kSlots && (size_t)k < n_chunks; k++)
      PNB_CUDA(cudaStreamWaitEvent(P.cov_stream, P.slots[k].kernel_done, 0));
    cudaError_t e = pnb::trf_cov_launch(nfree, (long long)(v1 - v0), nb, nullptr, cov_dev, P.cov_stream,
                                        opt.absolute_sigma);
    if (e != cudaSuccess) return cuda_fail(e, "covariance kernel launch");
  }
  if (staged)
    for (int k = 0; k < Pipeline::kSlots; k++) {
      // drain in submission order so the copies overlap the chunks still running
      if (int rc = drain(P.slots[(slot + k) % Pipeline::kSlots])) return rc;
    }
  for (auto &s : P.slots) PNB_CUDA(cudaStreamSynchronize(s.stream));
  if (cov_dev) PNB_CUDA(cudaStreamSynchronize(P.cov_stream));
  unsigned long long nf = 0;
  PNB_CUDA(cudaMemcpy(&nf, P.n_failed, sizeof(nf), cudaMemcpyDeviceToHost));
  g_last_failed.fetch_add((long long)nf);
  (void)any_cov;
  return 0;
}

// true when `ptr` is device memory (the covariance may be left on the GPU)
bool is_device_ptr(const void *ptr, int *dev_out) {
  if (!ptr) return false;
  cudaPointerAttributes attr;
  if (cudaPointerGetAttributes(&attr, ptr) != cudaSuccess) { cudaGetLastError(); return false; }
  if (attr.type != cudaMemoryTypeDevice) return false;
  if (dev_out) *dev_out = attr.device;
  return true;
}
}  // namespace

extern "C" int64_t pnb_trf_last_failed_count(void) { return (int64_t)g_last_failed.load(); }

extern "C" int pnb_trf_fit_host(const pnb_trf_problem *p, int device, int64_t chunk_vox) {
  if (int rc = check_problem(p)) return rc;
  g_last_failed.store(0);
  if (p->n_vox == 0) return 0;
  int cov_device = -1;
  double *cov_dev = is_device_ptr(p->cov, &cov_device) ? p->cov : nullptr;
  if (cov_dev && cov_device != device) return fail(PNB_E_BADARG, "cov is device memory of another GPU");
  return trf_host_range(p, device, chunk_vox, 0, (size_t)p->n_vox, cov_dev);
}

// One host pipeline per GPU, run concurrently by one thread each; the voxels are cut into
// contiguous, balanced ranges and every GPU writes straight into its part of the caller's arrays
// (no gather step: the "parameter maps to rank 0" of the multi-process mode is the D2H copy itself).
extern "C" int pnb_trf_fit_host_multi(const pnb_trf_problem *p, const int32_t *devices, int32_t n_devices,
                                      int64_t chunk_vox, double *const *cov_per_device) {
  if (int rc = check_problem(p)) return rc;
  if (n_devices < 1 || n_devices > 16) return fail(PNB_E_BADARG, "n_devices must be in [1, 16]");
  g_last_failed.store(0);
  if (p->n_vox == 0) return 0;
  const size_t NV = (size_t)p->n_vox;
  std::vector<int> rcs(n_devices, 0);
  std::vector<std::string> errs(n_devices);
  std::vector<std::thread> workers;
  const size_t base = NV / n_devices, rem = NV % n_devices;
  size_t start = 0;
  for (int i = 0; i < n_devices; i++) {
    const size_t stop = start + base + ((size_t)i < rem ? 1 : 0);
    const int dev = devices ? devices[i] : i;
    double *cov_dev = cov_per_device ? cov_per_device[i] : nullptr;
    workers.emplace_back([=, &rcs, &errs] {
      rcs[i] = trf_host_range(p, dev, chunk_vox, start, stop, cov_dev);
      if (rcs[i]) errs[i] = g_err;  // thread-local text of this worker
    });
    start = stop;
  }
  for (auto &w : workers) w.join();
  for (int i = 0; i < n_devices; i++)
    if (rcs[i]) { g_err = "device " + std::to_string(devices ? devices[i] : i) + ": " + errs[i]; return rcs[i]; }
  return 0;
}

// ---------------------------------------------------------------------
// Staged transfers between device memory and PAGEABLE host memory.
// cudaMemcpy on pageable memory goes through the driver's own bounce buffer at
// 2-3 GB/s; here the bytes travel in 32 MB pieces through two page-locked
// blocks, the DMA of one piece overlapping the multi-threaded host copy of the
// other (~12-15 GB/s, bound by the host copy).  Page-locked host memory is
// copied directly.
// ---------------------------------------------------------------------
namespace {
struct Bounce {
  char *blk[2] = {nullptr, nullptr};
  cudaStream_t st[2] = {nullptr, nullptr};
  static constexpr size_t kPiece = 32u << 20;
};
Bounce g_bounce[16];
std::mutex g_bounce_mu;

int bounce_for(int dev, Bounce **out) {
  Bounce &b = g_bounce[dev & 15];
  for (int i = 0; i < 2; i++) {
    if (!b.blk[i]) PNB_CUDA(cudaHostAlloc((void **)&b.blk[i], Bounce::kPiece, cudaHostAllocDefault));
    if (!b.st[i]) PNB_CUDA(cudaStreamCreateWithFlags(&b.st[i], cudaStreamNonBlocking));
  }
  *out = &b;
  return 0;
}
}  // namespace

// Device-to-device copy enqueued on `cuda_stream` of the CURRENT device (a peer copy when dst lives
// on another GPU, e.g. memory mapped through CUDA IPC): the source GPU's copy engine pushes the
// bytes over NVLink, nothing runs in the destination GPU's context.
extern "C" int pnb_copy_d2d(void *dst, const void *src, int64_t bytes, void *cuda_stream) {
  if (bytes <= 0) return 0;
  if (!dst || !src) return fail(PNB_E_BADARG, "null pointer");
  int cur = 0;
  PNB_CUDA(cudaGetDevice(&cur));
  cudaPointerAttributes ad;
  PNB_CUDA(cudaPointerGetAttributes(&ad, dst));
  if (ad.type == cudaMemoryTypeDevice && ad.device != cur) {
    // without peer access a copy into another GPU's memory is routed over PCIe (~50 GB/s measured,
    // against 700 GB/s over NVLink); memory mapped through CUDA IPC does not enable it by itself
    int can = 0;
    PNB_CUDA(cudaDeviceCanAccessPeer(&can, cur, ad.device));
    if (can) {
      cudaError_t e = cudaDeviceEnablePeerAccess(ad.device, 0);
      if (e != cudaSuccess && e != cudaErrorPeerAccessAlreadyEnabled) return cuda_fail(e, "cudaDeviceEnablePeerAccess");
      cudaGetLastError();
    }
    PNB_CUDA(cudaMemcpyPeerAsync(dst, ad.device, src, cur, (size_t)bytes, (cudaStream_t)cuda_stream));
    return 0;
  }
  PNB_CUDA(cudaMemcpyAsync(dst, src, (size_t)bytes, cudaMemcpyDefault, (cudaStream_t)cuda_stream));
  return 0;
}

extern "C" int pnb_download(void *host_dst, const void *dev_src, int64_t bytes, void *after_stream) {
  if (bytes <= 0) return 0;
  if (!host_dst || !dev_src) return fail(PNB_E_BADARG, "null pointer");
  // the producer's work must be finished before the side streams read it
  PNB_CUDA(cudaStreamSynchronize((cudaStream_t)after_stream));
  if (!pnbi::is_pageable(host_dst)) {
    PNB_CUDA(cudaMemcpy(host_dst, dev_src, (size_t)bytes, cudaMemcpyDeviceToHost));
    return 0;
  }
  int dev = 0;
  PNB_CUDA(cudaGetDevice(&dev));
  std::lock_guard<std::mutex> lk(g_bounce_mu);
  Bounce *b = nullptr;
  if (int rc = bounce_for(dev, &b)) return rc;
  const size_t n = (size_t)bytes, P = Bounce::kPiece;
  const size_t pieces = (n + P - 1) / P;
  auto len = [&](size_t i) { return (i + 1) * P <= n ? P : n - i * P; };
  PNB_CUDA(cudaMemcpyAsync(b->blk[0], dev_src, len(0), cudaMemcpyDeviceToHost, b->st[0]));
  for (size_t i = 0; i < pieces; i++) {
    if (i + 1 < pieces)
      PNB_CUDA(cudaMemcpyAsync(b->blk[(i + 1) & 1], (const char *)dev_src + (i + 1) * P, len(i + 1),
                               cudaMemcpyDeviceToHost, b->st[(i + 1) & 1]));
    PNB_CUDA(cudaStreamSynchronize(b->st[i & 1]));
    pnbi::parallel_memcpy((char *)host_dst + i * P, b->blk[i & 1], len(i));
  }
  return 0;
}

extern "C" int pnb_upload(void *dev_dst, const void *host_src, int64_t bytes, void *then_stream) {
  if (bytes <= 0) return 0;
  if (!dev_dst || !host_src) return fail(PNB_E_BADARG, "null pointer");
  if (!pnbi::is_pageable(host_src)) {
    PNB_CUDA(cudaMemcpyAsync(dev_dst, host_src, (size_t)bytes, cudaMemcpyHostToDevice, (cudaStream_t)then_stream));
    PNB_CUDA(cudaStreamSynchronize((cudaStream_t)then_stream));
    return 0;
  }
  int dev = 0;
  PNB_CUDA(cudaGetDevice(&dev));
  std::lock_guard<std::mutex> lk(g_bounce_mu);
  Bounce *b = nullptr;
  if (int rc = bounce_for(dev, &b)) return rc;
  const size_t n = (size_t)bytes, P = Bounce::kPiece;
  const size_t pieces = (n + P - 1) / P;
  auto len = [&](size_t i) { return (i + 1) * P <= n ? P : n - i * P; };
  // dev_dst may be a block the caller's allocator has just recycled: work queued on `then_stream`
  // (a kernel of an earlier chunk, say) can still be reading it.  The side streams write it, so they
  // must run after everything `then_stream` holds at this point (the host-side staging copies of
  // the first two pieces still overlap that work).
  {
    cudaEvent_t ev;
    PNB_CUDA(cudaEventCreateWithFlags(&ev, cudaEventDisableTiming));
    cudaError_t e = cudaEventRecord(ev, (cudaStream_t)then_stream);
    if (e == cudaSuccess) e = cudaStreamWaitEvent(b->st[0], ev, 0);
    if (e == cudaSuccess) e = cudaStreamWaitEvent(b->st[1], ev, 0);
    cudaEventDestroy(ev);
    if (e != cudaSuccess) return cuda_fail(e, "ordering the upload after then_stream");
  }
  for (size_t i = 0; i < pieces; i++) {
    PNB_CUDA(cudaStreamSynchronize(b->st[i & 1]));  // the block's previous piece has left
    pnbi::parallel_memcpy(b->blk[i & 1], (const char *)host_src + i * P, len(i));
    PNB_CUDA(cudaMemcpyAsync((char *)dev_dst + i * P, b->blk[i & 1], len(i), cudaMemcpyHostToDevice, b->st[i & 1]));
  }
  PNB_CUDA(cudaStreamSynchronize(b->st[0]));
  PNB_CUDA(cudaStreamSynchronize(b->st[1]));
  (void)then_stream;  // both copies have completed: any stream may use the data
  return 0;
}

// ---------------------------------------------------------------------
// FP64 FMA peak (roofline denominator; MEASURED_PEAKS.json has no FP64 figure)
// ---------------------------------------------------------------------
namespace {
__global__ void __launch_bounds__(256) dfma_peak_kernel(double *out, int iters, double seed) {
  double a0 = seed + threadIdx.x, a1 = a0 + 1, a2 = a0 + 2, a3 = a0 + 3, a4 = a0 + 4, a5 = a0 + 5,
         a6 = a0 + 6, a7 = a0 + 7;
  const double m = 0.999999, c = 1e-9;
  for (int i = 0; i < iters; i++) {
    a0 = fma(a0, m, c); a1 = fma(a1, m, c); a2 = fma(a2, m, c); a3 = fma(a3, m, c);
    a4 = fma(a4, m, c); a5 = fma(a5, m, c); a6 = fma(a6, m, c); a7 = fma(a7, m, c);
  }
  out[blockIdx.x * blockDim.x + threadIdx.x] = a0 + a1 + a2 + a3 + a4 + a5 + a6 + a7;
}
}  // namespace

extern "C" int pnb_measure_fp64_peak(int device, double *tflops) {
  if (!tflops) return fail(PNB_E_BADARG, "null output");
  if (pnb_device_count() <= device || device < 0) return fail(PNB_E_NODEVICE, "no such CUDA device");
  pnbi::DeviceScope dev_scope(device);
  PNB_CUDA(dev_scope.error());
  int sms = 0;
  PNB_CUDA(cudaDeviceGetAttribute(&sms, cudaDevAttrMultiProcessorCount, device));
  const int blocks = sms * 8, threads = 256, iters = 1 << 15;
  double *out = nullptr;
  PNB_CUDA(cudaMalloc(&out, sizeof(double) * blocks * threads));
  cudaEvent_t e0, e1;
  PNB_CUDA(cudaEventCreate(&e0));
  PNB_CUDA(cudaEventCreate(&e1));
  double best = 0.0;
  for (int rep = 0; rep < 6; rep++) {
    PNB_CUDA(cudaEventRecord(e0));
    dfma_peak_kernel<<<blocks, threads>>>(out, iters, 1.0 + rep);
    PNB_CUDA(cudaEventRecord(e1));
    PNB_CUDA(cudaEventSynchronize(e1));
    g_launches.fetch_add(1);
    float ms = 0;
    PNB_CUDA(cudaEventElapsedTime(&ms, e0, e1));
    const double flops = 2.0 * 8.0 * (double)iters * blocks * threads;
    const double tf = flops / (ms * 1e-3) / 1e12;
    if (rep > 0 && tf > best) best = tf;
  }
  cudaEventDestroy(e0); cudaEventDestroy(e1); cudaFree(out);
  *tflops = best;
  return 0;
}
