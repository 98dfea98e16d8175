// C ABI of the batched NNLS solver (see include/pyneapple_b200.h).
#include <cuda_runtime.h>

#include <cstdlib>
#include <cstring>
#include <mutex>
#include <string>
#include <thread>
#include <vector>

#include "../../include/pyneapple_b200.h"
#include "pnb_internal.h"
#include "pnb_nnls_fast.cuh"
#include "pnb_nnls_gemm.cuh"

namespace {

#ifndef PNB_NNLS_WARPS
#define PNB_NNLS_WARPS 8
#endif
constexpr int kWarps = PNB_NNLS_WARPS;        // robust kernel
#ifndef PNB_NNLS_FAST_WARPS
#define PNB_NNLS_FAST_WARPS 8
#endif
constexpr int kFastWarps = PNB_NNLS_FAST_WARPS;  // fast kernel
constexpr size_t kSmemBudget = 220 * 1024;
// host pipeline depth: with three chunks in flight the next chunk's fast kernel fills the SMs while
// the few voxels a chunk hands to the robust kernel (one long voxel = milliseconds) are re-solved
constexpr int kSlots = 3;

struct NnlsCtx {
  unsigned long long *counters = nullptr;  // ring
  int next = 0;
  // one set per concurrently running launch (the host pipeline rotates over kSlots streams)
  double *scratch[kSlots] = {};
  size_t scratch_cap[kSlots] = {};
  int *redo_list[kSlots] = {};
  size_t redo_cap[kSlots] = {};
  double *h0[kSlots] = {};  // dual_init = 1: H0 = Y B materialised by the tensor-core GEMM
  size_t h0_cap[kSlots] = {};
  unsigned long long *last_redo = nullptr;  // device counter of the most recent auto-mode launch
  // host pipeline
  cudaStream_t streams[kSlots] = {};
  double *y[kSlots] = {}, *coef[kSlots] = {}, *rn[kSlots] = {}, *r2[kSlots] = {};
  int *st[kSlots] = {}, *it[kSlots] = {};
  size_t cap_y[kSlots] = {}, cap_coef[kSlots] = {}, cap_rn[kSlots] = {}, cap_r2[kSlots] = {};
  size_t cap_st[kSlots] = {}, cap_it[kSlots] = {};
  // device-path launches share slot 0's scratch: each one waits for the previous one's event
  cudaEvent_t dev_done = nullptr;
  double *B = nullptr, *rtr = nullptr;
  size_t cap_B = 0, cap_rtr = 0;
  // page-locked staging blocks for pageable caller memory
  char *pin[kSlots] = {};
  size_t cap_pin[kSlots] = {};
  size_t pend_start[kSlots] = {}, pend_n[kSlots] = {};
  // host pipeline: ONE hand-over list for all chunks of a call (see nnls_host_range)
  int *defer_list = nullptr;
  size_t defer_cap = 0;
  unsigned long long *defer_count = nullptr;
};
NnlsCtx g_ctx[16];

// Hand-over of a host-pipeline call: the chunks' fast kernels append to one list, the robust pass runs
// once after the last chunk.
struct Deferred {
  int *list;
  unsigned long long *count;
  int base;
};
// PNB_NNLS_NO_V3=1 in the environment keeps the second-generation fast kernel (A/B measurements)
const bool g_disable_v3 = [] { const char *e = std::getenv("PNB_NNLS_NO_V3"); return e && e[0] == '1'; }();
// certification threshold of the fast path (NnlsDeviceArgs::cert_ztol): a would-be coefficient above
// it sends the voxel to the robust path.  2e-7 keeps the worst case a factor 5 inside the 1e-6
// absolute parity tolerance; PNB_NNLS_CERT_ZTOL overrides it (measurements).
const double g_cert_ztol = [] { const char *e = std::getenv("PNB_NNLS_CERT_ZTOL"); return e ? std::atof(e) : 2e-7; }();
// PNB_NNLS_SCREEN=1 switches the FP32 screening of the fast kernel's dual pass on; it only has an effect in
// a library built with -DPNB_V3_SCREEN (an experiment that lost, see pnb_nnls_v3.cuh)
const int g_screen = [] { const char *e = std::getenv("PNB_NNLS_SCREEN"); return (e && e[0] == '1') ? 1 : 0; }();
std::mutex g_mu[16];  // per device: the host pipelines of different GPUs run concurrently

// (re)allocate a device buffer of `need` elements; pointer and capacity stay consistent on failure
template <class T> int grow(T **ptr, size_t *cap, size_t need) {
  if (need <= *cap) return 0;
  if (*ptr) cudaFree(*ptr);
  *ptr = nullptr; *cap = 0;
  PNBI_CUDA(cudaMalloc(ptr, need * sizeof(T)));
  *cap = need;
  return 0;
}

int check(const pnb_nnls_problem *p) {
  if (!p) return pnbi::fail(PNB_E_BADARG, "null problem");
  if (p->n_b < 1 || p->n_b > 512) return pnbi::fail(PNB_E_BADARG, "n_b must be in [1, 512]");
  if (p->n_bins < 1 || p->n_bins > 1024) return pnbi::fail(PNB_E_BADARG, "n_bins must be in [1, 1024]");
  if (p->rtr_halfband < 0 || p->rtr_halfband > 8) return pnbi::fail(PNB_E_BADARG, "rtr_halfband must be in [0, 8]");
  if (p->max_iter < 1) return pnbi::fail(PNB_E_BADARG, "max_iter must be positive");
  if (p->algorithm != 0 && p->algorithm != 1) return pnbi::fail(PNB_E_BADARG, "algorithm must be 0 or 1");
  if (p->dual_init != 0 && p->dual_init != 1) return pnbi::fail(PNB_E_BADARG, "dual_init must be 0 (fused) or 1 (GEMM)");
  if (p->n_vox < 0) return pnbi::fail(PNB_E_BADARG, "n_vox < 0");
  if (p->n_vox > 0 && (!p->basis || !p->rtr_band || !p->signal || !p->coefficients || !p->residual ||
                       !p->status || !p->iterations))
    return pnbi::fail(PNB_E_BADARG, "null array pointer");
  return 0;
}

// the largest active-set size whose factor fits shared memory next to everything else
int pick_kmax(int m, int n, int W, int warps) {
  int k = n;
  while (k > 8 && pnb::nnls_smem_bytes(m, n, W, k, warps) > kSmemBudget) k--;
  return k;
}
constexpr int kRedoWarps = 2;  // the few voxels the fast path hands over have large active sets:
                               // fewer warps per SM, but a factor that stays in shared memory

int pick_mt(int m) { return m <= 8 ? 8 : (m <= 16 ? 16 : (m <= 24 ? 24 : (m <= 32 ? 32 : 0))); }

int pick_kcap_fast(int mt, int n, int W) {
  int k = n < 96 ? n : 96;
  while (k > 8 && pnb::nnls_fast_smem_bytes(mt, n, W, k, kFastWarps) > kSmemBudget) k--;
  return k;
}

// v3 fast kernel (n_bins <= 256): one translation unit per (MT, WK), see pnb_nnls_v3_inst.cu
#define PNB_V3_DECL(mt, wk) extern "C" cudaError_t pnb_nnls_v3_launch_##mt##_##wk(const pnb::NnlsDeviceArgs *, cudaStream_t);
#define PNB_V3_ALL(X) X(8, 0) X(8, 2) X(8, 4) X(16, 0) X(16, 2) X(16, 4) X(24, 0) X(24, 2) X(24, 4) X(32, 0) X(32, 2) X(32, 4)
PNB_V3_ALL(PNB_V3_DECL)
using V3Launch = cudaError_t (*)(const pnb::NnlsDeviceArgs *, cudaStream_t);
V3Launch v3_launcher_for(int mt, int n, int W) {
  if (n > 256 || W > 4 || mt == 0) return nullptr;
  const int wk = W == 0 ? 0 : (W <= 2 ? 2 : 4);
#define PNB_V3_PICK(m_, w_) if (mt == m_ && wk == w_) return pnb_nnls_v3_launch_##m_##_##w_;
  PNB_V3_ALL(PNB_V3_PICK)
  return nullptr;
}

using FastKernel = void (*)(const pnb::NnlsDeviceArgs);
FastKernel fast_kernel_for(int mt) {
  switch (mt) {
    case 8: return pnb::nnls_fast_kernel<kFastWarps, 8>;
    case 16: return pnb::nnls_fast_kernel<kFastWarps, 16>;
    case 24: return pnb::nnls_fast_kernel<kFastWarps, 24>;
    case 32: return pnb::nnls_fast_kernel<kFastWarps, 32>;
  }
  return nullptr;
}

int h0_gemm(int mt, const double *y, const double *B, double *h0, long long n_vox, int m, int n, cudaStream_t stream) {
  cudaError_t e = cudaErrorInvalidValue;
  switch (mt) {
    case 8: e = pnb::nnls_h0_dmma_launch<8>(y, B, h0, n_vox, m, n, stream); break;
    case 16: e = pnb::nnls_h0_dmma_launch<16>(y, B, h0, n_vox, m, n, stream); break;
    case 24: e = pnb::nnls_h0_dmma_launch<24>(y, B, h0, n_vox, m, n, stream); break;
    case 32: e = pnb::nnls_h0_dmma_launch<32>(y, B, h0, n_vox, m, n, stream); break;
  }
  if (e != cudaSuccess) return pnbi::cuda_fail(e, "nnls_h0_dmma_kernel launch");
  pnbi::count_launch();
  return 0;
}

int launch(NnlsCtx &C, const pnb_nnls_problem *p, const double *B, const double *rtr, const double *y,
           long long n_vox, double *coef, double *rn, int *st, int *it, double *r2, cudaStream_t stream,
           int slot, const Deferred *defer = nullptr, int algorithm = -1) {
  if (algorithm < 0) algorithm = p->algorithm;
  const int m = p->n_b, n = p->n_bins, W = p->rtr_halfband;
  const int kmax = pick_kmax(m, n, W, kWarps);
  const size_t smem = pnb::nnls_smem_bytes(m, n, W, kmax, kWarps);
  if (smem > 227 * 1024) return pnbi::fail(PNB_E_UNSUPPORTED, "n_b x n_bins too large for shared memory");
  auto robust = pnb::nnls_kernel<kWarps>;
  const int mt = pick_mt(m);
  FastKernel fast = fast_kernel_for(mt);
  const bool use_fast = algorithm == 0 && fast != nullptr;  // more than 32 measurements: robust only
  V3Launch v3 = (use_fast && !g_disable_v3) ? v3_launcher_for(mt, n, W) : nullptr;
  const int kcap = use_fast ? pick_kcap_fast(mt, n, W) : 0;
  const size_t smem_fast = use_fast ? pnb::nnls_fast_smem_bytes(mt, n, W, kcap, kFastWarps) : 0;
  PNBI_CUDA(cudaFuncSetAttribute(robust, cudaFuncAttributeMaxDynamicSharedMemorySize, (int)smem));
  if (use_fast) PNBI_CUDA(cudaFuncSetAttribute(fast, cudaFuncAttributeMaxDynamicSharedMemorySize, (int)smem_fast));
  int dev = 0, sms = 0, bps = 0, bps_fast = 0;
  PNBI_CUDA(cudaGetDevice(&dev));
  PNBI_CUDA(cudaDeviceGetAttribute(&sms, cudaDevAttrMultiProcessorCount, dev));
  PNBI_CUDA(cudaOccupancyMaxActiveBlocksPerMultiprocessor(&bps, robust, kWarps * 32, smem));
  if (use_fast)
    PNBI_CUDA(cudaOccupancyMaxActiveBlocksPerMultiprocessor(&bps_fast, fast, kFastWarps * 32, smem_fast));
  if (bps < 1 || (use_fast && bps_fast < 1)) return pnbi::fail(PNB_E_UNSUPPORTED, "NNLS kernel does not fit on this device");
  const long long want = (n_vox + kWarps - 1) / kWarps;
  const size_t per_warp = (size_t)n * (n + 1) / 2 + 5 * (size_t)n;
  const size_t need = (size_t)bps * sms * kWarps * per_warp;
  if (int rc = grow(&C.scratch[slot], &C.scratch_cap[slot], need)) return rc;
  if (int rc = grow(&C.redo_list[slot], &C.redo_cap[slot], (size_t)n_vox)) return rc;
  if (!C.counters) PNBI_CUDA(cudaMalloc(&C.counters, 64 * sizeof(unsigned long long)));
  // three consecutive counters per call: fast work queue, redo count, robust work queue
  unsigned long long *ctr = C.counters + C.next;
  C.next = (C.next + 3) % 60;
  PNBI_CUDA(cudaMemsetAsync(ctr, 0, 3 * sizeof(unsigned long long), stream));
  pnb::NnlsDeviceArgs a;
  a.m = m; a.n = n; a.W = W; a.maxiter = p->max_iter; a.n_vox = n_vox;
  a.B = B; a.rtr = rtr; a.y = y; a.coef = coef; a.rnorm = rn; a.status = st; a.iters = it; a.r2 = r2;
  a.scratch = C.scratch[slot];
  a.cert_ztol = g_cert_ztol;
  a.h0 = nullptr;
  a.screen = g_screen;
  if (p->dual_init == 1 && v3) {
    if (int rc = grow(&C.h0[slot], &C.h0_cap[slot], (size_t)n_vox * n)) return rc;
    if (int rc = h0_gemm(mt, y, B, C.h0[slot], n_vox, m, n, stream)) return rc;
    a.h0 = C.h0[slot];
  }
  a.redo_count = ctr + 1; a.redo_list = C.redo_list[slot];
  if (defer && use_fast) { a.redo_count = defer->count; a.redo_list = defer->list; a.redo_base = defer->base; }
  C.last_redo = use_fast ? a.redo_count : nullptr;
  if (use_fast) {
    long long grid = (long long)bps_fast * sms;
    const long long want_fast = (n_vox + kFastWarps - 1) / kFastWarps;
    if (want_fast < grid) grid = want_fast;
    if (grid < 1) grid = 1;
    a.counter = ctr; a.kmax = kcap; a.work_list = nullptr; a.work_count = nullptr;
    a.work_min = 0; a.work_max = ~0ULL;
    if (v3) {
      PNBI_CUDA(v3(&a, stream));
    } else {
      fast<<<(unsigned)grid, kFastWarps * 32, smem_fast, stream>>>(a);
      PNBI_CUDA(cudaGetLastError());
    }
    pnbi::count_launch();
    if (defer) return 0;
    // voxels the fast path could not certify (the count stays on the device)
    auto redo = pnb::nnls_kernel<kRedoWarps>;
    const int kmax_redo = pick_kmax(m, n, W, kRedoWarps);
    const size_t smem_redo = pnb::nnls_smem_bytes(m, n, W, kmax_redo, kRedoWarps);
    PNBI_CUDA(cudaFuncSetAttribute(redo, cudaFuncAttributeMaxDynamicSharedMemorySize, (int)smem_redo));
    int bps_redo = 0;
    PNBI_CUDA(cudaOccupancyMaxActiveBlocksPerMultiprocessor(&bps_redo, redo, kRedoWarps * 32, smem_redo));
    if (bps_redo < 1) return pnbi::fail(PNB_E_UNSUPPORTED, "NNLS redo kernel does not fit on this device");
    // few voxels: 2-warp CTAs whose factor stays in shared memory; many: the 8-warp shape
    const unsigned long long few = 8ULL * bps_redo * sms * kRedoWarps;
    a.counter = ctr + 2; a.kmax = kmax_redo; a.work_list = C.redo_list[slot]; a.work_count = ctr + 1;
    a.work_min = 1; a.work_max = few;
    redo<<<(unsigned)((long long)bps_redo * sms), kRedoWarps * 32, smem_redo, stream>>>(a);
    PNBI_CUDA(cudaGetLastError());
    a.kmax = kmax; a.work_min = few + 1; a.work_max = ~0ULL;
    robust<<<(unsigned)((long long)bps * sms), kWarps * 32, smem, stream>>>(a);
    PNBI_CUDA(cudaGetLastError());
    pnbi::count_launch();
    pnbi::count_launch();
  } else {
    a.work_list = nullptr; a.work_count = nullptr; a.work_min = 0; a.work_max = ~0ULL;
    if (algorithm == 2) {
      // the compact hand-over pass of the host pipeline: same choice of shape as above, the count is
      // known on the host here
      auto redo = pnb::nnls_kernel<kRedoWarps>;
      const int kmax_redo = pick_kmax(m, n, W, kRedoWarps);
      const size_t smem_redo = pnb::nnls_smem_bytes(m, n, W, kmax_redo, kRedoWarps);
      PNBI_CUDA(cudaFuncSetAttribute(redo, cudaFuncAttributeMaxDynamicSharedMemorySize, (int)smem_redo));
      int bps_redo = 0;
      PNBI_CUDA(cudaOccupancyMaxActiveBlocksPerMultiprocessor(&bps_redo, redo, kRedoWarps * 32, smem_redo));
      if (bps_redo < 1) return pnbi::fail(PNB_E_UNSUPPORTED, "NNLS redo kernel does not fit on this device");
      if ((unsigned long long)n_vox <= 8ULL * bps_redo * sms * kRedoWarps) {
        long long grid = (long long)bps_redo * sms;
        const long long want_redo = (n_vox + kRedoWarps - 1) / kRedoWarps;
        if (want_redo < grid) grid = want_redo;
        a.counter = ctr; a.kmax = kmax_redo;
        redo<<<(unsigned)grid, kRedoWarps * 32, smem_redo, stream>>>(a);
        PNBI_CUDA(cudaGetLastError());
        pnbi::count_launch();
        return 0;
      }
    }
    long long grid = (long long)bps * sms;
    if (want < grid) grid = want;
    if (grid < 1) grid = 1;
    a.counter = ctr; a.kmax = kmax;
    robust<<<(unsigned)grid, kWarps * 32, smem, stream>>>(a);
    PNBI_CUDA(cudaGetLastError());
    pnbi::count_launch();
  }
  return 0;
}

}  // namespace

extern "C" int pnb_nnls_fit_device(const pnb_nnls_problem *p, void *cuda_stream) {
  if (int rc = check(p)) return rc;
  if (p->n_vox == 0) return 0;
  int dev = 0;
  PNBI_CUDA(cudaGetDevice(&dev));
  std::lock_guard<std::mutex> lk(g_mu[dev & 15]);
  NnlsCtx &C = g_ctx[dev & 15];
  // Device-path launches use slot 0's scratch, redo list and counters whatever stream they are given:
  // order each launch after the previous one (a no-op on the same stream), so two launches in flight
  // on different streams of one device cannot race on them.
  cudaStream_t stream = (cudaStream_t)cuda_stream;
  if (!C.dev_done) PNBI_CUDA(cudaEventCreateWithFlags(&C.dev_done, cudaEventDisableTiming));
  else PNBI_CUDA(cudaStreamWaitEvent(stream, C.dev_done, 0));
  const int rc = launch(C, p, p->basis, p->rtr_band, p->signal, p->n_vox, p->coefficients,
                        p->residual, p->status, p->iterations, p->r_squared, stream, 0);
  PNBI_CUDA(cudaEventRecord(C.dev_done, stream));
  return rc;
}

namespace {
// voxels [v0, v1) of the problem through the host pipeline of `device` (see pnb_api.cu: trf_host_range)
int nnls_host_range(const pnb_nnls_problem *p, int device, int64_t chunk_vox, size_t v0, size_t v1) {
  if (v1 <= v0) return 0;
  if (pnb_device_count() <= device || device < 0) return pnbi::fail(PNB_E_NODEVICE, "no such CUDA device");
  pnbi::DeviceScope dev_scope(device);
  PNBI_CUDA(dev_scope.error());
  std::lock_guard<std::mutex> lk(g_mu[device & 15]);
  NnlsCtx &C = g_ctx[device & 15];
  if (C.dev_done) PNBI_CUDA(cudaEventSynchronize(C.dev_done));  // a device-path launch may still use slot 0
  const int m = p->n_b, n = p->n_bins, BW = 2 * p->rtr_halfband + 1;
  const bool staged = pnbi::is_pageable(p->signal) || pnbi::is_pageable(p->coefficients);
  // default chunk: long enough that the tail of a launch (its slowest voxels) stays small; with
  // pageable caller memory the call is bound by the host copies anyway and a quarter of that keeps
  // the page-locked staging blocks (3 x chunk x 8 (n_b + n_bins + 3) bytes, ~1 s per GB to lock) small
  if (chunk_vox <= 0) chunk_vox = staged ? (1 << 16) : (1 << 18);
  if ((size_t)chunk_vox > v1 - v0) chunk_vox = (int64_t)(v1 - v0);
  const size_t Cn = (size_t)chunk_vox;
  if (!C.streams[0])
    for (auto &s : C.streams) PNBI_CUDA(cudaStreamCreateWithFlags(&s, cudaStreamNonBlocking));
  for (int s = 0; s < kSlots; s++) {
    if (int rc = grow(&C.y[s], &C.cap_y[s], Cn * m)) return rc;
    if (int rc = grow(&C.coef[s], &C.cap_coef[s], Cn * n)) return rc;
    if (int rc = grow(&C.rn[s], &C.cap_rn[s], Cn)) return rc;
    if (int rc = grow(&C.r2[s], &C.cap_r2[s], Cn)) return rc;
    if (int rc = grow(&C.st[s], &C.cap_st[s], Cn)) return rc;
    if (int rc = grow(&C.it[s], &C.cap_it[s], Cn)) return rc;
  }
  if (int rc = grow(&C.B, &C.cap_B, (size_t)m * n)) return rc;
  if (int rc = grow(&C.rtr, &C.cap_rtr, (size_t)n * BW)) return rc;
  PNBI_CUDA(cudaMemcpy(C.B, p->basis, (size_t)m * n * sizeof(double), cudaMemcpyHostToDevice));
  PNBI_CUDA(cudaMemcpy(C.rtr, p->rtr_band, (size_t)n * BW * sizeof(double), cudaMemcpyHostToDevice));
  const size_t D = sizeof(double), I = sizeof(int);
  const size_t o_y = 0, o_coef = o_y + Cn * m * D, o_rn = o_coef + Cn * n * D, o_r2 = o_rn + Cn * D;
  const size_t o_st = o_r2 + Cn * D, o_it = o_st + Cn * I, pin_bytes = o_it + Cn * I;
  if (staged)
    for (int k = 0; k < kSlots; k++) {
      C.pend_n[k] = 0;
      if (pin_bytes > C.cap_pin[k]) {
        if (C.pin[k]) PNBI_CUDA(cudaFreeHost(C.pin[k]));
        C.pin[k] = nullptr; C.cap_pin[k] = 0;
        PNBI_CUDA(cudaHostAlloc((void **)&C.pin[k], pin_bytes, cudaHostAllocDefault));
        C.cap_pin[k] = pin_bytes;
      }
    }
  auto drain = [&](int k) -> int {
    if (!C.pend_n[k]) return 0;
    PNBI_CUDA(cudaStreamSynchronize(C.streams[k]));
    const size_t st0 = C.pend_start[k], nv = C.pend_n[k];
    pnbi::parallel_memcpy(p->coefficients + st0 * n, C.pin[k] + o_coef, nv * n * D);
    std::memcpy(p->residual + st0, C.pin[k] + o_rn, nv * D);
    if (p->r_squared) std::memcpy(p->r_squared + st0, C.pin[k] + o_r2, nv * D);
    std::memcpy(p->status + st0, C.pin[k] + o_st, nv * I);
    std::memcpy(p->iterations + st0, C.pin[k] + o_it, nv * I);
    C.pend_n[k] = 0;
    return 0;
  };
  // The few voxels a chunk's fast kernel cannot certify go to ONE list for the whole call and the
  // robust pass runs once at the end, on a compact copy of their signals.  Queued behind every chunk
  // instead, the small robust launches could only start when the NEXT chunk's fast kernel (one CTA
  // per SM, all of its shared memory) had drained, which held the chunk's downloads and its slot back:
  // 537 instead of 452 ms for the 4.19 M voxel volume (profiles/r2_nnls_e2e_probe.log).
  const bool defer_ok = p->algorithm == 0 && (v1 - v0) < (size_t)0x7fffffff;
  Deferred defer{nullptr, nullptr, 0};
  if (defer_ok) {
    if (int rc = grow(&C.defer_list, &C.defer_cap, v1 - v0)) return rc;
    if (!C.defer_count) PNBI_CUDA(cudaMalloc(&C.defer_count, sizeof(unsigned long long)));
    PNBI_CUDA(cudaMemset(C.defer_count, 0, sizeof(unsigned long long)));
    defer.list = C.defer_list; defer.count = C.defer_count;
  }
  int s = 0;
  for (size_t start = v0; start < v1; start += Cn, s = (s + 1) % kSlots) {
    const size_t nv = (v1 - start < Cn) ? v1 - start : Cn;
    cudaStream_t st = C.streams[s];
    const double *src_y = p->signal + start * m;
    if (staged) {
      if (int rc = drain(s)) return rc;
      pnbi::parallel_memcpy(C.pin[s] + o_y, src_y, nv * m * D);
      src_y = reinterpret_cast<const double *>(C.pin[s] + o_y);
    }
    PNBI_CUDA(cudaMemcpyAsync(C.y[s], src_y, nv * m * D, cudaMemcpyHostToDevice, st));
    defer.base = (int)(start - v0);
    if (int rc = launch(C, p, C.B, C.rtr, C.y[s], (long long)nv, C.coef[s], C.rn[s], C.st[s], C.it[s],
                        p->r_squared ? C.r2[s] : nullptr, st, s, defer_ok ? &defer : nullptr))
      return rc;
    if (staged) {
      PNBI_CUDA(cudaMemcpyAsync(C.pin[s] + o_coef, C.coef[s], nv * n * D, cudaMemcpyDeviceToHost, st));
      PNBI_CUDA(cudaMemcpyAsync(C.pin[s] + o_rn, C.rn[s], nv * D, cudaMemcpyDeviceToHost, st));
      if (p->r_squared) PNBI_CUDA(cudaMemcpyAsync(C.pin[s] + o_r2, C.r2[s], nv * D, cudaMemcpyDeviceToHost, st));
      PNBI_CUDA(cudaMemcpyAsync(C.pin[s] + o_st, C.st[s], nv * I, cudaMemcpyDeviceToHost, st));
      PNBI_CUDA(cudaMemcpyAsync(C.pin[s] + o_it, C.it[s], nv * I, cudaMemcpyDeviceToHost, st));
      C.pend_start[s] = start; C.pend_n[s] = nv;
      continue;
    }
    PNBI_CUDA(cudaMemcpyAsync(p->coefficients + start * n, C.coef[s], nv * n * sizeof(double), cudaMemcpyDeviceToHost, st));
    PNBI_CUDA(cudaMemcpyAsync(p->residual + start, C.rn[s], nv * sizeof(double), cudaMemcpyDeviceToHost, st));
    if (p->r_squared)
      PNBI_CUDA(cudaMemcpyAsync(p->r_squared + start, C.r2[s], nv * sizeof(double), cudaMemcpyDeviceToHost, st));
    PNBI_CUDA(cudaMemcpyAsync(p->status + start, C.st[s], nv * sizeof(int), cudaMemcpyDeviceToHost, st));
    PNBI_CUDA(cudaMemcpyAsync(p->iterations + start, C.it[s], nv * sizeof(int), cudaMemcpyDeviceToHost, st));
  }
  if (staged)
    for (int k = 0; k < kSlots; k++)  // oldest first
      if (int rc = drain((s + k) % kSlots)) return rc;
  for (auto &st : C.streams) PNBI_CUDA(cudaStreamSynchronize(st));
  if (!defer_ok) return 0;
  // ---- the deferred robust pass
  unsigned long long n_redo = 0;
  PNBI_CUDA(cudaMemcpy(&n_redo, C.defer_count, sizeof(n_redo), cudaMemcpyDeviceToHost));
  if (n_redo == 0) return 0;
  if (n_redo > v1 - v0) return pnbi::fail(PNB_E_BADARG, "NNLS hand-over list overflow (internal)");
  const size_t R = (size_t)n_redo;
  std::vector<int> idx(R);
  PNBI_CUDA(cudaMemcpy(idx.data(), C.defer_list, R * I, cudaMemcpyDeviceToHost));
  // in batches of at most one chunk: slot 0's buffers are free again and hold exactly that (a weakly
  // regularised problem can hand over most of its voxels)
  const size_t batch = R < Cn ? R : Cn;
  std::vector<double> ybuf(batch * m), cbuf(batch * n), rnbuf(batch), r2buf(p->r_squared ? batch : 0);
  std::vector<int> stbuf(batch), itbuf(batch);
  cudaStream_t st0 = C.streams[0];
  for (size_t q0 = 0; q0 < R; q0 += batch) {
    const size_t nb = (R - q0 < batch) ? R - q0 : batch;
    for (size_t q = 0; q < nb; q++)
      std::memcpy(&ybuf[q * m], p->signal + (v0 + (size_t)idx[q0 + q]) * m, m * D);
    PNBI_CUDA(cudaMemcpyAsync(C.y[0], ybuf.data(), nb * m * D, cudaMemcpyHostToDevice, st0));
    if (int rc = launch(C, p, C.B, C.rtr, C.y[0], (long long)nb, C.coef[0], C.rn[0], C.st[0], C.it[0],
                        p->r_squared ? C.r2[0] : nullptr, st0, 0, nullptr, 2))
      return rc;
    PNBI_CUDA(cudaMemcpyAsync(cbuf.data(), C.coef[0], nb * n * D, cudaMemcpyDeviceToHost, st0));
    PNBI_CUDA(cudaMemcpyAsync(rnbuf.data(), C.rn[0], nb * D, cudaMemcpyDeviceToHost, st0));
    if (p->r_squared) PNBI_CUDA(cudaMemcpyAsync(r2buf.data(), C.r2[0], nb * D, cudaMemcpyDeviceToHost, st0));
    PNBI_CUDA(cudaMemcpyAsync(stbuf.data(), C.st[0], nb * I, cudaMemcpyDeviceToHost, st0));
    PNBI_CUDA(cudaMemcpyAsync(itbuf.data(), C.it[0], nb * I, cudaMemcpyDeviceToHost, st0));
    PNBI_CUDA(cudaStreamSynchronize(st0));
    for (size_t q = 0; q < nb; q++) {
      const size_t v = v0 + (size_t)idx[q0 + q];
      std::memcpy(p->coefficients + v * n, &cbuf[q * n], n * D);
      p->residual[v] = rnbuf[q];
      if (p->r_squared) p->r_squared[v] = r2buf[q];
      p->status[v] = stbuf[q];
      p->iterations[v] = itbuf[q];
    }
  }
  C.last_redo = C.defer_count;  // what pnb_nnls_last_redo_count reports for this call
  return 0;
}
}  // namespace

extern "C" int pnb_nnls_fit_host(const pnb_nnls_problem *p, int device, int64_t chunk_vox) {
  if (int rc = check(p)) return rc;
  return nnls_host_range(p, device, chunk_vox, 0, (size_t)p->n_vox);
}

extern "C" int pnb_nnls_fit_host_multi(const pnb_nnls_problem *p, const int32_t *devices, int32_t n_devices,
                                       int64_t chunk_vox) {
  if (int rc = check(p)) return rc;
  if (n_devices < 1 || n_devices > 16) return pnbi::fail(PNB_E_BADARG, "n_devices must be in [1, 16]");
  const size_t NV = (size_t)p->n_vox;
  std::vector<int> rcs(n_devices, 0);
  std::vector<std::string> errs(n_devices);
  std::vector<std::thread> workers;
  const size_t base = NV / n_devices, rem = NV % n_devices;
  size_t start = 0;
  for (int i = 0; i < n_devices; i++) {
    const size_t stop = start + base + ((size_t)i < rem ? 1 : 0);
    const int dev = devices ? devices[i] : i;
    workers.emplace_back([=, &rcs, &errs] {
      rcs[i] = nnls_host_range(p, dev, chunk_vox, start, stop);
      if (rcs[i]) errs[i] = pnb_last_error();
    });
    start = stop;
  }
  for (auto &w : workers) w.join();
  for (int i = 0; i < n_devices; i++)
    if (rcs[i]) return pnbi::fail(rcs[i], "device " + std::to_string(devices ? devices[i] : i) + ": " + errs[i]);
  return 0;
}

extern "C" int pnb_sizeof_nnls_problem(void) { return (int)sizeof(pnb_nnls_problem); }

extern "C" int pnb_nnls_dual_gemm_device(int32_t n_b, int32_t n_bins, int64_t n_vox, const double *basis,
                                         const double *signal, double *h0, void *cuda_stream) {
  if (n_b < 1 || n_b > 32) return pnbi::fail(PNB_E_UNSUPPORTED, "the tensor-core GEMM is built for n_b <= 32");
  if (n_bins < 1 || n_bins > 1024 || n_vox < 0) return pnbi::fail(PNB_E_BADARG, "bad sizes");
  if (n_vox == 0) return 0;
  if (!basis || !signal || !h0) return pnbi::fail(PNB_E_BADARG, "null array pointer");
  return h0_gemm(pick_mt(n_b), signal, basis, h0, n_vox, n_b, n_bins, (cudaStream_t)cuda_stream);
}

extern "C" int64_t pnb_nnls_last_redo_count(int device) {
  if (device < 0 || device > 15) return -1;
  std::lock_guard<std::mutex> lk(g_mu[device]);
  NnlsCtx &C = g_ctx[device];
  if (!C.last_redo) return 0;
  unsigned long long v = 0;
  pnbi::DeviceScope dev_scope(device);
  if (dev_scope.error() != cudaSuccess) return -1;
  if (cudaMemcpy(&v, C.last_redo, sizeof(v), cudaMemcpyDeviceToHost) != cudaSuccess) return -1;
  return (int64_t)v;
}
