// Persistent one-lane-per-voxel TRF kernel.
//
// Execution model (B200, FP64 CUDA-core bound — there is no FP64 tcgen05 path):
//   * grid = resident CTAs per SM x 148 SMs, every thread is a lane that pulls
//     voxel indices from a global counter (warp-aggregated atomicAdd) until the
//     volume is exhausted.  A lane whose voxel converges refills immediately,
//     so the 4..30 iterations different voxels need do not idle the other 31
//     lanes of the warp the way a one-voxel-per-thread grid would.
//   * each pass of the main loop is the same straight-line sequence for every
//     lane — [refill] -> [prologue + trust-region solve + step selection] ->
//     [model / Jacobian / normal-equation accumulation over the b-values] ->
//     [accept / reject bookkeeping] -> [finalise] — guarded by per-lane
//     predicates, so lanes at different iterations of different voxels still
//     execute the FP64-heavy middle part convergently.
//   * shared memory holds the b-value vector (broadcast reads) and, per lane,
//     its voxel's signal and bounds in [row][lane] order (conflict-free 8-byte
//     accesses); the signal row of a voxel is fetched from the voxel-major
//     (n_vox, n_b) array with 16-byte loads.  All solver state is in registers.
#pragma once
#include <cuda_runtime.h>

#include "pnb_trf_core.cuh"

namespace pnb {

struct TrfDeviceArgs {
  int n_b;
  long long n_vox;
  const double *b;       // (n_b)
  const double *y;       // (n_vox, n_b) voxel-major
  // p0 / lb / ub over ALL model parameters: element (k, v) at ptr[k * row_stride + v * vox_stride]
  const double *p0, *lb, *ub;
  long long p0_row_stride, p0_vox_stride;
  long long bd_row_stride, bd_vox_stride;
  TrfOptions opt;
  // outputs
  double *params;        // (NP, n_vox) parameter-major, like the reference's popt
  double *cov;           // (n_vox, n_free, n_free) or nullptr
  int *status;           // (n_vox) SciPy status, <0: input rejected
  int *nfev;             // (n_vox)
  int *njev;             // (n_vox) or nullptr
  double *cost;          // (n_vox) or nullptr
  unsigned long long *counter;  // work counter, zeroed before launch
};

template <class M, int BLOCK>
__global__ void __launch_bounds__(BLOCK) trf_kernel(const TrfDeviceArgs a) {
  constexpr int N = M::NP;
  extern __shared__ double smem[];
  const int m = a.n_b;
  const int tid = threadIdx.x;
  double *b_s = smem;                       // [m]
  double *y_s = b_s + ((m + 1) & ~1);       // [m][BLOCK]
  double *lb_s = y_s + (size_t)m * BLOCK;   // [N][BLOCK]
  double *ub_s = lb_s + N * BLOCK;          // [N][BLOCK]
  for (int i = tid; i < m; i += BLOCK) b_s[i] = a.b[i];
  __syncthreads();
  double *my_y = y_s + tid;
  const double *my_lb = lb_s + tid;
  const double *my_ub = ub_s + tid;
  const TrfOptions &O = a.opt;
  const unsigned lane = tid & 31;

  TrfLane<M> S;
  long long vox = -1;
  bool first_eval = false;
  auto yb = [&](int r, double &yv, double &bv) { yv = my_y[r * BLOCK]; bv = b_s[r]; };

  for (;;) {
    bool finished = false;
    bool do_eval = false;
    // ---- refill -----------------------------------------------------
    if (vox < 0) {
      const unsigned mask = __activemask();
      const int leader = __ffs(mask) - 1;
      unsigned long long base = 0;
      if ((int)lane == leader) base = atomicAdd(a.counter, (unsigned long long)__popc(mask));
      base = __shfl_sync(mask, base, leader);
      vox = (long long)(base + __popc(mask & ((1u << lane) - 1)));
      if (vox >= a.n_vox) break;
      // signal row -> shared (16-byte loads when the row is 16-byte aligned)
      const double *yrow = a.y + vox * m;
      bool yfin = true;
      if ((m & 1) == 0 && ((reinterpret_cast<uintptr_t>(yrow) & 15) == 0)) {
        const double2 *y2 = reinterpret_cast<const double2 *>(yrow);
        for (int r = 0; r < m / 2; r++) {
          const double2 v = __ldg(y2 + r);
          my_y[(2 * r) * BLOCK] = v.x;
          my_y[(2 * r + 1) * BLOCK] = v.y;
          yfin = yfin && finite_d(v.x) && finite_d(v.y);
        }
      } else {
        for (int r = 0; r < m; r++) {
          const double v = __ldg(yrow + r);
          my_y[r * BLOCK] = v;
          yfin = yfin && finite_d(v);
        }
      }
      double p0v[N];
#pragma unroll
      for (int k = 0; k < N; k++) {
        p0v[k] = __ldg(a.p0 + k * a.p0_row_stride + vox * a.p0_vox_stride);
        lb_s[k * BLOCK + tid] = __ldg(a.lb + k * a.bd_row_stride + vox * a.bd_vox_stride);
        ub_s[k * BLOCK + tid] = __ldg(a.ub + k * a.bd_row_stride + vox * a.bd_vox_stride);
      }
      if (trf_begin<M>(S, O, p0v, my_lb, my_ub, BLOCK, yfin)) {
        first_eval = true;
        do_eval = true;
#pragma unroll
        for (int k = 0; k < N; k++) S.x_new[k] = S.x[k];
      } else {
        finished = true;
      }
    } else {
      // ---- prepare a trial step ---------------------------------------
      bool go = true;
      if (S.need_prologue) {
        go = trf_prologue<M>(S, O, my_lb, my_ub, BLOCK);
        S.need_prologue = false;
      }
      if (go) {
        double p_h[N];
        trf_solve_tr<M>(S, p_h);
        trf_select_step<M>(S, p_h, my_lb, my_ub, BLOCK, O.frozen);
        do_eval = true;
      } else {
        finished = true;
      }
    }
    // ---- model, Jacobian and normal equations at x_new ------------------
    if (do_eval) {
      double c, g[N], A[N][N];
      trf_evaluate<M>(S.x_new, O, m, yb, my_lb, my_ub, BLOCK, c, g, A);
      if (first_eval) {
        first_eval = false;
        if (!trf_after_first_eval<M>(S, O, c, g, A, my_lb, my_ub, BLOCK)) finished = true;
      } else {
        S.need_prologue = trf_after_trial<M>(S, O, c, g, A);
      }
    }
    // ---- write results ------------------------------------------------
    if (finished) {
      const bool ok = S.status > 0;
      int n_free = 0;
#pragma unroll
      for (int k = 0; k < N; k++) {
        n_free += ((O.frozen >> k) & 1u) ? 0 : 1;
        const double v = ok ? S.x[k] : __ldg(a.p0 + k * a.p0_row_stride + vox * a.p0_vox_stride);
        a.params[(long long)k * a.n_vox + vox] = v;
      }
      a.status[vox] = S.status;
      a.nfev[vox] = S.nfev;
      if (a.njev) a.njev[vox] = S.njev;
      if (a.cost) a.cost[vox] = ok || S.status == kStMaxNfev ? S.cost : nan("");
      if (a.cov) {
        double *cv = a.cov + vox * (long long)(n_free * n_free);
        if (ok) {
          trf_covariance<M>(S, O, m, cv);
        } else {
          const double qnan = nan("");
          for (int i = 0; i < n_free * n_free; i++) cv[i] = qnan;
        }
      }
      vox = -1;
    }
  }
}

template <class M, int BLOCK> size_t trf_smem_bytes(int n_b) {
  return sizeof(double) * (((n_b + 1) & ~1) + (size_t)n_b * BLOCK + 2 * M::NP * BLOCK);
}

// Launch configuration: persistent grid, as many CTAs as are resident.
template <class M, int BLOCK> cudaError_t trf_launch(const TrfDeviceArgs &a, cudaStream_t stream) {
  const size_t smem = trf_smem_bytes<M, BLOCK>(a.n_b);
  auto kern = trf_kernel<M, BLOCK>;
  static int blocks_per_sm_cache = -1;
  static size_t smem_cache = 0;
  static int sm_count = 0;
  cudaError_t err;
  if (smem > 48 * 1024) {
    err = cudaFuncSetAttribute(kern, cudaFuncAttributeMaxDynamicSharedMemorySize, (int)smem);
    if (err != cudaSuccess) return err;
  }
  if (blocks_per_sm_cache < 0 || smem_cache != smem) {
    int dev = 0;
    err = cudaGetDevice(&dev);
    if (err != cudaSuccess) return err;
    err = cudaDeviceGetAttribute(&sm_count, cudaDevAttrMultiProcessorCount, dev);
    if (err != cudaSuccess) return err;
    int bps = 0;
    err = cudaOccupancyMaxActiveBlocksPerMultiprocessor(&bps, kern, BLOCK, smem);
    if (err != cudaSuccess) return err;
    if (bps < 1) return cudaErrorInvalidConfiguration;
    blocks_per_sm_cache = bps;
    smem_cache = smem;
  }
  long long want = (a.n_vox + BLOCK - 1) / BLOCK;
  long long grid = (long long)blocks_per_sm_cache * sm_count;
  if (want < grid) grid = want;
  if (grid < 1) grid = 1;
  err = cudaMemsetAsync(a.counter, 0, sizeof(unsigned long long), stream);
  if (err != cudaSuccess) return err;
  kern<<<(unsigned)grid, BLOCK, smem, stream>>>(a);
  return cudaGetLastError();
}

}  // namespace pnb
