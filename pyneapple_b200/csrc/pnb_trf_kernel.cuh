// Persistent one-lane-per-voxel TRF kernel.
//
// Execution model (B200, FP64 CUDA-core bound — there is no FP64 tcgen05 path):
//   * grid = resident CTAs per SM x 148 SMs; every thread is a lane that works
//     through voxels until the volume is exhausted.  A lane whose voxel
//     converges starts its next voxel on the following pass, so the 4..30
//     iterations different voxels need do not idle the other lanes of the warp.
//   * voxel indices come from a global counter, claimed 64 at a time per warp
//     (one atomicAdd by lane 0, issued one range ahead of need so its latency
//     is never waited for) and handed to lanes with ballot / popc arithmetic.
//   * a lane always holds one voxel *ahead*: as soon as it starts voxel v it
//     claims the next index and pulls that voxel's signal row (and per-voxel
//     p0 / bounds) into the other half of a double-buffered shared-memory
//     column with cp.async, ~8 passes before the data is used, so HBM/L2
//     latency is hidden even at 8 warps per SM.
//   * every pass of the main loop is the same sequence for all 32 lanes —
//     [claim/prefetch] [start] [prologue + trust-region solve + step selection]
//     [model / Jacobian / normal equations over the b-values] [accept/reject]
//     [write results] — with a warp-wide vote at the top of the pass, which
//     re-converges the warp: without it the lanes drift apart after the first
//     divergent branch and the FP64-heavy evaluation runs at half width.
//   * shared memory: the b-value vector (broadcast reads) and, per lane, signal
//     and bounds columns in [row][lane] order (conflict-free 8-byte accesses).
//     All solver state is in registers.
#pragma once
#include <cuda_runtime.h>

#include "pnb_dogbox_core.cuh"
#include "pnb_lm_core.cuh"

namespace pnb {

struct TrfDeviceArgs {
  int n_b;
  long long n_vox;
  const double *b;       // (n_b)
  const double *y;       // (n_vox, n_b) voxel-major
  const double *w = nullptr;  // (n_b) 1 / sigma of curve_fit, or nullptr (EXTRAS kernels only)
  // p0 / lb / ub over ALL model parameters: element (k, v) at ptr[k * row_stride + v * vox_stride]
  const double *p0, *lb, *ub;
  long long p0_row_stride, p0_vox_stride;
  long long bd_row_stride, bd_vox_stride;
  TrfOptions opt;
  // outputs
  double *params;        // (NP, n_vox) parameter-major, like the reference's popt
  double *cov;           // (n_vox, n_free, n_free) or nullptr
  int defer_cov = 0;     // host pipeline: leave the slots parked, one trf_cov_launch(status = nullptr) follows
  int *status;           // (n_vox) SciPy status, <0: input rejected
  int *nfev;             // (n_vox)
  int *njev;             // (n_vox) or nullptr
  double *cost;          // (n_vox) or nullptr
  double *r2;            // (n_vox) or nullptr: 1 - SS_res / SS_tot at the returned parameters
  unsigned long long *counter;  // work counter, zeroed before launch
  unsigned long long *n_failed; // += voxels that end with status <= 0 (failures are rare), or nullptr
};

__device__ __forceinline__ void cp_async8(double *smem_dst, const double *gsrc) {
  const unsigned s = (unsigned)__cvta_generic_to_shared(smem_dst);
  asm volatile("cp.async.ca.shared.global [%0], [%1], 8;\n" ::"r"(s), "l"(gsrc) : "memory");
}
__device__ __forceinline__ void cp_async_commit() { asm volatile("cp.async.commit_group;\n" ::: "memory"); }
__device__ __forceinline__ void cp_async_wait_all() { asm volatile("cp.async.wait_group 0;\n" ::: "memory"); }

constexpr int kTrfClaim = 64;  // voxel indices claimed per warp-level atomicAdd
// Batched write-out of converged lanes (see `pending` in trf_kernel).  A finished lane waits at most
// PNB_TRF_FINISH_WAIT passes (the default of pnb_trf_problem.finish_wait / TrfOptions::finish_wait); the gate also opens when PNB_TRF_FINISH_BATCH lanes are ready (32 = never
// by count: measured, a count trigger below 32 splits a warp into groups that drift apart and pay the
// start / finish code separately - profiles/r2_trf_finish_batch.log) or nothing else runs.
// BATCH 1 + WAIT 0 = every lane writes at once (the old behaviour).
#ifndef PNB_TRF_FINISH_BATCH
#define PNB_TRF_FINISH_BATCH 32
#endif
#ifndef PNB_TRF_FINISH_WAIT
#define PNB_TRF_FINISH_WAIT 3
#endif
constexpr int kTrfFinishBatch = PNB_TRF_FINISH_BATCH;
constexpr int kTrfFinishWait = PNB_TRF_FINISH_WAIT;

#ifndef PNB_TRF_MINBLOCKS
#define PNB_TRF_MINBLOCKS 1
#endif

// EXTRAS: the instantiation that honours curve_fit's `sigma` (a.w) and least_squares' robust `loss`
// (trf_evaluate<M, true>); the plain kernels do not carry that code.
template <class M, int BLOCK, int METHOD = 0, bool EXTRAS = false>
__global__ void __launch_bounds__(BLOCK, PNB_TRF_MINBLOCKS) trf_kernel(const TrfDeviceArgs a) {
  constexpr int N = M::NP;
  constexpr unsigned FULL = 0xffffffffu;
  extern __shared__ double smem[];
  const int m = a.n_b;
  const int tid = threadIdx.x;
  const unsigned lane = tid & 31;
  const bool pv_p0 = a.p0_vox_stride != 0;
  const bool pv_bd = a.bd_vox_stride != 0;
  // layout: b[m'] | (EXTRAS: w[m']) | y[2][m][BLOCK] | lb[2][N][BLOCK] | ub[2][N][BLOCK] | p0[2][N][BLOCK]
  double *b_s = smem;
  double *w_s = b_s + ((m + 1) & ~1);
  double *y_s = w_s + (EXTRAS ? ((m + 1) & ~1) : 0);
  double *lb_s = y_s + (size_t)2 * m * BLOCK;
  double *ub_s = lb_s + 2 * N * BLOCK;
  double *p0_s = ub_s + 2 * N * BLOCK;
  exp_tab_init(tid, BLOCK);
  for (int i = tid; i < m; i += BLOCK) b_s[i] = a.b[i];
  if (EXTRAS)
    for (int i = tid; i < m; i += BLOCK) w_s[i] = a.w ? a.w[i] : 1.0;
  if (!pv_bd) {
#pragma unroll
    for (int k = 0; k < N; k++) {
      const double l = a.lb[k * a.bd_row_stride], u = a.ub[k * a.bd_row_stride];
      lb_s[k * BLOCK + tid] = l; lb_s[(N + k) * BLOCK + tid] = l;
      ub_s[k * BLOCK + tid] = u; ub_s[(N + k) * BLOCK + tid] = u;
    }
  }
  if (!pv_p0) {
#pragma unroll
    for (int k = 0; k < N; k++) {
      const double v = a.p0[k * a.p0_row_stride];
      p0_s[k * BLOCK + tid] = v; p0_s[(N + k) * BLOCK + tid] = v;
    }
  }
  __syncthreads();
  const TrfOptions &O = a.opt;

  // warp-level queue of claimed voxel indices: two contiguous ranges
  long long qa = 0, qb = 0;
  int na = 0, nb = 0;
  auto claim_range = [&](long long &base, int &cnt) {
    unsigned long long v = 0;
    if (lane == 0) v = atomicAdd(a.counter, (unsigned long long)kTrfClaim);
    v = __shfl_sync(FULL, v, 0);
    base = (long long)v;
    long long left = a.n_vox - base;
    cnt = left <= 0 ? 0 : (left < kTrfClaim ? (int)left : kTrfClaim);
  };
  claim_range(qa, na);
  claim_range(qb, nb);

  TrfLane<M> S;
  DogboxLane<M> DB;  // METHOD == 1 only
  LmLane<M> LM;      // METHOD == 2 only
  long long cur = -1, nxt = -1;
  int buf = 0, nxt_buf = 0;  // which half of the double buffers holds `cur` / receives `nxt`
  bool first_eval = false;
  // A lane whose voxel has converged does not write its results at once: the once-per-voxel code
  // (result stores, R^2, and on the next pass the preamble of the lane's next voxel) ran with the
  // ~4 lanes of 32 that happened to finish in a pass and took ~30 % of the kernel's warp samples
  // (profiles/r2_trf_hotspots.txt).  Finished lanes wait until one of them has waited kTrfFinishWait
  // passes (or nothing else is running) and go through that code together; they also start their
  // next voxels together, and voxels that start together mostly finish within a pass or two of each
  // other, so a warp settles into near-synchronous batches with a bounded wait for stragglers.
  bool pending = false;
  int waited = 0;

  for (;;) {
    // ---- claim + prefetch the voxel after the current one ----------------
    const bool want = nxt < 0;
    const unsigned wmask = __ballot_sync(FULL, want);
    if (wmask && (na + nb) > 0) {
      const int k = __popc(wmask);
      const int i = __popc(wmask & ((1u << lane) - 1));
      if (want) {
        if (i < na) nxt = qa + i;
        else if (i - na < nb) nxt = qb + (i - na);
      }
      if (k >= na) {  // range A used up: B becomes A, claim a new B (needed a few passes from now)
        const int used_b = k - na;
        qa = qb + used_b;
        na = nb > used_b ? nb - used_b : 0;
        claim_range(qb, nb);
      } else {
        qa += k;
        na -= k;
      }
      if (want && nxt >= 0) {
        nxt_buf = (cur >= 0) ? (buf ^ 1) : buf;
        const double *yrow = a.y + nxt * m;
        double *ycol = y_s + (size_t)nxt_buf * m * BLOCK + tid;
        for (int r = 0; r < m; r++) cp_async8(ycol + r * BLOCK, yrow + r);
        if (pv_bd) {
#pragma unroll
          for (int kk = 0; kk < N; kk++) {
            cp_async8(lb_s + (nxt_buf * N + kk) * BLOCK + tid, a.lb + kk * a.bd_row_stride + nxt);
            cp_async8(ub_s + (nxt_buf * N + kk) * BLOCK + tid, a.ub + kk * a.bd_row_stride + nxt);
          }
        }
        if (pv_p0) {
#pragma unroll
          for (int kk = 0; kk < N; kk++)
            cp_async8(p0_s + (nxt_buf * N + kk) * BLOCK + tid, a.p0 + kk * a.p0_row_stride + nxt);
        }
        cp_async_commit();
      }
    }
    // ---- every lane out of work: done --------------------------------------
    if (__all_sync(FULL, cur < 0 && nxt < 0)) break;

    bool finished = false;
    bool do_eval = false;
    bool starting = false;
    if (cur < 0 && nxt >= 0) {
      cp_async_wait_all();  // issued ~a whole voxel ago, except for the very first voxel
      cur = nxt;
      nxt = -1;
      buf = nxt_buf;
      starting = true;
    }
    const double *my_y = y_s + (size_t)buf * m * BLOCK + tid;
    const double *my_lb = lb_s + buf * N * BLOCK + tid;
    const double *my_ub = ub_s + buf * N * BLOCK + tid;
    const double *my_p0 = p0_s + buf * N * BLOCK + tid;
    auto yb = [&](int r, double &yv, double &bv) { yv = my_y[r * BLOCK]; bv = b_s[r]; };

    if (starting) {
      // ---- least_squares() preamble -------------------------------------------
      bool yfin = true;
      for (int r = 0; r < m; r++) yfin = yfin && finite_d(my_y[r * BLOCK]);
      double p0v[N];
#pragma unroll
      for (int k = 0; k < N; k++) p0v[k] = my_p0[k * BLOCK];
      if (trf_begin<M>(S, O, p0v, my_lb, my_ub, BLOCK, yfin)) {
        first_eval = true;
        do_eval = true;
#pragma unroll
        for (int k = 0; k < N; k++) S.x_new[k] = S.x[k];
      } else {
        finished = true;
      }
    } else if (cur >= 0 && !pending) {
      // ---- prepare a trial step (the prologue of this outer iteration ran at the end of the
      // previous pass, see below) -------------------------------------------------
      if (METHOD == 2) {
        lm_trial<M>(S, LM, O);
      } else if (METHOD == 1) {
        dbx_trial<M>(S, DB, O, my_lb, my_ub, BLOCK);
      } else {
        double p_h[N];
        trf_solve_tr<M>(S, p_h);
        trf_select_step<M>(S, p_h, my_lb, my_ub, BLOCK, O.frozen);
      }
      do_eval = true;
    }
    __syncwarp();
    // ---- model, Jacobian and normal equations at x_new --------------------------
    if (do_eval) {
      double c, g[N], A[N][N];
      trf_evaluate<M, EXTRAS>(S.x_new, O, m, yb, my_lb, my_ub, BLOCK, c, g, A, EXTRAS ? w_s : nullptr);
      if (first_eval) {
        first_eval = false;
        const bool ok0 = (METHOD == 2) ? lm_after_first_eval<M>(S, LM, O, c, g, A)
                         : (METHOD == 1) ? dbx_after_first_eval<M>(S, DB, O, c, g, A, my_lb, my_ub, BLOCK)
                                         : trf_after_first_eval<M>(S, O, c, g, A, my_lb, my_ub, BLOCK);
        if (!ok0) finished = true;
      } else {
        S.need_prologue = (METHOD == 2) ? lm_after_trial<M>(S, LM, O, c, g, A)
                          : (METHOD == 1) ? dbx_after_trial<M>(S, DB, O, c, g, A, my_lb, my_ub, BLOCK)
                                          : trf_after_trial<M>(S, O, c, g, A);
      }
      // The head of the next outer iteration (scaling vector, termination tests, trust-region model)
      // runs in THIS pass, right after the step has been accepted: a converged voxel is recognised in
      // the pass of its last evaluation instead of spending one more pass to find out.
      if (!finished && S.need_prologue) {
        const bool go = (METHOD == 2) ? lm_prologue<M>(S, LM, O)
                        : (METHOD == 1) ? dbx_prologue<M>(S, DB, O) : trf_prologue<M>(S, O, my_lb, my_ub, BLOCK);
        S.need_prologue = false;
        if (!go) finished = true;
      }
    }
    __syncwarp();
    // ---- write results ------------------------------------------------------------
    if (pending) waited++;
    pending = pending || finished;
    const unsigned pend_mask = __ballot_sync(FULL, pending);
    const bool running_any = __any_sync(FULL, cur >= 0 && !pending);
    const bool overdue = __any_sync(FULL, pending && waited >= O.finish_wait);
    const bool open = __popc(pend_mask) >= kTrfFinishBatch || overdue || !running_any;
    if (pending && open) {
      pending = false;
      waited = 0;
      const long long vox = cur;
      const bool ok = S.status > 0;
      int n_free = 0;
#pragma unroll
      for (int k = 0; k < N; k++) {
        n_free += ((O.frozen >> k) & 1u) ? 0 : 1;
        a.params[(long long)k * a.n_vox + vox] = ok ? S.x[k] : my_p0[k * BLOCK];
      }
      a.status[vox] = S.status;
      if (S.status <= 0 && a.n_failed) atomicAdd(a.n_failed, 1ULL);
      a.nfev[vox] = S.nfev;
      if (a.njev) a.njev[vox] = S.njev;
      if (a.cost) a.cost[vox] = (ok || S.status == kStMaxNfev) ? S.cost : nan("");
      if (a.r2) {
        // fitters/base.py:142-186: R^2 = 1 - SS_res / SS_tot, NaN when the signal is constant
        double mean = 0.0;
        for (int r = 0; r < m; r++) mean += my_y[r * BLOCK];
        mean /= (double)m;
        double ss_tot = 0.0;
        for (int r = 0; r < m; r++) { const double d = my_y[r * BLOCK] - mean; ss_tot += d * d; }
        double ss_res = 2.0 * S.cost;
        if (!ok || EXTRAS) {
          // the returned parameters are p0: evaluate the model there.  EXTRAS: S.cost is the weighted /
          // robust cost, R^2 wants the plain residuals at the solution
          double p0v[N], c0, g0[N], A0[N][N];
#pragma unroll
          for (int k = 0; k < N; k++) p0v[k] = ok ? S.x[k] : my_p0[k * BLOCK];
          TrfOptions O2 = O;
          O2.jac_mode = 0;
          trf_evaluate<M>(p0v, O2, m, yb, my_lb, my_ub, BLOCK, c0, g0, A0);
          ss_res = 2.0 * c0;
        }
        a.r2[vox] = (ss_tot > 0.0) ? 1.0 - ss_res / ss_tot : nan("");
      }
      if (a.cov) {
        double *cv = a.cov + vox * (long long)(n_free * n_free);
        if (!ok) {
          const double qnan = nan("");
          for (int i = 0; i < n_free * n_free; i++) cv[i] = qnan;
        } else if (n_free < 2) {
          trf_covariance<M>(S, O, m, cv);
        } else {
          // The few lanes that finish in a pass would run the LDL^T / pseudo-inverse at ~4 of 32
          // lanes (15 % of the kernel for 3 % of the arithmetic).  They only park J^T J over the
          // free parameters (packed lower triangle) and the cost in the voxel's covariance slot;
          // cov_kernel (below) turns every slot into pinv(J^T J) * 2 cost / (m - n) with all
          // lanes busy.  n (n + 1) / 2 + 1 <= n^2 for n >= 2.
          int q = 0;
#pragma unroll
          for (int i = 0; i < N; i++) {
            if ((O.frozen >> i) & 1u) continue;
#pragma unroll
            for (int j = 0; j <= i; j++) {
              if ((O.frozen >> j) & 1u) continue;
              cv[q++] = S.A[i][j];
            }
          }
          cv[q] = S.cost;
        }
      }
      cur = -1;
    }
  }
}

// curve_fit's covariance from the parked normal matrix: pinv(J^T J) * 2 cost / (m - n), inf when
// m <= n (same rules as trf_covariance).  One thread per voxel, every lane busy.
template <int NF> struct CovTile { static constexpr int TV = NF <= 4 ? 256 : 64; };  // voxels per CTA (<= 48 KB tile)

template <int NF>
__global__ void __launch_bounds__(CovTile<NF>::TV) cov_kernel(long long n_vox, int m, const int *__restrict__ status,
                                                              double *__restrict__ cov, int absolute_sigma) {
  constexpr int TV = CovTile<NF>::TV;
  // a tile of TV voxel slots goes through shared memory so that global memory is read and written
  // with consecutive lanes on consecutive doubles (each thread's slot is NF^2 doubles long)
  constexpr int SL = NF * NF, LDS = SL | 1;  // odd row stride: conflict-free per-thread rows
  __shared__ double tile[TV * LDS];
  const long long v0 = (long long)blockIdx.x * TV;
  const int nv = (int)((n_vox - v0) < TV ? (n_vox - v0) : TV);
  double *base = cov + v0 * SL;
  for (int i = threadIdx.x; i < nv * SL; i += TV) tile[(i / SL) * LDS + (i % SL)] = base[i];
  __syncthreads();
  // status == nullptr (one pass over a whole range after the fact): a failed voxel's slot is all NaN
  const bool live = (int)threadIdx.x < nv &&
                    (status ? status[v0 + threadIdx.x] > 0 : tile[threadIdx.x * LDS] == tile[threadIdx.x * LDS]);
  if (live) {
    double *cv = tile + threadIdx.x * LDS;
    double A[NF][NF], C[NF][NF], L[NF][NF], dinv[NF];
    int q = 0;
#pragma unroll
    for (int i = 0; i < NF; i++)
#pragma unroll
      for (int j = 0; j <= i; j++) A[i][j] = cv[q++];
    const double cost = cv[q];
    if (m <= NF && !absolute_sigma) {
#pragma unroll
      for (int i = 0; i < NF; i++)
#pragma unroll
        for (int j = 0; j < NF; j++) C[i][j] = kInf;
    } else {
      const double s_sq = absolute_sigma ? 1.0 : 2.0 * cost / (double)(m - NF);
      if (ldlt<NF>(A, 0.0, L, dinv)) {
#pragma unroll
        for (int c = 0; c < NF; c++) {
          double e[NF], x[NF];
#pragma unroll
          for (int i = 0; i < NF; i++) e[i] = (i == c) ? 1.0 : 0.0;
          ldlt_solve<NF>(L, dinv, e, x);
#pragma unroll
          for (int i = 0; i < NF; i++) C[i][c] = x[i] * s_sq;
        }
      } else {
        // rank deficient: Moore-Penrose inverse, singular values s <= eps * max(m, n) * s_max dropped
        double Asym[NF][NF], V[NF][NF], w[NF];
#pragma unroll
        for (int i = 0; i < NF; i++)
#pragma unroll
          for (int j = 0; j < NF; j++) Asym[i][j] = (j <= i) ? A[i][j] : A[j][i];
        jacobi_eig<NF>(Asym, V, w);
        double wmax = 0.0;
#pragma unroll
        for (int i = 0; i < NF; i++) wmax = dmax(wmax, w[i]);
        const double thr = kEps * (double)(m > NF ? m : NF);
        const double wthr = thr * thr * wmax, noise = 4.0 * NF * kEps * wmax;
#pragma unroll
        for (int i = 0; i < NF; i++)
#pragma unroll
          for (int j = 0; j < NF; j++) {
            double t = 0.0;
#pragma unroll
            for (int kk = 0; kk < NF; kk++)
              if (w[kk] > wthr && w[kk] > noise) t += V[i][kk] * V[j][kk] / w[kk];
            C[i][j] = t * s_sq;
          }
      }
    }
#pragma unroll
    for (int i = 0; i < NF; i++)
#pragma unroll
      for (int j = 0; j < NF; j++) cv[i * NF + j] = C[i][j];
  }
  __syncthreads();
  for (int i = threadIdx.x; i < nv * SL; i += TV) base[i] = tile[(i / SL) * LDS + (i % SL)];
}

// defined once in pnb_api.cu (six small instantiations)
cudaError_t trf_cov_launch(int n_free, long long n_vox, int m, const int *status, double *cov, cudaStream_t stream,
                           int absolute_sigma = 0);

template <class M, int BLOCK> size_t trf_smem_bytes(int n_b, bool extras = false) {
  return sizeof(double) * ((extras ? 2 : 1) * ((n_b + 1) & ~1) + (size_t)2 * n_b * BLOCK + 6 * M::NP * BLOCK);
}

// Launch configuration: persistent grid, as many CTAs as are resident.
template <class M, int BLOCK, int METHOD = 0, bool EXTRAS = false>
cudaError_t trf_launch(const TrfDeviceArgs &a, cudaStream_t stream) {
  const size_t smem = trf_smem_bytes<M, BLOCK>(a.n_b, EXTRAS);
  auto kern = trf_kernel<M, BLOCK, METHOD, EXTRAS>;
  // per thread: the multi-GPU host entry drives one device from each of its threads
  static thread_local int blocks_per_sm_cache = -1;
  static thread_local size_t smem_cache = 0;
  static thread_local int sm_count = 0;
  cudaError_t err;
  if (smem > 200 * 1024) return cudaErrorInvalidValue;  // n_b too large for this block size
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
  // one voxel in flight + one prefetched per lane: do not launch more lanes than that needs
  long long want = (a.n_vox + BLOCK - 1) / BLOCK;
  long long grid = (long long)blocks_per_sm_cache * sm_count;
  if (want < grid) grid = want;
  if (grid < 1) grid = 1;
  err = cudaMemsetAsync(a.counter, 0, sizeof(unsigned long long), stream);
  if (err != cudaSuccess) return err;
  kern<<<(unsigned)grid, BLOCK, smem, stream>>>(a);
  err = cudaGetLastError();
  if (err != cudaSuccess || !a.cov || a.defer_cov) return err;
  int n_free = 0;
  for (int i = 0; i < M::NP; i++) n_free += ((a.opt.frozen >> i) & 1u) ? 0 : 1;
  if (n_free < 2) return err;
  return trf_cov_launch(n_free, a.n_vox, a.n_b, a.status, a.cov, stream, a.opt.absolute_sigma);
}

}  // namespace pnb
