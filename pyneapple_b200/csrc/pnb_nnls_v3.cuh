// Third generation of the fast NNLS path (n_bins <= 256, n_b <= 32, half-bandwidth <= 4).
//
// Same mathematics as pnb_nnls_fast.cuh — Lawson-Hanson with the active block
// carried as its explicit inverse H = (G_PP)^-1, bordering update when a bin
// enters, rank-one downdate when one leaves, polish by iterative refinement and
// self-certification with hand-over to the robust Cholesky kernel — rewritten
// around what the ncu source profile of that kernel showed (profiles/r1_summary.md:
// 84 k warp instructions per voxel spread evenly over a dozen generic loops,
// 8 warps per SM because every warp reserved a 64 x 64 inverse):
//
//   * lane l owns the 8 CONSECUTIVE bins 8l .. 8l+7.  Their duals live in
//     registers; the banded regulariser reads one window x[8l-W .. 8l+7+W] of the
//     coefficient vector instead of 2W+1 values per bin (the vector is stored
//     lane-major so these loads are bank-conflict free) and takes the interior
//     (Toeplitz) weights of mu^2 R^T R from registers; the whole dual pass is
//     one fully unrolled block (8 x (MT/2 + 3) LDS.128, 8 x (MT + 2W+1) DFMA) that
//     also tracks the lane's best candidate, and the warp-wide arg-max is two
//     redux.sync on the halves of the (positive) double instead of a 5-step
//     (value, index) shuffle butterfly.
//     The dictionary is stored bin-permuted (row q*32 + l holds bin 8l + q) so the
//     128-bit loads of consecutive lanes stay bank-conflict free.
//   * L&H's z-test numerator h_j - G_jP z is the dual w_j that selected the
//     candidate, so no second dot product / reduction is needed, and the
//     independence test only takes its two square roots when the pivot is within
//     1e-18 of being absorbed.
//   * rows >= k of H and z are kept at zero, which turns the bordering update into
//     the same rank-one update as every other row (u = [v; -1]).
//   * shared memory per warp holds rows 0 .. KB-1 of the packed inverse (KB = 36
//     covers 81 % of the voxels of config C3); a warp whose active set outgrows
//     that borrows extension areas (rows KB .. K1-1, then K1 .. 95) from a CTA-wide
//     pool and returns them when the voxel is done.  12 warps per SM instead of 8,
//     96 slots instead of 64.
//
// Status / hand-over protocol, iteration counter, certification thresholds:
// identical to pnb_nnls_fast.cuh (see there).
#pragma once
#include "pnb_nnls_fast.cuh"

namespace pnb {

#ifndef PNB_V3_KB
#define PNB_V3_KB 36
#endif
#ifndef PNB_V3_K1
#define PNB_V3_K1 56
#endif
#ifndef PNB_V3_MAXWARPS
#define PNB_V3_MAXWARPS 12
#endif
// unroll factors of the dual pass over the lane's eight bins (code size against instruction-level
// parallelism: see the note on the instruction cache in the kernel)
#ifndef PNB_V3_UQ32
#define PNB_V3_UQ32 2
#endif
#ifndef PNB_V3_UQX
#define PNB_V3_UQX 8
#endif
// FP32 screening of the dual pass (-DPNB_V3_SCREEN): built and measured in round 2, off by default.
// It does what it was designed for — ncu: 16 % fewer shared-memory wavefronts, FP64 pipe 22 % -> 11 %,
// bit-identical spectra — and the kernel gets SLOWER (full C3 volume: 474 -> 538 ms): the FP32 copy of
// the dictionary costs two of the twelve warps per SM (-10 % on its own) and the added serial work per
// trip costs more than the saved bandwidth returns; the kernel is bound by dependent-instruction
// latency at three warps per scheduler, not by the shared-memory pipe alone
// (profiles/r2_nnls_experiments.md).
#ifndef PNB_V3_SCREEN
#define PNB_V3_NO_SCREEN 1
#endif
#define PNB_PRAGMA_(x) _Pragma(#x)
#define PNB_UNROLL(n) PNB_PRAGMA_(unroll n)

template <int MT, int WK> struct NnlsV3Cfg {
  static constexpr int NQ = 8;                 // bins per lane
  static constexpr int NR = 32 * NQ;           // dictionary rows
  static constexpr int LD = MT + 2;            // row stride of the dictionary (doubles)
  static constexpr int BWK = 2 * WK + 1;       // band width the kernel is compiled for
  static constexpr int LB = (WK == 0) ? 0 : ((BWK + 1) & ~1);  // padded band row
  static constexpr int XW = NQ + 2 * WK;       // coefficient window of one lane: bins 8l-WK .. 8l+7+WK
  static constexpr int XR = 34;                // row stride of the coefficient scratch: column 0 / 33 stay zero
  static constexpr int NX = NQ * XR;           // x[j] lives at (j % 8) * XR + j / 8 + 1 (lane-major, conflict-free)
  static constexpr int KC = 96;                // slots
  static constexpr int KB = PNB_V3_KB;         // rows of H in the warp's own shared memory
  static constexpr int K1 = PNB_V3_K1;         // rows KB .. K1-1: tier-1 extension, K1 .. KC-1: tier 2
  static constexpr int TB = (KB * (KB + 1) / 2 + 1) & ~1;
  static constexpr int T1 = ((K1 * (K1 + 1) / 2 - KB * (KB + 1) / 2) + 1) & ~1;  // one tier-1 area
  static constexpr int T2 = ((KC * (KC + 1) / 2 - K1 * (K1 + 1) / 2) + 1) & ~1;  // one tier-2 area
  static constexpr int PER_WARP = NX + 3 * KC + MT + KC + TB;
  static constexpr int LDF = MT + 4;           // row stride (floats) of the FP32 copy of the dictionary: 128-bit
                                               // loads of consecutive lanes stay bank-conflict free
  static constexpr int BF = (NR * LDF + 1) / 2;  // doubles taken by the FP32 copy (only with screening)
  static constexpr int SHARED = NR * LD + NR * LB + NR + 2;
  __host__ __device__ static constexpr size_t smem_doubles(int warps, int e1, int e2, bool screen) {
    return (size_t)SHARED + (screen ? BF : 0) + (size_t)e1 * T1 + (size_t)e2 * T2 + (size_t)warps * PER_WARP;
  }
};

template <int MT, int WK>
__global__ void __launch_bounds__(PNB_V3_MAXWARPS * 32, 1) nnls_v3_kernel(const NnlsDeviceArgs a, const int n_e1, const int n_e2) {
  using C = NnlsV3Cfg<MT, WK>;
  constexpr int NQ = C::NQ, NR = C::NR, LD = C::LD, LB = C::LB, KC = C::KC, KB = C::KB, K1 = C::K1;
  extern __shared__ __align__(16) double smem_v3[];
  double *smem = smem_v3;
  const int m = a.m, n = a.n, W = a.W;
  const int lane = threadIdx.x & 31, wid = threadIdx.x >> 5;
  const int nthreads = blockDim.x;
  const unsigned FULL = 0xffffffffu;
  // ---- CTA-shared data ----------------------------------------------------------
  double *Bt = smem;                       // [NR][LD], bin-permuted rows
  double *rtrs = Bt + NR * LD;             // [NR][LB], same row order, taps centred on WK
  double *gdiag = rtrs + NR * LB;          // [NR], natural bin order
  int *pool_mask = reinterpret_cast<int *>(gdiag + NR);
  float *Bf = reinterpret_cast<float *>(gdiag + NR + 2);  // [NR][LDF] FP32 copy of Bt (screening pass)
  double *pool1 = gdiag + NR + 2 + (a.screen ? C::BF : 0);  // n_e1 areas of T1 doubles (rows KB .. K1-1)
  double *pool2 = pool1 + (size_t)n_e1 * C::T1;        // n_e2 areas of T2 doubles (rows K1 .. KC-1)
  double *wbase = pool2 + (size_t)n_e2 * C::T2 + (size_t)wid * C::PER_WARP;
  double *xs_raw = wbase;                  // NX, x[j] at xs_raw[xpos(j)]
  double *gsm = xs_raw + C::NX;            // KC
  double *usm = gsm + KC;                  // KC
  double *zs = usm + KC;                   // KC
  double *rs = zs + KC;                    // MT: residual y - B x (or y)
  int *ro = reinterpret_cast<int *>(rs + MT);  // KC: offset of the slot's dictionary row in Bt
  int *Pb = ro + KC;                           // KC: bin of the slot
  double *Hb = rs + MT + KC;               // packed rows 0 .. KB-1
  auto xpos = [](int j) -> int { return (j & 7) * C::XR + (j >> 3) + 1; };

  for (int i = threadIdx.x; i < NR * LD; i += nthreads) {
    const int row = i / LD, b = i - row * LD;
    const int j = NQ * (row & 31) + (row >> 5);
    Bt[i] = (j < n && b < m) ? a.B[(size_t)b * n + j] : 0.0;
  }
  for (int i = threadIdx.x; a.screen && i < NR * C::LDF; i += nthreads) {
    const int row = i / C::LDF, b = i - row * C::LDF;
    const int j = NQ * (row & 31) + (row >> 5);
    Bf[i] = (j < n && b < m) ? (float)a.B[(size_t)b * n + j] : 0.0f;
  }
  if (LB > 0) {
    for (int i = threadIdx.x; i < NR * LB; i += nthreads) {
      const int row = i / LB, t = i - row * LB;
      const int j = NQ * (row & 31) + (row >> 5);
      const int d = t - WK;  // column offset of this tap
      rtrs[i] = (j < n && d >= -W && d <= W) ? a.rtr[(size_t)j * (2 * W + 1) + d + W] : 0.0;
    }
  }
  if (threadIdx.x == 0) { pool_mask[0] = 0; pool_mask[1] = 0; }
  __syncthreads();
  for (int j = threadIdx.x; j < NR; j += nthreads) {
    const int row = (j & 7) * 32 + (j >> 3);
    double acc = (j < n) ? a.rtr[(size_t)j * (2 * W + 1) + W] : 0.0;
    for (int b = 0; b < MT; b++) acc += Bt[row * LD + b] * Bt[row * LD + b];
    gdiag[j] = acc;
  }
  // interior rows of mu^2 R^T R are identical for every regularisation order of the reference
  // (model_functions/nnls.py:46-85): their 2 WK + 1 weights live in registers, only the first and
  // last W bins read their own row (12 % of all shared-memory wavefronts went into those rows)
  double wt[C::BWK];
  unsigned bmask = 0xffu;
  // bins >= n do not exist (zero dictionary rows): they are never candidates, whatever the
  // Toeplitz weights make of the neighbouring coefficients
  unsigned nomask = 0;
#pragma unroll
  for (int q = 0; q < NQ; q++)
    if (NQ * lane + q >= n) nomask |= 1u << q;
  if (LB > 0) {
    const int jmid = n / 2, rmid = (jmid & 7) * 32 + (jmid >> 3);
    bool same = true;
    for (int j = W + threadIdx.x; j < n - W; j += nthreads) {
      const int row = (j & 7) * 32 + (j >> 3);
      for (int t = 0; t < C::BWK; t++) same = same && (rtrs[row * LB + t] == rtrs[rmid * LB + t]);
    }
    const bool toep = __syncthreads_and(same) && n > 2 * W + 1;
#pragma unroll
    for (int t = 0; t < C::BWK; t++) wt[t] = rtrs[rmid * LB + t];
    if (toep) {
      bmask = 0;
#pragma unroll
      for (int q = 0; q < NQ; q++) {
        const int j = NQ * lane + q;
        // only bins that exist: with n = 250 the six padding bins of lane 31 used to take the slow
        // row-from-shared-memory path alone, the other 31 lanes idle (9 % of the kernel's samples)
        if (j < n && (j < W || j >= n - W)) bmask |= 1u << q;
      }
    }
  }
  for (int i = lane; i < C::NX; i += 32) xs_raw[i] = 0.0;
  for (int i = lane; i < KC; i += 32) zs[i] = 0.0;
  for (int i = lane; i < C::TB; i += 32) Hb[i] = 0.0;
  __syncthreads();

  const int kcap = n_e1 > 0 ? (n_e2 > 0 ? KC : K1) : KB;
  enum { PH_INIT = 0, PH_ITER = 1, PH_POLISH = 2, PH_VERIFY = 3, PH_CHECK = 4 };

  // Every voxel is a sequence of TRIPS through one loop body whose big blocks (residual, dual,
  // H * vector, rank-one update of H) exist exactly once in the instruction stream: the first
  // version of this kernel inlined them at every use, grew to 178 KB of SASS and spent 44 % of
  // its warp samples waiting for instruction fetch (profiles/r1_nnls_v3.md).
  for (;;) {
    unsigned long long vq = 0;
    if (lane == 0) vq = atomicAdd(a.counter, 1ULL);
    const long long vox = (long long)__shfl_sync(FULL, vq, 0);
    if (vox >= a.n_vox) break;

    const double yv = (lane < m) ? a.y[vox * m + lane] : 0.0;  // lane b holds y_b
    const bool fin = __all_sync(FULL, finite_d(yv));

    unsigned inP = 0;    // bit q: bin 8 lane + q is active
    unsigned rej = 0;    // bit q: bin 8 lane + q was rejected as a candidate since the last addition
    int k = 0, iter = 0, mode = fin ? 1 : 2;
    int ext1 = -1, ext2 = -1;
    double *H1 = nullptr, *H2 = nullptr;  // extension areas, biased so that H? + i (i + 1) / 2 is row i
    int phase = PH_INIT, pass = 0;
    bool do_dual = true;
    double hmax = 0.0, rel = 1.0, zscale = 0.0;
#ifdef PNB_V3_NO_SCREEN
    const bool screen_ok = false;    // (variant without the screening code, for measurements)
#else
    bool screen_ok = a.screen != 0;  // FP32 screening of the dual pass still pays for this voxel
#endif
    int jc = -1;          // PH_CHECK: the candidate whose would-be coefficient is being computed
    double wjc = 0.0;     // ... and its dual

    auto Htier = [&](int i) -> double * { return i < KB ? Hb : (i < K1 ? H1 : H2); };
    auto Hrow = [&](int i) -> double * { return Htier(i) + (i * (i + 1)) / 2; };
    // one dictionary row dotted with a register vector
    auto col_dot = [&](int row, const double (&vec)[MT]) -> double {
      const double2 *bp = reinterpret_cast<const double2 *>(Bt + row * LD);
      double a0 = 0.0, a1 = 0.0;
#pragma unroll
      for (int t = 0; t < MT / 2; t++) {
        const double2 v = bp[t];
        a0 += v.x * vec[2 * t];
        a1 += v.y * vec[2 * t + 1];
      }
      return a0 + a1;
    };
    // (mu^2 R^T R x)_j for an arbitrary bin
    auto band_at = [&](int row, int j) -> double {
      double acc = 0.0;
      if (LB > 0) {
        const double *rp = rtrs + row * LB;
#pragma unroll
        for (int t = 0; t < C::BWK; t++) {
          int jj = j - WK + t;  // taps outside [0, NR) carry a zero weight: any valid address will do
          jj = jj < 0 ? 0 : (jj >= NR ? NR - 1 : jj);
          acc += rp[t] * xs_raw[xpos(jj)];
        }
      }
      return acc;
    };

    // h = B^T y materialised by the tensor-core GEMM (dual_init = 1): the lane's best candidate of the
    // first trip comes from there.  Kept out of the trip loop: global loads inside it cost the
    // whole kernel 50 % (measured, round 2) although they execute once per voxel.
    const bool have_h0 = a.h0 != nullptr;
    double best_h0 = 0.0;
    int bq_h0 = -1;
    if (have_h0 && mode == 1) {
      const double *hp = a.h0 + vox * (long long)n + NQ * lane;
#pragma unroll
      for (int q = 0; q < NQ; q++) {
        const double acc = (NQ * lane + q < n) ? hp[q] : 0.0;
        if (acc > best_h0) { best_h0 = acc; bq_h0 = q; }
      }
    }
    if (mode == 1) for (;;) {
      // ---- rs = y - B_P z (k = 0: rs = y) ----------------------------------------------------
      {
        constexpr int G = (MT <= 8) ? 4 : (MT <= 16 ? 2 : 1);  // lane groups sharing the slots
        const int b = lane % MT, grp = lane / MT;
        double a0 = 0.0, a1 = 0.0;
        if (grp < G) {
          int i = grp;
#pragma unroll 1
          for (; i + 3 * G < k; i += 4 * G) {
            const int r0 = ro[i], r1 = ro[i + G], r2 = ro[i + 2 * G], r3 = ro[i + 3 * G];
            const double z0 = zs[i], z1 = zs[i + G], z2 = zs[i + 2 * G], z3 = zs[i + 3 * G];
            const double b0 = Bt[r0 + b], b1 = Bt[r1 + b], b2 = Bt[r2 + b], b3 = Bt[r3 + b];
            a0 += b0 * z0;
            a1 += b1 * z1;
            a0 += b2 * z2;
            a1 += b3 * z3;
          }
#pragma unroll 1
          for (; i < k; i += G) a0 += Bt[ro[i] + b] * zs[i];
        }
        double acc = a0 + a1;
        if (G >= 2) acc += __shfl_xor_sync(FULL, acc, MT);
        if (G >= 4) acc += __shfl_xor_sync(FULL, acc, 2 * MT);
        __syncwarp();
        if (lane < MT) rs[lane] = yv - acc;
        __syncwarp();
      }
      double rr[MT];  // the residual, replicated
      {
        const double2 *rp = reinterpret_cast<const double2 *>(rs);
#pragma unroll
        for (int t = 0; t < MT / 2; t++) { const double2 v = rp[t]; rr[2 * t] = v.x; rr[2 * t + 1] = v.y; }
      }
      // ---- duals of this lane's bins, best positive one ----------------------------------------
      double best = 0.0;
      int bq = -1;
      if (phase == PH_INIT && have_h0) {
        best = best_h0; bq = bq_h0;  // x = 0: the dual is h = B^T y, read before the loop
      } else if (do_dual) {
        double xw[C::XW > 0 ? C::XW : 1];
        if (LB > 0) {
          // x[8 lane - WK .. 8 lane + 7 + WK]: own bins from the lane's column, the halo from the neighbours'
#pragma unroll
          for (int t = 0; t < C::XW; t++) {
            const int qq = t - WK;  // bin offset relative to 8 lane
            const int q8 = qq < 0 ? qq + NQ : (qq >= NQ ? qq - NQ : qq);
            const int dl = qq < 0 ? -1 : (qq >= NQ ? 1 : 0);
            xw[t] = xs_raw[q8 * C::XR + lane + 1 + dl];
          }
        }
        const unsigned skip = inP | rej | nomask;
        // (mu^2 R^T R x)_j of the lane's bin q
        auto band_q = [&](int q) -> double {
          if (LB == 0) return 0.0;
          const int row = q * 32 + lane;
          double b0 = 0.0, b1 = 0.0;
          if ((bmask >> q) & 1u) {
            // first / last W bins (or a regulariser that is not Toeplitz): the row from shared memory
            const double2 *rp = reinterpret_cast<const double2 *>(rtrs + row * LB);
#pragma unroll
            for (int t = 0; t < LB / 2; t++) {
              const double2 v = rp[t];
              b0 += v.x * xw[q + 2 * t];
              if (2 * t + 1 < C::BWK) b1 += v.y * xw[q + 2 * t + 1];
            }
          } else {
#pragma unroll
            for (int t = 0; t < C::BWK; t += 2) {
              b0 += wt[t] * xw[q + t];
              if (t + 1 < C::BWK) b1 += wt[t + 1] * xw[q + t + 1];
            }
          }
          return b0 + b1;
        };
        // the regulariser term of the lane's eight bins, once per trip (every use below reads these
        // registers: inlining band_q at each use tripled the loop body and the kernel started to wait
        // for instruction fetch — 2.3 no_instruction stalls per issue, 50 % slower)
        double bandv[NQ];
#pragma unroll
        for (int q = 0; q < NQ; q++) bandv[q] = band_q(q);
        // element q (a run-time index) of a register array
        auto pick = [](const double (&v)[NQ], int q) -> double {
          double r = v[0];
#pragma unroll
          for (int t = 1; t < NQ; t++) r = (q == t) ? v[t] : r;
          return r;
        };
        bool exact_pass = true;
        if (screen_ok && phase != PH_VERIFY) {
          // ---- FP32 screening.  Half of this kernel's shared-memory traffic was the FP64 dictionary in
          // this pass, and all the pass has to deliver is the arg-max.  The dictionary products are taken
          // in FP32 first (half the bytes, the idle FP32 pipe): |w32_j - w_j| <= E = 2e-6 ||r||_1 (18
          // roundings of 2^-24 on sum_b |B_bj r_b|, B <= 1, doubled), so the exact arg-max is among the bins
          // with w32_j >= max w32 - 2 E — 1.3 bins on average — and only those are evaluated in FP64,
          // with the same instructions as the full pass: the result is bit-identical.  When the largest
          // dual has come down to the FP32 noise (the last few additions) the voxel returns to the full
          // FP64 pass for good.
          float r32[MT];
          double rn1 = 0.0;
#pragma unroll
          for (int b = 0; b < MT; b++) { r32[b] = (float)rr[b]; rn1 += fabs(rr[b]); }
          const double E = 2e-6 * rn1;
          // rolled loops from here on: the whole kernel has to stay under 4096 instructions (64 KB), the
          // instruction cache it is served from — 88 instructions above that it ran 50 % slower
          float w32[NQ];
          double lmax = 0.0;
          PNB_UNROLL(PNB_V3_UQ32)
          for (int q = 0; q < NQ; q++) {
            const float4 *bp = reinterpret_cast<const float4 *>(Bf + (q * 32 + lane) * C::LDF);
            float s0 = 0.0f, s1 = 0.0f;
#pragma unroll
            for (int t = 0; t < MT / 4; t++) {
              const float4 v = bp[t];
              s0 = fmaf(v.x, r32[4 * t], s0);
              s1 = fmaf(v.y, r32[4 * t + 1], s1);
              s0 = fmaf(v.z, r32[4 * t + 2], s0);
              s1 = fmaf(v.w, r32[4 * t + 3], s1);
            }
            const double w = (double)(s0 + s1) - pick(bandv, q);
#pragma unroll
            for (int t = 0; t < NQ; t++)
              if (q == t) w32[t] = (float)w;
            if (!((skip >> q) & 1u) && w > lmax) lmax = w;
          }
          const unsigned hi = (unsigned)__double2hiint(lmax);
          const unsigned mhi = __reduce_max_sync(FULL, hi);
          const unsigned lo = (hi == mhi) ? (unsigned)__double2loint(lmax) : 0u;
          const unsigned mlo = __reduce_max_sync(FULL, lo);
          const double wmax32 = __hiloint2double((int)mhi, (int)mlo);
          if (wmax32 > 32.0 * E) {
            exact_pass = false;
            const float thr = (float)(wmax32 - 2.0 * E - 2e-7 * wmax32);  // w32 itself is rounded to float
            // the candidates (1.3 per trip on average), one rolled loop: q selects its registers
            unsigned cand = 0;
#pragma unroll
            for (int q = 0; q < NQ; q++)
              if (!((skip >> q) & 1u) && w32[q] >= thr) cand |= 1u << q;
#pragma unroll 1
            while (cand) {
              const int q = __ffs(cand) - 1;
              cand &= cand - 1;
              const double acc = col_dot(q * 32 + lane, rr) - pick(bandv, q);
              if (acc > best) { best = acc; bq = q; }
            }
          } else {
#ifndef PNB_V3_NO_SCREEN
            screen_ok = false;
#endif
          }
        }
        if (exact_pass) {
          PNB_UNROLL(PNB_V3_UQX)
          for (int q = 0; q < NQ; q++) {
            const double acc = col_dot(q * 32 + lane, rr) - pick(bandv, q);
            if (!((skip >> q) & 1u) && acc > best) { best = acc; bq = q; }
          }
        }
      }
      int j = -1, jrow = 0;
      double wj = 0.0;
      if (phase != PH_POLISH) {
        if (phase == PH_CHECK) {
          j = jc; wj = wjc;
        } else {
          // warp-wide arg-max of the positive duals; the lowest bin wins ties
          const unsigned hi = (unsigned)__double2hiint(best);
          const unsigned mhi = __reduce_max_sync(FULL, hi);
          const unsigned lo = (hi == mhi) ? (unsigned)__double2loint(best) : 0u;
          const unsigned mlo = __reduce_max_sync(FULL, lo);
          const unsigned win = __ballot_sync(FULL, bq >= 0 && hi == mhi && lo == mlo);
          if (win) {
            j = __shfl_sync(FULL, NQ * lane + bq, __ffs(win) - 1);
            wj = __hiloint2double((int)mhi, (int)mlo);
          }
        }
        if (phase == PH_INIT) { hmax = wj; phase = PH_ITER; }
        if (phase == PH_VERIFY) {
          // duals of bins that sit at the optimum are zero to rounding (+-1e-13 relative) in every
          // voxel; anything clearly positive means this is not a Kuhn-Tucker point
          if (wj > 1e-12 * hmax) { mode = kNnlsRedo; break; }
          // A positive dual below that threshold may be rounding noise or the start of one more
          // Lawson-Hanson step, and SciPy takes that step for ANY positive dual.  What matters is
          // the coefficient the bin would enter with, dual / pivot^2 — large when the bin is nearly
          // dependent on the active ones.  One more trip (no dual pass) computes the pivot.
          if (j < 0 || k == 0 || k == kcap) break;
          phase = PH_CHECK; do_dual = false; jc = j; wjc = wj;
          continue;
        }
        if (j < 0) {  // Kuhn-Tucker point of the carried solution: polish it
          if (k == 0) break;
          phase = PH_POLISH; pass = 0; do_dual = false; rej = 0;
          continue;
        }
        // candidate column in registers, gsm = G_Pj
        jrow = (j & 7) * 32 + (j >> 3);
        double cj[MT];
        {
          const double2 *bp = reinterpret_cast<const double2 *>(Bt + jrow * LD);
#pragma unroll
          for (int t = 0; t < MT / 2; t++) { const double2 c2 = bp[t]; cj[2 * t] = c2.x; cj[2 * t + 1] = c2.y; }
        }
#pragma unroll 1
        for (int i = lane; i < k; i += 32) {
          const int p = Pb[i];
          const int row = (p & 7) * 32 + (p >> 3);
          double acc = col_dot(row, cj);
          if (LB > 0) {
            const int d = j - p;
            if (d >= -WK && d <= WK) acc += rtrs[row * LB + d + WK];
          }
          gsm[i] = acc;
        }
      } else {
        // gradient on P from the true residual
#pragma unroll 1
        for (int i = lane; i < k; i += 32) {
          const int p = Pb[i];
          const int row = (p & 7) * 32 + (p >> 3);
          gsm[i] = col_dot(row, rr) - band_at(row, p);
        }
      }
      __syncwarp();
      // ---- usm = H gsm, p0 = gsm . usm ----------------------------------------------------------
      double p0 = 0.0;
#pragma unroll 1
      for (int i = lane; i < k; i += 32) {
        // element (max(i, c), min(i, c)) of the packed lower triangle, c = 0 .. k-1: the row of i up
        // to the diagonal, then down its column — one loop, so a lane never waits for the longer
        // row or column of another lane
        double a0 = 0.0, a1 = 0.0;
        const double *prow = Hrow(i);  // &H(i, c)
        int c = 0, inc = 1;            // inc = c + 1: distance from H(c, i) to H(c + 1, i)
        int off = i;                   // c (c + 1) / 2 + i
#pragma unroll 1
        for (int seg = 0; seg < 3; seg++) {
          const int lim = seg == 0 ? KB : (seg == 1 ? K1 : KC);
          const int ke = k < lim ? k : lim;
          if (c >= ke) continue;
          const double *pcol = (seg == 0 ? Hb : (seg == 1 ? H1 : H2)) + off;  // &H(c, i), used when c > i
#pragma unroll 1
          for (; c + 3 < ke; c += 4) {
            const double *q0 = (c <= i) ? prow : pcol;
            const double *q1 = (c + 1 <= i) ? prow + 1 : pcol + inc;
            const double *q2 = (c + 2 <= i) ? prow + 2 : pcol + 2 * inc + 1;
            const double *q3 = (c + 3 <= i) ? prow + 3 : pcol + 3 * inc + 3;
            const double h0 = *q0, h1 = *q1, h2 = *q2, h3 = *q3;
            const double g0 = gsm[c], g1 = gsm[c + 1], g2 = gsm[c + 2], g3 = gsm[c + 3];
            a0 += h0 * g0;
            a1 += h1 * g1;
            a0 += h2 * g2;
            a1 += h3 * g3;
            pcol += 4 * inc + 6;
            off += 4 * inc + 6;
            inc += 4;
            prow += 4;
          }
#pragma unroll 1
          for (; c < ke; c++) {
            const double *q0 = (c <= i) ? prow : pcol;
            a0 += *q0 * gsm[c];
            pcol += inc;
            off += inc;
            inc += 1;
            prow += 1;
          }
        }
        const double vi = a0 + a1;
        usm[i] = vi;
        p0 += gsm[i] * vi;
      }
      if (phase == PH_POLISH) {
        // one refinement step: z += H grad
        bool pos = true;
        double dmax_ = 0.0, zmax = 0.0;
#pragma unroll 1
        for (int i = lane; i < k; i += 32) {
          const double z = zs[i], dz = usm[i];
          pos = pos && (z + dz > 0.0);
          dmax_ = fmax(dmax_, fabs(dz));
          zmax = fmax(zmax, fabs(z));
        }
        pos = __all_sync(FULL, pos);
#pragma unroll
        for (int o = 16; o > 0; o >>= 1) {
          dmax_ = fmax(dmax_, __shfl_xor_sync(FULL, dmax_, o));
          zmax = fmax(zmax, __shfl_xor_sync(FULL, zmax, o));
        }
        rel = dmax_ / zmax;
        zscale = zmax;
        // the first correction measures how far the carried solution had drifted while the
        // active-set decisions were being made; wrong results only appeared above 3e-4
        if ((pass == 0 && rel > 5e-5) || !pos) { mode = kNnlsRedo; break; }
#pragma unroll 1
        for (int i = lane; i < k; i += 32) { const double z = zs[i] + usm[i]; zs[i] = z; xs_raw[xpos(Pb[i])] = z; }
        __syncwarp();
        pass += 1;
        if (rel < 1e-13 || pass == 4) {
          if (rel > 1e-10) { mode = kNnlsRedo; break; }  // refinement did not converge
          phase = PH_VERIFY;
          do_dual = k < n;
        }
        continue;
      }
      // ---- PH_ITER: Lawson-Hanson's independence and z tests for candidate j ------------------
      p0 = warp_sum(p0);
      if (phase == PH_CHECK) {
        // would-be coefficient of the sub-threshold candidate at the polished point: negligible ->
        // the voxel is certified; otherwise the robust path (SciPy's own arithmetic) decides
        const double piv2 = gdiag[j] - (p0 > 0.0 ? p0 : 0.0);
        if (!(piv2 > 0.0) || wj > a.cert_ztol * fmax(1.0, 1e-4 * zscale) * piv2) mode = kNnlsRedo;
        break;
      }
      double sinv = 0.0, zeta = 0.0;
      {
        const double unorm2 = p0 > 0.0 ? p0 : 0.0;
        const double piv2 = gdiag[j] - unorm2;
        bool ok = piv2 > 1e-18 * unorm2 && piv2 > 0.0;
        if (!ok && piv2 > 0.0) {
          // the test taken literally when the pivot is close to being absorbed
          const double av = sqrt(piv2), unorm = sqrt(unorm2);
          ok = ((unorm + av * 0.01) - unorm) > 0.0;
        }
        if (ok) {
          sinv = 1.0 / piv2;
          zeta = wj * sinv;  // h_j - G_jP z is the dual that selected j
          ok = zeta > 0.0;
        }
        if (!ok) {
          // not a candidate again before the next addition
          if ((j >> 3) == lane) rej |= 1u << (j & 7);
          continue;
        }
      }
      rej = 0;
      if (k == kcap) { mode = kNnlsRedo; break; }
      if ((k == KB && ext1 < 0) || (k == K1 && ext2 < 0)) {
        // borrow an extension area (holders always finish, so waiting is safe)
        const bool first = k == KB;
        int *mask = pool_mask + (first ? 0 : 1);
        int e = 0;
        if (lane == 0) {
          const int all = (1 << (first ? n_e1 : n_e2)) - 1;
          for (;;) {
            const int cur = *reinterpret_cast<volatile int *>(mask);
            const int freeb = ~cur & all;
            if (freeb) {
              e = __ffs(freeb) - 1;
              if (atomicCAS(mask, cur, cur | (1 << e)) == cur) break;
            } else {
              __nanosleep(200);
            }
          }
        }
        e = __shfl_sync(FULL, e, 0);
        double *area = first ? pool1 + (size_t)e * C::T1 : pool2 + (size_t)e * C::T2;
        const int na = first ? C::T1 : C::T2;
#pragma unroll 1
        for (int i = lane; i < na; i += 32) area[i] = 0.0;
        if (first) { ext1 = e; H1 = area - (KB * (KB + 1)) / 2; }
        else { ext2 = e; H2 = area - (K1 * (K1 + 1)) / 2; }
      }
      // ---- rank-one operations on H and z: the bordering update H += u u^T / s with
      //      u = [H g; -1], then Lawson-Hanson's secondary loop (each removal is a downdate) ---
      {
        if (lane == 0) { usm[k] = -1.0; Pb[k] = j; ro[k] = jrow * LD; }
        if ((j >> 3) == lane) inP |= 1u << (j & 7);
        double scale = sinv, zfac = zeta;
        int kk = k + 1;  // rows the operation touches
        int q = -1;      // slot being removed; -1: the bordering update prepared above
        for (;;) {
          if (q >= 0) {
            // usm = column q of H; H -= usm usm^T / H_qq, z -= usm z_q / H_qq
            const int idx = Pb[q];
            if ((idx >> 3) == lane) inP &= ~(1u << (idx & 7));
            const double dq = Hrow(q)[q], zq = zs[q];
#pragma unroll 1
            for (int i = lane; i < k; i += 32) usm[i] = (i >= q) ? Hrow(i)[q] : Hrow(q)[i];
            if (lane == 0) xs_raw[xpos(idx)] = 0.0;
            const double dinv = 1.0 / dq;
            scale = -dinv;
            zfac = zq * dinv;
            kk = k;
          }
          __syncwarp();
          // gsm = scale * usm, z -= zfac * usm; then every packed element (i, c) += gsm[i] * usm[c],
          // the elements dealt out to all 32 lanes
#pragma unroll 1
          for (int i = lane; i < kk; i += 32) {
            const double ui = usm[i];
            gsm[i] = ui * scale;
            zs[i] -= ui * zfac;
          }
          __syncwarp();
          {
            // row i of H += gsm[i] * usm[0 .. i]: lane l owns columns l, l + 32, l + 64, the rows
            // are walked by the whole warp (conflict-free, no per-lane row lengths)
            const double u0 = usm[lane];
            const double u1 = (kk > 32) ? usm[lane + 32] : 0.0;
            const double u2 = (kk > 64) ? usm[lane + 64] : 0.0;
            int i = 0;
            {
              const int ke = kk < 32 ? kk : 32;
              double *hp = Hb + lane;
#pragma unroll 1
              for (; i + 3 < ke; i += 4) {
                // four rows per trip, every load issued before the first store (the rows are
                // distinct memory, which the compiler cannot know)
                double *h1 = hp + i + 1, *h2 = h1 + i + 2, *h3 = h2 + i + 3;
                const double c0 = gsm[i], c1 = gsm[i + 1], c2 = gsm[i + 2], c3 = gsm[i + 3];
                const double v0 = *hp, v1 = *h1, v2 = *h2, v3 = *h3;  // lanes past the row end read the next rows
                const double w0 = v0 + c0 * u0, w1 = v1 + c1 * u0, w2 = v2 + c2 * u0, w3 = v3 + c3 * u0;
                if (lane <= i) *hp = w0;
                if (lane <= i + 1) *h1 = w1;
                if (lane <= i + 2) *h2 = w2;
                if (lane <= i + 3) *h3 = w3;
                hp = h3 + i + 4;
              }
#pragma unroll 1
              for (; i < ke; i++) {
                if (lane <= i) *hp += gsm[i] * u0;
                hp += i + 1;
              }
            }
#pragma unroll 1
            for (; i < kk; i++) {
              const double c0 = gsm[i];
              double *hp = Hrow(i) + lane;
              *hp += c0 * u0;
              if (lane + 32 <= i) hp[32] += c0 * u1;
              if (lane + 64 <= i) hp[64] += c0 * u2;
            }
          }
          __syncwarp();
          if (q < 0) {
            k = kk;
          } else {
            // the last slot takes the place of q; the vacated last row / z entry return to zero
            const int last = k - 1;
            double mv[3];  // element (last, c) for c = lane, lane + 32, lane + 64
#pragma unroll
            for (int t = 0; t < 3; t++) {
              const int c = lane + 32 * t;
              mv[t] = (c <= last) ? Hrow(last)[c] : 0.0;
            }
            const double zl = zs[last];
            const int pl = Pb[last], rl = ro[last];
            __syncwarp();
#pragma unroll
            for (int t = 0; t < 3; t++) {
              const int c = lane + 32 * t;
              if (c <= last) {
                Hrow(last)[c] = 0.0;
                if (q != last) {
                  if (c == last) Hrow(q)[q] = mv[t];
                  else if (c < q) Hrow(q)[c] = mv[t];
                  else if (c > q) Hrow(c)[q] = mv[t];
                }
              }
            }
            if (lane == 0) {
              zs[last] = 0.0;
              if (q != last) { zs[q] = zl; Pb[q] = pl; ro[q] = rl; }
            }
            k -= 1;
            __syncwarp();
            // anything else pushed to (or below) zero by the step?
            int bad = KC;
#pragma unroll 1
            for (int i = lane; i < k; i += 32)
              if (xs_raw[xpos(Pb[i])] <= 0.0 && i < bad) bad = i;
            bad = __reduce_min_sync(FULL, bad);
            if (bad < KC) { q = bad; continue; }
          }
          iter += 1;
          // reaching the cap is only believed from the robust path, whose iteration count is SciPy's
          if (iter >= a.maxiter) { mode = kNnlsRedo; break; }
          bool neg = false;
#pragma unroll 1
          for (int i = lane; i < k; i += 32) neg = neg || (zs[i] <= 0.0);
          if (!__any_sync(FULL, neg)) break;
          double alpha = 2.0;
          int jj = -1;
#pragma unroll 1
          for (int i = lane; i < k; i += 32) {
            const double z = zs[i];
            if (z <= 0.0) {
              const double xv = xs_raw[xpos(Pb[i])];
              const double t = -xv / (z - xv);
              if (alpha > t) { alpha = t; jj = i; }
            }
          }
#pragma unroll
          for (int o = 16; o > 0; o >>= 1) {
            const double oa = __shfl_xor_sync(FULL, alpha, o);
            const int oj = __shfl_xor_sync(FULL, jj, o);
            if (oj >= 0 && (jj < 0 || oa < alpha || (oa == alpha && oj < jj))) { alpha = oa; jj = oj; }
          }
#pragma unroll 1
          for (int i = lane; i < k; i += 32) {
            const int p = xpos(Pb[i]);
            xs_raw[p] += alpha * (zs[i] - xs_raw[p]);
          }
          __syncwarp();
          q = jj;
        }
      }
      if (mode != 1) break;
#pragma unroll 1
      for (int i = lane; i < k; i += 32) xs_raw[xpos(Pb[i])] = zs[i];
      __syncwarp();
    }

    // rs is the residual of the final coefficients (last trip: verification, or k = 0)
    double *out = a.coef + vox * (long long)n;
    if (mode != kNnlsRedo) {
      double ss = 0.0, part = 0.0;
      if (mode == 1) {
#pragma unroll
        for (int b = 0; b < MT; b++) ss += rs[b] * rs[b];
#pragma unroll 2
        for (int jx = lane; jx < n; jx += 32) {
          const double xj = xs_raw[xpos(jx)];
          out[jx] = xj;
          if (xj != 0.0) part += xj * band_at((jx & 7) * 32 + (jx >> 3), jx);
        }
      } else {
        part = yv * yv;  // failure: zeros and ||y||
#pragma unroll 2
        for (int jx = lane; jx < n; jx += 32) out[jx] = 0.0;
      }
      const double mean = warp_sum(yv) / (double)m;
      const double d = (lane < m) ? yv - mean : 0.0;
      double red[2] = {part, d * d};
#pragma unroll
      for (int o = 16; o > 0; o >>= 1) {
        red[0] += __shfl_xor_sync(FULL, red[0], o);
        red[1] += __shfl_xor_sync(FULL, red[1], o);
      }
      const double tot = ss + red[0];
      const double ss_res = (mode == 1) ? ss : tot;
      if (lane == 0) {
        a.rnorm[vox] = sqrt(tot > 0.0 ? tot : 0.0);
        if (a.r2) a.r2[vox] = (red[1] > 0.0) ? 1.0 - ss_res / red[1] : nan("");
      }
    }
    // leave the scratch clean for the next voxel: x = 0, z = 0, rows < k of H = 0
    __syncwarp();
#pragma unroll 1
    for (int i = lane; i < k; i += 32) { xs_raw[xpos(Pb[i])] = 0.0; zs[i] = 0.0; }
    {
      const int kb = k < KB ? k : KB;
      const int nb = (kb * (kb + 1)) / 2;
#pragma unroll 1
      for (int i = lane; i < nb; i += 32) Hb[i] = 0.0;
    }
    if (lane == 0) {
      if (ext1 >= 0) atomicAnd(pool_mask, ~(1 << ext1));
      if (ext2 >= 0) atomicAnd(pool_mask + 1, ~(1 << ext2));
      a.status[vox] = mode;
      a.iters[vox] = iter;
      if (mode == kNnlsRedo) {
        const unsigned long long slot = atomicAdd(a.redo_count, 1ULL);
        a.redo_list[slot] = (int)vox + a.redo_base;
      }
    }
    __syncwarp();
  }
}

}  // namespace pnb

namespace pnb {
// Launch one CTA per SM with as many warps (<= PNB_V3_MAXWARPS) and extension areas (3 + 1) as
// shared memory holds.  Returns cudaErrorInvalidConfiguration when not even 4 warps fit.
template <int MT, int WK>
cudaError_t nnls_v3_launch(const NnlsDeviceArgs &a_in, cudaStream_t stream) {
  using C = NnlsV3Cfg<MT, WK>;
  NnlsDeviceArgs a = a_in;
#ifdef PNB_V3_NO_SCREEN
  a.screen = 0;  // the screening code is not in this build: no shared memory for the FP32 dictionary either
#endif
  constexpr size_t budget = 227 * 1024;
  int warps = PNB_V3_MAXWARPS, e1 = 3, e2 = 1;
  const bool sc = a.screen != 0;
  while (warps > 4 && C::smem_doubles(warps, e1, e2, sc) * sizeof(double) > budget) warps--;
  while (e1 > 1 && C::smem_doubles(warps, e1, e2, sc) * sizeof(double) > budget) e1--;
  if (C::smem_doubles(warps, e1, e2, sc) * sizeof(double) > budget) e2 = 0;
  if (C::smem_doubles(warps, e1, e2, sc) * sizeof(double) > budget) e1 = 0;
  const size_t smem = C::smem_doubles(warps, e1, e2, sc) * sizeof(double);
  if (smem > budget) return cudaErrorInvalidConfiguration;
  auto kern = nnls_v3_kernel<MT, WK>;
  cudaError_t err = cudaFuncSetAttribute(kern, cudaFuncAttributeMaxDynamicSharedMemorySize, (int)smem);
  if (err != cudaSuccess) return err;
  static thread_local int sms = 0;
  if (!sms) {
    int dev = 0;
    err = cudaGetDevice(&dev);
    if (err != cudaSuccess) return err;
    err = cudaDeviceGetAttribute(&sms, cudaDevAttrMultiProcessorCount, dev);
    if (err != cudaSuccess) return err;
  }
  long long grid = sms;
  const long long want = (a.n_vox + warps - 1) / warps;
  if (want < grid) grid = want;
  if (grid < 1) grid = 1;
  kern<<<(unsigned)grid, warps * 32, smem, stream>>>(a, e1, e2);
  return cudaGetLastError();
}
}  // namespace pnb
