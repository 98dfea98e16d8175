// Batched Tikhonov-regularised NNLS: one warp per voxel, Lawson-Hanson active set.
//
// Replaces the per-voxel scipy.optimize.nnls(A, b, maxiter) call of
// solvers/nnls_solver.py:195-197 with A = [B; mu R] ((n_b + n_bins) x n_bins,
// nnls_solver.py:61-73) and b = [y; 0] (:75-86).
//
// SciPy's routine is the classical Lawson-Hanson algorithm on a Householder /
// Givens QR of the active columns of a *private, transformed copy of A*
// (532 KB per voxel at 266 x 250) — unusable for millions of voxels.  The
// kernel runs the same active-set iteration (same candidate selection, same
// independence / z-test with the 0.01 factor, same feasibility line search,
// same removal order, same iteration counter and maxiter rule) on the normal
// equations instead, which only need data shared by all voxels:
//
//   h   = A^T b   = B^T y                          (per voxel, n_bins)
//   G   = A^T A   = B^T B + mu^2 R^T R             (never stored: an entry is a
//                   16-term dot product of two basis columns held in shared
//                   memory plus one element of the banded mu^2 R^T R)
//   L L^T = G_PP  Cholesky factor of the active block, kept per warp in shared
//                   memory; a column entering P appends a row (forward
//                   substitution), a column leaving P deletes a row and
//                   re-triangularises with Givens rotations (exactly the
//                   G1/G2 step of Lawson-Hanson applied to R = L^T),
//   u   = L^-1 h_P carried along (L&H's transformed right-hand side),
//   z   = L^-T u   the least-squares solution on P,
//   w   = A^T (b - A x) = B^T (y - B x) - mu^2 R^T R x   the dual.
//
// Forming G squares the condition number, so after every outer iteration the
// solution on P gets one step of iterative refinement with the true residual
// (corrected semi-normal equations: dz = (L L^T)^-1 w_P, w_P being the part of
// the freshly computed dual that should vanish).  Measured against SciPy on the
// golden cases this brings the coefficient difference from ~2e-8 to ~2e-11 and
// reproduces SciPy's iteration counts exactly (scripts/proto_nnls_gram.py).
#pragma once
#include <cuda_runtime.h>

#include "pnb_hd.cuh"

namespace pnb {

struct NnlsDeviceArgs {
  int m, n, W, maxiter;  // measurements, bins, half-bandwidth of RtR, L&H iteration cap
  long long n_vox;
  const double *B;       // (m, n) row-major basis exp(-b D)
  const double *rtr;     // (n, 2W+1): rtr[j*(2W+1) + d + W] = (mu^2 R^T R)[j][j+d], 0 outside
  const double *y;       // (n_vox, m)
  double *coef;          // (n_vox, n)
  double *rnorm;         // (n_vox)
  int *status;           // (n_vox) 1 ok, 3 iteration cap reached (-> zeros, ||y||), 2 non-finite input
  int *iters;            // (n_vox)
  double *r2;            // (n_vox) or nullptr: 1 - ||y - B x||^2 / SS_tot
  unsigned long long *counter;
  double *scratch;       // per-warp overflow storage, (n (n+1) / 2 + 5 n) doubles per warp
  int kmax;              // active-set size that fits the shared-memory factor
  // fast path -> robust path hand-over: the fast kernel appends voxels it gives up on,
  // the robust kernel (work_list != nullptr) processes exactly work_count[0] of them
  unsigned long long *redo_count;
  int *redo_list;
  int redo_base = 0;     // added to the voxel index an entry records (host pipeline: chunk offset)
  const int *work_list;
  const unsigned long long *work_count;
  // a work-list launch only runs when work_count lies in [work_min, work_max] (lets the host
  // enqueue two differently shaped launches and have the device pick one without a sync)
  unsigned long long work_min, work_max;
  // fast path, certification: a bin whose dual is positive but below the rounding threshold at the
  // polished point would enter with the coefficient dual / pivot^2; above cert_ztol the voxel is
  // handed to the robust path (see nnls_v3_kernel, PH_CHECK)
  double cert_ztol;
  // fast path: h = B^T y of every voxel when it has been materialised by the tensor-core GEMM
  // (pnb_nnls_gemm.cuh); nullptr: the kernel computes it in its first dual pass
  const double *h0;
  int screen;  // fast path: FP32 screening of the dual pass (1 = on; exact either way, see nnls_v3_kernel)
};

// column-major packed lower triangle with leading dimension ld: element (i, c), i >= c
__device__ __forceinline__ int lpos(int i, int c, int ld) { return c * ld - (c * (c - 1)) / 2 + (i - c); }

__device__ __forceinline__ double warp_sum(double v) {
#pragma unroll
  for (int o = 16; o > 0; o >>= 1) v += __shfl_xor_sync(0xffffffffu, v, o);
  return v;
}

template <int WARPS>
__global__ void __launch_bounds__(WARPS * 32) nnls_kernel(const NnlsDeviceArgs a) {
  extern __shared__ double smem[];
  const int m = a.m, n = a.n, W = a.W, BW = 2 * a.W + 1;
  const int lane = threadIdx.x & 31, wid = threadIdx.x >> 5;
  const int kmax = a.kmax;
  const unsigned FULL = 0xffffffffu;
  // CTA-shared: basis and banded regulariser Gram
  double *B_s = smem;
  double *rtr_s = B_s + (size_t)m * n;
  const int ltri = kmax * (kmax + 1) / 2;
  const int per_warp = 3 * n + 2 * m + ltri + 5 * kmax + (n + 1) / 2 + 2;
  double *base = rtr_s + (size_t)n * BW + (size_t)wid * per_warp;
  double *xs = base, *ws = xs + n, *hs = ws + n, *ys = hs + n, *rs = ys + m;
  double *Ls_sm = rs + m;
  double *vec_sm = Ls_sm + ltri;                       // us, zs, gs, ls, dinv (kmax each)
  int *P = reinterpret_cast<int *>(vec_sm + 5 * kmax);  // active list, n ints
  for (int i = threadIdx.x; i < m * n; i += WARPS * 32) B_s[i] = a.B[i];
  for (int i = threadIdx.x; i < n * BW; i += WARPS * 32) rtr_s[i] = a.rtr[i];
  __syncthreads();
  if (a.work_list) {
    const unsigned long long cnt = a.work_count[0];
    if (cnt < a.work_min || cnt > a.work_max) return;
  }
  const long long gwarp = (long long)blockIdx.x * WARPS + wid;
  double *scratch = a.scratch + gwarp * ((size_t)n * (n + 1) / 2 + 5 * (size_t)n);

  for (;;) {
    unsigned long long vq = 0;
    if (lane == 0) vq = atomicAdd(a.counter, 1ULL);
    long long vox = (long long)__shfl_sync(FULL, vq, 0);
    if (a.work_list) {
      if (vox >= (long long)a.work_count[0]) break;
      vox = a.work_list[vox];
    } else if (vox >= a.n_vox) {
      break;
    }

    // ---- set-up: y, h = B^T y, w = h, x = 0 ----------------------------------
    bool fin = true;
    for (int b = lane; b < m; b += 32) {
      const double v = a.y[vox * m + b];
      ys[b] = v;
      fin = fin && finite_d(v);
    }
    fin = __all_sync(FULL, fin);
    __syncwarp();
    for (int j = lane; j < n; j += 32) {
      double acc = 0.0;
      for (int b = 0; b < m; b++) acc += B_s[b * n + j] * ys[b];
      hs[j] = acc; ws[j] = acc; xs[j] = 0.0;
    }
    unsigned inP = 0;  // bit q: bin lane + 32 q is in P (n <= 1024)
    int k = 0, iter = 0, mode = fin ? 1 : 2;
    double *Lp = Ls_sm;
    int ld = kmax, kcap = kmax;
    double *us = vec_sm, *zs = us + kmax, *gs = zs + kmax, *ls = gs + kmax, *dinv = ls + kmax;
    __syncwarp();

    // back substitution  L^T z = u  (column sweeps, pivot broadcast through shared memory)
    auto back_solve = [&](const double *rhs, double *z) {
      for (int i = lane; i < k; i += 32) z[i] = rhs[i];
      __syncwarp();
      for (int r = k - 1; r >= 0; r--) {
        const double zr = z[r] * dinv[r];
        __syncwarp();
        if (lane == 0) z[r] = zr;
        for (int i = lane; i < r; i += 32) z[i] -= Lp[lpos(r, i, ld)] * zr;
        __syncwarp();
      }
    };
    // forward substitution  L t = rhs  (rhs destroyed), result in out
    auto fwd_solve = [&](double *rhs, double *out, double &sumsq) {
      double ss = 0.0;
      for (int c = 0; c < k; c++) {
        const double lc = rhs[c] * dinv[c];
        ss += lc * lc;
        if (lane == 0) out[c] = lc;
        for (int i = c + 1 + lane; i < k; i += 32) rhs[i] -= Lp[lpos(i, c, ld)] * lc;
        __syncwarp();
      }
      sumsq = ss;
    };
    // delete position q from the active set (L&H: move index to Z, Givens re-triangularisation)
    auto remove_at = [&](int q) {
      const int idx = P[q];
      if ((idx & 31) == lane) inP &= ~(1u << (idx >> 5));
      if (lane == 0) xs[idx] = 0.0;
      // columns c < q: rows below q move up by one
      for (int c = 0; c < q; c++) {
        for (int i0 = q; i0 < k - 1; i0 += 32) {
          const int i = i0 + lane;
          double v = 0.0;
          if (i < k - 1) v = Lp[lpos(i + 1, c, ld)];
          __syncwarp();
          if (i < k - 1) Lp[lpos(i, c, ld)] = v;
          __syncwarp();
        }
      }
      // columns i >= q: rotate (i, i+1) so that the bump (row i, column i+1) vanishes
      for (int i = q; i < k - 1; i++) {
        // new row i is old row i+1: diagonal candidate (i+1, i) and bump (i+1, i+1)
        const double av = Lp[lpos(i + 1, i, ld)], bv = Lp[lpos(i + 1, i + 1, ld)];
        double c, s, sig;
        if (fabs(av) > fabs(bv)) {
          const double xr = bv / av, yr = sqrt(1.0 + xr * xr);
          c = copysign(1.0 / yr, av); s = c * xr; sig = fabs(av) * yr;
        } else if (bv != 0.0) {
          const double xr = av / bv, yr = sqrt(1.0 + xr * xr);
          s = copysign(1.0 / yr, bv); c = s * xr; sig = fabs(bv) * yr;
        } else {
          sig = 0.0; c = 0.0; s = 1.0;
        }
        __syncwarp();
        // rows r > i (new numbering) <-> old rows r+1 >= i+2
        for (int r0 = i + 1; r0 < k - 1; r0 += 32) {
          const int r = r0 + lane;
          double xv = 0.0, yv = 0.0;
          if (r < k - 1) { xv = Lp[lpos(r + 1, i, ld)]; yv = Lp[lpos(r + 1, i + 1, ld)]; }
          __syncwarp();
          if (r < k - 1) {
            Lp[lpos(r, i, ld)] = c * xv + s * yv;          // final value, new row numbering
            Lp[lpos(r + 1, i + 1, ld)] = -s * xv + c * yv; // updated column i+1, still old numbering
          }
          __syncwarp();
        }
        if (lane == 0) {
          Lp[lpos(i, i, ld)] = sig;
          dinv[i] = 1.0 / sig;
          const double ui = us[i], uj = us[i + 1];
          us[i] = c * ui + s * uj;
          us[i + 1] = -s * ui + c * uj;
        }
        __syncwarp();
      }
      for (int i0 = q; i0 < k - 1; i0 += 32) {
        const int i = i0 + lane;
        int v = 0;
        if (i < k - 1) v = P[i + 1];
        __syncwarp();
        if (i < k - 1) P[i] = v;
        __syncwarp();
      }
      k -= 1;
      __syncwarp();
    };

    while (mode == 1 && k < n) {
      // ---- pick the candidate with the largest positive dual --------------------
      bool accepted = false;
      int jsel = -1;
      double a_new = 0.0, t_new = 0.0;
      for (;;) {
        double best = 0.0;
        int bj = -1;
        for (int j = lane, q = 0; j < n; j += 32, q++) {
          const double v = ws[j];
          if (!((inP >> q) & 1u) && v > best) { best = v; bj = j; }
        }
#pragma unroll
        for (int o = 16; o > 0; o >>= 1) {
          const double ob = __shfl_xor_sync(FULL, best, o);
          const int oj = __shfl_xor_sync(FULL, bj, o);
          if (oj >= 0 && (bj < 0 || ob > best || (ob == best && oj < bj))) { best = ob; bj = oj; }
        }
        if (bj < 0) break;  // wmax <= 0: Kuhn-Tucker conditions hold
        const int j = bj;
        // g = G[P, j],  g_jj
        for (int i = lane; i < k; i += 32) {
          const int p = P[i];
          double acc = 0.0;
          for (int b = 0; b < m; b++) acc += B_s[b * n + p] * B_s[b * n + j];
          const int d = j - p;
          if (d >= -W && d <= W) acc += rtr_s[p * BW + d + W];
          gs[i] = acc;
        }
        double gjj = rtr_s[j * BW + W];
        for (int b = 0; b < m; b++) gjj += B_s[b * n + j] * B_s[b * n + j];
        __syncwarp();
        double unorm2;
        fwd_solve(gs, ls, unorm2);
        const double piv2 = gjj - unorm2;
        const double av = sqrt(piv2 > 0.0 ? piv2 : 0.0);
        const double unorm = sqrt(unorm2);
        bool ok = ((unorm + av * 0.01) - unorm) > 0.0;
        double t = 0.0;
        if (ok) {
          double part = 0.0;
          for (int c = lane; c < k; c += 32) part += ls[c] * us[c];
          t = (hs[j] - warp_sum(part)) / av;
          ok = (t / av) > 0.0;
        }
        if (ok) { accepted = true; jsel = j; a_new = av; t_new = t; break; }
        if (lane == 0) ws[j] = 0.0;
        __syncwarp();
      }
      if (!accepted) break;
      // ---- move jsel into P: append a row to the factor -----------------------
      if (k == kcap) {
        // the shared-memory factor is full: continue in the per-warp global scratch
        double *Lg = scratch;
        double *vg = scratch + (size_t)n * (n + 1) / 2;
        for (int c = 0; c < k; c++)
          for (int i = c + lane; i < k; i += 32) Lg[lpos(i, c, n)] = Lp[lpos(i, c, ld)];
        for (int i = lane; i < k; i += 32) { vg[i] = us[i]; vg[4 * n + i] = dinv[i]; vg[3 * n + i] = ls[i]; }
        __syncwarp();
        Lp = Lg; ld = n; kcap = n;
        us = vg; zs = vg + n; gs = vg + 2 * n; ls = vg + 3 * n; dinv = vg + 4 * n;
        __syncwarp();
      }
      for (int c = lane; c < k; c += 32) Lp[lpos(k, c, ld)] = ls[c];
      if (lane == 0) {
        Lp[lpos(k, k, ld)] = a_new;
        dinv[k] = 1.0 / a_new;
        us[k] = t_new;
        P[k] = jsel;
        ws[jsel] = 0.0;
      }
      if ((jsel & 31) == lane) inP |= 1u << (jsel >> 5);
      k += 1;
      __syncwarp();
      // ---- secondary loop: keep the least-squares solution on P feasible -----------
      back_solve(us, zs);
      for (;;) {
        iter += 1;
        if (iter >= a.maxiter) { mode = 3; break; }  // SciPy fails once the count reaches maxiter
        double alpha = 2.0;
        int jj = -1;
        for (int i = lane; i < k; i += 32) {
          const double z = zs[i];
          if (z <= 0.0) {
            const double xv = xs[P[i]];
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
        if (jj < 0) break;
        for (int i = lane; i < k; i += 32) {
          const int p = P[i];
          xs[p] += alpha * (zs[i] - xs[p]);
        }
        __syncwarp();
        remove_at(jj);
        for (;;) {  // round-off guard of L&H: every coefficient left in P must be positive
          int bad = n;
          for (int i = lane; i < k; i += 32)
            if (xs[P[i]] <= 0.0 && i < bad) bad = i;
#pragma unroll
          for (int o = 16; o > 0; o >>= 1) bad = min(bad, __shfl_xor_sync(FULL, bad, o));
          if (bad >= n) break;
          remove_at(bad);
        }
        back_solve(us, zs);
      }
      if (mode != 1) break;
      for (int i = lane; i < k; i += 32) xs[P[i]] = zs[i];
      __syncwarp();
      // ---- dual  w = B^T (y - B x) - mu^2 R^T R x  --------------------------------
      for (int b = lane; b < m; b += 32) {
        double acc = ys[b];
        for (int i = 0; i < k; i++) acc -= B_s[b * n + P[i]] * zs[i];
        rs[b] = acc;
      }
      __syncwarp();
      for (int j = lane; j < n; j += 32) {
        double acc = 0.0;
        for (int b = 0; b < m; b++) acc += B_s[b * n + j] * rs[b];
        const int lo = (j - W < 0) ? 0 : j - W, hi = (j + W > n - 1) ? n - 1 : j + W;
        for (int jn = lo; jn <= hi; jn++) acc -= rtr_s[j * BW + (jn - j) + W] * xs[jn];
        ws[j] = acc;
      }
      __syncwarp();
      // ---- one refinement step on P with the true residual ------------------------------
      for (int i = lane; i < k; i += 32) gs[i] = ws[P[i]];
      __syncwarp();
      double dummy;
      fwd_solve(gs, ls, dummy);
      back_solve(ls, gs);
      bool pos = true;
      for (int i = lane; i < k; i += 32) pos = pos && (zs[i] + gs[i] > 0.0);
      pos = __all_sync(FULL, pos);
      for (int i = lane; i < k; i += 32) {
        const int p = P[i];
        if (pos) xs[p] = zs[i] + gs[i];
        ws[p] = 0.0;
      }
      __syncwarp();
    }

    // ---- results -------------------------------------------------------------------------
    double *out = a.coef + vox * (long long)n;
    if (mode == 1) {
      for (int b = lane; b < m; b += 32) {
        double acc = ys[b];
        for (int i = 0; i < k; i++) { const int p = P[i]; acc -= B_s[b * n + p] * xs[p]; }
        rs[b] = acc;
      }
      __syncwarp();
      double part = 0.0;
      for (int b = lane; b < m; b += 32) part += rs[b] * rs[b];
      for (int j = lane; j < n; j += 32) {
        const double xj = xs[j];
        out[j] = xj;
        if (xj != 0.0) {
          const int lo = (j - W < 0) ? 0 : j - W, hi = (j + W > n - 1) ? n - 1 : j + W;
          double acc = 0.0;
          for (int jn = lo; jn <= hi; jn++) acc += rtr_s[j * BW + (jn - j) + W] * xs[jn];
          part += xj * acc;
        }
      }
      double top = 0.0;
      for (int b = lane; b < m; b += 32) top += rs[b] * rs[b];
      top = warp_sum(top);
      const double tot = warp_sum(part);
      if (lane == 0) a.rnorm[vox] = sqrt(tot > 0.0 ? tot : 0.0);
      if (a.r2) {
        double sm = 0.0;
        for (int b = lane; b < m; b += 32) sm += ys[b];
        const double mean = warp_sum(sm) / (double)m;
        double st = 0.0;
        for (int b = lane; b < m; b += 32) { const double d = ys[b] - mean; st += d * d; }
        st = warp_sum(st);
        if (lane == 0) a.r2[vox] = (st > 0.0) ? 1.0 - top / st : nan("");
      }
    } else {
      // failure (nnls_solver.py:201-210): zeros and ||[y; 0]||
      double part = 0.0;
      for (int b = lane; b < m; b += 32) part += ys[b] * ys[b];
      for (int j = lane; j < n; j += 32) out[j] = 0.0;
      const double tot = warp_sum(part);
      if (lane == 0) a.rnorm[vox] = sqrt(tot);
      if (a.r2) {  // prediction is zero
        double sm = 0.0;
        for (int b = lane; b < m; b += 32) sm += ys[b];
        const double mean = warp_sum(sm) / (double)m;
        double st = 0.0;
        for (int b = lane; b < m; b += 32) { const double d = ys[b] - mean; st += d * d; }
        st = warp_sum(st);
        if (lane == 0) a.r2[vox] = (st > 0.0) ? 1.0 - tot / st : nan("");
      }
    }
    if (lane == 0) { a.status[vox] = mode; a.iters[vox] = iter; }
    __syncwarp();
  }
}

inline size_t nnls_smem_bytes(int m, int n, int W, int kmax, int warps) {
  const size_t per_warp = 3 * (size_t)n + 2 * m + (size_t)kmax * (kmax + 1) / 2 + 5 * (size_t)kmax + (n + 1) / 2 + 2;
  return sizeof(double) * ((size_t)m * n + (size_t)n * (2 * W + 1) + warps * per_warp);
}

}  // namespace pnb
