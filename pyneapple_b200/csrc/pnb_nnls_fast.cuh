// Fast path of the batched NNLS: the same Lawson-Hanson iteration as
// pnb_nnls_kernel.cuh (same candidate rule, independence / z test, line search,
// removal rule, iteration counter), but the active block is carried as its
// explicit inverse  H = (G_PP)^-1  instead of a Cholesky factor:
//
//   column j enters:   v = H g,  s = G_jj - g.v  (= L&H's pivot^2),  zeta = (h_j - g.z) / s
//                      H <- [H + v v^T / s, -v / s; -v^T / s, 1 / s],   z <- [z - zeta v; zeta]
//   slot q leaves:     H <- H - H_q H_q^T / H_qq  (row / column q dropped),  z <- z - H_q z_q / H_qq
//
// Every step is a rank-one update or a matrix-vector product that all 32 lanes
// work on at once: no triangular solves, i.e. none of the k-step dependent
// chains (one shared-memory round trip and one warp barrier per step) that make
// the Cholesky kernel latency bound, and slots need no ordering, so a leaving
// variable is replaced by the last slot instead of shifting a factor.
//
// Updating an inverse is less stable than updating a factor.  For the
// regularised dictionaries this path is meant for (cond(G_PP) ~ 1e6..1e8) the
// solution is polished with up to four steps of iterative refinement against the true
// residual at the end and agrees with SciPy to ~2e-11 with identical iteration
// counts (scripts/proto_nnls_inverse.py).  A voxel is handed to the robust
// Cholesky kernel instead (status kNnlsRedo) when
//   * its active set outgrows the shared-memory inverse,
//   * the first polish step moves the solution by more than 5e-5 relative or the
//     refinement has not converged to 1e-10 after four steps, or
//   * the polished point violates the Kuhn-Tucker conditions,
// which is what happens for weakly / un-regularised problems (mu <~ 5e-4).
#pragma once
#include "pnb_nnls_kernel.cuh"

namespace pnb {

constexpr int kNnlsRedo = 4;  // status: re-run this voxel with the robust kernel

__device__ __forceinline__ int hpos(int i, int c) { return (i * (i + 1)) / 2 + c; }  // i >= c

template <int WARPS>
__global__ void __launch_bounds__(WARPS * 32) nnls_fast_kernel(const NnlsDeviceArgs a) {
  extern __shared__ double smem[];
  const int m = a.m, n = a.n, W = a.W, BW = 2 * a.W + 1;
  const int lane = threadIdx.x & 31, wid = threadIdx.x >> 5;
  const int kcap = a.kmax;
  const unsigned FULL = 0xffffffffu;
  double *B_s = smem;
  double *rtr_s = B_s + (size_t)m * n;
  const int htri = kcap * (kcap + 1) / 2;
  const int per_warp = 2 * n + 2 * m + htri + 3 * kcap + (kcap + 1) / 2 + 2;
  double *base = rtr_s + (size_t)n * BW + (size_t)wid * per_warp;
  double *xs = base, *ws = xs + n, *ys = ws + n, *rs = ys + m;
  double *H = rs + m;
  double *zs = H + htri, *gs = zs + kcap, *vs = gs + kcap;
  int *P = reinterpret_cast<int *>(vs + kcap);
  for (int i = threadIdx.x; i < m * n; i += WARPS * 32) B_s[i] = a.B[i];
  for (int i = threadIdx.x; i < n * BW; i += WARPS * 32) rtr_s[i] = a.rtr[i];
  __syncthreads();

  for (;;) {
    unsigned long long vq = 0;
    if (lane == 0) vq = atomicAdd(a.counter, 1ULL);
    const long long vox = (long long)__shfl_sync(FULL, vq, 0);
    if (vox >= a.n_vox) break;

    bool fin = true;
    for (int b = lane; b < m; b += 32) {
      const double v = a.y[vox * m + b];
      ys[b] = v;
      fin = fin && finite_d(v);
    }
    fin = __all_sync(FULL, fin);
    __syncwarp();
    double hmax = 0.0;
    for (int j = lane; j < n; j += 32) {
      double acc = 0.0;
      for (int b = 0; b < m; b++) acc += B_s[b * n + j] * ys[b];
      ws[j] = acc; xs[j] = 0.0;
      hmax = fmax(hmax, fabs(acc));
    }
#pragma unroll
    for (int o = 16; o > 0; o >>= 1) hmax = fmax(hmax, __shfl_xor_sync(FULL, hmax, o));
    unsigned inP = 0;
    int k = 0, iter = 0, mode = fin ? 1 : 2;
    __syncwarp();

    // dst = H src over the k active slots (symmetric packed storage, two accumulators per row)
    auto matvec = [&](const double *src, double *dst) {
      for (int i = lane; i < k; i += 32) {
        double a0 = 0.0, a1 = 0.0, a2 = 0.0, a3 = 0.0;
        const int rb = (i * (i + 1)) / 2;
        int c = 0;
        for (; c + 3 <= i; c += 4) {
          a0 += H[rb + c] * src[c]; a1 += H[rb + c + 1] * src[c + 1];
          a2 += H[rb + c + 2] * src[c + 2]; a3 += H[rb + c + 3] * src[c + 3];
        }
        for (; c <= i; c++) a0 += H[rb + c] * src[c];
        int idx = ((i + 1) * (i + 2)) / 2 + i;
        int r = i + 1;
        for (; r + 1 < k; r += 2) {
          a1 += H[idx] * src[r];
          a2 += H[idx + r + 1] * src[r + 1];
          idx += 2 * r + 3;
        }
        if (r < k) a3 += H[idx] * src[r];
        dst[i] = (a0 + a1) + (a2 + a3);
      }
      __syncwarp();
    };
    // drop slot q: rank-one downdate of H and z, then the last slot takes its place
    auto remove_at = [&](int q) {
      const int idx = P[q];
      if ((idx & 31) == lane) inP &= ~(1u << (idx >> 5));
      const double dq = H[hpos(q, q)], zq = zs[q];
      __syncwarp();
      for (int i = lane; i < k; i += 32) gs[i] = (i >= q) ? H[hpos(i, q)] : H[hpos(q, i)];
      if (lane == 0) xs[idx] = 0.0;
      __syncwarp();
      const double dinv = 1.0 / dq;
      for (int i = lane; i < k; i += 32) {
        const double gi = gs[i] * dinv;
        const int rb = (i * (i + 1)) / 2;
#pragma unroll 4
        for (int c = 0; c <= i; c++) H[rb + c] -= gi * gs[c];
        zs[i] -= gi * zq;
      }
      __syncwarp();
      const int last = k - 1;
      if (q != last) {
        for (int c = lane; c < last; c += 32) {
          if (c == q) continue;
          const double v = H[hpos(last, c)];
          if (c < q) H[hpos(q, c)] = v; else H[hpos(c, q)] = v;
        }
        if (lane == 0) { H[hpos(q, q)] = H[hpos(last, last)]; zs[q] = zs[last]; P[q] = P[last]; }
      }
      k -= 1;
      __syncwarp();
    };
    // r = y - B_P z ;  returns nothing, rs filled
    auto residual = [&]() {
      if (m <= 16) {
        const int b = lane & 15, half = lane >> 4;
        const int mid = (k + 1) >> 1;
        const int i0 = half ? mid : 0, i1 = half ? k : mid;
        double a0 = 0.0, a1 = 0.0;
        if (b < m) {
          int i = i0;
          for (; i + 1 < i1; i += 2) {
            a0 += B_s[b * n + P[i]] * zs[i];
            a1 += B_s[b * n + P[i + 1]] * zs[i + 1];
          }
          if (i < i1) a0 += B_s[b * n + P[i]] * zs[i];
        }
        double acc = a0 + a1;
        acc += __shfl_xor_sync(FULL, acc, 16);
        if (half == 0 && b < m) rs[b] = ys[b] - acc;
      } else {
        for (int b = lane; b < m; b += 32) {
          double a0 = 0.0, a1 = 0.0;
          int i = 0;
          for (; i + 1 < k; i += 2) {
            a0 += B_s[b * n + P[i]] * zs[i];
            a1 += B_s[b * n + P[i + 1]] * zs[i + 1];
          }
          if (i < k) a0 += B_s[b * n + P[i]] * zs[i];
          rs[b] = ys[b] - (a0 + a1);
        }
      }
      __syncwarp();
    };
    // duals of four bins of this lane at once (independent accumulation chains)
    auto dual4 = [&](int j0, double (&w4)[4]) {
      double a[4] = {0.0, 0.0, 0.0, 0.0};
      int jj[4];
#pragma unroll
      for (int t = 0; t < 4; t++) jj[t] = (j0 + 32 * t < n) ? j0 + 32 * t : j0;
#pragma unroll 4
      for (int b = 0; b < m; b++) {
        const double rb = rs[b];
        const double *row = B_s + b * n;
#pragma unroll
        for (int t = 0; t < 4; t++) a[t] += row[jj[t]] * rb;
      }
#pragma unroll
      for (int t = 0; t < 4; t++) {
        const int j = jj[t];
        const int lo = (j - W < 0) ? 0 : j - W, hi = (j + W > n - 1) ? n - 1 : j + W;
        double acc = 0.0;
        for (int jn = lo; jn <= hi; jn++) acc += rtr_s[j * BW + (jn - j) + W] * xs[jn];
        w4[t] = a[t] - acc;
      }
    };
    auto dual_of = [&](int j) -> double {
      double a0 = 0.0, a1 = 0.0;
      int b = 0;
      for (; b + 1 < m; b += 2) { a0 += B_s[b * n + j] * rs[b]; a1 += B_s[(b + 1) * n + j] * rs[b + 1]; }
      if (b < m) a0 += B_s[b * n + j] * rs[b];
      const int lo = (j - W < 0) ? 0 : j - W, hi = (j + W > n - 1) ? n - 1 : j + W;
      for (int jn = lo; jn <= hi; jn++) a1 -= rtr_s[j * BW + (jn - j) + W] * xs[jn];
      return a0 + a1;
    };

    while (mode == 1 && k < n) {
      bool accepted = false;
      int jsel = -1;
      double s_new = 0.0, zeta = 0.0;
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
        if (bj < 0) break;
        const int j = bj;
        for (int i = lane; i < k; i += 32) {
          const int p = P[i];
          double a0 = 0.0, a1 = 0.0, a2 = 0.0, a3 = 0.0;
          int b = 0;
          for (; b + 3 < m; b += 4) {
            a0 += B_s[b * n + p] * B_s[b * n + j];
            a1 += B_s[(b + 1) * n + p] * B_s[(b + 1) * n + j];
            a2 += B_s[(b + 2) * n + p] * B_s[(b + 2) * n + j];
            a3 += B_s[(b + 3) * n + p] * B_s[(b + 3) * n + j];
          }
          for (; b < m; b++) a0 += B_s[b * n + p] * B_s[b * n + j];
          double acc = (a0 + a1) + (a2 + a3);
          const int d = j - p;
          if (d >= -W && d <= W) acc += rtr_s[p * BW + d + W];
          gs[i] = acc;
        }
        double gjj = rtr_s[j * BW + W], hj = 0.0;
        for (int b = 0; b < m; b++) { const double bj2 = B_s[b * n + j]; gjj += bj2 * bj2; hj += bj2 * ys[b]; }
        __syncwarp();
        matvec(gs, vs);
        double p0 = 0.0, p1 = 0.0;
        for (int i = lane; i < k; i += 32) { p0 += gs[i] * vs[i]; p1 += gs[i] * zs[i]; }
#pragma unroll
        for (int o = 16; o > 0; o >>= 1) {
          p0 += __shfl_xor_sync(FULL, p0, o);
          p1 += __shfl_xor_sync(FULL, p1, o);
        }
        const double unorm2 = p0 > 0.0 ? p0 : 0.0;
        const double piv2 = gjj - unorm2;
        const double av = sqrt(piv2 > 0.0 ? piv2 : 0.0), unorm = sqrt(unorm2);
        bool ok = ((unorm + av * 0.01) - unorm) > 0.0;
        double zt = 0.0;
        if (ok) { zt = (hj - p1) / piv2; ok = zt > 0.0; }
        if (ok) { accepted = true; jsel = j; s_new = piv2; zeta = zt; break; }
        if (lane == 0) ws[j] = 0.0;
        __syncwarp();
      }
      if (!accepted) break;
      if (k == kcap) { mode = kNnlsRedo; break; }
      // ---- bordering update ----------------------------------------------------
      {
        const double sinv = 1.0 / s_new;
        for (int i = lane; i < k; i += 32) {
          const double vi = vs[i] * sinv;
          const int rb = (i * (i + 1)) / 2;
#pragma unroll 4
          for (int c = 0; c <= i; c++) H[rb + c] += vi * vs[c];
          zs[i] -= vs[i] * zeta;
        }
        const int rb = (k * (k + 1)) / 2;
        for (int c = lane; c < k; c += 32) H[rb + c] = -vs[c] * sinv;
        if (lane == 0) { H[rb + k] = sinv; zs[k] = zeta; P[k] = jsel; ws[jsel] = 0.0; }
        if ((jsel & 31) == lane) inP |= 1u << (jsel >> 5);
        k += 1;
        __syncwarp();
      }
      // ---- secondary loop ----------------------------------------------------------
      for (;;) {
        iter += 1;
        if (iter >= a.maxiter) { mode = 3; break; }
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
        for (;;) {
          int bad = n;
          for (int i = lane; i < k; i += 32)
            if (xs[P[i]] <= 0.0 && i < bad) bad = i;
#pragma unroll
          for (int o = 16; o > 0; o >>= 1) bad = min(bad, __shfl_xor_sync(FULL, bad, o));
          if (bad >= n) break;
          remove_at(bad);
        }
      }
      if (mode != 1) break;
      for (int i = lane; i < k; i += 32) xs[P[i]] = zs[i];
      __syncwarp();
      residual();
      for (int j0 = lane, q0 = 0; j0 < n; j0 += 128, q0 += 4) {
        double w4[4];
        dual4(j0, w4);
#pragma unroll
        for (int t = 0; t < 4; t++)
          if (j0 + 32 * t < n) ws[j0 + 32 * t] = ((inP >> (q0 + t)) & 1u) ? 0.0 : w4[t];
      }
      __syncwarp();
    }

    // ---- polish: two refinement steps with the true residual, then verify ------------
    if (mode == 1 && k > 0) {
      double rel = 1.0;
      for (int pass = 0; pass < 4 && mode == 1; pass++) {
        residual();
        for (int i = lane; i < k; i += 32) gs[i] = dual_of(P[i]);
        __syncwarp();
        matvec(gs, vs);
        bool pos = true;
        double dmax = 0.0, zmax = 0.0;
        for (int i = lane; i < k; i += 32) {
          pos = pos && (zs[i] + vs[i] > 0.0);
          dmax = fmax(dmax, fabs(vs[i]));
          zmax = fmax(zmax, fabs(zs[i]));
        }
        pos = __all_sync(FULL, pos);
#pragma unroll
        for (int o = 16; o > 0; o >>= 1) {
          dmax = fmax(dmax, __shfl_xor_sync(FULL, dmax, o));
          zmax = fmax(zmax, __shfl_xor_sync(FULL, zmax, o));
        }
        rel = dmax / zmax;
        // the first correction measures how far the carried solution had drifted while the
        // active-set decisions were being made; wrong results only appeared above 3e-4
        if ((pass == 0 && rel > 5e-5) || !pos) { mode = kNnlsRedo; break; }
        for (int i = lane; i < k; i += 32) { zs[i] += vs[i]; xs[P[i]] = zs[i]; }
        __syncwarp();
        if (rel < 1e-13) break;
      }
      if (mode == 1 && rel > 1e-10) mode = kNnlsRedo;  // refinement did not converge
      if (mode == 1) {
        residual();
        double wmax = 0.0;
        for (int j0 = lane, q0 = 0; j0 < n; j0 += 128, q0 += 4) {
          double w4[4];
          dual4(j0, w4);
#pragma unroll
          for (int t = 0; t < 4; t++)
            if (j0 + 32 * t < n && !((inP >> (q0 + t)) & 1u)) wmax = fmax(wmax, w4[t]);
        }
#pragma unroll
        for (int o = 16; o > 0; o >>= 1) wmax = fmax(wmax, __shfl_xor_sync(FULL, wmax, o));
        if (wmax > 1e-12 * hmax) mode = kNnlsRedo;  // not (to rounding) a Kuhn-Tucker point
      }
    }

    double *out = a.coef + vox * (long long)n;
    if (mode == 1) {
      if (k == 0) residual();
      double part = 0.0, top = 0.0;
      for (int b = lane; b < m; b += 32) top += rs[b] * rs[b];
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
      top = warp_sum(top);
      const double tot = top + warp_sum(part);
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
    } else if (mode != kNnlsRedo) {
      double part = 0.0;
      for (int b = lane; b < m; b += 32) part += ys[b] * ys[b];
      for (int j = lane; j < n; j += 32) out[j] = 0.0;
      const double tot = warp_sum(part);
      if (lane == 0) a.rnorm[vox] = sqrt(tot);
      if (a.r2) {
        double sm = 0.0;
        for (int b = lane; b < m; b += 32) sm += ys[b];
        const double mean = warp_sum(sm) / (double)m;
        double st = 0.0;
        for (int b = lane; b < m; b += 32) { const double d = ys[b] - mean; st += d * d; }
        st = warp_sum(st);
        if (lane == 0) a.r2[vox] = (st > 0.0) ? 1.0 - tot / st : nan("");
      }
    }
    if (lane == 0) {
      a.status[vox] = mode;
      a.iters[vox] = iter;
      if (mode == kNnlsRedo) {
        const unsigned long long slot = atomicAdd(a.redo_count, 1ULL);
        a.redo_list[slot] = (int)vox;
      }
    }
    __syncwarp();
  }
}

inline size_t nnls_fast_smem_bytes(int m, int n, int W, int kcap, int warps) {
  const size_t per_warp = 2 * (size_t)n + 2 * m + (size_t)kcap * (kcap + 1) / 2 + 3 * (size_t)kcap + (kcap + 1) / 2 + 2;
  return sizeof(double) * ((size_t)m * n + (size_t)n * (2 * W + 1) + warps * per_warp);
}

}  // namespace pnb
