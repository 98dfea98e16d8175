// Per-voxel bounded Trust-Region-Reflective least squares, register resident.
//
// One *lane* (one CUDA thread) owns one voxel.  The algorithm is SciPy's
// `trf_bounds` with `tr_solver='exact'`, `loss='linear'` — the routine the
// reference reaches through solvers/curvefit.py:295-306 ->
// scipy.optimize.curve_fit -> least_squares(method='trf') — restated so that
// nothing larger than N x N (N = number of model parameters, <= 7) is ever
// stored:
//
//   * SciPy forms the (m+n) x n augmented Jacobian J_aug = [J D; diag(C)^1/2]
//     and takes its SVD each iteration (scipy/optimize/_lsq/trf.py, the
//     `tr_solver == 'exact'` branch).  Every quantity it derives from that SVD
//     is a function of the n x n matrix  B = J_aug^T J_aug = D (J^T J) D + C
//     and the vector g_h = D J^T f:
//        Gauss-Newton step            p    = -B^-1 g_h
//        Levenberg-Marquardt step     p(a) = -(B + a I)^-1 g_h
//        phi(a)  = ||p(a)|| - Delta,  phi'(a) = -(p^T (B + a I)^-1 p) / ||p||
//        ||s * uf|| (alpha upper bound) = ||g_h||
//        evaluate_quadratic / build_quadratic_1d:  ||J_h s||^2 + s^T C s = s^T B s
//     so the lane accumulates J^T J and J^T f while it streams over the
//     b-values (the Jacobian itself is never materialised) and solves the
//     trust-region sub-problem with LDL^T factorisations of B + a I held in
//     registers.  LDL^T is invariant under the symmetric diagonal scaling
//     that makes J badly conditioned here (x_scale = 1 with S0 ~ 1e3 and
//     D ~ 1e-3), so no accuracy is lost to that scaling.
//   * the control flow (Coleman-Li scaling, More iteration on alpha carried
//     between iterations, reflective / gradient candidate steps, radius and
//     termination rules, nfev accounting) follows trf.py / common.py line
//     by line; see the function comments.
//
// Fixed ("frozen") parameters stay in the vector: their Jacobian column is
// zeroed, their row/column of B is the identity, so every norm and step the
// algorithm sees equals that of the reduced problem SciPy is given by
// solvers/curvefit.py:274-288.
#pragma once
#include "pnb_models.cuh"

// unroll factor of the row loop of the finite-difference evaluation (1 = rolled, the default:
// profiles/r2_trf_experiments.md)
#ifndef PNB_TRF_ROW_UNROLL
#define PNB_TRF_ROW_UNROLL 1
#endif
#define PNB_PRAGMA_(x) _Pragma(#x)
#define PNB_PRAGMA(x) PNB_PRAGMA_(x)
#define PNB_TRF_ROW_PRAGMA PNB_PRAGMA(unroll PNB_TRF_ROW_UNROLL)

namespace pnb {

struct TrfOptions {
  double ftol, xtol, gtol;
  int max_nfev;
  int jac_mode;      // 0 analytic Jacobian, 1 SciPy '2-point' finite differences, 2 MINPACK forward differences (fdjac2)
  int x_scale_jac;   // x_scale == 'jac'
  int method;        // 0 'trf', 1 'dogbox' (pnb_dogbox_core.cuh), 2 'lm' (pnb_lm_core.cuh)
  unsigned frozen;   // bit j set: parameter j is fixed at its p0 value
  double x_scale[8];
  double tr, tm;     // repetition / mixing time of the T1 variants
  int finish_wait = 3;  // kernel scheduling only (trf_kernel: passes a converged lane waits for its warp)
  // curve_fit / least_squares extras (solvers/curvefit.py:305 forwards them)
  double diff_step[8] = {0.0, 0.0, 0.0, 0.0, 0.0, 0.0, 0.0, 0.0};  // relative 2-point step per parameter, 0 = SciPy's default
  int loss = 0;            // 0 linear, 1 soft_l1, 2 huber, 3 cauchy, 4 arctan (EXTRAS evaluation only)
  double f_scale = 1.0;    // soft margin of the robust losses
  int absolute_sigma = 0;  // curve_fit(absolute_sigma=True): the covariance is not scaled by 2 cost / (m - n)
};

// least_squares' robust losses (scipy/optimize/_lsq/least_squares.py: soft_l1, huber, cauchy, arctan and
// construct_loss_function): rho(z), rho'(z), rho''(z) at z = (f / f_scale)^2, with rho scaled by
// f_scale^2 and rho'' by 1 / f_scale^2.
PNB_HD void trf_rho(int loss, double f, double f_scale, double &r0, double &r1, double &r2) {
  const double q = f / f_scale;
  const double z = q * q;
  if (loss == 1) {
    const double t = 1.0 + z, st = sqrt(t);
    r0 = 2.0 * (st - 1.0); r1 = 1.0 / st; r2 = -0.5 / (t * st);
  } else if (loss == 2) {
    if (z <= 1.0) { r0 = z; r1 = 1.0; r2 = 0.0; }
    else { const double sz = sqrt(z); r0 = 2.0 * sz - 1.0; r1 = 1.0 / sz; r2 = -0.5 / (z * sz); }
  } else if (loss == 3) {
    const double t = 1.0 + z;
    r0 = log1p(z); r1 = 1.0 / t; r2 = -1.0 / (t * t);
  } else if (loss == 4) {
    const double t = 1.0 + z * z;
    r0 = atan(z); r1 = 1.0 / t; r2 = -2.0 * z / (t * t);
  } else {
    r0 = z; r1 = 1.0; r2 = 0.0;
  }
  r0 *= f_scale * f_scale;
  r2 /= f_scale * f_scale;
}

enum TrfStatus {
  kStRunning = -99,
  kStBadBounds = -1,   // "Each lower bound must be strictly less than each upper bound."
  kStInfeasible = -2,  // "Initial guess is outside of provided bounds"
  kStNonFiniteY = -3,  // asarray_chkfinite(ydata)
  kStNonFiniteF0 = -4, // "Residuals are not finite in the initial point."
  kStMaxNfev = 0, kStGtol = 1, kStFtol = 2, kStXtol = 3, kStBoth = 4
};

template <int N> PNB_HD double norm2(const double (&v)[N]) {
  double t = 0.0;
#pragma unroll
  for (int i = 0; i < N; i++) t += v[i] * v[i];
  return sqrt(t);
}
template <int N> PNB_HD double dot(const double (&a)[N], const double (&b)[N]) {
  double t = 0.0;
#pragma unroll
  for (int i = 0; i < N; i++) t += a[i] * b[i];
  return t;
}

// s^T B s for symmetric B stored in the lower triangle
template <int N> PNB_HD double quad_form(const double (&B)[N][N], const double (&s)[N]) {
  double q = 0.0;
#pragma unroll
  for (int i = 0; i < N; i++) {
    double t = 0.5 * B[i][i] * s[i];
#pragma unroll
    for (int j = 0; j < i; j++) t += B[i][j] * s[j];
    q += s[i] * t;
  }
  return 2.0 * q;
}
// a^T B b
template <int N>
PNB_HD double bilinear(const double (&B)[N][N], const double (&a)[N], const double (&b)[N]) {
  double q = 0.0;
#pragma unroll
  for (int i = 0; i < N; i++) {
    double t = 0.0;
#pragma unroll
    for (int j = 0; j < N; j++) t += (j <= i ? B[i][j] : B[j][i]) * b[j];
    q += a[i] * t;
  }
  return q;
}

// LDL^T of (B + alpha I).  L unit lower (strict part stored), dinv = 1 / D.
// Returns false when a pivot is not safely positive (numerically rank deficient).
template <int N>
PNB_HD bool ldlt(const double (&B)[N][N], double alpha, double (&L)[N][N], double (&dinv)[N]) {
  bool ok = true;
  double dd[N];  // D itself (dinv holds 1 / D)
#pragma unroll
  for (int j = 0; j < N; j++) {
    const double ajj = B[j][j] + alpha;
    double dj = ajj;
    double ld[N];  // L[j][k] * D[k]
#pragma unroll
    for (int k = 0; k < j; k++) {
      ld[k] = L[j][k] * dd[k];
      dj -= L[j][k] * ld[k];
    }
    if (!(dj > 4.0 * N * kEps * ajj)) { ok = false; dj = (ajj > 0.0) ? ajj : 1.0; }
    dd[j] = dj;
    dinv[j] = 1.0 / dj;
#pragma unroll
    for (int i = j + 1; i < N; i++) {
      double t = B[i][j];
#pragma unroll
      for (int k = 0; k < j; k++) t -= L[i][k] * ld[k];
      L[i][j] = t * dinv[j];
    }
  }
  return ok;
}

// q = (L D L^T)^-1 r ; also returns  q^T (L D L^T)^-1 q  when want_curv (for phi')
template <int N>
PNB_HD void ldlt_solve(const double (&L)[N][N], const double (&dinv)[N], const double (&r)[N],
                       double (&q)[N]) {
  double z[N];
#pragma unroll
  for (int i = 0; i < N; i++) {
    double t = r[i];
#pragma unroll
    for (int k = 0; k < i; k++) t -= L[i][k] * z[k];
    z[i] = t;
  }
#pragma unroll
  for (int i = N - 1; i >= 0; i--) {
    double t = z[i] * dinv[i];
#pragma unroll
    for (int k = i + 1; k < N; k++) t -= L[k][i] * q[k];
    q[i] = t;
  }
}
// q^T (L D L^T)^-1 q = sum_i (L^-1 q)_i^2 / D_i
template <int N>
PNB_HD double ldlt_curv(const double (&L)[N][N], const double (&dinv)[N], const double (&q)[N]) {
  double z[N];
  double c = 0.0;
#pragma unroll
  for (int i = 0; i < N; i++) {
    double t = q[i];
#pragma unroll
    for (int k = 0; k < i; k++) t -= L[i][k] * z[k];
    z[i] = t;
    c += t * t * dinv[i];
  }
  return c;
}

// scipy/optimize/_lsq/common.py: find_active_constraints + make_strictly_feasible
PNB_HD double strictly_feasible(double x, double lb, double ub, double rstep) {
  int act = 0;
  if (rstep == 0.0) {
    if (x <= lb) act = -1;
    if (x >= ub) act = 1;
  } else {
    const double ld = x - lb, ud = ub - x;
    const double lt = rstep * dmax(1.0, fabs(lb)), ut = rstep * dmax(1.0, fabs(ub));
    if (finite_d(lb) && ld <= dmin(ud, lt)) act = -1;
    if (finite_d(ub) && ud <= dmin(ld, ut)) act = 1;
  }
  if (act == -1) x = (rstep == 0.0) ? nextafter(lb, ub) : lb + rstep * dmax(1.0, fabs(lb));
  if (act == 1) x = (rstep == 0.0) ? nextafter(ub, lb) : ub - rstep * dmax(1.0, fabs(ub));
  if (x < lb || x > ub) x = 0.5 * (lb + ub);
  return x;
}

// common.py: minimize_quadratic_1d  (np.argmin picks the first minimum)
PNB_HD double min_quad_1d(double a, double b, double lo, double hi, double c, double &yv) {
  double t = lo;
  double y = lo * (a * lo + b) + c;
  const double yh = hi * (a * hi + b) + c;
  if (yh < y) { y = yh; t = hi; }
  if (a != 0.0) {
    const double ext = -0.5 * b / a;
    if (lo < ext && ext < hi) {
      const double ye = ext * (a * ext + b) + c;
      if (ye < y) { y = ye; t = ext; }
    }
  }
  yv = y;
  return t;
}

// The state one lane carries for its voxel.  Everything is indexed with
// compile-time constants after unrolling, so it lives in registers.
template <class M> struct TrfLane {
  static constexpr int N = M::NP;
  double x[N];        // current iterate (frozen entries hold the fixed value)
  double g[N];        // J^T f
  double A[N][N];     // J^T J (lower triangle)
  double cost;        // 0.5 ||f||^2
  double Delta, alpha;
  double scale_inv[N];
  int nfev, njev, status;
  // hat-space quantities of the current outer iteration
  double d[N], g_h[N], B[N][N];
  double theta;
  // the trial being evaluated
  double x_new[N], step_h_norm, step_norm, predicted;
  bool need_prologue;
};

// common.py: step_size_to_bound; hits returned as a bit mask of +-1 entries (sign not needed
// by the caller beyond "this component hit", see select_step)
template <int N>
PNB_HD double step_to_bound(const double (&x)[N], const double (&s)[N], const double *lb,
                            const double *ub, int lbs, unsigned &hits) {
  double steps[N];
  double mn = kInf;
#pragma unroll
  for (int i = 0; i < N; i++) {
    steps[i] = kInf;
    // max((lb - x) / s, (ub - x) / s) is the quotient whose numerator has the sign of s
    // (lb - x <= 0 <= ub - x; for an x a rounding error outside the box the max still picks the
    // same quotient): one division instead of two, same bits
    if (s[i] != 0.0) steps[i] = ((s[i] > 0.0 ? ub[i * lbs] : lb[i * lbs]) - x[i]) / s[i];
    mn = dmin(mn, steps[i]);
  }
  hits = 0;
#pragma unroll
  for (int i = 0; i < N; i++)
    if (steps[i] == mn && s[i] != 0.0) hits |= 1u << i;
  return mn;
}

// Outer-iteration prologue (top of the `while True` in trf_bounds): Coleman-Li
// scaling, first-order optimality test, hat-space system.  Returns false when
// the voxel terminates here (status set).
template <class M>
PNB_HD bool trf_prologue(TrfLane<M> &S, const TrfOptions &O, const double *lb, const double *ub,
                         int lbs) {
  constexpr int N = M::NP;
  double v[N], dv[N];
  double g_norm = 0.0;
#pragma unroll
  for (int i = 0; i < N; i++) {
    v[i] = 1.0; dv[i] = 0.0;
    const double l = lb[i * lbs], u = ub[i * lbs];
    if (S.g[i] < 0.0 && finite_d(u)) { v[i] = u - S.x[i]; dv[i] = -1.0; }
    if (S.g[i] > 0.0 && finite_d(l)) { v[i] = S.x[i] - l; dv[i] = 1.0; }
    g_norm = dmax(g_norm, fabs(S.g[i] * v[i]));
  }
  if (g_norm < O.gtol) S.status = kStGtol;
  if (S.status != kStRunning) return false;
  if (S.nfev == O.max_nfev) { S.status = kStMaxNfev; return false; }
  double diag_h[N];
#pragma unroll
  for (int i = 0; i < N; i++) {
    // trf.py: scale = x_scale as given, or 1 / scale_inv when it comes from the Jacobian
    const double sc = O.x_scale_jac ? 1.0 / S.scale_inv[i] : O.x_scale[i];
    if (dv[i] != 0.0) v[i] *= S.scale_inv[i];
    S.d[i] = sqrt(v[i]) * sc;
    diag_h[i] = S.g[i] * dv[i] * sc;
    S.g_h[i] = S.d[i] * S.g[i];
  }
#pragma unroll
  for (int i = 0; i < N; i++) {
    const bool fi = (O.frozen >> i) & 1u;
#pragma unroll
    for (int j = 0; j <= i; j++) {
      const bool fj = (O.frozen >> j) & 1u;
      double t = S.d[i] * S.A[i][j] * S.d[j];
      if (i == j) t += diag_h[i];
      if (fi || fj) t = (i == j) ? 1.0 : 0.0;
      S.B[i][j] = t;
    }
    if (fi) S.g_h[i] = 0.0;
  }
  S.theta = dmax(0.995, 1.0 - g_norm);
  return true;
}

// common.py: solve_lsq_trust_region, on B instead of the SVD (see header).
template <class M> PNB_HD void trf_solve_tr(TrfLane<M> &S, double (&p_h)[M::NP]) {
  constexpr int N = M::NP;
  double L[N][N], dinv[N], q[N];
  const double Delta = S.Delta;
  const bool full_rank = ldlt<N>(S.B, 0.0, L, dinv);
  double pn = 0.0;
  if (full_rank) {
    ldlt_solve<N>(L, dinv, S.g_h, q);
    pn = norm2<N>(q);
    if (pn <= Delta) {
#pragma unroll
      for (int i = 0; i < N; i++) p_h[i] = -q[i];
      S.alpha = 0.0;
      return;
    }
  }
  double alpha_upper = norm2<N>(S.g_h) / Delta;
  double alpha_lower = 0.0;
  if (full_rank) {
    const double phi = pn - Delta;
    const double phi_prime = -ldlt_curv<N>(L, dinv, q) / pn;
    alpha_lower = -phi / phi_prime;
  }
  double alpha = S.alpha;
  if (!full_rank && alpha == 0.0)
    alpha = dmax(0.001 * alpha_upper, sqrt(alpha_lower * alpha_upper));
  for (int it = 0; it < 10; it++) {
    if (alpha < alpha_lower || alpha > alpha_upper)
      alpha = dmax(0.001 * alpha_upper, sqrt(alpha_lower * alpha_upper));
    ldlt<N>(S.B, alpha, L, dinv);
    ldlt_solve<N>(L, dinv, S.g_h, q);
    pn = norm2<N>(q);
    const double phi = pn - Delta;
    const double phi_prime = -ldlt_curv<N>(L, dinv, q) / pn;
    if (phi < 0.0) alpha_upper = alpha;
    const double ratio = phi / phi_prime;
    alpha_lower = dmax(alpha_lower, alpha - ratio);
    alpha -= (phi + Delta) * ratio / Delta;
    if (fabs(phi) < 0.01 * Delta) break;
  }
  ldlt<N>(S.B, alpha, L, dinv);
  ldlt_solve<N>(L, dinv, S.g_h, q);
  const double sc = -Delta / norm2<N>(q);
#pragma unroll
  for (int i = 0; i < N; i++) p_h[i] = q[i] * sc;
  S.alpha = alpha;
}

// trf.py: select_step.  On entry p_h is the trust-region solution; on exit
// S.x_new, S.predicted, S.step_h_norm, S.step_norm describe the chosen step.
template <class M>
PNB_HD void trf_select_step(TrfLane<M> &S, double (&p_h)[M::NP], const double *lb,
                            const double *ub, int lbs, unsigned frozen) {
  constexpr int N = M::NP;
  double p[N], step[N], step_h[N];
  bool inside = true;
#pragma unroll
  for (int i = 0; i < N; i++) {
    p[i] = S.d[i] * p_h[i];
    const double xn = S.x[i] + p[i];
    if (!((frozen >> i) & 1u)) inside = inside && (xn >= lb[i * lbs]) && (xn <= ub[i * lbs]);
  }
  if (inside) {
    S.predicted = -(0.5 * quad_form<N>(S.B, p_h) + dot<N>(p_h, S.g_h));
#pragma unroll
    for (int i = 0; i < N; i++) { step[i] = p[i]; step_h[i] = p_h[i]; }
  } else {
    unsigned hits, dummy;
    const double p_stride = step_to_bound<N>(S.x, p, lb, ub, lbs, hits);
    double r_h[N], r[N], x_on_bound[N];
#pragma unroll
    for (int i = 0; i < N; i++) {
      r_h[i] = ((hits >> i) & 1u) ? -p_h[i] : p_h[i];
      r[i] = S.d[i] * r_h[i];
      p[i] *= p_stride;
      p_h[i] *= p_stride;
      x_on_bound[i] = S.x[i] + p[i];
    }
    // common.py: intersect_trust_region(p_h, r_h, Delta), larger root
    double to_tr;
    {
      const double a = dot<N>(r_h, r_h), b = dot<N>(p_h, r_h);
      const double c = dot<N>(p_h, p_h) - S.Delta * S.Delta;
      const double dd = sqrt(b * b - a * c);
      const double qq = -(b + copysign(dd, b));
      const double t1 = qq / a, t2 = c / qq;
      to_tr = (t1 < t2) ? t2 : t1;
    }
    double to_bound = step_to_bound<N>(x_on_bound, r, lb, ub, lbs, dummy);
    double r_stride = dmin(to_bound, to_tr), r_stride_l, r_stride_u;
    if (r_stride > 0.0) {
      r_stride_l = (1.0 - S.theta) * p_stride / r_stride;
      r_stride_u = (r_stride == to_bound) ? S.theta * to_bound : to_tr;
    } else {
      r_stride_l = 0.0;
      r_stride_u = -1.0;
    }
    double r_value = kInf;
    if (r_stride_l <= r_stride_u) {
      // build_quadratic_1d(J_h, g_h, r_h, s0=p_h, diag=diag_h)
      const double a = 0.5 * quad_form<N>(S.B, r_h);
      const double b = dot<N>(S.g_h, r_h) + bilinear<N>(S.B, p_h, r_h);
      const double c = 0.5 * quad_form<N>(S.B, p_h) + dot<N>(S.g_h, p_h);
      r_stride = min_quad_1d(a, b, r_stride_l, r_stride_u, c, r_value);
#pragma unroll
      for (int i = 0; i < N; i++) {
        r_h[i] = r_h[i] * r_stride + p_h[i];
        r[i] = r_h[i] * S.d[i];
      }
    }
#pragma unroll
    for (int i = 0; i < N; i++) { p[i] *= S.theta; p_h[i] *= S.theta; }
    const double p_value = 0.5 * quad_form<N>(S.B, p_h) + dot<N>(p_h, S.g_h);
    double ag_h[N], ag[N];
#pragma unroll
    for (int i = 0; i < N; i++) { ag_h[i] = -S.g_h[i]; ag[i] = S.d[i] * ag_h[i]; }
    to_tr = S.Delta / norm2<N>(ag_h);
    to_bound = step_to_bound<N>(S.x, ag, lb, ub, lbs, dummy);
    double ag_stride = (to_bound < to_tr) ? S.theta * to_bound : to_tr;
    double ag_value;
    {
      const double a = 0.5 * quad_form<N>(S.B, ag_h);
      const double b = dot<N>(S.g_h, ag_h);
      ag_stride = min_quad_1d(a, b, 0.0, ag_stride, 0.0, ag_value);
    }
    if (p_value < r_value && p_value < ag_value) {
      S.predicted = -p_value;
#pragma unroll
      for (int i = 0; i < N; i++) { step[i] = p[i]; step_h[i] = p_h[i]; }
    } else if (r_value < p_value && r_value < ag_value) {
      S.predicted = -r_value;
#pragma unroll
      for (int i = 0; i < N; i++) { step[i] = r[i]; step_h[i] = r_h[i]; }
    } else {
      S.predicted = -ag_value;
#pragma unroll
      for (int i = 0; i < N; i++) { step[i] = ag[i] * ag_stride; step_h[i] = ag_h[i] * ag_stride; }
    }
  }
  S.step_h_norm = norm2<N>(step_h);
  S.step_norm = norm2<N>(step);
  // x_new = make_strictly_feasible(x + step, lb, ub, rstep=0)
#pragma unroll
  for (int i = 0; i < N; i++)
    S.x_new[i] = ((frozen >> i) & 1u)
                     ? S.x[i]
                     : strictly_feasible(S.x[i] + step[i], lb[i * lbs], ub[i * lbs], 0.0);
}

// Evaluate cost, J^T f and J^T J at xe, streaming over the measurements.
// yb(i) -> (y_i, b_i).  jac_mode 1 reproduces SciPy's 2-point differences
// (_numdiff.py: h = sqrt(eps) * sign(x) * max(1, |x|), flipped / shrunk at the
// bounds by _adjust_scheme_to_bounds, column = (f(x + h e_k) - f(x)) / dx).
// EXTRAS (a separate instantiation, the plain one stays as it is): `w` = 1 / sigma of curve_fit per
// row (or nullptr) scales residual and Jacobian rows (_wrap_func / _wrap_jac), and a robust loss
// rescales them as scale_for_robust_loss_function does: with J_s = max(rho' + 2 rho'' f^2, eps),
// J^T J -> sum J_s row row^T, J^T f -> sum rho' f row, cost -> 0.5 sum rho.
template <class M, bool EXTRAS = false, class RowFn>
PNB_HD void trf_evaluate(const double (&xe)[M::NP], const TrfOptions &O, int m, RowFn yb,
                         const double *lb, const double *ub, int lbs, double &cost,
                         double (&g)[M::NP], double (&A)[M::NP][M::NP], const double *w = nullptr) {
  constexpr int N = M::NP;
  typename M::Point pt;
  M::prepare(xe, O.tr, O.tm, pt);
  double c = 0.0;
#pragma unroll
  for (int i = 0; i < N; i++) {
    g[i] = 0.0;
#pragma unroll
    for (int j = 0; j <= i; j++) A[i][j] = 0.0;
  }
  if (O.jac_mode == 0) {
    for (int r = 0; r < m; r++) {
      double yv, bv;
      yb(r, yv, bv);
      double gr[N];
      double f = M::value_grad(pt, bv, gr) - yv;
      if (EXTRAS) {
        const double wv = w ? w[r] : 1.0;
        f *= wv;
        double r0, r1, r2;
        trf_rho(O.loss, f, O.f_scale, r0, r1, r2);
        double js = r1 + 2.0 * r2 * f * f;
        if (js < kEps) js = kEps;
        c += r0;
        const double fs = f * r1;
#pragma unroll
        for (int i = 0; i < N; i++) {
          gr[i] = ((O.frozen >> i) & 1u) ? 0.0 : gr[i] * wv;
          g[i] += gr[i] * fs;
#pragma unroll
          for (int j = 0; j <= i; j++) A[i][j] += js * (gr[i] * gr[j]);
        }
        continue;
      }
      c += f * f;
#pragma unroll
      for (int i = 0; i < N; i++) {
        if ((O.frozen >> i) & 1u) gr[i] = 0.0;
        g[i] += gr[i] * f;
#pragma unroll
        for (int j = 0; j <= i; j++) A[i][j] += gr[i] * gr[j];
      }
    }
  } else {
    typename M::Point pk[N];
    double dx[N];
    const double rstep = sqrt(kEps);
#pragma unroll
    for (int k = 0; k < N; k++) {
      const double xk = xe[k];
      double h = rstep * (xk >= 0.0 ? 1.0 : -1.0) * dmax(1.0, fabs(xk));
      if (EXTRAS && O.diff_step[k] > 0.0) {
        // _numdiff.py: _compute_absolute_step with rel_step: rel_step * sign(x) * |x|, the default where
        // that step vanishes in floating point
        const double hr = O.diff_step[k] * (xk >= 0.0 ? 1.0 : -1.0) * fabs(xk);
        if ((xk + hr) - xk != 0.0) h = hr;
      }
      const double l = lb[k * lbs], u = ub[k * lbs];
      const double lower = xk - l, upper = u - xk;
      const double xh = xk + h;
      const bool violated = (xh < l) || (xh > u);
      const bool fitting = fabs(h) <= dmax(lower, upper);
      if (violated && fitting) h = -h;
      if (!fitting) h = (upper >= lower) ? upper : -lower;
      if (O.jac_mode == 2) {
        // minpack/fdjac2.f: h = eps * |x|, eps = sqrt(max(epsfcn, epsmch)); h = eps when x = 0
        h = rstep * fabs(xk);
        if (h == 0.0) h = rstep;
      }
      double xt[N];
#pragma unroll
      for (int j = 0; j < N; j++) xt[j] = xe[j];
      xt[k] = xk + h;
      // the quotient below multiplies (<= 1 ulp from SciPy's division); MINPACK divides by h itself
      dx[k] = (O.jac_mode == 2) ? 1.0 / h : 1.0 / (xt[k] - xk);
      M::prepare(xt, O.tr, O.tm, pk[k]);
    }
    PNB_TRF_ROW_PRAGMA
    for (int r = 0; r < m; r++) {
      double yv, bv;
      yb(r, yv, bv);
      double e[M::K];
      M::exps(pt, bv, e);
      if (EXTRAS) {
        double gr[N], ep[M::K];
        M::perturbed_exps(pk, pt, bv, e, ep);
        const double wv = w ? w[r] : 1.0;
        const double f = (M::combine(pt, e) - yv) * wv;
        double r0, r1, r2;
        trf_rho(O.loss, f, O.f_scale, r0, r1, r2);
        double js = r1 + 2.0 * r2 * f * f;
        if (js < kEps) js = kEps;
        c += r0;
        const double fs = f * r1;
#pragma unroll
        for (int k = 0; k < N; k++) {
          gr[k] = 0.0;
          if (!((O.frozen >> k) & 1u))
            gr[k] = (((M::value_perturbed(pk[k], e, ep, k) - yv) * wv) - f) * dx[k];
        }
#pragma unroll
        for (int i = 0; i < N; i++) {
          g[i] += gr[i] * fs;
#pragma unroll
          for (int j = 0; j <= i; j++) A[i][j] += js * (gr[i] * gr[j]);
        }
        continue;
      }
      const double f = M::combine(pt, e) - yv;
      c += f * f;
      double gr[N], ep[M::K];
      M::perturbed_exps(pk, pt, bv, e, ep);
#pragma unroll
      for (int k = 0; k < N; k++) {
#if defined(PNB_FD_SELECT_ON_FROZEN)
        // measured slower than the (warp-uniform) branch: 10.95 vs 10.63 ms, profiles/r2_trf_experiments.md
        const double col = ((M::value_perturbed(pk[k], e, ep, k) - yv) - f) * dx[k];
        gr[k] = ((O.frozen >> k) & 1u) ? 0.0 : col;
#else
        gr[k] = 0.0;
        if (!((O.frozen >> k) & 1u))
          gr[k] = ((M::value_perturbed(pk[k], e, ep, k) - yv) - f) * dx[k];
#endif
      }
#pragma unroll
      for (int i = 0; i < N; i++) {
        g[i] += gr[i] * f;
#pragma unroll
        for (int j = 0; j <= i; j++) A[i][j] += gr[i] * gr[j];
      }
    }
  }
  cost = 0.5 * c;
}

// x_scale='jac' bookkeeping (common.py: compute_jac_scale)
template <class M>
PNB_HD void trf_update_jac_scale(TrfLane<M> &S, const TrfOptions &O, bool first) {
  constexpr int N = M::NP;
#pragma unroll
  for (int i = 0; i < N; i++) {
    if (O.x_scale_jac && !((O.frozen >> i) & 1u)) {
      double s = sqrt(S.A[i][i]);
      if (first) { if (s == 0.0) s = 1.0; }
      else s = dmax(s, S.scale_inv[i]);
      S.scale_inv[i] = s;
    } else if (first) {
      S.scale_inv[i] = 1.0 / O.x_scale[i];
    }
  }
}

// Validate and start a voxel (least_squares() preamble).  Returns false when
// SciPy would raise (status set, nothing to evaluate).
template <class M>
PNB_HD bool trf_begin(TrfLane<M> &S, const TrfOptions &O, const double (&p0)[M::NP],
                      const double *lb, const double *ub, int lbs, bool y_finite) {
  constexpr int N = M::NP;
  S.status = kStRunning;
  S.nfev = 0; S.njev = 0; S.alpha = 0.0; S.cost = 0.0; S.Delta = 0.0;
  S.need_prologue = true;
  bool bad_bounds = false, outside = false;
#pragma unroll
  for (int i = 0; i < N; i++) {
    S.x[i] = p0[i];
    if ((O.frozen >> i) & 1u) continue;
    const double l = lb[i * lbs], u = ub[i * lbs];
    if (!(l < u)) bad_bounds = true;
    if (!(p0[i] >= l && p0[i] <= u)) outside = true;
  }
  if (!y_finite) { S.status = kStNonFiniteY; return false; }
  if (bad_bounds) { S.status = kStBadBounds; return false; }
  if (outside) { S.status = kStInfeasible; return false; }
  if (O.method == 0) {  // least_squares(): x0 = make_strictly_feasible(x0, lb, ub) for 'trf' only
#pragma unroll
    for (int i = 0; i < N; i++)
      if (!((O.frozen >> i) & 1u)) S.x[i] = strictly_feasible(S.x[i], lb[i * lbs], ub[i * lbs], 1e-10);
  }
  return true;
}

// After the evaluation at the (strictly feasible) start point.
template <class M>
PNB_HD bool trf_after_first_eval(TrfLane<M> &S, const TrfOptions &O, double cost,
                                 const double (&g)[M::NP], const double (&A)[M::NP][M::NP],
                                 const double *lb, const double *ub, int lbs) {
  constexpr int N = M::NP;
  if (!finite_d(cost)) { S.status = kStNonFiniteF0; return false; }
  S.cost = cost;
  S.nfev = 1; S.njev = 1;
#pragma unroll
  for (int i = 0; i < N; i++) {
    S.g[i] = g[i];
#pragma unroll
    for (int j = 0; j <= i; j++) S.A[i][j] = A[i][j];
  }
  trf_update_jac_scale<M>(S, O, true);
  // Delta = norm(x0 * scale_inv / v**0.5)
  double t = 0.0;
#pragma unroll
  for (int i = 0; i < N; i++) {
    if ((O.frozen >> i) & 1u) continue;
    double v = 1.0;
    bool moved = false;
    const double l = lb[i * lbs], u = ub[i * lbs];
    if (S.g[i] < 0.0 && finite_d(u)) { v = u - S.x[i]; moved = true; }
    if (S.g[i] > 0.0 && finite_d(l)) { v = S.x[i] - l; moved = true; }
    if (moved) v *= S.scale_inv[i];
    const double q = S.x[i] * S.scale_inv[i] / sqrt(v);
    t += q * q;
  }
  S.Delta = sqrt(t);
  if (S.Delta == 0.0) S.Delta = 1.0;
  return true;
}

// Body of the inner `while actual_reduction <= 0 and nfev < max_nfev` loop after
// f(x_new) is known.  Returns true when the outer iteration is over (the next
// pass must run the prologue), false when another trial from the same point
// is needed.
template <class M>
PNB_HD bool trf_after_trial(TrfLane<M> &S, const TrfOptions &O, double cost_new,
                            const double (&g_new)[M::NP], const double (&A_new)[M::NP][M::NP]) {
  constexpr int N = M::NP;
  S.nfev += 1;
  if (!finite_d(cost_new)) {
    S.Delta = 0.25 * S.step_h_norm;
    // inner loop condition: actual_reduction (still -1) <= 0 and nfev < max_nfev
    return !(S.nfev < O.max_nfev);
  }
  const double actual = S.cost - cost_new;
  // common.py: update_tr_radius
  double ratio;
  if (S.predicted > 0.0) ratio = actual / S.predicted;
  else if (S.predicted == 0.0 && actual == 0.0) ratio = 1.0;
  else ratio = 0.0;
  double Delta_new = S.Delta;
  if (ratio < 0.25) Delta_new = 0.25 * S.step_h_norm;
  else if (ratio > 0.75 && S.step_h_norm > 0.95 * S.Delta) Delta_new = 2.0 * S.Delta;
  // common.py: check_termination
  double xn = 0.0;
#pragma unroll
  for (int i = 0; i < N; i++)
    if (!((O.frozen >> i) & 1u)) xn += S.x[i] * S.x[i];
  xn = sqrt(xn);
  const bool ft = (actual < O.ftol * S.cost) && (ratio > 0.25);
  const bool xt = S.step_norm < O.xtol * (O.xtol + xn);
  if (ft && xt) S.status = kStBoth;
  else if (ft) S.status = kStFtol;
  else if (xt) S.status = kStXtol;
  const bool terminated = S.status != kStRunning;
  if (!terminated) {
    S.alpha *= S.Delta / Delta_new;
    S.Delta = Delta_new;
  }
  if (actual > 0.0) {
#pragma unroll
    for (int i = 0; i < N; i++) {
      S.x[i] = S.x_new[i];
      S.g[i] = g_new[i];
#pragma unroll
      for (int j = 0; j <= i; j++) S.A[i][j] = A_new[i][j];
    }
    S.cost = cost_new;
    S.njev += 1;
    trf_update_jac_scale<M>(S, O, false);
    return true;
  }
  // rejected step: leave the inner loop only on termination or nfev exhaustion
  return terminated || !(S.nfev < O.max_nfev);
}

// Symmetric Jacobi eigen-decomposition, used only for the rank-deficient
// covariance fall-back (Moore-Penrose inverse as in curve_fit).
template <int N>
PNB_HD void jacobi_eig(double (&A)[N][N], double (&V)[N][N], double (&w)[N]) {
#pragma unroll
  for (int i = 0; i < N; i++)
#pragma unroll
    for (int j = 0; j < N; j++) V[i][j] = (i == j) ? 1.0 : 0.0;
  for (int sweep = 0; sweep < 30; sweep++) {
    double off = 0.0;
#pragma unroll
    for (int i = 0; i < N; i++)
#pragma unroll
      for (int j = 0; j < i; j++) off += A[i][j] * A[i][j];
    if (off == 0.0) break;
#pragma unroll
    for (int p = 0; p < N - 1; p++)
#pragma unroll
      for (int q = p + 1; q < N; q++) {
        const double apq = A[q][p];
        if (apq == 0.0) continue;
        const double zeta = (A[q][q] - A[p][p]) / (2.0 * apq);
        const double t = (zeta >= 0.0 ? 1.0 : -1.0) / (fabs(zeta) + sqrt(1.0 + zeta * zeta));
        const double c = 1.0 / sqrt(1.0 + t * t), s = c * t;
        // A <- R^T A R on the full symmetric matrix (kept symmetric explicitly)
#pragma unroll
        for (int k = 0; k < N; k++) {
          const double akp = A[k][p], akq = A[k][q];
          A[k][p] = c * akp - s * akq;
          A[k][q] = s * akp + c * akq;
        }
#pragma unroll
        for (int k = 0; k < N; k++) {
          const double apk = A[p][k], aqk = A[q][k];
          A[p][k] = c * apk - s * aqk;
          A[q][k] = s * apk + c * aqk;
        }
#pragma unroll
        for (int k = 0; k < N; k++) {
          const double vkp = V[k][p], vkq = V[k][q];
          V[k][p] = c * vkp - s * vkq;
          V[k][q] = s * vkp + c * vkq;
        }
      }
  }
#pragma unroll
  for (int i = 0; i < N; i++) w[i] = A[i][i];
}

// curve_fit's covariance: pinv(J^T J) * 2 cost / (m - n); inf when m <= n.
// cov is written densely over the FREE parameters (row-major n_free x n_free).
template <class M>
PNB_HD void trf_covariance(const TrfLane<M> &S, const TrfOptions &O, int m, double *cov) {
  constexpr int N = M::NP;
  int n_free = 0;
#pragma unroll
  for (int i = 0; i < N; i++) n_free += ((O.frozen >> i) & 1u) ? 0 : 1;
  double Af[N][N];
#pragma unroll
  for (int i = 0; i < N; i++) {
    const bool fi = (O.frozen >> i) & 1u;
#pragma unroll
    for (int j = 0; j <= i; j++) {
      const bool fj = (O.frozen >> j) & 1u;
      Af[i][j] = (fi || fj) ? ((i == j) ? 1.0 : 0.0) : S.A[i][j];
    }
  }
  double C[N][N];
  double L[N][N], dinv[N];
  if (m <= n_free && !O.absolute_sigma) {
#pragma unroll
    for (int i = 0; i < N; i++)
#pragma unroll
      for (int j = 0; j < N; j++) C[i][j] = kInf;
  } else {
    const double s_sq = O.absolute_sigma ? 1.0 : 2.0 * S.cost / (double)(m - n_free);
    if (ldlt<N>(Af, 0.0, L, dinv)) {
#pragma unroll
      for (int c = 0; c < N; c++) {
        double e[N], q[N];
#pragma unroll
        for (int i = 0; i < N; i++) e[i] = (i == c) ? 1.0 : 0.0;
        ldlt_solve<N>(L, dinv, e, q);
#pragma unroll
        for (int i = 0; i < N; i++) C[i][c] = q[i] * s_sq;
      }
    } else {
      // rank deficient: discard singular values s <= eps * max(m, n) * s_max, s^2 = eigenvalue
      double Asym[N][N], V[N][N], w[N];
#pragma unroll
      for (int i = 0; i < N; i++)
#pragma unroll
        for (int j = 0; j < N; j++) Asym[i][j] = (j <= i) ? Af[i][j] : Af[j][i];
      jacobi_eig<N>(Asym, V, w);
      double wmax = 0.0;
#pragma unroll
      for (int i = 0; i < N; i++) {
        const bool fi = (O.frozen >> i) & 1u;
        (void)fi;
        wmax = dmax(wmax, w[i]);
      }
      const double thr = kEps * (double)(m > n_free ? m : n_free);
      const double wthr = thr * thr * wmax;
      // Gram-matrix eigenvalues below ~eps * wmax are rounding noise
      const double noise = 4.0 * N * kEps * wmax;
#pragma unroll
      for (int i = 0; i < N; i++)
#pragma unroll
        for (int j = 0; j < N; j++) {
          double t = 0.0;
#pragma unroll
          for (int k = 0; k < N; k++)
            if (w[k] > wthr && w[k] > noise) t += V[i][k] * V[j][k] / w[k];
          C[i][j] = t * s_sq;
        }
    }
  }
  int ri = 0;
#pragma unroll
  for (int i = 0; i < N; i++) {
    if ((O.frozen >> i) & 1u) continue;
    int ci = 0;
#pragma unroll
    for (int j = 0; j < N; j++) {
      if ((O.frozen >> j) & 1u) continue;
      cov[ri * n_free + ci] = C[i][j];
      ci++;
    }
    ri++;
  }
}

}  // namespace pnb
