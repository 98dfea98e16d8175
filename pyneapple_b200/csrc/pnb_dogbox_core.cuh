// Per-voxel `dogbox` least squares (rectangular trust regions), register resident.
//
// SciPy's scipy/optimize/_lsq/dogbox.py (`dogbox`, `dogleg_step`,
// `find_intersection`) with tr_solver='exact', loss='linear' — what the reference
// reaches with `method = "dogbox"` (solvers/curvefit.py:295-306, the keyword is
// forwarded from the TOML, io/toml.py:328-332) — restated on the same per-lane
// state as the TRF kernel: J is never stored, only A = J^T J and g = J^T f.
//
//   * Gauss-Newton step on the free set:  lstsq(J_free, -f) = -(A_ff)^-1 g_f  by the
//     register LDL^T (invariant under the diagonal scaling that makes J badly
//     conditioned here); a numerically rank-deficient A_ff falls back to the
//     Moore-Penrose solution through a Jacobi eigen-decomposition, which is
//     lstsq's minimum-norm answer (a zero Jacobian column gets a zero step);
//   * build_quadratic_1d(J_free, g_free, -g_free):  a = g_f^T A_ff g_f / 2, b = -|g_f|^2;
//     evaluate_quadratic(J_free, g_free, s) = s^T A_ff s / 2 + g_f . s;
//   * everything else (active set on_bound * g < 0, intersection of the trust box
//     with the bounds, constrained Cauchy step, the walk from it towards the
//     Newton step, radius update with tr_hit, termination tests, nfev accounting,
//     variables put exactly on a bound they hit) follows dogbox.py line by line.
//
// least_squares() does not move x0 strictly inside the bounds for this method
// (that is TRF only), so trf_begin skips it.  SciPy evaluates the next Jacobian at
// the point with the bound-hitting variables snapped onto the bound while keeping
// f of the clipped trial point; both are evaluated at the trial point here — the
// two points differ by rounding (x + step vs. the bound itself).
#pragma once
#include "pnb_trf_core.cuh"

namespace pnb {

// The dogbox state of one lane lives in the TrfLane fields TRF does not need at
// the same time:  d <- Newton step, g_h <- scale (not its inverse), theta <- a,
// alpha <- b; plus the masks below.
template <class M> struct DogboxLane {
  unsigned on_lo, on_hi;    // on_bound == -1 / +1
  unsigned free_set;        // ~active_set of the current outer iteration, frozen parameters excluded
  unsigned hit_lo, hit_hi;  // on_bound_free of the trial
  bool tr_hit;
};

template <class M>
PNB_HD double dbx_scale(const TrfLane<M> &S, const TrfOptions &O, int i) {
  return O.x_scale_jac ? 1.0 / S.scale_inv[i] : O.x_scale[i];
}

// After the evaluation at x0: Delta = ||x0 * scale_inv||_inf, on_bound from equality with the bounds.
template <class M>
PNB_HD bool dbx_after_first_eval(TrfLane<M> &S, DogboxLane<M> &D, const TrfOptions &O, double cost,
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
  double t = 0.0;
  D.on_lo = 0; D.on_hi = 0;
#pragma unroll
  for (int i = 0; i < N; i++) {
    if ((O.frozen >> i) & 1u) continue;
    t = dmax(t, fabs(S.x[i] * S.scale_inv[i]));
    if (S.x[i] == lb[i * lbs]) D.on_lo |= 1u << i;
    if (S.x[i] == ub[i * lbs]) { D.on_hi |= 1u << i; D.on_lo &= ~(1u << i); }
  }
  S.Delta = (t == 0.0) ? 1.0 : t;
  return true;
}

// Top of dogbox's `while True`: active set, first-order optimality, Gauss-Newton step and the
// quadratic model along the anti-gradient.  Returns false when the voxel terminates here.
template <class M>
PNB_HD bool dbx_prologue(TrfLane<M> &S, DogboxLane<M> &D, const TrfOptions &O) {
  constexpr int N = M::NP;
  unsigned free_set = 0;
  double g_norm = 0.0;
#pragma unroll
  for (int i = 0; i < N; i++) {
    if ((O.frozen >> i) & 1u) continue;
    const double ob = ((D.on_lo >> i) & 1u) ? -1.0 : (((D.on_hi >> i) & 1u) ? 1.0 : 0.0);
    const bool active = ob * S.g[i] < 0.0;
    if (!active) {
      free_set |= 1u << i;
      g_norm = dmax(g_norm, fabs(S.g[i]));
    }
  }
  D.free_set = free_set;
  if (g_norm < O.gtol) S.status = kStGtol;
  if (S.status != kStRunning) return false;
  if (S.nfev == O.max_nfev) { S.status = kStMaxNfev; return false; }
  // newton_step = lstsq(J_free, -f)
  double Mx[N][N], gf[N];
#pragma unroll
  for (int i = 0; i < N; i++) {
    const bool fi = (free_set >> i) & 1u;
    gf[i] = fi ? S.g[i] : 0.0;
#pragma unroll
    for (int j = 0; j <= i; j++) {
      const bool fj = (free_set >> j) & 1u;
      Mx[i][j] = (fi && fj) ? S.A[i][j] : ((i == j) ? 1.0 : 0.0);
    }
  }
  double L[N][N], dinv[N], q[N];
  if (ldlt<N>(Mx, 0.0, L, dinv)) {
    ldlt_solve<N>(L, dinv, gf, q);
  } else {
    // minimum-norm solution: drop the directions J_free cannot see
    double Asym[N][N], V[N][N], w[N];
#pragma unroll
    for (int i = 0; i < N; i++)
#pragma unroll
      for (int j = 0; j < N; j++) {
        const bool fr = ((free_set >> i) & 1u) && ((free_set >> j) & 1u);
        Asym[i][j] = fr ? ((j <= i) ? S.A[i][j] : S.A[j][i]) : 0.0;
      }
    jacobi_eig<N>(Asym, V, w);
    double wmax = 0.0;
#pragma unroll
    for (int i = 0; i < N; i++) wmax = dmax(wmax, w[i]);
    const double noise = 4.0 * N * kEps * wmax;
#pragma unroll
    for (int i = 0; i < N; i++) q[i] = 0.0;
#pragma unroll
    for (int k = 0; k < N; k++) {
      if (!(w[k] > noise)) continue;
      double t = 0.0;
#pragma unroll
      for (int j = 0; j < N; j++) t += V[j][k] * gf[j];
      t /= w[k];
#pragma unroll
      for (int i = 0; i < N; i++) q[i] += V[i][k] * t;
    }
  }
  double gg = 0.0;
#pragma unroll
  for (int i = 0; i < N; i++) {
    S.d[i] = ((free_set >> i) & 1u) ? -q[i] : 0.0;  // Newton step
    S.g_h[i] = dbx_scale<M>(S, O, i);               // scale
    gg += gf[i] * gf[i];
  }
  // build_quadratic_1d(J_free, g_free, -g_free)
  double Az[N][N];
#pragma unroll
  for (int i = 0; i < N; i++)
#pragma unroll
    for (int j = 0; j <= i; j++) Az[i][j] = (((free_set >> i) & 1u) && ((free_set >> j) & 1u)) ? S.A[i][j] : 0.0;
  S.theta = 0.5 * quad_form<N>(Az, gf);  // a
  S.alpha = -gg;                         // b
  return true;
}

// common.py: step_size_to_bound restricted to the free set; hits as two masks
template <int N>
PNB_HD double dbx_step_to_bound(const double (&x)[N], const double (&s)[N], const double (&lb)[N],
                                const double (&ub)[N], unsigned free_set, unsigned &neg, unsigned &pos) {
  double steps[N];
  double mn = kInf;
#pragma unroll
  for (int i = 0; i < N; i++) {
    steps[i] = kInf;
    if (((free_set >> i) & 1u) && s[i] != 0.0) steps[i] = dmax((lb[i] - x[i]) / s[i], (ub[i] - x[i]) / s[i]);
    if ((free_set >> i) & 1u) mn = dmin(mn, steps[i]);
  }
  neg = 0; pos = 0;
#pragma unroll
  for (int i = 0; i < N; i++) {
    if (!((free_set >> i) & 1u) || steps[i] != mn) continue;
    if (s[i] < 0.0) neg |= 1u << i;
    if (s[i] > 0.0) pos |= 1u << i;
  }
  return mn;
}

// Body of the inner loop up to the function evaluation: dogleg_step, predicted reduction, x_new.
template <class M>
PNB_HD void dbx_trial(TrfLane<M> &S, DogboxLane<M> &D, const TrfOptions &O, const double *lb,
                      const double *ub, int lbs) {
  constexpr int N = M::NP;
  const unsigned fs = D.free_set;
  double lbt[N], ubt[N], zero[N], step[N];
  unsigned orig_l = 0, orig_u = 0, tr_l = 0, tr_u = 0;
  bool inside = true;
#pragma unroll
  for (int i = 0; i < N; i++) {
    zero[i] = 0.0; step[i] = 0.0; lbt[i] = 0.0; ubt[i] = 0.0;
    if (!((fs >> i) & 1u)) continue;
    const double trb = S.Delta * S.g_h[i];
    const double lc = lb[i * lbs] - S.x[i], uc = ub[i * lbs] - S.x[i];
    lbt[i] = dmax(lc, -trb);
    ubt[i] = dmin(uc, trb);
    if (lbt[i] == lc) orig_l |= 1u << i;
    if (ubt[i] == uc) orig_u |= 1u << i;
    if (lbt[i] == -trb) tr_l |= 1u << i;
    if (ubt[i] == trb) tr_u |= 1u << i;
    inside = inside && (S.d[i] >= lbt[i]) && (S.d[i] <= ubt[i]);
  }
  D.hit_lo = 0; D.hit_hi = 0; D.tr_hit = false;
  if (inside) {
#pragma unroll
    for (int i = 0; i < N; i++) step[i] = S.d[i];
  } else {
    double ng[N], cauchy[N], diff[N];
#pragma unroll
    for (int i = 0; i < N; i++) ng[i] = ((fs >> i) & 1u) ? -S.g[i] : 0.0;
    unsigned dn, dp;
    const double to_bounds = dbx_step_to_bound<N>(zero, ng, lbt, ubt, fs, dn, dp);
    double yv;
    const double t = min_quad_1d(S.theta, S.alpha, 0.0, to_bounds, 0.0, yv);
#pragma unroll
    for (int i = 0; i < N; i++) {
      cauchy[i] = ((fs >> i) & 1u) ? -t * S.g[i] : 0.0;
      diff[i] = S.d[i] - cauchy[i];
    }
    unsigned hn, hp;
    const double step_size = dbx_step_to_bound<N>(cauchy, diff, lbt, ubt, fs, hn, hp);
    D.hit_lo = hn & orig_l;
    D.hit_hi = hp & orig_u;
    D.tr_hit = ((hn & tr_l) | (hp & tr_u)) != 0;
#pragma unroll
    for (int i = 0; i < N; i++) step[i] = ((fs >> i) & 1u) ? cauchy[i] + step_size * diff[i] : 0.0;
  }
  // predicted_reduction = -evaluate_quadratic(J_free, g_free, step_free)
  double Az[N][N], gs = 0.0, sh = 0.0, sn = 0.0;
#pragma unroll
  for (int i = 0; i < N; i++) {
#pragma unroll
    for (int j = 0; j <= i; j++) Az[i][j] = (((fs >> i) & 1u) && ((fs >> j) & 1u)) ? S.A[i][j] : 0.0;
    if ((fs >> i) & 1u) gs += step[i] * S.g[i];
    sh = dmax(sh, fabs(step[i] * S.scale_inv[i]));
    sn += step[i] * step[i];
  }
  S.predicted = -(0.5 * quad_form<N>(Az, step) + gs);
  S.step_h_norm = sh;       // ||step * scale_inv||_inf
  S.step_norm = sqrt(sn);
  // x_new = clip(x + step, lb, ub)
#pragma unroll
  for (int i = 0; i < N; i++) {
    double xn = S.x[i] + step[i];
    if (!((O.frozen >> i) & 1u)) xn = dmin(dmax(xn, lb[i * lbs]), ub[i * lbs]);
    S.x_new[i] = xn;
  }
}

// Rest of the inner loop once f(x_new) is known.  Returns true when the outer iteration is over.
template <class M>
PNB_HD bool dbx_after_trial(TrfLane<M> &S, DogboxLane<M> &D, const TrfOptions &O, double cost_new,
                            const double (&g_new)[M::NP], const double (&A_new)[M::NP][M::NP],
                            const double *lb, const double *ub, int lbs) {
  constexpr int N = M::NP;
  S.nfev += 1;
  if (!finite_d(cost_new)) {
    S.Delta = 0.25 * S.step_h_norm;
    return !(S.nfev < O.max_nfev);
  }
  const double actual = S.cost - cost_new;
  // common.py: update_tr_radius(Delta, actual, predicted, step_h_norm, tr_hit)
  double ratio;
  if (S.predicted > 0.0) ratio = actual / S.predicted;
  else if (S.predicted == 0.0 && actual == 0.0) ratio = 1.0;
  else ratio = 0.0;
  if (ratio < 0.25) S.Delta = 0.25 * S.step_h_norm;
  else if (ratio > 0.75 && D.tr_hit) S.Delta *= 2.0;
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
  if (actual > 0.0) {
    // on_bound[free_set] = on_bound_free; variables on a bound are set exactly onto it
    D.on_lo = (D.on_lo & ~D.free_set) | D.hit_lo;
    D.on_hi = (D.on_hi & ~D.free_set) | D.hit_hi;
#pragma unroll
    for (int i = 0; i < N; i++) {
      double xi = S.x_new[i];
      if ((D.on_lo >> i) & 1u) xi = lb[i * lbs];
      if ((D.on_hi >> i) & 1u) xi = ub[i * lbs];
      S.x[i] = xi;
      S.g[i] = g_new[i];
#pragma unroll
      for (int j = 0; j <= i; j++) S.A[i][j] = A_new[i][j];
    }
    S.cost = cost_new;
    S.njev += 1;
    trf_update_jac_scale<M>(S, O, false);
    return true;
  }
  return terminated || !(S.nfev < O.max_nfev);
}

}  // namespace pnb
