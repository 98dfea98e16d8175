// Per-voxel unbounded Levenberg-Marquardt: MINPACK's lmdif / lmder, register resident.
//
// The reference reaches MINPACK through solvers/curvefit.py:295-306 -> scipy.optimize.curve_fit(
// method="lm", maxfev=max_iter, ftol=tol) -> leastsq -> _minpack._lmdif (no Jacobian given: forward
// differences, fdjac2) or _lmder (analytic Jacobian, which the reference passes when a parameter is
// fixed).  curve_fit only accepts "lm" for problems without bounds.
//
// MINPACK factorises the m x n Jacobian (qrfac, column pivoting) and runs everything on R and Q^T f.
// Every quantity it derives from them is a function of the n x n normal matrix A = J^T J and
// g = J^T f — the same two things the TRF lane accumulates while it streams over the b-values — so
// the restatement below needs nothing larger than N x N:
//
//   acnorm_j = ||J e_j|| = sqrt(A_jj)                      column norms (diag scaling, mode 1)
//   gnorm    = max_j |g_j| / (||f|| acnorm_j)              cosine test
//   lmpar    : x(par) = (A + par D^2)^-1 g;  phi = ||D x|| - delta;
//              Newton correction  parc = (phi / delta) / (q^T (A + par D^2)^-1 q),  q = D^2 x / ||D x||
//              (MINPACK: || S^-T P^T q ||^2 with S^T S = P^T (A + par D^2) P),
//              lower bound  parl = (phi(0) / delta) / (q0^T A^-1 q0),  upper bound  paru = ||D^-1 g|| / delta
//   ||J p||^2 = p^T A p                                    predicted reduction
//
// with LDL^T factorisations in registers (pnb_trf_core.cuh).  Control flow, constants (0.1, 0.25,
// 0.75, 0.5, 1e-4, p1 = 0.1, p001 = 0.001, ten lmpar iterations), the nfev accounting (lmdif counts
// its n forward-difference evaluations per Jacobian, lmder does not) and the termination codes
// follow minpack/lmdif.f, lmder.f and lmpar.f statement by statement.
//
// Frozen parameters stay in the vector (as in the TRF lane): their row / column of A is the
// identity, their gradient and diagonal scale are zero, so every norm and step equals that of the
// reduced problem the reference hands to leastsq.
#pragma once
#include "pnb_trf_core.cuh"

namespace pnb {

// leastsq's info codes that are failures for curve_fit (ier not in 1..4), as kernel statuses
enum LmStatus {
  kLmFtolTooSmall = -6,  // info 6: "ftol=... is too small, no further reduction in the sum of squares is possible."
  kLmXtolTooSmall = -7,  // info 7: "xtol=... is too small, no further improvement in the approximate solution is possible."
  kLmGtolTooSmall = -8   // info 8: "gtol=... is too small, func(x) is orthogonal to the columns of the Jacobian ..."
};

template <class M> struct LmLane {
  static constexpr int N = M::NP;
  double diag[N];   // D (0 for frozen parameters)
  double par;       // Levenberg-Marquardt parameter, carried between iterations
  double delta;     // step bound
  double fnorm, xnorm, gnorm;
  double p[N];      // the trial step
  double pnorm, prered, dirder;
  int iter;         // MINPACK's iter (1-based count of outer iterations)
};

constexpr double kLmFactor = 100.0;  // leastsq default `factor`
constexpr double kLmDwarf = 2.2250738585072014e-308;

template <class M> PNB_HD unsigned lm_free_mask(const TrfOptions &O) {
  return ~O.frozen & ((1u << M::NP) - 1u);
}

// After the evaluation at x0 (MINPACK: first fcn call, fnorm = enorm(fvec); the Jacobian of the
// first outer iteration comes with the same pass here).
template <class M>
PNB_HD bool lm_after_first_eval(TrfLane<M> &S, LmLane<M> &LM, const TrfOptions &O, double cost,
                                const double (&g)[M::NP], const double (&A)[M::NP][M::NP]) {
  constexpr int N = M::NP;
  if (!finite_d(cost)) { S.status = kStNonFiniteF0; return false; }
  S.cost = cost;
  S.nfev = 1; S.njev = 0;
#pragma unroll
  for (int i = 0; i < N; i++) {
    S.g[i] = g[i];
#pragma unroll
    for (int j = 0; j <= i; j++) S.A[i][j] = A[i][j];
  }
  LM.fnorm = sqrt(2.0 * cost);
  LM.par = 0.0;
  LM.iter = 1;
  LM.delta = 0.0; LM.xnorm = 0.0; LM.gnorm = 0.0;
  return true;
}

// Top of the outer loop: Jacobian bookkeeping, scaling, cosine test.  Returns false when the voxel
// terminates here.
template <class M>
PNB_HD bool lm_prologue(TrfLane<M> &S, LmLane<M> &LM, const TrfOptions &O) {
  constexpr int N = M::NP;
  if (S.status != kStRunning) return false;  // the inner loop ended the fit
  const unsigned freem = lm_free_mask<M>(O);
  int n_free = 0;
#pragma unroll
  for (int i = 0; i < N; i++) n_free += (freem >> i) & 1u;
  // lmdif: nfev = nfev + n for the forward differences; lmder: njev = njev + 1
  if (O.jac_mode == 0) S.njev += 1;
  else { S.nfev += n_free; S.njev += 1; }
  double acnorm[N];
#pragma unroll
  for (int i = 0; i < N; i++) acnorm[i] = ((freem >> i) & 1u) ? sqrt(S.A[i][i]) : 0.0;
  if (LM.iter == 1) {
    double t = 0.0;
#pragma unroll
    for (int i = 0; i < N; i++) {
      LM.diag[i] = 0.0;
      if ((freem >> i) & 1u) {
        LM.diag[i] = (acnorm[i] == 0.0) ? 1.0 : acnorm[i];
        const double q = LM.diag[i] * S.x[i];
        t += q * q;
      }
    }
    LM.xnorm = sqrt(t);
    LM.delta = kLmFactor * LM.xnorm;
    if (LM.delta == 0.0) LM.delta = kLmFactor;
  }
  // norm of the scaled gradient
  double gnorm = 0.0;
  if (LM.fnorm != 0.0) {
#pragma unroll
    for (int i = 0; i < N; i++)
      if (((freem >> i) & 1u) && acnorm[i] != 0.0) gnorm = dmax(gnorm, fabs(S.g[i] / LM.fnorm / acnorm[i]));
  }
  LM.gnorm = gnorm;
  if (gnorm <= O.gtol) { S.status = kStGtol; return false; }  // info = 4 (gtol = 0 unless given)
#pragma unroll
  for (int i = 0; i < N; i++)
    if ((freem >> i) & 1u) LM.diag[i] = dmax(LM.diag[i], acnorm[i]);
  return true;
}

// M(par) = A + par D^2 over the free parameters, identity on frozen ones (lower triangle)
template <class M>
PNB_HD void lm_matrix(const TrfLane<M> &S, const LmLane<M> &LM, unsigned freem, double par,
                      double (&Mx)[M::NP][M::NP]) {
  constexpr int N = M::NP;
#pragma unroll
  for (int i = 0; i < N; i++) {
    const bool fi = (freem >> i) & 1u;
#pragma unroll
    for (int j = 0; j <= i; j++) {
      const bool fj = (freem >> j) & 1u;
      double t = S.A[i][j];
      if (i == j) t += par * LM.diag[i] * LM.diag[i];
      if (!fi || !fj) t = (i == j) ? 1.0 : 0.0;
      Mx[i][j] = t;
    }
  }
}

// minpack/lmpar.f on the normal matrix; on exit LM.p = -x(par) (the step), LM.par updated.
template <class M>
PNB_HD void lm_par(TrfLane<M> &S, LmLane<M> &LM, const TrfOptions &O) {
  constexpr int N = M::NP;
  const unsigned freem = lm_free_mask<M>(O);
  const double delta = LM.delta;
  double gf[N];
#pragma unroll
  for (int i = 0; i < N; i++) gf[i] = ((freem >> i) & 1u) ? S.g[i] : 0.0;
  double Mx[N][N], L[N][N], dinv[N], x[N], q[N];
  auto dnorm = [&](const double (&v)[N]) {
    double t = 0.0;
#pragma unroll
    for (int i = 0; i < N; i++) { const double w = LM.diag[i] * v[i]; t += w * w; }
    return sqrt(t);
  };
  // Gauss-Newton direction
  lm_matrix<M>(S, LM, freem, 0.0, Mx);
  const bool full_rank = ldlt<N>(Mx, 0.0, L, dinv);
  double dxnorm = kInf, fp = kInf;
  if (full_rank) {
    ldlt_solve<N>(L, dinv, gf, x);
    dxnorm = dnorm(x);
    fp = dxnorm - delta;
    if (fp <= 0.1 * delta) {  // the Gauss-Newton step is inside the region: par = 0
      LM.par = 0.0;
#pragma unroll
      for (int i = 0; i < N; i++) LM.p[i] = -x[i];
      return;
    }
  }
  // bounds of the zero of phi
  double parl = 0.0;
  if (full_rank) {
#pragma unroll
    for (int i = 0; i < N; i++) q[i] = LM.diag[i] * (LM.diag[i] * x[i] / dxnorm);
    parl = ((fp / delta) / ldlt_curv<N>(L, dinv, q));
  }
  double gn = 0.0;
#pragma unroll
  for (int i = 0; i < N; i++)
    if ((freem >> i) & 1u) { const double w = gf[i] / LM.diag[i]; gn += w * w; }
  gn = sqrt(gn);
  double paru = gn / delta;
  if (paru == 0.0) paru = kLmDwarf / dmin(delta, 0.1);
  double par = dmax(LM.par, parl);
  par = dmin(par, paru);
  if (par == 0.0) par = gn / dxnorm;
  for (int it = 1;; it++) {
    if (par == 0.0) par = dmax(kLmDwarf, 0.001 * paru);
    lm_matrix<M>(S, LM, freem, par, Mx);
    ldlt<N>(Mx, 0.0, L, dinv);
    ldlt_solve<N>(L, dinv, gf, x);
    dxnorm = dnorm(x);
    const double temp = fp;
    fp = dxnorm - delta;
    if (fabs(fp) <= 0.1 * delta || (parl == 0.0 && fp <= temp && temp < 0.0) || it == 10) break;
#pragma unroll
    for (int i = 0; i < N; i++) q[i] = LM.diag[i] * (LM.diag[i] * x[i] / dxnorm);
    const double parc = ((fp / delta) / ldlt_curv<N>(L, dinv, q));
    if (fp > 0.0) parl = dmax(parl, par);
    if (fp < 0.0) paru = dmin(paru, par);
    par = dmax(parl, par + parc);
  }
  LM.par = par;
#pragma unroll
  for (int i = 0; i < N; i++) LM.p[i] = -x[i];
}

// Top of the inner loop: the step and the trial point.
template <class M>
PNB_HD void lm_trial(TrfLane<M> &S, LmLane<M> &LM, const TrfOptions &O) {
  constexpr int N = M::NP;
  const unsigned freem = lm_free_mask<M>(O);
  lm_par<M>(S, LM, O);
  double t = 0.0;
#pragma unroll
  for (int i = 0; i < N; i++) {
    if (!((freem >> i) & 1u)) LM.p[i] = 0.0;
    S.x_new[i] = S.x[i] + LM.p[i];
    const double w = LM.diag[i] * LM.p[i];
    t += w * w;
  }
  LM.pnorm = sqrt(t);
  if (LM.iter == 1) LM.delta = dmin(LM.delta, LM.pnorm);  // on the first iteration, adjust the initial step bound
  // scaled predicted reduction and directional derivative: ||J p||^2 = p^T A p
  double Af[N][N];
  lm_matrix<M>(S, LM, freem, 0.0, Af);
  double pf[N];
#pragma unroll
  for (int i = 0; i < N; i++) pf[i] = ((freem >> i) & 1u) ? LM.p[i] : 0.0;
  const double jp2 = quad_form<N>(Af, pf);
  const double temp1 = sqrt(jp2 > 0.0 ? jp2 : 0.0) / LM.fnorm;
  const double temp2 = sqrt(LM.par) * LM.pnorm / LM.fnorm;
  LM.prered = temp1 * temp1 + temp2 * temp2 / 0.5;
  LM.dirder = -(temp1 * temp1 + temp2 * temp2);
}

// After f (and the Jacobian) at the trial point.  Returns true when the outer iteration is over
// (accepted step or termination), false when the inner loop repeats from the same point.
template <class M>
PNB_HD bool lm_after_trial(TrfLane<M> &S, LmLane<M> &LM, const TrfOptions &O, double cost_new,
                           const double (&g_new)[M::NP], const double (&A_new)[M::NP][M::NP]) {
  constexpr int N = M::NP;
  const unsigned freem = lm_free_mask<M>(O);
  S.nfev += 1;
  const double fnorm1 = sqrt(2.0 * cost_new);  // NaN / inf propagate like in MINPACK: the step is rejected
  double actred = -1.0;
  if (0.1 * fnorm1 < LM.fnorm) { const double r = fnorm1 / LM.fnorm; actred = 1.0 - r * r; }
  double ratio = 0.0;
  if (LM.prered != 0.0) ratio = actred / LM.prered;
  // update the step bound
  if (ratio <= 0.25) {
    double temp = 0.5;
    if (actred < 0.0) temp = 0.5 * LM.dirder / (LM.dirder + 0.5 * actred);
    if (0.1 * fnorm1 >= LM.fnorm || temp < 0.1) temp = 0.1;
    LM.delta = temp * dmin(LM.delta, LM.pnorm / 0.1);
    LM.par = LM.par / temp;
  } else if (LM.par == 0.0 || ratio >= 0.75) {
    LM.delta = LM.pnorm / 0.5;
    LM.par = 0.5 * LM.par;
  }
  const bool accept = ratio >= 1e-4;
  if (accept) {
    double t = 0.0;
#pragma unroll
    for (int i = 0; i < N; i++) {
      S.x[i] = S.x_new[i];
      const double w = LM.diag[i] * S.x[i];
      t += w * w;
    }
    LM.xnorm = sqrt(t);
    LM.fnorm = fnorm1;
    S.cost = cost_new;
    LM.iter += 1;
  }
  // tests for convergence
  int info = 0;
  if (fabs(actred) <= O.ftol && LM.prered <= O.ftol && 0.5 * ratio <= 1.0) info = 1;
  if (LM.delta <= O.xtol * LM.xnorm) info = 2;
  if (fabs(actred) <= O.ftol && LM.prered <= O.ftol && 0.5 * ratio <= 1.0 && info == 2) info = 3;
  if (info == 0) {
    // tests for termination and stringent tolerances
    if (S.nfev >= O.max_nfev) info = 5;
    if (fabs(actred) <= kEps && LM.prered <= kEps && 0.5 * ratio <= 1.0) info = 6;
    if (LM.delta <= kEps * LM.xnorm) info = 7;
    if (LM.gnorm <= kEps) info = 8;
  }
  if (info != 0) {
    // S.A stays the Jacobian of THIS outer iteration: leastsq builds cov_x from the last fjac,
    // which MINPACK computed before the final step
    S.status = info == 1 ? (int)kStFtol : info == 2 ? (int)kStXtol : info == 3 ? (int)kStBoth
               : info == 5 ? (int)kStMaxNfev : info == 6 ? (int)kLmFtolTooSmall
               : info == 7 ? (int)kLmXtolTooSmall : (int)kLmGtolTooSmall;
    (void)freem;
    return true;
  }
  if (!accept) return false;
#pragma unroll
  for (int i = 0; i < N; i++) {
    S.g[i] = g_new[i];
#pragma unroll
    for (int j = 0; j <= i; j++) S.A[i][j] = A_new[i][j];
  }
  return true;
}

}  // namespace pnb
