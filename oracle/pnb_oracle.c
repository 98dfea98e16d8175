/*
 * ORACLE (test infrastructure, not product code).
 *
 * Plain-C, FP64, CPU restatement of the arithmetic behind Pyneapple's
 * voxel-wise fitting hot path.  The reference (darksim33/Pyneapple, pure
 * Python) delegates that arithmetic to SciPy, a third-party dependency that
 * is not vendored in the reference tree (pyproject.toml:34-36 pins
 * scipy>=1.17.1; the image has SciPy 1.18.1).  This file restates the
 * published algorithms as implemented there:
 *
 *   pnbo_trf_fit   <- scipy/optimize/_minpack_py.py  curve_fit (bounds, pcov)
 *                     scipy/optimize/_lsq/least_squares.py:least_squares
 *                     scipy/optimize/_lsq/trf.py        trf_bounds, select_step
 *                     scipy/optimize/_lsq/common.py     solve_lsq_trust_region,
 *                        CL_scaling_vector, step_size_to_bound,
 *                        intersect_trust_region, build_quadratic_1d,
 *                        minimize_quadratic_1d, evaluate_quadratic,
 *                        update_tr_radius, check_termination,
 *                        make_strictly_feasible, find_active_constraints
 *                     scipy/optimize/_numdiff.py       2-point differences
 *                        (_compute_absolute_step, _adjust_scheme_to_bounds)
 *                     called from solvers/curvefit.py:295-306
 *   pnbo_lsq_fit(method = 1)
 *                  <- scipy/optimize/_lsq/dogbox.py     dogbox, dogleg_step,
 *                        find_intersection (method = "dogbox" of the same call)
 *   pnbo_nnls      <- Lawson & Hanson, "Solving Least Squares Problems",
 *                     ch. 23 algorithm NNLS (what scipy.optimize.nnls wraps),
 *                     called from solvers/nnls_solver.py:195-197
 *   models         <- model_functions/multiexp.py:35-302
 *
 * The dense SVD that SciPy takes from LAPACK (gesdd) is replaced by a
 * one-sided Jacobi SVD, which is at least as accurate for these tiny,
 * badly column-scaled matrices.
 *
 * PINNING: tests/test_oracle_golden.py checks this restatement against the
 * golden vectors in tests/golden/ (outputs of the real reference, generated
 * by oracle/make_golden.py) and against SciPy itself on the GPU box.
 *
 * Only tests/, __graft_entry__.smoke() and bench.py's cpu_baseline leg may
 * load the library built from this file (oracle/_build/libpnb_oracle.so).
 */
#include <float.h>
#include <math.h>
#include <stdlib.h>
#include <string.h>

#define MAXN 8    /* max free parameters */
#define MAXM 256  /* max measurements */
#define EPS DBL_EPSILON

enum { M_MONO = 0, M_BI_RED, M_BI_FULL, M_BI_S0, M_TRI_RED, M_TRI_FULL, M_TRI_S0 };

typedef struct {
  int model_id, t1_mode; /* t1_mode: 0 none, 1 standard, 2 STEAM */
  double tr, tm;
  int n_all;
} model_t;

static int base_count(int id) {
  static const int n[] = {2, 3, 4, 4, 5, 6, 6};
  return n[id];
}

/* signal for the FULL parameter vector p (model_functions/multiexp.py) */
static void model_forward(const model_t *M, int nb, const double *b, const double *p, double *s) {
  for (int i = 0; i < nb; i++) {
    double bi = b[i], v;
    switch (M->model_id) {
      case M_MONO: v = p[0] * exp(-bi * p[1]); break;
      case M_BI_RED: v = p[0] * exp(-bi * p[1]) + (1 - p[0]) * exp(-bi * p[2]); break;
      case M_BI_FULL: v = p[0] * exp(-bi * p[1]) + p[2] * exp(-bi * p[3]); break;
      case M_BI_S0: v = p[3] * (p[0] * exp(-bi * p[1]) + (1 - p[0]) * exp(-bi * p[2])); break;
      case M_TRI_RED:
        v = p[0] * exp(-bi * p[1]) + p[2] * exp(-bi * p[3]) + (1 - p[0] - p[2]) * exp(-bi * p[4]);
        break;
      case M_TRI_FULL:
        v = p[0] * exp(-bi * p[1]) + p[2] * exp(-bi * p[3]) + p[4] * exp(-bi * p[5]);
        break;
      default: /* M_TRI_S0 */
        v = p[5] * (p[0] * exp(-bi * p[1]) + p[2] * exp(-bi * p[3]) +
                    (1 - p[0] - p[2]) * exp(-bi * p[4]));
    }
    if (M->t1_mode) {
      double t1 = p[M->n_all - 1];
      v = v * (1 - exp(-M->tr / t1));
      if (M->t1_mode == 2) v = v * exp(-M->tm / t1);
    }
    s[i] = v;
  }
}

/* analytic Jacobian, row-major nb x n_all (models/*.py jacobian + apply_t1_jacobian) */
static void model_jacobian(const model_t *M, int nb, const double *b, const double *p, double *J) {
  int na = M->n_all, nbase = base_count(M->model_id);
  for (int i = 0; i < nb; i++) {
    double bi = b[i], base;
    double *r = J + (size_t)i * na;
    switch (M->model_id) {
      case M_MONO: {
        double e = exp(-bi * p[1]);
        r[0] = e; r[1] = -bi * p[0] * e; base = p[0] * e;
      } break;
      case M_BI_RED: {
        double e1 = exp(-bi * p[1]), e2 = exp(-bi * p[2]);
        r[0] = e1 - e2; r[1] = -bi * p[0] * e1; r[2] = -bi * (1 - p[0]) * e2;
        base = p[0] * e1 + (1 - p[0]) * e2;
      } break;
      case M_BI_FULL: {
        double e1 = exp(-bi * p[1]), e2 = exp(-bi * p[3]);
        r[0] = e1; r[1] = -bi * p[0] * e1; r[2] = e2; r[3] = -bi * p[2] * e2;
        base = p[0] * e1 + p[2] * e2;
      } break;
      case M_BI_S0: {
        double e1 = exp(-bi * p[1]), e2 = exp(-bi * p[2]), s0 = p[3];
        r[0] = s0 * (e1 - e2); r[1] = -bi * s0 * p[0] * e1; r[2] = -bi * s0 * (1 - p[0]) * e2;
        r[3] = p[0] * e1 + (1 - p[0]) * e2;
        base = s0 * r[3];
      } break;
      case M_TRI_FULL: {
        double e1 = exp(-bi * p[1]), e2 = exp(-bi * p[3]), e3 = exp(-bi * p[5]);
        r[0] = e1; r[1] = -bi * p[0] * e1; r[2] = e2; r[3] = -bi * p[2] * e2;
        r[4] = e3; r[5] = -bi * p[4] * e3;
        base = p[0] * e1 + p[2] * e2 + p[4] * e3;
      } break;
      default: { /* TRI reduced / S0 */
        double e1 = exp(-bi * p[1]), e2 = exp(-bi * p[3]), e3 = exp(-bi * p[4]);
        double f3 = 1 - p[0] - p[2], s0 = (M->model_id == M_TRI_S0) ? p[5] : 1.0;
        double shape = p[0] * e1 + p[2] * e2 + f3 * e3;
        r[0] = s0 * (e1 - e3); r[1] = -bi * s0 * p[0] * e1; r[2] = s0 * (e2 - e3);
        r[3] = -bi * s0 * p[2] * e2; r[4] = -bi * s0 * f3 * e3;
        if (M->model_id == M_TRI_S0) r[5] = shape;
        base = s0 * shape;
      }
    }
    if (M->t1_mode) {
      double t1 = p[na - 1], e_tr = exp(-M->tr / t1), a = 1 - e_tr, factor, d;
      if (M->t1_mode == 2) {
        double e_tm = exp(-M->tm / t1);
        factor = a * e_tm;
        d = base * e_tm / (t1 * t1) * (-M->tr * e_tr + M->tm * a);
      } else {
        factor = a;
        d = base * (-e_tr * M->tr / (t1 * t1));
      }
      for (int k = 0; k < nbase; k++) r[k] *= factor;
      r[nbase] = d;
    }
  }
}

/* ------------------------------------------------------------------ */
/* the per-voxel problem: residual f(x) = model(x) - y over FREE params  */
/* ------------------------------------------------------------------ */
typedef struct {
  const model_t *M;
  int nb, n;               /* measurements, free parameters */
  const double *b, *y;
  int free_idx[MAXN];      /* free slot -> full index */
  double pfull[MAXN];      /* full vector with fixed values filled in */
  int jac_mode;            /* 0 analytic, 1 scipy '2-point' */
  const double *lb, *ub;
} prob_t;

static void prob_fun(prob_t *P, const double *x, double *f) {
  for (int k = 0; k < P->n; k++) P->pfull[P->free_idx[k]] = x[k];
  model_forward(P->M, P->nb, P->b, P->pfull, f);
  for (int i = 0; i < P->nb; i++) f[i] -= P->y[i];
}

/* J row-major nb x n */
static void prob_jac(prob_t *P, const double *x, const double *f0, double *J) {
  int n = P->n, m = P->nb;
  if (P->jac_mode == 0) {
    double Jf[MAXM * MAXN];
    for (int k = 0; k < n; k++) P->pfull[P->free_idx[k]] = x[k];
    model_jacobian(P->M, m, P->b, P->pfull, Jf);
    for (int i = 0; i < m; i++)
      for (int k = 0; k < n; k++) J[i * n + k] = Jf[i * P->M->n_all + P->free_idx[k]];
    return;
  }
  /* _numdiff.py: _compute_absolute_step + _adjust_scheme_to_bounds('1-sided') + _dense_difference */
  const double rstep = sqrt(EPS);
  double xt[MAXN], f1[MAXM];
  for (int k = 0; k < n; k++) {
    double sign = (x[k] >= 0) ? 1.0 : -1.0;
    double h = rstep * sign * fmax(1.0, fabs(x[k]));
    double lower = x[k] - P->lb[k], upper = P->ub[k] - x[k];
    int unbounded = isinf(P->lb[k]) && isinf(P->ub[k]);
    /* scipy skips the adjustment only when ALL bounds are infinite; per-component
       the formulas below reduce to "no change" for an infinite pair anyway */
    (void)unbounded;
    double xh = x[k] + h;
    int violated = (xh < P->lb[k]) || (xh > P->ub[k]);
    int fitting = fabs(h) <= fmax(lower, upper);
    if (violated && fitting) h = -h;
    if (!fitting) h = (upper >= lower) ? upper : -lower;
    memcpy(xt, x, sizeof(double) * n);
    xt[k] = x[k] + h;
    double dx = xt[k] - x[k];
    prob_fun(P, xt, f1);
    for (int i = 0; i < m; i++) J[i * n + k] = (f1[i] - f0[i]) / dx;
  }
}

/* ------------------------------------------------------------------ */
/* one-sided Jacobi SVD of A (rows x n, row-major): A = U diag(s) V^T    */
/* on exit A holds U*diag(s) columns; s sorted descending with V, A cols */
/* ------------------------------------------------------------------ */
static void jacobi_svd(int rows, int n, double *A, double *s, double *V) {
  for (int i = 0; i < n; i++)
    for (int j = 0; j < n; j++) V[i * n + j] = (i == j);
  for (int sweep = 0; sweep < 60; sweep++) {
    int rotated = 0;
    for (int p = 0; p < n - 1; p++)
      for (int q = p + 1; q < n; q++) {
        double app = 0, aqq = 0, apq = 0;
        for (int i = 0; i < rows; i++) {
          double ap = A[i * n + p], aq = A[i * n + q];
          app += ap * ap; aqq += aq * aq; apq += ap * aq;
        }
        if (apq == 0.0 || fabs(apq) <= 1e-300) continue;
        if (fabs(apq) <= 0.5 * EPS * sqrt(app * aqq)) continue;
        rotated = 1;
        double zeta = (aqq - app) / (2.0 * apq);
        double t = (zeta >= 0 ? 1.0 : -1.0) / (fabs(zeta) + sqrt(1.0 + zeta * zeta));
        double c = 1.0 / sqrt(1.0 + t * t), sn = c * t;
        for (int i = 0; i < rows; i++) {
          double ap = A[i * n + p], aq = A[i * n + q];
          A[i * n + p] = c * ap - sn * aq;
          A[i * n + q] = sn * ap + c * aq;
        }
        for (int i = 0; i < n; i++) {
          double vp = V[i * n + p], vq = V[i * n + q];
          V[i * n + p] = c * vp - sn * vq;
          V[i * n + q] = sn * vp + c * vq;
        }
      }
    if (!rotated) break;
  }
  for (int j = 0; j < n; j++) {
    double t = 0;
    for (int i = 0; i < rows; i++) t += A[i * n + j] * A[i * n + j];
    s[j] = sqrt(t);
  }
  /* sort descending (selection sort on columns) */
  for (int j = 0; j < n - 1; j++) {
    int k = j;
    for (int l = j + 1; l < n; l++) if (s[l] > s[k]) k = l;
    if (k != j) {
      double t = s[j]; s[j] = s[k]; s[k] = t;
      for (int i = 0; i < rows; i++) { t = A[i * n + j]; A[i * n + j] = A[i * n + k]; A[i * n + k] = t; }
      for (int i = 0; i < n; i++) { t = V[i * n + j]; V[i * n + j] = V[i * n + k]; V[i * n + k] = t; }
    }
  }
}

static double vnorm(int n, const double *v) {
  double t = 0;
  for (int i = 0; i < n; i++) t += v[i] * v[i];
  return sqrt(t);
}
static double vdot(int n, const double *a, const double *b) {
  double t = 0;
  for (int i = 0; i < n; i++) t += a[i] * b[i];
  return t;
}

/* common.py:find_active_constraints / make_strictly_feasible */
static void make_strictly_feasible(int n, double *x, const double *lb, const double *ub, double rstep) {
  for (int i = 0; i < n; i++) {
    int act = 0;
    if (rstep == 0) {
      if (x[i] <= lb[i]) act = -1;
      if (x[i] >= ub[i]) act = 1;
    } else {
      double ld = x[i] - lb[i], ud = ub[i] - x[i];
      double lt = rstep * fmax(1.0, fabs(lb[i])), ut = rstep * fmax(1.0, fabs(ub[i]));
      if (isfinite(lb[i]) && ld <= fmin(ud, lt)) act = -1;
      if (isfinite(ub[i]) && ud <= fmin(ld, ut)) act = 1;
    }
    if (act == -1) x[i] = (rstep == 0) ? nextafter(lb[i], ub[i]) : lb[i] + rstep * fmax(1.0, fabs(lb[i]));
    if (act == 1) x[i] = (rstep == 0) ? nextafter(ub[i], lb[i]) : ub[i] - rstep * fmax(1.0, fabs(ub[i]));
    if (x[i] < lb[i] || x[i] > ub[i]) x[i] = 0.5 * (lb[i] + ub[i]);
  }
}

static void cl_scaling(int n, const double *x, const double *g, const double *lb, const double *ub,
                       double *v, double *dv) {
  for (int i = 0; i < n; i++) {
    v[i] = 1.0; dv[i] = 0.0;
    if (g[i] < 0 && isfinite(ub[i])) { v[i] = ub[i] - x[i]; dv[i] = -1; }
    if (g[i] > 0 && isfinite(lb[i])) { v[i] = x[i] - lb[i]; dv[i] = 1; }
  }
}

static double step_size_to_bound(int n, const double *x, const double *s, const double *lb,
                                 const double *ub, int *hits) {
  double steps[MAXN], mn = INFINITY;
  for (int i = 0; i < n; i++) {
    steps[i] = INFINITY;
    if (s[i] != 0) steps[i] = fmax((lb[i] - x[i]) / s[i], (ub[i] - x[i]) / s[i]);
    if (steps[i] < mn) mn = steps[i];
  }
  if (hits)
    for (int i = 0; i < n; i++) hits[i] = (steps[i] == mn) ? ((s[i] > 0) - (s[i] < 0)) : 0;
  return mn;
}

/* J_h . s ; J_h = J * d given as J (row-major m x n) and d */
static void jh_dot(int m, int n, const double *J, const double *d, const double *s, double *out) {
  for (int i = 0; i < m; i++) {
    double t = 0;
    for (int k = 0; k < n; k++) t += J[i * n + k] * d[k] * s[k];
    out[i] = t;
  }
}

static double evaluate_quadratic(int m, int n, const double *J, const double *d, const double *g_h,
                                 const double *s, const double *diag) {
  double Js[MAXM];
  jh_dot(m, n, J, d, s, Js);
  double q = vdot(m, Js, Js);
  for (int k = 0; k < n; k++) q += s[k] * diag[k] * s[k];
  return 0.5 * q + vdot(n, s, g_h);
}

static void build_quadratic_1d(int m, int n, const double *J, const double *d, const double *g,
                               const double *s, const double *diag, const double *s0, double *a,
                               double *b, double *c) {
  double v[MAXM], u[MAXM];
  jh_dot(m, n, J, d, s, v);
  double aa = vdot(m, v, v);
  for (int k = 0; k < n; k++) aa += s[k] * diag[k] * s[k];
  aa *= 0.5;
  double bb = vdot(n, g, s), cc = 0;
  if (s0) {
    jh_dot(m, n, J, d, s0, u);
    bb += vdot(m, u, v);
    cc = 0.5 * vdot(m, u, u) + vdot(n, g, s0);
    for (int k = 0; k < n; k++) { bb += s0[k] * diag[k] * s[k]; cc += 0.5 * s0[k] * diag[k] * s0[k]; }
  }
  *a = aa; *b = bb; *c = cc;
}

static double minimize_quadratic_1d(double a, double b, double lo, double hi, double c, double *yv) {
  double t[3] = {lo, hi, 0};
  int nt = 2;
  if (a != 0) {
    double ext = -0.5 * b / a;
    if (lo < ext && ext < hi) t[nt++] = ext;
  }
  int best = 0;
  double ybest = 0;
  for (int i = 0; i < nt; i++) {
    double y = t[i] * (a * t[i] + b) + c;
    if (i == 0 || y < ybest) { ybest = y; best = i; } /* np.argmin: first minimum */
  }
  *yv = ybest;
  return t[best];
}

/* common.py:solve_lsq_trust_region */
static void solve_lsq_trust_region(int n, int m, const double *uf, const double *s, const double *V,
                                   double Delta, double *alpha_io, double *p) {
  double suf[MAXN], tmp[MAXN];
  for (int i = 0; i < n; i++) suf[i] = s[i] * uf[i];
  int full_rank = 0;
  if (m >= n) full_rank = s[n - 1] > EPS * m * s[0];
  if (full_rank) {
    for (int i = 0; i < n; i++) tmp[i] = uf[i] / s[i];
    for (int i = 0; i < n; i++) { double t = 0; for (int j = 0; j < n; j++) t += V[i * n + j] * tmp[j]; p[i] = -t; }
    if (vnorm(n, p) <= Delta) { *alpha_io = 0.0; return; }
  }
  double alpha_upper = vnorm(n, suf) / Delta, alpha_lower = 0.0, alpha = *alpha_io;
#define PHI(al, phi, phip) do { double pn2 = 0, sp = 0; \
    for (int i_ = 0; i_ < n; i_++) { double den = s[i_] * s[i_] + (al); double q = suf[i_] / den; pn2 += q * q; \
      sp += suf[i_] * suf[i_] / (den * den * den); } \
    double pn = sqrt(pn2); phi = pn - Delta; phip = -sp / pn; } while (0)
  if (full_rank) {
    double phi, phip;
    PHI(0.0, phi, phip);
    alpha_lower = -phi / phip;
  }
  if (!full_rank && alpha == 0)
    alpha = fmax(0.001 * alpha_upper, sqrt(alpha_lower * alpha_upper));
  for (int it = 0; it < 10; it++) {
    if (alpha < alpha_lower || alpha > alpha_upper)
      alpha = fmax(0.001 * alpha_upper, sqrt(alpha_lower * alpha_upper));
    double phi, phip;
    PHI(alpha, phi, phip);
    if (phi < 0) alpha_upper = alpha;
    double ratio = phi / phip;
    alpha_lower = fmax(alpha_lower, alpha - ratio);
    alpha -= (phi + Delta) * ratio / Delta;
    if (fabs(phi) < 0.01 * Delta) break;
  }
#undef PHI
  for (int i = 0; i < n; i++) tmp[i] = suf[i] / (s[i] * s[i] + alpha);
  for (int i = 0; i < n; i++) { double t = 0; for (int j = 0; j < n; j++) t += V[i * n + j] * tmp[j]; p[i] = -t; }
  double sc = Delta / vnorm(n, p);
  for (int i = 0; i < n; i++) p[i] *= sc;
  *alpha_io = alpha;
}

static int in_bounds(int n, const double *x, const double *lb, const double *ub) {
  for (int i = 0; i < n; i++) if (!(x[i] >= lb[i] && x[i] <= ub[i])) return 0;
  return 1;
}

/* trf.py:select_step ; p, p_h are modified in place like the original */
static double select_step(int m, int n, const double *x, const double *J, const double *d,
                          const double *diag_h, const double *g_h, double *p, double *p_h,
                          double Delta, const double *lb, const double *ub, double theta,
                          double *step, double *step_h) {
  double xt[MAXN];
  for (int i = 0; i < n; i++) xt[i] = x[i] + p[i];
  if (in_bounds(n, xt, lb, ub)) {
    memcpy(step, p, sizeof(double) * n); memcpy(step_h, p_h, sizeof(double) * n);
    return -evaluate_quadratic(m, n, J, d, g_h, p_h, diag_h);
  }
  int hits[MAXN];
  double p_stride = step_size_to_bound(n, x, p, lb, ub, hits);
  double r_h[MAXN], r[MAXN], x_on_bound[MAXN];
  for (int i = 0; i < n; i++) { r_h[i] = hits[i] ? -p_h[i] : p_h[i]; r[i] = d[i] * r_h[i]; }
  for (int i = 0; i < n; i++) { p[i] *= p_stride; p_h[i] *= p_stride; x_on_bound[i] = x[i] + p[i]; }
  /* intersect_trust_region(p_h, r_h, Delta) -> upper root */
  double to_tr;
  {
    double a = vdot(n, r_h, r_h), b = vdot(n, p_h, r_h), c = vdot(n, p_h, p_h) - Delta * Delta;
    double dd = sqrt(b * b - a * c);
    double q = -(b + copysign(dd, b));
    double t1 = q / a, t2 = c / q;
    to_tr = (t1 < t2) ? t2 : t1;
  }
  double to_bound = step_size_to_bound(n, x_on_bound, r, lb, ub, NULL);
  double r_stride = fmin(to_bound, to_tr), r_stride_l, r_stride_u;
  if (r_stride > 0) {
    r_stride_l = (1 - theta) * p_stride / r_stride;
    r_stride_u = (r_stride == to_bound) ? theta * to_bound : to_tr;
  } else { r_stride_l = 0; r_stride_u = -1; }
  double r_value;
  if (r_stride_l <= r_stride_u) {
    double a, b, c;
    build_quadratic_1d(m, n, J, d, g_h, r_h, diag_h, p_h, &a, &b, &c);
    r_stride = minimize_quadratic_1d(a, b, r_stride_l, r_stride_u, c, &r_value);
    for (int i = 0; i < n; i++) { r_h[i] = r_h[i] * r_stride + p_h[i]; r[i] = r_h[i] * d[i]; }
  } else r_value = INFINITY;
  for (int i = 0; i < n; i++) { p[i] *= theta; p_h[i] *= theta; }
  double p_value = evaluate_quadratic(m, n, J, d, g_h, p_h, diag_h);
  double ag_h[MAXN], ag[MAXN];
  for (int i = 0; i < n; i++) { ag_h[i] = -g_h[i]; ag[i] = d[i] * ag_h[i]; }
  to_tr = Delta / vnorm(n, ag_h);
  to_bound = step_size_to_bound(n, x, ag, lb, ub, NULL);
  double ag_stride = (to_bound < to_tr) ? theta * to_bound : to_tr;
  double a, b, c, ag_value;
  build_quadratic_1d(m, n, J, d, g_h, ag_h, diag_h, NULL, &a, &b, &c);
  ag_stride = minimize_quadratic_1d(a, b, 0, ag_stride, 0, &ag_value);
  for (int i = 0; i < n; i++) { ag_h[i] *= ag_stride; ag[i] *= ag_stride; }
  if (p_value < r_value && p_value < ag_value) {
    memcpy(step, p, sizeof(double) * n); memcpy(step_h, p_h, sizeof(double) * n); return -p_value;
  } else if (r_value < p_value && r_value < ag_value) {
    memcpy(step, r, sizeof(double) * n); memcpy(step_h, r_h, sizeof(double) * n); return -r_value;
  }
  memcpy(step, ag, sizeof(double) * n); memcpy(step_h, ag_h, sizeof(double) * n); return -ag_value;
}

typedef struct {
  double ftol, xtol, gtol;
  int max_nfev;
  int jac_mode;      /* 0 analytic, 1 2-point */
  int x_scale_jac;   /* 1: x_scale='jac' */
  double x_scale[MAXN];
} trf_opts;

/*
 * One voxel through curve_fit(method='trf').  status: scipy's 0..4, or
 *  -1 lb>=ub, -2 x0 out of bounds, -3 non-finite ydata, -4 non-finite f(x0).
 * Returns status; x (n), cov (n*n), nfev, cost.
 */
static int trf_one(prob_t *P, const trf_opts *O, const double *x0, double *x, double *cov,
                   int *nfev_out, int *njev_out, double *cost_out, double *opt_out) {
  int n = P->n, m = P->nb;
  const double *lb = P->lb, *ub = P->ub;
  *nfev_out = 0; *njev_out = 0; *cost_out = NAN; *opt_out = NAN;
  for (int i = 0; i < m; i++) if (!isfinite(P->y[i])) return -3;
  for (int i = 0; i < n; i++) if (!(lb[i] < ub[i])) return -1;
  if (!in_bounds(n, x0, lb, ub)) return -2;
  memcpy(x, x0, sizeof(double) * n);
  make_strictly_feasible(n, x, lb, ub, 1e-10);

  double f[MAXM], f_new[MAXM], J[MAXM * MAXN], g[MAXN];
  prob_fun(P, x, f);
  for (int i = 0; i < m; i++) if (!isfinite(f[i])) return -4;
  prob_jac(P, x, f, J);
  int nfev = 1, njev = 1;
  double cost = 0.5 * vdot(m, f, f);
  for (int k = 0; k < n; k++) { double t = 0; for (int i = 0; i < m; i++) t += J[i * n + k] * f[i]; g[k] = t; }

  double scale[MAXN], scale_inv[MAXN];
  if (O->x_scale_jac) {
    for (int k = 0; k < n; k++) {
      double t = 0; for (int i = 0; i < m; i++) t += J[i * n + k] * J[i * n + k];
      scale_inv[k] = sqrt(t); if (scale_inv[k] == 0) scale_inv[k] = 1; scale[k] = 1 / scale_inv[k];
    }
  } else for (int k = 0; k < n; k++) { scale[k] = O->x_scale[k]; scale_inv[k] = 1 / scale[k]; }

  double v[MAXN], dv[MAXN], d[MAXN], diag_h[MAXN], g_h[MAXN];
  cl_scaling(n, x, g, lb, ub, v, dv);
  for (int k = 0; k < n; k++) if (dv[k] != 0) v[k] *= scale_inv[k];
  double Delta;
  { double t = 0; for (int k = 0; k < n; k++) { double q = x[k] * scale_inv[k] / sqrt(v[k]); t += q * q; } Delta = sqrt(t); }
  if (Delta == 0) Delta = 1.0;
  double g_norm = 0, alpha = 0.0;
  int status = -99; /* None */
  double Jaug[(MAXM + MAXN) * MAXN], s[MAXN], V[MAXN * MAXN], uf[MAXN];

  for (;;) {
    cl_scaling(n, x, g, lb, ub, v, dv);
    g_norm = 0;
    for (int k = 0; k < n; k++) g_norm = fmax(g_norm, fabs(g[k] * v[k]));
    if (g_norm < O->gtol) status = 1;
    if (status != -99 || nfev == O->max_nfev) break;
    for (int k = 0; k < n; k++) if (dv[k] != 0) v[k] *= scale_inv[k];
    for (int k = 0; k < n; k++) { d[k] = sqrt(v[k]) * scale[k]; diag_h[k] = g[k] * dv[k] * scale[k]; g_h[k] = d[k] * g[k]; }
    /* augmented system and its SVD */
    for (int i = 0; i < m; i++) for (int k = 0; k < n; k++) Jaug[i * n + k] = J[i * n + k] * d[k];
    for (int i = 0; i < n; i++) for (int k = 0; k < n; k++) Jaug[(m + i) * n + k] = (i == k) ? sqrt(diag_h[k]) : 0.0;
    /* uf = U^T f_aug = (U s)^T f_aug / s  -- computed from the rotated columns */
    jacobi_svd(m + n, n, Jaug, s, V);
    for (int k = 0; k < n; k++) {
      double t = 0; for (int i = 0; i < m; i++) t += Jaug[i * n + k] * f[i];
      uf[k] = (s[k] > 0) ? t / s[k] : 0.0;
    }
    double theta = fmax(0.995, 1 - g_norm);
    double actual_reduction = -1, cost_new = cost, x_new[MAXN];
    while (actual_reduction <= 0 && nfev < O->max_nfev) {
      double p_h[MAXN], p[MAXN], step[MAXN], step_h[MAXN];
      solve_lsq_trust_region(n, m, uf, s, V, Delta, &alpha, p_h);
      for (int k = 0; k < n; k++) p[k] = d[k] * p_h[k];
      double predicted = select_step(m, n, x, J, d, diag_h, g_h, p, p_h, Delta, lb, ub, theta, step, step_h);
      for (int k = 0; k < n; k++) x_new[k] = x[k] + step[k];
      make_strictly_feasible(n, x_new, lb, ub, 0.0);
      prob_fun(P, x_new, f_new);
      nfev++;
      double step_h_norm = vnorm(n, step_h);
      int finite = 1;
      for (int i = 0; i < m; i++) if (!isfinite(f_new[i])) finite = 0;
      if (!finite) { Delta = 0.25 * step_h_norm; continue; }
      cost_new = 0.5 * vdot(m, f_new, f_new);
      actual_reduction = cost - cost_new;
      /* update_tr_radius */
      double ratio, Delta_new = Delta;
      if (predicted > 0) ratio = actual_reduction / predicted;
      else if (predicted == 0 && actual_reduction == 0) ratio = 1;
      else ratio = 0;
      if (ratio < 0.25) Delta_new = 0.25 * step_h_norm;
      else if (ratio > 0.75 && step_h_norm > 0.95 * Delta) Delta_new = Delta * 2.0;
      double step_norm = vnorm(n, step), x_norm = vnorm(n, x);
      int ft = (actual_reduction < O->ftol * cost) && ratio > 0.25;
      int xt = step_norm < O->xtol * (O->xtol + x_norm);
      if (ft && xt) status = 4; else if (ft) status = 2; else if (xt) status = 3;
      if (status != -99) break;
      alpha *= Delta / Delta_new;
      Delta = Delta_new;
    }
    if (actual_reduction > 0) {
      memcpy(x, x_new, sizeof(double) * n);
      memcpy(f, f_new, sizeof(double) * m);
      cost = cost_new;
      prob_jac(P, x, f, J);
      njev++;
      for (int k = 0; k < n; k++) { double t = 0; for (int i = 0; i < m; i++) t += J[i * n + k] * f[i]; g[k] = t; }
      if (O->x_scale_jac)
        for (int k = 0; k < n; k++) {
          double t = 0; for (int i = 0; i < m; i++) t += J[i * n + k] * J[i * n + k];
          scale_inv[k] = fmax(sqrt(t), scale_inv[k]); scale[k] = 1 / scale_inv[k];
        }
    }
  }
  if (status == -99) status = 0;
  *nfev_out = nfev; *njev_out = njev; *cost_out = cost; *opt_out = g_norm;

  /* curve_fit: pcov = pinv(J^T J) * 2*cost/(m-n), inf when m <= n */
  {
    double Jc[MAXM * MAXN];
    memcpy(Jc, J, sizeof(double) * m * n);
    jacobi_svd(m, n, Jc, s, V);
    double thr = EPS * (m > n ? m : n) * s[0];
    for (int i = 0; i < n; i++)
      for (int j = 0; j < n; j++) {
        double t = 0;
        for (int k = 0; k < n; k++) if (s[k] > thr) t += V[i * n + k] * V[j * n + k] / (s[k] * s[k]);
        cov[i * n + j] = t;
      }
    if (m > n) { double s_sq = 2 * cost / (m - n); for (int i = 0; i < n * n; i++) cov[i] *= s_sq; }
    else for (int i = 0; i < n * n; i++) cov[i] = INFINITY;
  }
  return status;
}

/* ------------------------------------------------------------------ */
/* method='dogbox': scipy/optimize/_lsq/dogbox.py (dogbox, dogleg_step, */
/* find_intersection), tr_solver='exact', loss='linear' — reached by     */
/* solvers/curvefit.py:295-306 when the TOML sets method = "dogbox".     */
/* J is kept explicitly here; lstsq(J_free, -f, rcond=-1) through the    */
/* one-sided Jacobi SVD (singular values <= eps * s_max are dropped).    */
/* ------------------------------------------------------------------ */
static double step_size_to_bound_hits(int n, const double *x, const double *s, const double *lb,
                                      const double *ub, int *hits) {
  double steps[MAXN], mn = INFINITY;
  for (int i = 0; i < n; i++) {
    steps[i] = INFINITY;
    if (s[i] != 0) steps[i] = fmax((lb[i] - x[i]) / s[i], (ub[i] - x[i]) / s[i]);
    if (steps[i] < mn) mn = steps[i];
  }
  for (int i = 0; i < n; i++) hits[i] = (steps[i] == mn) ? ((s[i] > 0) - (s[i] < 0)) : 0;
  return mn;
}

static int dogbox_one(prob_t *P, const trf_opts *O, const double *x0, double *x, double *cov,
                      int *nfev_out, int *njev_out, double *cost_out, double *opt_out) {
  int n = P->n, m = P->nb;
  const double *lb = P->lb, *ub = P->ub;
  *nfev_out = 0; *njev_out = 0; *cost_out = NAN; *opt_out = NAN;
  for (int i = 0; i < m; i++) if (!isfinite(P->y[i])) return -3;
  for (int i = 0; i < n; i++) if (!(lb[i] < ub[i])) return -1;
  if (!in_bounds(n, x0, lb, ub)) return -2;
  memcpy(x, x0, sizeof(double) * n); /* no make_strictly_feasible for dogbox (least_squares.py) */

  double f[MAXM], f_new[MAXM], J[MAXM * MAXN], g[MAXN];
  prob_fun(P, x, f);
  for (int i = 0; i < m; i++) if (!isfinite(f[i])) return -4;
  prob_jac(P, x, f, J);
  int nfev = 1, njev = 1;
  double cost = 0.5 * vdot(m, f, f);
  for (int k = 0; k < n; k++) { double t = 0; for (int i = 0; i < m; i++) t += J[i * n + k] * f[i]; g[k] = t; }
  double scale[MAXN], scale_inv[MAXN];
  if (O->x_scale_jac) {
    for (int k = 0; k < n; k++) {
      double t = 0; for (int i = 0; i < m; i++) t += J[i * n + k] * J[i * n + k];
      scale_inv[k] = sqrt(t); if (scale_inv[k] == 0) scale_inv[k] = 1; scale[k] = 1 / scale_inv[k];
    }
  } else for (int k = 0; k < n; k++) { scale[k] = O->x_scale[k]; scale_inv[k] = 1 / scale[k]; }
  double Delta = 0;
  for (int k = 0; k < n; k++) Delta = fmax(Delta, fabs(x0[k] * scale_inv[k]));
  if (Delta == 0) Delta = 1.0;
  int on_bound[MAXN];
  for (int k = 0; k < n; k++) { on_bound[k] = 0; if (x0[k] == lb[k]) on_bound[k] = -1; if (x0[k] == ub[k]) on_bound[k] = 1; }
  int status = -99;
  double g_norm = 0;

  for (;;) {
    int fr[MAXN], nf = 0; /* free_set: indices of the free variables */
    g_norm = 0;
    for (int k = 0; k < n; k++) {
      int active = on_bound[k] * g[k] < 0;
      if (!active) { fr[nf++] = k; g_norm = fmax(g_norm, fabs(g[k])); }
    }
    if (g_norm < O->gtol) status = 1;
    if (status != -99 || nfev == O->max_nfev) break;
    double Jf[MAXM * MAXN], gf[MAXN], xf[MAXN], lf[MAXN], uf_[MAXN], sf[MAXN];
    for (int a = 0; a < nf; a++) {
      gf[a] = g[fr[a]]; xf[a] = x[fr[a]]; lf[a] = lb[fr[a]]; uf_[a] = ub[fr[a]]; sf[a] = scale[fr[a]];
      for (int i = 0; i < m; i++) Jf[i * nf + a] = J[i * n + fr[a]];
    }
    /* newton_step = lstsq(J_free, -f, rcond=-1)[0] */
    double newton[MAXN];
    {
      double A[MAXM * MAXN], s[MAXN], V[MAXN * MAXN];
      memcpy(A, Jf, sizeof(double) * m * nf);
      jacobi_svd(m, nf, A, s, V);
      double coef[MAXN];
      for (int k = 0; k < nf; k++) {
        double t = 0;
        for (int i = 0; i < m; i++) t += A[i * nf + k] * (-f[i]); /* (U s)^T b */
        coef[k] = (nf > 0 && s[k] > EPS * s[0]) ? t / (s[k] * s[k]) : 0.0;
      }
      for (int a = 0; a < nf; a++) { double t = 0; for (int k = 0; k < nf; k++) t += V[a * nf + k] * coef[k]; newton[a] = t; }
    }
    /* a, b = build_quadratic_1d(J_free, g_free, -g_free) */
    double qa = 0, qb = 0;
    for (int i = 0; i < m; i++) { double t = 0; for (int a = 0; a < nf; a++) t += Jf[i * nf + a] * (-gf[a]); qa += t * t; }
    qa *= 0.5;
    for (int a = 0; a < nf; a++) qb += gf[a] * (-gf[a]);

    double actual_reduction = -1, cost_new = cost, x_new[MAXN], step[MAXN];
    int on_free[MAXN];
    while (actual_reduction <= 0 && nfev < O->max_nfev) {
      /* dogleg_step */
      double lbt[MAXN], ubt[MAXN], sfree[MAXN];
      int orig_l[MAXN], orig_u[MAXN], tr_l[MAXN], tr_u[MAXN], tr_hit = 0, inside = 1;
      for (int a = 0; a < nf; a++) {
        double trb = Delta * sf[a], lc = lf[a] - xf[a], uc = uf_[a] - xf[a];
        lbt[a] = fmax(lc, -trb); ubt[a] = fmin(uc, trb);
        orig_l[a] = lbt[a] == lc; orig_u[a] = ubt[a] == uc; tr_l[a] = lbt[a] == -trb; tr_u[a] = ubt[a] == trb;
        on_free[a] = 0;
        if (!(newton[a] >= lbt[a] && newton[a] <= ubt[a])) inside = 0;
      }
      if (inside) {
        memcpy(sfree, newton, sizeof(double) * nf);
      } else {
        double zero[MAXN] = {0}, ng[MAXN], cauchy[MAXN], diff[MAXN], yv;
        int hits[MAXN];
        for (int a = 0; a < nf; a++) ng[a] = -gf[a];
        double to_bounds = step_size_to_bound_hits(nf, zero, ng, lbt, ubt, hits);
        double t = minimize_quadratic_1d(qa, qb, 0.0, to_bounds, 0.0, &yv);
        for (int a = 0; a < nf; a++) { cauchy[a] = -t * gf[a]; diff[a] = newton[a] - cauchy[a]; }
        double step_size = step_size_to_bound_hits(nf, cauchy, diff, lbt, ubt, hits);
        for (int a = 0; a < nf; a++) {
          if (hits[a] < 0 && orig_l[a]) on_free[a] = -1;
          if (hits[a] > 0 && orig_u[a]) on_free[a] = 1;
          if ((hits[a] < 0 && tr_l[a]) || (hits[a] > 0 && tr_u[a])) tr_hit = 1;
          sfree[a] = cauchy[a] + step_size * diff[a];
        }
      }
      for (int k = 0; k < n; k++) step[k] = 0;
      for (int a = 0; a < nf; a++) step[fr[a]] = sfree[a];
      /* predicted_reduction = -evaluate_quadratic(J_free, g_free, step_free) */
      double q = 0, l = 0;
      for (int i = 0; i < m; i++) { double t = 0; for (int a = 0; a < nf; a++) t += Jf[i * nf + a] * sfree[a]; q += t * t; }
      for (int a = 0; a < nf; a++) l += sfree[a] * gf[a];
      double predicted = -(0.5 * q + l);
      for (int k = 0; k < n; k++) x_new[k] = fmin(fmax(x[k] + step[k], lb[k]), ub[k]);
      prob_fun(P, x_new, f_new);
      nfev++;
      double step_h_norm = 0;
      for (int k = 0; k < n; k++) step_h_norm = fmax(step_h_norm, fabs(step[k] * scale_inv[k]));
      int finite = 1;
      for (int i = 0; i < m; i++) if (!isfinite(f_new[i])) finite = 0;
      if (!finite) { Delta = 0.25 * step_h_norm; continue; }
      cost_new = 0.5 * vdot(m, f_new, f_new);
      actual_reduction = cost - cost_new;
      double ratio;
      if (predicted > 0) ratio = actual_reduction / predicted;
      else if (predicted == 0 && actual_reduction == 0) ratio = 1;
      else ratio = 0;
      if (ratio < 0.25) Delta = 0.25 * step_h_norm;
      else if (ratio > 0.75 && tr_hit) Delta *= 2.0;
      double step_norm = vnorm(n, step), x_norm = vnorm(n, x);
      int ft = (actual_reduction < O->ftol * cost) && ratio > 0.25;
      int xt = step_norm < O->xtol * (O->xtol + x_norm);
      if (ft && xt) status = 4; else if (ft) status = 2; else if (xt) status = 3;
      if (status != -99) break;
    }
    if (actual_reduction > 0) {
      for (int a = 0; a < nf; a++) on_bound[fr[a]] = on_free[a];
      memcpy(x, x_new, sizeof(double) * n);
      for (int k = 0; k < n; k++) { if (on_bound[k] == -1) x[k] = lb[k]; if (on_bound[k] == 1) x[k] = ub[k]; }
      memcpy(f, f_new, sizeof(double) * m);
      cost = cost_new;
      prob_jac(P, x, f, J);
      njev++;
      for (int k = 0; k < n; k++) { double t = 0; for (int i = 0; i < m; i++) t += J[i * n + k] * f[i]; g[k] = t; }
      if (O->x_scale_jac)
        for (int k = 0; k < n; k++) {
          double t = 0; for (int i = 0; i < m; i++) t += J[i * n + k] * J[i * n + k];
          scale_inv[k] = fmax(sqrt(t), scale_inv[k]); scale[k] = 1 / scale_inv[k];
        }
    }
  }
  if (status == -99) status = 0;
  *nfev_out = nfev; *njev_out = njev; *cost_out = cost; *opt_out = g_norm;
  {
    double Jc[MAXM * MAXN], s[MAXN], V[MAXN * MAXN];
    memcpy(Jc, J, sizeof(double) * m * n);
    jacobi_svd(m, n, Jc, s, V);
    double thr = EPS * (m > n ? m : n) * s[0];
    for (int i = 0; i < n; i++)
      for (int j = 0; j < n; j++) {
        double t = 0;
        for (int k = 0; k < n; k++) if (s[k] > thr) t += V[i * n + k] * V[j * n + k] / (s[k] * s[k]);
        cov[i * n + j] = t;
      }
    if (m > n) { double s_sq = 2 * cost / (m - n); for (int i = 0; i < n * n; i++) cov[i] *= s_sq; }
    else for (int i = 0; i < n * n; i++) cov[i] = INFINITY;
  }
  return status;
}

/*
 * Batch entry point.  Layouts: y (n_vox, nb) row-major; p0/lb/ub (n_vox, n_all)
 * row-major over the model's FULL parameter list (entries of frozen parameters
 * in p0 carry the fixed value; their bounds are ignored); frozen[n_all] flags.
 * Outputs: params (n_vox, n_all) (frozen entries = fixed value), cov
 * (n_vox, n_free, n_free), status, nfev, cost.  On failure (status <= 0)
 * params = p0 and cov = NaN, as solvers/curvefit.py:308-317 does.
 */
int pnbo_lsq_fit(int method, int model_id, int t1_mode, double tr, double tm, int nb, const double *b,
                 long n_vox, const double *y, const double *p0, const double *lb, const double *ub,
                 const int *frozen, double ftol, double xtol, double gtol, int max_nfev,
                 int jac_mode, int x_scale_jac, const double *x_scale, double *params, double *cov,
                 int *status, int *nfev, double *cost) {
  model_t M = {model_id, t1_mode, tr, tm, base_count(model_id) + (t1_mode ? 1 : 0)};
  int na = M.n_all, n = 0, fidx[MAXN];
  if (nb > MAXM || na > MAXN) return -1;
  for (int k = 0; k < na; k++) if (!frozen || !frozen[k]) fidx[n++] = k;
  trf_opts O;
  O.ftol = ftol; O.xtol = xtol; O.gtol = gtol; O.max_nfev = max_nfev; O.jac_mode = jac_mode;
  O.x_scale_jac = x_scale_jac;
  for (int k = 0; k < n; k++) O.x_scale[k] = x_scale ? x_scale[fidx[k]] : 1.0;
#pragma omp parallel for schedule(dynamic, 64)
  for (long vx = 0; vx < n_vox; vx++) {
    prob_t P;
    P.M = &M; P.nb = nb; P.n = n; P.b = b; P.y = y + vx * nb; P.jac_mode = jac_mode;
    double x0[MAXN], l[MAXN], u[MAXN], x[MAXN], c[MAXN * MAXN];
    for (int k = 0; k < na; k++) P.pfull[k] = p0[vx * na + k];
    for (int k = 0; k < n; k++) {
      P.free_idx[k] = fidx[k];
      x0[k] = p0[vx * na + fidx[k]]; l[k] = lb[vx * na + fidx[k]]; u[k] = ub[vx * na + fidx[k]];
    }
    P.lb = l; P.ub = u;
    int nf, nj; double cs, op;
    int st = method == 1 ? dogbox_one(&P, &O, x0, x, c, &nf, &nj, &cs, &op)
                         : trf_one(&P, &O, x0, x, c, &nf, &nj, &cs, &op);
    status[vx] = st; nfev[vx] = nf; cost[vx] = cs;
    for (int k = 0; k < na; k++) params[vx * na + k] = p0[vx * na + k];
    if (st > 0) {
      for (int k = 0; k < n; k++) params[vx * na + fidx[k]] = x[k];
      memcpy(cov + vx * n * n, c, sizeof(double) * n * n);
    } else
      for (int k = 0; k < n * n; k++) cov[vx * n * n + k] = NAN;
  }
  return 0;
}

/* method = 'trf' (the historical entry point) */
int pnbo_trf_fit(int model_id, int t1_mode, double tr, double tm, int nb, const double *b,
                 long n_vox, const double *y, const double *p0, const double *lb, const double *ub,
                 const int *frozen, double ftol, double xtol, double gtol, int max_nfev,
                 int jac_mode, int x_scale_jac, const double *x_scale, double *params, double *cov,
                 int *status, int *nfev, double *cost) {
  return pnbo_lsq_fit(0, model_id, t1_mode, tr, tm, nb, b, n_vox, y, p0, lb, ub, frozen, ftol, xtol, gtol,
                      max_nfev, jac_mode, x_scale_jac, x_scale, params, cov, status, nfev, cost);
}

/* ------------------------------------------------------------------ */
/* Lawson-Hanson NNLS (L&H 1974 ch. 23; Householder / Givens form)      */
/* ------------------------------------------------------------------ */

/* Householder: construct (mode 1) and/or apply (mode 2), L&H H12.  Pivot row
   lpivot, rows l1..m-1 are zeroed.  u is column `u` with stride ue. */
static void h12(int mode, int lpivot, int l1, int m, double *u, int ue, double *up, double *c,
                int ce, int cv, int ncv) {
  if (lpivot < 0 || lpivot >= l1 || l1 >= m) return;
  double cl = fabs(u[lpivot * ue]);
  if (mode == 1) {
    for (int j = l1; j < m; j++) cl = fmax(cl, fabs(u[j * ue]));
    if (cl <= 0) return;
    double clinv = 1.0 / cl, sm = (u[lpivot * ue] * clinv) * (u[lpivot * ue] * clinv);
    for (int j = l1; j < m; j++) sm += (u[j * ue] * clinv) * (u[j * ue] * clinv);
    cl *= sqrt(sm);
    if (u[lpivot * ue] > 0) cl = -cl;
    *up = u[lpivot * ue] - cl;
    u[lpivot * ue] = cl;
  } else if (cl <= 0) return;
  if (ncv <= 0) return;
  double b = (*up) * u[lpivot * ue];
  if (b >= 0) return;
  b = 1.0 / b;
  for (int j = 0; j < ncv; j++) {
    double *cj = c + (size_t)j * cv;
    double sm = cj[lpivot * ce] * (*up);
    for (int i = l1; i < m; i++) sm += cj[i * ce] * u[i * ue];
    if (sm != 0) {
      sm *= b;
      cj[lpivot * ce] += sm * (*up);
      for (int i = l1; i < m; i++) cj[i * ce] += sm * u[i * ue];
    }
  }
}

static void g1(double a, double b, double *c, double *s, double *sig) {
  if (fabs(a) > fabs(b)) {
    double xr = b / a, yr = sqrt(1 + xr * xr);
    *c = copysign(1.0 / yr, a); *s = (*c) * xr; *sig = fabs(a) * yr;
  } else if (b != 0) {
    double xr = a / b, yr = sqrt(1 + xr * xr);
    *s = copysign(1.0 / yr, b); *c = (*s) * xr; *sig = fabs(b) * yr;
  } else { *sig = 0; *c = 0; *s = 1; }
}

/*
 * A: m x n column-major working copy (destroyed); b: m (destroyed).
 * Returns mode: 1 ok, 3 iteration count exceeded.  x (n), rnorm.
 */
static int nnls_one(int m, int n, double *A, double *b, int itmax, double *x, double *rnorm,
                    double *w, double *zz, int *index, int *iters) {
  const double factor = 0.01;
  int mode = 1, iter = 0, nsetp = 0, iz1 = 0, iz2 = n - 1;
#define AA(i, j) A[(size_t)(j) * m + (i)]
  for (int i = 0; i < n; i++) { x[i] = 0; index[i] = i; }
  for (;;) {
    if (iz1 > iz2 || nsetp >= m) break;
    for (int iz = iz1; iz <= iz2; iz++) {
      int j = index[iz];
      double sm = 0;
      for (int l = nsetp; l < m; l++) sm += AA(l, j) * b[l];
      w[j] = sm;
    }
    int izmax = -1, j = -1;
    double up = 0;
    for (;;) {
      double wmax = 0;
      izmax = -1;
      for (int iz = iz1; iz <= iz2; iz++) {
        int jj = index[iz];
        if (w[jj] > wmax) { wmax = w[jj]; izmax = iz; }
      }
      if (wmax <= 0 || izmax < 0) goto terminate;
      j = index[izmax];
      double asave = AA(nsetp, j);
      h12(1, nsetp, nsetp + 1, m, &AA(0, j), 1, &up, NULL, 1, 1, 0);
      double unorm = 0;
      for (int l = 0; l < nsetp; l++) unorm += AA(l, j) * AA(l, j);
      unorm = sqrt(unorm);
      if ((unorm + fabs(AA(nsetp, j)) * factor) - unorm > 0) {
        memcpy(zz, b, sizeof(double) * m);
        h12(2, nsetp, nsetp + 1, m, &AA(0, j), 1, &up, zz, 1, 1, 1);
        double ztest = zz[nsetp] / AA(nsetp, j);
        if (ztest > 0) break;
      }
      AA(nsetp, j) = asave;
      w[j] = 0;
    }
    memcpy(b, zz, sizeof(double) * m);
    index[izmax] = index[iz1];
    index[iz1] = j;
    iz1++;
    nsetp++;
    for (int jz = iz1; jz <= iz2; jz++) {
      int jj = index[jz];
      h12(2, nsetp - 1, nsetp, m, &AA(0, j), 1, &up, &AA(0, jj), 1, m, 1);
    }
    for (int l = nsetp; l < m; l++) AA(l, j) = 0;
    w[j] = 0;
    /* solve the triangular system */
    for (int l = 0; l < nsetp; l++) {
      int ip = nsetp - 1 - l;
      if (l != 0) { int jj2 = index[ip + 1]; for (int ii = 0; ii <= ip; ii++) zz[ii] -= AA(ii, jj2) * zz[ip + 1]; }
      zz[ip] /= AA(ip, index[ip]);
    }
    /* secondary loop */
    for (;;) {
      iter++;
      if (iter >= itmax) { mode = 3; goto terminate; } /* SciPy 1.18: fails once the count reaches maxiter */
      double alpha = 2.0;
      int jj = -1;
      for (int ip = 0; ip < nsetp; ip++) {
        int l = index[ip];
        if (zz[ip] <= 0) {
          double t = -x[l] / (zz[ip] - x[l]);
          if (alpha > t) { alpha = t; jj = ip; }
        }
      }
      if (alpha == 2.0) break;
      for (int ip = 0; ip < nsetp; ip++) { int l = index[ip]; x[l] += alpha * (zz[ip] - x[l]); }
      int i = index[jj];
      for (;;) {
        x[i] = 0;
        if (jj != nsetp - 1) {
          jj++;
          for (int jc = jj; jc < nsetp; jc++) {
            int ii = index[jc];
            index[jc - 1] = ii;
            double cc, ss, sig;
            g1(AA(jc - 1, ii), AA(jc, ii), &cc, &ss, &sig);
            AA(jc - 1, ii) = sig;
            AA(jc, ii) = 0;
            for (int l = 0; l < n; l++)
              if (l != ii) {
                double t = AA(jc - 1, l);
                AA(jc - 1, l) = cc * t + ss * AA(jc, l);
                AA(jc, l) = -ss * t + cc * AA(jc, l);
              }
            double t = b[jc - 1];
            b[jc - 1] = cc * t + ss * b[jc];
            b[jc] = -ss * t + cc * b[jc];
          }
        }
        nsetp--;
        iz1--;
        index[iz1] = i;
        /* all coefficients in P must stay feasible (round-off guard) */
        int again = 0;
        for (jj = 0; jj < nsetp; jj++) {
          i = index[jj];
          if (x[i] <= 0) { again = 1; break; }
        }
        if (!again) break;
      }
      memcpy(zz, b, sizeof(double) * m);
      for (int l = 0; l < nsetp; l++) {
        int ip = nsetp - 1 - l;
        if (l != 0) { int jj2 = index[ip + 1]; for (int ii = 0; ii <= ip; ii++) zz[ii] -= AA(ii, jj2) * zz[ip + 1]; }
        zz[ip] /= AA(ip, index[ip]);
      }
    }
    for (int ip = 0; ip < nsetp; ip++) x[index[ip]] = zz[ip];
  }
terminate: {
    double sm = 0;
    for (int i = nsetp; i < m; i++) sm += b[i] * b[i];
    *rnorm = sqrt(sm);
  }
#undef AA
  *iters = iter;
  return mode;
}

/*
 * Batch NNLS on a shared matrix.  A: (m, n) ROW-major (numpy C order);
 * B: (n_vox, m) right-hand sides.  Failure (mode 3) -> zeros and ||b||, as
 * solvers/nnls_solver.py:201-210 does.
 */
int pnbo_nnls(int m, int n, const double *A, long n_vox, const double *B, int maxiter, double *X,
              double *rnorm, int *status, int *iters) {
#pragma omp parallel
  {
    double *Aw = malloc(sizeof(double) * m * n), *bw = malloc(sizeof(double) * m);
    double *w = malloc(sizeof(double) * n), *zz = malloc(sizeof(double) * m);
    int *idx = malloc(sizeof(int) * n);
#pragma omp for schedule(dynamic, 16)
    for (long v = 0; v < n_vox; v++) {
      for (int i = 0; i < m; i++) for (int j = 0; j < n; j++) Aw[(size_t)j * m + i] = A[(size_t)i * n + j];
      memcpy(bw, B + v * m, sizeof(double) * m);
      int it;
      int mode = nnls_one(m, n, Aw, bw, maxiter, X + v * n, rnorm + v, w, zz, idx, &it);
      status[v] = mode; iters[v] = it;
      if (mode == 3) {
        double sm = 0;
        for (int i = 0; i < m; i++) sm += B[v * m + i] * B[v * m + i];
        rnorm[v] = sqrt(sm);
        for (int j = 0; j < n; j++) X[v * n + j] = 0;
      }
    }
    free(Aw); free(bw); free(w); free(zz); free(idx);
  }
  return 0;
}
