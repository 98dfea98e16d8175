// TEST INFRASTRUCTURE ONLY.  Compiles the per-voxel device mathematics
// (pyneapple_b200/csrc/pnb_trf_core.cuh) with g++ so that it can be compared
// with the oracle in the GPU-less authoring container.  The product never
// loads this library; pyneapple_b200 only ever calls the CUDA build.
#define PNB_HOST_SIM 1
#include <cstring>
#include <cmath>
using std::exp; using std::sqrt; using std::fabs; using std::copysign; using std::nextafter;
#include "../../pyneapple_b200/csrc/pnb_dogbox_core.cuh"
#include "../../pyneapple_b200/csrc/pnb_lm_core.cuh"

using namespace pnb;

template <class M, bool X = false>
static void run_one(const TrfOptions &O, int m, const double *b, const double *y, const double *p0,
                    const double *lb, const double *ub, double *params, double *cov, int *status,
                    int *nfev, double *cost, const double *w = nullptr) {
  constexpr int N = M::NP;
  TrfLane<M> S;
  double p0v[N];
  for (int i = 0; i < N; i++) p0v[i] = p0[i];
  bool yfin = true;
  for (int r = 0; r < m; r++) yfin = yfin && finite_d(y[r]);
  auto yb = [&](int r, double &yv, double &bv) { yv = y[r]; bv = b[r]; };
  double c, g[N], A[N][N];
  bool running = false;
  S.status = kStRunning;
  if (O.method == 0) {
    running = trf_begin<M>(S, O, p0v, lb, ub, 1, yfin);
    if (running) {
      trf_evaluate<M, X>(S.x, O, m, yb, lb, ub, 1, c, g, A, w);
      running = trf_after_first_eval<M>(S, O, c, g, A, lb, ub, 1);
    }
  }
  DogboxLane<M> DB;
  if (O.method == 1 && S.status == kStRunning) {
    // the same calls, in the same order, as trf_kernel<M, BLOCK, 1>
    running = trf_begin<M>(S, O, p0v, lb, ub, 1, yfin);
    if (running) {
      trf_evaluate<M, X>(S.x, O, m, yb, lb, ub, 1, c, g, A, w);
      running = dbx_after_first_eval<M>(S, DB, O, c, g, A, lb, ub, 1);
    }
    while (running) {
      if (S.need_prologue) {
        if (!dbx_prologue<M>(S, DB, O)) break;
        S.need_prologue = false;
      }
      dbx_trial<M>(S, DB, O, lb, ub, 1);
      trf_evaluate<M, X>(S.x_new, O, m, yb, lb, ub, 1, c, g, A, w);
      S.need_prologue = dbx_after_trial<M>(S, DB, O, c, g, A, lb, ub, 1);
    }
  }
  LmLane<M> LM;
  if (O.method == 2 && S.status == kStRunning) {
    // the same calls, in the same order, as trf_kernel<M, BLOCK, 2>
    running = trf_begin<M>(S, O, p0v, lb, ub, 1, yfin);
    if (running) {
      trf_evaluate<M, X>(S.x, O, m, yb, lb, ub, 1, c, g, A, w);
      running = lm_after_first_eval<M>(S, LM, O, c, g, A);
    }
    while (running) {
      if (S.need_prologue) {
        if (!lm_prologue<M>(S, LM, O)) break;
        S.need_prologue = false;
      }
      lm_trial<M>(S, LM, O);
      trf_evaluate<M, X>(S.x_new, O, m, yb, lb, ub, 1, c, g, A, w);
      S.need_prologue = lm_after_trial<M>(S, LM, O, c, g, A);
    }
  }
  while (running && O.method == 0) {
    if (S.need_prologue) {
      if (!trf_prologue<M>(S, O, lb, ub, 1)) break;
      S.need_prologue = false;
    }
    double p_h[N];
    trf_solve_tr<M>(S, p_h);
    trf_select_step<M>(S, p_h, lb, ub, 1, O.frozen);
    trf_evaluate<M, X>(S.x_new, O, m, yb, lb, ub, 1, c, g, A, w);
    S.need_prologue = trf_after_trial<M>(S, O, c, g, A);
  }
  *status = S.status; *nfev = S.nfev; *cost = S.cost;
  int nf = 0;
  for (int i = 0; i < N; i++) nf += ((O.frozen >> i) & 1u) ? 0 : 1;
  if (S.status > 0) {
    for (int i = 0; i < N; i++) params[i] = S.x[i];
    trf_covariance<M>(S, O, m, cov);
  } else {
    for (int i = 0; i < N; i++) params[i] = p0[i];
    for (int i = 0; i < nf * nf; i++) cov[i] = NAN;
  }
}

template <class M>
static void run_all(const TrfOptions &O, int m, const double *b, long n_vox, const double *y,
                    const double *p0, const double *lb, const double *ub, double *params, double *cov,
                    int *status, int *nfev, double *cost, const double *w = nullptr, bool extras = false) {
  constexpr int N = M::NP;
  int nf = 0;
  for (int i = 0; i < N; i++) nf += ((O.frozen >> i) & 1u) ? 0 : 1;
  for (long v = 0; v < n_vox; v++) {
    if (extras)  // the EXTRAS instantiation of trf_evaluate (weights, robust loss), as trf_kernel<M, BLOCK, METHOD, true>
      run_one<M, true>(O, m, b, y + v * m, p0 + v * N, lb + v * N, ub + v * N, params + v * N,
                       cov + v * nf * nf, status + v, nfev + v, cost + v, w);
    else
      run_one<M>(O, m, b, y + v * m, p0 + v * N, lb + v * N, ub + v * N, params + v * N,
                 cov + v * nf * nf, status + v, nfev + v, cost + v);
  }
}

extern "C" int pnbh_trf_fit(int model_id, int t1_mode, double tr, double tm, int nb, const double *b,
                            long n_vox, const double *y, const double *p0, const double *lb,
                            const double *ub, const int *frozen, double ftol, double xtol, double gtol,
                            int max_nfev, int jac_mode, int x_scale_jac, const double *x_scale,
                            double *params, double *cov, int *status, int *nfev, double *cost, int method) {
  TrfOptions O;
  O.ftol = ftol; O.xtol = xtol; O.gtol = gtol; O.max_nfev = max_nfev; O.jac_mode = jac_mode;
  O.x_scale_jac = x_scale_jac; O.frozen = 0; O.tr = tr; O.tm = tm; O.method = method;
  for (int i = 0; i < 8; i++) { O.x_scale[i] = x_scale ? x_scale[i] : 1.0; if (frozen && i < 7 && frozen[i]) O.frozen |= 1u << i; }
#define CASE(ID, T) if (model_id == ID && t1_mode == T) { run_all<Model<ID, T>>(O, nb, b, n_vox, y, p0, lb, ub, params, cov, status, nfev, cost); return 0; }
  CASE(0, 0) CASE(1, 0) CASE(2, 0) CASE(3, 0) CASE(4, 0) CASE(5, 0) CASE(6, 0)
  CASE(0, 1) CASE(1, 1) CASE(3, 1) CASE(0, 2) CASE(3, 2) CASE(4, 1) CASE(6, 2)
  CASE(2, 1) CASE(4, 2) CASE(6, 1) CASE(5, 2)
  return -1;
}

// the same with the curve_fit extras: weights = 1 / sigma per row (or null), robust loss, relative
// finite-difference steps, absolute_sigma; mono-exponential, bi-exponential S0 and reduced tri-exponential models
extern "C" int pnbh_trf_fit_extras(int model_id, int nb, const double *b, long n_vox, const double *y, const double *p0,
                                   const double *lb, const double *ub, double ftol, double xtol, double gtol,
                                   int max_nfev, int jac_mode, int method, const double *w, int loss, double f_scale,
                                   const double *diff_step, int absolute_sigma, int use_extras_kernel,
                                   double *params, double *cov, int *status, int *nfev, double *cost) {
  TrfOptions O;
  O.ftol = ftol; O.xtol = xtol; O.gtol = gtol; O.max_nfev = max_nfev; O.jac_mode = jac_mode;
  O.x_scale_jac = 0; O.frozen = 0; O.tr = 0.0; O.tm = 0.0; O.method = method;
  for (int i = 0; i < 8; i++) { O.x_scale[i] = 1.0; O.diff_step[i] = diff_step ? diff_step[i] : 0.0; }
  O.loss = loss; O.f_scale = f_scale; O.absolute_sigma = absolute_sigma;
#define XCASE(ID) if (model_id == ID) { run_all<Model<ID, 0>>(O, nb, b, n_vox, y, p0, lb, ub, params, cov, status, nfev, cost, w, use_extras_kernel != 0); return 0; }
  XCASE(0) XCASE(3) XCASE(4)
  return -1;
}

// pnb_exp (pnb_hd.cuh) on an array — checked against libm in tests/test_hostsim_core.py
extern "C" void pnbh_exp(long n, const double *x, double *out) {
  for (long i = 0; i < n; i++) out[i] = pnb::pnb_exp(x[i]);
}
