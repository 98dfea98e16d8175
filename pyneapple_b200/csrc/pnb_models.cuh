// Signal models of the multi-exponential family, one row (one b-value) at a time.
//
//   S(b) = A * C(T1) * sum_k w_k * exp(-b * D_k)
//
// A   : amplitude, parameter S0 or the constant 1
// w_k : a free fraction parameter, the implied fraction 1 - sum(free fractions)
//       (last component of the "reduced"/"S0" layouts), or the constant 1 (mono)
// C   : optional T1 / STEAM factor, (1 - exp(-TR/T1)) [* exp(-TM/T1)]
//
// Replaces the per-call numpy evaluation of
//   reference model_functions/multiexp.py:35-202 (forward equations),
//   models/monoexp.py:120-163, models/biexp.py:151-221, models/triexp.py:170-247
//   (analytic Jacobians), model_functions/multiexp.py:240-302 (T1 Jacobian).
// Parameter order per layout is the reference's `_all_param_names`.
#pragma once
#include "pnb_hd.cuh"

namespace pnb {

enum ModelId {
  kMono = 0,       // [S0, D]
  kBiReduced = 1,  // [f1, D1, D2]
  kBiFull = 2,     // [f1, D1, f2, D2]
  kBiS0 = 3,       // [f1, D1, D2, S0]
  kTriReduced = 4, // [f1, D1, f2, D2, D3]
  kTriFull = 5,    // [f1, D1, f2, D2, f3, D3]
  kTriS0 = 6,      // [f1, D1, f2, D2, D3, S0]
  kNumModels = 7
};

constexpr int kFracImplied = -1;
constexpr int kFracOne = -2;

template <int ID> struct Layout;
template <> struct Layout<kMono> {
  static constexpr int K = 1, NBASE = 2, AMP = 0;
  PNB_HD static constexpr int d(int) { return 1; }
  PNB_HD static constexpr int f(int) { return kFracOne; }
};
template <> struct Layout<kBiReduced> {
  static constexpr int K = 2, NBASE = 3, AMP = -1;
  PNB_HD static constexpr int d(int k) { return k == 0 ? 1 : 2; }
  PNB_HD static constexpr int f(int k) { return k == 0 ? 0 : kFracImplied; }
};
template <> struct Layout<kBiFull> {
  static constexpr int K = 2, NBASE = 4, AMP = -1;
  PNB_HD static constexpr int d(int k) { return k == 0 ? 1 : 3; }
  PNB_HD static constexpr int f(int k) { return k == 0 ? 0 : 2; }
};
template <> struct Layout<kBiS0> {
  static constexpr int K = 2, NBASE = 4, AMP = 3;
  PNB_HD static constexpr int d(int k) { return k == 0 ? 1 : 2; }
  PNB_HD static constexpr int f(int k) { return k == 0 ? 0 : kFracImplied; }
};
template <> struct Layout<kTriReduced> {
  static constexpr int K = 3, NBASE = 5, AMP = -1;
  PNB_HD static constexpr int d(int k) { return k == 0 ? 1 : (k == 1 ? 3 : 4); }
  PNB_HD static constexpr int f(int k) { return k == 0 ? 0 : (k == 1 ? 2 : kFracImplied); }
};
template <> struct Layout<kTriFull> {
  static constexpr int K = 3, NBASE = 6, AMP = -1;
  PNB_HD static constexpr int d(int k) { return k == 0 ? 1 : (k == 1 ? 3 : 5); }
  PNB_HD static constexpr int f(int k) { return k == 0 ? 0 : (k == 1 ? 2 : 4); }
};
template <> struct Layout<kTriS0> {
  static constexpr int K = 3, NBASE = 6, AMP = 5;
  PNB_HD static constexpr int d(int k) { return k == 0 ? 1 : (k == 1 ? 3 : 4); }
  PNB_HD static constexpr int f(int k) { return k == 0 ? 0 : (k == 1 ? 2 : kFracImplied); }
};

// T1MODE: 0 none, 1 standard, 2 STEAM.  The T1 parameter is the last one.
template <int ID, int T1MODE> struct Model {
  using L = Layout<ID>;
  static constexpr int K = L::K;
  static constexpr int NBASE = L::NBASE;
  static constexpr int NP = NBASE + (T1MODE ? 1 : 0);
  static constexpr bool kHasImplied = (L::f(K - 1) == kFracImplied);

  // row-independent quantities of one parameter vector
  struct Point {
    double amp;       // A
    double w[K];      // component weights
    double nd[K];     // -D_k
    double t1fac;     // C(T1)
    double t1dfac;    // dC/dT1
  };

  PNB_HD static void prepare(const double (&p)[NP], double tr, double tm, Point &pt) {
    pt.amp = (L::AMP >= 0) ? p[L::AMP >= 0 ? L::AMP : 0] : 1.0;
    double implied = 1.0;
#pragma unroll
    for (int k = 0; k < K; k++)
      if (L::f(k) >= 0) implied = implied - p[L::f(k) >= 0 ? L::f(k) : 0];
#pragma unroll
    for (int k = 0; k < K; k++) {
      pt.w[k] = (L::f(k) >= 0) ? p[L::f(k) >= 0 ? L::f(k) : 0]
                               : (L::f(k) == kFracImplied ? implied : 1.0);
      pt.nd[k] = -p[L::d(k)];
    }
    pt.t1fac = 1.0;
    pt.t1dfac = 0.0;
    if (T1MODE) {
      const double t1 = p[NP - 1];
      const double e_tr = exp(-tr / t1);
      const double a = 1.0 - e_tr;
      if (T1MODE == 2) {
        const double e_tm = exp(-tm / t1);
        pt.t1fac = a * e_tm;
        pt.t1dfac = e_tm / (t1 * t1) * (-tr * e_tr + tm * a);
      } else {
        pt.t1fac = a;
        pt.t1dfac = -e_tr * tr / (t1 * t1);
      }
    }
  }

  // The K exponentials of a row are independent ~10-deep FP64 chains.  With the range check inside
  // each of them every chain sits in its own branch region and the compiler cannot interleave them;
  // here the table path runs unconditionally for all K, ONE test covers the row, and the rare row
  // with an argument outside [-690, 690] is redone by the full pnb_exp as a call (same bits for the
  // in-range arguments: it is the same code).
  PNB_HD static void exps(const Point &pt, double b, double (&e)[K]) {
#if defined(PNB_EXP_PER_ELEMENT_CHECK)
#pragma unroll
    for (int k = 0; k < K; k++) e[k] = pnb_exp(b * pt.nd[k]);
#else
    double x[K];
    bool in_range = true;
#pragma unroll
    for (int k = 0; k < K; k++) {
      x[k] = b * pt.nd[k];
      in_range = in_range && (fabs(x[k]) < 690.0);
    }
#pragma unroll
    for (int k = 0; k < K; k++) e[k] = pnb_exp_core(x[k]);
    if (__builtin_expect(!in_range, 0)) {
#pragma unroll
      for (int k = 0; k < K; k++) e[k] = pnb_exp_cold(x[k]);
    }
#endif
  }
  PNB_HD static double combine(const Point &pt, const double (&e)[K]) {
    double shape = 0.0;
#pragma unroll
    for (int k = 0; k < K; k++) shape = (k == 0) ? pt.w[k] * e[k] : shape + pt.w[k] * e[k];
    double v = (L::AMP >= 0) ? pt.amp * shape : shape;
    if (T1MODE) v = v * pt.t1fac;
    return v;
  }
  // signal only
  PNB_HD static double value(const Point &pt, double b) {
    double e[K];
    exps(pt, b, e);
    return combine(pt, e);
  }
  // Finite-difference Jacobian: the point pk[j] differs from p0 (exponentials e0) only in parameter
  // j, and only a changed D_k needs a new exponential.  For the ~1e-8 steps of the 2-point scheme
  // that one is e0 * exp(b * (nd' - nd)) with a 4-term series (|b dnd| < 1e-4: truncation < 5e-18
  // relative) instead of a second full exp().  As in exps(): the series of all K components run
  // unconditionally, ONE test covers the row, the rare large step is redone as a call.
  template <class PK>
  PNB_HD static void perturbed_exps(const PK &pk, const Point &p0, double b, const double (&e0)[K],
                                    double (&ep)[K]) {
    double t[K];
    bool small = true;
#pragma unroll
    for (int k = 0; k < K; k++) {
      t[k] = b * (pk[L::d(k)].nd[k] - p0.nd[k]);
      small = small && (fabs(t[k]) < 1e-4);
    }
#pragma unroll
    for (int k = 0; k < K; k++) ep[k] = e0[k] + e0[k] * (t[k] * (1.0 + t[k] * (0.5 + t[k] * (1.0 / 6.0))));
    if (__builtin_expect(!small, 0)) {
#pragma unroll
      for (int k = 0; k < K; k++)
        if (!(fabs(t[k]) < 1e-4)) ep[k] = pnb_exp_cold(b * pk[L::d(k)].nd[k]);
    }
  }
  // signal at pk = pk[j] from the exponentials of p0 (e0) and the perturbed ones (ep)
  PNB_HD static double value_perturbed(const Point &pk, const double (&e0)[K], const double (&ep)[K], int j) {
    double e[K];
#pragma unroll
    for (int k = 0; k < K; k++) e[k] = (L::d(k) == j) ? ep[k] : e0[k];
    return combine(pk, e);
  }

  // signal and d(signal)/d(p_j) for all NP parameters
  PNB_HD static double value_grad(const Point &pt, double b, double (&grad)[NP]) {
    double e[K];
    double shape = 0.0;
    exps(pt, b, e);
#pragma unroll
    for (int k = 0; k < K; k++) shape = (k == 0) ? pt.w[k] * e[k] : shape + pt.w[k] * e[k];
#pragma unroll
    for (int k = 0; k < K; k++) {
      grad[L::d(k)] = (L::AMP >= 0) ? -b * pt.amp * pt.w[k] * e[k] : -b * pt.w[k] * e[k];
      if (L::f(k) >= 0)
        grad[L::f(k) >= 0 ? L::f(k) : 0] =
            (L::AMP >= 0) ? (kHasImplied ? pt.amp * (e[k] - e[K - 1]) : pt.amp * e[k])
                          : (kHasImplied ? (e[k] - e[K - 1]) : e[k]);
    }
    if (L::AMP >= 0) grad[L::AMP >= 0 ? L::AMP : 0] = shape;
    const double base = (L::AMP >= 0) ? pt.amp * shape : shape;
    if (T1MODE) {
#pragma unroll
      for (int j = 0; j < NBASE; j++) grad[j] *= pt.t1fac;
      grad[NP - 1] = base * pt.t1dfac;
      return base * pt.t1fac;
    }
    return base;
  }
};

}  // namespace pnb
