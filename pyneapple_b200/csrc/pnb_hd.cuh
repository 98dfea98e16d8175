// Host/device portability shims.  The product library is built by nvcc only;
// the same headers also compile with g++ (PNB_HOST_SIM) so that the per-voxel
// mathematics can be unit-tested in the GPU-less authoring container
// (tests/hostsim, test infrastructure only — never loaded by pyneapple_b200).
#pragma once
#include <cfloat>
#include <cmath>
#include <cstdint>

#if defined(__CUDACC__)
#define PNB_HD __host__ __device__ __forceinline__
#define PNB_D __device__ __forceinline__
#else
#define PNB_HD inline
#define PNB_D inline
#endif

namespace pnb {
constexpr double kEps = DBL_EPSILON;
constexpr double kInf = __builtin_huge_val();

PNB_HD double dmax(double a, double b) { return a > b ? a : b; }  // np.maximum without NaN care
PNB_HD double dmin(double a, double b) { return a < b ? a : b; }
PNB_HD bool finite_d(double v) { return fabs(v) <= DBL_MAX; }

// ---- exp() for the signal models -------------------------------------------------------------
// The TRF kernel is bound by dependent FP64 chains (ncu: 51 % of the warp samples are "wait" at two
// warps per scheduler) and 23 % of its instructions are libdevice exp(), a 13-deep Horner chain.
// pnb_exp splits x = (64 e + j) ln2 / 64 + r, |r| <= ln2 / 128, takes 2^(j/64) from a 64-entry
// table (correctly rounded) and exp(r) - 1 from a degree-5 polynomial in Estrin form: 9 FP64
// operations, depth 7.  Error <= 1 ulp on [-690, 690] (checked against numpy on 2e6 arguments);
// everything else (overflow, underflow, NaN) goes to libm.
#define PNB_EXP_TABLE {                                                                          \
    0x1.0000000000000p+0, 0x1.02c9a3e778061p+0, 0x1.059b0d3158574p+0, 0x1.0874518759bc8p+0,  \
    0x1.0b5586cf9890fp+0, 0x1.0e3ec32d3d1a2p+0, 0x1.11301d0125b51p+0, 0x1.1429aaea92de0p+0,  \
    0x1.172b83c7d517bp+0, 0x1.1a35beb6fcb75p+0, 0x1.1d4873168b9aap+0, 0x1.2063b88628cd6p+0,  \
    0x1.2387a6e756238p+0, 0x1.26b4565e27cddp+0, 0x1.29e9df51fdee1p+0, 0x1.2d285a6e4030bp+0,  \
    0x1.306fe0a31b715p+0, 0x1.33c08b26416ffp+0, 0x1.371a7373aa9cbp+0, 0x1.3a7db34e59ff7p+0,  \
    0x1.3dea64c123422p+0, 0x1.4160a21f72e2ap+0, 0x1.44e086061892dp+0, 0x1.486a2b5c13cd0p+0,  \
    0x1.4bfdad5362a27p+0, 0x1.4f9b2769d2ca7p+0, 0x1.5342b569d4f82p+0, 0x1.56f4736b527dap+0,  \
    0x1.5ab07dd485429p+0, 0x1.5e76f15ad2148p+0, 0x1.6247eb03a5585p+0, 0x1.6623882552225p+0,  \
    0x1.6a09e667f3bcdp+0, 0x1.6dfb23c651a2fp+0, 0x1.71f75e8ec5f74p+0, 0x1.75feb564267c9p+0,  \
    0x1.7a11473eb0187p+0, 0x1.7e2f336cf4e62p+0, 0x1.82589994cce13p+0, 0x1.868d99b4492edp+0,  \
    0x1.8ace5422aa0dbp+0, 0x1.8f1ae99157736p+0, 0x1.93737b0cdc5e5p+0, 0x1.97d829fde4e50p+0,  \
    0x1.9c49182a3f090p+0, 0x1.a0c667b5de565p+0, 0x1.a5503b23e255dp+0, 0x1.a9e6b5579fdbfp+0,  \
    0x1.ae89f995ad3adp+0, 0x1.b33a2b84f15fbp+0, 0x1.b7f76f2fb5e47p+0, 0x1.bcc1e904bc1d2p+0,  \
    0x1.c199bdd85529cp+0, 0x1.c67f12e57d14bp+0, 0x1.cb720dcef9069p+0, 0x1.d072d4a07897cp+0,  \
    0x1.d5818dcfba487p+0, 0x1.da9e603db3285p+0, 0x1.dfc97337b9b5fp+0, 0x1.e502ee78b3ff6p+0,  \
    0x1.ea4afa2a490dap+0, 0x1.efa1bee615a27p+0, 0x1.f50765b6e4540p+0, 0x1.fa7c1819e90d8p+0,  \
}
#if defined(__CUDACC__)
static __device__ const double kExpTabGlobal[64] = PNB_EXP_TABLE;
// the kernel's copy in shared memory (one per CTA; fill it with exp_tab_init before the first use)
__device__ __forceinline__ double *exp_tab_shared() {
  __shared__ double tab[64];
  return tab;
}
__device__ __forceinline__ void exp_tab_init(int tid, int nthreads) {
  double *t = exp_tab_shared();
  for (int i = tid; i < 64; i += nthreads) t[i] = kExpTabGlobal[i];
  __syncthreads();
}
#endif
#if !defined(__CUDA_ARCH__)
static const double kExpTabHost[64] = PNB_EXP_TABLE;
#endif

// The out-of-range arguments must stay OUT of the instruction stream of the fast path: written
// inline (`if (...) return exp(x);`) the compiler if-converts the branch and every call executes
// libdevice's whole exp() next to the table version, selecting one result at the end (ncu, round 2:
// 30 % of the TRF kernel's executed instructions sat on that one source line).  A call the compiler
// cannot inline cannot be predicated either.
#if defined(__CUDA_ARCH__)
__device__ __noinline__ double pnb_exp_out_of_range(double x) { return exp(x); }
#else
inline double pnb_exp_out_of_range(double x) { return exp(x); }
#endif

// the table path alone: valid for |x| < 690, garbage (no trap) outside
PNB_HD double pnb_exp_core(double x) {
  const double kShift = 6755399441055744.0;  // 1.5 * 2^52: the integer lands in the low mantissa bits
  const double t = fma(x, 92.33248261689366, kShift);
  const double nf = t - kShift;
  double r = fma(nf, -0x1.62e42ff000000p-7, x);  // ln2 / 64, 32-bit head: nf * head is exact
  r = fma(nf, 0x1.718432a1b0e26p-41, r);
  const double r2 = r * r;
  const double a = fma(r, 1.0 / 6.0, 0.5), b = fma(r, 1.0 / 120.0, 1.0 / 24.0);
  const double q = fma(r2, b, a);
  const double p = fma(r2, q, r);  // exp(r) - 1
#if defined(__CUDA_ARCH__)
  const int n = __double2loint(t);
  const double T = exp_tab_shared()[n & 63];
  const double v = fma(T, p, T);
  return __hiloint2double(__double2hiint(v) + (n >> 6) * (1 << 20), __double2loint(v));
#else
  long long bits;
  __builtin_memcpy(&bits, &t, 8);
  const int n = (int)(unsigned)(bits & 0xffffffffLL);
  const double T = kExpTabHost[n & 63];
  const double v = fma(T, p, T);
  __builtin_memcpy(&bits, &v, 8);
  bits += (long long)(n >> 6) * (1LL << 52);
  double out;
  __builtin_memcpy(&out, &bits, 8);
  return out;
#endif
}

PNB_HD double pnb_exp(double x) {
  if (__builtin_expect(!(fabs(x) < 690.0), 0)) return pnb_exp_out_of_range(x);
  return pnb_exp_core(x);
}

// pnb_exp as a real call, for cold paths that must not be inlined into (and if-converted with) a hot one
#if defined(__CUDA_ARCH__)
__device__ __noinline__ double pnb_exp_cold(double x) { return pnb_exp(x); }
#else
inline double pnb_exp_cold(double x) { return pnb_exp(x); }
#endif
}  // namespace pnb
