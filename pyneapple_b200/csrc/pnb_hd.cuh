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
}  // namespace pnb
