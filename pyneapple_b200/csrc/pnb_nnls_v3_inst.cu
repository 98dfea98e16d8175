// One translation unit per (MT, WK) instantiation of the v3 NNLS kernel:
//   -DPNB_V3_MT=<8|16|24|32> -DPNB_V3_WK=<0|2|4>
#include "pnb_nnls_v3.cuh"

#if !defined(PNB_V3_MT) || !defined(PNB_V3_WK)
#error "compile with -DPNB_V3_MT=<mt> -DPNB_V3_WK=<wk>"
#endif

#define PNB_CAT_(a, b) pnb_nnls_v3_launch_##a##_##b
#define PNB_CAT(a, b) PNB_CAT_(a, b)

extern "C" cudaError_t PNB_CAT(PNB_V3_MT, PNB_V3_WK)(const pnb::NnlsDeviceArgs *a, cudaStream_t stream) {
  return pnb::nnls_v3_launch<PNB_V3_MT, PNB_V3_WK>(*a, stream);
}
