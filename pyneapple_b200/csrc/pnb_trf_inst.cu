// One translation unit per (model, T1 mode, method): compiled with
//   -DPNB_MODEL_ID=<0..6> -DPNB_T1MODE=<0..2> [-DPNB_METHOD=1 for dogbox, 2 for lm]
//   [-DPNB_EXTRAS=1: the trf kernel that honours curve_fit's sigma / least_squares' robust loss]
// so the seven-plus register-heavy kernels build in parallel.
#include "pnb_trf_kernel.cuh"

#ifndef PNB_MODEL_ID
#error "compile with -DPNB_MODEL_ID=<id> -DPNB_T1MODE=<mode>"
#endif
#ifndef PNB_TRF_BLOCK
#define PNB_TRF_BLOCK 128
#endif

#ifndef PNB_METHOD
#define PNB_METHOD 0
#endif
#ifndef PNB_EXTRAS
#define PNB_EXTRAS 0
#endif
#if PNB_EXTRAS
#define PNB_CAT_(a, b, c) pnb_trfx_launch_##a##_##b
#elif PNB_METHOD == 2
#define PNB_CAT_(a, b, c) pnb_lm_launch_##a##_##b
#elif PNB_METHOD == 1
#define PNB_CAT_(a, b, c) pnb_dogbox_launch_##a##_##b
#else
#define PNB_CAT_(a, b, c) pnb_trf_launch_##a##_##b
#endif
#define PNB_CAT(a, b) PNB_CAT_(a, b, 0)

extern "C" cudaError_t PNB_CAT(PNB_MODEL_ID, PNB_T1MODE)(const pnb::TrfDeviceArgs *a,
                                                          cudaStream_t stream) {
  return pnb::trf_launch<pnb::Model<PNB_MODEL_ID, PNB_T1MODE>, PNB_TRF_BLOCK, PNB_METHOD, (PNB_EXTRAS != 0)>(*a, stream);
}
