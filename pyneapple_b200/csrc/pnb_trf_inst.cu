// One translation unit per (model, T1 mode): compiled with
//   -DPNB_MODEL_ID=<0..6> -DPNB_T1MODE=<0..2>
// so the seven-plus register-heavy kernels build in parallel.
#include "pnb_trf_kernel.cuh"

#ifndef PNB_MODEL_ID
#error "compile with -DPNB_MODEL_ID=<id> -DPNB_T1MODE=<mode>"
#endif
#ifndef PNB_TRF_BLOCK
#define PNB_TRF_BLOCK 128
#endif

#define PNB_CAT_(a, b, c) pnb_trf_launch_##a##_##b
#define PNB_CAT(a, b) PNB_CAT_(a, b, 0)

extern "C" cudaError_t PNB_CAT(PNB_MODEL_ID, PNB_T1MODE)(const pnb::TrfDeviceArgs *a,
                                                          cudaStream_t stream) {
  return pnb::trf_launch<pnb::Model<PNB_MODEL_ID, PNB_T1MODE>, PNB_TRF_BLOCK>(*a, stream);
}
