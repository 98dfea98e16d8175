// C ABI of the result-side helpers (see include/pyneapple_b200.h): model prediction for every fitted
// voxel, and the row gather / scatter that moves voxels between a masked list and the volume.
//
//   predict  replaces the per-voxel model.forward loop of BaseFitter.predict (fitters/base.py:93-128)
//   gather   replaces image[segmentation != 0] of BaseFitter._extract_pixel_data (fitters/base.py:280-308)
//   scatter  replaces vol[idx] = values of BaseFitter._reconstruct_volume (fitters/base.py:310-330) and of
//            reconstruct_maps (io/nifti.py:279-312, float32 output)
//
// All three are HBM-bound: 8 n_params bytes in, 8 n_b bytes out per voxel for predict.
#include <cuda_runtime.h>

#include "../../include/pyneapple_b200.h"
#include "pnb_internal.h"
#include "pnb_models.cuh"

namespace {

struct PredictArgs {
  int n_b;
  long long n_vox;
  const double *b, *params;   // params (NP, n_vox)
  const long long *index;     // (n_vox) output row of each voxel, or nullptr: identity
  double *signal;             // (n_out, n_b)
  double tr, tm;
};

// One thread per voxel computes its n_b signal values into a shared-memory tile (row stride padded:
// conflict-free); the tile is then written with n_b consecutive threads per row, so every row —
// wherever the scatter index sends it — leaves as one contiguous segment.
template <class M, int TV>
__global__ void __launch_bounds__(TV) predict_kernel(const PredictArgs a) {
  extern __shared__ double tile[];
  const int nb = a.n_b, ld = nb | 1;
  double *bs = tile + (size_t)TV * ld;
  pnb::exp_tab_init(threadIdx.x, TV);  // pnb_exp's table lives in shared memory
  for (int i = threadIdx.x; i < nb; i += TV) bs[i] = a.b[i];
  __syncthreads();
  for (long long base = (long long)blockIdx.x * TV; base < a.n_vox; base += (long long)gridDim.x * TV) {
    const long long v = base + threadIdx.x;
    if (v < a.n_vox) {
      double p[M::NP];
#pragma unroll
      for (int j = 0; j < M::NP; j++) p[j] = a.params[(size_t)j * a.n_vox + v];
      typename M::Point pt;
      M::prepare(p, a.tr, a.tm, pt);
      double *row = tile + (size_t)threadIdx.x * ld;
      for (int r = 0; r < nb; r++) row[r] = M::value(pt, bs[r]);
    }
    __syncthreads();
    const long long left = a.n_vox - base;
    const int rows = left < TV ? (int)left : TV;
    for (int i = threadIdx.x; i < rows * nb; i += TV) {
      const int rr = i / nb, c = i - rr * nb;
      const long long dst = a.index ? a.index[base + rr] : base + rr;
      a.signal[(size_t)dst * nb + c] = tile[(size_t)rr * ld + c];
    }
    __syncthreads();
  }
}

using PredictFn = cudaError_t (*)(const PredictArgs &, cudaStream_t);

template <class M> cudaError_t predict_launch(const PredictArgs &a, cudaStream_t stream) {
  constexpr int TV = 128;
  const size_t smem = ((size_t)TV * (a.n_b | 1) + a.n_b) * sizeof(double);
  auto kern = predict_kernel<M, TV>;
  if (smem > 48 * 1024) {
    cudaError_t e = cudaFuncSetAttribute(kern, cudaFuncAttributeMaxDynamicSharedMemorySize, (int)smem);
    if (e != cudaSuccess) return e;
  }
  long long blocks = (a.n_vox + TV - 1) / TV;
  if (blocks > 148LL * 8) blocks = 148LL * 8;
  kern<<<(unsigned)blocks, TV, smem, stream>>>(a);
  return cudaGetLastError();
}

PredictFn predict_for(int model_id, int t1) {
#define PNB_PRED_ROW(t) {predict_launch<pnb::Model<0, t>>, predict_launch<pnb::Model<1, t>>, predict_launch<pnb::Model<2, t>>, \
                         predict_launch<pnb::Model<3, t>>, predict_launch<pnb::Model<4, t>>, predict_launch<pnb::Model<5, t>>, \
                         predict_launch<pnb::Model<6, t>>}
  static const PredictFn table[3][7] = {PNB_PRED_ROW(0), PNB_PRED_ROW(1), PNB_PRED_ROW(2)};
  if (model_id < 0 || model_id > 6 || t1 < 0 || t1 > 2) return nullptr;
  return table[t1][model_id];
}

// rows of `width` elements: dst[i] = src[index[i]] (gather) or dst[index[i]] = src[i] (scatter),
// width consecutive threads per row; TOUT = float converts on the way (reconstruct_maps is float32)
template <class TOUT, bool SCATTER>
__global__ void __launch_bounds__(256) move_rows_kernel(const double *src, TOUT *dst, const long long *index,
                                                        long long n_rows, int width) {
  const long long total = n_rows * width;
  for (long long i = (long long)blockIdx.x * blockDim.x + threadIdx.x; i < total; i += (long long)gridDim.x * blockDim.x) {
    const long long r = i / width;
    const int c = (int)(i - r * width);
    const long long other = index[r];
    if (SCATTER) dst[(size_t)other * width + c] = (TOUT)src[i];
    else dst[i] = (TOUT)src[(size_t)other * width + c];
  }
}

int model_np(int model_id, int t1) {
  static const int base[7] = {2, 3, 4, 4, 5, 6, 6};
  return base[model_id] + (t1 ? 1 : 0);
}

int check_predict(const pnb_predict_problem *p) {
  if (!p) return pnbi::fail(PNB_E_BADARG, "null problem");
  if (!predict_for(p->model_id, p->t1_mode)) return pnbi::fail(PNB_E_UNSUPPORTED, "unknown model_id / t1_mode");
  if (p->n_params != model_np(p->model_id, p->t1_mode)) return pnbi::fail(PNB_E_BADARG, "n_params does not match the model");
  if (p->n_b < 1 || p->n_b > 512) return pnbi::fail(PNB_E_BADARG, "n_b must be in [1, 512]");
  if (p->n_vox < 0 || p->n_out < 0) return pnbi::fail(PNB_E_BADARG, "negative size");
  if (!p->flat_index && p->n_out != p->n_vox) return pnbi::fail(PNB_E_BADARG, "n_out must equal n_vox without a scatter index");
  if (p->n_vox > 0 && (!p->xdata || !p->params || !p->signal)) return pnbi::fail(PNB_E_BADARG, "null array pointer");
  return 0;
}

int check_rows(const pnb_rows_problem *p) {
  if (!p) return pnbi::fail(PNB_E_BADARG, "null problem");
  if (p->n_rows < 0 || p->n_other < 0 || p->width < 1) return pnbi::fail(PNB_E_BADARG, "bad sizes");
  if (p->direction != 0 && p->direction != 1) return pnbi::fail(PNB_E_BADARG, "direction must be 0 (gather) or 1 (scatter)");
  if (p->out_dtype != 0 && p->out_dtype != 1) return pnbi::fail(PNB_E_BADARG, "out_dtype must be 0 (f64) or 1 (f32)");
  if (p->n_rows > 0 && (!p->src || !p->dst || !p->index)) return pnbi::fail(PNB_E_BADARG, "null array pointer");
  return 0;
}

}  // namespace

extern "C" int pnb_predict_device(const pnb_predict_problem *p, void *cuda_stream) {
  if (int rc = check_predict(p)) return rc;
  cudaStream_t stream = (cudaStream_t)cuda_stream;
  if (p->flat_index && p->n_out > 0)  // voxels nobody fitted read as zero, like np.zeros + assignment
    PNBI_CUDA(cudaMemsetAsync(p->signal, 0, (size_t)p->n_out * p->n_b * sizeof(double), stream));
  if (p->n_vox == 0) return 0;
  PredictArgs a;
  a.n_b = p->n_b; a.n_vox = p->n_vox; a.b = p->xdata; a.params = p->params;
  a.index = reinterpret_cast<const long long *>(p->flat_index); a.signal = p->signal;
  a.tr = p->repetition_time; a.tm = p->mixing_time;
  cudaError_t e = predict_for(p->model_id, p->t1_mode)(a, stream);
  if (e != cudaSuccess) return pnbi::cuda_fail(e, "predict kernel launch");
  pnbi::count_launch();
  return 0;
}

extern "C" int pnb_move_rows_device(const pnb_rows_problem *p, void *cuda_stream) {
  if (int rc = check_rows(p)) return rc;
  cudaStream_t stream = (cudaStream_t)cuda_stream;
  const size_t esz = p->out_dtype == 0 ? 8 : 4;
  if (p->direction == 1 && p->zero_fill && p->n_other > 0)
    PNBI_CUDA(cudaMemsetAsync(p->dst, 0, (size_t)p->n_other * p->width * esz, stream));
  if (p->n_rows == 0) return 0;
  const long long total = p->n_rows * (long long)p->width;
  long long blocks = (total + 255) / 256;
  if (blocks > 148LL * 16) blocks = 148LL * 16;
  const long long *idx = reinterpret_cast<const long long *>(p->index);
  const unsigned g = (unsigned)blocks;
  if (p->direction == 0) {
    if (p->out_dtype == 0) move_rows_kernel<double, false><<<g, 256, 0, stream>>>(p->src, (double *)p->dst, idx, p->n_rows, p->width);
    else move_rows_kernel<float, false><<<g, 256, 0, stream>>>(p->src, (float *)p->dst, idx, p->n_rows, p->width);
  } else {
    if (p->out_dtype == 0) move_rows_kernel<double, true><<<g, 256, 0, stream>>>(p->src, (double *)p->dst, idx, p->n_rows, p->width);
    else move_rows_kernel<float, true><<<g, 256, 0, stream>>>(p->src, (float *)p->dst, idx, p->n_rows, p->width);
  }
  PNBI_CUDA(cudaGetLastError());
  pnbi::count_launch();
  return 0;
}

extern "C" int pnb_sizeof_predict_problem(void) { return (int)sizeof(pnb_predict_problem); }
extern "C" int pnb_sizeof_rows_problem(void) { return (int)sizeof(pnb_rows_problem); }
