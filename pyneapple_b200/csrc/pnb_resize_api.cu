// C ABI of the OpenCV-faithful in-plane resampler (see include/pyneapple_b200.h).
#include <cuda_runtime.h>

#include "../../include/pyneapple_b200.h"
#include "pnb_internal.h"
#include "pnb_resize_kernel.cuh"

namespace {
int check(const pnb_resize_problem *p) {
  if (!p) return pnbi::fail(PNB_E_BADARG, "null problem");
  if (p->dtype != 0 && p->dtype != 1) return pnbi::fail(PNB_E_BADARG, "dtype must be 0 (f64) or 1 (f32)");
  if (p->method != 0 && p->method != 1) return pnbi::fail(PNB_E_BADARG, "method must be 0 (linear) or 1 (cubic)");
  if (p->src_h < 1 || p->src_w < 1 || p->dst_h < 1 || p->dst_w < 1 || p->inner < 1)
    return pnbi::fail(PNB_E_BADARG, "sizes must be positive");
  if (!p->src || !p->dst) return pnbi::fail(PNB_E_BADARG, "null array pointer");
  return 0;
}
int launch(const pnb_resize_problem *p, const void *src, void *dst, cudaStream_t stream) {
  pnb::ResizeArgs a;
  a.src_h = p->src_h; a.src_w = p->src_w; a.dst_h = p->dst_h; a.dst_w = p->dst_w; a.inner = p->inner;
  // cv::resize: inv_scale = dsize / ssize; hal::resize: scale = 1. / inv_scale
  a.scale_y = 1.0 / ((double)p->dst_h / (double)p->src_h);
  a.scale_x = 1.0 / ((double)p->dst_w / (double)p->src_w);
  a.method = p->method;
  // INTER_LINEAR with an exact 2x decimation is replaced by INTER_AREA in cv::resize
  if (p->method == 0 && p->src_h == 2 * p->dst_h && p->src_w == 2 * p->dst_w) a.method = 2;
  a.src = src; a.dst = dst;
  const long long total = (long long)p->dst_h * p->dst_w * p->inner;
  long long blocks = (total + 255) / 256;
  if (blocks > 148LL * 32) blocks = 148LL * 32;
  if (p->dtype == 0) pnb::resize_kernel<double><<<(unsigned)blocks, 256, 0, stream>>>(a);
  else pnb::resize_kernel<float><<<(unsigned)blocks, 256, 0, stream>>>(a);
  PNBI_CUDA(cudaGetLastError());
  pnbi::count_launch();
  return 0;
}
}  // namespace

extern "C" int pnb_resize2d_device(const pnb_resize_problem *p, void *cuda_stream) {
  if (int rc = check(p)) return rc;
  return launch(p, p->src, p->dst, (cudaStream_t)cuda_stream);
}

extern "C" int pnb_resize2d_host(const pnb_resize_problem *p, int device) {
  if (int rc = check(p)) return rc;
  if (pnb_device_count() <= device || device < 0) return pnbi::fail(PNB_E_NODEVICE, "no such CUDA device");
  pnbi::DeviceScope dev_scope(device);
  PNBI_CUDA(dev_scope.error());
  const size_t esz = p->dtype == 0 ? 8 : 4;
  const size_t nsrc = (size_t)p->src_h * p->src_w * p->inner * esz;
  const size_t ndst = (size_t)p->dst_h * p->dst_w * p->inner * esz;
  void *ds = nullptr, *dd = nullptr;
  PNBI_CUDA(cudaMalloc(&ds, nsrc));
  if (cudaMalloc(&dd, ndst) != cudaSuccess) { cudaFree(ds); return pnbi::fail(PNB_E_BADARG, "out of device memory"); }
  int rc = 0;
  cudaError_t e = cudaMemcpy(ds, p->src, nsrc, cudaMemcpyHostToDevice);
  if (e == cudaSuccess) rc = launch(p, ds, dd, nullptr);
  if (e == cudaSuccess && rc == 0) e = cudaMemcpy(p->dst, dd, ndst, cudaMemcpyDeviceToHost);
  cudaFree(ds); cudaFree(dd);
  if (e != cudaSuccess) return pnbi::cuda_fail(e, "pnb_resize2d_host");
  return rc;
}
