// In-plane resampling of (H, W, inner) arrays, bit-faithful to the OpenCV code
// path the reference's IDEAL fitter uses: fitters/ideal.py:299-320 calls
// cv2.resize(array[..., slice, channel], (W', H'), interpolation) once per z-slice
// and channel, i.e. purely 2-D, here done for all slices / channels at once.
//
// What is reproduced (cv::resize generic path, modules/imgproc/src/resize.cpp):
//   scale = 1 / (dst / src) (double); source coordinate and fractional offset
//   in FLOAT: fx = (float)((d + 0.5) * scale - 0.5), s = floor(fx), t = fx - s;
//   cubic weights in float with A = -0.75 (interpolateCubic); taps s-1..s+2
//   clamped to the image; horizontal pass rounded to the working type, then
//   vertical pass; left-to-right accumulation, no fused multiply-add (all
//   products / sums below use the _rn intrinsics so nvcc cannot contract them).
//   Linear: horizontal taps collapse at the borders (t = 0), vertical taps are
//   clamped; an exact 2x down-scale is INTER_AREA's 2x2 mean, as in OpenCV.
// FP64 cubic (image and parameter maps, the IDEAL default) is bit-identical to
// cv2.resize; FP64 linear is bit-identical to OpenCV's own implementation (the
// opencv-python wheel routes large linear resizes to Intel IPP, whose arithmetic
// is not public); FP32 is used for the label mask only.
#pragma once
#include <cuda_runtime.h>

namespace pnb {

struct ResizeArgs {
  int src_h, src_w, dst_h, dst_w;
  long long inner;
  double scale_y, scale_x;  // for rows (h) and columns (w)
  int method;               // 0 linear, 1 cubic, 2 area 2x2
  const void *src;
  void *dst;
};

__device__ __forceinline__ void src_coord(int d, double scale, int &s, float &t) {
  const double f = __dadd_rn(__dmul_rn(__dadd_rn((double)d, 0.5), scale), -0.5);
  const float fx = (float)f;
  const float fl = floorf(fx);
  s = (int)fl;
  t = __fsub_rn(fx, fl);
}

__device__ __forceinline__ void cubic_weights(float x, float (&w)[4]) {
  const float A = -0.75f;
  const float x1 = __fadd_rn(x, 1.0f);
  // ((A*(x+1) - 5A)*(x+1) + 8A)*(x+1) - 4A
  float v = __fsub_rn(__fmul_rn(A, x1), 5.0f * A);
  v = __fadd_rn(__fmul_rn(v, x1), 8.0f * A);
  w[0] = __fsub_rn(__fmul_rn(v, x1), 4.0f * A);
  // ((A+2)*x - (A+3))*x*x + 1
  v = __fsub_rn(__fmul_rn(A + 2.0f, x), A + 3.0f);
  w[1] = __fadd_rn(__fmul_rn(__fmul_rn(v, x), x), 1.0f);
  const float u = __fsub_rn(1.0f, x);
  v = __fsub_rn(__fmul_rn(A + 2.0f, u), A + 3.0f);
  w[2] = __fadd_rn(__fmul_rn(__fmul_rn(v, u), u), 1.0f);
  w[3] = __fsub_rn(__fsub_rn(__fsub_rn(1.0f, w[0]), w[1]), w[2]);
}

template <class T> __device__ __forceinline__ T mul_rn(T a, float w);
template <> __device__ __forceinline__ double mul_rn<double>(double a, float w) { return __dmul_rn(a, (double)w); }
template <> __device__ __forceinline__ float mul_rn<float>(float a, float w) { return __fmul_rn(a, w); }
template <class T> __device__ __forceinline__ T add_rn(T a, T b);
template <> __device__ __forceinline__ double add_rn<double>(double a, double b) { return __dadd_rn(a, b); }
template <> __device__ __forceinline__ float add_rn<float>(float a, float b) { return __fadd_rn(a, b); }

__device__ __forceinline__ int clampi(int v, int lo, int hi) { return v < lo ? lo : (v > hi ? hi : v); }

template <class T> __global__ void __launch_bounds__(256) resize_kernel(const ResizeArgs a) {
  const long long total = (long long)a.dst_h * a.dst_w * a.inner;
  const T *src = static_cast<const T *>(a.src);
  T *dst = static_cast<T *>(a.dst);
  const long long row_stride = (long long)a.src_w * a.inner;
  for (long long idx = blockIdx.x * (long long)blockDim.x + threadIdx.x; idx < total;
       idx += (long long)gridDim.x * blockDim.x) {
    const long long q = idx % a.inner;
    const long long pix = idx / a.inner;
    const int dx = (int)(pix % a.dst_w);  // column
    const int dy = (int)(pix / a.dst_w);  // row
    T out;
    if (a.method == 2) {
      const T *p = src + (long long)(2 * dy) * row_stride + (long long)(2 * dx) * a.inner + q;
      T sum = add_rn<T>(add_rn<T>(add_rn<T>(p[0], p[a.inner]), p[row_stride]), p[row_stride + a.inner]);
      out = sum * (T)0.25;
    } else if (a.method == 1) {
      int sx, sy;
      float tx, ty, wx[4], wy[4];
      src_coord(dx, a.scale_x, sx, tx);
      src_coord(dy, a.scale_y, sy, ty);
      cubic_weights(tx, wx);
      cubic_weights(ty, wy);
      T acc = 0;
#pragma unroll
      for (int r = 0; r < 4; r++) {
        const T *prow = src + (long long)clampi(sy - 1 + r, 0, a.src_h - 1) * row_stride + q;
        T h = 0;
#pragma unroll
        for (int c = 0; c < 4; c++) {
          const T v = mul_rn<T>(prow[(long long)clampi(sx - 1 + c, 0, a.src_w - 1) * a.inner], wx[c]);
          h = (c == 0) ? v : add_rn<T>(h, v);
        }
        const T v = mul_rn<T>(h, wy[r]);
        acc = (r == 0) ? v : add_rn<T>(acc, v);
      }
      out = acc;
    } else {
      int sx, sy;
      float tx, ty;
      src_coord(dx, a.scale_x, sx, tx);
      src_coord(dy, a.scale_y, sy, ty);
      bool single = false;  // horizontal: one tap with weight exactly 1 beyond the right border
      if (sx < 0) { sx = 0; tx = 0.0f; }
      if (sx >= a.src_w - 1) { sx = a.src_w - 1; tx = 0.0f; single = true; }
      const int sx1 = sx + 1 < a.src_w ? sx + 1 : a.src_w - 1;
      const float ax0 = __fsub_rn(1.0f, tx), ay0 = __fsub_rn(1.0f, ty);
      const int r0 = clampi(sy, 0, a.src_h - 1), r1 = clampi(sy + 1, 0, a.src_h - 1);
      const T *p0 = src + (long long)r0 * row_stride + q;
      const T *p1 = src + (long long)r1 * row_stride + q;
      T h0, h1;
      if (single) {
        h0 = p0[(long long)sx * a.inner];
        h1 = p1[(long long)sx * a.inner];
      } else {
        h0 = add_rn<T>(mul_rn<T>(p0[(long long)sx * a.inner], ax0), mul_rn<T>(p0[(long long)sx1 * a.inner], tx));
        h1 = add_rn<T>(mul_rn<T>(p1[(long long)sx * a.inner], ax0), mul_rn<T>(p1[(long long)sx1 * a.inner], tx));
      }
      out = add_rn<T>(mul_rn<T>(h0, ay0), mul_rn<T>(h1, ty));
    }
    dst[idx] = out;
  }
}

}  // namespace pnb
