// h = B^T y for all voxels as ONE dense FP64 GEMM on the tensor cores:
//
//   H0 (n_vox x n_bins) = Y (n_vox x n_b) * B (n_b x n_bins)
//
// — the first dual of Lawson-Hanson (x = 0: w = A^T b = B^T y, solvers/nnls_solver.py:61-86,
// 195-197), the one place on this path that is a contraction (BASELINE.json north_star (2)).
// mma.sync.aligned.m8n8k4.f64: one warp tile = 8 voxels x 8 bins, K = 4 b-values per instruction.
// The dictionary sits in shared memory (row stride chosen so the B-fragment loads of a warp are
// conflict-free), the 8 x n_b signal fragment of the tile in registers; each quad of lanes writes 64
// contiguous bytes of a voxel's row.  The kernel is HBM-write bound (8 n_bins bytes out, 8 n_b in
// per voxel); bench.py measures it against the fused form, where the same products are the first
// trip of nnls_v3_kernel's dual pass and nothing is written.
#pragma once
#include <cuda_runtime.h>

namespace pnb {

__device__ __forceinline__ void dmma_m8n8k4(double &c0, double &c1, double a, double b) {
  asm volatile("mma.sync.aligned.m8n8k4.row.col.f64.f64.f64.f64 {%0, %1}, {%2}, {%3}, {%0, %1};\n"
               : "+d"(c0), "+d"(c1)
               : "d"(a), "d"(b));
}

__host__ __device__ inline int nnls_gemm_ldb(int n) { return ((n + 7) & ~7) + 8; }  // 2 * ldb = 16 (mod 32) banks

template <int MT>
__global__ void __launch_bounds__(256) nnls_h0_dmma_kernel(const double *__restrict__ Y, const double *__restrict__ B,
                                                           double *__restrict__ H0, long long n_vox, int m, int n) {
  extern __shared__ double bsm[];  // [MT][ldb], zero padded
  const int ldb = nnls_gemm_ldb(n);
  for (int i = threadIdx.x; i < MT * ldb; i += blockDim.x) {
    const int row = i / ldb, col = i - row * ldb;
    bsm[i] = (row < m && col < n) ? B[(size_t)row * n + col] : 0.0;
  }
  __syncthreads();
  const int lane = threadIdx.x & 31, g = lane >> 2, t = lane & 3;
  const long long warps = (long long)gridDim.x * (blockDim.x >> 5);
  const long long w0 = (long long)blockIdx.x * (blockDim.x >> 5) + (threadIdx.x >> 5);
  const int n_tiles = (n + 7) >> 3;
  const bool vec_ok = (n & 1) == 0;
  for (long long tile = w0; tile * 8 < n_vox; tile += warps) {
    const long long v = tile * 8 + g;
    double a[MT / 4];
#pragma unroll
    for (int s = 0; s < MT / 4; s++) a[s] = (v < n_vox && 4 * s + t < m) ? Y[v * m + 4 * s + t] : 0.0;
    for (int bt = 0; bt < n_tiles; bt++) {
      double c0 = 0.0, c1 = 0.0;
#pragma unroll
      for (int s = 0; s < MT / 4; s++) dmma_m8n8k4(c0, c1, a[s], bsm[(4 * s + t) * ldb + 8 * bt + g]);
      const int bin = 8 * bt + 2 * t;
      if (v < n_vox) {
        double *dst = H0 + v * n + bin;
        if (vec_ok && bin + 1 < n) {
          *reinterpret_cast<double2 *>(dst) = make_double2(c0, c1);
        } else {
          if (bin < n) dst[0] = c0;
          if (bin + 1 < n) dst[1] = c1;
        }
      }
    }
  }
}

template <int MT>
cudaError_t nnls_h0_dmma_launch(const double *Y, const double *B, double *H0, long long n_vox, int m, int n,
                                cudaStream_t stream) {
  const size_t smem = (size_t)MT * nnls_gemm_ldb(n) * sizeof(double);
  auto kern = nnls_h0_dmma_kernel<MT>;
  if (smem > 48 * 1024) {
    cudaError_t e = cudaFuncSetAttribute(kern, cudaFuncAttributeMaxDynamicSharedMemorySize, (int)smem);
    if (e != cudaSuccess) return e;
  }
  long long blocks = (n_vox + 63) / 64;  // 8 warps x 8 voxels per CTA pass
  if (blocks > 148LL * 4) blocks = 148LL * 4;
  if (blocks < 1) blocks = 1;
  kern<<<(unsigned)blocks, 256, smem, stream>>>(Y, B, H0, n_vox, m, n);
  return cudaGetLastError();
}

}  // namespace pnb
