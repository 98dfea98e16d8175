// Per-label mean signals of a volume (SegmentationWiseFitter._extract_segmentation_mean_signals,
// fitters/segmentationwise.py:112-137: np.mean(image[segmentation == seg], axis=0) for every label).
//
// Deterministic two-level reduction (no floating-point atomics, so the result does not depend
// on scheduling): every warp owns a contiguous slice of the voxels and adds them, in voxel order,
// into its own (n_labels, n_b) table in shared memory — lane = b-value, a run of equal labels is
// summed in registers and flushed once —, writes the table to global memory, and a second kernel
// adds the per-warp tables in warp order and divides by the counts.  The order of the additions
// differs from NumPy's (voxel order over the whole volume), i.e. results agree to a few ulp.
// HBM-bound: 8 n_b + 4 bytes per voxel in, nothing out.
#include <cuda_runtime.h>

#include "../../include/pyneapple_b200.h"
#include "pnb_internal.h"

namespace {

constexpr int kWarps = 8;

__global__ void __launch_bounds__(kWarps * 32) segmeans_partial_kernel(
    int n_b, int n_labels, long long n_vox, long long vox_per_warp, const double *__restrict__ image,
    const int *__restrict__ label, double *__restrict__ partial_sum, long long *__restrict__ partial_cnt) {
  extern __shared__ double sm[];
  const int lane = threadIdx.x & 31, wid = threadIdx.x >> 5;
  const long long gw = (long long)blockIdx.x * (blockDim.x >> 5) + wid;
  const int T = n_labels * n_b;
  double *tab = sm + (size_t)wid * (T + n_labels);
  long long *cnt = reinterpret_cast<long long *>(tab + T);
  for (int i = lane; i < T; i += 32) tab[i] = 0.0;
  for (int i = lane; i < n_labels; i += 32) cnt[i] = 0;
  __syncwarp();
  const long long v0 = gw * vox_per_warp;
  long long v1 = v0 + vox_per_warp;
  if (v1 > n_vox) v1 = n_vox;
  for (int b0 = 0; b0 < n_b; b0 += 32) {
    const int b = b0 + lane;
    const bool act = b < n_b;
    int cur = -1;
    double acc = 0.0;
    long long run = 0;
    for (long long v = v0; v < v1; v += 4) {
      // four voxels per trip: the loads are issued together, the sums stay in voxel order
      int l[4];
      double x[4];
#pragma unroll
      for (int k = 0; k < 4; k++) {
        const bool in = v + k < v1;
        l[k] = in ? label[v + k] : -1;
        x[k] = (in && act) ? image[(v + k) * n_b + b] : 0.0;
      }
#pragma unroll
      for (int k = 0; k < 4; k++) {
        if (v + k >= v1) break;
        if (l[k] != cur) {
          if (cur >= 0 && cur < n_labels) {
            if (act) tab[cur * n_b + b] += acc;
            if (b0 == 0 && lane == 0) cnt[cur] += run;
          }
          cur = l[k]; acc = 0.0; run = 0;
        }
        acc += x[k];
        run += 1;
      }
    }
    if (cur >= 0 && cur < n_labels) {
      if (act) tab[cur * n_b + b] += acc;
      if (b0 == 0 && lane == 0) cnt[cur] += run;
    }
    __syncwarp();
  }
  for (int i = lane; i < T; i += 32) partial_sum[gw * T + i] = tab[i];
  for (int i = lane; i < n_labels; i += 32) partial_cnt[gw * n_labels + i] = cnt[i];
}

__global__ void segmeans_final_kernel(int n_b, int n_labels, long long n_warps, const double *partial_sum,
                                      const long long *partial_cnt, double *means, long long *counts) {
  const int i = blockIdx.x * blockDim.x + threadIdx.x;
  const int T = n_labels * n_b;
  if (i >= T) return;
  const int lab = i / n_b;
  double s = 0.0;
  long long c = 0;
  for (long long w = 0; w < n_warps; w++) {
    s += partial_sum[w * T + i];
    c += partial_cnt[w * n_labels + lab];
  }
  means[i] = s / (double)c;  // 0 / 0 = NaN for an empty label, like np.mean of an empty selection
  if (i % n_b == 0 && counts) counts[lab] = c;
}

int check(const pnb_segmeans_problem *p) {
  if (!p) return pnbi::fail(PNB_E_BADARG, "null problem");
  if (p->n_b < 1 || p->n_b > 512) return pnbi::fail(PNB_E_BADARG, "n_b must be in [1, 512]");
  if (p->n_labels < 1) return pnbi::fail(PNB_E_BADARG, "n_labels must be positive");
  if (p->n_vox < 0) return pnbi::fail(PNB_E_BADARG, "n_vox < 0");
  if ((size_t)p->n_labels * (p->n_b + 1) * sizeof(double) > 220 * 1024)
    return pnbi::fail(PNB_E_UNSUPPORTED, "n_labels x n_b too large for shared memory");
  if (!p->image || !p->label || !p->means) return pnbi::fail(PNB_E_BADARG, "null array pointer");
  return 0;
}

struct Plan { int warps_per_cta; long long n_warps, vox_per_warp; unsigned grid; size_t smem; };

Plan plan_for(const pnb_segmeans_problem *p) {
  Plan pl;
  const size_t per_warp = (size_t)p->n_labels * (p->n_b + 1) * sizeof(double);
  int warps = kWarps;
  while (warps > 1 && warps * per_warp > 220 * 1024) warps--;  // many labels: fewer warps per CTA
  pl.smem = warps * per_warp;
  long long grid = 148LL * (pl.smem > 100 * 1024 ? 1 : 4);
  const long long want = (p->n_vox + warps * 256 - 1) / (warps * 256);
  if (want < grid) grid = want;
  if (grid < 1) grid = 1;
  pl.grid = (unsigned)grid;
  pl.n_warps = grid * warps;
  pl.vox_per_warp = (p->n_vox + pl.n_warps - 1) / pl.n_warps;
  pl.warps_per_cta = warps;
  return pl;
}

int launch(const pnb_segmeans_problem *p, const double *image, const int *label, double *means,
           long long *counts, cudaStream_t stream) {
  Plan pl = plan_for(p);
  if (pl.smem > 220 * 1024) return pnbi::fail(PNB_E_UNSUPPORTED, "n_labels x n_b too large for shared memory");
  if (pl.smem > 48 * 1024)
    PNBI_CUDA(cudaFuncSetAttribute(segmeans_partial_kernel, cudaFuncAttributeMaxDynamicSharedMemorySize, (int)pl.smem));
  const size_t T = (size_t)p->n_labels * p->n_b;
  double *ps = nullptr;
  long long *pc = nullptr;
  PNBI_CUDA(cudaMallocAsync(&ps, pl.n_warps * T * sizeof(double), stream));
  PNBI_CUDA(cudaMallocAsync(&pc, pl.n_warps * p->n_labels * sizeof(long long), stream));
  segmeans_partial_kernel<<<pl.grid, pl.warps_per_cta * 32, pl.smem, stream>>>(p->n_b, p->n_labels, p->n_vox, pl.vox_per_warp,
                                                                    image, label, ps, pc);
  PNBI_CUDA(cudaGetLastError());
  pnbi::count_launch();
  segmeans_final_kernel<<<(unsigned)((T + 127) / 128), 128, 0, stream>>>(p->n_b, p->n_labels, pl.n_warps, ps, pc,
                                                                         means, counts);
  PNBI_CUDA(cudaGetLastError());
  pnbi::count_launch();
  PNBI_CUDA(cudaFreeAsync(ps, stream));
  PNBI_CUDA(cudaFreeAsync(pc, stream));
  return 0;
}
}  // namespace

extern "C" int pnb_segment_means_device(const pnb_segmeans_problem *p, void *cuda_stream) {
  if (int rc = check(p)) return rc;
  return launch(p, p->image, p->label, p->means, reinterpret_cast<long long *>(p->counts), (cudaStream_t)cuda_stream);
}

extern "C" int pnb_segment_means_host(const pnb_segmeans_problem *p, int device) {
  if (int rc = check(p)) return rc;
  if (pnb_device_count() <= device || device < 0) return pnbi::fail(PNB_E_NODEVICE, "no such CUDA device");
  pnbi::DeviceScope dev_scope(device);
  PNBI_CUDA(dev_scope.error());
  const size_t nv = (size_t)p->n_vox, T = (size_t)p->n_labels * p->n_b;
  double *img = nullptr, *means = nullptr;
  int *lab = nullptr;
  long long *cnt = nullptr;
  cudaStream_t st = nullptr;
  int rc = 0;
  cudaError_t e = cudaMalloc(&img, (nv ? nv : 1) * p->n_b * sizeof(double));
  if (e == cudaSuccess) e = cudaMalloc(&lab, (nv ? nv : 1) * sizeof(int));
  if (e == cudaSuccess) e = cudaMalloc(&means, T * sizeof(double));
  if (e == cudaSuccess) e = cudaMalloc(&cnt, p->n_labels * sizeof(long long));
  if (e == cudaSuccess) e = cudaMemcpyAsync(img, p->image, nv * p->n_b * sizeof(double), cudaMemcpyHostToDevice, st);
  if (e == cudaSuccess) e = cudaMemcpyAsync(lab, p->label, nv * sizeof(int), cudaMemcpyHostToDevice, st);
  if (e == cudaSuccess) {
    pnb_segmeans_problem q = *p;
    rc = launch(&q, img, lab, means, cnt, st);
  }
  if (e == cudaSuccess && rc == 0) e = cudaMemcpyAsync(p->means, means, T * sizeof(double), cudaMemcpyDeviceToHost, st);
  if (e == cudaSuccess && rc == 0 && p->counts)
    e = cudaMemcpyAsync(p->counts, cnt, p->n_labels * sizeof(long long), cudaMemcpyDeviceToHost, st);
  if (e == cudaSuccess) e = cudaStreamSynchronize(st);
  cudaFree(img); cudaFree(lab); cudaFree(means); cudaFree(cnt);
  if (e != cudaSuccess) return pnbi::cuda_fail(e, "pnb_segment_means_host");
  return rc;
}
