// C ABI of the batched NNLS-spectrum post-processing (see include/pyneapple_b200.h).
#include <cuda_runtime.h>

#include "../../include/pyneapple_b200.h"
#include "pnb_internal.h"
#include "pnb_spectrum_kernel.cuh"

namespace {
int check(const pnb_spectrum_problem *p) {
  if (!p) return pnbi::fail(PNB_E_BADARG, "null problem");
  if (p->n_vox < 0) return pnbi::fail(PNB_E_BADARG, "n_vox < 0");
  if (p->max_peaks < 1 || p->max_peaks > 32) return pnbi::fail(PNB_E_BADARG, "max_peaks must be in [1, 32]");
  if (p->n_cutoffs < 0 || p->n_cutoffs > 32) return pnbi::fail(PNB_E_BADARG, "n_cutoffs must be in [0, 32]");
  if (p->n_vox == 0) return 0;
  if (p->detect || p->areas) {
    if (p->n_bins < 3 || p->n_bins > 4096) return pnbi::fail(PNB_E_BADARG, "n_bins must be in [3, 4096]");
    if (!p->spectrum) return pnbi::fail(PNB_E_BADARG, "spectrum is needed to detect peaks or to measure their areas");
  }
  if (p->detect && !p->bins) return pnbi::fail(PNB_E_BADARG, "bins is needed to detect peaks");
  if (!p->n_peaks) return pnbi::fail(PNB_E_BADARG, "null n_peaks");
  if (!p->detect && p->areas && !p->peak_index) return pnbi::fail(PNB_E_BADARG, "areas of given peaks need peak_index");
  if (!p->detect && !p->f_values) return pnbi::fail(PNB_E_BADARG, "given peaks need f_values");
  if (!p->detect && !p->peak_index && !p->d_values) return pnbi::fail(PNB_E_BADARG, "given peaks need d_values or peak_index");
  if (p->n_cutoffs > 0 && (!p->cutoffs || !p->d_cut || !p->f_cut)) return pnbi::fail(PNB_E_BADARG, "null cut-off arrays");
  return 0;
}

pnb::SpectrumArgs args_of(const pnb_spectrum_problem *p) {
  pnb::SpectrumArgs a;
  a.n = p->n_bins; a.max_peaks = p->max_peaks; a.detect = p->detect; a.areas = p->areas;
  a.normalize = p->normalize; a.n_cut = p->n_cutoffs; a.cut_normalize = p->cut_normalize;
  a.height = p->height; a.rel_height = p->rel_height; a.n_vox = p->n_vox;
  a.bins = p->bins; a.cutoffs = p->cutoffs; a.x = p->spectrum; a.n_peaks = p->n_peaks; a.idx = p->peak_index;
  a.d = p->d_values; a.f = p->f_values; a.d_cut = p->d_cut; a.f_cut = p->f_cut;
  return a;
}

int launch(const pnb::SpectrumArgs &a, cudaStream_t stream) {
  const size_t smem = pnb::spectrum_smem_bytes(a.n);  // the kernel lays its per-warp arrays out behind n doubles
  if (smem > 200 * 1024) return pnbi::fail(PNB_E_UNSUPPORTED, "n_bins too large for shared memory");
  if (smem > 48 * 1024)
    PNBI_CUDA(cudaFuncSetAttribute(pnb::spectrum_kernel, cudaFuncAttributeMaxDynamicSharedMemorySize, (int)smem));
  long long blocks = (a.n_vox + pnb::kSpecWarps - 1) / pnb::kSpecWarps;
  if (blocks > 148LL * 8) blocks = 148LL * 8;
  pnb::spectrum_kernel<<<(unsigned)blocks, pnb::kSpecWarps * 32, smem, stream>>>(a);
  PNBI_CUDA(cudaGetLastError());
  pnbi::count_launch();
  return 0;
}

template <class T> struct DevBuf {
  T *p = nullptr;
  ~DevBuf() { if (p) cudaFree(p); }
  cudaError_t alloc(size_t n) { return cudaMalloc(&p, (n ? n : 1) * sizeof(T)); }
};
}  // namespace

extern "C" int pnb_spectrum_peaks_device(const pnb_spectrum_problem *p, void *cuda_stream) {
  if (int rc = check(p)) return rc;
  if (p->n_vox == 0) return 0;
  return launch(args_of(p), (cudaStream_t)cuda_stream);
}

// Host arrays: voxels are processed in chunks (upload, kernel, download on one stream; the spectra
// dominate the traffic, 8 n_bins bytes per voxel, so the chunk loop runs at PCIe speed).
extern "C" int pnb_spectrum_peaks_host(const pnb_spectrum_problem *p, int device, int64_t chunk_vox) {
  if (int rc = check(p)) return rc;
  if (p->n_vox == 0) return 0;
  if (pnb_device_count() <= device || device < 0) return pnbi::fail(PNB_E_NODEVICE, "no such CUDA device");
  pnbi::DeviceScope dev_scope(device);
  PNBI_CUDA(dev_scope.error());
  if (chunk_vox <= 0) chunk_vox = 1 << 18;
  if (chunk_vox > p->n_vox) chunk_vox = p->n_vox;
  const size_t C = (size_t)chunk_vox, n = (size_t)p->n_bins, P = (size_t)p->max_peaks, K = (size_t)p->n_cutoffs;
  const bool need_x = p->spectrum != nullptr;
  DevBuf<double> x, bins, cut, d, f, dc, fc;
  DevBuf<int> np_, idx;
  if (need_x) PNBI_CUDA(x.alloc(C * n));
  if (p->bins) {
    PNBI_CUDA(bins.alloc(n));
    PNBI_CUDA(cudaMemcpy(bins.p, p->bins, n * sizeof(double), cudaMemcpyHostToDevice));
  }
  if (K) {
    PNBI_CUDA(cut.alloc(2 * K));
    PNBI_CUDA(cudaMemcpy(cut.p, p->cutoffs, 2 * K * sizeof(double), cudaMemcpyHostToDevice));
    PNBI_CUDA(dc.alloc(C * K));
    PNBI_CUDA(fc.alloc(C * K));
  }
  PNBI_CUDA(np_.alloc(C));
  PNBI_CUDA(idx.alloc(C * P));
  PNBI_CUDA(d.alloc(C * P));
  PNBI_CUDA(f.alloc(C * P));
  cudaStream_t st = nullptr;
  for (size_t s0 = 0; s0 < (size_t)p->n_vox; s0 += C) {
    const size_t nv = ((size_t)p->n_vox - s0 < C) ? (size_t)p->n_vox - s0 : C;
    if (need_x) PNBI_CUDA(cudaMemcpyAsync(x.p, p->spectrum + s0 * n, nv * n * sizeof(double), cudaMemcpyHostToDevice, st));
    if (!p->detect) {
      PNBI_CUDA(cudaMemcpyAsync(np_.p, p->n_peaks + s0, nv * sizeof(int), cudaMemcpyHostToDevice, st));
      if (p->peak_index) PNBI_CUDA(cudaMemcpyAsync(idx.p, p->peak_index + s0 * P, nv * P * sizeof(int), cudaMemcpyHostToDevice, st));
      if (p->d_values) PNBI_CUDA(cudaMemcpyAsync(d.p, p->d_values + s0 * P, nv * P * sizeof(double), cudaMemcpyHostToDevice, st));
      PNBI_CUDA(cudaMemcpyAsync(f.p, p->f_values + s0 * P, nv * P * sizeof(double), cudaMemcpyHostToDevice, st));
    }
    pnb::SpectrumArgs a = args_of(p);
    a.n_vox = (long long)nv;
    a.bins = p->bins ? bins.p : nullptr; a.cutoffs = cut.p; a.x = need_x ? x.p : nullptr;
    a.n_peaks = np_.p; a.idx = (p->detect || p->peak_index) ? idx.p : nullptr;
    a.d = (p->detect || p->d_values) ? d.p : nullptr; a.f = f.p; a.d_cut = dc.p; a.f_cut = fc.p;
    if (int rc = launch(a, st)) return rc;
    if (p->detect) PNBI_CUDA(cudaMemcpyAsync(p->n_peaks + s0, np_.p, nv * sizeof(int), cudaMemcpyDeviceToHost, st));
    if (p->detect && p->peak_index) PNBI_CUDA(cudaMemcpyAsync(p->peak_index + s0 * P, idx.p, nv * P * sizeof(int), cudaMemcpyDeviceToHost, st));
    if (p->d_values && a.d) PNBI_CUDA(cudaMemcpyAsync(p->d_values + s0 * P, d.p, nv * P * sizeof(double), cudaMemcpyDeviceToHost, st));
    if (p->f_values) PNBI_CUDA(cudaMemcpyAsync(p->f_values + s0 * P, f.p, nv * P * sizeof(double), cudaMemcpyDeviceToHost, st));
    if (K) {
      PNBI_CUDA(cudaMemcpyAsync(p->d_cut + s0 * K, dc.p, nv * K * sizeof(double), cudaMemcpyDeviceToHost, st));
      PNBI_CUDA(cudaMemcpyAsync(p->f_cut + s0 * K, fc.p, nv * K * sizeof(double), cudaMemcpyDeviceToHost, st));
    }
    PNBI_CUDA(cudaStreamSynchronize(st));
  }
  return 0;
}

extern "C" int pnb_sizeof_spectrum_problem(void) { return (int)sizeof(pnb_spectrum_problem); }
