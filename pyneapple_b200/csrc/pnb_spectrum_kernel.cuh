// Batched post-processing of NNLS spectra: peaks, Gaussian peak areas, cut-off ranges.
//
// Replaces the per-voxel Python loop a user of the reference writes around
// utility/spectrum.py:13-206 (find_spectrum_peaks -> calculate_peak_area ->
// apply_cutoffs -> geometric_mean_peak), whose arithmetic is SciPy's
// scipy.signal.find_peaks(x, height=h) and peak_widths(x, peaks, rel_height)
// (scipy/signal/_peak_finding.py, _peak_finding_utils.pyx: _local_maxima_1d,
// _peak_prominences, _peak_widths).  Restated here per voxel:
//
//   * a local maximum is a rising edge x[i-1] < x[i] followed by a plateau of equal
//     samples and a falling edge; the peak is the plateau's midpoint (integer
//     division).  Every rising edge can be judged on its own, so the 32 lanes test
//     32 positions at once and a ballot keeps the peaks in index order;
//   * height filter  x[peak] >= height;
//   * prominence: walk left / right from the peak while x[i] <= x[peak], remember the
//     lowest sample (first occurrence); prominence = x[peak] - max(left_min, right_min);
//   * width at rel_height: evaluation height = x[peak] - prominence * rel_height, walk
//     from the peak towards each base while the height is below the signal,
//     interpolate linearly between the two samples that bracket it;
//   * area = height * width / (2 sqrt(2 ln 2)) * sqrt(2 pi)   (spectrum.py:46),
//     fractions normalised to sum 1 (spectrum.py:97-99);
//   * cut-off ranges (spectrum.py:139-206): no peak -> NaN, one -> kept, several ->
//     (log10 of the weighted geometric mean, summed weight), then nansum-normalised.
//
// One warp per spectrum: the row is staged in shared memory with coalesced loads
// (the kernel is HBM-bound: 8 n_bins bytes per voxel in, a few dozen out), one lane
// per peak does the two walks, one lane per cut-off range does the merge.
#pragma once
#include <cuda_runtime.h>

#include "pnb_hd.cuh"

namespace pnb {

struct SpectrumArgs {
  int n;              // bins
  int max_peaks;      // P <= 32: peaks stored per voxel
  int detect;         // 1: find the peaks in x; 0: peak list given (n_peaks, idx and / or d, f)
  int areas;          // 1: f = Gaussian area of each peak (needs x and idx); 0: f = x[peak] (detect) or as given
  int normalize;      // 1: fractions divided by their sum
  int n_cut;          // cut-off ranges (<= 32), 0: stage skipped
  int cut_normalize;  // 1: nansum-normalise the merged fractions (apply_cutoffs); 0: raw (geometric_mean_peak)
  double height, rel_height;
  long long n_vox;
  const double *bins;     // (n) or nullptr when d is given
  const double *cutoffs;  // (n_cut, 2)
  const double *x;        // (n_vox, n) or nullptr
  int *n_peaks;           // (n_vox)           out (detect) / in
  int *idx;               // (n_vox, P)        out (detect) / in, or nullptr
  double *d;              // (n_vox, P)        out / in, NaN padded
  double *f;              // (n_vox, P)        out / in (areas = 1, detect = 0: heights in, areas out)
  double *d_cut, *f_cut;  // (n_vox, n_cut)
};

constexpr int kSpecWarps = 8;

__global__ void __launch_bounds__(kSpecWarps * 32) spectrum_kernel(const SpectrumArgs a) {
  extern __shared__ double smem_spec[];
  const int lane = threadIdx.x & 31, wid = threadIdx.x >> 5;
  const unsigned FULL = 0xffffffffu;
  const int n = a.n, P = a.max_peaks;
  const int nx = (n + 1) & ~1;
  double *xs = smem_spec + (size_t)wid * (nx + 3 * 32);
  double *ds = xs + nx, *fs = ds + 32, *cs = fs + 32;  // peak positions, fractions, merged fractions
  __shared__ int idx_s[kSpecWarps][32];
  const double qnan = nan("");
  // np.sqrt(2 * np.log(2)) * 2 and np.sqrt(2 * np.pi), the doubles NumPy produces
  const double kFwhm = 2.3548200450309493, kSqrt2Pi = 2.5066282746310002;

  for (long long vox = (long long)blockIdx.x * kSpecWarps + wid; vox < a.n_vox;
       vox += (long long)gridDim.x * kSpecWarps) {
    if (a.x) {
      const double *row = a.x + vox * (long long)n;
      for (int i = lane; i < n; i += 32) xs[i] = row[i];
    }
    __syncwarp();
    int count = 0;
    if (a.detect) {
      // ---- _local_maxima_1d + the height condition of find_peaks -------------------------
      for (int base = 0; base < n; base += 32) {
        const int i = base + lane;
        bool pk = false;
        int mid = 0;
        if (i >= 1 && i < n - 1) {
          const double xi = xs[i];
          if (xs[i - 1] < xi) {
            int ia = i + 1;
            while (ia < n - 1 && xs[ia] == xi) ia++;
            if (xs[ia] < xi) {
              mid = (i + ia - 1) / 2;
              pk = a.height <= xs[mid];
            }
          }
        }
        const unsigned mask = __ballot_sync(FULL, pk);
        const int slot = count + __popc(mask & ((1u << lane) - 1u));
        if (pk && slot < P) idx_s[wid][slot] = mid;
        count += __popc(mask);
      }
      __syncwarp();
    } else {
      count = a.n_peaks[vox];
      if (a.idx && lane < P && lane < count) idx_s[wid][lane] = a.idx[vox * P + lane];
      __syncwarp();
    }
    const int np = count < P ? count : P;
    // ---- one lane per peak: position and weight ----------------------------------------------
    double dv = qnan, fv = qnan;
    if (lane < np) {
      const bool have_idx = a.detect || a.idx != nullptr;
      const int pk = have_idx ? idx_s[wid][lane] : 0;
      if (have_idx && a.bins) dv = a.bins[pk];
      else if (a.d) dv = a.d[vox * P + lane];
      fv = a.detect ? xs[pk] : a.f[vox * P + lane];
      if (a.areas) {
        const double xp = xs[pk];
        // _peak_prominences (wlen = None)
        int i = pk, lbase = pk, rbase = pk;
        double lmin = xp, rmin = xp;
        while (0 <= i && xs[i] <= xp) {
          if (xs[i] < lmin) { lmin = xs[i]; lbase = i; }
          i--;
        }
        i = pk;
        while (i <= n - 1 && xs[i] <= xp) {
          if (xs[i] < rmin) { rmin = xs[i]; rbase = i; }
          i++;
        }
        const double prom = xp - (lmin > rmin ? lmin : rmin);
        // _peak_widths
        const double h = __dsub_rn(xp, __dmul_rn(prom, a.rel_height));  // no FMA: SciPy's C does not contract
        i = pk;
        while (lbase < i && h < xs[i]) i--;
        double left_ip = (double)i;
        if (xs[i] < h) left_ip += (h - xs[i]) / (xs[i + 1] - xs[i]);
        i = pk;
        while (i < rbase && h < xs[i]) i++;
        double right_ip = (double)i;
        if (xs[i] < h) right_ip -= (h - xs[i]) / (xs[i - 1] - xs[i]);
        const double width = right_ip - left_ip;
        fv = __dmul_rn(__ddiv_rn(__dmul_rn(fv, width), kFwhm), kSqrt2Pi);
      }
    }
    if (lane < 32) { ds[lane] = dv; fs[lane] = fv; }
    __syncwarp();
    if (a.normalize) {
      double total = 0.0;
      for (int p = 0; p < np; p++) total += fs[p];
      if (total > 0.0 && lane < np) fv = fv / total;
      __syncwarp();
      fs[lane] = fv;
      __syncwarp();
    }
    if (a.detect && lane == 0) a.n_peaks[vox] = count;
    if (lane < P) {
      if (a.detect && a.idx) a.idx[vox * P + lane] = lane < np ? idx_s[wid][lane] : -1;
      if (a.d) a.d[vox * P + lane] = dv;
      if (a.f) a.f[vox * P + lane] = fv;
    }
    // ---- apply_cutoffs: one lane per range ---------------------------------------------------
    if (a.n_cut > 0) {
      double dc = qnan, fc = qnan;
      if (lane < a.n_cut) {
        const double lo = a.cutoffs[2 * lane], hi = a.cutoffs[2 * lane + 1];
        int cnt = 0;
        double sum = 0.0, d1 = 0.0, f1 = 0.0;
        for (int p = 0; p < np; p++) {
          const double dp = ds[p];
          if (dp >= lo && dp <= hi) {
            if (cnt == 0) { d1 = dp; f1 = fs[p]; }
            sum += fs[p];
            cnt++;
          }
        }
        if (cnt == 1) {
          dc = d1;
          fc = f1;
        } else if (cnt > 1) {
          // geometric_mean_peak: log10(prod(d ** (f / sum f))), sum f
          double prod = 1.0;
          for (int p = 0; p < np; p++) {
            const double dp = ds[p];
            if (dp >= lo && dp <= hi) prod *= pow(dp, fs[p] / sum);
          }
          dc = log10(prod);
          fc = sum;
        }
      }
      cs[lane] = fc;
      __syncwarp();
      if (a.cut_normalize) {
        double total = 0.0;
        for (int c = 0; c < a.n_cut; c++) {
          const double v = cs[c];
          if (v == v) total += v;
        }
        if (total > 0.0) fc = fc / total;
      }
      if (lane < a.n_cut) {
        a.d_cut[vox * a.n_cut + lane] = dc;
        a.f_cut[vox * a.n_cut + lane] = fc;
      }
    }
    __syncwarp();
  }
}

inline size_t spectrum_smem_bytes(int n) { return sizeof(double) * kSpecWarps * (((n + 1) & ~1) + 3 * 32); }

}  // namespace pnb
