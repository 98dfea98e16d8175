/*
 * pyneapple_b200 — C ABI of the B200 voxel-fitting engine (libpnb200.so).
 *
 * This is the drop-in boundary: plain C, plain pointers and sizes, no torch
 * types.  Every entry point names the reference interface it replaces
 * (paths relative to darksim33/Pyneapple, src/pyneapple/).  Array layouts are
 * the reference's own, so a binding needs no repacking:
 *
 *   ydata / signal   (n_vox, n_b)       row-major  -- fitters/base.py:297-308
 *   p0, lb, ub       (n_params, n_vox)  row-major  -- utility/validation.py:177-203, 248-297
 *   params (popt)    (n_params, n_vox)  row-major  -- solvers/curvefit.py:231-233
 *   cov (pcov)       (n_vox, n_free, n_free)       -- solvers/curvefit.py:234-243
 *   coefficients     (n_vox, n_bins)               -- solvers/nnls_solver.py:175
 *
 * `*_host` entry points take HOST pointers and run the whole
 * upload / solve / download pipeline (chunked, copies overlapped with the
 * kernels on separate streams); `*_device` entry points take DEVICE pointers
 * and only enqueue work on the given CUDA stream.
 *
 * Return value: 0 on success, otherwise a CUDA error code (>0) or a
 * PNB_E_* code (<0); pnb_last_error() gives the text.  There is no CPU
 * fallback anywhere in this library.
 */
#ifndef PYNEAPPLE_B200_H
#define PYNEAPPLE_B200_H

#include <stdint.h>

#ifdef __cplusplus
extern "C" {
#endif

#define PNB_ABI_VERSION 4

/* model_id: parameter order is the reference's `_all_param_names`
 * (models/monoexp.py:91-105, models/biexp.py:103-126, models/triexp.py:103-128) */
enum {
  PNB_MODEL_MONO = 0,        /* [S0, D]                     */
  PNB_MODEL_BI_REDUCED = 1,  /* [f1, D1, D2]                */
  PNB_MODEL_BI_FULL = 2,     /* [f1, D1, f2, D2]            */
  PNB_MODEL_BI_S0 = 3,       /* [f1, D1, D2, S0]            */
  PNB_MODEL_TRI_REDUCED = 4, /* [f1, D1, f2, D2, D3]        */
  PNB_MODEL_TRI_FULL = 5,    /* [f1, D1, f2, D2, f3, D3]    */
  PNB_MODEL_TRI_S0 = 6       /* [f1, D1, f2, D2, D3, S0]    */
};
/* t1_mode appends the parameter T1 (model_functions/multiexp.py:210-302) */
enum { PNB_T1_NONE = 0, PNB_T1_STANDARD = 1, PNB_T1_STEAM = 2 };
/* SciPy least_squares method behind curve_fit (solvers/curvefit.py:295-306, `method` keyword):
 * scipy/optimize/_lsq/trf.py (trf_bounds) or scipy/optimize/_lsq/dogbox.py */
enum { PNB_METHOD_TRF = 0, PNB_METHOD_DOGBOX = 1,
       /* curve_fit(method="lm") -> leastsq -> MINPACK lmdif / lmder (scipy/optimize/_minpack_py.py);
        * unbounded problems only (the caller passes -inf / +inf bounds); per-voxel status: 1 gtol,
        * 2 ftol, 3 xtol, 4 both (success); 0 maxfev reached, -6 / -7 / -8 = leastsq's info 6 / 7 / 8
        * ("ftol / xtol / gtol is too small"): failures, params = p0 */
       PNB_METHOD_LM = 2 };
/* least_squares(loss=...) (scipy/optimize/_lsq/least_squares.py: soft_l1, huber, cauchy, arctan), forwarded by the
 * reference from solver_kwargs (solvers/curvefit.py:70-73, 305) */
enum { PNB_LOSS_LINEAR = 0, PNB_LOSS_SOFT_L1 = 1, PNB_LOSS_HUBER = 2, PNB_LOSS_CAUCHY = 3, PNB_LOSS_ARCTAN = 4 };

enum {
  PNB_E_BADARG = -1,      /* inconsistent sizes / null pointers */
  PNB_E_UNSUPPORTED = -2, /* model / option without a device implementation */
  PNB_E_NODEVICE = -3     /* no CUDA device */
};

/* per-voxel status written by pnb_trf_fit_*: SciPy's least_squares status
 * (0 max_nfev reached = failure, 1 gtol, 2 ftol, 3 xtol, 4 ftol and xtol), or
 * the input error SciPy would raise for that voxel.  success <=> status > 0;
 * on failure params = p0 and cov = NaN (solvers/curvefit.py:308-317). */
enum {
  PNB_ST_BAD_BOUNDS = -1,  /* "Each lower bound must be strictly less than each upper bound." */
  PNB_ST_INFEASIBLE = -2,  /* "Initial guess is outside of provided bounds" */
  PNB_ST_NONFINITE_Y = -3, /* "array must not contain infs or NaNs" */
  PNB_ST_NONFINITE_F0 = -4, /* "Residuals are not finite in the initial point." */
  PNB_ST_LM_FTOL_SMALL = -6, PNB_ST_LM_XTOL_SMALL = -7, PNB_ST_LM_GTOL_SMALL = -8 /* PNB_METHOD_LM */
};

/*
 * Bounded non-linear least squares for n_vox voxels.
 * Replaces CurveFitSolver._fit_data / _fit_single_pixel
 * (solvers/curvefit.py:171-317), i.e. one scipy.optimize.curve_fit(
 * method="trf", maxfev=max_nfev, ftol=ftol) call per voxel.
 */
typedef struct pnb_trf_problem {
  int32_t model_id;          /* PNB_MODEL_*                                   */
  int32_t t1_mode;           /* PNB_T1_*                                      */
  double repetition_time;    /* TR, used when t1_mode != 0                    */
  double mixing_time;        /* TM, used when t1_mode == PNB_T1_STEAM         */
  int32_t n_b;               /* number of measurements (b-values)             */
  int32_t n_params;          /* number of model parameters incl. fixed ones   */
  int64_t n_vox;
  const double *xdata;       /* (n_b)                                          */
  const double *ydata;       /* (n_vox, n_b)                                   */
  const double *p0;          /* (n_params) or (n_params, n_vox)                */
  const double *lb;          /* like p0                                        */
  const double *ub;
  int32_t p0_per_voxel;      /* 0: one vector for all voxels, 1: per voxel     */
  int32_t bounds_per_voxel;
  uint32_t frozen_mask;      /* bit j: parameter j is fixed at p0[j] (per voxel
                                when p0_per_voxel) -- model.fixed_params and
                                pixel_fixed_params of curvefit.py:274-288      */
  int32_t max_nfev;          /* CurveFitSolver.max_iter                        */
  double ftol;               /* CurveFitSolver.tol                             */
  double xtol;               /* SciPy default 1e-8                             */
  double gtol;               /* SciPy default 1e-8                             */
  int32_t jac_mode;          /* 0 analytic, 1 SciPy '2-point' finite differences (trf / dogbox),
                                2 MINPACK forward differences, fdjac2 (lm)        */
  int32_t x_scale_jac;       /* 1: x_scale='jac'                               */
  int32_t method;            /* PNB_METHOD_TRF (curve_fit's default with bounds) or
                                PNB_METHOD_DOGBOX: least_squares(method=...)      */
  int32_t finish_wait;       /* scheduling hint, results do not depend on it: how many passes a converged
                                lane of the kernel waits for the other lanes of its warp before it writes
                                its results and takes the next voxel (0 = the default, 3; fits that take
                                many evaluations per voxel, like the tight tolerances of the constrained
                                solver, run ~6 % faster with 6)                                   */
  double x_scale[8];         /* per parameter, 1.0 = SciPy default             */
  /* curve_fit extras (same forwarding): any of weights / loss / diff_step selects a separate kernel instantiation,
     built for method trf without a T1 parameter; PNB_E_UNSUPPORTED otherwise */
  const double *weights;     /* (n_b) 1 / sigma of curve_fit(sigma=<1-D>), or NULL; host memory for the
                                _host entry points, device memory for _device          */
  double diff_step[8];       /* relative step of the 2-point Jacobian per parameter (least_squares
                                diff_step), 0 = SciPy's default sqrt(eps) max(1, |x|)    */
  int32_t loss;              /* PNB_LOSS_*                                      */
  int32_t absolute_sigma;    /* curve_fit(absolute_sigma=True): cov not scaled by 2 cost / (m - n) */
  double f_scale;            /* least_squares f_scale (soft margin of the robust losses), > 0; 0 = 1.0 */
  /* outputs */
  double *params;            /* (n_params, n_vox); fixed rows repeat the fixed value */
  double *cov;               /* (n_vox, n_free, n_free) or NULL.  pnb_trf_fit_host also accepts
                                DEVICE memory of the fitting GPU here: the covariances then stay
                                on the GPU (no D2H traffic) for the caller to fetch on demand;
                                they are complete when the call returns (one covariance pass
                                over the range follows the last chunk's solver kernel)         */
  int32_t *status;           /* (n_vox)                                        */
  int32_t *nfev;             /* (n_vox) residual evaluations, SciPy's count     */
  int32_t *njev;             /* (n_vox) or NULL                                */
  double *cost;              /* (n_vox) 0.5 * ||f||^2 at the solution, or NULL */
  double *r_squared;         /* (n_vox) R^2 of fitters/base.py:142-186 at the returned
                                parameters (NaN for a constant signal), or NULL  */
} pnb_trf_problem;

int pnb_trf_fit_device(const pnb_trf_problem *prob, void *cuda_stream);
int pnb_trf_fit_host(const pnb_trf_problem *prob, int device, int64_t chunk_vox);
/* The same call spread over n_devices GPUs of the node (devices[i], or 0 .. n_devices-1 when NULL):
 * the voxels are cut into contiguous, balanced ranges (range i = voxels [i n / N, ...), the z-slabs
 * of SURVEY.md §8e), one host pipeline per GPU runs concurrently, and every GPU copies its results
 * straight into its part of the caller's arrays — one process, no gather step.  Replaces the joblib
 * pool of CurveFitSolver._fit_data (solvers/curvefit.py:201-229) at node scale.
 * cov_per_device: NULL (covariances go to prob->cov on the host, or nowhere when that is NULL), or
 * n_devices DEVICE buffers, buffer i on GPU i holding its range's (n_range, n_free, n_free). */
int pnb_trf_fit_host_multi(const pnb_trf_problem *prob, const int32_t *devices, int32_t n_devices,
                           int64_t chunk_vox, double *const *cov_per_device);
/* voxels of the most recent pnb_trf_fit_host / _host_multi call of this process that ended with
 * status <= 0, counted by the kernels (a pass over the status array costs the host milliseconds) */
int64_t pnb_trf_last_failed_count(void);

/*
 * Tikhonov-regularised non-negative least squares for n_vox voxels on a shared
 * dictionary.  Replaces NNLSSolver._fit_data / _fit_single_pixel
 * (solvers/nnls_solver.py:129-210), i.e. one scipy.optimize.nnls([basis; mu R],
 * [signal; 0], maxiter=max_iter) call per voxel.  The caller passes the plain
 * basis (model_functions/nnls.py:31-43) and the band of mu^2 R^T R of the
 * regularisation matrix (model_functions/nnls.py:46-85); both are built once
 * per fit on the host with the reference's own formulas.
 */
typedef struct pnb_nnls_problem {
  int32_t n_b;               /* measurements                                    */
  int32_t n_bins;            /* dictionary size                                 */
  int32_t rtr_halfband;      /* W: (R^T R)[i][j] = 0 for |i-j| > W              */
  int32_t max_iter;          /* NNLSSolver.max_iter (Lawson-Hanson iteration cap) */
  int32_t algorithm;         /* 0 auto: inverse-update fast path, voxels it cannot certify
                                are re-solved by the robust path; 1 robust (Cholesky +
                                refinement) only -- use for un-regularised problems */
  int32_t dual_init;         /* how h = B^T y (the first dual) is computed on the fast path:
                                0 fused into the solver kernel's first dual pass (default),
                                1 materialised for all voxels by one dense FP64 tensor-core GEMM
                                  (mma.sync m8n8k4) that the solver kernel then reads — the two
                                  forms BASELINE.json's north_star (2) asks to be measured        */
  int64_t n_vox;
  const double *basis;       /* (n_b, n_bins)                                   */
  const double *rtr_band;    /* (n_bins, 2W+1): [j][d+W] = (mu^2 R^T R)[j][j+d] */
  const double *signal;      /* (n_vox, n_b)                                    */
  double *coefficients;      /* (n_vox, n_bins)                                 */
  double *residual;          /* (n_vox) ||[basis; mu R] x - [signal; 0]||_2     */
  int32_t *status;           /* (n_vox) 1 converged; 3 iteration cap, 2 non-finite
                                signal: coefficients = 0, residual = ||signal|| */
  int32_t *iterations;       /* (n_vox) Lawson-Hanson iteration count           */
  double *r_squared;         /* (n_vox) R^2 of the un-regularised prediction, or NULL */
} pnb_nnls_problem;

/* Device-path launches on one device share one set of scratch buffers: the library orders each
 * launch after the previous one on that device (event wait on the given stream), so launches on
 * different streams are safe but do not overlap. */
int pnb_nnls_fit_device(const pnb_nnls_problem *prob, void *cuda_stream);
int pnb_nnls_fit_host(const pnb_nnls_problem *prob, int device, int64_t chunk_vox);
/* multi-GPU form, see pnb_trf_fit_host_multi (replaces the joblib pool of NNLSSolver._fit_data,
 * solvers/nnls_solver.py:153-172) */
int pnb_nnls_fit_host_multi(const pnb_nnls_problem *prob, const int32_t *devices, int32_t n_devices,
                            int64_t chunk_vox);
int pnb_sizeof_nnls_problem(void);
/* H0 (n_vox, n_bins) = signal (n_vox, n_b) x basis (n_b, n_bins) on the FP64 tensor cores: the
 * batched A^T y of NNLSSolver._fit_single_pixel's first Lawson-Hanson step (solvers/nnls_solver.py:
 * 195-197) as one GEMM; DEVICE pointers, n_b <= 32 */
int pnb_nnls_dual_gemm_device(int32_t n_b, int32_t n_bins, int64_t n_vox, const double *basis,
                              const double *signal, double *h0, void *cuda_stream);
/* voxels of the most recent auto-mode launch on `device` that were re-solved by the robust path
 * (synchronises the device) */
int64_t pnb_nnls_last_redo_count(int device);

/*
 * In-plane (x, y) resampling of an (H, W, inner) array for all `inner` =
 * slices x channels planes at once, bit-faithful to the per-slice, per-channel
 * cv2.resize(plane, (W', H'), interpolation) loop of IDEALFitter._interpolate_array
 * (fitters/ideal.py:299-320).  FP64 cubic is bit-identical to OpenCV.
 */
typedef struct pnb_resize_problem {
  int32_t dtype;             /* 0 float64, 1 float32                            */
  int32_t method;            /* 0 cv2.INTER_LINEAR, 1 cv2.INTER_CUBIC           */
  int32_t src_h, src_w;      /* source extent of the two leading axes (x, y)    */
  int32_t dst_h, dst_w;      /* target extent                                   */
  int64_t inner;             /* product of the trailing axes (z * channels)     */
  const void *src;           /* (src_h, src_w, inner) C-contiguous              */
  void *dst;                 /* (dst_h, dst_w, inner)                           */
} pnb_resize_problem;

int pnb_resize2d_device(const pnb_resize_problem *prob, void *cuda_stream);
int pnb_resize2d_host(const pnb_resize_problem *prob, int device);

/*
 * Post-processing of NNLS spectra for n_vox voxels: peak detection, Gaussian peak areas and
 * cut-off ranges.  Replaces the per-voxel loop over find_spectrum_peaks / calculate_peak_area /
 * apply_cutoffs / geometric_mean_peak (utility/spectrum.py:13-206), i.e. one
 * scipy.signal.find_peaks(x, height=h) + peak_widths(x, peaks, rel_height) call per voxel.
 * Stages (each optional):
 *   detect = 1   peaks of `spectrum` with x[peak] >= height -> n_peaks, peak_index, d_values =
 *                bins[peak], f_values = x[peak];  detect = 0: the peak list is an input
 *                (n_peaks, and peak_index and / or d_values, f_values)
 *   areas = 1    f_values = height * FWHM-at-rel_height / (2 sqrt(2 ln 2)) * sqrt(2 pi)
 *                (`regularized=True` of find_spectrum_peaks)
 *   normalize    f_values /= sum(f_values)  (when the sum is positive)
 *   n_cutoffs>0  d_cut, f_cut per range (lo, hi): NaN without a peak, the peak itself, or
 *                (log10 of the weighted geometric mean, summed weight) of several;
 *                cut_normalize: f_cut /= nansum(f_cut)
 * At most max_peaks (<= 32) peaks are stored per voxel, NaN / -1 padded; n_peaks holds the
 * true count, so a caller can see truncation.
 */
typedef struct pnb_spectrum_problem {
  int32_t n_bins;
  int32_t max_peaks;
  int32_t detect, areas, normalize;
  int32_t n_cutoffs, cut_normalize;
  int32_t reserved;
  double height;             /* find_peaks(height=...)                          */
  double rel_height;         /* peak_widths(rel_height=...), 0.5 = FWHM         */
  int64_t n_vox;
  const double *bins;        /* (n_bins) or NULL when d_values is an input       */
  const double *cutoffs;     /* (n_cutoffs, 2) or NULL                           */
  const double *spectrum;    /* (n_vox, n_bins) or NULL (cut-offs of given peaks) */
  int32_t *n_peaks;          /* (n_vox)                                          */
  int32_t *peak_index;       /* (n_vox, max_peaks) or NULL                       */
  double *d_values;          /* (n_vox, max_peaks) or NULL                       */
  double *f_values;          /* (n_vox, max_peaks)                               */
  double *d_cut;             /* (n_vox, n_cutoffs) or NULL                       */
  double *f_cut;             /* (n_vox, n_cutoffs) or NULL                       */
} pnb_spectrum_problem;

int pnb_spectrum_peaks_device(const pnb_spectrum_problem *prob, void *cuda_stream);
int pnb_spectrum_peaks_host(const pnb_spectrum_problem *prob, int device, int64_t chunk_vox);
int pnb_sizeof_spectrum_problem(void);

/*
 * Mean signal of every segmentation label: replaces the per-label
 * np.mean(image[segmentation == seg], axis=0) loop of
 * SegmentationWiseFitter._extract_segmentation_mean_signals (fitters/segmentationwise.py:112-137).
 * `label` holds the dense index of each voxel's label (0 .. n_labels-1, the position in
 * np.unique(segmentation)); voxels with an index outside that range are skipped.  An empty label
 * gets NaN, like np.mean of an empty selection.  Deterministic (no floating-point atomics); the
 * order of the additions differs from NumPy's, so the means agree to a few ulp.
 */
typedef struct pnb_segmeans_problem {
  int32_t n_b;
  int32_t n_labels;
  int64_t n_vox;             /* all voxels of the volume, C order                 */
  const double *image;       /* (n_vox, n_b)                                      */
  const int32_t *label;      /* (n_vox)                                           */
  double *means;             /* (n_labels, n_b)                                   */
  int64_t *counts;           /* (n_labels) voxels per label, or NULL              */
} pnb_segmeans_problem;

int pnb_segment_means_device(const pnb_segmeans_problem *prob, void *cuda_stream);
int pnb_segment_means_host(const pnb_segmeans_problem *prob, int device);

/*
 * Model signal for every fitted voxel: replaces the per-voxel model.forward loop of
 * BaseFitter.predict (fitters/base.py:93-128).  With flat_index the rows are scattered into an
 * (n_out, n_b) volume that is zero where nothing was fitted (BaseFitter._reconstruct_volume,
 * fitters/base.py:310-330).  All pointers are DEVICE pointers.
 */
typedef struct pnb_predict_problem {
  int32_t model_id;          /* PNB_MODEL_*                                       */
  int32_t t1_mode;           /* PNB_T1_*                                          */
  double repetition_time, mixing_time;
  int32_t n_b;               /* len(xdata)                                        */
  int32_t n_params;          /* rows of params (all model parameters, fixed ones included) */
  int64_t n_vox;             /* fitted voxels                                     */
  int64_t n_out;             /* rows of signal: n_vox, or the volume's voxel count with flat_index */
  const double *xdata;       /* (n_b)                                             */
  const double *params;      /* (n_params, n_vox)                                 */
  const int64_t *flat_index; /* (n_vox) C-order position of each voxel in the volume, or NULL */
  double *signal;            /* (n_out, n_b)                                      */
} pnb_predict_problem;

int pnb_predict_device(const pnb_predict_problem *prob, void *cuda_stream);
int pnb_sizeof_predict_problem(void);

/*
 * Row gather / scatter between a masked voxel list and a volume (DEVICE pointers):
 *   direction 0  dst[i, :] = src[index[i], :]   image[segmentation != 0] of
 *                BaseFitter._extract_pixel_data (fitters/base.py:280-308)
 *   direction 1  dst[index[i], :] = src[i, :]   vol[idx] = values of _reconstruct_volume
 *                (fitters/base.py:310-330) and reconstruct_maps (io/nifti.py:279-312; out_dtype 1
 *                writes float32 like that function); zero_fill clears dst first.
 */
typedef struct pnb_rows_problem {
  int32_t direction;         /* 0 gather, 1 scatter                               */
  int32_t out_dtype;         /* 0 float64, 1 float32 (source is always float64)   */
  int32_t width;             /* elements per row                                  */
  int32_t zero_fill;         /* scatter: memset dst (n_other rows) first          */
  int64_t n_rows;            /* rows moved = len(index)                           */
  int64_t n_other;           /* rows of the array on the indexed side             */
  const double *src;
  void *dst;
  const int64_t *index;      /* (n_rows)                                          */
} pnb_rows_problem;

int pnb_move_rows_device(const pnb_rows_problem *prob, void *cuda_stream);
int pnb_sizeof_rows_problem(void);

/* housekeeping */
int pnb_abi_version(void);
/* sizeof(struct pnb_trf_problem) as compiled, for binding self-checks */
int pnb_sizeof_trf_problem(void);
const char *pnb_last_error(void);
int pnb_device_count(void);
/* number of kernels this library has launched since load (bench.py: gpu_launches) */
int64_t pnb_launch_count(void);
/* pinned host memory for callers that want full-speed transfers */
int pnb_host_alloc(void **ptr, int64_t bytes);
int pnb_host_free(void *ptr);
/* Synchronous copies between device memory and host memory of any kind.  Pageable host memory —
 * the numpy arrays the reference's fitters hold (fitters/base.py:280-330) — is staged in 32 MB
 * pieces through two page-locked blocks with multi-threaded host copies (~5x cudaMemcpy on such
 * memory); page-locked memory is copied directly.  `after_stream`: the stream whose work produces
 * dev_src (synchronised first); `then_stream`: the stream that uses dev_dst — the copies are ordered
 * after the work already queued on it (dev_dst may be recycled memory that work still reads), and
 * the call returns when the data is on the device. */
int pnb_download(void *host_dst, const void *dev_src, int64_t bytes, void *after_stream);
int pnb_upload(void *dev_dst, const void *host_src, int64_t bytes, void *then_stream);
/* Asynchronous device-to-device copy on a stream of the current device; dst may be peer memory
 * (another GPU of the node, e.g. rank 0's gather buffer mapped through CUDA IPC): the copy engines
 * move the block over NVLink — the one data-path exchange of the multi-GPU mode. */
int pnb_copy_d2d(void *dst, const void *src, int64_t bytes, void *cuda_stream);
/* FP64 FMA micro-benchmark: achieved TFLOP/s of dependent-chain-free DFMA (roofline denominator) */
int pnb_measure_fp64_peak(int device, double *tflops);

#ifdef __cplusplus
}
#endif
#endif /* PYNEAPPLE_B200_H */
