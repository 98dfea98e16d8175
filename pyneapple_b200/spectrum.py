"""NNLS spectrum post-processing on the GPU.

Mirror of the reference's ``pyneapple.utility.spectrum`` (utility/spectrum.py:13-206:
``calculate_peak_area``, ``find_spectrum_peaks``, ``geometric_mean_peak``,
``apply_cutoffs`` — same names, arguments and return values for one spectrum) plus
the batched form a volume needs, :func:`find_spectrum_peaks_batch`, which does all
voxels in one kernel launch (``pnb_spectrum_peaks_*``).  Feeding it the CUDA tensor
``engine.nnls_fit`` returns keeps the 2000-byte-per-voxel spectra on the device: only
a few peaks per voxel travel to the host.

There is no CPU implementation here: every function calls the CUDA library.
"""

from __future__ import annotations

import ctypes as C

import numpy as np

from . import _lib
from .engine import _is_torch_cuda

__all__ = ["calculate_peak_area", "find_spectrum_peaks", "geometric_mean_peak", "apply_cutoffs",
           "find_spectrum_peaks_batch"]


def _ptr(a):
    return None if a is None else (a.data_ptr() if _is_torch_cuda(a) else a.ctypes.data)


def _run(*, spectrum=None, bins=None, n_peaks=None, peak_index=None, d_values=None, f_values=None,
         detect, areas, normalize, height=0.0, rel_height=0.5, cutoffs=None, cut_normalize=True,
         max_peaks, device=0, chunk_vox=0):
    """One call of ``pnb_spectrum_peaks_{host,device}``; arrays are numpy (host path) or CUDA tensors."""
    _lib.require_device()
    lib = _lib.load()
    on_device = _is_torch_cuda(spectrum) or _is_torch_cuda(f_values) or _is_torch_cuda(d_values)
    if on_device:
        import torch

        ref = next(t for t in (spectrum, f_values, d_values) if _is_torch_cuda(t))
        dev = ref.device

        def arr(x, dtype):
            if x is None:
                return None
            return torch.as_tensor(x, dtype=dtype, device=dev).contiguous()

        def empty(shape, dtype):
            return torch.empty(shape, dtype=dtype, device=dev)

        f64, i32 = torch.float64, torch.int32
    else:
        def arr(x, dtype):
            if x is None:
                return None
            return np.ascontiguousarray(x, dtype=dtype)

        def empty(shape, dtype):
            return np.empty(shape, dtype=dtype)

        f64, i32 = np.float64, np.int32
    spectrum = arr(spectrum, f64)
    if spectrum is not None and spectrum.ndim != 2:
        raise ValueError(f"spectrum must be (n_vox, n_bins), got {tuple(spectrum.shape)}")
    bins = arr(bins, f64)
    P = int(max_peaks)
    if spectrum is not None:
        n_vox, n_bins = int(spectrum.shape[0]), int(spectrum.shape[1])
        if bins is not None and int(bins.shape[0]) != n_bins:
            raise ValueError(f"bins has {int(bins.shape[0])} entries, the spectrum {n_bins} bins")
    else:
        n_vox, n_bins = int(f_values.shape[0]), (0 if bins is None else int(bins.shape[0]))
    cut = None if cutoffs is None or len(cutoffs) == 0 else arr(np.asarray(cutoffs, dtype=np.float64).reshape(-1, 2), f64)
    K = 0 if cut is None else int(cut.shape[0])
    if detect:
        n_peaks, peak_index = empty((n_vox,), i32), empty((n_vox, P), i32)
        d_values, f_values = empty((n_vox, P), f64), empty((n_vox, P), f64)
    else:
        n_peaks = arr(n_peaks, i32)
        peak_index = arr(peak_index, i32)
        f_values = arr(f_values, f64)
        f_values = f_values.clone() if on_device else f_values.copy()
        if d_values is None and bins is not None and peak_index is not None:
            d_values = empty((n_vox, P), f64)
        elif d_values is not None:
            d_values = arr(d_values, f64)
            d_values = d_values.clone() if on_device else d_values.copy()
    d_cut = empty((n_vox, K), f64) if K else None
    f_cut = empty((n_vox, K), f64) if K else None
    prob = _lib.SpectrumProblem()
    prob.n_bins, prob.max_peaks = n_bins, P
    prob.detect, prob.areas, prob.normalize = int(detect), int(areas), int(normalize)
    prob.n_cutoffs, prob.cut_normalize = K, int(cut_normalize)
    prob.height, prob.rel_height, prob.n_vox = float(height), float(rel_height), n_vox
    prob.bins, prob.cutoffs, prob.spectrum = _ptr(bins), _ptr(cut), _ptr(spectrum)
    prob.n_peaks, prob.peak_index = _ptr(n_peaks), _ptr(peak_index)
    prob.d_values, prob.f_values, prob.d_cut, prob.f_cut = _ptr(d_values), _ptr(f_values), _ptr(d_cut), _ptr(f_cut)
    if on_device:
        import torch

        with torch.cuda.device(dev):
            stream = torch.cuda.current_stream(dev).cuda_stream
            _lib.check(lib.pnb_spectrum_peaks_device(C.byref(prob), C.c_void_p(stream)), "pnb_spectrum_peaks_device")
    else:
        _lib.check(lib.pnb_spectrum_peaks_host(C.byref(prob), int(device), int(chunk_vox)), "pnb_spectrum_peaks_host")
    return dict(n_peaks=n_peaks, peak_index=peak_index, d_values=d_values, f_values=f_values, d_cut=d_cut, f_cut=f_cut)


def find_spectrum_peaks_batch(spectra, bins, height: float = 0.1, regularized: bool = False, *,
                              cutoffs=None, max_peaks: int = 8, device: int = 0, chunk_vox: int = 0) -> dict:
    """``find_spectrum_peaks`` (and, with ``cutoffs``, ``apply_cutoffs``) for every row of ``spectra``.

    ``spectra``: ``(n_vox, n_bins)`` numpy array or CUDA tensor (results then stay on that device).
    Returns ``n_peaks (n_vox,)`` — the number of peaks found, which may exceed ``max_peaks`` —,
    ``peak_index``, ``d_values``, ``f_values`` ``(n_vox, max_peaks)`` padded with -1 / NaN and, with
    ``cutoffs``, ``d_cut``, ``f_cut`` ``(n_vox, len(cutoffs))``.  When a voxel has more than
    ``max_peaks`` peaks only the first ``max_peaks`` (in bin order) are stored and normalised.
    """
    if not 1 <= int(max_peaks) <= 32:
        raise ValueError("max_peaks must be in [1, 32]")
    return _run(spectrum=spectra, bins=bins, detect=True, areas=bool(regularized), normalize=True, height=height,
                cutoffs=cutoffs, max_peaks=max_peaks, device=device, chunk_vox=chunk_vox)


def find_spectrum_peaks(spectrum, bins, height: float = 0.1, regularized: bool = False):
    """Peaks of one spectrum: ``(d_values, f_values)``, fractions normalised (spectrum.py:50-103)."""
    spectrum = np.asarray(spectrum, dtype=np.float64)
    if spectrum.ndim != 1:
        raise ValueError("spectrum must be 1-D; use find_spectrum_peaks_batch for (n_vox, n_bins)")
    r = _run(spectrum=spectrum[None, :], bins=np.asarray(bins, dtype=np.float64), detect=True, areas=bool(regularized),
             normalize=True, height=height, max_peaks=32)
    k = int(r["n_peaks"][0])
    if k > 32:
        raise ValueError(f"{k} peaks in one spectrum; at most 32 are supported")
    if k == 0:
        return np.array([]), np.array([])
    return r["d_values"][0, :k].copy(), r["f_values"][0, :k].copy()


def calculate_peak_area(spectrum, peak_indices, peak_heights, rel_height: float = 0.5) -> list:
    """Gaussian area of the given peaks of one spectrum (spectrum.py:13-47)."""
    spectrum = np.asarray(spectrum, dtype=np.float64)
    idx = np.asarray(peak_indices, dtype=np.int32).ravel()
    k = idx.size
    if k == 0:
        return []
    if k > 32:
        raise ValueError("at most 32 peaks per spectrum are supported")
    if idx.min() < 0 or idx.max() >= spectrum.size:
        raise ValueError("peak index outside the spectrum")
    r = _run(spectrum=spectrum[None, :], n_peaks=np.array([k], np.int32), peak_index=idx[None, :],
             f_values=np.asarray(peak_heights, dtype=np.float64).reshape(1, k), detect=False, areas=True,
             normalize=False, rel_height=rel_height, max_peaks=k)
    return [float(v) for v in r["f_values"][0]]


def geometric_mean_peak(positions, heights):
    """``(log10 of the weighted geometric mean position, summed height)`` (spectrum.py:106-136)."""
    d = np.asarray(positions, dtype=np.float64).ravel()
    f = np.asarray(heights, dtype=np.float64).ravel()
    k = d.size
    if k == 0 or k > 32:
        raise ValueError("between 1 and 32 peaks are supported")
    if k == 1:
        # one peak is its own mean; the kernel's cut-off stage keeps single peaks untouched
        # (apply_cutoffs semantics), so the logarithm is taken through a two-entry list
        d, f, k = np.array([d[0], d[0]]), np.array([f[0], 0.0]), 2
    r = _run(n_peaks=np.array([k], np.int32), d_values=d[None, :], f_values=f[None, :], detect=False, areas=False,
             normalize=False, cutoffs=[(-np.inf, np.inf)], cut_normalize=False, max_peaks=k)
    return float(r["d_cut"][0, 0]), float(r["f_cut"][0, 0])


def apply_cutoffs(d_values, f_values, cutoffs):
    """Merge the peaks inside each cut-off range (spectrum.py:139-206): ``(d_new, f_new)``."""
    d = np.asarray(d_values, dtype=np.float64).ravel()
    f = np.asarray(f_values, dtype=np.float64).ravel()
    k = d.size
    if k > 32:
        raise ValueError("at most 32 peaks per spectrum are supported")
    if len(cutoffs) == 0:
        return np.array([]), np.array([])
    P = max(k, 1)
    dd, ff = np.full((1, P), np.nan), np.full((1, P), np.nan)
    dd[0, :k], ff[0, :k] = d, f
    r = _run(n_peaks=np.array([k], np.int32), d_values=dd, f_values=ff, detect=False, areas=False, normalize=False,
             cutoffs=cutoffs, cut_normalize=True, max_peaks=P)
    return r["d_cut"][0].copy(), r["f_cut"][0].copy()
