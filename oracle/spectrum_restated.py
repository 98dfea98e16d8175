"""ORACLE (test infrastructure) — plain-Python restatement of the SciPy routines behind
``pyneapple.utility.spectrum`` (utility/spectrum.py:13-206), independent of ``scipy.signal``.

The reference's spectrum post-processing delegates its arithmetic to SciPy
(``scipy/signal/_peak_finding.py``: ``find_peaks``, ``peak_widths``;
``scipy/signal/_peak_finding_utils.pyx``: ``_local_maxima_1d``, ``_peak_prominences``,
``_peak_widths``).  This module restates those published algorithms as loops, one function per
SciPy routine, and the reference's own formulas on top of them.  It is pinned bit for bit to the
outputs of the real reference (``tests/golden/spectrum_peaks.npz``) in
``tests/test_oracle_golden.py``.  Product code never imports it.
"""

from __future__ import annotations

import numpy as np


def local_maxima_1d(x):
    """``_local_maxima_1d``: midpoints of all strict local maxima, plateaus included."""
    n = len(x)
    midpoints = []
    i, i_max = 1, n - 1
    while i < i_max:
        if x[i - 1] < x[i]:
            i_ahead = i + 1
            while i_ahead < i_max and x[i_ahead] == x[i]:
                i_ahead += 1
            if x[i_ahead] < x[i]:
                left_edge, right_edge = i, i_ahead - 1
                midpoints.append((left_edge + right_edge) // 2)
                i = i_ahead
        i += 1
    return np.array(midpoints, dtype=np.intp)


def find_peaks_height(x, height):
    """``find_peaks(x, height=height)``: local maxima with ``x[peak] >= height``."""
    peaks = local_maxima_1d(x)
    keep = np.array([height <= x[p] for p in peaks], dtype=bool)
    peaks = peaks[keep] if len(peaks) else peaks
    return peaks, np.array([x[p] for p in peaks], dtype=np.float64)


def peak_prominences(x, peaks):
    """``_peak_prominences`` with ``wlen=None``: prominence, left base, right base."""
    n = len(x)
    prom, lbs, rbs = [], [], []
    for peak in peaks:
        i_min, i_max = 0, n - 1
        i = left_base = peak
        left_min = x[peak]
        while i_min <= i and x[i] <= x[peak]:
            if x[i] < left_min:
                left_min = x[i]
                left_base = i
            i -= 1
        i = right_base = peak
        right_min = x[peak]
        while i <= i_max and x[i] <= x[peak]:
            if x[i] < right_min:
                right_min = x[i]
                right_base = i
            i += 1
        prom.append(x[peak] - max(left_min, right_min))
        lbs.append(left_base)
        rbs.append(right_base)
    return np.array(prom, dtype=np.float64), lbs, rbs


def peak_widths(x, peaks, rel_height=0.5):
    """``peak_widths(x, peaks, rel_height)[0]`` (``_peak_widths`` on the prominence data)."""
    prom, lbs, rbs = peak_prominences(x, peaks)
    widths = []
    for p, peak in enumerate(peaks):
        i_min, i_max = lbs[p], rbs[p]
        height = x[peak] - prom[p] * rel_height
        i = peak
        while i_min < i and height < x[i]:
            i -= 1
        left_ip = float(i)
        if x[i] < height:
            left_ip += (height - x[i]) / (x[i + 1] - x[i])
        i = peak
        while i < i_max and height < x[i]:
            i += 1
        right_ip = float(i)
        if x[i] < height:
            right_ip -= (height - x[i]) / (x[i - 1] - x[i])
        widths.append(right_ip - left_ip)
    return np.array(widths, dtype=np.float64)


def find_spectrum_peaks(spectrum, bins, height=0.1, regularized=False):
    """utility/spectrum.py:50-103 (and :13-47 for the Gaussian areas)."""
    x = np.asarray(spectrum, dtype=np.float64)
    peaks, raw = find_peaks_height(x, height)
    if len(peaks) == 0:
        return peaks, np.array([]), np.array([])
    if regularized:
        fw = peak_widths(x, peaks, 0.5)
        f = np.array([float(h * w / (2 * np.sqrt(2 * np.log(2))) * np.sqrt(2 * np.pi)) for h, w in zip(raw, fw)])
    else:
        f = raw.copy()
    total = np.sum(f)
    if total > 0:
        f = f / total
    return peaks, np.asarray(bins)[peaks], f


def apply_cutoffs(d_values, f_values, cutoffs):
    """utility/spectrum.py:139-206 with geometric_mean_peak (:106-136) inlined."""
    d = np.asarray(d_values, dtype=float)
    f = np.asarray(f_values, dtype=float)
    new_d, new_f = [], []
    for lo, hi in cutoffs:
        mask = (d >= lo) & (d <= hi)
        _d, _f = d[mask], f[mask]
        if len(_d) == 0:
            new_d.append(float("nan")); new_f.append(float("nan"))
        elif len(_d) == 1:
            new_d.append(float(_d[0])); new_f.append(float(_f[0]))
        else:
            new_d.append(float(np.log10(np.prod(_d ** (_f / np.sum(_f)))))); new_f.append(float(np.sum(_f)))
    d_out, f_out = np.array(new_d), np.array(new_f)
    total = np.nansum(f_out)
    if total > 0:
        f_out = f_out / total
    return d_out, f_out
