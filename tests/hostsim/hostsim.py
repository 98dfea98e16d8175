"""TEST INFRASTRUCTURE ONLY — g++ build of the device mathematics for CPU-side checks."""

from __future__ import annotations

import ctypes as C
import os
import subprocess

import numpy as np

_HERE = os.path.dirname(os.path.abspath(__file__))
_SO = os.path.join(_HERE, "_build", "libpnb_hostsim.so")
_LIB = None


def build(force=False):
    srcs = [os.path.join(_HERE, "trf_hostsim.cpp")] + [
        os.path.join(_HERE, "..", "..", "pyneapple_b200", "csrc", f)
        for f in ("pnb_trf_core.cuh", "pnb_dogbox_core.cuh", "pnb_lm_core.cuh", "pnb_models.cuh", "pnb_hd.cuh")
    ]
    newest = max(os.path.getmtime(s) for s in srcs)
    if force or not os.path.exists(_SO) or os.path.getmtime(_SO) < newest:
        os.makedirs(os.path.dirname(_SO), exist_ok=True)
        subprocess.check_call(
            ["g++", "-O1", "-fPIC", "-shared", "-std=c++17", "-ffp-contract=fast", "-o", _SO, srcs[0]]
        )
    return _SO


def lib():
    global _LIB
    if _LIB is None:
        _LIB = C.CDLL(build())
    return _LIB


def _p(a, t=C.c_double):
    return a.ctypes.data_as(C.POINTER(t))


def trf_fit(model_id, b, y, p0, lb, ub, frozen=None, t1_mode=0, tr=0.0, tm=0.0, ftol=1e-8,
            xtol=1e-8, gtol=1e-8, max_nfev=250, jac_mode=0, x_scale_jac=False, x_scale=None, method=0):
    b = np.ascontiguousarray(b, np.float64)
    y = np.ascontiguousarray(np.atleast_2d(y), np.float64)
    n_vox, nb = y.shape
    p0 = np.ascontiguousarray(p0, np.float64)
    lb = np.ascontiguousarray(lb, np.float64)
    ub = np.ascontiguousarray(ub, np.float64)
    na = p0.shape[1]
    fr = np.zeros(8, np.int32)
    if frozen is not None:
        fr[:na] = frozen
    n = int(na - fr.sum())
    params = np.empty((n_vox, na))
    cov = np.empty((n_vox, n, n))
    status = np.empty(n_vox, np.int32)
    nfev = np.empty(n_vox, np.int32)
    cost = np.empty(n_vox)
    xs = np.ones(8)
    if x_scale is not None:
        xs[:na] = x_scale
    rc = lib().pnbh_trf_fit(
        C.c_int(model_id), C.c_int(t1_mode), C.c_double(tr), C.c_double(tm), C.c_int(nb), _p(b),
        C.c_long(n_vox), _p(y), _p(p0), _p(lb), _p(ub), _p(fr, C.c_int), C.c_double(ftol),
        C.c_double(xtol), C.c_double(gtol), C.c_int(max_nfev), C.c_int(jac_mode),
        C.c_int(int(x_scale_jac)), _p(xs), _p(params), _p(cov), _p(status, C.c_int),
        _p(nfev, C.c_int), _p(cost), C.c_int(method),
    )
    if rc != 0:
        raise RuntimeError(f"hostsim: unsupported model {model_id}/{t1_mode}")
    return dict(params=params, cov=cov, status=status, nfev=nfev, cost=cost)


def trf_fit_extras(model_id, b, y, p0, lb, ub, ftol=1e-8, xtol=1e-8, gtol=1e-8, max_nfev=250, jac_mode=1, method=0,
                   weights=None, loss=0, f_scale=1.0, diff_step=None, absolute_sigma=False, extras_kernel=True):
    """The core with the curve_fit extras (``weights`` = 1 / sigma per b-value)."""
    b = np.ascontiguousarray(b, np.float64)
    y = np.ascontiguousarray(np.atleast_2d(y), np.float64)
    n_vox, nb = y.shape
    p0, lb, ub = (np.ascontiguousarray(a, np.float64) for a in (p0, lb, ub))
    na = p0.shape[1]
    params = np.empty((n_vox, na)); cov = np.empty((n_vox, na, na))
    status = np.empty(n_vox, np.int32); nfev = np.empty(n_vox, np.int32); cost = np.empty(n_vox)
    w = None if weights is None else np.ascontiguousarray(np.broadcast_to(np.asarray(weights, float), (nb,)))
    ds = None
    if diff_step is not None:
        ds = np.zeros(8); ds[:na] = np.broadcast_to(np.asarray(diff_step, float), (na,))
    rc = lib().pnbh_trf_fit_extras(
        C.c_int(model_id), C.c_int(nb), _p(b), C.c_long(n_vox), _p(y), _p(p0), _p(lb), _p(ub), C.c_double(ftol),
        C.c_double(xtol), C.c_double(gtol), C.c_int(max_nfev), C.c_int(jac_mode), C.c_int(method),
        None if w is None else _p(w), C.c_int(loss), C.c_double(f_scale), None if ds is None else _p(ds),
        C.c_int(int(absolute_sigma)), C.c_int(int(extras_kernel)), _p(params), _p(cov), _p(status, C.c_int),
        _p(nfev, C.c_int), _p(cost))
    if rc != 0:
        raise RuntimeError(f"hostsim: unsupported model {model_id}")
    return dict(params=params, cov=cov, status=status, nfev=nfev, cost=cost)


def exp(x):
    """``pnb_exp`` of pnb_hd.cuh (host build of the same code the kernel runs)."""
    x = np.ascontiguousarray(x, np.float64)
    out = np.empty_like(x)
    lib().pnbh_exp(C.c_long(x.size), _p(x), _p(out))
    return out
