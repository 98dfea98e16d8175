"""ORACLE (test infrastructure) — ctypes front-end to oracle/_build/libpnb_oracle.so.

See the header of ``oracle/pnb_oracle.c`` for what is restated and how it is
pinned.  Product code never imports this.
"""

from __future__ import annotations

import ctypes as C
import os
import subprocess

import numpy as np

_HERE = os.path.dirname(os.path.abspath(__file__))
_LIB = None

MODEL_IDS = {
    ("monoexp", "s0"): 0, ("biexp", "reduced"): 1, ("biexp", "full"): 2, ("biexp", "s0"): 3,
    ("triexp", "reduced"): 4, ("triexp", "full"): 5, ("triexp", "s0"): 6,
}


def build(force: bool = False) -> str:
    so = os.path.join(_HERE, "_build", "libpnb_oracle.so")
    src = os.path.join(_HERE, "pnb_oracle.c")
    if force or not os.path.exists(so) or os.path.getmtime(so) < os.path.getmtime(src):
        subprocess.check_call(["make", "-s", "-C", _HERE])
    return so


def lib():
    global _LIB
    if _LIB is None:
        _LIB = C.CDLL(build())
    return _LIB


def _p(a, t=C.c_double):
    return a.ctypes.data_as(C.POINTER(t))


def trf_fit(model_id, b, y, p0, lb, ub, frozen=None, t1_mode=0, tr=0.0, tm=0.0, ftol=1e-8,
            xtol=1e-8, gtol=1e-8, max_nfev=250, jac_mode=1, x_scale_jac=False, x_scale=None, method="trf"):
    """p0/lb/ub: (n_vox, n_all) row-major over the FULL parameter list."""
    b = np.ascontiguousarray(b, np.float64)
    y = np.ascontiguousarray(np.atleast_2d(y), np.float64)
    n_vox, nb = y.shape
    p0 = np.ascontiguousarray(p0, np.float64)
    lb = np.ascontiguousarray(lb, np.float64)
    ub = np.ascontiguousarray(ub, np.float64)
    na = p0.shape[1]
    fr = np.zeros(na, np.int32) if frozen is None else np.ascontiguousarray(frozen, np.int32)
    n = int(na - fr.sum())
    params = np.empty((n_vox, na))
    cov = np.empty((n_vox, n, n))
    status = np.empty(n_vox, np.int32)
    nfev = np.empty(n_vox, np.int32)
    cost = np.empty(n_vox)
    xs = None if x_scale is None else np.ascontiguousarray(x_scale, np.float64)
    rc = lib().pnbo_lsq_fit(
        C.c_int({"trf": 0, "dogbox": 1}[method]), C.c_int(model_id), C.c_int(t1_mode), C.c_double(tr), C.c_double(tm), C.c_int(nb), _p(b),
        C.c_long(n_vox), _p(y), _p(p0), _p(lb), _p(ub), _p(fr, C.c_int), C.c_double(ftol),
        C.c_double(xtol), C.c_double(gtol), C.c_int(max_nfev), C.c_int(jac_mode),
        C.c_int(int(x_scale_jac)), None if xs is None else _p(xs), _p(params), _p(cov),
        _p(status, C.c_int), _p(nfev, C.c_int), _p(cost),
    )
    if rc != 0:
        raise RuntimeError(f"pnbo_trf_fit failed: {rc}")
    return dict(params=params, cov=cov, status=status, nfev=nfev, cost=cost)


def nnls(A, B, maxiter):
    A = np.ascontiguousarray(A, np.float64)
    B = np.ascontiguousarray(np.atleast_2d(B), np.float64)
    m, n = A.shape
    n_vox = B.shape[0]
    X = np.empty((n_vox, n))
    rnorm = np.empty(n_vox)
    status = np.empty(n_vox, np.int32)
    iters = np.empty(n_vox, np.int32)
    lib().pnbo_nnls(C.c_int(m), C.c_int(n), _p(A), C.c_long(n_vox), _p(B), C.c_int(maxiter), _p(X),
                    _p(rnorm), _p(status, C.c_int), _p(iters, C.c_int))
    return dict(x=X, rnorm=rnorm, status=status, iters=iters)
