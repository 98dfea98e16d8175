"""ORACLE (test infrastructure, not product code) — CPU restatement of Pyneapple's solver loops.

Only ``tests/``, ``__graft_entry__.smoke()`` and ``bench.py``'s CPU-baseline /
``--impl reference`` legs may import this module.  Nothing under
``pyneapple_b200/`` does.

What it restates (reference file:line, /root/reference/src/pyneapple):
  * ``solvers/curvefit.py:171-317``   per-voxel loop around
    ``scipy.optimize.curve_fit(method="trf", maxfev=max_iter, ftol=tol)``;
    analytic Jacobian only when some parameter is fixed (``:274-293``);
    any exception -> result = p0, NaN covariance, success=False (``:308-317``)
  * ``solvers/constrained_curvefit.py:125-305``  per-voxel
    ``scipy.optimize.minimize(method="SLSQP")`` on 0.5*||y-f(p)||^2 with box
    bounds and ``1 - sum(f_i) >= 0``; Gauss-Newton covariance
  * ``solvers/nnls_solver.py:61-210``  [basis; mu*R] stacking, zero-padded
    signal, per-voxel ``scipy.optimize.nnls(A, b, maxiter=max_iter)``;
    exception -> zeros, residual = ||b||
  * ``model_functions/multiexp.py:35-302`` signal equations,
    ``model_functions/nnls.py:17-85`` bins / basis / regulariser
  * joblib process pool (``curvefit.py:201-213``, ``nnls_solver.py:153-162``)

The arithmetic itself lives in a third-party dependency that is not vendored
in the reference: SciPy (``pyproject.toml:34-36`` pins ``scipy>=1.17.1``; this
image has 1.18.1) — ``scipy/optimize/_lsq/trf.py`` (trf_bounds),
``_lsq/common.py``, ``_numdiff.py``, ``_nnls.py`` -> ``_slsqplib.nnls``,
``_slsqp_py.py``.  SciPy is part of the image on the GPU box too, so this port
calls the very same SciPy routines the reference calls; a plain-C restatement
of those SciPy algorithms is in ``oracle/pnb_oracle.c``.

PINNING: ``oracle/make_golden.py`` runs the *real* reference
(``/root/reference/src``) and this port on the same seeded inputs in the
authoring container and stores the reference's outputs in ``tests/golden/``;
``tests/test_oracle_golden.py`` asserts this port reproduces them bit for bit
(same SciPy calls on the same doubles).  Parity is therefore pinned to
reference outputs, not to the reference's own (property-only) tests.
"""

from __future__ import annotations

import os
import warnings

import numpy as np
from scipy.optimize import curve_fit, minimize, nnls

# --------------------------------------------------------------------------
# models (model_functions/multiexp.py)
# --------------------------------------------------------------------------

PARAM_NAMES = {
    ("monoexp", "s0"): ["S0", "D"],
    ("biexp", "reduced"): ["f1", "D1", "D2"],
    ("biexp", "full"): ["f1", "D1", "f2", "D2"],
    ("biexp", "s0"): ["f1", "D1", "D2", "S0"],
    ("triexp", "reduced"): ["f1", "D1", "f2", "D2", "D3"],
    ("triexp", "full"): ["f1", "D1", "f2", "D2", "f3", "D3"],
    ("triexp", "s0"): ["f1", "D1", "f2", "D2", "D3", "S0"],
}


class Model:
    """Minimal stand-in for MonoExp/BiExp/TriExpModel (models/*.py)."""

    def __init__(self, kind, mode="reduced", t1=None, tr=None, tm=None, fixed=None):
        # kind: monoexp|biexp|triexp ; mode: reduced|full|s0 ; t1: None|"t1"|"steam"
        if kind == "monoexp":
            mode = "s0"
        self.kind, self.mode, self.t1, self.tr, self.tm = kind, mode, t1, tr, tm
        self.all_names = list(PARAM_NAMES[(kind, mode)]) + (["T1"] if t1 else [])
        self.fixed = dict(fixed or {})

    @property
    def param_names(self):
        return [n for n in self.all_names if n not in self.fixed]

    def forward(self, b, *p):
        k, m = self.kind, self.mode
        if k == "monoexp":
            s = p[0] * np.exp(-b * p[1])
        elif k == "biexp":
            if m == "s0":
                s = p[3] * (p[0] * np.exp(-b * p[1]) + (1 - p[0]) * np.exp(-b * p[2]))
            elif m == "reduced":
                s = p[0] * np.exp(-b * p[1]) + (1 - p[0]) * np.exp(-b * p[2])
            else:
                s = p[0] * np.exp(-b * p[1]) + p[2] * np.exp(-b * p[3])
        else:
            if m == "s0":
                s = p[5] * (
                    p[0] * np.exp(-b * p[1])
                    + p[2] * np.exp(-b * p[3])
                    + (1 - p[0] - p[2]) * np.exp(-b * p[4])
                )
            elif m == "reduced":
                s = (
                    p[0] * np.exp(-b * p[1])
                    + p[2] * np.exp(-b * p[3])
                    + (1 - p[0] - p[2]) * np.exp(-b * p[4])
                )
            else:
                s = (
                    p[0] * np.exp(-b * p[1])
                    + p[2] * np.exp(-b * p[3])
                    + p[4] * np.exp(-b * p[5])
                )
        if self.t1:
            t1 = p[len(self.all_names) - 1]
            s = s * (1 - np.exp(-self.tr / t1))
            if self.t1 == "steam":
                s = s * np.exp(-self.tm / t1)
        return s

    def jacobian(self, b, *p):
        k, m = self.kind, self.mode
        if k == "monoexp":
            e = np.exp(-b * p[1])
            cols = [e, -b * p[0] * e]
            base = p[0] * e
        elif k == "biexp":
            f1, d1 = p[0], p[1]
            e1 = np.exp(-b * d1)
            if m == "s0":
                e2 = np.exp(-b * p[2])
                s0 = p[3]
                cols = [s0 * (e1 - e2), -b * s0 * f1 * e1, -b * s0 * (1 - f1) * e2,
                        f1 * e1 + (1 - f1) * e2]
                base = s0 * (f1 * e1 + (1 - f1) * e2)
            elif m == "reduced":
                e2 = np.exp(-b * p[2])
                cols = [e1 - e2, -b * f1 * e1, -b * (1 - f1) * e2]
                base = f1 * e1 + (1 - f1) * e2
            else:
                f2 = p[2]
                e2 = np.exp(-b * p[3])
                cols = [e1, -b * f1 * e1, e2, -b * f2 * e2]
                base = f1 * e1 + f2 * e2
        else:
            f1, d1, f2, d2 = p[0], p[1], p[2], p[3]
            e1, e2 = np.exp(-b * d1), np.exp(-b * d2)
            if m == "full":
                f3 = p[4]
                e3 = np.exp(-b * p[5])
                cols = [e1, -b * f1 * e1, e2, -b * f2 * e2, e3, -b * f3 * e3]
                base = f1 * e1 + f2 * e2 + f3 * e3
            else:
                e3 = np.exp(-b * p[4])
                f3 = 1 - f1 - f2
                s0 = p[5] if m == "s0" else 1.0
                cols = [s0 * (e1 - e3), -b * s0 * f1 * e1, s0 * (e2 - e3),
                        -b * s0 * f2 * e2, -b * s0 * f3 * e3]
                shape = f1 * e1 + f2 * e2 + f3 * e3
                if m == "s0":
                    cols.append(shape)
                base = s0 * shape
        jac = np.column_stack(cols)
        if not self.t1:
            return jac
        t1 = p[len(self.all_names) - 1]
        e_tr = np.exp(-self.tr / t1)
        a = 1 - e_tr
        if self.t1 == "steam":
            e_tm = np.exp(-self.tm / t1)
            factor = a * e_tm
            d_t1 = base * e_tm / t1**2 * (-self.tr * e_tr + self.tm * a)
        else:
            factor = a
            d_t1 = base * (-e_tr * self.tr / t1**2)
        return np.column_stack((jac * factor, d_t1))

    # models/base.py:145-230
    def free_indices(self, fixed):
        return [i for i, n in enumerate(self.all_names) if n not in fixed]

    def inject(self, free, fixed):
        it = iter(free)
        return tuple(float(fixed[n]) if n in fixed else next(it) for n in self.all_names)


# --------------------------------------------------------------------------
# curve_fit path (solvers/curvefit.py:246-317)
# --------------------------------------------------------------------------


def _curvefit_one(model, xdata, y, p0, lb, ub, max_iter, tol, fixed, method, extra):
    if fixed:
        fwd = lambda x, *p, _f=fixed: model.forward(x, *model.inject(p, _f))
        idx = model.free_indices(fixed)
        jac = lambda x, *p, _f=fixed: model.jacobian(x, *model.inject(p, _f))[:, idx]
        if len(idx) < len(p0):
            p0, lb, ub = p0[idx], lb[idx], ub[idx]
    else:
        fwd, jac = model.forward, None
    n = len(p0)
    try:
        with warnings.catch_warnings():
            warnings.simplefilter("ignore")
            popt, pcov, info, msg, ier = curve_fit(
                f=fwd, xdata=xdata, ydata=y, p0=p0, bounds=(lb, ub), jac=jac,
                method=method, maxfev=max_iter, ftol=tol, full_output=True, **extra,
            )
        return popt, pcov, True, None, int(info["nfev"]), int(ier)
    except Exception as e:  # noqa: BLE001 - the reference catches everything
        return np.array(p0, float), np.full((n, n), np.nan), False, str(e), -1, 0


def _chunk_runner(fn, chunks, n_jobs):
    if n_jobs == 1:
        return [fn(c) for c in chunks]
    from joblib import Parallel, delayed

    return Parallel(n_jobs=n_jobs)(delayed(fn)(c) for c in chunks)


def curvefit_fit(model, xdata, ydata, p0, lb, ub, max_iter=250, tol=1e-8,
                 pixel_fixed=None, method="trf", n_jobs=1, per_voxel_tasks=False, **extra):
    """Batch fit.  ``p0, lb, ub``: ``(n_params, n_vox)`` over ``model.param_names``.

    Returns dict(params (n_free, n_vox), pcov (n_vox, n, n), success, messages,
    nfev, status).  ``n_jobs != 1`` uses a joblib (loky) process pool like the
    reference; ``per_voxel_tasks=True`` submits one task per voxel exactly as
    ``curvefit.py:201-213`` does, otherwise voxels are submitted in chunks
    (same arithmetic, less pickling) — the CPU baseline reports which it used.
    """
    xdata = np.asarray(xdata, float)
    ydata = np.atleast_2d(np.asarray(ydata, float))
    n_vox = ydata.shape[0]

    def one(i):
        fixed = dict(model.fixed) if model.fixed else None
        if pixel_fixed:
            fixed = {k: float(v[i]) for k, v in pixel_fixed.items()}
        return _curvefit_one(model, xdata, ydata[i], p0[:, i], lb[:, i], ub[:, i],
                             max_iter, tol, fixed, method, extra)

    if n_jobs == 1:
        res = [one(i) for i in range(n_vox)]
    else:
        size = 1 if per_voxel_tasks else max(1, min(256, n_vox // (4 * _n_workers(n_jobs)) or 1))
        chunks = [range(s, min(s + size, n_vox)) for s in range(0, n_vox, size)]
        parts = _chunk_runner(lambda c: [one(i) for i in c], chunks, n_jobs)
        res = [r for part in parts for r in part]
    return {
        "params": np.array([r[0] for r in res]).T,
        "pcov": np.array([r[1] for r in res]),
        "success": np.array([r[2] for r in res], bool),
        "messages": [r[3] for r in res],
        "nfev": np.array([r[4] for r in res]),
        "status": np.array([r[5] for r in res]),
    }


def _n_workers(n_jobs):
    return os.cpu_count() if n_jobs is None or n_jobs < 1 else n_jobs


# --------------------------------------------------------------------------
# SLSQP path (solvers/constrained_curvefit.py:125-305)
# --------------------------------------------------------------------------


def _constrained_one(model, xdata, y, p0, lb, ub, max_iter, tol, fixed, fraction_constraint):
    if fixed:
        fwd = lambda x, *p, _f=fixed: model.forward(x, *model.inject(p, _f))
        idx = model.free_indices(fixed)
        jac = lambda x, *p, _f=fixed: model.jacobian(x, *model.inject(p, _f))[:, idx]
        if len(idx) < len(p0):
            p0, lb, ub = p0[idx], lb[idx], ub[idx]
        free_names = [n for n in model.all_names if n not in fixed]
    else:
        fwd, jac = model.forward, model.jacobian
        free_names = list(model.param_names)
    frac_idx = [i for i, n in enumerate(free_names) if n.startswith("f")] if fraction_constraint else []

    def objective(p):
        r = y - fwd(xdata, *p)
        return 0.5 * np.dot(r, r)

    def gradient(p):
        r = y - fwd(xdata, *p)
        return -jac(xdata, *p).T @ r

    cons = []
    if fraction_constraint and len(frac_idx) >= 2:
        cons.append({"type": "ineq", "fun": lambda p, _i=frac_idx: 1.0 - sum(p[i] for i in _i)})
    n = len(p0)
    try:
        with warnings.catch_warnings():
            warnings.simplefilter("ignore")
            res = minimize(fun=objective, x0=p0, jac=gradient, method="SLSQP",
                           bounds=list(zip(lb, ub)), constraints=cons,
                           options={"maxiter": max_iter, "ftol": tol, "disp": False})
        popt = res.x
        try:
            J = jac(xdata, *popt)
            r = y - fwd(xdata, *popt)
            s_sq = np.dot(r, r) / max(len(y) - n, 1)
            pcov = s_sq * np.linalg.inv(J.T @ J)
        except (np.linalg.LinAlgError, ValueError):
            pcov = np.full((n, n), np.nan)
        return popt, pcov, bool(res.success), str(res.message), int(res.nit)
    except Exception as e:  # noqa: BLE001
        return np.array(p0, float), np.full((n, n), np.nan), False, str(e), -1


def constrained_fit(model, xdata, ydata, p0, lb, ub, max_iter=250, tol=1e-8,
                    pixel_fixed=None, fraction_constraint=True, n_jobs=1):
    xdata = np.asarray(xdata, float)
    ydata = np.atleast_2d(np.asarray(ydata, float))
    n_vox = ydata.shape[0]

    def one(i):
        fixed = dict(model.fixed) if model.fixed else None
        if pixel_fixed:
            fixed = {k: float(v[i]) for k, v in pixel_fixed.items()}
        return _constrained_one(model, xdata, ydata[i], p0[:, i], lb[:, i], ub[:, i],
                                max_iter, tol, fixed, fraction_constraint)

    if n_jobs == 1:
        res = [one(i) for i in range(n_vox)]
    else:
        size = max(1, min(256, n_vox // (4 * _n_workers(n_jobs)) or 1))
        chunks = [range(s, min(s + size, n_vox)) for s in range(0, n_vox, size)]
        parts = _chunk_runner(lambda c: [one(i) for i in c], chunks, n_jobs)
        res = [r for part in parts for r in part]
    return {
        "params": np.array([r[0] for r in res]).T,
        "pcov": np.array([r[1] for r in res]),
        "success": np.array([r[2] for r in res], bool),
        "messages": [r[3] for r in res],
        "nit": np.array([r[4] for r in res]),
    }


# --------------------------------------------------------------------------
# NNLS path (model_functions/nnls.py, solvers/nnls_solver.py)
# --------------------------------------------------------------------------


def nnls_bins(d_min, d_max, n_bins):
    return np.logspace(np.log10(d_min), np.log10(d_max), n_bins)


def nnls_basis(bvalues, d_values):
    return np.exp(-np.asarray(bvalues, float).reshape(-1, 1) * d_values.reshape(1, -1))


def regularization_matrix(n, order, mu):
    if order == 0:
        return np.zeros((n, n))
    if order == 1:
        return (np.diag(np.full(n, -1.0)) + np.diag(np.ones(n - 1), 1)) * mu
    if order == 2:
        return (np.diag(np.ones(n - 1), -1) + np.diag(np.full(n, -2.0)) + np.diag(np.ones(n - 1), 1)) * mu
    if order == 3:
        return (
            np.diag(np.ones(n - 2), -2) + np.diag(np.full(n - 1, 2.0), -1) + np.diag(np.full(n, -6.0))
            + np.diag(np.full(n - 1, 2.0), 1) + np.diag(np.ones(n - 2), 2)
        ) * mu
    raise NotImplementedError(f"Regularization order {order} not supported. Use 0-3.")


def _nnls_one(A, b, max_iter):
    try:
        x, r = nnls(A, b, maxiter=max_iter)
        return x, float(r), True
    except Exception:  # noqa: BLE001
        return np.zeros(A.shape[1]), float(np.linalg.norm(b)), False


def nnls_fit(xdata, signal, d_range, n_bins, reg_order=0, mu=0.02, max_iter=250, n_jobs=1):
    signal = np.atleast_2d(np.asarray(signal, float))
    bins = nnls_bins(d_range[0], d_range[1], n_bins)
    A = np.concatenate([nnls_basis(xdata, bins), regularization_matrix(n_bins, reg_order, mu)], axis=0)
    ext = np.concatenate((signal, np.zeros((signal.shape[0], n_bins))), axis=1)
    n_vox = signal.shape[0]
    if n_jobs == 1:
        res = [_nnls_one(A, ext[i], max_iter) for i in range(n_vox)]
    else:
        size = max(1, min(256, n_vox // (4 * _n_workers(n_jobs)) or 1))
        chunks = [range(s, min(s + size, n_vox)) for s in range(0, n_vox, size)]
        parts = _chunk_runner(lambda c: [_nnls_one(A, ext[i], max_iter) for i in c], chunks, n_jobs)
        res = [r for part in parts for r in part]
    return {
        "coefficients": np.array([r[0] for r in res]),
        "residual": np.array([r[1] for r in res]),
        "success": np.array([r[2] for r in res], bool),
    }


# --------------------------------------------------------------------------- NNLS spectrum post-processing
def spectrum_peaks(spectra, bins, height=0.1, regularized=False, cutoffs=None, max_peaks=8):
    """utility/spectrum.py:50-103 (find_spectrum_peaks -> calculate_peak_area :13-47) and :139-206
    (apply_cutoffs -> geometric_mean_peak :106-136) per row of ``spectra``, with the same
    scipy.signal.find_peaks / peak_widths calls, packed like ``find_spectrum_peaks_batch``."""
    from scipy import signal as scipy_signal

    spectra = np.atleast_2d(np.asarray(spectra, dtype=np.float64))
    bins = np.asarray(bins, dtype=np.float64)
    n = spectra.shape[0]
    K = 0 if cutoffs is None else len(cutoffs)
    out = dict(n_peaks=np.zeros(n, np.int32), peak_index=np.full((n, max_peaks), -1, np.int32),
               d_values=np.full((n, max_peaks), np.nan), f_values=np.full((n, max_peaks), np.nan),
               d_cut=np.full((n, K), np.nan), f_cut=np.full((n, K), np.nan))
    for v in range(n):
        x = spectra[v]
        idx, props = scipy_signal.find_peaks(x, height=height)
        k = len(idx)
        out["n_peaks"][v] = k
        if k:
            raw = props["peak_heights"]
            if regularized:
                fw = scipy_signal.peak_widths(x, idx, rel_height=0.5)[0]
                f = np.array([float(h * w / (2 * np.sqrt(2 * np.log(2))) * np.sqrt(2 * np.pi)) for h, w in zip(raw, fw)])
            else:
                f = raw.copy()
            total = np.sum(f)
            if total > 0:
                f = f / total
            d = bins[idx]
        else:
            d, f = np.array([]), np.array([])
        m = min(k, max_peaks)
        out["peak_index"][v, :m], out["d_values"][v, :m], out["f_values"][v, :m] = idx[:m], d[:m], f[:m]
        if K:
            new_d, new_f = [], []
            for lo, hi in cutoffs:
                mask = (d >= lo) & (d <= hi)
                _d, _f = d[mask], f[mask]
                if len(_d) == 0:
                    new_d.append(np.nan); new_f.append(np.nan)
                elif len(_d) == 1:
                    new_d.append(float(_d[0])); new_f.append(float(_f[0]))
                else:
                    new_d.append(float(np.log10(np.prod(_d ** (_f / np.sum(_f)))))); new_f.append(float(np.sum(_f)))
            new_d, new_f = np.array(new_d), np.array(new_f)
            total = np.nansum(new_f)
            if total > 0:
                new_f = new_f / total
            out["d_cut"][v], out["f_cut"][v] = new_d, new_f
    return out
