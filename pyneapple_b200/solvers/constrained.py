"""B200 drop-in for ``ConstrainedCurveFitSolver`` (solvers/constrained_curvefit.py:22-305).

The reference minimises ``0.5 * ||y - f(p)||^2`` per voxel with SciPy's SLSQP
under box bounds and one linear inequality ``sum(fractions) <= 1``.  SLSQP's
BFGS / line-search trajectory stops, at ``ftol = 1e-8``, a median 9e-4
(relative) away from the constrained minimiser (hard part 2), so
"parity" with it cannot mean reproducing its iterates.  The contract here is:

  (i)   residual norm <= the reference's, voxel for voxel;
  (ii)  bounds and ``sum(f) <= 1`` hold exactly (to rounding);
  (iii) parameters are the tightly converged minimiser of the same problem.

Method (active set over the single inequality, both phases are the CUDA TRF
kernel): phase 1 solves the box-bounded problem; voxels whose solution
violates ``f1 + f2 <= 1`` have the constraint active at their minimiser, and
on that face the reduced tri-exponential model with ``f3 = 1 - f1 - f2 = 0``
*is* the reduced bi-exponential model in ``(f1, D1, D2[, S0][, T1])`` with
``f2 = 1 - f1``, so phase 2 fits exactly that, with the bounds of ``f2``
folded into those of ``f1``.  ``D3`` is unidentifiable on the face (its
Jacobian column is zero) and keeps its phase-1 value.
"""

from __future__ import annotations

import logging
from typing import Any

import numpy as np

from .. import engine
from ..models import (MODEL_BI_REDUCED, MODEL_BI_S0, MODEL_TRI_REDUCED, MODEL_TRI_S0, ModelDesc,
                      _all_names)
from .base import PixelResults
from .curvefit import CurveFitSolver

log = logging.getLogger("pyneapple_b200")

_TIGHT = 1e-13


class ConstrainedCurveFitSolver(CurveFitSolver):
    """Constrained NLLS on the GPU (see module docstring for the parity contract)."""

    def __init__(self, model: Any, max_iter: int, tol: float, p0: dict, bounds: dict,
                 fraction_constraint: bool = True, verbose: bool = False, method: str = "SLSQP",
                 multi_threading: bool = False, use_jacobian: bool = True, **solver_kwargs):
        if fraction_constraint and not getattr(model, "fit_reduced", False):
            raise ValueError(
                "fraction_constraint=True requires fit_reduced=True. In reduced mode the signal is "
                "normalised to S0=1 before fitting, so the hard constraint sum(f_i) <= 1 is "
                "physically meaningful. Use fit_reduced=True (or fit_s0=True) with the constrained solver."
            )
        # constrained_curvefit.py:217-229 does not forward extra kwargs to minimize(): accept and ignore
        extras = {k: solver_kwargs.pop(k) for k in list(solver_kwargs)
                  if k not in ("n_pools", "device", "chunk_vox", "want_cov", "pinned_outputs")}
        super().__init__(model=model, max_iter=max_iter, tol=tol, p0=p0, bounds=bounds, verbose=verbose,
                         method="trf", multi_threading=multi_threading, use_jacobian=use_jacobian,
                         jac="analytic", **solver_kwargs)
        self.ignored_kwargs = extras
        self.method = "SLSQP"
        self.fraction_constraint = fraction_constraint
        if fraction_constraint:
            self._fraction_names = [n for n in self.model.param_names if n.startswith("f")]
            if len(self._fraction_names) < 2:
                raise ValueError(
                    "fraction_constraint=True requires at least 2 fraction parameters. "
                    f"Found: {self._fraction_names}"
                )
            self._fraction_indices = [self.model.param_names.index(n) for n in self._fraction_names]
            if self._desc.model_id not in (MODEL_TRI_REDUCED, MODEL_TRI_S0):
                raise NotImplementedError("fraction constraint is implemented for the reduced / S0 tri-exponential model")
        else:
            self._fraction_names = []
            self._fraction_indices = []
        self.n_active_ = 0

    def fit(self, xdata, ydata, p0=None, bounds=None, pixel_fixed_params=None, **fit_kwargs):
        from .. import validation as V

        self._reset_state()
        xdata = np.asarray(xdata)
        on_device = engine._is_torch_cuda(ydata)
        if on_device:
            ydata = ydata.cpu().numpy()  # orchestration of the two phases is done on host arrays
        ydata = np.asarray(ydata)
        V.validate_data_shapes(xdata, ydata)
        n_pixels = ydata.shape[0] if ydata.ndim > 1 else 1
        if ydata.ndim == 1:
            ydata = ydata[None, :]
        p0_m, lb_m, ub_m = self._validate_p0_and_bounds(p0, bounds, n_pixels)
        # phase 1: box-bounded problem, tight tolerances, analytic Jacobian
        saved = (self.tol, self.max_iter, dict(self.solver_kwargs))
        self.tol = min(self.tol, _TIGHT)
        self.solver_kwargs.update(xtol=_TIGHT, gtol=_TIGHT)
        self.max_iter = max(4 * saved[1], 1000)
        try:
            res, free_names = self._solve(xdata, ydata, p0_m, lb_m, ub_m, pixel_fixed_params, n_pixels)
        finally:
            self.tol, self.max_iter, self.solver_kwargs = saved
        desc = self._desc
        all_names = list(desc.all_names)
        pix_fixed = set(pixel_fixed_params or {}) & set(all_names)
        fracs = [n for n in free_names if n.startswith("f")] if self.fraction_constraint else []
        self.n_active_ = 0
        if len(fracs) >= 2:
            i1, i2 = all_names.index("f1"), all_names.index("f2")
            par = res["params"]
            viol = (par[i1] + par[i2] > 1.0) & (res["status"] > 0)
            idx = np.nonzero(viol)[0]
            self.n_active_ = int(idx.size)
            if idx.size:
                self._solve_on_face(xdata, ydata, idx, res, lb_m, ub_m, all_names, pixel_fixed_params, pix_fixed)
        # iteration limit: the reference reports result.x with success=False (not p0)
        self._store(res, free_names, n_pixels)
        self.pixel_results_.n_iterations = res["nfev"]
        return self

    def _solve_on_face(self, xdata, ydata, idx, res, lb_m, ub_m, all_names, pixel_fixed_params, pix_fixed):
        desc = self._desc
        model_names = list(self.model.param_names)
        face_id = MODEL_BI_S0 if desc.model_id == MODEL_TRI_S0 else MODEL_BI_REDUCED
        face_names = _all_names(face_id, desc.t1_mode)
        face_desc = ModelDesc(model_id=face_id, t1_mode=desc.t1_mode, repetition_time=desc.repetition_time,
                              mixing_time=desc.mixing_time, all_names=tuple(face_names), fixed={})
        nv = idx.size
        par = res["params"]

        def bound_of(arr, name):
            if name in desc.fixed or name in pix_fixed:
                return None
            row = arr[model_names.index(name)]
            return row[idx] if np.ndim(row) else np.full(nv, float(row))

        src = {"f1": "f1", "D1": "D1", "D2": "D2", "S0": "S0", "T1": "T1"}
        P0 = np.empty((len(face_names), nv))
        LB = np.full((len(face_names), nv), -np.inf)
        UB = np.full((len(face_names), nv), np.inf)
        frozen = 0
        for j, nm in enumerate(face_names):
            full_name = src[nm]
            P0[j] = par[all_names.index(full_name)][idx]
            lo, hi = bound_of(lb_m, full_name), bound_of(ub_m, full_name)
            if lo is None:
                frozen |= 1 << j
            else:
                LB[j], UB[j] = lo, hi
        # fold the bounds of f2 = 1 - f1 into those of f1 and start from the projection onto the face
        j1 = face_names.index("f1")
        lo2, hi2 = bound_of(lb_m, "f2"), bound_of(ub_m, "f2")
        LB[j1] = np.maximum(LB[j1], 1.0 - hi2)
        UB[j1] = np.minimum(UB[j1], 1.0 - lo2)
        f1, f2 = par[all_names.index("f1")][idx], par[all_names.index("f2")][idx]
        P0[j1] = np.clip(f1 - 0.5 * (f1 + f2 - 1.0), LB[j1], UB[j1])
        r2 = engine.trf_fit(face_desc, xdata, np.ascontiguousarray(ydata[idx]), P0, LB, UB, frozen,
                            max_nfev=max(4 * self.max_iter, 1000), ftol=_TIGHT, xtol=_TIGHT, gtol=_TIGHT,
                            jac_mode=engine.JAC_ANALYTIC, want_cov=False, device=self.device)
        ok = r2["status"] > 0
        for j, nm in enumerate(face_names):
            row = all_names.index(src[nm])
            par[row][idx[ok]] = r2["params"][j][ok]
        par[all_names.index("f2")][idx[ok]] = 1.0 - r2["params"][j1][ok]
        res["nfev"][idx] += r2["nfev"]
        res["status"][idx[~ok]] = 0
        res["cost"][idx[ok]] = r2["cost"][ok]
        if res.get("r2") is not None:
            res["r2"][idx[ok]] = r2["r2"][ok]  # same signal, same prediction on the face
        if res["cov"] is not None:
            # J^T J is singular on the face (dS/dD3 = -b f3 e3 = 0): np.linalg.inv raises ->
            # NaN covariance (constrained_curvefit.py:300-305)
            res["cov"][idx] = np.nan
