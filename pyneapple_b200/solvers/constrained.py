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
from ..validation import bounds_vectors as V_bounds
from .curvefit import CurveFitSolver

log = logging.getLogger("pyneapple_b200")

# ftol = xtol = gtol of the three phases.  Measured on the 4 096-voxel SLSQP golden against a 1e-15 solution
# (TRF stops on the first criterion met): 1e-13 leaves max 1e-5 / median 2e-8 relative in the parameters at 13.8
# evaluations per voxel, 1e-11 max 9e-5 at 11.9 — too close to the 1e-4 of the contract to buy 14 % of speed with
_TIGHT = 1e-13
# tight tolerances mean ~14 evaluations per voxel, spread widely: converged lanes of the TRF kernel wait longer for
# their warp before the once-per-voxel code runs (pnb_trf_problem.finish_wait; 48.9 -> 45.7 ms per 4.19 M voxel slab)
_FINISH_WAIT = 6


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
        self.n_released_ = 0

    def fit(self, xdata, ydata, p0=None, bounds=None, pixel_fixed_params=None, **fit_kwargs):
        """Both phases run on the GPU (``fit_device``); a numpy ``ydata`` is uploaded once (sharded over
        the solver's devices when it has several), results come back as numpy arrays."""
        from concurrent.futures import ThreadPoolExecutor

        from .. import _lib
        from .. import validation as V

        self._reset_state()
        _lib.require_device()
        import torch

        xdata = np.asarray(xdata)
        on_device = engine._is_torch_cuda(ydata)
        if not on_device:
            ydata = np.asarray(ydata)
        V.validate_data_shapes(xdata, ydata)
        n_pixels = ydata.shape[0] if ydata.ndim > 1 else 1
        if ydata.ndim == 1:
            ydata = ydata[None, :]
        p0_m, lb_m, ub_m = self._p0_and_bounds(p0, bounds, n_pixels)
        devices = [ydata.device.index] if on_device else _lib.resolve_devices(self.device)
        ranges = _lib.shard_ranges(n_pixels, len(devices)) if len(devices) > 1 else [(0, n_pixels)]

        def part(arr, a, z):  # per-voxel rows follow the voxels, vectors are shared
            if arr is None or np.ndim(arr) < 2:
                return arr
            return arr[:, a:z]

        # page-locked result arrays from the second fit of a shape on (CurveFitSolver._pinned_out): every GPU
        # downloads straight into its columns; a fresh numpy array per result costs its first-touch page faults
        n_all = len(self._desc.all_names)
        n_free_est = n_all - len(set(self._desc.fixed) | (set(pixel_fixed_params or {}) & set(self._desc.all_names)))
        pinned = None if on_device else self._pinned_out(n_all, n_free_est, n_pixels, ydata)

        not_pinned = set()   # result keys whose shape did not match the cached block (copied the plain way)

        def download(key, v, a, z):
            if pinned is None or key not in pinned:
                return engine.to_host(v)
            dst = pinned[key]
            if key == "params" and v.ndim == 2 and dst.shape[0] == v.shape[0] and v.shape[1] == z - a:
                for r in range(v.shape[0]):          # (n_all, n): row by row into the columns a..z
                    engine.to_host_into(v[r], dst[r, a:z])
                return dst[:, a:z]
            if key != "params" and tuple(dst[a:z].shape) == tuple(v.shape) and dst.itemsize == v.element_size():
                engine.to_host_into(v, dst[a:z])
                return dst[a:z]
            not_pinned.add(key)
            return engine.to_host(v)

        def run(k):
            (a, z), d = ranges[k], devices[k]
            if z <= a:
                return None
            with torch.cuda.device(d):
                y = ydata[a:z] if on_device else engine.to_device(np.ascontiguousarray(ydata[a:z], np.float64),
                                                                  torch.device("cuda", d))
                pf = None
                if pixel_fixed_params:
                    pf = {k_: torch.as_tensor(np.ascontiguousarray(np.asarray(v, float)[a:z])).to(y.device)
                          for k_, v in pixel_fixed_params.items()}
                up = lambda arr: arr if np.ndim(arr) < 2 else engine.to_device(np.ascontiguousarray(arr), y.device)  # noqa: E731
                res = self.fit_device(xdata, y, p0=up(part(p0_m, a, z)),
                                      bounds=(up(part(lb_m, a, z)), up(part(ub_m, a, z))), pixel_fixed_params=pf)
                n_failed = int((res["status"] <= 0).sum().item())
                lazy_cov = res.pop("cov") if (self.want_cov is True or self.want_cov == "lazy") else None
                host = {k_: (download(k_, v, a, z) if hasattr(v, "cpu") else v) for k_, v in res.items()}
                host["cov_dev"] = lazy_cov
                host["n_failed"] = n_failed
                return host

        if len(ranges) == 1:
            parts = [run(0)]
        else:
            with ThreadPoolExecutor(len(ranges)) as pool:
                parts = list(pool.map(run, range(len(ranges))))
        live = [(r, p) for r, p in zip(ranges, parts) if p is not None]
        first = live[0][1]
        res = {"free_names": first["free_names"], "free_rows": first["free_rows"]}
        for key in ("params", "status", "nfev", "njev", "cost", "r2"):
            vals = [p[key] for _, p in live]
            if pinned is not None and key in pinned and key not in not_pinned:
                res[key] = pinned[key]          # the parts are views of it
            else:
                res[key] = vals[0] if len(vals) == 1 else np.concatenate(vals, axis=-1)
        n_free = len(first["free_rows"])
        if first["cov_dev"] is not None:
            from .._lazy import LazyArray

            res["cov"] = LazyArray((n_pixels, n_free, n_free), [(a, z, p["cov_dev"]) for (a, z), p in live])
        elif first.get("cov") is not None:
            res["cov"] = (pinned["cov"] if pinned is not None and "cov" in pinned and "cov" not in not_pinned
                          else np.concatenate([p["cov"] for _, p in live], axis=0))
        else:
            res["cov"] = None
        res["n_failed"] = int(sum(p["n_failed"] for _, p in live))
        self.n_active_ = int(sum(p["n_active"] for _, p in live))      # box-bounded minimiser violated the constraint
        self.n_released_ = int(sum(p["n_released"] for _, p in live))  # ... and left the face again in phase 3
        self._free_rows = res["free_rows"]
        self._store(res, res["free_names"], n_pixels)
        # the reference reports SLSQP's iteration count (result.nit); the closest quantity here is the
        # number of residual evaluations of the two TRF phases
        self.pixel_results_.n_iterations = res["nfev"]
        return self

    def fit_device(self, xdata, y_dev, p0=None, bounds=None, pixel_fixed_params=None, want_cov=None):
        """Device-resident constrained fit: phase 1 (box-bounded TRF, tight tolerances), violation mask,
        compaction, re-fit on the face ``f1 + f2 = 1`` and scatter, all on the GPU of ``y_dev`` with one
        host synchronisation (the number of voxels on the face sizes the second launch).  Arguments and
        result as :meth:`CurveFitSolver.fit_device`, plus ``n_active``.

        A voxel that exhausts ``max(4 max_iter, 1000)`` evaluations in either phase is reported like a
        failed ``curve_fit``: ``status 0``, parameters = its start values.  (SLSQP would report its last
        iterate with ``success=False``; the two iteration limits are not comparable.)"""
        import torch

        saved = (self.tol, self.max_iter, dict(self.solver_kwargs), self.jac)
        self.tol = min(self.tol, _TIGHT)
        self.solver_kwargs.update(xtol=_TIGHT, gtol=_TIGHT)
        self.max_iter = max(4 * saved[1], 1000)
        self.jac = "analytic"
        self._finish_wait = _FINISH_WAIT
        try:
            res = super().fit_device(xdata, y_dev, p0=p0, bounds=bounds, pixel_fixed_params=pixel_fixed_params,
                                     want_cov=want_cov)
        finally:
            self.tol, self.max_iter, self.solver_kwargs, self.jac = saved
            self._finish_wait = 0
        res["n_active"] = 0
        res["n_released"] = 0
        free_names = res["free_names"]
        fracs = [n for n in free_names if n.startswith("f")] if self.fraction_constraint else []
        if len(fracs) < 2:
            return res
        desc = self._desc
        all_names = list(desc.all_names)
        model_names = list(self.model.param_names)
        par = res["params"]
        i1, i2 = all_names.index("f1"), all_names.index("f2")
        viol = (par[i1] + par[i2] > 1.0) & (res["status"] > 0)
        idx = viol.nonzero().squeeze(1)
        nv = int(idx.numel())  # the one host synchronisation
        res["n_active"] = nv
        if nv == 0:
            return res
        dev = y_dev.device
        pix_fixed = set(pixel_fixed_params or {}) & set(all_names)
        face_id = MODEL_BI_S0 if desc.model_id == MODEL_TRI_S0 else MODEL_BI_REDUCED
        face_names = _all_names(face_id, desc.t1_mode)
        face_desc = ModelDesc(model_id=face_id, t1_mode=desc.t1_mode, repetition_time=desc.repetition_time,
                              mixing_time=desc.mixing_time, all_names=tuple(face_names), fixed={})
        lb_def, ub_def = V_bounds(self.bounds, model_names)
        lb_src = bounds[0] if bounds is not None else lb_def
        ub_src = bounds[1] if bounds is not None else ub_def

        def bound_of(src, name):
            """Bound row of a model parameter on the face voxels: tensor (nv,) or float; None if fixed."""
            if name in desc.fixed or name in pix_fixed:
                return None
            row = src[model_names.index(name)]
            if isinstance(row, torch.Tensor) and row.ndim:
                return row.index_select(0, idx)
            return float(row)

        def as_row(v):
            return v if isinstance(v, torch.Tensor) else torch.full((nv,), float(v), dtype=torch.float64, device=dev)

        P0, LB, UB = [], [], []
        frozen = 0
        for j, nm in enumerate(face_names):
            P0.append(par[all_names.index(nm)].index_select(0, idx))
            lo, hi = bound_of(lb_src, nm), bound_of(ub_src, nm)
            if lo is None:
                frozen |= 1 << j
                lo, hi = -np.inf, np.inf
            LB.append(as_row(lo))
            UB.append(as_row(hi))
        # fold the bounds of f2 = 1 - f1 into those of f1 and start from the projection onto the face
        j1 = face_names.index("f1")
        lo2, hi2 = as_row(bound_of(lb_src, "f2")), as_row(bound_of(ub_src, "f2"))
        LB[j1] = torch.maximum(LB[j1], 1.0 - hi2)
        UB[j1] = torch.minimum(UB[j1], 1.0 - lo2)
        f1, f2 = P0[j1], par[i2].index_select(0, idx)
        P0[j1] = torch.minimum(torch.maximum(f1 - 0.5 * (f1 + f2 - 1.0), LB[j1]), UB[j1])
        r2 = engine.trf_fit(face_desc, np.asarray(xdata, float), y_dev.index_select(0, idx),
                            torch.stack(P0), torch.stack(LB), torch.stack(UB), frozen,
                            max_nfev=max(4 * self.max_iter, 1000), ftol=_TIGHT, xtol=_TIGHT, gtol=_TIGHT,
                            jac_mode=engine.JAC_ANALYTIC, finish_wait=_FINISH_WAIT, want_cov=False)
        ok = r2["status"] > 0
        # a face fit that fails (iteration limit) reports its start point — the projection of the phase-1
        # answer onto the face — so the voxel is returned feasible, with success = False, instead of keeping
        # the violating box-bounded answer
        for j, nm in enumerate(face_names):
            par[all_names.index(nm)].index_copy_(0, idx, r2["params"][j])
        par[i2].index_copy_(0, idx, 1.0 - r2["params"][j1])
        res["nfev"].index_add_(0, idx, r2["nfev"])
        res["status"].index_copy_(0, idx, torch.where(ok, res["status"].index_select(0, idx),
                                                      torch.zeros_like(r2["status"])))
        res["cost"].index_copy_(0, idx, torch.where(ok, r2["cost"], res["cost"].index_select(0, idx)))
        res["r2"].index_copy_(0, idx, torch.where(ok, r2["r2"], res["r2"].index_select(0, idx)))
        if res["cov"] is not None:
            # J^T J is singular on the face (dS/dD3 = -b f3 e3 = 0): np.linalg.inv raises ->
            # NaN covariance (constrained_curvefit.py:300-305)
            res["cov"].index_fill_(0, idx, float("nan"))
        res["n_released"] = self._release_from_face(xdata, y_dev, res, idx, face_desc, face_names, r2, lb_src, ub_src,
                                                    bound_of, pixel_fixed_params, bounds, p0)
        return res

    def _release_from_face(self, xdata, y_dev, res, idx, face_desc, face_names, r2, lb_src, ub_src, bound_of,
                           pixel_fixed_params, bounds, p0):
        """Phase 3: is the face point a constrained minimiser?  On the face the third component has no
        weight, so ``D3`` is free — and the Kuhn-Tucker multiplier of ``f1 + f2 <= 1`` depends on it:
        ``dF/df1 = r . A C (e1 - e3(D3))``.  The point is optimal only if no ``D3`` within its bounds
        offers a descent direction into the interior.  Voxels where one does (the box-bounded phase had
        ended in a different basin of this non-convex problem) are re-fitted with the full model from
        the face point with the most promising ``D3``; the result is kept when it is feasible and better.
        Returns the number of voxels released."""
        import torch

        desc = self._desc
        all_names = list(desc.all_names)
        if "D3" in desc.fixed or "D3" in (pixel_fixed_params or {}) or int(idx.numel()) == 0:
            return 0
        dev = y_dev.device
        nv = int(idx.numel())
        b = torch.as_tensor(np.asarray(xdata, float), device=dev)
        y_face = y_dev.index_select(0, idx)
        resid = engine.predict_device(face_desc, xdata, r2["params"]) - y_face      # (nv, n_b)
        d1 = r2["params"][face_names.index("D1")]
        re1 = (resid * torch.exp(-b[None, :] * d1[:, None])).sum(dim=1)             # r . e1
        lo3, hi3 = bound_of(lb_src, "D3"), bound_of(ub_src, "D3")
        K = 16
        t = torch.linspace(0.0, 1.0, K, dtype=torch.float64, device=dev)

        def as_col(v):
            return v[:, None] if isinstance(v, torch.Tensor) else torch.full((1, 1), float(v), dtype=torch.float64, device=dev)

        lo, hi = as_col(lo3), as_col(hi3)
        grid = torch.where((lo > 0) & torch.isfinite(hi), lo * (hi / lo.clamp_min(1e-300)) ** t[None, :],
                           lo + (hi - lo) * t[None, :])                                   # (nv | 1, K)
        if grid.shape[0] == 1:
            re3 = resid @ torch.exp(-b[:, None] * grid[0][None, :])                       # (nv, K): one small GEMM
        else:
            re3 = torch.stack([(resid * torch.exp(-b[None, :] * grid[:, k:k + 1])).sum(dim=1) for k in range(K)], dim=1)
        gain, kbest = (re1[:, None] - re3).max(dim=1)          # > 0: moving into the interior lowers the cost
        scale = resid.norm(dim=1) * float(np.sqrt(b.shape[0]))
        pick = (gain > 1e-9 * scale) & (r2["status"] > 0)
        sub = pick.nonzero().squeeze(1)
        n_rel = int(sub.numel())
        if n_rel == 0:
            return 0
        vox = idx.index_select(0, sub)
        par = res["params"]
        d3_start = (grid.expand(nv, K) if grid.shape[0] == 1 else grid).gather(1, kbest[:, None]).squeeze(1).index_select(0, sub)
        P0 = torch.stack([par[j].index_select(0, vox) for j in range(len(all_names))])
        P0[all_names.index("D3")] = d3_start

        def rows_of(src, default):
            if src is None:
                return default
            rows = []
            for n_ in all_names:
                if n_ in self.model.param_names:
                    v = src[list(self.model.param_names).index(n_)]
                    rows.append(v.index_select(0, vox) if isinstance(v, torch.Tensor) and v.ndim else
                                torch.full((n_rel,), float(v), dtype=torch.float64, device=dev))
                else:
                    rows.append(torch.full((n_rel,), float(default), dtype=torch.float64, device=dev))
            return torch.stack(rows)

        model_names = list(self.model.param_names)
        from .. import validation as V

        lb_def, ub_def = V.bounds_vectors(self.bounds, model_names)
        LB = rows_of(bounds[0] if bounds is not None else lb_def, -np.inf)
        UB = rows_of(bounds[1] if bounds is not None else ub_def, np.inf)
        fixed_names = set(desc.fixed) | (set(pixel_fixed_params or {}) & set(all_names))
        frozen = engine.frozen_mask(desc, fixed_names)
        r3 = engine.trf_fit(desc, np.asarray(xdata, float), y_dev.index_select(0, vox), P0, LB, UB, frozen,
                            max_nfev=max(4 * self.max_iter, 1000), ftol=_TIGHT, xtol=_TIGHT, gtol=_TIGHT,
                            jac_mode=engine.JAC_ANALYTIC, finish_wait=_FINISH_WAIT, want_cov=res["cov"] is not None)
        i1, i2 = all_names.index("f1"), all_names.index("f2")
        better = ((r3["status"] > 0) & (r3["params"][i1] + r3["params"][i2] <= 1.0)
                  & (r3["cost"] < res["cost"].index_select(0, vox)))
        for j in range(len(all_names)):
            row = par[j]
            row.index_copy_(0, vox, torch.where(better, r3["params"][j], row.index_select(0, vox)))
        res["nfev"].index_add_(0, vox, r3["nfev"])
        res["cost"].index_copy_(0, vox, torch.where(better, r3["cost"], res["cost"].index_select(0, vox)))
        res["r2"].index_copy_(0, vox, torch.where(better, r3["r2"], res["r2"].index_select(0, vox)))
        if res["cov"] is not None:
            keep = res["cov"].index_select(0, vox)
            res["cov"].index_copy_(0, vox, torch.where(better[:, None, None], r3["cov"], keep))
        return int(better.sum().item())
