"""B200 drop-in for the reference's ``CurveFitSolver`` (solvers/curvefit.py:16-392).

Same constructor, same ``fit`` signature, same result attributes and error
behaviour; the per-voxel ``scipy.optimize.curve_fit`` loop / joblib pool is
replaced by one batched CUDA solve (``pnb_trf_fit_*``).
"""

from __future__ import annotations

import logging
from typing import Any

import numpy as np

from .. import engine
from ..models import describe_model
from .. import validation as V
from .base import BaseSolver, PixelResults, _PixelFitResult

log = logging.getLogger("pyneapple_b200")

# extra keyword arguments the reference would splat into curve_fit /
# least_squares (curvefit.py:70-73, 305) and that the device solver honours
_HONOURED_KWARGS = {"xtol", "gtol", "x_scale", "sigma", "absolute_sigma", "loss", "f_scale", "diff_step"}
# accepted for TOML compatibility, meaningless on the GPU
_IGNORED_KWARGS = {"n_pools"}


def _unreferenced(arrays: dict) -> bool:
    """True when nothing but ``arrays`` itself references its values (views count: they keep their
    base alive).  Expected count per array: the dict, the loop variable, getrefcount's argument."""
    import sys

    for a in arrays.values():
        if sys.getrefcount(a) > 3:
            return False
    return True


class CurveFitSolver(BaseSolver):
    """Bounded non-linear least squares for every voxel on the GPU.

    Args mirror ``CurveFitSolver.__init__`` (curvefit.py:36-89).  Additional
    keyword ``jac``:

    * ``"reference"`` (default) — do what the reference does: SciPy's 2-point
      finite-difference Jacobian when no parameter is fixed, the analytic
      Jacobian when any parameter is fixed (curvefit.py:274-293);
    * ``"analytic"`` — always the analytic Jacobian (faster, differs from the
      reference by ~1e-6 relative);
    * ``"2-point"`` — always finite differences.
    """

    def __init__(
        self,
        model: Any,
        max_iter: int,
        tol: float,
        p0: dict[str, float],
        bounds: dict[str, tuple[float, float]],
        verbose: bool = False,
        method: str = "trf",
        multi_threading: bool = False,
        use_jacobian: bool = True,
        **solver_kwargs,
    ):
        super().__init__(model=model, max_iter=max_iter, tol=tol, verbose=verbose)
        self.method = method
        self.multi_threading = multi_threading
        self.use_jacobian = use_jacobian and hasattr(model, "jacobian")
        self.n_pools = solver_kwargs.pop("n_pools", None)
        self.jac = solver_kwargs.pop("jac", "reference")
        # GPU(s) of the host path: an ordinal, a list of ordinals or "all" (one call, voxels sharded
        # over the node's GPUs); the device-resident drivers use the first one
        self.device = solver_kwargs.pop("device", 0)
        self.chunk_vox = solver_kwargs.pop("chunk_vox", 0)
        # True / "lazy": covariances are computed and stay on the GPU until somebody reads them
        # (diagnostics_["pcov"] is then a LazyArray); "eager": shipped with the other results; False: skipped
        self.want_cov = solver_kwargs.pop("want_cov", True)
        # result arrays in page-locked memory (no staging copy on the way back).  "auto": from the
        # second fit of the same shape on (locking ~1 GB costs more than one staged download);
        # True: always; False: never.  A cached block is only reused when nothing outside the solver
        # still references the arrays of the previous fit.
        self.pinned_outputs = solver_kwargs.pop("pinned_outputs", "auto")
        self._out_cache = None
        self._last_out_key = None
        if self.jac not in ("reference", "analytic", "2-point"):
            raise ValueError("jac must be 'reference', 'analytic' or '2-point'")
        unknown = set(solver_kwargs) - _HONOURED_KWARGS - _IGNORED_KWARGS
        if unknown:
            raise NotImplementedError(
                f"curve_fit option(s) {sorted(unknown)} have no B200 implementation "
                f"(honoured: {sorted(_HONOURED_KWARGS)})"
            )
        self.solver_kwargs = solver_kwargs
        self._desc = describe_model(model)  # raises NotImplementedError for unknown models

        names = self.model.param_names
        if isinstance(p0[names[0]], (int, float, np.ndarray)):
            V.validate_parameter_names(p0, names)
            self.p0 = p0
        else:
            raise ValueError(
                "p0 must be a dict with parameter names as keys and initial values as values."
            )
        if isinstance(bounds[names[0]], tuple):
            V.validate_parameter_names(bounds, names)
            self.bounds = bounds
        else:
            raise ValueError(
                "bounds must be a dict with parameter names as keys and (lower, upper) tuples as values."
            )
        self.status_ = None
        self.nfev_ = None
        self.cost_ = None
        self.r_squared_ = None

    def _ls_method(self) -> str:
        """least_squares method the engine runs (subclasses with their own outer method use 'trf')."""
        return self.method if self.method in engine.METHODS else "trf"

    # ------------------------------------------------------------------
    def fit(self, xdata, ydata, p0=None, bounds=None, pixel_fixed_params=None, **fit_kwargs):
        """Fit all voxels (curvefit.py:91-159).  Unknown keyword arguments are
        swallowed like the reference does (``fixed_param_maps=None`` arrives
        here from ``run_pipeline`` through ``IDEALFitter.fit``)."""
        self._reset_state()
        if self.method not in engine.METHODS and self.method != "lm":
            raise NotImplementedError(
                f"method={self.method!r}: SciPy's 'trf', 'dogbox' and 'lm' have a B200 implementation"
            )
        xdata = np.asarray(xdata)
        on_device = engine._is_torch_cuda(ydata)
        if not on_device:
            ydata = np.asarray(ydata)
        V.validate_data_shapes(xdata, ydata)
        n_pixels = ydata.shape[0] if ydata.ndim > 1 else 1
        if ydata.ndim == 1:
            ydata = ydata[None, :]
        p0_m, lb_m, ub_m = self._p0_and_bounds(p0, bounds, n_pixels)
        lm = self.method == "lm"
        bounded = bool(np.isfinite(np.asarray(lb_m)).any() or np.isfinite(np.asarray(ub_m)).any())
        lm_rejected = lm and bounded
        n_free_model = len(self.model.param_names) - len(set(pixel_fixed_params or {}) & set(self.model.param_names))
        lm_too_few = lm and not bounded and n_free_model > xdata.shape[0]
        if lm_rejected or lm_too_few:
            # curve_fit raises before it looks at the data; one dogbox evaluation supplies R^2 at p0
            res, free_names = self._solve(xdata, ydata, p0_m, lb_m, ub_m, pixel_fixed_params, n_pixels,
                                          max_nfev=1, method="dogbox")
        else:
            res, free_names = self._solve(xdata, ydata, p0_m, lb_m, ub_m, pixel_fixed_params, n_pixels)
        if on_device:
            cov_dev = None
            if (self.want_cov is True or self.want_cov == "lazy") and res.get("cov") is not None and not (lm_rejected or lm_too_few):
                cov_dev = res.pop("cov")  # stays on the GPU until somebody reads it
            res = {k: (engine.to_host(v) if v is not None else None) for k, v in res.items()}
            res.pop("n_failed", None)
            if cov_dev is not None:
                from .._lazy import LazyArray

                res["cov"] = LazyArray(tuple(cov_dev.shape), [(0, int(cov_dev.shape[0]), cov_dev)])
        if lm_rejected or lm_too_few:
            # curve_fit rejects 'lm' for a bounded problem before it looks at the data, for every voxel
            # (scipy/optimize/_minpack_py.py: "Method 'lm' only works for unconstrained problems."): the
            # reference's solver turns that into success=False, params = p0, NaN covariance
            # (solvers/curvefit.py:308-317).  The single evaluation above (dogbox leaves x0 = p0 untouched)
            # only supplies R^2 at p0.
            res["status"][...] = engine.ST_LM_BOUNDED if lm_rejected else engine.ST_LM_TOO_FEW_DATA
            res["n_failed"] = n_pixels
            if res.get("cov") is not None:
                res["cov"] = np.full(tuple(res["cov"].shape), np.nan)
        self._store(res, free_names, n_pixels)
        return self

    # ------------------------------------------------------------------
    # The reference's two internal hooks, with its signatures (curvefit.py:171-244, 246-317).  They are
    # exercised directly by its own tests; here they run on the GPU like everything else (the class the
    # plugin registers also derives from Pyneapple's CurveFitSolver, whose SciPy versions these override).
    def _fit_data(self, xdata, ydata, p0, bounds, n_pixels, pixel_fixed_params=None):
        """``(popt (n_params, n_pixels), pcov (n_pixels, n_params, n_params))``; sets ``pixel_results_``."""
        ydata = np.asarray(ydata)
        if ydata.ndim == 1:
            ydata = ydata[None, :]
        res, free_names = self._solve(np.asarray(xdata), ydata, np.asarray(p0, float), np.asarray(bounds[0], float),
                                      np.asarray(bounds[1], float), pixel_fixed_params, n_pixels)
        self._store(res, free_names, n_pixels)
        popt = np.stack([np.asarray(res["params"][r]) for r in self._free_rows])
        pcov = res["cov"]
        pcov = np.full((n_pixels, popt.shape[0], popt.shape[0]), np.nan) if pcov is None else np.asarray(pcov)
        return popt, pcov

    def _fit_single_pixel(self, xdata, ydata, p0, bounds, pixel_idx=None, pixel_fixed=None):
        """One voxel -> ``_PixelFitResult`` (failure: params = p0, NaN covariance, SciPy's message)."""
        p0 = np.asarray(p0, float)
        pf = {k: np.array([float(v)]) for k, v in pixel_fixed.items()} if pixel_fixed else None
        keep = (self.params_, self.diagnostics_, self.pixel_results_)
        try:
            self._fit_data(xdata, np.asarray(ydata)[None, :], p0[:, None], (np.asarray(bounds[0], float)[:, None],
                                                                            np.asarray(bounds[1], float)[:, None]), 1, pf)
            return self.pixel_results_[0]
        finally:
            self.params_, self.diagnostics_, self.pixel_results_ = keep

    # ------------------------------------------------------------------
    def _solve(self, xdata, ydata, p0_m, lb_m, ub_m, pixel_fixed_params, n_pixels, max_nfev=None, method=None):
        """Assemble full-parameter arrays and call the engine.

        ``p0_m`` / ``lb_m`` / ``ub_m`` are ``(n_model,)`` or ``(n_model,
        n_pixels)`` over ``model.param_names``.
        """
        desc = self._desc
        model_names = list(self.model.param_names)
        all_names = list(desc.all_names)
        pix_fixed = {k: v for k, v in (pixel_fixed_params or {}).items() if k in all_names}
        fixed_names = set(desc.fixed) | set(pix_fixed)
        frozen = engine.frozen_mask(desc, fixed_names)
        free_names = [n for n in model_names if n not in pix_fixed]

        per_voxel_p0 = p0_m.ndim == 2 or bool(pix_fixed)
        per_voxel_bd = lb_m.ndim == 2

        def full(src, per_voxel, filler, is_p0):
            rows = []
            for name in all_names:
                if name in desc.fixed:
                    v = float(desc.fixed[name]) if is_p0 else filler
                elif name in pix_fixed:
                    v = pix_fixed[name] if is_p0 else filler
                else:
                    v = src[model_names.index(name)]
                rows.append(v)
            if not per_voxel:
                return np.array([float(r) for r in rows])
            out = np.empty((len(all_names), n_pixels))
            for j, r in enumerate(rows):
                r = np.asarray(r, dtype=float)
                if r.ndim == 1 and r.shape[0] != n_pixels:
                    raise ValueError(
                        f"per-pixel values for '{all_names[j]}' have length {r.shape[0]}, expected {n_pixels}"
                    )
                out[j] = r
            return out

        P0 = full(p0_m, per_voxel_p0, 0.0, True)
        LB = full(lb_m, per_voxel_bd, -np.inf, False)
        UB = full(ub_m, per_voxel_bd, np.inf, False)

        jac_mode, xtol, gtol = self._jac_and_tols(fixed_names, method or self._ls_method())
        xs_full, x_scale_jac = self._x_scale_args(all_names, model_names)
        self._fit_opts = dict(method=method or self._ls_method(), xtol=xtol, gtol=gtol, n_data=int(np.asarray(xdata).shape[0]),
                              n_params=len(all_names) - len(fixed_names))
        res = engine.trf_fit(
            desc, xdata, ydata, P0, LB, UB, frozen,
            max_nfev=self.max_iter if max_nfev is None else max_nfev, ftol=self.tol,
            xtol=xtol, gtol=gtol,
            jac_mode=jac_mode, x_scale=xs_full, x_scale_jac=x_scale_jac,
            want_cov=self.want_cov, device=self.device, chunk_vox=self.chunk_vox,
            out=self._pinned_out(len(all_names), len(all_names) - len(fixed_names), n_pixels, ydata),
            method=method or self._ls_method(), finish_wait=getattr(self, "_finish_wait", 0),
            **self._extras_for(all_names, model_names, int(np.asarray(xdata).shape[0])),
        )
        self._free_rows = [all_names.index(n) for n in free_names]
        return res, free_names

    def _jac_and_tols(self, fixed_names, method):
        """Jacobian mode and xtol / gtol for the engine: the reference's rule (analytic Jacobian only
        when a parameter is fixed, curvefit.py:274-293) and SciPy's defaults for the method
        (least_squares: 1e-8 / 1e-8; leastsq behind method="lm": 1.49012e-8 / 0)."""
        lm = method == "lm"
        fd = engine.JAC_MINPACK_FORWARD if lm else engine.JAC_TWO_POINT
        if self.jac == "reference":
            jac_mode = engine.JAC_ANALYTIC if fixed_names else fd
        else:
            jac_mode = engine.JAC_ANALYTIC if self.jac == "analytic" else fd
        xtol = self.solver_kwargs.get("xtol", engine.LM_XTOL if lm else 1e-8)
        gtol = self.solver_kwargs.get("gtol", engine.LM_GTOL if lm else 1e-8)
        return jac_mode, float(xtol), float(gtol)

    def _x_scale_args(self, all_names, model_names):
        """``x_scale`` of ``least_squares`` (forwarded by the reference, curvefit.py:305) as the
        engine wants it: a vector over ``all_names`` (1.0 for fixed parameters) and the 'jac' flag."""
        x_scale = self.solver_kwargs.get("x_scale", 1.0)
        x_scale_jac = isinstance(x_scale, str)
        if x_scale_jac and x_scale != "jac":
            raise ValueError("`x_scale` must be 'jac' or array_like with positive numbers.")
        xs_full = None
        if not x_scale_jac:
            xs = np.broadcast_to(np.asarray(x_scale, float), (len(model_names),))
            if not np.all(np.isfinite(xs)) or np.any(xs <= 0):
                raise ValueError("`x_scale` must be 'jac' or array_like with positive numbers.")
            xs_full = np.ones(len(all_names))
            for i, n in enumerate(model_names):
                xs_full[all_names.index(n)] = xs[i]
        return xs_full, x_scale_jac

    def _extras_for(self, all_names, model_names, n_b):
        out = self._extras_args(all_names, model_names)
        w = out.get("weights")
        if w is not None and np.ndim(w) == 0:
            out["weights"] = np.full(n_b, 1.0 / float(w))   # scalar sigma
        return out

    def _extras_args(self, all_names, model_names):
        """``sigma`` / ``absolute_sigma`` of ``curve_fit`` and ``loss`` / ``f_scale`` / ``diff_step`` of
        ``least_squares`` (all forwarded by the reference, curvefit.py:70-73, 305) as engine arguments.  The same
        values apply to every voxel, as in the reference's loop."""
        kw = self.solver_kwargs
        out = {}
        sigma = kw.get("sigma")
        if sigma is not None:
            sigma = np.asarray(sigma, float)
            if sigma.ndim == 2:
                raise NotImplementedError("a 2-D `sigma` (covariance matrix of the data) has no B200 implementation")
            # scipy/optimize/_minpack_py.py: transform = 1.0 / sigma (a scalar sigma applies to every point)
            out["weights"] = 1.0 / sigma if sigma.ndim == 1 else sigma.reshape(())
        if kw.get("absolute_sigma"):
            out["absolute_sigma"] = True
        loss = kw.get("loss", "linear")
        if callable(loss):
            raise NotImplementedError("a callable `loss` has no B200 implementation")
        if loss != "linear":
            out["loss"] = loss
        if "f_scale" in kw:
            out["f_scale"] = float(kw["f_scale"])
        ds = kw.get("diff_step")
        if ds is not None:
            ds = np.broadcast_to(np.asarray(ds, float), (len(model_names),))
            full = np.zeros(len(all_names))
            for i, n in enumerate(model_names):
                full[all_names.index(n)] = ds[i]
            out["diff_step"] = full
        return out

    # ------------------------------------------------------------------
    def fit_device(self, xdata, y_dev, p0=None, bounds=None, pixel_fixed_params=None, want_cov=None):
        """Device-resident fit for the IDEAL / segmented drivers: nothing leaves the GPU.

        ``y_dev (n_pix, n_b)``, optional ``p0 (n_params, n_pix)``, ``bounds = (lb, ub)`` and
        ``pixel_fixed_params {name: (n_pix,)}`` are CUDA tensors (rows over
        ``model.param_names``).  Returns the engine's dict of CUDA tensors plus
        ``free_names``; call :meth:`store_device_result` to publish it as ``params_`` etc.
        """
        from .. import _lib

        _lib.require_device()
        import torch

        desc = self._desc
        model_names = list(self.model.param_names)
        all_names = list(desc.all_names)
        n_pix = y_dev.shape[0]
        dev = y_dev.device
        pix_fixed = {k: v for k, v in (pixel_fixed_params or {}).items() if k in all_names}
        fixed_names = set(desc.fixed) | set(pix_fixed)
        frozen = engine.frozen_mask(desc, fixed_names)
        free_names = [n for n in model_names if n not in pix_fixed]
        p0_def = V.p0_vector(self.p0, model_names)
        lb_def, ub_def = V.bounds_vectors(self.bounds, model_names)

        def assemble(src, default, filler, is_p0):
            rows, per_voxel = [], False
            for name in all_names:
                if name in desc.fixed:
                    rows.append(float(desc.fixed[name]) if is_p0 else filler)
                elif name in pix_fixed:
                    if is_p0:
                        rows.append(pix_fixed[name].to(device=dev, dtype=torch.float64))
                        per_voxel = True
                    else:
                        rows.append(filler)
                elif src is not None:
                    v = src[model_names.index(name)]
                    if isinstance(v, torch.Tensor) and v.ndim >= 1:
                        rows.append(v)
                        per_voxel = True
                    else:  # a vector over the parameters: one value for all voxels
                        rows.append(float(v))
                else:
                    rows.append(float(default[model_names.index(name)]))
            if not per_voxel:
                return np.array(rows, dtype=float)
            return torch.stack([r if isinstance(r, torch.Tensor)
                                else torch.full((n_pix,), float(r), dtype=torch.float64, device=dev)
                                for r in rows])

        P0 = assemble(p0, p0_def, 0.0, True)
        LB = assemble(None if bounds is None else bounds[0], lb_def, -np.inf, False)
        UB = assemble(None if bounds is None else bounds[1], ub_def, np.inf, False)
        jac_mode, xtol, gtol = self._jac_and_tols(fixed_names, self._ls_method())
        xs_full, x_scale_jac = self._x_scale_args(all_names, model_names)
        self._fit_opts = dict(method=self._ls_method(), xtol=xtol, gtol=gtol, n_data=int(np.asarray(xdata).shape[0]),
                              n_params=len(all_names) - len(fixed_names))
        res = engine.trf_fit(
            desc, np.asarray(xdata, float), y_dev, P0, LB, UB, frozen, max_nfev=self.max_iter,
            ftol=self.tol, xtol=xtol,
            gtol=gtol, jac_mode=jac_mode,
            x_scale=xs_full, x_scale_jac=x_scale_jac,
            want_cov=self.want_cov if want_cov is None else want_cov, device=self.primary_device,
            method=self._ls_method(), finish_wait=getattr(self, "_finish_wait", 0),
            **self._extras_for(all_names, model_names, int(np.asarray(xdata).shape[0])),
        )
        res["free_names"] = free_names
        res["free_rows"] = [all_names.index(n) for n in free_names]
        return res

    def store_device_result(self, res):
        """Copy a :meth:`fit_device` result to the host and publish it like ``fit`` does."""
        self._reset_state()
        free_names, self._free_rows = res["free_names"], res["free_rows"]
        cov_dev = None
        if (self.want_cov is True or self.want_cov == "lazy") and res.get("cov") is not None:
            cov_dev = res["cov"]  # stays on the GPU until somebody reads it
        host = {k: (engine.to_host(v) if hasattr(v, "cpu") else v) for k, v in res.items()
                if k not in ("free_names", "free_rows", "n_active", "n_released") and not (k == "cov" and cov_dev is not None)}
        if cov_dev is not None:
            from .._lazy import LazyArray

            host["cov"] = LazyArray(tuple(cov_dev.shape), [(0, int(cov_dev.shape[0]), cov_dev)])
        self._store(host, free_names, host["params"].shape[1])
        return self

    @property
    def primary_device(self) -> int:
        from .. import _lib

        return _lib.resolve_devices(self.device)[0]

    def _pinned_out(self, n_all, n_free, n_pixels, ydata):
        if not self.pinned_outputs or engine._is_torch_cuda(ydata) or n_pixels < 65536:
            return None
        from .. import _lib

        eager = bool(self.want_cov) and self.want_cov is not True and self.want_cov != "lazy"
        key = (n_all, n_free, n_pixels, eager)
        nbytes = n_pixels * (8 * n_all + 28 + (8 * n_free * n_free if eager else 0))
        if self.pinned_outputs == "auto" and (self._last_out_key != key or nbytes > (1 << 30)):
            # first fit of a shape, or a block so large that locking its pages (~0.6 s per GB) never pays
            # for itself: plain numpy arrays
            self._last_out_key = key
            return None
        self._last_out_key = key
        cached = self._out_cache
        if cached is not None and cached[0] == key:
            # reuse only when the previous fit's arrays are referenced by nobody else (views held by
            # a caller keep their base alive: count = cache dict + getrefcount's argument)
            if _unreferenced(cached[1]):
                return cached[1]
        out = dict(
            params=_lib.pinned_empty((n_all, n_pixels)),
            status=_lib.pinned_empty((n_pixels,), np.int32),
            nfev=_lib.pinned_empty((n_pixels,), np.int32),
            njev=_lib.pinned_empty((n_pixels,), np.int32),
            cost=_lib.pinned_empty((n_pixels,)),
            r2=_lib.pinned_empty((n_pixels,)),
        )
        if eager:
            out["cov"] = _lib.pinned_empty((n_pixels, n_free, n_free))
        self._out_cache = (key, out)
        return out

    def _reset_state(self):
        super()._reset_state()
        # drop the solver's own references to the previous fit's arrays (see _pinned_out)
        self.status_ = self.nfev_ = self.njev_ = self.cost_ = self.r_squared_ = None

    def _store(self, res, free_names, n_pixels):
        rows = [res["params"][r] for r in self._free_rows]  # views of the (n_all, n_vox) output
        pcov = res["cov"]
        status = res["status"]
        self.status_, self.nfev_, self.cost_ = status, res["nfev"], res["cost"]
        self.njev_ = res["njev"]
        self.r_squared_ = res.get("r2")
        if pcov is None:
            # covariance not requested (want_cov=False): a read-only NaN view of the right shape instead of
            # filling n_pixels x n x n doubles (0.84 GB / 113 ms for a 4.19 M-voxel tri-exponential slab)
            pcov = np.broadcast_to(np.nan, (n_pixels, len(free_names), len(free_names)))
        self.pixel_results_ = PixelResults(
            params=rows, covariance=pcov, status=status,
            messages=lambda i, s=status, o=dict(getattr(self, "_fit_opts", {})): engine.status_message(
                s[i], "lm" if self.method == "lm" else o.get("method", "trf"), max_nfev=self.max_iter, ftol=self.tol,
                xtol=o.get("xtol", 0.0), gtol=o.get("gtol", 0.0), n_params=o.get("n_params", 0), n_data=o.get("n_data", 0)),
        )
        self.params_ = {
            name: [float(rows[i][0])] if n_pixels == 1 else rows[i]
            for i, name in enumerate(free_names)
        }
        self.diagnostics_ = {"pcov": pcov[0] if n_pixels == 1 else pcov, "n_pixels": n_pixels}
        # the kernels count the failed voxels (host path); a pass over 4 M status words costs milliseconds
        n_fail = res.get("n_failed")
        if n_fail is None:
            n_fail = int(np.count_nonzero(np.asarray(status) <= 0))
        self.n_failed_ = int(n_fail)
        if n_fail:
            log.warning("%d of %d voxel fits failed (see pixel_results_[i].message)", n_fail, n_pixels)

    # ------------------------------------------------------------------
    def _validate_p0_and_bounds(self, p0, bounds, n_pixels):
        """The reference's ``_validate_p0_and_bounds`` (curvefit.py:319-392): ``(p0, (lower, upper))``,
        each ``(n_params, n_pixels)``.  Scalars are broadcast views (read-only, no ``np.tile`` copy)."""
        p0_m, lb_m, ub_m = self._p0_and_bounds(p0, bounds, n_pixels)

        def wide(a):
            return a if a.ndim == 2 else np.broadcast_to(a[:, None], (a.shape[0], n_pixels))

        return wide(p0_m), (wide(lb_m), wide(ub_m))

    def _p0_and_bounds(self, p0, bounds, n_pixels):
        """Same validation, without the widening: ``(p0, lb, ub)`` as ``(n_params,)`` vectors or
        ``(n_params, n_pixels)`` arrays (what the engine takes)."""
        names = self.model.param_names
        if p0 is not None:
            if isinstance(p0, dict):
                if isinstance(p0[names[0]], np.ndarray):
                    raise ValueError(
                        "p0 should either be a basic dict with scalar values or a single np.ndarray "
                        "of initial values for all parameters. Spatial non-uniform p0 should be "
                        "handled separately before calling fit()."
                    )
                elif isinstance(p0[names[0]], (int, float)):
                    p0 = V.p0_vector(p0, names)
                else:
                    raise ValueError(
                        "p0 dict values must be either all scalars or a single np.ndarray of "
                        "initial values for all parameters."
                    )
            elif isinstance(p0, np.ndarray):
                if p0.ndim != 2 or p0.shape[1] != n_pixels:
                    raise ValueError(
                        f"p0 shape {p0.shape} does not match number of voxels in ydata {n_pixels}."
                    )
            else:
                raise ValueError("p0 must be either a dict or a single np.ndarray.")
        else:
            p0 = V.p0_vector(self.p0, names)

        if bounds is not None:
            if isinstance(bounds, dict):
                if isinstance(bounds[names[0]], tuple):
                    lb, ub = V.bounds_vectors(bounds, names)
                else:
                    raise ValueError(
                        "bounds dict values must be tuples of (lower, upper) for each parameter."
                    )
            elif isinstance(bounds, tuple) and len(bounds) == 2:
                if not all(isinstance(b, np.ndarray) for b in bounds):
                    raise ValueError(
                        "bounds tuple must contain two np.ndarrays (lower, upper) of shape (n_params, n_pixels)."
                    )
                lb, ub = bounds
                if lb.ndim != 2 or ub.ndim != 2 or lb.shape[1] != n_pixels or ub.shape[1] != n_pixels:
                    raise ValueError(
                        f"bounds shape {lb.shape} and {ub.shape} do not match number of voxels in ydata {n_pixels}."
                    )
            else:
                raise ValueError(
                    "bounds must be either a dict with parameter names as keys and (lower, upper) "
                    "tuples as values, or a tuple of (lower, upper) np.ndarrays."
                )
        else:
            lb, ub = V.bounds_vectors(self.bounds, names)
        return np.asarray(p0, float), np.asarray(lb, float), np.asarray(ub, float)
