"""IDEAL multi-resolution fitting, device resident
(mirror of reference fitters/ideal.py:23-326).

Per level the reference resamples image, segmentation and the previous
level's parameter maps slice by slice with ``cv2.resize`` on the host, builds
per-voxel ``p0`` and ``p0 * (1 +- tol)`` bounds with numpy, calls the solver and
scatters the result through Python lists.  Here the image and mask are
uploaded once; every level runs resample -> mask -> gather -> clamp / bounds
-> bounded TRF fit -> scatter entirely on the GPU (``pnb_resize2d_device`` is
bit-faithful to ``cv2.resize``; the element-wise steps are the same IEEE
operations), and only the finished parameter maps come back to the host.
"""

from __future__ import annotations

import time

import numpy as np

from .. import _lib, engine
from .. import validation as V
from ..resize import METHODS, interpolate_array
from .base import BaseFitter, LazyVolumes, PixelIndices


class IDEALFitter(BaseFitter):
    def __init__(self, solver, dim_steps, step_tol, ideal_dims: int = 2,
                 segmentation_threshold: float = 0.2, interpolation_method: str = "cubic",
                 **fitter_kwargs):
        super().__init__(solver=solver, **fitter_kwargs)
        self.dim_steps = np.asarray(dim_steps)
        self.step_tol = step_tol
        self.ideal_dims = ideal_dims
        self.segmentation_threshold = segmentation_threshold
        if interpolation_method not in METHODS:
            raise ValueError(
                f"Invalid interpolation method: {interpolation_method}. Must be one of {tuple(METHODS)}."
            )
        self.interpolation_method = interpolation_method
        self.step_params: list[np.ndarray] = []
        self.step_pixel_counts: list[int] = []

    # ---- validation (ideal.py:78-95, 265-297) -------------------------------------
    def _validate_fitter_inputs(self, dim_steps, ideal_dims):
        if dim_steps.ndim != 2:
            raise ValueError("dim_steps must be a 2D array of shape (n_steps, ideal_dims).")
        if dim_steps.shape[1] != ideal_dims:
            raise ValueError(f"dim_steps must have {ideal_dims} columns corresponding to ideal_dims.")
        for i in range(dim_steps.shape[0] - 1):
            if not np.all(dim_steps[i + 1] > dim_steps[i]):
                raise ValueError(f"dim_steps row {i + 1} must be greater than row {i} (monotonic increase).")

    def _validate_step_tol(self):
        if not isinstance(self.step_tol, dict):
            raise ValueError(
                "step_tol must be a dict mapping parameter names to tolerance fractions, "
                f"e.g. {{'S0': 0.5, 'D': 0.2}}. Got {type(self.step_tol).__name__}."
            )
        try:
            V.validate_parameter_names(self.step_tol, self.solver.model.param_names)
        except ValueError as exc:
            raise ValueError(
                f"step_tol keys {set(self.step_tol.keys())} do not match model parameter names "
                f"{self.solver.model.param_names}: {exc}"
            ) from exc

    def _validate_image_dims(self, image):
        if image.ndim == 4:
            return image
        if image.ndim == 3:
            if self.ideal_dims == 3:
                raise ValueError(
                    f"Image dimension ({image.ndim}) not sufficient for 3D interpolation (ideal_dims={self.ideal_dims})"
                )
            return np.expand_dims(image, axis=-2)
        raise ValueError(f"Image Array needs to be 3 or 4 not {image.ndim}")

    def _interpolate_array(self, array, target_shape):
        """Host-callable resampler with the reference's signature (ideal.py:299-320)."""
        return interpolate_array(array, target_shape, self.interpolation_method,
                                 device=getattr(self.solver, "primary_device", 0))

    # ---- fit -----------------------------------------------------------------------------
    def fit(self, xdata, image, segmentation=None, z_range=None, **fit_kwargs):
        """``z_range=(z0, z1)`` restricts the fit to a z-slab (multi-GPU sharding: the
        in-plane resampling never couples slices)."""
        if not hasattr(self.solver, "fit_device"):
            raise TypeError("IDEALFitter needs a B200 CurveFitSolver (no CPU fallback)")
        # fit_kwargs: the reference forwards them to solver.fit at every level (ideal.py:240-242).
        # run_pipeline passes fixed_param_maps=None, which the solver swallows; per-voxel arrays cannot
        # be forwarded (their length would have to match every level's voxel count — the reference
        # raises a shape error at the first level that differs), so they are rejected up front.
        for key, val in fit_kwargs.items():
            if val is not None:
                raise NotImplementedError(
                    f"IDEALFitter.fit(..., {key}=...): per-call solver arguments are not supported by the "
                    "device-resident IDEAL driver (per-voxel arrays cannot match every resolution level)"
                )
        _lib.require_device()  # EngineError, not a torch error, when there is no GPU
        import torch

        xdata = np.asarray(xdata)
        V.validate_xdata(xdata)
        V.validate_data_shapes(xdata, image)
        self._validate_step_tol()
        self._validate_fitter_inputs(self.dim_steps, self.ideal_dims)
        image = self._validate_image_dims(np.asarray(image))
        self.n_measurements = len(xdata)
        _t0 = time.perf_counter()
        self._drop_previous_results()
        if not np.allclose(self.dim_steps[-1], image.shape[: self.ideal_dims]):
            raise ValueError("The last step in dim_steps must match the spatial dimensions of the image.")
        if segmentation is not None:
            segmentation = V.validate_segmentation(np.asarray(segmentation), image.shape)
        else:
            segmentation = np.ones(image.shape[:3], dtype=int)
        if z_range is not None:
            z0, z1 = z_range
            image = image[:, :, z0:z1]
            segmentation = segmentation[:, :, z0:z1]
        self.image_shape = image.shape
        Z = image.shape[2]
        if self.ideal_dims == 2:
            dim_steps = np.hstack([self.dim_steps, np.full((self.dim_steps.shape[0], 1), Z)])
        else:
            # ideal_dims = 3: the reference still resamples in-plane only (ideal.py:313-319) and then
            # indexes slices that do not exist when a level's Z differs from the image's
            if not np.all(self.dim_steps[:, 2] == Z):
                raise ValueError(
                    "ideal_dims=3 with a slice count that differs from the image's is not supported: "
                    "the resampling is in-plane (x, y) only"
                )
            dim_steps = self.dim_steps
        solver = self.solver
        names = solver.model.param_names
        n_params = len(names)
        dev = torch.device("cuda", solver.primary_device)
        f64 = dict(dtype=torch.float64, device=dev)
        p0_vals = torch.tensor([solver.p0[n] for n in names], **f64)
        lo_vals = torch.tensor([solver.bounds[n][0] for n in names], **f64)
        hi_vals = torch.tensor([solver.bounds[n][1] for n in names], **f64)
        tol_vals = torch.tensor([self.step_tol[n] for n in names], **f64)
        # resample in the image's own dtype like cv2.resize does (float32 stays float32, anything that is
        # not a float becomes float32, ideal.py:310-311); the fit upcasts to float64
        img_dtype = image.dtype if image.dtype in (np.float32, np.float64) else np.dtype(np.float32)
        img_d = engine.to_device(np.ascontiguousarray(image, dtype=img_dtype), dev)
        seg = segmentation[..., None] if segmentation.ndim == 3 else segmentation
        seg_d = engine.to_device(np.ascontiguousarray(seg), dev)
        if not seg_d.dtype.is_floating_point:
            seg_d = seg_d.to(torch.float32)  # ideal.py:310-311
        thr = torch.tensor(self.segmentation_threshold, dtype=seg_d.dtype, device=dev)

        step_maps = []
        self.step_pixel_counts = []
        res = None
        coords = None
        for step_index, step in enumerate(dim_steps):
            shape = tuple(int(s) for s in step)
            if step_index == 0:
                p0 = lb = ub = None  # solver defaults, broadcast on the device
            else:
                p0_map = interpolate_array(step_maps[-1], shape, self.interpolation_method)
                p0_map = torch.minimum(torch.maximum(p0_map, lo_vals), hi_vals)
                lb_map = torch.minimum(torch.maximum(p0_map * (1 - tol_vals), lo_vals), hi_vals)
                ub_map = torch.minimum(torch.maximum(p0_map * (1 + tol_vals), lo_vals), hi_vals)
            img_l = interpolate_array(img_d, shape, self.interpolation_method)
            seg_l = interpolate_array(seg_d, shape, self.interpolation_method)
            mask = seg_l[..., 0] > thr
            coords = mask.nonzero()  # (n_pix, 3), C order over (x, y, z) like np.where; the level's one host sync
            if coords.shape[0] == 0:
                mask = torch.ones(shape, dtype=torch.bool, device=dev)
                coords = mask.nonzero()
            y = img_l[mask].to(torch.float64)  # (n_pix, n_b)
            if step_index > 0:
                p0 = p0_map[mask].T.contiguous()
                lb = lb_map[mask].T.contiguous()
                ub = ub_map[mask].T.contiguous()
            last = step_index == len(dim_steps) - 1
            res = solver.fit_device(xdata, y, p0=p0, bounds=None if lb is None else (lb, ub),
                                    want_cov=None if last else False)
            param_map = torch.zeros(shape + (n_params,), **f64)
            rows = torch.stack([res["params"][r] for r in res["free_rows"]], dim=1)  # (n_pix, n_params)
            param_map[mask] = rows
            step_maps.append(param_map)
            self.step_pixel_counts.append(int(coords.shape[0]))
        solver.store_device_result(res)
        # every level's map stays on the GPU until it is looked at; the voxel positions likewise
        self.step_params = LazyVolumes(step_maps)
        sy, sz = int(shape[1]), int(shape[2])
        flat_dev = (coords[:, 0] * sy + coords[:, 1]) * sz + coords[:, 2]
        self.pixel_indices = PixelIndices(full_shape=shape, flat=lambda t=flat_dev: engine.to_host(t),
                                          n=int(coords.shape[0]))
        self.fitted_params_ = {}
        for param, values in solver.params_.items():
            self.fitted_params_[param] = values
        fit_time = time.perf_counter() - _t0
        self.results_ = self._assemble_fit_result(xdata, None, fit_time)
        return self
