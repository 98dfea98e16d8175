"""Two-step "segmented" fitting, device resident
(mirror of reference fitters/segmented.py:23-256).

Step 1 fits a simple model on a b-value subset, step 2 fits the full model
with selected step-1 parameters fixed *per voxel*.  The reference does this
with two PixelWiseFitters and a host-side volume reconstruction in between;
here the masked signal matrix is uploaded once, step 1's parameter vectors
stay on the GPU and become the frozen rows of step 2's parameter matrix — no
host round trip between the steps.
"""

from __future__ import annotations

import time
from dataclasses import replace

import numpy as np

from .. import engine
from .. import validation as V
from .base import BaseFitter
from .pixelwise import PixelWiseFitter


class SegmentedFitter(BaseFitter):
    def __init__(self, step1_solver, step2_solver, step1_bvalue_range=None, fixed_from_step1=None,
                 param_mapping=None, **fitter_kwargs):
        super().__init__(solver=step2_solver, **fitter_kwargs)
        self.step1_solver = step1_solver
        self.step2_solver = step2_solver
        self.step1_bvalue_range = step1_bvalue_range
        self.fixed_from_step1 = fixed_from_step1 or []
        self.param_mapping = param_mapping or {}
        self.step1_params_: dict = {}
        self.step1_result_ = None
        self._validate_init()

    def _validate_init(self) -> None:
        step1_all = self.step1_solver.model._all_param_names
        for name in self.fixed_from_step1:
            if name not in step1_all:
                raise ValueError(
                    f"fixed_from_step1 name {name!r} is not a parameter of the Step 1 model. "
                    f"Available: {step1_all}"
                )
        step2_all = self.step2_solver.model._all_param_names
        for src in self.fixed_from_step1:
            dst = self.param_mapping.get(src, src)
            if dst not in step2_all:
                raise ValueError(
                    f"Mapped parameter {dst!r} (from {src!r}) is not a parameter of the Step 2 model. "
                    f"Available: {step2_all}"
                )

    def _bvalue_mask(self, xdata):
        if self.step1_bvalue_range is None:
            return np.ones(len(xdata), dtype=bool)
        lo, hi = self.step1_bvalue_range
        mask = np.ones(len(xdata), dtype=bool)
        if lo is not None:
            mask &= xdata >= lo
        if hi is not None:
            mask &= xdata <= hi
        if not np.any(mask):
            raise ValueError(
                f"No b-values fall within the range {self.step1_bvalue_range}. Available b-values: {xdata}"
            )
        if int(mask.sum()) < 3:
            raise ValueError(
                f"Step 1 requires at least 3 b-values for fitting, but only {int(mask.sum())} "
                f"fall within the range {self.step1_bvalue_range}."
            )
        return mask

    def _subset_bvalues(self, xdata, image):
        mask = self._bvalue_mask(np.asarray(xdata))
        if mask.all():
            return xdata, image
        return xdata[mask], image[..., mask]

    def fit(self, xdata, image, segmentation=None, **fit_kwargs):
        from .. import _lib

        _lib.require_device()  # EngineError, not a torch error, when there is no GPU
        import torch

        xdata = np.asarray(xdata)
        V.validate_xdata(xdata)
        V.validate_data_shapes(xdata, image)
        self.n_measurements = len(xdata)
        self.image_shape = image.shape
        _t0 = time.perf_counter()
        self._drop_previous_results()
        if segmentation is not None:
            segmentation = V.validate_segmentation(np.asarray(segmentation), image.shape)
        bmask = self._bvalue_mask(xdata)
        device_ok = all(hasattr(s, "fit_device") for s in (self.step1_solver, self.step2_solver))
        if not device_ok:
            raise TypeError("SegmentedFitter needs B200 CurveFitSolver instances for both steps")
        pixel_to_fit = self._extract_pixel_data(image, segmentation)
        dev = torch.device("cuda", self.step2_solver.primary_device)
        y_dev = engine.to_device(np.ascontiguousarray(pixel_to_fit), dev)
        # ---- step 1 on the b-value subset --------------------------------------
        sub_idx = torch.as_tensor(np.nonzero(bmask)[0], device=dev)
        y1 = y_dev if bmask.all() else y_dev.index_select(1, sub_idx).contiguous()
        res1 = self.step1_solver.fit_device(xdata[bmask], y1)
        # ---- step 2 with the step-1 parameters frozen per voxel --------------------
        fixed = {}
        for src in self.fixed_from_step1:
            dst = self.param_mapping.get(src, src)
            row = res1["free_rows"][res1["free_names"].index(src)]
            fixed[dst] = res1["params"][row]
        res2 = self.step2_solver.fit_device(xdata, y_dev, pixel_fixed_params=fixed or None)
        # ---- publish ------------------------------------------------------------------
        self.step1_solver.store_device_result(res1)
        self.step2_solver.store_device_result(res2)
        self.step1_params_ = dict(self.step1_solver.params_)
        step1 = PixelWiseFitter(solver=self.step1_solver)
        step1.n_measurements, step1.image_shape = int(bmask.sum()), image.shape[:-1] + (int(bmask.sum()),)
        step1.pixel_indices = self.pixel_indices
        step1.fitted_params_ = dict(self.step1_params_)
        self.step1_result_ = step1._assemble_fit_result(xdata[bmask], None, time.perf_counter() - _t0)
        self.fitted_params_ = dict(self.step2_solver.params_)
        for src in self.fixed_from_step1:
            self.fitted_params_[self.param_mapping.get(src, src)] = self.step1_params_[src]
        fit_time = time.perf_counter() - _t0
        self.results_ = replace(self._assemble_fit_result(xdata, None, fit_time), fit_time=fit_time)
        return self

    def _get_param_names(self):
        return self.solver.model._all_param_names
