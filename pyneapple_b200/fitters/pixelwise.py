"""Every masked voxel independently (mirror of reference fitters/pixelwise.py:16-110)."""

from __future__ import annotations

import time

import numpy as np

from .. import validation as V
from .base import BaseFitter


class PixelWiseFitter(BaseFitter):
    def fit(self, xdata, image, segmentation=None, fixed_param_maps=None, **fit_kwargs):
        _t0 = time.perf_counter()
        self._drop_previous_results()
        xdata = np.asarray(xdata)
        V.validate_xdata(xdata)
        V.validate_data_shapes(xdata, image)
        self.n_measurements = len(xdata)
        self.image_shape = image.shape
        if segmentation is not None:
            segmentation = V.validate_segmentation(np.asarray(segmentation), image.shape)
            pixel_to_fit = self._extract_pixel_data(image, segmentation)
        else:
            segmentation = None
            spatial = image.shape[:-1]
            pixel_to_fit = self._extract_pixel_data(image, None)  # zero-copy view, all voxels
            del spatial
        pixel_fixed_params = None
        if fixed_param_maps is not None:
            V.validate_fixed_param_maps(fixed_param_maps, image.shape[:-1],
                                        self.solver.model._all_param_names)
            pixel_fixed_params = {
                name: (vol[segmentation != 0] if segmentation is not None else np.asarray(vol).reshape(-1))
                for name, vol in fixed_param_maps.items()
            }
        self.solver.fit(xdata, pixel_to_fit, pixel_fixed_params=pixel_fixed_params, **fit_kwargs)
        self.fitted_params_ = {}
        for param, values in self.solver.params_.items():
            self.fitted_params_[param] = values
        fit_time = time.perf_counter() - _t0
        self.results_ = self._assemble_fit_result(xdata, pixel_to_fit, fit_time)
        return self
