"""One fit per segmentation label on the label's mean signal
(mirror of reference fitters/segmentationwise.py:19-179).  The per-label means — one
pass over the whole volume — are reduced on the device (``pnb_segment_means_*``); the
handful of fits go through the GPU solver for API uniformity, not for speed."""

from __future__ import annotations

import time

import numpy as np

from .. import engine
from .. import validation as V
from ..models import describe_model, family_forward
from .base import BaseFitter, PixelIndices


class SegmentationWiseFitter(BaseFitter):
    def fit(self, xdata, image, segmentation=None, fixed_param_maps=None, **fit_kwargs):
        if segmentation is None:
            raise ValueError("segmentation is required for segmentation-wise fitting")
        xdata = np.asarray(xdata)
        V.validate_xdata(xdata)
        V.validate_data_shapes(xdata, image)
        self.n_measurements = len(xdata)
        self.image_shape = image.shape
        segmentation = V.validate_segmentation(np.asarray(segmentation), image.shape)
        _t0 = time.perf_counter()
        self._drop_previous_results()
        segs_to_fit = self._extract_segmentation_mean_signals(image, segmentation)
        pixel_fixed_params = None
        if fixed_param_maps is not None:
            V.validate_fixed_param_maps(fixed_param_maps, image.shape[:-1],
                                        self.solver.model._all_param_names)
            dev = getattr(self.solver, "device", 0)
            pixel_fixed_params = {
                name: engine.segment_means(np.asarray(vol, dtype=np.float64)[..., None], segmentation, device=dev)[1][:, 0]
                for name, vol in fixed_param_maps.items()
            }
        self.solver.fit(xdata, segs_to_fit, pixel_fixed_params=pixel_fixed_params, **fit_kwargs)
        self.fitted_params_ = dict(self.solver.params_)
        fit_time = time.perf_counter() - _t0
        self.results_ = self._assemble_fit_result(xdata, segs_to_fit, fit_time)
        return self

    def _extract_segmentation_mean_signals(self, image, segmentation):
        # np.unique includes the background label 0, like the reference (segmentationwise.py:116)
        labels, means, _, inverse = engine.segment_means(image, segmentation, device=getattr(self.solver, "device", 0))
        inverse = inverse.reshape(segmentation.shape)
        order = np.argsort(inverse.ravel(), kind="stable")  # voxels grouped by label, C order within
        coords = np.stack(np.unravel_index(order, segmentation.shape), axis=1)
        self.segment_labels = labels
        self._segment_of_pixel = inverse.ravel()[order]
        self.pixel_indices = PixelIndices(coords)
        self.pixel_to_segment = _PixelToSegment(coords, self._segment_of_pixel)
        return means

    def predict(self, xdata, **predict_kwargs):
        self._check_fitted()
        xdata = np.asarray(xdata)
        if xdata.ndim != 1:
            raise ValueError(f"Expected xdata to be 1D array, but got shape {xdata.shape}")
        names = self.solver.model.param_names
        desc = describe_model(self.solver.model)
        full = [np.atleast_1d(np.asarray(self.fitted_params_[n], float)) if n in names
                else None for n in desc.all_names]
        nseg = max(v.shape[0] for v in full if v is not None)
        full = [np.full(nseg, float(desc.fixed[n])) if v is None else v for v, n in zip(full, desc.all_names)]
        per_segment = family_forward(desc, xdata, full)  # (n_segments, n_b)
        predictions = per_segment[self._segment_of_pixel]
        return self._reconstruct_volume(predictions, self.pixel_indices, self.image_shape[:-1] + (xdata.size,))


class _PixelToSegment:
    """Mapping ``(x, y, z) -> segment index`` without materialising a dict of millions of tuples."""

    def __init__(self, coords, seg):
        self._coords, self._seg = coords, seg
        self._lookup = None

    def __getitem__(self, key):
        if self._lookup is None:
            self._lookup = {tuple(int(v) for v in c): int(s) for c, s in zip(self._coords, self._seg)}
        return self._lookup[tuple(int(v) for v in key)]

    def __len__(self):
        return len(self._seg)
