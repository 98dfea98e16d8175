"""Fitter base: spatial bookkeeping around a batched GPU solve.

Mirrors the contract of reference fitters/base.py:19-351 (attributes
``solver``, ``results_``, ``fitted_params_``, ``image_shape``,
``pixel_indices``, ``n_measurements``; ``fit`` / ``predict`` /
``get_fitted_params``).  The reference spends ~24 us per voxel in Python loops
here (``_extract_pixel_data``, ``_assemble_fit_result``,
``_compute_r_squared``: ~100 s at 4.19 M voxels); everything below is
array-at-a-time and R^2 comes out of the fit kernel itself.
"""

from __future__ import annotations

from collections.abc import Sequence
from typing import Any

import numpy as np

from .. import engine
from ..models import describe_model, family_forward
from ..result import FitResult


class PixelIndices(Sequence):
    """``list[tuple[int, int, int]]`` look-alike backed by an ``(n, ndim)`` index array.

    For an unmasked volume only the spatial shape is kept and the coordinates are generated on
    demand (4.19 M voxels: a 100 MB table nobody may ever look at).
    """

    def __init__(self, coords: np.ndarray | None = None, full_shape: tuple | None = None,
                 flat=None, n: int | None = None):
        self._array = None if coords is None else np.ascontiguousarray(coords, dtype=np.int64)
        self._shape = None if full_shape is None else tuple(int(v) for v in full_shape)
        # masked volume given as C-order flat positions (what the device gather / scatter use):
        # the coordinate table is only built when somebody asks for it.  `flat` may be a callable
        # that fetches the positions (they may still be on the GPU); `n` is then their count.
        self._flat_loader = flat if callable(flat) else None
        self._n = n
        self._flat_store = None if (flat is None or callable(flat)) else np.ascontiguousarray(flat, dtype=np.int64)

    @property
    def _flat(self):
        if self._flat_store is None and self._flat_loader is not None:
            self._flat_store = np.ascontiguousarray(self._flat_loader(), dtype=np.int64)
            self._flat_loader = None
        return self._flat_store

    @_flat.setter
    def _flat(self, value):
        self._flat_store = value

    @property
    def array(self) -> np.ndarray:
        if self._array is None:
            flat = np.arange(int(np.prod(self._shape))) if self._flat is None else self._flat
            self._array = np.stack(np.unravel_index(flat, self._shape), axis=1).astype(np.int64)
        return self._array

    @property
    def flat(self) -> np.ndarray | None:
        """C-order flat position of every voxel in ``full_shape`` (``None`` when built from bare coordinates)."""
        if self._flat is None and self._shape is not None and self._array is not None:
            self._flat = np.ravel_multi_index(tuple(self._array.T), self._shape).astype(np.int64)
        return self._flat

    @property
    def is_full(self) -> bool:
        return self._array is None and self._flat_store is None and self._flat_loader is None

    def __len__(self):
        if self._array is not None:
            return self._array.shape[0]
        if self._n is not None:
            return int(self._n)
        return int(np.prod(self._shape)) if self.is_full else self._flat.shape[0]

    def __getitem__(self, i):
        if isinstance(i, slice):
            if self._array is None:
                idx = np.arange(len(self))[i]
                if self._flat is not None:
                    idx = self._flat[idx]
                return [tuple(int(v) for v in np.unravel_index(j, self._shape)) for j in idx]
            return [tuple(int(v) for v in row) for row in self._array[i]]
        if self._array is None:
            n = len(self)
            if i < 0:
                i += n
            if not 0 <= i < n:
                raise IndexError(i)
            j = i if self._flat is None else int(self._flat[i])
            return tuple(int(v) for v in np.unravel_index(j, self._shape))
        return tuple(int(v) for v in self._array[i])

    def __iter__(self):
        for row in self.array:
            yield tuple(int(v) for v in row)

    def __eq__(self, other):
        try:
            return len(other) == len(self) and all(tuple(a) == tuple(b) for a, b in zip(self, other))
        except TypeError:
            return NotImplemented


class LazyVolumes(Sequence):
    """``list[np.ndarray]`` look-alike over CUDA tensors: an element is downloaded when it is first
    indexed (IDEAL keeps every level's parameter map; most callers only look at the last one)."""

    def __init__(self, tensors):
        self._items = list(tensors)

    def __len__(self):
        return len(self._items)

    def __getitem__(self, i):
        if isinstance(i, slice):
            return [self[j] for j in range(*i.indices(len(self)))]
        v = self._items[i]
        if not isinstance(v, np.ndarray):
            v = self._items[i] = engine.to_host(v)
        return v

    def device_tensor(self, i):
        """Element ``i`` as it is stored (CUDA tensor until it has been read on the host)."""
        return self._items[i]


class BaseFitter:
    def __init__(self, solver: Any, verbose: bool = False, **fitter_kwargs):
        self.solver = solver
        self.verbose = verbose
        self.fitter_kwargs = fitter_kwargs
        self.results_: FitResult | None = None
        self.fitted_params_: dict = {}
        self.image_shape: tuple | None = None
        self.pixel_indices: Any = None
        self.n_measurements: int | None = None

    # ------------------------------------------------------------------
    def fit(self, xdata, image, segmentation=None, **fit_kwargs):  # pragma: no cover - abstract
        raise NotImplementedError

    def get_fitted_params(self):
        return self.fitted_params_

    def _drop_previous_results(self):
        """Release this fitter's references to the previous fit's arrays before fitting again: the
        solvers reuse their page-locked result blocks only when nobody else still holds them."""
        self.results_ = None
        self.fitted_params_ = {}

    def predict(self, xdata: np.ndarray, **predict_kwargs) -> np.ndarray:
        """Model signal for every fitted voxel, ``(X, Y, Z, len(xdata))`` (fitters/base.py:93-128)."""
        self._check_fitted()
        xdata = np.asarray(xdata)
        if xdata.ndim != 1:
            raise ValueError(
                f"xdata must be a 1D array of independent variable values, got shape {xdata.shape}."
            )
        output_shape = self.image_shape[:-1] + (xdata.size,)
        dev_pred = self._predict_device(xdata, output_shape)
        if dev_pred is not None:
            return dev_pred
        predictions = self._predict_flat(xdata)
        return self._reconstruct_volume(predictions, self.pixel_indices, output_shape)

    def _device_index(self):
        """(torch device, flat-index tensor | None) of the fitted voxels, or None without a usable index."""
        import torch

        pi = self.pixel_indices
        if not isinstance(pi, PixelIndices):
            return None
        dev = torch.device("cuda", getattr(self.solver, "primary_device", 0))
        if pi.is_full:
            return dev, None
        if pi._shape is None:
            pi._shape = tuple(self.image_shape[:-1])
        flat = pi.flat
        return dev, (None if flat is None else engine.to_device(flat, dev))

    def _predict_device(self, xdata, output_shape):
        """Prediction, scatter into the volume and nothing else on the GPU (``pnb_predict_device``; a
        dictionary model is one FP64 GEMM); the finished ``(X, Y, Z, n)`` volume comes back."""
        import torch

        from .. import _lib

        if _lib.load().pnb_device_count() < 1 or len(self.pixel_indices) < 4096:
            return None
        where = self._device_index()
        if where is None:
            return None
        dev, flat = where
        n_out = int(np.prod(output_shape[:-1]))
        model = self.solver.model
        with torch.cuda.device(dev):
            if hasattr(model, "get_basis") and "coefficients" in self.fitted_params_:
                coef = engine.to_device(np.ascontiguousarray(self.fitted_params_["coefficients"], np.float64), dev)
                basis_t = torch.as_tensor(np.ascontiguousarray(model.get_basis(xdata).T)).to(dev)
                pred = coef @ basis_t  # (n_vox, n_bins) x (n_bins, n): a plain library GEMM
                if flat is not None:
                    pred = engine.move_rows(pred, flat, n_out, scatter=True)
            else:
                desc = describe_model(model)
                rows = []
                n_pix = len(self.pixel_indices)
                for n in desc.all_names:
                    if n in self.fitted_params_:
                        rows.append(np.atleast_1d(np.asarray(self.fitted_params_[n], dtype=np.float64)))
                    else:
                        rows.append(np.full(n_pix, float(desc.fixed[n])))
                par = engine.to_device(np.ascontiguousarray(np.stack(rows)), dev)
                pred = engine.predict_device(desc, xdata, par, flat, n_out)
            return engine.to_host(pred).reshape(output_shape)

    def reconstruct_maps(self) -> dict:
        """Parameter volumes of the last fit, float32, zero where nothing was fitted — what the
        reference's ``reconstruct_maps(fitter.fitted_params_, fitter.pixel_indices, image_shape[:3])``
        (io/nifti.py:279-312) returns, without the per-voxel tuple list it builds."""
        from ..maps import reconstruct_maps

        self._check_fitted()
        return reconstruct_maps(self.fitted_params_, self.pixel_indices, tuple(self.image_shape[:-1]),
                                device=getattr(self.solver, "primary_device", 0))

    def _predict_flat(self, xdata):
        model = self.solver.model
        if hasattr(model, "get_basis") and "coefficients" in self.fitted_params_:
            return np.asarray(self.fitted_params_["coefficients"]) @ model.get_basis(xdata).T
        names = self._get_param_names()
        desc = describe_model(model)
        full = []
        for n in desc.all_names:
            if n in self.fitted_params_:
                full.append(np.atleast_1d(np.asarray(self.fitted_params_[n], dtype=float)))
            else:
                full.append(None)
        n_pix = max(v.shape[0] for v in full if v is not None)
        full = [np.full(n_pix, float(desc.fixed[n])) if v is None else v for v, n in zip(full, desc.all_names)]
        del names
        return family_forward(desc, xdata, full)

    # ------------------------------------------------------------------
    def _extract_pixel_data(self, image: np.ndarray, segmentation: np.ndarray) -> np.ndarray:
        """``image[segmentation != 0]`` in C order over (x, y, z) (fitters/base.py:280-308)."""
        if self.n_measurements is None:
            raise RuntimeError(
                "n_measurements must be set before extracting pixel data. Validate image data first."
            )
        if segmentation is not None:
            mask = segmentation != 0
            if self._gather_on_device(image):
                # image[mask] in numpy is a single-threaded pass at ~1 GB/s: upload the volume once
                # (staged, multi-threaded) and compact the masked rows on the GPU instead; the solver's
                # device path then fits the tensor where it is
                import torch

                flat = np.flatnonzero(mask.reshape(-1))
                self.pixel_indices = PixelIndices(full_shape=image.shape[:-1], flat=flat)
                dev = torch.device("cuda", getattr(self.solver, "primary_device", 0))
                img_d = engine.to_device(np.ascontiguousarray(image, dtype=np.float64).reshape(-1, self.n_measurements), dev)
                return engine.move_rows(img_d, engine.to_device(flat, dev), img_d.shape[0], scatter=False)
            pixel_to_fit = image[mask]
            self.pixel_indices = PixelIndices(np.argwhere(mask), full_shape=image.shape[:-1])
        else:
            pixel_to_fit = image.reshape(-1, self.n_measurements)
            self.pixel_indices = PixelIndices(full_shape=image.shape[:-1])
        return pixel_to_fit

    def _gather_on_device(self, image) -> bool:
        from .. import _lib

        return (isinstance(image, np.ndarray) and image.nbytes >= (32 << 20)
                and _lib.load().pnb_device_count() > 0 and hasattr(self.solver, "primary_device"))

    def _reconstruct_volume(self, flat_values, pixel_indices, spatial_shape) -> np.ndarray:
        if isinstance(pixel_indices, PixelIndices) and pixel_indices.is_full:
            # unmasked volume: the flat order is the C order of the volume
            return np.asarray(flat_values, dtype=np.float64).reshape(spatial_shape).copy()
        vol = np.zeros(spatial_shape, dtype=np.float64)
        coords = pixel_indices.array if isinstance(pixel_indices, PixelIndices) else np.asarray(list(pixel_indices))
        if coords.size:
            vol[tuple(coords.T)] = flat_values
        return vol

    def _compute_r_squared(self, xdata, pixel_signals) -> np.ndarray:
        """Host-side R^2 (fitters/base.py:142-186); the fit kernels normally provide it directly."""
        try:
            pixel_signals = engine.to_host(pixel_signals)
            predictions = self._predict_flat(np.asarray(xdata))
            ss_res = np.sum((pixel_signals - predictions) ** 2, axis=1)
            mean = pixel_signals.mean(axis=1, keepdims=True)
            ss_tot = np.sum((pixel_signals - mean) ** 2, axis=1)
            with np.errstate(divide="ignore", invalid="ignore"):
                return np.where(ss_tot > 0, 1.0 - ss_res / ss_tot, np.nan).astype(np.float64)
        except Exception:  # noqa: BLE001 - the reference also degrades to NaN
            return np.full(pixel_signals.shape[0], np.nan)

    def _assemble_fit_result(self, xdata, pixel_signals, fit_time, pixel_indices=None) -> FitResult:
        """Array-at-a-time version of fitters/base.py:188-274."""
        s = self.solver
        prs = s.pixel_results_
        n_pixels = len(prs)
        if hasattr(prs, "success"):  # PixelResults of the B200 solvers
            success = np.asarray(prs.success, bool)
            n_it = None if prs.n_iterations is None else np.asarray(prs.n_iterations).astype(np.intp)
            covariance = prs.covariance
            residuals = None if prs.residual is None else np.asarray(prs.residual, np.float64)
            status = getattr(s, "status_", None)
            messages = None
            if prs.messages is not None and not success.all():
                messages = [None] * n_pixels
                for i in np.nonzero(~success)[0]:
                    messages[int(i)] = prs.messages(int(i))
            del status
        else:  # a foreign solver following the reference's list-of-dataclasses contract
            success = np.array([pr.success for pr in prs], dtype=bool)
            its = [pr.n_iterations for pr in prs]
            n_it = (np.array([-1 if i is None else i for i in its], dtype=np.intp)
                    if any(i is not None for i in its) else None)
            msgs = [pr.message for pr in prs]
            messages = msgs if any(m is not None for m in msgs) else None
            covs = [pr.covariance for pr in prs]
            covariance = None
            if any(c is not None for c in covs):
                npar = prs[0].params.shape[0]
                covariance = np.array([np.full((npar, npar), np.nan) if c is None else c for c in covs])
            res = [pr.residual for pr in prs]
            residuals = (np.array([np.nan if r is None else r for r in res], dtype=np.float64)
                         if any(r is not None for r in res) else None)
        r2 = getattr(s, "r_squared_", None)
        if r2 is None or len(r2) != n_pixels:
            r2 = self._compute_r_squared(xdata, pixel_signals)
        return FitResult(
            params=dict(s.params_), success=success, n_iterations=n_it, messages=messages,
            covariance=covariance, residuals=residuals, r_squared=np.asarray(r2, np.float64),
            fit_time=fit_time, image_shape=self.image_shape,
            pixel_indices=pixel_indices if pixel_indices is not None else self.pixel_indices,
            n_pixels=n_pixels, solver_name=type(s).__name__, model_name=type(s.model).__name__,
        )

    def _check_fitted(self) -> None:
        if not self.fitted_params_:
            raise RuntimeError(
                f"{self.__class__.__name__} has not been fitted yet. "
                "Call fit() before predict() or get_fitted_params()."
            )
        if self.pixel_indices is None:
            raise RuntimeError(
                f"{self.__class__.__name__} has not extracted pixel data yet. "
                "Call _extract_pixel_data() before predict() or get_fitted_params()."
            )

    def _get_param_names(self) -> list[str]:
        return self.solver.model.param_names


_ = engine  # the engine is reached through the solvers; imported so a missing library fails early
