"""B200 drop-in for the reference's ``NNLSSolver`` (solvers/nnls_solver.py:16-210)."""

from __future__ import annotations

import logging
from typing import Any

import numpy as np

from .. import engine
from .base import BaseSolver, PixelResults, _PixelFitResult

log = logging.getLogger("pyneapple_b200")


# "auto" page-locks result blocks up to this many bytes (see NNLSSolver.fit)
_AUTO_PIN_LIMIT = 1 << 30


def regularization_matrix(n_bins: int, order: int, mu: float = 1.0) -> np.ndarray:
    """Tikhonov matrix, orders 0-3 (behaviour of model_functions/nnls.py:46-85).

    0: zeros; 1: forward difference; 2: second difference; 3: the reference's
    ``[1, 2, -6, 2, 1]`` pentadiagonal stencil; all scaled by ``mu``.
    """
    stencils = {0: {}, 1: {0: -1.0, 1: 1.0}, 2: {-1: 1.0, 0: -2.0, 1: 1.0},
                3: {-2: 1.0, -1: 2.0, 0: -6.0, 1: 2.0, 2: 1.0}}
    if order not in stencils:
        raise NotImplementedError(f"Regularization order {order} not supported. Use 0-3.")
    R = np.zeros((n_bins, n_bins))
    for off, val in stencils[order].items():
        R += np.diag(np.full(n_bins - abs(off), val), off)
    return R * mu


class NNLSSolver(BaseSolver):
    """Regularised NNLS for every voxel on the GPU.

    Constructor and ``fit`` as in nnls_solver.py:34-51, :88-127; ``tol`` is
    stored and unused exactly like there; ``multi_threading`` / ``n_pools``
    are accepted and ignored.
    """

    def __init__(self, model: Any, reg_order: int = 0, mu: float = 0.02, max_iter: int = 250,
                 tol: float = 1e-8, verbose=False, multi_threading: bool = False, **solver_kwargs: Any):
        super().__init__(model, max_iter, tol, verbose)
        self.multi_threading = multi_threading
        self.n_pools = solver_kwargs.pop("n_pools", None)
        self.device = solver_kwargs.pop("device", 0)
        self.chunk_vox = solver_kwargs.pop("chunk_vox", 0)
        self.pinned_outputs = solver_kwargs.pop("pinned_outputs", "auto")  # see CurveFitSolver
        self._last_out_key = None
        self.algorithm = solver_kwargs.pop("algorithm", "auto")
        # "fused": h = B^T y inside the solver kernel; "gemm": materialised first by the tensor-core GEMM
        self.dual_init = solver_kwargs.pop("dual_init", "fused")
        self._out_cache = None
        self._peaks_cache = None
        self.reg_order = reg_order
        self.mu = mu
        self.status_ = None
        self.iterations_ = None
        self.r_squared_ = None

    @property
    def primary_device(self) -> int:
        from .. import _lib

        return _lib.resolve_devices(self.device)[0]

    def get_regularization_matrix(self) -> np.ndarray:
        return regularization_matrix(self.model.n_bins, self.reg_order, self.mu)

    def _build_regularized_basis(self, xdata: np.ndarray) -> np.ndarray:
        return np.concatenate([self.model.get_basis(xdata), self.get_regularization_matrix()], axis=0)

    def _extend_signal(self, signal: np.ndarray) -> np.ndarray:
        return np.concatenate((signal, np.zeros((signal.shape[0], self.model.n_bins))), axis=1)

    # The reference's internal hooks (nnls_solver.py:129-210): they take the regularised system
    # ``A = [basis; mu R]`` and the zero-extended signals; split back into the structured form the kernel uses.
    def _fit_data(self, basis, signal):
        """``(coefficients (n_pixels, n_bins), residuals (n_pixels,))``; sets ``pixel_results_``."""
        A = np.asarray(basis, dtype=np.float64)
        n_bins = A.shape[1]
        n_b = A.shape[0] - n_bins
        if n_b < 1:
            raise ValueError(f"regularised basis must have more rows than columns, got {A.shape}")
        y = np.asarray(signal, dtype=np.float64)
        if y.ndim == 1:
            y = y[None, :]
        if np.any(y[:, n_b:] != 0.0):
            raise NotImplementedError("the right-hand side of the regulariser rows must be zero")
        res = engine.nnls_fit(A[:n_b], A[n_b:], np.ascontiguousarray(y[:, :n_b]), self.max_iter,
                              device=self.device, algorithm=self.algorithm, dual_init=self.dual_init)
        status = res["status"]
        self.pixel_results_ = PixelResults(params=res["coefficients"], covariance=None, success=status == 1,
                                           residual=res["residual"])
        return res["coefficients"], res["residual"]

    def _fit_single_pixel(self, basis, signal, pixel_idx=None):
        keep = self.pixel_results_
        try:
            self._fit_data(basis, np.asarray(signal)[None, :])
            pr = self.pixel_results_[0]
        finally:
            self.pixel_results_ = keep
        return _PixelFitResult(params=pr.params, residual=pr.residual, success=pr.success,
                               message=None if pr.success else "Maximum number of iterations reached.")

    def fit(self, xdata, signal, pixel_fixed_params=None) -> "NNLSSolver":
        self._reset_state()
        xdata = np.asarray(xdata)
        on_device = engine._is_torch_cuda(signal)
        if not on_device:
            signal = np.asarray(signal)
        basis = self.model.get_basis(xdata)
        reg = self.get_regularization_matrix()
        if signal.ndim == 1:
            signal = signal[None, :]
        self.n_pixels = signal.shape[0]
        out = None
        self.status_ = self.iterations_ = self.r_squared_ = None
        key = (self.n_pixels, basis.shape[1])
        use_pinned = bool(self.pinned_outputs) and self.n_pixels >= 16384
        if use_pinned and self.pinned_outputs == "auto" and (
                self._last_out_key != key or self.n_pixels * (basis.shape[1] + 4) * 8 > _AUTO_PIN_LIMIT):
            # first fit of this shape: locking the pages costs more than one staged download.  Large blocks
            # never pay for themselves in "auto" mode: page-locking takes ~0.6 s per GB (5 s for the 8.4 GB of
            # spectra of a 4.19 M voxel volume) and saves ~0.03 s per GB and fit; pinned_outputs=True asks for it
            use_pinned = False
        self._last_out_key = key
        if use_pinned:
            from .. import _lib
            from .curvefit import _unreferenced

            cached = self._out_cache
            if cached is not None and cached[0] == key and _unreferenced(cached[1]):
                out = cached[1]
            else:
                out = dict(
                    coefficients=_lib.pinned_empty((self.n_pixels, basis.shape[1])),
                    residual=_lib.pinned_empty((self.n_pixels,)),
                    status=_lib.pinned_empty((self.n_pixels,), np.int32),
                    iterations=_lib.pinned_empty((self.n_pixels,), np.int32),
                    r2=_lib.pinned_empty((self.n_pixels,)))
                self._out_cache = (key, out)
        res = engine.nnls_fit(basis, reg, signal, self.max_iter, device=self.device,
                              chunk_vox=self.chunk_vox, out=None if on_device else out, algorithm=self.algorithm,
                              dual_init=self.dual_init)
        if on_device and out is not None:
            # device-resident signal, host results: straight into the page-locked block (the 8.4 GB of
            # spectra of a 4.19 M voxel volume take 0.15 s that way, 0.46 s into a fresh numpy array)
            for k, v in res.items():
                engine.to_host_into(v, out[k])
            res = dict(out)
        elif on_device:
            res = {k: engine.to_host(v) for k, v in res.items()}
        status = res["status"]
        self.status_, self.iterations_ = status, res["iterations"]
        self.r_squared_ = res.get("r2")
        success = status == 1
        self.pixel_results_ = PixelResults(
            params=res["coefficients"], covariance=None, success=success,
            messages=lambda i, s=status: None if s[i] == 1 else (
                "Maximum number of iterations reached." if s[i] == 3
                else "array must not contain infs or NaNs"),
            residual=res["residual"],
        )
        self.params_["coefficients"] = res["coefficients"]
        self.diagnostics_["residual"] = res["residual"]
        n_fail = int((~success).sum())
        if n_fail:
            log.warning("%d of %d NNLS fits failed", n_fail, self.n_pixels)
        return self

    def fit_spectrum_peaks(self, xdata, signal, height: float = 0.1, regularized: bool | None = None,
                           cutoffs=None, max_peaks: int = 8, chunk_vox: int = 1 << 20) -> dict:
        """NNLS fit followed by the spectrum post-processing of utility/spectrum.py, all on the device.

        What a user of the reference writes as ``solver.fit(...)`` plus a Python loop of
        ``find_spectrum_peaks`` / ``apply_cutoffs`` over ``params_["coefficients"]``.  The spectra
        (8 n_bins bytes per voxel, 8.4 GB for a 256 x 256 x 64 volume) never leave the GPU: per chunk the
        signal goes up, ``pnb_nnls_fit_device`` and ``pnb_spectrum_peaks_device`` run back to back on the
        chunk's stream, and only the peaks come back.  ``regularized`` defaults to ``reg_order > 0``.
        Returns numpy arrays: the keys of :func:`pyneapple_b200.spectrum.find_spectrum_peaks_batch` plus
        ``residual``, ``status``, ``iterations``, ``r2``.  The solver's ``params_`` are not set.  The arrays live in
        page-locked memory that the solver reuses for the next call of the same shape: copy what must
        outlive it.
        """
        import torch

        from .. import _lib, spectrum

        _lib.require_device()
        xdata = np.asarray(xdata)
        signal = np.asarray(signal, dtype=np.float64)
        if signal.ndim == 1:
            signal = signal[None, :]
        if regularized is None:
            regularized = self.reg_order > 0
        basis = self.model.get_basis(xdata)
        reg = self.get_regularization_matrix()
        n_vox, K = signal.shape[0], (0 if cutoffs is None else len(cutoffs))
        dev = torch.device("cuda", self.primary_device)
        # page-locked result arrays: the downloads are asynchronous and run at PCIe speed (a pageable
        # destination costs ~100 ms per million voxels, most of the fit time)
        pe = _lib.pinned_empty
        key = ("peaks", n_vox, int(max_peaks), K)
        if self._peaks_cache is not None and self._peaks_cache[0] == key:
            out = self._peaks_cache[1]  # page-locking ~1 GB costs as much as the fit: keep it for the next volume
        else:
            out = None
        out = out or dict(n_peaks=pe((n_vox,), np.int32), peak_index=pe((n_vox, max_peaks), np.int32),
                   d_values=pe((n_vox, max_peaks)), f_values=pe((n_vox, max_peaks)),
                   d_cut=pe((n_vox, K)) if K else np.empty((n_vox, 0)), f_cut=pe((n_vox, K)) if K else np.empty((n_vox, 0)),
                   residual=pe((n_vox,)), status=pe((n_vox,), np.int32), iterations=pe((n_vox,), np.int32), r2=pe((n_vox,)))
        # one stream: the library keeps one set of per-device scratch buffers for device-path launches,
        # and the uploads (128 B per voxel) are small next to the fit itself
        for s0 in range(0, n_vox, chunk_vox):
            s1 = min(n_vox, s0 + chunk_vox)
            y = engine.to_device(signal[s0:s1], dev)
            fit = engine.nnls_fit(basis, reg, y, self.max_iter, device=self.device, algorithm=self.algorithm)
            pk = spectrum.find_spectrum_peaks_batch(fit.pop("coefficients"), self.model.bins, height, regularized,
                                                    cutoffs=cutoffs, max_peaks=max_peaks)
            for k, v in {**fit, **pk}.items():
                if v is not None:
                    torch.from_numpy(out[k][s0:s1]).copy_(v, non_blocking=True)
        torch.cuda.synchronize(dev)
        self._peaks_cache = (key, out)
        return out
