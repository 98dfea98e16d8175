"""B200 drop-in for the reference's ``NNLSSolver`` (solvers/nnls_solver.py:16-210)."""

from __future__ import annotations

import logging
from typing import Any

import numpy as np

from .. import engine
from .base import BaseSolver, PixelResults

log = logging.getLogger("pyneapple_b200")


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
        self.pinned_outputs = solver_kwargs.pop("pinned_outputs", False)
        self.algorithm = solver_kwargs.pop("algorithm", "auto")
        self._out_cache = None
        self.reg_order = reg_order
        self.mu = mu
        self.status_ = None
        self.iterations_ = None
        self.r_squared_ = None

    def get_regularization_matrix(self) -> np.ndarray:
        return regularization_matrix(self.model.n_bins, self.reg_order, self.mu)

    def _build_regularized_basis(self, xdata: np.ndarray) -> np.ndarray:
        return np.concatenate([self.model.get_basis(xdata), self.get_regularization_matrix()], axis=0)

    def _extend_signal(self, signal: np.ndarray) -> np.ndarray:
        return np.concatenate((signal, np.zeros((signal.shape[0], self.model.n_bins))), axis=1)

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
        if self.pinned_outputs and not on_device:
            from .. import _lib

            key = (self.n_pixels, basis.shape[1])
            if self._out_cache is None or self._out_cache[0] != key:
                self._out_cache = (key, dict(
                    coefficients=_lib.pinned_empty((self.n_pixels, basis.shape[1])),
                    residual=_lib.pinned_empty((self.n_pixels,)),
                    status=_lib.pinned_empty((self.n_pixels,), np.int32),
                    iterations=_lib.pinned_empty((self.n_pixels,), np.int32),
                    r2=_lib.pinned_empty((self.n_pixels,))))
            out = self._out_cache[1]
        res = engine.nnls_fit(basis, reg, signal, self.max_iter, device=self.device,
                              chunk_vox=self.chunk_vox, out=out, algorithm=self.algorithm)
        if on_device:
            res = {k: v.cpu().numpy() for k, v in res.items()}
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
