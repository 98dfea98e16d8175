"""Solver contract of the B200 engine (mirror of reference solvers/base.py:14-90).

State ownership is the reference's: a solver owns ``params_``, ``diagnostics_``
and ``pixel_results_`` after ``fit``.  ``pixel_results_`` is a *lazy* sequence
here — 4.19 M dataclass instances cost more host time than the whole GPU fit —
that materialises a per-voxel record only when it is indexed or iterated, so
code written against the reference (``fitters/base.py:211-253``: ``len()``,
iteration, ``[0].params.shape``) keeps working.
"""

from __future__ import annotations

from collections.abc import Sequence
from dataclasses import dataclass
from typing import Any

import numpy as np


@dataclass
class _PixelFitResult:
    """Per-voxel record, field for field the reference's ``_PixelFitResult``."""

    params: np.ndarray
    covariance: np.ndarray | None = None
    success: bool = True
    message: str | None = None
    n_iterations: int | None = None
    residual: float | None = None


class PixelResults(Sequence):
    """Array-backed, lazily materialised list of :class:`_PixelFitResult`.

    ``params`` is either an ``(n_vox, n_values)`` array or a list of ``n_values`` row vectors of
    length ``n_vox`` (views into the solver's parameter-major output: no copy is made).
    """

    def __init__(self, params, covariance=None, success=None, messages=None,
                 n_iterations=None, residual=None, status=None):
        if isinstance(params, (list, tuple)):
            self._rows = [np.asarray(r) for r in params]
            self._matrix = None
            self._n = self._rows[0].shape[0] if self._rows else 0
        else:
            self._rows = None
            self._matrix = np.asarray(params)
            self._n = self._matrix.shape[0]
        self.covariance = covariance      # (n_vox, n, n) | None
        self._success = None if success is None else np.asarray(success, bool)
        self._status = status             # success = status > 0 when given instead of `success`
        self.messages = messages          # callable(i) -> str | None, or None
        self.n_iterations = n_iterations  # (n_vox,) | None
        self.residual = residual          # (n_vox,) | None

    @property
    def success(self) -> np.ndarray:
        if self._success is None:
            self._success = (np.ones(self._n, bool) if self._status is None
                             else np.asarray(self._status) > 0)
        return self._success

    @property
    def params(self) -> np.ndarray:
        """``(n_vox, n_values)`` matrix (materialised on first use when built from rows)."""
        if self._matrix is None:
            self._matrix = np.stack(self._rows, axis=1)
        return self._matrix

    def __len__(self) -> int:
        return self._n

    def _one(self, i: int) -> _PixelFitResult:
        p = (self._matrix[i] if self._matrix is not None
             else np.array([r[i] for r in self._rows], dtype=np.float64))
        return _PixelFitResult(
            params=p,
            covariance=None if self.covariance is None else self.covariance[i],
            success=bool(self.success[i]),
            message=None if self.messages is None else self.messages(i),
            n_iterations=None if self.n_iterations is None else int(self.n_iterations[i]),
            residual=None if self.residual is None else float(self.residual[i]),
        )

    def __getitem__(self, i):
        if isinstance(i, slice):
            return [self._one(j) for j in range(*i.indices(len(self)))]
        if i < 0:
            i += len(self)
        if not 0 <= i < len(self):
            raise IndexError(i)
        return self._one(i)

    def __iter__(self):
        for i in range(len(self)):
            yield self._one(i)


class BaseSolver:
    """Common state / accessors (reference solvers/base.py:41-90)."""

    def __init__(self, model: Any, max_iter: int = 250, tol: float = 1e-8,
                 verbose: bool = False, **solver_kwargs):
        self.model = model
        self.max_iter = max_iter
        self.tol = tol
        self.verbose = verbose
        self.diagnostics_: dict[str, Any] = {}
        self.params_: dict[str, Any] = {}
        self.pixel_results_: Sequence = []

    def fit(self, *args, **kwargs) -> "BaseSolver":  # pragma: no cover - abstract
        raise NotImplementedError

    def get_diagnostics(self) -> dict[str, Any]:
        if len(self.diagnostics_) == 0:
            raise RuntimeError(
                "No diagnostics available. Ensure fit() has been called and diagnostics are stored."
            )
        return self.diagnostics_.copy()

    def get_params(self) -> dict[str, Any]:
        if len(self.params_) == 0:
            raise RuntimeError(
                "No parameters available. Ensure fit() has been called and parameters are stored."
            )
        return self.params_.copy()

    def _reset_state(self):
        self.diagnostics_ = {}
        self.params_ = {}
        self.pixel_results_ = []
