"""``FitResult`` — what every fitter stores in ``results_`` (mirror of reference result.py:11-106)."""

from __future__ import annotations

from dataclasses import dataclass
from typing import Any

import numpy as np


@dataclass
class FitResult:
    params: dict[str, np.ndarray]
    success: np.ndarray
    n_iterations: np.ndarray | None = None
    messages: list | None = None
    covariance: np.ndarray | None = None
    residuals: np.ndarray | None = None
    r_squared: np.ndarray | None = None
    fit_time: float = 0.0
    image_shape: tuple | None = None
    pixel_indices: Any = None
    n_pixels: int = 0
    solver_name: str = ""
    model_name: str = ""

    @property
    def n_converged(self) -> int:
        return int(np.sum(self.success))

    @property
    def convergence_rate(self) -> float:
        return 0.0 if self.n_pixels == 0 else float(self.n_converged / self.n_pixels)

    @property
    def mean_r_squared(self) -> float | None:
        if self.r_squared is None:
            return None
        if np.all(np.isnan(self.r_squared)):
            return float("nan")
        return float(np.nanmean(self.r_squared))

    def __repr__(self) -> str:  # pragma: no cover
        r2 = self.mean_r_squared
        return (f"FitResult(model={self.model_name!r}, solver={self.solver_name!r}, n_pixels={self.n_pixels}, "
                f"converged={self.n_converged}/{self.n_pixels}, "
                f"mean_R²={'n/a' if r2 is None else format(r2, '.4f')}, fit_time={self.fit_time:.3f}s)")
