"""B200 solvers with the reference's names and registry (solvers/__init__.py:8-35)."""

from __future__ import annotations

from .base import BaseSolver, PixelResults, _PixelFitResult
from .curvefit import CurveFitSolver
from .nnls import NNLSSolver
from .constrained import ConstrainedCurveFitSolver

_REGISTRY: dict[str, type] = {
    "curvefit": CurveFitSolver,
    "nnls": NNLSSolver,
    "constrained_curvefit": ConstrainedCurveFitSolver,
}


def get_solver(name: str, **kwargs) -> BaseSolver:
    key = name.lower()
    if key not in _REGISTRY:
        raise ValueError(f"Unknown solver: {name!r}. Available: {sorted(_REGISTRY)}")
    return _REGISTRY[key](**kwargs)


__all__ = ["BaseSolver", "ConstrainedCurveFitSolver", "CurveFitSolver", "NNLSSolver", "PixelResults", "get_solver"]
