"""GPU: constrained tri-exponential fits vs the reference's SLSQP outputs (contract in
pyneapple_b200/solvers/constrained.py): residual no worse, feasible, tightly converged."""

from __future__ import annotations

import numpy as np
import pytest
from scipy.optimize import minimize

from _util import load

pytestmark = pytest.mark.gpu

from oracle import ref_port  # noqa: E402
from pyneapple_b200 import models  # noqa: E402
from pyneapple_b200.solvers import ConstrainedCurveFitSolver  # noqa: E402

NAMES = ["f1", "D1", "f2", "D2", "D3"]


def _fit(g, **kw):
    p0 = {n: float(v) for n, v in zip(NAMES, g["p0"])}
    bounds = {n: (float(l), float(u)) for n, l, u in zip(NAMES, g["lb"], g["ub"])}
    s = ConstrainedCurveFitSolver(model=models.TriExpModel(), max_iter=250, tol=1e-8, p0=p0, bounds=bounds,
                                  fraction_constraint=True, **kw)
    return s.fit(g["b"], g["y"])


def _rnorm(b, y, p):
    m = ref_port.Model("triexp", "reduced")
    return np.array([np.linalg.norm(m.forward(b, *p[:, i]) - y[i]) for i in range(y.shape[0])])


def test_contract_vs_reference_slsqp():
    g = load("slsqp_triexp_c5")
    s = _fit(g)
    got = np.stack([s.params_[n] for n in NAMES])
    ref = g["params"]
    assert np.array([pr.success for pr in s.pixel_results_]).all() and g["success"].all()
    # (ii) feasibility
    assert (got[0] + got[2] <= 1.0 + 1e-12).all()
    assert (got >= g["lb"][:, None] - 1e-15).all() and (got <= g["ub"][:, None] + 1e-15).all()
    # (i) residual norm no worse than the reference's, voxel for voxel
    r_ours, r_ref = _rnorm(g["b"], g["y"], got), _rnorm(g["b"], g["y"], ref)
    assert (r_ours <= r_ref * (1 + 1e-9)).all()
    # the constraint is active somewhere in this sample, and those voxels sit on the face
    assert s.n_active_ >= 1
    # (iii) tightly converged: polishing our answer with SLSQP(ftol=1e-15) does not move it
    m = ref_port.Model("triexp", "reduced")
    worst = 0.0
    for i in range(0, got.shape[1], 4):
        y = g["y"][i]
        obj = lambda p: 0.5 * np.sum((y - m.forward(g["b"], *p)) ** 2)
        grad = lambda p: -m.jacobian(g["b"], *p).T @ (y - m.forward(g["b"], *p))
        res = minimize(obj, got[:, i], jac=grad, method="SLSQP", bounds=list(zip(g["lb"], g["ub"])),
                       constraints=[{"type": "ineq", "fun": lambda p: 1.0 - p[0] - p[2]}],
                       options={"maxiter": 500, "ftol": 1e-15})
        assert obj(res.x) >= obj(got[:, i]) * (1 - 1e-9) - 1e-18  # SLSQP cannot find anything better
        worst = max(worst, float(np.max(np.abs(res.x - got[:, i]) / np.abs(got[:, i]))))
    assert worst < 1e-4
    # (iv) distribution of the differences to the reference (informational, bounded loosely)
    rel = np.abs(got - ref) / np.abs(ref)
    assert np.median(rel.max(axis=0)) < 5e-2


def test_constructor_contract():
    with pytest.raises(ValueError):
        ConstrainedCurveFitSolver(model=models.TriExpModel(fit_reduced=False), max_iter=10, tol=1e-8,
                                  p0={}, bounds={}, fraction_constraint=True)
    with pytest.raises(ValueError):
        ConstrainedCurveFitSolver(model=models.BiExpModel(), max_iter=10, tol=1e-8,
                                  p0={"f1": 0.2, "D1": 1e-3, "D2": 1e-2},
                                  bounds={"f1": (0.0, 1.0), "D1": (1e-5, 1e-2), "D2": (1e-3, 1.0)},
                                  fraction_constraint=True)
    g = load("slsqp_triexp_c5")
    s = _fit(g)
    assert s.method == "SLSQP" and s._fraction_names == ["f1", "f2"] and s._fraction_indices == [0, 2]
    assert s.diagnostics_["pcov"].shape == (256, 5, 5)
