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


def _solver(g, **kw):
    p0 = {n: float(v) for n, v in zip(NAMES, g["p0"])}
    bounds = {n: (float(l), float(u)) for n, l, u in zip(NAMES, g["lb"], g["ub"])}
    return ConstrainedCurveFitSolver(model=models.TriExpModel(), max_iter=250, tol=1e-8, p0=p0, bounds=bounds,
                                     fraction_constraint=True, **kw)


def _fit(g, **kw):
    return _solver(g, **kw).fit(g["b"], g["y"])


def _rnorm(b, y, p):
    f1, d1, f2, d2, d3 = (v[:, None] for v in p)
    pred = f1 * np.exp(-b * d1) + f2 * np.exp(-b * d2) + (1 - f1 - f2) * np.exp(-b * d3)
    return np.linalg.norm(pred - y, axis=1)


def _polish_moves(g, got, names, fixed, stride):
    """Largest relative parameter change when SLSQP(ftol=1e-15) restarts from our answer."""
    m = ref_port.Model("triexp", "reduced")
    free = [NAMES.index(n) for n in names]
    lb, ub = g["lb"][free], g["ub"][free]
    fr = [i for i, n in enumerate(names) if n.startswith("f")]
    worst = 0.0
    for i in range(0, got.shape[1], stride):
        y = g["y"][i]
        fx = {k: float(v[i]) for k, v in fixed.items()}

        def full(p):
            it = iter(p)
            return [fx[n] if n in fx else next(it) for n in NAMES]

        obj = lambda p: 0.5 * np.sum((y - m.forward(g["b"], *full(p))) ** 2)  # noqa: E731
        grad = lambda p: -(m.jacobian(g["b"], *full(p))[:, free]).T @ (y - m.forward(g["b"], *full(p)))  # noqa: E731
        cons = [{"type": "ineq", "fun": lambda p: 1.0 - sum(p[k] for k in fr)}] if len(fr) >= 2 else []
        x0 = got[free, i]
        res = minimize(obj, x0, jac=grad, method="SLSQP", bounds=list(zip(lb, ub)), constraints=cons,
                       options={"maxiter": 500, "ftol": 1e-15})
        assert obj(res.x) >= obj(x0) * (1 - 1e-9) - 1e-18  # SLSQP cannot find anything better
        worst = max(worst, float(np.max(np.abs(res.x - x0) / np.abs(x0))))
    return worst


def test_contract_vs_reference_slsqp():
    g = load("slsqp_triexp_c5")
    assert g["y"].shape[0] == 4096
    s = _fit(g)
    got = np.stack([s.params_[n] for n in NAMES])
    ref = g["params"]
    assert np.asarray(s.pixel_results_.success).all() and g["success"].all()
    # (ii) feasibility
    assert (got[0] + got[2] <= 1.0 + 1e-12).all()
    assert (got >= g["lb"][:, None] - 1e-15).all() and (got <= g["ub"][:, None] + 1e-15).all()
    # (i) residual norm no worse than the reference's, voxel for voxel
    r_ours, r_ref = _rnorm(g["b"], g["y"], got), _rnorm(g["b"], g["y"], ref)
    assert (r_ours <= r_ref * (1 + 1e-9)).all()
    # the constraint is active on the planted sub-population, and those voxels sit on the face
    assert 0.01 * 4096 < s.n_active_ < 0.1 * 4096
    on_face = got[0] + got[2] > 1.0 - 1e-12
    assert on_face.sum() == s.n_active_ - s.n_released_ and 0 <= s.n_released_ <= 0.5 * s.n_active_
    assert np.isnan(np.asarray(s.diagnostics_["pcov"])[on_face]).all()
    assert np.isfinite(np.asarray(s.diagnostics_["pcov"])[~on_face]).all()
    # (iii) tightly converged: polishing our answer with SLSQP(ftol=1e-15) does not move it
    assert _polish_moves(g, got, NAMES, {}, 16) < 1e-4
    # (iv) distribution of the differences to the reference (informational, bounded loosely)
    rel = np.abs(got - ref) / np.abs(ref)
    assert np.median(rel.max(axis=0)) < 5e-2


def test_per_voxel_fixed_parameters():
    """constrained_curvefit.py:170-177: the fraction set is recomputed from the free parameters."""
    # fixed D3: f1 and f2 stay free, the constraint applies
    g = load("slsqp_triexp_c5_pixfixed_D3")
    s = _solver(g).fit(g["b"], g["y"], pixel_fixed_params={"D3": g["fixed_D3"]})
    names = ["f1", "D1", "f2", "D2"]
    assert list(s.params_) == names
    got = np.stack([s.params_[n] for n in names])
    full = np.insert(got, 4, g["fixed_D3"], axis=0)
    ref_full = np.insert(g["params"], 4, g["fixed_D3"], axis=0)
    assert (got[0] + got[2] <= 1.0 + 1e-12).all() and s.n_active_ >= 1
    assert (_rnorm(g["b"], g["y"], full) <= _rnorm(g["b"], g["y"], ref_full) * (1 + 1e-9)).all()
    assert _polish_moves(g, full, names, {"D3": g["fixed_D3"]}, 8) < 1e-4
    # fixed f1: one free fraction, the reference applies no constraint -> plain box-bounded minimiser
    g = load("slsqp_triexp_c5_pixfixed_f1")
    s = _solver(g).fit(g["b"], g["y"], pixel_fixed_params={"f1": g["fixed_f1"]})
    names = ["D1", "f2", "D2", "D3"]
    assert list(s.params_) == names and s.n_active_ == 0
    got = np.stack([s.params_[n] for n in names])
    full = np.insert(got, 0, g["fixed_f1"], axis=0)
    ref_full = np.insert(g["params"], 0, g["fixed_f1"], axis=0)
    assert (_rnorm(g["b"], g["y"], full) <= _rnorm(g["b"], g["y"], ref_full) * (1 + 1e-9)).all()
    assert _polish_moves(g, full, names, {"f1": g["fixed_f1"]}, 8) < 1e-4


def test_device_resident_call_equals_the_host_call_and_uses_no_host_arrays():
    import torch

    g = load("slsqp_triexp_c5")
    host = _fit(g)
    s = _solver(g)
    res = s.fit_device(g["b"], torch.as_tensor(g["y"]).cuda())
    assert all(v.is_cuda for k, v in res.items() if hasattr(v, "is_cuda"))
    assert res["n_active"] == host.n_active_
    for n, row in zip(res["free_names"], res["free_rows"]):
        assert np.array_equal(res["params"][row].cpu().numpy(), host.params_[n]), n
    dev = _solver(g).fit(g["b"], torch.as_tensor(g["y"]).cuda())
    for n in NAMES:
        assert np.array_equal(dev.params_[n], host.params_[n])
    # per-call p0 / bounds arrays take the same path
    n = g["y"].shape[0]
    P0 = np.tile(g["p0"][:, None], (1, n))
    LB, UB = np.tile(g["lb"][:, None], (1, n)), np.tile(g["ub"][:, None], (1, n))
    arr = _solver(g).fit(g["b"], g["y"], p0=P0, bounds=(LB, UB))
    for n_ in NAMES:
        assert np.array_equal(arr.params_[n_], host.params_[n_])


def test_page_locked_results_from_the_second_fit_on():
    """`pinned_outputs="auto"`: the first fit of a shape downloads into fresh numpy arrays, later ones into the
    solver's page-locked block — same values, and arrays a caller still holds are never overwritten."""
    from pyneapple_b200 import synth

    cfg = synth.CONFIGS["C5"]
    b, img, _ = synth.make_volume(cfg, 0, 1)
    y = np.ascontiguousarray(img.reshape(-1, img.shape[3]))
    assert y.shape[0] >= 65536
    s = ConstrainedCurveFitSolver(model=models.TriExpModel(), max_iter=250, tol=1e-8, p0=cfg.p0, bounds=cfg.bounds,
                                  fraction_constraint=True)
    s.fit(b, y)
    first = {n: s.params_[n].copy() for n in NAMES}
    nfev = s.nfev_.copy()
    assert s._out_cache is None                    # fresh numpy arrays
    s.fit(b, y)                                    # second fit of the shape: the page-locked block
    held = s.params_["D1"]
    assert np.shares_memory(held, s._out_cache[1]["params"])
    for n in NAMES:
        assert np.array_equal(s.params_[n], first[n]), n
    assert np.array_equal(s.nfev_, nfev)
    y2 = y * 1.01
    s.fit(b, y2)                                   # `held` is still referenced: this fit must not touch it
    assert np.array_equal(held, first["D1"])
    assert not np.shares_memory(s.params_["D1"], held)


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
    assert s.diagnostics_["pcov"].shape == (4096, 5, 5)
    assert np.array_equal(s.pixel_results_.n_iterations, s.nfev_)
