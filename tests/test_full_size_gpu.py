"""GPU: BASELINE.json configurations at their FULL sizes, checked through size-independent
properties (the oracle cannot finish these in seconds): feasibility, optimality, idempotence,
ground-truth recovery on the seeded synthetic volumes, and agreement between independent paths."""

from __future__ import annotations

import numpy as np
import pytest

pytestmark = pytest.mark.gpu

from pyneapple_b200 import models, synth  # noqa: E402
from pyneapple_b200.fitters import IDEALFitter, PixelWiseFitter  # noqa: E402
from pyneapple_b200.solvers import ConstrainedCurveFitSolver, CurveFitSolver, NNLSSolver  # noqa: E402


def _bounds_ok(solver, cfg, names, eps=0.0):
    for n in names:
        lo, hi = cfg.bounds[n]
        v = np.asarray(solver.params_[n])
        assert (v >= lo - eps).all() and (v <= hi + eps).all(), n


def test_c1_monoexp_full_volume():
    cfg = synth.CONFIGS["C1"]
    b, img, truth = synth.make_volume(cfg)
    s = CurveFitSolver(models.MonoExpModel(), p0=cfg.p0, bounds=cfg.bounds, **cfg.solver_kwargs)
    f = PixelWiseFitter(solver=s).fit(b, img)
    r = f.results_
    assert r.n_pixels == 128 * 128 * 32 and r.success.all()
    _bounds_ok(s, cfg, ["S0", "D"])
    # SNR ~ 100: the estimates scatter around the truth
    d_err = np.abs(r.params["D"] - truth["D"].reshape(-1)) / truth["D"].reshape(-1)
    assert np.median(d_err) < 0.02 and np.nanmean(r.r_squared) > 0.99
    # idempotence: restarting from the solution stays there
    p0 = np.stack([r.params["S0"], r.params["D"]])
    s.fit(b, img.reshape(-1, 16), p0=p0)
    assert np.abs(s.params_["D"] / r.params["D"] - 1).max() < 1e-5 and s.nfev_.max() <= 4


def test_c2_biexp_full_volume_properties():
    cfg = synth.CONFIGS["C2"]
    b, img, truth = synth.make_volume(cfg)
    y = img.reshape(-1, 16)
    names = ["f1", "D1", "D2", "S0"]
    s = CurveFitSolver(models.BiExpModel(fit_s0=True), p0=cfg.p0, bounds=cfg.bounds, **cfg.solver_kwargs)
    s.fit(b, y)
    assert y.shape[0] == 256 * 256 * 64
    assert (s.status_ > 0).all() and 4 <= s.nfev_.min() and s.nfev_.max() <= 40
    _bounds_ok(s, cfg, names)
    # optimality: the cost never exceeds the cost at the start point, R^2 = 1 - 2 cost / SS_tot
    desc = models.describe_model(s.model)
    sub = slice(0, None, 997)
    pred0 = models.family_forward(desc, b, [np.full(y[sub].shape[0], cfg.p0[n]) for n in names])
    cost0 = 0.5 * ((pred0 - y[sub]) ** 2).sum(axis=1)
    assert (s.cost_[sub] <= cost0 * (1 + 1e-12)).all()
    ss_tot = ((y[sub] - y[sub].mean(axis=1, keepdims=True)) ** 2).sum(axis=1)
    assert np.abs((1 - 2 * s.cost_[sub] / ss_tot) - s.r_squared_[sub]).max() < 1e-9
    # the analytic-Jacobian path lands on the same minimiser
    a = CurveFitSolver(models.BiExpModel(fit_s0=True), p0=cfg.p0, bounds=cfg.bounds, jac="analytic",
                       want_cov=False, **cfg.solver_kwargs).fit(b, y)
    rel = np.max([np.abs(a.params_[n] / s.params_[n] - 1) for n in names], axis=0)
    assert (rel > 1e-4).mean() < 2e-4 and np.median(rel) < 1e-6
    # ground truth within noise
    assert np.median(np.abs(s.params_["D1"] - truth["D1"].reshape(-1)) / truth["D1"].reshape(-1)) < 0.05


def test_c3_nnls_full_volume_kuhn_tucker():
    cfg = synth.CONFIGS["C3"]
    b, img, _ = synth.make_volume(cfg)
    y = img.reshape(-1, 16)
    model = models.NNLSModel(d_range=(0.0008, 0.5), n_bins=250)
    s = NNLSSolver(model=model, reg_order=2, mu=0.02, max_iter=250).fit(b, y)
    x, res = s.params_["coefficients"], s.diagnostics_["residual"]
    ok = s.status_ == 1
    assert ok.mean() > 0.9999 and (x >= 0).all() and (x[~ok] == 0).all()
    # Kuhn-Tucker conditions of min ||[B; mu R] x - [y; 0]|| on a strided sample:
    # w = A^T (b - A x) <= 0 off the support, = 0 on it
    sub = np.arange(0, y.shape[0], 1009)
    sub = sub[ok[sub]]
    B = model.get_basis(b)
    R = s.get_regularization_matrix()
    w = (y[sub] - x[sub] @ B.T) @ B - x[sub] @ (R.T @ R)
    scale = np.abs(y[sub] @ B).max(axis=1, keepdims=True)
    assert (w / scale <= 1e-10).all()
    assert (np.abs(w / scale)[x[sub] > 0] <= 1e-10).all()
    r_host = np.sqrt(((y[sub] - x[sub] @ B.T) ** 2).sum(axis=1) + ((x[sub] @ R.T) ** 2).sum(axis=1))
    assert np.abs(r_host - res[sub]).max() < 1e-8
    # the (very few) voxels reported as failed fail in SciPy too: same iteration-cap rule
    from oracle import ref_port

    bad = np.where(~ok)[0][:8]
    if bad.size:
        assert not ref_port.nnls_fit(b, y[bad], (0.0008, 0.5), 250, 2, 0.02, 250)["success"].any()
    # spectrum post-processing of the whole volume (8f N4): fractions sum to one, positions are bins
    from pyneapple_b200 import spectrum

    cut = [(0.0008, 0.003), (0.003, 0.05), (0.05, 0.5)]
    pk = spectrum.find_spectrum_peaks_batch(x, model.bins, 0.1, True, cutoffs=cut, max_peaks=8)
    has = pk["n_peaks"] > 0
    assert has.mean() > 0.999 and pk["n_peaks"].max() <= 8
    assert np.abs(np.nansum(pk["f_values"][has], axis=1) - 1).max() < 1e-12
    assert np.isin(pk["d_values"][has][:, 0], model.bins).all()
    assert np.abs(np.nansum(pk["f_cut"][has], axis=1) - 1).max() < 1e-12
    sub = np.arange(0, y.shape[0], 40009)
    ref = ref_port.spectrum_peaks(x[sub], model.bins, 0.1, True, cut, 8)
    assert np.array_equal(ref["n_peaks"], pk["n_peaks"][sub])
    np.testing.assert_allclose(pk["f_values"][sub], ref["f_values"], rtol=1e-12, equal_nan=True)


def test_c2_dogbox_full_volume_agrees_with_trf():
    """method = "dogbox" on the whole C2 volume: a different path to the same minimiser."""
    cfg = synth.CONFIGS["C2"]
    b, img, _ = synth.make_volume(cfg)
    y = img.reshape(-1, 16)
    names = ["f1", "D1", "D2", "S0"]
    kw = dict(p0=cfg.p0, bounds=cfg.bounds, want_cov=False, **cfg.solver_kwargs)
    t = CurveFitSolver(models.BiExpModel(fit_s0=True), **kw).fit(b, y)
    d = CurveFitSolver(models.BiExpModel(fit_s0=True), method="dogbox", **kw).fit(b, y)
    assert (d.status_ > 0).all() and d.nfev_.max() <= 60
    _bounds_ok(d, cfg, names)
    # both stop on ftol = 1e-8: the costs agree to that order, the parameters to its square root
    assert np.abs(d.cost_ / t.cost_ - 1).max() < 1e-6
    for n in names:
        assert np.median(np.abs(d.params_[n] / t.params_[n] - 1)) < 1e-6, n
        assert np.quantile(np.abs(d.params_[n] / t.params_[n] - 1), 0.999) < 1e-3, n


def test_c4_ideal_full_volume():
    cfg = synth.CONFIGS["C4"]
    b, img, _ = synth.make_volume(cfg)
    seg = synth.ellipsoid_mask(cfg.shape)
    ideal = synth.IDEAL_C4
    s = CurveFitSolver(models.BiExpModel(fit_s0=True), p0=cfg.p0, bounds=cfg.bounds, **cfg.solver_kwargs)
    f = IDEALFitter(s, np.array(ideal["dim_steps"]), ideal["step_tol"],
                    segmentation_threshold=ideal["segmentation_threshold"]).fit(b, img, seg)
    assert [m.shape for m in f.step_params] == [(n, n, 64, 4) for n in (16, 32, 64, 128, 256)]
    # the finest level fits exactly the masked voxels (cubic resampling at identical size is the identity)
    assert f.step_pixel_counts[-1] == int((seg != 0).sum()) == f.results_.n_pixels
    assert np.array_equal(np.array(list(f.pixel_indices[:5])), np.argwhere(seg != 0)[:5])
    names = s.model.param_names
    final = f.step_params[-1]
    fitted = final[seg != 0]
    lo = np.array([cfg.bounds[n][0] for n in names])
    hi = np.array([cfg.bounds[n][1] for n in names])
    assert (fitted >= lo).all() and (fitted <= hi).all() and (final[seg == 0] == 0).all()
    assert f.results_.success.mean() > 0.999 and np.nanmean(f.results_.r_squared) > 0.99
    # every level's solution stays inside the +-tol window around the resampled previous level
    from pyneapple_b200.resize import interpolate_array

    tol = np.array([ideal["step_tol"][n] for n in names])
    p0 = np.clip(interpolate_array(f.step_params[-2], (256, 256), "cubic"), lo, hi)[seg != 0]
    okv = f.results_.success
    assert (fitted[okv] >= np.clip(p0 * (1 - tol), lo, hi)[okv] - 1e-12).all()
    assert (fitted[okv] <= np.clip(p0 * (1 + tol), lo, hi)[okv] + 1e-12).all()


def test_c5_constrained_triexp_full_volume_in_slabs():
    cfg = synth.CONFIGS["C5"]
    names = ["f1", "D1", "f2", "D2", "D3"]
    s = ConstrainedCurveFitSolver(models.TriExpModel(), p0=cfg.p0, bounds=cfg.bounds, want_cov=False,
                                  **cfg.solver_kwargs)
    desc = models.describe_model(s.model)
    n_total = n_active = 0
    for z0 in range(0, 128, 16):  # 8 slabs of 4.19 M voxels x 24 b-values (0.8 GB each)
        b, img, _ = synth.make_volume(cfg, z0, z0 + 16)
        y = img.reshape(-1, 24)
        s.fit(b, y)
        n_total += y.shape[0]
        n_active += s.n_active_
        ok = s.status_ > 0
        assert ok.mean() > 0.995
        assert (s.params_["f1"] + s.params_["f2"] <= 1.0 + 1e-12).all()
        _bounds_ok(s, cfg, names)
        sub = slice(0, None, 4099)
        pred0 = models.family_forward(desc, b, [np.full(y[sub].shape[0], cfg.p0[n]) for n in names])
        cost0 = 0.5 * ((pred0 - y[sub]) ** 2).sum(axis=1)
        pred = models.family_forward(desc, b, [s.params_[n][sub] for n in names])
        cost = 0.5 * ((pred - y[sub]) ** 2).sum(axis=1)
        assert (cost[ok[sub]] <= cost0[ok[sub]] * (1 + 1e-12)).all()
    assert n_total == 512 * 512 * 128
    assert 0.005 < n_active / n_total < 0.06  # the constraint is active on the planted sub-population
