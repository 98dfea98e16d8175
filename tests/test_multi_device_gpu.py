"""One-call multi-GPU host entry points (``pnb_trf_fit_host_multi`` / ``pnb_nnls_fit_host_multi``):
results are bit-identical to the single-GPU call, whatever the sharding.  On a one-GPU box the
range logic is exercised by listing device 0 several times (the ranges then run one after the other
through the same pipeline); with two or more GPUs the real concurrent path runs as well."""

from __future__ import annotations

import numpy as np
import pytest

pytestmark = pytest.mark.gpu

from pyneapple_b200 import _lib, models, synth  # noqa: E402
from pyneapple_b200.solvers import ConstrainedCurveFitSolver, CurveFitSolver, NNLSSolver  # noqa: E402


def _n_gpus():
    return _lib.load().pnb_device_count()


def _device_sets():
    sets = [[0, 0, 0]]
    if _n_gpus() >= 2:
        sets += [[0, 1], "all"]
    return sets


def _c2(n_slices=2):
    cfg = synth.CONFIGS["C2"]
    b, img, _ = synth.make_volume(cfg, 0, n_slices)
    return cfg, b, np.ascontiguousarray(img.reshape(-1, 16)[: 100_003])  # odd count: uneven ranges


@pytest.mark.parametrize("want_cov", [True, "eager", False])
def test_trf_sharded_call_equals_the_single_device_call(want_cov):
    cfg, b, y = _c2()
    kw = dict(model=models.BiExpModel(fit_s0=True), p0=cfg.p0, bounds=cfg.bounds, max_iter=250, tol=1e-8,
              want_cov=want_cov, chunk_vox=8192)
    one = CurveFitSolver(device=0, **kw).fit(b, y)
    for devs in _device_sets():
        many = CurveFitSolver(device=devs, **kw).fit(b, y)
        for n in ("f1", "D1", "D2", "S0"):
            assert np.array_equal(one.params_[n], many.params_[n]), (devs, n)
        assert np.array_equal(one.status_, many.status_) and np.array_equal(one.nfev_, many.nfev_)
        assert np.array_equal(one.r_squared_, many.r_squared_) and np.array_equal(one.cost_, many.cost_)
        if want_cov:
            assert np.array_equal(np.asarray(one.diagnostics_["pcov"]), np.asarray(many.diagnostics_["pcov"]))
        else:
            assert np.isnan(many.diagnostics_["pcov"]).all()


def test_lazy_covariance_stays_on_the_gpu_until_it_is_read():
    from pyneapple_b200._lazy import LazyArray

    cfg, b, y = _c2(1)
    kw = dict(model=models.BiExpModel(fit_s0=True), p0=cfg.p0, bounds=cfg.bounds, max_iter=250, tol=1e-8)
    lazy = CurveFitSolver(**kw).fit(b, y)
    eager = CurveFitSolver(want_cov="eager", **kw).fit(b, y)
    cov = lazy.diagnostics_["pcov"]
    assert isinstance(cov, LazyArray) and cov.on_device and cov.shape == (y.shape[0], 4, 4)
    assert isinstance(eager.diagnostics_["pcov"], np.ndarray)
    assert np.array_equal(lazy.pixel_results_[17].covariance, eager.diagnostics_["pcov"][17]) and cov.on_device
    assert np.array_equal(np.asarray(cov), eager.diagnostics_["pcov"]) and not cov.on_device


def test_per_voxel_inputs_are_sharded_with_the_voxels():
    cfg, b, y = _c2(1)
    y = y[:30_001]
    names = ["f1", "D1", "D2", "S0"]
    rng = np.random.default_rng(0)
    p0 = np.array([cfg.p0[n] for n in names])[:, None] * rng.uniform(0.8, 1.2, (4, y.shape[0]))
    lo = np.array([cfg.bounds[n][0] for n in names])[:, None]
    hi = np.array([cfg.bounds[n][1] for n in names])[:, None]
    p0 = np.clip(p0, lo, hi)
    lb, ub = np.clip(p0 * 0.5, lo, hi), np.clip(p0 * 1.5, lo, hi)
    d1 = rng.uniform(8e-4, 2e-3, y.shape[0])
    kw = dict(model=models.BiExpModel(fit_s0=True), p0=cfg.p0, bounds=cfg.bounds, max_iter=250, tol=1e-8, chunk_vox=4096)
    one = CurveFitSolver(device=0, **kw).fit(b, y, p0=p0, bounds=(lb, ub), pixel_fixed_params={"D1": d1})
    for devs in _device_sets():
        many = CurveFitSolver(device=devs, **kw).fit(b, y, p0=p0, bounds=(lb, ub), pixel_fixed_params={"D1": d1})
        for n in ("f1", "D2", "S0"):
            assert np.array_equal(one.params_[n], many.params_[n]), (devs, n)
        assert np.array_equal(np.asarray(one.diagnostics_["pcov"]), np.asarray(many.diagnostics_["pcov"]))


def test_nnls_sharded_call_equals_the_single_device_call():
    cfg = synth.CONFIGS["C3"]
    b, img, _ = synth.make_volume(cfg, 0, 1)
    y = np.ascontiguousarray(img.reshape(-1, 16)[:20_011])
    kw = dict(model=models.NNLSModel(d_range=(0.0008, 0.5), n_bins=250), reg_order=2, mu=0.02, max_iter=250,
              chunk_vox=2048)
    one = NNLSSolver(device=0, **kw).fit(b, y)
    for devs in _device_sets():
        many = NNLSSolver(device=devs, **kw).fit(b, y)
        assert np.array_equal(one.params_["coefficients"], many.params_["coefficients"]), devs
        assert np.array_equal(one.diagnostics_["residual"], many.diagnostics_["residual"])
        assert np.array_equal(one.status_, many.status_) and np.array_equal(one.iterations_, many.iterations_)


def test_constrained_solver_on_several_devices():
    cfg = synth.CONFIGS["C5"]
    b, img, _ = synth.make_volume(cfg, 0, 1)
    y = np.ascontiguousarray(img.reshape(-1, 24)[:50_001])
    kw = dict(model=models.TriExpModel(), p0=cfg.p0, bounds=cfg.bounds, want_cov=False, **cfg.solver_kwargs)
    one = ConstrainedCurveFitSolver(device=0, **kw).fit(b, y)
    for devs in _device_sets():
        many = ConstrainedCurveFitSolver(device=devs, **kw).fit(b, y)
        for n in ("f1", "D1", "f2", "D2", "D3"):
            assert np.array_equal(one.params_[n], many.params_[n]), (devs, n)
        assert one.n_active_ == many.n_active_ > 0


def test_a_bad_device_is_reported_with_its_ordinal():
    cfg, b, y = _c2(1)
    s = CurveFitSolver(model=models.BiExpModel(fit_s0=True), p0=cfg.p0, bounds=cfg.bounds, max_iter=250, tol=1e-8,
                       device=[0, 99], want_cov=False)
    with pytest.raises(_lib.EngineError, match="device 99"):
        s.fit(b, y[:1000])


def test_caller_device_is_left_alone():
    import torch

    if _n_gpus() < 2:
        pytest.skip("needs two GPUs")
    cfg, b, y = _c2(1)
    torch.cuda.set_device(0)
    CurveFitSolver(model=models.BiExpModel(fit_s0=True), p0=cfg.p0, bounds=cfg.bounds, max_iter=250, tol=1e-8,
                   device=1).fit(b, y[:5000])
    assert torch.cuda.current_device() == 0
