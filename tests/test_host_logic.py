"""Host-side logic: models, validation / packing, solver and fitter contracts (no GPU)."""

from __future__ import annotations

import numpy as np
import pytest

from oracle import ref_port
from pyneapple_b200 import engine, models, validation as V
from pyneapple_b200.fitters import IDEALFitter, PixelIndices, SegmentedFitter, get_fitter
from pyneapple_b200.solvers import (ConstrainedCurveFitSolver, CurveFitSolver, NNLSSolver, PixelResults,
                                    get_solver)
from pyneapple_b200.solvers.nnls import regularization_matrix

B = np.array([0, 25, 50, 75, 100, 150, 200, 300, 400, 500, 600, 700, 800, 900, 1000, 1200], float)

MODEL_CASES = [
    (models.MonoExpModel, {}, ("monoexp", "s0"), [900.0, 1.2e-3]),
    (models.BiExpModel, {}, ("biexp", "reduced"), [0.25, 1.1e-3, 0.03]),
    (models.BiExpModel, {"fit_reduced": False}, ("biexp", "full"), [0.25, 1.1e-3, 0.7, 0.03]),
    (models.BiExpModel, {"fit_s0": True}, ("biexp", "s0"), [0.25, 1.1e-3, 0.03, 950.0]),
    (models.TriExpModel, {}, ("triexp", "reduced"), [0.2, 0.1, 0.3, 0.01, 1e-3]),
    (models.TriExpModel, {"fit_reduced": False}, ("triexp", "full"), [0.2, 0.1, 0.3, 0.01, 0.5, 1e-3]),
    (models.TriExpModel, {"fit_s0": True}, ("triexp", "s0"), [0.2, 0.1, 0.3, 0.01, 1e-3, 800.0]),
]


@pytest.mark.parametrize("cls,kw,key,p", MODEL_CASES)
@pytest.mark.parametrize("t1", [None, "t1", "steam"])
def test_models_match_reference_equations(cls, kw, key, p, t1):
    extra, pp = {}, list(p)
    if t1:
        extra = dict(fit_t1=True, repetition_time=2500.0)
        if t1 == "steam":
            extra.update(fit_t1_steam=True, mixing_time=40.0)
        pp = pp + [1300.0]
    m = cls(**kw, **extra)
    ref = ref_port.Model(key[0], key[1], t1=t1, tr=2500.0, tm=40.0)
    assert m.param_names == ref.param_names == m._all_param_names
    np.testing.assert_allclose(m.forward(B, *pp), ref.forward(B, *pp), rtol=1e-10)  # tests/test_models.py:93-134
    np.testing.assert_allclose(m.jacobian(B, *pp), ref.jacobian(B, *pp), rtol=1e-10, atol=1e-300)
    # the analytic Jacobian is the derivative of forward
    eps = 1e-6
    J = m.jacobian(B, *pp)
    for j in range(len(pp)):
        q = list(pp)
        h = eps * abs(q[j])
        q[j] += h
        fd = (m.forward(B, *q) - m.forward(B, *pp)) / h
        np.testing.assert_allclose(J[:, j], fd, rtol=2e-4, atol=1e-6 * np.abs(J[:, j]).max() + 1e-12)
    desc = models.describe_model(m)
    assert desc.all_names == tuple(m._all_param_names) and desc.n_all == len(pp)
    batched = models.family_forward(desc, B, [np.array([v, v]) for v in pp])
    np.testing.assert_allclose(batched[1], m.forward(B, *pp), rtol=1e-14)


def test_model_constructor_errors_and_fixed_params():
    with pytest.raises(ValueError):
        models.BiExpModel(fit_s0=True, fit_reduced=False)
    with pytest.raises(ValueError):
        models.MonoExpModel(fit_t1=True)
    with pytest.raises(ValueError):
        models.MonoExpModel(fit_t1_steam=True, repetition_time=100.0)
    with pytest.raises(ValueError):
        models.BiExpModel(fixed_params={"nope": 1.0})
    with pytest.raises(ValueError):
        models.MonoExpModel(fixed_params={"S0": 1.0, "D": 1.0})
    m = models.BiExpModel(fit_s0=True, fixed_params={"D2": 0.03})
    assert m.param_names == ["f1", "D1", "S0"] and m._free_indices({"D2": 0.03}) == [0, 1, 3]
    full = m._inject_fixed((0.2, 1e-3, 900.0), {"D2": 0.03})
    assert full == (0.2, 1e-3, 0.03, 900.0)
    np.testing.assert_allclose(m.forward_with_fixed(B, {"D2": 0.03}, 0.2, 1e-3, 900.0), m.forward(B, *full))
    assert m.jacobian_with_fixed(B, {"D2": 0.03}, 0.2, 1e-3, 900.0).shape == (16, 3)
    nn = models.NNLSModel((0.0008, 0.5), 250)
    assert np.array_equal(nn.bins, ref_port.nnls_bins(0.0008, 0.5, 250))
    assert np.array_equal(nn.get_basis(B), ref_port.nnls_basis(B, nn.bins))
    with pytest.raises(ValueError):
        nn.get_basis(B[None, :])


def test_describe_model_duck_types_foreign_objects():
    class BiExpModel:  # looks like pyneapple.models.BiExpModel
        fit_reduced, fit_s0, fit_t1, fit_t1_steam = True, True, False, False
        repetition_time = mixing_time = None
        fixed_params = {"D2": 0.02}
        _all_param_names = ["f1", "D1", "D2", "S0"]

    d = models.describe_model(BiExpModel())
    assert d.model_id == models.MODEL_BI_S0 and d.fixed == {"D2": 0.02}
    assert engine.frozen_mask(d, {"D2"}) == 0b0100

    class Strange:
        pass

    with pytest.raises(NotImplementedError):
        models.describe_model(Strange())


def _mono(**kw):
    return CurveFitSolver(models.MonoExpModel(), 250, 1e-8, {"S0": 1000.0, "D": 1e-3},
                          {"S0": (1.0, 5000.0), "D": (1e-5, 0.1)}, **kw)


def test_curvefit_solver_constructor_contract():
    s = _mono(multi_threading=True, n_pools=4, xtol=1e-9)
    assert (s.method, s.multi_threading, s.n_pools, s.use_jacobian) == ("trf", True, 4, True)
    assert s.solver_kwargs == {"xtol": 1e-9}
    with pytest.raises(RuntimeError):
        s.get_params()
    with pytest.raises(RuntimeError):
        s.get_diagnostics()
    with pytest.raises(ValueError):  # bounds must be tuples (curvefit.py:83-89)
        CurveFitSolver(models.MonoExpModel(), 250, 1e-8, {"S0": 1.0, "D": 1e-3}, {"S0": [1, 2], "D": [0, 1]})
    with pytest.raises(ValueError):  # missing name
        CurveFitSolver(models.MonoExpModel(), 250, 1e-8, {"S0": 1.0}, {"S0": (1, 2), "D": (0, 1)})
    with pytest.raises(ValueError):  # p0 list
        CurveFitSolver(models.MonoExpModel(), 250, 1e-8, {"S0": [1.0], "D": [1e-3]}, {"S0": (1, 2), "D": (0, 1)})
    with pytest.raises(NotImplementedError):  # curve_fit kwargs without a device implementation
        _mono(check_finite=False)
    x = _mono(loss="huber", f_scale=2.0, sigma=np.array([1.0, 2.0, 4.0]), absolute_sigma=True, diff_step=1e-6)
    args = x._extras_for(["D", "S0"], list(x.model.param_names), 3)   # curve_fit's extras as engine arguments
    assert args["loss"] == "huber" and args["f_scale"] == 2.0 and args["absolute_sigma"] is True
    assert np.array_equal(args["weights"], [1.0, 0.5, 0.25]) and np.array_equal(args["diff_step"], [1e-6, 1e-6])
    assert np.array_equal(_mono(sigma=4.0)._extras_for(["D", "S0"], ["S0", "D"], 3)["weights"], [0.25] * 3)
    with pytest.raises(NotImplementedError):  # full covariance matrix of the data
        _mono(sigma=np.eye(3))._extras_for(["D", "S0"], ["S0", "D"], 3)
    with pytest.raises(NotImplementedError):
        _mono(loss=lambda z: z)._extras_for(["D", "S0"], ["S0", "D"], 3)
    assert get_solver("curvefit", model=models.MonoExpModel(), max_iter=5, tol=1e-3, p0=s.p0,
                      bounds=s.bounds).max_iter == 5
    with pytest.raises(ValueError):
        get_solver("nope")


def test_p0_and_bounds_packing_errors():
    s = _mono()
    p0, lb, ub = s._p0_and_bounds(None, None, 7)
    assert p0.shape == lb.shape == ub.shape == (2,)  # broadcast vector, not tiled
    p0, lb, ub = s._p0_and_bounds({"S0": 900.0, "D": 2e-3}, {"S0": (2.0, 10.0), "D": (0.0, 1.0)}, 7)
    assert list(p0) == [900.0, 2e-3] and list(ub) == [10.0, 1.0]
    arr = np.ones((2, 7))
    p0, lb, ub = s._p0_and_bounds(arr, (arr * 0, arr * 2), 7)
    assert p0.shape == (2, 7)
    with pytest.raises(ValueError):
        s._p0_and_bounds(np.ones((2, 6)), None, 7)
    with pytest.raises(ValueError):
        s._p0_and_bounds({"S0": np.ones(7), "D": np.ones(7)}, None, 7)
    with pytest.raises(ValueError):
        s._p0_and_bounds([1.0, 2.0], None, 7)
    with pytest.raises(ValueError):
        s._p0_and_bounds(None, (arr, [1, 2]), 7)
    with pytest.raises(ValueError):
        s._p0_and_bounds(None, (np.ones((2, 3)), np.ones((2, 3))), 7)
    with pytest.raises(ValueError):
        V.validate_data_shapes(B, np.ones((4, 15)))
    with pytest.raises(ValueError):
        V.validate_data_shapes(B[None], np.ones((4, 16)))
    seg = V.validate_segmentation(np.ones((4, 4, 2, 1)), (4, 4, 2, 16))
    assert seg.shape == (4, 4, 2)
    with pytest.raises(ValueError):
        V.validate_segmentation(np.ones((4, 4)), (4, 4, 2, 16))


def test_constrained_and_nnls_constructor_contract():
    with pytest.raises(ValueError):
        ConstrainedCurveFitSolver(models.TriExpModel(fit_reduced=False), 250, 1e-8, {}, {})
    tri = ConstrainedCurveFitSolver(
        models.TriExpModel(), 250, 1e-8, {"f1": .15, "D1": .1, "f2": .25, "D2": .01, "D3": .001},
        {"f1": (0., 1.), "D1": (.03, .5), "f2": (0., 1.), "D2": (.003, .03), "D3": (1e-4, .003)},
        use_jacobian=True, some_unknown_option=3)
    assert tri.method == "SLSQP" and tri._fraction_indices == [0, 2] and tri.fraction_constraint
    n = NNLSSolver(models.NNLSModel((0.0008, 0.5), 250), reg_order=2, mu=0.02, multi_threading=True, n_pools=8)
    assert n.n_pools == 8 and n.tol == 1e-8 and n.max_iter == 250
    assert n.get_regularization_matrix().shape == (250, 250)
    assert n._extend_signal(np.ones((3, 16))).shape == (3, 266)
    with pytest.raises(NotImplementedError):
        regularization_matrix(10, 4, 1.0)
    R = regularization_matrix(6, 2, 0.5)  # stencils of tests/test_solver_nnls.py:282-307
    assert R[2, 1] == 0.5 and R[2, 2] == -1.0 and R[2, 3] == 0.5
    band, W = engine.rtr_band(regularization_matrix(9, 3, 0.1))
    assert W == 4 and band.shape == (9, 9)
    full = regularization_matrix(9, 3, 0.1).T @ regularization_matrix(9, 3, 0.1)
    for j in range(9):
        for d in range(-4, 5):
            if 0 <= j + d < 9:
                # same terms as the BLAS product, added in a position-independent order
                assert band[j, d + 4] == pytest.approx(full[j, j + d], rel=4e-16, abs=0)
    assert engine.rtr_band(np.zeros((5, 5)))[1] == 0
    # the interior rows of the band are bit-identical (the kernel keeps them in registers only then)
    for order, W in ((1, 1), (2, 2), (3, 4)):
        band, w = engine.rtr_band(regularization_matrix(250, order, 0.0337))
        assert w == W and all(np.array_equal(band[W + 1], band[i]) for i in range(W + 1, 250 - W - 1))


def test_lazy_containers():
    pr = PixelResults(params=np.arange(6.0).reshape(3, 2), covariance=np.zeros((3, 2, 2)),
                      success=[True, False, True], messages=lambda i: f"m{i}")
    assert len(pr) == 3 and pr[1].success is False and pr[1].message == "m1" and pr[-1].params[0] == 4.0
    assert [p.success for p in pr] == [True, False, True] and len(pr[0:2]) == 2
    with pytest.raises(IndexError):
        pr[3]
    pi = PixelIndices(np.array([[0, 1, 2], [3, 4, 5]]))
    assert list(pi) == [(0, 1, 2), (3, 4, 5)] and pi[1] == (3, 4, 5) and pi == [(0, 1, 2), (3, 4, 5)]


def test_fitter_construction_contract():
    s = CurveFitSolver(models.BiExpModel(fit_s0=True), 250, 1e-8, {"f1": .2, "D1": 1e-3, "D2": .02, "S0": 1e3},
                       {"f1": (.01, .99), "D1": (1e-5, .003), "D2": (.003, .3), "S0": (1., 5e3)})
    f = get_fitter("pixelwise", solver=s)
    assert f.results_ is None and f.fitted_params_ == {} and f.pixel_indices is None
    with pytest.raises(RuntimeError):
        f.predict(B)
    with pytest.raises(ValueError):
        get_fitter("nope")
    i = IDEALFitter(s, np.array([[2, 2], [4, 4]]), {"f1": .2, "D1": .2, "D2": .2, "S0": .5})
    assert i.step_params == [] and i.ideal_dims == 2 and i.segmentation_threshold == 0.2
    with pytest.raises(ValueError):
        i._validate_fitter_inputs(np.array([[4, 4], [2, 2]]), 2)
    with pytest.raises(ValueError):
        i._validate_fitter_inputs(np.array([2, 4]), 2)
    mono = _mono()
    sf = SegmentedFitter(mono, s, step1_bvalue_range=(200, None), fixed_from_step1=["D"], param_mapping={"D": "D1"})
    xb, img = sf._subset_bvalues(B, np.ones((2, 2, 1, 16)))
    assert xb.min() == 200 and img.shape[-1] == xb.size
    with pytest.raises(ValueError):
        SegmentedFitter(mono, s, fixed_from_step1=["D"])  # "D" is not a step-2 parameter without a mapping


def test_lazy_array_behaves_like_the_array_it_becomes():
    import torch

    from pyneapple_b200._lazy import LazyArray

    full = np.arange(7 * 2 * 2, dtype=np.float64).reshape(7, 2, 2)
    parts = [(0, 3, torch.as_tensor(full[:3])), (3, 7, torch.as_tensor(full[3:]))]
    lz = LazyArray(full.shape, parts)
    assert lz.shape == (7, 2, 2) and lz.ndim == 3 and len(lz) == 7 and lz.on_device
    assert np.array_equal(lz[4], full[4]) and np.array_equal(lz[-1], full[-1]) and lz.on_device  # one block only
    with pytest.raises(IndexError):
        lz[7]
    mask = np.array([1, 0, 0, 1, 0, 0, 1], bool)
    assert np.array_equal(lz[mask], full[mask]) and not lz.on_device  # anything else materialises
    assert np.array_equal(np.asarray(lz), full) and np.isnan(lz).sum() == 0
    assert np.array_equal(lz.reshape(7, 4), full.reshape(7, 4))
    assert np.array_equal(np.asarray(LazyArray(full.shape, parts), dtype=np.float32), full.astype(np.float32))


def test_pinned_output_blocks_are_only_reused_when_nobody_else_holds_them():
    from pyneapple_b200.solvers.curvefit import _unreferenced

    cache = {"params": np.zeros((4, 10)), "status": np.zeros(10, np.int32)}
    assert _unreferenced(cache)
    row = cache["params"][1]  # a view, like solver.params_["D1"]
    assert not _unreferenced(cache)
    del row
    assert _unreferenced(cache)
    held = cache["status"]
    assert not _unreferenced(cache)
    del held


def test_page_locked_block_policy(monkeypatch):
    """`pinned_outputs="auto"`: nothing on the first fit of a shape, the cached block from the second on, a fresh
    block while a caller holds the old arrays OR A VIEW of one, and never for blocks above 1 GB.  The page-locked
    allocator is replaced by one that builds its arrays the same way (`np.ndarray(buffer=...)`) from plain memory."""
    import ctypes

    from pyneapple_b200 import _lib, models
    from pyneapple_b200.solvers import CurveFitSolver

    def fake_pinned(shape, dtype=np.float64):
        shape = tuple(int(v) for v in np.atleast_1d(shape))
        dtype = np.dtype(dtype)
        if int(np.prod(shape)) * dtype.itemsize > (1 << 26):
            raise AssertionError("the policy asked for a page-locked block it should not want")
        buf = (ctypes.c_char * max(1, int(np.prod(shape)) * dtype.itemsize))()
        return np.ndarray(shape=shape, dtype=dtype, buffer=buf)

    monkeypatch.setattr(_lib, "pinned_empty", fake_pinned)
    s = CurveFitSolver(model=models.BiExpModel(fit_s0=True), max_iter=250, tol=1e-8,
                       p0={"f1": 0.2, "D1": 1e-3, "D2": 0.02, "S0": 1000.0},
                       bounds={"f1": (0.01, 0.99), "D1": (1e-5, 3e-3), "D2": (3e-3, 0.3), "S0": (1.0, 5e3)})
    y = np.zeros((70000, 16))
    assert s._pinned_out(4, 4, 70000, y) is None                 # first fit of the shape
    blk = s._pinned_out(4, 4, 70000, y)
    assert blk is not None and blk["params"].shape == (4, 70000) and "cov" not in blk
    assert s._pinned_out(4, 4, 70000, y) is blk                  # nobody holds anything: reused
    row = blk["params"][1]                                        # what solver.params_["D1"] is
    assert row.base is blk["params"]
    fresh = s._pinned_out(4, 4, 70000, y)
    assert fresh is not blk and not np.shares_memory(fresh["params"], row)
    del row, blk, fresh
    assert s._pinned_out(4, 4, 1000, y[:1000]) is None            # small problems stay pageable
    assert s._pinned_out(4, 4, 40_000_000, np.zeros((1, 16))) is None  # first fit of that shape ...
    assert s._pinned_out(4, 4, 40_000_000, np.zeros((1, 16))) is None  # ... and 2.4 GB is above the limit of "auto"


def test_device_argument_forms():
    from pyneapple_b200 import _lib

    assert _lib.resolve_devices(3) == [3] and _lib.resolve_devices([1, 0]) == [1, 0]
    assert _lib.resolve_devices("all")[0] == 0
    with pytest.raises(ValueError):
        _lib.resolve_devices("some")
    assert _lib.shard_ranges(10, 3) == [(0, 4), (4, 7), (7, 10)]
    assert _lib.shard_ranges(2, 4) == [(0, 1), (1, 2), (2, 2), (2, 2)]


def test_status_messages_are_scipys_texts():
    from pyneapple_b200 import engine

    m = engine.status_message
    assert m(2) is None and m(4, "lm") is None
    assert m(0) == "Optimal parameters not found: The maximum number of function evaluations is exceeded."
    assert m(-3, "lm") == "array must not contain infs or NaNs"
    assert m(0, "lm", max_nfev=8) == "Optimal parameters not found: Number of calls to function has reached maxfev = 8."
    assert m(-6, "lm", ftol=1e-8).startswith("Optimal parameters not found: ftol=0.000000 is too small, no further reduction")
    assert m(-7, "lm", xtol=1.49012e-8).startswith("Optimal parameters not found: xtol=0.000000 is too small")
    assert m(-5, "lm") == "Method 'lm' only works for unconstrained problems. Use 'trf' or 'dogbox' instead."
    assert m(engine.ST_LM_TOO_FEW_DATA, "lm", n_params=3, n_data=2) == (
        "The number of func parameters=3 must not exceed the number of data points=2")


def test_small_constants_are_uploaded_once(monkeypatch):
    """Device-path launches reuse the device copies of b-values / broadcast p0 / bounds (no stream
    synchronisation per call); checked on the CPU with torch's cpu device standing in."""
    import torch

    from pyneapple_b200 import engine

    engine._CONSTS.clear()
    a = engine._small_const(np.array([1.0, 2.0, 3.0]), torch.device("cpu"))
    b = engine._small_const([1.0, 2.0, 3.0], torch.device("cpu"))
    c = engine._small_const(np.array([1.0, 2.0, 4.0]), torch.device("cpu"))
    assert a is b and c is not a and torch.equal(c, torch.tensor([1.0, 2.0, 4.0], dtype=torch.float64))
    big = engine._small_const(np.zeros(1 << 14), torch.device("cpu"))
    assert big is not engine._small_const(np.zeros(1 << 14), torch.device("cpu"))


def test_validate_p0_and_bounds_has_the_references_contract():
    """``_validate_p0_and_bounds`` is exercised by the reference's own tests (tests/test_solver_curvefit.py:
    186-262): ``(p0, (lower, upper))``, each ``(n_params, n_pixels)``."""
    s = _mono()
    p0, (lo, hi) = s._validate_p0_and_bounds(None, None, 5)
    assert p0.shape == lo.shape == hi.shape == (2, 5) and isinstance(p0, np.ndarray)
    assert (p0[0] == p0[0, 0]).all()
    arr = np.tile(np.array([900.0, 0.0012])[:, None], (1, 4))
    got, _ = s._validate_p0_and_bounds(arr, None, 4)
    assert np.array_equal(got, arr)
    with pytest.raises(ValueError, match="shape"):
        s._validate_p0_and_bounds(np.zeros((2, 3)), None, 5)
