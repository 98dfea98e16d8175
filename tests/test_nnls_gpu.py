"""GPU parity: CUDA NNLS path vs the reference's golden outputs and the C restatement."""

from __future__ import annotations

import numpy as np
import pytest

from _util import load

pytestmark = pytest.mark.gpu

from pyneapple_b200 import models  # noqa: E402
from pyneapple_b200.solvers import NNLSSolver  # noqa: E402

CASES = ["nnls_c3_reg2", "nnls_c3_reg0", "nnls_c3_reg1", "nnls_c3_reg3", "nnls_c3_degenerate",
         "nnls_c3_maxiter5", "nnls_c3_maxiter20", "nnls_small_reg1"]


def _solver(g, **kw):
    model = models.NNLSModel(d_range=tuple(float(v) for v in g["d_range"]), n_bins=int(g["n_bins"]))
    return NNLSSolver(model=model, reg_order=int(g["reg_order"]), mu=float(g["mu"]),
                      max_iter=int(g["max_iter"]), **kw)


@pytest.mark.parametrize("name", CASES)
@pytest.mark.parametrize("algorithm", ["auto", "robust"])
def test_golden_parity(name, algorithm):
    g = load(name)
    s = _solver(g, algorithm=algorithm).fit(g["b"], g["y"])
    coef, res = s.params_["coefficients"], s.diagnostics_["residual"]
    success = np.array([pr.success for pr in s.pixel_results_])
    assert (success == g["success"]).all()
    assert coef.shape == g["coefficients"].shape
    assert (coef >= 0).all()
    # north_star tolerance: coefficients within 1e-6 absolute
    assert np.abs(coef - g["coefficients"]).max() <= 1e-6
    assert (np.abs(res - g["residual"]) <= 1e-9 * np.maximum(1.0, g["residual"])).all()
    fail = ~g["success"]
    if fail.any():
        assert (coef[fail] == 0).all()


def test_iteration_counts_match_lawson_hanson():
    """Same active-set path as the classical algorithm: identical iteration counts."""
    from oracle import c_oracle, ref_port

    g = load("nnls_c3_reg2")
    s = _solver(g).fit(g["b"], g["y"])
    bins = ref_port.nnls_bins(0.0008, 0.5, 250)
    A = np.concatenate([ref_port.nnls_basis(g["b"], bins), ref_port.regularization_matrix(250, 2, 0.02)])
    Bx = np.concatenate([g["y"], np.zeros((g["y"].shape[0], 250))], axis=1)
    ref = c_oracle.nnls(A, Bx, 250)
    assert (s.iterations_ == ref["iters"]).mean() > 0.99


def test_against_c_oracle_large():
    from oracle import c_oracle, ref_port
    from pyneapple_b200 import synth

    b, y, _ = synth.sample_voxels(synth.CONFIGS["C3"], 4096, z=7)
    model = models.NNLSModel(d_range=(0.0008, 0.5), n_bins=250)
    s = NNLSSolver(model=model, reg_order=2, mu=0.02, max_iter=250).fit(b, y)
    A = np.concatenate([ref_port.nnls_basis(b, model.bins), ref_port.regularization_matrix(250, 2, 0.02)])
    Bx = np.concatenate([y, np.zeros((y.shape[0], 250))], axis=1)
    ref = c_oracle.nnls(A, Bx, 250)
    assert ((s.status_ == 1) == (ref["status"] == 1)).all()
    assert np.abs(s.params_["coefficients"] - ref["x"]).max() <= 1e-6
    assert np.abs(s.diagnostics_["residual"] - ref["rnorm"]).max() <= 1e-8


def test_large_active_set_overflows_to_global_scratch():
    """Strong regularisation -> broad spectra -> active sets beyond the shared-memory factor."""
    from oracle import c_oracle, ref_port
    from pyneapple_b200 import synth

    b, y, _ = synth.sample_voxels(synth.CONFIGS["C3"], 64, z=3)
    model = models.NNLSModel(d_range=(0.0008, 0.5), n_bins=250)
    s = NNLSSolver(model=model, reg_order=2, mu=3.0, max_iter=750).fit(b, y)
    A = np.concatenate([ref_port.nnls_basis(b, model.bins), ref_port.regularization_matrix(250, 2, 3.0)])
    Bx = np.concatenate([y, np.zeros((y.shape[0], 250))], axis=1)
    ref = c_oracle.nnls(A, Bx, 750)
    assert (ref["x"] > 0).sum(axis=1).max() > 64  # the case does exercise the overflow path
    assert ((s.status_ == 1) == (ref["status"] == 1)).all()
    assert np.abs(s.params_["coefficients"] - ref["x"]).max() <= 1e-6


@pytest.mark.parametrize("order,mu", [(2, 1e-3), (2, 2e-4), (1, 2e-4), (2, 1e-5)])
def test_weak_regularisation_falls_back_to_the_robust_path(order, mu):
    """The inverse-update fast path certifies its own result; weakly regularised voxels are
    re-solved by the Cholesky kernel, so the 1e-6 tolerance holds for any mu."""
    from oracle import c_oracle, ref_port
    from pyneapple_b200 import synth

    b, y, _ = synth.sample_voxels(synth.CONFIGS["C3"], 256, z=11)
    model = models.NNLSModel(d_range=(0.0008, 0.5), n_bins=250)
    s = NNLSSolver(model=model, reg_order=order, mu=mu, max_iter=750).fit(b, y)
    A = np.concatenate([ref_port.nnls_basis(b, model.bins), ref_port.regularization_matrix(250, order, mu)])
    Bx = np.concatenate([y, np.zeros((y.shape[0], 250))], axis=1)
    ref = c_oracle.nnls(A, Bx, 750)
    assert ((s.status_ == 1) == (ref["status"] == 1)).all()
    assert np.abs(s.params_["coefficients"] - ref["x"]).max() <= 1e-6


@pytest.mark.parametrize("n_b", [7, 24, 30, 40])
def test_other_measurement_counts(n_b):
    """Fast-path specialisations for <= 8 / 16 / 24 / 32 measurements, robust path beyond."""
    from oracle import c_oracle, ref_port

    rng = np.random.default_rng(n_b)
    b = np.sort(rng.uniform(0, 1500, n_b))
    b[0] = 0.0
    n = 192
    f = rng.uniform(0.05, 0.4, n)[:, None]
    y = 1000 * (f * np.exp(-b * rng.uniform(0.01, 0.1, n)[:, None]) +
                (1 - f) * np.exp(-b * rng.uniform(5e-4, 2.5e-3, n)[:, None])) + rng.normal(0, 10, (n, n_b))
    model = models.NNLSModel(d_range=(0.0005, 0.3), n_bins=120)
    s = NNLSSolver(model=model, reg_order=2, mu=0.05, max_iter=360).fit(b, y)
    A = np.concatenate([ref_port.nnls_basis(b, model.bins), ref_port.regularization_matrix(120, 2, 0.05)])
    ref = c_oracle.nnls(A, np.concatenate([y, np.zeros((n, 120))], axis=1), 360)
    assert ((s.status_ == 1) == (ref["status"] == 1)).all()
    assert np.abs(s.params_["coefficients"] - ref["x"]).max() <= 1e-6
    assert np.abs(s.diagnostics_["residual"] - ref["rnorm"]).max() <= 1e-8


@pytest.mark.parametrize("n_bins", [9, 33, 64, 127, 249, 256, 257, 300])
@pytest.mark.parametrize("order", [1, 2, 3])
def test_other_dictionary_sizes(n_bins, order):
    """The fast kernel owns 8 consecutive bins per lane and handles the first / last W bins (and the
    bins beyond n_bins) separately: every size around its limits (256 bins, then the previous
    generation takes over) against the C restatement of Lawson-Hanson."""
    from oracle import c_oracle, ref_port

    rng = np.random.default_rng(1000 * order + n_bins)
    b = np.array([0, 25, 50, 75, 100, 150, 200, 300, 400, 500, 600, 700, 800, 900, 1000, 1200], float)
    n = 96
    f = rng.uniform(0.05, 0.4, n)[:, None]
    y = 1000 * (f * np.exp(-b * rng.uniform(0.01, 0.1, n)[:, None]) +
                (1 - f) * np.exp(-b * rng.uniform(5e-4, 2.5e-3, n)[:, None])) + rng.normal(0, 10, (n, 16))
    y[0] = 0.0
    y[1, 3] = np.inf
    model = models.NNLSModel(d_range=(0.0007, 0.4), n_bins=n_bins)
    s = NNLSSolver(model=model, reg_order=order, mu=0.03, max_iter=3 * n_bins).fit(b, y)
    A = np.concatenate([ref_port.nnls_basis(b, model.bins), ref_port.regularization_matrix(n_bins, order, 0.03)])
    fin = np.isfinite(y).all(axis=1)
    ref = c_oracle.nnls(A, np.concatenate([y[fin], np.zeros((fin.sum(), n_bins))], axis=1), 3 * n_bins)
    assert ((s.status_[fin] == 1) == (ref["status"] == 1)).all()
    assert (s.status_[~fin] != 1).all() and (s.params_["coefficients"][~fin] == 0).all()
    assert np.abs(s.params_["coefficients"][fin] - ref["x"]).max() <= 1e-6
    assert np.abs(s.diagnostics_["residual"][fin] - ref["rnorm"]).max() <= 1e-8


def test_chunked_host_pipeline_equals_single_launch():
    """Many small chunks on alternating streams (each with its own scratch / hand-over list)."""
    from pyneapple_b200 import synth

    b, y, _ = synth.sample_voxels(synth.CONFIGS["C3"], 8192, z=21)
    model = models.NNLSModel(d_range=(0.0008, 0.5), n_bins=250)
    one = NNLSSolver(model=model, reg_order=2, mu=0.02, max_iter=250).fit(b, y)
    many = NNLSSolver(model=model, reg_order=2, mu=0.02, max_iter=250, chunk_vox=257).fit(b, y)
    assert (one.status_ == 1).all() and (many.status_ == 1).all()
    assert np.array_equal(one.params_["coefficients"], many.params_["coefficients"])
    assert np.array_equal(one.iterations_, many.iterations_)


def test_host_pipeline_hands_over_once_per_call():
    """The host pipeline collects the voxels its chunks' fast kernels cannot certify in ONE list and
    re-solves them after the last chunk from a compact copy of their signals: same bits as the device
    path (hand-over inside the launch), for pageable and page-locked arrays, over ragged chunks."""
    import torch

    from pyneapple_b200 import _lib, synth

    b, y, _ = synth.sample_voxels(synth.CONFIGS["C3"], 6000, z=5)
    model = models.NNLSModel(d_range=(0.0008, 0.5), n_bins=250)
    kw = dict(model=model, reg_order=2, mu=1e-3, max_iter=750)   # weak regularisation: many hand-overs
    dev = NNLSSolver(**kw).fit(b, torch.as_tensor(y).cuda())
    ref_coef, ref_it, ref_res = dev.params_["coefficients"].copy(), dev.iterations_.copy(), dev.diagnostics_["residual"].copy()
    ypin = _lib.pinned_empty(y.shape); ypin[...] = y
    for sig, pinned in ((y, False), (ypin, True)):
        s = NNLSSolver(chunk_vox=701, pinned_outputs=pinned, **kw).fit(b, sig)
        n_redo = _lib.load().pnb_nnls_last_redo_count(0)
        assert 0 < n_redo < 6000
        assert np.array_equal(s.params_["coefficients"], ref_coef)
        assert np.array_equal(s.iterations_, ref_it) and np.array_equal(s.status_, dev.status_)
        assert np.array_equal(s.diagnostics_["residual"], ref_res)
        assert np.array_equal(s.r_squared_, dev.r_squared_, equal_nan=True)


def test_page_locked_spectra_are_never_overwritten_while_a_caller_holds_them():
    from pyneapple_b200 import synth

    b, y, _ = synth.sample_voxels(synth.CONFIGS["C3"], 16384, z=9)
    model = models.NNLSModel(d_range=(0.0008, 0.5), n_bins=250)
    s = NNLSSolver(model=model, reg_order=2, mu=0.02, max_iter=250)
    first = s.fit(b, y).params_["coefficients"].copy()
    s.fit(b, y)                                   # second fit of the shape: the page-locked block
    block = s._out_cache[1]["coefficients"]
    held = s.params_["coefficients"][:100]        # a view
    assert np.shares_memory(held, block)
    s.fit(b, 2.0 * y)
    assert np.array_equal(held, first[:100]) and not np.shares_memory(s.params_["coefficients"], block)
    assert np.allclose(s.params_["coefficients"], 2.0 * first, rtol=1e-9, atol=1e-9)
    # a device-resident signal downloads into the page-locked block too
    import torch

    del held
    s2 = NNLSSolver(model=model, reg_order=2, mu=0.02, max_iter=250)
    yd = torch.as_tensor(y).cuda()
    s2.fit(b, yd)
    s2.fit(b, yd)
    assert np.shares_memory(s2.params_["coefficients"], s2._out_cache[1]["coefficients"])
    assert np.array_equal(s2.params_["coefficients"], first) and np.array_equal(s2.status_, np.ones(16384, np.int32))


def test_single_voxel_and_device_path():
    import torch

    g = load("nnls_c3_reg2")
    s = _solver(g)
    s.fit(g["b"], g["y"][0])
    assert s.params_["coefficients"].shape == (1, 250)
    host = s.fit(g["b"], g["y"]).params_["coefficients"].copy()
    dev = s.fit(g["b"], torch.as_tensor(g["y"]).cuda()).params_["coefficients"]
    assert np.array_equal(host, dev)


def test_c3_full_volume_default_path_equals_robust_path_on_every_voxel():
    """`algorithm="auto"` (inverse-update fast kernel + certification + hand-over) against
    `algorithm="robust"` (Cholesky + refinement, the path pinned to SciPy by the goldens, iteration
    counts included) on ALL 4 194 304 voxels of config C3 — not a strided sample: coefficients within
    north_star's 1e-6 absolute, the same support above that tolerance, the same failures."""
    import torch

    from pyneapple_b200 import engine, models, synth
    from pyneapple_b200.solvers.nnls import regularization_matrix

    cfg = synth.CONFIGS["C3"]
    b, img, _ = synth.make_volume(cfg)
    y = torch.as_tensor(img.reshape(-1, 16)).cuda()
    del img
    model = models.NNLSModel(d_range=(0.0008, 0.5), n_bins=250)
    basis, R = model.get_basis(b), regularization_matrix(250, 2, 0.02)
    worst, n_support, n_status = 0.0, 0, 0
    step = 1 << 20
    for s0 in range(0, y.shape[0], step):  # 2 x 2.1 GB of spectra at a time
        fa = engine.nnls_fit(basis, R, y[s0:s0 + step], 250, algorithm="auto")
        fr = engine.nnls_fit(basis, R, y[s0:s0 + step], 250, algorithm="robust")
        d = (fa["coefficients"] - fr["coefficients"]).abs()
        worst = max(worst, float(d.max()))
        n_support += int((((fa["coefficients"] > 1e-6) != (fr["coefficients"] > 1e-6)) & (d > 1e-6)).sum())
        n_status += int((fa["status"] != fr["status"]).sum())
        assert float((fa["residual"] - fr["residual"]).abs().max()) <= 1e-9 * max(1.0, float(fr["residual"].max()))
        del fa, fr, d
    assert n_status == 0
    assert worst <= 1e-6, f"max |coef(auto) - coef(robust)| = {worst:.3e}"
    assert n_support == 0


def test_dual_gemm_on_the_fp64_tensor_cores_equals_the_matrix_product():
    """``pnb_nnls_dual_gemm_device`` (mma.sync m8n8k4.f64): H0 = Y B for odd sizes, ragged tiles and
    every supported measurement count."""
    import torch

    from pyneapple_b200 import engine

    rng = np.random.default_rng(5)
    for n_vox, m, n in ((1, 16, 250), (7, 16, 250), (4099, 16, 250), (1000, 11, 37), (513, 24, 96), (64, 32, 255), (9, 3, 5)):
        B = rng.normal(size=(m, n))
        Y = rng.normal(size=(n_vox, m)) * 100.0
        got = engine.nnls_dual_gemm(B, torch.as_tensor(Y).cuda()).cpu().numpy()
        want = Y @ B
        assert got.shape == want.shape
        scale = np.abs(Y) @ np.abs(B)
        assert (np.abs(got - want) <= 4e-16 * m * scale).all(), (n_vox, m, n)


def test_materialised_dual_gives_the_same_fits_as_the_fused_form():
    g = load("nnls_c3_reg2")
    from pyneapple_b200 import models
    from pyneapple_b200.solvers import NNLSSolver

    model = models.NNLSModel(d_range=tuple(g["d_range"]), n_bins=int(g["n_bins"]))
    kw = dict(model=model, reg_order=2, mu=0.02, max_iter=250)
    fused = NNLSSolver(**kw).fit(g["b"], g["y"])
    gemm = NNLSSolver(dual_init="gemm", **kw).fit(g["b"], g["y"])
    assert np.array_equal(fused.status_, gemm.status_) and np.array_equal(fused.iterations_, gemm.iterations_)
    assert np.abs(fused.params_["coefficients"] - gemm.params_["coefficients"]).max() <= 1e-9
    assert np.abs(gemm.params_["coefficients"] - g["coefficients"]).max() <= 1e-6
