"""Drop-in boundary against the real Pyneapple in ``baseline/_ref`` (scripts/install_reference.py):
registry replacement, unchanged TOML files, and — on the GPU — whole fits through
``load_config(...).build_fitter().fit(...)`` compared with the direct B200 call (bit for bit) and with
the reference's own fitter built from the same TOML (north_star tolerances)."""

import os

import numpy as np
import pytest

from oracle import reference

pytestmark = pytest.mark.skipif(not reference.available(),
                                reason="baseline/_ref is missing: run scripts/install_reference.py")

CFG = os.path.join(reference.EXAMPLES, "configs")
PAR = os.path.join(reference.EXAMPLES, "parameters")


@pytest.fixture()
def pyneapple_toml():
    reference.import_reference()
    import pyneapple.io.toml as t

    from pyneapple_b200 import plugin

    plugin.install()
    yield t
    plugin.uninstall()


def test_unchanged_toml_builds_b200_classes(pyneapple_toml):
    t = pyneapple_toml
    import pyneapple.fitters
    import pyneapple.solvers

    from pyneapple_b200 import fitters, solvers

    cfg = t.load_config(os.path.join(CFG, "monoexp_pixelwise.toml"))
    fitter = cfg.build_fitter()
    assert isinstance(fitter, fitters.PixelWiseFitter) and isinstance(fitter, pyneapple.fitters.PixelWiseFitter)
    assert isinstance(fitter.solver, solvers.CurveFitSolver)
    assert isinstance(fitter.solver, pyneapple.solvers.CurveFitSolver)
    assert fitter.solver.max_iter == 250 and fitter.solver.tol == 1e-8
    assert fitter.solver.p0 == {"S0": 1000.0, "D": 0.001}
    assert fitter.solver.bounds["D"] == (1e-5, 0.1)
    # the Pyneapple model object is translated to a device descriptor by duck typing
    assert fitter.solver._desc.model_id == 0 and fitter.solver._desc.all_names == ("S0", "D")

    cfg = t.load_config(os.path.join(CFG, "nnls_example.toml"))
    fitter = cfg.build_fitter()
    assert isinstance(fitter.solver, solvers.NNLSSolver) and isinstance(fitter.solver, pyneapple.solvers.NNLSSolver)
    assert fitter.solver.reg_order == 2 and fitter.solver.mu == 0.02 and fitter.solver.multi_threading is False
    A = fitter.solver._build_regularized_basis(np.linspace(0, 1000, 16))
    assert A.shape == (16 + 250, 250)

    cfg = t.load_config(os.path.join(PAR, "ideal_biexp.toml"))
    fitter = cfg.build_fitter()
    assert isinstance(fitter, fitters.IDEALFitter) and isinstance(fitter, pyneapple.fitters.IDEALFitter)
    assert fitter.dim_steps.shape == (4, 2) and fitter.step_tol["S0"] == 0.5
    assert t._SOLVER_REGISTRY["b200_curvefit"] is t._SOLVER_REGISTRY["curvefit"]
    assert pyneapple.solvers.get_solver.__module__ == "pyneapple.solvers"


def test_uninstall_restores_builtins():
    reference.import_reference()
    import pyneapple.io.toml as t
    import pyneapple.solvers as ref_solvers

    from pyneapple_b200 import plugin

    before = dict(t._SOLVER_REGISTRY)
    plugin.install()
    assert t._SOLVER_REGISTRY["curvefit"] is not before["curvefit"]
    plugin.uninstall()
    assert t._SOLVER_REGISTRY == before
    assert ref_solvers._REGISTRY["curvefit"] is ref_solvers.CurveFitSolver


def test_plugin_classes_refuse_to_run_without_a_gpu(pyneapple_toml):
    """No CPU fallback behind the plugin either: every fitter built from a TOML raises the library's
    EngineError (not a torch / CUDA-runtime error) when no device is visible."""
    from pyneapple_b200 import _lib

    if _lib.load().pnb_device_count() > 0:
        pytest.skip("a CUDA device is visible")
    t = pyneapple_toml
    b = np.linspace(0, 1000, 16)
    img = np.full((16, 16, 1, 16), 100.0)
    for path in (os.path.join(CFG, "monoexp_pixelwise.toml"), os.path.join(CFG, "nnls_example.toml")):
        with pytest.raises(_lib.EngineError):
            t.load_config(path).build_fitter().fit(b, img)
    ideal = t.load_config(os.path.join(PAR, "ideal_biexp.toml")).build_fitter()
    ideal.dim_steps = np.array([[4, 4], [8, 8], [16, 16]])
    with pytest.raises(_lib.EngineError):
        ideal.fit(b, img / 100.0)


# ----------------------------------------------------------------------------------------------
# GPU: whole fits through the unmodified TOML loader
# ----------------------------------------------------------------------------------------------
def _volume(kind, shape=(16, 16, 2)):
    from pyneapple_b200 import synth

    base = synth.CONFIGS["C1" if kind == "mono" else "C2"]
    cfg = synth.Config(**{**base.__dict__, "shape": shape})
    b, img, _ = synth.make_volume(cfg)
    if kind == "biexp_reduced":
        img = img / img[..., :1]  # S0-normalised signals for the reduced model of ideal_biexp.toml
    x = (np.arange(shape[0]) - (shape[0] - 1) / 2) / (shape[0] / 2)
    seg = ((x[:, None] ** 2 + x[None, :] ** 2) <= 0.9).astype(np.int64)
    return b, img, np.repeat(seg[:, :, None], shape[2], axis=2)


def _reference_fitter(path, edit=None):
    """The reference's own fitter from the same TOML (the plugin is not installed at this point)."""
    import pyneapple.io.toml as t

    f = t.load_config(path).build_fitter()
    if edit:
        edit(f)
    return f


@pytest.mark.gpu
def test_monoexp_pixelwise_toml_end_to_end():
    from pyneapple_b200 import fitters, models, plugin, solvers

    reference.import_reference()
    import pyneapple.io.toml as t

    path = os.path.join(CFG, "monoexp_pixelwise.toml")
    b, img, seg = _volume("mono")
    ref = _reference_fitter(path).fit(b, img, seg)
    plugin.install()
    try:
        f = t.load_config(path).build_fitter().fit(b, img, seg)
    finally:
        plugin.uninstall()
    assert isinstance(f, fitters.PixelWiseFitter)
    direct = fitters.PixelWiseFitter(solver=solvers.CurveFitSolver(
        model=models.MonoExpModel(), max_iter=250, tol=1e-8, p0={"S0": 1000.0, "D": 0.001},
        bounds={"S0": (1.0, 5000.0), "D": (1e-5, 0.1)})).fit(b, img, seg)
    for n in ("S0", "D"):
        assert np.array_equal(f.fitted_params_[n], direct.fitted_params_[n])
        rel = np.abs(f.fitted_params_[n] / ref.fitted_params_[n] - 1)
        assert rel.max() <= 1e-4, (n, rel.max())
    assert np.array_equal(f.results_.success, ref.results_.success)
    assert list(f.pixel_indices[:7]) == list(ref.pixel_indices[:7]) and len(f.pixel_indices) == len(ref.pixel_indices)
    np.testing.assert_allclose(f.results_.r_squared, ref.results_.r_squared, atol=1e-9)
    np.testing.assert_allclose(f.predict(b), ref.predict(b), rtol=1e-4)


@pytest.mark.gpu
def test_nnls_toml_end_to_end():
    from pyneapple_b200 import fitters, models, plugin, solvers

    reference.import_reference()
    import pyneapple.io.toml as t

    path = os.path.join(CFG, "nnls_example.toml")
    b, img, seg = _volume("biexp")
    ref = _reference_fitter(path).fit(b, img, seg)
    plugin.install()
    try:
        f = t.load_config(path).build_fitter().fit(b, img, seg)
    finally:
        plugin.uninstall()
    direct = fitters.PixelWiseFitter(solver=solvers.NNLSSolver(
        model=models.NNLSModel(d_range=(0.0008, 0.5), n_bins=250), reg_order=2, mu=0.02, max_iter=250)).fit(b, img, seg)
    c, cd, cr = (x.fitted_params_["coefficients"] for x in (f, direct, ref))
    assert np.array_equal(c, cd)
    assert np.abs(c - cr).max() <= 1e-6
    assert np.array_equal(f.results_.success, ref.results_.success)
    assert np.abs(f.results_.residuals - ref.results_.residuals).max() <= 1e-9 * max(1.0, ref.results_.residuals.max())


@pytest.mark.gpu
def test_ideal_toml_end_to_end():
    from pyneapple_b200 import fitters, plugin

    reference.import_reference()
    import pyneapple.io.toml as t

    path = os.path.join(PAR, "ideal_biexp.toml")
    b, img, seg = _volume("biexp_reduced")

    def small(f):  # three levels ending at the 16 x 16 test volume instead of the TOML's 128 x 128
        f.dim_steps = np.array([[4, 4], [8, 8], [16, 16]])

    ref = _reference_fitter(path, small).fit(b, img, seg)
    plugin.install()
    try:
        f = t.load_config(path).build_fitter()
        small(f)
        f.fit(b, img, seg)
    finally:
        plugin.uninstall()
    assert isinstance(f, fitters.IDEALFitter)
    assert len(f.step_params) == len(ref.step_params) == 3
    for lvl, (m, mr) in enumerate(zip(f.step_params, ref.step_params)):
        assert m.shape == mr.shape
        nz = mr != 0
        assert np.array_equal(nz, m != 0), lvl
        assert np.abs(m[nz] / mr[nz] - 1).max() <= 1e-4, lvl
    assert np.array_equal(f.results_.success, ref.results_.success)
    for n in ref.fitted_params_:
        assert np.abs(f.fitted_params_[n] / ref.fitted_params_[n] - 1).max() <= 1e-4, n
