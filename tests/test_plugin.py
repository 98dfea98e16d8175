"""Drop-in boundary against the real Pyneapple (only where /root/reference is mounted)."""

import os
import sys
import types

import numpy as np
import pytest

REF = "/root/reference/src"
pytestmark = pytest.mark.skipif(not os.path.isdir(REF), reason="reference tree not mounted (GPU box)")


@pytest.fixture()
def pyneapple_toml():
    sys.path.insert(0, REF)
    for mod in ("nibabel", "h5py"):  # file IO back-ends that are not installed here
        sys.modules.setdefault(mod, types.ModuleType(mod))
    import pyneapple.io.toml as t

    from pyneapple_b200 import plugin

    plugin.install()
    yield t
    plugin.uninstall()
    sys.path.remove(REF)


def test_unchanged_toml_builds_b200_classes(pyneapple_toml):
    t = pyneapple_toml
    import pyneapple.fitters
    import pyneapple.solvers

    from pyneapple_b200 import fitters, solvers

    cfg = t.load_config("/root/reference/examples/configs/monoexp_pixelwise.toml")
    fitter = cfg.build_fitter()
    assert isinstance(fitter, fitters.PixelWiseFitter) and isinstance(fitter, pyneapple.fitters.PixelWiseFitter)
    assert isinstance(fitter.solver, solvers.CurveFitSolver)
    assert isinstance(fitter.solver, pyneapple.solvers.CurveFitSolver)
    assert fitter.solver.max_iter == 250 and fitter.solver.tol == 1e-8
    assert fitter.solver.p0 == {"S0": 1000.0, "D": 0.001}
    assert fitter.solver.bounds["D"] == (1e-5, 0.1)
    # the Pyneapple model object is translated to a device descriptor by duck typing
    assert fitter.solver._desc.model_id == 0 and fitter.solver._desc.all_names == ("S0", "D")

    cfg = t.load_config("/root/reference/examples/configs/nnls_example.toml")
    fitter = cfg.build_fitter()
    assert isinstance(fitter.solver, solvers.NNLSSolver) and isinstance(fitter.solver, pyneapple.solvers.NNLSSolver)
    assert fitter.solver.reg_order == 2 and fitter.solver.mu == 0.02 and fitter.solver.multi_threading is False
    A = fitter.solver._build_regularized_basis(np.linspace(0, 1000, 16))
    assert A.shape == (16 + 250, 250)

    cfg = t.load_config("/root/reference/examples/parameters/ideal_biexp.toml")
    fitter = cfg.build_fitter()
    assert isinstance(fitter, fitters.IDEALFitter) and isinstance(fitter, pyneapple.fitters.IDEALFitter)
    assert fitter.dim_steps.shape == (4, 2) and fitter.step_tol["S0"] == 0.5
    assert t._SOLVER_REGISTRY["b200_curvefit"] is t._SOLVER_REGISTRY["curvefit"]
    assert pyneapple.solvers.get_solver.__module__ == "pyneapple.solvers"


def test_uninstall_restores_builtins():
    sys.path.insert(0, REF)
    for mod in ("nibabel", "h5py"):
        sys.modules.setdefault(mod, types.ModuleType(mod))
    import pyneapple.io.toml as t
    import pyneapple.solvers as ref_solvers

    from pyneapple_b200 import plugin

    before = dict(t._SOLVER_REGISTRY)
    plugin.install()
    assert t._SOLVER_REGISTRY["curvefit"] is not before["curvefit"]
    plugin.uninstall()
    assert t._SOLVER_REGISTRY == before
    assert ref_solvers._REGISTRY["curvefit"] is ref_solvers.CurveFitSolver
    sys.path.remove(REF)
