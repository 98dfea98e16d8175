"""The C-ABI library builds, loads and exports every symbol the header declares (no GPU needed)."""

import ctypes
import os

import pytest

from pyneapple_b200 import _lib


@pytest.fixture(scope="module")
def lib():
    if not os.path.exists(_lib.SO_PATH):
        _lib.build()
    return _lib.load()


def test_exports_every_declared_symbol(lib):
    names = _lib.exported_symbols()
    assert "pnb_trf_fit_host" in names and "pnb_trf_fit_device" in names
    for n in names:
        assert hasattr(lib, n), f"{n} declared in include/pyneapple_b200.h but not exported"


def test_abi_version(lib):
    assert lib.pnb_abi_version() == 4


def test_struct_layout_matches_header(lib):
    # sizeof(pnb_trf_problem) as compiled by nvcc must equal the ctypes mirror
    assert ctypes.sizeof(_lib.TrfProblem) == lib.pnb_sizeof_trf_problem()
    assert ctypes.sizeof(_lib.NnlsProblem) == lib.pnb_sizeof_nnls_problem()
    assert ctypes.sizeof(_lib.SpectrumProblem) == lib.pnb_sizeof_spectrum_problem()


def test_bad_arguments_are_rejected_without_a_device(lib):
    prob = _lib.TrfProblem()
    prob.model_id = 99
    rc = lib.pnb_trf_fit_host(ctypes.byref(prob), 0, 0)
    assert rc == -2
    assert b"model" in lib.pnb_last_error()


def test_curve_fit_extras_are_validated_by_the_library(lib):
    """`loss` outside PNB_LOSS_*, extras with a method / T1 mode they are not built for, `lm` with a robust loss:
    rejected by the argument check (before any CUDA call)."""
    def problem(**kw):
        prob = _lib.TrfProblem()
        prob.model_id, prob.n_params, prob.n_b, prob.n_vox, prob.max_nfev, prob.jac_mode = 3, 4, 16, 0, 250, 1
        for k, v in kw.items():
            setattr(prob, k, v)
        return prob

    assert lib.pnb_trf_fit_host(ctypes.byref(problem()), 0, 0) == 0          # n_vox = 0: nothing to do
    assert lib.pnb_trf_fit_host(ctypes.byref(problem(loss=2, f_scale=1.5)), 0, 0) == 0
    assert lib.pnb_trf_fit_host(ctypes.byref(problem(loss=9)), 0, 0) == -1
    assert b"PNB_LOSS" in lib.pnb_last_error()
    assert lib.pnb_trf_fit_host(ctypes.byref(problem(loss=2, method=1)), 0, 0) == -2      # dogbox + robust loss
    assert b"method trf" in lib.pnb_last_error()
    t1 = problem(loss=1, t1_mode=1, n_params=5)
    assert lib.pnb_trf_fit_host(ctypes.byref(t1), 0, 0) == -2
    step = problem(method=1)
    step.diff_step[0] = 1e-6
    assert lib.pnb_trf_fit_host(ctypes.byref(step), 0, 0) == -2


def test_no_cpu_fallback():
    import numpy as np
    from pyneapple_b200 import models
    from pyneapple_b200.solvers import CurveFitSolver

    if _lib.load().pnb_device_count() > 0:
        pytest.skip("a GPU is present")
    s = CurveFitSolver(models.MonoExpModel(), 250, 1e-8, {"S0": 1000.0, "D": 1e-3},
                       {"S0": (1.0, 5000.0), "D": (1e-5, 0.1)})
    b = np.linspace(0, 1000, 8)
    with pytest.raises(_lib.EngineError):
        s.fit(b, 1000 * np.exp(-b * 1e-3))


def test_no_cpu_fallback_for_the_new_entry_points():
    """Spectrum post-processing, segment means, staged transfers: they raise without a device too."""
    import numpy as np
    from pyneapple_b200 import engine, spectrum

    if _lib.load().pnb_device_count() > 0:
        pytest.skip("a GPU is present")
    x = np.zeros((4, 250))
    x[:, 100] = 1.0
    with pytest.raises(_lib.EngineError):
        spectrum.find_spectrum_peaks_batch(x, np.geomspace(8e-4, 0.5, 250))
    with pytest.raises(_lib.EngineError):
        spectrum.find_spectrum_peaks(x[0], np.geomspace(8e-4, 0.5, 250))
    with pytest.raises(_lib.EngineError):
        spectrum.apply_cutoffs([1e-3], [1.0], [(1e-4, 1e-2)])
    with pytest.raises(_lib.EngineError):
        engine.segment_means(np.zeros((4, 4, 2, 8)), np.zeros((4, 4, 2), int))


def test_bench_refuses_to_run_without_a_gpu():
    import os
    import subprocess
    import sys

    if _lib.load().pnb_device_count() > 0:
        pytest.skip("a GPU is present")
    root = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
    r = subprocess.run([sys.executable, os.path.join(root, "bench.py"), "--steps", "1", "--warmup", "0"],
                       capture_output=True, text=True, timeout=300)
    assert r.returncode != 0 and r.stdout.strip() == ""
    assert "no CPU fallback" in r.stderr


def test_spectrum_bad_arguments_are_rejected_without_a_device(lib):
    prob = _lib.SpectrumProblem()
    prob.max_peaks = 64
    assert lib.pnb_spectrum_peaks_host(ctypes.byref(prob), 0, 0) == -1
    assert b"max_peaks" in lib.pnb_last_error()
    prob.max_peaks, prob.n_vox, prob.detect, prob.n_bins = 8, 4, 1, 250
    assert lib.pnb_spectrum_peaks_host(ctypes.byref(prob), 0, 0) == -1
    assert b"spectrum" in lib.pnb_last_error()
