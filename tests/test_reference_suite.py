"""The reference's OWN solver / fitter test files (baseline/_ref/tests, unmodified) run against the B200 classes.

`tests/refsuite/refsuite_conftest.py` (copied next to them as conftest.py) swaps `pyneapple.solvers.*` / `pyneapple.fitters.*` for the plugin classes before the
reference's test modules import them, so every `CurveFitSolver(...)`, `PixelWiseFitter(...)` … in those files is the
GPU implementation.  On a box without a GPU only the tests that never fit (constructors, validation, error
behaviour) can pass — every other one must fail with the library's `EngineError` and nothing else; on the GPU the
whole selection must pass, except the listed tests that pin CPU-only internals of the reference."""

import os
import re
import shutil
import subprocess
import sys

import pytest

from oracle import reference

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
SUITE = os.path.join(reference.REF_DIR, "tests")
pytestmark = pytest.mark.skipif(not os.path.isdir(SUITE), reason="baseline/_ref/tests is missing: run scripts/install_reference.py")

# Tests of the reference that assert its CPU internals rather than the solver / fitter contract — they patch a
# SciPy function inside Pyneapple's modules (`curve_fit`, `minimize`, `nnls`) and expect the solver to run into the
# patch, or patch the per-voxel hook and expect the batch loop to call it once per voxel.  The GPU solvers make no
# SciPy call and have no per-voxel loop.  Everything else — including the tests that call the internal hooks
# `_fit_data` / `_fit_single_pixel` directly, which the GPU classes implement with the reference's signatures — runs.
CPU_INTERNALS = [
    "test_failed_fit_returns_p0_and_nan_cov",           # curvefit / constrained: scipy curve_fit / minimize patched to raise
    "test_failed_fit_returns_zeros_and_norm_residual",  # nnls: scipy nnls patched to raise
    "test_failed_pixel_produces_nan_in_popt",           # `_fit_single_pixel` patched on the instance
]


def _run(tmp_path, extra):
    work = tmp_path / "refsuite"
    work.mkdir()
    for name in os.listdir(SUITE):
        shutil.copy(os.path.join(SUITE, name), work / name)
    shutil.copy(os.path.join(ROOT, "tests", "refsuite", "refsuite_conftest.py"), work / "conftest.py")
    deselect = " and ".join(f"not {k}" for k in CPU_INTERNALS)
    cmd = [sys.executable, "-m", "pytest", "-c", os.devnull, "-p", "no:cacheprovider", "-q", "-k", deselect,
           "--rootdir", str(work), str(work)] + extra
    env = dict(os.environ, PYNEAPPLE_QUIET="1", PYTHONPATH=ROOT)
    return subprocess.run(cmd, capture_output=True, text=True, timeout=1500, cwd=str(work), env=env)


def _counts(out):
    m = re.search(r"(?:(\d+) failed, )?(\d+) passed", out)
    return (int(m.group(1) or 0), int(m.group(2))) if m else (None, None)


def test_without_a_gpu_everything_that_fits_fails_with_engine_error_only(tmp_path):
    from pyneapple_b200 import _lib

    if _lib.load().pnb_device_count() > 0:
        pytest.skip("a CUDA device is visible")
    r = _run(tmp_path, ["--tb=line"])
    failed, passed = _counts(r.stdout)
    assert passed is not None and passed >= 150, r.stdout[-2000:]
    reasons = [ln for ln in r.stdout.splitlines() if ln.startswith("/") or ln.startswith("E ")]
    other = [ln for ln in reasons if "EngineError" not in ln]
    assert not other, other[:10]


@pytest.mark.gpu
def test_the_references_own_tests_pass_on_the_gpu(tmp_path):
    r = _run(tmp_path, ["--tb=short"])
    failed, passed = _counts(r.stdout)
    assert failed == 0 and passed and passed >= 290, r.stdout[-6000:]
