"""conftest for running the REFERENCE's own solver / fitter tests (baseline/_ref/tests, copied there unmodified by
scripts/install_reference.py) against the B200 classes: before the test modules import anything,
``pyneapple.solvers.*`` / ``pyneapple.fitters.*`` are replaced by the plugin classes (each derives from both the B200
class and Pyneapple's class of the same name), and the registries likewise (``plugin.install()``).  Test infrastructure."""
import os
import sys
from unittest import mock

import pytest

ROOT = os.path.dirname(os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
if ROOT not in sys.path:
    sys.path.insert(0, ROOT)
from oracle import reference  # noqa: E402

reference.import_reference()
import pyneapple.fitters  # noqa: E402
import pyneapple.solvers  # noqa: E402

from pyneapple_b200 import plugin  # noqa: E402

_solvers, _fitters = plugin.plugin_classes()
for _key, _name in plugin.SOLVERS.items():
    setattr(pyneapple.solvers, _name, _solvers[_key])
for _key, _name in plugin.FITTERS.items():
    setattr(pyneapple.fitters, _name, _fitters[_key])
plugin.install()

# the spectrum post-processing functions (utility/spectrum.py) likewise: same names, same arguments, on the GPU
import pyneapple.utility.spectrum as _ref_spectrum  # noqa: E402

from pyneapple_b200 import spectrum as _our_spectrum  # noqa: E402

for _name in ("find_spectrum_peaks", "calculate_peak_area", "apply_cutoffs", "geometric_mean_peak"):
    setattr(_ref_spectrum, _name, getattr(_our_spectrum, _name))


class _Patcher:
    """The two entry points of pytest-mock's ``mocker.patch`` the reference's tests use (the plugin is not installed)."""

    def __init__(self):
        self._active = []

    def __call__(self, *a, **k):
        p = mock.patch(*a, **k)
        self._active.append(p)
        return p.start()

    def object(self, *a, **k):
        p = mock.patch.object(*a, **k)
        self._active.append(p)
        return p.start()

    def stopall(self):
        for p in reversed(self._active):
            p.stop()


class _Mocker:
    MagicMock, Mock = mock.MagicMock, mock.Mock

    def __init__(self):
        self.patch = _Patcher()


@pytest.fixture
def mocker():
    m = _Mocker()
    yield m
    m.patch.stopall()
