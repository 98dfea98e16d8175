"""TEST / BASELINE INFRASTRUCTURE ONLY: makes the UNMODIFIED reference (darksim33/Pyneapple)
importable from ``baseline/_ref`` (installed there by ``scripts/install_reference.py``; git-ignored,
shipped to the GPU box by gpurun).  Nothing under ``pyneapple_b200/`` imports this module; it is used
by ``tests/`` (parity against the real fitters / TOML loader) and by ``bench.py --impl reference`` /
the ``cpu_baseline`` leg (the reference's own multiprocessing CPU path, timed beside the GPU path).
"""

from __future__ import annotations

import os
import sys
import types

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
REF_DIR = os.path.join(ROOT, "baseline", "_ref")
EXAMPLES = os.path.join(REF_DIR, "examples")


def available() -> bool:
    return os.path.isdir(os.path.join(REF_DIR, "pyneapple"))


def import_reference(with_toml: bool = True):
    """Put ``baseline/_ref`` on ``sys.path`` and import ``pyneapple``.

    ``with_toml``: also import ``pyneapple.io.toml``; its package pulls in ``nibabel`` / ``h5py``
    (file IO back-ends that are not in this image and not on the fitting path), which are stubbed by
    empty modules first.  Returns the ``pyneapple`` module, or ``None`` when the
    reference has not been installed.
    """
    if not available():
        return None
    os.environ.setdefault("PYNEAPPLE_QUIET", "1")
    if REF_DIR not in sys.path:
        sys.path.insert(0, REF_DIR)
    if with_toml:
        for mod in ("nibabel", "h5py"):
            if mod not in sys.modules:
                try:
                    __import__(mod)
                except ImportError:
                    sys.modules[mod] = types.ModuleType(mod)
    import pyneapple

    if with_toml:
        import pyneapple.io.toml  # noqa: F401
    return pyneapple
