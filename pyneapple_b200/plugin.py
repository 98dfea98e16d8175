"""Hooking the B200 engine into an installed Pyneapple (the drop-in boundary).

Two routes:

1. **entry points** (non-invasive, the documented plugin route of
   ``io/toml.py:106-148``): this package declares ``b200_curvefit``,
   ``b200_constrained_curvefit``, ``b200_nnls`` in group ``pyneapple.solvers`` and
   ``b200_pixelwise``, ``b200_ideal``, ``b200_segmented``, ``b200_segmentationwise``
   in ``pyneapple.fitters`` (``pyproject.toml``); a TOML then says
   ``type = "b200_curvefit"``.
2. :func:`install` — *replace* the built-in keys (``curvefit``,
   ``constrained_curvefit``, ``nnls``; ``pixelwise``, ``ideal``, ``segmented``,
   ``segmentationwise``) in Pyneapple's registries so that an **unchanged** TOML
   runs on the GPU.  Direct registry insertion is what Pyneapple's own test does
   (``tests/test_io_plugin_discovery.py:131``).  :func:`uninstall` restores them.

The registered classes are subclasses of both the B200 class and the Pyneapple
class of the same name, so ``isinstance`` checks inside Pyneapple
(``fitters/base.py:164``, ``io/toml.py:203``) keep working.
"""

from __future__ import annotations

_SAVED: dict = {}

SOLVERS = {"curvefit": "CurveFitSolver", "constrained_curvefit": "ConstrainedCurveFitSolver",
           "nnls": "NNLSSolver"}
FITTERS = {"pixelwise": "PixelWiseFitter", "segmentationwise": "SegmentationWiseFitter",
           "ideal": "IDEALFitter", "segmented": "SegmentedFitter"}


def _mixed(ours: type, theirs: type) -> type:
    if issubclass(theirs, ours):  # somebody already put a mixed class where Pyneapple's was
        return theirs
    return type(ours.__name__, (ours, theirs), {"__module__": ours.__module__, "__doc__": ours.__doc__})


_CLASSES: tuple | None = None


def plugin_classes() -> tuple[dict, dict]:
    """B200 solver / fitter classes, mixed with Pyneapple's when it is importable (built once)."""
    global _CLASSES
    if _CLASSES is not None:
        return _CLASSES
    from . import fitters as our_fitters
    from . import solvers as our_solvers

    try:
        import pyneapple.fitters as ref_fitters
        import pyneapple.solvers as ref_solvers
    except ImportError:
        ref_fitters = ref_solvers = None
    s = {}
    for key, name in SOLVERS.items():
        cls = getattr(our_solvers, name)
        s[key] = _mixed(cls, getattr(ref_solvers, name)) if ref_solvers else cls
    f = {}
    for key, name in FITTERS.items():
        cls = getattr(our_fitters, name)
        f[key] = _mixed(cls, getattr(ref_fitters, name)) if ref_fitters else cls
    if ref_solvers is not None:
        _adopt_reference_records()
        _CLASSES = (s, f)
    return s, f


def _adopt_reference_records() -> None:
    """Per-voxel records handed out by the B200 solvers become instances of Pyneapple's own
    ``_PixelFitResult`` (same fields), so ``isinstance`` checks in code written against the reference hold."""
    from pyneapple.solvers import base as ref_base

    from .solvers import base as our_base
    from .solvers import curvefit as our_curvefit
    from .solvers import nnls as our_nnls

    for mod in (our_base, our_curvefit, our_nnls):
        mod._PixelFitResult = ref_base._PixelFitResult


def _registries():
    import pyneapple.fitters as ref_fitters
    import pyneapple.solvers as ref_solvers

    regs = {"solvers": [ref_solvers._REGISTRY], "fitters": [ref_fitters._REGISTRY]}
    try:  # the TOML loader keeps its own registries; it needs nibabel / h5py to import
        import pyneapple.io.toml as ref_toml

        regs["solvers"].append(ref_toml._SOLVER_REGISTRY)
        regs["fitters"].append(ref_toml._FITTER_REGISTRY)
    except ImportError:
        pass
    return regs


def install(replace_builtin: bool = True) -> None:
    """Register the B200 classes in Pyneapple's registries (see module docstring)."""
    solvers, fitters = plugin_classes()
    regs = _registries()
    for kind, classes in (("solvers", solvers), ("fitters", fitters)):
        for reg in regs[kind]:
            for key, cls in classes.items():
                reg["b200_" + key] = cls
                if replace_builtin:
                    _SAVED.setdefault((id(reg), key), (reg, reg.get(key)))
                    reg[key] = cls


def uninstall() -> None:
    """Undo :func:`install`."""
    for (_, key), (reg, old) in list(_SAVED.items()):
        if old is None:
            reg.pop(key, None)
        else:
            reg[key] = old
        reg.pop("b200_" + key, None)
    _SAVED.clear()
