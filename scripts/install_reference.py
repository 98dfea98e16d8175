#!/usr/bin/env python
"""Put the UNMODIFIED reference where the GPU box can import it: ``baseline/_ref/pyneapple``.

    python scripts/install_reference.py            # authoring container (needs /root/reference)

``baseline/_ref/`` is git-ignored (no reference source enters the history) but travels with the
``gpurun`` snapshot, so ``bench.py --impl reference`` and ``tests/test_plugin.py`` can run the real
Pyneapple on the box's host cores.  The documented route,

    python -m pip install --no-index --no-build-isolation --no-deps --find-links /opt/wheelhouse \
        --target baseline/_ref <copy of /root/reference>

fails in this image (the build backend ``hatchling`` is not installed and there is no index to get
it from).  Pyneapple is pure Python with ``packages = ["src/pyneapple"]`` (pyproject.toml:108-109), so
the wheel that install would produce is exactly the ``src/pyneapple`` tree: the fallback copies it.
Also copied: the example TOML files the parity tests load (examples/configs, examples/parameters) and the
reference's own solver / fitter / spectrum test files (tests/test_solver_*.py, tests/test_fitter_*.py,
test_utility_spectrum.py, test_toolbox.py), which
tests/test_reference_suite.py runs against the B200 classes.
"""

from __future__ import annotations

import json
import os
import shutil
import subprocess
import sys
import tempfile
import time

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
SRC = os.environ.get("PNB_REFERENCE", "/root/reference")
DST = os.path.join(ROOT, "baseline", "_ref")


def main() -> int:
    if not os.path.isdir(os.path.join(SRC, "src", "pyneapple")):
        print(f"{SRC} is not mounted: nothing to install (the GPU box uses the prebuilt baseline/_ref)")
        return 0
    how = None
    if shutil.which("python"):
        with tempfile.TemporaryDirectory() as tmp:
            copy = os.path.join(tmp, "reference")
            shutil.copytree(SRC, copy, ignore=shutil.ignore_patterns(".git", "docs", "tests"))
            target = os.path.join(tmp, "target")
            res = subprocess.run(
                [sys.executable, "-m", "pip", "install", "--no-index", "--no-build-isolation", "--no-deps",
                 "--find-links", "/opt/wheelhouse", "--target", target, copy],
                stdout=subprocess.PIPE, stderr=subprocess.STDOUT, text=True)
            if res.returncode == 0 and os.path.isdir(os.path.join(target, "pyneapple")):
                shutil.rmtree(DST, ignore_errors=True)
                shutil.copytree(target, DST)
                how = "pip install --target"
            else:
                tail = res.stdout.strip().splitlines()[-1] if res.stdout.strip() else "?"
                print(f"pip install failed ({tail}); copying the pure-Python package tree instead")
    if how is None:
        shutil.rmtree(DST, ignore_errors=True)
        os.makedirs(DST)
        shutil.copytree(os.path.join(SRC, "src", "pyneapple"), os.path.join(DST, "pyneapple"),
                        ignore=shutil.ignore_patterns("__pycache__"))
        how = "copy of src/pyneapple (pip needs hatchling, absent here)"
    # the reference's own solver / fitter tests: tests/test_reference_suite.py runs them against the B200 classes
    tdst = os.path.join(DST, "tests")
    shutil.rmtree(tdst, ignore_errors=True)
    os.makedirs(tdst)
    for name in sorted(os.listdir(os.path.join(SRC, "tests"))):
        if name.startswith(("test_solver_", "test_fitter_")) or name in ("test_toolbox.py", "test_utility_spectrum.py"):
            shutil.copy(os.path.join(SRC, "tests", name), os.path.join(tdst, name))
    ex = os.path.join(DST, "examples")
    shutil.rmtree(ex, ignore_errors=True)
    for sub in ("configs", "parameters"):
        shutil.copytree(os.path.join(SRC, "examples", sub), os.path.join(ex, sub))
    with open(os.path.join(DST, "PROVENANCE.json"), "w") as fh:
        json.dump({"source": SRC, "how": how, "when": time.strftime("%Y-%m-%dT%H:%M:%SZ", time.gmtime())}, fh)
    print(f"reference installed in {DST} ({how})")
    return 0


if __name__ == "__main__":
    sys.exit(main())
