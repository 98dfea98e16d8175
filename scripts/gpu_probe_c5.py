"""Where ConstrainedCurveFitSolver.fit (config C5, one slab) spends its wall clock (dev tool)."""
import os, sys, time, cProfile, pstats
sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
import numpy as np, torch
from pyneapple_b200 import models, synth
from pyneapple_b200.solvers import ConstrainedCurveFitSolver
cfg = synth.CONFIGS["C5"]
c = ConstrainedCurveFitSolver(models.TriExpModel(), p0=cfg.p0, bounds=cfg.bounds, want_cov=False, **cfg.solver_kwargs)
b, img, _ = synth.make_volume(cfg, 0, 16); y = img.reshape(-1, 24)
c.fit(b, y)
t = time.perf_counter(); c.fit(b, y); print(f"fit: {(time.perf_counter()-t)*1e3:.1f} ms for {y.shape[0]} voxels")
pr = cProfile.Profile(); pr.enable(); c.fit(b, y); pr.disable()
pstats.Stats(pr).sort_stats("tottime").print_stats(12)
