"""Where the IDEAL fitter (config C4) spends its wall clock (dev tool)."""
import os, sys, time, cProfile, pstats
sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
import numpy as np, torch
from pyneapple_b200 import models, synth
from pyneapple_b200.fitters import IDEALFitter
from pyneapple_b200.solvers import CurveFitSolver
cfg = synth.CONFIGS["C2"]; b, img, _ = synth.make_volume(cfg)
s = CurveFitSolver(models.BiExpModel(fit_s0=True), p0=cfg.p0, bounds=cfg.bounds, **cfg.solver_kwargs)
seg = synth.ellipsoid_mask(cfg.shape); ideal = synth.IDEAL_C4
f = IDEALFitter(s, np.array(ideal["dim_steps"]), ideal["step_tol"], segmentation_threshold=0.2)
f.fit(b, img, seg)
torch.cuda.synchronize()
t = time.perf_counter(); f.fit(b, img, seg); torch.cuda.synchronize(); print(f"fit: {(time.perf_counter()-t)*1e3:.1f} ms, fits {sum(f.step_pixel_counts)}")
pr = cProfile.Profile(); pr.enable(); f.fit(b, img, seg); torch.cuda.synchronize(); pr.disable()
pstats.Stats(pr).sort_stats("tottime").print_stats(14)
