"""Host path with pageable numpy arrays (what an unmodified Pyneapple script passes) (dev tool)."""
import os, sys, time
sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
import numpy as np, torch
from pyneapple_b200 import models, synth
from pyneapple_b200.solvers import CurveFitSolver, NNLSSolver
cfg = synth.CONFIGS["C2"]; b, img, _ = synth.make_volume(cfg); y = img.reshape(-1, 16)
s = CurveFitSolver(models.BiExpModel(fit_s0=True), p0=cfg.p0, bounds=cfg.bounds, **cfg.solver_kwargs)
for rep in range(3):
    t = time.perf_counter(); s.fit(b, y); dt = time.perf_counter() - t
print(f"threads {os.environ.get('PNB_COPY_THREADS', 'default')}: C2 pageable solver.fit {dt*1e3:.1f} ms -> {y.shape[0]/dt/1e6:.1f} Mvox/s")
n = NNLSSolver(models.NNLSModel((0.0008, 0.5), 250), reg_order=2, mu=0.02, max_iter=250)
for rep in range(3):
    t = time.perf_counter(); n.fit(b, y); dt = time.perf_counter() - t
    print(f'  C3 call {rep}: {dt*1e3:.0f} ms')
print(f"threads {os.environ.get('PNB_COPY_THREADS', 'default')}: C3 pageable solver.fit {dt*1e3:.1f} ms -> {y.shape[0]/dt/1e6:.2f} Mvox/s")
