"""Time the five BASELINE.json configurations end to end through the public API (dev tool).

Writes a markdown table to stdout; numbers are wall-clock of fit() with host (pageable numpy) inputs,
i.e. what a Pyneapple user would see after `plugin.install()`.
"""
import os, sys, time
sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
import numpy as np, torch
from pyneapple_b200 import models, synth
from pyneapple_b200.fitters import IDEALFitter, PixelWiseFitter
from pyneapple_b200.solvers import ConstrainedCurveFitSolver, CurveFitSolver, NNLSSolver

def timed(fn, reps=2):
    best = 1e9
    for _ in range(reps):
        torch.cuda.synchronize(); t = time.perf_counter(); out = fn(); torch.cuda.synchronize()
        best = min(best, time.perf_counter() - t)
    return best, out

rows = []
cfg = synth.CONFIGS["C1"]; b, img, _ = synth.make_volume(cfg)
s = CurveFitSolver(models.MonoExpModel(), p0=cfg.p0, bounds=cfg.bounds, **cfg.solver_kwargs)
t, _ = timed(lambda: s.fit(b, img.reshape(-1, 16))); rows.append(("C1 monoexp curvefit 128x128x32x16", img.size // 16, "solver.fit", t))
t, _ = timed(lambda: PixelWiseFitter(solver=s).fit(b, img)); rows.append(("C1 ... through PixelWiseFitter.fit", img.size // 16, "fitter.fit", t))
cfg = synth.CONFIGS["C2"]; b, img, _ = synth.make_volume(cfg); y = img.reshape(-1, 16)
s = CurveFitSolver(models.BiExpModel(fit_s0=True), p0=cfg.p0, bounds=cfg.bounds, **cfg.solver_kwargs)
t, _ = timed(lambda: s.fit(b, y)); rows.append(("C2 biexp(S0) curvefit 256x256x64x16", y.shape[0], "solver.fit", t))
t, _ = timed(lambda: PixelWiseFitter(solver=s).fit(b, img)); rows.append(("C2 ... through PixelWiseFitter.fit", y.shape[0], "fitter.fit", t))
n = NNLSSolver(models.NNLSModel((0.0008, 0.5), 250), reg_order=2, mu=0.02, max_iter=250)
t, _ = timed(lambda: n.fit(b, y), reps=1); rows.append(("C3 NNLS 250 bins reg 2 on the C2-shaped volume", y.shape[0], "solver.fit", t))
seg = synth.ellipsoid_mask(cfg.shape); ideal = synth.IDEAL_C4
f = IDEALFitter(s, np.array(ideal["dim_steps"]), ideal["step_tol"], segmentation_threshold=0.2)
t, _ = timed(lambda: f.fit(b, img, seg)); rows.append(("C4 IDEAL biexp 5 levels, ellipsoid mask", int(sum(f.step_pixel_counts)), "fitter.fit (all levels)", t))
del img, y
cfg = synth.CONFIGS["C5"]
c = ConstrainedCurveFitSolver(models.TriExpModel(), p0=cfg.p0, bounds=cfg.bounds, want_cov=False, **cfg.solver_kwargs)
b, img, _ = synth.make_volume(cfg, 0, 16); y = img.reshape(-1, 24)
t, _ = timed(lambda: c.fit(b, y)); rows.append(("C5 triexp constrained, one 512x512x16 slab x 24 b (1/8 of the volume)", y.shape[0], "solver.fit", t))
print("| configuration | fits | call | seconds | fits/s |\n|---|---|---|---|---|")
for name, nf, call, t in rows:
    print(f"| {name} | {nf} | {call} | {t:.3f} | {nf / t:.3e} |")
