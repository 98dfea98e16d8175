"""Where the device-resident constrained fit (config C5, 16-slice slab = 4.19 M voxels) spends its time
(dev tool): phase 1 at the solver's tight tolerances vs looser ones, the torch plumbing, the face re-fit."""
import os, sys, time
sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
import numpy as np, torch
from pyneapple_b200 import engine, models, synth
from pyneapple_b200.solvers import ConstrainedCurveFitSolver

cfg = synth.CONFIGS["C5"]
dev = torch.device("cuda", 0)
b, img = synth.make_volume_device(cfg, 0, 16, device=dev)
y = img.reshape(-1, 24)
s = ConstrainedCurveFitSolver(models.TriExpModel(), p0=cfg.p0, bounds=cfg.bounds, want_cov=False, **cfg.solver_kwargs)
desc = models.describe_model(s.model)
names = list(desc.all_names)
p0 = np.array([cfg.p0[n] for n in names]); lb = np.array([cfg.bounds[n][0] for n in names]); ub = np.array([cfg.bounds[n][1] for n in names])


def timed(fn, reps=3):
    fn(); torch.cuda.synchronize()
    e0, e1 = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
    e0.record()
    for _ in range(reps):
        r = fn()
    e1.record(); torch.cuda.synchronize()
    return e0.elapsed_time(e1) / reps, r


ms, r = timed(lambda: s.fit_device(b, y))
print(f"fit_device total: {ms:.1f} ms for {y.shape[0]} voxels, active {r['n_active']}, mean nfev {float(r['nfev'].double().mean()):.2f}")
for tol in (1e-13, 1e-11, 1e-10, 1e-8):
    for jm, jn in ((0, "analytic"), (1, "2-point")):
        ms, r1 = timed(lambda: engine.trf_fit(desc, b, y, p0, lb, ub, 0, max_nfev=1000, ftol=tol, xtol=tol, gtol=tol,
                                               jac_mode=jm, want_cov=False))
        print(f"phase 1 alone, tol {tol:g}, {jn}: {ms:.1f} ms, mean nfev {float(r1['nfev'].double().mean()):.2f}, "
              f"ok {float((r1['status'] > 0).double().mean()):.6f}")
par = r1["params"]
ms, _ = timed(lambda: ((par[0] + par[2] > 1.0) & (r1["status"] > 0)).nonzero().squeeze(1))
print(f"violation mask + nonzero: {ms:.2f} ms")
idx = ((par[0] + par[2] > 1.0) & (r1["status"] > 0)).nonzero().squeeze(1)
ms, _ = timed(lambda: y.index_select(0, idx))
print(f"gather of {idx.numel()} signal rows: {ms:.2f} ms")
