"""One process driving several GPUs (pnb_trf_fit_host_multi): where the time goes (dev tool)."""
import os, sys, time
sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
import numpy as np, torch
from pyneapple_b200 import _lib, models, synth
from pyneapple_b200.solvers import CurveFitSolver

n_gpu = _lib.load().pnb_device_count()
base = synth.CONFIGS["C2"]
cfg = synth.Config(**{**base.__dict__, "shape": (256, 256, 64 * max(1, min(n_gpu, 2)))})
b, img, _ = synth.make_volume(cfg)
y = _lib.pinned_empty((img.shape[0] * img.shape[1] * img.shape[2], 16)); y[...] = img.reshape(y.shape); del img
kw = dict(model=models.BiExpModel(fit_s0=True), p0=base.p0, bounds=base.bounds, max_iter=250, tol=1e-8)


def t(label, **extra):
    s = CurveFitSolver(**kw, **extra)
    s.fit(b, y); s.fit(b, y)
    t0 = time.perf_counter()
    for _ in range(3):
        s.fit(b, y)
    dt = (time.perf_counter() - t0) / 3
    print(f"{label:60s} {dt*1e3:8.1f} ms  {y.shape[0]/dt/1e6:8.1f} Mvox/s", flush=True)


print("voxels", y.shape[0], "gpus", n_gpu)
t("device 0, pinned out, lazy cov", device=0, pinned_outputs=True)
t("device 0, pinned out, no cov", device=0, pinned_outputs=True, want_cov=False)
if n_gpu >= 2:
    t("device 1, pinned out, no cov", device=1, pinned_outputs=True, want_cov=False)
    t("devices [0, 1], pinned out, no cov", device=[0, 1], pinned_outputs=True, want_cov=False)
    t("devices [0, 1], pinned out, lazy cov", device=[0, 1], pinned_outputs=True)
    t("devices [0, 1], pinned out, eager cov", device=[0, 1], pinned_outputs=True, want_cov="eager")
    t("devices [0, 1], pageable out, no cov", device=[0, 1], pinned_outputs=False, want_cov=False)
    t("devices [0, 1], chunk 65536", device=[0, 1], pinned_outputs=True, want_cov=False, chunk_vox=65536)
    t("devices [0, 1], chunk 1M", device=[0, 1], pinned_outputs=True, want_cov=False, chunk_vox=1 << 20)
