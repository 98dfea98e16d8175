"""End-to-end C2 fit through CurveFitSolver.fit (page-locked arrays) against the pipeline chunk size (dev tool)."""
import os, sys, time
sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
import numpy as np, torch
from pyneapple_b200 import _lib, models, synth
from pyneapple_b200.solvers import CurveFitSolver, NNLSSolver

base = synth.CONFIGS["C2"]
b, img, _ = synth.make_volume(base)
y = _lib.pinned_empty((img.shape[0] * img.shape[1] * img.shape[2], 16)); y[...] = img.reshape(y.shape); del img
kw = dict(model=models.BiExpModel(fit_s0=True), p0=base.p0, bounds=base.bounds, max_iter=250, tol=1e-8, pinned_outputs=True)
QUICK = os.environ.get('PNB_PROBE_QUICK') == '1'
for chunk in ((262144,) if QUICK else (65536, 131072, 262144, 393216, 524288, 1048576, 2097152)):
    s = CurveFitSolver(chunk_vox=chunk, **kw)
    s.fit(b, y); s.fit(b, y)
    t0 = time.perf_counter()
    for _ in range(5):
        s.fit(b, y)
    dt = (time.perf_counter() - t0) / 5
    print(f"TRF chunk {chunk:8d}: {dt*1e3:7.2f} ms  {y.shape[0]/dt/1e6:7.1f} Mvox/s", flush=True)
model = models.NNLSModel(d_range=(0.0008, 0.5), n_bins=250)
for chunk in (() if QUICK else (65536, 131072, 262144, 524288)):
    s = NNLSSolver(model=model, reg_order=2, mu=0.02, max_iter=250, pinned_outputs=True, chunk_vox=chunk)
    s.fit(b, y); s.fit(b, y)
    t0 = time.perf_counter()
    s.fit(b, y)
    dt = time.perf_counter() - t0
    print(f"NNLS chunk {chunk:8d}: {dt*1e3:7.1f} ms  {y.shape[0]/dt/1e6:7.2f} Mvox/s", flush=True)

# where the end-to-end time goes: the C call alone vs solver.fit, lazy covariance vs none
import ctypes as C
from pyneapple_b200 import engine
desc = models.describe_model(models.BiExpModel(fit_s0=True)); names = list(desc.all_names)
p0 = np.array([base.p0[n] for n in names]); lb = np.array([base.bounds[n][0] for n in names]); ub = np.array([base.bounds[n][1] for n in names])
n = y.shape[0]
out = dict(params=_lib.pinned_empty((4, n)), status=_lib.pinned_empty((n,), np.int32), nfev=_lib.pinned_empty((n,), np.int32),
           njev=_lib.pinned_empty((n,), np.int32), cost=_lib.pinned_empty((n,)), r2=_lib.pinned_empty((n,)))
for wc in (False, True):
    f = lambda: engine.trf_fit(desc, b, y, p0, lb, ub, 0, jac_mode=1, want_cov=wc, out=out)
    f(); f()
    t0 = time.perf_counter()
    for _ in range(5):
        f()
    dt = (time.perf_counter() - t0) / 5
    print(f"engine.trf_fit host path, want_cov={wc}: {dt*1e3:7.2f} ms", flush=True)
