"""Quick device-side timing of the TRF kernel on config C2/C1/C5-shaped samples (dev tool)."""
import sys, time, os
sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
import numpy as np, torch
from pyneapple_b200 import synth, models, engine, _lib

print("fp64 peak TFLOP/s:", _lib.measure_fp64_peak(0))
def run(cfgname, model, nrep_slices, jac_mode, want_cov=True):
    cfg = synth.CONFIGS[cfgname]
    b, img, _ = synth.make_volume(cfg, 0, nrep_slices)
    y = torch.as_tensor(img.reshape(-1, b.shape[0])).cuda()
    desc = models.describe_model(model)
    names = list(desc.all_names)
    p0 = np.array([cfg.p0[n] for n in names]); lb = np.array([cfg.bounds[n][0] for n in names]); ub = np.array([cfg.bounds[n][1] for n in names])
    for rep in range(3):
        torch.cuda.synchronize(); e0 = torch.cuda.Event(enable_timing=True); e1 = torch.cuda.Event(enable_timing=True)
        e0.record()
        r = engine.trf_fit(desc, b, y, p0, lb, ub, 0, max_nfev=250, ftol=1e-8, jac_mode=jac_mode, want_cov=want_cov)
        e1.record(); torch.cuda.synchronize()
        ms = e0.elapsed_time(e1)
    st = r["status"].cpu().numpy(); nf = r["nfev"].cpu().numpy()
    print(f"{cfgname} jac={jac_mode} cov={want_cov} n_vox={y.shape[0]} {ms:.2f} ms  {y.shape[0]/ms*1e3/1e6:.1f} Mvox/s  nfev mean {nf.mean():.2f} max {nf.max()} success {np.mean(st>0):.4f}")
    return r

run("C2", models.BiExpModel(fit_s0=True), 16, 0)
run("C2", models.BiExpModel(fit_s0=True), 16, 1)
run("C2", models.BiExpModel(fit_s0=True), 16, 0, want_cov=False)
run("C2", models.BiExpModel(fit_s0=True), 64, 1)
run("C1", models.MonoExpModel(), 32, 1)
run("C5", models.TriExpModel(), 8, 1)
run("C5", models.TriExpModel(), 8, 0)

# host path, pageable vs pinned
cfg = synth.CONFIGS["C2"]
b, img, _ = synth.make_volume(cfg, 0, 32)
y = np.ascontiguousarray(img.reshape(-1, 16))
desc = models.describe_model(models.BiExpModel(fit_s0=True)); names = list(desc.all_names)
p0 = np.array([cfg.p0[n] for n in names]); lb = np.array([cfg.bounds[n][0] for n in names]); ub = np.array([cfg.bounds[n][1] for n in names])
for label, yy in (("pageable", y),):
    for rep in range(3):
        t = time.perf_counter(); r = engine.trf_fit(desc, b, yy, p0, lb, ub, 0, jac_mode=1); dt = time.perf_counter() - t
    print(f"host path {label}: {y.shape[0]} vox in {dt*1e3:.1f} ms -> {y.shape[0]/dt/1e6:.2f} Mvox/s")
pin = _lib.pinned_empty(y.shape); pin[...] = y
n = y.shape[0]
outs = dict(params=_lib.pinned_empty((4, n)), cov=_lib.pinned_empty((n, 4, 4)), status=_lib.pinned_empty((n,), np.int32),
            nfev=_lib.pinned_empty((n,), np.int32), njev=_lib.pinned_empty((n,), np.int32), cost=_lib.pinned_empty((n,)))
for chunk in (1 << 16, 1 << 18, 1 << 20):
    for rep in range(3):
        t = time.perf_counter(); r = engine.trf_fit(desc, b, pin, p0, lb, ub, 0, jac_mode=1, out=outs, chunk_vox=chunk); dt = time.perf_counter() - t
    print(f"host path pinned chunk={chunk}: {n} vox in {dt*1e3:.1f} ms -> {n/dt/1e6:.2f} Mvox/s")
