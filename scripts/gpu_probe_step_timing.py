"""Why do 10 back-to-back device-path fits take longer per step than 5?  (dev tool)"""
import os, sys, time
sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
import numpy as np, torch
import bench
from pyneapple_b200 import engine, models, synth

cfg = synth.CONFIGS["C2"]
b, img, _ = synth.make_volume(cfg, 0, 64)
y = torch.as_tensor(img.reshape(-1, 16)).cuda()
desc = models.describe_model(models.BiExpModel(fit_s0=True)); names = list(desc.all_names)
p0 = np.array([cfg.p0[n] for n in names]); lb = np.array([cfg.bounds[n][0] for n in names]); ub = np.array([cfg.bounds[n][1] for n in names])
f = lambda: engine.trf_fit(desc, b, y, p0, lb, ub, 0, max_nfev=250, ftol=1e-8, jac_mode=1, want_cov="eager")

def loop(n, tag):
    torch.cuda.synchronize()
    e0, e1 = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
    t0 = time.perf_counter(); e0.record()
    for _ in range(n):
        r = f()
    t1 = time.perf_counter(); e1.record(); torch.cuda.synchronize()
    print(f"{tag}: {n} steps, {e0.elapsed_time(e1)/n:7.3f} ms/step on the device, host enqueue {1e3*(t1-t0)/n:6.3f} ms/step", flush=True)

for _ in range(3): f()
for rep in range(3): loop(10, "no sampler")
loop(30, "no sampler")
s = bench.ClockSampler(0, 1); s.start()
for rep in range(4): loop(10, "nvidia-smi -lms 200")
loop(30, "nvidia-smi -lms 200")
print(s.stop())
for rep in range(2): loop(10, "after sampler")
