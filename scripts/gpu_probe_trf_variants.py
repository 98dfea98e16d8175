"""Time the C2 TRF kernel for several builds of libpnb200 (PNB_LIB) (dev tool)."""
import glob, os, subprocess, sys
ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
CHILD = r'''
import sys, hashlib, numpy as np, torch
sys.path.insert(0, %r)
from pyneapple_b200 import engine, models, synth
cfg = synth.CONFIGS["C2"]
b, img, _ = synth.make_volume(cfg, 0, int(sys.argv[2]))
y = torch.as_tensor(img.reshape(-1, 16)).cuda()
desc = models.describe_model(models.BiExpModel(fit_s0=True)); names = list(desc.all_names)
p0 = np.array([cfg.p0[n] for n in names]); lb = np.array([cfg.bounds[n][0] for n in names]); ub = np.array([cfg.bounds[n][1] for n in names])
import os
if os.environ.get("HET") == "1":
    # heterogeneous difficulty: every other voxel gets heavy noise (its fit takes more, and more varied, evaluations)
    g = torch.Generator(device="cuda").manual_seed(7)
    y[1::2] += 0.08 * y.max() * torch.randn(y[1::2].shape, generator=g, device="cuda", dtype=torch.float64)
for jm in (1, 0, 1):
    f = lambda: engine.trf_fit(desc, b, y, p0, lb, ub, 0, jac_mode=jm, want_cov="eager")
    for _ in range(5):
        r = f()
    torch.cuda.synchronize()
    e0, e1 = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
    e0.record()
    for _ in range(20):
        r = f()
    e1.record(); torch.cuda.synchronize()
    digest = hashlib.sha1(r["params"].cpu().numpy().tobytes()).hexdigest()[:12]
    print(sys.argv[1], "jac", jm, "ms", round(e0.elapsed_time(e1) / 20, 3), "nfev", round(float(r["nfev"].double().mean()), 3), "digest", digest, flush=True)
'''
slices = sys.argv[1] if len(sys.argv) > 1 else "64"
libs = sorted(glob.glob(os.path.join(ROOT, "pyneapple_b200", "csrc", "libpnb200_*.so")))
for lib in [os.path.join(ROOT, "pyneapple_b200", "csrc", "libpnb200.so")] + libs:
    env = dict(os.environ, PNB_LIB=lib)
    tag = os.path.basename(lib).replace("libpnb200", "").replace(".so", "") or "_default"
    subprocess.run([sys.executable, "-c", CHILD % ROOT, tag, slices], env=env, check=False)
