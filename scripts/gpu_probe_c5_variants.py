"""Phase 1 of the constrained fit (tri-exponential, tight tolerances, analytic Jacobian) for several builds (dev tool)."""
import glob, os, subprocess, sys
ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
CHILD = r'''
import sys, hashlib, numpy as np, torch
sys.path.insert(0, %r)
from pyneapple_b200 import engine, models, synth
cfg = synth.CONFIGS["C5"]
b, img = synth.make_volume_device(cfg, 0, 16, device=torch.device("cuda", 0))
y = img.reshape(-1, 24)
desc = models.describe_model(models.TriExpModel()); names = list(desc.all_names)
p0 = np.array([cfg.p0[n] for n in names]); lb = np.array([cfg.bounds[n][0] for n in names]); ub = np.array([cfg.bounds[n][1] for n in names])
f = lambda: engine.trf_fit(desc, b, y, p0, lb, ub, 0, max_nfev=1000, ftol=1e-13, xtol=1e-13, gtol=1e-13, jac_mode=0, want_cov=False)
for _ in range(3): r = f()
torch.cuda.synchronize()
e0, e1 = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
e0.record()
for _ in range(8): r = f()
e1.record(); torch.cuda.synchronize()
print(sys.argv[1], "ms", round(e0.elapsed_time(e1) / 8, 2), "nfev", round(float(r["nfev"].double().mean()), 3),
      "digest", hashlib.sha1(r["params"].cpu().numpy().tobytes()).hexdigest()[:12], flush=True)
'''
libs = sorted(glob.glob(os.path.join(ROOT, "pyneapple_b200", "csrc", "libpnb200_*.so")))
for lib in [os.path.join(ROOT, "pyneapple_b200", "csrc", "libpnb200.so")] + libs:
    tag = os.path.basename(lib).replace("libpnb200", "").replace(".so", "") or "_default"
    subprocess.run([sys.executable, "-c", CHILD % ROOT, tag], env=dict(os.environ, PNB_LIB=lib), check=False)
