"""TRF vs dogbox timing on config C2 (dev tool)."""
import sys, os
sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
import numpy as np, torch
from pyneapple_b200 import synth, models, engine
cfg = synth.CONFIGS["C2"]
b, img, _ = synth.make_volume(cfg, 0, 64)
y = torch.as_tensor(img.reshape(-1, 16)).cuda()
desc = models.describe_model(models.BiExpModel(fit_s0=True)); names = list(desc.all_names)
p0 = np.array([cfg.p0[n] for n in names]); lb = np.array([cfg.bounds[n][0] for n in names]); ub = np.array([cfg.bounds[n][1] for n in names])
for method in ("trf", "dogbox"):
    for rep in range(3):
        torch.cuda.synchronize(); e0 = torch.cuda.Event(enable_timing=True); e1 = torch.cuda.Event(enable_timing=True)
        e0.record(); r = engine.trf_fit(desc, b, y, p0, lb, ub, 0, jac_mode=1, method=method); e1.record(); torch.cuda.synchronize()
    ms = e0.elapsed_time(e1)
    print(f"{method}: {y.shape[0]} vox {ms:.2f} ms -> {y.shape[0]/ms*1e3/1e6:.1f} Mvox/s; nfev mean {r['nfev'].double().mean().item():.2f}; success {(r['status']>0).double().mean().item():.5f}")
