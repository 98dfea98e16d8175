"""Heterogeneous difficulty (heavy noise on every other voxel): where does the time go?  (dev tool)"""
import os, sys
sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
import numpy as np, torch
from pyneapple_b200 import engine, models, synth
cfg = synth.CONFIGS["C2"]
b, img, _ = synth.make_volume(cfg, 0, 64)
y = torch.as_tensor(img.reshape(-1, 16)).cuda()
g = torch.Generator(device="cuda").manual_seed(7)
noise = 0.08 * y.max() * torch.randn(y[1::2].shape, generator=g, device="cuda", dtype=torch.float64)
desc = models.describe_model(models.BiExpModel(fit_s0=True)); names = list(desc.all_names)
p0 = np.array([cfg.p0[n] for n in names]); lb = np.array([cfg.bounds[n][0] for n in names]); ub = np.array([cfg.bounds[n][1] for n in names])
def run(tag, yy, **kw):
    f = lambda: engine.trf_fit(desc, b, yy, p0, lb, ub, 0, jac_mode=1, want_cov="eager", **kw)
    for _ in range(3): r = f()
    torch.cuda.synchronize()
    e0, e1 = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
    e0.record()
    for _ in range(5): r = f()
    e1.record(); torch.cuda.synchronize()
    nf = r["nfev"].double()
    q = torch.quantile(nf[::16], torch.tensor([0.5, 0.9, 0.99, 0.999], device="cuda", dtype=torch.float64)).tolist()
    at_bound = ((r["params"] <= torch.as_tensor(lb, device="cuda")[:, None] * (1 + 1e-9)) | (r["params"] >= torch.as_tensor(ub, device="cuda")[:, None] * (1 - 1e-9))).any(dim=0).double().mean().item()
    print(f"{tag}: {e0.elapsed_time(e1)/5:7.2f} ms  mean nfev {nf.mean().item():6.2f}  p50/p90/p99/p99.9 {q}  max {int(nf.max())}  failed {(r['status']<=0).double().mean().item():.4f}  at a bound {at_bound:.3f}", flush=True)
run("clean            ", y)
yh = y.clone(); yh[1::2] += noise
run("noisy every other", yh)
yn = y.clone(); yn[: y.shape[0] // 2] = yh[1::2][: y.shape[0] // 2]  # the same voxels, noisy half first (contiguous)
run("noisy first half ", yn)
run("noisy, max_nfev 50", yh, max_nfev=50)
run("all noisy        ", (y + 0.08 * y.max() * torch.randn(y.shape, generator=g, device="cuda", dtype=torch.float64)))
