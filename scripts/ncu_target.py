"""Fixed workloads for ncu captures (dev tool):

    python scripts/ncu_target.py trf|nnls [slices] [jac_mode]

`slices` = z-slices of the C2 / C3 volume per launch (64 = the bench-size launch, 4 194 304 voxels).
Three launches of the fit (profile the last: `ncu -k regex:<kernel> -s 2 -c 1`)."""
import os
import sys

sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
import numpy as np
import torch

from pyneapple_b200 import engine, models, synth
from pyneapple_b200.solvers.nnls import regularization_matrix

what = sys.argv[1]
slices = int(sys.argv[2]) if len(sys.argv) > 2 else 64
if what == "trf":
    cfg = synth.CONFIGS["C2"]
    b, img, _ = synth.make_volume(cfg, 0, slices)
    y = torch.as_tensor(img.reshape(-1, 16)).cuda()
    desc = models.describe_model(models.BiExpModel(fit_s0=True))
    names = list(desc.all_names)
    p0 = np.array([cfg.p0[n] for n in names])
    lb = np.array([cfg.bounds[n][0] for n in names])
    ub = np.array([cfg.bounds[n][1] for n in names])
    jm = int(sys.argv[3]) if len(sys.argv) > 3 else 1
    if len(sys.argv) > 4 and sys.argv[4] == "noisy":  # heavy noise on every voxel: widely spread difficulty
        g = torch.Generator(device="cuda").manual_seed(7)
        y = y + 0.08 * y.max() * torch.randn(y.shape, generator=g, device="cuda", dtype=torch.float64)
    for _ in range(3):
        r = engine.trf_fit(desc, b, y, p0, lb, ub, 0, jac_mode=jm, want_cov="eager")
    torch.cuda.synchronize()
    print("trf ok", y.shape[0], float(r["nfev"].double().mean()))
else:
    cfg = synth.CONFIGS["C3"]
    b, img, _ = synth.make_volume(cfg, 0, slices)
    y = torch.as_tensor(img.reshape(-1, 16)).cuda()
    model = models.NNLSModel((0.0008, 0.5), 250)
    basis, R = model.get_basis(b), regularization_matrix(250, 2, 0.02)
    for _ in range(3):
        r = None
        r = engine.nnls_fit(basis, R, y, 250)
    torch.cuda.synchronize()
    print("nnls ok", y.shape[0], float(r["iterations"].double().mean()))
