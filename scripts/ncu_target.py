"""Small fixed workloads for `ncu --set full` captures (dev tool): python scripts/ncu_target.py trf|nnls [jac]"""
import sys, os
sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
import numpy as np, torch
from pyneapple_b200 import synth, models, engine
from pyneapple_b200.solvers.nnls import regularization_matrix
what = sys.argv[1]
if what == "trf":
    cfg = synth.CONFIGS["C2"]
    b, img, _ = synth.make_volume(cfg, 0, 16)
    y = torch.as_tensor(img.reshape(-1, 16)).cuda()
    desc = models.describe_model(models.BiExpModel(fit_s0=True)); names = list(desc.all_names)
    p0 = np.array([cfg.p0[n] for n in names]); lb = np.array([cfg.bounds[n][0] for n in names]); ub = np.array([cfg.bounds[n][1] for n in names])
    jm = int(sys.argv[2]) if len(sys.argv) > 2 else 1
    for _ in range(3):
        r = engine.trf_fit(desc, b, y, p0, lb, ub, 0, jac_mode=jm)
    torch.cuda.synchronize(); print("trf ok", float(r["nfev"].double().mean()))
else:
    cfg = synth.CONFIGS["C3"]
    b, img, _ = synth.make_volume(cfg, 0, 2)
    y = torch.as_tensor(img.reshape(-1, 16)).cuda()
    model = models.NNLSModel((0.0008, 0.5), 250)
    for _ in range(2):
        r = engine.nnls_fit(model.get_basis(b), regularization_matrix(250, 2, 0.02), y, 250)
    torch.cuda.synchronize(); print("nnls ok", float(r["iterations"].double().mean()))
