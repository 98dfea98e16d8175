"""Small workloads through every kernel for compute-sanitizer runs (dev tool):
    compute-sanitizer --tool memcheck|racecheck python scripts/sanitize_target.py"""
import sys, os
sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
import numpy as np, torch
from pyneapple_b200 import synth, models, engine, spectrum
from pyneapple_b200.solvers.nnls import regularization_matrix
n = int(sys.argv[1]) if len(sys.argv) > 1 else 2048
cfg = synth.CONFIGS["C3"]
b, y, _ = synth.sample_voxels(cfg, n)
y[:4] = 0.0; y[4] = -5.0; y[5, 0] = np.nan
yd = torch.as_tensor(y).cuda()
model = models.NNLSModel((0.0008, 0.5), 250)
for order in (2, 1, 3, 0):
    r = engine.nnls_fit(model.get_basis(b), regularization_matrix(250, order, 0.02), yd, 250)
    print("nnls order", order, "ok", int((r["status"] == 1).sum()), "redo", engine._lib.load().pnb_nnls_last_redo_count(0))
r = engine.nnls_fit(model.get_basis(b), regularization_matrix(250, 2, 0.02), yd, 250)
pk = spectrum.find_spectrum_peaks_batch(r["coefficients"], model.bins, 0.1, True, cutoffs=[(0.0008, 0.003), (0.003, 0.05), (0.05, 0.5)])
print("peaks", int(pk["n_peaks"].sum()))
m37 = models.NNLSModel((0.001, 0.2), 37)
r = engine.nnls_fit(m37.get_basis(b[:11]), regularization_matrix(37, 1, 0.1), yd[:, :11].contiguous(), 250)
print("nnls 37 bins ok", int((r["status"] == 1).sum()))
cfg2 = synth.CONFIGS["C2"]
b2, y2, _ = synth.sample_voxels(cfg2, n)
desc = models.describe_model(models.BiExpModel(fit_s0=True)); names = list(desc.all_names)
p0 = np.array([cfg2.p0[k] for k in names]); lb = np.array([cfg2.bounds[k][0] for k in names]); ub = np.array([cfg2.bounds[k][1] for k in names])
for method in ("trf", "dogbox"):
    r = engine.trf_fit(desc, b2, torch.as_tensor(y2).cuda(), p0, lb, ub, 0, jac_mode=1, method=method)
    print(method, "ok", int((r["status"] > 0).sum()))
img = np.random.default_rng(0).uniform(0, 1000, (24, 24, 4, 16)); seg = np.random.default_rng(1).integers(0, 5, (24, 24, 4))
print("segmeans", engine.segment_means(torch.as_tensor(img).cuda(), torch.as_tensor(seg).cuda())[1].shape)
torch.cuda.synchronize()
