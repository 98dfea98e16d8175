import sys, os
sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
import numpy as np, torch
from pyneapple_b200 import synth, models, engine, _lib
from pyneapple_b200.solvers.nnls import regularization_matrix
cfg = synth.CONFIGS["C3"]
z0 = int(sys.argv[1]) if len(sys.argv) > 1 else 0
nz = int(sys.argv[2]) if len(sys.argv) > 2 else 16
b, img, _ = synth.make_volume(cfg, z0, z0 + nz)
y = torch.as_tensor(img.reshape(-1, 16)).cuda()
model = models.NNLSModel((0.0008, 0.5), 250); B = model.get_basis(b); R = regularization_matrix(250, 2, 0.02)
ra = engine.nnls_fit(B, R, y, 250, algorithm="auto"); redo = _lib.load().pnb_nnls_last_redo_count(0)
rr = engine.nnls_fit(B, R, y, 250, algorithm="robust")
sa = ra["status"].cpu().numpy(); sr = rr["status"].cpu().numpy()
print("redo", redo, "auto status hist", np.unique(sa, return_counts=True), "robust hist", np.unique(sr, return_counts=True))
bad = np.nonzero(sa != sr)[0]
print("n differing", bad.size)
ia = ra["iterations"].cpu().numpy(); ir = rr["iterations"].cpu().numpy()
ka = (ra["coefficients"] > 0).sum(1).cpu().numpy(); kr = (rr["coefficients"] > 0).sum(1).cpu().numpy()
for v in bad[:12]:
    print('slice', v % 64 if nz == 64 else '-', end=' ')
    print(v, "auto", sa[v], ia[v], ka[v], "robust", sr[v], ir[v], kr[v])
d = (ra["coefficients"] - rr["coefficients"]).abs().amax(1).cpu().numpy()
same = sa == sr
print("max coef diff where status same:", d[same].max(), "iters differ:", int((ia[same] != ir[same]).sum()))
from oracle import c_oracle, ref_port
worst = np.argsort(d)[-12:]
A = np.concatenate([ref_port.nnls_basis(b, model.bins), ref_port.regularization_matrix(250, 2, 0.02)])
yy = y[torch.as_tensor(worst).cuda()].cpu().numpy()
ref = c_oracle.nnls(A, np.concatenate([yy, np.zeros((len(worst), 250))], 1), 250)
ca = ra["coefficients"][torch.as_tensor(worst).cuda()].cpu().numpy(); cr = rr["coefficients"][torch.as_tensor(worst).cuda()].cpu().numpy()
for i, v in enumerate(worst):
    print(v, "d(auto,robust) %.1e  d(auto,oracle) %.1e  d(robust,oracle) %.1e  iters a/r/o %d %d %d  k a/r/o %d %d %d  redo-able? rnorm a-o %.1e" % (
        d[v], np.abs(ca[i]-ref["x"][i]).max(), np.abs(cr[i]-ref["x"][i]).max(), ia[v], ir[v], ref["iters"][i], ka[v], kr[v], (ref["x"][i]>0).sum(),
        ra["residual"][int(v)].item()-ref["rnorm"][i]))
