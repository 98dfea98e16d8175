"""GPU probe: how many C3 voxels the fast NNLS path hands to the robust kernel, and the kernel time,
as a function of the certification threshold (PNB_NNLS_CERT_ZTOL, read at library load)."""
import os
import subprocess
import sys

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
CHILD = r'''
import sys, time, numpy as np, torch
sys.path.insert(0, %r)
from pyneapple_b200 import _lib, engine, models, synth
from pyneapple_b200.solvers.nnls import regularization_matrix
cfg = synth.CONFIGS["C3"]
b, img, _ = synth.make_volume(cfg, 0, 16)
y = torch.as_tensor(img.reshape(-1, 16)).cuda()
model = models.NNLSModel(d_range=(0.0008, 0.5), n_bins=250)
basis, R = model.get_basis(b), regularization_matrix(250, 2, 0.02)
r = engine.nnls_fit(basis, R, y, 250); torch.cuda.synchronize()
e0, e1 = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
e0.record(); r = engine.nnls_fit(basis, R, y, 250); e1.record(); torch.cuda.synchronize()
redo = _lib.load().pnb_nnls_last_redo_count(0)
print("ztol", sys.argv[1], "ms", round(e0.elapsed_time(e1), 2), "redo", redo, "of", y.shape[0])
'''
for z in ("0", "1e-9", "1e-8", "5e-8", "2e-7", "1e-6", "1e30"):
    env = dict(os.environ, PNB_NNLS_CERT_ZTOL=z)
    subprocess.run([sys.executable, "-c", CHILD % ROOT, z], env=env, check=False)
