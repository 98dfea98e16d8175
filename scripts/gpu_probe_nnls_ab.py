"""NNLS fast kernel A/B on the full C3 volume (dev tool): FP32 screening of the dual pass on / off
(PNB_NNLS_SCREEN, read at library load), kernel time by CUDA events, and a digest of the coefficients —
the two settings must produce bit-identical spectra."""
import os
import subprocess
import sys

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
CHILD = r'''
import sys, hashlib, numpy as np, torch
sys.path.insert(0, %r)
from pyneapple_b200 import _lib, engine, models, synth
from pyneapple_b200.solvers.nnls import regularization_matrix
slices = int(sys.argv[2])
cfg = synth.CONFIGS["C3"]
b, img, _ = synth.make_volume(cfg, 0, slices)
y = torch.as_tensor(img.reshape(-1, 16)).cuda()
model = models.NNLSModel(d_range=(0.0008, 0.5), n_bins=250)
basis, R = model.get_basis(b), regularization_matrix(250, 2, 0.02)
r = engine.nnls_fit(basis, R, y, 250); torch.cuda.synchronize()
e0, e1 = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
e0.record()
for _ in range(3):
    r = None
    r = engine.nnls_fit(basis, R, y, 250)
e1.record(); torch.cuda.synchronize()
redo = _lib.load().pnb_nnls_last_redo_count(0)
c = r["coefficients"]
digest = hashlib.sha1(c[::97].cpu().numpy().tobytes()).hexdigest()[:16]
print("screen", sys.argv[1], "ms", round(e0.elapsed_time(e1) / 3, 2), "redo", redo, "of", y.shape[0],
      "sum", float(c.sum()), "digest", digest, "iters", float(r["iterations"].double().mean()))
'''
slices = sys.argv[1] if len(sys.argv) > 1 else "64"
for z in ("0", "1"):
    env = dict(os.environ, PNB_NNLS_SCREEN=z)
    subprocess.run([sys.executable, "-c", CHILD % ROOT, z, slices], env=env, check=False)
