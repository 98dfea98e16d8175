"""Time the NNLS fast kernel for several builds of libpnb200 (PNB_LIB) x screening on / off (dev tool)."""
import glob, os, subprocess, sys
ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
CHILD = r'''
import sys, hashlib, numpy as np, torch
sys.path.insert(0, %r)
from pyneapple_b200 import _lib, engine, models, synth
from pyneapple_b200.solvers.nnls import regularization_matrix
cfg = synth.CONFIGS["C3"]
b, img, _ = synth.make_volume(cfg, 0, int(sys.argv[2]))
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
digest = hashlib.sha1(r["coefficients"][::97].cpu().numpy().tobytes()).hexdigest()[:12]
print(sys.argv[1], "ms", round(e0.elapsed_time(e1) / 3, 2), "digest", digest, flush=True)
'''
slices = sys.argv[1] if len(sys.argv) > 1 else "16"
libs = sorted(glob.glob(os.path.join(ROOT, "pyneapple_b200", "csrc", "libpnb200_*.so")))
for lib in [os.path.join(ROOT, "pyneapple_b200", "csrc", "libpnb200.so")] + libs:
    for scr in ("0", "1"):
        env = dict(os.environ, PNB_LIB=lib, PNB_NNLS_SCREEN=scr)
        tag = os.path.basename(lib).replace("libpnb200", "").replace(".so", "") or "_default"
        subprocess.run([sys.executable, "-c", CHILD % ROOT, f"{tag} screen={scr}", slices], env=env, check=False)
