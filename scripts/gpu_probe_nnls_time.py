"""Device-resident C3 NNLS fit: time per launch and a digest of the spectra (dev tool)."""
import hashlib, os, sys
sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
import numpy as np, torch
from pyneapple_b200 import engine, models, synth
from pyneapple_b200.solvers.nnls import regularization_matrix
cfg = synth.CONFIGS["C3"]
b, img, _ = synth.make_volume(cfg)
y = torch.as_tensor(img.reshape(-1, 16)).cuda()
model = models.NNLSModel((0.0008, 0.5), 250)
basis, R = model.get_basis(b), regularization_matrix(250, 2, 0.02)
f = lambda: engine.nnls_fit(basis, R, y, 250)
r = f(); torch.cuda.synchronize()
n = int(sys.argv[1]) if len(sys.argv) > 1 else 3
e0, e1 = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
e0.record()
for _ in range(n): r = f()
e1.record(); torch.cuda.synchronize()
print("C3 device-resident:", round(e0.elapsed_time(e1) / n, 2), "ms  digest",
      hashlib.sha1(r["coefficients"].cpu().numpy().tobytes()).hexdigest()[:12], "(round-2 reference digest 631435c059d3)", flush=True)
