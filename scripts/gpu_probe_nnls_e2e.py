"""NNLS end to end (page-locked arrays) against the device-resident kernel time (dev tool)."""
import os, sys, time
sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
import numpy as np, torch
from pyneapple_b200 import _lib, models, synth, engine
from pyneapple_b200.solvers import NNLSSolver

base = synth.CONFIGS["C3"]
b, img, _ = synth.make_volume(base)
y = _lib.pinned_empty((img.shape[0] * img.shape[1] * img.shape[2], 16)); y[...] = img.reshape(y.shape); del img
model = models.NNLSModel(d_range=(0.0008, 0.5), n_bins=250)
yd = torch.as_tensor(y).cuda()
s = NNLSSolver(model=model, reg_order=2, mu=0.02, max_iter=250, pinned_outputs=True)
s.fit(b, yd[:1 << 20]); torch.cuda.synchronize()
t0 = time.perf_counter(); s.fit(b, yd); torch.cuda.synchronize(); t1 = time.perf_counter()
print(f"device-resident fit + download: {1e3*(t1-t0):7.1f} ms", flush=True)
for chunk in (int(c) for c in os.environ.get("CHUNKS", "262144").split(",")):
    s = NNLSSolver(model=model, reg_order=2, mu=0.02, max_iter=250, pinned_outputs=True, chunk_vox=chunk)
    s.fit(b, y); s.fit(b, y)
    t0 = time.perf_counter()
    s.fit(b, y)
    dt = time.perf_counter() - t0
    print(f"NNLS chunk {chunk:8d}: {dt*1e3:7.1f} ms  {y.shape[0]/dt/1e6:7.2f} Mvox/s  redo {_lib.load().pnb_nnls_last_redo_count(0)}", flush=True)
