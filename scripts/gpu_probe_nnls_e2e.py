"""NNLS host pipeline (pinned buffers) against the chunk size (dev tool)."""
import sys, os, time
sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
import numpy as np, torch
from pyneapple_b200 import synth, models, _lib
from pyneapple_b200.solvers import NNLSSolver
cfg = synth.CONFIGS["C3"]
b, img, _ = synth.make_volume(cfg, 0, 64)
y = img.reshape(-1, 16)
pin = _lib.pinned_empty(y.shape); pin[...] = y
model = models.NNLSModel(d_range=(0.0008, 0.5), n_bins=250)
for chunk in (0, 1 << 16, 1 << 17, 1 << 18, 1 << 19, 1 << 20):
    s = NNLSSolver(model=model, reg_order=2, mu=0.02, max_iter=250, pinned_outputs=True, chunk_vox=chunk)
    s.fit(b, pin)
    t = time.perf_counter(); s.fit(b, pin); dt = time.perf_counter() - t
    print(f"chunk {chunk}: {dt*1e3:.1f} ms -> {y.shape[0]/dt/1e6:.2f} Mvox/s", flush=True)
    del s
s = NNLSSolver(model=model, reg_order=2, mu=0.02, max_iter=250)
cut = [(0.0008, 0.003), (0.003, 0.05), (0.05, 0.5)]
t = time.perf_counter(); s.fit_spectrum_peaks(b, pin, cutoffs=cut); print(f'first call (page-locks the result arrays): {(time.perf_counter()-t)*1e3:.1f} ms')
t = time.perf_counter(); r = s.fit_spectrum_peaks(b, pin, cutoffs=cut); dt = time.perf_counter() - t
print(f"fit_spectrum_peaks (spectra stay on the device): {dt*1e3:.1f} ms -> {y.shape[0]/dt/1e6:.2f} Mvox/s; mean peaks {r['n_peaks'].mean():.2f}")
from pyneapple_b200 import spectrum
