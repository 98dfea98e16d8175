"""Timing of the spectrum kernel and of the fused NNLS + peaks path (dev tool)."""
import sys, os, time
sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
import numpy as np, torch
from pyneapple_b200 import synth, models, engine, spectrum, _lib
from pyneapple_b200.solvers.nnls import regularization_matrix
cfg = synth.CONFIGS["C3"]
b, img, _ = synth.make_volume(cfg, 0, 16)
y = torch.as_tensor(img.reshape(-1, 16)).cuda()
model = models.NNLSModel((0.0008, 0.5), 250)
B = model.get_basis(b); R = regularization_matrix(250, 2, 0.02)
cut = [(0.0008, 0.003), (0.003, 0.05), (0.05, 0.5)]
def t(fn, reps=3):
    for _ in range(reps):
        torch.cuda.synchronize(); e0 = torch.cuda.Event(enable_timing=True); e1 = torch.cuda.Event(enable_timing=True)
        w = time.perf_counter(); e0.record(); r = fn(); e1.record(); torch.cuda.synchronize(); w = time.perf_counter() - w
    return e0.elapsed_time(e1), w * 1e3, r
ms, wall, fit = t(lambda: engine.nnls_fit(B, R, y, 250))
print(f"nnls_fit device {y.shape[0]} vox: {ms:.1f} ms (wall {wall:.1f})")
coef = fit["coefficients"]
for reg in (True, False):
    ms, wall, pk = t(lambda: spectrum.find_spectrum_peaks_batch(coef, model.bins, 0.1, reg, cutoffs=cut))
    gb = coef.numel() * 8 / 1e9
    print(f"spectrum kernel regularized={reg}: {ms:.2f} ms (wall {wall:.1f}) -> {gb/ms*1e3:.0f} GB/s of spectra; mean peaks {pk['n_peaks'].double().mean().item():.2f}")
ms, wall, _ = t(lambda: {k: v.cpu() for k, v in pk.items()})
print(f"peaks D2H: wall {wall:.1f} ms")
