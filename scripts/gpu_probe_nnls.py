"""Quick device-side timing of the NNLS kernel on config C3 (dev tool)."""
import sys, time, os
sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
import numpy as np, torch
from pyneapple_b200 import synth, models, engine
from pyneapple_b200.solvers.nnls import regularization_matrix
cfg = synth.CONFIGS["C3"]
b, img, _ = synth.make_volume(cfg, 0, 4)
y = torch.as_tensor(img.reshape(-1, 16)).cuda()
model = models.NNLSModel((0.0008, 0.5), 250)
B = model.get_basis(b)
for order in (2, 0, 1, 3):
    R = regularization_matrix(250, order, 0.02)
    for rep in range(2):
        torch.cuda.synchronize(); e0 = torch.cuda.Event(enable_timing=True); e1 = torch.cuda.Event(enable_timing=True)
        e0.record(); r = engine.nnls_fit(B, R, y, 250); e1.record(); torch.cuda.synchronize()
    ms = e0.elapsed_time(e1)
    it = r["iterations"].cpu().numpy(); st = r["status"].cpu().numpy(); k = (r["coefficients"] > 0).sum(1).cpu().numpy()
    from pyneapple_b200 import _lib
    redo = _lib.load().pnb_nnls_last_redo_count(0)
    print(f"reg{order}: redo {redo} {y.shape[0]} vox {ms:.1f} ms -> {y.shape[0]/ms*1e3/1e6:.2f} Mvox/s; iters mean {it.mean():.1f} max {it.max()}; active mean {k.mean():.1f} max {k.max()}; ok {np.mean(st==1):.4f}")
