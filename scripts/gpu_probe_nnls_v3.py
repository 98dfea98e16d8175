"""v3 NNLS fast kernel against the robust kernel on a slab of config C3 (dev tool).

PNB_NNLS_NO_V3=1 runs the second-generation fast kernel instead (A/B)."""
import sys, os
sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
import numpy as np, torch
from pyneapple_b200 import synth, models, engine, _lib
from pyneapple_b200.solvers.nnls import regularization_matrix
slices = int(sys.argv[1]) if len(sys.argv) > 1 else 4
cfg = synth.CONFIGS["C3"]
b, img, _ = synth.make_volume(cfg, 0, slices)
y = torch.as_tensor(img.reshape(-1, 16)).cuda()
model = models.NNLSModel((0.0008, 0.5), 250)
B = model.get_basis(b)
orders = [int(c) for c in os.environ.get('ORDERS', '2130')]
for order in orders:
    R = regularization_matrix(250, order, 0.02)
    ref = engine.nnls_fit(B, R, y, 250, algorithm="robust") if not os.environ.get("NOREF") else None
    for rep in range(int(os.environ.get('REPS', '2'))):
        torch.cuda.synchronize(); e0 = torch.cuda.Event(enable_timing=True); e1 = torch.cuda.Event(enable_timing=True)
        e0.record(); r = engine.nnls_fit(B, R, y, 250); e1.record(); torch.cuda.synchronize()
    ms = e0.elapsed_time(e1)
    redo = _lib.load().pnb_nnls_last_redo_count(0)
    if ref is None:
        print(f"reg{order}: {y.shape[0]} vox {ms:.1f} ms -> {y.shape[0]/ms*1e3/1e6:.2f} Mvox/s redo {redo}", flush=True)
        continue
    d = (r["coefficients"] - ref["coefficients"]).abs().amax(dim=1)
    it_eq = (r["iterations"] == ref["iterations"]).double().mean().item()
    st_eq = (r["status"] == ref["status"]).double().mean().item()
    dr = (r["residual"] - ref["residual"]).abs().max().item()
    kf = (ref["coefficients"] > 0).sum(dim=1)
    print(f"   final active set: >64 {(kf > 64).sum().item()}, >60 {(kf > 60).sum().item()}, >96 {(kf > 96).sum().item()}, max {kf.max().item()}; iters >=200: {(ref['iterations'] >= 200).sum().item()}")
    print(f"reg{order}: {y.shape[0]} vox {ms:.1f} ms -> {y.shape[0]/ms*1e3/1e6:.2f} Mvox/s redo {redo}; "
          f"max|dc| {d.max().item():.2e} n(>1e-6) {(d > 1e-6).sum().item()} n(>1e-9) {(d > 1e-9).sum().item()}; "
          f"iters equal {it_eq:.5f} status equal {st_eq:.5f} max|dres| {dr:.2e}", flush=True)
