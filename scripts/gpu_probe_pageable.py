"""Where the pageable-input end-to-end time goes (dev tool): host staging copy bandwidth against the number of
copy threads, pnb_upload of the same bytes, and CurveFitSolver.fit on a plain numpy image."""
import os, subprocess, sys, time
ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
sys.path.insert(0, ROOT)

if len(sys.argv) > 1 and sys.argv[1] == "child":
    import numpy as np, torch
    from pyneapple_b200 import _lib, engine, models, synth
    from pyneapple_b200.solvers import CurveFitSolver
    base = synth.CONFIGS["C2"]
    b, img, _ = synth.make_volume(base)
    y = np.ascontiguousarray(img.reshape(-1, 16)); del img
    n = y.shape[0]
    pin = _lib.pinned_empty(y.shape)
    t0 = time.perf_counter(); np.copyto(pin, y); t1 = time.perf_counter()
    print(f"  numpy copy pageable->pinned (1 thread): {y.nbytes/(t1-t0)/1e9:6.2f} GB/s", flush=True)
    d = torch.empty(y.shape, dtype=torch.float64, device="cuda:0")
    for _ in range(2):
        engine.upload(d, y) if hasattr(engine, "upload") else None
    lib = _lib.load()
    import ctypes as C
    def up():
        rc = lib.pnb_upload(C.c_void_p(d.data_ptr()), C.c_void_p(y.ctypes.data), C.c_int64(y.nbytes), C.c_void_p(0))
        assert rc == 0
    up(); up()
    t0 = time.perf_counter()
    for _ in range(3): up()
    dt = (time.perf_counter() - t0) / 3
    print(f"  pnb_upload pageable 537 MB: {dt*1e3:6.1f} ms  {y.nbytes/dt/1e9:6.2f} GB/s", flush=True)
    kw = dict(model=models.BiExpModel(fit_s0=True), p0=base.p0, bounds=base.bounds, max_iter=250, tol=1e-8)
    s = CurveFitSolver(**kw)
    s.fit(b, y); s.fit(b, y)
    t0 = time.perf_counter()
    for _ in range(3): s.fit(b, y)
    dt = (time.perf_counter() - t0) / 3
    print(f"  CurveFitSolver.fit(pageable): {dt*1e3:6.1f} ms  {n/dt/1e6:6.1f} Mvox/s", flush=True)
    sys.exit(0)

print("host cores:", os.cpu_count(), flush=True)
for nt in (4, 8, 12, 16, 24, 32):
    print(f"PNB_COPY_THREADS={nt}", flush=True)
    env = dict(os.environ, PNB_COPY_THREADS=str(nt))
    subprocess.run([sys.executable, os.path.abspath(__file__), "child"], env=env, timeout=300)
