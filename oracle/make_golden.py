"""Generate ``tests/golden/*.npz`` by running the REAL reference (authoring container only).

    PYNEAPPLE_QUIET=1 python oracle/make_golden.py [case ...]

Imports Pyneapple from ``/root/reference/src`` (read-only mount; it does not
exist on the GPU box), runs its own solvers / fitters on seeded synthetic
inputs from ``pyneapple_b200.synth`` and stores inputs + outputs.  The
fixtures are the parity pins for the oracle port (``oracle/ref_port.py``),
the C restatement (``oracle/pnb_oracle.c``) and the CUDA path.
"""

from __future__ import annotations

import os
import sys
import time

os.environ.setdefault("PYNEAPPLE_QUIET", "1")
ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
sys.path.insert(0, ROOT)
sys.path.insert(0, "/root/reference/src")

import numpy as np  # noqa: E402

from pyneapple.models import BiExpModel, MonoExpModel, NNLSModel, TriExpModel  # noqa: E402
from pyneapple.solvers import (  # noqa: E402
    ConstrainedCurveFitSolver,
    CurveFitSolver,
    NNLSSolver,
)

from pyneapple_b200 import synth  # noqa: E402

GOLD = os.path.join(ROOT, "tests", "golden")
MODELS = {"monoexp": MonoExpModel, "biexp": BiExpModel, "triexp": TriExpModel}


def _save(name, **arrays):
    path = os.path.join(GOLD, name + ".npz")
    np.savez_compressed(path, **arrays)
    print(f"  wrote {path} ({os.path.getsize(path) / 1024:.0f} KiB)")


def _curvefit_outputs(solver, n_free):
    prs = solver.pixel_results_
    return dict(
        params=np.array([pr.params for pr in prs]).T,
        pcov=np.array([pr.covariance for pr in prs]),
        success=np.array([pr.success for pr in prs]),
        messages=np.array([pr.message or "" for pr in prs]),
    )


def _run_curvefit(name, model_kind, model_kwargs, b, y, p0, bounds, pixel_fixed=None,
                  p0_arr=None, bounds_arr=None, cls=CurveFitSolver, **solver_kwargs):
    model = MODELS[model_kind](**model_kwargs)
    kw = dict(max_iter=250, tol=1e-8)
    kw.update(solver_kwargs)
    solver = cls(model=model, p0=p0, bounds=bounds, **kw)
    t = time.time()
    solver.fit(b, y, p0=p0_arr, bounds=bounds_arr, pixel_fixed_params=pixel_fixed)
    dt = time.time() - t
    out = _curvefit_outputs(solver, None)
    if cls is ConstrainedCurveFitSolver:
        out["nit"] = np.array(
            [pr.n_iterations if pr.n_iterations is not None else -1 for pr in solver.pixel_results_]
        )
    print(f"  {name}: {y.shape[0]} voxels in {dt:.1f}s, success {out['success'].mean():.3f}")
    extra = {}
    if pixel_fixed:
        for k, v in pixel_fixed.items():
            extra["fixed_" + k] = v
    if p0_arr is not None:
        extra.update(p0_arr=p0_arr, lb_arr=bounds_arr[0], ub_arr=bounds_arr[1])
    _save(
        name, b=b, y=y, param_names=np.array(model.param_names),
        all_names=np.array(model._all_param_names),
        p0=np.array([p0[n] for n in model.param_names], float),
        lb=np.array([bounds[n][0] for n in model.param_names], float),
        ub=np.array([bounds[n][1] for n in model.param_names], float),
        max_iter=kw["max_iter"], tol=kw["tol"], **out, **extra,
    )
    return solver


def _cfg_args(c):
    cfg = synth.CONFIGS[c]
    return cfg, cfg.p0, cfg.bounds


def case_mono_c1():
    cfg, p0, bounds = _cfg_args("C1")
    b, y, _ = synth.sample_voxels(cfg, 512)
    _run_curvefit("trf_mono_c1", "monoexp", {}, b, y, p0, bounds)


def case_biexp_c2():
    cfg, p0, bounds = _cfg_args("C2")
    b, y, _ = synth.sample_voxels(cfg, 2048)
    _run_curvefit("trf_biexp_s0_c2", "biexp", {"fit_s0": True}, b, y, p0, bounds)


def case_biexp_modes():
    cfg, p0, bounds = _cfg_args("C2")
    b, y, _ = synth.sample_voxels(cfg, 256, z=5)
    yn = y / y[:, :1]  # normalised signals for the S0-free modes
    p0r = {k: p0[k] for k in ("f1", "D1", "D2")}
    br = {k: bounds[k] for k in ("f1", "D1", "D2")}
    _run_curvefit("trf_biexp_reduced", "biexp", {}, b, yn, p0r, br)
    p0f = {"f1": 0.2, "D1": 0.001, "f2": 0.8, "D2": 0.02}
    bf = {"f1": (0.0, 1.5), "D1": (1e-5, 0.003), "f2": (0.0, 1.5), "D2": (0.003, 0.3)}
    _run_curvefit("trf_biexp_full", "biexp", {"fit_reduced": False}, b, yn, p0f, bf)


def case_triexp():
    cfg, p0, bounds = _cfg_args("C5")
    b, y, _ = synth.sample_voxels(cfg, 256)
    _run_curvefit("trf_triexp_reduced", "triexp", {}, b, y, p0, bounds)
    p0s = dict(p0, S0=900.0)
    bs = dict(bounds, S0=(1.0, 5000.0))
    _run_curvefit("trf_triexp_s0", "triexp", {"fit_s0": True}, b, y[:128] * 1000.0, p0s, bs)
    p0f = {"f1": 0.15, "D1": 0.1, "f2": 0.25, "D2": 0.01, "f3": 0.6, "D3": 0.001}
    bf = {"f1": (0.0, 1.0), "D1": (0.03, 0.5), "f2": (0.0, 1.0), "D2": (0.003, 0.03),
          "f3": (0.0, 1.0), "D3": (1e-4, 0.003)}
    _run_curvefit("trf_triexp_full", "triexp", {"fit_reduced": False}, b, y[:128], p0f, bf)


def case_constrained_c5():
    """SLSQP goldens: 4 096 voxels of config C5 (the planted 5 % sub-population has the constraint
    active), and per-voxel fixed parameters (constrained_curvefit.py:170-177): a fixed D3 keeps both
    fractions free (constraint applies), a fixed f1 leaves one fraction (the reference then applies no
    constraint at all)."""
    cfg, p0, bounds = _cfg_args("C5")
    b, y, _ = synth.sample_voxels(cfg, 4096)
    _run_curvefit("slsqp_triexp_c5", "triexp", {}, b, y, p0, bounds,
                  cls=ConstrainedCurveFitSolver, fraction_constraint=True)
    b, y, _ = synth.sample_voxels(cfg, 512, z=7)
    rng = np.random.default_rng(55)
    d3 = rng.uniform(5e-4, 2e-3, size=y.shape[0])
    _run_curvefit("slsqp_triexp_c5_pixfixed_D3", "triexp", {}, b, y, p0, bounds, pixel_fixed={"D3": d3},
                  cls=ConstrainedCurveFitSolver, fraction_constraint=True)
    f1 = rng.uniform(0.05, 0.25, size=y.shape[0])
    _run_curvefit("slsqp_triexp_c5_pixfixed_f1", "triexp", {}, b, y, p0, bounds, pixel_fixed={"f1": f1},
                  cls=ConstrainedCurveFitSolver, fraction_constraint=True)


def case_fixed():
    cfg, p0, bounds = _cfg_args("C2")
    b, y, _ = synth.sample_voxels(cfg, 256, z=9)
    # segmented step 2: per-voxel fixed slow diffusion (analytic-Jacobian path)
    rng = np.random.default_rng(77)
    d1 = rng.uniform(8e-4, 2e-3, size=y.shape[0])
    _run_curvefit("trf_biexp_s0_pixfixed_D1", "biexp", {"fit_s0": True}, b, y, p0, bounds,
                  pixel_fixed={"D1": d1})
    # model-level scalar fixed parameter
    p0m = {k: v for k, v in p0.items() if k != "D2"}
    bm = {k: v for k, v in bounds.items() if k != "D2"}
    _run_curvefit("trf_biexp_s0_modelfixed_D2", "biexp",
                  {"fit_s0": True, "fixed_params": {"D2": 0.03}}, b, y, p0m, bm)


def case_per_voxel_p0():
    """IDEAL-style call: per-voxel p0 and bounds arrays (fitters/ideal.py:240-242)."""
    cfg, p0, bounds = _cfg_args("C2")
    b, y, _ = synth.sample_voxels(cfg, 256, z=12)
    names = ["f1", "D1", "D2", "S0"]
    rng = np.random.default_rng(5)
    lo = np.array([bounds[n][0] for n in names])[:, None]
    hi = np.array([bounds[n][1] for n in names])[:, None]
    centre = np.array([p0[n] for n in names])[:, None] * rng.uniform(0.7, 1.3, size=(4, y.shape[0]))
    centre = np.clip(centre, lo, hi)
    tol = np.array([0.2, 0.2, 0.2, 0.5])[:, None]
    lb = np.clip(centre * (1 - tol), lo, hi)
    ub = np.clip(centre * (1 + tol), lo, hi)
    _run_curvefit("trf_biexp_s0_pervoxel", "biexp", {"fit_s0": True}, b, y, p0, bounds,
                  p0_arr=centre, bounds_arr=(lb, ub))


def case_degenerate():
    """Appendix A.12 voxels + failure modes (max_iter exhaustion, lb == ub, x0 outside)."""
    cfg, p0, bounds = _cfg_args("C2")
    b = cfg.bvalues
    good = synth.sample_voxels(cfg, 4, z=3)[1]
    y = np.stack([
        np.zeros(16), np.full(16, -5.0), np.full(16, 1e-12), np.full(16, 300.0),
        np.r_[np.nan, good[0][1:]], np.r_[good[1][:5], np.inf, good[1][6:]],
        good[2], good[3], 1e6 * good[2] / good[2][0], 1e-3 * good[3],
    ])
    _run_curvefit("trf_biexp_s0_degenerate", "biexp", {"fit_s0": True}, b, y, p0, bounds)
    c1, p01, b1 = _cfg_args("C1")
    _run_curvefit("trf_mono_degenerate", "monoexp", {}, b, y, p01, b1)
    # nfev exhaustion -> status 0 -> failure -> p0
    y2 = synth.sample_voxels(cfg, 64, z=3)[1]
    for mi in (1, 2, 3, 5):
        _run_curvefit(f"trf_biexp_s0_maxiter{mi}", "biexp", {"fit_s0": True}, b, y2, p0, bounds,
                      max_iter=mi)
    # invalid per-voxel bounds / p0
    names = ["f1", "D1", "D2", "S0"]
    n = 8
    P0 = np.tile(np.array([p0[k] for k in names])[:, None], (1, n))
    LB = np.tile(np.array([bounds[k][0] for k in names])[:, None], (1, n))
    UB = np.tile(np.array([bounds[k][1] for k in names])[:, None], (1, n))
    UB[1, 0] = LB[1, 0]            # lb == ub
    LB[2, 1] = 0.5; UB[2, 1] = 0.4  # lb > ub
    P0[0, 2] = 1.5                  # x0 above ub
    P0[3, 3] = 0.0                  # x0 below lb
    P0[0, 4] = 0.01                 # x0 on the lower bound -> make_strictly_feasible
    P0[1, 5] = 0.003                # x0 on the upper bound
    LB[3, 6] = -np.inf              # half-open
    UB[3, 7] = np.inf
    _run_curvefit("trf_biexp_s0_badbounds", "biexp", {"fit_s0": True}, b, y2[:n], p0, bounds,
                  p0_arr=P0, bounds_arr=(LB, UB))


def case_nnls_c3():
    cfg = synth.CONFIGS["C3"]
    b, y, _ = synth.sample_voxels(cfg, 256)
    for order, n_vox in ((2, 256), (0, 64), (1, 64), (3, 64)):
        model = NNLSModel(d_range=(0.0008, 0.5), n_bins=250)
        solver = NNLSSolver(model=model, reg_order=order, mu=0.02, max_iter=250)
        t = time.time()
        solver.fit(b, y[:n_vox])
        print(f"  nnls reg{order}: {n_vox} voxels in {time.time() - t:.1f}s")
        _save(
            f"nnls_c3_reg{order}", b=b, y=y[:n_vox], d_range=np.array([0.0008, 0.5]), n_bins=250,
            reg_order=order, mu=0.02, max_iter=250,
            coefficients=solver.params_["coefficients"],
            residual=solver.diagnostics_["residual"],
            success=np.array([pr.success for pr in solver.pixel_results_]),
        )
    # degenerate signals and a tight iteration cap (failure -> zeros, ||b||)
    yd = np.stack([np.zeros(16), np.full(16, -5.0), np.full(16, 300.0), 1e-12 * np.ones(16),
                   np.r_[1000.0, np.zeros(15)], y[0], y[1]])
    for mi, tag in ((250, "degenerate"), (5, "maxiter5"), (20, "maxiter20")):
        yy = yd if tag == "degenerate" else y[:32]
        model = NNLSModel(d_range=(0.0008, 0.5), n_bins=250)
        solver = NNLSSolver(model=model, reg_order=2, mu=0.02, max_iter=mi)
        solver.fit(b, yy)
        _save(
            f"nnls_c3_{tag}", b=b, y=yy, d_range=np.array([0.0008, 0.5]), n_bins=250,
            reg_order=2, mu=0.02, max_iter=mi,
            coefficients=solver.params_["coefficients"],
            residual=solver.diagnostics_["residual"],
            success=np.array([pr.success for pr in solver.pixel_results_]),
        )
    # a small, odd-shaped problem
    model = NNLSModel(d_range=(0.001, 0.2), n_bins=37)
    solver = NNLSSolver(model=model, reg_order=1, mu=0.1, max_iter=250)
    solver.fit(b[:11], y[:64, :11])
    _save("nnls_small_reg1", b=b[:11], y=y[:64, :11], d_range=np.array([0.001, 0.2]), n_bins=37,
          reg_order=1, mu=0.1, max_iter=250, coefficients=solver.params_["coefficients"],
          residual=solver.diagnostics_["residual"],
          success=np.array([pr.success for pr in solver.pixel_results_]))




def case_resize():
    """cv2.resize outputs for the IDEAL resampler (fitters/ideal.py:299-320).

    Linear goldens are taken with IPP disabled: the opencv-python wheel routes large
    INTER_LINEAR resizes to Intel IPP (closed arithmetic); OpenCV's own code is the pin.
    """
    import cv2

    rng = np.random.default_rng(11)
    shapes = [(8, 8, 16, 16), (16, 16, 8, 8), (32, 32, 64, 64), (64, 64, 16, 16), (5, 7, 10, 14),
              (10, 10, 24, 24), (3, 3, 7, 7), (24, 24, 10, 10), (40, 40, 72, 72), (7, 9, 33, 20),
              (64, 64, 8, 8), (16, 16, 32, 32), (4, 4, 2, 2)]
    out = {}
    for n, (h, w, H, W) in enumerate(shapes):
        a = rng.normal(size=(h, w, 1, 2)) * 1000.0
        cub = np.zeros((H, W, 1, 2))
        lin = np.zeros((H, W, 1, 2))
        for z in range(1):
            for c in range(2):
                cub[..., z, c] = cv2.resize(a[..., z, c], (W, H), interpolation=cv2.INTER_CUBIC)
        cv2.ipp.setUseIPP(False)
        for z in range(1):
            for c in range(2):
                lin[..., z, c] = cv2.resize(a[..., z, c], (W, H), interpolation=cv2.INTER_LINEAR)
        cv2.ipp.setUseIPP(True)
        out[f"src{n}"], out[f"cubic{n}"], out[f"linear{n}"] = a, cub, lin
    # integer label mask on dyadic pyramids (the FP32 path of the IDEAL fitter)
    seg = synth.ellipsoid_mask((64, 64, 4))[..., None] * 2
    for n, t in enumerate((4, 8, 16, 32)):
        s32 = seg.astype(np.float32)
        o = np.zeros((t, t, 4, 1), np.float32)
        for z in range(4):
            o[..., z, 0] = cv2.resize(s32[..., z, 0], (t, t), interpolation=cv2.INTER_CUBIC)
        out[f"seg{n}"] = o
    out["seg_src"] = seg
    out["n_cases"] = len(shapes)
    _save("resize_cv2", **out)


def _small_volume(shape=(32, 32, 2), seed=4):
    """A C4-style biexp volume with smooth parameter fields, small enough for the CPU reference."""
    cfg = synth.Config(**{**synth.CONFIGS["C4"].__dict__, "shape": shape, "seed": seed})
    b, img, truth = synth.make_volume(cfg)
    x = (np.arange(shape[0]) - (shape[0] - 1) / 2) / (shape[0] / 2)
    y = (np.arange(shape[1]) - (shape[1] - 1) / 2) / (shape[1] / 2)
    mask = ((x[:, None] ** 2 + y[None, :] ** 2) <= 0.85).astype(np.int64)
    seg = np.repeat(mask[:, :, None], shape[2], axis=2)
    return cfg, b, img, seg


def case_fitters():
    """PixelWise / IDEAL / Segmented fitters of the reference on a 32 x 32 x 2 volume."""
    from pyneapple.fitters import IDEALFitter, PixelWiseFitter, SegmentedFitter

    cfg, b, img, seg = _small_volume()
    p0, bounds = cfg.p0, cfg.bounds

    def solver():
        return CurveFitSolver(model=BiExpModel(fit_s0=True), max_iter=250, tol=1e-8, p0=p0, bounds=bounds)

    # pixelwise with a mask
    f = PixelWiseFitter(solver=solver()).fit(b, img, seg)
    r = f.results_
    _save("fitter_pixelwise", b=b, image=img, seg=seg, names=np.array(list(r.params)),
          params=np.stack([r.params[n] for n in r.params]), success=r.success, r_squared=r.r_squared,
          covariance=r.covariance, pixel_indices=np.array(r.pixel_indices),
          predict=f.predict(b))
    # IDEAL, cubic, 4 levels
    steps = np.array([[4, 4], [8, 8], [16, 16], [32, 32]])
    tol = {"S0": 0.5, "f1": 0.2, "D1": 0.2, "D2": 0.2}
    fi = IDEALFitter(solver=solver(), dim_steps=steps, step_tol=tol, ideal_dims=2,
                     segmentation_threshold=0.2, interpolation_method="cubic").fit(b, img, seg)
    ri = fi.results_
    out = {f"step{i}": m for i, m in enumerate(fi.step_params)}
    _save("fitter_ideal", b=b, image=img, seg=seg, dim_steps=steps, names=np.array(list(ri.params)),
          params=np.stack([ri.params[n] for n in ri.params]), success=ri.success, r_squared=ri.r_squared,
          pixel_indices=np.array(ri.pixel_indices), n_steps=len(fi.step_params), **out)
    # segmented: monoexp on b >= 200 -> D fixed as the slow D1 of the biexp
    s1 = CurveFitSolver(model=MonoExpModel(), max_iter=250, tol=1e-8, p0={"S0": 1000.0, "D": 0.001},
                        bounds={"S0": (1.0, 5000.0), "D": (1e-5, 0.003)})
    fs = SegmentedFitter(step1_solver=s1, step2_solver=solver(), step1_bvalue_range=(200, None),
                         fixed_from_step1=["D"], param_mapping={"D": "D1"}).fit(b, img, seg)
    rs = fs.results_
    _save("fitter_segmented", b=b, image=img, seg=seg,
          names=np.array(list(fs.fitted_params_)), params=np.stack([fs.fitted_params_[n] for n in fs.fitted_params_]),
          step1_names=np.array(list(fs.step1_params_)),
          step1_params=np.stack([fs.step1_params_[n] for n in fs.step1_params_]),
          success=rs.success, r_squared=rs.r_squared, step1_success=fs.step1_result_.success)


def case_spectrum():
    """utility/spectrum.py on real regularised spectra (golden NNLS cases) and on hand-made edge cases."""
    from pyneapple.utility.spectrum import apply_cutoffs, find_spectrum_peaks
    from scipy import signal as scipy_signal

    g2 = np.load(os.path.join(GOLD, "nnls_c3_reg2.npz"))
    g0 = np.load(os.path.join(GOLD, "nnls_c3_reg0.npz"))
    bins = NNLSModel(d_range=(0.0008, 0.5), n_bins=250).bins
    rng = np.random.default_rng(7)
    n = 250
    edge = np.zeros((24, n))
    edge[1, 100] = 3.0                                   # single spike
    edge[2, 100:104] = 2.0                               # flat-topped peak (even plateau -> lower midpoint)
    edge[3, 100:105] = 2.0                               # odd plateau
    edge[4, 0] = 5.0; edge[4, n - 1] = 5.0               # maxima on the borders are not peaks
    edge[5, 1] = 1.0; edge[5, n - 2] = 1.0               # next to the borders they are
    edge[6, 50] = 1.0; edge[6, 52] = 1.0                 # two equal spikes one bin apart
    edge[7, 60:70] = np.array([1, 2, 3, 4, 5, 5, 4, 3, 2, 1.0])      # plateau on a ramp
    edge[8, :] = 1.0                                     # constant
    edge[9, 40:60] = np.hanning(20) * 4; edge[9, 55:80] += np.hanning(25) * 2   # overlapping lobes
    edge[10, 30:50] = np.hanning(20); edge[10, 120:150] = 0.5 * np.hanning(30); edge[10, 200:230] = 2 * np.hanning(30)
    edge[11, 10:240] = np.abs(np.sin(np.arange(230) / 3.0))          # many peaks (> max_peaks)
    edge[12, 100:110] = np.array([1, 3, 1, 3, 1, 3, 1, 3, 1, 0.0])   # equal heights, prominence ties
    edge[13, 100] = 0.1                                  # exactly the height threshold
    edge[14, 100] = 0.1 - 1e-12                          # just below it
    edge[15, 80:120] = np.hanning(40) + 0.5              # peak on a pedestal with cliffs
    edge[16, :] = np.linspace(0, 1, n)                   # monotone: no peak
    edge[17, :] = np.linspace(1, 0, n)
    edge[18, 100:103] = np.array([1.0, 1.0, 2.0])        # plateau followed by a rise, then a fall
    edge[19, 5:15] = np.hanning(10) * 3; edge[19, 235:245] = np.hanning(10) * 3
    edge[20:24] = rng.random((4, n)) * (rng.random((4, n)) > 0.93)   # sparse random
    spectra = np.concatenate([g2["coefficients"][:96], g0["coefficients"][:32], edge])
    cutoffs = [(0.0008, 0.003), (0.003, 0.05), (0.05, 0.5)]
    out = {}
    for tag, height, reg in (("h0p1_reg", 0.1, True), ("h0p1_raw", 0.1, False), ("h5_reg", 5.0, True)):
        P = 32
        nv = spectra.shape[0]
        n_peaks = np.zeros(nv, np.int32); idx = np.full((nv, P), -1, np.int32)
        d = np.full((nv, P), np.nan); f = np.full((nv, P), np.nan)
        dc = np.full((nv, len(cutoffs)), np.nan); fc = np.full((nv, len(cutoffs)), np.nan)
        for v in range(nv):
            dv, fv = find_spectrum_peaks(spectra[v], bins, height=height, regularized=reg)
            pk = scipy_signal.find_peaks(spectra[v], height=height)[0]
            assert len(pk) == len(dv) and len(dv) <= P
            n_peaks[v] = len(dv); idx[v, :len(dv)] = pk; d[v, :len(dv)] = dv; f[v, :len(dv)] = fv
            dc[v], fc[v] = apply_cutoffs(dv, fv, cutoffs)
        out.update({f"{tag}_n_peaks": n_peaks, f"{tag}_idx": idx, f"{tag}_d": d, f"{tag}_f": f,
                    f"{tag}_d_cut": dc, f"{tag}_f_cut": fc})
    _save("spectrum_peaks", spectra=spectra, bins=bins, cutoffs=np.array(cutoffs), **out)


def case_dogbox():
    """method = "dogbox" (forwarded from the TOML to curve_fit, solvers/curvefit.py:295-306)."""
    cfg, p0, bounds = _cfg_args("C2")
    b, y, _ = synth.sample_voxels(cfg, 512, z=3)
    _run_curvefit("dbx_biexp_s0_c2", "biexp", {"fit_s0": True}, b, y, p0, bounds, method="dogbox")
    # tight bounds and a start on a bound: active-set logic, variables snapped onto bounds
    bt = dict(bounds, D1=(1e-5, 0.0012), f1=(0.01, 0.25))
    pt = dict(p0, D1=1e-5, f1=0.25)
    _run_curvefit("dbx_biexp_s0_onbound", "biexp", {"fit_s0": True}, b, y[:256], pt, bt, method="dogbox")
    for mi in (1, 2, 3, 6):
        _run_curvefit(f"dbx_biexp_s0_maxiter{mi}", "biexp", {"fit_s0": True}, b, y[:48], p0, bounds,
                      method="dogbox", max_iter=mi)
    rng = np.random.default_rng(78)
    d1 = rng.uniform(8e-4, 2e-3, size=128)
    _run_curvefit("dbx_biexp_s0_pixfixed_D1", "biexp", {"fit_s0": True}, b, y[:128], p0, bounds,
                  pixel_fixed={"D1": d1}, method="dogbox")
    cfg1, p01, b1 = _cfg_args("C1")
    bb, yy, _ = synth.sample_voxels(cfg1, 256, z=2)
    _run_curvefit("dbx_mono_c1", "monoexp", {}, bb, yy, p01, b1, method="dogbox")
    cfg5, p05, b5 = _cfg_args("C5")
    bb, yy, _ = synth.sample_voxels(cfg5, 128, z=2)
    _run_curvefit("dbx_triexp_reduced", "triexp", {}, bb, yy, p05, b5, method="dogbox")


def _t1_signal(model, b, truth, sigma, rng):
    """Noisy signals of a reference model for per-voxel truth vectors (n_vox, n_all)."""
    y = np.stack([model.forward(b, *row) for row in truth])
    return y + rng.normal(0.0, sigma, y.shape)


def case_t1():
    """T1 / STEAM variants (models/monoexp.py:120-163, model_functions/multiexp.py:210-302), T1 fitted,
    model-fixed and per-voxel fixed, for method trf and dogbox.  TR / TM / T1 in ms."""
    rng = np.random.default_rng(21)
    tr, tm = 2500.0, 30.0
    b16 = synth.CONFIGS["C2"].bvalues
    b24 = synth.CONFIGS["C5"].bvalues
    n = 128

    def extras(model):
        return dict(t1_mode=2 if model.fit_t1_steam else 1, tr=tr, tm=tm if model.fit_t1_steam else 0.0)

    def run(name, kind, mk, b, y, p0, bounds, **kw):
        model = MODELS[kind](**mk)
        _run_curvefit(name, kind, mk, b, y, p0, bounds, **kw)
        # re-save with the T1 description added (np.savez has no append)
        path = os.path.join(GOLD, name + ".npz")
        g = dict(np.load(path))
        g.update({k: np.asarray(v) for k, v in extras(model).items()})
        np.savez_compressed(path, **g)

    # ---- mono-exponential -------------------------------------------------------------------
    s0, d, t1 = rng.uniform(800, 1200, n), rng.uniform(8e-4, 2e-3, n), rng.uniform(900, 1500, n)
    p0 = {"S0": 1000.0, "D": 1e-3, "T1": 1200.0}
    bd = {"S0": (1.0, 5000.0), "D": (1e-5, 0.1), "T1": (100.0, 5000.0)}
    for tag, mk in (("t1", dict(fit_t1=True, repetition_time=tr)),
                    ("steam", dict(fit_t1_steam=True, repetition_time=tr, mixing_time=tm))):
        m = MonoExpModel(**mk)
        y = _t1_signal(m, b16, np.stack([s0, d, t1], 1), 3.0, rng)
        run(f"trf_mono_{tag}", "monoexp", mk, b16, y, p0, bd)                      # T1 fitted (S0 x C(T1) only)
        run(f"trf_mono_{tag}_pixfixed", "monoexp", mk, b16, y, p0, bd, pixel_fixed={"T1": t1})
        mkf = dict(mk, fixed_params={"T1": 1200.0})
        p0f = {k: v for k, v in p0.items() if k != "T1"}
        bdf = {k: v for k, v in bd.items() if k != "T1"}
        run(f"trf_mono_{tag}_modelfixed", "monoexp", mkf, b16, y, p0f, bdf)
        run(f"dbx_mono_{tag}_pixfixed", "monoexp", mk, b16, y, p0, bd, pixel_fixed={"T1": t1}, method="dogbox")
    # ---- bi-exponential ----------------------------------------------------------------------
    f1, d1, d2 = rng.uniform(0.05, 0.4, n), rng.uniform(5e-4, 2.5e-3, n), rng.uniform(0.01, 0.1, n)
    s0 = rng.uniform(800, 1200, n)
    # reduced mode: the amplitude is fixed at 1, so T1 is identifiable from C(T1) alone
    mk = dict(fit_t1=True, repetition_time=tr)
    m = BiExpModel(**mk)
    y = _t1_signal(m, b16, np.stack([f1, d1, d2, t1], 1), 0.005, rng)
    p0 = {"f1": 0.2, "D1": 0.001, "D2": 0.02, "T1": 1200.0}
    bd = {"f1": (0.01, 0.99), "D1": (1e-5, 0.003), "D2": (0.003, 0.3), "T1": (100.0, 5000.0)}
    run("trf_biexp_reduced_t1", "biexp", mk, b16, y, p0, bd)
    run("dbx_biexp_reduced_t1", "biexp", mk, b16, y, p0, bd, method="dogbox")
    mk = dict(fit_s0=True, fit_t1_steam=True, repetition_time=tr, mixing_time=tm)
    m = BiExpModel(**mk)
    y = _t1_signal(m, b16, np.stack([f1, d1, d2, s0, t1], 1), 3.0, rng)
    p0 = dict(p0, S0=1000.0)
    bd = dict(bd, S0=(1.0, 5000.0))
    run("trf_biexp_s0_steam_pixfixed", "biexp", mk, b16, y, p0, bd, pixel_fixed={"T1": t1})
    run("dbx_biexp_s0_steam_pixfixed", "biexp", mk, b16, y, p0, bd, pixel_fixed={"T1": t1}, method="dogbox")
    mk = dict(fit_reduced=False, fit_t1=True, repetition_time=tr, fixed_params={"T1": 1300.0})
    m = BiExpModel(fit_reduced=False, fit_t1=True, repetition_time=tr)
    y = _t1_signal(m, b16, np.stack([f1, d1, 1 - f1, d2, np.full(n, 1300.0)], 1), 0.005, rng)
    p0 = {"f1": 0.2, "D1": 0.001, "f2": 0.8, "D2": 0.02}
    bd = {"f1": (0.0, 1.5), "D1": (1e-5, 0.003), "f2": (0.0, 1.5), "D2": (0.003, 0.3)}
    run("trf_biexp_full_t1_modelfixed", "biexp", mk, b16, y, p0, bd)
    # ---- tri-exponential ---------------------------------------------------------------------
    n3 = 96
    f1, f2 = rng.uniform(0.05, 0.25, n3), rng.uniform(0.1, 0.35, n3)
    d1, d2, d3 = rng.uniform(0.05, 0.2, n3), rng.uniform(5e-3, 2e-2, n3), rng.uniform(5e-4, 2e-3, n3)
    t13 = rng.uniform(900, 1500, n3)
    p0 = {"f1": 0.15, "D1": 0.1, "f2": 0.25, "D2": 0.01, "D3": 0.001, "T1": 1200.0}
    bd = {"f1": (0.0, 1.0), "D1": (0.03, 0.5), "f2": (0.0, 1.0), "D2": (0.003, 0.03), "D3": (1e-4, 0.003),
          "T1": (100.0, 5000.0)}
    mk = dict(fit_t1_steam=True, repetition_time=tr, mixing_time=tm)
    m = TriExpModel(**mk)
    y = _t1_signal(m, b24, np.stack([f1, d1, f2, d2, d3, t13], 1), 0.004, rng)
    run("trf_triexp_reduced_steam", "triexp", mk, b24, y, p0, bd)
    run("dbx_triexp_reduced_steam_pixfixed", "triexp", mk, b24, y, p0, bd, pixel_fixed={"T1": t13}, method="dogbox")
    mk = dict(fit_s0=True, fit_t1=True, repetition_time=tr)
    m = TriExpModel(**mk)
    s03 = rng.uniform(800, 1200, n3)
    y = _t1_signal(m, b24, np.stack([f1, d1, f2, d2, d3, s03, t13], 1), 3.0, rng)
    run("trf_triexp_s0_t1_pixfixed", "triexp", mk, b24, y, dict(p0, S0=900.0), dict(bd, S0=(1.0, 5000.0)),
        pixel_fixed={"T1": t13})
    mk = dict(fit_reduced=False, fit_t1_steam=True, repetition_time=tr, mixing_time=tm, fixed_params={"T1": 1100.0})
    m = TriExpModel(fit_reduced=False, fit_t1_steam=True, repetition_time=tr, mixing_time=tm)
    y = _t1_signal(m, b24, np.stack([f1, d1, f2, d2, 1 - f1 - f2, d3, np.full(n3, 1100.0)], 1), 0.004, rng)
    p0f = {"f1": 0.15, "D1": 0.1, "f2": 0.25, "D2": 0.01, "f3": 0.6, "D3": 0.001}
    bf = {"f1": (0.0, 1.0), "D1": (0.03, 0.5), "f2": (0.0, 1.0), "D2": (0.003, 0.03), "f3": (0.0, 1.0),
          "D3": (1e-4, 0.003)}
    run("trf_triexp_full_steam_modelfixed", "triexp", mk, b24, y, p0f, bf)


def case_lm():
    """method = "lm" on problems without bounds: curve_fit -> leastsq -> MINPACK lmdif (forward
    differences) or, with a fixed parameter, lmder (analytic Jacobian); maxfev = max_iter."""
    inf = float("inf")

    def free(bounds):
        return {k: (-inf, inf) for k in bounds}

    cfg, p0, bounds = _cfg_args("C1")
    b, y, _ = synth.sample_voxels(cfg, 256, z=4)
    _run_curvefit("lm_mono_c1", "monoexp", {}, b, y, p0, free(bounds), method="lm")
    cfg, p0, bounds = _cfg_args("C2")
    b, y, _ = synth.sample_voxels(cfg, 512, z=6)
    _run_curvefit("lm_biexp_s0_c2", "biexp", {"fit_s0": True}, b, y, p0, free(bounds), method="lm")
    for mi in (3, 8, 14, 30):  # maxfev counts the forward-difference evaluations too
        _run_curvefit(f"lm_biexp_s0_maxiter{mi}", "biexp", {"fit_s0": True}, b, y[:64], p0, free(bounds),
                      method="lm", max_iter=mi)
    rng = np.random.default_rng(79)
    d1 = rng.uniform(8e-4, 2e-3, size=128)
    _run_curvefit("lm_biexp_s0_pixfixed_D1", "biexp", {"fit_s0": True}, b, y[:128], p0, free(bounds),
                  pixel_fixed={"D1": d1}, method="lm")
    yn = y[:128] / y[:128, :1]
    p0r = {k: p0[k] for k in ("f1", "D1", "D2")}
    _run_curvefit("lm_biexp_reduced", "biexp", {}, b, yn, p0r, free(p0r), method="lm")
    cfg5, p05, b5 = _cfg_args("C5")
    bb, yy, _ = synth.sample_voxels(cfg5, 128, z=3)
    _run_curvefit("lm_triexp_reduced", "triexp", {}, bb, yy, p05, free(b5), method="lm")
    # degenerate signals: zeros, constants, a NaN
    good = synth.sample_voxels(cfg, 4, z=3)[1]
    yd = np.stack([np.zeros(16), np.full(16, 300.0), np.r_[np.nan, good[0][1:]], good[1], 1e-3 * good[2]])
    _run_curvefit("lm_biexp_s0_degenerate", "biexp", {"fit_s0": True}, b, yd, p0, free(bounds), method="lm")


CASES = {k[5:]: v for k, v in globals().items() if k.startswith("case_")}

if __name__ == "__main__":
    os.makedirs(GOLD, exist_ok=True)
    todo = sys.argv[1:] or list(CASES)
    for c in todo:
        print(f"[{c}]")
        CASES[c]()
