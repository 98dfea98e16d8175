"""The per-voxel device mathematics (pnb_trf_core.cuh) compiled with g++ against the goldens.

This is how the kernel arithmetic is checked in the GPU-less container; the product only ever
runs the nvcc build (tests/test_trf_gpu.py checks that one on the B200).
"""

from __future__ import annotations

import numpy as np
import pytest

from _util import (DBX_CASES, DBX_T1_CASES, LM_CASES, T1_AMPLITUDE_ONLY, TRF_CASES, TRF_T1_CASES, full_problem,
                   rel_err)
from hostsim import hostsim
from oracle import c_oracle


@pytest.mark.parametrize("name", sorted(TRF_CASES))
@pytest.mark.parametrize("mode", ["reference", "analytic"])
def test_core_matches_reference(name, mode):
    kind, m = TRF_CASES[name]
    P = full_problem(name)
    jm = (1 if P["uses_fd"] else 0) if mode == "reference" else 0
    r = hostsim.trf_fit(c_oracle.MODEL_IDS[(kind, m)], P["b"], P["y"], P["P0"], P["LB"], P["UB"],
                        frozen=P["frozen"], ftol=P["tol"], max_nfev=P["max_iter"], jac_mode=jm)
    free = [i for i in range(len(P["all_names"])) if not P["frozen"][i]]
    par = r["params"][:, free]
    assert ((r["status"] > 0) == P["ref_success"]).all()
    fail = ~P["ref_success"]
    assert np.array_equal(par[fail], P["ref_params"][fail])
    ok = P["ref_success"]
    if ok.any():
        err = rel_err(par[ok], P["ref_params"][ok]).max(axis=1)
        allowed = 2 if name == "trf_triexp_full" else 0
        assert (err > 1e-4).sum() <= allowed
        if mode == "reference":
            assert np.median(err) < 1e-7


def test_t1_variants_against_c_oracle():
    """T1 / STEAM models (model_functions/multiexp.py:210-302): core vs the C restatement."""
    rng = np.random.default_rng(3)
    b = np.array([0, 50, 100, 200, 400, 600, 800, 1000, 1200, 1500], float)
    n = 64
    s0, d, t1 = rng.uniform(800, 1200, n), rng.uniform(8e-4, 2e-3, n), rng.uniform(900, 1500, n)
    for t1_mode, tm in ((1, 0.0), (2, 30.0)):
        tr = 2500.0
        y = s0[:, None] * np.exp(-b * d[:, None]) * (1 - np.exp(-tr / t1[:, None]))
        if t1_mode == 2:
            y = y * np.exp(-tm / t1[:, None])
        y = y + rng.normal(0, 3, y.shape)
        P0 = np.tile([1000.0, 1e-3, 1200.0], (n, 1))
        LB = np.tile([1.0, 1e-5, 100.0], (n, 1))
        UB = np.tile([5000.0, 0.1, 5000.0], (n, 1))
        a = hostsim.trf_fit(0, b, y, P0, LB, UB, t1_mode=t1_mode, tr=tr, tm=tm, jac_mode=0)
        c = c_oracle.trf_fit(0, b, y, P0, LB, UB, t1_mode=t1_mode, tr=tr, tm=tm, jac_mode=0)
        assert ((a["status"] > 0) == (c["status"] > 0)).all()
        ok = c["status"] > 0
        # S0 and T1 enter only through the product S0 * C(T1): compare the identifiable quantities
        def amp(p):
            f = 1 - np.exp(-tr / p[:, 2])
            return p[:, 0] * (f * np.exp(-tm / p[:, 2]) if t1_mode == 2 else f)
        assert rel_err(a["params"][ok, 1], c["params"][ok, 1]).max() < 1e-7
        assert rel_err(amp(a["params"][ok]), amp(c["params"][ok])).max() < 1e-7
        assert rel_err(a["cost"][ok], c["cost"][ok]).max() < 1e-9


def test_x_scale_jac_against_scipy():
    from scipy.optimize import least_squares

    P = full_problem("trf_biexp_s0_c2")
    n = 24
    r = hostsim.trf_fit(3, P["b"], P["y"][:n], P["P0"][:n], P["LB"][:n], P["UB"][:n], jac_mode=1,
                        x_scale_jac=True)
    from oracle import ref_port

    m = ref_port.Model("biexp", "s0")
    for i in range(n):
        ref = least_squares(lambda p: m.forward(P["b"], *p) - P["y"][i], P["P0"][i],
                            bounds=(P["LB"][i], P["UB"][i]), method="trf", x_scale="jac", ftol=1e-8,
                            max_nfev=250)
        assert rel_err(r["params"][i], ref.x).max() < 1e-5


def test_model_exp_is_within_one_ulp_of_libm():
    """pnb_exp (table + degree-5 polynomial) replaces libdevice exp() in the signal models."""
    rng = np.random.default_rng(11)
    x = np.concatenate([
        -rng.uniform(0.0, 690.0, 400_000), rng.uniform(0.0, 690.0, 100_000), -rng.uniform(0.0, 1e-3, 50_000),
        np.array([0.0, -0.0, -1e-300, 689.999, -689.999, -690.0, -745.0, -800.0, 709.0, 710.0, np.inf, -np.inf]),
    ])
    got = hostsim.exp(x)
    with np.errstate(over="ignore"):
        ref = np.exp(x)
    fin = np.isfinite(ref) & (ref > 0)
    ulp = np.abs(got[fin] - ref[fin]) / np.spacing(ref[fin])
    assert ulp.max() <= 1.0, ulp.max()
    assert np.array_equal(got[~fin], ref[~fin])
    assert np.isnan(hostsim.exp(np.array([np.nan]))[0])


@pytest.mark.parametrize("name", sorted(DBX_CASES))
def test_dogbox_core_matches_reference(name):
    """pnb_dogbox_core.cuh (method = "dogbox") against the reference's outputs."""
    kind, m = DBX_CASES[name]
    P = full_problem(name)
    jm = 1 if P["uses_fd"] else 0
    r = hostsim.trf_fit(c_oracle.MODEL_IDS[(kind, m)], P["b"], P["y"], P["P0"], P["LB"], P["UB"],
                        frozen=P["frozen"], ftol=P["tol"], max_nfev=P["max_iter"], jac_mode=jm, method=1)
    free = [i for i in range(len(P["all_names"])) if not P["frozen"][i]]
    par = r["params"][:, free]
    assert ((r["status"] > 0) == P["ref_success"]).all()
    fail = ~P["ref_success"]
    assert np.array_equal(par[fail], P["ref_params"][fail])
    ok = P["ref_success"]
    if ok.any():
        err = rel_err(par[ok], P["ref_params"][ok]).max(axis=1)
        assert (err > 1e-4).sum() == 0, (err > 1e-4).sum()
        assert np.median(err) < 1e-7


def test_dogbox_t1_variants_and_x_scale_jac_against_c_oracle():
    """method = "dogbox" with the T1 / STEAM models and with x_scale='jac': device core vs the plain-C
    restatement of scipy's dogbox (both pinned to the reference goldens for the plain models)."""
    rng = np.random.default_rng(5)
    b = np.array([0, 50, 100, 200, 400, 600, 800, 1000, 1200, 1500], float)
    n = 48
    s0, d, t1 = rng.uniform(800, 1200, n), rng.uniform(8e-4, 2e-3, n), rng.uniform(900, 1500, n)
    for t1_mode, tm in ((1, 0.0), (2, 30.0)):
        tr = 2500.0
        y = s0[:, None] * np.exp(-b * d[:, None]) * (1 - np.exp(-tr / t1[:, None]))
        if t1_mode == 2:
            y = y * np.exp(-tm / t1[:, None])
        y = y + rng.normal(0, 3, y.shape)
        P0 = np.tile([1000.0, 1e-3, 1200.0], (n, 1))
        LB = np.tile([1.0, 1e-5, 100.0], (n, 1))
        UB = np.tile([5000.0, 0.1, 5000.0], (n, 1))
        a = hostsim.trf_fit(0, b, y, P0, LB, UB, t1_mode=t1_mode, tr=tr, tm=tm, jac_mode=0, method=1)
        c = c_oracle.trf_fit(0, b, y, P0, LB, UB, t1_mode=t1_mode, tr=tr, tm=tm, jac_mode=0, method="dogbox")
        assert ((a["status"] > 0) == (c["status"] > 0)).all()
        ok = c["status"] > 0
        assert rel_err(a["params"][ok, 1], c["params"][ok, 1]).max() < 1e-6
        assert rel_err(a["cost"][ok], c["cost"][ok]).max() < 1e-8
    P = full_problem("dbx_biexp_s0_c2")
    m = 32
    a = hostsim.trf_fit(3, P["b"], P["y"][:m], P["P0"][:m], P["LB"][:m], P["UB"][:m], jac_mode=1, x_scale_jac=True, method=1)
    c = c_oracle.trf_fit(3, P["b"], P["y"][:m], P["P0"][:m], P["LB"][:m], P["UB"][:m], jac_mode=1, x_scale_jac=True,
                         method="dogbox")
    assert ((a["status"] > 0) == (c["status"] > 0)).all()
    assert rel_err(a["params"], c["params"]).max() < 1e-5


def _t1_amplitude(par, t1_mode, tr, tm, i_s0, i_t1):
    f = 1.0 - np.exp(-tr / par[:, i_t1])
    if t1_mode == 2:
        f = f * np.exp(-tm / par[:, i_t1])
    return par[:, i_s0] * f


@pytest.mark.parametrize("name", sorted(TRF_T1_CASES) + sorted(DBX_T1_CASES))
def test_t1_goldens_core_matches_reference(name):
    """T1 / STEAM variants against goldens produced by the reference's own models and solver
    (models/monoexp.py:120-163, model_functions/multiexp.py:210-302)."""
    cases = TRF_T1_CASES if name in TRF_T1_CASES else DBX_T1_CASES
    kind, m = cases[name]
    P = full_problem(name)
    jm = 1 if P["uses_fd"] else 0
    r = hostsim.trf_fit(c_oracle.MODEL_IDS[(kind, m)], P["b"], P["y"], P["P0"], P["LB"], P["UB"],
                        frozen=P["frozen"], ftol=P["tol"], max_nfev=P["max_iter"], jac_mode=jm,
                        t1_mode=P["t1_mode"], tr=P["tr"], tm=P["tm"], method=1 if name.startswith("dbx") else 0)
    free = [i for i in range(len(P["all_names"])) if not P["frozen"][i]]
    par = r["params"][:, free]
    assert ((r["status"] > 0) == P["ref_success"]).all()
    ok = P["ref_success"]
    if name in T1_AMPLITUDE_ONLY:
        ref = P["ref_params"]
        assert rel_err(par[ok, 1], ref[ok, 1]).max() < 1e-4
        amp = _t1_amplitude(par, P["t1_mode"], P["tr"], P["tm"], 0, 2)
        amp_ref = _t1_amplitude(ref, P["t1_mode"], P["tr"], P["tm"], 0, 2)
        assert rel_err(amp[ok], amp_ref[ok]).max() < 1e-4
        return
    err = rel_err(par[ok], P["ref_params"][ok]).max(axis=1)
    assert (err > 1e-4).sum() == 0, (int((err > 1e-4).sum()), float(err.max()))


@pytest.mark.parametrize("name", sorted(LM_CASES))
def test_lm_core_matches_reference(name):
    """pnb_lm_core.cuh (MINPACK lmdif / lmder restated on the normal matrix) against the reference run
    with method="lm" on unbounded problems: same success flags (maxfev counts the forward-difference
    evaluations), parameters within north_star's 1e-4 (measured ~1e-8), failures return p0."""
    kind, m = LM_CASES[name]
    P = full_problem(name)
    jm = 2 if P["uses_fd"] else 0
    r = hostsim.trf_fit(c_oracle.MODEL_IDS[(kind, m)], P["b"], P["y"], P["P0"], P["LB"], P["UB"],
                        frozen=P["frozen"], ftol=P["tol"], xtol=1.49012e-8, gtol=0.0, max_nfev=P["max_iter"],
                        jac_mode=jm, method=2)
    free = [i for i in range(len(P["all_names"])) if not P["frozen"][i]]
    par = r["params"][:, free]
    assert ((r["status"] > 0) == P["ref_success"]).all(), (r["status"], P["ref_success"])
    fail = ~P["ref_success"]
    assert np.array_equal(par[fail], P["ref_params"][fail])
    ok = P["ref_success"].copy()
    if name == "lm_biexp_s0_degenerate":
        # all-zero and constant signals have no unique minimiser (S0 -> 0 frees everything else):
        # both answers must reproduce the signal equally well
        from oracle import ref_port

        mdl = ref_port.Model(kind, m)
        for v in (0, 1):
            ours = np.linalg.norm(mdl.forward(P["b"], *r["params"][v]) - P["y"][v])
            ref = np.linalg.norm(mdl.forward(P["b"], *P["ref_params"][v]) - P["y"][v])
            assert ours <= ref * (1 + 1e-6) + 1e-6
            ok[v] = False
    if ok.any():
        err = rel_err(par[ok], P["ref_params"][ok]).max(axis=1)
        off = np.where(ok)[0][err > 1e-4]
        # without bounds a few voxels have a flat direction (a fast component that has decayed before the
        # first non-zero b-value leaves its D free): the parameter gate applies to identifiable voxels,
        # the residual gate to all
        assert off.size <= max(0, int(0.02 * ok.sum())), (off, err.max())
        if off.size:
            from oracle import ref_port

            mdl = ref_port.Model(kind, m)
            for v in off:
                ours = np.linalg.norm(mdl.forward(P["b"], *r["params"][v]) - P["y"][v])
                ref = np.linalg.norm(mdl.forward(P["b"], *P["ref_params"][v]) - P["y"][v])
                assert abs(ours - ref) <= 1e-6 * ref
        assert np.median(err) < 1e-6
        cerr = rel_err(r["cov"][ok], P["ref_pcov"][ok]).reshape(int(ok.sum()), -1).max(axis=1)
        assert np.nanmedian(cerr) < 1e-3


# --------------------------------------------------------------------------- curve_fit extras
_EXTRA_CASES = {
    "sigma": (dict(sigma="S"), dict(weights="W")),
    "sigma_absolute": (dict(sigma="S", absolute_sigma=True), dict(weights="W", absolute_sigma=True)),
    "diff_step": (dict(diff_step=1e-6), dict(diff_step=1e-6)),
    "soft_l1": (dict(loss="soft_l1", f_scale=20.0), dict(loss=1, f_scale=20.0)),
    "huber": (dict(loss="huber", f_scale=20.0), dict(loss=2, f_scale=20.0)),
    "cauchy": (dict(loss="cauchy", f_scale=30.0), dict(loss=3, f_scale=30.0)),
    "arctan": (dict(loss="arctan", f_scale=60.0), dict(loss=4, f_scale=60.0)),
    "sigma_huber": (dict(sigma="S", loss="huber", f_scale=1.5), dict(weights="W", loss=2, f_scale=1.5)),
}


def extras_problem(n=48, seed=3):
    """Bi-exponential (S0) voxels with 2 % noise and a planted outlier on every seventh voxel."""
    rng = np.random.default_rng(seed)
    b = np.array([0, 5, 10, 20, 30, 40, 50, 75, 100, 150, 200, 300, 400, 500, 650, 800.0])
    P = np.stack([rng.uniform(0.1, 0.4, n), rng.uniform(5e-4, 2e-3, n), rng.uniform(0.01, 0.08, n),
                  rng.uniform(500, 1500, n)], 1)
    y = np.stack([p[3] * (p[0] * np.exp(-b * p[1]) + (1 - p[0]) * np.exp(-b * p[2])) for p in P])
    y *= 1 + 0.02 * rng.standard_normal(y.shape)
    y[::7, 5] *= 1.6
    return b, y, np.array([0.2, 1e-3, 0.02, 1000.0]), np.array([0.01, 1e-5, 3e-3, 1.0]), np.array([0.99, 3e-3, 0.3, 5e3]), \
        np.linspace(5, 40, 16)


def scipy_extras(b, y, p0, lb, ub, kw):
    from scipy.optimize import curve_fit

    f = lambda b_, f1, D1, D2, S0: S0 * (f1 * np.exp(-b_ * D1) + (1 - f1) * np.exp(-b_ * D2))  # noqa: E731
    fits = [curve_fit(f, b, yi, p0=p0, bounds=(lb, ub), method="trf", maxfev=250, ftol=1e-8, **kw) for yi in y]
    return np.array([p for p, _ in fits]), np.array([c for _, c in fits])


@pytest.mark.parametrize("case", sorted(_EXTRA_CASES))
def test_curve_fit_extras_core_equals_scipy(case):
    """`sigma` / `absolute_sigma` (curve_fit) and `loss` / `f_scale` / `diff_step` (least_squares), which the
    reference forwards from its solver kwargs (solvers/curvefit.py:70-73, 305): the EXTRAS instantiation of the
    device core, compiled for the host, against SciPy itself."""
    b, y, p0, lb, ub, sigma = extras_problem()
    kw_scipy, kw_core = _EXTRA_CASES[case]
    kw_scipy = {k: (sigma if isinstance(v, str) and v == "S" else v) for k, v in kw_scipy.items()}
    kw_core = {k: (1.0 / sigma if isinstance(v, str) and v == "W" else v) for k, v in kw_core.items()}
    ref_p, ref_c = scipy_extras(b, y, p0, lb, ub, kw_scipy)
    n = y.shape[0]
    r = hostsim.trf_fit_extras(3, b, y, np.tile(p0, (n, 1)), np.tile(lb, (n, 1)), np.tile(ub, (n, 1)), **kw_core)
    assert (r["status"] > 0).all()
    assert (np.abs(r["params"] - ref_p) / np.abs(ref_p)).max() <= 1e-4
    d = np.sqrt(np.einsum("vii->vi", ref_c))
    assert (np.abs(r["cov"] - ref_c) / (d[:, :, None] * d[:, None, :])).max() <= 1e-4   # on the scale of the variances


@pytest.mark.parametrize("model_id,case", [(0, "sigma_huber"), (0, "diff_step"), (4, "huber"), (4, "sigma")])
def test_curve_fit_extras_core_other_models(model_id, case):
    """The EXTRAS evaluation is a template over the model: mono-exponential [S0, D] (id 0) and the reduced
    tri-exponential (id 4) against SciPy, like the bi-exponential above."""
    from scipy.optimize import curve_fit

    rng = np.random.default_rng(11)
    b = np.array([0, 5, 10, 20, 30, 40, 50, 75, 100, 150, 200, 300, 400, 500, 650, 800.0])
    n = 24
    if model_id == 0:   # [S0, D]
        f = lambda b_, S0, D: S0 * np.exp(-b_ * D)  # noqa: E731
        P = np.stack([rng.uniform(500, 1500, n), rng.uniform(5e-4, 3e-3, n)], 1)
        p0, lb, ub = np.array([1000.0, 1e-3]), np.array([1.0, 1e-5]), np.array([5e3, 0.1])
        scale = 1000.0
    else:               # [f1, D1, f2, D2, D3]
        f = lambda b_, f1, D1, f2, D2, D3: f1 * np.exp(-b_ * D1) + f2 * np.exp(-b_ * D2) + (1 - f1 - f2) * np.exp(-b_ * D3)  # noqa: E731
        P = np.stack([rng.uniform(0.3, 0.5, n), rng.uniform(5e-4, 1.5e-3, n), rng.uniform(0.2, 0.3, n),
                      rng.uniform(5e-3, 1e-2, n), rng.uniform(0.05, 0.15, n)], 1)
        p0 = np.array([0.4, 1e-3, 0.25, 8e-3, 0.1])
        lb, ub = np.array([0.01, 1e-4, 0.01, 3e-3, 0.03]), np.array([0.9, 3e-3, 0.9, 3e-2, 0.5])
        scale = 1.0
    y = np.stack([f(b, *p) for p in P]) * (1 + 0.01 * rng.standard_normal((n, 16)))
    y[::5, 4] *= 1.3
    sigma = np.linspace(0.005, 0.04, 16) * scale
    kw_scipy, kw_core = _EXTRA_CASES[case]
    fs = {"f_scale": 0.02 * scale} if "loss" in kw_scipy and "sigma" not in kw_scipy else {}
    kw_scipy = {**{k: (sigma if isinstance(v, str) and v == "S" else v) for k, v in kw_scipy.items()}, **fs}
    kw_core = {**{k: (1.0 / sigma if isinstance(v, str) and v == "W" else v) for k, v in kw_core.items()}, **fs}
    fits = [curve_fit(f, b, yi, p0=p0, bounds=(lb, ub), method="trf", maxfev=250, ftol=1e-8, **kw_scipy) for yi in y]
    ref = np.array([p for p, _ in fits])
    r = hostsim.trf_fit_extras(model_id, b, y, np.tile(p0, (n, 1)), np.tile(lb, (n, 1)), np.tile(ub, (n, 1)), **kw_core)
    ok = r["status"] > 0
    assert ok.all()
    # residual norms agree; parameters agree wherever the minimiser is well determined
    res = lambda P_: np.array([np.linalg.norm(f(b, *p) - yi) for p, yi in zip(P_, y)])  # noqa: E731
    assert np.allclose(res(r["params"]), res(ref), rtol=1e-6)
    assert np.median(np.abs(r["params"] - ref) / np.abs(ref)) <= 1e-6
