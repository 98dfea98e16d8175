"""Shared helpers for the parity tests."""

from __future__ import annotations

import os

import numpy as np

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
GOLD = os.path.join(ROOT, "tests", "golden")

# golden file -> (kind, mode) of the model that produced it
TRF_CASES = {
    "trf_mono_c1": ("monoexp", "s0"),
    "trf_biexp_s0_c2": ("biexp", "s0"),
    "trf_biexp_reduced": ("biexp", "reduced"),
    "trf_biexp_full": ("biexp", "full"),
    "trf_triexp_reduced": ("triexp", "reduced"),
    "trf_triexp_s0": ("triexp", "s0"),
    "trf_triexp_full": ("triexp", "full"),
    "trf_biexp_s0_pixfixed_D1": ("biexp", "s0"),
    "trf_biexp_s0_modelfixed_D2": ("biexp", "s0"),
    "trf_biexp_s0_pervoxel": ("biexp", "s0"),
    "trf_biexp_s0_degenerate": ("biexp", "s0"),
    "trf_mono_degenerate": ("monoexp", "s0"),
    "trf_biexp_s0_maxiter1": ("biexp", "s0"),
    "trf_biexp_s0_maxiter2": ("biexp", "s0"),
    "trf_biexp_s0_maxiter3": ("biexp", "s0"),
    "trf_biexp_s0_maxiter5": ("biexp", "s0"),
    "trf_biexp_s0_badbounds": ("biexp", "s0"),
}
# method = "dogbox" goldens (oracle/make_golden.py: case_dogbox)
DBX_CASES = {
    "dbx_biexp_s0_c2": ("biexp", "s0"),
    "dbx_biexp_s0_onbound": ("biexp", "s0"),
    "dbx_biexp_s0_maxiter1": ("biexp", "s0"),
    "dbx_biexp_s0_maxiter2": ("biexp", "s0"),
    "dbx_biexp_s0_maxiter3": ("biexp", "s0"),
    "dbx_biexp_s0_maxiter6": ("biexp", "s0"),
    "dbx_biexp_s0_pixfixed_D1": ("biexp", "s0"),
    "dbx_mono_c1": ("monoexp", "s0"),
    "dbx_triexp_reduced": ("triexp", "reduced"),
}
# T1 / STEAM goldens (oracle/make_golden.py: case_t1); t1_mode / tr / tm are stored in the file
TRF_T1_CASES = {
    "trf_mono_t1": ("monoexp", "s0"),
    "trf_mono_t1_pixfixed": ("monoexp", "s0"),
    "trf_mono_t1_modelfixed": ("monoexp", "s0"),
    "trf_mono_steam": ("monoexp", "s0"),
    "trf_mono_steam_pixfixed": ("monoexp", "s0"),
    "trf_mono_steam_modelfixed": ("monoexp", "s0"),
    "trf_biexp_reduced_t1": ("biexp", "reduced"),
    "trf_biexp_s0_steam_pixfixed": ("biexp", "s0"),
    "trf_biexp_full_t1_modelfixed": ("biexp", "full"),
    "trf_triexp_reduced_steam": ("triexp", "reduced"),
    "trf_triexp_s0_t1_pixfixed": ("triexp", "s0"),
    "trf_triexp_full_steam_modelfixed": ("triexp", "full"),
}
DBX_T1_CASES = {
    "dbx_mono_t1_pixfixed": ("monoexp", "s0"),
    "dbx_mono_steam_pixfixed": ("monoexp", "s0"),
    "dbx_biexp_reduced_t1": ("biexp", "reduced"),
    "dbx_biexp_s0_steam_pixfixed": ("biexp", "s0"),
    "dbx_triexp_reduced_steam_pixfixed": ("triexp", "reduced"),
}
# method = "lm" goldens (oracle/make_golden.py: case_lm): unbounded problems through MINPACK
LM_CASES = {
    "lm_mono_c1": ("monoexp", "s0"),
    "lm_biexp_s0_c2": ("biexp", "s0"),
    "lm_biexp_s0_maxiter3": ("biexp", "s0"),
    "lm_biexp_s0_maxiter8": ("biexp", "s0"),
    "lm_biexp_s0_maxiter14": ("biexp", "s0"),
    "lm_biexp_s0_maxiter30": ("biexp", "s0"),
    "lm_biexp_s0_pixfixed_D1": ("biexp", "s0"),
    "lm_biexp_reduced": ("biexp", "reduced"),
    "lm_triexp_reduced": ("triexp", "reduced"),
    "lm_biexp_s0_degenerate": ("biexp", "s0"),
}
# T1 fitted next to a free amplitude: S0 and T1 enter only through S0 * C(T1), so the minimiser is a
# curve; parity is checked on the identifiable quantities (D, S0 * C(T1)) and the residual
T1_AMPLITUDE_ONLY = {"trf_mono_t1", "trf_mono_steam"}
MODEL_FIXED = {"trf_biexp_s0_modelfixed_D2": {"D2": 0.03},
               "trf_mono_t1_modelfixed": {"T1": 1200.0}, "trf_mono_steam_modelfixed": {"T1": 1200.0},
               "trf_biexp_full_t1_modelfixed": {"T1": 1300.0}, "trf_triexp_full_steam_modelfixed": {"T1": 1100.0}}


def t1_kwargs(name):
    """Model keyword arguments (fit_t1 / fit_t1_steam / repetition_time / mixing_time) of a golden."""
    g = load(name)
    if "t1_mode" not in g.files:
        return {}
    kw = {"repetition_time": float(g["tr"])}
    if int(g["t1_mode"]) == 2:
        kw.update(fit_t1_steam=True, mixing_time=float(g["tm"]))
    else:
        kw["fit_t1"] = True
    return kw


def load(name):
    return np.load(os.path.join(GOLD, name + ".npz"), allow_pickle=False)


def full_problem(name):
    """Expand a golden TRF case into full-parameter-vector arrays.

    Returns dict(b, y, all_names, free_names, P0/LB/UB (n_vox, n_all), frozen
    (n_all,), max_iter, tol, ref_params (n_vox, n_free), ref_pcov, ref_success,
    messages, uses_fd) where frozen entries of P0 carry the fixed value.
    """
    g = load(name)
    all_names = [str(s) for s in g["all_names"]]
    model_free = [str(s) for s in g["param_names"]]
    y = g["y"]
    n = y.shape[0]
    na = len(all_names)
    P0 = np.zeros((n, na))
    LB = np.full((n, na), -np.inf)
    UB = np.full((n, na), np.inf)
    frozen = np.zeros(na, np.int32)
    pix_fixed = {k[6:]: g[k] for k in g.files if k.startswith("fixed_")}
    mfixed = MODEL_FIXED.get(name, {})
    for j, nm in enumerate(model_free):
        col = all_names.index(nm)
        if "p0_arr" in g.files:
            P0[:, col], LB[:, col], UB[:, col] = g["p0_arr"][j], g["lb_arr"][j], g["ub_arr"][j]
        else:
            P0[:, col], LB[:, col], UB[:, col] = g["p0"][j], g["lb"][j], g["ub"][j]
    for nm, val in mfixed.items():
        col = all_names.index(nm)
        P0[:, col] = val
        frozen[col] = 1
    for nm, arr in pix_fixed.items():
        col = all_names.index(nm)
        P0[:, col] = arr
        frozen[col] = 1
    free_names = [nm for j, nm in enumerate(all_names) if not frozen[j]]
    return dict(
        b=g["b"], y=y, all_names=all_names, free_names=free_names, P0=P0, LB=LB, UB=UB,
        frozen=frozen, max_iter=int(g["max_iter"]), tol=float(g["tol"]),
        ref_params=g["params"].T, ref_pcov=g["pcov"], ref_success=g["success"],
        messages=[str(s) for s in g["messages"]],
        uses_fd=not (pix_fixed or mfixed), pix_fixed=pix_fixed, mfixed=mfixed,
        p0_vec=g["p0"], lb_vec=g["lb"], ub_vec=g["ub"],
        per_voxel="p0_arr" in g.files,
        t1_mode=int(g["t1_mode"]) if "t1_mode" in g.files else 0,
        tr=float(g["tr"]) if "tr" in g.files else 0.0, tm=float(g["tm"]) if "tm" in g.files else 0.0,
    )


def rel_err(a, ref):
    return np.abs(a - ref) / np.maximum(np.abs(ref), 1e-300)
