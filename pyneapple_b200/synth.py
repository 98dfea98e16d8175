"""Seeded synthetic DWI volumes for the five BASELINE.json configurations.

Recipe: parameter fields are smooth in space (a low-order
polynomial of the normalised voxel coordinates) plus 5 % uniform jitter, so
that the IDEAL multi-resolution driver sees something worth interpolating;
Gaussian noise sigma = 10 on S0-scaled data (SNR ~ 100 at b = 0) or 0.01 on
normalised data.  Everything is FP64 and generated from
``np.random.default_rng(seed)`` in a fixed draw order, so any sub-range of
z-slices can be generated independently (each slab has its own child seed)
and the full 4.19 M-voxel volumes never need to be shipped anywhere.
"""

from __future__ import annotations

from dataclasses import dataclass

import numpy as np

B16 = np.array(
    [0, 25, 50, 75, 100, 150, 200, 300, 400, 500, 600, 700, 800, 900, 1000, 1200],
    dtype=np.float64,
)
B24 = np.array(
    [0, 5, 10, 15, 20, 30, 40, 50, 60, 75, 100, 125, 150, 200, 250, 300, 400, 500,
     600, 700, 800, 1000, 1200, 1500],
    dtype=np.float64,
)


@dataclass(frozen=True)
class Config:
    """One BASELINE.json workload: shape, model, p0/bounds and truth ranges."""

    name: str
    seed: int
    shape: tuple  # (X, Y, Z)
    bvalues: np.ndarray
    model: str  # "monoexp" | "biexp" | "triexp" | "nnls"
    model_kwargs: dict
    solver: str  # "curvefit" | "constrained_curvefit" | "nnls"
    p0: dict
    bounds: dict
    truth: dict  # name -> (lo, hi)
    noise_sigma: float
    solver_kwargs: dict


CONFIGS = {
    # examples/configs/monoexp_pixelwise.toml
    "C1": Config(
        "C1", 1, (128, 128, 32), B16, "monoexp", {}, "curvefit",
        {"S0": 1000.0, "D": 0.001},
        {"S0": (1.0, 5000.0), "D": (1e-5, 0.1)},
        {"S0": (500.0, 1500.0), "D": (5e-4, 3e-3)},
        10.0, {"max_iter": 250, "tol": 1e-8},
    ),
    # examples/parameters/ideal_biexp.toml:27-37 with fit_s0 = true
    "C2": Config(
        "C2", 2, (256, 256, 64), B16, "biexp", {"fit_s0": True}, "curvefit",
        {"f1": 0.2, "D1": 0.001, "D2": 0.02, "S0": 1000.0},
        {"f1": (0.01, 0.99), "D1": (1e-5, 0.003), "D2": (0.003, 0.3), "S0": (1.0, 5000.0)},
        {"f1": (0.05, 0.4), "D1": (5e-4, 2.5e-3), "D2": (0.01, 0.1), "S0": (500.0, 1500.0)},
        10.0, {"max_iter": 250, "tol": 1e-8},
    ),
    # examples/configs/nnls_example.toml on the C2 signals
    "C3": Config(
        "C3", 3, (256, 256, 64), B16, "nnls", {"d_range": (0.0008, 0.5), "n_bins": 250}, "nnls",
        {}, {},
        {"f1": (0.05, 0.4), "D1": (5e-4, 2.5e-3), "D2": (0.01, 0.1), "S0": (500.0, 1500.0)},
        10.0, {"reg_order": 2, "mu": 0.02, "max_iter": 250, "tol": 1e-8},
    ),
    # C2 volume + ellipsoid mask, IDEAL driver
    "C4": Config(
        "C4", 4, (256, 256, 64), B16, "biexp", {"fit_s0": True}, "curvefit",
        {"f1": 0.2, "D1": 0.001, "D2": 0.02, "S0": 1000.0},
        {"f1": (0.01, 0.99), "D1": (1e-5, 0.003), "D2": (0.003, 0.3), "S0": (1.0, 5000.0)},
        {"f1": (0.05, 0.4), "D1": (5e-4, 2.5e-3), "D2": (0.01, 0.1), "S0": (500.0, 1500.0)},
        10.0, {"max_iter": 250, "tol": 1e-8},
    ),
    # triexp reduced, constrained, 24 b-values, normalised signal
    "C5": Config(
        "C5", 5, (512, 512, 128), B24, "triexp", {}, "constrained_curvefit",
        {"f1": 0.15, "D1": 0.1, "f2": 0.25, "D2": 0.01, "D3": 0.001},
        {"f1": (0.0, 1.0), "D1": (0.03, 0.5), "f2": (0.0, 1.0), "D2": (0.003, 0.03),
         "D3": (1e-4, 0.003)},
        {"f1": (0.05, 0.25), "D1": (0.05, 0.2), "f2": (0.1, 0.35), "D2": (5e-3, 2e-2),
         "D3": (5e-4, 2e-3)},
        0.01, {"max_iter": 250, "tol": 1e-8, "fraction_constraint": True},
    ),
}

IDEAL_C4 = {
    "dim_steps": [[16, 16], [32, 32], [64, 64], [128, 128], [256, 256]],
    "step_tol": {"S0": 0.5, "f1": 0.2, "D1": 0.2, "D2": 0.2},
    "segmentation_threshold": 0.2,
    "interpolation_method": "cubic",
    "ideal_dims": 2,
}


def _smooth_field(rng, xs, ys, zs, lo, hi):
    """A quadratic polynomial of the normalised coordinates mapped into [lo, hi]."""
    c = rng.uniform(-1.0, 1.0, size=10)
    x, y, z = xs[:, None, None], ys[None, :, None], zs[None, None, :]
    poly = (
        c[0] + c[1] * x + c[2] * y + c[3] * z + c[4] * x * y + c[5] * y * z
        + c[6] * x * z + c[7] * x * x + c[8] * y * y + c[9] * z * z
    )
    # the polynomial's range over the unit cube is bounded by sum |c_i|
    span = np.abs(c).sum()
    u = 0.5 + 0.5 * poly / span  # in [0, 1]
    return lo + (hi - lo) * u


def truth_fields(cfg: Config, z0: int = 0, z1: int | None = None, replica: int = 0) -> dict:
    """Ground-truth parameter maps ``name -> (X, Y, z1-z0)`` for a z-slab.

    ``replica > 0``: another realisation of the same volume — the same smooth parameter fields, its
    own jitter (and, in :func:`make_volume`, noise) streams; replica 0 is the configuration itself."""
    X, Y, Z = cfg.shape
    z1 = Z if z1 is None else z1
    xs = np.linspace(0.0, 1.0, X)
    ys = np.linspace(0.0, 1.0, Y)
    zs = np.linspace(0.0, 1.0, Z)[z0:z1]
    field_rng = np.random.default_rng([cfg.seed, 0])
    out = {}
    for name, (lo, hi) in cfg.truth.items():
        base = _smooth_field(field_rng, xs, ys, zs, lo, hi)
        out[name] = base
    # per-slice jitter streams so that slabs are reproducible independently
    for zi in range(z0, z1):
        jr = np.random.default_rng([cfg.seed, 1, zi] + ([replica] if replica else []))
        for name, (lo, hi) in cfg.truth.items():
            jit = 1.0 + 0.05 * jr.uniform(-1.0, 1.0, size=(X, Y))
            out[name][:, :, zi - z0] = np.clip(out[name][:, :, zi - z0] * jit, lo, hi)
    if cfg.name == "C5":
        # 5 % sub-population with f1 + f2 -> 1 so the inequality is active
        for zi in range(z0, z1):
            jr = np.random.default_rng([cfg.seed, 2, zi] + ([replica] if replica else []))
            hit = jr.uniform(size=(X, Y)) < 0.05
            s = out["f1"][:, :, zi - z0] + out["f2"][:, :, zi - z0]
            scale = np.where(hit, 1.0 / s, 1.0)
            out["f1"][:, :, zi - z0] *= scale
            out["f2"][:, :, zi - z0] *= scale
    return out


def make_volume(cfg: Config | str, z0: int = 0, z1: int | None = None, noise: bool = True, replica: int = 0):
    """Return ``(bvalues, image[X, Y, z1-z0, n_b], truth)`` for a z-slab of a config (``replica``:
    see :func:`truth_fields`)."""
    if isinstance(cfg, str):
        cfg = CONFIGS[cfg]
    X, Y, Z = cfg.shape
    z1 = Z if z1 is None else z1
    t = truth_fields(cfg, z0, z1, replica)
    b = cfg.bvalues
    if cfg.model == "monoexp":
        img = t["S0"][..., None] * np.exp(-b * t["D"][..., None])
    elif cfg.model in ("biexp", "nnls"):
        f1 = t["f1"][..., None]
        img = t["S0"][..., None] * (
            f1 * np.exp(-b * t["D1"][..., None]) + (1 - f1) * np.exp(-b * t["D2"][..., None])
        )
    elif cfg.model == "triexp":
        f1 = t["f1"][..., None]
        f2 = t["f2"][..., None]
        img = (
            f1 * np.exp(-b * t["D1"][..., None])
            + f2 * np.exp(-b * t["D2"][..., None])
            + (1 - f1 - f2) * np.exp(-b * t["D3"][..., None])
        )
    else:
        raise ValueError(cfg.model)
    if noise:
        for zi in range(z0, z1):
            nr = np.random.default_rng([cfg.seed, 3, zi] + ([replica] if replica else []))
            img[:, :, zi - z0, :] += nr.normal(0.0, cfg.noise_sigma, size=(X, Y, b.shape[0]))
    return b.copy(), np.ascontiguousarray(img), t


def ellipsoid_mask(shape, z0: int = 0, z1: int | None = None) -> np.ndarray:
    """Integer-label ellipsoid covering ~50 % of the volume (config C4)."""
    X, Y, Z = shape
    z1 = Z if z1 is None else z1
    x = (np.arange(X) - (X - 1) / 2) / (X / 2)
    y = (np.arange(Y) - (Y - 1) / 2) / (Y / 2)
    z = (np.arange(Z)[z0:z1] - (Z - 1) / 2) / (Z / 2)
    r2 = (x[:, None, None] / 0.98) ** 2 + (y[None, :, None] / 0.98) ** 2 + (z[None, None, :] / 1.0) ** 2
    return (r2 <= 1.0).astype(np.int64)


def sample_voxels(cfg: Config | str, n_vox: int, z: int = 0):
    """A deterministic ``n_vox``-voxel sample of a config: every k-th voxel of slice ``z``.

    Returns ``(bvalues, ydata[n_vox, n_b], flat_index)``; only that slice is
    generated, so the sample is cheap at any volume size.  Voxels of one slice
    are i.i.d. with the rest of the volume up to the smooth z-trend.
    """
    if isinstance(cfg, str):
        cfg = CONFIGS[cfg]
    b, img, _ = make_volume(cfg, z, z + 1)
    flat = img.reshape(-1, b.shape[0])
    if n_vox > flat.shape[0]:
        raise ValueError(f"slice holds {flat.shape[0]} voxels, asked for {n_vox}")
    idx = np.arange(n_vox) * (flat.shape[0] // n_vox)
    return b, np.ascontiguousarray(flat[idx]), idx


def make_volume_device(cfg: Config | str, z0: int = 0, z1: int | None = None, device="cuda:0", replica: int = 0):
    """The recipe of :func:`make_volume` evaluated on a GPU (torch, float64): the same smooth
    parameter fields (identical polynomial coefficients), jitter and noise from torch's generator
    seeded per slab — the same distribution, not the same random numbers.  For workloads too large
    to synthesise on the host in reasonable time (config C5: 805 M samples); benchmarks only, the
    parity tests use :func:`make_volume`.  Returns ``(bvalues, image (X, Y, z1-z0, n_b) tensor)``.
    """
    import torch

    if isinstance(cfg, str):
        cfg = CONFIGS[cfg]
    X, Y, Z = cfg.shape
    z1 = Z if z1 is None else z1
    dev = torch.device(device)
    f64 = dict(dtype=torch.float64, device=dev)
    xs = torch.linspace(0.0, 1.0, X, **f64)[:, None, None]
    ys = torch.linspace(0.0, 1.0, Y, **f64)[None, :, None]
    zs = torch.linspace(0.0, 1.0, Z, **f64)[z0:z1][None, None, :]
    field_rng = np.random.default_rng([cfg.seed, 0])
    gen = torch.Generator(device=dev)
    gen.manual_seed(int(cfg.seed) * 1_000_003 + 7919 * z0 + 104_729 * replica + 1)
    t = {}
    for name, (lo, hi) in cfg.truth.items():
        c = field_rng.uniform(-1.0, 1.0, size=10)
        poly = (c[0] + c[1] * xs + c[2] * ys + c[3] * zs + c[4] * xs * ys + c[5] * ys * zs
                + c[6] * xs * zs + c[7] * xs * xs + c[8] * ys * ys + c[9] * zs * zs)
        u = 0.5 + 0.5 * poly / float(np.abs(c).sum())
        jit = 1.0 + 0.05 * (2.0 * torch.rand(u.shape, generator=gen, **f64) - 1.0)
        t[name] = torch.clamp((lo + (hi - lo) * u) * jit, lo, hi)
    if cfg.name == "C5":
        hit = torch.rand(t["f1"].shape, generator=gen, **f64) < 0.05
        s = t["f1"] + t["f2"]
        scale = torch.where(hit, 1.0 / s, torch.ones_like(s))
        t["f1"] = t["f1"] * scale
        t["f2"] = t["f2"] * scale
    b = torch.as_tensor(cfg.bvalues, **f64)
    e = lambda name: torch.exp(-b * t[name][..., None])  # noqa: E731
    if cfg.model == "monoexp":
        img = t["S0"][..., None] * e("D")
    elif cfg.model in ("biexp", "nnls"):
        f1 = t["f1"][..., None]
        img = t["S0"][..., None] * (f1 * e("D1") + (1 - f1) * e("D2"))
    elif cfg.model == "triexp":
        f1, f2 = t["f1"][..., None], t["f2"][..., None]
        img = f1 * e("D1") + f2 * e("D2") + (1 - f1 - f2) * e("D3")
    else:
        raise ValueError(cfg.model)
    img = img + cfg.noise_sigma * torch.randn(img.shape, generator=gen, **f64)
    return cfg.bvalues.copy(), img.contiguous()
