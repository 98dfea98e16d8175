"""Host-side signal models and the device model descriptor.

The B200 engine does not evaluate models on the host during a fit; these
classes exist so that (a) the solvers can be constructed and driven exactly
like the reference's (``model.param_names``, ``model.fixed_params``,
``forward`` / ``jacobian`` for ``predict`` and single-voxel checks) on a box
where Pyneapple itself is not installed, and (b) a Pyneapple model object
handed to a B200 solver (drop-in use) can be translated to the same device
descriptor by duck typing (:func:`describe_model`).

Reference interface mirrored here (behaviour, names, parameter order):
  * ``models/base.py:96-230``  ParametricModel (fixed-parameter injection)
  * ``models/monoexp.py:91-163``, ``models/biexp.py:103-221``,
    ``models/triexp.py:103-247``  parameter order per mode, forward, Jacobian
  * ``model_functions/multiexp.py:35-302``  signal equations and T1 factors
  * ``models/nnls.py`` + ``model_functions/nnls.py:17-43``  bins and basis

All multi-exponential variants are one family here,

    S(b) = A * C(T1) * sum_k w_k * exp(-b * D_k)

with ``A`` the amplitude (parameter ``S0`` or the constant 1), ``w_k`` either
free fractions or ``1 - sum(others)`` for the last component of a *reduced*
model, and ``C`` the optional T1 / STEAM factor.  A :class:`ModelDesc` holds
the slot table that says which entry of the full parameter vector plays
which role; the CUDA kernels are specialised on exactly this table
(``csrc/pnb_models.cuh``).
"""

from __future__ import annotations

from dataclasses import dataclass, field
from typing import Any, Sequence

import numpy as np

# device model ids -- keep in sync with include/pyneapple_b200.h
MODEL_MONO = 0  # [S0, D]
MODEL_BI_REDUCED = 1  # [f1, D1, D2]
MODEL_BI_FULL = 2  # [f1, D1, f2, D2]
MODEL_BI_S0 = 3  # [f1, D1, D2, S0]
MODEL_TRI_REDUCED = 4  # [f1, D1, f2, D2, D3]
MODEL_TRI_FULL = 5  # [f1, D1, f2, D2, f3, D3]
MODEL_TRI_S0 = 6  # [f1, D1, f2, D2, D3, S0]

T1_NONE = 0
T1_STANDARD = 1  # * (1 - exp(-TR/T1))
T1_STEAM = 2  # * (1 - exp(-TR/T1)) * exp(-TM/T1)

_BASE_NAMES = {
    MODEL_MONO: ["S0", "D"],
    MODEL_BI_REDUCED: ["f1", "D1", "D2"],
    MODEL_BI_FULL: ["f1", "D1", "f2", "D2"],
    MODEL_BI_S0: ["f1", "D1", "D2", "S0"],
    MODEL_TRI_REDUCED: ["f1", "D1", "f2", "D2", "D3"],
    MODEL_TRI_FULL: ["f1", "D1", "f2", "D2", "f3", "D3"],
    MODEL_TRI_S0: ["f1", "D1", "f2", "D2", "D3", "S0"],
}

# slot tables: (#components, index of D_k per component, index of the free
# fraction per component (-1 = implied 1 - sum), index of the amplitude (-1 = 1))
_SLOTS = {
    MODEL_MONO: (1, (1,), (None,), 0),
    MODEL_BI_REDUCED: (2, (1, 2), (0, -1), -1),
    MODEL_BI_FULL: (2, (1, 3), (0, 2), -1),
    MODEL_BI_S0: (2, (1, 2), (0, -1), 3),
    MODEL_TRI_REDUCED: (3, (1, 3, 4), (0, 2, -1), -1),
    MODEL_TRI_FULL: (3, (1, 3, 5), (0, 2, 4), -1),
    MODEL_TRI_S0: (3, (1, 3, 4), (0, 2, -1), 5),
}


@dataclass(frozen=True)
class ModelDesc:
    """What the device needs to know about a parametric model."""

    model_id: int
    t1_mode: int = T1_NONE
    repetition_time: float = 0.0
    mixing_time: float = 0.0
    all_names: tuple = ()
    # model-level scalar fixed parameters, name -> value
    fixed: dict = field(default_factory=dict)

    @property
    def n_all(self) -> int:
        return len(self.all_names)

    @property
    def free_names(self) -> list[str]:
        return [n for n in self.all_names if n not in self.fixed]


def _all_names(model_id: int, t1_mode: int) -> list[str]:
    names = list(_BASE_NAMES[model_id])
    if t1_mode != T1_NONE:
        names.append("T1")
    return names


def family_forward(desc: ModelDesc, xdata: np.ndarray, params: Sequence) -> np.ndarray:
    """Evaluate the signal for full parameter vector(s).

    ``params[i]`` may be scalars (one voxel -> ``(n_b,)``) or arrays of shape
    ``(n_vox,)`` (batched -> ``(n_vox, n_b)``).
    """
    n_comp, d_idx, f_idx, a_idx = _SLOTS[desc.model_id]
    p = [np.asarray(v, dtype=np.float64) for v in params]
    batched = p[0].ndim > 0
    b = np.asarray(xdata, dtype=np.float64)
    if batched:
        p = [v[:, None] for v in p]
        b = b[None, :]
    if desc.model_id == MODEL_MONO:
        sig = p[0] * np.exp(-b * p[1])
    else:
        implied = 1.0
        for k in range(n_comp):
            if f_idx[k] >= 0:
                implied = implied - p[f_idx[k]]
        sig = 0.0
        for k in range(n_comp):
            w = p[f_idx[k]] if f_idx[k] >= 0 else implied
            sig = sig + w * np.exp(-b * p[d_idx[k]])
        if a_idx >= 0:
            sig = p[a_idx] * sig
    if desc.t1_mode != T1_NONE:
        t1 = p[-1]
        sig = sig * (1 - np.exp(-desc.repetition_time / t1))
        if desc.t1_mode == T1_STEAM:
            sig = sig * np.exp(-desc.mixing_time / t1)
    return sig


def family_jacobian(desc: ModelDesc, xdata: np.ndarray, params: Sequence) -> np.ndarray:
    """Analytic Jacobian ``(n_b, n_all)`` for one voxel's full parameter vector."""
    n_comp, d_idx, f_idx, a_idx = _SLOTS[desc.model_id]
    b = np.asarray(xdata, dtype=np.float64)
    p = [float(v) for v in params]
    n_base = len(_BASE_NAMES[desc.model_id])
    jac = np.zeros((b.shape[0], n_base))
    if desc.model_id == MODEL_MONO:
        e = np.exp(-b * p[1])
        jac[:, 0] = e
        jac[:, 1] = -b * p[0] * e
        base = p[0] * e
    else:
        amp = p[a_idx] if a_idx >= 0 else 1.0
        exps = [np.exp(-b * p[d_idx[k]]) for k in range(n_comp)]
        implied = 1.0 - sum(p[f_idx[k]] for k in range(n_comp) if f_idx[k] >= 0)
        w = [p[f_idx[k]] if f_idx[k] >= 0 else implied for k in range(n_comp)]
        has_implied = f_idx[-1] < 0
        shape = sum(w[k] * exps[k] for k in range(n_comp))
        for k in range(n_comp):
            jac[:, d_idx[k]] = -b * amp * w[k] * exps[k]
            if f_idx[k] >= 0:
                jac[:, f_idx[k]] = amp * (exps[k] - exps[-1]) if has_implied else exps[k]
        if a_idx >= 0:
            jac[:, a_idx] = shape
        base = amp * shape
    if desc.t1_mode == T1_NONE:
        return jac
    t1 = p[-1]
    tr = desc.repetition_time
    e_tr = np.exp(-tr / t1)
    a = 1 - e_tr
    if desc.t1_mode == T1_STEAM:
        tm = desc.mixing_time
        e_tm = np.exp(-tm / t1)
        factor = a * e_tm
        d_t1 = base * e_tm / t1**2 * (-tr * e_tr + tm * a)
    else:
        factor = a
        d_t1 = base * (-e_tr * tr / t1**2)
    return np.column_stack((jac * factor, d_t1))


class ParametricModel:
    """Common behaviour of the mono/bi/tri-exponential host models."""

    _model_id: int = -1

    def __init__(
        self,
        fit_t1: bool = False,
        fit_t1_steam: bool = False,
        repetition_time: float | None = None,
        mixing_time: float | None = None,
        fixed_params: dict[str, float] | None = None,
        **model_kwargs: Any,
    ):
        if fit_t1_steam:
            fit_t1 = True
        if fit_t1 and repetition_time is None:
            raise ValueError("repetition_time is required when fit_t1=True.")
        if fit_t1_steam and mixing_time is None:
            raise ValueError("mixing_time is required when fit_t1_steam=True.")
        self.fit_t1 = fit_t1
        self.fit_t1_steam = fit_t1_steam
        self.repetition_time = repetition_time
        self.mixing_time = mixing_time
        self.model_kwargs = model_kwargs
        self.fixed_params: dict[str, float] = dict(fixed_params) if fixed_params else {}
        self._check_fixed()

    # -- descriptor ---------------------------------------------------
    @property
    def _t1_mode(self) -> int:
        if self.fit_t1_steam:
            return T1_STEAM
        return T1_STANDARD if self.fit_t1 else T1_NONE

    def _desc(self) -> ModelDesc:
        return ModelDesc(
            model_id=self._model_id,
            t1_mode=self._t1_mode,
            repetition_time=float(self.repetition_time or 0.0),
            mixing_time=float(self.mixing_time or 0.0),
            all_names=tuple(self._all_param_names),
            fixed=dict(self.fixed_params),
        )

    # -- names --------------------------------------------------------
    @property
    def _all_param_names(self) -> list[str]:
        return _all_names(self._model_id, self._t1_mode)

    @property
    def param_names(self) -> list[str]:
        return [n for n in self._all_param_names if n not in self.fixed_params]

    @property
    def n_params(self) -> int:
        return len(self.param_names)

    def _check_fixed(self) -> None:
        if not self.fixed_params:
            return
        names = self._all_param_names
        unknown = set(self.fixed_params) - set(names)
        if unknown:
            raise ValueError(
                f"Unknown fixed parameter(s): {sorted(unknown)}. Valid names: {names}"
            )
        if len(self.fixed_params) >= len(names):
            raise ValueError(
                "Cannot fix all parameters — at least one must remain free. "
                f"Fixed: {sorted(self.fixed_params)}, all: {names}"
            )

    # -- evaluation ---------------------------------------------------
    def forward(self, xdata: np.ndarray, *params: float) -> np.ndarray:
        return family_forward(self._desc(), xdata, params)

    def jacobian(self, xdata: np.ndarray, *params: float) -> np.ndarray:
        return family_jacobian(self._desc(), xdata, params)

    def residual(self, xdata, measured_signal, params):
        return measured_signal - self.forward(xdata, *params)

    def _free_indices(self, fixed: dict[str, float]) -> list[int]:
        return [i for i, n in enumerate(self._all_param_names) if n not in fixed]

    def _inject_fixed(self, free_params: tuple, fixed: dict[str, float]) -> tuple:
        if not fixed:
            return tuple(free_params)
        it = iter(free_params)
        return tuple(
            float(fixed[n]) if n in fixed else next(it) for n in self._all_param_names
        )

    def forward_with_fixed(self, xdata, fixed_dict, *free_params):
        return self.forward(xdata, *self._inject_fixed(free_params, fixed_dict))

    def jacobian_with_fixed(self, xdata, fixed_dict, *free_params):
        full = self.jacobian(xdata, *self._inject_fixed(free_params, fixed_dict))
        return full[:, self._free_indices(fixed_dict)]


class MonoExpModel(ParametricModel):
    """``S0 * exp(-b D)``; parameters ``[S0, D(, T1)]`` (models/monoexp.py:91-105)."""

    _model_id = MODEL_MONO


class BiExpModel(ParametricModel):
    """Bi-exponential IVIM model (models/biexp.py:103-126).

    reduced ``[f1, D1, D2]`` (default), S0 ``[f1, D1, D2, S0]``, full
    ``[f1, D1, f2, D2]``; ``+T1`` appended when T1 fitting is on.
    """

    def __init__(self, fit_reduced: bool = True, fit_s0: bool = False, **kw: Any):
        if fit_s0 and not fit_reduced:
            raise ValueError(
                "fit_s0=True requires fit_reduced=True. Full model with independent "
                "fractions and S0 is over-parameterized."
            )
        self.fit_reduced = fit_reduced
        self.fit_s0 = fit_s0
        super().__init__(**kw)

    @property
    def _model_id(self) -> int:  # type: ignore[override]
        if self.fit_s0:
            return MODEL_BI_S0
        return MODEL_BI_REDUCED if self.fit_reduced else MODEL_BI_FULL


class TriExpModel(ParametricModel):
    """Tri-exponential model (models/triexp.py:103-128).

    reduced ``[f1, D1, f2, D2, D3]`` (default), S0 ``[..., D3, S0]``, full
    ``[f1, D1, f2, D2, f3, D3]``.
    """

    def __init__(self, fit_reduced: bool = True, fit_s0: bool = False, **kw: Any):
        if fit_s0 and not fit_reduced:
            raise ValueError(
                "fit_s0=True requires fit_reduced=True. Full model with independent "
                "fractions and S0 is over-parameterized."
            )
        self.fit_reduced = fit_reduced
        self.fit_s0 = fit_s0
        super().__init__(**kw)

    @property
    def _model_id(self) -> int:  # type: ignore[override]
        if self.fit_s0:
            return MODEL_TRI_S0
        return MODEL_TRI_REDUCED if self.fit_reduced else MODEL_TRI_FULL


class NNLSModel:
    """Log-spaced diffusion spectrum model (models/nnls.py, model_functions/nnls.py:17-43)."""

    def __init__(self, d_range: tuple[float, float], n_bins: int, **model_kwargs: Any):
        self.d_range = d_range
        self.n_bins = n_bins
        self.model_kwargs = model_kwargs

    @property
    def bins(self) -> np.ndarray:
        return np.logspace(
            np.log10(self.d_range[0]), np.log10(self.d_range[1]), self.n_bins
        )

    def get_basis(self, xdata: np.ndarray) -> np.ndarray:
        xdata = np.asarray(xdata)
        if xdata.ndim != 1:
            raise ValueError(
                "xdata must be a 1D array of shape (n_measurements,), "
                f"but got shape {xdata.shape}"
            )
        return np.exp(-xdata.reshape(-1, 1) * self.bins.reshape(1, -1))

    def forward(self, xdata: np.ndarray, *spectrum: float) -> np.ndarray:
        return self.get_basis(xdata) @ np.asarray(spectrum)


def describe_model(model: Any) -> ModelDesc:
    """Translate a model object (ours or Pyneapple's) into a :class:`ModelDesc`.

    Pyneapple models are recognised by class name and their mode flags
    (``fit_reduced``, ``fit_s0``, ``fit_t1``, ``fit_t1_steam``,
    ``repetition_time``, ``mixing_time``, ``fixed_params``) — the attributes
    ``models/biexp.py:86-93`` sets — so no import of Pyneapple is needed.
    """
    if isinstance(model, ParametricModel):
        return model._desc()
    cls = type(model).__name__
    reduced = bool(getattr(model, "fit_reduced", True))
    s0 = bool(getattr(model, "fit_s0", False))
    if cls == "MonoExpModel":
        mid = MODEL_MONO
    elif cls == "BiExpModel":
        mid = MODEL_BI_S0 if s0 else (MODEL_BI_REDUCED if reduced else MODEL_BI_FULL)
    elif cls == "TriExpModel":
        mid = MODEL_TRI_S0 if s0 else (MODEL_TRI_REDUCED if reduced else MODEL_TRI_FULL)
    else:
        raise NotImplementedError(
            f"model {cls!r} has no B200 device implementation "
            "(supported: MonoExpModel, BiExpModel, TriExpModel)"
        )
    steam = bool(getattr(model, "fit_t1_steam", False))
    t1 = T1_STEAM if steam else (T1_STANDARD if getattr(model, "fit_t1", False) else T1_NONE)
    names = tuple(_all_names(mid, t1))
    ref_names = tuple(getattr(model, "_all_param_names", names))
    if ref_names != names:
        raise NotImplementedError(
            f"parameter layout {ref_names} of {cls} does not match device layout {names}"
        )
    return ModelDesc(
        model_id=mid,
        t1_mode=t1,
        repetition_time=float(getattr(model, "repetition_time", None) or 0.0),
        mixing_time=float(getattr(model, "mixing_time", None) or 0.0),
        all_names=names,
        fixed=dict(getattr(model, "fixed_params", None) or {}),
    )
