"""Thin Python layer over the C ABI: array checking, pointer passing, nothing numeric.

Two data paths, chosen by the type of ``ydata``:

* ``numpy.ndarray``  -> ``pnb_trf_fit_host``: host pointers, the library runs
  the chunked upload / solve / download pipeline; outputs are numpy arrays.
* ``torch.Tensor`` on a CUDA device -> ``pnb_trf_fit_device``: device pointers,
  work is enqueued on torch's current stream; outputs are CUDA tensors (this
  is what the IDEAL / segmented drivers and the multi-GPU path use, so that
  nothing round-trips through the host between levels / steps).
"""

from __future__ import annotations

import ctypes as C

import numpy as np

from . import _lib
from .models import ModelDesc

ST_LM_BOUNDED = -5  # not a kernel status: method='lm' on a bounded problem (see CurveFitSolver.fit)
STATUS_MESSAGES = {
    -5: "Method 'lm' only works for unconstrained problems. Use 'trf' or 'dogbox' instead.",
    -4: "Residuals are not finite in the initial point.",
    -3: "array must not contain infs or NaNs",
    -2: "Initial guess is outside of provided bounds",
    -1: "Each lower bound must be strictly less than each upper bound.",
    0: "Optimal parameters not found: The maximum number of function evaluations is exceeded.",
    1: "`gtol` termination condition is satisfied.",
    2: "`ftol` termination condition is satisfied.",
    3: "`xtol` termination condition is satisfied.",
    4: "Both `ftol` and `xtol` termination conditions are satisfied.",
}

JAC_ANALYTIC = 0
JAC_TWO_POINT = 1
JAC_MINPACK_FORWARD = 2  # fdjac2 of MINPACK's lmdif (method "lm" without an analytic Jacobian)
# PNB_METHOD_*: scipy.optimize.least_squares(method=...) for "trf" / "dogbox", MINPACK through leastsq for "lm"
METHODS = {"trf": 0, "dogbox": 1, "lm": 2}
LOSSES = {"linear": 0, "soft_l1": 1, "huber": 2, "cauchy": 3, "arctan": 4}  # least_squares(loss=...), PNB_LOSS_*
ST_LM_TOO_FEW_DATA = -9  # not a kernel status: method='lm' with more parameters than measurements
# leastsq's xtol / gtol defaults (curve_fit(method="lm") does not override them)
LM_XTOL, LM_GTOL = 1.49012e-8, 0.0


def status_message(status: int, method: str = "trf", *, max_nfev: int = 0, ftol: float = 0.0, xtol: float = 0.0,
                   gtol: float = 0.0, n_params: int = 0, n_data: int = 0):
    """The text SciPy's exception carries for a failed voxel (``None`` for a success), as the
    reference stores it in ``_PixelFitResult.message`` (solvers/curvefit.py:308-317)."""
    status = int(status)
    if status > 0:
        return None
    if method == "lm" and status in (0, -6, -7, -8, ST_LM_TOO_FEW_DATA):
        # scipy/optimize/_minpack_py.py: leastsq's `errors` table, raised by curve_fit as RuntimeError
        if status == ST_LM_TOO_FEW_DATA:
            return f"The number of func parameters={n_params} must not exceed the number of data points={n_data}"
        text = {
            0: f"Number of calls to function has reached maxfev = {max_nfev}.",
            -6: f"ftol={ftol:f} is too small, no further reduction in the sum of squares\n  is possible.",
            -7: f"xtol={xtol:f} is too small, no further improvement in the approximate\n  solution is possible.",
            -8: f"gtol={gtol:f} is too small, func(x) is orthogonal to the columns of\n  the Jacobian to machine precision.",
        }[status]
        return "Optimal parameters not found: " + text
    return STATUS_MESSAGES[status]


def _is_torch_cuda(x) -> bool:
    try:
        import torch
    except ImportError:  # pragma: no cover
        return False
    return isinstance(x, torch.Tensor) and x.is_cuda


def _as_f64(a) -> np.ndarray:
    return np.ascontiguousarray(a, dtype=np.float64)


def frozen_mask(desc: ModelDesc, fixed_names) -> int:
    mask = 0
    for j, name in enumerate(desc.all_names):
        if name in fixed_names:
            mask |= 1 << j
    return mask


def trf_fit(
    desc: ModelDesc,
    xdata,
    ydata,
    p0,
    lb,
    ub,
    frozen: int = 0,
    *,
    max_nfev: int = 250,
    ftol: float = 1e-8,
    xtol: float = 1e-8,
    gtol: float = 1e-8,
    jac_mode: int = JAC_ANALYTIC,
    x_scale=None,
    x_scale_jac: bool = False,
    want_cov=True,
    device=0,
    chunk_vox: int = 0,
    out: dict | None = None,
    method: str = "trf",
    finish_wait: int = 0,
    weights=None,
    loss: str = "linear",
    f_scale: float = 1.0,
    diff_step=None,
    absolute_sigma: bool = False,
):
    """Fit all voxels.  ``p0``/``lb``/``ub``: ``(n_all,)`` or ``(n_all, n_vox)`` over
    ``desc.all_names`` (frozen rows of ``p0`` carry the fixed values).

    ``device``: a CUDA ordinal, a list of ordinals or ``"all"`` — with several GPUs the host path
    shards the voxels into contiguous ranges, one pipeline per GPU in one call
    (``pnb_trf_fit_host_multi``).  ``want_cov``: ``True`` / ``"lazy"`` computes the covariances and
    leaves them on the GPU(s) behind a :class:`~pyneapple_b200._lazy.LazyArray` (host path), ``"eager"``
    copies them to the host with the other results, ``False`` skips them.

    Returns ``dict(params (n_all, n_vox), cov (n_vox, n_free, n_free) | None,
    status, nfev, njev, cost)``.
    """
    _lib.require_device()
    lib = _lib.load()
    n_all = desc.n_all
    n_free = n_all - bin(frozen).count("1")
    prob = _lib.TrfProblem()
    prob.model_id = desc.model_id
    prob.t1_mode = desc.t1_mode
    prob.repetition_time = desc.repetition_time
    prob.mixing_time = desc.mixing_time
    prob.n_params = n_all
    prob.frozen_mask = frozen
    prob.max_nfev = int(max_nfev)
    prob.ftol, prob.xtol, prob.gtol = float(ftol), float(xtol), float(gtol)
    prob.jac_mode = int(jac_mode)
    prob.x_scale_jac = int(bool(x_scale_jac))
    if method not in METHODS:
        raise NotImplementedError(f"method={method!r}: SciPy's 'trf', 'dogbox' and 'lm' have a B200 implementation")
    prob.method = METHODS[method]
    prob.finish_wait = int(finish_wait)  # kernel scheduling hint (pnb_trf_problem.finish_wait), 0 = default
    # curve_fit extras (`weights` = 1 / sigma per b-value; `diff_step` over desc.all_names)
    if loss not in LOSSES:
        raise ValueError(f"`loss` must be one of {sorted(LOSSES)} (callables have no B200 implementation).")
    if loss != "linear" and method == "lm":
        raise ValueError("method='lm' supports only 'linear' loss function.")
    if not float(f_scale) > 0.0:
        raise ValueError("`f_scale` must be positive.")
    prob.loss, prob.f_scale, prob.absolute_sigma = LOSSES[loss], float(f_scale), int(bool(absolute_sigma))
    ds = np.zeros(8)
    if diff_step is not None:
        ds[:n_all] = np.broadcast_to(np.asarray(diff_step, float), (n_all,))
        if not np.all(np.isfinite(ds)) or np.any(ds < 0):
            raise ValueError("`diff_step` must be non-negative and finite.")
    for i in range(8):
        prob.diff_step[i] = ds[i]
    w_host = None
    if weights is not None:
        w_host = np.ascontiguousarray(np.asarray(weights, np.float64))
        if w_host.ndim != 1 or not np.all(np.isfinite(w_host)):
            raise ValueError("weights (1 / sigma) must be a finite vector over the b-values")
    extras = w_host is not None or loss != "linear" or bool(np.any(ds > 0))
    if extras and (method != "trf" or desc.t1_mode != 0):
        raise NotImplementedError("sigma / loss / diff_step are implemented for method='trf' on models without a T1 parameter")
    xs = np.ones(8)
    if x_scale is not None:
        xs[:n_all] = np.broadcast_to(np.asarray(x_scale, float), (n_all,))
    for i in range(8):
        prob.x_scale[i] = xs[i]

    if _is_torch_cuda(ydata):
        if w_host is not None:
            if w_host.shape != (int(ydata.shape[1]),):
                raise ValueError("`sigma` has incorrect shape.")
            w_dev = _small_const(w_host, ydata.device)
            prob.weights = w_dev.data_ptr()
        return _trf_fit_device(lib, prob, desc, xdata, ydata, p0, lb, ub, n_free, want_cov)

    y = _as_f64(ydata)
    if y.ndim != 2:
        raise ValueError(f"ydata must be (n_vox, n_b), got {y.shape}")
    n_vox, n_b = y.shape
    b = _as_f64(xdata)
    if b.shape != (n_b,):
        raise ValueError(f"xdata length {b.shape} does not match ydata {y.shape}")
    p0 = _as_f64(p0)
    lb = _as_f64(lb)
    ub = _as_f64(ub)
    for name, arr in (("p0", p0), ("lb", lb), ("ub", ub)):
        if arr.shape not in ((n_all,), (n_all, n_vox)):
            raise ValueError(f"{name} must have shape ({n_all},) or ({n_all}, {n_vox}), got {arr.shape}")
    if lb.shape != ub.shape:
        raise ValueError("lb and ub must have the same shape")
    prob.n_b, prob.n_vox = n_b, n_vox
    prob.p0_per_voxel = int(p0.ndim == 2)
    prob.bounds_per_voxel = int(lb.ndim == 2)

    devices = _lib.resolve_devices(device)
    if len(devices) > n_vox:
        devices = devices[: max(1, n_vox)]
    o = out or {}
    params = o.get("params")
    if params is None:
        params = np.empty((n_all, n_vox))
    lazy = want_cov is True or want_cov == "lazy"
    cov = cov_parts = None
    if want_cov and not lazy:
        cov = o.get("cov")
        if cov is None:
            cov = np.empty((n_vox, n_free, n_free))
    elif lazy and n_vox:
        import torch

        cov_parts = [(a, z, torch.empty((z - a, n_free, n_free), dtype=torch.float64, device=f"cuda:{d}"))
                     for (a, z), d in zip(_lib.shard_ranges(n_vox, len(devices)), devices)]
        for d in set(devices):
            # the blocks may be recycled memory with work still queued on torch's stream; the library
            # writes them from its own streams
            torch.cuda.current_stream(d).synchronize()
    status = o.get("status") if o.get("status") is not None else np.empty(n_vox, np.int32)
    nfev = o.get("nfev") if o.get("nfev") is not None else np.empty(n_vox, np.int32)
    njev = o.get("njev") if o.get("njev") is not None else np.empty(n_vox, np.int32)
    cost = o.get("cost") if o.get("cost") is not None else np.empty(n_vox)
    r2 = o.get("r2") if o.get("r2") is not None else np.empty(n_vox)
    if w_host is not None:
        if w_host.shape != (n_b,):
            raise ValueError("`sigma` has incorrect shape.")
        prob.weights = w_host.ctypes.data
    keep = (b, y, p0, lb, ub, params, cov, status, nfev, njev, cost, r2, w_host)
    prob.xdata, prob.ydata = b.ctypes.data, y.ctypes.data
    prob.p0, prob.lb, prob.ub = p0.ctypes.data, lb.ctypes.data, ub.ctypes.data
    prob.params = params.ctypes.data
    prob.cov = cov.ctypes.data if cov is not None else None
    prob.status, prob.nfev = status.ctypes.data, nfev.ctypes.data
    prob.njev, prob.cost = njev.ctypes.data, cost.ctypes.data
    prob.r_squared = r2.ctypes.data
    if len(devices) == 1:
        if cov_parts:
            prob.cov = cov_parts[0][2].data_ptr()
        _lib.check(lib.pnb_trf_fit_host(C.byref(prob), devices[0], int(chunk_vox)), "pnb_trf_fit_host")
    else:
        dev_arr = (C.c_int32 * len(devices))(*devices)
        cov_ptrs = None
        if cov_parts:
            cov_ptrs = (C.c_void_p * len(devices))(*[t.data_ptr() for _, _, t in cov_parts])
        _lib.check(lib.pnb_trf_fit_host_multi(C.byref(prob), dev_arr, len(devices), int(chunk_vox), cov_ptrs),
                   "pnb_trf_fit_host_multi")
    del keep
    if cov_parts:
        from ._lazy import LazyArray

        cov = LazyArray((n_vox, n_free, n_free), cov_parts)
    return dict(params=params, cov=cov, status=status, nfev=nfev, njev=njev, cost=cost, r2=r2,
                n_failed=int(lib.pnb_trf_last_failed_count()))


_CONSTS: dict = {}


def _small_const(a, dev):
    """Device copy of a small host vector (b-values, broadcast p0 / bounds, the NNLS dictionary),
    cached by content.  Uploading it on every call is a pageable H2D copy, i.e. a synchronisation of the
    stream: back-to-back device-path launches then wait for each other on the host, and every host
    hiccup lands in the step time (measured: 13.4 -> 18.9 ms per C2 step while nvidia-smi was polling)."""
    import torch

    arr = np.ascontiguousarray(a, np.float64)
    if arr.nbytes > (1 << 16):
        return torch.as_tensor(arr).to(dev)
    key = (str(dev), arr.shape, arr.tobytes())
    t = _CONSTS.get(key)
    if t is None:
        if len(_CONSTS) > 256:
            _CONSTS.clear()
        t = _CONSTS[key] = torch.as_tensor(arr).to(dev)
    return t


def _trf_fit_device(lib, prob, desc, xdata, ydata, p0, lb, ub, n_free, want_cov):
    import torch

    dev = ydata.device
    y = ydata.contiguous().to(torch.float64)
    if y.ndim != 2:
        raise ValueError(f"ydata must be (n_vox, n_b), got {tuple(y.shape)}")
    n_vox, n_b = y.shape
    n_all = desc.n_all

    def dev_f64(a):
        if isinstance(a, torch.Tensor):
            return a.to(device=dev, dtype=torch.float64).contiguous()
        return _small_const(a, dev)

    b = dev_f64(xdata)
    p0, lb, ub = dev_f64(p0), dev_f64(lb), dev_f64(ub)
    for name, arr in (("p0", p0), ("lb", lb), ("ub", ub)):
        if tuple(arr.shape) not in ((n_all,), (n_all, n_vox)):
            raise ValueError(f"{name} must have shape ({n_all},) or ({n_all}, {n_vox}), got {tuple(arr.shape)}")
    prob.n_b, prob.n_vox = n_b, n_vox
    prob.p0_per_voxel = int(p0.ndim == 2)
    prob.bounds_per_voxel = int(lb.ndim == 2)
    params = torch.empty((n_all, n_vox), dtype=torch.float64, device=dev)
    cov = torch.empty((n_vox, n_free, n_free), dtype=torch.float64, device=dev) if want_cov else None
    status = torch.empty(n_vox, dtype=torch.int32, device=dev)
    nfev = torch.empty(n_vox, dtype=torch.int32, device=dev)
    njev = torch.empty(n_vox, dtype=torch.int32, device=dev)
    cost = torch.empty(n_vox, dtype=torch.float64, device=dev)
    r2 = torch.empty(n_vox, dtype=torch.float64, device=dev)
    prob.xdata, prob.ydata = b.data_ptr(), y.data_ptr()
    prob.p0, prob.lb, prob.ub = p0.data_ptr(), lb.data_ptr(), ub.data_ptr()
    prob.params = params.data_ptr()
    prob.cov = cov.data_ptr() if cov is not None else None
    prob.status, prob.nfev = status.data_ptr(), nfev.data_ptr()
    prob.njev, prob.cost = njev.data_ptr(), cost.data_ptr()
    prob.r_squared = r2.data_ptr()
    with torch.cuda.device(dev):
        stream = torch.cuda.current_stream(dev).cuda_stream
        _lib.check(lib.pnb_trf_fit_device(C.byref(prob), C.c_void_p(stream)), "pnb_trf_fit_device")
        # inputs must outlive the enqueued kernel
        for t in (b, y, p0, lb, ub):
            t.record_stream(torch.cuda.current_stream(dev))
    return dict(params=params, cov=cov, status=status, nfev=nfev, njev=njev, cost=cost, r2=r2)


# ---------------------------------------------------------------------------
# NNLS
# ---------------------------------------------------------------------------


def rtr_band(reg_matrix: np.ndarray):
    """Band storage of ``R^T R`` for a (banded) regularisation matrix ``R`` (n x n).

    Returns ``(band (n, 2W+1), W)`` with ``band[j, d + W] = (R^T R)[j, j + d]``.
    """
    R = np.asarray(reg_matrix, dtype=np.float64)
    m_r, n = R.shape
    nzr = np.nonzero(R)
    if nzr[0].size == 0:
        return np.zeros((n, 1)), 0
    lo, hi = int((nzr[0] - nzr[1]).min()), int((nzr[0] - nzr[1]).max())  # rows k - column j of the stencil
    W = hi - lo
    if W > 8:
        raise NotImplementedError(f"regularisation matrix with R^T R half-bandwidth {W} > 8")
    # (R^T R)[j, j+d] = sum_k R[k, j] R[k, j+d], added over the stencil rows k = j + r in the same
    # order r = lo .. hi for every column.  A BLAS product would add the same terms in an order that
    # depends on where j falls in its blocking, so the interior rows of the band could differ in the
    # last bit — the kernel keeps them in registers only when they are identical.
    band = np.zeros((n, 2 * W + 1))
    cols = np.arange(n)
    for d in range(-W, W + 1):
        j = cols[max(0, -d): min(n, n - d)]
        acc = np.zeros(j.shape[0])
        for r in range(lo, hi + 1):
            k = j + r
            ok = (k >= 0) & (k < m_r)
            term = np.zeros(j.shape[0])
            term[ok] = R[k[ok], j[ok]] * R[k[ok], j[ok] + d]
            acc = acc + term
        band[j, d + W] = acc
    # trim offsets that are zero everywhere (e.g. a diagonal R)
    while W > 0 and not band[:, 0].any() and not band[:, -1].any():
        band = band[:, 1:-1]
        W -= 1
    band = np.ascontiguousarray(band)
    return band, W


def nnls_fit(basis, reg_matrix, signal, max_iter: int, *, device=0, chunk_vox: int = 0,
             out: dict | None = None, algorithm: str = "auto", dual_init: str = "fused"):
    """Batched ``scipy.optimize.nnls([basis; reg_matrix], [signal; 0], maxiter=max_iter)``.

    ``signal``: numpy ``(n_vox, n_b)`` (host path) or CUDA tensor (device path).
    Returns ``dict(coefficients (n_vox, n_bins), residual, status, iterations)``.
    """
    _lib.require_device()
    lib = _lib.load()
    B = _as_f64(basis)
    n_b, n_bins = B.shape
    band, W = rtr_band(reg_matrix)
    prob = _lib.NnlsProblem()
    prob.n_b, prob.n_bins, prob.rtr_halfband, prob.max_iter = n_b, n_bins, W, int(max_iter)
    if algorithm not in ("auto", "robust"):
        raise ValueError("algorithm must be 'auto' or 'robust'")
    # an un- or barely regularised dictionary is too ill-conditioned for the inverse-update fast
    # path (the minimiser is then numerically non-unique and only the Cholesky path follows
    # SciPy's choice): go straight to the robust kernel below a relative weight of 1e-8
    reg_weight = float(np.abs(band[:, W]).max()) / float((B * B).sum(axis=0).max())
    prob.algorithm = 1 if (algorithm == "robust" or reg_weight < 1e-8) else 0
    if dual_init not in ("fused", "gemm"):
        raise ValueError("dual_init must be 'fused' or 'gemm'")
    prob.dual_init = int(dual_init == "gemm")
    if _is_torch_cuda(signal):
        import torch

        dev = signal.device
        y = signal.contiguous().to(torch.float64)
        if y.ndim != 2 or y.shape[1] != n_b:
            raise ValueError(f"signal must be (n_vox, {n_b}), got {tuple(y.shape)}")
        n_vox = y.shape[0]
        Bd = _small_const(B, dev)
        bd = _small_const(band, dev)
        coef = torch.empty((n_vox, n_bins), dtype=torch.float64, device=dev)
        res = torch.empty(n_vox, dtype=torch.float64, device=dev)
        status = torch.empty(n_vox, dtype=torch.int32, device=dev)
        iters = torch.empty(n_vox, dtype=torch.int32, device=dev)
        r2 = torch.empty(n_vox, dtype=torch.float64, device=dev)
        prob.n_vox = n_vox
        prob.basis, prob.rtr_band, prob.signal = Bd.data_ptr(), bd.data_ptr(), y.data_ptr()
        prob.coefficients, prob.residual = coef.data_ptr(), res.data_ptr()
        prob.status, prob.iterations = status.data_ptr(), iters.data_ptr()
        prob.r_squared = r2.data_ptr()
        with torch.cuda.device(dev):
            stream = torch.cuda.current_stream(dev)
            _lib.check(lib.pnb_nnls_fit_device(C.byref(prob), C.c_void_p(stream.cuda_stream)),
                       "pnb_nnls_fit_device")
            for t in (Bd, bd, y):
                t.record_stream(stream)
        return dict(coefficients=coef, residual=res, status=status, iterations=iters, r2=r2)
    y = _as_f64(signal)
    if y.ndim != 2 or y.shape[1] != n_b:
        raise ValueError(f"signal must be (n_vox, {n_b}), got {y.shape}")
    n_vox = y.shape[0]
    o = out or {}
    coef = o.get("coefficients") if o.get("coefficients") is not None else np.empty((n_vox, n_bins))
    res = o.get("residual") if o.get("residual") is not None else np.empty(n_vox)
    status = o.get("status") if o.get("status") is not None else np.empty(n_vox, np.int32)
    iters = o.get("iterations") if o.get("iterations") is not None else np.empty(n_vox, np.int32)
    r2 = o.get("r2") if o.get("r2") is not None else np.empty(n_vox)
    prob.n_vox = n_vox
    prob.basis, prob.rtr_band, prob.signal = B.ctypes.data, band.ctypes.data, y.ctypes.data
    prob.coefficients, prob.residual = coef.ctypes.data, res.ctypes.data
    prob.status, prob.iterations = status.ctypes.data, iters.ctypes.data
    prob.r_squared = r2.ctypes.data
    devices = _lib.resolve_devices(device)
    if len(devices) == 1:
        _lib.check(lib.pnb_nnls_fit_host(C.byref(prob), devices[0], int(chunk_vox)), "pnb_nnls_fit_host")
    else:
        dev_arr = (C.c_int32 * len(devices))(*devices)
        _lib.check(lib.pnb_nnls_fit_host_multi(C.byref(prob), dev_arr, len(devices), int(chunk_vox)),
                   "pnb_nnls_fit_host_multi")
    return dict(coefficients=coef, residual=res, status=status, iterations=iters, r2=r2)


def nnls_dual_gemm(basis, signal):
    """``signal (n_vox, n_b) @ basis (n_b, n_bins)`` on the FP64 tensor cores (``pnb_nnls_dual_gemm_device``):
    ``h = B^T y`` of every voxel, the first dual of Lawson-Hanson.  CUDA tensors in, CUDA tensor out."""
    import torch

    _lib.require_device()
    dev = signal.device
    y = signal.to(torch.float64).contiguous()
    Bd = torch.as_tensor(_as_f64(basis)).to(dev) if not _is_torch_cuda(basis) else basis.to(torch.float64).contiguous()
    out = torch.empty((y.shape[0], Bd.shape[1]), dtype=torch.float64, device=dev)
    with torch.cuda.device(dev):
        stream = torch.cuda.current_stream(dev)
        _lib.check(_lib.load().pnb_nnls_dual_gemm_device(int(Bd.shape[0]), int(Bd.shape[1]), int(y.shape[0]),
                                                          Bd.data_ptr(), y.data_ptr(), out.data_ptr(),
                                                          C.c_void_p(stream.cuda_stream)), "pnb_nnls_dual_gemm_device")
        y.record_stream(stream)
        Bd.record_stream(stream)
    return out


def segment_means(image, segmentation, *, device: int = 0):
    """Mean signal of every label of ``segmentation`` (``fitters/segmentationwise.py:112-137``).

    ``image``: ``(..., n_b)`` numpy array or CUDA tensor, ``segmentation``: its spatial shape.
    Returns ``(labels, means (n_labels, n_b), counts, inverse)``: ``labels = np.unique(segmentation)``
    (the background label included, like the reference), ``inverse`` the dense label index of
    every voxel (C order).  One pass over the volume on the GPU (``pnb_segment_means_*``).
    """
    _lib.require_device()
    lib = _lib.load()
    n_b = int(image.shape[-1])
    prob = _lib.SegmeansProblem()
    prob.n_b = n_b
    if _is_torch_cuda(image):
        import torch

        dev = image.device
        seg = torch.as_tensor(segmentation, device=dev)
        labels_t, inverse = torch.unique(seg, return_inverse=True)
        img = image.reshape(-1, n_b).contiguous().to(torch.float64)
        lab = inverse.reshape(-1).to(torch.int32).contiguous()
        L = int(labels_t.numel())
        means = torch.empty((L, n_b), dtype=torch.float64, device=dev)
        counts = torch.empty((L,), dtype=torch.int64, device=dev)
        prob.n_labels, prob.n_vox = L, int(lab.numel())
        prob.image, prob.label = img.data_ptr(), lab.data_ptr()
        prob.means, prob.counts = means.data_ptr(), counts.data_ptr()
        with torch.cuda.device(dev):
            stream = torch.cuda.current_stream(dev).cuda_stream
            _lib.check(lib.pnb_segment_means_device(C.byref(prob), C.c_void_p(stream)), "pnb_segment_means_device")
        return labels_t.cpu().numpy(), means, counts, inverse.reshape(-1)
    segmentation = np.asarray(segmentation)
    labels, inverse = np.unique(segmentation, return_inverse=True)
    img = np.ascontiguousarray(np.asarray(image, dtype=np.float64).reshape(-1, n_b))
    lab = np.ascontiguousarray(inverse.reshape(-1), dtype=np.int32)
    if lab.shape[0] != img.shape[0]:
        raise ValueError(f"segmentation has {lab.shape[0]} voxels, the image {img.shape[0]}")
    L = int(labels.shape[0])
    means = np.empty((L, n_b))
    counts = np.empty(L, np.int64)
    prob.n_labels, prob.n_vox = L, int(lab.shape[0])
    prob.image, prob.label = img.ctypes.data, lab.ctypes.data
    prob.means, prob.counts = means.ctypes.data, counts.ctypes.data
    _lib.check(lib.pnb_segment_means_host(C.byref(prob), int(device)), "pnb_segment_means_host")
    return labels, means, counts, inverse.reshape(-1)


def predict_device(desc: ModelDesc, xdata, params, flat_index=None, n_out: int | None = None):
    """Model signal of every voxel on the GPU (``pnb_predict_device``).

    ``params``: CUDA tensor ``(n_all, n_vox)`` over ``desc.all_names``.  With ``flat_index`` (CUDA int64
    tensor, the C-order position of each voxel in a volume of ``n_out`` voxels) the rows are scattered
    into an ``(n_out, n_b)`` tensor that is zero elsewhere; otherwise the result is ``(n_vox, n_b)``.
    """
    import torch

    _lib.require_device()
    lib = _lib.load()
    dev = params.device
    par = params.to(torch.float64).contiguous()
    b = torch.as_tensor(np.ascontiguousarray(xdata, np.float64)).to(dev)
    n_vox = int(par.shape[1])
    rows = n_vox if flat_index is None else int(n_out)
    out = torch.empty((rows, b.shape[0]), dtype=torch.float64, device=dev)
    prob = _lib.PredictProblem()
    prob.model_id, prob.t1_mode = desc.model_id, desc.t1_mode
    prob.repetition_time, prob.mixing_time = desc.repetition_time, desc.mixing_time
    prob.n_b, prob.n_params, prob.n_vox, prob.n_out = int(b.shape[0]), int(par.shape[0]), n_vox, rows
    idx = None
    if flat_index is not None:
        idx = flat_index.to(device=dev, dtype=torch.int64).contiguous()
    prob.xdata, prob.params, prob.signal = b.data_ptr(), par.data_ptr(), out.data_ptr()
    prob.flat_index = idx.data_ptr() if idx is not None else None
    with torch.cuda.device(dev):
        stream = torch.cuda.current_stream(dev)
        _lib.check(lib.pnb_predict_device(C.byref(prob), C.c_void_p(stream.cuda_stream)), "pnb_predict_device")
        for t in (b, par) + ((idx,) if idx is not None else ()):
            t.record_stream(stream)
    return out


def move_rows(src, index, n_other: int, scatter: bool, out_float32: bool = False, zero_fill: bool = True):
    """Row gather (``dst[i] = src[index[i]]``) or scatter (``dst[index[i]] = src[i]``) on the GPU
    (``pnb_move_rows_device``).  ``src``: CUDA float64 tensor ``(rows, ...)``; ``index``: CUDA int64
    ``(n_rows,)``; ``n_other``: rows of the array on the indexed side.  Returns the new tensor."""
    import torch

    _lib.require_device()
    lib = _lib.load()
    dev = src.device
    s = src.to(torch.float64).contiguous()
    idx = index.to(device=dev, dtype=torch.int64).contiguous()
    n_rows = int(idx.shape[0])
    tail = tuple(s.shape[1:])
    width = int(np.prod(tail)) if tail else 1
    dt = torch.float32 if out_float32 else torch.float64
    out = torch.empty(((int(n_other) if scatter else n_rows),) + tail, dtype=dt, device=dev)
    prob = _lib.RowsProblem()
    prob.direction, prob.out_dtype, prob.width = int(scatter), int(out_float32), width
    prob.zero_fill = int(zero_fill)
    prob.n_rows, prob.n_other = n_rows, int(n_other)
    prob.src, prob.dst, prob.index = s.data_ptr(), out.data_ptr(), idx.data_ptr()
    with torch.cuda.device(dev):
        stream = torch.cuda.current_stream(dev)
        _lib.check(lib.pnb_move_rows_device(C.byref(prob), C.c_void_p(stream.cuda_stream)), "pnb_move_rows_device")
        s.record_stream(stream)
        idx.record_stream(stream)
    return out


_TORCH_NP = None


def to_host(tensor) -> np.ndarray:
    """CUDA tensor -> fresh numpy array through ``pnb_download`` (staged, multi-threaded: about five
    times ``tensor.cpu()`` for the pageable memory numpy allocates)."""
    import torch

    global _TORCH_NP
    if _TORCH_NP is None:
        _TORCH_NP = {torch.float64: np.float64, torch.float32: np.float32, torch.int32: np.int32,
                     torch.int64: np.int64, torch.bool: np.bool_, torch.uint8: np.uint8}
    if not _is_torch_cuda(tensor):
        return np.asarray(tensor)
    if tensor.dtype not in _TORCH_NP or tensor.numel() * tensor.element_size() < (1 << 20):
        return tensor.cpu().numpy()
    t = tensor.contiguous()
    out = np.empty(tuple(t.shape), dtype=_TORCH_NP[t.dtype])
    with torch.cuda.device(t.device):
        stream = torch.cuda.current_stream(t.device).cuda_stream
        _lib.check(_lib.load().pnb_download(out.ctypes.data, t.data_ptr(), out.nbytes, C.c_void_p(stream)), "pnb_download")
    return out


def to_host_into(tensor, out: np.ndarray) -> None:
    """CUDA tensor -> an existing C-contiguous numpy array (or contiguous slice of one) of the same shape and
    item size, through ``pnb_download`` on the current stream of the tensor's device."""
    import torch

    t = tensor.contiguous()
    if not out.flags["C_CONTIGUOUS"] or tuple(out.shape) != tuple(t.shape) or out.itemsize != t.element_size():
        raise ValueError("to_host_into needs a C-contiguous destination of the tensor's shape and item size")
    if out.nbytes == 0:
        return
    with torch.cuda.device(t.device):
        stream = torch.cuda.current_stream(t.device).cuda_stream
        _lib.check(_lib.load().pnb_download(out.ctypes.data, t.data_ptr(), out.nbytes, C.c_void_p(stream)), "pnb_download")


def to_device(array, device):
    """numpy array -> CUDA tensor through ``pnb_upload`` (staged like :func:`to_host`)."""
    import torch

    a = np.ascontiguousarray(array)
    if a.nbytes < (1 << 20) or a.dtype.type not in (np.float64, np.float32, np.int32, np.int64, np.uint8, np.bool_):
        return torch.as_tensor(a).to(device)
    dev = torch.device(device)
    t = torch.empty(a.shape, dtype=torch.from_numpy(a[:0].reshape(-1)).dtype, device=dev)
    with torch.cuda.device(dev):
        stream = torch.cuda.current_stream(dev).cuda_stream
        _lib.check(_lib.load().pnb_upload(t.data_ptr(), a.ctypes.data, a.nbytes, C.c_void_p(stream)), "pnb_upload")
    return t
