"""OpenCV-faithful in-plane resampling on the GPU (``pnb_resize2d_*``).

Device replacement for ``IDEALFitter._interpolate_array`` (fitters/ideal.py:299-320):
the leading two axes are resampled, every trailing (slice, channel) plane
independently; non-float input is cast to float32 first, like the reference.
"""

from __future__ import annotations

import ctypes as C

import numpy as np

from . import _lib

METHODS = {"linear": 0, "cubic": 1}


def interpolate_array(array, target_shape, method: str = "cubic", device: int = 0):
    """Resample ``array (X, Y, ...)`` to ``(target_shape[0], target_shape[1], ...)``.

    numpy in -> numpy out (host entry point); CUDA tensor in -> CUDA tensor out.
    """
    if method not in METHODS:
        raise ValueError(f"Invalid interpolation method: {method}. Must be one of {sorted(METHODS)}.")
    _lib.require_device()
    lib = _lib.load()
    th, tw = int(target_shape[0]), int(target_shape[1])
    prob = _lib.ResizeProblem()
    prob.method = METHODS[method]
    prob.dst_h, prob.dst_w = th, tw
    try:
        import torch
    except ImportError:  # pragma: no cover
        torch = None
    if torch is not None and isinstance(array, torch.Tensor) and array.is_cuda:
        a = array
        if not a.dtype.is_floating_point:
            a = a.to(torch.float32)
        if a.dtype not in (torch.float64, torch.float32):
            a = a.to(torch.float32)
        a = a.contiguous()
        out = torch.empty((th, tw) + tuple(a.shape[2:]), dtype=a.dtype, device=a.device)
        prob.dtype = 0 if a.dtype == torch.float64 else 1
        prob.src_h, prob.src_w = a.shape[0], a.shape[1]
        prob.inner = int(np.prod(a.shape[2:])) if a.ndim > 2 else 1
        prob.src, prob.dst = a.data_ptr(), out.data_ptr()
        with torch.cuda.device(a.device):
            stream = torch.cuda.current_stream(a.device)
            _lib.check(lib.pnb_resize2d_device(C.byref(prob), C.c_void_p(stream.cuda_stream)),
                       "pnb_resize2d_device")
            a.record_stream(stream)
        return out
    a = np.asarray(array)
    if a.dtype.kind != "f":
        a = a.astype(np.float32)
    if a.dtype not in (np.float64, np.float32):
        a = a.astype(np.float32)
    a = np.ascontiguousarray(a)
    out = np.empty((th, tw) + a.shape[2:], dtype=a.dtype)
    prob.dtype = 0 if a.dtype == np.float64 else 1
    prob.src_h, prob.src_w = a.shape[0], a.shape[1]
    prob.inner = int(np.prod(a.shape[2:])) if a.ndim > 2 else 1
    prob.src, prob.dst = a.ctypes.data, out.ctypes.data
    _lib.check(lib.pnb_resize2d_host(C.byref(prob), int(device)), "pnb_resize2d_host")
    return out
