"""ctypes binding of ``csrc/libpnb200.so`` (the C ABI in ``include/pyneapple_b200.h``).

The library is built in-tree by ``__graft_entry__.build()`` / ``make -C
pyneapple_b200/csrc``.  There is no fallback: if the shared object or a CUDA
device is missing every compute entry point raises.
"""

from __future__ import annotations

import ctypes as C
import os
import subprocess

import numpy as np

_HERE = os.path.dirname(os.path.abspath(__file__))
CSRC = os.path.join(_HERE, "csrc")
SO_PATH = os.environ.get("PNB_LIB") or os.path.join(CSRC, "libpnb200.so")
HEADER = os.path.join(os.path.dirname(_HERE), "include", "pyneapple_b200.h")

_lib = None


class EngineError(RuntimeError):
    """The CUDA engine could not run (library / device missing, CUDA error)."""


class TrfProblem(C.Structure):
    """Mirror of ``struct pnb_trf_problem``."""

    _fields_ = [
        ("model_id", C.c_int32),
        ("t1_mode", C.c_int32),
        ("repetition_time", C.c_double),
        ("mixing_time", C.c_double),
        ("n_b", C.c_int32),
        ("n_params", C.c_int32),
        ("n_vox", C.c_int64),
        ("xdata", C.c_void_p),
        ("ydata", C.c_void_p),
        ("p0", C.c_void_p),
        ("lb", C.c_void_p),
        ("ub", C.c_void_p),
        ("p0_per_voxel", C.c_int32),
        ("bounds_per_voxel", C.c_int32),
        ("frozen_mask", C.c_uint32),
        ("max_nfev", C.c_int32),
        ("ftol", C.c_double),
        ("xtol", C.c_double),
        ("gtol", C.c_double),
        ("jac_mode", C.c_int32),
        ("x_scale_jac", C.c_int32),
        ("method", C.c_int32),
        ("finish_wait", C.c_int32),
        ("x_scale", C.c_double * 8),
        ("weights", C.c_void_p),
        ("diff_step", C.c_double * 8),
        ("loss", C.c_int32),
        ("absolute_sigma", C.c_int32),
        ("f_scale", C.c_double),
        ("params", C.c_void_p),
        ("cov", C.c_void_p),
        ("status", C.c_void_p),
        ("nfev", C.c_void_p),
        ("njev", C.c_void_p),
        ("cost", C.c_void_p),
        ("r_squared", C.c_void_p),
    ]


class NnlsProblem(C.Structure):
    """Mirror of ``struct pnb_nnls_problem``."""

    _fields_ = [
        ("n_b", C.c_int32),
        ("n_bins", C.c_int32),
        ("rtr_halfband", C.c_int32),
        ("max_iter", C.c_int32),
        ("algorithm", C.c_int32),
        ("dual_init", C.c_int32),
        ("n_vox", C.c_int64),
        ("basis", C.c_void_p),
        ("rtr_band", C.c_void_p),
        ("signal", C.c_void_p),
        ("coefficients", C.c_void_p),
        ("residual", C.c_void_p),
        ("status", C.c_void_p),
        ("iterations", C.c_void_p),
        ("r_squared", C.c_void_p),
    ]


class SpectrumProblem(C.Structure):
    """Mirror of ``struct pnb_spectrum_problem``."""

    _fields_ = [
        ("n_bins", C.c_int32),
        ("max_peaks", C.c_int32),
        ("detect", C.c_int32),
        ("areas", C.c_int32),
        ("normalize", C.c_int32),
        ("n_cutoffs", C.c_int32),
        ("cut_normalize", C.c_int32),
        ("reserved", C.c_int32),
        ("height", C.c_double),
        ("rel_height", C.c_double),
        ("n_vox", C.c_int64),
        ("bins", C.c_void_p),
        ("cutoffs", C.c_void_p),
        ("spectrum", C.c_void_p),
        ("n_peaks", C.c_void_p),
        ("peak_index", C.c_void_p),
        ("d_values", C.c_void_p),
        ("f_values", C.c_void_p),
        ("d_cut", C.c_void_p),
        ("f_cut", C.c_void_p),
    ]


class SegmeansProblem(C.Structure):
    """Mirror of ``struct pnb_segmeans_problem``."""

    _fields_ = [
        ("n_b", C.c_int32),
        ("n_labels", C.c_int32),
        ("n_vox", C.c_int64),
        ("image", C.c_void_p),
        ("label", C.c_void_p),
        ("means", C.c_void_p),
        ("counts", C.c_void_p),
    ]


class PredictProblem(C.Structure):
    """Mirror of ``struct pnb_predict_problem``."""

    _fields_ = [
        ("model_id", C.c_int32),
        ("t1_mode", C.c_int32),
        ("repetition_time", C.c_double),
        ("mixing_time", C.c_double),
        ("n_b", C.c_int32),
        ("n_params", C.c_int32),
        ("n_vox", C.c_int64),
        ("n_out", C.c_int64),
        ("xdata", C.c_void_p),
        ("params", C.c_void_p),
        ("flat_index", C.c_void_p),
        ("signal", C.c_void_p),
    ]


class RowsProblem(C.Structure):
    """Mirror of ``struct pnb_rows_problem``."""

    _fields_ = [
        ("direction", C.c_int32),
        ("out_dtype", C.c_int32),
        ("width", C.c_int32),
        ("zero_fill", C.c_int32),
        ("n_rows", C.c_int64),
        ("n_other", C.c_int64),
        ("src", C.c_void_p),
        ("dst", C.c_void_p),
        ("index", C.c_void_p),
    ]


class ResizeProblem(C.Structure):
    """Mirror of ``struct pnb_resize_problem``."""

    _fields_ = [
        ("dtype", C.c_int32),
        ("method", C.c_int32),
        ("src_h", C.c_int32),
        ("src_w", C.c_int32),
        ("dst_h", C.c_int32),
        ("dst_w", C.c_int32),
        ("inner", C.c_int64),
        ("src", C.c_void_p),
        ("dst", C.c_void_p),
    ]


def build(verbose: bool = False, t1: bool = True) -> str:
    """Compile ``libpnb200.so`` for sm_100a with nvcc (works without a GPU)."""
    cmd = ["make", "-C", CSRC, "-j", str(min(16, os.cpu_count() or 4))]
    if t1:
        cmd.append("T1MODES=0 1 2")
    res = subprocess.run(cmd, stdout=subprocess.PIPE, stderr=subprocess.STDOUT, text=True)
    if verbose or res.returncode != 0:
        print(res.stdout)
    if res.returncode != 0:
        raise EngineError("building libpnb200.so failed")
    return SO_PATH


def load():
    """Load the shared library (no device needed for this)."""
    global _lib
    if _lib is not None:
        return _lib
    if not os.path.exists(SO_PATH):
        raise EngineError(
            f"{SO_PATH} is missing: build it with `python -c 'import __graft_entry__ as g; g.build()'` "
            "or `make -C pyneapple_b200/csrc`. pyneapple_b200 has no CPU fallback."
        )
    lib = C.CDLL(SO_PATH)
    lib.pnb_abi_version.restype = C.c_int
    lib.pnb_last_error.restype = C.c_char_p
    lib.pnb_device_count.restype = C.c_int
    lib.pnb_launch_count.restype = C.c_int64
    lib.pnb_trf_fit_device.argtypes = [C.POINTER(TrfProblem), C.c_void_p]
    lib.pnb_trf_fit_device.restype = C.c_int
    lib.pnb_trf_fit_host.argtypes = [C.POINTER(TrfProblem), C.c_int, C.c_int64]
    lib.pnb_trf_fit_host.restype = C.c_int
    lib.pnb_trf_fit_host_multi.argtypes = [C.POINTER(TrfProblem), C.POINTER(C.c_int32), C.c_int32, C.c_int64,
                                           C.POINTER(C.c_void_p)]
    lib.pnb_trf_fit_host_multi.restype = C.c_int
    lib.pnb_nnls_fit_host_multi.argtypes = [C.POINTER(NnlsProblem), C.POINTER(C.c_int32), C.c_int32, C.c_int64]
    lib.pnb_nnls_fit_host_multi.restype = C.c_int
    lib.pnb_trf_last_failed_count.restype = C.c_int64
    lib.pnb_nnls_fit_device.argtypes = [C.POINTER(NnlsProblem), C.c_void_p]
    lib.pnb_nnls_fit_device.restype = C.c_int
    lib.pnb_nnls_fit_host.argtypes = [C.POINTER(NnlsProblem), C.c_int, C.c_int64]
    lib.pnb_nnls_fit_host.restype = C.c_int
    lib.pnb_resize2d_device.argtypes = [C.POINTER(ResizeProblem), C.c_void_p]
    lib.pnb_resize2d_device.restype = C.c_int
    lib.pnb_resize2d_host.argtypes = [C.POINTER(ResizeProblem), C.c_int]
    lib.pnb_resize2d_host.restype = C.c_int
    lib.pnb_spectrum_peaks_device.argtypes = [C.POINTER(SpectrumProblem), C.c_void_p]
    lib.pnb_spectrum_peaks_device.restype = C.c_int
    lib.pnb_spectrum_peaks_host.argtypes = [C.POINTER(SpectrumProblem), C.c_int, C.c_int64]
    lib.pnb_spectrum_peaks_host.restype = C.c_int
    lib.pnb_segment_means_device.argtypes = [C.POINTER(SegmeansProblem), C.c_void_p]
    lib.pnb_segment_means_device.restype = C.c_int
    lib.pnb_segment_means_host.argtypes = [C.POINTER(SegmeansProblem), C.c_int]
    lib.pnb_segment_means_host.restype = C.c_int
    lib.pnb_predict_device.argtypes = [C.POINTER(PredictProblem), C.c_void_p]
    lib.pnb_predict_device.restype = C.c_int
    lib.pnb_move_rows_device.argtypes = [C.POINTER(RowsProblem), C.c_void_p]
    lib.pnb_move_rows_device.restype = C.c_int
    lib.pnb_nnls_dual_gemm_device.argtypes = [C.c_int32, C.c_int32, C.c_int64, C.c_void_p, C.c_void_p, C.c_void_p,
                                              C.c_void_p]
    lib.pnb_nnls_dual_gemm_device.restype = C.c_int
    lib.pnb_nnls_last_redo_count.argtypes = [C.c_int]
    lib.pnb_nnls_last_redo_count.restype = C.c_int64
    lib.pnb_host_alloc.argtypes = [C.POINTER(C.c_void_p), C.c_int64]
    lib.pnb_host_free.argtypes = [C.c_void_p]
    lib.pnb_download.argtypes = [C.c_void_p, C.c_void_p, C.c_int64, C.c_void_p]
    lib.pnb_download.restype = C.c_int
    lib.pnb_upload.argtypes = [C.c_void_p, C.c_void_p, C.c_int64, C.c_void_p]
    lib.pnb_upload.restype = C.c_int
    lib.pnb_copy_d2d.argtypes = [C.c_void_p, C.c_void_p, C.c_int64, C.c_void_p]
    lib.pnb_copy_d2d.restype = C.c_int
    lib.pnb_measure_fp64_peak.argtypes = [C.c_int, C.POINTER(C.c_double)]
    _lib = lib
    return lib


def require_device() -> None:
    lib = load()
    if lib.pnb_device_count() < 1:
        raise EngineError(
            "no CUDA device visible: pyneapple_b200 runs on B200 GPUs only and has no CPU fallback"
        )


def resolve_devices(device) -> list[int]:
    """``device`` as the solvers accept it -> list of CUDA ordinals: an int, a sequence of ints, or
    ``"all"`` (every visible GPU of the node)."""
    if isinstance(device, str):
        if device != "all":
            raise ValueError("device must be an int, a sequence of ints or 'all'")
        n = load().pnb_device_count()
        return list(range(max(1, n)))
    if isinstance(device, (list, tuple, np.ndarray)):
        out = [int(d) for d in device]
        if not out:
            raise ValueError("empty device list")
        return out
    return [int(device)]


def shard_ranges(n: int, parts: int) -> list[tuple[int, int]]:
    """The contiguous ranges pnb_*_fit_host_multi assigns to its devices."""
    base, rem = divmod(int(n), int(parts))
    out, start = [], 0
    for i in range(parts):
        stop = start + base + (1 if i < rem else 0)
        out.append((start, stop))
        start = stop
    return out


def check(rc: int, what: str) -> None:
    if rc != 0:
        msg = load().pnb_last_error().decode("utf-8", "replace")
        if rc == -1:
            raise ValueError(f"{what}: {msg}")
        if rc == -2:
            raise NotImplementedError(f"{what}: {msg}")
        raise EngineError(f"{what} failed (code {rc}): {msg}")


def launch_count() -> int:
    return int(load().pnb_launch_count())


def exported_symbols() -> list[str]:
    """Function names declared in ``include/pyneapple_b200.h``."""
    import re

    text = open(HEADER).read()
    text = re.sub(r"/\*.*?\*/", "", text, flags=re.S)
    return sorted(set(re.findall(r"\b(pnb_[a-z0-9_]+)\s*\(", text)))


class _PinnedBlock:
    """Owner of one cudaHostAlloc block; freed when the last array viewing it dies."""

    def __init__(self, nbytes: int):
        ptr = C.c_void_p()
        check(load().pnb_host_alloc(C.byref(ptr), nbytes), "pnb_host_alloc")
        self.ptr = ptr

    def __del__(self):  # pragma: no cover - best effort
        try:
            if self.ptr is not None:
                load().pnb_host_free(self.ptr)
                self.ptr = None
        except Exception:
            pass


def pinned_empty(shape, dtype=np.float64) -> np.ndarray:
    """A numpy array in page-locked host memory (full-speed, truly asynchronous H2D / D2H).

    The block is released when the array (and every view of it) is garbage collected.
    """
    shape = tuple(int(s) for s in np.atleast_1d(shape))
    dtype = np.dtype(dtype)
    count = int(np.prod(shape))
    nbytes = max(1, count * dtype.itemsize)
    block = _PinnedBlock(nbytes)
    buf = (C.c_char * nbytes).from_address(block.ptr.value)
    buf._pnb_block = block  # the ctypes buffer is the ndarray's base and keeps the block alive
    # np.ndarray(buffer=...) and not np.frombuffer(...).reshape(...): views of the latter reference the
    # hidden 1-D frombuffer array, not the array returned here, and the solvers decide from the returned
    # array's reference count whether a caller still holds (a view of) the previous fit's results
    return np.ndarray(shape=shape, dtype=dtype, buffer=buf)


def measure_fp64_peak(device: int = 0) -> float:
    out = C.c_double()
    check(load().pnb_measure_fp64_peak(device, C.byref(out)), "pnb_measure_fp64_peak")
    return float(out.value)
