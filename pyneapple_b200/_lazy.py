"""Arrays that are computed during a fit but only cross PCIe when somebody looks at them.

The covariance of a 4-parameter fit is 128 of the 188 result bytes per voxel; the reference's fitters
store it (``diagnostics_["pcov"]``, ``FitResult.covariance``) and most scripts never read it.  The
TRF kernels always compute it; with ``want_cov=True`` (the default) the host path leaves it in
device memory (``pnb_trf_fit_host`` with a device pointer in ``cov``) behind this ndarray look-alike,
which downloads on first use (``pnb_download``) and then behaves like the array it has become.
"""

from __future__ import annotations

import numpy as np


class LazyArray:
    """``(n, ...)`` float64 array whose rows still live on one or several GPUs.

    ``parts``: list of ``(start, stop, cuda_tensor)`` covering rows ``[0, n)``.
    """

    __array_priority__ = 0.0

    def __init__(self, shape, parts, dtype=np.float64):
        self.shape = tuple(int(s) for s in shape)
        self.dtype = np.dtype(dtype)
        self._parts = list(parts)
        self._host = None

    # -- ndarray surface -------------------------------------------------------------------
    @property
    def ndim(self):
        return len(self.shape)

    @property
    def size(self):
        return int(np.prod(self.shape))

    @property
    def nbytes(self):
        return self.size * self.dtype.itemsize

    @property
    def on_device(self) -> bool:
        return self._host is None

    def __len__(self):
        return self.shape[0]

    def materialize(self) -> np.ndarray:
        """Download (once) and return the host array; the device copies are released."""
        if self._host is None:
            from . import engine

            out = np.empty(self.shape, dtype=self.dtype)
            for start, stop, t in self._parts:
                if stop > start:
                    out[start:stop] = engine.to_host(t)
            self._host = out
            self._parts = []
        return self._host

    def __array__(self, dtype=None, copy=None):
        a = self.materialize()
        if dtype is not None and np.dtype(dtype) != a.dtype:
            return a.astype(dtype)
        return a.copy() if copy else a

    def __getitem__(self, idx):
        if self._host is None and isinstance(idx, (int, np.integer)):
            i = int(idx) + (self.shape[0] if idx < 0 else 0)
            if not 0 <= i < self.shape[0]:
                raise IndexError(idx)
            for start, stop, t in self._parts:  # one voxel: fetch just that block
                if start <= i < stop:
                    return t[i - start].cpu().numpy()
        return self.materialize()[idx]

    def __iter__(self):
        return iter(self.materialize())

    def __repr__(self):
        where = "device" if self._host is None else "host"
        return f"LazyArray(shape={self.shape}, dtype={self.dtype}, on {where})"

    def __getattr__(self, name):
        # everything else (reshape, mean, T, ...) is the materialised array's business
        if name.startswith("_"):
            raise AttributeError(name)
        return getattr(self.materialize(), name)
