"""Per-voxel results back into image space (mirror of ``reconstruct_maps``, io/nifti.py:279-312).

The reference turns ``pixel_indices`` (a list of one tuple per voxel) into index arrays with
``tuple(zip(*pixel_indices))`` — seconds for a 4 M-voxel volume — and assigns on the host.  Here the
indices are an array (``PixelIndices``); values that still live on the GPU (CUDA tensors, lazy arrays)
are scattered and converted to float32 there (``pnb_move_rows_device``), so only the float32 volume
crosses PCIe, and host arrays take one vectorised numpy assignment.
"""

from __future__ import annotations

import numpy as np

from . import engine


def _flat_index(pixel_indices, spatial_shape):
    flat = getattr(pixel_indices, "flat", None)
    if flat is not None and getattr(pixel_indices, "_shape", None) in (None, tuple(spatial_shape)):
        return flat
    if getattr(pixel_indices, "is_full", False):
        return None
    coords = pixel_indices.array if hasattr(pixel_indices, "array") else np.asarray(list(pixel_indices))
    if coords.size == 0:
        return np.zeros(0, np.int64)
    return np.ravel_multi_index(tuple(np.asarray(coords).T), spatial_shape).astype(np.int64)


def reconstruct_maps(fitted_params: dict, pixel_indices, spatial_shape: tuple, device: int = 0) -> dict:
    """``name -> float32 volume`` of shape ``spatial_shape (+ extra dims)``, zero where no voxel was fitted."""
    spatial_shape = tuple(int(s) for s in spatial_shape)
    n_out = int(np.prod(spatial_shape))
    full = getattr(pixel_indices, "is_full", False)
    flat = None if full else _flat_index(pixel_indices, spatial_shape)
    maps = {}
    flat_dev = None
    for name, values in fitted_params.items():
        if engine._is_torch_cuda(values):
            import torch

            v = values.to(torch.float64)
            extra = tuple(v.shape[1:])
            if flat is None:
                vol = v.to(torch.float32)
            else:
                if flat_dev is None or flat_dev.device != v.device:
                    flat_dev = engine.to_device(flat, v.device)
                vol = engine.move_rows(v, flat_dev, n_out, scatter=True, out_float32=True)
            maps[name] = engine.to_host(vol).reshape(spatial_shape + extra)
            continue
        v = np.asarray(values)
        extra = v.shape[1:] if v.ndim > 1 else ()
        if flat is None:
            maps[name] = v.astype(np.float32).reshape(spatial_shape + extra)
            continue
        vol = np.zeros((n_out,) + extra, dtype=np.float32)
        vol[flat] = v.astype(np.float32)
        maps[name] = vol.reshape(spatial_shape + extra)
    del device
    return maps
