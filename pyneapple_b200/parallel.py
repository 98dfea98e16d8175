"""Multi-GPU plumbing: one process per GPU, voxels sharded, one gather to rank 0.

Voxels are independent optimisation problems, so the only
communication is the final gather of the parameter maps; IDEAL couples voxels
within a z-slice only, so volumes are cut into contiguous z-slabs.  The
functions work with any ``torch.distributed`` backend (``nccl`` on the GPUs,
``gloo`` in the CPU tests).
"""

from __future__ import annotations

import numpy as np


def shard_bounds(n: int, world: int) -> list[tuple[int, int]]:
    """Contiguous, balanced ``[start, stop)`` ranges of ``n`` units over ``world`` ranks."""
    base, rem = divmod(int(n), int(world))
    out, start = [], 0
    for r in range(world):
        stop = start + base + (1 if r < rem else 0)
        out.append((start, stop))
        start = stop
    return out


def slab_bounds(mask_per_slice: np.ndarray, world: int) -> list[tuple[int, int]]:
    """Contiguous z-slabs balanced by the number of masked voxels per slice.

    ``mask_per_slice[z]`` = voxels to fit in slice ``z``.  Every rank gets at
    least one slice when ``Z >= world``.
    """
    counts = np.asarray(mask_per_slice, dtype=np.float64)
    Z = counts.shape[0]
    if world >= Z:
        return [(min(r, Z), min(r + 1, Z)) for r in range(world)]
    cum = np.concatenate([[0.0], np.cumsum(counts)])
    total = cum[-1]
    cuts = [0]
    for r in range(1, world):
        target = total * r / world
        z = int(np.searchsorted(cum, target, side="left"))
        z = max(z, cuts[-1] + 1)
        z = min(z, Z - (world - r))
        cuts.append(z)
    cuts.append(Z)
    return [(cuts[r], cuts[r + 1]) for r in range(world)]


def gather_to_rank0(local, sizes: list[int], dim: int = -1, dst: int = 0, concat: bool = True, out=None):
    """Gather per-rank tensors that differ in length along ``dim`` to ``dst``.

    Returns the concatenated tensor on ``dst`` and ``None`` elsewhere.  One
    ``torch.distributed.gather`` (ranks pad to the largest shard).
    ``concat=False`` returns the blocks stacked along a new leading axis, ``(world, *padded shape)``
    — what the collective delivers, without the extra pass over the data that joining them along
    ``dim`` costs (the caller scatters every rank's block into its z-slab of the volume anyway);
    ``out`` is an optional preallocated buffer of that stacked shape on ``dst``.
    """
    import torch
    import torch.distributed as dist

    world = dist.get_world_size()
    rank = dist.get_rank()
    dim = dim % local.ndim
    big = max(sizes)
    if local.shape[dim] != sizes[rank]:
        raise ValueError(f"rank {rank}: local size {local.shape[dim]} != declared {sizes[rank]}")
    if local.shape[dim] < big:
        pad_shape = list(local.shape)
        pad_shape[dim] = big - local.shape[dim]
        local = torch.cat([local, local.new_zeros(pad_shape)], dim=dim)
    local = local.contiguous()
    bufs = None
    if rank == dst:
        if out is None or tuple(out.shape) != (world,) + tuple(local.shape) or out.dtype != local.dtype:
            out = torch.empty((world,) + tuple(local.shape), dtype=local.dtype, device=local.device)
        bufs = list(out.unbind(0))
    dist.gather(local, bufs, dst=dst)
    if rank != dst:
        return None
    if not concat:
        return out
    return torch.cat([b.narrow(dim, 0, s) for b, s in zip(bufs, sizes)], dim=dim)
