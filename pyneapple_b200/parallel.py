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


class PeerGather:
    """Gather equally shaped per-rank blocks into one buffer on rank ``dst`` through peer memory.

    Rank ``dst`` allocates ``(world, *shape)`` and publishes it with a CUDA IPC handle; every other
    rank maps it and writes its block with a device-to-device ``cudaMemcpyAsync`` over NVLink.  The
    copy engines do the transfer — no SM is needed, so it overlaps whatever kernel the rank launches
    next (the NCCL gather is a kernel and cannot share an SM with the register-bound TRF kernel).
    ``push`` only enqueues; ``wait`` makes the gathered data visible on ``dst`` (stream
    synchronisation + barrier).  Needs the ``nccl`` backend's world to live on one node.

    Measured on this pool's B200 boxes (virtualised, ``scripts/gpu_probe_peer.py``): a push through
    the cudaIpc mapping runs at 36 GB/s per rank — PCIe-class, although the same copy inside one
    process reaches 700 GB/s over NVLink and peer access is enabled — so it hides behind a 13 ms
    kernel for two to four ranks but not for eight.  ``bench.py`` therefore uses the NCCL gather
    (``gather_to_rank0``, ~570 GB/s) unless ``PNB_GATHER=peer``.
    """

    def __init__(self, shape, dtype, device, dst: int = 0):
        import torch
        import torch.distributed as dist

        self.world, self.rank, self.dst = dist.get_world_size(), dist.get_rank(), dst
        self.shape = tuple(shape)
        self.out = None
        self._side = None
        payload = [None]
        if self.rank == dst:
            try:
                self.out = torch.empty((self.world,) + self.shape, dtype=dtype, device=device)
                payload = [self.out.untyped_storage()._share_cuda_()]
            except Exception:  # the other ranks are waiting in the broadcast: tell them
                payload = [None]
        dist.broadcast_object_list(payload, src=dst)
        if payload[0] is None:
            raise RuntimeError("CUDA IPC export of the gather buffer failed on the destination rank")
        if self.rank == dst:
            self._remote = self.out
        else:
            storage = torch.UntypedStorage._new_shared_cuda(*payload[0])
            self._remote = torch.empty(0, dtype=dtype, device=storage.device).set_(
                storage, 0, (self.world,) + self.shape)

    def push(self, local):
        """Enqueue the transfer of this rank's block: a copy into one of two persistent staging buffers
        on the current stream, then the peer copy from there on a side stream, ordered after it (so the
        next kernel does not wait for the transfer).  Staging keeps the caller's tensor out of the
        picture: holding it until the peer copy is done (``record_stream``) made the caching allocator
        run out of reusable blocks and call ``cudaMalloc`` — a device-wide synchronisation — once the
        transfers got slower (eight ranks pushing into one GPU)."""
        import ctypes as C

        import torch

        from . import _lib

        dev = local.device
        if self._side is None:
            self._side = torch.cuda.Stream(dev)
            self._stage = [torch.empty(self.shape, dtype=local.dtype, device=dev) for _ in range(2)]
            self._done = [None, None]
            self._turn = 0
        t = self._turn
        self._turn ^= 1
        cur = torch.cuda.current_stream(dev)
        if self._done[t] is not None:
            cur.wait_event(self._done[t])  # the peer copy that last read this staging buffer (two pushes ago)
        self._stage[t].copy_(local, non_blocking=True)
        dst = self._remote[self.rank]
        with torch.cuda.device(dev):
            self._side.wait_stream(cur)
            _lib.check(_lib.load().pnb_copy_d2d(dst.data_ptr(), self._stage[t].data_ptr(),
                                                self._stage[t].numel() * self._stage[t].element_size(),
                                                C.c_void_p(self._side.cuda_stream)), "pnb_copy_d2d")
            self._done[t] = self._side.record_event()

    def flush(self):
        """Make the current stream wait for the pushes enqueued so far (timing: an event recorded
        afterwards covers the transfers)."""
        import torch

        if self._side is not None:
            torch.cuda.current_stream(self._side.device).wait_stream(self._side)

    def wait(self):
        import torch
        import torch.distributed as dist

        torch.cuda.synchronize()
        dist.barrier()
        return self.out
