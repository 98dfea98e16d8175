"""Host-side sharding logic and the world_size-2 gather on the gloo backend (CPU)."""

import os

import numpy as np
import pytest
import torch
import torch.distributed as dist
import torch.multiprocessing as mp

from pyneapple_b200 import parallel


def test_shard_bounds_cover_and_balance():
    for n in (0, 1, 7, 8, 4194304, 33554433):
        for w in (1, 2, 3, 8):
            b = parallel.shard_bounds(n, w)
            assert b[0][0] == 0 and b[-1][1] == n
            assert all(b[i][1] == b[i + 1][0] for i in range(w - 1))
            sizes = [e - s for s, e in b]
            assert max(sizes) - min(sizes) <= 1


def test_slab_bounds_balance_masked_voxels():
    z = np.arange(64)
    counts = np.maximum(0, 1000 - (z - 31.5) ** 2)  # ellipsoid-like profile
    slabs = parallel.slab_bounds(counts, 8)
    assert slabs[0][0] == 0 and slabs[-1][1] == 64
    assert all(e > s for s, e in slabs)
    per = [counts[s:e].sum() for s, e in slabs]
    assert max(per) < 1.35 * np.mean(per)
    assert parallel.slab_bounds(np.ones(4), 8)[:4] == [(0, 1), (1, 2), (2, 3), (3, 4)]


def _worker(rank, world, port, q):
    os.environ["MASTER_ADDR"] = "127.0.0.1"
    os.environ["MASTER_PORT"] = str(port)
    dist.init_process_group("gloo", rank=rank, world_size=world)
    n = 11
    bounds = parallel.shard_bounds(n, world)
    s, e = bounds[rank]
    full = torch.arange(4 * n, dtype=torch.float64).reshape(4, n)
    local = full[:, s:e].clone()
    sizes = [b[1] - b[0] for b in bounds]
    out = parallel.gather_to_rank0(local, sizes, dim=1)
    # stacked form: what the collective delivers (ranks padded to the largest shard), buffer reused
    buf = torch.empty((world, 4, max(sizes)), dtype=torch.float64) if rank == 0 else None
    blocks = parallel.gather_to_rank0(local, sizes, dim=1, concat=False, out=buf)
    if rank == 0:
        ok = bool(torch.equal(out, full)) and blocks.data_ptr() == buf.data_ptr()
        for r, (bs, be) in enumerate(bounds):
            ok = ok and bool(torch.equal(blocks[r][:, : be - bs], full[:, bs:be]))
        q.put(ok)
    else:
        assert out is None and blocks is None
    dist.destroy_process_group()


def test_gather_to_rank0_gloo_world2():
    ctx = mp.get_context("spawn")
    q = ctx.Queue()
    port = 29500 + (os.getpid() % 2000)
    procs = [ctx.Process(target=_worker, args=(r, 2, port, q)) for r in range(2)]
    for p in procs:
        p.start()
    for p in procs:
        p.join(120)
        assert p.exitcode == 0
    assert q.get(timeout=5) is True
