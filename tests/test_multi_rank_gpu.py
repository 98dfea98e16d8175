"""One process per GPU over NCCL (the `torchrun` mode of SURVEY.md §8e): every rank fits its z-slab on its own
GPU, the parameter maps are gathered to rank 0, and the assembled volume equals the single-GPU fit bit for bit —
for the TRF path, the constrained path (device-resident) and an IDEAL fit cut into slabs.  Needs two GPUs."""

from __future__ import annotations

import os
import subprocess
import sys
import textwrap

import pytest

pytestmark = pytest.mark.gpu

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))

WORKER = textwrap.dedent(r'''
    import os, sys
    sys.path.insert(0, %r)
    import numpy as np, torch, torch.distributed as dist
    from pyneapple_b200 import models, parallel, synth
    from pyneapple_b200.fitters import IDEALFitter
    from pyneapple_b200.solvers import ConstrainedCurveFitSolver, CurveFitSolver

    rank, world, lr = int(os.environ["RANK"]), int(os.environ["WORLD_SIZE"]), int(os.environ["LOCAL_RANK"])
    torch.cuda.set_device(lr)
    dev = torch.device("cuda", lr)
    dist.init_process_group("nccl", device_id=dev)
    ok = True

    # --- pixelwise biexp TRF over z-slabs -------------------------------------------------------
    cfg = synth.Config(**{**synth.CONFIGS["C2"].__dict__, "shape": (64, 64, 6)})
    b, img, _ = synth.make_volume(cfg)
    slabs = parallel.shard_bounds(cfg.shape[2], world)
    z0, z1 = slabs[rank]
    kw = dict(model=models.BiExpModel(fit_s0=True), p0=cfg.p0, bounds=cfg.bounds, max_iter=250, tol=1e-8)
    y_slab = torch.as_tensor(np.ascontiguousarray(img[:, :, z0:z1]).reshape(-1, 16)).to(dev)
    res = CurveFitSolver(device=lr, **kw).fit_device(b, y_slab)
    sizes = [(e - s) * 64 * 64 for s, e in slabs]
    blocks = parallel.gather_to_rank0(res["params"], sizes, dim=1, concat=False)
    if rank == 0:
        whole = CurveFitSolver(device=0, **kw).fit(b, img.reshape(-1, 16))
        vol = np.stack([whole.params_[n] for n in ("f1", "D1", "D2", "S0")]).reshape(4, 64, 64, 6)
        for r, (s, e) in enumerate(slabs):
            got = blocks[r][:, : sizes[r]].cpu().numpy().reshape(4, 64, 64, e - s)
            ok = ok and np.array_equal(got, vol[:, :, :, s:e])

    # --- constrained tri-exponential, device-resident -------------------------------------------
    c5 = synth.Config(**{**synth.CONFIGS["C5"].__dict__, "shape": (64, 64, 4)})
    b5, img5, _ = synth.make_volume(c5)
    slabs5 = parallel.shard_bounds(4, world)
    s5, e5 = slabs5[rank]
    ckw = dict(model=models.TriExpModel(), p0=c5.p0, bounds=c5.bounds, want_cov=False, **c5.solver_kwargs)
    y5 = torch.as_tensor(np.ascontiguousarray(img5[:, :, s5:e5]).reshape(-1, 24)).to(dev)
    r5 = ConstrainedCurveFitSolver(device=lr, **ckw).fit_device(b5, y5)
    sizes5 = [(e - s) * 64 * 64 for s, e in slabs5]
    blk5 = parallel.gather_to_rank0(r5["params"], sizes5, dim=1, concat=False)
    if rank == 0:
        whole5 = ConstrainedCurveFitSolver(device=0, **ckw).fit(b5, img5.reshape(-1, 24))
        vol5 = np.stack([whole5.params_[n] for n in ("f1", "D1", "f2", "D2", "D3")]).reshape(5, 64, 64, 4)
        for r, (s, e) in enumerate(slabs5):
            got = blk5[r][:, : sizes5[r]].cpu().numpy().reshape(5, 64, 64, e - s)
            ok = ok and np.array_equal(got, vol5[:, :, :, s:e])

    # --- IDEAL: a slab fit equals the same slab of the full fit ---------------------------------------
    seg = synth.ellipsoid_mask(cfg.shape)
    steps = np.array([[8, 8], [16, 16], [32, 32], [64, 64]])
    tol = {"S0": 0.5, "f1": 0.2, "D1": 0.2, "D2": 0.2}
    mine = IDEALFitter(CurveFitSolver(device=lr, **kw), steps, tol).fit(b, img, seg, z_range=(z0, z1))
    last = torch.as_tensor(np.ascontiguousarray(np.moveaxis(mine.step_params[-1], 2, 0))).to(dev)  # (z, x, y, p)
    gathered = parallel.gather_to_rank0(last, [e - s for s, e in slabs], dim=0)
    if rank == 0:
        full = IDEALFitter(CurveFitSolver(device=0, **kw), steps, tol).fit(b, img, seg)
        ok = ok and np.array_equal(np.moveaxis(gathered.cpu().numpy(), 0, 2), full.step_params[-1])

    flag = torch.tensor([1 if ok else 0], device=dev)
    dist.all_reduce(flag, op=dist.ReduceOp.MIN)
    dist.destroy_process_group()
    if rank == 0:
        print("MULTI_RANK_OK" if int(flag.item()) == 1 else "MULTI_RANK_MISMATCH")
''')


def test_two_ranks_over_nccl_reproduce_the_single_gpu_fit(tmp_path):
    from pyneapple_b200 import _lib

    if _lib.load().pnb_device_count() < 2:
        pytest.skip("needs two GPUs")
    script = tmp_path / "worker.py"
    script.write_text(WORKER % ROOT)
    r = subprocess.run(
        [sys.executable, "-m", "torch.distributed.run", "--nnodes=1", "--nproc-per-node", "2", "--master-addr",
         "127.0.0.1", "--master-port", "29541", str(script)],
        capture_output=True, text=True, timeout=600)
    assert r.returncode == 0, r.stderr[-2000:]
    assert "MULTI_RANK_OK" in r.stdout, (r.stdout[-500:], r.stderr[-1500:])
