#!/usr/bin/env python
"""Benchmark of the voxel-fitting hot path (BASELINE.json metric: voxel fits/sec).

    python bench.py [--gpus N] [--steps K] [--warmup W] [--impl reference]

A *step* is one pass of the hot path over one full synthetic volume of the
workload BASELINE.json's metric is quoted on (configs[1]: pixelwise biexp
bounded curvefit, S0 mode, 256 x 256 x 64 voxels x 16 b-values = 4 194 304
fits), i.e. what ``CurveFitSolver.fit`` does for that volume.

* ``value``      — voxels/s with the volume already resident in HBM
                   (``pnb_trf_fit_device``), timed with CUDA events on the
                   launching stream; multi-GPU: every rank fits its own volume
                   (weak scaling, z-slabs of an N-times deeper volume), the
                   parameter maps are gathered to rank 0 (NCCL) inside the
                   timed region, time = max over ranks.
* ``e2e``        — the same metric through the reference-facing call with HOST
                   buffers (``CurveFitSolver.fit`` -> ``pnb_trf_fit_host``):
                   pinned host input, H2D / kernels / D2H inside the timed region.
* ``roofline``   — FP64-arithmetic roofline of the TRF kernel (it is FP64-CUDA-core
                   bound, not HBM- or tensor-bound) plus the HBM
                   fraction for completeness.
* ``cpu_baseline`` — the oracle port (same SciPy calls as the reference, joblib
                   over all host cores) on a bounded sample of the same workload.

``--impl reference`` times only that CPU path and prints the same JSON shape.
"""

from __future__ import annotations

import argparse
import json
import os
import subprocess
import sys
import threading
import time

ROOT = os.path.dirname(os.path.abspath(__file__))
sys.path.insert(0, ROOT)
os.environ.setdefault("PYNEAPPLE_QUIET", "1")

import numpy as np  # noqa: E402

METRIC = "voxel fits/sec (biexp NLLS, 250-bin NNLS) at 1/2/4/8 B200 vs scipy CPU"
UNIT = "voxel fits/s"
WORKLOAD = "C2: pixelwise biexp (fit_s0) bounded curvefit, 256x256x64 voxels x 16 b-values, TRF, max_iter 250, tol 1e-8"
C_EXP = 25  # FP64 flop per exp (SURVEY.md §8d convention)


def algorithmic_flops(n, m, k_exp, nfev_sum, njev_sum):
    """F = nfev*F_f + njev*F_J + nit*F_it (nit ~ njev)."""
    f_f = m * (k_exp * (C_EXP + 3) + 2) + 2 * m
    f_j = 3 * m * n
    f_svd = 6 * (n * (n - 1) // 2) * (6 * (m + n) + 12)
    f_it = 2 * m * n + 3 * m * n + 2 * n + f_svd + 10 * 6 * n + 3 * (2 * m * n + 6 * n) + 12 * n
    return nfev_sum * f_f + njev_sum * (f_j + f_it)


class ClockSampler:
    """nvidia-smi clocks / throttle reasons sampled every 200 ms while the timed region runs."""

    Q = ("index,clocks.sm,clocks.max.sm,power.draw,clocks_event_reasons.active,"
         "clocks_event_reasons.hw_slowdown,clocks_event_reasons.hw_thermal_slowdown,"
         "clocks_event_reasons.sw_thermal_slowdown,clocks_event_reasons.sw_power_cap")

    def __init__(self, gpu_index=0, n_gpus=1):
        self.rows = []
        self.gpu = ",".join(str(gpu_index + i) for i in range(max(1, n_gpus)))  # every GPU of the job
        self.proc = None

    def start(self):
        try:
            self.proc = subprocess.Popen(
                ["nvidia-smi", f"--query-gpu={self.Q}", "--format=csv,noheader,nounits", "-lms", "200",
                 "-i", str(self.gpu)],
                stdout=subprocess.PIPE, stderr=subprocess.DEVNULL, text=True)
            self.t = threading.Thread(target=self._read, daemon=True)
            self.t.start()
        except OSError:
            self.proc = None

    def _read(self):
        for line in self.proc.stdout:
            self.rows.append([c.strip() for c in line.split(",")])

    def stop(self):
        if not self.proc:
            return {"sm_mhz": None, "sm_max_mhz": None, "reasons": ["nvidia-smi unavailable"]}
        time.sleep(0.25)
        self.proc.terminate()
        try:
            self.proc.wait(timeout=2)
        except subprocess.TimeoutExpired:
            self.proc.kill()
        sm, mx, reasons, per_gpu = [], [], set(), {}
        names = ["hw_slowdown", "hw_thermal_slowdown", "sw_thermal_slowdown", "sw_power_cap"]
        for r in self.rows:
            if len(r) < 9:
                continue
            try:
                sm.append(float(r[1])); mx.append(float(r[2]))
                per_gpu.setdefault(r[0], []).append(float(r[1]))
            except ValueError:
                continue
            for nm, v in zip(names, r[5:9]):
                if v.lower().startswith("active"):
                    reasons.add(nm)
        med = {g: float(np.median(v)) for g, v in per_gpu.items()}
        out = {
            # median under load of the SLOWEST GPU of the job (the timing is the max over ranks)
            "sm_mhz": min(med.values()) if med else None,
            "sm_max_mhz": float(max(mx)) if mx else None,
            "reasons": sorted(reasons),
            "samples": len(sm),
        }
        if len(med) > 1:
            out["sm_mhz_per_gpu"] = med
        return out


def nnls_algorithmic_flops(m, n_bins, w, iters, k_final):
    """SURVEY.md §8(d) K4 model, evaluated from the device counters (iterations, final active size).

    adds = (iters + k_final) / 2 outer iterations, mean active size ~0.6 k_final.
    """
    iters = np.asarray(iters, np.float64)
    k = np.asarray(k_final, np.float64)
    n_outer = 0.5 * (iters + k)
    kbar = 0.6 * k
    per_outer = 2 * m * n_bins + 2 * (2 * w + 1) * n_bins + 2 * m * kbar + (2 * w + 1) * kbar
    per_inner = 3 * kbar * kbar + 2 * m * kbar
    return float(np.sum(2 * m * n_bins + n_outer * per_outer + iters * per_inner))


def _traffic(name):
    """DRAM bytes per launch of the dominant kernel from the committed ncu capture (profiles/), or None."""
    try:
        return json.load(open(os.path.join(ROOT, "profiles", name))).get("dram_bytes_per_launch")
    except (OSError, ValueError):
        return None


def bench_nnls(args, world, rank, local_rank, dev):
    """250-bin NNLS half of the metric: config C3 (d_range [0.0008, 0.5], reg_order 2, mu 0.02)."""
    import torch
    import torch.distributed as dist

    from pyneapple_b200 import _lib, engine, models, synth
    from pyneapple_b200.solvers import NNLSSolver
    from pyneapple_b200.solvers.nnls import regularization_matrix

    base = synth.CONFIGS["C3"]
    # weak scaling: every rank fits its own volume of the configuration (same parameter fields, its
    # own noise), i.e. slab `rank` of a stack of C3 volumes
    cfg = synth.Config(**{**base.__dict__, "shape": (base.shape[0], base.shape[1], args.slices)})
    b, img, _ = synth.make_volume(cfg, 0, args.slices, replica=rank)
    y_host = img.reshape(-1, b.shape[0])
    del img
    n_vox, n_b = y_host.shape
    model = models.NNLSModel(d_range=(0.0008, 0.5), n_bins=250)
    basis = model.get_basis(b)
    R = regularization_matrix(250, 2, 0.02)
    y_dev = torch.as_tensor(y_host).to(dev)
    steps = max(1, min(args.steps, 3))
    r = engine.nnls_fit(basis, R, y_dev, 250, device=local_rank)  # warm-up
    torch.cuda.synchronize()
    if world > 1:
        dist.barrier()
    e0, e1 = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
    nsampler = ClockSampler(0, world)
    if rank == 0:
        nsampler.start()
    l0 = _lib.launch_count()
    e0.record()
    for _ in range(steps):
        r = None  # release the previous step's 8.4 GB of coefficients first: no cudaMalloc in the timed region
        r = engine.nnls_fit(basis, R, y_dev, 250, device=local_rank)
    e1.record()
    torch.cuda.synchronize()
    launches = _lib.launch_count() - l0
    nclocks = nsampler.stop() if rank == 0 else None
    ms = torch.tensor([e0.elapsed_time(e1)], dtype=torch.float64, device=dev)
    if world > 1:
        dist.all_reduce(ms, op=dist.ReduceOp.MAX)
    ms = float(ms.item())
    value = n_vox * world * steps / (ms * 1e-3)
    iters = r["iterations"].cpu().numpy()
    k_final = (r["coefficients"] > 0).sum(dim=1).cpu().numpy()
    ok_rate = float((r["status"] == 1).double().mean().item())
    del r
    torch.cuda.empty_cache()
    # e2e through NNLSSolver.fit with pinned host buffers
    y_pin = _lib.pinned_empty(y_host.shape)
    y_pin[...] = y_host
    solver = NNLSSolver(model=model, reg_order=2, mu=0.02, max_iter=250, device=local_rank, pinned_outputs=True)
    solver.fit(b, y_pin)
    torch.cuda.synchronize()
    if world > 1:
        dist.barrier()
    t0 = time.perf_counter()
    solver.fit(b, y_pin)
    e2e_s = torch.tensor([time.perf_counter() - t0], dtype=torch.float64, device=dev)
    if world > 1:
        dist.all_reduce(e2e_s, op=dist.ReduceOp.MAX)
    e2e_value = n_vox * world / float(e2e_s.item())
    if rank != 0:
        return None
    flops = nnls_algorithmic_flops(n_b, 250, 2, iters, k_final)
    kernel_ms = ms / steps
    fp64_peak = _lib.measure_fp64_peak(local_rank)
    alg_bytes = n_vox * (8 * n_b + 8 * 250 + 8 + 4 + 4 + 8)
    out = {
        "workload": "C3: pixelwise NNLS n_bins=250, d_range [0.0008, 0.5], reg_order=2, mu=0.02, max_iter=250, "
                    "256x256x64 voxels x 16 b-values",
        "value": value, "unit": UNIT, "steps": steps, "ms_per_step": ms / steps, "gpu_launches": int(launches),
        "clocks": nclocks,
        "voxels_per_gpu": n_vox, "success_rate": ok_rate, "mean_iterations": float(iters.mean()),
        "mean_active_set": float(k_final.mean()), "max_active_set": int(k_final.max()),
        "e2e": {"value": e2e_value, "unit": UNIT, "h2d_bytes_per_step": int(y_host.nbytes),
                "d2h_bytes_per_step": int(n_vox * (250 * 8 + 8 + 4 + 4 + 8)),
                "api": "NNLSSolver.fit(numpy pinned) -> pnb_nnls_fit_host"},
        "roofline": {"bound": "fp64", "kernel": "nnls_v3_kernel<16,2> (+ nnls_kernel<2> for the voxels it hands over)", "achieved": flops / (kernel_ms * 1e-3) / 1e12,
                     "peak": fp64_peak, "unit": "TFLOP/s", "frac": flops / (kernel_ms * 1e-3) / 1e12 / fp64_peak,
                     "flops_per_launch": flops, "flop_model": "SURVEY.md §8(d) K4, from device iteration counters",
                     "kernel_ms": kernel_ms, "traffic": _traffic("nnls_traffic.json"),
                     "hbm": {"achieved": alg_bytes / (kernel_ms * 1e-3) / 1e9, "unit": "GB/s",
                             "algorithmic_bytes_per_launch": alg_bytes},
                     "shared_memory": "the fast kernel moves 0.80 shared-memory wavefronts per clock and SM (1.0 = "
                                      "the pipe's limit; profiles/r1_final_summary.md): half of them are the "
                                      "dictionary columns of the dual pass"},
    }
    if not args.no_cpu_baseline:
        from oracle import ref_port

        cores = os.cpu_count() or 1
        sb, sy, _ = synth.sample_voxels(base, 32768, z=0)
        ref_port.nnls_fit(sb, sy[: 8 * cores], (0.0008, 0.5), 250, 2, 0.02, 250, n_jobs=-1)
        t0 = time.perf_counter()
        probe = ref_port.nnls_fit(sb, sy[: 64 * cores], (0.0008, 0.5), 250, 2, 0.02, 250, n_jobs=-1)
        rate = 64 * cores / (time.perf_counter() - t0)
        n = int(min(32768, max(64 * cores, rate * 10.0)))
        t0 = time.perf_counter()
        ref = ref_port.nnls_fit(sb, sy[:n], (0.0008, 0.5), 250, 2, 0.02, 250, n_jobs=-1)
        dt = time.perf_counter() - t0
        del probe
        out["cpu_baseline"] = {"value": n / dt, "unit": UNIT, "cores": cores, "kind": "port",
                               "sample": f"{n} voxels, scipy.optimize.nnls per voxel via oracle/ref_port.py, "
                                         f"joblib n_jobs=-1, {dt:.1f} s"}
        chk = NNLSSolver(model=model, reg_order=2, mu=0.02, max_iter=250, device=local_rank).fit(sb, sy[:n])
        out["parity_vs_cpu_sample"] = {
            "voxels": n,
            "max_abs_coefficient_diff": float(np.abs(chk.params_["coefficients"] - ref["coefficients"]).max()),
            "max_abs_residual_diff": float(np.abs(chk.diagnostics_["residual"] - ref["residual"]).max()),
            "success_flags_equal": bool(((chk.status_ == 1) == ref["success"]).all()),
        }
    return out


def problem_arrays(cfg):
    names = ["f1", "D1", "D2", "S0"]
    return (names, np.array([cfg.p0[n] for n in names]), np.array([cfg.bounds[n][0] for n in names]),
            np.array([cfg.bounds[n][1] for n in names]))


def cpu_reference_run(b, y, cfg, n_jobs):
    """The oracle port: one scipy curve_fit(method='trf') per voxel, joblib pool over all cores."""
    from oracle import ref_port

    names, p0, lb, ub = problem_arrays(cfg)
    n = y.shape[0]
    model = ref_port.Model("biexp", "s0")
    P0, LB, UB = (np.tile(v[:, None], (1, n)) for v in (p0, lb, ub))
    t = time.perf_counter()
    res = ref_port.curvefit_fit(model, b, y, P0, LB, UB, max_iter=250, tol=1e-8, n_jobs=n_jobs)
    return time.perf_counter() - t, res


def cpu_baseline(cfg, target_seconds=15.0, max_vox=65536):
    """Bounded sample of the workload on all host cores; returns (vox/s, description dict)."""
    from pyneapple_b200 import synth

    cores = os.cpu_count() or 1
    b, y, _ = synth.sample_voxels(cfg, min(max_vox, 65536), z=0)
    # warm the loky pool and probe the rate
    cpu_reference_run(b, y[: 8 * cores], cfg, -1)
    dt, _ = cpu_reference_run(b, y[: 32 * cores], cfg, -1)
    rate = 32 * cores / dt
    n = int(min(max_vox, max(64 * cores, rate * target_seconds)))
    n = min(n, y.shape[0])
    dt, res = cpu_reference_run(b, y[:n], cfg, -1)
    return n / dt, {
        "value": n / dt, "unit": UNIT, "cores": cores, "kind": "port",
        "sample": f"{n} voxels (every k-th voxel of slice 0 of the C2 volume), scipy {__import__('scipy').__version__} "
                  f"curve_fit per voxel via oracle/ref_port.py, joblib loky pool n_jobs=-1, {dt:.1f} s",
    }, (b, y[:n], res)


def run_reference(args):
    from pyneapple_b200 import synth

    rank = int(os.environ.get("RANK", "0"))
    if rank != 0:
        return 0
    cfg = synth.CONFIGS["C2"]
    cores = os.cpu_count() or 1
    b, y, _ = synth.sample_voxels(cfg, 65536, z=0)
    cpu_reference_run(b, y[: 8 * cores], cfg, -1)  # pool warm-up
    dt, _ = cpu_reference_run(b, y[: 32 * cores], cfg, -1)
    rate = 32 * cores / dt
    per_step = int(min(65536, max(64 * cores, rate * max(2.0, 60.0 / max(1, args.steps + args.warmup)))))
    for _ in range(args.warmup):
        cpu_reference_run(b, y[:per_step], cfg, -1)
    t = time.perf_counter()
    for _ in range(args.steps):
        cpu_reference_run(b, y[:per_step], cfg, -1)
    total = time.perf_counter() - t
    value = per_step * args.steps / total
    line = {
        "impl": "reference", "metric": METRIC, "value": value, "unit": UNIT, "n_gpus": args.gpus,
        "steps": args.steps, "warmup": args.warmup, "ms_per_step": 1e3 * total / args.steps,
        "higher_is_better": True, "scaling": "weak", "vs_baseline": None, "dtype": "f64",
        "data": "synthetic", "config": {"workload": WORKLOAD, "sample_per_step": per_step},
        "cpu_baseline": {"value": value, "unit": UNIT, "cores": cores, "kind": "port",
                         "sample": f"{per_step} voxels per step, scipy curve_fit per voxel, joblib n_jobs=-1"},
        "e2e": {"value": value, "unit": UNIT, "h2d_bytes_per_step": 0, "d2h_bytes_per_step": 0},
        "gpu_launches": 0,
    }
    _emit(line)
    return 0


_REAL_STDOUT = None


def _quiet_stdout():
    """Route file descriptor 1 to stderr for the duration of the run: stdout must carry exactly one
    JSON line, and libraries (NCCL with NCCL_DEBUG=VERSION, for one) print there."""
    global _REAL_STDOUT
    if _REAL_STDOUT is None:
        sys.stdout.flush()
        _REAL_STDOUT = os.dup(1)
        os.dup2(2, 1)


def _emit(line: dict):
    sys.stdout.flush()
    data = (json.dumps(line) + "\n").encode()
    if _REAL_STDOUT is None:
        os.write(1, data)
    else:
        os.write(_REAL_STDOUT, data)


def main():
    _quiet_stdout()
    ap = argparse.ArgumentParser()
    ap.add_argument("--gpus", type=int, default=1)
    ap.add_argument("--steps", type=int, default=10)
    ap.add_argument("--warmup", type=int, default=3)
    ap.add_argument("--impl", default="b200", choices=["b200", "reference"])
    ap.add_argument("--jac", default="reference", choices=["reference", "analytic"])
    ap.add_argument("--slices", type=int, default=64, help="z-slices per GPU (64 = the full C2 volume)")
    ap.add_argument("--no-cpu-baseline", action="store_true")
    ap.add_argument("--no-nnls", action="store_true")
    args = ap.parse_args()
    if args.impl == "reference":
        return run_reference(args)

    import torch
    import torch.distributed as dist

    from pyneapple_b200 import _lib, engine, models, parallel, synth
    from pyneapple_b200.solvers import CurveFitSolver

    world = int(os.environ.get("WORLD_SIZE", "1"))
    rank = int(os.environ.get("RANK", "0"))
    local_rank = int(os.environ.get("LOCAL_RANK", "0"))
    if not torch.cuda.is_available():
        raise SystemExit("bench.py needs a CUDA device: pyneapple_b200 has no CPU fallback")
    torch.cuda.set_device(local_rank)
    dev = torch.device("cuda", local_rank)
    if world > 1:
        dist.init_process_group("nccl", device_id=dev)

    base = synth.CONFIGS["C2"]
    # weak scaling: every rank fits its own C2 volume (same parameter fields, its own noise), i.e.
    # z-slab `rank` of a stack of `world` C2 volumes.  Cutting one world-times deeper volume instead
    # would give every rank a different part of the parameter ranges and a different amount of work.
    cfg = synth.Config(**{**base.__dict__, "shape": (base.shape[0], base.shape[1], args.slices)})
    b, img, _ = synth.make_volume(cfg, 0, args.slices, replica=rank)
    n_b = b.shape[0]
    y_host = img.reshape(-1, n_b)
    del img
    n_vox = y_host.shape[0]
    names, p0, lb, ub = problem_arrays(cfg)
    model = models.BiExpModel(fit_s0=True)
    desc = models.describe_model(model)
    jac_mode = engine.JAC_TWO_POINT if args.jac == "reference" else engine.JAC_ANALYTIC

    # ------------------------------------------------------------ device-resident steps
    y_dev = torch.as_tensor(y_host).to(dev)

    gathered = [None]
    peer = None
    # PNB_GATHER=peer: CUDA-IPC peer-memory pushes instead of the NCCL gather (slower on this pool:
    # 36 GB/s per pusher through cudaIpc mappings against ~570 GB/s for NCCL, profiles/r1_final_multi_gpu.md)
    if world > 1 and os.environ.get("PNB_GATHER", "nccl") == "peer":
        try:
            peer = parallel.PeerGather((4, n_vox), torch.float64, dev)
        except Exception as exc:  # no CUDA IPC in this environment: NCCL gather instead
            print(f"rank {rank}: peer-memory gather unavailable ({exc}); using the NCCL gather", file=sys.stderr)
        ok = torch.tensor([1 if peer is not None else 0], device=dev)
        dist.all_reduce(ok, op=dist.ReduceOp.MIN)
        if int(ok.item()) == 0:
            peer = None

    def device_step():
        r = engine.trf_fit(desc, b, y_dev, p0, lb, ub, 0, max_nfev=250, ftol=1e-8, jac_mode=jac_mode,
                           device=local_rank)
        if world > 1:
            # the blocks arrive stacked (world, n_params, n_vox): every rank's block is the parameter
            # map of its z-slab
            if peer is not None:
                peer.push(r["params"])  # copy engines over NVLink peer memory; completed by the sync below
            else:
                gathered[0] = parallel.gather_to_rank0(r["params"], [n_vox] * world, dim=1, concat=False,
                                                       out=gathered[0])
        return r

    for _ in range(args.warmup):
        r = device_step()
    torch.cuda.synchronize()
    if world > 1:
        dist.barrier()
    sampler = ClockSampler(0, world)
    if rank == 0:
        sampler.start()
    launches0 = _lib.launch_count()
    ev = [torch.cuda.Event(enable_timing=True) for _ in range(args.steps + 1)]
    torch.cuda.synchronize()
    ev[0].record()
    for i in range(args.steps):
        r = device_step()
        if peer is not None and i == args.steps - 1:
            peer.flush()  # the last event must cover the transfers still in flight on the side stream
        ev[i + 1].record()
    torch.cuda.synchronize()
    if world > 1:
        dist.barrier()
    total_ms = ev[0].elapsed_time(ev[-1])
    kernel_launches = _lib.launch_count() - launches0
    if peer is not None:
        # the blocks really are on rank 0: compare a checksum of every rank's parameters
        mine = r["params"].sum().reshape(1)
        sums = [torch.empty_like(mine) for _ in range(world)]
        dist.all_gather(sums, mine)
        if rank == 0:
            got = peer.out.sum(dim=(1, 2))
            want = torch.cat(sums)
            if not torch.allclose(got, want, rtol=1e-12, atol=0):
                raise SystemExit(f"peer-memory gather delivered wrong data: {got.tolist()} vs {want.tolist()}")
    t = torch.tensor([total_ms], dtype=torch.float64, device=dev)
    if world > 1:
        dist.all_reduce(t, op=dist.ReduceOp.MAX)
    total_ms = float(t.item())
    value = n_vox * world * args.steps / (total_ms * 1e-3)

    # per-launch kernel duration for the roofline (kernel only, no gather): CUDA events around
    # five back-to-back launches on the launching stream
    kev0, kev1 = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
    r = engine.trf_fit(desc, b, y_dev, p0, lb, ub, 0, max_nfev=250, ftol=1e-8, jac_mode=jac_mode, device=local_rank)
    torch.cuda.synchronize()
    kev0.record()
    for _ in range(5):
        r = engine.trf_fit(desc, b, y_dev, p0, lb, ub, 0, max_nfev=250, ftol=1e-8, jac_mode=jac_mode,
                           device=local_rank)
    kev1.record()
    torch.cuda.synchronize()
    kernel_ms = kev0.elapsed_time(kev1) / 5
    nfev_sum = int(r["nfev"].sum().item())
    njev_sum = int(r["njev"].sum().item())
    success = float((r["status"] > 0).double().mean().item())

    # ------------------------------------------------------------ e2e through the solver API, host buffers
    y_pin = _lib.pinned_empty(y_host.shape)
    y_pin[...] = y_host
    solver = CurveFitSolver(model=model, p0=base.p0, bounds=base.bounds, max_iter=250, tol=1e-8,
                            jac=args.jac, device=local_rank, pinned_outputs=True)
    e2e_steps = max(2, min(args.steps, 5))
    for _ in range(2):
        solver.fit(b, y_pin)
    torch.cuda.synchronize()
    if world > 1:
        dist.barrier()
    t0 = time.perf_counter()
    for _ in range(e2e_steps):
        solver.fit(b, y_pin)
    e2e_s = time.perf_counter() - t0
    t = torch.tensor([e2e_s], dtype=torch.float64, device=dev)
    if world > 1:
        dist.all_reduce(t, op=dist.ReduceOp.MAX)
    e2e_s = float(t.item())
    e2e_value = n_vox * world * e2e_steps / e2e_s
    clocks = sampler.stop() if rank == 0 else None
    h2d = y_host.nbytes + b.nbytes + 3 * 4 * 8
    d2h = n_vox * (4 * 8 + 16 * 8 + 4 + 4 + 4 + 8 + 8)

    del y_pin, solver, y_dev
    torch.cuda.empty_cache()
    nnls = None if args.no_nnls else bench_nnls(args, world, rank, local_rank, dev)
    if rank != 0:
        if world > 1:
            dist.destroy_process_group()
        return 0

    # ------------------------------------------------------------ roofline + baselines (rank 0)
    try:
        peaks = json.load(open(os.path.join(ROOT, "MEASURED_PEAKS.json")))
        hbm_peak, hbm_src = float(peaks["hbm_gbs"]), "MEASURED_PEAKS.json"
    except (OSError, KeyError, ValueError):
        hbm_peak, hbm_src = 6650.0, "fallback (B200_PROFILING.md)"
    fp64_peak = _lib.measure_fp64_peak(local_rank)
    flops = algorithmic_flops(4, n_b, 2, nfev_sum, njev_sum)
    alg_bytes = n_vox * (8 * n_b + 8 * 4 + 8 * 16 + 4 + 4 + 4 + 8 + 8)
    achieved_tf = flops / (kernel_ms * 1e-3) / 1e12
    achieved_gbs = alg_bytes / (kernel_ms * 1e-3) / 1e9
    traffic = _traffic("trf_traffic.json")
    roofline = {
        "bound": "fp64", "kernel": "trf_kernel<Model<BiS0>,128>", "achieved": achieved_tf, "peak": fp64_peak,
        "unit": "TFLOP/s", "frac": achieved_tf / fp64_peak,
        "peak_source": "DFMA micro-benchmark run by this process (MEASURED_PEAKS.json holds no FP64 figure)",
        "flops_per_launch": flops, "flop_model": "SURVEY.md §8(d): nfev*F_f + njev*(F_J+F_it), C_exp=25",
        "kernel_ms": kernel_ms, "traffic": traffic,
        "hbm": {"achieved": achieved_gbs, "peak": hbm_peak, "unit": "GB/s", "frac": achieved_gbs / hbm_peak,
                "algorithmic_bytes_per_launch": alg_bytes, "peak_source": hbm_src},
    }
    extra = {}
    cpu = None
    if not args.no_cpu_baseline:
        _, cpu, (sb, sy, sres) = cpu_baseline(base)
        # parity of the GPU path against that same CPU sample, voxel for voxel
        solver2 = CurveFitSolver(model=model, p0=base.p0, bounds=base.bounds, max_iter=250, tol=1e-8,
                                 jac=args.jac, device=local_rank)
        solver2.fit(sb, sy)
        got = np.stack([solver2.params_[n] for n in names], axis=0)
        ok = sres["success"]
        rel = (np.abs(got - sres["params"]) / np.abs(sres["params"]))[:, ok].max(axis=0)
        extra["parity_vs_cpu_sample"] = {
            "voxels": int(ok.sum()), "max_rel_param_diff": float(rel.max()),
            "frac_within_1e-4": float((rel <= 1e-4).mean()),
            "success_flags_equal": bool(((solver2.status_ > 0) == ok).all()),
        }
    line = {
        "metric": METRIC, "value": value, "unit": UNIT, "n_gpus": world, "steps": args.steps,
        "warmup": args.warmup, "ms_per_step": total_ms / args.steps, "higher_is_better": True,
        "scaling": "weak", "vs_baseline": None, "dtype": "f64", "data": "synthetic",
        "config": {"workload": WORKLOAD, "voxels_per_gpu": n_vox, "jacobian": args.jac,
                   "l2": "inputs (537 MB per GPU) exceed the 126 MB L2, no flush needed",
                   "multi_gpu": "z-slabs of a stack of C2 volumes, one volume (same parameter fields, own noise) per rank, parameter maps gathered to "
                                "rank 0 in the timed region ("
                                + ("peer-memory copies over NVLink" if peer is not None else "NCCL gather") + ")",
                   "success_rate": success, "mean_nfev": nfev_sum / n_vox},
        "clocks": clocks,
        "e2e": {"value": e2e_value, "unit": UNIT, "h2d_bytes_per_step": int(h2d), "d2h_bytes_per_step": int(d2h),
                "steps": e2e_steps, "api": "CurveFitSolver.fit(numpy pinned) -> pnb_trf_fit_host"},
        "gpu_launches": int(kernel_launches),
        "roofline": roofline,
        "cpu_baseline": cpu,
    }
    line.update(extra)
    line["nnls"] = nnls
    _emit(line)
    if world > 1:
        dist.destroy_process_group()
    return 0


if __name__ == "__main__":
    sys.exit(main())
