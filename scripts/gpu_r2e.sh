#!/bin/bash
# 2-GPU run: single-process multi-GPU probe, constrained test, 1-GPU bench TRF part (exp fix), 2-GPU bench
python scripts/gpu_probe_multi.py > gpurun_out/r2e_multi.log 2>&1
python -m pytest tests/test_constrained_gpu.py tests/test_trf_gpu.py -m gpu -q 2>&1 | tail -15 > gpurun_out/r2e_tests.log
python bench.py --steps 10 --warmup 3 --no-c5 --no-cpu-baseline > gpurun_out/r2e_bench_g1.json 2> gpurun_out/r2e_bench_g1.err
PNB_BENCH_GATHER=serial python -m torch.distributed.run --nnodes=1 --nproc-per-node 2 --master-addr 127.0.0.1 --master-port 29512 \
    bench.py --gpus 2 --steps 10 --warmup 3 --no-nnls --no-c5 --no-extras > gpurun_out/r2e_bench_g2_serial.json 2> gpurun_out/r2e_bench_g2_serial.err
python -m torch.distributed.run --nnodes=1 --nproc-per-node 2 --master-addr 127.0.0.1 --master-port 29513 \
    bench.py --gpus 2 --steps 10 --warmup 3 --no-nnls --no-c5 --no-extras > gpurun_out/r2e_bench_g2_overlap.json 2> gpurun_out/r2e_bench_g2_overlap.err
cat gpurun_out/r2e_multi.log; tail -4 gpurun_out/r2e_tests.log
python - <<'PY'
import json
for f in ("r2e_bench_g1", "r2e_bench_g2_serial", "r2e_bench_g2_overlap"):
    try:
        d = json.load(open(f"gpurun_out/{f}.json"))
        print(f, d["value"], d["ms_per_step"], d["roofline"]["kernel_ms"], d.get("kernel_ms_per_rank"), d["e2e"]["value"], d.get("nnls_ms_per_step"), (d.get("nnls") or {}).get("dual_init_ab"))
    except Exception as e:
        print(f, "failed", e)
PY
