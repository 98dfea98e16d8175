#!/bin/bash
# 2-GPU run: full -m gpu suite (multi-device tests really use two GPUs), then the 2-GPU bench
python -m pytest tests -m gpu -q 2>&1 | tail -120 > gpurun_out/r2d_tests.log
python -m torch.distributed.run --nnodes=1 --nproc-per-node 2 --master-addr 127.0.0.1 --master-port 29511 \
    bench.py --gpus 2 --steps 10 --warmup 3 > gpurun_out/r2d_bench_g2.json 2> gpurun_out/r2d_bench_g2.err
echo "bench rc $?"
tail -6 gpurun_out/r2d_tests.log; tail -3 gpurun_out/r2d_bench_g2.err; head -c 400 gpurun_out/r2d_bench_g2.json
