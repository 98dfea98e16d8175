#!/bin/bash
mkdir -p gpurun_out
timeout 1200 python -m pytest tests/test_nnls_gpu.py tests/test_multi_device_gpu.py tests/test_spectrum_gpu.py -x -q -m gpu > gpurun_out/r2q_tests.log 2>&1; echo "tests rc=$?" >> gpurun_out/r2q_tests.log
tail -8 gpurun_out/r2q_tests.log
timeout 400 python scripts/gpu_probe_nnls_e2e.py 2>&1 | grep -v "fits failed" > gpurun_out/r2q_nnls_e2e.log
cat gpurun_out/r2q_nnls_e2e.log
