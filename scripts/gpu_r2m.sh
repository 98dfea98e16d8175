#!/bin/bash
# deferred covariance pass: parity tests that read covariances + the end-to-end probe
mkdir -p gpurun_out
timeout 900 python -m pytest tests/test_trf_gpu.py tests/test_multi_device_gpu.py tests/test_fitters_gpu.py -x -q -m gpu > gpurun_out/r2m_tests.log 2>&1; echo "tests rc=$?" >> gpurun_out/r2m_tests.log
tail -5 gpurun_out/r2m_tests.log
PNB_PROBE_QUICK=1 timeout 600 python scripts/gpu_probe_e2e_chunks.py > gpurun_out/r2m_chunks.log 2>&1
cat gpurun_out/r2m_chunks.log
