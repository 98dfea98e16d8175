#!/bin/bash
# GPU run: full -m gpu test suite with complete failure output; IDEAL / C5 phase timing probes
python -m pytest tests -m gpu -q 2>&1 | tail -150 > gpurun_out/r2c_tests.log
python scripts/gpu_probe_ideal.py > gpurun_out/r2c_ideal.log 2>&1
python scripts/gpu_probe_c5.py > gpurun_out/r2c_c5.log 2>&1
tail -5 gpurun_out/r2c_tests.log; head -5 gpurun_out/r2c_ideal.log; tail -15 gpurun_out/r2c_c5.log
