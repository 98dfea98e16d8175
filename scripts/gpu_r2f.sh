#!/bin/bash
# 1-GPU run: NNLS screening A/B, full test suite, ncu of the NNLS kernel with screening at bench size
python scripts/gpu_probe_nnls_ab.py 64 > gpurun_out/r2f_nnls_ab.log 2>&1
python -m pytest tests -m gpu -q 2>&1 | tail -60 > gpurun_out/r2f_tests.log
python scripts/ncu_target.py nnls 64 > gpurun_out/r2f_plain_nnls.log 2>&1 && \
  ncu --set full --clock-control none --import-source on -k regex:nnls_v3 -s 2 -c 1 -o gpurun_out/r2f_prof_nnls \
      python scripts/ncu_target.py nnls 64 > gpurun_out/r2f_ncu_nnls.log 2>&1
cat gpurun_out/r2f_nnls_ab.log; tail -5 gpurun_out/r2f_tests.log
