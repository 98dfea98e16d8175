#!/bin/bash
# full -m gpu suite, then ncu --set full of the TRF kernel (branch-free exp build) at bench size
mkdir -p gpurun_out
timeout 1500 python -m pytest tests -m gpu -q -x 2>&1 | tail -15 > gpurun_out/r2t_tests.log
tail -5 gpurun_out/r2t_tests.log
timeout 300 python scripts/ncu_target.py trf 64 > gpurun_out/r2t_plain_trf.log 2>&1 && \
  timeout 900 ncu --set full --clock-control none --import-source on -k regex:trf_kernel -s 2 -c 1 -f -o gpurun_out/r2t_prof_trf \
      python scripts/ncu_target.py trf 64 > gpurun_out/r2t_ncu_trf.log 2>&1
tail -2 gpurun_out/r2t_ncu_trf.log
