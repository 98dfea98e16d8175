#!/bin/bash
# GPU run: full -m gpu test suite, 1-GPU bench, ncu --set full captures of the two kernels at bench size
python -m pytest tests -m gpu -q 2>&1 | tail -30 > gpurun_out/r2b_tests.log
python bench.py --steps 10 --warmup 3 > gpurun_out/r2b_bench.json 2> gpurun_out/r2b_bench.err; echo "bench rc $?"
python scripts/ncu_target.py trf 64 > gpurun_out/r2b_plain_trf.log 2>&1 && \
  ncu --set full --clock-control none --import-source on -k regex:trf_kernel -s 2 -c 1 -o gpurun_out/r2b_prof_trf \
      python scripts/ncu_target.py trf 64 > gpurun_out/r2b_ncu_trf.log 2>&1
python scripts/ncu_target.py nnls 64 > gpurun_out/r2b_plain_nnls.log 2>&1 && \
  ncu --set full --clock-control none --import-source on -k regex:nnls_v3 -s 2 -c 1 -o gpurun_out/r2b_prof_nnls \
      python scripts/ncu_target.py nnls 64 > gpurun_out/r2b_ncu_nnls.log 2>&1
tail -12 gpurun_out/r2b_tests.log; tail -3 gpurun_out/r2b_bench.err; head -c 600 gpurun_out/r2b_bench.json
