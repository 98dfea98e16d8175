#!/bin/bash
# evidence: ncu --set full of the NNLS fast kernel at bench size; launch list of the bench command
mkdir -p gpurun_out
timeout 300 python scripts/ncu_target.py nnls 64 > gpurun_out/r2ae_plain_nnls.log 2>&1 && \
  timeout 1200 ncu --set full --clock-control none --import-source on -k regex:nnls_v3 -s 2 -c 1 -f -o gpurun_out/r2ae_prof_nnls \
      python scripts/ncu_target.py nnls 64 > gpurun_out/r2ae_ncu_nnls.log 2>&1
tail -2 gpurun_out/r2ae_ncu_nnls.log
timeout 600 python bench.py --steps 2 --warmup 1 --no-cpu-baseline --no-extras > gpurun_out/r2ae_bench_short.json 2> gpurun_out/r2ae_bench_short.err && \
  timeout 1500 ncu --metrics gpu__time_duration.sum --clock-control none -c 600 --csv --log-file gpurun_out/r2ae_launches.csv \
      python bench.py --steps 2 --warmup 1 --no-cpu-baseline --no-extras > gpurun_out/r2ae_ncu_bench.log 2>&1
echo "launch list rc=$?"; wc -l gpurun_out/r2ae_launches.csv
