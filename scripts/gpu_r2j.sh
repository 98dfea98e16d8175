#!/bin/bash
# 1-GPU: full -m gpu suite, smoke, full bench, reference arm, launch list of the bench
python -m pytest tests -m gpu -q 2>&1 | tail -40 > gpurun_out/r2j_tests.log
python __graft_entry__.py smoke > gpurun_out/r2j_smoke.log 2>&1
python bench.py --steps 10 --warmup 3 > gpurun_out/r2j_bench_g1.json 2> gpurun_out/r2j_bench_g1.err; echo "bench rc $?"
python bench.py --impl reference --steps 3 --warmup 1 > gpurun_out/r2j_bench_ref.json 2> gpurun_out/r2j_bench_ref.err; echo "ref rc $?"
python bench.py --steps 2 --warmup 1 --no-cpu-baseline --no-c5 --no-extras > gpurun_out/r2j_plain.log 2>&1 && \
  ncu --metrics gpu__time_duration.sum --clock-control none -c 400 --csv --log-file gpurun_out/r2j_launches.csv \
      python bench.py --steps 2 --warmup 1 --no-cpu-baseline --no-c5 --no-extras > gpurun_out/r2j_ncu.log 2>&1
tail -4 gpurun_out/r2j_tests.log; cat gpurun_out/r2j_smoke.log | tail -2; head -c 300 gpurun_out/r2j_bench_g1.json; echo; head -c 400 gpurun_out/r2j_bench_ref.json
