#!/bin/bash
# batched finish: full -m gpu suite, 1-GPU bench, ncu --set full of the TRF kernel at bench size
mkdir -p gpurun_out
timeout 1500 python -m pytest tests -m gpu -q -x 2>&1 | tail -15 > gpurun_out/r2aa_tests.log
tail -3 gpurun_out/r2aa_tests.log
timeout 900 python bench.py --steps 10 --warmup 3 > gpurun_out/r2aa_bench.json 2> gpurun_out/r2aa_bench.err; echo "bench rc=$?"
timeout 300 python scripts/ncu_target.py trf 64 > gpurun_out/r2aa_plain_trf.log 2>&1 && \
  timeout 900 ncu --set full --clock-control none --import-source on -k regex:trf_kernel -s 2 -c 1 -f -o gpurun_out/r2aa_prof_trf \
      python scripts/ncu_target.py trf 64 > gpurun_out/r2aa_ncu_trf.log 2>&1
tail -2 gpurun_out/r2aa_ncu_trf.log
python - <<'PY'
import json
d=json.load(open('gpurun_out/r2aa_bench.json'))
print({k:d[k] for k in ('value','ms_per_step')}, d['e2e']['value'], d['e2e_fitter']['value'], d['e2e_pageable']['value'], d['c5']['value'], d['c5']['e2e']['value'], d['nnls_value'], d['nnls_e2e']['value'])
PY
