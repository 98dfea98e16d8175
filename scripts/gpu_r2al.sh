#!/bin/bash
# final state of round 2: full -m gpu suite, smoke(), 1-GPU bench, ncu --set full of the TRF kernel
mkdir -p gpurun_out
timeout 1500 python -m pytest tests -m gpu -q -x 2>&1 | tail -4 > gpurun_out/r2al_tests.log; cat gpurun_out/r2al_tests.log
timeout 600 python -c "import __graft_entry__ as g; g.smoke()" 2>&1 | tail -1
timeout 900 python bench.py --steps 10 --warmup 3 > gpurun_out/r2al_bench.json 2> gpurun_out/r2al_bench.err; echo "bench rc=$?"
timeout 300 python scripts/ncu_target.py trf 64 > gpurun_out/r2al_plain_trf.log 2>&1 && \
  timeout 900 ncu --set full --clock-control none --import-source on -k regex:trf_kernel -s 2 -c 1 -f -o gpurun_out/r2al_prof_trf \
      python scripts/ncu_target.py trf 64 > gpurun_out/r2al_ncu_trf.log 2>&1
tail -1 gpurun_out/r2al_ncu_trf.log
python - <<'PY'
import json
d=json.load(open("gpurun_out/r2al_bench.json"))
print({k:d[k] for k in ("value","ms_per_step")}, "e2e", d["e2e"]["value"], "fitter", d["e2e_fitter"]["value"], "pageable", d["e2e_pageable"]["value"])
print("c5", d["c5"]["value"], d["c5"]["e2e"]["value"], "nnls", d["nnls_value"], d["nnls_e2e"]["value"], d["nnls_e2e_pageable"]["value"])
PY
