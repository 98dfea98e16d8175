#!/bin/bash
# validation of the round's final state: full -m gpu suite, smoke(), IDEAL probe, 1-GPU bench
mkdir -p gpurun_out
timeout 1500 python -m pytest tests -m gpu -q -x 2>&1 | tail -4 > gpurun_out/r2ai_tests.log; cat gpurun_out/r2ai_tests.log
timeout 600 python -c "import __graft_entry__ as g; g.smoke(); print('smoke ok')" 2>&1 | tail -2
timeout 600 python scripts/gpu_probe_ideal.py 2>&1 | head -1 | tee gpurun_out/r2ai_ideal.log
timeout 900 python bench.py --steps 10 --warmup 3 > gpurun_out/r2ai_bench.json 2> gpurun_out/r2ai_bench.err; echo "bench rc=$?"
python - <<'PY'
import json
d=json.load(open("gpurun_out/r2ai_bench.json"))
print({k:d[k] for k in ("value","ms_per_step")}, "e2e", d["e2e"]["value"], "fitter", d["e2e_fitter"]["value"], "pageable", d["e2e_pageable"]["value"])
print("c5", d["c5"]["value"], d["c5"]["e2e"]["value"], "nnls", d["nnls_value"], d["nnls_e2e"]["value"])
PY
