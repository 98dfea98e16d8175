#!/bin/bash
# 8-GPU bench (as the driver launches it)
python -m torch.distributed.run --nnodes=1 --nproc-per-node 8 --master-addr 127.0.0.1 --master-port 29520 \
    bench.py --gpus 8 --steps 10 --warmup 3 > gpurun_out/r2l_g8.json 2> gpurun_out/r2l_g8.err
echo "rc $?"
python - <<'PY'
import json
d = json.load(open("gpurun_out/r2l_g8.json"))
print(round(d["value"] / 1e6, 1), round(d["ms_per_step"], 3), [round(k, 2) for k in d["kernel_ms_per_rank"]], round(d["e2e"]["value"] / 1e6, 1),
      d.get("e2e_one_process"), d.get("nnls_value"), d.get("nnls_e2e"), (d.get("c5") or {}).get("value"), (d.get("c5") or {}).get("ms_per_step"))
PY
tail -3 gpurun_out/r2l_g8.err
