#!/bin/bash
# 2-GPU run: TRF value after the constant-cache fix (N=1), gather A/B at N=2, full 2-GPU bench
python bench.py --steps 10 --warmup 3 --no-nnls --no-c5 --no-extras --no-cpu-baseline > gpurun_out/r2k_g1_trf.json 2> gpurun_out/r2k_g1_trf.err
PNB_BENCH_GATHER=serial python -m torch.distributed.run --nnodes=1 --nproc-per-node 2 --master-addr 127.0.0.1 --master-port 29512 \
    bench.py --gpus 2 --steps 10 --warmup 3 --no-nnls --no-c5 --no-extras > gpurun_out/r2k_g2_serial.json 2> gpurun_out/r2k_g2_serial.err
python -m torch.distributed.run --nnodes=1 --nproc-per-node 2 --master-addr 127.0.0.1 --master-port 29513 \
    bench.py --gpus 2 --steps 10 --warmup 3 > gpurun_out/r2k_g2.json 2> gpurun_out/r2k_g2.err
python - <<'PY'
import json
for f in ("r2k_g1_trf", "r2k_g2_serial", "r2k_g2"):
    try:
        d = json.load(open(f"gpurun_out/{f}.json"))
        print(f, round(d["value"] / 1e6, 1), round(d["ms_per_step"], 3), [round(k, 3) for k in d["kernel_ms_per_rank"]], round(d["e2e"]["value"] / 1e6, 1),
              d.get("e2e_one_process"), d.get("nnls_value"), (d.get("c5") or {}).get("value"))
    except Exception as e:
        print(f, "failed", e)
PY
tail -3 gpurun_out/r2k_g2.err
