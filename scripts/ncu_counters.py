"""Extract the roofline counters of ONE profiled launch from an .ncu-rep into profiles/<out>.json
(read by bench.py: roofline.traffic, roofline.frac_pipe):

    python scripts/ncu_counters.py <rep> <kernel substring> <voxels in the launch> <out.json> [note]
"""
import csv
import io
import json
import subprocess
import sys

rep, kern, voxels, out = sys.argv[1], sys.argv[2], int(sys.argv[3]), sys.argv[4]
note = sys.argv[5] if len(sys.argv) > 5 else ""
raw = subprocess.run(["ncu", "-i", rep, "--page", "raw", "--csv"], capture_output=True, text=True).stdout
rows = list(csv.reader(io.StringIO(raw)))
hdr, units, data = rows[0], rows[1], rows[2:]
name_col = hdr.index("Kernel Name")
rows_k = [r for r in data if kern in r[name_col]]
if not rows_k:
    raise SystemExit(f"no launch of a kernel matching {kern!r} in {rep}")
row = rows_k[-1]


def val(metric, default=None):
    if metric not in hdr:
        return default
    i = hdr.index(metric)
    v = float(row[i].replace(",", ""))
    u = units[i].lower()
    scale = {"kbyte": 1e3, "mbyte": 1e6, "gbyte": 1e9, "tbyte": 1e12, "byte": 1.0, "ms": 1e-3, "us": 1e-6, "ns": 1e-9,
             "s": 1.0, "second": 1.0, "msecond": 1e-3, "usecond": 1e-6, "nsecond": 1e-9}
    for key, f in scale.items():
        if u == key:
            return v * f
    return v


dram = (val("dram__bytes_read.sum", 0.0) or 0.0) + (val("dram__bytes_write.sum", 0.0) or 0.0)
pipe = val("sm__inst_executed_pipe_fp64.avg.pct_of_peak_sustained_active")
smem = val("l1tex__data_pipe_lsu_wavefronts_mem_shared.sum")
res = {
    "kernel": row[name_col], "voxels_per_launch": voxels,
    "dram_bytes_per_launch": dram, "dram_bytes_per_voxel": dram / voxels,
    "fp64_pipe_frac": None if pipe is None else pipe / 100.0,
    "smem_wavefronts_per_voxel": None if smem is None else smem / voxels,
    "smem_bank_conflicts_per_voxel": (val("l1tex__data_bank_conflicts_pipe_lsu_mem_shared.sum", 0.0) or 0.0) / voxels,
    "duration_under_ncu_s": val("gpu__time_duration.sum"),
    "warps_active_pct": val("sm__warps_active.avg.pct_of_peak_sustained_active"),
    "registers_per_thread": val("launch__registers_per_thread"),
    "issue_active_pct": val("smsp__issue_active.avg.pct_of_peak_sustained_active"),
    "source": f"ncu --set full --clock-control none, one launch of the bench size ({voxels} voxels): {rep.split('/')[-1]}" + (f"; {note}" if note else ""),
}
json.dump(res, open(out, "w"), indent=1)
print(json.dumps(res, indent=1))
