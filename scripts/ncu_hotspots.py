"""Aggregate an ncu SASS source-page CSV by source line using nvdisasm -g line info (dev tool).

usage: ncu_hotspots.py <sass.csv from `ncu -i rep --page source --csv --print-source sass`> <nvdisasm -g -c output>
"""
import csv, re, sys, collections

sass_csv, disasm = sys.argv[1], sys.argv[2]
func = sys.argv[3] if len(sys.argv) > 3 else None  # substring of the mangled kernel name
# address -> (file, line) with inline context collapsed to innermost
addr2line = {}
cur = None
in_func = func is None
for ln in open(disasm):
    if ln.startswith("//--------------------- .text."):
        in_func = func is None or func in ln
        continue
    if not in_func:
        continue
    m = re.search(r'//## File "([^"]+)", line (\d+)', ln)
    if m:
        cur = (m.group(1).split("/")[-1], int(m.group(2)))
        continue
    m = re.match(r"\s*/\*([0-9a-f]{4,})\*/\s+(\S+)", ln)
    if m and cur:
        addr2line[int(m.group(1), 16)] = cur
rows = list(csv.reader(open(sass_csv)))
hdr = rows[1]
ix = {h: i for i, h in enumerate(hdr)}
agg = collections.defaultdict(lambda: [0, 0, 0, 0, 0, 0])
tot_i = tot_t = tot_s = tot_w = 0
base = None
for r in rows[2:]:
    try:
        a = int(r[ix["Address"]], 16) if r[ix["Address"]].startswith("0x") else int(r[ix["Address"]])
    except ValueError:
        continue
    if base is None:
        base = a
    key = addr2line.get(a - base, ("?", 0))
    ie = int(float(r[ix["Instructions Executed"]] or 0)); te = int(float(r[ix["Thread Instructions Executed"]] or 0))
    smp = int(float(r[ix["# Samples"]] or 0))
    op = r[ix["Source"]].split()[0] if r[ix["Source"]] else ""
    fp64 = op.startswith(("DFMA", "DMUL", "DADD", "DSETP", "MUFU"))
    wf = int(float(r[ix["L1 Wavefronts Shared"]] or 0)) if "L1 Wavefronts Shared" in ix else 0
    wfi = int(float(r[ix["L1 Wavefronts Shared Ideal"]] or 0)) if "L1 Wavefronts Shared Ideal" in ix else 0
    agg[key][0] += ie; agg[key][1] += te; agg[key][2] += smp; agg[key][3] += ie if fp64 else 0
    agg[key][4] += wf; agg[key][5] += wfi
    tot_i += ie; tot_t += te; tot_s += smp; tot_w += wf
print(f"total warp-inst {tot_i:.3e} thread-inst {tot_t:.3e} avg active {tot_t/max(tot_i,1):.2f} samples {tot_s}")
# group by function-ish ranges: print top lines
top = sorted(agg.items(), key=lambda kv: -kv[1][2])[:45]
print(f"{'file:line':34s} {'samples%':>8s} {'inst%':>7s} {'active':>6s} {'fp64%':>6s}")
for (f, l), (ie, te, smp, f64, wf, wfi) in top:
    print(f"{f+':'+str(l):34s} {100*smp/tot_s:8.2f} {100*ie/tot_i:7.2f} {te/max(ie,1):6.1f} {100*f64/max(ie,1):6.1f}")
if tot_w:
    print(f"shared-memory wavefronts: total {tot_w:.3e}")
    print(f"{'file:line':34s} {'wavefronts%':>11s} {'excess%':>8s} {'samples%':>8s}")
    for (f, l), (ie, te, smp, f64, wf, wfi) in sorted(agg.items(), key=lambda kv: -kv[1][4])[:40]:
        print(f"{f+':'+str(l):34s} {100*wf/tot_w:11.2f} {100*(wf-wfi)/max(wf,1):8.1f} {100*smp/tot_s:8.2f}")
# by file
byfile = collections.defaultdict(lambda: [0, 0, 0])
for (f, l), (ie, te, smp, f64, wf, wfi) in agg.items():
    byfile[f][0] += ie; byfile[f][1] += te; byfile[f][2] += smp
for f, (ie, te, smp) in byfile.items():
    print(f"FILE {f:28s} samples {100*smp/tot_s:6.2f}% inst {100*ie/tot_i:6.2f}% active {te/max(ie,1):5.1f}")
