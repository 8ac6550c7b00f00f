#!/usr/bin/env python3
"""Per CUDA source line: warp instructions executed, average active threads and stall samples, from
`ncu --page source --csv --print-source sass,cuda` (needs -lineinfo and --import-source on).
usage: tools/ncu_source_lines.py <report.ncu-rep> [top-n]"""
import collections, csv, io, subprocess, sys
rep = sys.argv[1]
top = int(sys.argv[2]) if len(sys.argv) > 2 else 40
raw = subprocess.run(["ncu", "-i", rep, "--page", "source", "--csv", "--print-source", "sass,cuda"], capture_output=True, text=True).stdout
agg = collections.defaultdict(lambda: [0, 0, 0, ""])
path, hdr = None, None
for r in csv.reader(io.StringIO(raw)):
    if not r: continue
    if r[0] == "File Path": path = r[1].split("/")[-1]; hdr = None; continue
    if r[0] == "Function Name": continue
    if r[0] == "Line No": hdr = {h: i for i, h in enumerate(r) if h not in ("Source",)}; src_i = 1; continue
    if hdr is None or not r[0].isdigit(): continue
    try:
        n = int(r[hdr["Instructions Executed"]] or 0); t = int(r[hdr["Thread Instructions Executed"]] or 0); s = int(r[hdr["# Samples"]] or 0)
    except (ValueError, IndexError):
        continue
    a = agg[(path, int(r[0]))]
    a[0] += n; a[1] += t; a[2] += s; a[3] = a[3] or r[src_i].strip()
tot = sum(a[0] for a in agg.values()); tots = sum(a[2] for a in agg.values())
print(f"total warp-inst {tot:.4g}, samples {tots}")
for (f, ln), a in sorted(agg.items(), key=lambda kv: -kv[1][2])[:top]:
    if a[0] == 0: continue
    print(f"{f}:{ln:5d}  inst {100*a[0]/tot:5.1f}%  samples {100*a[2]/tots:5.1f}%  thr {a[1]/a[0]:5.1f} | {a[3][:90]}")
print("per file:")
pf = collections.defaultdict(lambda: [0, 0, 0])
for (f, ln), a in agg.items():
    pf[f][0] += a[0]; pf[f][1] += a[1]; pf[f][2] += a[2]
for f, a in sorted(pf.items(), key=lambda kv: -kv[1][0]):
    if a[0]: print(f"  {f:28s} inst {100*a[0]/tot:5.1f}%  samples {100*a[2]/tots:5.1f}%  thr {a[1]/a[0]:5.1f}")
if len(sys.argv) > 4:
    f0 = sys.argv[3]
    print("ranges of", f0)
    for rg in sys.argv[4:]:
        lo, hi = map(int, rg.split("-"))
        n = sum(a[0] for (f, ln), a in agg.items() if f == f0 and lo <= ln <= hi)
        t = sum(a[1] for (f, ln), a in agg.items() if f == f0 and lo <= ln <= hi)
        s = sum(a[2] for (f, ln), a in agg.items() if f == f0 and lo <= ln <= hi)
        print(f"  {rg:12s} inst {100*n/tot:5.1f}%  samples {100*s/tots:5.1f}%  thr {t/max(n,1):5.1f}")
