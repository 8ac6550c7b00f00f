#!/usr/bin/env python3
"""Per-region thread efficiency / instruction counts from `ncu --page source --csv` (SASS view).
usage: tools/ncu_source_hot.py <report.ncu-rep> [bucket]   -- aggregates consecutive SASS lines with equal exec count."""
import csv, io, subprocess, sys
rep = sys.argv[1]
raw = subprocess.run(["ncu", "-i", rep, "--page", "source", "--csv"], capture_output=True, text=True).stdout
rows = list(csv.reader(io.StringIO(raw)))
hdr = rows[1]
ix = {h: i for i, h in enumerate(hdr)}
data = rows[2:]
tot_i = sum(int(r[ix["Instructions Executed"]]) for r in data)
tot_t = sum(int(r[ix["Thread Instructions Executed"]]) for r in data)
print(f"total warp-inst {tot_i:.4g}, thread-inst {tot_t:.4g}, avg threads {tot_t/tot_i:.2f}")
# group consecutive lines by executed count
groups = []
cur = None
for k, r in enumerate(data):
    n = int(r[ix["Instructions Executed"]]); t = int(r[ix["Thread Instructions Executed"]])
    s = int(r[ix["# Samples"]])
    if cur and cur["n"] == n:
        cur["lines"] += 1; cur["t"] += t; cur["tot"] += n; cur["samples"] += s; cur["end"] = k
    else:
        cur = {"n": n, "lines": 1, "t": t, "tot": n, "start": k, "end": k, "samples": s, "first": r[ix["Source"]].strip()}
        groups.append(cur)
groups = [g for g in groups if g["tot"] > 0]
groups.sort(key=lambda g: -g["tot"])
print("top regions by warp-instructions executed:")
for g in groups[:18]:
    print(f"  lines {g['start']:5d}-{g['end']:5d} ({g['lines']:4d} instr) x {g['n']:.3g} = {100*g['tot']/tot_i:5.1f}% of issue, "
          f"avg threads {g['t']/g['tot']:5.2f}, samples {g['samples']:6d}  | {g['first'][:50]}")
