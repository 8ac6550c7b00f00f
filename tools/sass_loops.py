#!/usr/bin/env python3
"""Opcode histogram of the loops of one kernel (cuobjdump -sass output), to see what the hot loop issues.

usage: tools/sass_loops.py <object-or-so> <mangled-kernel-name-substring>
"""
import collections, re, subprocess, sys

obj, pat = sys.argv[1], sys.argv[2]
txt = subprocess.run(["cuobjdump", "-sass", obj], capture_output=True, text=True).stdout
blocks = txt.split("Function : ")
body = next(b for b in blocks if pat in b.split("\n", 1)[0])
ins = []
for m in re.finditer(r"^\s+/\*([0-9a-f]{4,6})\*/\s+(.*?);", body, re.M):
    ins.append((int(m.group(1), 16), m.group(2).strip()))
print("kernel:", body.split("\n", 1)[0], "instructions:", len(ins))
loops = []
for addr, t in ins:
    m = re.search(r"\bBRA(?:\.\w+)*\s+(?:[A-Za-z0-9!,]+\s*,\s*)?(0x[0-9a-f]+)", t)
    if m:
        tgt = int(m.group(1), 16)
        if tgt <= addr:
            loops.append((tgt, addr))
loops = sorted(set(loops), key=lambda l: l[0] - l[1])
print('back-edges (instructions):', [(hex(t), (a - t) // 16 + 1) for t, a in loops])
lo = int(sys.argv[3]) if len(sys.argv) > 3 else 200
for tgt, addr in [l for l in loops if (l[1] - l[0]) // 16 + 1 >= lo][-3:]:
    seg = [t for a, t in ins if tgt <= a <= addr]
    hist = collections.Counter()
    for t in seg:
        t = re.sub(r"^@!?U?P\d+\s+", "", t)
        hist[t.split()[0].split(".")[0] + ("." + t.split()[0].split(".")[1] if t.startswith("IMAD.MOV") else "")] += 1
    fp64 = sum(v for k, v in hist.items() if k in ("DFMA", "DMUL", "DADD", "DSETP"))
    print(f"\nloop 0x{tgt:x}..0x{addr:x}: {len(seg)} instructions, fp64 pipe {fp64} ({100*fp64/len(seg):.0f}%)")
    print("  " + ", ".join(f"{k}:{v}" for k, v in hist.most_common(24)))
