#!/usr/bin/env python
"""Opcode mix of one kernel from `ncu -i X.ncu-rep --page source --csv`: executed warp instructions and stall
samples per SASS opcode (development helper; the summaries that matter go to profiles/)."""
import csv
import sys
from collections import Counter

rows = list(csv.reader(open(sys.argv[1])))
hdr = rows[1]
ix = {h: i for i, h in enumerate(hdr)}
ex, smp = Counter(), Counter()
tot = 0
for r in rows[2:]:
    if len(r) < len(hdr):
        continue
    op = r[ix["Source"]].split()
    if not op:
        continue
    o = op[1] if op[0].startswith("@") else op[0]
    o = o.split(".")[0] + ("." + o.split(".")[1] if o.startswith(("LD", "ST", "ATOM", "RED", "SHFL", "MUFU")) and "." in o else "")
    n = int(r[ix["Instructions Executed"]] or 0)
    ex[o] += n
    smp[o] += int(r[ix["# Samples"]] or 0)
    tot += n
div = float(sys.argv[2]) if len(sys.argv) > 2 else 1.0
print(f"total executed {tot} (/{div:g} = {tot / div:.1f})")
ts = sum(smp.values())
for o, n in ex.most_common(45):
    print(f"{o:14s} {n:12d} {n / div:9.2f} {100 * n / tot:6.2f}%   samples {100 * smp[o] / ts:6.2f}%")
