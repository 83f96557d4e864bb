#!/usr/bin/env python
"""Dynamic SASS opcode mix of the cluster-scheme kernels from an `ncu --set full --import-source on` capture:

    python profiles/sass_mix_md.py gpurun_out/prof_r02_cluster.ncu-rep profiles/r02_sass_mix_cluster_kernels.md

Reads `ncu -i REP --page source --csv --print-source sass` per kernel (no GPU needed) and writes executed warp
instructions per opcode; for `cl_pair_kernel` also per trip (MUFU.RSQ64H executes once per chain = 4 per trip)."""
import csv
import io
import subprocess
import sys
from collections import Counter

rep, dst = sys.argv[1], sys.argv[2]
FP64 = ("DMUL", "DFMA", "DADD", "DSETP")


def mix(kernel):
    out = subprocess.run(["ncu", "-i", rep, "--page", "source", "--csv", "--print-source", "sass", "-k", "regex:" + kernel],
                         capture_output=True, text=True).stdout
    rows = list(csv.reader(io.StringIO(out)))
    hi = [i for i, r in enumerate(rows) if r and r[0] == "Address"][0]
    hdr = rows[hi]
    ix = {h: i for i, h in enumerate(hdr)}
    ex, seen = Counter(), set()
    for r in rows[hi + 1:]:
        if len(r) < len(hdr) or not r[0].startswith("0x") or r[0] in seen:
            continue
        seen.add(r[0])
        op = [t for t in r[ix["Source"]].split() if not t.startswith("@")]
        if not op:
            continue
        o = op[0]
        parts = o.split(".")
        o = parts[0] + ("." + parts[1] if parts[0] in ("LDS", "STS", "LDG", "STG", "SHFL", "MUFU", "ATOMG", "RED") and len(parts) > 1 else "")
        ex[o] += int(r[ix["Instructions Executed"]] or 0)
    return ex


with open(dst, "w") as f:
    f.write(f"# Dynamic SASS opcode mix of the cluster-scheme kernels (ions1m), from `ncu --page source` of `{rep}`\n")
    f.write("Executed warp instructions per opcode for one launch; for `cl_pair_kernel` also per trip (one trip = 32 lanes x 4 pair "
            "slots; `MUFU.RSQ64H` executes once per chain = 4 per trip).  Produced by `profiles/sass_mix_md.py`.\n")
    for k in ("cl_list_kernel<", "cl_pair_kernel<", "cl_reduce_kernel"):
        ex = mix(k.rstrip("<"))
        tot = sum(ex.values())
        trips = ex.get("MUFU.RSQ64H", 0) / 4 if k.startswith("cl_pair") else 0
        head = f"\n## `{k}` - {tot} warp instructions"
        if trips:
            head += f", {trips:.0f} trips, {tot / trips:.0f} per trip"
        f.write(head + "\n")
        f.write("| opcode | executed | share |" + (" per trip |" if trips else "") + "\n|---|---:|---:|" + ("---:|" if trips else "") + "\n")
        for o, n in ex.most_common(28):
            f.write(f"| `{o}` | {n} | {100 * n / tot:.1f} % |" + (f" {n / trips:.1f} |" if trips else "") + "\n")
        fp = sum(n for o, n in ex.items() if o in FP64)
        f.write(f"\nFP64 pipe instructions (DMUL + DFMA + DADD + DSETP): {fp} = {100 * fp / tot:.1f} %" +
                (f" = {fp / trips:.0f} per trip = {fp / trips / 4:.1f} per pair slot" if trips else "") + "\n")
