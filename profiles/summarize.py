#!/usr/bin/env python
"""Turn the raw ncu artefacts that `gpurun` brings back (gpurun_out/) into the small, committed
summaries under profiles/.

    python profiles/summarize.py launches gpurun_out/launches_ions1m.csv profiles/r01_launches_ions1m.md
    python profiles/summarize.py full gpurun_out/prof_cells.ncu-rep profiles/r01_ncu_cells_pair_kernel.md

`launches`: per-kernel totals of an `ncu --metrics gpu__time_duration.sum` launch list (cold-cache,
serialised times: compare SHARES, not absolutes).  `full`: the headline metrics of one `ncu --set full`
capture (needs the `ncu` binary, which this image has; no GPU needed to read a report).
"""
from __future__ import annotations

import csv
import json
import re
import subprocess
import sys
from collections import OrderedDict


def short(name: str) -> str:
    name = re.sub(r"^void ", "", name)
    name = re.sub(r"\(.*$", "", name)
    name = re.sub(r"jme::", "", name)
    return name[:90]


def launches(src: str, dst: str) -> None:
    rows = []
    with open(src) as fh:
        lines = [l for l in fh if l.startswith('"')]
    for r in csv.DictReader(lines):
        if r.get("Metric Name") == "gpu__time_duration.sum":
            rows.append((short(r["Kernel Name"]), float(r["Metric Value"]), r["Grid Size"], r["Block Size"]))
    agg = OrderedDict()
    for name, ns, grid, block in rows:
        a = agg.setdefault(name, {"launches": 0, "ns": 0.0, "grid": grid, "block": block})
        a["launches"] += 1
        a["ns"] += ns
    total = sum(a["ns"] for a in agg.values())
    mine = {k: v for k, v in agg.items() if not k.startswith("at::") and "nccl" not in k.lower()}
    with open(dst, "w") as fh:
        fh.write(f"# ncu launch list summary - {src}\n\n")
        fh.write("`ncu --metrics gpu__time_duration.sum --clock-control none` over `bench.py --steps 3 --warmup 3`;\n"
                 "times are cold-cache and serialised: the kernel SHARES are the evidence, not the absolutes.\n\n")
        fh.write(f"launches captured: {len(rows)}, total {total / 1e3:.1f} us, of which this repo's kernels "
                 f"{sum(v['ns'] for v in mine.values()) / 1e3:.1f} us\n\n")
        fh.write("| kernel | launches | total us | avg us | share of all | grid | block |\n|---|---:|---:|---:|---:|---|---|\n")
        for k, v in sorted(agg.items(), key=lambda kv: -kv[1]["ns"]):
            fh.write(f"| `{k}` | {v['launches']} | {v['ns'] / 1e3:.1f} | {v['ns'] / 1e3 / v['launches']:.2f} | "
                     f"{100 * v['ns'] / total:.1f} % | {v['grid']} | {v['block']} |\n")


WANTED = OrderedDict([
    ("gpu__time_duration.sum", "duration"),
    ("sm__cycles_active.avg", "SM active cycles"),
    ("sm__inst_executed.avg.per_cycle_active", "IPC per SM (max 4)"),
    ("smsp__inst_executed.sum", "warp instructions executed"),
    ("sm__inst_executed_pipe_fp64.avg.pct_of_peak_sustained_active", "FP64 pipe % of peak"),
    ("sm__inst_executed_pipe_alu.avg.pct_of_peak_sustained_active", "ALU pipe % of peak"),
    ("sm__inst_executed_pipe_fma.avg.pct_of_peak_sustained_active", "FMA (FP32) pipe % of peak"),
    ("sm__inst_executed_pipe_lsu.avg.pct_of_peak_sustained_active", "LSU pipe % of peak"),
    ("sm__inst_executed_pipe_xu.avg.pct_of_peak_sustained_active", "XU pipe % of peak"),
    ("smsp__sass_thread_inst_executed_op_dfma_pred_on.sum", "DFMA thread instructions"),
    ("smsp__sass_thread_inst_executed_op_dmul_pred_on.sum", "DMUL thread instructions"),
    ("smsp__sass_thread_inst_executed_op_dadd_pred_on.sum", "DADD thread instructions"),
    ("sm__warps_active.avg.pct_of_peak_sustained_active", "achieved occupancy %"),
    ("launch__registers_per_thread", "registers / thread"),
    ("launch__grid_size", "grid"),
    ("launch__block_size", "block"),
    ("dram__bytes_read.sum", "DRAM bytes read"),
    ("dram__bytes_write.sum", "DRAM bytes written"),
    ("gpu__dram_throughput.avg.pct_of_peak_sustained_elapsed", "DRAM throughput % of peak"),
    ("l1tex__t_sector_hit_rate.pct", "L1 hit rate %"),
    ("lts__t_sector_hit_rate.pct", "L2 hit rate %"),
    ("l1tex__data_pipe_lsu_wavefronts.avg", "L1 data-pipe wavefronts per SM"),
    ("smsp__average_warp_latency_issue_stalled_wait_per_warp_active.pct", "stall: wait (fixed latency)"),
    ("smsp__average_warp_latency_issue_stalled_long_scoreboard_per_warp_active.pct", "stall: long scoreboard (global/local loads)"),
    ("smsp__average_warp_latency_issue_stalled_short_scoreboard_per_warp_active.pct", "stall: short scoreboard (MUFU/shared)"),
    ("smsp__average_warp_latency_issue_stalled_math_pipe_throttle_per_warp_active.pct", "stall: math pipe throttle"),
])


def full(src: str, dst: str, kernel_filter: str = "") -> None:
    raw = subprocess.run(["ncu", "-i", src, "--page", "raw", "--csv"], capture_output=True, text=True, check=True).stdout
    rows = list(csv.reader(raw.splitlines()))
    hdr, units = rows[0], rows[1]
    out = []
    for vals in rows[2:]:
        rec = dict(zip(hdr, vals))
        if kernel_filter and kernel_filter not in rec.get("Kernel Name", ""):
            continue
        out.append(rec)
    unit_of = dict(zip(hdr, units))
    with open(dst, "w") as fh:
        fh.write(f"# ncu --set full summary - {src}\n\n")
        fh.write("`ncu --set full --clock-control none --import-source on` (one GPU, profiled after the same command "
                 "exited 0 without ncu).  Values are per launch.\n\n")
        for rec in out:
            fh.write(f"## `{short(rec.get('Kernel Name', '?'))}` (launch id {rec.get('ID')})\n\n| metric | value | unit |\n|---|---:|---|\n")
            for key, label in WANTED.items():
                if key in rec and rec[key] != "":
                    fh.write(f"| {label} (`{key}`) | {rec[key]} | {unit_of.get(key, '')} |\n")
            rd, wr = rec.get("dram__bytes_read.sum"), rec.get("dram__bytes_write.sum")
            if rd and wr:
                # ncu prints these in the unit of the units row
                fh.write(f"| DRAM traffic read + write | {float(rd) + float(wr):.6g} | {unit_of.get('dram__bytes_read.sum', '')} |\n")
            fh.write("\n")


if __name__ == "__main__":
    mode = sys.argv[1]
    if mode == "launches":
        launches(sys.argv[2], sys.argv[3])
    elif mode == "full":
        full(sys.argv[2], sys.argv[3], sys.argv[4] if len(sys.argv) > 4 else "")
    else:
        raise SystemExit(__doc__)
