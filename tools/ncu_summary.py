#!/usr/bin/env python3
"""Summarise an .ncu-rep (ncu --set full) and a launch list CSV into a small markdown file for profiles/.

    python tools/ncu_summary.py gpurun_out/prof.ncu-rep gpurun_out/launches.csv profiles/r01_summary.md "title"
"""
import collections
import csv
import io
import subprocess
import sys

rep, launches, out, title = sys.argv[1], sys.argv[2], sys.argv[3], (sys.argv[4] if len(sys.argv) > 4 else "ncu summary")
METRICS = [
    ("gpu__time_duration.sum", "duration"),
    ("launch__registers_per_thread", "regs/thread"),
    ("launch__occupancy_limit_registers", "occupancy limit (regs), blocks/SM"),
    ("sm__warps_active.avg.pct_of_peak_sustained_active", "achieved occupancy %"),
    ("sm__inst_executed.avg.per_cycle_elapsed", "warp inst / cycle / SM (max 4)"),
    ("smsp__thread_inst_executed_per_inst_executed.ratio", "active threads / warp inst"),
    ("sm__throughput.avg.pct_of_peak_sustained_elapsed", "SM throughput %"),
    ("gpu__dram_throughput.avg.pct_of_peak_sustained_elapsed", "DRAM throughput %"),
    ("dram__bytes_read.sum", "DRAM read"),
    ("dram__bytes_write.sum", "DRAM write"),
    ("l1tex__t_sector_hit_rate.pct", "L1 hit %"),
    ("lts__t_sector_hit_rate.pct", "L2 hit %"),
    ("sm__inst_executed_pipe_fp64.avg.pct_of_peak_sustained_active", "FP64 pipe %"),
    ("sm__inst_executed_pipe_lsu.avg.pct_of_peak_sustained_active", "LSU pipe %"),
    ("smsp__inst_executed.sum", "warp instructions"),
]
raw = subprocess.run(["ncu", "-i", rep, "--page", "raw", "--csv"], capture_output=True, text=True).stdout
rows = list(csv.reader(io.StringIO(raw)))
h, units = rows[0], rows[1]
lines = [f"# {title}", "", f"Source: `{rep}` (ncu --set full --clock-control none), launch list `{launches}`.", ""]
lines += ["## Launch list (gpu__time_duration.sum; cold-cache, serialised: compare shares)", "", "| kernel | launches | mean us | share |", "|---|---|---|---|"]
agg = collections.OrderedDict()
if launches != "-":
    lr = list(csv.reader(open(launches)))
    hi = [i for i, r in enumerate(lr) if r and r[0] == "ID"][0]
    lh = lr[hi]
    for r in lr[hi + 1:]:
        if len(r) > lh.index("Metric Value"):
            agg.setdefault(r[lh.index("Kernel Name")].split("(")[0], []).append(float(r[lh.index("Metric Value")].replace(",", "")))
else:       # no separate launch list: the captured launches themselves
    for r in rows[2:]:
        agg.setdefault(r[h.index("Kernel Name")].split("(")[0], []).append(
            float(r[h.index("gpu__time_duration.sum")].replace(",", "")) * {"us": 1e3, "ms": 1e6, "ns": 1.0}.get(units[h.index("gpu__time_duration.sum")], 1.0))
tot = sum(sum(v) for v in agg.values())
for k, v in agg.items():
    lines.append(f"| `{k}` | {len(v)} | {sum(v) / len(v) / 1e3:.1f} | {100 * sum(v) / tot:.1f}% |")
lines += ["", "## Per-kernel metrics (one captured launch each)", ""]
seen = set()
for r in rows[2:]:
    name = r[h.index("Kernel Name")].split("(")[0]
    if name in seen:
        continue
    seen.add(name)
    lines += [f"### `{name}`  grid {r[h.index('Grid Size')]} block {r[h.index('Block Size')]}", "", "| metric | value |", "|---|---|"]
    for m, label in METRICS:
        if m in h:
            lines.append(f"| {label} (`{m}`) | {r[h.index(m)]} {units[h.index(m)]} |")
    lines.append("")
open(out, "w").write("\n".join(lines) + "\n")
print("wrote", out)
