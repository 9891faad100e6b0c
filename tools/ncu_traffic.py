#!/usr/bin/env python3
"""DRAM traffic, duration and issue rate per kernel from an .ncu-rep (ncu --set full) -> the JSON bench.py reads for
roofline.traffic:   python tools/ncu_traffic.py gpurun_out/prof.ncu-rep profiles/r01d_traffic.json c2 "source note" """
import csv, io, json, subprocess, sys

rep, out, config, note = sys.argv[1], sys.argv[2], sys.argv[3], sys.argv[4]
raw = subprocess.run(["ncu", "-i", rep, "--page", "raw", "--csv"], capture_output=True, text=True).stdout
rows = list(csv.reader(io.StringIO(raw)))
h, units = rows[0], rows[1]


def val(r, name, unit_scale=True):
    v = float(r[h.index(name)].replace(",", ""))
    u = units[h.index(name)]
    if unit_scale:
        v *= {"Kbyte": 1e3, "Mbyte": 1e6, "Gbyte": 1e9, "byte": 1.0, "us": 1e-3, "ms": 1.0, "ns": 1e-6, "s": 1e3}.get(u, 1.0)
    return v


kernels = {}
for r in rows[2:]:
    name = r[h.index("Kernel Name")].split("(")[0].replace("void ", "").split("<")[0].replace("romis::", "")
    if name in kernels:
        continue
    kernels[name] = {"dram_bytes_read": val(r, "dram__bytes_read.sum"), "dram_bytes_write": val(r, "dram__bytes_write.sum"),
                     "duration_ms": val(r, "gpu__time_duration.sum"),
                     "warp_inst_per_cycle_per_sm": val(r, "sm__inst_executed.avg.per_cycle_elapsed", False),
                     "thread_inst_executed": val(r, "smsp__thread_inst_executed_per_inst_executed.ratio", False) * val(r, "smsp__inst_executed.sum", False),
                     "warp_inst_executed": val(r, "smsp__inst_executed.sum", False),
                     "registers_per_thread": val(r, "launch__registers_per_thread", False),
                     "achieved_occupancy_pct": val(r, "sm__warps_active.avg.pct_of_peak_sustained_active", False)}
json.dump({"source": note, "config": config, "kernels": kernels}, open(out, "w"), indent=1)
print(json.dumps(kernels, indent=1))
