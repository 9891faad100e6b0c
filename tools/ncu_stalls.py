#!/usr/bin/env python3
"""Per-source-line stall samples of one kernel, by stall reason, from an .ncu-rep captured with --import-source on:

    python tools/ncu_stalls.py gpurun_out/prof.ncu-rep spatial_kernel [reason=stall_long_sb] [top]

Companion of tools/ncu_hot.py (which ranks lines by instructions executed)."""
import collections, csv, io, subprocess, sys

rep, kern = sys.argv[1], sys.argv[2]
reason = sys.argv[3] if len(sys.argv) > 3 else "stall_long_sb"
top = int(sys.argv[4]) if len(sys.argv) > 4 else 25
raw = subprocess.run(["ncu", "-i", rep, "--page", "source", "--csv", "--print-source", "sass,cuda", "--kernel-name", "regex:" + kern],
                     capture_output=True, text=True).stdout
rows = list(csv.reader(io.StringIO(raw)))
hi = [i for i, r in enumerate(rows) if r and r[0] == "Line No"]
h = rows[hi[0]]
iR, iS, iI = h.index(reason), h.index("# Samples"), h.index("Instructions Executed")


def num(s):
    try:
        return int(s)
    except ValueError:
        return 0


agg = collections.OrderedDict()
cur_file = ""
for r in rows[hi[0] + 1:]:          # one block per source file and captured launch: summed
    if len(r) < len(h) or not r[0] or r[0] == "Line No":
        if r and r[0] in ("File Name", "File Path"): cur_file = r[1].split("/")[-1]
        continue
    a = agg.setdefault((r[0], cur_file[:18] + ": " + r[1].strip()[:100]), [0, 0, 0])
    a[0] += num(r[iR]); a[1] += num(r[iS]); a[2] += num(r[iI])
tot = sum(a[0] for a in agg.values()); ts = sum(a[1] for a in agg.values())
print(f"kernel {kern}: {reason} samples {tot} of {ts} ({100 * tot / max(ts, 1):.1f} %)")
print(f"{reason}%  samples  inst  line  source")
for (ln, src), a in sorted(agg.items(), key=lambda kv: -kv[1][0])[:top]:
    print(f"{100 * a[0] / max(tot, 1):6.1f} {a[1]:7d} {a[2]:10d} {ln:>5s}  {src}")
