#!/usr/bin/env python3
"""Per-source-line hot spots of one kernel from an .ncu-rep captured with --import-source on (-lineinfo build):

    python tools/ncu_hot.py gpurun_out/prof.ncu-rep initial_kernel [top]

Aggregates the SASS rows of `ncu --page source --csv --print-source sass,cuda` by CUDA-C source line: warp instructions
executed and stall samples, so that the share of every source construct in the kernel's issue slots can be read off."""
import collections, csv, io, subprocess, sys

rep, kern = sys.argv[1], sys.argv[2]
top = int(sys.argv[3]) if len(sys.argv) > 3 else 25
raw = subprocess.run(["ncu", "-i", rep, "--page", "source", "--csv", "--print-source", "sass,cuda", "--kernel-name", "regex:" + kern],
                     capture_output=True, text=True).stdout
rows = list(csv.reader(io.StringIO(raw)))
hi = [i for i, r in enumerate(rows) if r and r[0] == "Line No"][0]
h = rows[hi]
iL, iS, iI, iN = 0, 1, h.index("Instructions Executed"), h.index("# Samples")
iT = h.index("Thread Instructions Executed")
agg = collections.OrderedDict()
sass = [0, 0, 0]
cur_file = ""
for r in rows[hi + 1:]:
    if len(r) < len(h):
        if r and r[0] in ("File Name", "File Path"): cur_file = r[1].split("/")[-1]
        continue
    try:
        inst, smp, thr = int(r[iI] or 0), int(r[iN] or 0), int(r[iT] or 0)
    except ValueError:
        continue
    if not r[iL]:                   # a SASS row: counts towards the kernel totals only
        sass[0] += inst; sass[1] += smp; sass[2] += thr
        continue
    key = (r[iL], r[iS].strip()[:110])
    a = agg.setdefault(key, [0, 0, 0, 0]); a[0] += inst; a[1] += smp; a[2] += thr; a[3] += 1
ti, ts, tt = sass
print(f"kernel {kern}: warp inst {ti}, samples {ts}, active threads / warp inst {tt / max(ti, 1):.2f}, source lines {len(agg)}")
print("inst%  samp%  thr/inst  sass  line  source")
for (ln, src), a in sorted(agg.items(), key=lambda kv: -kv[1][0])[:top]:
    print(f"{100 * a[0] / ti:5.1f} {100 * a[1] / max(ts, 1):6.1f} {a[2] / max(a[0], 1):8.1f} {a[3]:5d} {ln:>5s}  {src}")
