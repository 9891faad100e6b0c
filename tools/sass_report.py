#!/usr/bin/env python3
"""Per-kernel SASS evidence from the shipped library -> a markdown table for profiles/.

    python tools/sass_report.py romis_b200/libromis_gpu.so profiles/r02_sass.md

Registers / stack / spill / shared memory come from `cuobjdump -res-usage`, instruction counts from `cuobjdump -sass`
(static counts of the emitted code, not executed counts): shared-memory traffic (LDS/STS), warp shuffles / votes,
local-memory traffic (LDL/STL), binary64 arithmetic (DMUL/DADD/DFMA), special-function unit (MUFU), global loads/stores."""
import collections
import re
import subprocess
import sys

lib, out = sys.argv[1], sys.argv[2]
only = re.compile(sys.argv[3]) if len(sys.argv) > 3 else None
GROUPS = [("LDS", r"^LDS"), ("STS", r"^STS"), ("SHFL", r"^SHFL"), ("VOTE/MATCH", r"^(VOTE|MATCH|WARPSYNC)"), ("LDL", r"^LDL"), ("STL", r"^STL"),
          ("LDG", r"^LDG"), ("STG", r"^STG"), ("DMUL/DADD/DFMA", r"^(DMUL|DADD|DFMA)"), ("MUFU", r"^MUFU"), ("FFMA", r"^FFMA"),
          ("FMUL/FADD", r"^(FMUL|FADD)"), ("IMAD/LOP3/SHF", r"^(IMAD|LOP3|SHF|IADD3)"), ("BSSY/BSYNC", r"^(BSSY|BSYNC)"), ("CALL", r"^CALL"),
          ("UTMALDG/UTMASTG", r"^UTMA"), ("MEMBAR.SC", r"^MEMBAR\.SC"), ("MEMBAR.ALL", r"^MEMBAR\.ALL"), ("CCTL.IVALL", r"^CCTL\.IVALL"),
          ("total", r"^[A-Z]")]


def demangle(names):
    r = subprocess.run(["c++filt"], input="\n".join(names), capture_output=True, text=True).stdout.split("\n")
    return dict(zip(names, r))


res = subprocess.run(["cuobjdump", "-res-usage", lib], capture_output=True, text=True).stdout
usage = {}
cur = None
for line in res.splitlines():
    m = re.match(r"\s*Function (\S+):", line)
    if m:
        cur = m.group(1); continue
    if cur and "REG:" in line:
        usage[cur] = {k: int(v) for k, v in re.findall(r"(REG|STACK|SHARED|LOCAL|CONSTANT\[0\]):(\d+)", line)}
        cur = None
sass = subprocess.run(["cuobjdump", "-sass", lib], capture_output=True, text=True).stdout
counts = collections.OrderedDict()
cur = None
for line in sass.splitlines():
    m = re.match(r"\s*Function : (\S+)", line)
    if m:
        cur = m.group(1); counts[cur] = collections.Counter(); continue
    m = re.match(r"\s*/\*[0-9a-f]{4,}\*/\s+(?:@!?U?P\d+\s+)?([A-Z0-9_.]+)", line)
    if m and cur:
        op = m.group(1)
        for g, pat in GROUPS:
            if re.match(pat, op):
                counts[cur][g] += 1
names = demangle(list(counts))


def short(n):
    d = names.get(n, n)
    d = re.sub(r"\(.*", "", d).replace("void ", "").replace("romis::", "")
    return d


rows = []
for n, c in counts.items():
    s = short(n)
    if only and not only.search(s):
        continue
    u = usage.get(n, {})
    rows.append((s, u.get("REG", 0), u.get("STACK", 0), u.get("LOCAL", 0), u.get("SHARED", 0), c))
rows.sort(key=lambda r: r[0])
hdr = ["kernel", "regs", "stack B", "local B", "static smem B"] + [g for g, _ in GROUPS]
lines = ["# SASS / resource listing of `%s`" % lib, "",
         "Static instruction counts from `cuobjdump -sass` (sm_100a cubins), resources from `cuobjdump -res-usage`; produced by",
         "`tools/sass_report.py`.  stack B = per-thread local memory the compiler reserved (traversal stacks, generic-N arrays, spills).",
         "MEMBAR.SC / MEMBAR.ALL / CCTL.IVALL: what the fences of the row-group counters and of the halo tokens compile to -- `__threadfence()`",
         "is MEMBAR.SC *and* CCTL.IVALL (the SM's whole L1 invalidated); `fence.release` is MEMBAR.ALL alone, `fence.acquire` CCTL.IVALL alone",
         "(DESIGN.md 4, 6).  The MEMBAR.SC left are the once-per-edge-and-pass token publishers and the stand-alone flag kernels.", "",
         "| " + " | ".join(hdr) + " |", "|" + "---|" * len(hdr)]
for s, reg, st, loc, sh, c in rows:
    lines.append("| `%s` | %d | %d | %d | %d | " % (s, reg, st, loc, sh) + " | ".join(str(c[g]) for g, _ in GROUPS) + " |")
open(out, "w").write("\n".join(lines) + "\n")
print("wrote", out, len(rows), "kernels")
