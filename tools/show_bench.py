#!/usr/bin/env python3
"""One-line digest of a bench.py JSON line: python tools/show_bench.py gpurun_out/bench_x.json"""
import json, sys
for f in sys.argv[1:]:
    try:
        d = json.loads(open(f).read().strip().splitlines()[-1])
    except Exception as e:
        print(f, "unreadable:", e); continue
    p = d.get("roofline", {}).get("passes", {})
    print(f, "fps", round(d["value"], 1), "ms", round(d["ms_per_step"], 4), "e2e", round(d["e2e"]["value"], 1),
          "static", round(d["e2e"].get("static_lights", {}).get("value", 0), 1), {k: v["ms_per_launch"] for k, v in p.items()},
          "exch", d.get("exchange_share_of_frame"), "clk", (d.get("clocks") or {}).get("sm_mhz"), (d.get("clocks") or {}).get("samples"))
