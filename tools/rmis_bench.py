#!/usr/bin/env python3
"""R-MIS frame time on one GPU (romis_render_frame_rmis, image left on the device): C2 scene at 1920x1080, M=32, N=2, k=5,
r=10, the reference's default 5 iterations, per neighbour-selection strategy and MIS weight."""
import json
import os
import sys
import time

sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
sys.path.insert(0, os.path.join(os.path.dirname(os.path.dirname(os.path.abspath(__file__))), "tests"))

from romis_b200 import abi  # noqa: E402
from romis_b200.api import RestirRenderer  # noqa: E402
from romis_b200.scene import Camera, Features, RmisParams  # noqa: E402
from common import load_scene  # noqa: E402


def main():
    W, H = 1920, 1080
    r = RestirRenderer(0)
    r.upload_scene(load_scene("CornellNightClub"))
    cam = Camera().to_abi(W, H)
    feat = Features(initialSamplesVisibilityCheck=True)
    rows = []
    for strat, sname in ((abi.ROMIS_NEIGHBOURS_RANDOM, "random"), (abi.ROMIS_NEIGHBOURS_SIMILAR, "similar"),
                         (abi.ROMIS_NEIGHBOURS_EQUAL_SIMILAR_DISSIMILAR, "equal-similar-dissimilar")):
        for mis, mname in ((abi.ROMIS_MIS_EQUAL, "equal"), (abi.ROMIS_MIS_BALANCE, "balance")):
            rp = RmisParams(maxIterationsMIS=5, misWeightRMIS=mis, neighbourSelectionStrategy=strat)
            for i in range(2):
                r.render_frame_rmis(feat, rp, cam, W, H, 1, i, want_image=False)
            ts = []
            for i in range(5):
                r.render_frame_rmis(feat, rp, cam, W, H, 1, 2 + i, want_image=False)
                ts.append(r.timings().total_ms)
            ts.sort()
            rows.append({"strategy": sname, "mis": mname, "ms_per_frame": ts[len(ts) // 2], "iterations": 5})
            print(rows[-1], flush=True)
    out = os.path.join("gpurun_out", "rmis_bench.json")
    os.makedirs("gpurun_out", exist_ok=True)
    json.dump(rows, open(out, "w"), indent=1)


if __name__ == "__main__":
    main()
