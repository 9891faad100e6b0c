#!/usr/bin/env python3
"""How much of the multi-GPU gap is the per-kernel tail of a thin band?  On ONE GPU: per-stage device time of the full C2
frame against the SUM over the G equal row bands rendered one after another (no exchange, stale halos: only the
timing matters).  ROMIS_BLOCK_Y selects the block shape."""
import os, sys
import numpy as np
sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
from romis_b200.api import RestirRenderer
from romis_b200.scene import Camera, Features, Scene

G = int(sys.argv[1]) if len(sys.argv) > 1 else 8
W, H = 1920, 1080
scene = Scene.load(os.path.join(os.path.dirname(__file__), "..", "tests", "golden", "scenes", "CornellNightClub.npz"))
feat = Features(spatialResamplingPasses=3, initialSamplesVisibilityCheck=True); cam = Camera()
r = RestirRenderer(0); r.upload_scene(scene); r.set_stage_timing(os.environ.get("STAGE_TIMING", "0") == "1")


def run(y0, y1):
    r.set_band(y0, y1)
    acc = []
    for fr in range(10):
        r.frame_begin(feat, cam, W, H, fr > 0, 1, fr)
        for p in range(3):
            r.frame_spatial_pass(p)
        r.frame_end(None); r.synchronize()
        t = r.timings()
        if fr >= 3:
            acc.append([t.total_ms, t.primary_ms, t.initial_ms, t.temporal_ms, sum(t.spatial_ms[:3]), t.shade_ms])
    return np.median(np.array(acc), axis=0)


full = run(0, 0)
names = ["total", "primary", "initial", "temporal", "spatial x3", "shade"]
bands = np.array([run(g * H // G, (g + 1) * H // G) for g in range(G)])
print(f"block_y={os.environ.get('ROMIS_BLOCK_Y', '8')} strips={os.environ.get('ROMIS_STRIPS', 'auto')} stage_timing={os.environ.get('STAGE_TIMING', '0')}  G={G}")
print("stage       full_ms  sum_bands_ms  ratio   max_band_ms  ideal(full/G)")
for i, n in enumerate(names):
    if full[i] <= 0:
        continue                    # per-stage times exist only with STAGE_TIMING=1 (which also forces the in-order sequence)
    print(f"{n:10s} {full[i]:8.3f} {bands[:, i].sum():10.3f} {bands[:, i].sum() / full[i]:8.2f} {bands[:, i].max():10.3f} {full[i] / G:10.3f}")
print("per-band totals:", np.round(bands[:, 0], 3))
