#!/usr/bin/env python3
"""Per-stage time against the number of rows of a band centred in the C2 frame (one GPU, in-order sequence):
time = a * rows + b; b is the fixed cost (tail + launch) a thin band pays per kernel."""
import os, sys
import numpy as np
sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
from romis_b200.api import RestirRenderer
from romis_b200.scene import Camera, Features, Scene

W, H = 1920, 1080
M = int(os.environ.get("M", "32"))
scene = Scene.load(os.path.join(os.path.dirname(__file__), "..", "tests", "golden", "scenes", "CornellNightClub.npz"))
feat = Features(spatialResamplingPasses=3, initialSamplesVisibilityCheck=True, initialLightSamples=M); cam = Camera()
r = RestirRenderer(0); r.upload_scene(scene); r.set_stage_timing(True)
print(f"M={M} block_y={os.environ.get('ROMIS_BLOCK_Y', '8')} split={os.environ.get('ROMIS_INIT_SPLIT', 'auto')}")
print("rows  primary initial temporal spatial(avg) shade   | per-135-rows: initial spatial")
for rows in (34, 68, 135, 270, 540, 1080):
    y0 = (H - rows) // 2
    r.set_band(y0, y0 + rows) if rows < H else r.set_band(0, 0)
    acc = []
    for fr in range(9):
        r.frame_begin(feat, cam, W, H, fr > 0, 1, fr)
        for p in range(3):
            r.frame_spatial_pass(p)
        r.frame_end(None); r.synchronize()
        t = r.timings()
        if fr >= 3:
            acc.append([t.primary_ms, t.initial_ms, t.temporal_ms, sum(t.spatial_ms[:3]) / 3, t.shade_ms])
    m = np.median(np.array(acc), axis=0) * 1e3
    print(f"{rows:5d} {m[0]:7.1f} {m[1]:7.1f} {m[2]:7.1f} {m[3]:9.1f} {m[4]:8.1f}   | {m[1] * 135 / rows:8.1f} {m[3] * 135 / rows:8.1f}")
