#!/usr/bin/env python3
"""Device time of the C2 frame (nightclub 1080p, M=32, temporal + 3 spatial, visibility reuse) WITHOUT per-stage events between
the kernels (they would serialise launches that programmatic dependent launch overlaps); optionally one row band only.

    python tools/frame_time.py [y0 y1]        # frames back to back, mean / min of romis_timings.total_ms over 40 frames
"""
import os, sys
import numpy as np
sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
from romis_b200.api import RestirRenderer
from romis_b200.scene import Camera, Features, Scene

W, H = 1920, 1080
scene = Scene.load(os.path.join(os.path.dirname(__file__), "..", "tests", "golden", "scenes", "CornellNightClub.npz"))
r = RestirRenderer(0); r.upload_scene(scene)
if len(sys.argv) > 2: r.set_band(int(sys.argv[1]), int(sys.argv[2]))
feat = Features(spatialResamplingPasses=3, initialSamplesVisibilityCheck=True)
cam = Camera()
ts = []
for fr in range(48):
    r.render_frame(feat, cam, W, H, fr > 0, 1, fr, want_image=False)
    t = r.timings()
    if fr >= 8: ts.append(t.total_ms)
print(f"band {sys.argv[1:3] or 'full'}: frame mean {np.mean(ts):.4f} ms  min {np.min(ts):.4f}  max {np.max(ts):.4f}  launches {t.n_launches}")
