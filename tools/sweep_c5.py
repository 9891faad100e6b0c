#!/usr/bin/env python3
"""BASELINE C5: spatial-reuse sweep k in {3,5,10} x radius in {10,30} x passes 1..4 at 3840x2160 on a synthetic 2^20-light
scene (1 GPU here; device time per frame and per spatial pass).  Writes a markdown table to stdout."""
import os, sys
import numpy as np
sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
from romis_b200.api import RestirRenderer
from romis_b200.scene import Camera, Features, Scene, synthetic_lights

L = int(sys.argv[1]) if len(sys.argv) > 1 else 1 << 20
W, H = 3840, 2160
scene = Scene.load(os.path.join(os.path.dirname(__file__), "..", "tests", "golden", "scenes", "CornellNightClub.npz"))
scene.lights = synthetic_lights(L, seed=1, intensity=4000.0)
scene.lights["p0"] += np.array([2.5, 2.0, -1.0], np.float32)
r = RestirRenderer(0); r.upload_scene(scene); r.set_stage_timing(True)
cam = Camera()
print(f"| k | radius | passes | frame ms | initial ms | spatial ms / pass | frames/s | G-candidates/s |\n|---|---|---|---|---|---|---|---|")
for k in (3, 5, 10):
    for rad in (10, 30):
        for P in (1, 2, 3, 4):
            feat = Features(numNeighboursToSample=k, spatialResampleRadius=rad, spatialResamplingPasses=P, initialSamplesVisibilityCheck=True)
            tot = []; ini = []; sp = []
            for fr in range(5):
                r.render_frame(feat, cam, W, H, fr > 0, 7, fr, want_image=False)
                t = r.timings()
                if fr >= 2:
                    tot.append(t.total_ms); ini.append(t.initial_ms); sp.append(sum(t.spatial_ms[:P]) / P)
            ms = float(np.mean(tot))
            print(f"| {k} | {rad} | {P} | {ms:.2f} | {np.mean(ini):.2f} | {np.mean(sp):.2f} | {1e3 / ms:.1f} | {W * H * 32 / ms / 1e6:.1f} |", flush=True)
