#!/usr/bin/env python3
"""Per-stage device timings of the C2 configuration (nightclub 1080p, M=32, temporal + 3 spatial, visibility reuse)."""
import os, sys, time
import numpy as np
sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
from romis_b200.api import RestirRenderer
from romis_b200.scene import Camera, Features, Scene

W, H = (int(sys.argv[1]), int(sys.argv[2])) if len(sys.argv) > 2 else (1920, 1080)
N = int(sys.argv[3]) if len(sys.argv) > 3 else 2
scene = Scene.load(os.path.join(os.path.dirname(__file__), "..", "tests", "golden", "scenes", "CornellNightClub.npz"))
r = RestirRenderer(0); r.upload_scene(scene); r.set_stage_timing(True)
feat = Features(spatialResamplingPasses=3, initialSamplesVisibilityCheck=True, numSamplesInReservoir=N)
cam = Camera()
for fr in range(6):
    t0 = time.time()
    r.render_frame(feat, cam, W, H, fr > 0, 1, fr, want_image=False)
    t = r.timings()
    print(f"frame {fr}: total {t.total_ms:.3f} ms  primary {t.primary_ms:.3f} initial {t.initial_ms:.3f} temporal {t.temporal_ms:.3f} "
          f"spatial {[round(x,3) for x in t.spatial_ms[:t.n_spatial]]} shade {t.shade_ms:.3f}  wall {1e3*(time.time()-t0):.2f} ms launches {t.n_launches}")
