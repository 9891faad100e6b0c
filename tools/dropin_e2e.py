#!/usr/bin/env python3
"""End-to-end frames/s of the drop-in itself: integration/render_restir_gpu.cpp driven through the reference's own Scene /
Trackball / Screen objects (oracle/_ref/libromis_dropin.so, ref_dropin_bench), wall clock around every renderReSTIR call --
scene hash, light hand-over, frame, image read-back into Screen::pixels().  C2 workload.

    [ROMIS_DEVICES=0,1] python tools/dropin_e2e.py [frames]

Prints one JSON object: page-locked vs pageable Screen, static lights vs one light edited per frame."""
import ctypes as C, json, os, sys
import numpy as np
sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
from oracle import pyoracle
from romis_b200.scene import Camera, Features, Scene

frames = int(sys.argv[1]) if len(sys.argv) > 1 else 60
lib = pyoracle.DropinLib()
scene = Scene.load(os.path.join(os.path.dirname(os.path.abspath(__file__)), "..", "tests", "golden", "scenes", "CornellNightClub.npz"))
lib.set_scene(scene)
feat = Features(spatialResamplingPasses=3, initialSamplesVisibilityCheck=True).to_abi()
cam = lib._cam(Camera())
W, H = 1920, 1080
lib.lib.ref_dropin_bench.argtypes = [C.c_void_p, C.c_void_p, C.c_int, C.c_int, C.c_int, C.c_int, C.c_int, C.c_void_p]
out = {"workload": "C2 cornell-nightclub 1920x1080 through renderReSTIR_gpu (drop-in body) on the reference's Scene / Trackball / Screen",
       "devices": os.environ.get("ROMIS_DEVICES", "0"), "frames": frames}
for name, pin, edit in (("pinned_screen", 1, 0), ("pageable_screen", 0, 0), ("pinned_screen_light_edit", 1, 1)):
    ms = np.zeros(frames, np.float64)
    rc = lib.lib.ref_dropin_bench(C.byref(feat), C.byref(cam), W, H, frames, pin, edit, ms.ctypes.data)
    if rc != 0:
        raise SystemExit(lib.lib.ref_last_error().decode())
    steady = ms[5:]
    out[name] = {"frames_per_s": round(1e3 / float(np.mean(steady)), 1), "ms_per_frame_mean": round(float(np.mean(steady)), 3),
                 "ms_per_frame_median": round(float(np.median(steady)), 3)}
print(json.dumps(out))
