#!/usr/bin/env python3
"""Generates tests/golden/ from the REFERENCE itself (run in the dev container, where /root/reference exists):

    make -C oracle ref && python tests/golden/gen_golden.py

* scenes/<SceneType>.npz -- what the reference's own loader produces for its prebuilt scenes
  (loadScenePrebuilt, reference src/scene/scene.cpp:68-132): meshes after OBJ parsing / normalisation,
  materials, textures, lights.  These are the inputs the C-ABI takes; they travel to the GPU box.
* <case>.npz -- per frame: romis_camera as the reference's Trackball computed it, G-buffer, reservoir
  state after every stage (position, colour, W, M, wSum), final float image, all produced by the
  reference's translation units (oracle/_ref/libromis_ref.so) with the injected counter-based RNG.
"""
import os
import sys

import numpy as np

ROOT = os.path.dirname(os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
sys.path.insert(0, ROOT)
sys.path.insert(0, os.path.dirname(os.path.abspath(__file__)))

from oracle.pyoracle import RefLib, REF_FLAG_SPLIT_SPATIAL, build_ref  # noqa: E402
from romis_b200 import abi  # noqa: E402
from cases import CASES, RMIS_CASES, ROMIS_CASES, SCENES  # noqa: E402

OUT = os.path.dirname(os.path.abspath(__file__))


def cam_array(c: abi.romis_camera) -> np.ndarray:
    return np.array([*c.origin, *c.quat, c.half_width, c.half_height], np.float32)


def main():
    build_ref()
    ref = RefLib()
    os.makedirs(os.path.join(OUT, "scenes"), exist_ok=True)
    for s in SCENES:
        ref.load_prebuilt(s)
        ref.export_scene(s).save(os.path.join(OUT, "scenes", s + ".npz"))
    if "--romis-only" in sys.argv:
        return romis_cases(ref)
    only_rmis = "--rmis-only" in sys.argv     # the ReSTIR vectors are left as committed
    for name, (scene, W, H, feat, cam, frames, seed) in ({} if only_rmis else CASES).items():
        ref.load_prebuilt(scene)
        d = {"camera": cam_array(ref.make_camera(cam, W, H))}
        for fr in range(frames):
            rf = ref.render_frame(feat, cam, W, H, fr > 0, seed, fr, REF_FLAG_SPLIT_SPATIAL)
            p = f"f{fr}_"
            d[p + "t"] = rf.gbuffer.t; d[p + "normal"] = rf.gbuffer.normal
            d[p + "mesh"] = rf.gbuffer.mesh; d[p + "texcoord"] = rf.gbuffer.texcoord
            d[p + "image"] = rf.image
            for pid, st in rf.stages.items():
                if pid == abi.ROMIS_PASS_TEMPORAL and (fr == 0 or not feat.temporalReuse):
                    continue
                if pid >= abi.ROMIS_PASS_SPATIAL0 and pid != abi.ROMIS_PASS_FINAL and not feat.spatialReuse:
                    continue
                for fld in ("position", "color", "W", "M", "wSum"):
                    d[f"{p}s{pid}_{fld}"] = getattr(st, fld)
        np.savez_compressed(os.path.join(OUT, name + ".npz"), **d)
        print(name, "ok")
    for name, (scene, W, H, feat, rmis, cam, seed, frame) in RMIS_CASES.items():
        ref.load_prebuilt(scene)
        img, xy, cnt = ref.render_frame_rmis(feat, rmis, cam, W, H, seed, frame)
        np.savez_compressed(os.path.join(OUT, name + ".npz"), camera=cam_array(ref.make_camera(cam, W, H)), image=img,
                            neighbours=xy.astype(np.int16), count=cnt.astype(np.uint8))
        print(name, "ok")
    romis_cases(ref)


def romis_cases(ref):
    for name, (scene, W, H, feat, rmis, cam, seed, frame) in ROMIS_CASES.items():
        ref.load_prebuilt(scene)
        img, A, B = ref.render_frame_romis(feat, rmis, cam, W, H, seed, frame)
        np.savez_compressed(os.path.join(OUT, name + ".npz"), camera=cam_array(ref.make_camera(cam, W, H)), image=img, matrices=A, contributions=B)
        print(name, "ok")


if __name__ == "__main__":
    main()
