"""Live comparison of the C restatement with the compiled reference (oracle/_ref/libromis_ref.so), beyond the
committed golden vectors: other sizes, seeds, the reference's own renderReSTIR entry point, both tracer modes."""
import os

import numpy as np
import pytest

from oracle import pyoracle
from oracle.pyoracle import REF_FLAG_SPLIT_SPATIAL, REF_FLAG_WHOLE_FRAME
from romis_b200 import abi
from romis_b200.scene import Features, synthetic_lights
from cases import CORNELL_CAM, NIGHTCLUB_CAM
from common import assert_bits_equal, load_scene

pytestmark = pytest.mark.skipif(not os.path.exists(pyoracle.REF_SO), reason="oracle/_ref not built (needs /root/reference: make -C oracle ref)")


@pytest.fixture(scope="module")
def ref():
    return pyoracle.RefLib()


@pytest.mark.parametrize("scene_name,cam,feat,W,H,seed", [
    ("CornellNightClub", NIGHTCLUB_CAM, Features(spatialResamplingPasses=3, initialSamplesVisibilityCheck=True), 72, 41, 1001),
    ("CornellNightClub", NIGHTCLUB_CAM, Features(unbiasedCombination=True, spatialReuseVisibilityCheck=True, spatialResampleRadius=3), 40, 32, 1002),
    ("Monkey", CORNELL_CAM, Features(numSamplesInReservoir=1, temporalClampM=1), 56, 48, 1003),
])
def test_restatement_equals_compiled_reference(ref, oracle_factory, scene_name, cam, feat, W, H, seed):
    scene = load_scene(scene_name)
    ref.set_scene(scene)                         # the reference's Scene rebuilt from the fixture arrays
    orc = oracle_factory(); orc.upload_scene(scene); orc.reset_history(); ref.reset_history()
    rcam = ref.make_camera(cam, W, H)
    for fr in range(3):
        rf = ref.render_frame(feat, cam, W, H, fr > 0, seed, fr, REF_FLAG_SPLIT_SPATIAL)
        img = orc.render_frame(feat, rcam, W, H, fr > 0, seed, fr)
        assert_bits_equal(orc.ray_dirs(rcam, W, H), rf.ray_dir, "camera rays (Trackball::generateRay)")
        g = orc.gbuffer()
        assert_bits_equal(g.t, rf.gbuffer.t, "t"); assert_bits_equal(g.normal, rf.gbuffer.normal, "normal"); assert_bits_equal(g.mesh, rf.gbuffer.mesh, "mesh")
        for pid, st in rf.stages.items():
            if pid == abi.ROMIS_PASS_TEMPORAL and fr == 0:
                continue
            o = orc.reservoirs(pid)
            for fld in ("position", "color", "W", "M", "wSum", "chosenW"):
                assert_bits_equal(getattr(o, fld), getattr(st, fld), f"frame {fr} stage {pid} {fld}")
        assert_bits_equal(img, rf.image, f"frame {fr} image")


def test_reference_entry_point_equals_stagewise_calls(ref):
    """renderReSTIR called as a whole (reference src/rendering/render.cpp:28-62) gives the same image and grid as the
    stage functions called one by one, so per-stage dumps are faithful."""
    scene = load_scene("CornellNightClub"); ref.set_scene(scene)
    feat = Features(spatialResamplingPasses=2)
    out = []
    for flags in (0, REF_FLAG_WHOLE_FRAME):
        ref.reset_history()
        for fr in range(2):
            rf = ref.render_frame(feat, NIGHTCLUB_CAM, 48, 36, fr > 0, 5, fr, flags)
        out.append(rf)
    assert_bits_equal(out[0].image, out[1].image, "image")
    assert_bits_equal(out[0].stages[abi.ROMIS_PASS_FINAL].W, out[1].stages[abi.ROMIS_PASS_FINAL].W, "final W")
    assert_bits_equal(out[0].stages[abi.ROMIS_PASS_FINAL].M, out[1].stages[abi.ROMIS_PASS_FINAL].M, "final M")


def test_many_light_scene_and_bruteforce_tracer(ref, oracle_factory):
    scene = load_scene("Monkey"); scene.lights = synthetic_lights(2048, seed=9)
    ref.set_scene(scene); ref.set_tracer_mode(0)
    try:
        orc = oracle_factory(1); orc.upload_scene(scene)
        feat = Features(initialSamplesVisibilityCheck=True, spatialResamplingPasses=1)
        rcam = ref.make_camera(CORNELL_CAM, 40, 40)
        for fr in range(2):
            rf = ref.render_frame(feat, CORNELL_CAM, 40, 40, fr > 0, 3, fr)
            img = orc.render_frame(feat, rcam, 40, 40, fr > 0, 3, fr)
            assert_bits_equal(img, rf.image, f"frame {fr} image (reference: brute-force tracer, oracle: BVH)")
    finally:
        ref.set_tracer_mode(1)
