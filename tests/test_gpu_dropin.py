"""The drop-in as a maintainer would build it: the reference's own translation units (Scene, Trackball, Screen, Features)
with the body of renderReSTIR replaced by integration/render_restir_gpu.cpp -> libromis_gpu.so.  The Screen it fills must
equal, bit for bit, the Screen the reference's CPU renderReSTIR fills for the same injected random stream."""
import os

import pytest

from oracle import pyoracle
from romis_b200 import abi
from romis_b200.scene import Features, RmisParams
from cases import CORNELL_CAM, NIGHTCLUB_CAM
from common import assert_bits_equal, assert_image_rmse, assert_solve_tolerance, load_scene

pytestmark = [pytest.mark.gpu,
              pytest.mark.skipif(not os.path.exists(pyoracle.DROPIN_SO), reason="oracle/_ref/libromis_dropin.so not built (make -C oracle dropin)")]


@pytest.mark.parametrize("scene_name,cam,feat,W,H", [
    ("CornellNightClub", NIGHTCLUB_CAM, Features(spatialResamplingPasses=3, initialSamplesVisibilityCheck=True), 96, 54),
    ("CornellBoxParallelogramLight", CORNELL_CAM, Features(spatialResamplingPasses=1), 64, 64),
    ("CubeTextured", CORNELL_CAM, Features(unbiasedCombination=True, spatialReuseVisibilityCheck=True), 40, 40),
])
def test_dropin_fills_the_reference_screen_identically(scene_name, cam, feat, W, H):
    lib = pyoracle.DropinLib()
    lib.set_scene(load_scene(scene_name))
    lib.reset_history()
    for fr in range(3):
        cpu = lib.render_frame(feat, cam, W, H, fr > 0, 314, fr, pyoracle.REF_FLAG_WHOLE_FRAME, dump=False).image
        gpu = lib.render_frame_gpu(feat, cam, W, H, fr > 0, 314, fr)
        assert_bits_equal(gpu, cpu, f"{scene_name} frame {fr}: Screen::pixels() of the GPU drop-in vs the reference's renderReSTIR")


def test_dropin_rmis_and_romis_fill_the_reference_screen():
    """renderRMIS / renderROMIS replaced by their GPU bodies: R-MIS fills the Screen the reference's renderRMIS fills bit for bit;
    R-OMIS to the solve's tolerance (tests/test_romis_oracle.py)."""
    lib = pyoracle.DropinLib()
    lib.set_scene(load_scene("CornellNightClub"))
    W, H = 64, 36
    feat = Features(initialSamplesVisibilityCheck=True)
    for rp in (RmisParams(maxIterationsMIS=2), RmisParams(maxIterationsMIS=2, misWeightRMIS=abi.ROMIS_MIS_BALANCE,
                                                          neighbourSelectionStrategy=abi.ROMIS_NEIGHBOURS_EQUAL_SIMILAR_DISSIMILAR)):
        cpu, _xy, _cnt = lib.render_frame_rmis(feat, rp, NIGHTCLUB_CAM, W, H, 2718, 1, False)
        gpu = lib.render_frame_mis_gpu(False, feat, rp, NIGHTCLUB_CAM, W, H, 2718, 1)
        assert_bits_equal(gpu, cpu, "Screen::pixels() of the GPU renderRMIS vs the reference's")
    rp = RmisParams(maxIterationsMIS=2)
    cpu, _A, _B = lib.render_frame_romis(feat, rp, NIGHTCLUB_CAM, W, H, 2718, 2, False)
    gpu = lib.render_frame_mis_gpu(True, feat, rp, NIGHTCLUB_CAM, W, H, 2718, 2)
    assert_solve_tolerance(gpu, cpu, "Screen::pixels() of the GPU renderROMIS vs the reference's")
    assert_image_rmse(gpu, cpu, 1e-3, "Screen::pixels() of the GPU renderROMIS vs the reference's (RMSE)")
