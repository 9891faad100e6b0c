"""R-OMIS mode on the GPU (romis_render_frame_romis, through the C-ABI; reference renderROMIS, src/rendering/render.cpp:121-265).
Technique matrices and contribution vectors: bit-exact against the oracle AND against the reference's golden dumps.  Image:
bit-exact against the oracle (both run include/romis_cod.h, no FMA on either side), tolerance against the reference (Eigen's
SIMD summation order inside the solve, see tests/test_romis_oracle.py)."""
import numpy as np
import pytest

from romis_b200 import abi
from romis_b200.scene import Features, RmisParams, synthetic_lights
from cases import NIGHTCLUB_CAM, ROMIS_CASES
from common import assert_bits_equal, assert_image_rmse, assert_solve_tolerance, camera_from_array, load_golden, load_scene

pytestmark = pytest.mark.gpu


@pytest.fixture(scope="module")
def renderer():
    from romis_b200.api import RestirRenderer
    r = RestirRenderer(0)
    yield r
    r.close()


@pytest.mark.parametrize("case", sorted(ROMIS_CASES))
def test_romis_gpu_matches_oracle_and_golden(case, renderer, oracle_factory):
    scene_name, W, H, feat, rmis, _cam, seed, frame = ROMIS_CASES[case]
    g = load_golden(case)
    scene = load_scene(scene_name)
    cam = camera_from_array(g["camera"])
    orc = oracle_factory(); orc.upload_scene(scene)
    oimg, oA, oB = orc.render_frame_romis(feat, rmis, cam, W, H, seed, frame)
    renderer.upload_scene(scene)
    gimg = renderer.render_frame_romis(feat, rmis, cam, W, H, seed, frame)
    gA, gB = renderer.romis_system()
    assert_bits_equal(gA, oA, f"{case} technique matrices vs oracle"); assert_bits_equal(gB, oB, f"{case} contribution vectors vs oracle")
    assert_bits_equal(gA, g["matrices"], f"{case} technique matrices vs reference")
    assert_bits_equal(gB, g["contributions"], f"{case} contribution vectors vs reference")
    assert_bits_equal(gimg, oimg, f"{case} image vs oracle")
    assert_solve_tolerance(gimg, g["image"], f"{case} image vs reference")
    assert_image_rmse(gimg, g["image"], 1e-3, f"{case} image vs reference (north_star: RMSE <= 1e-3)")


def test_romis_larger_frame_many_lights(renderer, oracle_factory):
    scene = load_scene("CornellNightClub"); scene.lights = synthetic_lights(4096, seed=6)
    W, H = 192, 108
    cam = NIGHTCLUB_CAM.to_abi(W, H)
    orc = oracle_factory(); orc.upload_scene(scene); renderer.upload_scene(scene)
    feat = Features(initialSamplesVisibilityCheck=True)
    for strategy in (abi.ROMIS_NEIGHBOURS_RANDOM, abi.ROMIS_NEIGHBOURS_SIMILAR):
        rp = RmisParams(maxIterationsMIS=2, neighbourSelectionStrategy=strategy)
        oimg, oA, oB = orc.render_frame_romis(feat, rp, cam, W, H, 55, 3)
        gimg = renderer.render_frame_romis(feat, rp, cam, W, H, 55, 3)
        gA, gB = renderer.romis_system()
        assert_bits_equal(gA, oA, f"strategy {strategy} technique matrices"); assert_bits_equal(gB, oB, f"strategy {strategy} contribution vectors")
        assert_bits_equal(gimg, oimg, f"strategy {strategy} image")


def test_romis_error_paths(renderer):
    from romis_b200.api import RomisError
    renderer.upload_scene(load_scene("Cube"))
    cam = NIGHTCLUB_CAM.to_abi(16, 16)
    for feat, rp in ((Features(), RmisParams(useProgressiveROMIS=True, progressiveUpdateMod=0)),
                     (Features(spatialResampleRadius=1), RmisParams()),                 # corner windows hold 3 < k pixels
                     (Features(numNeighboursToSample=11, spatialResampleRadius=8), RmisParams()),
                     (Features(), RmisParams(neighbourSelectionStrategy=abi.ROMIS_NEIGHBOURS_DISSIMILAR)),
                     (Features(), RmisParams(maxIterationsMIS=0))):
        with pytest.raises(RomisError):
            renderer.render_frame_romis(feat, rp, cam, 16, 16, 1, 0)
