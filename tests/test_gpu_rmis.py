"""R-MIS mode on the GPU (romis_render_frame_rmis, through the C-ABI) against the oracle and the golden vectors the
REFERENCE's renderRMIS produced (reference src/rendering/render.cpp:64-119).  Neighbour grid and counts bit-exact
(integer work); image within 1e-4 relative / 1e-3 RMSE (north_star) and, as built, bit-exact."""
import numpy as np
import pytest

from romis_b200 import abi
from romis_b200.scene import Features, RmisParams, synthetic_lights
from cases import NIGHTCLUB_CAM, RMIS_CASES
from common import assert_bits_equal, assert_rel_close, camera_from_array, load_golden, load_scene

pytestmark = pytest.mark.gpu


@pytest.fixture(scope="module")
def renderer():
    from romis_b200.api import RestirRenderer
    r = RestirRenderer(0)
    yield r
    r.close()


@pytest.mark.parametrize("case", sorted(RMIS_CASES))
def test_rmis_gpu_matches_oracle_and_golden(case, renderer, oracle_factory):
    scene_name, W, H, feat, rmis, _cam, seed, frame = RMIS_CASES[case]
    g = load_golden(case)
    scene = load_scene(scene_name)
    cam = camera_from_array(g["camera"])
    orc = oracle_factory(); orc.upload_scene(scene)
    oimg, oxy, ocnt = orc.render_frame_rmis(feat, rmis, cam, W, H, seed, frame)
    renderer.upload_scene(scene)
    gimg = renderer.render_frame_rmis(feat, rmis, cam, W, H, seed, frame)
    gxy, gcnt = renderer.rmis_neighbours()
    assert_bits_equal(gcnt, ocnt, f"{case} neighbour count vs oracle")
    assert_bits_equal(gxy, oxy, f"{case} neighbour grid vs oracle")
    assert_bits_equal(gcnt, g["count"].astype(np.uint32), f"{case} neighbour count vs reference")
    assert_bits_equal(gxy, g["neighbours"].astype(np.int32), f"{case} neighbour grid vs reference")
    assert_rel_close(gimg, oimg, 1e-4, f"{case} image vs oracle")
    rmse = float(np.sqrt(np.mean((gimg.astype(np.float64) - g["image"]) ** 2)))
    assert rmse <= 1e-3, f"{case} image RMSE vs reference {rmse}"
    assert_bits_equal(gimg, g["image"], f"{case} image vs reference")


def test_rmis_larger_frame_many_lights(renderer, oracle_factory):
    """256x144 nightclub geometry with 4096 synthetic lights, balance heuristic, every strategy: GPU = oracle bit for bit."""
    scene = load_scene("CornellNightClub"); scene.lights = synthetic_lights(4096, seed=5)
    W, H = 256, 144
    cam = NIGHTCLUB_CAM.to_abi(W, H)
    orc = oracle_factory(); orc.upload_scene(scene); renderer.upload_scene(scene)
    feat = Features(initialSamplesVisibilityCheck=True)
    for strategy in (abi.ROMIS_NEIGHBOURS_RANDOM, abi.ROMIS_NEIGHBOURS_SIMILAR, abi.ROMIS_NEIGHBOURS_EQUAL_SIMILAR_DISSIMILAR):
        rp = RmisParams(maxIterationsMIS=2, misWeightRMIS=abi.ROMIS_MIS_BALANCE, neighbourSelectionStrategy=strategy)
        oimg, oxy, ocnt = orc.render_frame_rmis(feat, rp, cam, W, H, 99, 4)
        gimg = renderer.render_frame_rmis(feat, rp, cam, W, H, 99, 4)
        gxy, gcnt = renderer.rmis_neighbours()
        assert_bits_equal(gcnt, ocnt, f"strategy {strategy} count"); assert_bits_equal(gxy, oxy, f"strategy {strategy} grid")
        assert_bits_equal(gimg, oimg, f"strategy {strategy} image")


def test_rmis_no_neighbours(renderer, oracle_factory):
    """numNeighboursToSample = 0 (the reference's UI allows it, ui.cpp:307): random and similar strategies resample nothing but
    the pixel itself; EqualSimilarDissimilar is rejected (its unsigned count wraps in the reference and takes the whole window,
    neighbour_selection.cpp:95-98 -- the k + 1 planes of the grid cannot hold that)."""
    from romis_b200.api import RomisError
    scene = load_scene("CornellNightClub"); W, H = 40, 24
    cam = NIGHTCLUB_CAM.to_abi(W, H)
    orc = oracle_factory(); orc.upload_scene(scene); renderer.upload_scene(scene)
    feat = Features(numNeighboursToSample=0, initialSamplesVisibilityCheck=True)
    for strategy in (abi.ROMIS_NEIGHBOURS_RANDOM, abi.ROMIS_NEIGHBOURS_SIMILAR):
        for weights in (abi.ROMIS_MIS_EQUAL, abi.ROMIS_MIS_BALANCE):
            rp = RmisParams(maxIterationsMIS=2, misWeightRMIS=weights, neighbourSelectionStrategy=strategy)
            oimg, oxy, ocnt = orc.render_frame_rmis(feat, rp, cam, W, H, 12, 0)
            gimg = renderer.render_frame_rmis(feat, rp, cam, W, H, 12, 0)
            gxy, gcnt = renderer.rmis_neighbours()
            assert (gcnt == 1).all()
            assert_bits_equal(gcnt, ocnt, f"k=0 strategy {strategy} count"); assert_bits_equal(gxy, oxy, f"k=0 strategy {strategy} grid")
            assert_bits_equal(gimg, oimg, f"k=0 strategy {strategy} weights {weights} image")
    for render in (renderer.render_frame_rmis, renderer.render_frame_romis):
        with pytest.raises(RomisError, match="EqualSimilarDissimilar"):
            render(feat, RmisParams(neighbourSelectionStrategy=abi.ROMIS_NEIGHBOURS_EQUAL_SIMILAR_DISSIMILAR), cam, W, H, 12, 0)


def test_rmis_leaves_restir_history_alone(renderer, oracle_factory):
    """An R-MIS frame between two ReSTIR frames must not disturb the temporal history (it works in a scratch buffer)."""
    scene = load_scene("CornellNightClub"); W, H = 48, 32
    cam = NIGHTCLUB_CAM.to_abi(W, H); feat = Features()
    renderer.upload_scene(scene); renderer.reset_history()
    renderer.render_frame(feat, cam, W, H, False, 5, 0)
    a = renderer.render_frame(feat, cam, W, H, True, 5, 1)
    renderer.reset_history()
    renderer.render_frame(feat, cam, W, H, False, 5, 0)
    renderer.render_frame_rmis(feat, RmisParams(maxIterationsMIS=1), cam, W, H, 5, 7)
    b = renderer.render_frame(feat, cam, W, H, True, 5, 1)
    assert_bits_equal(a, b, "ReSTIR frame 1 with and without an R-MIS frame in between")


def test_rmis_error_paths(renderer):
    from romis_b200.api import RomisError
    scene = load_scene("Cube"); renderer.upload_scene(scene)
    cam = NIGHTCLUB_CAM.to_abi(16, 16)
    with pytest.raises(RomisError):
        renderer.render_frame_rmis(Features(), RmisParams(neighbourSelectionStrategy=abi.ROMIS_NEIGHBOURS_DISSIMILAR), cam, 16, 16, 1, 0)
    with pytest.raises(RomisError):
        renderer.render_frame_rmis(Features(), RmisParams(maxIterationsMIS=0), cam, 16, 16, 1, 0)
    with pytest.raises(RomisError):
        renderer.render_frame_rmis(Features(), RmisParams(misWeightRMIS=7), cam, 16, 16, 1, 0)
    with pytest.raises(RomisError):
        renderer.render_frame_rmis(Features(spatialResampleRadius=31), RmisParams(), cam, 16, 16, 1, 0)
    renderer.set_band(0, 24)                        # a band beyond the image
    try:
        with pytest.raises(RomisError):
            renderer.render_frame_rmis(Features(), RmisParams(), cam, 16, 16, 1, 0)
    finally:
        renderer.set_band(0, 0)


@pytest.mark.parametrize("mode", ["rmis", "romis", "romis_progressive"])
def test_mis_frames_as_row_bands_equal_the_whole_frame(mode, renderer):
    """R-MIS / R-OMIS frames sharded into row bands (romis_set_band): a band renders its halo rows itself -- primary rays and
    every iteration's initial reservoirs are functions of the pixel alone -- so the bands, assembled, equal the undivided frame
    bit for bit, with uneven bands, bands of fewer rows than the radius apart from the image edge, and for all three estimators."""
    scene = load_scene("CornellNightClub"); scene.lights = synthetic_lights(512, seed=3)
    W, H = 192, 108
    cam = NIGHTCLUB_CAM.to_abi(W, H)
    renderer.upload_scene(scene)
    feat = Features(initialSamplesVisibilityCheck=True, numSamplesInReservoir=6 if mode == "romis_progressive" else 2)
    rp = RmisParams(maxIterationsMIS=3, misWeightRMIS=abi.ROMIS_MIS_BALANCE, neighbourSelectionStrategy=abi.ROMIS_NEIGHBOURS_SIMILAR,
                    useProgressiveROMIS=mode == "romis_progressive", progressiveUpdateMod=1)
    render = renderer.render_frame_rmis if mode == "rmis" else renderer.render_frame_romis
    renderer.set_band(0, 0)
    whole = render(feat, rp, cam, W, H, 5, 2)
    banded = np.full((H, W, 3), np.nan, np.float32)
    for y0, y1 in ((0, 37), (37, 49), (49, 95), (95, H)):
        renderer.set_band(y0, y1)
        render(feat, rp, cam, W, H, 5, 2, out=banded)
    renderer.set_band(0, 0)
    assert not np.isnan(banded).any(), "rows left unwritten by the bands"
    assert_bits_equal(banded, whole, f"{mode}: bands vs whole frame")
