"""R-MIS mode (renderRMIS, reference src/rendering/render.cpp:64-119): the C restatement (oracle/restir_oracle.c
orc_render_frame_rmis) against golden vectors produced by the REFERENCE's own renderRMIS /
generateResampleIndicesGrid (tests/golden/gen_golden.py), and live against the compiled reference where it is built.
Bit-exact on the neighbour index grid, the neighbour counts and the image."""
import os

import numpy as np
import pytest

from oracle import pyoracle
from romis_b200 import abi
from romis_b200.scene import Features, RmisParams
from cases import CORNELL_CAM, NIGHTCLUB_CAM, RMIS_CASES
from common import assert_bits_equal, camera_from_array, load_golden, load_scene


@pytest.mark.parametrize("case", sorted(RMIS_CASES))
def test_rmis_oracle_matches_reference_golden(case, oracle_factory):
    scene_name, W, H, feat, rmis, _cam, seed, frame = RMIS_CASES[case]
    g = load_golden(case)
    orc = oracle_factory(); orc.upload_scene(load_scene(scene_name))
    img, xy, cnt = orc.render_frame_rmis(feat, rmis, camera_from_array(g["camera"]), W, H, seed, frame)
    assert_bits_equal(cnt, g["count"].astype(np.uint32), f"{case} neighbour count")
    assert_bits_equal(xy, g["neighbours"].astype(np.int32), f"{case} neighbour grid")
    assert_bits_equal(img, g["image"], f"{case} image")


def test_rmis_dissimilar_is_rejected(oracle_factory):
    """NeighbourSelectionStrategy::Dissimilar hands std::sample a negative count (neighbour_selection.cpp:88-93)."""
    orc = oracle_factory(); orc.upload_scene(load_scene("Cube"))
    with pytest.raises(RuntimeError):
        orc.render_frame_rmis(Features(), RmisParams(neighbourSelectionStrategy=abi.ROMIS_NEIGHBOURS_DISSIMILAR),
                              CORNELL_CAM.to_abi(8, 8), 8, 8, 1, 0)


@pytest.mark.skipif(not os.path.exists(pyoracle.REF_SO), reason="oracle/_ref not built (needs /root/reference: make -C oracle ref)")
@pytest.mark.parametrize("strategy", [abi.ROMIS_NEIGHBOURS_RANDOM, abi.ROMIS_NEIGHBOURS_SIMILAR, abi.ROMIS_NEIGHBOURS_EQUAL_SIMILAR_DISSIMILAR])
@pytest.mark.parametrize("mis", [abi.ROMIS_MIS_EQUAL, abi.ROMIS_MIS_BALANCE])
def test_rmis_restatement_equals_compiled_reference(oracle_factory, strategy, mis):
    ref = pyoracle.RefLib()
    scene = load_scene("CornellNightClub"); ref.set_scene(scene)
    orc = oracle_factory(); orc.upload_scene(scene)
    W, H = 44, 31
    rcam = ref.make_camera(NIGHTCLUB_CAM, W, H)
    for k, r, n in ((5, 10, 2), (9, 1, 1), (2, 3, 4)):
        feat = Features(numNeighboursToSample=k, spatialResampleRadius=r, numSamplesInReservoir=n)
        rp = RmisParams(maxIterationsMIS=2, misWeightRMIS=mis, neighbourSelectionStrategy=strategy)
        ri, rxy, rcnt = ref.render_frame_rmis(feat, rp, NIGHTCLUB_CAM, W, H, 1234 + k, 2)
        oi, oxy, ocnt = orc.render_frame_rmis(feat, rp, rcam, W, H, 1234 + k, 2)
        assert_bits_equal(ocnt, rcnt, f"k={k} r={r} count"); assert_bits_equal(oxy, rxy, f"k={k} r={r} grid")
        assert_bits_equal(oi, ri, f"k={k} r={r} image")
