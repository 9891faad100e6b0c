"""R-OMIS mode (renderROMIS, reference src/rendering/render.cpp:121-265), direct and progressive estimator.

What is pinned how:
* neighbour grid: bit-exact (tests/test_rmis_oracle.py, same generateResampleIndicesGrid);
* the per-pixel technique matrices and contribution vectors after the last iteration: BIT-EXACT against the reference --
  renderROMIS hands them to visualiseAlphas after every iteration (render.cpp:227-229), which the harness receives
  (oracle/ref_harness/ref_api.cpp), so golden vectors of them exist;
* the image: the per-pixel solves go through include/romis_cod.h, a restatement of Eigen's COD whose reductions are summed
  front to back, while Eigen sums them in alignment-dependent SSE packets.  The systems are ill-conditioned by construction
  (cond 1e5 .. 1e6), so the two fp32 solves agree to cond * eps: bar = common.assert_solve_tolerance (97 % of the channels
  within 1e-3 relative, 99.7 % within 1e-1), stated as "tolerance-pinned" in DESIGN.md.
"""
import os

import numpy as np
import pytest

from oracle import pyoracle
from romis_b200 import abi
from romis_b200.scene import Features, RmisParams
from cases import CORNELL_CAM, NIGHTCLUB_CAM, ROMIS_CASES
from common import assert_bits_equal, assert_image_rmse, assert_solve_tolerance, camera_from_array, load_golden, load_scene


@pytest.mark.parametrize("case", sorted(ROMIS_CASES))
def test_romis_oracle_matches_reference_golden(case, oracle_factory):
    scene_name, W, H, feat, rmis, _cam, seed, frame = ROMIS_CASES[case]
    g = load_golden(case)
    orc = oracle_factory(); orc.upload_scene(load_scene(scene_name))
    img, A, B = orc.render_frame_romis(feat, rmis, camera_from_array(g["camera"]), W, H, seed, frame)
    assert_bits_equal(A, g["matrices"], f"{case} technique matrices")
    assert_bits_equal(B, g["contributions"], f"{case} contribution vectors")
    assert_solve_tolerance(img, g["image"], f"{case} image")
    assert_image_rmse(img, g["image"], 1e-3, f"{case} image vs reference (north_star: RMSE <= 1e-3)")


def test_cod_solver_against_float64_least_squares(oracle_factory):
    """include/romis_cod.h on its own: full-rank systems reproduce the float64 solution, rank-deficient ones (the common
    R-OMIS case: neighbouring techniques are nearly identical) the minimum-norm least-squares solution."""
    orc = oracle_factory()
    rng = np.random.default_rng(11)
    for n in (1, 2, 3, 6, 8, 11):
        for rank in sorted({n, max(1, n // 2), 1}):
            for _ in range(20):
                V = rng.normal(size=(rank, n))
                A = (V.T @ V).astype(np.float32)                   # symmetric PSD of the given rank, like sum of v v^T
                b = (A.astype(np.float64) @ rng.normal(size=n)).astype(np.float32)      # consistent right-hand side
                x, r = orc.cod_solve(A, b)
                assert r == rank, (n, rank, r)
                ref = np.linalg.pinv(A.astype(np.float64), rcond=1e-6) @ b.astype(np.float64)
                scale = max(1.0, np.abs(ref).max())
                cond = np.linalg.cond(V @ V.T)
                assert np.abs(x - ref).max() <= 2e-5 * cond * scale, (n, rank, np.abs(x - ref).max(), cond)
    x, r = orc.cod_solve(np.zeros((4, 4), np.float32), np.ones(4, np.float32))
    assert r == 0 and not x.any()                                   # rank 0 -> zero solution (_solve_impl)


def test_romis_unsupported_parameters_are_rejected(oracle_factory):
    orc = oracle_factory(); orc.upload_scene(load_scene("Cube")); cam = CORNELL_CAM.to_abi(8, 8)
    with pytest.raises(RuntimeError):   # iteration % progressiveUpdateMod with a zero modulus
        orc.render_frame_romis(Features(), RmisParams(useProgressiveROMIS=True, progressiveUpdateMod=0), cam, 8, 8, 1, 0)
    with pytest.raises(RuntimeError):   # window of 3 other pixels at the corners, k = 5: the reference reads out of bounds
        orc.render_frame_romis(Features(spatialResampleRadius=1), RmisParams(), cam, 8, 8, 1, 0)


@pytest.mark.skipif(not os.path.exists(pyoracle.REF_SO), reason="oracle/_ref not built (needs /root/reference: make -C oracle ref)")
@pytest.mark.parametrize("strategy", [abi.ROMIS_NEIGHBOURS_RANDOM, abi.ROMIS_NEIGHBOURS_SIMILAR])
def test_romis_restatement_equals_compiled_reference(oracle_factory, strategy):
    ref = pyoracle.RefLib()
    scene = load_scene("CornellNightClub"); ref.set_scene(scene)
    orc = oracle_factory(); orc.upload_scene(scene)
    W, H = 40, 29
    rcam = ref.make_camera(NIGHTCLUB_CAM, W, H)
    for k, r, n in ((5, 10, 2), (2, 3, 4)):
        feat = Features(numNeighboursToSample=k, spatialResampleRadius=r, numSamplesInReservoir=n)
        rp = RmisParams(maxIterationsMIS=2, neighbourSelectionStrategy=strategy)
        ri, rA, rB = ref.render_frame_romis(feat, rp, NIGHTCLUB_CAM, W, H, 4321 + k, 1)
        oi, oA, oB = orc.render_frame_romis(feat, rp, rcam, W, H, 4321 + k, 1)
        assert_bits_equal(oA, rA, f"k={k} technique matrices"); assert_bits_equal(oB, rB, f"k={k} contribution vectors")
        assert_solve_tolerance(oi, ri, f"k={k} image")
        assert_image_rmse(oi, ri, 1e-3, f"k={k} image (RMSE)")
