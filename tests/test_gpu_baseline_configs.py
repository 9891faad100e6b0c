"""BASELINE.json's other configurations as parity cases (reduced sizes, same structure), CUDA path vs the oracle:
  C3  monkey.obj + synthetic many point/area lights, row-band shards with halo exchange
  C4  temporal sequence with an orbiting camera, temporal M-clamp 20x
  C5  spatial-reuse sweep k in {3,5,10} x radius in {10,30} x iterations 1..4 on a synthetic many-light scene
Bars as in test_gpu_parity.py: light indices / M bit-exact, weights and image within 1e-4 relative (compared bit for bit)."""
import ctypes as C

import numpy as np
import pytest

from romis_b200 import abi
from romis_b200.scene import Camera, Features, synthetic_lights
from cases import CORNELL_CAM, NIGHTCLUB_CAM
from common import assert_bits_equal, load_scene

pytestmark = pytest.mark.gpu


def final_state_equal(tag, renderer, orc):
    g, o = renderer.reservoirs(abi.ROMIS_PASS_FINAL), orc.reservoirs(abi.ROMIS_PASS_FINAL)
    for fld in ("light_id", "M", "u", "v", "W"):
        assert_bits_equal(getattr(g, fld), getattr(o, fld), f"{tag} final {fld}")


def test_c4_orbiting_camera_sequence(oracle_factory):
    """64-frame orbit of BASELINE C4, here 10 frames at 96x54: rotation.y = 30 deg + 360 deg * f / 64."""
    from romis_b200.api import RestirRenderer
    scene = load_scene("CornellNightClub")
    feat = Features(spatialResamplingPasses=3, initialSamplesVisibilityCheck=True, temporalClampM=20)
    W, H = 96, 54
    r = RestirRenderer(0); r.upload_scene(scene)
    orc = oracle_factory(); orc.upload_scene(scene); orc.reset_history()
    try:
        for f in range(10):
            cam = Camera(rotation_deg=(10.3, 30.0 + 360.0 * f / 64.0, 0.0)).to_abi(W, H)
            oimg = orc.render_frame(feat, cam, W, H, f > 0, 64, f)
            gimg = r.render_frame(feat, cam, W, H, f > 0, 64, f)
            final_state_equal(f"orbit frame {f}", r, orc)
            assert_bits_equal(gimg, oimg, f"orbit frame {f} image")
    finally:
        r.close()


@pytest.mark.parametrize("k,radius,passes", [(3, 10, 1), (5, 10, 2), (10, 10, 4), (3, 30, 3), (5, 30, 4), (10, 30, 2)])
def test_c5_spatial_sweep(oracle_factory, k, radius, passes):
    from romis_b200.api import RestirRenderer
    scene = load_scene("CornellNightClub")
    scene.lights = synthetic_lights(8192, seed=11, intensity=4000.0)
    scene.lights["p0"] += np.array([2.5, 2.0, -1.0], np.float32)        # move the light shell into the room
    feat = Features(numNeighboursToSample=k, spatialResampleRadius=radius, spatialResamplingPasses=passes)
    W, H = 80, 64
    cam = NIGHTCLUB_CAM.to_abi(W, H)
    r = RestirRenderer(0); r.upload_scene(scene)
    orc = oracle_factory(); orc.upload_scene(scene); orc.reset_history()
    try:
        for f in range(2):
            oimg = orc.render_frame(feat, cam, W, H, f > 0, 5, f)
            gimg = r.render_frame(feat, cam, W, H, f > 0, 5, f)
            final_state_equal(f"k={k} r={radius} P={passes} frame {f}", r, orc)
            assert_bits_equal(gimg, oimg, f"k={k} r={radius} P={passes} frame {f} image")
    finally:
        r.close()


@pytest.mark.parametrize("unbiased", [False, True])
def test_c3_many_lights_row_bands(oracle_factory, unbiased):
    """monkey + 16384 synthetic point / parallelogram lights, 4 row bands with hand-moved halos, biased and unbiased + visibility."""
    from romis_b200.api import RestirRenderer
    cudart = C.CDLL("libcudart.so")
    scene = load_scene("Monkey")
    scene.lights = synthetic_lights(16384, seed=1)
    feat = Features(spatialResamplingPasses=2, spatialResampleRadius=8, initialSamplesVisibilityCheck=True,
                    unbiasedCombination=unbiased, spatialReuseVisibilityCheck=unbiased)
    W, H = 120, 96
    cam = CORNELL_CAM.to_abi(W, H)
    orc = oracle_factory(); orc.upload_scene(scene); orc.reset_history()
    edges = [0, 20, 47, 70, H]
    bands = [RestirRenderer(0) for _ in range(4)]
    for i, b in enumerate(bands):
        b.set_band(edges[i], edges[i + 1]); b.upload_scene(scene)
    try:
        for f in range(2):
            oimg = orc.render_frame(feat, cam, W, H, f > 0, 3, f)
            img = np.zeros((H, W, 3), np.float32)
            for b in bands:
                b.frame_begin(feat, cam, W, H, f > 0, 3, f)
            for p in range(feat.spatialResamplingPasses):
                for b in bands:
                    b.synchronize()
                for i in range(3):
                    s, n = bands[i].halo_region(abi.ROMIS_HALO_SEND_HIGH); d, m = bands[i + 1].halo_region(abi.ROMIS_HALO_RECV_LOW)
                    assert n == m > 0 and cudart.cudaMemcpy(C.c_void_p(d), C.c_void_p(s), C.c_size_t(n), 3) == 0
                    s, n = bands[i + 1].halo_region(abi.ROMIS_HALO_SEND_LOW); d, m = bands[i].halo_region(abi.ROMIS_HALO_RECV_HIGH)
                    assert n == m > 0 and cudart.cudaMemcpy(C.c_void_p(d), C.c_void_p(s), C.c_size_t(n), 3) == 0
                assert cudart.cudaDeviceSynchronize() == 0
                for b in bands:
                    b.frame_spatial_pass(p)
            for b in bands:
                b.frame_end(img)
            assert_bits_equal(img, oimg, f"C3-style banded frame {f} (unbiased={unbiased}) vs oracle")
    finally:
        for b in bands:
            b.close()


@pytest.mark.parametrize("n_lights", [65536, 1 << 20, 100003])
@pytest.mark.parametrize("unbiased", [False, True])
def test_c3_c5_full_light_counts(oracle_factory, n_lights, unbiased):
    """The light counts BASELINE C3 (65 536) and C5 (2^20) name, at a small resolution: the power-of-two shortcut of the initial
    weight (pdf * L instead of pdf / (1 / L)), the light pick at large ranges and the 96-byte record gather, biased and unbiased
    + visibility; 100 003 lights take the division route."""
    from romis_b200.api import RestirRenderer
    scene = load_scene("Monkey")
    scene.lights = synthetic_lights(n_lights, seed=3, intensity=24.0)
    feat = Features(spatialResamplingPasses=2, initialSamplesVisibilityCheck=True,
                    unbiasedCombination=unbiased, spatialReuseVisibilityCheck=unbiased)
    W, H = 96, 64
    cam = CORNELL_CAM.to_abi(W, H)
    r = RestirRenderer(0); r.upload_scene(scene)
    orc = oracle_factory(); orc.upload_scene(scene); orc.reset_history()
    try:
        for f in range(2):
            oimg = orc.render_frame(feat, cam, W, H, f > 0, 17, f)
            gimg = r.render_frame(feat, cam, W, H, f > 0, 17, f)
            final_state_equal(f"L={n_lights} unbiased={unbiased} frame {f}", r, orc)
            assert_bits_equal(gimg, oimg, f"L={n_lights} unbiased={unbiased} frame {f} image")
        ids = r.reservoirs(abi.ROMIS_PASS_FINAL).light_id
        held = ids[ids != 0xFFFFFFFF]
        assert held.size and held.max() < n_lights and held.max() > n_lights // 2     # picks span the whole table
    finally:
        r.close()


def test_mid_size_unbiased_multi_frame(oracle_factory):
    """A larger unbiased + visibility run (rare paths: Z = 0, empty sub-reservoirs, generic N)."""
    from romis_b200.api import RestirRenderer
    scene = load_scene("CornellNightClub")
    W, H = 160, 90
    cam = NIGHTCLUB_CAM.to_abi(W, H)
    for feat in (Features(unbiasedCombination=True, spatialReuseVisibilityCheck=True, initialSamplesVisibilityCheck=True, spatialResamplingPasses=2),
                 Features(numSamplesInReservoir=6, initialLightSamples=8, spatialResamplingPasses=2, temporalClampM=1)):
        r = RestirRenderer(0); r.upload_scene(scene)
        orc = oracle_factory(); orc.upload_scene(scene); orc.reset_history()
        try:
            for f in range(3):
                oimg = orc.render_frame(feat, cam, W, H, f > 0, 21, f)
                gimg = r.render_frame(feat, cam, W, H, f > 0, 21, f)
                final_state_equal(f"N={feat.numSamplesInReservoir} unbiased={feat.unbiasedCombination} frame {f}", r, orc)
                assert_bits_equal(gimg, oimg, f"frame {f} image")
        finally:
            r.close()
