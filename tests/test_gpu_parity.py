"""CUDA path (libromis_gpu.so, through the C-ABI) against the oracle and the reference's golden vectors.

Bars (BASELINE.json north_star): selected light indices and M counts bit-exact; reservoir weights and
radiance within 1e-4 relative (they are in fact expected bit-exact too and are compared that way
against the golden vectors, with the 1e-4 bar as the stated contract); G-buffer bit-exact.
"""
import numpy as np
import pytest

from romis_b200 import abi
from romis_b200.scene import Features, synthetic_lights
from cases import CASES, NIGHTCLUB_CAM
from common import (assert_bits_equal, assert_rel_close, camera_from_array, load_golden, load_scene, stage_ids)

pytestmark = pytest.mark.gpu
REL = 1e-4      # north_star: weights and radiance within 1e-4 relative


@pytest.fixture(scope="module")
def renderer():
    from romis_b200.api import RestirRenderer
    r = RestirRenderer(0)
    r.set_capture(True)
    yield r
    r.close()


def compare_stage(case, tag, gpu_st, orc_st):
    assert_bits_equal(gpu_st.light_id, orc_st.light_id, f"{case} {tag} light index")
    assert_bits_equal(gpu_st.M, orc_st.M, f"{case} {tag} M")
    assert_bits_equal(gpu_st.u, orc_st.u, f"{case} {tag} u")
    assert_bits_equal(gpu_st.v, orc_st.v, f"{case} {tag} v")
    assert_rel_close(gpu_st.W, orc_st.W, REL, f"{case} {tag} W")
    assert_rel_close(gpu_st.position, orc_st.position, REL, f"{case} {tag} position")
    assert_rel_close(gpu_st.color, orc_st.color, REL, f"{case} {tag} color")


@pytest.mark.parametrize("case", sorted(CASES))
def test_gpu_matches_oracle_and_golden(case, renderer, oracle_factory):
    scene_name, W, H, feat, _cam, frames, seed = CASES[case]
    g = load_golden(case)
    scene = load_scene(scene_name)
    orc = oracle_factory(); orc.upload_scene(scene); orc.reset_history()
    renderer.upload_scene(scene); renderer.reset_history()
    cam = camera_from_array(g["camera"])
    for fr in range(frames):
        oimg = orc.render_frame(feat, cam, W, H, fr > 0, seed, fr)
        gimg = renderer.render_frame(feat, cam, W, H, fr > 0, seed, fr)
        ggb, ogb = renderer.gbuffer(), orc.gbuffer()
        p = f"f{fr}_"
        assert_bits_equal(ggb.t, ogb.t, f"{case} {p}t")
        assert_bits_equal(ggb.normal, ogb.normal, f"{case} {p}normal")
        assert_bits_equal(ggb.mesh, ogb.mesh, f"{case} {p}mesh")
        if scene.textures:
            assert_bits_equal(ggb.texcoord, ogb.texcoord, f"{case} {p}texcoord")
        for pid in stage_ids(feat, fr):
            gst, ost = renderer.reservoirs(pid), orc.reservoirs(pid)
            compare_stage(case, f"{p}stage {pid}", gst, ost)
            # and directly against the reference's own dump
            assert_bits_equal(gst.M, g[f"{p}s{pid}_M"], f"{case} {p}stage {pid} M vs reference")
            assert_bits_equal(gst.position, g[f"{p}s{pid}_position"], f"{case} {p}stage {pid} position vs reference")
            assert_bits_equal(gst.W, g[f"{p}s{pid}_W"], f"{case} {p}stage {pid} W vs reference")
        assert_rel_close(gimg, oimg, REL, f"{case} {p}image")
        rmse = float(np.sqrt(np.mean((gimg.astype(np.float64) - g[p + "image"]) ** 2)))
        assert rmse <= 1e-3, f"{case} {p}image RMSE vs reference {rmse}"      # north_star image bar
        assert_bits_equal(gimg, g[p + "image"], f"{case} {p}image vs reference")


def test_tracer_matches_oracle_bruteforce(renderer, oracle_factory):
    """closestHit / anyHit: the GPU BVH must agree bit for bit with the oracle's brute-force loop."""
    rng = np.random.default_rng(5)
    for scene_name in ("Monkey", "CornellNightClub", "Cube"):
        scene = load_scene(scene_name)
        orc = oracle_factory(0); orc.upload_scene(scene)
        renderer.upload_scene(scene)
        n = 20000
        verts = np.concatenate([m.vertices["position"] for m in scene.meshes])
        lo, hi = verts.min(0), verts.max(0)
        o = rng.uniform(lo - 1.0, hi + 1.0, size=(n, 3)).astype(np.float32)
        tgt = rng.uniform(lo, hi, size=(n, 3)).astype(np.float32)
        d = tgt - o
        d /= np.linalg.norm(d, axis=1, keepdims=True)
        d = d.astype(np.float32)
        d[:50, 0] = 0.0                                   # axis-parallel rays: 1/0 in the slab test
        d[50:100, 1] = 0.0
        d[100:130, 2] = np.float32(1e-35) * np.where(rng.random(30) < 0.5, -1, 1)   # below the slab test's 2^-100 threshold
        d[130:160, 0] = np.float32(-1e-42)                # subnormal: 1/d overflows
        d[160:180, 1] = np.float32(-0.0)
        d[180:200, 2] = np.float32(2.0 ** -99)            # just above the threshold: 1/d = 2^99
        o[200:230, 0] = verts[rng.integers(0, len(verts), 30), 0]       # origins exactly on a vertex plane, axis-parallel
        d[200:230, 0] = 0.0
        tfar = np.where(rng.random(n) < 0.5, np.float32(3.4e38), rng.uniform(0.1, 5.0, n)).astype(np.float32)
        gh, gt, gu, gv, gtri = renderer.trace_rays(o, d, tfar, any_hit=False)
        oh, ot, ou, ov, otri = orc.trace_rays(o, d, tfar, any_hit=False)
        assert_bits_equal(gh, oh, f"{scene_name} closest hit flag")
        m = oh.astype(bool)
        assert m.sum() > n // 10
        assert_bits_equal(gt[m], ot[m], f"{scene_name} t"); assert_bits_equal(gtri[m], otri[m], f"{scene_name} tri")
        assert_bits_equal(gu[m], ou[m], f"{scene_name} u"); assert_bits_equal(gv[m], ov[m], f"{scene_name} v")
        ga = renderer.trace_rays(o, d, tfar, any_hit=True)[0]
        oa = orc.trace_rays(o, d, tfar, any_hit=True)[0]
        assert_bits_equal(ga, oa, f"{scene_name} any hit")


def test_many_lights_and_edge_sizes(renderer, oracle_factory):
    """Synthetic many-light set (C3-style lights, reduced count) on the monkey, ragged resolution."""
    scene = load_scene("Monkey")
    scene.lights = synthetic_lights(4096, seed=3)
    feat = Features(spatialResamplingPasses=2, initialSamplesVisibilityCheck=True)
    orc = oracle_factory(); orc.upload_scene(scene); orc.reset_history()
    renderer.upload_scene(scene); renderer.reset_history()
    from cases import CORNELL_CAM
    W, H = 37, 29
    cam = CORNELL_CAM.to_abi(W, H)
    for fr in range(2):
        oimg = orc.render_frame(feat, cam, W, H, fr > 0, 5, fr)
        gimg = renderer.render_frame(feat, cam, W, H, fr > 0, 5, fr)
        compare_stage("manylights", f"f{fr} final", renderer.reservoirs(abi.ROMIS_PASS_FINAL), orc.reservoirs(abi.ROMIS_PASS_FINAL))
        assert_rel_close(gimg, oimg, REL, f"manylights f{fr} image")
    # no lights at all: genCanonicalSamples returns early (reference src/scene/light.cpp:46)
    scene.lights = scene.lights[:0]
    orc.upload_scene(scene); renderer.upload_scene(scene)
    oimg = orc.render_frame(feat, cam, W, H, False, 5, 0)
    gimg = renderer.render_frame(feat, cam, W, H, False, 5, 0)
    compare_stage("nolights", "final", renderer.reservoirs(abi.ROMIS_PASS_FINAL), orc.reservoirs(abi.ROMIS_PASS_FINAL))
    assert_bits_equal(gimg, oimg, "nolights image")
    # 1x1 image
    scene = load_scene("CornellNightClub")
    orc.upload_scene(scene); renderer.upload_scene(scene)
    cam1 = NIGHTCLUB_CAM.to_abi(1, 1)
    oimg = orc.render_frame(Features(), cam1, 1, 1, False, 1, 0)
    gimg = renderer.render_frame(Features(), cam1, 1, 1, False, 1, 0)
    assert_bits_equal(gimg, oimg, "1x1 image")


def test_row_bands_reproduce_full_frame(oracle_factory):
    """Two contexts, each a row band with halo rows moved by hand, reproduce the single-context frame
    bit for bit (RNG is keyed by global pixel)."""
    import ctypes as C
    from romis_b200.api import RestirRenderer
    cudart = C.CDLL("libcudart.so")
    scene = load_scene("CornellNightClub")
    W, H = 48, 40
    feat = Features(spatialResamplingPasses=3, spatialResampleRadius=6, initialSamplesVisibilityCheck=True)
    cam = NIGHTCLUB_CAM.to_abi(W, H)
    full = RestirRenderer(0); full.upload_scene(scene)
    bands = [RestirRenderer(0), RestirRenderer(0)]
    split = 17
    bands[0].set_band(0, split); bands[1].set_band(split, H)
    for b in bands:
        b.upload_scene(scene)
    try:
        for fr in range(3):
            ref_img = full.render_frame(feat, cam, W, H, fr > 0, 9, fr)
            img = np.zeros((H, W, 3), np.float32)
            for b in bands:
                b.frame_begin(feat, cam, W, H, fr > 0, 9, fr)
            for p in range(feat.spatialResamplingPasses):
                for b in bands:
                    b.synchronize()
                # band 0 (low rows) sends its top rows up to band 1 and receives band 1's bottom rows
                s_ptr, s_n = bands[0].halo_region(abi.ROMIS_HALO_SEND_HIGH); r_ptr, r_n = bands[1].halo_region(abi.ROMIS_HALO_RECV_LOW)
                assert s_n == r_n and s_n > 0
                assert cudart.cudaMemcpy(C.c_void_p(r_ptr), C.c_void_p(s_ptr), C.c_size_t(s_n), 3) == 0
                s_ptr, s_n = bands[1].halo_region(abi.ROMIS_HALO_SEND_LOW); r_ptr, r_n = bands[0].halo_region(abi.ROMIS_HALO_RECV_HIGH)
                assert s_n == r_n and s_n > 0
                assert cudart.cudaMemcpy(C.c_void_p(r_ptr), C.c_void_p(s_ptr), C.c_size_t(s_n), 3) == 0
                assert bands[0].halo_region(abi.ROMIS_HALO_SEND_LOW)[1] == 0 and bands[1].halo_region(abi.ROMIS_HALO_SEND_HIGH)[1] == 0
                assert cudart.cudaDeviceSynchronize() == 0      # D2D cudaMemcpy is asynchronous w.r.t. the non-blocking context streams
                for b in bands:
                    b.frame_spatial_pass(p)
            for b in bands:
                b.frame_end(img)
            assert_bits_equal(img, ref_img, f"banded frame {fr}")
    finally:
        full.close()
        for b in bands:
            b.close()


def test_error_paths(renderer):
    """C-ABI error behaviour: status codes + message, never an exception across the boundary."""
    from romis_b200.api import RomisError, RestirRenderer
    r = RestirRenderer(0)
    try:
        with pytest.raises(RomisError, match="no scene"):
            r.render_frame(Features(), NIGHTCLUB_CAM, 8, 8, False, 1, 0)
        r.upload_scene(load_scene("Cube"))
        with pytest.raises(RomisError, match="numSamplesInReservoir"):
            r.render_frame(Features(numSamplesInReservoir=0), NIGHTCLUB_CAM, 8, 8, False, 1, 0)
        with pytest.raises(RomisError, match="numNeighboursToSample"):
            r.render_frame(Features(numNeighboursToSample=99), NIGHTCLUB_CAM, 8, 8, False, 1, 0)
        with pytest.raises(RomisError):
            r.frame_spatial_pass(0)
        with pytest.raises(RomisError):
            r.reservoirs(abi.ROMIS_PASS_INITIAL)          # nothing rendered yet
        r.render_frame(Features(), NIGHTCLUB_CAM, 8, 8, False, 1, 0)
        with pytest.raises(RomisError, match="not captured"):
            r.reservoirs(abi.ROMIS_PASS_INITIAL)          # capture is off on this context
        r.reservoirs(abi.ROMIS_PASS_FINAL)
    finally:
        r.close()


def test_light_table_dirty_tracking(renderer, oracle_factory):
    """romis_upload_lights every frame (as the drop-in does): an unchanged table is a no-op; a shorter table keeps the history
    (the reference's reservoirs hold their samples by value, reservoir.h:18-26 -- tests/test_gpu_light_edits.py has the rest)."""
    scene = load_scene("CornellNightClub")
    feat = Features(spatialResamplingPasses=1)
    W, H = 48, 32
    cam = NIGHTCLUB_CAM.to_abi(W, H)
    orc = oracle_factory(); orc.upload_scene(scene); orc.reset_history()
    renderer.upload_scene(scene); renderer.reset_history()
    for fr in range(2):
        renderer.upload_lights(scene.lights)                      # unchanged: must not disturb anything
        oimg = orc.render_frame(feat, cam, W, H, fr > 0, 8, fr)
        gimg = renderer.render_frame(feat, cam, W, H, fr > 0, 8, fr)
        assert_bits_equal(gimg, oimg, f"frame {fr}")
    fewer = scene.lights[:100].copy()
    renderer.upload_lights(fewer); orc.upload_lights(fewer)
    oimg = orc.render_frame(feat, cam, W, H, True, 8, 2)          # history samples of the removed lights live on
    gimg = renderer.render_frame(feat, cam, W, H, True, 8, 2)
    assert_bits_equal(gimg, oimg, "frame after the light table shrank")
    compare_stage("lights", "after shrink", renderer.reservoirs(abi.ROMIS_PASS_FINAL), orc.reservoirs(abi.ROMIS_PASS_FINAL))


@pytest.mark.gpu
def test_shared_reciprocal_division():
    """div3_shared (three divisions by one denominator behind one refined reciprocal, csrc/device_common.cuh) returns the bits of
    the plain IEEE division for every operand: random bit patterns (all exponents, subnormals, infinities, NaNs), the ranges the
    shading produces, exact zeros of both signs, operands on the edges of the fast path's exponent window, near-overflow quotients."""
    from romis_b200.api import RestirRenderer
    rng = np.random.default_rng(11)
    r = RestirRenderer(0)
    n = 1 << 22

    def check(num, den, what):
        fast, ref = r.selftest_division(num, den)
        same = fast.view(np.uint32) == ref.view(np.uint32)
        both_nan = np.isnan(fast) & np.isnan(ref)
        bad = ~(same | both_nan)
        assert not bad.any(), f"{what}: {int(bad.sum())} of {bad.size} quotients differ, first at {np.argwhere(bad)[0]}"

    bits = lambda k: rng.integers(0, 1 << 32, size=k, dtype=np.uint64).astype(np.uint32).view(np.float32)
    check(bits(3 * n).reshape(n, 3), bits(n), "random bit patterns")
    check(bits(3 * n).reshape(n, 3), np.abs(bits(n)), "random bit patterns, positive denominators")
    # what computeShading divides: radiance-like numerators (some channels exactly zero), squared distances
    num = (rng.random((n, 3), dtype=np.float32) ** 4 * np.float32(50.0)).astype(np.float32)
    num[rng.random((n, 3)) < 0.2] = 0.0
    num[rng.random((n, 3)) < 0.01] = -0.0
    den = (rng.random(n, dtype=np.float32) * np.float32(30.0) + np.float32(1e-4)).astype(np.float32) ** 2
    check(num, den, "shading range")
    check(-num, den, "shading range, negative numerators")
    # exponent edges of the fast path (2^-60 .. 2^60 numerators, 2^-40 .. 2^40 denominators), a binade either side
    e_num = rng.integers(-63, 64, size=(n, 3)); e_den = rng.integers(-43, 44, size=n)
    mant = lambda shape: (1.0 + rng.random(shape)).astype(np.float32)
    check(np.ldexp(mant((n, 3)), e_num).astype(np.float32), np.ldexp(mant(n), e_den).astype(np.float32), "window edges")
    # quotients next to overflow / underflow
    e_num = rng.integers(60, 128, size=(n, 3)); e_den = rng.integers(-60, -30, size=n)
    check(np.ldexp(mant((n, 3)), e_num).astype(np.float32), np.ldexp(mant(n), e_den).astype(np.float32), "towards overflow")
    check(np.ldexp(mant((n, 3)), -e_num).astype(np.float32), np.ldexp(mant(n), -e_den).astype(np.float32), "towards underflow")
    r.close()
