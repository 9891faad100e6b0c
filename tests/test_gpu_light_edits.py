"""Light edits between frames with temporal reuse on.

The reference's reservoirs hold LightSample{position, color} by value (reference src/rendering/reservoir.h:18-26) and
temporalReuse streams the predecessor's samples as stored (src/rendering/render_utils.cpp:154-170): a history sample keeps the
position / colour its light had when it was drawn, whatever the UI (src/ui/ui.cpp:172-261) does to scene.lights afterwards.
The CUDA path stores (light, u, v) and keeps the OLD record of every edited / removed light in a device-side archive
(romis_upload_lights).  Checked here against the oracle (which holds samples by value like the reference) stage by stage, and
against the compiled reference's renderReSTIR through the drop-in."""
import ctypes as C
import os

import numpy as np
import pytest

from oracle import pyoracle
from romis_b200 import abi
from romis_b200.scene import Features, synthetic_lights
from cases import NIGHTCLUB_CAM
from common import assert_bits_equal, load_scene, stage_ids

pytestmark = pytest.mark.gpu


def edit_schedule(base):
    """frame -> (light table, dirty range or None): move / recolour, nothing, shrink, grow, replace everything, range edit."""
    t = {0: (base.copy(), None), 1: (base.copy(), None)}
    a = base.copy(); a["p0"][3] += np.float32(0.25); a["c0"][7] *= np.float32(0.5); a["c2"][7] *= np.float32(2.0)
    t[2] = (a, None)
    t[3] = (a.copy(), None)                                     # unchanged
    t[4] = (a[:300].copy(), None)                               # 212 lights removed
    g = np.concatenate([a[:300], base[100:200]]); g["p0"][300:] += np.float32(0.1)
    t[5] = (g, None)                                            # 100 lights added
    r = g.copy(); r["p0"] += np.float32(0.05); r["c1"] *= np.float32(0.9)
    t[6] = (r, None)                                            # the whole table moves
    e = r.copy(); e["e1"][10:13] *= np.float32(1.5)
    t[7] = (e, (10, 3))                                         # caller-supplied dirty range
    b = e.copy(); b["p0"][3] = base["p0"][3]                    # a light goes back to where it once was
    t[8] = (b, (3, 1))
    t[9] = (b.copy(), (0, 0))
    return t


def test_light_edits_keep_history_samples_as_drawn(oracle_factory):
    from romis_b200.api import RestirRenderer
    scene = load_scene("CornellNightClub")
    feat = Features(spatialResamplingPasses=2, initialSamplesVisibilityCheck=True)
    W, H = 64, 40
    cam = NIGHTCLUB_CAM.to_abi(W, H)
    r = RestirRenderer(0); r.set_capture(True); r.upload_scene(scene)
    orc = oracle_factory(); orc.upload_scene(scene); orc.reset_history()
    try:
        for fr, (lights, dirty) in sorted(edit_schedule(scene.lights).items()):
            r.upload_lights(lights, dirty); orc.upload_lights(lights)
            oimg = orc.render_frame(feat, cam, W, H, fr > 0, 77, fr)
            gimg = r.render_frame(feat, cam, W, H, fr > 0, 77, fr)
            for pid in stage_ids(feat, fr):
                g, o = r.reservoirs(pid), orc.reservoirs(pid)
                for fld in ("light_id", "M", "u", "v", "W", "position", "color"):
                    assert_bits_equal(getattr(g, fld), getattr(o, fld), f"frame {fr} stage {pid} {fld}")
            assert_bits_equal(gimg, oimg, f"frame {fr} image")
        n_slots, held = r.light_archive_size()
        assert 0 < held <= n_slots <= 2 * W * H * 2, (n_slots, held)
    finally:
        r.close()


def test_archive_slots_are_recycled(oracle_factory):
    """Every light edited every frame: the archive stays bounded by what the history can still hold."""
    from romis_b200.api import RestirRenderer
    scene = load_scene("CornellNightClub")
    feat = Features(spatialResamplingPasses=1, numSamplesInReservoir=1)
    W, H = 32, 20
    cam = NIGHTCLUB_CAM.to_abi(W, H)
    r = RestirRenderer(0); r.upload_scene(scene)
    orc = oracle_factory(); orc.upload_scene(scene); orc.reset_history()
    try:
        lights = scene.lights.copy()
        for fr in range(12):
            lights["c0"] *= np.float32(0.99); lights["p0"][:, 1] += np.float32(0.01)
            r.upload_lights(lights); orc.upload_lights(lights)
            oimg = orc.render_frame(feat, cam, W, H, fr > 0, 5, fr)
            gimg = r.render_frame(feat, cam, W, H, fr > 0, 5, fr)
            assert_bits_equal(gimg, oimg, f"frame {fr} image")
        n_slots, held = r.light_archive_size()
        assert n_slots <= W * H + 2 * len(lights), (n_slots, held)          # held <= one per history record, + the last edit's batch
        r.reset_history()
        r.upload_lights(scene.lights)                                       # no history: nothing to keep
        assert r.light_archive_size()[1] == 0
    finally:
        r.close()


def test_many_lights_edit(oracle_factory):
    """65 536 lights, a block of them edited, then the table grows past its capacity."""
    from romis_b200.api import RestirRenderer
    scene = load_scene("Monkey")
    scene.lights = synthetic_lights(65536, seed=4)
    feat = Features(spatialResamplingPasses=1, initialSamplesVisibilityCheck=True)
    W, H = 48, 40
    from cases import CORNELL_CAM
    cam = CORNELL_CAM.to_abi(W, H)
    r = RestirRenderer(0); r.upload_scene(scene)
    orc = oracle_factory(); orc.upload_scene(scene); orc.reset_history()
    try:
        lights = scene.lights.copy()
        for fr in range(4):
            if fr == 1:
                lights["c0"][1000:3000] *= np.float32(0.5)
            if fr == 2:
                lights = np.concatenate([lights, synthetic_lights(40000, seed=9)])
            if fr == 3:
                lights = lights[:50000].copy()
            r.upload_lights(lights); orc.upload_lights(lights)
            oimg = orc.render_frame(feat, cam, W, H, fr > 0, 11, fr)
            gimg = r.render_frame(feat, cam, W, H, fr > 0, 11, fr)
            g, o = r.reservoirs(abi.ROMIS_PASS_FINAL), orc.reservoirs(abi.ROMIS_PASS_FINAL)
            for fld in ("light_id", "M", "u", "v", "W", "position", "color"):
                assert_bits_equal(getattr(g, fld), getattr(o, fld), f"frame {fr} final {fld}")
            assert_bits_equal(gimg, oimg, f"frame {fr} image")
    finally:
        r.close()


def test_banded_contexts_share_one_archive(oracle_factory):
    """Two row bands of one frame (two contexts on one GPU, halos copied by hand) with light edits: the hosts OR the archive
    marks of the bands and release the same slots on both (include/romis_gpu.h), so halo rows keep meaning the same lights."""
    from romis_b200.api import RestirRenderer
    cudart = C.CDLL("libcudart.so")
    scene = load_scene("CornellNightClub")
    feat = Features(spatialResamplingPasses=2, spatialResampleRadius=6)
    W, H = 56, 48
    cam = NIGHTCLUB_CAM.to_abi(W, H)
    orc = oracle_factory(); orc.upload_scene(scene); orc.reset_history()
    edges = [0, 22, H]
    bands = [RestirRenderer(0) for _ in range(2)]
    for i, b in enumerate(bands):
        b.set_band(edges[i], edges[i + 1]); b.upload_scene(scene); b.set_light_archive_auto(False)
    try:
        lights = scene.lights.copy()
        for fr in range(7):
            if fr >= 2:
                lights["c0"] *= np.float32(0.97); lights["p0"][fr::7, 0] += np.float32(0.02)
            marks = [b.light_archive_marks() for b in bands]
            n = max(len(m) for m in marks)
            keep = np.zeros(n, np.uint8)
            for m in marks:
                keep[:len(m)] |= m
            for b in bands:
                b.light_archive_release(keep)
                b.upload_lights(lights)
            orc.upload_lights(lights)
            oimg = orc.render_frame(feat, cam, W, H, fr > 0, 9, fr)
            img = np.zeros((H, W, 3), np.float32)
            for b in bands:
                b.frame_begin(feat, cam, W, H, fr > 0, 9, fr)
            for p in range(feat.spatialResamplingPasses):
                for b in bands:
                    b.synchronize()
                s, n1 = bands[0].halo_region(abi.ROMIS_HALO_SEND_HIGH); d, m1 = bands[1].halo_region(abi.ROMIS_HALO_RECV_LOW)
                assert n1 == m1 > 0 and cudart.cudaMemcpy(C.c_void_p(d), C.c_void_p(s), C.c_size_t(n1), 3) == 0
                s, n1 = bands[1].halo_region(abi.ROMIS_HALO_SEND_LOW); d, m1 = bands[0].halo_region(abi.ROMIS_HALO_RECV_HIGH)
                assert n1 == m1 > 0 and cudart.cudaMemcpy(C.c_void_p(d), C.c_void_p(s), C.c_size_t(n1), 3) == 0
                assert cudart.cudaDeviceSynchronize() == 0
                for b in bands:
                    b.frame_spatial_pass(p)
            for b in bands:
                b.frame_end(img)
            assert_bits_equal(img, oimg, f"banded frame {fr} with edited lights vs oracle")
            assert bands[0].light_archive_size()[0] == bands[1].light_archive_size()[0]
    finally:
        for b in bands:
            b.close()


@pytest.mark.skipif(not os.path.exists(pyoracle.DROPIN_SO), reason="oracle/_ref/libromis_dropin.so not built (make -C oracle dropin)")
def test_dropin_light_edit_matches_the_reference_screen():
    """The reference's own Scene with a light moved / recoloured / removed between frames: Screen::pixels() of the GPU drop-in
    equals the compiled reference's renderReSTIR bit for bit, frame after frame."""
    lib = pyoracle.DropinLib()
    scene = load_scene("CornellNightClub")
    lib.set_scene(scene)
    lib.reset_history()
    feat = Features(spatialResamplingPasses=2, initialSamplesVisibilityCheck=True)
    W, H = 72, 44
    lights = scene.lights.copy()
    for fr in range(6):
        if fr == 2:
            lights["c0"][5] = np.float32([4.0, 0.5, 0.5]); lights["p0"][9] += np.float32(0.3)
        if fr == 3:
            lights["c0"][:] *= np.float32(0.8)
        if fr == 4:
            lights = lights[:256].copy()
        lib.set_lights(lights)
        cpu = lib.render_frame(feat, NIGHTCLUB_CAM, W, H, fr > 0, 2718, fr, pyoracle.REF_FLAG_WHOLE_FRAME, dump=False).image
        gpu = lib.render_frame_gpu(feat, NIGHTCLUB_CAM, W, H, fr > 0, 2718, fr)
        assert_bits_equal(gpu, cpu, f"frame {fr}: Screen::pixels() of the GPU drop-in vs the reference's renderReSTIR after light edits")
