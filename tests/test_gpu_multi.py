"""Row bands on SEVERAL GPUs (one process per GPU under torchrun, peer-mapped halos with the exchange fused into the spatial
pass) reproduce the single-GPU frame bit for bit.  Skipped on a box with one GPU; the CPU-side plumbing is covered by
tests/test_bands_gloo.py and the band arithmetic by the single-GPU band tests (tests/test_gpu_full_size.py)."""
import os
import subprocess
import sys

import pytest

pytestmark = pytest.mark.gpu
ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))


def _run(nproc, extra, port):
    import torch
    if torch.cuda.device_count() < nproc:
        pytest.skip(f"needs {nproc} GPUs, this box has {torch.cuda.device_count()}")
    cmd = [sys.executable, "-m", "torch.distributed.run", "--nnodes=1", "--nproc-per-node", str(nproc), "--master-addr", "127.0.0.1",
           "--master-port", str(port), os.path.join(ROOT, "tools", "check_bands_multi_gpu.py")] + extra
    r = subprocess.run(cmd, cwd=ROOT, capture_output=True, text=True, timeout=900)
    assert r.returncode == 0, r.stdout[-3000:] + r.stderr[-3000:]
    assert "== single GPU: True" in r.stdout and "False" not in r.stdout, r.stdout[-3000:]


def test_two_gpu_bands_equal_the_single_gpu_frame():
    _run(2, ["--width", "960", "--height", "540", "--frames", "4"], 29541)


def test_two_gpu_bands_with_light_edits():
    _run(2, ["--width", "640", "--height", "360", "--frames", "5", "--edit-lights"], 29542)


def test_two_gpu_bands_equal_rows_and_nccl_transport():
    _run(2, ["--width", "640", "--height", "360", "--frames", "3", "--equal-rows", "--halo", "nccl"], 29543)


def test_four_gpu_bands_c3_many_lights():
    _run(4, ["--width", "768", "--height", "432", "--frames", "3", "--config", "c3"], 29544)


# ---- one process, several devices: romis_create(device_ids, n_devices > 1) ----
def _devices(n):
    """n distinct GPUs when the box has them, else the same GPU n times (two bands on one device: same code path, the
    neighbours' buffers are then reached without peer mapping)."""
    import torch
    have = torch.cuda.device_count()
    return list(range(n)) if have >= n else [0] * n


@pytest.mark.parametrize("n", [2, 3])
def test_multi_device_context_equals_single_device_frame(n, oracle_factory):
    """The reference's caller is one thread of one process (main.cpp:164): one context over n devices renders the frame as n row
    bands into the caller's single image, bit-identical to the one-device frame and to the oracle, light edits included."""
    import numpy as np
    from romis_b200 import abi
    from romis_b200.api import RestirRenderer
    from romis_b200.scene import Features
    from cases import NIGHTCLUB_CAM
    from common import assert_bits_equal, load_scene
    scene = load_scene("CornellNightClub")
    feat = Features(spatialResamplingPasses=3, initialSamplesVisibilityCheck=True)
    W, H = 160, 120
    cam = NIGHTCLUB_CAM.to_abi(W, H)
    multi = RestirRenderer(_devices(n)); multi.upload_scene(scene)
    orc = oracle_factory(); orc.upload_scene(scene); orc.reset_history()
    lights = scene.lights.copy()
    try:
        for fr in range(5):
            if fr >= 2:
                lights["c0"][fr::3] *= np.float32(0.9); lights["p0"][fr::7, 2] += np.float32(0.03)
                multi.upload_lights(lights); orc.upload_lights(lights)
            oimg = orc.render_frame(feat, cam, W, H, fr > 0, 31, fr)
            gimg = multi.render_frame(feat, cam, W, H, fr > 0, 31, fr)
            assert_bits_equal(gimg, oimg, f"{n}-device context, frame {fr} image vs oracle")
            g, o = multi.reservoirs(abi.ROMIS_PASS_FINAL), orc.reservoirs(abi.ROMIS_PASS_FINAL)
            for fld in ("light_id", "M", "u", "v", "W"):
                assert_bits_equal(getattr(g, fld), getattr(o, fld), f"{n}-device context, frame {fr} final {fld}")
        t = multi.timings()
        assert t.total_ms > 0 and t.n_launches > 0
    finally:
        multi.close()


def test_multi_device_context_rejects_per_device_calls():
    from romis_b200.api import RestirRenderer, RomisError
    from romis_b200.scene import Features
    from cases import NIGHTCLUB_CAM
    from common import load_scene
    multi = RestirRenderer(_devices(2)); multi.upload_scene(load_scene("Cube"))
    try:
        for call in (lambda: multi.set_band(0, 8), lambda: multi.frame_begin(Features(), NIGHTCLUB_CAM, 16, 16, False, 1, 0),
                     lambda: multi.stream()):
            with pytest.raises(RomisError, match="multi-device"):
                call()
        multi.render_frame(Features(), NIGHTCLUB_CAM, 32, 32, False, 1, 0)     # tiny frame, radius 10: fewer bands than devices
    finally:
        multi.close()


@pytest.mark.parametrize("n", [2, 3])
def test_multi_device_context_renders_mis_frames_and_keeps_the_restir_history(n):
    """renderRMIS / renderROMIS through one context over n devices (one row band per device, halo rows rendered by the band itself)
    equal the one-device frames bit for bit; and as in the reference, where previousFrameGrid outlives an R-MIS / R-OMIS frame
    (render.cpp:268-280, main.cpp:164-165), a ReSTIR sequence interrupted by such frames continues from its temporal history."""
    import numpy as np
    from romis_b200 import abi
    from romis_b200.api import RestirRenderer
    from romis_b200.scene import Features, RmisParams, synthetic_lights
    from cases import NIGHTCLUB_CAM
    from common import assert_bits_equal, load_scene
    scene = load_scene("CornellNightClub"); scene.lights = synthetic_lights(512, seed=3)
    W, H = 176, 99
    cam = NIGHTCLUB_CAM.to_abi(W, H)
    feat = Features(spatialResamplingPasses=2, initialSamplesVisibilityCheck=True)
    rp = RmisParams(maxIterationsMIS=2, misWeightRMIS=abi.ROMIS_MIS_BALANCE)
    renderer = RestirRenderer(0); renderer.upload_scene(scene)
    multi = RestirRenderer(_devices(n)); multi.upload_scene(scene)
    try:
        # before any ReSTIR frame: the context cuts bands for the R-MIS frame itself
        assert_bits_equal(multi.render_frame_rmis(feat, rp, cam, W, H, 9, 0), renderer.render_frame_rmis(feat, rp, cam, W, H, 9, 0), "rmis, fresh context")
        for fr in range(5):
            if fr in (2, 4):      # laid out for ReSTIR by now: same bands, history untouched
                assert_bits_equal(multi.render_frame_rmis(feat, rp, cam, W, H, 9, fr), renderer.render_frame_rmis(feat, rp, cam, W, H, 9, fr), f"rmis before frame {fr}")
                assert_bits_equal(multi.render_frame_romis(feat, rp, cam, W, H, 9, fr), renderer.render_frame_romis(feat, rp, cam, W, H, 9, fr), f"romis before frame {fr}")
            a = multi.render_frame(feat, cam, W, H, fr > 0, 31, fr)
            b = renderer.render_frame(feat, cam, W, H, fr > 0, 31, fr)
            assert_bits_equal(a, b, f"ReSTIR frame {fr} after R-MIS / R-OMIS frames")
            g, o = multi.reservoirs(abi.ROMIS_PASS_FINAL), renderer.reservoirs(abi.ROMIS_PASS_FINAL)
            for fld in ("light_id", "M", "W"):
                assert_bits_equal(getattr(g, fld), getattr(o, fld), f"frame {fr} final {fld}")
        # another radius than the layout was made for: own bands for the R-MIS frame, ReSTIR restarts
        feat2 = Features(spatialResamplingPasses=2, initialSamplesVisibilityCheck=True, spatialResampleRadius=14)
        assert_bits_equal(multi.render_frame_romis(feat2, rp, cam, W, H, 9, 7), renderer.render_frame_romis(feat2, rp, cam, W, H, 9, 7), "romis, radius 14")
        assert_bits_equal(multi.render_frame(feat, cam, W, H, False, 31, 8), renderer.render_frame(feat, cam, W, H, False, 31, 8), "ReSTIR after re-layout")
    finally:
        multi.close(); renderer.close()


@pytest.mark.skipif(not os.path.exists(os.path.join(ROOT, "oracle", "_ref", "libromis_dropin.so")), reason="oracle/_ref/libromis_dropin.so not built")
def test_dropin_on_two_devices_fills_the_reference_screen():
    """integration/render_restir_gpu.cpp with ROMIS_DEVICES=a,b (own process: the drop-in creates its context once per thread):
    Screen::pixels() equals the compiled reference's renderReSTIR bit for bit; so does renderRMIS through the same context, and
    the ReSTIR sequence continues from its history afterwards as the reference's does."""
    devs = ",".join(str(d) for d in _devices(2))
    code = ("import sys; sys.path[:0] = ['tests', 'tests/golden']\n"
            "from oracle import pyoracle; from romis_b200.scene import Features; from cases import NIGHTCLUB_CAM; from common import assert_bits_equal, load_scene\n"
            "lib = pyoracle.DropinLib(); lib.set_scene(load_scene('CornellNightClub')); lib.reset_history()\n"
            "feat = Features(spatialResamplingPasses=3, initialSamplesVisibilityCheck=True)\n"
            "for fr in range(3):\n"
            "    cpu = lib.render_frame(feat, NIGHTCLUB_CAM, 96, 80, fr > 0, 314, fr, pyoracle.REF_FLAG_WHOLE_FRAME, dump=False).image\n"
            "    gpu = lib.render_frame_gpu(feat, NIGHTCLUB_CAM, 96, 80, fr > 0, 314, fr)\n"
            "    assert_bits_equal(gpu, cpu, f'frame {fr}')\n"
            "from romis_b200.scene import RmisParams\n"
            "rp = RmisParams(maxIterationsMIS=2)\n"
            "cpu = lib.render_frame_rmis(feat, rp, NIGHTCLUB_CAM, 96, 80, 2718, 1, False)[0]\n"
            "assert_bits_equal(lib.render_frame_mis_gpu(False, feat, rp, NIGHTCLUB_CAM, 96, 80, 2718, 1), cpu, 'renderRMIS on the device group')\n"
            "cpu = lib.render_frame(feat, NIGHTCLUB_CAM, 96, 80, True, 314, 3, pyoracle.REF_FLAG_WHOLE_FRAME, dump=False).image\n"
            "assert_bits_equal(lib.render_frame_gpu(feat, NIGHTCLUB_CAM, 96, 80, True, 314, 3), cpu, 'ReSTIR frame 3 after the R-MIS frame: history kept')\n"
            "print('dropin on devices " + devs + " ok')\n")
    r = subprocess.run([sys.executable, "-c", code], cwd=ROOT, capture_output=True, text=True, timeout=600, env=dict(os.environ, ROMIS_DEVICES=devs))
    assert r.returncode == 0 and "ok" in r.stdout, r.stdout[-2000:] + r.stderr[-3000:]
