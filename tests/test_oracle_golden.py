"""The C restatement (oracle/restir_oracle.c) against the golden vectors generated from the REFERENCE
itself (tests/golden/gen_golden.py -> oracle/_ref/libromis_ref.so).  Bit-exact on every dumped field:
G-buffer, per-stage LightSample position / colour, outputWeight, sampleNums, wSums, final image.
This is what pins the oracle (DESIGN.md "oracle")."""
import pytest

from cases import CASES
from common import assert_bits_equal, camera_from_array, load_golden, load_scene, stage_ids


@pytest.mark.parametrize("case", sorted(CASES))
def test_oracle_matches_reference_golden(case, oracle_factory):
    scene_name, W, H, feat, _cam, frames, seed = CASES[case]
    g = load_golden(case)
    orc = oracle_factory()
    orc.upload_scene(load_scene(scene_name))
    orc.reset_history()
    cam = camera_from_array(g["camera"])
    for fr in range(frames):
        img = orc.render_frame(feat, cam, W, H, fr > 0, seed, fr)
        gb = orc.gbuffer()
        p = f"f{fr}_"
        assert_bits_equal(gb.t, g[p + "t"], f"{case} {p}t")
        assert_bits_equal(gb.normal, g[p + "normal"], f"{case} {p}normal")
        assert_bits_equal(gb.mesh, g[p + "mesh"], f"{case} {p}mesh")
        assert_bits_equal(gb.texcoord, g[p + "texcoord"], f"{case} {p}texcoord")
        for pid in stage_ids(feat, fr):
            st = orc.reservoirs(pid)
            for fld in ("position", "color", "W", "M", "wSum"):
                assert_bits_equal(getattr(st, fld), g[f"{p}s{pid}_{fld}"], f"{case} {p}stage {pid} {fld}")
        assert_bits_equal(img, g[p + "image"], f"{case} {p}image")
