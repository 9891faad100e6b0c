"""Size-independent properties of the ReSTIR frame, checked on the oracle (CPU) -- the same properties are
checked on the CUDA path at BASELINE's full size in test_gpu_full_size.py."""
import numpy as np

from romis_b200 import abi
from romis_b200.scene import Features
from cases import NIGHTCLUB_CAM
from common import FLT_MAX, load_scene


def check_frame_invariants(gb, stages, feat, n_lights, prev_total=None):
    """stages: {pass_id: ReservoirState}.  Returns total M of the final stage."""
    miss = gb.t == FLT_MAX
    N, Mc = feat.numSamplesInReservoir, feat.initialLightSamples
    ini = stages[abi.ROMIS_PASS_INITIAL]
    tot = ini.M.astype(np.int64).sum(0)
    assert (tot == Mc).all(), "initial RIS must process exactly M candidates per pixel (light.cpp:63)"
    assert (ini.M[0][miss] == Mc).all() and (ini.W[:, miss] == 0).all(), "miss pixels: sub-reservoir 0 takes all, W = 0 (SURVEY A.3/A.4)"
    for pid, st in stages.items():
        assert np.isfinite(st.W).all() and (st.W >= 0).all(), f"stage {pid}: W finite and non-negative"
        ok = (st.light_id < n_lights) | (st.light_id == 0xFFFFFFFF)
        assert ok.all(), f"stage {pid}: light index in range"
        assert ((st.u >= 0) & (st.u <= 1) & (st.v >= 0) & (st.v <= 1)).all()
        assert (st.W[st.light_id == 0xFFFFFFFF] == 0).all(), "a sub-reservoir without a sample has W = 0"
    if abi.ROMIS_PASS_TEMPORAL in stages and prev_total is not None:
        tmp = stages[abi.ROMIS_PASS_TEMPORAL].M.astype(np.int64).sum(0)
        cap = feat.temporalClampM * Mc + 1
        assert (tmp >= Mc).all() and (tmp <= Mc + N * cap).all(), "temporal M = current + clamped predecessor (render_utils.cpp:156-163)"
    if not feat.unbiasedCombination and feat.spatialReuse:
        # biased spatial reuse rejects every neighbour of a miss pixel, so its M only carries over
        before = stages.get(abi.ROMIS_PASS_TEMPORAL, ini).M.astype(np.int64).sum(0)
        after = stages[abi.ROMIS_PASS_SPATIAL0].M.astype(np.int64).sum(0)
        assert (after[miss] == before[miss]).all()
        assert (after >= before).all(), "merging never loses samples"
    return stages[abi.ROMIS_PASS_FINAL].M.astype(np.int64).sum(0)


def test_oracle_invariants_and_determinism(oracle_factory):
    scene = load_scene("CornellNightClub")
    feat = Features(spatialResamplingPasses=2, initialSamplesVisibilityCheck=True)
    W, H = 64, 40
    cam = NIGHTCLUB_CAM.to_abi(W, H)
    imgs = []
    for run in range(2):
        orc = oracle_factory(); orc.upload_scene(scene); orc.reset_history()
        prev = None
        for fr in range(3):
            img = orc.render_frame(feat, cam, W, H, fr > 0, 77, fr)
            ids = [abi.ROMIS_PASS_INITIAL] + ([abi.ROMIS_PASS_TEMPORAL] if fr else []) + [abi.ROMIS_PASS_SPATIAL0, abi.ROMIS_PASS_SPATIAL0 + 1, abi.ROMIS_PASS_FINAL]
            prev = check_frame_invariants(orc.gbuffer(), {i: orc.reservoirs(i) for i in ids}, feat, len(scene.lights), prev)
            assert np.isfinite(img).all() and (img >= 0).all() and (img <= 1).all(), "tone-mapped image lies in [0, 1]"
        imgs.append(img)
    assert np.array_equal(imgs[0].view(np.uint32), imgs[1].view(np.uint32)), "same seed, same frame index -> same bits"


def test_oracle_bvh_equals_bruteforce(oracle_factory):
    """The oracle's own BVH and its brute-force loop agree bit for bit (order independence of the tracer rules)."""
    rng = np.random.default_rng(1)
    for name in ("Monkey", "CornellNightClub"):
        scene = load_scene(name)
        a, b = oracle_factory(0), oracle_factory(1)
        a.upload_scene(scene); b.upload_scene(scene)
        n = 20000
        o = rng.uniform(-4, 4, (n, 3)).astype(np.float32); d = rng.normal(size=(n, 3)).astype(np.float32)
        d /= np.linalg.norm(d, axis=1, keepdims=True)
        tf = np.full(n, 3.0e38, np.float32)
        ra, rb = a.trace_rays(o, d, tf), b.trace_rays(o, d, tf)
        assert ra[0].sum() > 100
        for x, y in zip(ra, rb):
            assert np.array_equal(x.view(np.uint8), y.view(np.uint8))
        assert np.array_equal(a.trace_rays(o, d, tf, True)[0], b.trace_rays(o, d, tf, True)[0])
