"""Host-side logic: parameter / camera / scene marshalling, row-band partitioning, synthetic lights."""
import os
import tempfile

import numpy as np
import pytest

from romis_b200 import abi
from romis_b200.bands import balanced_band_edges, band_rows, check_bands
from romis_b200.scene import Camera, Features, LIGHT_DTYPE, Scene, synthetic_lights
from cases import CASES
from common import load_golden, load_scene


def test_features_defaults_mirror_reference():
    f = Features()                               # reference src/utils/common.h:103-136
    assert (f.numSamplesInReservoir, f.initialLightSamples, f.numNeighboursToSample, f.spatialResampleRadius) == (2, 32, 5, 10)
    assert (f.spatialResamplingPasses, f.temporalClampM, f.gamma, f.exposure) == (2, 20, 1.0, 1.5)
    assert f.enableShading and f.spatialReuse and f.temporalReuse and f.enableToneMapping
    assert not (f.unbiasedCombination or f.spatialReuseVisibilityCheck or f.initialSamplesVisibilityCheck)
    a = f.to_abi()
    assert a.initialLightSamples == 32 and a.exposure == 1.5 and a.unbiasedCombination == 0


@pytest.mark.parametrize("case", ["nightclub_default", "cornell_c1"])
def test_camera_marshalling_matches_reference_trackball(case):
    """Camera.to_abi restates Trackball::position / glm::quat(euler) (trackball.cpp:75-78, type_quat.inl:208-217); the golden
    camera was computed by the reference's own Trackball.  numpy's sin/cos may differ from glibc's in the last bit."""
    _s, W, H, _f, cam, _n, _seed = CASES[case]
    ref = load_golden(case)["camera"]
    c = cam.to_abi(W, H)
    mine = np.array([*c.origin, *c.quat, c.half_width, c.half_height], np.float32)
    assert np.allclose(mine, ref, rtol=2e-6, atol=2e-6)


def test_scene_roundtrip_and_fixture_contents():
    s = load_scene("CornellNightClub")
    assert len(s.meshes) == 18 and s.n_triangles == 166 and len(s.lights) == 512
    assert (s.lights["type"] == abi.ROMIS_LIGHT_PARALLELOGRAM).all()
    with tempfile.TemporaryDirectory() as d:
        p = os.path.join(d, "s.npz"); s.save(p); t = Scene.load(p)
    assert len(t.meshes) == 18 and np.array_equal(t.lights, s.lights)
    assert all(np.array_equal(a.vertices, b.vertices) and np.array_equal(a.triangles, b.triangles) for a, b in zip(s.meshes, t.meshes))
    assert load_scene("Monkey").n_triangles == 968 and len(load_scene("CubeTextured").textures) == 1
    descs, nm, texs, nt, keep = s.to_abi()
    assert nm == 18 and descs[0].n_triangles == len(s.meshes[0].triangles) and descs[3].material.shininess == 250.0


def test_synthetic_lights_are_seed_fixed():
    a, b = synthetic_lights(1000, seed=4), synthetic_lights(1000, seed=4)
    assert a.dtype == LIGHT_DTYPE and np.array_equal(a, b)
    assert not np.array_equal(a, synthetic_lights(1000, seed=5))
    r = np.linalg.norm(a["p0"], axis=1)
    assert (r >= 1.49).all() and (r <= 3.01).all()
    assert (a["type"] == abi.ROMIS_LIGHT_POINT).sum() == 500
    assert np.abs(a["e1"]).max() <= 0.05


def test_band_partition():
    for H in (1, 7, 1080, 2160):
        for G in (1, 2, 3, 4, 8):
            rows = [band_rows(H, G, r) for r in range(G)]
            assert rows[0][0] == 0 and rows[-1][1] == H
            assert all(rows[i][1] == rows[i + 1][0] for i in range(G - 1))
            sizes = [b - a for a, b in rows]
            assert max(sizes) - min(sizes) <= 1
    check_bands(1080, 8, 30)
    with pytest.raises(ValueError):
        check_bands(64, 8, 10)


def test_equal_cost_band_edges():
    """balanced_band_edges: contiguous, deterministic, each band >= min_rows, costs nearly equal."""
    rng = np.random.default_rng(2)
    for H, G, r in ((1080, 8, 10), (2160, 8, 30), (360, 2, 10), (97, 3, 7)):
        cost = rng.uniform(0.0, 1.0, H) ** 3 * 1000 + 40.0
        cost[: H // 5] = 40.0                         # a stretch of cheap (all-miss) rows
        e = balanced_band_edges(cost, G, r)
        assert e == balanced_band_edges(cost, G, r)
        assert e[0] == 0 and e[-1] == H and len(e) == G + 1
        sizes = np.diff(e)
        assert (sizes >= r).all()
        band_cost = np.add.reduceat(cost, e[:-1])
        assert band_cost.max() <= band_cost.mean() * 1.25
    # degenerate: everything in one row-range still leaves room for every band
    cost = np.zeros(100); cost[95:] = 1.0
    e = balanced_band_edges(cost, 4, 10)
    assert (np.diff(e) >= 10).all() and e[-1] == 100
    with pytest.raises(ValueError):
        balanced_band_edges(np.ones(30), 4, 10)


def test_refine_band_edges_converges_with_latency_floor():
    """Band times follow t = a + sum(row cost) with a large fixed part `a` (the thin-band latency floor): the refinement
    must still converge to near-equal times (the naive density t / rows over-corrects), and keep every band >= min_rows."""
    from romis_b200.bands import refine_band_edges
    rng = np.random.default_rng(3)
    H, world, a = 1080, 8, 0.25
    y = np.arange(H)
    true_cost = (0.2 + np.exp(-((y - 600) / 250.0) ** 2)) * (1.0 + 0.1 * rng.random(H))
    true_cost *= 2.2 / true_cost.sum()
    profile = np.maximum(true_cost * (1.0 + 0.3 * np.sin(y / 40.0)), 1e-3)      # the hit profile only resembles the true cost
    edges = [g * H // world for g in range(world)] + [H]
    times_of = lambda e: [a + true_cost[e[g]:e[g + 1]].sum() for g in range(world)]
    spread0 = max(times_of(edges)) / np.mean(times_of(edges))
    for _ in range(6):
        edges = refine_band_edges(edges, times_of(edges), profile, 10)
    t = times_of(edges)
    assert max(t) / np.mean(t) < 1.03 < spread0
    assert min(np.diff(edges)) >= 10 and edges[0] == 0 and edges[-1] == H


def test_balance_over_a_camera_path_cuts_for_the_mean_profile():
    """BASELINE C4 (orbiting camera) on several GPUs: the bands are cut once for the mean hit profile over cameras sampled along
    the path, so that no row changes owner (and loses its temporal history) mid-sequence."""
    import torch
    from romis_b200.bands import BandedRenderer

    class Stub:                                         # row_hit_counts of two very different views
        def row_hit_counts(self, cam, W, H):
            top = np.zeros(H); top[: H // 4] = W
            bottom = np.zeros(H); bottom[H // 2:] = W
            return top if cam == "up" else bottom

    W, H = 64, 240
    br = BandedRenderer.__new__(BandedRenderer)
    br.r = Stub(); br.rank = 0; br.world_size = 3; br.edges = None; br._height = None
    br.balance("up", W, H, 10)
    only_up = list(br.edges)
    br.balance(["up", "down"], W, H, 10)
    mean = 0.5 * (Stub().row_hit_counts("up", W, H) + Stub().row_hit_counts("down", W, H))
    assert br.edges == balanced_band_edges(mean + 0.04 * (W - mean), 3, 10)
    assert br.edges != only_up and br.edges[0] == 0 and br.edges[-1] == H


def test_search_band_edges_minimises_the_frame_time_not_the_stage_balance():
    """The calibration's objective is the measured frame (slowest rank, passes overlapping), the per-band compute times only
    propose cuts: on a model where equal compute times are NOT the fastest frame (one band hides part of its work behind the
    overlap of its passes) and every measurement carries noise, the search must end near the true optimum, never stop at the
    first cut because one noisy measurement looked balanced, and never return a cut slower than the one it started from."""
    from romis_b200.bands import search_band_edges
    H, world, a = 1080, 4, 0.12
    y = np.arange(H)
    cost = 0.3 + np.exp(-((y - 650) / 220.0) ** 2); cost *= 2.0 / cost.sum()
    profile = np.maximum(cost * (1.0 + 0.25 * np.sin(y / 55.0)), 1e-3)
    overlap = np.array([0.00, 0.06, 0.00, 0.03])           # part of a band's compute that the production frame hides
    rng = np.random.default_rng(11)
    calls = []

    def evaluate(e, with_compute):
        compute = np.array([a + cost[e[g]:e[g + 1]].sum() for g in range(world)])
        frame = float(np.max(compute * (1.0 - overlap)))
        calls.append(list(e))
        return frame * (1.0 + 0.002 * rng.standard_normal()), list(compute * (1.0 + 0.004 * rng.standard_normal(world)))

    true_frame = lambda e: max((a + cost[e[g]:e[g + 1]].sum()) * (1.0 - overlap[g]) for g in range(world))
    start = [g * H // world for g in range(world)] + [H]
    edges, ms, log = search_band_edges(start, evaluate, profile, 10)
    # brute-force optimum of the noiseless model on a coarse grid around the answer is not needed: compare with the two natural cuts
    from romis_b200.bands import refine_band_edges
    balanced = list(start)
    for _ in range(8):
        balanced = refine_band_edges(balanced, [a + cost[balanced[g]:balanced[g + 1]].sum() for g in range(world)], profile, 10)
    assert true_frame(edges) <= true_frame(start) and true_frame(edges) <= true_frame(balanced) * 1.002
    assert true_frame(edges) < true_frame(balanced) * 0.995, "the polish must find what the stage balance cannot see"
    assert len(log) == len(set(tuple(e) for e, _ in log)) <= 96 and len(calls) <= 2 * len(log)
    assert min(np.diff(edges)) >= 10 and edges[0] == 0 and edges[-1] == H


def test_search_band_edges_is_not_ended_by_a_lucky_first_measurement():
    from romis_b200.bands import search_band_edges
    H, world = 1080, 2
    cost = np.where(np.arange(H) < 400, 2.0, 1.0); cost = cost / cost.sum()
    first = [True]

    def evaluate(e, with_compute):
        t = [0.1 + cost[e[g]:e[g + 1]].sum() for g in range(world)]
        if first[0]:                                    # the first measurement claims perfect balance
            first[0] = False
            return max(t), [np.mean(t)] * world
        return max(t), t

    edges, ms, log = search_band_edges([0, 540, H], evaluate, cost, 10)
    best = min(range(10, H - 10), key=lambda c: max(cost[:c].sum(), cost[c:].sum()))
    assert abs(edges[1] - best) <= 4 and len(log) > 1
