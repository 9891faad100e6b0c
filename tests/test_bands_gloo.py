"""The N > 1 path on CPU: world_size-2 and -3 process groups over gloo exercise the halo exchange plumbing of
romis_b200/bands.py (partitioning, neighbour ranks, the four posted transfers, edge ranks) on stand-in band buffers."""
import os
import socket

import numpy as np
import pytest
import torch
import torch.distributed as dist
import torch.multiprocessing as mp

from romis_b200.bands import band_rows, exchange_halos


def _free_port():
    s = socket.socket(); s.bind(("127.0.0.1", 0)); p = s.getsockname()[1]; s.close(); return p


def _worker(rank, world, port, H, W, radius, passes, out_dir):
    os.environ["MASTER_ADDR"] = "127.0.0.1"; os.environ["MASTER_PORT"] = str(port)
    dist.init_process_group("gloo", rank=rank, world_size=world)
    try:
        y0, y1 = band_rows(H, world, rank)
        ey0, ey1 = max(0, y0 - radius), min(H, y1 + radius)
        stride = W * 5                                  # stand-in for row_stride bytes
        # the full "previous iteration" buffer every rank would agree on, and this rank's banded copy of it
        full = (torch.arange(H * stride, dtype=torch.int64) * 2654435761 % 251).to(torch.uint8).reshape(H, stride)
        for p in range(passes):
            full = (full.to(torch.int64) * 7 + p + 1).remainder(251).to(torch.uint8)
            buf = torch.zeros(ey1 - ey0, stride, dtype=torch.uint8)
            buf[y0 - ey0:y1 - ey0] = full[y0:y1]        # only own rows are valid before the exchange
            lo = lambda a, b: buf[a - ey0:b - ey0].reshape(-1)
            send_low = lo(y0, y0 + radius) if rank > 0 else None
            send_high = lo(y1 - radius, y1) if rank < world - 1 else None
            recv_low = lo(ey0, y0); recv_high = lo(y1, ey1)
            exchange_halos(send_low, send_high, recv_low, recv_high, rank, world)
            assert torch.equal(buf, full[ey0:ey1]), f"rank {rank} pass {p}: halo rows differ from the neighbours' rows"
        np.save(os.path.join(out_dir, f"ok{rank}.npy"), np.array([1]))
    finally:
        dist.destroy_process_group()


@pytest.mark.parametrize("world,H,radius", [(2, 40, 6), (3, 50, 10), (2, 21, 10)])
def test_halo_exchange_over_gloo(world, H, radius, tmp_path):
    mp.spawn(_worker, args=(world, _free_port(), H, 17, radius, 3, str(tmp_path)), nprocs=world, join=True)
    assert all((tmp_path / f"ok{r}.npy").exists() for r in range(world))


# ---- the band calibration as a collective: every rank measures its own band, all gather, all take the same decisions ----
class _Timings:
    def __init__(self, compute, total):
        self.primary_ms = 0.0; self.initial_ms = compute; self.temporal_ms = 0.0; self.shade_ms = 0.0
        self.spatial_ms = [0.0]; self.n_spatial = 0; self.total_ms = total


class _StubRenderer:
    """Stands in for RestirRenderer: a band's time is a latency floor plus the cost of its rows, with rank-dependent noise."""

    def __init__(self, rank, cost):
        self.rank, self.cost, self.band, self.stage = rank, cost, (0, len(cost)), False
        self.rng = np.random.default_rng(100 + rank); self.frames = 0

    def row_hit_counts(self, cam, W, H):
        return np.asarray(self.cost) * W

    def set_stage_timing(self, on): self.stage = on
    def synchronize(self): pass
    def reset_history(self): pass

    def timings(self):
        t = 0.15 + float(np.sum(self.cost[self.band[0]:self.band[1]])) * (1.0 + 0.003 * self.rng.standard_normal())
        return _Timings(t, t * (0.97 if self.rank == 0 else 1.0))     # rank 0 hides part of its work in the production frame


def _calibrate_worker(rank, world, port, out_dir):
    os.environ["MASTER_ADDR"] = "127.0.0.1"; os.environ["MASTER_PORT"] = str(port)
    dist.init_process_group("gloo", rank=rank, world_size=world)
    try:
        from romis_b200.bands import BandedRenderer
        from romis_b200.scene import Features
        H, W = 360, 64
        y = np.arange(H)
        cost = (0.3 + np.exp(-((y - 230) / 70.0) ** 2)); cost = cost / cost.sum()

        class Stub(BandedRenderer):
            def render_frame(self, features, camera, W, H, history_valid, seed, frame, out=None):
                self.r.band = self.band(H); self.r.frames += 1

        br = Stub.__new__(Stub)
        br.r = _StubRenderer(rank, cost); br.rank, br.world_size = rank, world
        br._attached = None; br._height = None; br.edges = None; br.transport = "none"
        feat = Features(spatialResamplingPasses=1)
        br.balance("cam", W, H, feat.spatialResampleRadius)
        first = list(br.edges)
        calls = []
        br.calibrate(feat, "cam", W, H, before_frame=lambda: calls.append(1))
        every = [None] * world
        dist.all_gather_object(every, (br.edges, br.calibration["frame_ms"], br.r.frames, len(calls)))
        assert all(e[0] == every[0][0] for e in every), f"ranks ended on different cuts: {every}"
        assert all(e[2] == every[0][2] for e in every), "ranks rendered different numbers of frames (a collective would hang)"
        assert every[0][3] == br.r.frames, "before_frame runs ahead of every measured frame"
        e = br.edges
        assert e[0] == 0 and e[-1] == H and min(np.diff(e)) >= feat.spatialResampleRadius
        t = [0.15 + cost[e[g]:e[g + 1]].sum() for g in range(world)]
        t0 = [0.15 + cost[first[g]:first[g + 1]].sum() for g in range(world)]
        assert max(t) <= max(t0) * 1.002 and max(t) / np.mean(t) < 1.06, (e, t)
        np.save(os.path.join(out_dir, f"cal{rank}.npy"), np.array(e))
    finally:
        dist.destroy_process_group()


@pytest.mark.parametrize("world", [2, 3])
def test_band_calibration_is_a_consistent_collective(world, tmp_path):
    mp.spawn(_calibrate_worker, args=(world, _free_port(), str(tmp_path)), nprocs=world, join=True)
    cuts = [np.load(tmp_path / f"cal{r}.npy") for r in range(world)]
    assert all(np.array_equal(c, cuts[0]) for c in cuts)
