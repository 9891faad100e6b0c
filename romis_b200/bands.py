"""Row-band sharding of one frame over several GPUs (one process per GPU) and the reservoir halo exchange.

The reference renders the whole frame in one address space; only `spatialReuse` reads other pixels, inside a
+-radius window of the previous iteration's grid (reference src/rendering/render_utils.cpp:91,109-111).  Rank g
owns a contiguous band of rows; before every spatial pass it sends its `radius` boundary rows of the reservoir
buffer to the neighbouring bands and receives theirs.  Transport is torch.distributed point to
point (NCCL over NVLink on GPUs, gloo in the CPU tests); the buffers are the renderer's own device memory
aliased as tensors, so nothing is staged.
"""
from __future__ import annotations

from typing import List, Optional, Tuple

import torch
import torch.distributed as dist

from . import abi


def band_rows(height: int, world_size: int, rank: int) -> Tuple[int, int]:
    """Contiguous, balanced row bands: rank g owns [g*H/G, (g+1)*H/G)."""
    return (rank * height) // world_size, ((rank + 1) * height) // world_size


def balanced_band_edges(row_cost, world_size: int, min_rows: int):
    """Cuts rows into `world_size` contiguous bands of (nearly) equal total cost, each at least `min_rows` tall.
    Deterministic, so every rank derives the same edges from the same cost profile.  Returns world_size + 1 edges."""
    import numpy as np
    cost = np.asarray(row_cost, np.float64)
    H = len(cost)
    min_rows = max(1, int(min_rows))
    if world_size * min_rows > H:
        raise ValueError(f"{world_size} bands of at least {min_rows} rows do not fit {H} rows")
    prefix = np.concatenate([[0.0], np.cumsum(cost)])
    edges = [0]
    for g in range(1, world_size):
        target = prefix[-1] * g / world_size
        e = int(np.searchsorted(prefix, target))
        e = max(e, edges[-1] + min_rows)                        # this band tall enough
        e = min(e, H - (world_size - g) * min_rows)             # room for the remaining bands
        edges.append(e)
    edges.append(H)
    return edges


def refine_band_edges(edges, times, row_profile, min_rows: int, fixed_frac: float = 0.4):
    """One refinement step of the band edges from measured per-band times (BandedRenderer.calibrate).

    A band's time is not proportional to its rows: every pass kernel has a latency floor (launch, one wave of blocks whose
    threads run ~1e4 dependent instructions), so t = a + sum of the rows' costs with `a` a sizeable part of a thin band's
    time.  Taking t / rows as the cost density makes narrow bands look expensive and the iteration over-correct; the part
    `fixed_frac * mean(t)` is therefore taken off first.  Inside a band the remaining cost is spread like `row_profile`
    (hit pixels per row: the only shape information there is), not uniformly.  Returns the new edges."""
    import numpy as np
    world = len(edges) - 1
    t = np.asarray(times, np.float64)
    prof = np.maximum(np.asarray(row_profile, np.float64), 1e-9)
    a = fixed_frac * float(t.mean())
    cost = np.zeros(len(prof))
    for g in range(world):
        lo, hi = edges[g], edges[g + 1]
        var = max(t[g] - a, 0.25 * t[g])
        cost[lo:hi] = var * prof[lo:hi] / prof[lo:hi].sum()
    return balanced_band_edges(cost, world, min_rows)


def search_band_edges(edges, evaluate, row_profile, min_rows: int, rounds: int = 8, polish_steps=(16, 8, 4), gain: float = 0.003,
                      max_evaluations: int = 96):
    """The band edges with the smallest measured FRAME time (BandedRenderer.calibrate; pure, so that it is testable on a cost model).

    `evaluate(edges, with_compute)` -> (frame_ms, per-band compute times or None): the objective is the frame as the caller sees
    it -- slowest rank, passes overlapping as they do in production -- and the per-band compute times only steer the proposals
    (`refine_band_edges`).  Measured facts that shape this (B200, C2 on 2 GPUs): bands whose per-stage compute times agree to
    1 % can still be 4 % off the best frame time (the per-stage events serialise what the frame overlaps); one noisy measurement
    of the first cut used to end the calibration there; and the frame time is not smooth in the edge position (a band's passes
    are a whole number of block rows and a fractional number of waves), so a +-1-step descent stalls 16 rows short of a cut that
    is 3.5 % faster.  So: (1) up to `rounds` refinement proposals, each one evaluated, none ending the search early; (2) from the
    best cut seen, a pattern search over the interior edges -- per edge the offsets -2s, -s, +s, +2s for s in `polish_steps`
    (rows), the best of them taken if the frame gets faster by more than `gain` (relative; measurement noise), swept until
    nothing moves.  Deterministic given `evaluate`, so ranks that see the same measurements walk the same path.
    Returns (edges, frame_ms, log)."""
    world = len(edges) - 1
    min_rows = max(1, int(min_rows))
    seen = {}
    log = []

    def ev(e, with_compute=False):
        key = tuple(int(v) for v in e)
        if key not in seen or (with_compute and seen[key][1] is None):
            fresh = key not in seen
            seen[key] = evaluate(list(key), with_compute)
            if fresh:
                log.append((list(key), float(seen[key][0])))
        return seen[key]

    def valid(e):
        return all(e[g + 1] - e[g] >= min_rows for g in range(world))

    cur = [int(v) for v in edges]
    best = (ev(cur, True)[0], list(cur))
    for _ in range(rounds):
        if len(seen) >= max_evaluations:
            break
        new = refine_band_edges(cur, ev(cur, True)[1], row_profile, min_rows)
        if tuple(new) in seen:
            break
        cur = new
        f = ev(cur, True)[0]
        if f < best[0]:
            best = (f, list(cur))
    for step in polish_steps:
        for _sweep in range(3):
            moved = False
            for i in range(1, world):
                trial = None
                for d in (-2 * step, -step, step, 2 * step):
                    cand = list(best[1]); cand[i] += d
                    if not valid(cand) or (tuple(cand) not in seen and len(seen) >= max_evaluations):
                        continue
                    f = ev(cand)[0]
                    if trial is None or f < trial[0]:
                        trial = (f, cand)
                if trial is not None and trial[0] < best[0] * (1.0 - gain):
                    best = trial; moved = True
            if not moved:
                break
    return best[1], best[0], log


def check_bands(height: int, world_size: int, radius: int) -> None:
    smallest = min(band_rows(height, world_size, r)[1] - band_rows(height, world_size, r)[0] for r in range(world_size))
    if world_size > 1 and smallest < radius:
        raise ValueError(f"bands of {smallest} rows are thinner than the spatial radius {radius}: halos would span two bands")


class _CudaAlias:
    """Exposes a raw device range through __cuda_array_interface__ so torch can alias it without a copy."""

    def __init__(self, ptr: int, nbytes: int):
        self.__cuda_array_interface__ = {"shape": (nbytes,), "typestr": "|u1", "data": (ptr, False), "version": 2}


def alias_device_bytes(ptr: int, nbytes: int, device: torch.device) -> torch.Tensor:
    if nbytes == 0:
        return torch.empty(0, dtype=torch.uint8, device=device)
    return torch.as_tensor(_CudaAlias(ptr, nbytes), device=device)


def exchange_halos(send_low: Optional[torch.Tensor], send_high: Optional[torch.Tensor],
                   recv_low: Optional[torch.Tensor], recv_high: Optional[torch.Tensor],
                   rank: int, world_size: int, group=None) -> None:
    """One halo exchange: rows go to rank-1 ("low", smaller y) and rank+1 ("high"); all four transfers are
    posted together (batch_isend_irecv) and waited for.  Tensors may be empty/None at the image edges."""
    ops: List[dist.P2POp] = []
    if rank > 0:
        if recv_low is not None and recv_low.numel():
            ops.append(dist.P2POp(dist.irecv, recv_low, rank - 1, group))
        if send_low is not None and send_low.numel():
            ops.append(dist.P2POp(dist.isend, send_low, rank - 1, group))
    if rank < world_size - 1:
        if recv_high is not None and recv_high.numel():
            ops.append(dist.P2POp(dist.irecv, recv_high, rank + 1, group))
        if send_high is not None and send_high.numel():
            ops.append(dist.P2POp(dist.isend, send_high, rank + 1, group))
    if not ops:
        return
    for w in dist.batch_isend_irecv(ops):
        w.wait()


class BandedRenderer:
    """Drives one `RestirRenderer` as rank `rank` of `world_size` row bands.

    transport "peer" (default): halo rows are pushed into the neighbours' buffers through CUDA-IPC mapped memory and
    ordered by device-side flags inside romis_frame_spatial_pass (include/romis_gpu.h romis_peer_*): torch.distributed is
    used once, to swap the 512-byte IPC blobs.  transport "nccl": batch_isend_irecv per pass (portable baseline)."""

    def __init__(self, renderer, rank: int, world_size: int, device: torch.device, transport: str = "peer"):
        self.r = renderer
        self.rank, self.world_size, self.device = rank, world_size, device
        self.stream = torch.cuda.ExternalStream(renderer.stream(), device=device)
        self._height = None
        self.transport = transport if world_size > 1 else "none"
        self._attached = None
        self.fallback_reason = None
        self.edges = None           # explicit band edges (balanced_band_edges); None = equal row counts
        self._archive_manual = False

    def _attach_peers(self, features, W, H):
        """prepare -> export -> swap blobs -> attach; redone when the frame geometry changes."""
        key = (W, H, features.numSamplesInReservoir, features.spatialResampleRadius)
        if self._attached == key:
            return
        if self._attached is not None:
            self._detach_all()
        ok, err = 1, ""
        try:
            self.r.band_prepare(features, W, H)
            blob = self.r.peer_export()
        except Exception as e:                          # noqa: BLE001 -- every rank must still reach the collectives below
            ok, err, blob = 0, str(e), b""
        blobs = [None] * self.world_size
        dist.all_gather_object(blobs, blob)
        if ok and all(blobs):
            try:
                self.r.peer_attach(blobs[self.rank - 1] if self.rank > 0 else None,
                                   blobs[self.rank + 1] if self.rank < self.world_size - 1 else None)
            except Exception as e:                      # noqa: BLE001
                ok, err = 0, str(e)
        else:
            ok = 0
        flag = torch.tensor([ok], dtype=torch.int32, device=self.device)
        dist.all_reduce(flag, op=dist.ReduceOp.MIN)
        if int(flag.item()) == 0:
            # some rank cannot map its neighbour (no peer access, IPC disabled, ...): everybody falls back to NCCL p2p
            try:
                self.r.peer_detach()
            except Exception:                           # noqa: BLE001
                pass
            self.transport = "nccl"
            self.fallback_reason = err or "a neighbouring rank could not attach"
            self._attached = None
            return
        self._attached = key

    def _detach_all(self):
        """Every rank drains its stream and closes its mappings of the neighbours' buffers; only after ALL ranks have done
        so may any of them re-allocate or free the exported buffers (CUDA: freeing exported memory an importer still has
        open is undefined) -- hence the barrier between detach and whatever follows (band_prepare, destroy)."""
        self.r.peer_detach()
        self._attached = None
        if self.world_size > 1:
            dist.barrier()

    def close(self):
        """Collective: detach, barrier, then destroy the context."""
        if self._attached is not None:
            self._detach_all()
        self.r.close()

    def upload_lights(self, lights, dirty=None):
        """scene.lights for the next frame, on every rank.  Halo rows carry light-archive slots from band to band, so all
        ranks recycle the same slots: the marks of every band's history are OR-ed before the release (romis_gpu.h)."""
        import numpy as np
        if self.world_size > 1:
            if not self._archive_manual:
                self.r.set_light_archive_auto(False); self._archive_manual = True
            marks = self.r.light_archive_marks()
            n = torch.tensor([len(marks)], dtype=torch.int64, device=self.device)
            dist.all_reduce(n, op=dist.ReduceOp.MAX)
            n = int(n.item())
            if n:
                keep = torch.zeros(n, dtype=torch.uint8, device=self.device)
                keep[:len(marks)] = torch.from_numpy(np.ascontiguousarray(marks)).to(self.device)
                dist.all_reduce(keep, op=dist.ReduceOp.MAX)
                self.r.light_archive_release(keep.cpu().numpy())
        self.r.upload_lights(lights, dirty)

    def balance(self, camera, W: int, H: int, radius: int, miss_cost: float = 0.04):
        """Equal-COST bands for this camera: hit pixels carry the work of every pass, miss pixels short-circuit
        (cost ratio measured on B200).  Every rank derives the same edges from romis_row_hit_counts; call it before the
        first frame (moving the edges later drops the temporal history of the rows that change owner).
        A moving camera whose path is known (BASELINE C4: the 64-frame orbit) passes a list of cameras sampled along the path:
        the bands are cut for the MEAN profile and stay put for the whole sequence, so no history changes owner."""
        hits = self._hit_profile(camera, W, H)
        self.edges = balanced_band_edges(hits + miss_cost * (W - hits), self.world_size, max(radius, 1))
        self._height = None

    def _hit_profile(self, camera, W: int, H: int):
        """Per-row primary-ray hit counts; for a camera PATH (list of cameras) their mean over the path."""
        cams = list(camera) if isinstance(camera, (list, tuple)) else [camera]
        return sum(self.r.row_hit_counts(c, W, H).astype("float64") for c in cams) / len(cams)

    def calibrate(self, features, camera, W: int, H: int, seed: int = 1, rounds: int = 8, frames: int = 6, before_frame=None):
        """Measured refinement of the band edges (static camera, or a list of cameras along a known path): `search_band_edges`
        with this renderer as the measuring device -- per cut a few frames with per-stage events (every rank's own compute time,
        which steers the proposals) and a few frames as they run in production (the objective: the slowest rank's frame time,
        median over the frames).  Cost per hit pixel varies across the image (e.g. surfaces facing away from most lights take the
        `NL < 0` early exit), which the hit-count profile cannot see.  Moving an edge re-allocates the band, so history is
        dropped and peers are re-attached: call this before the frames that matter.  Collective: every rank calls it and, seeing
        the same gathered measurements, takes the same decisions.  `before_frame()` runs ahead of every measured frame: a caller
        that times its frames under special conditions (bench.py: L2 flushed before every step) calibrates under the same ones --
        the best cut differs by ~15 rows of 1080 between warm and cold caches."""
        import numpy as np
        if self.world_size == 1:
            return
        radius = features.spatialResampleRadius if features.spatialReuse else 0
        if self.edges is None:
            self.edges = [band_rows(H, self.world_size, g)[0] for g in range(self.world_size)] + [H]
        cams = list(camera) if isinstance(camera, (list, tuple)) else [camera]     # a camera path: timed over one pass along it
        hits = self._hit_profile(cams, W, H)
        profile = hits + 0.04 * (W - hits)
        frames = max(frames, len(cams)) if len(cams) > 1 else frames

        def run(stage_timing):
            """frames + 1 frames of the current cut (the first one builds the history); every rank's per-frame values"""
            self.r.set_stage_timing(stage_timing)
            vals = []
            for fr in range(frames + 1):
                if before_frame is not None:
                    before_frame()
                self.render_frame(features, cams[fr % len(cams)], W, H, fr > 0, seed, fr, out=None)
                self.r.synchronize()
                t = self.r.timings()
                if fr > 0:       # own compute only: the wait for the neighbours is reported separately (exchange_ms)
                    vals.append(t.primary_ms + t.initial_ms + t.temporal_ms + t.shade_ms + sum(t.spatial_ms[:t.n_spatial])
                                if stage_timing else t.total_ms)
            self.r.set_stage_timing(False)
            every = [None] * self.world_size
            dist.all_gather_object(every, vals)
            return every

        def evaluate(edges, with_compute):
            if list(edges) != list(self.edges):
                if self._attached is not None:
                    self._detach_all()
                self.edges = list(edges)
                self._height = None
            compute = [float(np.mean(v)) for v in run(True)] if with_compute else None
            whole = run(False)                                  # the frame as it runs in production: no events between the passes
            frame_ms = float(np.median([max(v[i] for v in whole) for i in range(len(whole[0]))]))
            return frame_ms, compute

        best_edges, best_ms, log = search_band_edges(self.edges, evaluate, profile, max(radius, 1), rounds=rounds)
        if list(best_edges) != list(self.edges):
            if self._attached is not None:
                self._detach_all()
            self.edges = list(best_edges)
            self._height = None
        self.calibration = {"frame_ms": best_ms, "edges": list(best_edges), "evaluations": len(log)}
        self.r.reset_history()

    def band(self, height: int):
        if self.edges is not None and self.edges[-1] == height:
            return self.edges[self.rank], self.edges[self.rank + 1]
        return band_rows(height, self.world_size, self.rank)

    def set_height(self, height: int, radius: int):
        if self._height != height:
            if self.edges is None or self.edges[-1] != height:
                check_bands(height, self.world_size, radius)
            y0, y1 = self.band(height)
            self.r.set_band(y0, y1)
            self._height = height

    def render_frame(self, features, camera, W, H, history_valid, seed, frame, out=None):
        """The banded frame.  All work (kernels and transfers) is ordered on the renderer's own stream."""
        self.set_height(H, features.spatialResampleRadius if features.spatialReuse else 0)
        r = self.r
        if self.transport == "peer" and features.spatialReuse:
            self._attach_peers(features, W, H)          # may switch self.transport to "nccl" on every rank
        with torch.cuda.stream(self.stream):
            r.frame_begin(features, camera, W, H, history_valid, seed, frame)
            if features.spatialReuse:
                for p in range(features.spatialResamplingPasses):
                    if self.transport == "nccl":
                        t = [alias_device_bytes(*r.halo_region(w), self.device) for w in
                             (abi.ROMIS_HALO_SEND_LOW, abi.ROMIS_HALO_SEND_HIGH, abi.ROMIS_HALO_RECV_LOW, abi.ROMIS_HALO_RECV_HIGH)]
                        exchange_halos(t[0], t[1], t[2], t[3], self.rank, self.world_size)
                    r.frame_spatial_pass(p)
            r.frame_end(out)
