#!/usr/bin/env python3
"""bench.py -- the ReSTIR frame benchmark (BASELINE.json metric).

    python bench.py [--gpus N] [--steps K] [--warmup W] [--impl ours|reference] [--config c1|c2|c2u|c3|c4|c4k|rmis|romis]

A "step" is one ReSTIR frame (renderReSTIR, reference src/rendering/render.cpp:28-62) of the workload:
  c2 (default)  cornell-nightclub, 1920x1080, M=32, N=2, temporal + 3 spatial passes k=5 r=10, visibility reuse
                (BASELINE.json configs[1], the configuration the metric is quoted on)
Frame 0 has no temporal history, so warm-up frames establish it; every timed frame runs all seven passes.
  rmis / romis  the same scene and resolution through the reference's other two estimators:
                one step = one renderRMIS / renderROMIS frame with the reference's defaults (5 iterations, k=5, r=10,
                similarity-based neighbours; R-MIS with equal weights, R-OMIS direct estimator).  One GPU.

  value  frames/s with everything resident in HBM, device time (CUDA events on the launching stream), image left
         on the device.  For N > 1 the frame is split into N row bands, one process per GPU; before each spatial pass the
         boundary reservoir rows are pushed into the neighbouring bands' halo rows over NVLink (peer-mapped CUDA-IPC
         memory, flag-ordered; NCCL only bootstraps -- `--halo nccl` selects NCCL send/recv instead); time = max over
         ranks; "scaling": "strong" (one frame).
  e2e    the same metric through the reference-facing call romis_render_frame with HOST buffers: per step the
         scene's lights are handed over again (the reference reads scene.lights fresh every frame, light.cpp:46-66;
         the whole table is compared with what the device holds) and the float RGB image is read back into page-locked
         host memory, all inside the timed region.  `e2e.light_edit` is the same with one light edited per frame (its
         old record is archived for the history samples drawn from it, tests/test_gpu_light_edits.py).
  roofline / cpu_baseline: see DESIGN.md "measurement".

--impl reference times the reference's own CPU implementation (oracle/_ref/libromis_ref.so: the reference's
translation units compiled here, OpenMP on ALL host cores set explicitly -- torchrun's OMP_NUM_THREADS=1 is ignored --,
thread-safe RNG shim) on full-size frames of the same workload; rank 0 only.
"""
from __future__ import annotations

import argparse
import json
import os
import subprocess
import sys
import threading
import time

import numpy as np

ROOT = os.path.dirname(os.path.abspath(__file__))
sys.path.insert(0, ROOT)

from romis_b200.scene import Camera, Features, Scene, synthetic_lights  # noqa: E402

SCENES = os.path.join(ROOT, "tests", "golden", "scenes")
SEED = 20240229


def workload(name: str):
    """-> (label, scene, W, H, Features, Camera)."""
    if name == "c1":
        s = Scene.load(os.path.join(SCENES, "CornellBoxParallelogramLight.npz"))
        return ("CornellBox-Mirror-Rotated 512x512 M=32 N=2 temporal + 1 spatial k=5 r=10", s, 512, 512,
                Features(spatialResamplingPasses=1), Camera(50.0, 3.0, (0.0, 0.0, 0.0), (20.0, 20.0, 0.0)))
    if name == "c2":
        s = Scene.load(os.path.join(SCENES, "CornellNightClub.npz"))
        return ("cornell-nightclub 1920x1080 M=32 N=2 temporal + 3 spatial k=5 r=10, visibility reuse", s, 1920, 1080,
                Features(spatialResamplingPasses=3, initialSamplesVisibilityCheck=True), Camera())
    if name == "c3":
        s = Scene.load(os.path.join(SCENES, "Monkey.npz"))
        s.lights = synthetic_lights(65536, seed=1)
        return ("monkey + 65536 synthetic point/parallelogram lights 3840x2160 M=32 N=2 temporal + 3 spatial k=5 r=10", s, 3840, 2160,
                Features(spatialResamplingPasses=3, initialSamplesVisibilityCheck=True), Camera(50.0, 3.0, (0.0, 0.0, 0.0), (20.0, 20.0, 0.0)))
    if name == "c4k":
        s = Scene.load(os.path.join(SCENES, "CornellNightClub.npz"))
        return ("cornell-nightclub 3840x2160 M=32 N=2 temporal + 3 spatial k=5 r=10, visibility reuse", s, 3840, 2160,
                Features(spatialResamplingPasses=3, initialSamplesVisibilityCheck=True), Camera())
    if name == "c2u":    # combineUnbiased + spatialReuseVisibilityCheck, (k+1)*N more shadow rays per pixel and pass
        label, s, W, H, f, cam = workload("c2")
        f.unbiasedCombination = True; f.spatialReuseVisibilityCheck = True
        return (label.replace("visibility reuse", "visibility reuse, unbiased combination with spatial visibility"), s, W, H, f, cam)
    if name == "c4":     # the 64-frame orbit: same as c2 with the camera moving every frame (see camera_for_frame)
        label, s, W, H, f, cam = workload("c2")
        return (label + ", camera orbiting 360 deg / 64 frames", s, W, H, f, cam)
    raise SystemExit(f"unknown config {name}")


def camera_for_frame(config: str, cam: Camera, frame: int) -> Camera:
    """BASELINE C4: rotation.y = 30 deg + 360 deg * f / 64 via Trackball::setCamera; other configs: static."""
    if config != "c4":
        return cam
    return Camera(cam.fov_deg, cam.distance, cam.look_at, (cam.rotation_deg[0], 30.0 + 360.0 * (frame % 64) / 64.0, cam.rotation_deg[2]))


# algorithmic bytes per pixel per pass: S = 20 B per sub-reservoir, G = 20 B per pixel
def pass_bytes(N: int):
    return {"primary": 20, "initial": 20 + 20 * N, "temporal": 20 + 60 * N, "spatial": 20 + 40 * N, "shade": 32 + 20 * N}


CLOCK_POLL_S = float(os.environ.get("ROMIS_CLOCK_POLL_MS", "2")) * 1e-3


class ClockSampler:
    """SM clock and throttle reasons DURING the timed legs, polled through NVML every ~2 ms from a thread (nvidia-smi -lms
    needs ~1 s to start and then misses a 12 ms region).  Only samples taken between begin() and end() count."""

    def __init__(self, gpu_index: int):
        self.gpu = gpu_index; self.rows = []; self.on = False; self.stop_flag = False; self.th = None; self.h = None
        try:
            import pynvml
            pynvml.nvmlInit()
            self.nv = pynvml
            self.h = pynvml.nvmlDeviceGetHandleByIndex(gpu_index)
            self.sm_max = float(pynvml.nvmlDeviceGetMaxClockInfo(self.h, pynvml.NVML_CLOCK_SM))
        except Exception:       # noqa: BLE001 -- no NVML: the line says so (samples: 0)
            self.h = None; self.sm_max = None

    def start(self):
        if self.h is None:
            return
        self.th = threading.Thread(target=self._poll, daemon=True); self.th.start()

    def begin(self):
        self.on = True

    def end(self):
        self.on = False

    def _poll(self):
        nv = self.nv
        while not self.stop_flag:
            if self.on:
                try:
                    reasons = getattr(nv, "nvmlDeviceGetCurrentClocksEventReasons", None) or nv.nvmlDeviceGetCurrentClocksThrottleReasons
                    self.rows.append((float(nv.nvmlDeviceGetClockInfo(self.h, nv.NVML_CLOCK_SM)), int(reasons(self.h))))
                except Exception:   # noqa: BLE001
                    pass
            time.sleep(CLOCK_POLL_S)

    def stop(self):
        self.stop_flag = True
        if self.th:
            self.th.join(timeout=1)
        if self.h is None:
            return {"sm_mhz": None, "sm_max_mhz": None, "reasons": [], "samples": 0, "source": "NVML unavailable"}
        nv = self.nv
        names = (("hw_slowdown", "nvmlClocksEventReasonHwSlowdown", 0x8), ("hw_thermal_slowdown", "nvmlClocksEventReasonHwThermalSlowdown", 0x40),
                 ("sw_thermal_slowdown", "nvmlClocksEventReasonSwThermalSlowdown", 0x20), ("sw_power_cap", "nvmlClocksEventReasonSwPowerCap", 0x4))
        reasons = set()
        for _, bits in self.rows:
            for name, attr, dflt in names:
                if bits & int(getattr(nv, attr, dflt)):
                    reasons.add(name)
        sm = [r[0] for r in self.rows]
        return {"sm_mhz": float(np.median(sm)) if sm else None, "sm_max_mhz": self.sm_max, "reasons": sorted(reasons), "samples": len(sm),
                "source": "NVML polled every ~2 ms during the timed legs (value, per-stage repeat, e2e)"}


def measured_peak_gbs():
    p = os.path.join(ROOT, "MEASURED_PEAKS.json")
    if os.path.exists(p):
        try:
            return float(json.load(open(p))["hbm_gbs"]), "measured (MEASURED_PEAKS.json hbm_gbs)"
        except Exception:
            pass
    return 6650.0, "fallback (B200_PROFILING.md)"


def host_cores() -> int:
    try:
        return len(os.sched_getaffinity(0))
    except AttributeError:
        return os.cpu_count() or 1


REF_FULL_BUDGET_S = 240.0       # full-size reference frames per run: as many of the K steps as fit this budget


def run_reference(args, label, scene, W, H, feat, cam, rank, world):
    """The reference's own CPU path (oracle/_ref: its translation units compiled here) on ALL host cores -- set explicitly,
    torchrun exports OMP_NUM_THREADS=1 -- at the workload's FULL size: one step = one full frame (about 8.5 s at C2 on 16
    cores), so steps x ms_per_step is the time this run really took.  Warm-up: reduced-size frames, then one full-size
    frame (a resolution change drops the history, so the timed frames need a full-size predecessor).  If K full frames do
    not fit REF_FULL_BUDGET_S the remaining steps are reduced-size samples scaled by pixel count (stated in `sample`)."""
    if rank != 0:
        return
    from oracle.pyoracle import RefLib, REF_FLAG_TIMING_RNG
    ref = RefLib()
    cores = host_cores()
    ref.lib.ref_set_num_threads(cores)
    ref.set_scene(scene)
    div = 4 if W * H >= 1920 * 1080 else 2
    w, h = W // div, H // div
    scale = (w * h) / float(W * H)

    last = [None]

    def frame(fr, ww, hh):
        # history only from a frame of the same size: the reference indexes previousFrameGrid with the new frame's pixels
        # (render_utils.cpp:154), so a predecessor of another resolution is an out-of-bounds read, not a reset
        hist = last[0] == (ww, hh)
        last[0] = (ww, hh)
        t0 = time.perf_counter()
        ref.render_frame(feat, camera_for_frame(args.config, cam, fr), ww, hh, hist, SEED, fr, REF_FLAG_TIMING_RNG, dump=False, want_image=True)
        return time.perf_counter() - t0
    fr = 0
    for _ in range(max(0, args.warmup - 1)):
        frame(fr, w, h); fr += 1
    est = frame(fr, W, H); fr += 1                          # full-size warm-up frame (no history of its size yet: slightly cheaper)
    n_full = max(1, min(args.steps, int(REF_FULL_BUDGET_S / max(est, 1e-3))))
    times = []
    for _ in range(n_full):
        times.append(frame(fr, W, H)); fr += 1
    if n_full < args.steps:
        frame(fr, w, h); fr += 1                            # history at the sample size
        for _ in range(args.steps - n_full):
            times.append(frame(fr, w, h) / scale); fr += 1
    ms = 1e3 * float(np.mean(times))
    fps = 1e3 / ms
    sample = (f"{n_full} full {W}x{H} frames" + (f" + {args.steps - n_full} frames at {w}x{h} scaled by pixel count" if n_full < args.steps else "") +
              f"; OpenMP {cores} threads (set explicitly), thread-safe RNG shim, harness BVH instead of Embree")
    line = {"impl": "reference", "metric": "ReSTIR frames/s", "value": fps, "unit": "frames/s", "n_gpus": args.gpus, "steps": args.steps,
            "warmup": args.warmup, "ms_per_step": ms, "higher_is_better": True, "scaling": "strong", "vs_baseline": None,
            "dtype": "f32", "data": "synthetic", "config": {"workload": label, "width": W, "height": H},
            "gcandidates_per_s": W * H * feat.initialLightSamples * fps / 1e9,
            "cpu_baseline": {"value": fps, "unit": "frames/s", "cores": cores, "kind": "reference", "sample": sample},
            "e2e": {"value": fps, "unit": "frames/s", "h2d_bytes_per_step": 0, "d2h_bytes_per_step": 0}}
    print(json.dumps(line), flush=True)


def cpu_baseline_leg(scene, W, H, feat, cam, config="c2"):
    """Reported baseline next to the GPU number (about 25 s of CPU work): oracle/_ref (the compiled reference) when present,
    else the C port.  One full-size frame is timed, after a full-size predecessor; the reference's own random sources
    ("as-is": glibc rand() behind its lock, std::random_device + std::mt19937 constructed per pixel) are
    timed on a 1/64-pixel sample with all cores and with one."""
    cores = host_cores()
    try:
        from oracle.pyoracle import RefLib, REF_FLAG_TIMING_RNG, REF_FLAG_ASIS_RNG
        ref = RefLib(); ref.lib.ref_set_num_threads(cores); ref.set_scene(scene)
        kind = "reference"

        def frame(fr, ww, hh, flags=REF_FLAG_TIMING_RNG):
            t0 = time.perf_counter()
            ref.render_frame(feat, camera_for_frame(config, cam, fr), ww, hh, fr > 0, SEED, fr, flags, dump=False, want_image=True)
            return time.perf_counter() - t0
    except (FileNotFoundError, OSError):
        from oracle.pyoracle import Oracle
        ref = None
        orc = Oracle(); orc.upload_scene(scene)
        kind = "port"

        def frame(fr, ww, hh, flags=0):
            t0 = time.perf_counter()
            orc.render_frame(feat, camera_for_frame(config, cam, fr).to_abi(ww, hh), ww, hh, fr > 0, SEED, fr)
            return time.perf_counter() - t0
    frame(0, W, H)
    t = frame(1, W, H)
    out = {"value": 1.0 / t, "unit": "frames/s", "cores": cores, "kind": kind,
           "sample": f"1 full {W}x{H} frame (frame 1, after a full-size frame 0), same scene/Features; " +
                     ("OpenMP on all host cores (set explicitly), thread-safe RNG shim, harness BVH instead of Embree" if ref else "C restatement, OpenMP")}
    if ref is not None:
        w, h = max(16, W // 8), max(16, H // 8)
        scale = (w * h) / float(W * H)
        as_is = []
        for n_thr in (cores, 1):
            ref.lib.ref_set_num_threads(n_thr)
            frame(0, w, h, REF_FLAG_ASIS_RNG)
            ts = [frame(1 + i, w, h, REF_FLAG_ASIS_RNG) for i in range(2)]
            as_is.append({"value": scale / float(np.median(ts)), "unit": "frames/s", "cores": n_thr, "kind": "reference",
                          "rng": "as-is: glibc rand(), std::random_device + std::mt19937 per pixel",
                          "sample": f"2 frames at {w}x{h} ({scale:.4f} of the pixels), median, scaled by pixel count"})
        ref.lib.ref_set_num_threads(cores)
        out["as_is"] = as_is
    return out


MIS_DEFAULT_STEPS = 10


def mis_pass_bytes(mode: str, N: int, K1: int):
    """Algorithmic bytes per pixel PER ITERATION of the dominant kernel: compulsory reads/writes, every logical
    array once; S = 20 B per sub-reservoir record, G = 20 B per pixel."""
    if mode == "rmis":      # gather: own G, K1 grid entries, K1*N neighbour records, accumulator read + write
        return 20 + 4 * K1 + 20 * K1 * N + 24
    # accumulate: own G, grid, K1*N records + their wSum/chosen, the K1 distributions' G, the symmetric technique matrix
    # (K1 (K1 + 1) / 2 independent elements) and the 3 contribution vectors read + written
    return 20 + 4 * K1 + (20 + 8) * K1 * N + 20 * K1 + 8 * (K1 * (K1 + 1) // 2 + 3 * K1)


def run_mis(args, mode, rank, world):
    """R-MIS / R-OMIS frames.  N > 1 (torchrun, one process per GPU): row bands of equal cost by the per-row hit profile; a band
    renders its halo rows itself (no collective), every rank's rows land in one shared host image."""
    import torch
    from romis_b200.api import PinnedImage, RestirRenderer
    from romis_b200.scene import RmisParams
    label, scene, W, H, feat, cam = workload("c2")
    label = label.split(" M=")[0] + f" {'R-MIS equal weights' if mode == 'rmis' else 'R-OMIS direct estimator'}, 5 iterations, M=32 N=2 k=5 r=10, similar neighbours, visibility reuse"
    rp = RmisParams()                                   # the reference's defaults (common.h:110-121)
    K1 = feat.numNeighboursToSample + 1; N = feat.numSamplesInReservoir
    if args.impl == "reference":
        if rank != 0:
            return                                      # rank 0 alone runs the CPU reference
        from oracle.pyoracle import RefLib
        ref = RefLib(); ref.lib.ref_set_num_threads(host_cores()); ref.set_scene(scene); ref.set_mis_timing(True)   # torchrun exports OMP_NUM_THREADS=1
        div = 8
        w, h = W // div, H // div
        scale = (w * h) / float(W * H)
        fn = ref.render_frame_rmis if mode == "rmis" else ref.render_frame_romis
        times = []
        for fr in range(max(1, args.warmup) + args.steps):
            t0 = time.perf_counter()
            fn(feat, rp, cam, w, h, SEED, fr, False)
            if fr >= max(1, args.warmup):
                times.append(time.perf_counter() - t0)
        ms = 1e3 * float(np.mean(times)) / scale; fps = 1e3 / ms; cores = host_cores()
        sample = f"{w}x{h} frames of the same scene/Features ({scale:.4f} of the pixels), time scaled by pixel count; OpenMP on all {cores} host cores, thread-safe RNG shim"
        print(json.dumps({"impl": "reference", "metric": f"{mode.upper()} frames/s", "value": fps, "unit": "frames/s", "n_gpus": args.gpus, "steps": args.steps,
                          "warmup": args.warmup, "ms_per_step": ms, "higher_is_better": True, "scaling": "strong", "vs_baseline": None, "dtype": "f32",
                          "data": "synthetic", "config": {"workload": label, "width": W, "height": H},
                          "cpu_baseline": {"value": fps, "unit": "frames/s", "cores": cores, "kind": "reference", "sample": sample},
                          "e2e": {"value": fps, "unit": "frames/s", "h2d_bytes_per_step": 0, "d2h_bytes_per_step": 0}}), flush=True)
        return
    local_rank = int(os.environ.get("LOCAL_RANK", 0))
    torch.cuda.set_device(local_rank)
    dev = torch.device("cuda", local_rank)
    edges = None
    if world > 1:
        import torch.distributed as dist
        from romis_b200.api import SharedImage
        from romis_b200.bands import balanced_band_edges
        os.environ.setdefault("NCCL_DEBUG_FILE", "/dev/stderr")
        dist.init_process_group("nccl", device_id=dev)
    r = RestirRenderer(local_rank); r.upload_scene(scene)
    if world > 1:
        hits = r.row_hit_counts(cam, W, H).astype("float64")
        edges = balanced_band_edges(hits + 0.04 * (W - hits), world, 1)
        r.set_band(edges[rank], edges[rank + 1])
        names = [None]
        if rank == 0:
            host_img = SharedImage(H, W); names[0] = host_img.name
        dist.broadcast_object_list(names, src=0)
        if rank != 0:
            host_img = SharedImage(H, W, name=names[0])
    else:
        host_img = PinnedImage(H, W)
    render = r.render_frame_rmis if mode == "rmis" else r.render_frame_romis
    flush = torch.empty(256 << 20, dtype=torch.uint8, device=dev)
    if world > 1 and not args.equal_rows:
        # measured refinement of the cut (untimed; the same routine as the ReSTIR bands, romis_b200/bands.py): a band's cost is
        # not its hit count alone
        from romis_b200.bands import refine_band_edges
        best = None
        for _ in range(5):
            render(feat, rp, cam, W, H, SEED, 0, want_image=False)
            render(feat, rp, cam, W, H, SEED, 1, want_image=False)
            times = [None] * world
            dist.all_gather_object(times, float(r.timings().total_ms))
            if best is None or max(times) < best[0]:
                best = (max(times), list(edges))
            if max(times) <= 1.01 * sum(times) / world:
                break
            edges = refine_band_edges(edges, times, hits + 0.04 * (W - hits), 1, fixed_frac=0.1)
            r.set_band(edges[rank], edges[rank + 1])
        edges = best[1]
        r.set_band(edges[rank], edges[rank + 1])

    def run_steps(n, first, host=False, stage=False):
        r.set_stage_timing(stage)
        tot = 0.0; acc = {}
        for i in range(n):
            flush.fill_(i & 0xff); torch.cuda.synchronize()
            if world > 1:
                dist.barrier()
            if host:
                scene.lights["c0"][0, 0] = np.float32(0.65 + 1e-4 * ((first + i) % 7)); r.upload_lights(scene.lights)
                t0 = time.perf_counter(); render(feat, rp, cam, W, H, SEED, first + i, out=host_img.array)
                if world > 1:
                    dist.barrier()                  # the frame is there when every band is
                tot += 1e3 * (time.perf_counter() - t0)
            else:
                render(feat, rp, cam, W, H, SEED, first + i, want_image=False)
                t = r.timings(); tot += t.total_ms
                if stage:
                    for k in ("primary_ms", "neighbours_ms", "initial_ms", "gather_ms", "resolve_ms"):
                        acc[k] = acc.get(k, 0.0) + getattr(t, k)
                    acc["launches"] = t.n_launches
        return tot, acc

    run_steps(args.warmup, 0)
    clocks = ClockSampler(0); clocks.start()
    dev_ms, _ = run_steps(args.steps, args.warmup)
    stage_ms, st = run_steps(args.steps, args.warmup + args.steps, stage=True)
    e2e_ms, _ = run_steps(args.steps, args.warmup + 2 * args.steps, host=True)
    clk = clocks.stop()
    one_image = None
    if world > 1:
        if rank == 0:
            host_img.array[...] = np.float32(np.nan)
        dist.barrier()
        run_steps(1, args.warmup + 3 * args.steps, host=True)
        if rank == 0:
            one_image = bool(not np.isnan(host_img.array).any())
        t = torch.tensor([dev_ms, e2e_ms, stage_ms] + [st.get(k, 0.0) for k in ("primary_ms", "neighbours_ms", "initial_ms", "gather_ms", "resolve_ms")],
                         dtype=torch.float64, device=dev)
        dist.all_reduce(t, op=dist.ReduceOp.MAX)
        v = [float(a) for a in t.tolist()]
        dev_ms, e2e_ms, stage_ms = v[:3]
        for k, a in zip(("primary_ms", "neighbours_ms", "initial_ms", "gather_ms", "resolve_ms"), v[3:]):
            st[k] = a
        dist.barrier()
        if rank != 0:
            host_img.free(); r.close(); dist.destroy_process_group()
            return
    ms = dev_ms / args.steps; fps = 1e3 / ms
    iters = rp.maxIterationsMIS
    peak, peak_src = measured_peak_gbs()
    g_ms = st["gather_ms"] / args.steps / iters
    gb = mis_pass_bytes(mode, N, K1)
    ach = gb * W * H / (g_ms * 1e-3) / 1e9
    kern = "rmis_gather_kernel" if mode == "rmis" else "romis_accumulate_kernel"
    line = {"metric": f"{mode.upper()} frames/s", "value": fps, "unit": "frames/s", "n_gpus": world, "steps": args.steps, "warmup": args.warmup,
            "ms_per_step": ms, "higher_is_better": True, "scaling": "strong", "vs_baseline": None, "dtype": "f32", "data": "synthetic", "impl": "ours",
            "config": {"workload": label, "width": W, "height": H, "seed": SEED, "l2": "256 MiB memset before every step (outside the timed region)",
                       "sharding": "single GPU" if world == 1 else f"{world} row bands, halo rows re-rendered per band (no collective)", "band_edges": edges},
            "gcandidates_per_s": W * H * feat.initialLightSamples * iters * fps / 1e9,
            "e2e": {"value": 1e3 / (e2e_ms / args.steps), "unit": "frames/s", "h2d_bytes_per_step": int(6 * 16 * len(scene.lights) + 256), "d2h_bytes_per_step": int(W * H * 12),
                    "note": "host wall clock around romis_render_frame_" + mode + " with a host image; lights re-uploaded (one edited) per frame", "one_host_image": one_image},
            "gpu_launches": int(st["launches"] * args.steps), "clocks": clk,
            "roofline": {"bound": "hbm", "kernel": kern, "achieved": round(ach, 1), "peak": peak, "unit": "GB/s", "frac": round(ach / peak, 4), "traffic": None,
                         "peak_source": peak_src, "algorithmic_bytes_per_px_per_iteration": gb,
                         "note": "issue-bound like the ReSTIR passes: " + ("(k+1) N shading evaluations + shadow rays" if mode == "rmis" else "(k+1)^2 N target-pdf evaluations") + " per pixel per iteration",
                         "stages_ms_per_frame": {k: round(v / args.steps, 4) for k, v in st.items() if k.endswith("_ms")}}}
    if world == 1 and not args.no_cpu_baseline:
        from oracle.pyoracle import RefLib
        ref = RefLib(); ref.set_scene(scene); ref.set_mis_timing(True)
        w, h = W // 8, H // 8; scale = (w * h) / float(W * H)
        fn = ref.render_frame_rmis if mode == "rmis" else ref.render_frame_romis
        fn(feat, rp, cam, w, h, SEED, 0, False)
        ts = []; t_start = time.perf_counter()
        while len(ts) < 3 and time.perf_counter() - t_start < 25.0:
            t0 = time.perf_counter(); fn(feat, rp, cam, w, h, SEED, 1 + len(ts), False); ts.append(time.perf_counter() - t0)
        cms = 1e3 * float(np.median(ts)) / scale
        line["cpu_baseline"] = {"value": 1e3 / cms, "unit": "frames/s", "cores": host_cores(), "kind": "reference",
                                "sample": f"{len(ts)} frames at {w}x{h} ({scale:.4f} of the pixels, same scene/Features), median, scaled by pixel count"}
    print(json.dumps(line), flush=True)
    host_img.free(); r.close()
    if world > 1:
        dist.destroy_process_group()


def main():
    ap = argparse.ArgumentParser()
    ap.add_argument("--gpus", type=int, default=1)
    ap.add_argument("--steps", type=int, default=100)
    ap.add_argument("--warmup", type=int, default=5)
    ap.add_argument("--impl", default="ours", choices=["ours", "reference"])
    ap.add_argument("--config", default="c2")
    ap.add_argument("--no-cpu-baseline", action="store_true")
    ap.add_argument("--equal-rows", action="store_true", help="N > 1: equal row counts instead of equal-cost bands")
    ap.add_argument("--halo", default="peer", choices=["peer", "nccl"], help="halo transport for N > 1")
    args = ap.parse_args()
    args.warmup = max(args.warmup, 3) if args.impl == "ours" else max(args.warmup, 1)

    rank = int(os.environ.get("RANK", "0")); world = int(os.environ.get("WORLD_SIZE", "1"))
    local_rank = int(os.environ.get("LOCAL_RANK", "0"))
    if args.config in ("rmis", "romis"):
        if args.steps == 100:
            args.steps = MIS_DEFAULT_STEPS
        return run_mis(args, args.config, rank, world)
    label, scene, W, H, feat, cam = workload(args.config)

    if args.impl == "reference":
        run_reference(args, label, scene, W, H, feat, cam, rank, world)
        return

    import torch
    import torch.distributed as dist
    from romis_b200.api import PinnedImage, RestirRenderer
    from romis_b200.bands import BandedRenderer, band_rows

    if not torch.cuda.is_available():
        raise SystemExit("bench.py --impl ours needs a CUDA device (there is no CPU fallback)")
    torch.cuda.set_device(local_rank)
    device = torch.device("cuda", local_rank)
    if world > 1:
        os.environ.setdefault("NCCL_DEBUG_FILE", "/dev/stderr")      # NCCL's own banner must not land on stdout next to the JSON line
        dist.init_process_group("nccl", device_id=device)
    r = RestirRenderer(local_rank)
    r.upload_scene(scene)
    br = BandedRenderer(r, rank, world, device, transport=args.halo)
    if world > 1 and not args.equal_rows:
        # equal-COST bands: first cut from the primary-ray hit profile, then refined from measured per-band compute times
        # (untimed, before the warm-up; the camera of this workload is static)
        # (C4, the orbit: cut for the mean over 8 cameras along the path and left alone -- no row changes owner mid-sequence)
        path = [camera_for_frame(args.config, cam, f) for f in range(0, 64, 8)] if args.config == "c4" else cam
        br.balance(path, W, H, feat.spatialResampleRadius if feat.spatialReuse else 0)
    N = feat.numSamplesInReservoir
    flush = torch.empty(256 << 20, dtype=torch.uint8, device=device)     # > 126 MB L2
    if world > 1 and not args.equal_rows:
        def cold_l2():                  # the timed steps start from a flushed L2: so do the calibration's frames
            with torch.cuda.stream(br.stream):
                flush.fill_(1)
        br.calibrate(feat, path, W, H, seed=SEED, before_frame=cold_l2)

    def barrier():
        if world > 1:
            dist.barrier()
        torch.cuda.synchronize(device)

    clocks = ClockSampler(local_rank)
    if rank == 0:
        clocks.start()
    edit_base = float(scene.lights["c0"][0, 0]) if len(scene.lights) else 0.0

    def run_steps(n_steps, first_frame, host_out=None, lights=None, stage_timing=False):
        """Returns (sum of per-step device ms, per-stage sums).  L2 is flushed before every step, outside the events.
        lights: None = not handed over; "same" = the unchanged table handed over every step (compared on the host);
        "edit" = one light edited every step (as from the reference's UI)."""
        r.set_stage_timing(stage_timing)
        total = 0.0; stages = {}
        ev0 = torch.cuda.Event(enable_timing=True); ev1 = torch.cuda.Event(enable_timing=True)
        clocks.begin()
        for i in range(n_steps):
            fr = first_frame + i
            with torch.cuda.stream(br.stream):
                flush.fill_(i & 0xff)
                ev0.record(br.stream)
            if lights is not None:
                if lights == "edit" and len(scene.lights):
                    scene.lights["c0"][0, 0] = np.float32(edit_base * (1.0 + 1e-3 * (1 + fr % 7)))
                br.upload_lights(scene.lights)
            br.render_frame(feat, camera_for_frame(args.config, cam, fr), W, H, fr > 0, SEED, fr, out=host_out)
            with torch.cuda.stream(br.stream):
                ev1.record(br.stream)
            ev1.synchronize()
            total += ev0.elapsed_time(ev1)
            if stage_timing:
                t = r.timings()
                for k in ("primary_ms", "initial_ms", "temporal_ms", "shade_ms"):
                    stages[k] = stages.get(k, 0.0) + getattr(t, k)
                stages["spatial_ms"] = stages.get("spatial_ms", 0.0) + sum(t.spatial_ms[:t.n_spatial])
                stages["exchange_ms"] = stages.get("exchange_ms", 0.0) + sum(t.exchange_ms[:t.n_spatial])
                stages["n_spatial"] = t.n_spatial
                stages["launches"] = t.n_launches
        clocks.end()
        return total, stages

    # ---- warm-up (establishes temporal history), then the timed K steps: device-resident ----
    run_steps(args.warmup, 0)
    barrier(); wall0 = time.perf_counter()
    dev_ms, _ = run_steps(args.steps, args.warmup)
    barrier(); wall_ms = 1e3 * (time.perf_counter() - wall0)
    launches_per_frame = r.timings().n_launches

    # ---- per-stage times of the same steps (stage events split the frame; not used for `value`) ----
    stage_ms, stages = run_steps(args.steps, args.warmup + args.steps, stage_timing=True)

    # ---- e2e: host buffers through romis_render_frame semantics ----
    # One host image for the whole job: at N > 1 it lives in shared memory that every rank page-locks, each rank's band rows land
    # in it over that GPU's own PCIe link, and rank 0 -- the caller -- holds the complete frame (checked below).
    if world > 1:
        from romis_b200.api import SharedImage
        names = [None]
        if rank == 0:
            pinned = SharedImage(H, W); names[0] = pinned.name
        dist.broadcast_object_list(names, src=0)
        if rank != 0:
            pinned = SharedImage(H, W, name=names[0])
    else:
        pinned = PinnedImage(H, W)
    nxt = args.warmup + 2 * args.steps
    run_steps(2, nxt, host_out=pinned.array, lights="same"); nxt += 2
    barrier()
    e2e_ms, _ = run_steps(args.steps, nxt, host_out=pinned.array, lights="same"); nxt += args.steps
    barrier()
    run_steps(2, nxt, host_out=pinned.array, lights="edit"); nxt += 2
    barrier()
    edit_ms, _ = run_steps(args.steps, nxt, host_out=pinned.array, lights="edit"); nxt += args.steps
    barrier()
    if len(scene.lights):
        scene.lights["c0"][0, 0] = np.float32(edit_base)
    clk = clocks.stop() if rank == 0 else None
    # the caller's single image really is the whole frame: poison it, render one more frame, nothing of the poison may be left
    one_image = None
    if world > 1:
        if rank == 0:
            pinned.array[...] = np.float32(np.nan)
        barrier()
        run_steps(1, nxt, host_out=pinned.array); nxt += 1
        barrier()
        if rank == 0:
            one_image = bool(not np.isnan(pinned.array).any())
        barrier()

    exch_share = stages.get("exchange_ms", 0.0) / max(stage_ms, 1e-9)
    if world > 1:
        t = torch.tensor([dev_ms, e2e_ms, stage_ms, edit_ms, exch_share], dtype=torch.float64, device=device)
        dist.all_reduce(t, op=dist.ReduceOp.MAX)
        dev_ms, e2e_ms, stage_ms, edit_ms, exch_share = [float(x) for x in t.tolist()]
    if rank != 0:
        pinned.free()
        br.close()
        if world > 1:
            dist.destroy_process_group()
        return

    ms_per_step = dev_ms / args.steps
    fps = 1e3 / ms_per_step
    e2e_fps = 1e3 / (e2e_ms / args.steps)
    y0, y1 = br.band(H)
    px = (y1 - y0) * W
    peak, peak_src = measured_peak_gbs()
    pb = pass_bytes(N)
    per_pass = {}
    for name, key, cnt in (("primary", "primary_ms", 1), ("initial", "initial_ms", 1), ("temporal", "temporal_ms", 1),
                           ("spatial", "spatial_ms", max(1, stages.get("n_spatial", 1))), ("shade", "shade_ms", 1)):
        ms = stages.get(key, 0.0) / args.steps / cnt
        if ms > 0:
            gbs = pb[name] * px / (ms * 1e-3) / 1e9
            per_pass[name] = {"ms_per_launch": round(ms, 4), "algorithmic_bytes_per_px": pb[name], "achieved_gbs": round(gbs, 1), "frac": round(gbs / peak, 4)}
    # DRAM traffic per launch from the committed ncu --set full capture of this same command (profiles/): only meaningful for
    # the configuration and GPU count it was taken on
    traffic = {}
    tfiles = sorted(f for f in os.listdir(os.path.join(ROOT, "profiles")) if f.endswith("_traffic.json"))
    tpath = os.path.join(ROOT, "profiles", tfiles[-1]) if tfiles else ""          # the newest committed capture
    tj = {}
    if world == 1 and os.path.exists(tpath):
        tj = json.load(open(tpath))
        if tj.get("config") == args.config:
            for name in per_pass:
                k = tj["kernels"].get(name + "_kernel")
                if k:
                    per_pass[name]["dram_traffic_bytes"] = int(k["dram_bytes_read"] + k["dram_bytes_write"])
                    per_pass[name]["algorithmic_bytes"] = int(pb[name] * px)
                    per_pass[name]["ncu_warp_inst_per_cycle_per_sm_of_4"] = round(k["warp_inst_per_cycle_per_sm"], 2)
                    traffic[name] = per_pass[name]["dram_traffic_bytes"]
                    if k.get("thread_inst_executed"):
                        # the limiter these passes actually sit on: instruction issue.  achieved = thread instructions of one
                        # launch (ncu, same command) / the live CUDA-event launch time; peak = 148 SMs x 128 lanes x SM clock
                        sm_mhz = (clk or {}).get("sm_mhz") or 1965.0
                        peak_issue = 148 * 128 * sm_mhz * 1e6 / 1e12
                        ach = k["thread_inst_executed"] / (per_pass[name]["ms_per_launch"] * 1e-3) / 1e12
                        per_pass[name]["issue"] = {"bound": "issue", "achieved": round(ach, 3), "peak": round(peak_issue, 3), "unit": "T thread-inst/s",
                                                   "frac": round(ach / peak_issue, 4), "thread_inst_per_launch": int(k["thread_inst_executed"]),
                                                   "active_threads_per_warp_inst": round(k["thread_inst_executed"] / max(1.0, k.get("warp_inst_executed", 0.0)), 2)}
    dominant = max(per_pass, key=lambda k: per_pass[k]["ms_per_launch"] * (stages.get("n_spatial", 1) if k == "spatial" else 1))
    roofline = {"bound": "hbm", "kernel": dominant + "_kernel", "achieved": per_pass[dominant]["achieved_gbs"], "peak": peak, "unit": "GB/s",
                "frac": per_pass[dominant]["frac"], "traffic": traffic.get(dominant), "traffic_source": (tj.get("source") if traffic else None),
                "peak_source": peak_src,
                "note": "algorithmic bytes (SURVEY 8d) / CUDA-event launch time; the pass kernels are issue-bound by the parity-exact fp32/fp64 arithmetic, see DESIGN.md",
                "passes": per_pass}
    if per_pass[dominant].get("issue"):
        roofline["issue"] = dict(per_pass[dominant]["issue"], kernel=dominant + "_kernel",
                                 note="second roofline: thread instructions per launch (ncu capture of this command) / live launch time, against 148 SM x 128 lanes x measured SM clock")
    n_slots = r.light_archive_size()[0]
    edit_fps = 1e3 / (edit_ms / args.steps)
    line = {"metric": "ReSTIR frames/s", "value": fps, "unit": "frames/s", "n_gpus": world, "steps": args.steps, "warmup": args.warmup,
            "ms_per_step": ms_per_step, "higher_is_better": True, "scaling": "strong", "vs_baseline": None, "dtype": "f32",
            "data": "synthetic", "impl": "ours",
            "config": {"workload": label, "width": W, "height": H, "seed": SEED, "l2": "256 MiB memset before every step (outside the step's events)",
                       "sharding": (f"{world} row bands, halo = radius rows, " + ("pushed into peer-mapped (CUDA IPC) buffers over NVLink, flag-ordered" if br.transport == "peer" else "NCCL p2p" + (f" (peer mapping unavailable: {br.fallback_reason})" if br.fallback_reason else ""))) if world > 1 else "single GPU",
                       "band_edges": br.edges if br.edges is not None else "equal rows",
                       "band_calibration": getattr(br, "calibration", None)},
            "gcandidates_per_s": W * H * feat.initialLightSamples * fps / 1e9,
            # the same count against the initial RIS pass alone (rank 0's band when the frame is sharded)
            "gcandidates_per_s_initial_pass": (round(px * feat.initialLightSamples / (per_pass["initial"]["ms_per_launch"] * 1e-3) / 1e9, 2)
                                               if world == 1 and per_pass.get("initial", {}).get("ms_per_launch") else None),
            "wall_ms_per_step_incl_flush": wall_ms / args.steps,
            # headline e2e: the costlier of the two host sequences -- one light edited per frame (its old record archived for the
            # history, the history re-pointed, the new record sent); `static_lights`: the same table handed over and compared
            "e2e": {"value": edit_fps, "unit": "frames/s", "h2d_bytes_per_step": int(96 + 8 + 4 * n_slots + 256) * world,
                    "d2h_bytes_per_step": int(W * H * 12 + n_slots * world), "gcandidates_per_s": W * H * feat.initialLightSamples * edit_fps / 1e9,
                    "sequence": "per frame: scene.lights handed over with one light edited (romis_upload_lights: compare, archive, re-point history, send), frame, float RGB image read back to page-locked host memory",
                    "static_lights": {"value": e2e_fps, "unit": "frames/s", "h2d_bytes_per_step": 256 * world, "d2h_bytes_per_step": int(W * H * 12)},
                    # N > 1: every rank's band lands in ONE shared, page-locked host image held by rank 0 (poison check: no pixel left unwritten)
                    "one_host_image": one_image},
            "exchange_share_of_frame": (round(exch_share, 4) if world > 1 else None),
            "gpu_launches": int(launches_per_frame * args.steps),
            "clocks": clk, "roofline": roofline}
    if world == 1 and not args.no_cpu_baseline:
        line["cpu_baseline"] = cpu_baseline_leg(scene, W, H, feat, cam, args.config)
    print(json.dumps(line), flush=True)
    pinned.free()
    br.close()
    if world > 1:
        dist.destroy_process_group()


if __name__ == "__main__":
    main()
