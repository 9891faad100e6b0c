#!/usr/bin/env python3
"""BASELINE C5: spatial-reuse sweep k in {3,5,10} x radius in {10,30} x passes 1..4 at 3840x2160 on a synthetic 2^20-light
scene.  Alone: one GPU.  Under torchrun (one process per GPU): the frame as row bands with the fused halo exchange, edges cut
by the per-row hit profile once; time = device time per frame, max over the ranks.  Writes a markdown table to stdout (rank 0).

    python tools/sweep_c5.py [lights]
    python -m torch.distributed.run --nproc-per-node 8 --master-addr 127.0.0.1 tools/sweep_c5.py [lights]
"""
import os, sys
import numpy as np
sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
from romis_b200.api import RestirRenderer
from romis_b200.scene import Camera, Features, Scene, synthetic_lights

rank, world, local = int(os.environ.get("RANK", 0)), int(os.environ.get("WORLD_SIZE", 1)), int(os.environ.get("LOCAL_RANK", 0))
L = int(sys.argv[1]) if len(sys.argv) > 1 else 1 << 20
W, H = 3840, 2160
scene = Scene.load(os.path.join(os.path.dirname(__file__), "..", "tests", "golden", "scenes", "CornellNightClub.npz"))
scene.lights = synthetic_lights(L, seed=1, intensity=4000.0)
scene.lights["p0"] += np.array([2.5, 2.0, -1.0], np.float32)
cam = Camera()
br = None
if world > 1:
    import torch, torch.distributed as dist
    from romis_b200.bands import BandedRenderer
    torch.cuda.set_device(local); device = torch.device("cuda", local)
    dist.init_process_group("nccl", device_id=device)
r = RestirRenderer(local); r.upload_scene(scene); r.set_stage_timing(True)
if world > 1:
    br = BandedRenderer(r, rank, world, device)
    br.balance(cam, W, H, 30)          # one cut for the whole sweep (bands of >= 30 rows: the largest radius)
if rank == 0:
    print(f"C5 sweep, {L} lights, {W}x{H}, M=32, N=2, {world} GPU(s)" + (f", band edges {br.edges}" if br else ""))
    print("| k | radius | passes | frame ms | initial ms | spatial ms / pass | stage-0 halo push ms | frames/s | G-candidates/s |\n|---|---|---|---|---|---|---|---|---|", flush=True)
for k in (3, 5, 10):
    for rad in (10, 30):
        for P in (1, 2, 3, 4):
            feat = Features(numNeighboursToSample=k, spatialResampleRadius=rad, spatialResamplingPasses=P, initialSamplesVisibilityCheck=True)
            acc = []
            for fr in range(6):
                if br:
                    dist.barrier(); torch.cuda.synchronize()
                    br.render_frame(feat, cam, W, H, fr > 0, 7, fr, out=None); r.synchronize()
                else:
                    r.render_frame(feat, cam, W, H, fr > 0, 7, fr, want_image=False)
                t = r.timings()
                if fr >= 2: acc.append([t.total_ms, t.initial_ms + t.temporal_ms, sum(t.spatial_ms[:P]) / P, sum(t.exchange_ms[:P])])
            row = np.mean(np.array(acc), axis=0)
            if br:
                v = torch.tensor(row, dtype=torch.float64, device=device)
                dist.all_reduce(v, op=dist.ReduceOp.MAX)
                row = v.cpu().numpy()
            ms = float(row[0])
            if rank == 0:
                print(f"| {k} | {rad} | {P} | {ms:.2f} | {row[1]:.2f} | {row[2]:.2f} | {row[3]:.3f} | {1e3 / ms:.1f} | {W * H * 32 / ms / 1e6:.1f} |", flush=True)
if br:
    br.close(); dist.destroy_process_group()
