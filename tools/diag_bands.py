#!/usr/bin/env python3
"""Per-rank per-stage timings of the banded C2 frame (diagnostic). Run under torchrun, or alone (single band = half frame)."""
import os, sys, time
import numpy as np, torch, torch.distributed as dist
sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
from romis_b200.api import RestirRenderer
from romis_b200.bands import BandedRenderer
from romis_b200.scene import Camera, Features, Scene
rank, world, local = int(os.environ.get("RANK", 0)), int(os.environ.get("WORLD_SIZE", 1)), int(os.environ.get("LOCAL_RANK", 0))
torch.cuda.set_device(local); device = torch.device("cuda", local)
if world > 1: dist.init_process_group("nccl", device_id=device)
scene = Scene.load(os.path.join(os.path.dirname(__file__), "..", "tests", "golden", "scenes", "CornellNightClub.npz"))
feat = Features(spatialResamplingPasses=3, initialSamplesVisibilityCheck=True); cam = Camera(); W, H = 1920, 1080
r = RestirRenderer(local); r.upload_scene(scene); r.set_stage_timing(True)
br = BandedRenderer(r, rank, world, device, transport=os.environ.get("HALO", "peer"))
if world > 1 and os.environ.get("BALANCE", "1") == "1":
    br.balance(cam, W, H, 10); br.calibrate(feat, cam, W, H); r.set_stage_timing(True)
if world == 1 and len(sys.argv) > 2: r.set_band(int(sys.argv[1]), int(sys.argv[2])); br._height = H
acc = None
for fr in range(12):
    if world > 1: dist.barrier()
    torch.cuda.synchronize()
    br.render_frame(feat, cam, W, H, fr > 0, 1, fr, out=None); r.synchronize()
    t = r.timings()
    row = np.array([t.total_ms, t.primary_ms, t.initial_ms, t.temporal_ms, *t.spatial_ms[:3], t.shade_ms, *t.exchange_ms[:3]])
    if fr >= 4: acc = row if acc is None else acc + row
acc /= 8
print(f"rank {rank} band {br.band(H) if world > 1 else sys.argv[1:3]}: total {acc[0]:.3f} primary {acc[1]:.3f} initial {acc[2]:.3f} temporal {acc[3]:.3f} spatial {acc[4]:.3f} {acc[5]:.3f} {acc[6]:.3f} shade {acc[7]:.3f} exchange {acc[8]:.3f} {acc[9]:.3f} {acc[10]:.3f}", flush=True)
if world > 1: dist.destroy_process_group()
