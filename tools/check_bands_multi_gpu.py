#!/usr/bin/env python3
"""Run under torchrun on N GPUs: renders frames as N row bands (peer-mapped or NCCL halos, equal-cost or equal-row bands)
and checks on rank 0 that the assembled image equals the single-GPU frame bit for bit.

    python -m torch.distributed.run --nproc-per-node 2 --master-addr 127.0.0.1 tools/check_bands_multi_gpu.py [--halo nccl] [--equal-rows]
"""
import argparse, os, sys
import numpy as np
import torch
import torch.distributed as dist
sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
from romis_b200.api import RestirRenderer
from romis_b200.bands import BandedRenderer
from romis_b200.scene import Camera, Features, Scene

ap = argparse.ArgumentParser(); ap.add_argument("--halo", default="peer"); ap.add_argument("--equal-rows", action="store_true")
ap.add_argument("--width", type=int, default=640); ap.add_argument("--height", type=int, default=360); ap.add_argument("--frames", type=int, default=4)
ap.add_argument("--edit-lights", action="store_true", help="edit lights between frames (the archive slots must stay in step across the bands)")
ap.add_argument("--config", default="c2", choices=["c2", "c3"], help="c3: monkey + 65 536 synthetic lights (BASELINE configs[2])")
args = ap.parse_args()
rank, world, local = int(os.environ["RANK"]), int(os.environ["WORLD_SIZE"]), int(os.environ["LOCAL_RANK"])
torch.cuda.set_device(local); device = torch.device("cuda", local)
dist.init_process_group("nccl", device_id=device)
from romis_b200.scene import synthetic_lights
scenes = os.path.join(os.path.dirname(os.path.abspath(__file__)), "..", "tests", "golden", "scenes")
if args.config == "c3":
    scene = Scene.load(os.path.join(scenes, "Monkey.npz")); scene.lights = synthetic_lights(65536, seed=1)
    cam = Camera(50.0, 3.0, (0.0, 0.0, 0.0), (20.0, 20.0, 0.0))
else:
    scene = Scene.load(os.path.join(scenes, "CornellNightClub.npz")); cam = Camera()
feat = Features(spatialResamplingPasses=3, initialSamplesVisibilityCheck=True)
W, H = args.width, args.height
r = RestirRenderer(local); r.upload_scene(scene)
br = BandedRenderer(r, rank, world, device, transport=args.halo)
if not args.equal_rows:
    br.balance(cam, W, H, feat.spatialResampleRadius)
full = RestirRenderer(local) if rank == 0 else None
if full: full.upload_scene(scene)
ok = True
lights = scene.lights.copy()
for fr in range(args.frames):
    if args.edit_lights and fr >= 1:
        lights["c0"][fr::5] *= np.float32(0.9); lights["p0"][fr::11, 1] += np.float32(0.02)
        br.upload_lights(lights)
        if full: full.upload_lights(lights)
    img = np.zeros((H, W, 3), np.float32)
    br.render_frame(feat, cam, W, H, fr > 0, 123, fr, out=img)
    t = torch.from_numpy(img).to(device)
    dist.all_reduce(t)                       # bands are disjoint and zero elsewhere: the sum assembles the frame exactly
    if rank == 0:
        ref = full.render_frame(feat, cam, W, H, fr > 0, 123, fr)
        same = np.array_equal(t.cpu().numpy().view(np.uint32), ref.view(np.uint32))
        print(f"frame {fr}: banded ({world} GPUs, {args.halo}, edges {br.edges or 'equal'}) == single GPU: {same}", flush=True)
        ok &= same
if r.peer_timed_out():
    print(f"rank {rank}: a halo flag wait timed out", flush=True); ok = False
flag = torch.tensor([int(ok)], device=device); dist.all_reduce(flag, op=dist.ReduceOp.MIN)
br.close()
dist.destroy_process_group()
sys.exit(0 if int(flag.item()) == 1 else 1)
