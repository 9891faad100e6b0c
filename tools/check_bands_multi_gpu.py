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
args = ap.parse_args()
rank, world, local = int(os.environ["RANK"]), int(os.environ["WORLD_SIZE"]), int(os.environ["LOCAL_RANK"])
torch.cuda.set_device(local); device = torch.device("cuda", local)
dist.init_process_group("nccl", device_id=device)
scene = Scene.load(os.path.join(os.path.dirname(__file__), "..", "tests", "golden", "scenes", "CornellNightClub.npz"))
feat = Features(spatialResamplingPasses=3, initialSamplesVisibilityCheck=True)
cam = Camera(); W, H = args.width, args.height
r = RestirRenderer(local); r.upload_scene(scene)
br = BandedRenderer(r, rank, world, device, transport=args.halo)
if not args.equal_rows:
    br.balance(cam, W, H, feat.spatialResampleRadius)
full = RestirRenderer(local) if rank == 0 else None
if full: full.upload_scene(scene)
ok = True
for fr in range(args.frames):
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
dist.destroy_process_group()
sys.exit(0 if int(flag.item()) == 1 else 1)
