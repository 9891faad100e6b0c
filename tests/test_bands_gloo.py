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
