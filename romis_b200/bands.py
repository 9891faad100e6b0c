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
    """Drives one `RestirRenderer` as rank `rank` of `world_size` row bands."""

    def __init__(self, renderer, rank: int, world_size: int, device: torch.device):
        self.r = renderer
        self.rank, self.world_size, self.device = rank, world_size, device
        self.stream = torch.cuda.ExternalStream(renderer.stream(), device=device)
        self._height = None

    def set_height(self, height: int, radius: int):
        if self._height != height:
            check_bands(height, self.world_size, radius)
            y0, y1 = band_rows(height, self.world_size, self.rank)
            self.r.set_band(y0, y1)
            self._height = height

    def render_frame(self, features, camera, W, H, history_valid, seed, frame, out=None):
        """The banded frame.  All work (kernels and transfers) is ordered on the renderer's own stream."""
        self.set_height(H, features.spatialResampleRadius if features.spatialReuse else 0)
        r = self.r
        with torch.cuda.stream(self.stream):
            r.frame_begin(features, camera, W, H, history_valid, seed, frame)
            if features.spatialReuse:
                for p in range(features.spatialResamplingPasses):
                    if self.world_size > 1:
                        t = [alias_device_bytes(*r.halo_region(w), self.device) for w in
                             (abi.ROMIS_HALO_SEND_LOW, abi.ROMIS_HALO_SEND_HIGH, abi.ROMIS_HALO_RECV_LOW, abi.ROMIS_HALO_RECV_HIGH)]
                        exchange_halos(t[0], t[1], t[2], t[3], self.rank, self.world_size)
                    r.frame_spatial_pass(p)
            r.frame_end(out)
