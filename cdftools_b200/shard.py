"""Host-side sharding of the MOC hot path over the GPUs of one box (one process per GPU, torch.distributed).

The path has NO exchange step: every output element (basin, j, k|bin, record) depends on one latitude row of one
record.  Work is therefore partitioned, never reduced:

  * time sharding      rank r owns records jt = r, r+n, r+2n, ...          (BASELINE config 4)
  * latitude bands     rank r owns rows [j0_r, j1_r) of every record        (BASELINE config 5)

and the only communication is the gather of the per-rank result slabs to rank 0, which writes the output file.
The gather runs on whatever backend the process group has: NCCL over NVLink on the GPU box, gloo in the CPU tests.
"""
from __future__ import annotations

from typing import List, Sequence, Tuple

import torch
import torch.distributed as dist


def time_plan(nrec: int, world: int, rank: int) -> List[int]:
    """Records (0-based jt) owned by `rank`: round-robin, so every rank streams the file front to back."""
    return list(range(rank, nrec, world))


def band_plan(ny: int, world: int) -> List[Tuple[int, int]]:
    """Balanced latitude bands [j0, j1) for every rank; bands differ by at most one row and cover [0, ny)."""
    base, extra = divmod(ny, world)
    out, j0 = [], 0
    for r in range(world):
        j1 = j0 + base + (1 if r < extra else 0)
        out.append((j0, j1))
        j0 = j1
    return out


def _world(group=None):
    if not dist.is_available() or not dist.is_initialized():
        return 1, 0
    return dist.get_world_size(group), dist.get_rank(group)


def gather_time_slabs(local: torch.Tensor, nrec: int, dst: int = 0, group=None):
    """local: [len(time_plan(nrec, world, rank)), ...] results of this rank's records.
    Returns on `dst` the [nrec, ...] tensor in record order, None elsewhere."""
    world, rank = _world(group)
    if world == 1:
        return local
    per = (nrec + world - 1) // world
    pad = torch.zeros((per,) + tuple(local.shape[1:]), dtype=local.dtype, device=local.device)
    pad[: local.shape[0]] = local
    bufs = [torch.empty_like(pad) for _ in range(world)] if rank == dst else None
    dist.gather(pad, bufs, dst=dst, group=group)
    if rank != dst:
        return None
    out = torch.empty((nrec,) + tuple(local.shape[1:]), dtype=local.dtype, device=local.device)
    for r in range(world):
        idx = time_plan(nrec, world, r)
        out[idx] = bufs[r][: len(idx)]
    return out


def gather_band_slabs(local: torch.Tensor, ny: int, row_dim: int, dst: int = 0, group=None):
    """local: this rank's band with its rows along `row_dim` (size j1-j0).  Returns on `dst` the tensor with all
    ny rows (bands concatenated along row_dim), None elsewhere.  Bands are padded to equal size for the collective."""
    world, rank = _world(group)
    if world == 1:
        return local
    bands = band_plan(ny, world)
    mx = max(j1 - j0 for j0, j1 in bands)
    shape = list(local.shape)
    shape[row_dim] = mx
    pad = torch.zeros(shape, dtype=local.dtype, device=local.device)
    pad.narrow(row_dim, 0, local.shape[row_dim]).copy_(local)
    bufs = [torch.empty_like(pad) for _ in range(world)] if rank == dst else None
    dist.gather(pad, bufs, dst=dst, group=group)
    if rank != dst:
        return None
    return torch.cat([bufs[r].narrow(row_dim, 0, bands[r][1] - bands[r][0]) for r in range(world)], dim=row_dim)


def band_slice(arr, j0: int, j1: int, row_axis: int):
    """Contiguous copy of rows [j0, j1) of a numpy array along row_axis (what a rank hands to *_setup / *_submit)."""
    import numpy as np
    sl = [slice(None)] * arr.ndim
    sl[row_axis] = slice(j0, j1)
    return np.ascontiguousarray(arr[tuple(sl)])
