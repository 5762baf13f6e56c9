"""Frame-chunk data parallelism.

The reference runs contiguous frame batches sequentially and independently (preprocess/fit_SMPLH_kpts.py:106-112,
recon/recon_fit_triplane.py:47) and recommends separate jobs per chunk (README.md:52); results meet on the file system, where
``smoothnet/smooth_smplt.py:122-143`` re-reads one pickle per frame.  Here one process per GPU owns a contiguous run of WHOLE
reference batches (temporal terms couple frames only inside a batch, so per-batch results are unchanged), and the per-chunk
trajectories are stitched with a single ``all_gather`` (NCCL over NVLink on the GPU box, gloo in the CPU tests): fp32
[T_local, 169] = pose 156 | betas 10 | trans 3 for SMPL-T, padded to the longest shard plus a length vector.
"""
from __future__ import annotations

from typing import List, Tuple

import torch
import torch.distributed as dist

SMPLT_WIDTH = 156 + 10 + 3


def batch_bounds(start: int, end: int, bs: int) -> List[Tuple[int, int]]:
    """The reference's batch boundaries for frames [start, end): ``range(start, end, bs)`` (fit_SMPLH_kpts.py:106-110)."""
    return [(b, min(end, b + bs)) for b in range(start, end, bs)]


def shard_batches(n_batches: int, world: int, rank: int) -> range:
    """Contiguous, near-equal split of whole batches: rank r gets batches [lo, hi)."""
    base, extra = divmod(n_batches, world)
    lo = rank * base + min(rank, extra)
    return range(lo, lo + base + (1 if rank < extra else 0))


def rank_frames(start: int, end: int, bs: int, world: int, rank: int) -> List[Tuple[int, int]]:
    bounds = batch_bounds(start, end, bs)
    return [bounds[i] for i in shard_batches(len(bounds), world, rank)]


def pack_smplt(pose: torch.Tensor, betas: torch.Tensor, trans: torch.Tensor) -> torch.Tensor:
    return torch.cat([pose, betas, trans], 1).float().contiguous()


def unpack_smplt(traj: torch.Tensor):
    return traj[:, :156], traj[:, 156:166], traj[:, 166:169]


def gather_trajectory(local: torch.Tensor, group=None) -> torch.Tensor:
    """All ranks contribute their [T_r, W] block (frame order = rank order, ragged T_r) and every rank receives the full
    [sum T_r, W] trajectory.  One collective for the lengths (8 bytes per rank) and one for the padded payload."""
    if not dist.is_available() or not dist.is_initialized() or dist.get_world_size(group) == 1:
        return local
    world = dist.get_world_size(group)
    n = torch.tensor([local.shape[0]], dtype=torch.int64, device=local.device)
    lens = [torch.zeros_like(n) for _ in range(world)]
    dist.all_gather(lens, n, group=group)
    lens = [int(x.item()) for x in lens]
    width, tmax = local.shape[1], max(max(lens), 1)
    padded = torch.zeros(tmax, width, dtype=local.dtype, device=local.device)
    padded[: local.shape[0]] = local
    out = torch.empty(world * tmax, width, dtype=local.dtype, device=local.device)
    dist.all_gather_into_tensor(out, padded, group=group)
    return torch.cat([out[r * tmax: r * tmax + lens[r]] for r in range(world)], 0)
