"""Frame sharding over ranks (one process per GPU, `torch.distributed`; NCCL on GPUs, gloo in CPU tests).

The hot path has no exchange step: frames are independent given the immutable map, which every rank
builds for itself from the same host arrays.  The only collectives are the optional ones a caller needs
when a batch lives on rank 0: scatter the frame batch, gather the poses / statuses.
Shards are contiguous; uneven batches are padded to the largest shard for the collective and trimmed after.
"""
from __future__ import annotations

import torch
import torch.distributed as dist


def shard_range(B: int, rank: int, world: int) -> tuple[int, int]:
    """Contiguous [lo, hi) of rank `rank` when B frames are split over `world` ranks (remainder to the first ranks)."""
    base, rem = divmod(B, world)
    lo = rank * base + min(rank, rem)
    return lo, lo + base + (1 if rank < rem else 0)


def _max_shard(B: int, world: int) -> int:
    return (B + world - 1) // world


def scatter_frames(batch: dict | None, B: int, src: int = 0, device=None, group=None) -> dict:
    """Rank `src` holds `batch` = dict of tensors whose leading dimension is B (pose0, pts_w, obs, octave);
    every rank returns its own contiguous shard.  Other ranks pass batch=None plus a `like` spec through
    the broadcast of shapes/dtypes done here."""
    rank, world = dist.get_rank(group), dist.get_world_size(group)
    spec = [None]
    if rank == src:
        spec = [{k: (tuple(v.shape[1:]), v.dtype) for k, v in batch.items()}]
    dist.broadcast_object_list(spec, src=src, group=group)
    S = _max_shard(B, world)
    lo, hi = shard_range(B, rank, world)
    out = {}
    for key, (tail, dtype) in spec[0].items():
        recv = torch.empty((S, *tail), dtype=dtype, device=device)
        chunks = None
        if rank == src:
            chunks = []
            full = batch[key].to(device) if device is not None else batch[key]
            for r in range(world):
                a, b = shard_range(B, r, world)
                c = torch.zeros((S, *tail), dtype=dtype, device=full.device)
                c[: b - a] = full[a:b]
                chunks.append(c)
        dist.scatter(recv, chunks, src=src, group=group)
        out[key] = recv[: hi - lo]
    return out


def gather_rows(local: torch.Tensor, B: int, dst: int = 0, group=None) -> torch.Tensor | None:
    """Inverse of the sharding for per-frame results (poses B x 7, status B, ...): rank `dst` gets the
    full tensor in frame order, the others None."""
    rank, world = dist.get_rank(group), dist.get_world_size(group)
    S = _max_shard(B, world)
    pad = torch.zeros((S, *local.shape[1:]), dtype=local.dtype, device=local.device)
    pad[: local.shape[0]] = local
    bufs = [torch.empty_like(pad) for _ in range(world)] if rank == dst else None
    dist.gather(pad, bufs, dst=dst, group=group)
    if rank != dst:
        return None
    parts = []
    for r in range(world):
        a, b = shard_range(B, r, world)
        parts.append(bufs[r][: b - a])
    return torch.cat(parts, dim=0)


def max_over_ranks(value: float, device=None, group=None) -> float:
    """Timing rule of the bench: the step time of a multi-GPU run is the max over ranks."""
    t = torch.tensor([value], dtype=torch.float64, device=device)
    if dist.is_initialized() and dist.get_world_size(group) > 1:
        dist.all_reduce(t, op=dist.ReduceOp.MAX, group=group)
    return float(t.item())
