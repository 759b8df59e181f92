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


class FrameSharder:
    """Pre-allocated plan for BASELINE configs[3]: ONE batch of B frames resident on rank `src` is scattered to the
    ranks (contiguous shards, `dist.scatter` = grouped ncclSend / ncclRecv over NVLink on GPUs), every rank solves its
    shard, and the per-frame results are gathered back to `src`.  All buffers are allocated once, so `scatter()` /
    `gather()` only enqueue collectives (they may sit inside a timed region).  Works with NCCL on device tensors and
    with gloo on CPU tensors (tests/test_dist_gloo.py).

        plan = FrameSharder(B, N, device)
        shard = plan.scatter(full)            # full: dict of B-leading tensors on rank src, None elsewhere
        ... solve shard[...][:plan.count] ...
        pose, status = plan.gather(pose_shard, status_shard)   # full-size tensors on rank src, (None, None) elsewhere
    """

    KEYS = {"pose0": ((7,), torch.float64), "pts_w": (None, torch.float32), "obs": (None, torch.float32), "octave": (None, torch.uint8)}

    def __init__(self, B: int, N: int, device=None, src: int = 0, group=None):
        self.B, self.N, self.src, self.group, self.device = B, N, src, group, device
        self.rank, self.world = dist.get_rank(group), dist.get_world_size(group)
        self.S = _max_shard(B, self.world)
        self.lo, self.hi = shard_range(B, self.rank, self.world)
        self.count = self.hi - self.lo
        tails = {"pose0": (7,), "pts_w": (N, 3), "obs": (N, 3), "octave": (N,)}
        self.recv = {k: torch.empty((self.S, *tails[k]), dtype=self.KEYS[k][1], device=device) for k in tails}
        self._chunks = None
        self._chunks_of = None
        self._g_pose = self._g_status = None
        if self.rank == src:
            self._g_pose = [torch.empty((self.S, 7), dtype=torch.float64, device=device) for _ in range(self.world)]
            self._g_status = [torch.empty((self.S,), dtype=torch.int32, device=device) for _ in range(self.world)]

    def _make_chunks(self, full: dict):
        chunks = {}
        for k, buf in self.recv.items():
            lst = []
            for r in range(self.world):
                a, b = shard_range(self.B, r, self.world)
                if b - a == self.S:
                    lst.append(full[k][a:b])  # a view of the resident batch: no copy on the source side
                else:  # ragged tail: padded staging copy (made once per batch object)
                    c = torch.zeros_like(buf)
                    c[: b - a] = full[k][a:b]
                    lst.append(c)
            chunks[k] = lst
        return chunks

    def scatter(self, full: dict | None) -> dict:
        if self.rank == self.src:
            if self._chunks_of is not full:
                self._chunks, self._chunks_of = self._make_chunks(full), full
        for k, buf in self.recv.items():
            dist.scatter(buf, self._chunks[k] if self.rank == self.src else None, src=self.src, group=self.group)
        return self.recv

    def sub_ranges(self, nsub: int) -> list[tuple[int, int]]:
        """[lo, hi) rows of a shard buffer (S rows) cut into `nsub` pieces of equal size (the last takes the remainder)."""
        nsub = max(1, min(int(nsub), self.S))
        step = -(-self.S // nsub)
        return [(a, min(a + step, self.S)) for a in range(0, self.S, step)]

    def scatter_piece(self, full: dict | None, lo: int, hi: int) -> None:
        """Scatters rows [lo, hi) of every rank's shard only (into the same receive buffers as scatter()).  A caller that
        cuts the shard into pieces can enqueue piece i + 1 on a communication stream while piece i is being solved:
        the batch leaves rank 0 through ONE NVLink port, so at 8 GPUs the scatter is a third of the configs[3] step."""
        if self.rank == self.src and self._chunks_of is not full:
            self._chunks, self._chunks_of = self._make_chunks(full), full
        for k, buf in self.recv.items():
            dist.scatter(buf[lo:hi], [c[lo:hi] for c in self._chunks[k]] if self.rank == self.src else None, src=self.src,
                         group=self.group)

    # ---- peer-memory scatter: every rank PULLS its shard out of rank src's HBM over NVLink with the copy engines ----
    def share(self, full: dict | None) -> None:
        """Rank `src` exports the resident batch through CUDA IPC handles (broadcast as picklable objects), the other ranks
        map it.  One process per GPU, one node.  After this, pull_piece() needs no collective and no SM: the copies are
        peer-to-peer cudaMemcpyAsync on the copy engines, so — unlike NCCL's send / receive kernels — they overlap a
        persistent compute kernel that holds every SM."""
        from torch.multiprocessing.reductions import reduce_tensor
        objs = [None]
        if self.rank == self.src:
            objs = [{k: reduce_tensor(full[k]) for k in self.recv}]
        dist.broadcast_object_list(objs, src=self.src, group=self.group)
        if self.rank == self.src:
            self._peer = {k: full[k] for k in self.recv}
        else:
            self._peer = {k: fn(*args) for k, (fn, args) in objs[0].items()}

    def pull_piece(self, lo: int, hi: int) -> None:
        """Copies rows [lo, hi) of THIS rank's shard from the shared batch into the receive buffers (current stream)."""
        hi = min(hi, self.count)
        if hi <= lo:
            return
        a = self.lo
        for k, buf in self.recv.items():
            buf[lo:hi].copy_(self._peer[k][a + lo:a + hi], non_blocking=True)

    def gather(self, pose: torch.Tensor, status: torch.Tensor):
        """pose S x 7 f64, status S i32 (rows >= count are padding).  Returns (B x 7, B) on rank src."""
        dist.gather(pose, self._g_pose, dst=self.src, group=self.group)
        dist.gather(status, self._g_status, dst=self.src, group=self.group)
        if self.rank != self.src:
            return None, None
        spans = [shard_range(self.B, r, self.world) for r in range(self.world)]
        return (torch.cat([self._g_pose[r][: b - a] for r, (a, b) in enumerate(spans)]),
                torch.cat([self._g_status[r][: b - a] for r, (a, b) in enumerate(spans)]))


def max_over_ranks(value: float, device=None, group=None) -> float:
    """Timing rule of the bench: the step time of a multi-GPU run is the max over ranks."""
    t = torch.tensor([value], dtype=torch.float64, device=device)
    if dist.is_initialized() and dist.get_world_size(group) > 1:
        dist.all_reduce(t, op=dist.ReduceOp.MAX, group=group)
    return float(t.item())
