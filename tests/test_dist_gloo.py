"""N > 1 host-side path on CPU: world_size-2 gloo run of the frame sharding (scatter a batch from rank 0,
gather per-frame results back, max-over-ranks timing).  No compute call — the kernels need a GPU."""
import os
import socket

import numpy as np
import pytest
import torch
import torch.distributed as dist
import torch.multiprocessing as mp

from gmmloc_b200 import dist as gd
from gmmloc_b200 import synth


def _free_port():
    with socket.socket() as s:
        s.bind(("127.0.0.1", 0))
        return s.getsockname()[1]


def _worker(rank, world, port, B, N, q):
    os.environ["MASTER_ADDR"] = "127.0.0.1"
    os.environ["MASTER_PORT"] = str(port)
    dist.init_process_group("gloo", rank=rank, world_size=world)
    try:
        batch = None
        if rank == 0:
            mean, cov = synth.make_map(200, 0.5, seed=1)
            fr = synth.make_frames(mean, cov, B, N, seed=2)
            batch = {k: torch.from_numpy(fr[k]) for k in ("pose0", "pts_w", "obs", "octave")}
        shard = gd.scatter_frames(batch, B, src=0)
        lo, hi = gd.shard_range(B, rank, world)
        assert shard["pts_w"].shape == (hi - lo, N, 3) and shard["pts_w"].dtype == torch.float32
        assert shard["octave"].dtype == torch.uint8 and shard["pose0"].dtype == torch.float64
        # a per-frame "result": the frame's global index and a checksum of its landmarks
        res = torch.stack([torch.arange(lo, hi, dtype=torch.float64), shard["pts_w"].double().sum(dim=(1, 2))], dim=1)
        full = gd.gather_rows(res, B, dst=0)
        tmax = gd.max_over_ranks(float(rank + 1))
        assert tmax == float(world)
        # the pre-allocated plan bench.py / configs[3] uses: same shards, results back in frame order
        plan = gd.FrameSharder(B, N)
        for _ in range(2):  # reusable: the second round goes through the cached chunk views
            sh = plan.scatter(batch)
            assert plan.count == hi - lo
            for k in ("pose0", "pts_w", "obs", "octave"):
                assert torch.equal(sh[k][: plan.count], shard[k]), k
            pose = torch.zeros((plan.S, 7), dtype=torch.float64)
            pose[: plan.count] = sh["pose0"][: plan.count] * 2.0
            status = torch.full((plan.S,), -7, dtype=torch.int32)
            status[: plan.count] = torch.arange(lo, hi, dtype=torch.int32)
            gp, gs = plan.gather(pose, status)
            if rank == 0:
                assert torch.equal(gp, batch["pose0"] * 2.0) and torch.equal(gs, torch.arange(B, dtype=torch.int32))
            else:
                assert gp is None and gs is None
        # the same shards scattered piece by piece (what bench.py --workload c4 pipelines against the solve)
        plan2 = gd.FrameSharder(B, N)
        for nsub in (1, 2, 3):
            for buf in plan2.recv.values():
                buf.zero_()
            pieces = plan2.sub_ranges(nsub)
            assert pieces[0][0] == 0 and pieces[-1][1] == plan2.S and all(a[1] == b[0] for a, b in zip(pieces, pieces[1:]))
            for a, b in pieces:
                plan2.scatter_piece(batch, a, b)
            for k in ("pose0", "pts_w", "obs", "octave"):
                assert torch.equal(plan2.recv[k][: plan2.count], shard[k]), (k, nsub)
        if rank == 0:
            assert full.shape == (B, 2)
            assert torch.equal(full[:, 0], torch.arange(B, dtype=torch.float64))
            assert torch.equal(full[:, 1], batch["pts_w"].double().sum(dim=(1, 2)))
            q.put("ok")
        else:
            assert full is None
    finally:
        dist.destroy_process_group()


@pytest.mark.parametrize("B", [8, 7, 1])
def test_scatter_gather_world2(B):
    ctx = mp.get_context("spawn")
    q = ctx.SimpleQueue()
    port = _free_port()
    procs = [ctx.Process(target=_worker, args=(r, 2, port, B, 16, q)) for r in range(2)]
    [p.start() for p in procs]
    [p.join(120) for p in procs]
    assert all(p.exitcode == 0 for p in procs)
    assert q.get() == "ok"


def test_shard_range_partitions_everything():
    for B in (0, 1, 5, 8, 8192, 8191):
        for world in (1, 2, 3, 8):
            spans = [gd.shard_range(B, r, world) for r in range(world)]
            assert spans[0][0] == 0 and spans[-1][1] == B
            assert all(spans[i][1] == spans[i + 1][0] for i in range(world - 1))
            sizes = [b - a for a, b in spans]
            assert max(sizes) - min(sizes) <= 1
