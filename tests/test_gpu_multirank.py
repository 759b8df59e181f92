"""BASELINE configs[3] on real GPUs (needs >= 2; skipped on a 1-GPU box): a batch resident on rank 0 is scattered over NCCL
(gmmloc_b200.dist.FrameSharder), every rank solves its shard through the C ABI, poses / status are gathered, and rank 0
checks the gathered result bit for bit against its own solve of the whole batch (and a few frames against the oracle)."""
import os
import socket

import numpy as np
import pytest
import torch

pytestmark = pytest.mark.gpu


def _free_port():
    with socket.socket() as s:
        s.bind(("127.0.0.1", 0))
        return s.getsockname()[1]


def _worker(rank, world, port, B, N, q, mode="nccl"):
    import torch.distributed as dist
    import gmmloc_b200 as g
    from gmmloc_b200 import dist as gd
    from gmmloc_b200 import synth
    os.environ["MASTER_ADDR"] = "127.0.0.1"
    os.environ["MASTER_PORT"] = str(port)
    torch.cuda.set_device(rank)
    dev = torch.device("cuda", rank)
    dist.init_process_group("nccl", rank=rank, world_size=world, device_id=dev)
    try:
        mean, cov = synth.make_map(3000, 0.5, seed=21)  # the map is replicated: every rank builds it from the same arrays
        gmm, ctx = g.GMM(mean, cov, device=rank), g.Context(rank)
        cam, prm = g.v1_camera(), g.default_params()
        full, fr = None, None
        if rank == 0:
            fr = synth.make_frames(mean, cov, B, N, seed=22)
            full = {k: torch.from_numpy(fr[k]).to(dev) for k in ("pose0", "pts_w", "obs", "octave")}
        plan = gd.FrameSharder(B, N, device=dev)
        stream = torch.cuda.ExternalStream(ctx.stream, device=dev)
        o_pose = torch.zeros((plan.S, 7), dtype=torch.float64, device=dev)
        o_status = torch.zeros((plan.S,), dtype=torch.int32, device=dev)
        o_assign = torch.empty((plan.S, N), dtype=torch.int32, device=dev)
        if mode == "p2p":  # peer-memory pull over CUDA IPC instead of the NCCL scatter (same receive buffers)
            plan.share(full)
            torch.cuda.synchronize()
            dist.barrier()
        with torch.cuda.stream(stream):
            if mode == "p2p":
                for lo, hi in plan.sub_ranges(3):
                    plan.pull_piece(lo, hi)
                sh = plan.recv
            else:
                sh = plan.scatter(full)
            if plan.count:
                ctx.frames_solve_async(gmm, cam, prm, sh["pose0"].data_ptr(), sh["pts_w"].data_ptr(), sh["obs"].data_ptr(),
                                       sh["octave"].data_ptr(), plan.count, N, o_pose.data_ptr(), 0, o_assign.data_ptr(), 0,
                                       o_status.data_ptr())
            gp, gs = plan.gather(o_pose, o_status)
        torch.cuda.synchronize()
        if rank == 0:
            one = ctx.frames_solve(gmm, fr["pose0"], fr["pts_w"], fr["obs"], fr["octave"])
            assert np.array_equal(gp.cpu().numpy(), one["pose"]), "gathered poses differ from the 1-GPU solve"
            assert np.array_equal(gs.cpu().numpy(), one["status"])
            from oracle import pyoracle as O
            comps = O.map_build(mean, cov)
            exp = O.frames_solve(comps, fr["pose0"][:4], fr["pts_w"][:4], fr["obs"][:4], fr["octave"][:4])
            assert np.abs(gp.cpu().numpy()[:4] - exp["pose"]).max() < 1e-5
            q.put("ok")
    finally:
        dist.destroy_process_group()


@pytest.mark.parametrize("mode", ["nccl", "p2p"])
@pytest.mark.parametrize("B", [64, 37])
def test_scatter_solve_gather_equals_single_gpu(B, mode):
    if torch.cuda.device_count() < 2:
        pytest.skip("needs 2 GPUs (gpurun --gpus 2)")
    import torch.multiprocessing as mp
    ctx = mp.get_context("spawn")
    q = ctx.SimpleQueue()
    port = _free_port()
    procs = [ctx.Process(target=_worker, args=(r, 2, port, B, 600, q, mode)) for r in range(2)]
    [p.start() for p in procs]
    [p.join(600) for p in procs]
    assert all(p.exitcode == 0 for p in procs)
    assert q.get() == "ok"
