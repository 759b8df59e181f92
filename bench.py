#!/usr/bin/env python
"""bench.py — frames/s of the GMMLoc hot path (GMM association + structure-constrained GN solve).

    python bench.py [--gpus N] [--steps K] [--warmup W] [--impl reference]
    python -m torch.distributed.run --nnodes=1 --nproc-per-node N ... bench.py --gpus N ...

One "step" = one pass of the hot path (association scan + exact re-rank + 10-iteration frame GN) over one
batch of B synthetic frames per GPU.  Workload = BASELINE.json configs[1]: 2 000 landmarks x 10 000
components, 10 GN iterations, fp32 scan (fp64 re-rank + fp64 solve), frames sharded over ranks with no
data-path collective (weak scaling: B frames per GPU).

 value : device-resident throughput (inputs already in HBM), CUDA events on the launching stream,
         max over ranks.
 e2e   : the same metric through the host C ABI (gmmloc_frames_solve) with page-locked HOST buffers:
         H2D of every input and D2H of poses / assignments / outlier flags / status inside the timed region.
 roofline      : association scan (the dominant kernel) against the measured HBM copy peak, using the
                 algorithmic bytes of SURVEY §8d (40 K + 16 N per frame) — the figure BASELINE.json asks
                 for.  The kernel is FP32-issue-bound, not HBM-bound (arithmetic intensity ~1e3 flop/B),
 roofline_fp32 : so the meaningful roof is reported next to it: 23 N K flop per frame (as the reference
                 writes the primitive, gaussian.cpp:65-70) over the measured FFMA peak of tools/ubench_pipes.
 cpu_baseline  : the CPU oracle ("port": the reference cannot be compiled here) on a bounded sample.

--impl reference times the oracle (the reference's algorithm on the host cores) on the same config.
"""
from __future__ import annotations

import argparse
import json
import os
import statistics
import subprocess
import sys
import time
from pathlib import Path

import numpy as np

ROOT = Path(__file__).resolve().parent
sys.path.insert(0, str(ROOT))

N_LM, K_COMP, GN_ITERS = 2000, 10000, 10
B_DEFAULT = 4096
FFMA_PEAK_TFLOPS = 68.97  # measured: tools/ubench_pipes.cu on this pool's B200 (profiles/ubench_pipes_r01.txt)


def algorithmic_bytes_per_frame(n, k):
    return 40 * k + 16 * n  # mu 12 + sym cov_inv 24 + det 4 per component; 12 in + 4 out per landmark


def algorithmic_flops_per_frame(n, k):
    return 23 * n * k  # gaussian.cpp:67-69 as written: 3 sub, 15 mat-vec, 5 dot


def measured_peaks():
    p = ROOT / "MEASURED_PEAKS.json"
    if p.exists():
        d = json.loads(p.read_text())
        return float(d["hbm_gbs"]), "measured"
    return 6650.0, "fallback"


def config_dict(B, world):
    return {"workload": "configs[1]: 2000 landmarks x 10000 GMM components, 10 GN iters, fp32 scan + fp64 re-rank/solve",
            "landmarks": N_LM, "components": K_COMP, "gn_iters": GN_ITERS, "frames_per_gpu_per_step": B,
            "global_frames_per_step": B * world, "parallelism": f"frame-sharded x{world}, map replicated, no data-path collective",
            "l2": f"inputs {B * (N_LM * 25 + 56) / 1e6:.0f} MB per step > 126 MB L2 (no flush needed)"}


class ClockSampler:
    Q = ("clocks.sm,clocks.max.sm,power.draw,clocks_event_reasons.hw_slowdown,clocks_event_reasons.hw_thermal_slowdown,"
         "clocks_event_reasons.sw_thermal_slowdown,clocks_event_reasons.sw_power_cap")

    def __init__(self, device):
        self.proc = None
        try:
            self.proc = subprocess.Popen(["nvidia-smi", f"--query-gpu={self.Q}", "--format=csv,noheader,nounits", "-lms", "100",
                                          "-i", str(device)], stdout=subprocess.PIPE, stderr=subprocess.DEVNULL, text=True)
        except OSError:
            pass

    def stop(self):
        if self.proc is None:
            return {"sm_mhz": None, "sm_max_mhz": None, "reasons": ["nvidia-smi unavailable"]}
        self.proc.terminate()
        try:
            out, _ = self.proc.communicate(timeout=5)
        except subprocess.TimeoutExpired:
            self.proc.kill()
            out, _ = self.proc.communicate()
        sm, mx, reasons = [], [], set()
        names = ["hw_slowdown", "hw_thermal_slowdown", "sw_thermal_slowdown", "sw_power_cap"]
        for line in out.splitlines():
            f = [x.strip() for x in line.split(",")]
            if len(f) < 7:
                continue
            try:
                sm.append(float(f[0])); mx.append(float(f[1]))
            except ValueError:
                continue
            for n, v in zip(names, f[3:7]):
                if v.lower().startswith("active"):
                    reasons.add(n)
        return {"sm_mhz": statistics.median(sm) if sm else None, "sm_max_mhz": max(mx) if mx else None,
                "reasons": sorted(reasons), "samples": len(sm)}


def cpu_sample(mean, cov, fr, cores, target_s=15.0):
    """Times the oracle on a bounded sample of the same workload; returns (frames/s, n_frames, seconds)."""
    from oracle import pyoracle as O
    comps = O.map_build(mean, cov)
    t0 = time.perf_counter()
    O.frames_solve(comps, fr["pose0"][:1], fr["pts_w"][:1], fr["obs"][:1], fr["octave"][:1], nthreads=1)
    t1 = time.perf_counter() - t0
    n = int(min(len(fr["pose0"]), max(cores, target_s * cores / max(t1, 1e-3))))
    n = max(cores, n - n % cores) if n >= cores else n
    n = min(n, len(fr["pose0"]))
    t0 = time.perf_counter()
    O.frames_solve(comps, fr["pose0"][:n], fr["pts_w"][:n], fr["obs"][:n], fr["octave"][:n], nthreads=cores)
    dt = time.perf_counter() - t0
    return n / dt, n, dt


def run_reference(args, rank, world):
    if rank != 0:
        return
    from gmmloc_b200 import synth
    cores = os.cpu_count() or 1
    mean, cov = synth.make_map(K_COMP, 0.5)
    per_step = max(cores, 16)
    fr = synth.make_frames(mean, cov, per_step, N_LM, seed=0)
    from oracle import pyoracle as O
    comps = O.map_build(mean, cov)
    # size each step to ~2 s so the whole run stays within minutes
    t0 = time.perf_counter()
    O.frames_solve(comps, fr["pose0"][:1], fr["pts_w"][:1], fr["obs"][:1], fr["octave"][:1], nthreads=1)
    t1 = time.perf_counter() - t0
    n = int(min(per_step, max(1, 2.0 * cores / t1)))

    def step():
        O.frames_solve(comps, fr["pose0"][:n], fr["pts_w"][:n], fr["obs"][:n], fr["octave"][:n], nthreads=cores)

    for _ in range(args.warmup):
        step()
    t0 = time.perf_counter()
    for _ in range(args.steps):
        step()
    dt = time.perf_counter() - t0
    fps = n * args.steps / dt
    line = {"impl": "reference", "metric": "frames/s (assoc+GN solve)", "value": fps, "unit": "frames/s", "n_gpus": args.gpus,
            "steps": args.steps, "warmup": args.warmup, "ms_per_step": 1e3 * dt / args.steps, "higher_is_better": True,
            "scaling": "weak", "vs_baseline": None, "dtype": "f64", "data": "synthetic",
            "config": config_dict(B_DEFAULT, 1),
            "cpu_baseline": {"value": fps, "unit": "frames/s", "cores": cores, "kind": "port",
                             "sample": f"{n} frames/step x {args.steps} steps of configs[1] (2000 x 10000, 10 GN iters), "
                                       f"{cores} threads sharded over frames"},
            "e2e": {"value": fps, "unit": "frames/s", "h2d_bytes_per_step": 0, "d2h_bytes_per_step": 0}}
    print(json.dumps(line), flush=True)


def main():
    ap = argparse.ArgumentParser()
    ap.add_argument("--gpus", type=int, default=1)
    ap.add_argument("--steps", type=int, default=10)
    ap.add_argument("--warmup", type=int, default=3)
    ap.add_argument("--impl", default="b200", choices=["b200", "reference"])
    ap.add_argument("--frames", type=int, default=B_DEFAULT, help="frames per GPU per step")
    ap.add_argument("--no-cpu-baseline", action="store_true")
    args = ap.parse_args()
    rank = int(os.environ.get("RANK", "0"))
    world = int(os.environ.get("WORLD_SIZE", "1"))
    local = int(os.environ.get("LOCAL_RANK", "0"))
    if args.impl == "reference":
        run_reference(args, rank, world)
        return
    if args.warmup < 3:
        args.warmup = 3

    import torch
    import torch.distributed as dist
    if not torch.cuda.is_available():
        raise SystemExit("bench.py needs a CUDA device: gmmloc_b200 has no CPU fallback (use --impl reference for the CPU arm)")
    torch.cuda.set_device(local)
    if world > 1:
        os.environ.setdefault("MASTER_ADDR", "127.0.0.1")
        dist.init_process_group("nccl", device_id=torch.device("cuda", local))

    import gmmloc_b200 as g
    from gmmloc_b200 import synth

    B = args.frames
    mean, cov = synth.make_map(K_COMP, 0.5)
    fr = synth.make_frames(mean, cov, B, N_LM, seed=0, first_frame=rank * B)
    gmm = g.GMM(mean, cov, device=local)
    ctx = g.Context(local)
    cam, prm = g.v1_camera(), g.default_params()
    assert prm.gn_iters == GN_ITERS
    dev = torch.device("cuda", local)

    # ---- device-resident arm ----
    d_pose0 = torch.from_numpy(fr["pose0"]).to(dev)
    d_pts = torch.from_numpy(fr["pts_w"]).to(dev)
    d_obs = torch.from_numpy(fr["obs"]).to(dev)
    d_oct = torch.from_numpy(fr["octave"]).to(dev)
    d_pose = torch.empty((B, 7), dtype=torch.float64, device=dev)
    d_assign = torch.empty((B, N_LM), dtype=torch.int32, device=dev)
    d_outl = torch.empty((B, N_LM), dtype=torch.uint8, device=dev)
    d_status = torch.empty((B,), dtype=torch.int32, device=dev)
    torch.cuda.synchronize()
    stream = torch.cuda.ExternalStream(ctx.stream, device=dev)

    def step():
        ctx.frames_solve_async(gmm, cam, prm, d_pose0.data_ptr(), d_pts.data_ptr(), d_obs.data_ptr(), d_oct.data_ptr(), B, N_LM,
                               d_pose.data_ptr(), 0, d_assign.data_ptr(), d_outl.data_ptr(), d_status.data_ptr())

    def barrier():
        if world > 1:
            dist.barrier()
        torch.cuda.synchronize()

    sampler = ClockSampler(local) if rank == 0 else None  # nvidia-smi takes ~1 s to start: begin before the warm-up
    for _ in range(args.warmup):
        step()
    ctx.synchronize()
    launches0 = ctx.stats()["kernel_launches"]
    ev0, ev1 = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
    kt = []
    barrier()
    ev0.record(stream)
    for _ in range(args.steps):
        step()
        kt.append(ctx.kernel_times())  # syncs the stream: per-kernel CUDA-event times of this step
    ev1.record(stream)
    barrier()
    clocks = sampler.stop() if sampler else None
    ms = ev0.elapsed_time(ev1)
    launches = ctx.stats()["kernel_launches"] - launches0
    flagged = ctx.stats()["flagged"]
    n_fail = int((d_status != 0).sum().item())
    t = torch.tensor([ms], dtype=torch.float64, device=dev)
    if world > 1:
        dist.all_reduce(t, op=dist.ReduceOp.MAX)
    ms_max = float(t.item())
    value = world * B * args.steps / (ms_max * 1e-3)

    # ---- end-to-end arm: host C ABI with page-locked host buffers ----
    def pinned(a):
        tt = torch.from_numpy(np.ascontiguousarray(a)).pin_memory()
        return tt, tt.numpy()

    keep = []
    h = {}
    for k in ("pose0", "pts_w", "obs", "octave"):
        tt, h[k] = pinned(fr[k]); keep.append(tt)
    o_pose_t = torch.empty((B, 7), dtype=torch.float64).pin_memory(); o_pose = o_pose_t.numpy()
    o_assign_t = torch.empty((B, N_LM), dtype=torch.int32).pin_memory(); o_assign = o_assign_t.numpy()
    o_outl_t = torch.empty((B, N_LM), dtype=torch.uint8).pin_memory(); o_outl = o_outl_t.numpy()
    o_status_t = torch.empty((B,), dtype=torch.int32).pin_memory(); o_status = o_status_t.numpy()

    def step_e2e():
        ctx.frames_solve_into(gmm, cam, prm, h["pose0"], h["pts_w"], h["obs"], h["octave"], o_pose, o_assign, o_status,
                              pts_out=None, outlier=o_outl)

    for _ in range(2):
        step_e2e()
    barrier()
    t0 = time.perf_counter()
    for _ in range(args.steps):
        step_e2e()
    torch.cuda.synchronize()
    e2e_s = time.perf_counter() - t0
    t = torch.tensor([e2e_s], dtype=torch.float64, device=dev)
    if world > 1:
        dist.all_reduce(t, op=dist.ReduceOp.MAX)
    e2e_value = world * B * args.steps / float(t.item())
    h2d = B * (56 + N_LM * (12 + 12 + 1))
    d2h = B * (56 + 4 + N_LM * (4 + 1))
    same = bool(np.array_equal(o_assign, d_assign.cpu().numpy()) and np.array_equal(o_pose, d_pose.cpu().numpy()))

    if rank == 0:
        hbm_peak, which = measured_peaks()
        scan_ms = statistics.mean(k["scan_ms"] for k in kt)
        rerank_ms = statistics.mean(k["rerank_ms"] for k in kt)
        gn_ms = statistics.mean(k["gn_ms"] for k in kt)
        bytes_launch = algorithmic_bytes_per_frame(N_LM, K_COMP) * B
        flops_launch = algorithmic_flops_per_frame(N_LM, K_COMP) * B
        ach_gbs = bytes_launch / (scan_ms * 1e-3) / 1e9
        ach_tf = flops_launch / (scan_ms * 1e-3) / 1e12
        line = {"metric": "frames/s (assoc+GN solve)", "value": value, "unit": "frames/s", "n_gpus": world, "steps": args.steps,
                "warmup": args.warmup, "ms_per_step": ms_max / args.steps, "higher_is_better": True, "scaling": "weak",
                "vs_baseline": None, "dtype": "f32", "data": "synthetic", "config": config_dict(B, world),
                "e2e": {"value": e2e_value, "unit": "frames/s", "h2d_bytes_per_step": h2d, "d2h_bytes_per_step": d2h,
                        "matches_device_arm": same},
                "gpu_launches": int(launches),
                "clocks": clocks,
                "roofline": {"bound": "hbm", "kernel": "assoc_scan_kernel", "achieved": ach_gbs, "peak": hbm_peak, "unit": "GB/s",
                             "frac": ach_gbs / hbm_peak, "traffic": None, "peak_source": f"{which} (MEASURED_PEAKS.json hbm_gbs)",
                             "algorithmic_bytes_per_launch": bytes_launch, "avg_launch_ms": scan_ms,
                             "note": "FP32-issue-bound kernel (≈1e3 flop/B); see roofline_fp32"},
                "roofline_fp32": {"bound": "fp32_fma", "kernel": "assoc_scan_kernel", "achieved": ach_tf, "peak": FFMA_PEAK_TFLOPS,
                                  "unit": "TFLOP/s", "frac": ach_tf / FFMA_PEAK_TFLOPS,
                                  "algorithmic_flops_per_launch": flops_launch,
                                  "peak_source": "measured FFMA micro-benchmark (tools/ubench_pipes.cu)"},
                "kernel_ms": {"assoc_scan": scan_ms, "assoc_rerank": rerank_ms, "frame_gn": gn_ms},
                "rerank_flagged_frac": flagged / float(B * N_LM), "frames_numeric_failure": n_fail}
        if world == 1 and not args.no_cpu_baseline:
            cores = os.cpu_count() or 1
            fps, n, dt = cpu_sample(mean, cov, fr, cores)
            line["cpu_baseline"] = {"value": fps, "unit": "frames/s", "cores": cores, "kind": "port",
                                    "sample": f"{n} frames of the same workload in {dt:.1f} s, {cores} threads sharded over frames"}
        print(json.dumps(line), flush=True)
    if world > 1:
        dist.destroy_process_group()


if __name__ == "__main__":
    main()
