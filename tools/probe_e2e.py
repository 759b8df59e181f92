"""GPU probe: PCIe copy bandwidth with page-locked buffers and the e2e frames/s of gmmloc_frames_solve for several
pipeline chunk sizes (gmmloc_ctx_set_option GMMLOC_OPT_E2E_CHUNK; each chunk size runs in a child process)."""
import os
import subprocess
import sys
import time
from pathlib import Path

ROOT = Path(__file__).resolve().parent.parent
sys.path.insert(0, str(ROOT))


def child(chunk):
    import numpy as np
    import torch
    import gmmloc_b200 as g
    from gmmloc_b200 import synth
    B, N, K = 4096, 2000, 10000
    mean, cov = synth.make_map(K, 0.5)
    fr = synth.make_frames(mean, cov, 256, N, seed=0)
    fr = {k: np.concatenate([v] * (B // 256)) for k, v in fr.items() if k in ("pose0", "pts_w", "obs", "octave")}
    gmm, ctx = g.GMM(mean, cov), g.Context(0)
    ctx.set_option(g.capi.OPT_E2E_CHUNK, chunk)
    cam, prm = g.v1_camera(), g.default_params()
    keep, h = [], {}
    for k in ("pose0", "pts_w", "obs", "octave"):
        t = torch.from_numpy(np.ascontiguousarray(fr[k])).pin_memory()
        keep.append(t)
        h[k] = t.numpy()
    o_pose = torch.empty((B, 7), dtype=torch.float64).pin_memory()
    o_assign = torch.empty((B, N), dtype=torch.int32).pin_memory()
    o_outl = torch.empty((B, N), dtype=torch.uint8).pin_memory()
    o_status = torch.empty((B,), dtype=torch.int32).pin_memory()

    def step():
        ctx.frames_solve_into(gmm, cam, prm, h["pose0"], h["pts_w"], h["obs"], h["octave"], o_pose.numpy(), o_assign.numpy(),
                              o_status.numpy(), pts_out=None, outlier=o_outl.numpy())
    for _ in range(3):
        step()
    t0 = time.perf_counter()
    for _ in range(8):
        step()
    dt = (time.perf_counter() - t0) / 8
    print(f"chunk={chunk:5d}  e2e {1e3 * dt:7.3f} ms/step  {B / dt:10.0f} frames/s", flush=True)


def main():
    if len(sys.argv) > 1:
        child(int(sys.argv[1]))
        return
    import torch
    n = 205 << 20
    hbuf = torch.empty(n, dtype=torch.uint8).pin_memory()
    d = torch.empty(n, dtype=torch.uint8, device="cuda")
    for name, fn in (("H2D", lambda: d.copy_(hbuf, non_blocking=True)), ("D2H", lambda: hbuf.copy_(d, non_blocking=True))):
        fn()
        torch.cuda.synchronize()
        a, b = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
        a.record()
        for _ in range(5):
            fn()
        b.record()
        torch.cuda.synchronize()
        print(f"{name} pinned 205 MiB: {5 * n / (a.elapsed_time(b) * 1e-3) / 1e9:.1f} GB/s", flush=True)
    for chunk in [int(x) for x in os.environ.get("PROBE_CHUNKS", "256,444,512,888,1024,1332,2048,4096").split(",")]:
        subprocess.run([sys.executable, __file__, str(chunk)], check=False)


if __name__ == "__main__":
    main()
