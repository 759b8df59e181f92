"""Times K9 (local bundle adjustment) against the CPU oracle on synthetic covisibility windows.

    python tools/bench_ba.py            # on the GPU box
"""
import json
import sys
import time
from pathlib import Path

import numpy as np

ROOT = Path(__file__).resolve().parent.parent
sys.path.insert(0, str(ROOT))
import gmmloc_b200 as g  # noqa: E402
from gmmloc_b200 import synth  # noqa: E402
from oracle import pyoracle as O  # noqa: E402

mean, cov = synth.make_map(3000, 0.7, seed=5)
m = g.GMM(mean, cov)
comps = O.map_build(mean, cov)
ctx = g.Context(0)
res = []
CASES = [(8, 2, 600, 1, 0), (12, 3, 1500, 1, 1), (12, 3, 1500, 1, 2), (12, 3, 1500, 1, 4), (12, 3, 1500, 1, 8), (20, 4, 4000, 1, 1),
         (20, 4, 4000, 1, 8), (12, 3, 1500, 18, 0), (12, 3, 1500, 148, 0), (12, 3, 1500, 592, 0)]
for (nk, nfix, npt, B, cluster) in CASES:  # cluster: CTAs per window (0 = the library's automatic choice)
    pbs = [synth.make_ba_problem(mean, cov, nk, nfix, npt, seed=100 + j) for j in range(min(B, 8))]
    pbs = [pbs[j % len(pbs)] for j in range(B)]
    ctx.set_option(g.capi.OPT_BA_CLUSTER, cluster)
    ctx.jointOptimization(m, pbs)
    ts = []
    for _ in range(3):
        t0 = time.perf_counter()
        out = ctx.jointOptimization(m, pbs)
        ts.append(time.perf_counter() - t0)
    t_gpu = min(ts)
    t0 = time.perf_counter()
    want = O.local_ba(comps, *(pbs[0][k] for k in ("pose", "fixed", "pts", "assoc", "obs_pt", "obs_kf", "obs", "octave")))
    t_cpu = time.perf_counter() - t0
    res.append(dict(keyframes=nk, fixed=nfix, points=npt, observations=int(len(pbs[0]["obs_pt"])), windows=B, cluster=cluster,
                    gpu_ms_per_call=1e3 * t_gpu, gpu_windows_per_s=B / t_gpu, cpu_ms_per_window_1core=1e3 * t_cpu,
                    iters=[int(x) for x in out[0]["iters"]], dpose_max=float(np.abs(out[0]["pose"] - want["pose"]).max())))
    print(json.dumps(res[-1]), flush=True)
