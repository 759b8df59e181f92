"""GPU-box helper: a small tour of the hot path for `compute-sanitizer --tool memcheck python tools/memcheck_smoke.py`
(or `--tool racecheck` with MEMCHECK_LIGHT=1; every launch variant of the mixed frame GN with and without the preparation kernel, odd frame sizes, both association
kernels with out-of-grid / non-finite landmarks, the pipelined host path)."""
import os
import sys
from pathlib import Path

sys.path.insert(0, str(Path(__file__).resolve().parent.parent))
import numpy as np
import gmmloc_b200 as g
from gmmloc_b200 import synth

mean, cov = synth.make_map(1500, 0.5)
m = g.GMM(mean, cov, device=0)
ctx = g.Context(0)
# the default kernel (K3c + K3q + K3s: TMA bulk copies, mbarrier rings) on both launch variants, chunk-size edges, a frame with a
# non-finite landmark (fails in K3s -> fp64 retry launch), the world-frame kernel, then the camera-frame kernels below
for kind in (g.capi.GN_STREAM, g.capi.GN_WORLD):
    ctx.set_gn_precision(kind)
    for B, N in ((3, 33), (2, 1999), (2, 2137), (1, 4289), (40, 601)):
        fr = synth.make_frames(mean, cov, B, N, seed=B + N, mono_frac=0.2)
        fr["pts_w"][0, 7] = np.nan
        out = ctx.frames_solve(m, fr["pose0"], fr["pts_w"], fr["obs"], fr["octave"], params=g.default_params(robust=1))
        print(kind, B, N, ctx.last_gn_kernel(), out["status"][:3], out["assign"][0, :5], flush=True)
ctx.set_gn_precision(g.capi.GN_MIXED)
for prep in (1, 0):
    ctx.set_option(g.capi.OPT_GN_PREP, prep)
    for B, N in ((3, 501), (2, 1999), (2, 2451), (2, 5001), (40, 1201)):
        fr = synth.make_frames(mean, cov, B, N, seed=B, mono_frac=0.2)
        out = ctx.frames_solve(m, fr["pose0"], fr["pts_w"], fr["obs"], fr["octave"], params=g.default_params(robust=1))
        print(prep, B, N, out["status"][:3], out["assign"][0, :5], flush=True)
ctx.set_option(g.capi.OPT_GN_PREP, -1)
ctx.set_gn_precision(g.capi.GN_DEFAULT)
LIGHT = os.environ.get("MEMCHECK_LIGHT") == "1"  # racecheck / initcheck runs: smaller workloads
pts, _ = synth.make_points_near(mean, cov, 20000 if LIGHT else 300000, seed=3, spread=2.0)
pts[::1000] += 50
pts[5::5000] = np.nan
for kind in (g.capi.ASSOC_GRID, g.capi.ASSOC_FFMA2):
    ctx.set_assoc_kernel(kind)
    a, d = ctx.associate(m, pts)
    print(kind, ctx.last_assoc_kernel(), a[:5], flush=True)
ctx.set_assoc_kernel(g.capi.ASSOC_GRID)
fr = synth.make_frames(mean, cov, 300 if LIGHT else 1100, 61, seed=9)
out = ctx.frames_solve(m, fr["pose0"], fr["pts_w"], fr["obs"], fr["octave"])
print(out["status"].sum())
print("memcheck script done")
