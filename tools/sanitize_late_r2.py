"""GPU-box helper for `compute-sanitizer --tool racecheck|memcheck python tools/sanitize_late_r2.py`: a short tour of the kernels
changed late in round 2 — the sorted grid scan with its warp-cooperative rounds and level loop (dense cells, sparse cells,
out-of-grid / non-finite landmarks, landmarks that need the higher levels), the pack kernel (one warp per class, MATCH) and the
streamed solver's vector shared-memory reads."""
import sys
from pathlib import Path

sys.path.insert(0, str(Path(__file__).resolve().parent.parent))
import numpy as np
import gmmloc_b200 as g
from gmmloc_b200 import synth

mean, cov = synth.make_map(1500, 0.5)
m = g.GMM(mean, cov, device=0)
ctx = g.Context(0)
ctx.set_assoc_kernel(g.capi.ASSOC_GRID)
for n, spread in ((40000, 1.0), (6000, 4.0)):  # many landmarks per cell / few per cell and far from every component
    pts, _ = synth.make_points_near(mean, cov, n, seed=3, spread=spread)
    pts[::1000] += 50
    pts[5::5000] = np.nan
    for want in (True, False):  # without d2 the conservative fp32 certificate decides first
        r = ctx.associate(m, pts, want_d2=want)
        print(n, want, ctx.last_assoc_kernel(), r[0][:5], flush=True)
for B, N in ((6, 601), (2, 1999)):
    fr = synth.make_frames(mean, cov, B, N, seed=B + N, mono_frac=0.2)
    out = ctx.frames_solve(m, fr["pose0"], fr["pts_w"], fr["obs"], fr["octave"], params=g.default_params(robust=1))
    print(B, N, ctx.last_gn_kernel(), out["status"][:3], out["assign"][0, :5], flush=True)
print("sanitize script done")
