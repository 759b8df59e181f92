import sys; sys.path.insert(0,'/root/repo')
import numpy as np
import gmmloc_b200 as g
from gmmloc_b200 import synth
mean, cov = synth.make_map(1500, 0.5)
m = g.GMM(mean, cov, device=0)
ctx = g.Context(0)
for B,N in ((3,500),(2,2000),(40,1200)):
    fr = synth.make_frames(mean, cov, B, N, seed=B, mono_frac=0.2)
    out = ctx.frames_solve(m, fr["pose0"], fr["pts_w"], fr["obs"], fr["octave"])
    print(B,N,out["status"][:3], out["assign"][0,:5])
pts,_ = synth.make_points_near(mean, cov, 300000, seed=3, spread=2.0)
pts[::1000]+=50; pts[5::5000]=np.nan
for kind in (g.capi.ASSOC_GRID, g.capi.ASSOC_FFMA2):
    ctx.set_assoc_kernel(kind); a,d = ctx.associate(m, pts); print(kind, ctx.last_assoc_kernel(), a[:5])
fr = synth.make_frames(mean, cov, 1100, 60, seed=9)
out = ctx.frames_solve(m, fr["pose0"], fr["pts_w"], fr["obs"], fr["octave"]); print(out["status"].sum())
print("memcheck script done")
