"""debug helper (GPU box): which bench frames report GMMLOC_FRAME_NUMERIC and what the oracle says about them."""
import sys
from pathlib import Path
import numpy as np
sys.path.insert(0, str(Path(__file__).resolve().parent.parent))
import gmmloc_b200 as g
from gmmloc_b200 import synth
from oracle import pyoracle as O

B = int(sys.argv[1]) if len(sys.argv) > 1 else 1024
mean, cov = synth.make_map(10000, 0.5)
fr = synth.make_frames(mean, cov, B, 2000)
m = g.GMM(mean, cov)
ctx = g.Context(0)
out = ctx.frames_solve(m, fr["pose0"], fr["pts_w"], fr["obs"], fr["octave"])
bad = np.flatnonzero(out["status"])
print("gpu failed frames", bad)
comps = O.map_build(mean, cov)
for f in bad[:6]:
    r = O.frame_solve(comps, fr["pose0"][f], fr["pts_w"][f], fr["obs"][f], fr["octave"][f], want_trace=True)
    print(f, "oracle status", r["status"], "trace", np.abs(r["trace"][:, :6]).max(axis=1))
    print("   assign equal", np.array_equal(r["assign"], out["assign"][f]), "pose moved", np.abs(r["pose"] - fr["pose0"][f]).max())
    for it in (1, 2, 5):
        p = g.default_params(gn_iters=it)
        o = ctx.frames_solve(m, fr["pose0"][f:f + 1], fr["pts_w"][f:f + 1], fr["obs"][f:f + 1], fr["octave"][f:f + 1], params=p)
        po = O.default_params(gn_iters=it)
        ro = O.frame_solve(comps, fr["pose0"][f], fr["pts_w"][f], fr["obs"][f], fr["octave"][f], prm=po)
        print("   iters", it, "gpu status", o["status"], "dpose vs oracle", np.abs(o["pose"][0] - ro["pose"]).max(),
              "dpts", np.abs(o["pts"][0] - ro["pts"]).max())
