"""Which frames of the C2 bench batch end with GMMLOC_FRAME_NUMERIC under each frame-GN kernel, and what the oracle says
about exactly those frames (diagnostic; run on the GPU box)."""
import sys, json
from pathlib import Path
import numpy as np
ROOT = Path(__file__).resolve().parent.parent
sys.path.insert(0, str(ROOT))
import gmmloc_b200 as g
from gmmloc_b200 import synth
from oracle import pyoracle as O

B = int(sys.argv[1]) if len(sys.argv) > 1 else 4096
mean, cov = synth.make_map(10000, 0.5)
fr = synth.make_frames(mean, cov, B, 2000, seed=0)
m = g.GMM(mean, cov)
ctx = g.Context(0)
outs = {}
for name, kind in (("fp64", g.capi.GN_FP64), ("mixed", g.capi.GN_MIXED), ("world", g.capi.GN_WORLD), ("stream", g.capi.GN_STREAM)):
    ctx.set_gn_precision(kind)
    outs[name] = ctx.frames_solve(m, fr["pose0"], fr["pts_w"], fr["obs"], fr["octave"])
    print(name, "failed frames:", np.flatnonzero(outs[name]["status"]).tolist())
bad = sorted(set(np.concatenate([np.flatnonzero(o["status"]) for o in outs.values()]).tolist()))
if bad:
    comps = O.map_build(mean, cov)
    idx = np.array(bad)
    exp = O.frames_solve(comps, fr["pose0"][idx], fr["pts_w"][idx], fr["obs"][idx], fr["octave"][idx], nthreads=8, want_pts=True, want_conv=True)
    for i, f in enumerate(bad):
        print("frame", f, "oracle status", int(exp["status"][i]), "conv", exp["conv"][i].tolist(), "max pt_last", float(np.nanmax(exp["pt_last"][i])),
              {k: int(o["status"][f]) for k, o in outs.items()})
