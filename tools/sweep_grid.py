"""GPU sweep of the association grid's options (gmmloc_grid_options_t: level thresholds T, cell size) on a BASELINE workload:
time of the association stage through gmmloc_associate (host call, kernel times from the context) and identity of the
assignments with the default grid.   python tools/sweep_grid.py [c2|c3|c5] [frames]"""
import sys
from pathlib import Path

import numpy as np

ROOT = Path(__file__).resolve().parent.parent
sys.path.insert(0, str(ROOT))
import gmmloc_b200 as g  # noqa: E402
from gmmloc_b200 import synth  # noqa: E402

wl = sys.argv[1] if len(sys.argv) > 1 else "c2"
B = int(sys.argv[2]) if len(sys.argv) > 2 else 1024
if wl == "c2":
    mean, cov = synth.make_map(10000, 0.5)
    fr = synth.make_frames(mean, cov, 128, 2000, seed=0)
elif wl == "c5":
    mean, cov = synth.make_map(50000, 0.95)
    fr = synth.make_frames(mean, cov, 32, 5000, seed=0, mono_frac=0.3)
else:
    z = np.load(ROOT / "tests" / "golden" / "v1_gmm.npz")
    gt = np.load(ROOT / "tests" / "golden" / "v1_03_gt_sync.npz")["gt"]
    mean, cov = z["mean"], z["cov"]
    fr = synth.make_frames(mean, cov, 256, 1200, seed=0, poses_wc=gt[np.arange(256) * 8, 1:8])
pts = np.concatenate([fr["pts_w"]] * max(1, B // len(fr["pts_w"])))
ctx = g.Context(0)
ref = None
cfgs = [(0.0, (0, 0, 0))]
for t0 in (9.0, 16.0, 25.0, 36.0):
    for cs in (0.0,):
        cfgs.append((cs, (t0, 144.0, 400.0)))
cfgs += [(0.0, (16.0, 64.0, 400.0)), (0.0, (16.0, 36.0, 400.0)), (0.0, (25.0, 100.0, 400.0))]
for cs, tm in cfgs:
    opt = g.capi.GridOptions(enable=1, levels=3 if tm[0] else 0, cell_size=cs)
    for i in range(3):
        opt.tmax[i] = tm[i]
    m = g.GMM(mean, cov, grid=opt)
    for _ in range(2):
        a, _d = ctx.associate(m, pts, want_d2=False)
    t = []
    for _ in range(5):
        a, _d = ctx.associate(m, pts, want_d2=False)
        t.append(ctx.kernel_times())
    if ref is None:
        ref = a
    print(f"cell {cs:4.2f} T {tm}: scan {np.median([x['scan_ms'] for x in t]):.3f} ms rerank {np.median([x['rerank_ms'] for x in t]):.3f} ms "
          f"identical {bool(np.array_equal(a, ref))}", flush=True)
