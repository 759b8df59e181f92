"""Bring-up check of the tensor-core pruning scan: assignments vs the FFMA2 scan and the oracle, timing."""
import sys
import time
from pathlib import Path

import numpy as np

ROOT = Path(__file__).resolve().parent.parent
sys.path.insert(0, str(ROOT))
import gmmloc_b200 as g  # noqa: E402
from gmmloc_b200 import capi, synth  # noqa: E402

K = int(sys.argv[1]) if len(sys.argv) > 1 else 10000
M = int(sys.argv[2]) if len(sys.argv) > 2 else 400000
mean, cov = synth.make_map(K, 0.5)
m = g.GMM(mean, cov)
ctx = g.Context(0)
pts, _ = synth.make_points_near(mean, cov, M, seed=3, spread=3.0)
ctx.set_assoc_kernel(capi.ASSOC_FFMA2)
a0, d0 = ctx.associate(m, pts)
f0 = ctx.stats()["flagged"]
ctx.set_assoc_kernel(capi.ASSOC_TENSOR)
t0 = time.perf_counter()
a1, d1 = ctx.associate(m, pts)
bnd = m.scan_bound(); print("alpha, xmax", bnd); print("tensor kernel ran:", ctx.last_assoc_kernel(), "first call s", time.perf_counter() - t0)
f1 = ctx.stats()["flagged"]
print("K", K, "M", M, "equal assign:", np.array_equal(a0, a1), "mismatches", int((a0 != a1).sum()), "d2 equal", np.array_equal(d0, d1),
      "flagged ffma2", f0, "tensor", f1)
if not np.array_equal(a0, a1):
    bad = np.flatnonzero(a0 != a1)[:10]
    print(bad, a0[0, bad] if a0.ndim == 2 else a0[bad], a1[0, bad] if a1.ndim == 2 else a1[bad])
for kind, name in ((capi.ASSOC_FFMA2, "ffma2"), (capi.ASSOC_TENSOR, "tensor")):
    ctx.set_assoc_kernel(kind)
    ctx.associate(m, pts)
    t = ctx.kernel_times()
    print(name, t)
