"""One 12-keyframe / 1 500-point window through gmmloc_local_ba on a cluster of 8 CTAs (for ncu)."""
import sys
from pathlib import Path
ROOT = Path(__file__).resolve().parent.parent
sys.path.insert(0, str(ROOT))
import gmmloc_b200 as g
from gmmloc_b200 import synth
mean, cov = synth.make_map(3000, 0.7, seed=5)
m = g.GMM(mean, cov)
ctx = g.Context(0)
pbs = [synth.make_ba_problem(mean, cov, 12, 3, 1500, seed=100)]
ctx.set_option(g.capi.OPT_BA_CLUSTER, int(sys.argv[1]) if len(sys.argv) > 1 else 8)
for _ in range(2):
    ctx.jointOptimization(m, pbs)
