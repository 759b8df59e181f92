"""How many assignments change when the 3-term sums of d2 = delta^T cov_inv delta (gaussian.cpp:65-70) are associated
a + (b + c) instead of the oracle's (a + b) + c?  Eigen's unrolled 3-element reduction order cannot be verified in this
box (no Eigen), so the sensitivity of the INTEGER output to that choice is measured instead.

    python tools/chi2_order_flips.py            # shipped v1.gmm (C3 landmarks) and v2.gmm, prints a JSON line per map
"""
import json
import sys
from pathlib import Path

import numpy as np

ROOT = Path(__file__).resolve().parent.parent
sys.path.insert(0, str(ROOT))
from gmmloc_b200 import synth  # noqa: E402
from oracle import pyoracle as O  # noqa: E402


def d2_both_orders(x, mu, A):
    """x n x 3, mu K x 3, A K x 9 -> (d2 with (a+b)+c, d2 with a+(b+c)), each n x K, fp64, no fused operations."""
    d = x[:, None, :] - mu[None, :, :]
    d0, d1, d2 = d[..., 0], d[..., 1], d[..., 2]
    t = [[d0 * A[None, :, c], d1 * A[None, :, 3 + c], d2 * A[None, :, 6 + c]] for c in range(3)]
    rL = [(t[c][0] + t[c][1]) + t[c][2] for c in range(3)]
    rR = [t[c][0] + (t[c][1] + t[c][2]) for c in range(3)]
    L = (rL[0] * d0 + rL[1] * d1) + rL[2] * d2
    R = rR[0] * d0 + (rR[1] * d1 + rR[2] * d2)
    return L, R


def run(name, pts):
    z = np.load(ROOT / "tests" / "golden" / f"{name}_gmm.npz")
    comps = O.map_build(z["mean"], z["cov"])
    mu, A = np.ascontiguousarray(comps["mean"]), np.ascontiguousarray(comps["cov_inv"])
    flips = ties = 0
    worst_gap = np.inf
    for i in range(0, len(pts), 1024):
        x = pts[i:i + 1024].astype(np.float64)
        L, R = d2_both_orders(x, mu, A)
        aL, aR = L.argmin(axis=1), R.argmin(axis=1)  # first minimum = lowest index on ties, as the reference's '<' scan
        flips += int((aL != aR).sum())
        part = np.partition(L, 1, axis=1)
        ties += int((part[:, 0] == part[:, 1]).sum())
        gap = (part[:, 1] - part[:, 0]) / np.maximum(part[:, 1], 1e-300)
        worst_gap = min(worst_gap, float(gap.min()))
    # the oracle's own order must be the (a + b) + c column
    chk = O.associate(comps, pts[:2000].astype(np.float32))[0].ravel()
    L, _ = d2_both_orders(pts[:2000].astype(np.float32).astype(np.float64), mu, A)
    assert np.array_equal(chk, L.argmin(axis=1))
    return {"map": name, "landmarks": int(len(pts)), "components": int(len(mu)), "assignment_flips": flips,
            "exact_ties_between_best_and_second": ties, "smallest_relative_gap_best_to_second": worst_gap}


if __name__ == "__main__":
    gt = np.load(ROOT / "tests" / "golden" / "v1_03_gt_sync.npz")["gt"]
    z = np.load(ROOT / "tests" / "golden" / "v1_gmm.npz")
    fr = synth.make_frames(z["mean"], z["cov"], len(gt[::8]), 1200, seed=0, poses_wc=gt[::8, 1:8])
    print(json.dumps(run("v1", fr["pts_w"].reshape(-1, 3))))
    z2 = np.load(ROOT / "tests" / "golden" / "v2_gmm.npz")
    pts, _ = synth.make_points_near(z2["mean"], z2["cov"], 300000, seed=5, spread=1.5)
    print(json.dumps(run("v2", pts)))
