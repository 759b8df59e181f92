"""GPU-box diagnostic: where does the frame GN deviate from the oracle, and does non-convergence explain it?

For BASELINE configs[1] (4 096 frames) and configs[4] (256 frames): solve with the mixed-precision kernel and with the fp64
kernel, replay every frame on the oracle (with the GPU's assignments: they are identical to the oracle's) recording
the per-iteration pose steps, and tabulate the pose deviation of both kernels against the oracle next to the
oracle's own convergence (last step, non-monotone steps).  Writes gpurun_out/diag_conv_<wl>.npz + a text summary."""
import sys
from concurrent.futures import ThreadPoolExecutor
from pathlib import Path

import numpy as np

ROOT = Path(__file__).resolve().parent.parent
sys.path.insert(0, str(ROOT))
import gmmloc_b200 as g  # noqa: E402
from gmmloc_b200 import synth  # noqa: E402
from oracle import pyoracle as O  # noqa: E402
from bench import pose_errors  # noqa: E402


def run(name, K, pdeg, N, mono, robust, B):
    mean, cov = synth.make_map(K, pdeg)
    fr = synth.make_frames(mean, cov, B, N, seed=0, mono_frac=mono)
    m, ctx = g.GMM(mean, cov), g.Context(0)
    prm = g.default_params(robust=robust)
    outs = {}
    for nm, prec in (("mixed", g.capi.GN_MIXED), ("fp64", g.capi.GN_FP64)):
        ctx.set_gn_precision(prec)
        outs[nm] = ctx.frames_solve(m, fr["pose0"], fr["pts_w"], fr["obs"], fr["octave"], params=prm)
    comps = O.map_build(mean, cov)
    oprm = O.default_params(robust=robust)

    def one(f):
        r = O.frame_solve(comps, fr["pose0"][f], fr["pts_w"][f], fr["obs"][f], fr["octave"][f], prm=oprm,
                          assign_in=outs["mixed"]["assign"][f], want_trace=True)
        return r["pose"], r["status"], np.abs(r["trace"][:, :6]).max(axis=1), r["trace"][:, 6], r["pts"], r["outlier"]

    with ThreadPoolExecutor(16) as ex:
        res = list(ex.map(one, range(B)))
    pose = np.array([r[0] for r in res]); status = np.array([r[1] for r in res])
    steps = np.array([r[2] for r in res]); ptsteps = np.array([r[3] for r in res])
    pts = np.array([r[4] for r in res]); outl = np.array([r[5] for r in res])
    dev = {}
    for nm in outs:
        rad = np.array([pose_errors(outs[nm]["pose"][f:f + 1], pose[f:f + 1])[0] for f in range(B)])
        met = np.abs(outs[nm]["pose"][:, 4:] - pose[:, 4:]).max(axis=1)
        dpt = np.abs(outs[nm]["pts"].astype(np.float64) - pts).reshape(B, -1).max(axis=1)
        dev[nm] = (rad, met, dpt)
        bad = (rad > 1e-5) | (met > 1e-4)
        print(f"{name} {nm}: status mismatch {(outs[nm]['status'] != status).sum()}, outlier mismatch "
              f"{(outs[nm]['outlier'] != outl).sum()}, frames beyond 1e-5 rad / 1e-4 m: {bad.sum()} {np.flatnonzero(bad)[:20]}, "
              f"max rad {rad.max():.2e} max m {met.max():.2e}; frames with a point beyond 1e-4 m: {(dpt > 1e-4).sum()}, max {dpt.max():.2e}")
    grow = np.zeros(B, bool)
    for i in range(2, steps.shape[1]):
        grow |= (steps[:, i] > steps[:, i - 1]) & (steps[:, i] > 1e-7)
    last = steps[:, -1]
    for nm in outs:
        rad, met, dpt = dev[nm]
        bad = (rad > 1e-5) | (met > 1e-4)
        print(f"{name} {nm}: bad&grow {(bad & grow).sum()} bad&~grow {(bad & ~grow).sum()} grow {grow.sum()} | bad&last>1e-6 {(bad & (last > 1e-6)).sum()} "
              f"bad&last<=1e-6 {(bad & (last <= 1e-6)).sum()} last>1e-6 {(last > 1e-6).sum()} | flag=grow&last>1e-6: {(grow & (last > 1e-6)).sum()} "
              f"bad outside flag {(bad & ~(grow & (last > 1e-6))).sum()}")
        order = np.argsort(-rad)[:6]
        for f in order:
            print(f"   frame {f}: {rad[f]:.2e} rad {met[f]:.2e} m pts {dpt[f]:.2e} | steps " + " ".join(f"{v:.1e}" for v in steps[f]) +
                  f" | max pt step last {ptsteps[f, -1]:.1e}")
    np.savez_compressed(ROOT / "gpurun_out" / f"diag_conv_{name}.npz", steps=steps, ptsteps=ptsteps, status=status,
                        **{f"{nm}_{k}": v for nm in dev for k, v in zip(("rad", "m", "pts"), dev[nm])})


if __name__ == "__main__":
    (ROOT / "gpurun_out").mkdir(exist_ok=True)
    run("c2", 10000, 0.5, 2000, 0.0, 0, 4096)
    run("c5", 50000, 0.95, 5000, 0.3, 1, 256)
