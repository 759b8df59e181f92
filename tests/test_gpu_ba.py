"""GPU parity of K9 (local bundle adjustment, Localization::jointOptimization) through the C ABI against the
CPU oracle.  Bar: poses within 1e-5 rad / 1e-4 m (BASELINE.json), outlier / cleared-association flags and LM
iteration counts identical."""
import numpy as np
import pytest

from gmmloc_b200 import synth
from tests import np_reference as ref

pytestmark = pytest.mark.gpu
TOL_RAD, TOL_M = 1e-5, 1e-4


def _pose_err(pa, pb):
    Ra, ta = ref.pose_Rt(pa)
    Rb, tb = ref.pose_Rt(pb)
    return ref.rot_angle(Ra, Rb), float(np.abs(ta - tb).max())


@pytest.fixture(scope="module")
def ba_map(gpu, oracle):
    g, _ = gpu
    mean, cov = synth.make_map(3000, 0.7, seed=5)
    return g.GMM(mean, cov), oracle.map_build(mean, cov), mean, cov


def _check(out, want, pb, tol_pts=1e-4):
    # stages 1 / 2 always run their 5 iterations; stage 3 ends when LM stalls at convergence (rho == 0 or ten
    # rejected trials), which depends on the last bits of the chi2 sums: compare the estimates, not that count
    assert np.array_equal(out["iters"][:2], want["iters"][:2]), (out["iters"], want["iters"])
    assert 1 <= out["iters"][2] <= 40
    assert np.array_equal(out["obs_outlier"], want["obs_outlier"])
    assert np.array_equal(out["assoc_cleared"], want["assoc_cleared"])
    for i in range(len(pb["pose"])):
        ang, dt = _pose_err(out["pose"][i], want["pose"][i])
        assert ang < TOL_RAD and dt < TOL_M, (i, ang, dt)
    assert np.abs(out["pts"] - want["pts"]).max() < tol_pts


@pytest.mark.parametrize("shape", [(4, 1, 80, -1), (8, 2, 600, -1), (5, 2, 150, 0), (12, 3, 1500, -1)])
def test_local_ba_matches_oracle(gpu, oracle, ba_map, shape):
    g, ctx = gpu
    m, comps, mean, cov = ba_map
    nk, nfix, npt, prior = shape
    pb = synth.make_ba_problem(mean, cov, nk, nfix, npt, seed=nk)
    pb["prior_idx"] = prior
    want = oracle.local_ba(comps, pb["pose"], pb["fixed"], pb["pts"], pb["assoc"], pb["obs_pt"], pb["obs_kf"], pb["obs"],
                           pb["octave"], prior_idx=prior)
    out = ctx.jointOptimization(m, [pb])[0]
    assert out["status"] == 0
    _check(out, want, pb)
    # fixed poses come back normalised but otherwise untouched
    for i in np.flatnonzero(pb["fixed"]):
        assert _pose_err(out["pose"][i], pb["pose"][i])[1] < 1e-15
    # the window moved towards the ground truth
    free = np.flatnonzero(pb["fixed"] == 0)
    e0 = np.abs(pb["pose"][free, 4:] - pb["pose_gt"][free, 4:]).max()
    e1 = np.abs(out["pose"][free, 4:] - pb["pose_gt"][free, 4:]).max()
    assert e1 < e0


def test_local_ba_batch_of_windows(gpu, oracle, ba_map):
    g, ctx = gpu
    m, comps, mean, cov = ba_map
    shapes = [(3, 1, 60), (6, 2, 300), (2, 2, 50), (5, 0, 200)]  # the third has no free pose: points only
    pbs = [synth.make_ba_problem(mean, cov, nk, nf, npt, seed=40 + j) for j, (nk, nf, npt) in enumerate(shapes)]
    outs = ctx.jointOptimization(m, pbs)
    for pb, out in zip(pbs, outs):
        want = oracle.local_ba(comps, pb["pose"], pb["fixed"], pb["pts"], pb["assoc"], pb["obs_pt"], pb["obs_kf"], pb["obs"],
                               pb["octave"])
        _check(out, want, pb)
    # deterministic: same bits on a second run, and batching does not change a window's result
    again = ctx.jointOptimization(m, pbs)
    single = ctx.jointOptimization(m, [pbs[1]])[0]
    assert np.array_equal(again[1]["pose"], outs[1]["pose"]) and np.array_equal(single["pose"], outs[1]["pose"])
    assert np.array_equal(single["pts"], outs[1]["pts"])


def test_local_ba_argument_checks(gpu, ba_map):
    g, ctx = gpu
    m, comps, mean, cov = ba_map
    pb = synth.make_ba_problem(mean, cov, 3, 1, 40, seed=3)
    bad = dict(pb)
    bad["obs_pt"] = pb["obs_pt"][::-1].copy()  # not grouped by point
    with pytest.raises(g.GmmlocError):
        ctx.jointOptimization(m, [bad])
    bad = dict(pb)
    bad["obs_kf"] = pb["obs_kf"] + 7
    with pytest.raises(g.GmmlocError):
        ctx.jointOptimization(m, [bad])
    assert ctx.jointOptimization(m, []) == []


@pytest.mark.parametrize("stop", [1, -1, -3, -7, -12, -14, -30, -1000])
def test_local_ba_stop_flag(gpu, oracle, ba_map, stop):
    """Localization::jointOptimization's pbStopFlag (localization_opt.cpp:541-542, 765-767, 791-795): raised before the call
    (1, -1) the window comes back untouched; -n replays a flag that rises at its n-th poll — inside stage 1 (-3), at the
    first iteration of stage 2 (-7), at the check after stage 2 (-12 when no LM trial was rejected: 1 + 5 + 5 polls, then
    the 12th), inside stage 3 (-14, -30), or never (-1000)."""
    g, ctx = gpu
    m, comps, mean, cov = ba_map
    pb = synth.make_ba_problem(mean, cov, 6, 2, 300, seed=77)
    want = oracle.local_ba(comps, pb["pose"], pb["fixed"], pb["pts"], pb["assoc"], pb["obs_pt"], pb["obs_kf"], pb["obs"],
                           pb["octave"], stop=stop)
    out = ctx.jointOptimization(m, [pb], stop=stop)[0]
    assert np.array_equal(out["iters"][:2], want["iters"][:2]), (out["iters"], want["iters"])
    if stop in (1, -1):
        assert list(out["iters"]) == [0, 0, 0]
        assert np.array_equal(out["pose"], pb["pose"]) and np.array_equal(out["pts"], pb["pts"])  # untouched, not even normalised
        assert not out["obs_outlier"].any() and not out["assoc_cleared"].any()
        return
    if stop > -1000 and want["iters"][2] < 40 and stop != -30:
        assert out["iters"][2] == want["iters"][2]
    assert np.array_equal(out["obs_outlier"], want["obs_outlier"])
    assert np.array_equal(out["assoc_cleared"], want["assoc_cleared"])
    for i in range(len(pb["pose"])):
        ang, dt = _pose_err(out["pose"][i], want["pose"][i])
        assert ang < TOL_RAD and dt < TOL_M, (i, ang, dt)
    assert np.abs(out["pts"] - want["pts"]).max() < 1e-4
    if stop == -3:
        assert list(want["iters"]) == [1, 0, 0]  # poll 1 = entry, poll 2 admits iteration 0 of stage 1, poll 3 stops it
    if stop == -12:
        assert list(want["iters"][:2]) == [5, 5]


def test_local_ba_stop_flag_raised_by_another_thread(gpu, ba_map):
    """The asynchronous use the reference makes of the flag: another host thread raises it while the BA runs
    (Localization::insertKeyFrame, localization.cpp:401-405).  Timing is not deterministic; the call must return, early."""
    import threading
    import time
    g, ctx = gpu
    m, comps, mean, cov = ba_map
    pbs = [synth.make_ba_problem(mean, cov, 12, 3, 1500, seed=90 + j) for j in range(8)]
    full = ctx.jointOptimization(m, pbs)
    flag = np.zeros(1, np.int32)

    def raise_it():
        time.sleep(0.02)
        flag[0] = 1

    th = threading.Thread(target=raise_it)
    th.start()
    outs = ctx.jointOptimization(m, pbs, stop=flag)
    th.join()
    assert sum(int(o["iters"].sum()) for o in outs) < sum(int(o["iters"].sum()) for o in full)


@pytest.mark.parametrize("cluster", [1, 2, 4, 8])
def test_local_ba_cluster_variants(gpu, oracle, ba_map, cluster):
    """A window spread over a thread-block cluster of 2 / 4 / 8 CTAs (distributed-shared-memory reductions, cluster barriers
    between the sweeps) against the oracle and against the single-CTA kernel: same flags, poses within tolerance."""
    g, ctx = gpu
    m, comps, mean, cov = ba_map
    pbs = [synth.make_ba_problem(mean, cov, 12, 3, 1500, seed=60), synth.make_ba_problem(mean, cov, 5, 2, 150, seed=61)]
    pbs[1]["prior_idx"] = 0
    ctx.set_option(g.capi.OPT_BA_CLUSTER, cluster)
    try:
        outs = ctx.jointOptimization(m, pbs)
        again = ctx.jointOptimization(m, pbs)
        stopped = ctx.jointOptimization(m, pbs[:1], stop=-9)[0]
    finally:
        ctx.set_option(g.capi.OPT_BA_CLUSTER, 0)
    for pb, out in zip(pbs, outs):
        want = oracle.local_ba(comps, pb["pose"], pb["fixed"], pb["pts"], pb["assoc"], pb["obs_pt"], pb["obs_kf"], pb["obs"],
                               pb["octave"], prior_idx=pb.get("prior_idx", -1))
        assert out["status"] == 0
        _check(out, want, pb)
    assert np.array_equal(again[0]["pose"], outs[0]["pose"]) and np.array_equal(again[0]["pts"], outs[0]["pts"])  # deterministic
    want = oracle.local_ba(comps, pbs[0]["pose"], pbs[0]["fixed"], pbs[0]["pts"], pbs[0]["assoc"], pbs[0]["obs_pt"], pbs[0]["obs_kf"],
                           pbs[0]["obs"], pbs[0]["octave"], stop=-9)
    assert np.array_equal(stopped["iters"], want["iters"])
