"""Oracle GN solves vs the independent dense NumPy restatement (CPU only)."""
import numpy as np
import pytest
from scipy.linalg import expm

from gmmloc_b200 import synth
from tests import np_reference as ref


@pytest.fixture(scope="module")
def small_scene(oracle):
    mean, cov = synth.make_map(400, 0.6, seed=11)
    comps = oracle.map_build(mean, cov)
    fr = synth.make_frames(mean, cov, 3, 60, seed=4, mono_frac=0.3)
    return comps, fr


def test_se3_exp_matches_matrix_exponential(oracle):
    rng = np.random.default_rng(1)
    for scale in (1e-7, 1e-3, 0.3, 2.0):
        xi = rng.normal(size=6) * scale
        pose = oracle.se3_oplus(xi, np.array([0, 0, 0, 1.0, 0, 0, 0]))
        tw = np.zeros((4, 4))
        tw[:3, :3] = ref.skew(xi[:3])
        tw[:3, 3] = xi[3:]
        T = expm(tw)
        np.testing.assert_allclose(synth.quat_to_rot(pose[:4]), T[:3, :3], atol=1e-12)
        np.testing.assert_allclose(pose[4:], T[:3, 3], atol=1e-12)


def test_se3_oplus_is_left_multiplication(oracle):
    rng = np.random.default_rng(2)
    q = rng.normal(size=4); q /= np.linalg.norm(q)
    pose = np.concatenate([q, rng.normal(size=3)])
    xi = rng.normal(size=6) * 0.1
    out = oracle.se3_oplus(xi, pose)
    dR, dt = ref.se3_exp(xi)
    np.testing.assert_allclose(synth.quat_to_rot(out[:4]), dR @ synth.quat_to_rot(pose[:4]), atol=1e-12)
    np.testing.assert_allclose(out[4:], dR @ pose[4:] + dt, atol=1e-12)
    assert out[3] >= 0 and abs(np.linalg.norm(out[:4]) - 1) < 1e-14


def test_sigma2_inv_table(oracle):
    for o in range(8):
        assert oracle.sigma2_inv(o) == ref.s2inv(o)
    assert abs(oracle.sigma2_inv(3) - 1.2 ** -6) < 1e-6


def test_projection_jacobians_vs_finite_differences():
    rng = np.random.default_rng(3)
    for _ in range(20):
        pc = np.array([rng.uniform(-1, 1), rng.uniform(-1, 1), rng.uniform(0.6, 4)])
        P = ref.dproj(pc, True)
        num = np.empty((3, 3))
        for j in range(3):
            d = np.zeros(3); d[j] = 1e-6
            num[:, j] = (ref.project(pc + d, True) - ref.project(pc - d, True)) / 2e-6
        np.testing.assert_allclose(P, num, rtol=1e-5, atol=1e-4)


def test_optimize_point_vs_numpy(oracle, small_scene):
    comps, fr = small_scene
    deg = np.flatnonzero(comps["is_degenerated"])
    rng = np.random.default_rng(5)
    n_acc = 0
    for f in range(3):
        pose = fr["pose0"][f]
        R = synth.quat_to_rot(pose[:4]); t = pose[4:]
        a, _ = oracle.associate(comps, fr["pts_w"][f])
        for i in range(0, 60, 3):
            if fr["obs"][f, i, 2] < 0:
                continue
            k = a[i] if comps["is_degenerated"][a[i]] else int(rng.choice(deg))
            x0 = fr["pts_w"][f, i].astype(np.float64)
            pz = min((R @ x0 + t)[2], 1.0)
            res, est, c2p, c2s = oracle.optimize_point(comps, k, x0, fr["obs"][f, i], fr["octave"][f, i], pose, pz * pz)
            c = comps[k]
            x, r2p, r2s = ref.optimize_point(x0, fr["obs"][f, i].astype(np.float64), fr["octave"][f, i], R, t, c["mean"],
                                             c["axis"].reshape(3, 3)[:, 0], 400.0 * pz * pz)
            np.testing.assert_allclose(est, x, atol=1e-9)
            np.testing.assert_allclose([c2p, c2s], [r2p, r2s], rtol=1e-7, atol=1e-9)
            assert res == (not (r2p > 7.815) and not (r2s > float(np.float32(0.0064)) * 400.0))
            n_acc += res
    assert n_acc > 0


@pytest.mark.parametrize("mode,robust", [(0, False), (0, True), (1, False), (2, False)])
def test_frame_solve_vs_dense_numpy(oracle, small_scene, mode, robust):
    comps, fr = small_scene
    prm = oracle.default_params(factor_mode=mode, robust=int(robust))
    for f in range(3):
        r = oracle.frame_solve(comps, fr["pose0"][f], fr["pts_w"][f], fr["obs"][f], fr["octave"][f], prm=prm)
        assert r["status"] == 0
        R, t, x, str_on, free = ref.frame_solve_dense(comps, fr["pose0"][f], fr["pts_w"][f], fr["obs"][f], fr["octave"][f],
                                                      r["assign"], mode=mode, robust=robust)
        assert ref.rot_angle(R, synth.quat_to_rot(r["pose"][:4])) < 1e-8
        np.testing.assert_allclose(r["pose"][4:], t, atol=1e-8)
        np.testing.assert_allclose(r["pts"], x, atol=1e-7)
        if mode != 1:
            assert str_on.sum() > 0  # the structure edges are really exercised


def test_frame_solve_gates_fix_the_active_set(oracle, small_scene):
    comps, fr = small_scene
    r = oracle.frame_solve(comps, fr["pose0"][0], fr["pts_w"][0], fr["obs"][0], fr["octave"][0], want_trace=True)
    inc = np.abs(r["trace"][:, :6]).max(axis=1)
    assert inc[-1] < 1e-6 * max(inc[0], 1e-12) + 1e-9  # converges to a fixed point


def test_frame_solve_numeric_failure_returns_frame_unsolved(oracle, small_scene):
    comps, fr = small_scene
    pts = fr["pts_w"][0].copy()
    pts[7] = np.nan
    r = oracle.frame_solve(comps, fr["pose0"][0], pts, fr["obs"][0], fr["octave"][0])
    assert r["status"] == 1
    q = fr["pose0"][0][:4] / np.linalg.norm(fr["pose0"][0][:4])
    np.testing.assert_allclose(r["pose"][:4], q if q[3] >= 0 else -q, atol=1e-15)
    np.testing.assert_array_equal(r["pose"][4:], fr["pose0"][0][4:])
    assert r["assign"][7] == -1


def test_batched_driver_equals_single_calls(oracle, small_scene):
    comps, fr = small_scene
    rb = oracle.frames_solve(comps, fr["pose0"], fr["pts_w"], fr["obs"], fr["octave"], nthreads=3, want_pts=True)
    for f in range(3):
        r = oracle.frame_solve(comps, fr["pose0"][f], fr["pts_w"][f], fr["obs"][f], fr["octave"][f])
        assert np.array_equal(rb["pose"][f], r["pose"]) and np.array_equal(rb["assign"][f], r["assign"])
        assert np.array_equal(rb["pts"][f], r["pts"]) and np.array_equal(rb["outlier"][f], r["outlier"])


def test_optimize_pose_lm_vs_independent_numpy(oracle):
    """f1: the 4-round Levenberg-Marquardt of Tracking::optimizeCurrentPose (oracle) against an independent
    NumPy restatement; poses agree to 1e-8, inlier sets exactly; the solve really recovers the GT pose."""
    mean, cov = synth.make_map(1500, 0.5, seed=21)
    fr = synth.make_tracking_frames(mean, cov, 3, 120, seed=21)
    for f in range(3):
        valid = np.ones(120, np.uint8)
        valid[::11] = 0
        n, pose, outl = oracle.optimize_pose_lm(fr["pose0"][f], fr["pts_w"][f], fr["obs"][f], fr["octave"][f], valid)
        n2, R, t, o2 = ref.optimize_pose_lm(fr["pose0"][f], fr["pts_w"][f], fr["obs"][f], fr["octave"][f], valid)
        assert n == n2 and np.array_equal(outl, o2)
        assert ref.rot_angle(R, synth.quat_to_rot(pose[:4])) < 1e-8
        np.testing.assert_allclose(pose[4:], t, atol=1e-8)
        assert not outl[::11].any() and 0 < outl.sum() < 60  # gross outliers found, unmatched features untouched
        d0 = np.abs(fr["pose0"][f][4:] - fr["pose_gt"][f][4:]).max()
        d1 = np.abs(pose[4:] - fr["pose_gt"][f][4:]).max()
        assert d1 < 0.2 * d0


def test_optimize_pose_lm_degenerate_inputs(oracle):
    mean, cov = synth.make_map(500, 0.5, seed=22)
    fr = synth.make_tracking_frames(mean, cov, 1, 40, seed=22)
    valid = np.zeros(40, np.uint8)
    valid[:2] = 1  # fewer than 3 correspondences: returns 0, pose untouched (tracking_opt.cpp:139-140)
    n, pose, outl = oracle.optimize_pose_lm(fr["pose0"][0], fr["pts_w"][0], fr["obs"][0], fr["octave"][0], valid)
    assert n == 0 and np.array_equal(pose, fr["pose0"][0]) and not outl.any()
    valid[:8] = 1  # fewer than 10 edges: a single round (tracking_opt.cpp:205-206)
    n, pose, outl = oracle.optimize_pose_lm(fr["pose0"][0], fr["pts_w"][0], fr["obs"][0], fr["octave"][0], valid)
    n2, R, t, o2 = ref.optimize_pose_lm(fr["pose0"][0], fr["pts_w"][0], fr["obs"][0], fr["octave"][0], valid)
    assert n == n2 and np.array_equal(outl, o2) and ref.rot_angle(R, synth.quat_to_rot(pose[:4])) < 1e-8


def test_optimize_triangulation_vs_numpy(oracle):
    mean, cov = synth.make_map(1500, 0.7, seed=41)
    comps = oracle.map_build(mean, cov)
    pr = synth.make_triangulation_problems(mean, cov, 60, seed=41)
    n_ok = 0
    for i in range(60):
        k = pr["comp"][i]
        res, est, e1, e2, cs = oracle.optimize_triangulation(comps, k, pr["x0"][i], pr["obs1"][i], pr["obs2"][i],
                                                             pr["octave1"][i], pr["pose1"][i], pr["pose2"][i])
        c = comps[k]
        x, r1, r2, rs = ref.optimize_triangulation(pr["x0"][i], pr["obs1"][i].astype(np.float64), pr["obs2"][i].astype(np.float64),
                                                   pr["octave1"][i], synth.quat_to_rot(pr["pose1"][i][:4]), pr["pose1"][i][4:],
                                                   synth.quat_to_rot(pr["pose2"][i][:4]), pr["pose2"][i][4:], c["mean"],
                                                   c["axis"].reshape(3, 3)[:, 0])
        np.testing.assert_allclose(est, x, atol=1e-8)
        np.testing.assert_allclose([e1, e2, cs], [r1, r2, rs], rtol=1e-6, atol=1e-9)
        th1 = 7.8 if pr["obs1"][i][2] >= 0 else 5.991
        th2 = 7.8 if pr["obs2"][i][2] >= 0 else 5.991
        assert res == (not rs > float(np.float32(0.0064)) * 400.0 and not (r1 > th1 or r2 > th2))
        n_ok += res
    assert n_ok > 20


def test_world_frame_schur_formulation_equals_camera_frame():
    """Design note for the next step of the frame GN, pinned numerically.

    The kernels evaluate the per-point Schur complement in the CAMERA frame: G = w P^T P, h = w P^T e,
    S_c = R S R^T, s_c = -R S (x - mu), D = G + S_c, T = G D^-1, A = [-[p_c]x  I], pose system sum A^T (T S_c) A and
    sum A^T (h - T (h + s_c)), point step R^T D^-1 (h + s_c - G A xi).  Rotating S into the camera frame is a third of
    the per-point work and has to be redone every sweep.  In the WORLD frame the structure block is constant:
    with J = P R, G_w = w J^T J, h_w = w J^T e, s_w = -S (x - mu), D_w = G_w + S, T_w = G_w D_w^-1, A_w = [-[x_w]x  I]
    the system sum A_w^T (T_w S) A_w, sum A_w^T (h_w - T_w (h_w + s_w)) is the camera-frame one expressed in the
    world twist xi_w = (R^T omega, R^T (omega x t + upsilon)), because A(p_c) xi = R A_w(x_w) xi_w; the point step is
    D_w^-1 (h_w + s_w - G_w A_w xi_w) with no rotation at all."""
    rng = np.random.default_rng(11)

    def skew(v):
        return np.array([[0, -v[2], v[1]], [v[2], 0, -v[0]], [-v[1], v[0], 0]])

    fx, fy, bf = 435.2, 435.2, 47.9
    q = rng.normal(size=4)
    q /= np.linalg.norm(q)
    x, y, z, w_ = q
    R = np.array([[1 - 2 * (y * y + z * z), 2 * (x * y - z * w_), 2 * (x * z + y * w_)],
                  [2 * (x * y + z * w_), 1 - 2 * (x * x + z * z), 2 * (y * z - x * w_)],
                  [2 * (x * z - y * w_), 2 * (y * z + x * w_), 1 - 2 * (x * x + y * y)]])
    t = rng.normal(size=3)
    n = 40
    pc = np.column_stack([rng.uniform(-1.5, 1.5, n), rng.uniform(-1, 1, n), rng.uniform(0.6, 3.5, n)])
    xw = (pc - t) @ R  # x_w = R^T (p_c - t)
    Hc, bc, Hw, bw = np.zeros((6, 6)), np.zeros(6), np.zeros((6, 6)), np.zeros(6)
    per_point = []
    for i in range(n):
        X, Y, Z = pc[i]
        P = np.array([[fx / Z, 0, -fx * X / Z**2], [0, fy / Z, -fy * Y / Z**2], [fx / Z, 0, -fx * X / Z**2 + bf / Z**2]])
        e = rng.normal(size=3)
        wgt = 1.0 / 1.44 ** rng.integers(0, 8)
        Lm = rng.normal(size=(3, 3))
        S = Lm @ Lm.T * 50 + np.eye(3)
        d = rng.normal(size=3) * 0.02
        # camera frame
        G, h = wgt * P.T @ P, wgt * P.T @ e
        Sc, sc = R @ S @ R.T, -R @ S @ d
        D = G + Sc
        T = G @ np.linalg.inv(D)
        A = np.hstack([-skew(pc[i]), np.eye(3)])
        Hc += A.T @ (T @ Sc) @ A
        bc += A.T @ (h - T @ (h + sc))
        # world frame
        J = P @ R
        Gw, hw, sw = wgt * J.T @ J, wgt * J.T @ e, -S @ d
        Dw = Gw + S
        Tw = Gw @ np.linalg.inv(Dw)
        Aw = np.hstack([-skew(xw[i]), np.eye(3)])
        Hw += Aw.T @ (Tw @ S) @ Aw
        bw += Aw.T @ (hw - Tw @ (hw + sw))
        per_point.append((G, h, Sc, sc, D, A, Gw, hw, sw, Dw, Aw))
    xi = np.linalg.solve(Hc, bc)
    xiw = np.linalg.solve(Hw, bw)
    om = R @ xiw[:3]
    up = R @ xiw[3:] - np.cross(om, t)
    assert np.allclose(np.concatenate([om, up]), xi, rtol=1e-9, atol=1e-12)
    for (G, h, Sc, sc, D, A, Gw, hw, sw, Dw, Aw) in per_point:
        dx_c = R.T @ np.linalg.solve(D, h + sc - G @ (A @ xi))
        dx_w = np.linalg.solve(Dw, hw + sw - Gw @ (Aw @ xiw))
        assert np.allclose(dx_c, dx_w, rtol=1e-8, atol=1e-12)
