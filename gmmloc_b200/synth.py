"""Deterministic synthetic maps and frames for parity tests and bench.py.

Pure NumPy (host side, no kernels).  Map statistics follow the shipped maps
(gmmloc_ros/data/map/v1.gmm: 95.6 % planar components, sigma_min ~1.6 mm, in-plane 5-13 cm); frames
follow the reference's landmark generation: a stereo keypoint (u, v, u_r, octave) is unprojected at
the current pose estimate (Frame::unproject3, gmmloc/src/types/frame.cpp:27-36; depth z = bf / (u - u_r),
frame.cpp:330) and handed to the association as an fp32 world point.
"""
from __future__ import annotations

import numpy as np

SEED = 0x9E3779B97F4A7C15
# gmmloc_ros/cfg/v1.yaml:5-17
FX = FY = 435.2046959714599
CX, CY = 367.4517211914062, 252.2008514404297
BF = 47.90639384423901
W, H = 752, 480
Z_NEAR, Z_FAR = 0.5, 3.85  # th_depth = bf * 35 / fx (init_config.hpp:61, v1.yaml:23)
BOX_LO = np.array([-5.0, -5.0, 0.0])
BOX_HI = np.array([5.0, 5.0, 3.5])


def _rng(*stream):
    return np.random.Generator(np.random.Philox(key=SEED % (1 << 64), counter=[*stream, 0, 0, 0][:4]))


def _rand_rot(rng, n):
    q = rng.normal(size=(n, 4))
    q /= np.linalg.norm(q, axis=1, keepdims=True)
    return quat_to_rot(q)


def quat_to_rot(q):
    """q = (x, y, z, w) -> R, vectorised."""
    q = np.asarray(q, np.float64)
    x, y, z, w = q[..., 0], q[..., 1], q[..., 2], q[..., 3]
    R = np.empty(q.shape[:-1] + (3, 3))
    R[..., 0, 0] = 1 - 2 * (y * y + z * z); R[..., 0, 1] = 2 * (x * y - z * w); R[..., 0, 2] = 2 * (x * z + y * w)
    R[..., 1, 0] = 2 * (x * y + z * w); R[..., 1, 1] = 1 - 2 * (x * x + z * z); R[..., 1, 2] = 2 * (y * z - x * w)
    R[..., 2, 0] = 2 * (x * z - y * w); R[..., 2, 1] = 2 * (y * z + x * w); R[..., 2, 2] = 1 - 2 * (x * x + y * y)
    return R


def rot_to_quat(R):
    """single 3x3 -> (x, y, z, w), w >= 0."""
    t = np.trace(R)
    if t > 0:
        s = np.sqrt(t + 1.0) * 2
        q = np.array([(R[2, 1] - R[1, 2]) / s, (R[0, 2] - R[2, 0]) / s, (R[1, 0] - R[0, 1]) / s, 0.25 * s])
    else:
        i = int(np.argmax(np.diag(R)))
        j, k = (i + 1) % 3, (i + 2) % 3
        s = np.sqrt(R[i, i] - R[j, j] - R[k, k] + 1.0) * 2
        q = np.empty(4)
        q[i] = 0.25 * s
        q[j] = (R[j, i] + R[i, j]) / s
        q[k] = (R[k, i] + R[i, k]) / s
        q[3] = (R[k, j] - R[j, k]) / s
    if q[3] < 0:
        q = -q
    return q / np.linalg.norm(q)


def make_map(K: int, p_deg: float = 0.5, seed: int = 0):
    """K components: mu ~ U(box), Sigma = R diag(s) R^T, planar (s1 in [1e-6,1e-4)) with prob p_deg."""
    rng = _rng(1, seed)
    mean = rng.uniform(BOX_LO, BOX_HI, size=(K, 3))
    R = _rand_rot(rng, K)
    deg = rng.random(K) < p_deg
    s1 = np.where(deg, 10 ** rng.uniform(-6, -4, K), 10 ** rng.uniform(-4, -2, K))
    s23 = 10 ** rng.uniform(-3, -1, size=(K, 2))
    s = np.concatenate([s1[:, None], s23], axis=1)
    cov = np.einsum("kij,kj,klj->kil", R, s, R)
    cov = 0.5 * (cov + cov.transpose(0, 2, 1))
    return mean, cov.reshape(K, 9)


def _sample_pose(rng):
    pos = rng.uniform(BOX_LO + 1.0, BOX_HI - 1.0)
    yaw = rng.uniform(-np.pi, np.pi)
    pitch, roll = np.deg2rad(rng.uniform(-20, 20, 2))
    cy_, sy_ = np.cos(yaw), np.sin(yaw)
    Rz = np.array([[cy_, -sy_, 0], [sy_, cy_, 0], [0, 0, 1]])
    # camera looks along +z_c; world z is up: base orientation maps z_c -> world x, x_c -> -world y, y_c -> -world z
    base = np.array([[0, 0, 1], [-1, 0, 0], [0, -1, 0]], float)
    cp, sp = np.cos(pitch), np.sin(pitch)
    cr, sr = np.cos(roll), np.sin(roll)
    Rx = np.array([[1, 0, 0], [0, cp, -sp], [0, sp, cp]])
    Rr = np.array([[cr, -sr, 0], [sr, cr, 0], [0, 0, 1]])
    R_wc = Rz @ base @ Rx @ Rr
    return R_wc, pos


def _se3_exp(xi):
    w, u = xi[:3], xi[3:]
    th = np.linalg.norm(w)
    Om = np.array([[0, -w[2], w[1]], [w[2], 0, -w[0]], [-w[1], w[0], 0]])
    if th < 1e-9:
        Rm, V = np.eye(3) + Om, np.eye(3) + 0.5 * Om
    else:
        Rm = np.eye(3) + np.sin(th) / th * Om + (1 - np.cos(th)) / th ** 2 * Om @ Om
        V = np.eye(3) + (1 - np.cos(th)) / th ** 2 * Om + (th - np.sin(th)) / th ** 3 * Om @ Om
    return Rm, V @ u


def make_frames(mean, cov, B: int, N: int, seed: int = 0, mono_frac: float = 0.0, first_frame: int = 0, poses_wc=None):
    """B frames of N landmarks.  Returns dict(pose0 Bx7 (T_cw: qx qy qz qw tx ty tz), pts_w BxNx3 f32,
    obs BxNx3 f32 (u v u_r; u_r = -1 => mono), octave BxN u8, pose_gt Bx7).
    poses_wc (optional, B x 7: x y z qx qy qz qw of T_wc, the layout of the reference's gt_sync files,
    dataloader.cpp:118-138) replaces the random ground-truth poses (EuRoC replay)."""
    mean = np.asarray(mean, np.float64).reshape(-1, 3)
    cov = np.asarray(cov, np.float64).reshape(-1, 3, 3)
    K = mean.shape[0]
    Lc = np.linalg.cholesky(cov)
    pose0 = np.empty((B, 7))
    pose_gt = np.empty((B, 7))
    pts_w = np.empty((B, N, 3), np.float32)
    obs = np.empty((B, N, 3), np.float32)
    octave = np.empty((B, N), np.uint8)
    for f in range(B):
        rng = _rng(2, seed, first_frame + f)
        for attempt in range(64):
            if poses_wc is not None and attempt == 0:
                t_wc = np.asarray(poses_wc[f][:3], np.float64)
                R_wc = quat_to_rot(np.asarray(poses_wc[f][3:7], np.float64) / np.linalg.norm(poses_wc[f][3:7]))
            elif poses_wc is not None:
                # the recorded pose sees (almost) no component in the stereo depth range: relax the range
                vis = np.flatnonzero((z > 0.2) & (u >= 0) & (u < W) & (v >= 0) & (v < H))
                if vis.size == 0:
                    vis = np.argsort(np.linalg.norm(mean - t_wc, axis=1))[:64]
                break
            else:
                R_wc, t_wc = _sample_pose(rng)
            R_cw = R_wc.T
            t_cw = -R_cw @ t_wc
            mc = mean @ R_cw.T + t_cw
            z = mc[:, 2]
            with np.errstate(divide="ignore", invalid="ignore"):
                u = FX * mc[:, 0] / z + CX
                v = FY * mc[:, 1] / z + CY
            vis = np.flatnonzero((z > Z_NEAR) & (z < Z_FAR) & (u >= 0) & (u < W) & (v >= 0) & (v < H))
            if vis.size >= 8:
                break
        else:
            raise RuntimeError("no pose with visible components found")
        k = vis[rng.integers(0, vis.size, N)]
        x = mean[k] + np.einsum("nij,nj->ni", Lc[k], rng.normal(size=(N, 3)))
        xc = x @ R_cw.T + t_cw
        zc = np.maximum(xc[:, 2], 0.2)
        octv = rng.integers(0, 8, N)
        sig = 1.2 ** octv
        uu = FX * xc[:, 0] / zc + CX + rng.normal(size=N) * sig
        vv = FY * xc[:, 1] / zc + CY + rng.normal(size=N) * sig
        ur = uu - BF / zc + rng.normal(size=N) * 0.25
        ur = np.minimum(ur, uu - BF / 40.0)  # keep disparity positive (depth < 40 m)
        # initial pose = exp(xi) * GT
        xi = np.concatenate([rng.normal(size=3) * np.deg2rad(0.5), rng.normal(size=3) * 0.02])
        dR, dt = _se3_exp(xi)
        R0 = dR @ R_cw
        t0 = dR @ t_cw + dt
        # landmark = unproject3 of the noisy observation at the INITIAL pose (gmmloc_opt.cpp:73)
        zz = BF / (uu - ur)
        Xc = np.stack([zz * (uu - CX) / FX, zz * (vv - CY) / FY, zz], axis=1)
        Xw = (Xc - t0) @ R0  # R0^T (Xc - t0)
        mono = rng.random(N) < mono_frac
        if mono.any():
            # a monocular point stands in for a triangulated one: true point + 2 cm noise
            Xw[mono] = x[mono] + rng.normal(size=(int(mono.sum()), 3)) * 0.02
            ur = np.where(mono, -1.0, ur)
        pose0[f, :4] = rot_to_quat(R0)
        pose0[f, 4:] = t0
        pose_gt[f, :4] = rot_to_quat(R_cw)
        pose_gt[f, 4:] = t_cw
        pts_w[f] = Xw
        obs[f, :, 0], obs[f, :, 1], obs[f, :, 2] = uu, vv, ur
        octave[f] = octv
    return dict(pose0=pose0, pts_w=pts_w, obs=obs, octave=octave, pose_gt=pose_gt)


def make_points_near(mean, cov, M: int, seed: int = 0, spread: float = 1.0):
    """M fp32 points drawn from the mixture itself (x ~ N(mu_k, spread^2 Sigma_k), k uniform)."""
    mean = np.asarray(mean, np.float64).reshape(-1, 3)
    cov = np.asarray(cov, np.float64).reshape(-1, 3, 3)
    rng = _rng(3, seed)
    k = rng.integers(0, mean.shape[0], M)
    Lc = np.linalg.cholesky(cov)
    x = mean[k] + spread * np.einsum("nij,nj->ni", Lc[k], rng.normal(size=(M, 3)))
    return x.astype(np.float32), k


def make_tracking_frames(mean, cov, B: int, N: int, seed: int = 0, mono_frac: float = 0.2, outlier_frac: float = 0.1,
                         lm_noise: float = 0.01):
    """Frames for the pose-only solve (Tracking::optimizeCurrentPose): landmarks are MAP POINTS (true point
    + lm_noise), observations are noisy projections at the ground-truth pose with a fraction of gross
    outliers (wrong matches), the initial pose is the ground truth perturbed by ~1 deg / 5 cm."""
    fr = make_frames(mean, cov, B, N, seed=seed, mono_frac=1.0)  # mono branch: landmark = true point + 2 cm
    rng = _rng(4, seed)
    pts = fr["pts_w"].astype(np.float64)
    pts += rng.normal(size=pts.shape) * (lm_noise - 0.02 if lm_noise > 0.02 else 0.0)
    obs = np.empty_like(fr["obs"])
    pose0 = np.empty_like(fr["pose0"])
    for f in range(B):
        q, t = fr["pose_gt"][f, :4], fr["pose_gt"][f, 4:]
        R = quat_to_rot(q)
        pc = pts[f] @ R.T + t
        z = np.maximum(pc[:, 2], 0.2)
        sig = 1.2 ** fr["octave"][f].astype(np.float64)
        u = FX * pc[:, 0] / z + CX + rng.normal(size=N) * sig
        v = FY * pc[:, 1] / z + CY + rng.normal(size=N) * sig
        ur = u - BF / z + rng.normal(size=N) * 0.25
        bad = rng.random(N) < outlier_frac
        u = np.where(bad, u + rng.normal(size=N) * 30.0, u)
        v = np.where(bad, v + rng.normal(size=N) * 30.0, v)
        mono = rng.random(N) < mono_frac
        ur = np.where(mono, -1.0, ur)
        obs[f, :, 0], obs[f, :, 1], obs[f, :, 2] = u, v, ur
        xi = np.concatenate([rng.normal(size=3) * np.deg2rad(1.0), rng.normal(size=3) * 0.05])
        dR, dt = _se3_exp(xi)
        pose0[f, :4] = rot_to_quat(dR @ R)
        pose0[f, 4:] = dR @ t + dt
    return dict(pose0=pose0, pts_w=pts.astype(np.float32), obs=obs, octave=fr["octave"], pose_gt=fr["pose_gt"])


def make_triangulation_problems(mean, cov, M: int, seed: int = 0, mono_frac: float = 0.5):
    """M two-view problems for Localization::optimizeTriangulationVec: a true point on a planar component,
    seen from a keyframe pair ~20 cm apart; view 1 / view 2 stereo or mono; x0 = true point + 3 cm noise.
    Returns dict(x0 Mx3 f64, obs1/obs2 Mx3 f32, octave1 M u8, pose1/pose2 Mx7, comp M i32 (generating comp))."""
    mean = np.asarray(mean, np.float64).reshape(-1, 3)
    cov = np.asarray(cov, np.float64).reshape(-1, 3, 3)
    rng = _rng(5, seed)
    w_, _v = np.linalg.eigh(cov)
    planar = np.flatnonzero(w_[:, 0] < 1e-4)
    Lc = np.linalg.cholesky(cov)
    out = dict(x0=np.empty((M, 3)), obs1=np.empty((M, 3), np.float32), obs2=np.empty((M, 3), np.float32),
               octave1=np.empty(M, np.uint8), pose1=np.empty((M, 7)), pose2=np.empty((M, 7)), comp=np.empty(M, np.int32))
    i = 0
    while i < M:
        R_wc, t_wc = _sample_pose(rng)
        R1 = R_wc.T
        t1 = -R1 @ t_wc
        dR, dt = _se3_exp(np.concatenate([rng.normal(size=3) * np.deg2rad(3.0), rng.normal(size=3) * 0.15]))
        R2, t2 = dR @ R1, dR @ t1 + dt
        mc = mean[planar] @ R1.T + t1
        vis = planar[(mc[:, 2] > Z_NEAR) & (mc[:, 2] < Z_FAR) & (np.abs(mc[:, 0] / mc[:, 2]) < 0.7) & (np.abs(mc[:, 1] / mc[:, 2]) < 0.45)]
        if vis.size == 0:
            continue
        for k in vis[rng.integers(0, vis.size, min(8, M - i))]:
            x = mean[k] + Lc[k] @ rng.normal(size=3)
            oc = int(rng.integers(0, 8))
            sig = 1.2 ** oc
            obs = []
            for (R, t) in ((R1, t1), (R2, t2)):
                pc = R @ x + t
                u = FX * pc[0] / pc[2] + CX + rng.normal() * sig
                v = FY * pc[1] / pc[2] + CY + rng.normal() * sig
                ur = u - BF / pc[2] + rng.normal() * 0.25 if rng.random() >= mono_frac else -1.0
                obs.append((u, v, ur))
            out["x0"][i] = x + rng.normal(size=3) * 0.03
            out["obs1"][i], out["obs2"][i] = obs
            out["octave1"][i] = oc
            out["pose1"][i, :4], out["pose1"][i, 4:] = rot_to_quat(R1), t1
            out["pose2"][i, :4], out["pose2"][i, 4:] = rot_to_quat(R2), t2
            out["comp"][i] = k
            i += 1
    return out


def make_ba_problem(mean, cov, n_kf: int = 6, n_fixed: int = 2, n_pts: int = 300, seed: int = 0, mono_frac: float = 0.2,
                    outlier_frac: float = 0.03, wrong_assoc_frac: float = 0.05, assoc_frac: float = 0.8):
    """A local-BA window (Localization::jointOptimization): n_kf keyframes along a short trajectory looking at the
    same part of the map (the last n_fixed are "fixed_kfs"), n_pts map points drawn from visible components and
    observed where visible.  Returns dict(pose P x 7 initial T_cw (free ones perturbed ~0.5 deg / 2 cm), pose_gt,
    fixed P u8, pts L x 3 initial (true + 2 cm), pts_gt, assoc L i32, obs_pt / obs_kf E i32 (sorted by point),
    obs E x 3 f32, octave E u8)."""
    mean = np.asarray(mean, np.float64).reshape(-1, 3)
    cov = np.asarray(cov, np.float64).reshape(-1, 3, 3)
    rng = _rng(6, seed)
    Lc = np.linalg.cholesky(cov)
    while True:
        R_wc, t_wc = _sample_pose(rng)
        R0 = R_wc.T
        t0 = -R0 @ t_wc
        mc = mean @ R0.T + t0
        vis = np.flatnonzero((mc[:, 2] > Z_NEAR + 0.3) & (mc[:, 2] < Z_FAR) & (np.abs(mc[:, 0] / mc[:, 2]) < 0.6) &
                             (np.abs(mc[:, 1] / mc[:, 2]) < 0.4))
        if vis.size >= 8:
            break
    poses_gt = []
    for j in range(n_kf):  # a short arc: ~8 cm and ~1.5 deg per keyframe
        xi = np.concatenate([rng.normal(size=3) * np.deg2rad(1.5), rng.normal(size=3) * 0.08]) * j
        dR, dt = _se3_exp(xi)
        poses_gt.append((dR @ R0, dR @ t0 + dt))
    k = vis[rng.integers(0, vis.size, n_pts)]
    pts_gt = mean[k] + np.einsum("nij,nj->ni", Lc[k], rng.normal(size=(n_pts, 3)))
    assoc = np.where(rng.random(n_pts) < assoc_frac, k, -1).astype(np.int32)
    wrong = (rng.random(n_pts) < wrong_assoc_frac) & (assoc >= 0)
    assoc[wrong] = rng.integers(0, mean.shape[0], int(wrong.sum()))
    obs_pt, obs_kf, obs, octave = [], [], [], []
    for l in range(n_pts):
        for j, (R, t) in enumerate(poses_gt):
            pc = R @ pts_gt[l] + t
            if pc[2] < 0.3:
                continue
            oc = int(rng.integers(0, 8))
            sig = 1.2 ** oc
            u = FX * pc[0] / pc[2] + CX + rng.normal() * sig
            v = FY * pc[1] / pc[2] + CY + rng.normal() * sig
            if not (0 <= u < W and 0 <= v < H) or rng.random() < 0.15:
                continue
            ur = u - BF / pc[2] + rng.normal() * 0.25 if rng.random() >= mono_frac else -1.0
            if rng.random() < outlier_frac:
                u += rng.normal() * 25.0
                v += rng.normal() * 25.0
            obs_pt.append(l); obs_kf.append(j); obs.append((u, v, ur)); octave.append(oc)
    P = n_kf
    pose = np.empty((P, 7))
    pose_gt = np.empty((P, 7))
    fixed = np.zeros(P, np.uint8)
    fixed[P - n_fixed:] = 1
    for j, (R, t) in enumerate(poses_gt):
        pose_gt[j, :4], pose_gt[j, 4:] = rot_to_quat(R), t
        if fixed[j]:
            pose[j] = pose_gt[j]
        else:
            dR, dt = _se3_exp(np.concatenate([rng.normal(size=3) * np.deg2rad(0.5), rng.normal(size=3) * 0.02]))
            pose[j, :4], pose[j, 4:] = rot_to_quat(dR @ R), dR @ t + dt
    pts = pts_gt + rng.normal(size=pts_gt.shape) * 0.02
    return dict(pose=pose, pose_gt=pose_gt, fixed=fixed, pts=pts, pts_gt=pts_gt, assoc=assoc,
                obs_pt=np.asarray(obs_pt, np.int32), obs_kf=np.asarray(obs_kf, np.int32),
                obs=np.asarray(obs, np.float32).reshape(-1, 3), octave=np.asarray(octave, np.uint8))
