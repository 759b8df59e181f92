"""Independent NumPy restatement of the solves (test helper; fp64, dense, NO Schur complement, NO
shared code with oracle/ or the CUDA kernels).  The frame solve assembles the full (6 + 3 N_free)
normal equations and solves them with numpy.linalg — if the oracle's per-point Schur elimination or the
kernel's camera-frame formulation were wrong, the poses would not agree."""
import numpy as np

from gmmloc_b200 import synth

CAM = dict(fx=synth.FX, fy=synth.FY, cx=synth.CX, cy=synth.CY, bf=synth.BF)


def s2inv(octave):
    sf = np.float32(1.0)
    out = np.float32(1.0)
    for _ in range(int(octave)):
        sf = np.float32(sf * np.float32(1.2))
        out = np.float32(1.0) / np.float32(sf * sf)
    return float(out)


def skew(v):
    return np.array([[0, -v[2], v[1]], [v[2], 0, -v[0]], [-v[1], v[0], 0]])


def se3_exp(xi):
    """g2o SE3Quat::exp, [omega; upsilon] (SURVEY Appx B.1) via the closed form."""
    w, u = xi[:3], xi[3:]
    th = np.linalg.norm(w)
    Om = skew(w)
    if th < 1e-5:
        R = np.eye(3) + Om + 0.5 * Om @ Om
        V = np.eye(3) + 0.5 * Om + Om @ Om / 6
    else:
        R = np.eye(3) + np.sin(th) / th * Om + (1 - np.cos(th)) / th ** 2 * Om @ Om
        V = np.eye(3) + (1 - np.cos(th)) / th ** 2 * Om + (th - np.sin(th)) / th ** 3 * Om @ Om
    return R, V @ u


def project(pc, stereo):
    u = CAM["fx"] * pc[0] / pc[2] + CAM["cx"]
    v = CAM["fy"] * pc[1] / pc[2] + CAM["cy"]
    if stereo:
        return np.array([u, v, u - CAM["bf"] / pc[2]])
    return np.array([u, v])


def dproj(pc, stereo):
    x, y, z = pc
    P = [[CAM["fx"] / z, 0, -CAM["fx"] * x / z ** 2], [0, CAM["fy"] / z, -CAM["fy"] * y / z ** 2]]
    if stereo:
        P.append([CAM["fx"] / z, 0, -CAM["fx"] * x / z ** 2 + CAM["bf"] / z ** 2])
    return np.array(P)


def optimize_point(x0, obs, octave, R, t, mu, n, w_str, iters=5):
    x = np.array(x0, float)
    w = s2inv(octave)
    for _ in range(iters):
        pc = R @ x + t
        e = obs - project(pc, True)
        J = -dproj(pc, True) @ R
        es = n @ (x - mu)
        c2p, c2s = w * e @ e, w_str * es * es
        H = w * J.T @ J + w_str * np.outer(n, n)
        b = -(w * J.T @ e + w_str * n * es)
        x = x + np.linalg.solve(H, b)
    return x, c2p, c2s


def frame_solve_dense(comps, pose0, pts_w, obs, octave, assign, iters=10, mode=0, robust=False, ba_lambda2=400.0,
                      tri_str_thresh=float(np.float32(0.0064)), gate=9.0):
    q = pose0[:4] / np.linalg.norm(pose0[:4])
    if q[3] < 0:
        q = -q
    R = synth.quat_to_rot(q)
    t = np.array(pose0[4:], float)
    N = len(pts_w)
    x = np.asarray(pts_w, np.float64).copy()
    stereo = obs[:, 2] >= 0
    pose_free, pts_free = mode != 2, mode != 1
    str_on = np.zeros(N, bool)
    for i in range(N):
        k = assign[i]
        if k >= 0 and pts_free:
            c = comps[k]
            d = x[i] - c["mean"]
            if c["is_degenerated"]:
                n = c["axis"].reshape(3, 3)[:, 0]
                po = c["mean"] - (-R.T @ t)
                view_cos = abs(po @ n) / np.linalg.norm(po)
                str_on[i] = (n @ d) ** 2 <= tri_str_thresh and not view_cos < np.cos(np.deg2rad(78.0))
            else:
                str_on[i] = d @ c["cov_inv"].reshape(3, 3) @ d <= gate
    free = pts_free & (stereo | str_on)
    idx = np.full(N, -1)
    idx[free] = np.arange(free.sum())
    npose = 6 if pose_free else 0
    dim = npose + 3 * int(free.sum())
    deltas = {True: float(np.float32(np.sqrt(7.815))), False: float(np.float32(np.sqrt(5.991)))}
    for _ in range(iters):
        H = np.zeros((dim, dim))
        b = np.zeros(dim)
        for i in range(N):
            pc = R @ x[i] + t
            st = bool(stereo[i])
            m = 3 if st else 2
            e = obs[i, :m].astype(np.float64) - project(pc, st)
            w = s2inv(octave[i])
            if robust:
                chi = w * e @ e
                if chi > deltas[st] ** 2:
                    w *= deltas[st] / np.sqrt(chi)
            P = dproj(pc, st)
            cols, Js = [], []
            if pose_free:
                Js.append(-P @ np.hstack([-skew(pc), np.eye(3)]))
                cols.append(np.arange(6))
            if free[i]:
                Js.append(-P @ R)
                cols.append(npose + 3 * idx[i] + np.arange(3))
            if Js:
                J = np.hstack(Js)
                c = np.concatenate(cols)
                H[np.ix_(c, c)] += w * J.T @ J
                b[c] -= w * J.T @ e
            if free[i] and str_on[i]:
                cmp_ = comps[assign[i]]
                d = x[i] - cmp_["mean"]
                c = npose + 3 * idx[i] + np.arange(3)
                if cmp_["is_degenerated"]:
                    n = cmp_["axis"].reshape(3, 3)[:, 0]
                    H[np.ix_(c, c)] += ba_lambda2 * np.outer(n, n)
                    b[c] -= ba_lambda2 * n * (n @ d)
                else:
                    L = cmp_["sqrt_info"].reshape(3, 3)
                    H[np.ix_(c, c)] += L @ L.T
                    b[c] -= L @ (L.T @ d)
        if dim == 0:
            break
        dx = np.linalg.solve(H, b)
        if pose_free:
            dR, dt = se3_exp(dx[:6])
            R, t = dR @ R, dR @ t + dt
        for i in np.flatnonzero(free):
            x[i] += dx[npose + 3 * idx[i]: npose + 3 * idx[i] + 3]
    return R, t, x, str_on, free


def rot_angle(Ra, Rb):
    """geodesic angle between two rotations (atan2 form: accurate for tiny angles, unlike arccos)."""
    M = Ra.T @ Rb
    sin = 0.5 * np.linalg.norm([M[2, 1] - M[1, 2], M[0, 2] - M[2, 0], M[1, 0] - M[0, 1]])
    cos = 0.5 * (np.trace(M) - 1)
    return float(np.arctan2(sin, cos))
