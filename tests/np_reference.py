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


def huber(c2, delta):
    """g2o RobustKernelHuber: (rho0, rho1)."""
    if c2 <= delta * delta:
        return c2, 1.0
    s = np.sqrt(c2)
    return 2 * s * delta - delta * delta, delta / s


def optimize_pose_lm(pose0, pts_w, obs, octave, valid=None):
    """Independent NumPy restatement of Tracking::optimizeCurrentPose (tracking_opt.cpp:21-217) with g2o's
    Levenberg-Marquardt (SURVEY Appx B.7).  Returns (n_inliers, R, t, is_outlier)."""
    N = len(pts_w)
    idx = [i for i in range(N) if valid is None or valid[i]]
    q0 = np.asarray(pose0[:4], float)
    q0 = q0 / np.linalg.norm(q0) * (1 if q0[3] >= 0 else -1)
    R0, t0 = synth.quat_to_rot(q0), np.asarray(pose0[4:], float)
    outl = np.zeros(N, np.uint8)
    if len(idx) < 3:
        return 0, R0, t0, outl
    X = np.asarray(pts_w, np.float64)
    ob = np.asarray(obs, np.float64)
    st = {i: bool(obs[i][2] >= 0) for i in idx}
    w = {i: s2inv(octave[i]) for i in idx}
    dl = {True: float(np.float32(np.sqrt(7.815))), False: float(np.float32(np.sqrt(5.991)))}
    th = {True: np.float32(7.815), False: np.float32(5.991)}
    level = {i: 0 for i in idx}
    is_out = {i: False for i in idx}
    err = {i: np.zeros(3) for i in idx}
    robust = True

    def resid(i, R, t):
        m = 3 if st[i] else 2
        return ob[i, :m] - project(R @ X[i] + t, st[i])

    def active(R, t):
        chi = 0.0
        for i in idx:
            if level[i]:
                continue
            e = resid(i, R, t)
            err[i] = e
            c2 = w[i] * (e @ e)
            chi += huber(c2, dl[st[i]])[0] if robust else c2
        return chi

    n_bad = 0
    for rnd in range(4):
        R, t = R0.copy(), t0.copy()
        lam, ni, ok = 0.0, 2.0, True
        for it in range(10):
            if not ok:
                break
            cur = active(R, t)
            H, b = np.zeros((6, 6)), np.zeros(6)
            for i in idx:
                if level[i]:
                    continue
                pc = R @ X[i] + t
                e = resid(i, R, t)
                J = -dproj(pc, st[i]) @ np.hstack([-skew(pc), np.eye(3)])
                r1 = huber(w[i] * (e @ e), dl[st[i]])[1] if robust else 1.0
                H += r1 * w[i] * J.T @ J
                b -= r1 * w[i] * J.T @ e
            if it == 0:
                lam, ni = 1e-5 * np.abs(np.diag(H)).max(), 2.0
            q = 0
            while True:
                dx = np.linalg.solve(H + lam * np.eye(6), b)
                dR, dt = se3_exp(dx)
                Rt, tt = dR @ R, dR @ t + dt
                tmp = active(Rt, tt)
                rho = (cur - tmp) / (dx @ (lam * dx + b) + 1e-3)
                if rho > 0 and np.isfinite(tmp):
                    lam *= max(1 / 3, min(1 - (2 * rho - 1) ** 3, 2 / 3))
                    ni, cur, R, t = 2.0, tmp, Rt, tt
                else:
                    lam *= ni
                    ni *= 2
                q += 1
                if not (rho < 0 and q < 10):
                    break
            if q == 10 or rho == 0:
                ok = False
        n_bad = 0
        for i in idx:
            if is_out[i]:
                err[i] = resid(i, R, t)
            c2 = np.float32(w[i] * (err[i] @ err[i]))
            if c2 > th[st[i]]:
                is_out[i], level[i] = True, 1
                n_bad += 1
            else:
                is_out[i], level[i] = False, 0
        if rnd == 2:
            robust = False
        if len(idx) < 10:
            break
    for i in idx:
        outl[i] = is_out[i]
    return len(idx) - n_bad, R, t, outl


def optimize_triangulation(x0, obs1, obs2, octave1, R1, t1, R2, t2, mu, n, w_str=400.0, iters=20):
    """Independent restatement of one optimizeTriangulationVec candidate (localization_opt.cpp:27-204)."""
    x = np.array(x0, float)
    w = s2inv(octave1)
    views = [(R1, t1, np.asarray(obs1, float)), (R2, t2, np.asarray(obs2, float))]
    errs = [0.0, 0.0]
    for _ in range(iters):
        H = w_str * np.outer(n, n)
        es = n @ (x - mu)
        b = -w_str * n * es
        for v, (R, t, ob) in enumerate(views):
            st = ob[2] >= 0
            m = 3 if st else 2
            pc = R @ x + t
            e = ob[:m] - project(pc, st)
            J = -dproj(pc, st) @ R
            errs[v] = w * e @ e
            H = H + w * J.T @ J
            b = b - w * J.T @ e
        c2s = w_str * es * es
        x = x + np.linalg.solve(H, b)
    return x, errs[0], errs[1], c2s


# ---------------------------------------------------------------- keyframe association pipeline
# Independent NumPy restatement of renderView / searchCorrespondence / checkMapAssociation: matrix algebra
# with numpy.linalg (eigvalsh, inv, det, argsort), no shared code with oracle/ or the CUDA kernels.
W_IMG, H_IMG = synth.W, synth.H


def pose_Rt(pose):
    q = np.asarray(pose[:4], float) / np.linalg.norm(pose[:4])
    if q[3] < 0:
        q = -q
    return synth.quat_to_rot(q), np.asarray(pose[4:], float)


def project_gaussian(mean, cov, pose):
    """(visible, mean2 (2), cov2 (2x2), depth) of one component."""
    R, t = pose_Rt(pose)
    pc = R @ mean + t
    if not pc[2] > 0:
        return False, None, None, pc[2]
    uv = project(pc, False)
    if not (0 <= uv[0] < W_IMG and 0 <= uv[1] < H_IMG):
        return False, None, None, pc[2]
    J = dproj(pc, False)
    return True, uv, J @ R @ cov.reshape(3, 3) @ R.T @ J.T, pc[2]


def bh(mu0, c0, mu1, c1):
    c = 0.5 * (c0 + c1)
    d = mu1 - mu0
    return d @ np.linalg.inv(c) @ d / 8.0 + 0.5 * np.log(np.linalg.det(c) / np.sqrt(np.linalg.det(c0) * np.linalg.det(c1)))


def render_view(comps, pose):
    """list of dict(id, mean, cov, depth), depth descending."""
    R, t = pose_Rt(pose)
    cam_pos = -R.T @ t
    kept = []
    for k in range(len(comps)):
        c = comps[k]
        if c["is_degenerated"]:
            po = c["mean"] - cam_pos
            if abs(po @ c["axis"].reshape(3, 3)[:, 0]) / np.linalg.norm(po) < np.cos(np.deg2rad(78.0)):
                continue
        vis, m2, c2, depth = project_gaussian(c["mean"], c["cov"], pose)
        if not vis:
            continue
        ev = np.linalg.eigvalsh(0.5 * (c2 + c2.T))
        if ev[0] < 4.0 and ev[1] < 4.0:
            continue
        g = dict(id=k, mean=m2, cov=c2, depth=depth)
        if kept:
            d = np.array([bh(h["mean"], h["cov"], m2, c2) for h in kept])
            j = int(np.argmin(d))
            if d[j] < 0.8:
                if depth < kept[j]["depth"]:
                    kept[j] = g
                continue
        kept.append(g)
    order = np.argsort([-g["depth"] for g in kept], kind="stable")
    return [kept[i] for i in order]


def search_correspondence(view, uv, num=5):
    """N x num positions in `view` (-1 padded)."""
    uv = np.asarray(uv, float).reshape(-1, 2)
    out = np.full((len(uv), num), -1, np.int32)
    if not view:
        return out
    means = np.array([g["mean"] for g in view])
    for i, p in enumerate(uv):
        d = ((p - means) ** 2).sum(1)
        nn = np.argsort(d, kind="stable")[:num]
        m = 0
        for j in nn:
            dl = p - view[j]["mean"]
            if dl @ np.linalg.inv(view[j]["cov"]) @ dl < 9.0:
                out[i, m] = j
                m += 1
    return out


def check_map_association(comps, nbr, pose, obs, octave, cand, pt3d, tri_lambda2=400.0,
                          tri_str_thresh=float(np.float32(0.0064))):
    """(component id or -1, point after the call)."""
    row_ptr, col = nbr[0], nbr[1]
    cand = [int(c) for c in cand if c >= 0]
    pt_init = np.array(pt3d, float)
    if not cand:
        return -1, pt_init
    R, t = pose_Rt(pose)
    obs = np.asarray(obs, float)
    pz = min((R @ pt_init + t)[2], 1.0)
    w_str = tri_lambda2 * pz * pz

    def solve(k):
        c = comps[k]
        x, c2p, c2s = optimize_point(pt_init, obs, octave, R, t, c["mean"], c["axis"].reshape(3, 3)[:, 0], w_str)
        ok = not (c2p > 7.815) and not (c2s > tri_str_thresh * tri_lambda2)
        return ok, x, c2p

    def chi2(k, x):
        d = x - comps[k]["mean"]
        return d @ comps[k]["cov_inv"].reshape(3, 3) @ d

    best, best_v, best_x = -1, np.inf, None
    for k in cand:
        ok, x, c2p = solve(k)
        if ok and c2p < best_v:
            best, best_v, best_x = k, c2p, x
    if best >= 0:
        ll, s = chi2(best, best_x), best
        for e in range(row_ptr[best], row_ptr[best + 1]):
            ln = chi2(col[e], best_x)
            if ln < ll:
                ll, s = ln, int(col[e])
        if s != best:
            ok, x, _ = solve(s)
            if ok:
                best_x = x
            else:
                s, ll = best, chi2(best, best_x)
        if ll > 9.0:
            return -1, pt_init
        return s, best_x
    means = np.array([c["mean"] for c in comps]) if not isinstance(comps, np.ndarray) else comps["mean"]
    nn = int(np.argmin(((pt_init - means) ** 2).sum(1)))
    if comps[nn]["is_degenerated"]:
        ok, x, _ = solve(nn)
        if ok:
            return -1, x
    return -1, pt_init


# ---------------------------------------------------------------- local bundle adjustment (dense, no Schur)
def se3_log_np(R, t):
    """log of (R, t) -> [omega; upsilon] via scipy-free closed form (Rodrigues)."""
    c = 0.5 * (np.trace(R) - 1)
    w = 0.5 * np.array([R[2, 1] - R[1, 2], R[0, 2] - R[2, 0], R[1, 0] - R[0, 1]])
    s = np.linalg.norm(w)
    th = np.arctan2(s, c)
    om = w if th < 1e-10 else w * (th / s)
    Om = skew(om)
    if th < 1e-5:
        Vi = np.eye(3) - 0.5 * Om + Om @ Om / 12
    else:
        Vi = np.eye(3) - 0.5 * Om + (1 - th / (2 * np.tan(th / 2))) / th ** 2 * Om @ Om
    return np.concatenate([om, Vi @ t])


def local_ba_dense(comps, pose, fixed, pts, assoc, obs_pt, obs_kf, obs, octave, prior_idx=-1, ba_lambda2=400.0,
                   tri_str_thresh=float(np.float32(0.0064))):
    """Independent NumPy restatement of Localization::jointOptimization (localization_opt.cpp:456-925): the FULL
    (6 nf + 3 L) normal equations are assembled and solved with numpy.linalg (no Schur complement), g2o's
    Levenberg-Marquardt schedule (SURVEY Appx B.7), stages 5 / 5 / 40 with the reference's edge gating.
    Returns dict(R list, t list, pts, obs_outlier, assoc_cleared, iters)."""
    P, L, E = len(pose), len(pts), len(obs_pt)
    Rs, ts = [], []
    for p in pose:
        R, t = pose_Rt(p)
        Rs.append(R); ts.append(t)
    x = np.array(pts, float)
    fidx = np.full(P, -1)
    free = [i for i in range(P) if not fixed[i]]
    fidx[free] = np.arange(len(free))
    nf = len(free)
    dim = 6 * nf + 3 * L
    st = [bool(obs[e][2] >= 0) for e in range(E)]
    w = [s2inv(octave[e]) for e in range(E)]
    ob = np.asarray(obs, np.float64)
    dl = {True: float(np.float32(np.sqrt(7.815))), False: float(np.float32(np.sqrt(5.991)))}
    level = np.zeros(E, int)
    err = [np.zeros(3) for _ in range(E)]
    stype = [0 if assoc[l] < 0 else (1 if comps[assoc[l]]["is_degenerated"] else 2) for l in range(L)]
    slevel = np.zeros(L, int)
    state = dict(robust=True)
    if prior_idx >= 0:
        Rm, tm = Rs[prior_idx].copy(), ts[prior_idx].copy()
        sr = np.deg2rad(2.0)
        Om_prior = np.diag([1 / sr ** 2] * 3 + [1 / 0.01 ** 2] * 3)

    def str_resid(l, xl):
        c = comps[assoc[l]]
        d = xl - c["mean"]
        if stype[l] == 1:
            n = c["axis"].reshape(3, 3)[:, 0]
            return np.array([n @ d]), n[None, :], np.array([[ba_lambda2]])
        Lm = c["sqrt_info"].reshape(3, 3)
        return Lm.T @ d, Lm.T, np.eye(3)

    def prior_resid(R, t):
        dR, dt = Rm.T @ R, Rm.T @ (t - tm)
        return se3_log_np(dR, dt)

    def active(Rs_, ts_, x_):
        chi = 0.0
        for e in range(E):
            if level[e]:
                continue
            i, l = obs_kf[e], obs_pt[e]
            m = 3 if st[e] else 2
            r = ob[e, :m] - project(Rs_[i] @ x_[l] + ts_[i], st[e])
            err[e] = r
            c2 = w[e] * (r @ r)
            chi += huber(c2, dl[st[e]])[0] if state["robust"] else c2
        for l in range(L):
            if stype[l] and not slevel[l]:
                r, _, Om = str_resid(l, x_[l])
                chi += r @ Om @ r
        if prior_idx >= 0:
            r = prior_resid(Rs_[prior_idx], ts_[prior_idx])
            chi += r @ Om_prior @ r
        return chi

    def build():
        H = np.zeros((dim, dim))
        b = np.zeros(dim)
        for e in range(E):
            if level[e]:
                continue
            i, l = obs_kf[e], obs_pt[e]
            pc = Rs[i] @ x[l] + ts[i]
            m = 3 if st[e] else 2
            r = ob[e, :m] - project(pc, st[e])
            we = w[e]
            if state["robust"]:
                we *= huber(w[e] * (r @ r), dl[st[e]])[1]
            Pm = dproj(pc, st[e])
            cols, Js = [], []
            if fidx[i] >= 0:
                Js.append(-Pm @ np.hstack([-skew(pc), np.eye(3)]))
                cols.append(6 * fidx[i] + np.arange(6))
            Js.append(-Pm @ Rs[i])
            cols.append(6 * nf + 3 * l + np.arange(3))
            J = np.hstack(Js)
            c = np.concatenate(cols)
            H[np.ix_(c, c)] += we * J.T @ J
            b[c] -= we * J.T @ r
        for l in range(L):
            if stype[l] and not slevel[l]:
                r, J, Om = str_resid(l, x[l])
                c = 6 * nf + 3 * l + np.arange(3)
                H[np.ix_(c, c)] += J.T @ Om @ J
                b[c] -= J.T @ Om @ r
        if prior_idx >= 0 and fidx[prior_idx] >= 0:
            # numerical Jacobian of log(M^-1 exp(d) T) wrt d (independent of the analytic form in factors.cpp:31-53)
            R, t = Rs[prior_idx], ts[prior_idx]
            r0 = prior_resid(R, t)
            J = np.zeros((6, 6))
            for a in range(6):
                d = np.zeros(6); d[a] = 1e-6
                dRp, dtp = se3_exp(d)
                dRm, dtm = se3_exp(-d)
                J[:, a] = (prior_resid(dRp @ R, dRp @ t + dtp) - prior_resid(dRm @ R, dRm @ t + dtm)) / 2e-6
            c = 6 * fidx[prior_idx] + np.arange(6)
            H[np.ix_(c, c)] += J.T @ Om_prior @ J
            b[c] -= J.T @ Om_prior @ r0
        return H, b

    def optimize(n):
        nonlocal Rs, ts, x
        lam, ni, ok, it = 0.0, 2.0, True, 0
        while it < n and ok:
            cur = active(Rs, ts, x)
            H, b = build()
            if it == 0:
                lam, ni = 1e-5 * np.abs(np.diag(H)).max(), 2.0
            rho, q = 0.0, 0
            while True:
                try:
                    dx = np.linalg.solve(H + lam * np.eye(dim), b)
                    good = bool(np.all(np.isfinite(dx)))
                except np.linalg.LinAlgError:
                    good = False
                if good:
                    Rt, tt = list(Rs), list(ts)
                    for i in free:
                        dR, dt = se3_exp(dx[6 * fidx[i]: 6 * fidx[i] + 6])
                        Rt[i], tt[i] = dR @ Rs[i], dR @ ts[i] + dt
                    xt = x + dx[6 * nf:].reshape(L, 3)
                    tmp = active(Rt, tt, xt)
                else:
                    dx = np.zeros(dim)
                    tmp = np.finfo(float).max
                rho = (cur - tmp) / (dx @ (lam * dx + b) + 1e-3)
                if rho > 0 and np.isfinite(tmp) and good:
                    lam *= max(1 / 3, min(1 - (2 * rho - 1) ** 3, 2 / 3))
                    ni = 2.0
                    cur, Rs, ts, x = tmp, Rt, tt, xt
                else:
                    lam *= ni
                    ni *= 2
                    if not np.isfinite(lam):
                        break
                q += 1
                if not (rho < 0 and q < 10):
                    break
            if q == 10 or rho == 0 or not np.isfinite(lam):
                ok = False
            it += 1
        return it

    def str_chi2(l):
        r, _, Om = str_resid(l, x[l])
        return r @ Om @ r

    iters = [optimize(5)]
    for l in range(L):
        if stype[l] == 1 and str_chi2(l) > tri_str_thresh * ba_lambda2:
            slevel[l] = 1
    iters.append(optimize(5))

    def bad(e):
        i, l = obs_kf[e], obs_pt[e]
        return w[e] * (err[e] @ err[e]) > (7.815 if st[e] else 5.991) or not (Rs[i] @ x[l] + ts[i])[2] > 0

    for e in range(E):
        if bad(e):
            level[e] = 1
    state["robust"] = False
    iters.append(optimize(40))
    cleared = np.array([1 if stype[l] == 1 and str_chi2(l) > tri_str_thresh * ba_lambda2 else 0 for l in range(L)], np.uint8)
    outl = np.array([1 if bad(e) else 0 for e in range(E)], np.uint8)
    return dict(R=Rs, t=ts, pts=x, obs_outlier=outl, assoc_cleared=cleared, iters=iters)
