"""ctypes binding of the CPU ORACLE (oracle/gmmloc_oracle.cpp).

TEST INFRASTRUCTURE ONLY: imported by tests/, __graft_entry__.smoke() and bench.py's cpu_baseline /
--impl reference legs.  Nothing under gmmloc_b200/ imports this module.

The shared object is compiled on demand with the host's own `-march=native` (the reference builds with
-O3 -march=native, gmmloc/CMakeLists.txt:7) into oracle/_build/, keyed by a hash of the CPU flags so a
library built in the CPU-only container is never dlopen'ed on a different CPU on the GPU box.
"""
from __future__ import annotations

import ctypes as C
import hashlib
import os
import subprocess
from pathlib import Path

import numpy as np

_HERE = Path(__file__).resolve().parent
_SRC = [_HERE / "gmmloc_oracle.cpp", _HERE / "gmmloc_oracle.h"]


def _cpu_key() -> str:
    try:
        txt = Path("/proc/cpuinfo").read_text()
        flags = next((l for l in txt.splitlines() if l.startswith("flags")), "")
        model = next((l for l in txt.splitlines() if l.startswith("model name")), "")
    except OSError:
        flags = model = ""
    return hashlib.sha1((flags + model).encode()).hexdigest()[:10]


def build(force: bool = False) -> Path:
    out_dir = _HERE / "_build"
    out_dir.mkdir(exist_ok=True)
    so = out_dir / f"libgmmloc_oracle_{_cpu_key()}.so"
    newest = max(p.stat().st_mtime for p in _SRC)
    if force or not so.exists() or so.stat().st_mtime < newest:
        cxx = os.environ.get("CXX", "g++")
        cmd = [cxx, "-O3", "-march=native", "-std=c++17", "-fPIC", "-ffp-contract=off", "-fno-fast-math",
               "-pthread", "-shared", "-o", str(so) + ".tmp", str(_SRC[0])]
        subprocess.run(cmd, check=True, cwd=_HERE)
        os.replace(str(so) + ".tmp", so)
    return so


REFERENCE = Path("/root/reference")
_REF_SO = _HERE / "_ref" / "libref_nanoflann.so"


def build_ref(force: bool = False) -> Path | None:
    """oracle/_ref: the kNN harness compiled against the reference's own vendored nanoflann.hpp (the only part of the
    reference's path that builds without Eigen / g2o).  Built where /root/reference exists (this container); elsewhere
    (the GPU box) the prebuilt file that travelled with the snapshot is used.  Returns None when neither is possible."""
    hdr = REFERENCE / "gmmloc" / "include" / "gmmloc" / "utils" / "nanoflann.hpp"
    src = _HERE / "ref_nanoflann.cpp"
    if hdr.exists() and (force or not _REF_SO.exists() or _REF_SO.stat().st_mtime < src.stat().st_mtime):
        _REF_SO.parent.mkdir(exist_ok=True)
        cxx = os.environ.get("CXX", "g++")
        subprocess.run([cxx, "-O2", "-std=c++14", "-fPIC", "-shared", f"-I{REFERENCE / 'gmmloc' / 'include'}", "-o",
                        str(_REF_SO) + ".tmp", str(src)], check=True, cwd=_HERE)
        os.replace(str(_REF_SO) + ".tmp", _REF_SO)
    return _REF_SO if _REF_SO.exists() else None


_ref_lib = None


def ref_knn(pts, queries, k: int):
    """knnSearch of the reference's nanoflann kd-tree (leaf size 5, L2_Simple_Adaptor) over `pts` (n x 2 or n x 3):
    returns (idx nq x k, squared distances nq x k, counts nq); idx is -1 padded.  None when oracle/_ref is unavailable."""
    global _ref_lib
    so = build_ref()
    if so is None:
        return None
    if _ref_lib is None:
        _ref_lib = C.CDLL(str(so))
    pts = np.ascontiguousarray(pts, np.float64)
    q = np.ascontiguousarray(queries, np.float64).reshape(-1, pts.shape[1])
    fn = {2: _ref_lib.ref_knn2d, 3: _ref_lib.ref_knn3d}[pts.shape[1]]
    idx = np.zeros((len(q), k), np.int64)
    dist = np.zeros((len(q), k))
    cnt = np.zeros(len(q), np.int32)
    fn(_p(pts), C.c_int64(len(pts)), _p(q), C.c_int64(len(q)), C.c_int(k), _p(idx), _p(dist), _p(cnt))
    return idx, dist, cnt


class Comp(C.Structure):
    _fields_ = [("mean", C.c_double * 3), ("cov", C.c_double * 9), ("cov_inv", C.c_double * 9),
                ("det", C.c_double), ("scale", C.c_double * 3), ("axis", C.c_double * 9),
                ("sqrt_info", C.c_double * 9), ("is_degenerated", C.c_int32), ("is_salient", C.c_int32)]


COMP_DTYPE = np.dtype([("mean", "f8", 3), ("cov", "f8", 9), ("cov_inv", "f8", 9), ("det", "f8"),
                       ("scale", "f8", 3), ("axis", "f8", 9), ("sqrt_info", "f8", 9),
                       ("is_degenerated", "i4"), ("is_salient", "i4")], align=True)
assert COMP_DTYPE.itemsize == C.sizeof(Comp)


class Camera(C.Structure):
    _fields_ = [("fx", C.c_double), ("fy", C.c_double), ("cx", C.c_double), ("cy", C.c_double),
                ("bf", C.c_double), ("width", C.c_int32), ("height", C.c_int32)]


class Params(C.Structure):
    _fields_ = [("tri_lambda2", C.c_double), ("ba_lambda2", C.c_double), ("tri_str_thresh", C.c_double),
                ("tri_check_str_chi2", C.c_int32), ("mdist2_gate", C.c_double), ("chi2_mono", C.c_double),
                ("chi2_stereo", C.c_double), ("gn_iters", C.c_int32), ("robust", C.c_int32),
                ("factor_mode", C.c_int32)]


_lib = None


def lib():
    global _lib
    if _lib is None:
        L = C.CDLL(str(build()))
        L.oracle_chi2.restype = C.c_double
        L.oracle_bh3.restype = C.c_double
        L.oracle_sigma2_inv.restype = C.c_double
        L.oracle_neighbors.restype = C.c_int64
        L.oracle_optimize_point.restype = C.c_int
        L.oracle_optimize_pose_lm.restype = C.c_int
        L.oracle_optimize_triangulation.restype = C.c_int
        L.oracle_gmm_read.restype = C.c_int
        L.oracle_gmm_write.restype = C.c_int
        _lib = L
    return _lib


def _p(a, t=None):
    if a is None:
        return None
    return a.ctypes.data_as(C.c_void_p)


def default_params(**kw) -> Params:
    p = Params()
    lib().oracle_default_params(C.byref(p))
    for k, v in kw.items():
        setattr(p, k, v)
    return p


def v1_camera() -> Camera:
    # gmmloc_ros/cfg/v1.yaml:5-17
    return Camera(435.2046959714599, 435.2046959714599, 367.4517211914062, 252.2008514404297,
                  47.90639384423901, 752, 480)


def map_build(mean, cov, deg_eps=1e-4) -> np.ndarray:
    mean = np.ascontiguousarray(mean, np.float64).reshape(-1, 3)
    cov = np.ascontiguousarray(cov, np.float64).reshape(-1, 9)
    K = mean.shape[0]
    out = np.zeros(K, COMP_DTYPE)
    lib().oracle_map_build(_p(mean), _p(cov), C.c_int(K), C.c_double(deg_eps), _p(out))
    return out


def associate(comps, pts, gate=0.0, nthreads=1):
    pts = np.ascontiguousarray(pts, np.float32).reshape(-1, 3)
    M = pts.shape[0]
    assign = np.empty(M, np.int32)
    d2 = np.empty(M, np.float64)
    lib().oracle_associate(_p(comps), C.c_int(len(comps)), _p(pts), C.c_int64(M), C.c_double(gate),
                           _p(assign), _p(d2), C.c_int(nthreads))
    return assign, d2


def chi2(comps, k, x):
    x = np.ascontiguousarray(x, np.float64)
    return lib().oracle_chi2(C.c_void_p(comps.ctypes.data + int(k) * COMP_DTYPE.itemsize), _p(x))


def bh3(comps, i, j):
    base = comps.ctypes.data
    return lib().oracle_bh3(C.c_void_p(base + int(i) * COMP_DTYPE.itemsize),
                            C.c_void_p(base + int(j) * COMP_DTYPE.itemsize))


def neighbors(comps, thresh):
    K = len(comps)
    row_ptr = np.zeros(K + 1, np.int64)
    n = lib().oracle_neighbors(_p(comps), C.c_int(K), C.c_double(thresh), _p(row_ptr), None, None, C.c_int64(0))
    col = np.empty(n, np.int32)
    dist = np.empty(n, np.float64)
    lib().oracle_neighbors(_p(comps), C.c_int(K), C.c_double(thresh), _p(row_ptr), _p(col), _p(dist), C.c_int64(n))
    return row_ptr, col, dist


def optimize_point(comps, k, x0, obs, octave, pose, proj_z2, cam=None, prm=None, iters=5):
    cam = cam or v1_camera()
    prm = prm or default_params()
    x0 = np.ascontiguousarray(x0, np.float64)
    obs = np.ascontiguousarray(obs, np.float64)
    pose = np.ascontiguousarray(pose, np.float64)
    est = np.zeros(3)
    c2p, c2s = C.c_double(), C.c_double()
    res = lib().oracle_optimize_point(_p(x0), _p(obs), C.c_int(int(octave)), _p(pose),
                                      C.c_void_p(comps.ctypes.data + int(k) * COMP_DTYPE.itemsize),
                                      C.c_double(proj_z2), C.byref(cam), C.byref(prm), C.c_int(iters), _p(est),
                                      C.byref(c2p), C.byref(c2s))
    return bool(res), est, c2p.value, c2s.value


def frame_solve(comps, pose0, pts_w, obs, octave, cam=None, prm=None, assign_in=None, want_trace=False):
    cam = cam or v1_camera()
    prm = prm or default_params()
    pose0 = np.ascontiguousarray(pose0, np.float64)
    pts_w = np.ascontiguousarray(pts_w, np.float32).reshape(-1, 3)
    obs = np.ascontiguousarray(obs, np.float32).reshape(-1, 3)
    octave = np.ascontiguousarray(octave, np.uint8)
    N = pts_w.shape[0]
    pose_out = np.zeros(7)
    pts_out = np.zeros((N, 3))
    assign = np.zeros(N, np.int32)
    outlier = np.zeros(N, np.uint8)
    status = C.c_int32(0)
    trace = np.zeros((prm.gn_iters, 8)) if want_trace else None  # dxi (6) | max point step | its index
    ai = None if assign_in is None else np.ascontiguousarray(assign_in, np.int32)
    lib().oracle_frame_solve(_p(comps), C.c_int(len(comps)), C.byref(cam), C.byref(prm), _p(pose0), _p(pts_w),
                             _p(obs), _p(octave), C.c_int(N), _p(ai), _p(pose_out), _p(pts_out), _p(assign),
                             _p(outlier), C.byref(status), _p(trace), None)
    out = dict(pose=pose_out, pts=pts_out, assign=assign, outlier=outlier, status=status.value)
    if want_trace:
        out["trace"] = trace
    return out


def frames_solve(comps, pose0, pts_w, obs, octave, cam=None, prm=None, nthreads=1, want_pts=False, want_conv=False):
    cam = cam or v1_camera()
    prm = prm or default_params()
    pose0 = np.ascontiguousarray(pose0, np.float64).reshape(-1, 7)
    B = pose0.shape[0]
    pts_w = np.ascontiguousarray(pts_w, np.float32).reshape(B, -1, 3)
    N = pts_w.shape[1]
    obs = np.ascontiguousarray(obs, np.float32).reshape(B, N, 3)
    octave = np.ascontiguousarray(octave, np.uint8).reshape(B, N)
    pose_out = np.zeros((B, 7))
    pts_out = np.zeros((B, N, 3)) if want_pts else None
    assign = np.zeros((B, N), np.int32)
    outlier = np.zeros((B, N), np.uint8)
    status = np.zeros(B, np.int32)
    conv = np.zeros((B, 3)) if want_conv else None
    pt_last = np.zeros((B, N)) if want_conv else None
    lib().oracle_frames_solve_diag(_p(comps), C.c_int(len(comps)), C.byref(cam), C.byref(prm), _p(pose0), _p(pts_w),
                                   _p(obs), _p(octave), C.c_int(B), C.c_int(N), _p(pose_out), _p(pts_out), _p(assign),
                                   _p(outlier), _p(status), C.c_int(nthreads), _p(conv), _p(pt_last))
    out = dict(pose=pose_out, pts=pts_out, assign=assign, outlier=outlier, status=status)
    if want_conv:
        # conv: |dxi|_inf of the last iteration | largest pose step that exceeded its predecessor | largest last point step
        out["conv"], out["pt_last"] = conv, pt_last
    return out


def optimize_triangulation(comps, k, x0, obs1, obs2, octave1, pose1, pose2, cam=None, prm=None, iters=20):
    """One candidate of Localization::optimizeTriangulationVec: (res, pt_est, err1, err2, chi2_str)."""
    cam = cam or v1_camera()
    prm = prm or default_params()
    a = [np.ascontiguousarray(v, np.float64) for v in (x0, obs1, obs2, pose1, pose2)]
    est = np.zeros(3)
    e1, e2, cs = C.c_double(), C.c_double(), C.c_double()
    res = lib().oracle_optimize_triangulation(_p(a[0]), _p(a[1]), _p(a[2]), C.c_int(int(octave1)), _p(a[3]), _p(a[4]),
                                              C.c_void_p(comps.ctypes.data + int(k) * COMP_DTYPE.itemsize), C.byref(cam),
                                              C.byref(prm), C.c_int(iters), _p(est), C.byref(e1), C.byref(e2), C.byref(cs))
    return bool(res), est, e1.value, e2.value, cs.value


def optimize_pose_lm(pose0, pts_w, obs, octave, valid=None, cam=None):
    """Tracking::optimizeCurrentPose (4-round LM): returns (n_inliers, pose, is_outlier)."""
    cam = cam or v1_camera()
    pose0 = np.ascontiguousarray(pose0, np.float64)
    pts_w = np.ascontiguousarray(pts_w, np.float32).reshape(-1, 3)
    N = pts_w.shape[0]
    obs = np.ascontiguousarray(obs, np.float32).reshape(N, 3)
    octave = np.ascontiguousarray(octave, np.uint8)
    v = None if valid is None else np.ascontiguousarray(valid, np.uint8)
    pose = np.zeros(7)
    outl = np.zeros(N, np.uint8)
    n = lib().oracle_optimize_pose_lm(C.byref(cam), _p(pose0), _p(pts_w), _p(obs), _p(octave), _p(v), C.c_int(N), _p(pose), _p(outl))
    return n, pose, outl


def sigma2_inv(octave: int) -> float:
    return lib().oracle_sigma2_inv(C.c_int(int(octave)))


def se3_oplus(xi, pose):
    xi = np.ascontiguousarray(xi, np.float64)
    pose = np.ascontiguousarray(pose, np.float64)
    out = np.zeros(7)
    lib().oracle_se3_oplus(_p(xi), _p(pose), _p(out))
    return out


def gmm_read(path):
    K = lib().oracle_gmm_read(str(path).encode(), None, None, None, C.c_int(0))
    if K < 0:
        raise IOError(f"oracle_gmm_read({path}) failed: {K}")
    mean = np.zeros((K, 3))
    cov = np.zeros((K, 9))
    flags = np.zeros((K, 2), np.uint8)
    lib().oracle_gmm_read(str(path).encode(), _p(mean), _p(cov), _p(flags), C.c_int(K))
    return mean, cov, flags


def gmm_write(path, mean, cov, flags=None):
    mean = np.ascontiguousarray(mean, np.float64).reshape(-1, 3)
    cov = np.ascontiguousarray(cov, np.float64).reshape(-1, 9)
    fl = None if flags is None else np.ascontiguousarray(flags, np.uint8).reshape(-1, 2)
    rc = lib().oracle_gmm_write(str(path).encode(), _p(mean), _p(cov), _p(fl), C.c_int(mean.shape[0]))
    if rc != 0:
        raise IOError(f"oracle_gmm_write failed: {rc}")


# ---- keyframe association pipeline (renderView / searchCorrespondence / checkMapAssociation) -------------
COMP2D_DTYPE = np.dtype([("id", "i4"), ("pad_", "i4"), ("mean", "f8", 2), ("cov", "f8", 4), ("cov_inv", "f8", 4),
                         ("det", "f8"), ("scale", "f8", 2), ("depth", "f8")], align=True)


def render_view(comps, pose, cam=None) -> np.ndarray:
    """GMM::renderView: the depth-sorted (descending) list of 2-D components of one view."""
    cam = cam or v1_camera()
    pose = np.ascontiguousarray(pose, np.float64)
    out = np.zeros(len(comps), COMP2D_DTYPE)
    L = lib()
    L.oracle_render_view.restype = C.c_int
    V = L.oracle_render_view(_p(comps), C.c_int(len(comps)), C.byref(cam), _p(pose), _p(out))
    return out[:V].copy()


def project_gaussian(comps, k, pose, cam=None):
    cam = cam or v1_camera()
    pose = np.ascontiguousarray(pose, np.float64)
    out = np.zeros(1, COMP2D_DTYPE)
    L = lib()
    L.oracle_project_gaussian.restype = C.c_int
    ok = L.oracle_project_gaussian(C.c_void_p(comps.ctypes.data + int(k) * COMP_DTYPE.itemsize), C.c_int(int(k)),
                                   C.byref(cam), _p(pose), _p(out))
    return (out[0] if ok else None)


def bh2(view, i, j) -> float:
    L = lib()
    L.oracle_bh2.restype = C.c_double
    base = view.ctypes.data
    return L.oracle_bh2(C.c_void_p(base + int(i) * COMP2D_DTYPE.itemsize), C.c_void_p(base + int(j) * COMP2D_DTYPE.itemsize))


def search_correspondence(view, uv, num=5) -> np.ndarray:
    """GMM::searchCorrespondence: N x num positions in `view` (-1 padded)."""
    view = np.ascontiguousarray(view)
    uv = np.ascontiguousarray(uv, np.float64).reshape(-1, 2)
    cand = np.full((uv.shape[0], num), -1, np.int32)
    lib().oracle_search_correspondence(_p(view), C.c_int(len(view)), _p(uv), C.c_int(uv.shape[0]), C.c_int(num), _p(cand))
    return cand


def query_point(comps, pt, num=5):
    """GMM::queryPoint: (returned index = Euclid-nearest mean, the num-NN list, min Mahalanobis d2 over it)."""
    pt = np.ascontiguousarray(pt, np.float64)
    knn = np.zeros(num, np.int32)
    md = C.c_double()
    L = lib()
    L.oracle_query_point.restype = C.c_int
    r = L.oracle_query_point(_p(comps), C.c_int(len(comps)), _p(pt), C.c_int(num), _p(knn), C.byref(md))
    return r, knn, md.value


def unproject3(pose, uv, depth, cam=None) -> np.ndarray:
    cam = cam or v1_camera()
    pose = np.ascontiguousarray(pose, np.float64)
    uv = np.ascontiguousarray(uv, np.float64)
    out = np.zeros(3)
    lib().oracle_unproject3(C.byref(cam), _p(pose), _p(uv), C.c_double(float(depth)), _p(out))
    return out


def check_map_association(comps, nbr, pose, obs, octave, cand, pt3d, cam=None, prm=None):
    """GMMLoc::checkMapAssociation for one feature: (component id or -1, pt3d after the call)."""
    cam = cam or v1_camera()
    prm = prm or default_params()
    row_ptr, col = np.ascontiguousarray(nbr[0], np.int64), np.ascontiguousarray(nbr[1], np.int32)
    pose = np.ascontiguousarray(pose, np.float64)
    obs = np.ascontiguousarray(obs, np.float64)
    cand = np.ascontiguousarray(cand, np.int32)
    pt = np.array(pt3d, np.float64)
    L = lib()
    L.oracle_check_map_association.restype = C.c_int
    k = L.oracle_check_map_association(_p(comps), C.c_int(len(comps)), _p(row_ptr), _p(col), C.byref(cam), C.byref(prm),
                                       _p(pose), _p(obs), C.c_int(int(octave)), _p(cand), C.c_int(len(cand)), _p(pt))
    return k, pt


def keyframe_associate(comps, nbr, pose, obs, depth, octave, num=5, cam=None, prm=None, nthreads=1):
    """associateMapElements + per-feature checkMapAssociation of one keyframe."""
    cam = cam or v1_camera()
    prm = prm or default_params()
    row_ptr, col = np.ascontiguousarray(nbr[0], np.int64), np.ascontiguousarray(nbr[1], np.int32)
    pose = np.ascontiguousarray(pose, np.float64)
    obs = np.ascontiguousarray(obs, np.float32).reshape(-1, 3)
    N = obs.shape[0]
    depth = np.ascontiguousarray(depth, np.float32)
    octave = np.ascontiguousarray(octave, np.uint8)
    assoc = np.zeros(N, np.int32)
    pt3d = np.zeros((N, 3))
    cand = np.zeros((N, num), np.int32)
    nview = C.c_int32(0)
    lib().oracle_keyframe_associate(_p(comps), C.c_int(len(comps)), _p(row_ptr), _p(col), C.byref(cam), C.byref(prm),
                                    _p(pose), _p(obs), _p(depth), _p(octave), C.c_int(N), C.c_int(num), _p(assoc),
                                    _p(pt3d), _p(cand), C.byref(nview), C.c_int(nthreads))
    return dict(assoc=assoc, pt3d=pt3d, cand=cand, nview=nview.value)


def local_ba(comps, pose, fixed, pts, assoc, obs_pt, obs_kf, obs, octave, prior_idx=-1, cam=None, prm=None, stop=0):
    """Localization::jointOptimization (3-stage LM local BA): dict(pose, pts, obs_outlier, assoc_cleared, iters)."""
    cam = cam or v1_camera()
    prm = prm or default_params()
    pose = np.array(pose, np.float64).reshape(-1, 7).copy()
    pts = np.array(pts, np.float64).reshape(-1, 3).copy()
    fixed = np.ascontiguousarray(fixed, np.uint8)
    assoc = np.ascontiguousarray(assoc, np.int32)
    obs_pt = np.ascontiguousarray(obs_pt, np.int32)
    obs_kf = np.ascontiguousarray(obs_kf, np.int32)
    obs = np.ascontiguousarray(obs, np.float32).reshape(-1, 3)
    octave = np.ascontiguousarray(octave, np.uint8)
    E, L, P = len(obs_pt), len(pts), len(pose)
    outl = np.zeros(E, np.uint8)
    cleared = np.zeros(L, np.uint8)
    iters = np.zeros(3, np.int32)
    fn = lib().oracle_local_ba_stop
    fn.restype = C.c_int
    rc = fn(_p(comps), C.c_int(len(comps)), C.byref(cam), C.byref(prm), C.c_int(P), _p(pose), _p(fixed), C.c_int(int(prior_idx)),
            C.c_int(L), _p(pts), _p(assoc), C.c_int(E), _p(obs_pt), _p(obs_kf), _p(obs), _p(octave), _p(outl), _p(cleared),
            _p(iters), C.c_int32(int(stop)))
    if rc != 0:
        raise ValueError("oracle_local_ba: invalid indices")
    return dict(pose=pose, pts=pts, obs_outlier=outl, assoc_cleared=cleared, iters=iters)


def se3_log(pose):
    pose = np.ascontiguousarray(pose, np.float64)
    out = np.zeros(6)
    lib().oracle_se3_log(_p(pose), _p(out))
    return out
