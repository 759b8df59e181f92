"""gmmloc_b200 — B200-native GMM map association + structure-constrained Gauss-Newton (GMMLoc hot path).

Host-side mirror of the reference's call surface over the C ABI (include/gmmloc_b200.h):

  GMM                     ≙ gmmloc::GMM + GMMUtility::loadGMMModel / saveGMMModel
                            (gmmloc/include/gmmloc/gmm/gaussian_mixture.h:98-170, gmm_utils.h:13-16)
  Context.associate       ≙ GaussianComponent::chi2 scanned over all components with the reference's
                            argmin convention (gaussian.cpp:65-70, gmmloc_opt.cpp:175-217,230)
  Context.optimizePoint   ≙ GMMLoc::optimizePoint, batched (gmmloc_opt.cpp:260-342)
  Context.frames_solve    ≙ single-pose Localization::jointOptimization / Tracking::optimizeCurrentPose
                            (localization_opt.cpp:456-925, tracking_opt.cpp:21-217)

Everything computes on the GPU through libgmmloc_b200.so; there is no CPU path in this package.
"""
from __future__ import annotations

import ctypes as C

import numpy as np

from . import capi
from .capi import (Camera, GmmlocError, Params, default_params, v1_camera, MODE_JOINT, MODE_POSE_ONLY,
                   MODE_POINTS_ONLY, FRAME_NUMERIC)

__all__ = ["GMM", "Context", "Camera", "Params", "GmmlocError", "default_params", "v1_camera", "MODE_JOINT",
           "MODE_POSE_ONLY", "MODE_POINTS_ONLY", "FRAME_NUMERIC", "version", "sigma2_inv"]


def version() -> str:
    return capi.lib().gmmloc_version().decode()


def sigma2_inv(octave: int) -> float:
    """frame::sigma2_inv[octave] (init_config.hpp:63-78)."""
    return float(capi.lib().gmmloc_sigma2_inv(int(octave)))


def frames_solve_chunks(B: int, sms: int = 148, chunk: int = 0, taper: int = 1, short_first: int = 0) -> list:
    """Chunk sizes of the host pipeline of gmmloc_frames_solve for a batch of B frames (gmmloc_frames_solve_chunks: pure
    host arithmetic, no device)."""
    cap = max(8, B)
    sizes = (C.c_int32 * cap)()
    n = capi.lib().gmmloc_frames_solve_chunks(int(B), int(sms), int(chunk), int(taper), int(short_first), sizes, cap)
    if n < 0:
        raise ValueError("gmmloc_frames_solve_chunks: bad argument")
    return [int(sizes[i]) for i in range(min(n, cap))]


def _ptr(a):
    return None if a is None else C.c_void_p(a.ctypes.data)


def _in(a, dtype, shape=None):
    a = np.ascontiguousarray(a, dtype=dtype)
    if shape is not None:
        a = a.reshape(shape)
    return a


class GMM:
    """Immutable GMM map resident on one GPU (shareable between contexts / host threads)."""

    def __init__(self, mean, cov, deg_eps: float = 1e-4, device: int = 0, grid: "capi.GridOptions | None" = None):
        mean = _in(mean, np.float64, (-1, 3))
        cov = _in(cov, np.float64, (-1, 9))
        if mean.shape[0] != cov.shape[0]:
            raise ValueError("mean / cov size mismatch")
        h = C.c_void_p()
        if grid is None:
            capi.check(capi.lib().gmmloc_map_create(_ptr(mean), _ptr(cov), mean.shape[0], float(deg_eps), int(device),
                                                    C.byref(h)))
        else:  # explicit association-grid tuning (A/B runs): gmmloc_map_create_ex
            capi.check(capi.lib().gmmloc_map_create_ex(_ptr(mean), _ptr(cov), mean.shape[0], float(deg_eps), int(device),
                                                       C.byref(grid), C.byref(h)))
        self._h = h

    @classmethod
    def load(cls, path, deg_eps: float = 1e-4, device: int = 0) -> "GMM":
        """GMMUtility::loadGMMModel (gmm_utils.cpp:9-67); raises GmmlocError(EIO) where the reference
        returns false."""
        h = C.c_void_p()
        capi.check(capi.lib().gmmloc_map_load_gmm(str(path).encode(), float(deg_eps), int(device), C.byref(h)))
        self = cls.__new__(cls)
        self._h = h
        return self

    def save(self, path) -> None:
        """GMMUtility::saveGMMModel (gmm_utils.cpp:69-119)."""
        capi.check(capi.lib().gmmloc_map_save_gmm(self._h, str(path).encode()))

    def countComponents(self) -> int:
        return int(capi.lib().gmmloc_map_size(self._h))

    __len__ = countComponents

    @property
    def device(self) -> int:
        return int(capi.lib().gmmloc_map_device(self._h))

    def getComponents(self) -> dict:
        """Per-component constants built on the device (≙ GaussianComponent public fields)."""
        K = self.countComponents()
        out = dict(mean=np.empty((K, 3)), cov=np.empty((K, 9)), cov_inv=np.empty((K, 9)), det=np.empty(K),
                   scale=np.empty((K, 3)), axis=np.empty((K, 9)), sqrt_info=np.empty((K, 9)),
                   is_degenerated=np.empty(K, np.uint8), is_salient=np.empty(K, np.uint8))
        capi.check(capi.lib().gmmloc_map_get_components(self._h, *[_ptr(out[k]) for k in (
            "mean", "cov", "cov_inv", "det", "scale", "axis", "sqrt_info", "is_degenerated", "is_salient")]))
        return out

    def neighbors(self, thresh: float = 2.5):
        """nbs_ of every component (GMM::GMM, gaussian_mixture.cpp:61-78) as CSR: (row_ptr K+1, col, dist)."""
        K = self.countComponents()
        row_ptr = np.zeros(K + 1, np.int64)
        n = C.c_int64(0)
        capi.check(capi.lib().gmmloc_map_neighbors(self._h, float(thresh), _ptr(row_ptr), None, None, 0, C.byref(n)))
        col = np.empty(n.value, np.int32)
        dist = np.empty(n.value, np.float64)
        if n.value:
            capi.check(capi.lib().gmmloc_map_neighbors(self._h, float(thresh), _ptr(row_ptr), _ptr(col), _ptr(dist), n.value,
                                                       C.byref(n)))
        return row_ptr, col, dist

    def buildNeighbors(self, thresh: float = 2.5) -> int:
        """Builds nbs_ once and keeps it on the device (needed by checkMapAssociation); returns the pair count."""
        n = C.c_int64(0)
        capi.check(capi.lib().gmmloc_map_build_neighbors(self._h, float(thresh), C.byref(n)))
        return int(n.value)

    def scan_bound(self):
        a, x = C.c_double(), C.c_double()
        capi.check(capi.lib().gmmloc_map_scan_bound(self._h, C.byref(a), C.byref(x)))
        return a.value, x.value

    def close(self):
        if getattr(self, "_h", None):
            capi.lib().gmmloc_map_destroy(self._h)
            self._h = None

    def __del__(self):
        try:
            self.close()
        except Exception:
            pass


class Context:
    """One CUDA stream + workspaces; use from one host thread at a time."""

    def __init__(self, device: int = 0):
        h = C.c_void_p()
        capi.check(capi.lib().gmmloc_ctx_create(int(device), C.byref(h)))
        self._h = h
        self.device = int(device)

    @property
    def stream(self) -> int:
        """cudaStream_t as an integer (wrap with torch.cuda.ExternalStream for event timing)."""
        return int(capi.lib().gmmloc_ctx_stream(self._h) or 0)

    def synchronize(self) -> None:
        capi.check(capi.lib().gmmloc_ctx_synchronize(self._h))

    def stats(self) -> dict:
        s = capi.Stats()
        capi.check(capi.lib().gmmloc_ctx_stats(self._h, C.byref(s)))
        return dict(points=s.points, flagged=s.flagged, kernel_launches=s.kernel_launches, flagged_brute=s.flagged_brute)

    def set_assoc_kernel(self, kind: int) -> None:
        """capi.ASSOC_AUTO / ASSOC_FFMA2 / ASSOC_TENSOR (tcgen05 pruning scan) / ASSOC_GRID (grid-indexed scan); identical results."""
        capi.check(capi.lib().gmmloc_ctx_set_assoc_kernel(self._h, int(kind)))

    def set_option(self, option: int, value: int) -> None:
        """capi.OPT_GN_PREP (-1 auto / 0 / 1), capi.OPT_E2E_CHUNK (frames, 0 auto), capi.OPT_E2E_SHORT_FIRST; identical results."""
        capi.check(capi.lib().gmmloc_ctx_set_option(self._h, int(option), int(value)))

    def queryPoint(self, gmm: "GMM", pts) -> dict:
        """Batched GMM::queryPoint (gaussian_mixture.cpp:545-576): nearest M (what the reference returns: the
        Euclid-nearest mean), knn M x 5, min_mdist2 M (the Mahalanobis minimum the reference computes and discards)."""
        pts = _in(pts, np.float64, (-1, 3))
        M = pts.shape[0]
        out = dict(nearest=np.empty(M, np.int32), knn=np.empty((M, 5), np.int32), min_mdist2=np.empty(M))
        capi.check(capi.lib().gmmloc_map_query_point(self._h, gmm._h, _ptr(pts), M, _ptr(out["nearest"]), _ptr(out["knn"]),
                                                     _ptr(out["min_mdist2"])))
        return out

    def set_gn_precision(self, precision: int) -> None:
        """capi.GN_STREAM (default), GN_WORLD, GN_MIXED (fp64 state / residuals, fp32 Jacobian algebra) or capi.GN_FP64."""
        capi.check(capi.lib().gmmloc_ctx_set_gn_precision(self._h, int(precision)))

    def last_gn_kernel(self) -> str:
        """Frame-GN kernel the last frames_solve ran: "fp64", "mixed", "world" or "stream" ("none" before the first call)."""
        k = capi.lib().gmmloc_ctx_last_gn_kernel(self._h)
        return ("fp64", "mixed", "world", "stream")[k] if k >= 0 else "none"

    def last_assoc_kernel(self) -> str:
        return ("ffma2", "tensor", "grid")[capi.lib().gmmloc_ctx_last_assoc_kernel(self._h)]

    def kernel_times(self) -> dict:
        """Device ms of the last call's kernels (association scan, exact re-rank, frame GN)."""
        ms = (C.c_float * 3)()
        capi.check(capi.lib().gmmloc_ctx_kernel_times(self._h, C.byref(ms)))
        return dict(scan_ms=ms[0], rerank_ms=ms[1], gn_ms=ms[2])

    # ---- host entry points (numpy in, numpy out) ----
    def associate(self, gmm: GMM, pts_w, gate: float = 0.0, want_d2: bool = True):
        pts = _in(pts_w, np.float32)
        if pts.ndim == 2:
            pts = pts[None]
        if pts.ndim != 3 or pts.shape[2] != 3:
            raise ValueError("pts_w must be B x N x 3 or N x 3")
        B, N = pts.shape[:2]
        assign = np.empty((B, N), np.int32)
        d2 = np.empty((B, N), np.float32) if want_d2 else None
        capi.check(capi.lib().gmmloc_associate(self._h, gmm._h, _ptr(pts), B, N, float(gate), _ptr(assign), _ptr(d2)))
        return assign, d2

    def frames_solve(self, gmm: GMM, pose0, pts_w, obs, octave, cam: Camera | None = None,
                     params: Params | None = None, want_pts: bool = True, want_outlier: bool = True) -> dict:
        cam = cam or v1_camera()
        params = params or default_params()
        pose0 = _in(pose0, np.float64, (-1, 7))
        B = pose0.shape[0]
        pts = _in(pts_w, np.float32).reshape(B, -1, 3)
        N = pts.shape[1]
        obs = _in(obs, np.float32, (B, N, 3))
        octave = _in(octave, np.uint8, (B, N))
        pose_out = np.empty((B, 7))
        pts_out = np.empty((B, N, 3), np.float32) if want_pts else None
        assign = np.empty((B, N), np.int32)
        outlier = np.empty((B, N), np.uint8) if want_outlier else None
        status = np.empty(B, np.int32)
        capi.check(capi.lib().gmmloc_frames_solve(self._h, gmm._h, C.byref(cam), C.byref(params), _ptr(pose0), _ptr(pts),
                                                  _ptr(obs), _ptr(octave), B, N, _ptr(pose_out), _ptr(pts_out),
                                                  _ptr(assign), _ptr(outlier), _ptr(status)))
        return dict(pose=pose_out, pts=pts_out, assign=assign, outlier=outlier, status=status)

    def frames_solve_into(self, gmm: GMM, cam: Camera, params: Params, pose0, pts_w, obs, octave, pose_out, assign,
                          status, pts_out=None, outlier=None) -> None:
        """gmmloc_frames_solve on caller-owned host arrays (e.g. page-locked), no allocation, no conversion."""
        B, N = pts_w.shape[0], pts_w.shape[1]
        for a, dt in ((pose0, np.float64), (pts_w, np.float32), (obs, np.float32), (octave, np.uint8),
                      (pose_out, np.float64), (assign, np.int32), (status, np.int32)):
            if a.dtype != dt or not a.flags.c_contiguous:
                raise ValueError("frames_solve_into needs C-contiguous arrays of the ABI dtypes")
        capi.check(capi.lib().gmmloc_frames_solve(self._h, gmm._h, C.byref(cam), C.byref(params), _ptr(pose0), _ptr(pts_w),
                                                  _ptr(obs), _ptr(octave), B, N, _ptr(pose_out), _ptr(pts_out),
                                                  _ptr(assign), _ptr(outlier), _ptr(status)))

    def optimizeCurrentPose(self, pose0, pts_w, obs, octave, valid=None, cam: Camera | None = None) -> dict:
        """Batched Tracking::optimizeCurrentPose (exact 4-round LM schedule): pose B x 7, is_outlier B x N,
        inliers B."""
        cam = cam or v1_camera()
        pose0 = _in(pose0, np.float64, (-1, 7))
        B = pose0.shape[0]
        pts = _in(pts_w, np.float32).reshape(B, -1, 3)
        N = pts.shape[1]
        obs = _in(obs, np.float32, (B, N, 3))
        octave = _in(octave, np.uint8, (B, N))
        v = None if valid is None else _in(valid, np.uint8, (B, N))
        pose_out = np.empty((B, 7))
        outlier = np.empty((B, N), np.uint8)
        inliers = np.empty(B, np.int32)
        capi.check(capi.lib().gmmloc_poses_optimize(self._h, C.byref(cam), _ptr(pose0), _ptr(pts), _ptr(obs), _ptr(octave),
                                                    _ptr(v), B, N, _ptr(pose_out), _ptr(outlier), _ptr(inliers)))
        return dict(pose=pose_out, outlier=outlier, inliers=inliers)

    def optimizePoint(self, gmm: GMM, x0, obs, octave, pose, comp, proj_z2, cam: Camera | None = None,
                      params: Params | None = None, iters: int = 5) -> dict:
        """Batched GMMLoc::optimizePoint: returns StrOptStat fields (res, chi2_proj, chi2_str, pt_est)."""
        cam = cam or v1_camera()
        params = params or default_params()
        x0 = _in(x0, np.float64, (-1, 3))
        M = x0.shape[0]
        obs = _in(obs, np.float32, (M, 3))
        octave = _in(octave, np.uint8, (M,))
        pose = _in(np.broadcast_to(np.asarray(pose, np.float64).reshape(-1, 7), (M, 7)), np.float64)
        comp = _in(comp, np.int32, (M,))
        proj_z2 = _in(proj_z2, np.float64, (M,))
        pt_est = np.empty((M, 3))
        c2p = np.empty(M)
        c2s = np.empty(M)
        res = np.empty(M, np.uint8)
        capi.check(capi.lib().gmmloc_points_solve(self._h, gmm._h, C.byref(cam), C.byref(params), _ptr(x0), _ptr(obs),
                                                  _ptr(octave), _ptr(pose), _ptr(comp), _ptr(proj_z2), M, int(iters),
                                                  _ptr(pt_est), _ptr(c2p), _ptr(c2s), _ptr(res)))
        return dict(res=res.astype(bool), chi2_proj=c2p, chi2_str=c2s, pt_est=pt_est)

    def optimizeTriangulation(self, gmm: GMM, x0, obs1, obs2, octave1, pose1, pose2, comp, cam: Camera | None = None,
                              params: Params | None = None, iters: int = 20) -> dict:
        """Batched candidate solves of Localization::optimizeTriangulationVec (localization_opt.cpp:27-204):
        one problem per (match, candidate component)."""
        cam = cam or v1_camera()
        params = params or default_params()
        x0 = _in(x0, np.float64, (-1, 3))
        M = x0.shape[0]
        obs1 = _in(obs1, np.float32, (M, 3))
        obs2 = _in(obs2, np.float32, (M, 3))
        octave1 = _in(octave1, np.uint8, (M,))
        pose1 = _in(np.broadcast_to(np.asarray(pose1, np.float64).reshape(-1, 7), (M, 7)), np.float64)
        pose2 = _in(np.broadcast_to(np.asarray(pose2, np.float64).reshape(-1, 7), (M, 7)), np.float64)
        comp = _in(comp, np.int32, (M,))
        pt_est = np.empty((M, 3))
        e1, e2, cs = np.empty(M), np.empty(M), np.empty(M)
        res = np.empty(M, np.uint8)
        capi.check(capi.lib().gmmloc_triangulation_solve(self._h, gmm._h, C.byref(cam), C.byref(params), _ptr(x0), _ptr(obs1),
                                                         _ptr(obs2), _ptr(octave1), _ptr(pose1), _ptr(pose2), _ptr(comp), M,
                                                         int(iters), _ptr(pt_est), _ptr(e1), _ptr(e2), _ptr(cs), _ptr(res)))
        return dict(res=res.astype(bool), err1=e1, err2=e2, chi2_str=cs, pt_est=pt_est)

    def optimizeTriangulationVec(self, gmm: GMM, x3d, obs1, obs2, octave1, pose1, pose2, candidates, is_degenerated,
                                 **kw):
        """Localization::optimizeTriangulationVec for one match: tries every DEGENERATE candidate component
        (ascending index; the reference iterates an unordered_set, SURVEY Appx A9), keeps the accepted one with
        the smallest err1 + err2 (strict '<').  Returns (component index or -1, point)."""
        cand = sorted({int(c) for c in candidates if c >= 0 and is_degenerated[int(c)]})
        x3d = np.asarray(x3d, np.float64)
        if not cand:
            return -1, x3d
        M = len(cand)
        r = self.optimizeTriangulation(gmm, np.tile(x3d, (M, 1)), np.tile(np.asarray(obs1, np.float32), (M, 1)),
                                       np.tile(np.asarray(obs2, np.float32), (M, 1)), np.full(M, octave1, np.uint8), pose1,
                                       pose2, np.asarray(cand, np.int32), **kw)
        best, best_v = -1, np.finfo(np.float64).max
        for j in range(M):
            s = r["err1"][j] + r["err2"][j]
            if r["res"][j] and s < best_v:
                best, best_v = j, s
        if best < 0:
            return -1, x3d
        return cand[best], r["pt_est"][best]

    # ---- keyframe association pipeline ----
    def renderView(self, gmm: GMM, pose, cap: int | None = None, cam: Camera | None = None) -> dict:
        """Batched GMM::renderView (gaussian_mixture.cpp:271-371): per view the depth-sorted 2-D components
        (count B, ids / mean / cov / depth padded to cap).  The views stay on this context for
        searchCorrespondence."""
        cam = cam or v1_camera()
        pose = _in(pose, np.float64, (-1, 7))
        B = pose.shape[0]
        cap = min(int(cap or len(gmm)), len(gmm))
        out = dict(count=np.zeros(B, np.int32), status=np.zeros(B, np.int32), ids=np.full((B, cap), -1, np.int32),
                   mean=np.zeros((B, cap, 2)), cov=np.zeros((B, cap, 4)), depth=np.zeros((B, cap)))
        capi.check(capi.lib().gmmloc_render_views(self._h, gmm._h, C.byref(cam), _ptr(pose), B, cap, _ptr(out["count"]),
                                                  _ptr(out["status"]), _ptr(out["ids"]), _ptr(out["mean"]), _ptr(out["cov"]),
                                                  _ptr(out["depth"])))
        for b in range(B):
            out["ids"][b, out["count"][b]:] = -1
        return out

    def searchCorrespondence(self, obs, num: int = 5, want_pos: bool = False):
        """Batched GMM::searchCorrespondence on the views last rendered here: cand B x N x num (component ids)."""
        obs = _in(obs, np.float32)
        if obs.ndim == 2:
            obs = obs[None]
        B, N = obs.shape[:2]
        cand = np.empty((B, N, num), np.int32)
        pos = np.empty((B, N, num), np.int32) if want_pos else None
        capi.check(capi.lib().gmmloc_search_correspondence(self._h, _ptr(obs), B, N, int(num), _ptr(cand), _ptr(pos)))
        return (cand, pos) if want_pos else cand

    def checkMapAssociation(self, gmm: GMM, pose, obs, octave, cand, pt3d, cam: Camera | None = None,
                            params: Params | None = None):
        """Batched GMMLoc::checkMapAssociation: returns (assoc B x N, pt3d B x N x 3 after the call)."""
        cam = cam or v1_camera()
        params = params or default_params()
        pose = _in(pose, np.float64, (-1, 7))
        B = pose.shape[0]
        obs = _in(obs, np.float32).reshape(B, -1, 3)
        N = obs.shape[1]
        octave = _in(octave, np.uint8, (B, N))
        cand = _in(cand, np.int32).reshape(B, N, -1)
        num = cand.shape[2]
        pt = np.array(pt3d, np.float64).reshape(B, N, 3).copy()
        assoc = np.empty((B, N), np.int32)
        capi.check(capi.lib().gmmloc_check_map_association(self._h, gmm._h, C.byref(cam), C.byref(params), _ptr(pose),
                                                           _ptr(obs), _ptr(octave), _ptr(cand), B, N, num, _ptr(pt),
                                                           _ptr(assoc)))
        return assoc, pt

    def associateMapElements(self, gmm: GMM, pose, obs, depth, octave, num: int = 5, cap: int | None = None,
                             cam: Camera | None = None, params: Params | None = None) -> dict:
        """GMMLoc::associateMapElements + the per-feature association of createMapPointsFromStereo for B
        keyframes, chained on the device (render -> search -> unproject3 -> checkMapAssociation)."""
        cam = cam or v1_camera()
        params = params or default_params()
        pose = _in(pose, np.float64, (-1, 7))
        B = pose.shape[0]
        obs = _in(obs, np.float32)
        if obs.ndim != 3:
            obs = obs.reshape(B, -1, 3)
        N = obs.shape[1]
        depth = _in(depth, np.float32, (B, N))
        octave = _in(octave, np.uint8, (B, N))
        cap = min(int(cap or len(gmm)), len(gmm))
        out = dict(assoc=np.empty((B, N), np.int32), pt3d=np.empty((B, N, 3)), cand=np.empty((B, N, num), np.int32),
                   nview=np.empty(B, np.int32), vstatus=np.empty(B, np.int32))
        capi.check(capi.lib().gmmloc_keyframes_associate(self._h, gmm._h, C.byref(cam), C.byref(params), _ptr(pose), _ptr(obs),
                                                         _ptr(depth), _ptr(octave), B, N, int(num), cap, _ptr(out["assoc"]),
                                                         _ptr(out["pt3d"]), _ptr(out["cand"]), _ptr(out["nview"]),
                                                         _ptr(out["vstatus"])))
        return out

    def jointOptimization(self, gmm: GMM, windows, cam: Camera | None = None, params: Params | None = None, stop=None):
        """Batched Localization::jointOptimization (local BA, 3-stage LM; localization_opt.cpp:456-925).
        `windows`: list of dicts with pose P x 7, fixed P, pts L x 3, assoc L, obs_pt / obs_kf E (window-local,
        obs_pt non-decreasing), obs E x 3, octave E and optionally prior_idx.  Returns one dict per window:
        pose, pts, obs_outlier, assoc_cleared, iters (3), status.  `stop`: the reference's pbStopFlag — an int, or a
        one-element int32 numpy array another thread may set while the call runs (see gmmloc_local_ba)."""
        cam = cam or v1_camera()
        params = params or default_params()
        B = len(windows)
        cat = lambda k, dt, shape: (np.concatenate([_in(w[k], dt, shape) for w in windows]) if B else np.zeros((0,) + shape[1:], dt))
        pose = np.ascontiguousarray(cat("pose", np.float64, (-1, 7)))
        fixed = np.ascontiguousarray(cat("fixed", np.uint8, (-1,)))
        pts = np.ascontiguousarray(cat("pts", np.float64, (-1, 3)))
        assoc = np.ascontiguousarray(cat("assoc", np.int32, (-1,)))
        obs_pt = np.ascontiguousarray(cat("obs_pt", np.int32, (-1,)))
        obs_kf = np.ascontiguousarray(cat("obs_kf", np.int32, (-1,)))
        obs = np.ascontiguousarray(cat("obs", np.float32, (-1, 3)))
        octave = np.ascontiguousarray(cat("octave", np.uint8, (-1,)))
        ptr = lambda k: np.concatenate([[0], np.cumsum([len(np.asarray(w[k]).reshape(len(w[k]), -1)) for w in windows])]).astype(np.int32)
        pose_ptr, pt_ptr, obs_ptr = ptr("pose"), ptr("pts"), ptr("obs_pt")
        prior = np.array([int(w.get("prior_idx", -1)) for w in windows], np.int32)
        outl = np.zeros(len(obs_pt), np.uint8)
        cleared = np.zeros(len(pts), np.uint8)
        iters = np.zeros((B, 3), np.int32)
        status = np.zeros(B, np.int32)
        if stop is not None and not isinstance(stop, np.ndarray):
            stop = np.array([int(stop)], np.int32)
        capi.check(capi.lib().gmmloc_local_ba(self._h, gmm._h, C.byref(cam), C.byref(params), B, _ptr(pose_ptr), _ptr(pose),
                                              _ptr(fixed), _ptr(prior), _ptr(pt_ptr), _ptr(pts), _ptr(assoc), _ptr(obs_ptr),
                                              _ptr(obs_pt), _ptr(obs_kf), _ptr(obs), _ptr(octave), _ptr(outl), _ptr(cleared),
                                              _ptr(iters), _ptr(status), _ptr(stop)))
        out = []
        for b in range(B):
            out.append(dict(pose=pose[pose_ptr[b]:pose_ptr[b + 1]], pts=pts[pt_ptr[b]:pt_ptr[b + 1]],
                            obs_outlier=outl[obs_ptr[b]:obs_ptr[b + 1]], assoc_cleared=cleared[pt_ptr[b]:pt_ptr[b + 1]],
                            iters=iters[b], status=int(status[b])))
        return out

    # ---- device entry points (raw device pointers, enqueue only) ----
    def associate_async(self, gmm: GMM, d_pts: int, M: int, gate: float, d_assign: int, d_d2: int = 0) -> None:
        capi.check(capi.lib().gmmloc_associate_async(self._h, gmm._h, d_pts, int(M), float(gate), d_assign, d_d2 or None))

    def frames_solve_async(self, gmm: GMM, cam: Camera, params: Params, d_pose0: int, d_pts: int, d_obs: int,
                           d_octave: int, B: int, N: int, d_pose_out: int, d_pts_out: int, d_assign: int,
                           d_outlier: int, d_status: int) -> None:
        capi.check(capi.lib().gmmloc_frames_solve_async(self._h, gmm._h, C.byref(cam), C.byref(params), d_pose0, d_pts,
                                                        d_obs, d_octave, int(B), int(N), d_pose_out, d_pts_out or None,
                                                        d_assign, d_outlier or None, d_status))

    def close(self):
        if getattr(self, "_h", None):
            capi.lib().gmmloc_ctx_destroy(self._h)
            self._h = None

    def __del__(self):
        try:
            self.close()
        except Exception:
            pass
