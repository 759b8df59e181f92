// local_ba.cu — K9: Localization::jointOptimization (local bundle adjustment), one CTA — or one thread-block CLUSTER of
// up to 8 CTAs — per window, fp64.
//
// Replaces gmmloc/src/modules/localization_opt.cpp:456-925: g2o Levenberg-Marquardt over the free keyframe
// poses (VertexSE3Expmap) and all marginalised map points (VertexSBAPointXYZ) of a covisibility window with
//   EdgeSE3ProjectXYZ / EdgeStereoSE3ProjectXYZ  per observation (information sigma2_inv[octave], Huber
//                                                sqrt(5.991) / sqrt(7.815) until stage 3),
//   EdgePt2GaussianDeg / EdgePt2Gaussian         per associated point (factors.cpp:5-17,55-64; information
//                                                ba_lambda2 / I),
//   EdgeSE3QuatPrior                             on keyframe 0 when ba_first_as_prior (factors.cpp:19-53),
// BlockSolver_6_3 (Schur complement over the points), stages optimize(5) / optimize(5) / optimize(40) with the
// reference's edge gating in between (:774-786, :797-823) and its final outlier / association tests (:842-881).
//
// The reference builds a heap g2o graph per call and runs a sparse Cholesky; here every pass is a block-wide
// sweep over flat arrays with OUTPUT-STATIONARY reductions (no floating-point atomics => deterministic):
//   * per observation: residual, Jacobians and its five normal-equation pieces (pose-pose 21, pose rhs 6,
//     pose-point 18, point-point 6, point rhs 3) are written once per linearisation;
//   * per point: a thread sums the pieces of its own observations (CSR lists built on the host); per free pose: a
//     warp strides over the pose's observation list and folds with a fixed shuffle tree;
//   * Schur complement: per trial every observation first gets T = Hpl Hll^-1 and Hpl Hll^-1 bl; then one thread
//     per ROW of every lower 6x6 block of the reduced system walks the block's precomputed contribution list
//     (pairs of observations of one point, sorted by block on the host) accumulating T_a W_b^T;
//   * reduced system: left-looking dense Cholesky with the right-hand side carried as an extra row.
// LM bookkeeping (g2o OptimizationAlgorithmLevenberg, SURVEY Appx B.7) is evaluated redundantly by all threads
// from block-reduced values, so control flow stays uniform.
//
// Cluster variant (local_ba_kernel<kCl>, kCl = 2 / 4 / 8).  The localisation thread solves ONE window at a time
// (Localization::run -> jointOptimization, localization.cpp:51,102), and a single CTA walks its 8 MB of per-observation
// workspace at one SM's L2 rate (67 ms for 12 keyframes / 1 500 points, profiles/bench_ba_r01.jsonl).  With few windows
// in flight a window is spread over a cluster: every sweep strides over the threads of ALL its CTAs, the sweeps are
// separated by hardware cluster barriers (barrier.cluster, release / acquire at cluster scope, which also orders the
// global-memory workspaces between the CTAs), block-wide reductions exchange one partial per CTA through distributed
// shared memory and add them in rank order in every CTA (so all CTAs take the same branches, bit for bit), and the
// dense Cholesky of the small reduced system stays inside CTA 0.
#include <cooperative_groups.h>
#include <cstdio>
#include <math_constants.h>

#include "gn_common.cuh"

namespace gmmloc {

namespace {

constexpr int kBaT = 512;
constexpr int kBaW = kBaT / 32;
// per-observation workspace layout (doubles)
constexpr int oA = 0, oG = 21, oW = 27, oHl = 45, oBl = 51, oErr = 54, oT = 57, oGy = 75, kObsWs = 81;
// per-point workspace layout
constexpr int pH = 0, pB = 6, pHi = 9, pY = 15, pDx = 18, pTr = 21, kPtWs = 24;
// per-pose workspace layout
constexpr int qR = 0, qRt = 9, qTr = 18, kPoseWs = 25;
// per-free-pose workspace layout
constexpr int fH = 0, fB = 21, kFreeWs = 27;
constexpr int kFoldParts = 8;                          // warps that share one free pose's observation list
constexpr int kFreeStride = kFreeWs * (1 + kFoldParts);  // per free pose: Hpp / bp, then the partial sums of the fold

__device__ __forceinline__ int sym6(int r, int c) {  // packed upper-triangle index of a symmetric 6x6
  if (r > c) { const int t = r; r = c; c = t; }
  return r * 6 - r * (r - 1) / 2 + (c - r);
}

namespace cg = cooperative_groups;

#ifdef GMMLOC_BA_PROFILE
// phase timers of one window (debug builds: tools/build_variant.sh baprof -DGMMLOC_BA_PROFILE): nanoseconds per phase,
// accumulated by thread 0 of the team, printed by the launcher
__device__ unsigned long long g_ba_prof[16];
__device__ __forceinline__ unsigned long long ba_now() {
  unsigned long long t;
  asm volatile("mov.u64 %0, %globaltimer;" : "=l"(t));
  return t;
}
#define BA_T(i)                                 \
  if (gtid == 0) {                              \
    const unsigned long long t1__ = ba_now();   \
    g_ba_prof[i] += t1__ - ba_t0;               \
    ba_t0 = t1__;                               \
  }
#else
#define BA_T(i)
#endif

// The threads that solve one window: one CTA (kCl = 1) or a thread-block cluster of kCl CTAs.
template <int kCl>
struct Team {
  static constexpr int kThreads = kCl * kBaT, kWarps = kCl * kBaW;
  double* s;      // kBaW doubles of this CTA's shared memory (warp partials)
  double* s_tot;  // 2 doubles of this CTA's shared memory: its block totals, double-buffered across reductions
  int* s_int;     // 2 ints, same use
  int crank;      // rank of this CTA in its cluster
  int gtid;       // thread index within the team
  mutable int par = 0;

  __device__ __forceinline__ void sync() const {
    if (kCl == 1) __syncthreads();
    else cg::this_cluster().sync();
  }
  // combine one value per CTA in rank order (identical in every CTA); `mine` is uniform within the CTA
  template <class Op>
  __device__ __forceinline__ double across(double mine, Op op) const {
    if (kCl == 1) return mine;
    cg::cluster_group cl = cg::this_cluster();
    if (threadIdx.x == 0) s_tot[par] = mine;
    cl.sync();
    double t = *cl.map_shared_rank(s_tot + par, 0);
#pragma unroll
    for (int r = 1; r < kCl; ++r) t = op(t, *cl.map_shared_rank(s_tot + par, r));
    par ^= 1;  // the next reduction writes the other slot: nobody can still be reading it (see the header comment)
    return t;
  }
  __device__ __forceinline__ double sum(double v) const {
#pragma unroll
    for (int off = 16; off > 0; off >>= 1) v += __shfl_xor_sync(0xffffffffu, v, off);
    __syncthreads();
    if ((threadIdx.x & 31) == 0) s[threadIdx.x >> 5] = v;
    __syncthreads();
    double t = s[0];
#pragma unroll
    for (int w = 1; w < kBaW; ++w) t += s[w];
    return across(t, [](double a, double b) { return a + b; });
  }
  __device__ __forceinline__ double max(double v) const {
#pragma unroll
    for (int off = 16; off > 0; off >>= 1) v = fmax(v, __shfl_xor_sync(0xffffffffu, v, off));
    __syncthreads();
    if ((threadIdx.x & 31) == 0) s[threadIdx.x >> 5] = v;
    __syncthreads();
    double t = s[0];
#pragma unroll
    for (int w = 1; w < kBaW; ++w) t = fmax(t, s[w]);
    return across(t, [](double a, double b) { return fmax(a, b); });
  }
  __device__ __forceinline__ int all_ok(bool ok) const {  // team-wide AND; also a team barrier
    const int r = __syncthreads_and(ok ? 1 : 0);
    if (kCl == 1) return r;
    return across(r ? 1.0 : 0.0, [](double a, double b) { return fmin(a, b); }) != 0.0;
  }
  // one value decided by thread 0 of the team, handed to every CTA (each CTA reads its own copy after the barrier)
  __device__ __forceinline__ int broadcast_from_lead(int v) const {
    if (kCl == 1) {
      __syncthreads();
      if (threadIdx.x == 0) s_int[0] = v;
      __syncthreads();
      return s_int[0];
    }
    cg::cluster_group cl = cg::this_cluster();
    cl.sync();
    if (gtid == 0)
      for (int r = 0; r < kCl; ++r) *cl.map_shared_rank(s_int, r) = v;
    cl.sync();
    return s_int[0];
  }
};

// ---- SE3 helpers for the prior edge (g2o se3quat.h) ----
__device__ void quat_rotate_v(const double* q, const double* v, double* out) {
  double uv[3] = {q[1] * v[2] - q[2] * v[1], q[2] * v[0] - q[0] * v[2], q[0] * v[1] - q[1] * v[0]};
  for (int i = 0; i < 3; ++i) uv[i] += uv[i];
  const double c[3] = {q[1] * uv[2] - q[2] * uv[1], q[2] * uv[0] - q[0] * uv[2], q[0] * uv[1] - q[1] * uv[0]};
  for (int i = 0; i < 3; ++i) out[i] = (v[i] + q[3] * uv[i]) + c[i];
}
__device__ void se3_inverse_d(const double* T, double* out) {
  const double qc[4] = {-T[0], -T[1], -T[2], T[3]};
  const double tn[3] = {-T[4], -T[5], -T[6]};
  quat_rotate_v(qc, tn, out + 4);
  for (int i = 0; i < 4; ++i) out[i] = qc[i];
}
__device__ void se3_mul_d(const double* A, const double* B, double* out) {
  double rt[3], q[4];
  quat_rotate_v(A, B + 4, rt);
  const double ax = A[0], ay = A[1], az = A[2], aw = A[3], bx = B[0], by = B[1], bz = B[2], bw = B[3];
  q[3] = aw * bw - ax * bx - ay * by - az * bz;
  q[0] = aw * bx + ax * bw + ay * bz - az * by;
  q[1] = aw * by + ay * bw + az * bx - ax * bz;
  q[2] = aw * bz + az * bw + ax * by - ay * bx;
  quat_normalize(q);
  for (int i = 0; i < 4; ++i) out[i] = q[i];
  for (int i = 0; i < 3; ++i) out[4 + i] = A[4 + i] + rt[i];
}
__device__ void se3_log_d(const double* T, double* xi) {
  double R[9];
  quat_to_rot(T, R);
  const double d = 0.5 * (R[0] + R[4] + R[8] - 1);
  const double dR[3] = {R[7] - R[5], R[2] - R[6], R[3] - R[1]};
  double om[3], coef;
  if (fabs(d) > 0.99999) {
    for (int i = 0; i < 3; ++i) om[i] = 0.5 * dR[i];
    coef = 1. / 12.;
  } else {
    const double theta = acos(d);
    for (int i = 0; i < 3; ++i) om[i] = theta / (2 * sqrt(1 - d * d)) * dR[i];
    coef = (1 - theta / (2 * tan(theta / 2))) / (theta * theta);
  }
  const double Om[9] = {0, -om[2], om[1], om[2], 0, -om[0], -om[1], om[0], 0};
  const double* t = T + 4;
  for (int r = 0; r < 3; ++r) {
    double acc = 0;
    for (int c = 0; c < 3; ++c) {
      const double o2 = Om[r * 3] * Om[c] + Om[r * 3 + 1] * Om[3 + c] + Om[r * 3 + 2] * Om[6 + c];
      const double vi = ((r == c) ? 1.0 : 0.0) - 0.5 * Om[r * 3 + c] + coef * o2;
      acc += vi * t[c];
    }
    xi[3 + r] = acc;
  }
  for (int i = 0; i < 3; ++i) xi[i] = om[i];
}

struct Prob {
  BaProb d;
  // problem-local views
  double* pose;
  const int32_t* fidx;
  double* pts;
  const int32_t* assoc;
  const int32_t *obs_pt, *obs_kf;
  const float* obs;
  const uint8_t* octave;
  const int32_t *pt_ptr, *kf_ptr, *kf_obs, *blk_ptr;
  const int2* contrib;
  double *wo, *wp, *wq, *wf, *S, *rhs, *dxp, *diag;
  uint8_t *level, *slevel;
};

// residual of a reprojection edge at camera-frame point pc
__device__ __forceinline__ void edge_err(const FrameParamsDev& P, const double* pc, const float* ob, bool stereo, double* e) {
  const double x = pc[0], y = pc[1], z = pc[2];
  if (stereo) {
    const double invz = 1.0 / z;
    const double u = x * invz * P.fx + P.cx, v = y * invz * P.fy + P.cy;
    e[0] = (double)ob[0] - u;
    e[1] = (double)ob[1] - v;
    e[2] = (double)ob[2] - (u - P.bf * invz);
  } else {
    e[0] = (double)ob[0] - (x / z * P.fx + P.cx);
    e[1] = (double)ob[1] - (y / z * P.fy + P.cy);
    e[2] = 0.0;
  }
}

// structure edge of point l: error (1 or 3 rows), chi2; type 1 = EdgePt2GaussianDeg, 2 = EdgePt2Gaussian
__device__ __forceinline__ int str_type_of(const BaArgs& a, int k) { return k < 0 ? 0 : (a.is_deg[k] ? 1 : 2); }
__device__ __forceinline__ double str_error(const BaArgs& a, int k, int type, const double* x, double* e) {
  const double* mu = a.mean + (size_t)k * 3;
  const double d[3] = {x[0] - mu[0], x[1] - mu[1], x[2] - mu[2]};
  if (type == 1) {
    const double* ax = a.axis + (size_t)k * 9;
    e[0] = (ax[0] * d[0] + ax[3] * d[1]) + ax[6] * d[2];
    e[1] = e[2] = 0.0;
    return a.prm.ba_lambda2 * (e[0] * e[0]);
  }
  const double* L = a.sqrt_info + (size_t)k * 9;
  for (int r = 0; r < 3; ++r) e[r] = (L[r] * d[0] + L[3 + r] * d[1]) + L[6 + r] * d[2];
  return (e[0] * e[0] + e[1] * e[1]) + e[2] * e[2];
}

// computeActiveErrors + activeRobustChi2 at (pose, R, pts); stores the per-edge errors
template <int kCl>
__device__ double active_errors(const BaArgs& a, const Prob& p, const Team<kCl>& red, const double* pose, int pose_stride,
                                const double* R, int r_stride, const double* pts, int pts_stride, bool robust,
                                const double* prior_inv, const double* prior_info) {
  const FrameParamsDev& P = a.prm;
  double part = 0.0;
  for (int e = red.gtid; e < p.d.E; e += Team<kCl>::kThreads) {
    if (p.level[e]) continue;
    const int i = p.obs_kf[e], l = p.obs_pt[e];
    const double* Ri = R + (size_t)i * r_stride;
    const double* t = pose + (size_t)i * pose_stride + 4;
    const double* x = pts + (size_t)l * pts_stride;
    const double pc[3] = {(Ri[0] * x[0] + Ri[1] * x[1]) + Ri[2] * x[2] + t[0], (Ri[3] * x[0] + Ri[4] * x[1]) + Ri[5] * x[2] + t[1],
                          (Ri[6] * x[0] + Ri[7] * x[1]) + Ri[8] * x[2] + t[2]};
    const float* ob = p.obs + (size_t)e * 3;
    const bool stereo = !(ob[2] < 0.0f);
    double* er = p.wo + (size_t)e * kObsWs + oErr;
    edge_err(P, pc, ob, stereo, er);
    const double c2 = (double)P.s2inv[min((int)p.octave[e], 7)] * ((er[0] * er[0] + er[1] * er[1]) + er[2] * er[2]);
    if (robust) {
      const double dl = stereo ? P.delta_stereo : P.delta_mono, dsqr = dl * dl;
      part += (c2 <= dsqr) ? c2 : 2 * sqrt(c2) * dl - dsqr;
    } else {
      part += c2;
    }
  }
  for (int l = red.gtid; l < p.d.L; l += Team<kCl>::kThreads) {
    const int k = p.assoc[l];
    const int type = str_type_of(a, k);
    if (!type || p.slevel[l]) continue;
    double es[3];
    part += str_error(a, k, type, pts + (size_t)l * pts_stride, es);
  }
  if (red.gtid == 0 && p.d.prior >= 0) {
    double delta[7], dv[6];
    se3_mul_d(prior_inv, pose + (size_t)p.d.prior * pose_stride, delta);
    se3_log_d(delta, dv);
    for (int q = 0; q < 6; ++q) part += prior_info[q] * (dv[q] * dv[q]);
  }
  return red.sum(part);
}

template <int kCl>
__global__ void __launch_bounds__(kBaT) local_ba_kernel(const BaArgs a) {
  __shared__ double s_red[kBaW], s_tot[2];
  __shared__ int s_int[2];
  __shared__ double s_prior_inv[7], s_prior_info[6];
  constexpr int kGT = Team<kCl>::kThreads, kGW = Team<kCl>::kWarps;
  const int crank = (kCl == 1) ? 0 : (int)cg::this_cluster().block_rank();
  const int win = blockIdx.x / kCl;  // window solved by this team
  const int tid = threadIdx.x, gtid = crank * kBaT + tid;
  Team<kCl> red;
  red.s = s_red; red.s_tot = s_tot; red.s_int = s_int; red.crank = crank; red.gtid = gtid;
  const FrameParamsDev& P = a.prm;
  Prob p;
  p.d = a.prob[win];
  const BaProb& d = p.d;
  p.pose = a.pose + d.pose_off * 7;
  p.fidx = a.fidx + d.pose_off;
  p.pts = a.pts + d.pt_off * 3;
  p.assoc = a.assoc + d.pt_off;
  p.obs_pt = a.obs_pt + d.obs_off;
  p.obs_kf = a.obs_kf + d.obs_off;
  p.obs = a.obs + d.obs_off * 3;
  p.octave = a.octave + d.obs_off;
  p.pt_ptr = a.pt_ptr + d.pt_off + win;  // L + 1 entries per problem
  p.kf_ptr = a.kf_ptr + d.free_off + win;
  p.kf_obs = a.kf_obs + d.kfo_off;
  p.blk_ptr = a.blk_ptr + d.blk_off;
  p.contrib = a.contrib + d.c_off;
  p.wo = a.w_obs + d.obs_off * kObsWs;
  p.level = a.level + d.obs_off;
  p.wp = a.w_pt + d.pt_off * kPtWs;
  p.slevel = a.str_level + d.pt_off;
  p.wq = a.w_pose + d.pose_off * kPoseWs;
  p.wf = a.w_free + d.free_off * kFreeStride;
  p.S = a.S + d.S_off;
  const int n = d.n, nf = d.nf;
  p.rhs = p.S + (size_t)n * n;
  p.dxp = p.rhs + n;
  p.diag = p.dxp + n;

  // g2o's force-stop flag (optimizer.setForceStopFlag(pbStopFlag), localization_opt.cpp:541-542).  a.stop points to
  // host-mapped memory the calling thread keeps in step with the caller's flag while the kernel runs; thread 0 reads it
  // wherever g2o calls terminate() and the CTA takes the same branch.  *stop > 0: raised; 0: not raised; -n: reads as
  // raised from the n-th poll on (the deterministic replay hook the parity tests use).
  int polls = 0;  // counted by every thread alike (uniform control flow), used by the team's lead thread
  auto terminate = [&]() -> bool {
    if (a.stop == nullptr) return false;
    ++polls;
    int v = 0;
    if (gtid == 0) {
      const int raw = *a.stop;
      v = (raw > 0 || (raw < 0 && polls >= -raw)) ? 1 : 0;
    }
    return red.broadcast_from_lead(v) != 0;  // one read of the flag per poll: every CTA of the team takes the same branch
  };
  // `if (pbStopFlag) if (*pbStopFlag) return;` ahead of the first optimize (:765-767): the window is left untouched
  if (terminate()) {
    if (a.assoc_cleared)
      for (int l = gtid; l < d.L; l += kGT) a.assoc_cleared[d.pt_off + l] = 0;
    if (a.obs_outlier)
      for (int e = gtid; e < d.E; e += kGT) a.obs_outlier[d.obs_off + e] = 0;
    if (gtid == 0) {
      if (a.iters)
        for (int q = 0; q < 3; ++q) a.iters[(size_t)win * 3 + q] = 0;
      if (a.status) a.status[win] = 0;
    }
    return;
  }

  // ---- init: normalise poses, rotation cache, levels, prior measurement ----
  for (int i = gtid; i < d.P; i += kGT) {
    quat_normalize(p.pose + (size_t)i * 7);
    quat_to_rot(p.pose + (size_t)i * 7, p.wq + (size_t)i * kPoseWs + qR);
  }
  for (int e = gtid; e < d.E; e += kGT) {
    p.level[e] = 0;
    for (int q = 0; q < 3; ++q) p.wo[(size_t)e * kObsWs + oErr + q] = 0.0;
  }
  for (int l = gtid; l < d.L; l += kGT) p.slevel[l] = 0;
  red.sync();
  if (tid == 0 && d.prior >= 0) {
    se3_inverse_d(p.pose + (size_t)d.prior * 7, s_prior_inv);  // setMeasurement(kfi_ptr->getTcw()), :566
    const double sigma_rot = 2.0 * CUDART_PI / 180.0;
    for (int q = 0; q < 3; ++q) {
      s_prior_info[q] = 1.0 / (sigma_rot * sigma_rot);
      s_prior_info[3 + q] = 1.0 / (0.01 * 0.01);
    }
  }
  red.sync();

  bool robust = true;
  int iters_run[3] = {0, 0, 0};
  bool numeric_trouble = false;

  for (int stage = 0; stage < 3; ++stage) {
    const int max_it = (stage == 2) ? 40 : 5;
    double lambda = 0.0, ni = 2.0;
    bool ok = true;
    int it = 0;
#ifdef GMMLOC_BA_PROFILE
    unsigned long long ba_t0 = ba_now();
#endif
    for (; it < max_it && !terminate() && ok; ++it) {  // SparseOptimizer::optimize: i < iterations && !terminate() && ok
      double cur_chi = active_errors(a, p, red, p.pose, 7, p.wq + qR, kPoseWs, p.pts, 3, robust, s_prior_inv, s_prior_info);
      BA_T(0)
      // ---- linearise every active edge at the current estimate ----
      for (int e = gtid; e < d.E; e += kGT) {
        if (p.level[e]) continue;
        const int i = p.obs_kf[e], l = p.obs_pt[e], f = p.fidx[i];
        const double* Ri = p.wq + (size_t)i * kPoseWs + qR;
        const double* t = p.pose + (size_t)i * 7 + 4;
        const double* xw = p.pts + (size_t)l * 3;
        const double x = (Ri[0] * xw[0] + Ri[1] * xw[1]) + Ri[2] * xw[2] + t[0];
        const double y = (Ri[3] * xw[0] + Ri[4] * xw[1]) + Ri[5] * xw[2] + t[1];
        const double z = (Ri[6] * xw[0] + Ri[7] * xw[1]) + Ri[8] * xw[2] + t[2];
        const float* ob = p.obs + (size_t)e * 3;
        const bool stereo = !(ob[2] < 0.0f);
        const int m = stereo ? 3 : 2;
        const double pc[3] = {x, y, z};
        double er[3];
        edge_err(P, pc, ob, stereo, er);
        const double z_2 = z * z;
        double Jl[9], Jp[18];
        if (stereo) {
          for (int j = 0; j < 3; ++j) {
            Jl[j] = -P.fx * Ri[j] / z + P.fx * x * Ri[6 + j] / z_2;
            Jl[3 + j] = -P.fy * Ri[3 + j] / z + P.fy * y * Ri[6 + j] / z_2;
            Jl[6 + j] = Jl[j] - P.bf * Ri[6 + j] / z_2;
          }
        } else {  // EdgeSE3ProjectXYZ: -1. / z * tmp * R (the expression of factors.cpp:84-107 and of g2o's stock edge)
          const double sc = -1. / z;
          const double T00 = sc * P.fx, T02 = sc * (-x / z * P.fx), T11 = sc * P.fy, T12 = sc * (-y / z * P.fy);
          for (int j = 0; j < 3; ++j) {
            Jl[j] = T00 * Ri[j] + T02 * Ri[6 + j];
            Jl[3 + j] = T11 * Ri[3 + j] + T12 * Ri[6 + j];
            Jl[6 + j] = 0.0;
          }
        }
        Jp[0] = x * y / z_2 * P.fx;
        Jp[1] = -(1 + (x * x / z_2)) * P.fx;
        Jp[2] = y / z * P.fx;
        Jp[3] = -1. / z * P.fx;
        Jp[4] = 0;
        Jp[5] = x / z_2 * P.fx;
        Jp[6] = (1 + y * y / z_2) * P.fy;
        Jp[7] = -x * y / z_2 * P.fy;
        Jp[8] = -x / z * P.fy;
        Jp[9] = 0;
        Jp[10] = -1. / z * P.fy;
        Jp[11] = y / z_2 * P.fy;
        if (stereo) {
          Jp[12] = Jp[0] - P.bf * y / z_2;
          Jp[13] = Jp[1] + P.bf * x / z_2;
          Jp[14] = Jp[2];
          Jp[15] = Jp[3];
          Jp[16] = 0;
          Jp[17] = Jp[5] - P.bf / z_2;
        } else {
          for (int j = 12; j < 18; ++j) Jp[j] = 0;
        }
        double we = (double)P.s2inv[min((int)p.octave[e], 7)];
        if (robust) {
          double c2 = 0;
          for (int r = 0; r < m; ++r) c2 += er[r] * er[r];
          c2 *= we;
          const double dl = stereo ? P.delta_stereo : P.delta_mono;
          if (c2 > dl * dl) we *= dl / sqrt(c2);
        }
        double* o = p.wo + (size_t)e * kObsWs;
        int q = 0;
        for (int r = 0; r < 3; ++r)
          for (int c = r; c < 3; ++c) {
            double s = 0;
            for (int k = 0; k < m; ++k) s += Jl[k * 3 + r] * Jl[k * 3 + c];
            o[oHl + q++] = we * s;
          }
        for (int r = 0; r < 3; ++r) {
          double s = 0;
          for (int k = 0; k < m; ++k) s += Jl[k * 3 + r] * er[k];
          o[oBl + r] = -(we * s);
        }
        if (f >= 0) {
          q = 0;
          for (int r = 0; r < 6; ++r)
            for (int c = r; c < 6; ++c) {
              double s = 0;
              for (int k = 0; k < m; ++k) s += Jp[k * 6 + r] * Jp[k * 6 + c];
              o[oA + q++] = we * s;
            }
          for (int r = 0; r < 6; ++r) {
            double s = 0;
            for (int k = 0; k < m; ++k) s += Jp[k * 6 + r] * er[k];
            o[oG + r] = -(we * s);
            for (int c = 0; c < 3; ++c) {
              double s2 = 0;
              for (int k = 0; k < m; ++k) s2 += Jp[k * 6 + r] * Jl[k * 3 + c];
              o[oW + r * 3 + c] = we * s2;
            }
          }
        }
      }
      red.sync();
      BA_T(1)
      // ---- per point: Hll, bl (reprojection pieces of its observations + structure edge) ----
      double md = 0.0;
      for (int l = gtid; l < d.L; l += kGT) {
        double H[6] = {0, 0, 0, 0, 0, 0}, b[3] = {0, 0, 0};
        for (int e = p.pt_ptr[l]; e < p.pt_ptr[l + 1]; ++e) {
          if (p.level[e]) continue;
          const double* o = p.wo + (size_t)e * kObsWs;
          for (int q = 0; q < 6; ++q) H[q] += o[oHl + q];
          for (int q = 0; q < 3; ++q) b[q] += o[oBl + q];
        }
        const int k = p.assoc[l];
        const int type = str_type_of(a, k);
        if (type && !p.slevel[l]) {
          double es[3];
          str_error(a, k, type, p.pts + (size_t)l * 3, es);
          if (type == 1) {
            const double* ax = a.axis + (size_t)k * 9;
            const double nn[3] = {ax[0], ax[3], ax[6]};
            int q = 0;
            for (int r = 0; r < 3; ++r) {
              for (int c = r; c < 3; ++c) H[q++] += P.ba_lambda2 * (nn[r] * nn[c]);
              b[r] -= P.ba_lambda2 * (nn[r] * es[0]);
            }
          } else {
            const double* L = a.sqrt_info + (size_t)k * 9;
            int q = 0;
            for (int r = 0; r < 3; ++r) {
              for (int c = r; c < 3; ++c) H[q++] += (L[r * 3] * L[c * 3] + L[r * 3 + 1] * L[c * 3 + 1]) + L[r * 3 + 2] * L[c * 3 + 2];
              b[r] -= (L[r * 3] * es[0] + L[r * 3 + 1] * es[1]) + L[r * 3 + 2] * es[2];
            }
          }
        }
        double* w = p.wp + (size_t)l * kPtWs;
        for (int q = 0; q < 6; ++q) w[pH + q] = H[q];
        for (int q = 0; q < 3; ++q) w[pB + q] = b[q];
        md = fmax(md, fmax(fabs(H[0]), fmax(fabs(H[3]), fabs(H[5]))));
      }
      // ---- per free pose: Hpp (21) and bp (6) from its observations.  The walk over a pose's observation list is a chain of
      // dependent L2 round trips (index -> level -> 27 doubles): with one warp per pose nine warps worked for 47 us of every
      // LM iteration of a lone window while the rest of the team waited.  kFoldParts warps now share a list (warp `part`
      // takes the entries part * 32 + lane, + 32 kFoldParts, ...), write their 27 partial sums, and after a team barrier
      // one thread per value adds the partials in part order: deterministic, independent of the cluster size ----
      for (int task = gtid >> 5; task < nf * kFoldParts; task += kGW) {
        const int f = task / kFoldParts, part = task % kFoldParts;
        double acc[kFreeWs];
#pragma unroll
        for (int q = 0; q < kFreeWs; ++q) acc[q] = 0.0;
        for (int j = p.kf_ptr[f] + part * 32 + (tid & 31); j < p.kf_ptr[f + 1]; j += 32 * kFoldParts) {
          const int e = p.kf_obs[j];
          const double* o = p.wo + (size_t)e * kObsWs;  // oA = 0 .. 20, oG = 21 .. 26: same order as fH, fB
          const bool on = !p.level[e];
#pragma unroll
          for (int q = 0; q < kFreeWs; ++q) acc[q] += on ? o[q] : 0.0;
        }
#pragma unroll
        for (int q = 0; q < kFreeWs; ++q) {
          double v = acc[q];
#pragma unroll
          for (int off = 16; off > 0; off >>= 1) v += __shfl_xor_sync(0xffffffffu, v, off);
          if ((tid & 31) == 0) p.wf[(size_t)f * kFreeStride + kFreeWs * (1 + part) + q] = v;
        }
      }
      red.sync();
      for (int i = gtid; i < nf * kFreeWs; i += kGT) {
        const int f = i / kFreeWs, q = i % kFreeWs;
        const double* pp = p.wf + (size_t)f * kFreeStride + kFreeWs + q;
        double v = pp[0];
#pragma unroll
        for (int part = 1; part < kFoldParts; ++part) v += pp[kFreeWs * part];
        p.wf[(size_t)f * kFreeStride + q] = v;
      }
      red.sync();
      BA_T(2)
      if (gtid == 0 && d.prior >= 0 && p.fidx[d.prior] >= 0) {  // EdgeSE3QuatPrior, factors.cpp:31-53
        const int f = p.fidx[d.prior];
        double delta[7], dv[6], Jr[36], Tinv[7], Ad[36], J[36], Rinv[9];
        se3_mul_d(s_prior_inv, p.pose + (size_t)d.prior * 7, delta);
        se3_log_d(delta, dv);
        for (int q = 0; q < 36; ++q) { Jr[q] = 0; Ad[q] = 0; }
        const double ph[9] = {0, -dv[2], dv[1], dv[2], 0, -dv[0], -dv[1], dv[0], 0};
        const double lu[9] = {0, -dv[5], dv[4], dv[5], 0, -dv[3], -dv[4], dv[3], 0};
        for (int r = 0; r < 3; ++r)
          for (int c = 0; c < 3; ++c) {
            Jr[r * 6 + c] = 0.5 * ph[r * 3 + c];
            Jr[(3 + r) * 6 + 3 + c] = 0.5 * ph[r * 3 + c];
            Jr[r * 6 + 3 + c] = 0.5 * lu[r * 3 + c];
          }
        for (int q = 0; q < 6; ++q) Jr[q * 6 + q] += 1.0;
        se3_inverse_d(p.pose + (size_t)d.prior * 7, Tinv);
        quat_to_rot(Tinv, Rinv);
        const double* tt = Tinv + 4;
        const double St[9] = {0, -tt[2], tt[1], tt[2], 0, -tt[0], -tt[1], tt[0], 0};
        for (int r = 0; r < 3; ++r)
          for (int c = 0; c < 3; ++c) {
            Ad[r * 6 + c] = Rinv[r * 3 + c];
            Ad[(3 + r) * 6 + 3 + c] = Rinv[r * 3 + c];
            Ad[(3 + r) * 6 + c] = (St[r * 3] * Rinv[c] + St[r * 3 + 1] * Rinv[3 + c]) + St[r * 3 + 2] * Rinv[6 + c];
          }
        for (int r = 0; r < 6; ++r)
          for (int c = 0; c < 6; ++c) {
            double s = 0;
            for (int k = 0; k < 6; ++k) s += Jr[r * 6 + k] * Ad[k * 6 + c];
            J[r * 6 + c] = s;
          }
        double* w = p.wf + (size_t)f * kFreeStride;
        int q = 0;
        for (int r = 0; r < 6; ++r) {
          for (int c = r; c < 6; ++c) {
            double s = 0;
            for (int k = 0; k < 6; ++k) s += J[k * 6 + r] * s_prior_info[k] * J[k * 6 + c];
            w[fH + q++] += s;
          }
          double s = 0;
          for (int k = 0; k < 6; ++k) s += J[k * 6 + r] * s_prior_info[k] * dv[k];
          w[fB + r] -= s;
        }
      }
      red.sync();
      if (it == 0) {  // computeLambdaInit: tau * max |diagonal|, tau = 1e-5
        for (int f = gtid; f < nf; f += kGT) {
          const double* w = p.wf + (size_t)f * kFreeStride + fH;
          for (int r = 0; r < 6; ++r) md = fmax(md, fabs(w[sym6(r, r)]));
        }
        lambda = 1e-5 * red.max(md);
        ni = 2.0;
      }
      double rho = 0.0;
      int qmax = 0;
      do {
        BA_T(3)
        // ---- Schur complement at damping lambda ----
        bool good = true;
        for (int l = gtid; l < d.L; l += kGT) {
          double* w = p.wp + (size_t)l * kPtWs;
          const double hxx = w[pH] + lambda, hxy = w[pH + 1], hxz = w[pH + 2], hyy = w[pH + 3] + lambda, hyz = w[pH + 4],
                       hzz = w[pH + 5] + lambda;
          // Eigen 3x3 inverse by cofactors (g2o: Hll.inverse())
          const double c00 = hyy * hzz - hyz * hyz, c10 = hyz * hxz - hxy * hzz, c20 = hxy * hyz - hyy * hxz;
          const double det = (c00 * hxx + c10 * hxy) + c20 * hxz;
          const double id = 1.0 / det;
          if (!(isfinite(id) && det != 0.0)) good = false;
          const double i0 = c00 * id, i1 = c10 * id, i2 = c20 * id, i3 = (hxx * hzz - hxz * hxz) * id,
                       i4 = (hxy * hxz - hxx * hyz) * id, i5 = (hxx * hyy - hxy * hxy) * id;
          w[pHi] = i0; w[pHi + 1] = i1; w[pHi + 2] = i2; w[pHi + 3] = i3; w[pHi + 4] = i4; w[pHi + 5] = i5;
          w[pY] = (i0 * w[pB] + i1 * w[pB + 1]) + i2 * w[pB + 2];
          w[pY + 1] = (i1 * w[pB] + i3 * w[pB + 1]) + i4 * w[pB + 2];
          w[pY + 2] = (i2 * w[pB] + i4 * w[pB + 1]) + i5 * w[pB + 2];
        }
        good = red.all_ok(good) != 0;  // also orders the Hinv writes before the block tasks
        if (good) {
          BA_T(4)
          // per observation on a free pose: T = Hpl Hll^-1 (6x3) and Hpl (Hll^-1 bl) (6)
          for (int e = gtid; e < d.E; e += kGT) {
            if (p.level[e] || p.fidx[p.obs_kf[e]] < 0) continue;
            double* o = p.wo + (size_t)e * kObsWs;
            const double* w = p.wp + (size_t)p.obs_pt[e] * kPtWs;
            const double* Hi = w + pHi;
            const double* yl = w + pY;
#pragma unroll
            for (int r = 0; r < 6; ++r) {
              const double w0 = o[oW + r * 3], w1 = o[oW + r * 3 + 1], w2 = o[oW + r * 3 + 2];
              o[oT + r * 3] = (w0 * Hi[0] + w1 * Hi[1]) + w2 * Hi[2];
              o[oT + r * 3 + 1] = (w0 * Hi[1] + w1 * Hi[3]) + w2 * Hi[4];
              o[oT + r * 3 + 2] = (w0 * Hi[2] + w1 * Hi[4]) + w2 * Hi[5];
              o[oGy + r] = (w0 * yl[0] + w1 * yl[1]) + w2 * yl[2];
            }
          }
          red.sync();
          BA_T(5)
          // reduced system: one WARP per HALF (three rows) of every lower 6x6 block; the lanes stride over the block's
          // contribution list and fold with a fixed shuffle tree.  (History: one thread per row walked ~2 500 pairs one
          // dependent L2 round trip at a time; one warp per row then read every pair's 6x3 W block six times over and still
          // took 8.9 of the 20.9 ms of a lone 12-keyframe window — the walk is latency-bound, so a task now carries three
          // rows' worth of independent loads per list entry and the whole system is one round of warps.)  The order of the
          // additions per entry of S is unchanged: lane l sums entries l, l + 32, ... in list order, then the tree.
          const int nblk = nf * (nf + 1) / 2;
          for (int task = gtid >> 5; task < nblk * 2; task += kGW) {
            const int blk = task >> 1, r0 = 3 * (task & 1);
            // blk = fi (fi + 1) / 2 + fj, fj <= fi
            int fi = (int)((sqrt(8.0 * blk + 1.0) - 1.0) * 0.5);
            while ((fi + 1) * (fi + 2) / 2 <= blk) ++fi;
            while (fi * (fi + 1) / 2 > blk) --fi;
            const int fj = blk - fi * (fi + 1) / 2;
            double acc[18];
#pragma unroll
            for (int q = 0; q < 18; ++q) acc[q] = 0.0;
            const int j1 = p.blk_ptr[blk + 1];
            // the next entry's observation pair is requested while this one is worked on, and the data are loaded next to
            // the level flags instead of behind them: one dependent round trip per entry instead of three
            int j = p.blk_ptr[blk] + (tid & 31);
            int2 ab = (j < j1) ? p.contrib[j] : make_int2(0, 0);
            while (j < j1) {
              const int jn = j + 32;
              const int2 abn = (jn < j1) ? p.contrib[jn] : ab;
              const bool on = !(p.level[ab.x] | p.level[ab.y]);
              const double* Ta = p.wo + (size_t)ab.x * kObsWs + oT + r0 * 3;
              const double* Wb = p.wo + (size_t)ab.y * kObsWs + oW;
              double wb[18], ta[9];
#pragma unroll
              for (int q = 0; q < 18; ++q) wb[q] = Wb[q];
#pragma unroll
              for (int q = 0; q < 9; ++q) ta[q] = Ta[q];
              if (on) {
#pragma unroll
                for (int rr = 0; rr < 3; ++rr) {
                  const double t0 = ta[rr * 3], t1 = ta[rr * 3 + 1], t2 = ta[rr * 3 + 2];
#pragma unroll
                  for (int c = 0; c < 6; ++c) acc[rr * 6 + c] -= (t0 * wb[c * 3] + t1 * wb[c * 3 + 1]) + t2 * wb[c * 3 + 2];
                }
              }
              j = jn;
              ab = abn;
            }
#pragma unroll
            for (int rr = 0; rr < 3; ++rr) {
              const int r = r0 + rr;
#pragma unroll
              for (int c = 0; c < 6; ++c) {
                double v = acc[rr * 6 + c];
#pragma unroll
                for (int off = 16; off > 0; off >>= 1) v += __shfl_xor_sync(0xffffffffu, v, off);
                if ((tid & 31) == 0) {
                  if (fi == fj) v += p.wf[(size_t)fi * kFreeStride + fH + sym6(r, c)] + (r == c ? lambda : 0.0);
                  p.S[(size_t)(6 * fi + r) * n + 6 * fj + c] = v;
                }
              }
            }
          }
          // right-hand side: kFoldParts warps per free pose sum their share of the pose's observation list (partials into
          // the fold scratch of w_free, free again by now); CTA 0 adds them in part order before it factorises
          for (int task = gtid >> 5; task < nblk * 2 + nf * kFoldParts; task += kGW) {  // one task list with the blocks above
            if (task < nblk * 2) continue;
            const int t = task - nblk * 2, f = t / kFoldParts, part = t % kFoldParts;
            double acc[6] = {0, 0, 0, 0, 0, 0};
            for (int j = p.kf_ptr[f] + part * 32 + (tid & 31); j < p.kf_ptr[f + 1]; j += 32 * kFoldParts) {
              const int e = p.kf_obs[j];
              const double* gy = p.wo + (size_t)e * kObsWs + oGy;
              const bool on = !p.level[e];
#pragma unroll
              for (int r = 0; r < 6; ++r) acc[r] += on ? gy[r] : 0.0;
            }
#pragma unroll
            for (int r = 0; r < 6; ++r) {
              double v = acc[r];
#pragma unroll
              for (int off = 16; off > 0; off >>= 1) v += __shfl_xor_sync(0xffffffffu, v, off);
              if ((tid & 31) == 0) p.wf[(size_t)f * kFreeStride + kFreeWs * (1 + part) + r] = v;
            }
          }
          red.sync();
          BA_T(6)
          // ---- left-looking Cholesky of the lower triangle; the rhs rides along as row n (forward solve) ----
          // The reduced system is small (n = 6 free poses): its ~3 n barriers stay inside CTA 0 of the team; the other
          // CTAs wait at the team-wide vote below.
          if (crank == 0) {
            for (int i = tid; i < n; i += kBaT) {
              const int f = i / 6, r = i % 6;
              const double* pp = p.wf + (size_t)f * kFreeStride + kFreeWs + r;
              double v = pp[0];
#pragma unroll
              for (int part = 1; part < kFoldParts; ++part) v += pp[kFreeWs * part];
              p.rhs[i] = p.wf[(size_t)f * kFreeStride + fB + r] - v;
            }
            __syncthreads();
            for (int k = 0; k < n && good; ++k) {
              for (int i = k + tid; i <= n; i += kBaT) {
                double* row = (i < n) ? p.S + (size_t)i * n : p.rhs;
                const double* rk = p.S + (size_t)k * n;
                double s = row[k];
                for (int j = 0; j < k; ++j) s -= row[j] * rk[j];
                row[k] = s;
              }
              __syncthreads();
              const double dkk = p.S[(size_t)k * n + k];
              good = (dkk > 0.0) && isfinite(dkk);  // uniform: every thread reads the same value
              const double sd = sqrt(dkk);
              __syncthreads();
              if (good) {
                for (int i = k + tid; i <= n; i += kBaT) {
                  double* row = (i < n) ? p.S + (size_t)i * n : p.rhs;
                  if (i == k) {
                    p.diag[k] = sd;
                    row[k] = sd;
                  } else {
                    row[k] = row[k] / sd;
                  }
                }
              }
              __syncthreads();
            }
            // ---- back substitution L^T x = y (y is in rhs) ----
            if (good) {
              for (int k = n - 1; k >= 0; --k) {
                if (tid == 0) p.dxp[k] = p.rhs[k] / p.diag[k];
                __syncthreads();
                const double xk = p.dxp[k];
                for (int i = tid; i < k; i += kBaT) p.rhs[i] -= p.S[(size_t)k * n + i] * xk;
                __syncthreads();
              }
              bool fin = true;
              for (int i = tid; i < n; i += kBaT) fin = fin && isfinite(p.dxp[i]);
              good = __syncthreads_and(fin ? 1 : 0) != 0;
            }
          }
          good = red.all_ok(good) != 0;  // CTA 0's verdict reaches the whole team; also publishes dxp
        }
        BA_T(7)
        // ---- landmark back-substitution, trial state ----
        if (good) {
          bool fin = true;
          for (int l = gtid; l < d.L; l += kGT) {
            double* w = p.wp + (size_t)l * kPtWs;
            double c[3] = {w[pB], w[pB + 1], w[pB + 2]};
            for (int e = p.pt_ptr[l]; e < p.pt_ptr[l + 1]; ++e) {
              if (p.level[e]) continue;
              const int f = p.fidx[p.obs_kf[e]];
              if (f < 0) continue;
              const double* W = p.wo + (size_t)e * kObsWs + oW;
              const double* dx = p.dxp + 6 * f;
              for (int q = 0; q < 3; ++q) {
                double s = 0;
                for (int r = 0; r < 6; ++r) s += W[r * 3 + q] * dx[r];
                c[q] -= s;
              }
            }
            const double* Hi = w + pHi;
            w[pDx] = (Hi[0] * c[0] + Hi[1] * c[1]) + Hi[2] * c[2];
            w[pDx + 1] = (Hi[1] * c[0] + Hi[3] * c[1]) + Hi[4] * c[2];
            w[pDx + 2] = (Hi[2] * c[0] + Hi[4] * c[1]) + Hi[5] * c[2];
            for (int q = 0; q < 3; ++q) {
              fin = fin && isfinite(w[pDx + q]);
              w[pTr + q] = p.pts[(size_t)l * 3 + q] + w[pDx + q];
            }
          }
          good = red.all_ok(fin) != 0;
        }
        BA_T(8)
        double tmp_chi = 1.7976931348623157e308, scale = 0.0;
        if (good) {
          for (int i = gtid; i < d.P; i += kGT) {
            double* w = p.wq + (size_t)i * kPoseWs;
            const int f = p.fidx[i];
            if (f >= 0) {
              se3_oplus(p.dxp + 6 * f, p.pose + (size_t)i * 7, w + qTr);
            } else {
              for (int q = 0; q < 7; ++q) w[qTr + q] = p.pose[(size_t)i * 7 + q];
            }
            quat_to_rot(w + qTr, w + qRt);
          }
          red.sync();
          tmp_chi = active_errors(a, p, red, p.wq + qTr, kPoseWs, p.wq + qRt, kPoseWs, p.wp + pTr, kPtWs, robust, s_prior_inv,
                                  s_prior_info);
          double part = 0.0;
          for (int i = gtid; i < n; i += kGT) {
            const double dx = p.dxp[i];
            part += dx * (lambda * dx + p.wf[(size_t)(i / 6) * kFreeStride + fB + i % 6]);
          }
          for (int l = gtid; l < d.L; l += kGT) {
            const double* w = p.wp + (size_t)l * kPtWs;
            for (int q = 0; q < 3; ++q) part += w[pDx + q] * (lambda * w[pDx + q] + w[pB + q]);
          }
          scale = red.sum(part);
        }
        BA_T(9)
        rho = cur_chi - tmp_chi;
        rho /= scale + 1e-3;
        if (rho > 0 && isfinite(tmp_chi) && good) {
          double alpha = 1. - pow(2 * rho - 1, 3.0);
          alpha = fmin(alpha, 2. / 3.);
          lambda *= fmax(1. / 3., alpha);
          ni = 2;
          cur_chi = tmp_chi;
          for (int i = gtid; i < d.P; i += kGT) {
            double* w = p.wq + (size_t)i * kPoseWs;
            for (int q = 0; q < 7; ++q) p.pose[(size_t)i * 7 + q] = w[qTr + q];
            for (int q = 0; q < 9; ++q) w[qR + q] = w[qRt + q];
          }
          for (int l = gtid; l < d.L; l += kGT)
            for (int q = 0; q < 3; ++q) p.pts[(size_t)l * 3 + q] = p.wp[(size_t)l * kPtWs + pTr + q];
          red.sync();
        } else {
          if (!good) numeric_trouble = true;
          lambda *= ni;
          ni *= 2;
          if (!isfinite(lambda)) break;
        }
        ++qmax;
        BA_T(10)
      } while (rho < 0 && qmax < 10 && !terminate());  // OptimizationAlgorithmLevenberg::solve
      if (qmax == 10 || rho == 0 || !isfinite(lambda)) ok = false;
    }
    iters_run[stage] = it;
    // ---- edge gating between the stages ----
    if (stage == 0) {  // :774-786
      for (int l = gtid; l < d.L; l += kGT) {
        const int k = p.assoc[l];
        if (str_type_of(a, k) != 1) continue;
        double es[3];
        if (str_error(a, k, 1, p.pts + (size_t)l * 3, es) > P.tri_str_thresh * P.ba_lambda2) p.slevel[l] = 1;
      }
    } else if (stage == 1) {  // :797-823
      if (terminate()) break;  // bDoMore = false (:791-795): no reprojection gating, no kernel removal, no stage 3
      for (int e = gtid; e < d.E; e += kGT) {
        const double* er = p.wo + (size_t)e * kObsWs + oErr;
        const float* ob = p.obs + (size_t)e * 3;
        const bool stereo = !(ob[2] < 0.0f);
        const double c2 = (double)P.s2inv[min((int)p.octave[e], 7)] * ((er[0] * er[0] + er[1] * er[1]) + er[2] * er[2]);
        const int i = p.obs_kf[e];
        const double* Ri = p.wq + (size_t)i * kPoseWs + qR;
        const double* xw = p.pts + (size_t)p.obs_pt[e] * 3;
        const double z = (Ri[6] * xw[0] + Ri[7] * xw[1]) + Ri[8] * xw[2] + p.pose[(size_t)i * 7 + 6];
        if (c2 > (stereo ? 7.815 : 5.991) || !(z > 0.0)) p.level[e] = 1;
      }
      robust = false;
    }
    red.sync();
  }

  // ---- results (:842-881) ----
  for (int l = gtid; l < d.L; l += kGT) {
    uint8_t cl = 0;
    const int k = p.assoc[l];
    if (str_type_of(a, k) == 1) {
      double es[3];
      cl = str_error(a, k, 1, p.pts + (size_t)l * 3, es) > P.tri_str_thresh * P.ba_lambda2;
    }
    if (a.assoc_cleared) a.assoc_cleared[d.pt_off + l] = cl;
  }
  if (a.obs_outlier) {
    for (int e = gtid; e < d.E; e += kGT) {
      const double* er = p.wo + (size_t)e * kObsWs + oErr;
      const float* ob = p.obs + (size_t)e * 3;
      const bool stereo = !(ob[2] < 0.0f);
      const double c2 = (double)P.s2inv[min((int)p.octave[e], 7)] * ((er[0] * er[0] + er[1] * er[1]) + er[2] * er[2]);
      const int i = p.obs_kf[e];
      const double* Ri = p.wq + (size_t)i * kPoseWs + qR;
      const double* xw = p.pts + (size_t)p.obs_pt[e] * 3;
      const double z = (Ri[6] * xw[0] + Ri[7] * xw[1]) + Ri[8] * xw[2] + p.pose[(size_t)i * 7 + 6];
      a.obs_outlier[d.obs_off + e] = (c2 > (stereo ? 7.815 : 5.991) || !(z > 0.0)) ? 1 : 0;
    }
  }
  if (gtid == 0) {
    if (a.iters)
      for (int q = 0; q < 3; ++q) a.iters[(size_t)win * 3 + q] = iters_run[q];
    if (a.status) a.status[win] = numeric_trouble ? GMMLOC_FRAME_NUMERIC : 0;
  }
}

}  // namespace

template <int kCl>
static cudaError_t launch_ba_t(const BaArgs& a, cudaStream_t stream) {
  if (kCl == 1) {
    local_ba_kernel<1><<<a.B, kBaT, 0, stream>>>(a);
    return cudaGetLastError();
  }
  cudaLaunchConfig_t cfg = {};
  cfg.gridDim = dim3((unsigned)a.B * kCl, 1, 1);
  cfg.blockDim = dim3(kBaT, 1, 1);
  cfg.dynamicSmemBytes = 0;
  cfg.stream = stream;
  cudaLaunchAttribute attr[1];
  attr[0].id = cudaLaunchAttributeClusterDimension;
  attr[0].val.clusterDim.x = kCl;
  attr[0].val.clusterDim.y = 1;
  attr[0].val.clusterDim.z = 1;
  cfg.attrs = attr;
  cfg.numAttrs = 1;
  return cudaLaunchKernelEx(&cfg, local_ba_kernel<kCl>, a);
}

cudaError_t launch_local_ba(const BaArgs& a, cudaStream_t stream, int cluster) {
  if (a.B <= 0) return cudaSuccess;
  // CTAs per window: as many as keep the machine busy with the windows of this call (148 SMs), at most the portable
  // cluster size 8 — a single window (what the localisation thread solves per keyframe) gets 8 SMs, a batch of a
  // hundred windows one CTA each (where the single-CTA kernel already fills the GPU)
  int dev = 0, sms = 148;
  cudaGetDevice(&dev);
  cudaDeviceGetAttribute(&sms, cudaDevAttrMultiProcessorCount, dev);
  int cl = cluster > 0 ? cluster : (a.B * 8 <= sms ? 8 : a.B * 4 <= sms ? 4 : a.B * 2 <= sms ? 2 : 1);
#ifdef GMMLOC_BA_PROFILE
  {
    unsigned long long z[16] = {0};
    cudaMemcpyToSymbol(g_ba_prof, z, sizeof z);
  }
#endif
  cudaError_t e;
  switch (cl) {
    case 8: e = launch_ba_t<8>(a, stream); break;
    case 4: e = launch_ba_t<4>(a, stream); break;
    case 2: e = launch_ba_t<2>(a, stream); break;
    default: e = launch_ba_t<1>(a, stream); break;
  }
#ifdef GMMLOC_BA_PROFILE
  if (e == cudaSuccess) {
    cudaStreamSynchronize(stream);
    unsigned long long t[16];
    cudaMemcpyFromSymbol(t, g_ba_prof, sizeof t);
    static const char* names[11] = {"poll+active_errors", "linearise", "Hll/Hpp", "prior+lambda0", "Hll inverse", "T per obs",
                                    "block rows+rhs", "cholesky+backsub", "landmark backsub", "trial errors+scale", "accept/reject"};
    std::fprintf(stderr, "BA profile (cluster %d):", cl);
    for (int i = 0; i < 11; ++i) std::fprintf(stderr, " %s %.3f ms |", names[i], t[i] * 1e-6);
    std::fprintf(stderr, "\n");
  }
#endif
  return e;
}

int local_ba_obs_ws() { return kObsWs; }
int local_ba_pt_ws() { return kPtWs; }
int local_ba_pose_ws() { return kPoseWs; }
int local_ba_free_ws() { return kFreeStride; }

}  // namespace gmmloc
