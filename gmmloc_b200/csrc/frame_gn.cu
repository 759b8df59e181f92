// frame_gn.cu — K3: per-frame structure-constrained Gauss-Newton (one CTA per frame, fp64).
//
// Replaces the single-free-pose specialisation of Localization::jointOptimization (reference
// gmmloc/src/modules/localization_opt.cpp:456-925): one VertexSE3Expmap + N marginalised
// VertexSBAPointXYZ, one reprojection edge per point (EdgeStereoSE3ProjectXYZ / EdgeSE3ProjectXYZ,
// point Jacobians as gmmloc/src/gmm/factors.cpp:84-107,133-168) and one structure edge per associated
// point (EdgePt2GaussianDeg factors.cpp:55-64, information ba_lambda2, localization_opt.cpp:657-669;
// EdgePt2Gaussian factors.cpp:5-17, information I, localization_opt.cpp:670-681).  factor_mode
// POSE_ONLY gives the system of Tracking::optimizeCurrentPose (gmmloc/src/modules/tracking_opt.cpp:21-217),
// POINTS_ONLY the per-point structure-constrained solves.
//
// Formulation.  BlockSolver_6_3's Schur complement is evaluated per point in the
// CAMERA frame, where every Jacobian factors through the 3x3 projection derivative P = d proj / d pc:
//     G = w P^T P,  h = w P^T e                       (reprojection, camera frame)
//     S_c = R S R^T, s_c = R b_s                      (structure edge rotated into the camera frame)
//     D = G + S_c  (= R Hll R^T),  T = G D^-1
//     pose block   += A^T (T S_c) A,   pose rhs += A^T (h - T (h + s_c)),   A = d pc / d xi = [-[pc]x  I]
//     point update  dx_w = R^T D^-1 (h + s_c - G A dxi)
// which equals Hpp - Hpl Hll^-1 Hpl^T etc. of the reference, with 3x3 work only.  A free point WITHOUT a
// structure edge has T = I: it absorbs its own reprojection error and contributes exactly nothing to
// the pose system, so it is skipped in the accumulation pass.
//
// Per-frame state (landmarks in fp64, edge flags) lives in shared memory for all iterations: HBM is
// read once per frame.  Reductions are fixed-shape (per-thread strided sum -> warp shuffle tree ->
// fixed-order cross-warp sum): deterministic, no float atomics.
#include <math_constants.h>

#include "common.cuh"
#include "gn_common.cuh"
#include "frame_gn_point.cuh"

namespace gmmloc {

namespace {

#ifndef GMMLOC_GN_THREADS
#define GMMLOC_GN_THREADS 128
#endif
#ifndef GMMLOC_GN_MINB
#define GMMLOC_GN_MINB 3
#endif
constexpr int kThreads = GMMLOC_GN_THREADS;
constexpr int kWarps = kThreads / 32;
constexpr int kAcc = 27;  // 21 upper-triangular entries of the 6x6 + 6 rhs

// numeric-failure contract (shared with the oracle): a frame whose solve hits a singular Hll / reduced
// system or a non-finite value is returned UNSOLVED (normalised pose0, input landmarks) with
// GMMLOC_FRAME_NUMERIC set.
__global__ void __launch_bounds__(kThreads, GMMLOC_GN_MINB) frame_gn_kernel(const FrameArgs a) {
  extern __shared__ __align__(16) unsigned char smem_raw[];
  __shared__ double s_red[kWarps][kAcc];
  __shared__ double s_acc[kAcc];
  __shared__ double s_pose[7], s_R[9], s_dxi[6], s_Ro[9], s_to[3];
  __shared__ int s_fail1, s_fail2;

  int f = blockIdx.x;
  if (a.retry_list != nullptr) {  // second opinion on the frames a mixed-precision kernel gave up on (frame_gn_stream.cu)
    if (f >= *a.retry_count) return;
    f = a.retry_list[f];
  }
  const int tid = threadIdx.x, lane = tid & 31, warp = tid >> 5;
  const int N = a.N;
  const FrameParamsDev& P = a.prm;
  double* xs;
  uint8_t* fl;
  if (a.xws != nullptr) {
    xs = a.xws + (size_t)f * N * 3;
    fl = a.fws + (size_t)f * N;
  } else {
    xs = reinterpret_cast<double*>(smem_raw);
    fl = reinterpret_cast<uint8_t*>(xs + (size_t)N * 3);
  }
  const float* pts = a.pts_w + (size_t)f * N * 3;
  const float* obs = a.obs + (size_t)f * N * 3;
  const uint8_t* oct = a.octave + (size_t)f * N;
  const int32_t* asg = a.assign + (size_t)f * N;
  const bool pose_free = P.factor_mode != GMMLOC_MODE_POINTS_ONLY;
  const bool pts_free = P.factor_mode != GMMLOC_MODE_POSE_ONLY;

  if (tid == 0) {
    double q[7];
    for (int i = 0; i < 7; ++i) q[i] = a.pose0[(size_t)f * 7 + i];
    quat_normalize(q);
    for (int i = 0; i < 7; ++i) s_pose[i] = q[i];
    quat_to_rot(q, s_R);
    for (int i = 0; i < 6; ++i) s_dxi[i] = 0.0;
    for (int i = 0; i < 9; ++i) s_Ro[i] = s_R[i];
    for (int i = 0; i < 3; ++i) s_to[i] = q[4 + i];
    s_fail1 = 0;
    s_fail2 = 0;
  }
  __syncthreads();
  // camera centre at the initial pose, t_w_c = -(R^T t)  (gaussian_mixture.cpp:289)
  const double twc0 = -(s_R[0] * s_pose[4] + s_R[3] * s_pose[5] + s_R[6] * s_pose[6]);
  const double twc1 = -(s_R[1] * s_pose[4] + s_R[4] * s_pose[5] + s_R[7] * s_pose[6]);
  const double twc2 = -(s_R[2] * s_pose[4] + s_R[5] * s_pose[5] + s_R[8] * s_pose[6]);
  const double view_cos_thresh = cos(78.0 * 3.14159265358979323846 / 180.0);  // gaussian_mixture.cpp:284
  // ---- landmarks to fp64; structure-edge gates fixed at iteration 0 (exact replay of the oracle's
  // comparisons so the active edge set is identical) ----
  for (int i = tid; i < N; i += kThreads) {
    const double x0 = (double)pts[i * 3], x1 = (double)pts[i * 3 + 1], x2 = (double)pts[i * 3 + 2];
    xs[i * 3] = x0;
    xs[i * 3 + 1] = x1;
    xs[i * 3 + 2] = x2;
    uint8_t b = (obs[i * 3 + 2] >= 0.0f) ? F_STEREO : 0;  // u_r < 0 => monocular (tracking_opt.cpp:67)
    const int k = asg[i];
    if (k >= 0 && pts_free) {
      if (a.is_deg[k]) {
        b |= F_DEG;
        const double* ax = a.axis + (size_t)k * 9;
        const double* mu = a.mean + (size_t)k * 3;
        const double e = dadd(dadd(dmul(ax[0], dsub(x0, mu[0])), dmul(ax[3], dsub(x1, mu[1]))), dmul(ax[6], dsub(x2, mu[2])));
        if (dmul(e, e) <= P.tri_str_thresh) b |= F_STR;  // gmmloc_opt.cpp:336-338, localization_opt.cpp:781
        // view-cos cull of planes seen closer than 12 deg to edge-on (GMM::renderView, gaussian_mixture.cpp:283-302)
        const double p0 = mu[0] - twc0, p1 = mu[1] - twc1, p2 = mu[2] - twc2;
        const double view_cos = fabs(p0 * ax[0] + p1 * ax[3] + p2 * ax[6]) / sqrt(p0 * p0 + p1 * p1 + p2 * p2);
        if (view_cos < view_cos_thresh) b &= ~F_STR;
      } else {
        if (chi2_exact(a.mean + (size_t)k * 3, a.cov_inv + (size_t)k * 9, x0, x1, x2) <= P.mdist2_gate) b |= F_STR;  // gmmloc_opt.cpp:230
      }
    }
    // a monocular point without a structure edge has a rank-2 Hll: held fixed
    if (pts_free && ((b & F_STEREO) || (b & F_STR))) b |= F_FREE;
    fl[i] = b;
  }
  __syncthreads();

  // Sweep schedule.  Iteration i of the reference is: linearise at state_i -> solve the reduced pose system ->
  // back-substitute the points (same linearisation point).  The back-substitution of iteration i and the
  // linearisation of iteration i+1 visit the same point with the same component constants, so they are FUSED
  // into one sweep: per point  [update x with dxi_i at (R_i, t_i)]  then  [accumulate at (R_{i+1}, t_{i+1}, x_{i+1})].
  // gn_iters + 1 sweeps instead of 2 gn_iters, half the component gathers and barriers; per-point arithmetic and
  // the reduction order are unchanged.
  //   sweep 0           : accumulate only (state_0)
  //   sweep 1..iters-1  : update with dxi_{s-1}, accumulate at state_s
  //   sweep iters       : update only
  const int iters = P.gn_iters;
  const int sweeps = (iters > 0) ? iters + 1 : 0;
  for (int sw = 0; sw < sweeps; ++sw) {
    const bool do_update = pts_free && sw > 0;
    const bool do_acc = pose_free && sw < iters;
    if (!do_update && !do_acc) continue;
    // s_Ro / s_to: linearisation point of the update (state sw-1); s_R / s_pose: state sw
    const Vec3 om = {s_dxi[0], s_dxi[1], s_dxi[2]}, up = {s_dxi[3], s_dxi[4], s_dxi[5]};
    double acc[kAcc];
#pragma unroll
    for (int i = 0; i < kAcc; ++i) acc[i] = 0.0;
    bool bad1 = false, bad2 = false;
    for (int i = tid; i < N; i += kThreads) {
      const uint8_t b = fl[i];
      const bool free_pt = b & F_FREE, str = b & F_STR;
      const bool upd = do_update && free_pt;
      const bool accp = do_acc && !(free_pt && !str);  // a free point without a structure edge contributes exactly zero
      if (!upd && !accp) continue;
      Vec3 xw = {xs[i * 3], xs[i * 3 + 1], xs[i * 3 + 2]};
      const float ob[3] = {obs[i * 3], obs[i * 3 + 1], obs[i * 3 + 2]};
      const double w = (double)P.s2inv[oct[i] & 7];
      StrConst sc;
      if (str) load_structure(a, asg[i], b, sc);
      if (upd) {
        if (!point_update(a, s_Ro, s_to, sc, b, om, up, ob, w, xw)) bad2 = true;
        xs[i * 3] = xw.x;
        xs[i * 3 + 1] = xw.y;
        xs[i * 3 + 2] = xw.z;
      }
      if (accp) {
        if (!pose_contribution(a, s_R, s_pose + 4, sc, b, xw, ob, w, acc)) bad1 = true;
      }
    }
    if (bad2) s_fail2 = 1;
    if (bad1) s_fail1 = 1;
    if (do_acc) {
#pragma unroll
      for (int i = 0; i < kAcc; ++i) {
        double s = acc[i];
#pragma unroll
        for (int off = 16; off > 0; off >>= 1) s += __shfl_xor_sync(0xffffffffu, s, off);
        if (lane == 0) s_red[warp][i] = s;
      }
    }
    __syncthreads();
    if (s_fail2) break;  // uniform: the update of iteration sw-1 failed
    if (do_acc) {
      if (tid < kAcc) {
        double s = s_red[0][tid];
#pragma unroll
        for (int w = 1; w < kWarps; ++w) s += s_red[w][tid];
        s_acc[tid] = s;
      }
      __syncthreads();
      if (tid == 0) {
        double dxi[6];
        if (s_fail1 || !solve6(s_acc, dxi)) {
          s_fail1 = 1;
        } else {
          double np[7];
          se3_oplus(dxi, s_pose, np);
          for (int i = 0; i < 9; ++i) s_Ro[i] = s_R[i];
          for (int i = 0; i < 3; ++i) s_to[i] = s_pose[4 + i];
          for (int i = 0; i < 6; ++i) s_dxi[i] = dxi[i];
          for (int i = 0; i < 7; ++i) s_pose[i] = np[i];
          quat_to_rot(np, s_R);
        }
      }
      __syncthreads();
      if (s_fail1) break;  // uniform
    }
  }
  __syncthreads();
  const bool fail = s_fail1 || s_fail2;
  __syncthreads();
  if (fail && tid == 0) {  // unsolved frame: back to the normalised input pose
    double q[7];
    for (int i = 0; i < 7; ++i) q[i] = a.pose0[(size_t)f * 7 + i];
    quat_normalize(q);
    for (int i = 0; i < 7; ++i) s_pose[i] = q[i];
    quat_to_rot(q, s_R);
  }
  __syncthreads();
  if (tid < 7) a.pose_out[(size_t)f * 7 + tid] = s_pose[tid];
  if (tid == 0) a.status[f] = fail ? GMMLOC_FRAME_NUMERIC : 0;
  if (a.pts_out != nullptr) {
    float* po = a.pts_out + (size_t)f * N * 3;
    for (int i = tid; i < N * 3; i += kThreads) po[i] = fail ? pts[i] : (float)xs[i];
  }
  if (a.outlier != nullptr) {
    // post-check at the final estimate: chi2 > 5.991 / 7.815 or non-positive depth
    // (localization_opt.cpp:857-881)
    double R[9], t[3];
#pragma unroll
    for (int i = 0; i < 9; ++i) R[i] = s_R[i];
#pragma unroll
    for (int i = 0; i < 3; ++i) t[i] = s_pose[4 + i];
    uint8_t* out = a.outlier + (size_t)f * N;
    for (int i = tid; i < N; i += kThreads) {
      Vec3 xw;
      if (fail) xw = {(double)pts[i * 3], (double)pts[i * 3 + 1], (double)pts[i * 3 + 2]};
      else xw = {xs[i * 3], xs[i * 3 + 1], xs[i * 3 + 2]};
      Vec3 pc = rot(R, xw);
      pc.x += t[0]; pc.y += t[1]; pc.z += t[2];
      const bool stereo = fl[i] & F_STEREO;
      const double invz = 1.0 / pc.z;
      const double u = pc.x * invz * P.fx + P.cx, v = pc.y * invz * P.fy + P.cy;
      const double e0 = (double)obs[i * 3] - u, e1 = (double)obs[i * 3 + 1] - v;
      double e2 = 0.0;
      if (stereo) e2 = (double)obs[i * 3 + 2] - (u - P.bf * invz);
      const double chi = (double)P.s2inv[oct[i] & 7] * (e0 * e0 + e1 * e1 + e2 * e2);
      const double th = stereo ? P.chi2_stereo : P.chi2_mono;
      out[i] = (chi > th || !(pc.z > 0.0)) ? 1 : 0;
    }
  }
}

}  // namespace

size_t frame_gn_smem_limit_points() { return (size_t)(200 * 1024) / 25; }

cudaError_t launch_frame_gn(const FrameArgs& a, cudaStream_t stream) {
  size_t smem = 0;
  if (a.xws == nullptr) {
    smem = (size_t)a.N * 25 + 16;
    cudaError_t e = cudaFuncSetAttribute(frame_gn_kernel, cudaFuncAttributeMaxDynamicSharedMemorySize, 208 * 1024);
    if (e != cudaSuccess) return e;
  }
  frame_gn_kernel<<<a.B, kThreads, smem, stream>>>(a);
  return cudaGetLastError();
}

}  // namespace gmmloc
