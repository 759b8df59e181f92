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

enum : uint8_t { F_STEREO = 1, F_STR = 2, F_FREE = 4, F_DEG = 8 };

// component constants of one structure edge, gathered ONCE per point and sweep (L2 latency is what bounds
// this kernel): planar => n = axis_.col(0) in c[0..2]; otherwise the 6 unique entries of cov_inv in c[0..5]
struct StrConst {
  double mu[3];
  double c[6];
};
__device__ __forceinline__ void load_structure(const FrameArgs& a, int k, uint8_t fl, StrConst& sc) {
  const double* mu = a.mean + (size_t)k * 3;
  sc.mu[0] = mu[0]; sc.mu[1] = mu[1]; sc.mu[2] = mu[2];
  if (fl & F_DEG) {
    const double* ax = a.axis + (size_t)k * 9;
    sc.c[0] = ax[0]; sc.c[1] = ax[3]; sc.c[2] = ax[6];  // axis_.col(0)
    sc.c[3] = sc.c[4] = sc.c[5] = 0.0;
  } else {
    const double* A = a.cov_inv + (size_t)k * 9;
    sc.c[0] = A[0]; sc.c[1] = A[1]; sc.c[2] = A[2]; sc.c[3] = A[4]; sc.c[4] = A[5]; sc.c[5] = A[8];
  }
}
__device__ __forceinline__ void structure(const FrameArgs& a, const double* R, const StrConst& sc, uint8_t fl, const Vec3& xw,
                                          Sym3& S, Vec3& s) {
  const Vec3 d = {xw.x - sc.mu[0], xw.y - sc.mu[1], xw.z - sc.mu[2]};
  if (fl & F_DEG) {
    const Vec3 n = {sc.c[0], sc.c[1], sc.c[2]};
    const double es = n.x * d.x + n.y * d.y + n.z * d.z;
    const Vec3 nc = rot(R, n);
    const double lam = a.prm.ba_lambda2;
    S = {lam * nc.x * nc.x, lam * nc.x * nc.y, lam * nc.x * nc.z, lam * nc.y * nc.y, lam * nc.y * nc.z, lam * nc.z * nc.z};
    const double c = -lam * es;
    s = {c * nc.x, c * nc.y, c * nc.z};
  } else {
    // EdgePt2Gaussian: J = L^T, Omega = I  =>  Hessian L L^T = cov_inv, rhs -cov_inv d
    const Sym3 Sw = {sc.c[0], sc.c[1], sc.c[2], sc.c[3], sc.c[4], sc.c[5]};
    const Vec3 bw = mul(Sw, d);
    s = rot(R, {-bw.x, -bw.y, -bw.z});
    // S_c = R Sw R^T
    double t[9];
#pragma unroll
    for (int r = 0; r < 3; ++r) {
      t[r * 3 + 0] = R[r * 3] * Sw.xx + R[r * 3 + 1] * Sw.xy + R[r * 3 + 2] * Sw.xz;
      t[r * 3 + 1] = R[r * 3] * Sw.xy + R[r * 3 + 1] * Sw.yy + R[r * 3 + 2] * Sw.yz;
      t[r * 3 + 2] = R[r * 3] * Sw.xz + R[r * 3 + 1] * Sw.yz + R[r * 3 + 2] * Sw.zz;
    }
    S.xx = t[0] * R[0] + t[1] * R[1] + t[2] * R[2];
    S.xy = t[0] * R[3] + t[1] * R[4] + t[2] * R[5];
    S.xz = t[0] * R[6] + t[1] * R[7] + t[2] * R[8];
    S.yy = t[3] * R[3] + t[4] * R[4] + t[5] * R[5];
    S.yz = t[3] * R[6] + t[4] * R[7] + t[5] * R[8];
    S.zz = t[6] * R[6] + t[7] * R[7] + t[8] * R[8];
  }
}

// Schur contribution of one point to the reduced pose system at linearisation point (R, t, xw)
// (pass-1 body): acc += A^T M A, A^T v.  Returns false when Hll is singular.
__device__ __forceinline__ bool pose_contribution(const FrameArgs& a, const double* R, const double* t, const StrConst& sc,
                                                  uint8_t b, const Vec3& xw, const float* ob, double w, double* acc) {
  const FrameParamsDev& P = a.prm;
  Vec3 pc = rot(R, xw);
  pc.x += t[0]; pc.y += t[1]; pc.z += t[2];
  Sym3 G;
  Vec3 h;
  reproj(P, pc, ob[0], ob[1], ob[2], b & F_STEREO, w, P.robust != 0, G, h);
  Sym3 M;
  Vec3 v;
  bool ok = true;
  if (!(b & F_FREE)) {
    M = G;
    v = h;
  } else {
    Sym3 S;
    Vec3 s;
    structure(a, R, sc, b, xw, S, s);
    const Sym3 D = {G.xx + S.xx, G.xy + S.xy, G.xz + S.xz, G.yy + S.yy, G.yz + S.yz, G.zz + S.zz};
    const Sym3 Di = inv_sym(D, ok);
    // T = G Di (general 3x3)
    const double T00 = G.xx * Di.xx + G.xy * Di.xy + G.xz * Di.xz, T01 = G.xx * Di.xy + G.xy * Di.yy + G.xz * Di.yz,
                 T02 = G.xx * Di.xz + G.xy * Di.yz + G.xz * Di.zz;
    const double T10 = G.xy * Di.xx + G.yy * Di.xy + G.yz * Di.xz, T11 = G.xy * Di.xy + G.yy * Di.yy + G.yz * Di.yz,
                 T12 = G.xy * Di.xz + G.yy * Di.yz + G.yz * Di.zz;
    const double T20 = G.xz * Di.xx + G.yz * Di.xy + G.zz * Di.xz, T21 = G.xz * Di.xy + G.yz * Di.yy + G.zz * Di.yz,
                 T22 = G.xz * Di.xz + G.yz * Di.yz + G.zz * Di.zz;
    const Vec3 bl = {h.x + s.x, h.y + s.y, h.z + s.z};
    v = {h.x - (T00 * bl.x + T01 * bl.y + T02 * bl.z), h.y - (T10 * bl.x + T11 * bl.y + T12 * bl.z),
         h.z - (T20 * bl.x + T21 * bl.y + T22 * bl.z)};
    // M = T S (symmetric in exact arithmetic; upper triangle taken)
    M.xx = T00 * S.xx + T01 * S.xy + T02 * S.xz;
    M.xy = T00 * S.xy + T01 * S.yy + T02 * S.yz;
    M.xz = T00 * S.xz + T01 * S.yz + T02 * S.zz;
    M.yy = T10 * S.xy + T11 * S.yy + T12 * S.yz;
    M.yz = T10 * S.xz + T11 * S.yz + T12 * S.zz;
    M.zz = T20 * S.xz + T21 * S.yz + T22 * S.zz;
  }
  accumulate_pose(acc, pc, M, v);
  return ok;
}

// point update at linearisation point (R, t, xw) for the pose step (om, up) (pass-2 body)
__device__ __forceinline__ bool point_update(const FrameArgs& a, const double* R, const double* t, const StrConst& sc, uint8_t b,
                                             const Vec3& om, const Vec3& up, const float* ob, double w, Vec3& xw) {
  const FrameParamsDev& P = a.prm;
  Vec3 pc = rot(R, xw);
  pc.x += t[0]; pc.y += t[1]; pc.z += t[2];
  Vec3 ad = cross(om, pc);  // A dxi = omega x pc + upsilon
  ad.x += up.x; ad.y += up.y; ad.z += up.z;
  // (a free stereo point without a structure edge has the closed form dc = P^-1 e - ad, but taking
  //  it as a separate branch diverges nearly every warp; measured slower than the uniform path)
  Sym3 G;
  Vec3 h;
  reproj(P, pc, ob[0], ob[1], ob[2], b & F_STEREO, w, P.robust != 0, G, h);
  Sym3 D = G;
  Vec3 bl = h;
  if (b & F_STR) {
    Sym3 S;
    Vec3 s;
    structure(a, R, sc, b, xw, S, s);
    D = {G.xx + S.xx, G.xy + S.xy, G.xz + S.xz, G.yy + S.yy, G.yz + S.yz, G.zz + S.zz};
    bl = {h.x + s.x, h.y + s.y, h.z + s.z};
  }
  bool ok;
  const Sym3 Di = inv_sym(D, ok);
  const Vec3 ga = mul(G, ad);
  const Vec3 dc = mul(Di, {bl.x - ga.x, bl.y - ga.y, bl.z - ga.z});
  const Vec3 dw = rot_t(R, dc);
  xw.x += dw.x; xw.y += dw.y; xw.z += dw.z;
  return ok && isfinite(xw.x) && isfinite(xw.y) && isfinite(xw.z);
}

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
  if (a.frame_list != nullptr) {  // only the frames the mixed-precision kernel routed here
    if (f >= *a.frame_count) return;
    f = a.frame_list[f];
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
