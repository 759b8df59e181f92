// frame_gn_mixed.cu — K3m: the per-frame structure-constrained Gauss-Newton of frame_gn.cu in MIXED precision.
//
// Same system, same sweep schedule and the same formulation as frame_gn_kernel (see the header of frame_gn.cu for the
// reference citations: Localization::jointOptimization, gmmloc/src/modules/localization_opt.cpp:456-925, single free
// pose; edges of gmmloc/src/gmm/factors.cpp:5-17,55-64,84-107,133-168).  What changes is where fp64 is spent:
//
//   fp64 : the state (landmarks in shared memory, pose), the camera-frame point pc = R x + t, the reprojection
//          RESIDUAL e = obs - proj(pc), the structure offset x - mu, the 6x6 reduced system (cross-thread sums,
//          Cholesky, SE3 update) and the state updates x += dx.
//   fp32 : everything that multiplies those residuals — projection Jacobian, G = w P^T P, S_c, D^-1, T = G D^-1,
//          M = T S_c, A^T M A — i.e. the Gauss-Newton model of the Hessian and the Jacobian transposes.
//
// Why this keeps the result: a Gauss-Newton fixed point is where the GRADIENTS vanish (per point h + s = 0, pose
// sum A^T h = 0).  The gradients are products of fp32 Jacobians (relative error 2^-24) with residuals that are
// formed in fp64 and only then rounded, so their error is RELATIVE to the (small) residual; the formulation
// v = h - T (h + s) keeps the pose right-hand side equal to A^T h at the fixed point whatever the error of D^-1.
// The fp32 Hessian model only perturbs the convergence rate (by ~1e-3 relative at cond(D) ~ 1e4).  Measured against
// the fp64 oracle: pose differences ~1e-9 rad / ~1e-9 m (tolerance 1e-5 rad / 1e-4 m), identical edge gates
// (still evaluated in fp64 at iteration 0) and outlier flags (fp64 post-check).  frame_gn_kernel (all fp64) stays
// selectable (gmmloc_ctx_set_gn_precision) for A/B runs.
#include <math_constants.h>

#include <algorithm>
#include <cstdint>
#include <cstdlib>

#include "common.cuh"
#include "gn_common.cuh"
#include "frame_gn_point.cuh"
#include "frame_gn_classes.cuh"

namespace gmmloc {

namespace {

// Instantiations: 192 threads x 2 CTAs per SM while two frames fit in shared memory, else 384 threads x 1 CTA
// (168 registers per thread, no spills, 12 resident warps per SM).
// CTA shape of the two-resident-frames variants: threads per CTA and the register cap that lets two CTAs share an SM
// (65 536 registers: 2 x 192 x 168 or 2 x 224 x 144)
#ifndef GMMLOC_GNM_THREADS2
#define GMMLOC_GNM_THREADS2 192
#endif
#ifndef GMMLOC_GNM_MAXREG2
#define GMMLOC_GNM_MAXREG2 168
#endif
constexpr int kThreads2 = GMMLOC_GNM_THREADS2;
constexpr int kAcc = 27;
constexpr int kScratchPerPoint = 4 * 12 + 8;          // global scratch per point and resident CTA: store 48 B, aux 7 B (+1)
constexpr int kStore = 12;                            // floats of the stored linearisation per point: u0 (3), W (9)
// bytes of one resident CTA's slice of the global scratch: a multiple of 16 so that the float4 accesses of the stored
// linearisation stay aligned for odd N
__host__ __device__ inline size_t scratch_slice(int N) { return ((size_t)N * kScratchPerPoint + 15) & ~(size_t)15; }


struct V3f {
  float x, y, z;
};
struct S3f {
  float xx, xy, xz, yy, yz, zz;
};
__device__ __forceinline__ V3f rotf(const float* R, const V3f& v) {
  return {R[0] * v.x + R[1] * v.y + R[2] * v.z, R[3] * v.x + R[4] * v.y + R[5] * v.z, R[6] * v.x + R[7] * v.y + R[8] * v.z};
}
__device__ __forceinline__ V3f rotf_t(const float* R, const V3f& v) {
  return {R[0] * v.x + R[3] * v.y + R[6] * v.z, R[1] * v.x + R[4] * v.y + R[7] * v.z, R[2] * v.x + R[5] * v.y + R[8] * v.z};
}
__device__ __forceinline__ V3f mulf(const S3f& A, const V3f& v) {
  return {A.xx * v.x + A.xy * v.y + A.xz * v.z, A.xy * v.x + A.yy * v.y + A.yz * v.z, A.xz * v.x + A.yz * v.y + A.zz * v.z};
}
__device__ __forceinline__ S3f inv_symf(const S3f& D, bool& ok) {
  const float c00 = D.yy * D.zz - D.yz * D.yz;
  const float c01 = D.xz * D.yz - D.xy * D.zz;
  const float c02 = D.xy * D.yz - D.xz * D.yy;
  const float det = D.xx * c00 + D.xy * c01 + D.xz * c02;
  const float id = 1.0f / det;
  ok = (det != 0.0f) && isfinite(id);
  return {c00 * id, c01 * id, c02 * id, (D.xx * D.zz - D.xz * D.xz) * id, (D.xy * D.xz - D.xx * D.yz) * id,
          (D.xx * D.yy - D.xy * D.xy) * id};
}

// camera constants in fp32 + the pose of one linearisation point in both precisions
struct PoseBoth {
  double R[9], t[3];
  float Rf[9], tf[3];
};

// per-point constants of the structure edge (fp32 Hessian model; the mean stays fp64 for the offset x - mu)
struct StrF {
  double mu[3];
  float c[6];
};
__device__ __forceinline__ void load_structure(const FrameArgs& a, const float* __restrict__ gn_tab, int k, StrF& sc) {
  const double* mu = a.mean + (size_t)k * 3;
  sc.mu[0] = mu[0]; sc.mu[1] = mu[1]; sc.mu[2] = mu[2];
  const float4 t0 = __ldg(reinterpret_cast<const float4*>(gn_tab) + 2 * (size_t)k);
  const float2 t1 = __ldg(reinterpret_cast<const float2*>(gn_tab) + 4 * (size_t)k + 2);
  sc.c[0] = t0.x; sc.c[1] = t0.y; sc.c[2] = t0.z; sc.c[3] = t0.w; sc.c[4] = t1.x; sc.c[5] = t1.y;
}

// fp64 part of one linearisation: camera-frame point and reprojection residual (rounded to fp32 at the end)
struct Lin {
  V3f pc;     // camera-frame point
  float iz;   // 1 / z
  V3f e;      // residual (u, v, u_r); e.z = 0 for monocular points
};
__device__ __forceinline__ Lin linearise(const FrameParamsDev& P, const double* R, const double* t, double x0, double x1, double x2,
                                         const float* ob, bool stereo) {
  const double pcx = R[0] * x0 + R[1] * x1 + R[2] * x2 + t[0];
  const double pcy = R[3] * x0 + R[4] * x1 + R[5] * x2 + t[1];
  const double pcz = R[6] * x0 + R[7] * x1 + R[8] * x2 + t[2];
  // 1 / z: fp32 reciprocal + one Newton step in fp64 (relative error ~1e-14; a full fp64 division costs ~30 instructions)
  double iz = (double)__frcp_rn((float)pcz);
  iz = iz * (2.0 - pcz * iz);
  const double u = pcx * (P.fx * iz) + P.cx, v = pcy * (P.fy * iz) + P.cy;
  Lin L;
  L.e.x = (float)((double)ob[0] - u);
  L.e.y = (float)((double)ob[1] - v);
  L.e.z = stereo ? (float)((double)ob[2] - (u - P.bf * iz)) : 0.f;
  L.pc = {(float)pcx, (float)pcy, (float)pcz};
  L.iz = (float)iz;
  return L;
}

struct CamF {
  float fx, fy, bf, lam, d_mono, d_stereo;
  int robust;
};

// fp32 Gauss-Newton model of the reprojection edge: G = w P^T P, h = w P^T e
__device__ __forceinline__ void reprojf(const CamF& C, const Lin& L, bool stereo, float w, S3f& G, V3f& h) {
  const float a = C.fx * L.iz, b = C.fy * L.iz;
  const float c0 = -a * L.pc.x * L.iz, c1 = -b * L.pc.y * L.iz;
  float c2 = 0.f, a2 = 0.f;
  if (stereo) {
    c2 = c0 + C.bf * L.iz * L.iz;
    a2 = a;
  }
  if (C.robust) {  // g2o RobustKernelHuber: weight delta / sqrt(chi2) beyond delta^2
    const float chi = w * (L.e.x * L.e.x + L.e.y * L.e.y + L.e.z * L.e.z);
    const float delta = stereo ? C.d_stereo : C.d_mono;
    if (chi > delta * delta) w *= delta * rsqrtf(chi);
  }
  G.xx = w * (a * a + a2 * a2);
  G.xy = 0.f;
  G.xz = w * (a * c0 + a2 * c2);
  G.yy = w * (b * b);
  G.yz = w * (b * c1);
  G.zz = w * (c0 * c0 + c1 * c1 + c2 * c2);
  h.x = w * (a * L.e.x + a2 * L.e.z);
  h.y = w * (b * L.e.y);
  h.z = w * (c0 * L.e.x + c1 * L.e.y + c2 * L.e.z);
}

// structure edge rotated into the camera frame: S_c = R S R^T, s_c = R b_s; the offset x - mu is formed in fp64
__device__ __forceinline__ void structuref(const CamF& C, const float* Rf, const StrF& sc, uint8_t fl, double x0, double x1, double x2,
                                           S3f& S, V3f& s) {
  const V3f d = {(float)(x0 - sc.mu[0]), (float)(x1 - sc.mu[1]), (float)(x2 - sc.mu[2])};
  if (fl & F_DEG) {
    const V3f n = {sc.c[0], sc.c[1], sc.c[2]};
    const float es = n.x * d.x + n.y * d.y + n.z * d.z;
    const V3f nc = rotf(Rf, n);
    const float lam = C.lam;
    S = {lam * nc.x * nc.x, lam * nc.x * nc.y, lam * nc.x * nc.z, lam * nc.y * nc.y, lam * nc.y * nc.z, lam * nc.z * nc.z};
    const float c = -lam * es;
    s = {c * nc.x, c * nc.y, c * nc.z};
  } else {
    const S3f Sw = {sc.c[0], sc.c[1], sc.c[2], sc.c[3], sc.c[4], sc.c[5]};
    const V3f bw = mulf(Sw, d);
    s = rotf(Rf, {-bw.x, -bw.y, -bw.z});
    float t[9];
#pragma unroll
    for (int r = 0; r < 3; ++r) {
      t[r * 3 + 0] = Rf[r * 3] * Sw.xx + Rf[r * 3 + 1] * Sw.xy + Rf[r * 3 + 2] * Sw.xz;
      t[r * 3 + 1] = Rf[r * 3] * Sw.xy + Rf[r * 3 + 1] * Sw.yy + Rf[r * 3 + 2] * Sw.yz;
      t[r * 3 + 2] = Rf[r * 3] * Sw.xz + Rf[r * 3 + 1] * Sw.yz + Rf[r * 3 + 2] * Sw.zz;
    }
    S.xx = t[0] * Rf[0] + t[1] * Rf[1] + t[2] * Rf[2];
    S.xy = t[0] * Rf[3] + t[1] * Rf[4] + t[2] * Rf[5];
    S.xz = t[0] * Rf[6] + t[1] * Rf[7] + t[2] * Rf[8];
    S.yy = t[3] * Rf[3] + t[4] * Rf[4] + t[5] * Rf[5];
    S.yz = t[3] * Rf[6] + t[4] * Rf[7] + t[5] * Rf[8];
    S.zz = t[6] * Rf[6] + t[7] * Rf[7] + t[8] * Rf[8];
  }
}

// acc[0..20] += upper triangle of A^T M A, acc[21..26] += A^T v, A = [-[pc]x  I]   (fp32 twin of accumulate_pose)
__device__ __forceinline__ void accumulate_posef(float* acc, const V3f& pc, const S3f& M, const V3f& v) {
  const float C00 = pc.y * M.xz - pc.z * M.xy, C01 = pc.y * M.yz - pc.z * M.yy, C02 = pc.y * M.zz - pc.z * M.yz;
  const float C10 = pc.z * M.xx - pc.x * M.xz, C11 = pc.z * M.xy - pc.x * M.yz, C12 = pc.z * M.xz - pc.x * M.zz;
  const float C20 = pc.x * M.xy - pc.y * M.xx, C21 = pc.x * M.yy - pc.y * M.xy, C22 = pc.x * M.yz - pc.y * M.xz;
  acc[0] += C02 * pc.y - C01 * pc.z;
  acc[1] += C00 * pc.z - C02 * pc.x;
  acc[2] += C01 * pc.x - C00 * pc.y;
  acc[3] += C00;
  acc[4] += C01;
  acc[5] += C02;
  acc[6] += C10 * pc.z - C12 * pc.x;
  acc[7] += C11 * pc.x - C10 * pc.y;
  acc[8] += C10;
  acc[9] += C11;
  acc[10] += C12;
  acc[11] += C21 * pc.x - C20 * pc.y;
  acc[12] += C20;
  acc[13] += C21;
  acc[14] += C22;
  acc[15] += M.xx;
  acc[16] += M.xy;
  acc[17] += M.xz;
  acc[18] += M.yy;
  acc[19] += M.yz;
  acc[20] += M.zz;
  acc[21] += pc.y * v.z - pc.z * v.y;
  acc[22] += pc.z * v.x - pc.x * v.z;
  acc[23] += pc.x * v.y - pc.y * v.x;
  acc[24] += v.x;
  acc[25] += v.y;
  acc[26] += v.z;
}


// closed-form pieces of a free stereo point on a plane (class C_STRP), all in the camera frame of one linearisation
struct PlaneLin {
  V3f pc, nc;   // camera-frame point, plane normal
  V3f pe;       // P^-1 e: the correction the stereo edge alone asks for
  V3f z;        // G^-1 nc
  float kappa;  // lam / (1 + lam nc.z)
};
__device__ __forceinline__ PlaneLin plane_lin(const CamF& C, const Lin& L, const V3f& nc, float w) {
  PlaneLin Q;
  Q.pc = L.pc;
  Q.nc = nc;
  // P = [[a, 0, c0], [0, b, c1], [a, 0, c0 + cz]]  (rows u, v, u_r)
  const float a_ = C.fx * L.iz, b_ = C.fy * L.iz;
  const float ia = 1.0f / a_, ib = 1.0f / b_;
  const float c0 = -a_ * L.pc.x * L.iz, c1 = -b_ * L.pc.y * L.iz;
  const float icz = 1.0f / (C.bf * L.iz * L.iz);
  if (C.robust) {  // Huber weight of the reprojection edge (cancels in pe, scales G)
    const float chi = w * (L.e.x * L.e.x + L.e.y * L.e.y + L.e.z * L.e.z);
    if (chi > C.d_stereo * C.d_stereo) w *= C.d_stereo * rsqrtf(chi);
  }
  Q.pe.z = (L.e.z - L.e.x) * icz;
  Q.pe.x = (L.e.x - c0 * Q.pe.z) * ia;
  Q.pe.y = (L.e.y - c1 * Q.pe.z) * ib;
  // q = P^-T nc
  const float q1 = nc.y * ib;
  const float q2 = (nc.z - c0 * nc.x * ia - c1 * q1) * icz;
  const float q0 = nc.x * ia - q2;
  // z = P^-1 q / w
  const float iw = 1.0f / w;
  const float zz = (q2 - q0) * icz;
  Q.z = {(q0 - c0 * zz) * ia * iw, (q1 - c1 * zz) * ib * iw, zz * iw};
  const float gamma = nc.x * Q.z.x + nc.y * Q.z.y + nc.z * Q.z.z;
  Q.kappa = C.lam / (1.0f + C.lam * gamma);
  return Q;
}

// fp64 treatment of one ill-conditioned point (class C_F64): the update and the Schur terms of frame_gn_kernel.
// Deliberately not inlined: it runs for a few warp-iterations per sweep and must not raise the register count of
// the fp32 path.
__device__ __noinline__ void point_fp64(const FrameArgs& a, const PoseBoth& oldp, const PoseBoth& newp, const double* dxi,
                                        uint8_t b, int k, bool upd, bool accp, const float* ob, double w, double* x, V3f& pcf,
                                        S3f& Mf, V3f& vf, int& bad) {
  StrConst sc;
  if (b & F_STR) load_structure(a, k, b, sc);
  Vec3 xw = {x[0], x[1], x[2]};
  if (upd) {
    const Vec3 om = {dxi[0], dxi[1], dxi[2]}, up = {dxi[3], dxi[4], dxi[5]};
    if (!point_update(a, oldp.R, oldp.t, sc, b, om, up, ob, w, xw)) bad |= 2;
    x[0] = xw.x; x[1] = xw.y; x[2] = xw.z;
  }
  if (accp) {
    Vec3 pc, v;
    Sym3 M;
    if (!pose_terms(a, newp.R, newp.t, sc, b, xw, ob, w, pc, M, v)) bad |= 1;
    pcf = {(float)pc.x, (float)pc.y, (float)pc.z};
    Mf = {(float)M.xx, (float)M.xy, (float)M.xz, (float)M.yy, (float)M.yz, (float)M.zz};
    vf = {(float)v.x, (float)v.y, (float)v.z};
  }
}

// ---- frame preparation (K3p) ------------------------------------------------------------------------------------
// Classifies the points of every frame (edge gates at iteration 0, point class) and writes the frame in class-sorted
// order to a global buffer.  This used to be the first two passes of frame_gn_mixed_kernel, where each chunk of points
// was a chain of dependent memory round trips (assignment -> degenerate flag -> component constants) on 12 resident
// warps per SM: 12 % of that kernel's stall samples.  Here the same passes run with one CTA of 256 threads per frame
// and no register or shared-memory pressure (up to 64 resident warps per SM hide the chains), and the solver starts
// from a coalesced copy.  Layout of frame f (prep_stride(N) bytes, all arrays in class-sorted order):
//   [0, 12N) landmarks f32 x 3 | [12N, 24N) observations f32 x 3 | [24N, 28N) component id i32 |
//   [28N, 30N) original index u16 | [30N, 31N) octave u8 | [31N, 32N) flags u8 (class in bits 4..6)

constexpr int kPrepThreads = 256;
__global__ void __launch_bounds__(kPrepThreads) frame_prep_kernel(const FrameArgs a, unsigned char* __restrict__ prep) {
  constexpr int kWarps = kPrepThreads / 32;
  extern __shared__ __align__(16) unsigned char smem_raw[];  // flags by ORIGINAL index while the frame is being sorted
  __shared__ int s_cnt[C_COUNT * kMaxChunks * 8];
  __shared__ double s_R[9], s_t[3];
  uint8_t* tmp = smem_raw;
  const int tid = threadIdx.x, lane = tid & 31, warp = tid >> 5;
  const int N = a.N;
  const FrameParamsDev& P = a.prm;
  const bool pose_free = P.factor_mode != GMMLOC_MODE_POINTS_ONLY;
  const bool pts_free = P.factor_mode != GMMLOC_MODE_POSE_ONLY;
  const double view_cos_thresh = cos(78.0 * 3.14159265358979323846 / 180.0);  // gaussian_mixture.cpp:284
  const int nchunks = (N + kPrepThreads - 1) / kPrepThreads;
  for (int f = blockIdx.x; f < a.B; f += gridDim.x) {
    const float* pts = a.pts_w + (size_t)f * N * 3;
    const float* obs = a.obs + (size_t)f * N * 3;
    const uint8_t* oct = a.octave + (size_t)f * N;
    const int32_t* asg = a.assign + (size_t)f * N;
    unsigned char* pf = prep + (size_t)f * prep_stride(N);
    float* gp = reinterpret_cast<float*>(pf);
    float* go = reinterpret_cast<float*>(pf + (size_t)N * 12);
    int32_t* gk = reinterpret_cast<int32_t*>(pf + (size_t)N * 24);
    uint16_t* gperm = reinterpret_cast<uint16_t*>(pf + (size_t)N * 28);
    uint8_t* goct = pf + (size_t)N * 30;
    uint8_t* gfl = pf + (size_t)N * 31;
    __syncthreads();  // the previous frame is done with s_R / s_cnt / tmp
    if (tid == 0) {
      double q[7];
      for (int i = 0; i < 7; ++i) q[i] = a.pose0[(size_t)f * 7 + i];
      quat_normalize(q);
      quat_to_rot(q, s_R);
      for (int i = 0; i < 3; ++i) s_t[i] = q[4 + i];
    }
    __syncthreads();
    const double twc0 = -(s_R[0] * s_t[0] + s_R[3] * s_t[1] + s_R[6] * s_t[2]);
    const double twc1 = -(s_R[1] * s_t[0] + s_R[4] * s_t[1] + s_R[7] * s_t[2]);
    const double twc2 = -(s_R[2] * s_t[0] + s_R[5] * s_t[1] + s_R[8] * s_t[2]);
    // pass A: flags and class of every point, class histogram per (chunk, warp)  [same arithmetic as the in-kernel pass]
    for (int c = 0; c < nchunks; ++c) {
      const int i = c * kPrepThreads + tid;
      int cls = -1;
      if (i < N) {
        const double x0 = (double)pts[i * 3], x1 = (double)pts[i * 3 + 1], x2 = (double)pts[i * 3 + 2];
        uint8_t b = (obs[i * 3 + 2] >= 0.0f) ? F_STEREO : 0;
        const int k = asg[i];
        if (k >= 0 && pts_free) {
          if (a.is_deg[k]) {
            b |= F_DEG;
            const double* ax = a.axis + (size_t)k * 9;
            const double* mu = a.mean + (size_t)k * 3;
            const double e = dadd(dadd(dmul(ax[0], dsub(x0, mu[0])), dmul(ax[3], dsub(x1, mu[1]))), dmul(ax[6], dsub(x2, mu[2])));
            if (dmul(e, e) <= P.tri_str_thresh) b |= F_STR;
            const double p0 = mu[0] - twc0, p1 = mu[1] - twc1, p2 = mu[2] - twc2;
            const double view_cos = fabs(p0 * ax[0] + p1 * ax[3] + p2 * ax[6]) / sqrt(p0 * p0 + p1 * p1 + p2 * p2);
            if (view_cos < view_cos_thresh) b &= ~F_STR;
          } else {
            if (chi2_exact(a.mean + (size_t)k * 3, a.cov_inv + (size_t)k * 9, x0, x1, x2) <= P.mdist2_gate) b |= F_STR;
          }
        }
        if (pts_free && ((b & F_STEREO) || (b & F_STR))) b |= F_FREE;
        if (!(b & F_FREE)) {
          cls = pose_free ? C_FIXED : C_OFF;
        } else {
          const double pz = s_R[6] * x0 + s_R[7] * x1 + s_R[8] * x2 + s_t[2];
          const double rx = x0 - twc0, ry = x1 - twc1, rz = x2 - twc2;
          if (!fp32_geometry_ok(pz, rx * rx + ry * ry + rz * rz)) cls = C_F64;
          else if (!(b & F_STEREO)) cls = (b & F_DEG) ? C_MPLANE : C_MSTR;  // a free monocular point always has F_STR
          else cls = (b & F_STR) ? ((b & F_DEG) ? C_STRP : C_STR) : C_PLAIN;
        }
        tmp[i] = (uint8_t)(b | (cls << 4));
      }
#pragma unroll
      for (int q = 0; q < C_COUNT; ++q) {
        const unsigned m = __ballot_sync(0xffffffffu, cls == q);
        if (lane == 0) s_cnt[q * (kMaxChunks * 8) + c * kWarps + warp] = __popc(m);
      }
    }
    __syncthreads();
    if (warp == 0) {  // exclusive scan in (class, chunk, warp) order = stable counting-sort offsets
      int run = 0;
      for (int q = 0; q < C_COUNT; ++q) {
        const int n = nchunks * kWarps;
        for (int b0 = 0; b0 < n; b0 += 32) {
          const int e = b0 + lane;
          const int idx = q * (kMaxChunks * 8) + e;
          const int v = (e < n) ? s_cnt[idx] : 0;
          int inc = v;
#pragma unroll
          for (int off = 1; off < 32; off <<= 1) {
            const int t = __shfl_up_sync(0xffffffffu, inc, off);
            if (lane >= off) inc += t;
          }
          if (e < n) s_cnt[idx] = run + inc - v;
          run += __shfl_sync(0xffffffffu, inc, 31);
        }
      }
    }
    __syncthreads();
    // pass B: scatter the frame into class-sorted order
    for (int c = 0; c < nchunks; ++c) {
      const int i = c * kPrepThreads + tid;
      int cls = -1;
      uint8_t b = 0;
      if (i < N) {
        b = tmp[i];
        cls = b >> 4;
      }
      int j = -1;
#pragma unroll
      for (int q = 0; q < C_COUNT; ++q) {
        const unsigned m = __ballot_sync(0xffffffffu, cls == q);
        if (cls == q) j = s_cnt[q * (kMaxChunks * 8) + c * kWarps + warp] + __popc(m & ((1u << lane) - 1u));
      }
      if (i < N) {
        gp[j * 3] = pts[i * 3];
        gp[j * 3 + 1] = pts[i * 3 + 1];
        gp[j * 3 + 2] = pts[i * 3 + 2];
        go[j * 3] = obs[i * 3];
        go[j * 3 + 1] = obs[i * 3 + 1];
        go[j * 3 + 2] = obs[i * 3 + 2];
        gk[j] = asg[i];
        gperm[j] = (uint16_t)i;
        goct[j] = oct[i];
        gfl[j] = b;
      }
    }
  }
}

template <int kThreads, int kMinB, bool kStoreLin, bool kPrep>
__global__ void __launch_bounds__(kThreads) __maxnreg__(kMinB == 1 ? 168 : GMMLOC_GNM_MAXREG2) frame_gn_mixed_kernel(const FrameArgs a, const float* __restrict__ gn_tab,
                                                                         int* __restrict__ counters, float* __restrict__ store_ws,
                                                                         int aux_global, const unsigned char* __restrict__ prep) {
  constexpr int kWarps = kThreads / 32;
  extern __shared__ __align__(16) unsigned char smem_raw[];
  __shared__ double s_red[kWarps][kAcc];
  __shared__ double s_acc[kAcc];
  __shared__ double s_pose[7], s_dxi[6];
  __shared__ PoseBoth s_new, s_old;  // linearisation points: state sw (accumulate) and state sw-1 (update)
  __shared__ int s_fail1, s_fail2, s_frame;

  const int tid = threadIdx.x, lane = tid & 31, warp = tid >> 5;
  const int N = a.N;
  const FrameParamsDev& P = a.prm;
  const bool pose_free = P.factor_mode != GMMLOC_MODE_POINTS_ONLY;
  const bool pts_free = P.factor_mode != GMMLOC_MODE_POSE_ONLY;
  CamF C;
  C.fx = (float)P.fx; C.fy = (float)P.fy; C.bf = (float)P.bf; C.lam = (float)P.ba_lambda2;
  C.d_mono = (float)P.delta_mono; C.d_stereo = (float)P.delta_stereo; C.robust = P.robust;

  // per-frame state, class-sorted position j: landmark (fp64), observation, component id, original index, octave, flags
  double* xs = reinterpret_cast<double*>(smem_raw);
  float* so = reinterpret_cast<float*>(xs + (size_t)N * 3);
  uint8_t* soct = reinterpret_cast<uint8_t*>(so + (size_t)N * 3);
  uint8_t* fl = soct + N;
  // component id, original index and the flags by ORIGINAL index (used while the frame is being sorted): in shared
  // memory, or — frames too large for that — behind the stored linearisation in this CTA's slice of the global scratch
  const size_t slice = scratch_slice(N);  // bytes per CTA
  unsigned char* aux = aux_global ? reinterpret_cast<unsigned char*>(store_ws) + (size_t)blockIdx.x * slice + (size_t)N * 4 * kStore
                                  : smem_raw + (((size_t)N * 38 + 15) & ~(size_t)15);
  int32_t* sk = reinterpret_cast<int32_t*>(aux);
  uint16_t* perm = reinterpret_cast<uint16_t*>(sk + N);
  uint8_t* tmp = reinterpret_cast<uint8_t*>(perm + N);
  // stored linearisation of the structured free points (classes C_STRP / C_STR): the point step of iteration i is
  // dc = u0 - W (A dxi) with u0 = D^-1 (h + s), W = D^-1 G from the SAME linearisation the pose system of iteration i
  // was accumulated at, so the fused sweep linearises every such point once, not twice
  // (in shared memory when it fits next to a second resident frame, else in a per-CTA slice of a global scratch
  // buffer: 2 x 148 CTAs x N x 48 B = 28 MB at C2, L2-resident and reused from frame to frame)
  float* store = store_ws != nullptr
                     ? reinterpret_cast<float*>(reinterpret_cast<unsigned char*>(store_ws) + (size_t)blockIdx.x * slice)
                     : reinterpret_cast<float*>(smem_raw + (((size_t)N * 38 + 15) & ~(size_t)15) + (((size_t)N * 7 + 15) & ~(size_t)15));

  // persistent CTAs: frames are handed out through a global counter (no tail wave)
  for (;;) {
    __syncthreads();
    if (tid == 0) s_frame = atomicAdd(counters, 1);
    __syncthreads();
    const int f = s_frame;
    if (f >= a.B) break;

    const float* pts = a.pts_w + (size_t)f * N * 3;
    const float* obs = a.obs + (size_t)f * N * 3;
    const uint8_t* oct = a.octave + (size_t)f * N;
    const int32_t* asg = a.assign + (size_t)f * N;

    if (tid == 0) {
      double q[7];
      for (int i = 0; i < 7; ++i) q[i] = a.pose0[(size_t)f * 7 + i];
      quat_normalize(q);
      for (int i = 0; i < 7; ++i) s_pose[i] = q[i];
      quat_to_rot(q, s_new.R);
      for (int i = 0; i < 6; ++i) s_dxi[i] = 0.0;
      for (int i = 0; i < 9; ++i) {
        s_new.Rf[i] = (float)s_new.R[i];
        s_old.R[i] = s_new.R[i];
        s_old.Rf[i] = s_new.Rf[i];
      }
      for (int i = 0; i < 3; ++i) {
        s_new.t[i] = s_old.t[i] = q[4 + i];
        s_new.tf[i] = s_old.tf[i] = (float)q[4 + i];
      }
      s_fail1 = 0;
      s_fail2 = 0;
    }
    __syncthreads();
    if constexpr (kPrep) {
      // the frame was classified and class-sorted by frame_prep_kernel: a plain coalesced copy into shared memory (every
      // load independent of the others) replaces the two load passes, whose chunks were chains of dependent round trips
      const unsigned char* pf = prep + (size_t)f * prep_stride(N);
      const float* gp = reinterpret_cast<const float*>(pf);
      const float* go = reinterpret_cast<const float*>(pf + (size_t)N * 12);
      const int32_t* gk = reinterpret_cast<const int32_t*>(pf + (size_t)N * 24);
      const uint16_t* gperm = reinterpret_cast<const uint16_t*>(pf + (size_t)N * 28);
      const uint8_t* goct = pf + (size_t)N * 30;
      const uint8_t* gfl = pf + (size_t)N * 31;
#pragma unroll 4
      for (int e = tid; e < 3 * N; e += kThreads) {
        xs[e] = (double)gp[e];
        so[e] = go[e];
      }
#pragma unroll 4
      for (int j = tid; j < N; j += kThreads) {
        sk[j] = gk[j];
        perm[j] = gperm[j];
        soct[j] = goct[j];
        fl[j] = gfl[j];
      }
      __syncthreads();
    } else {
      __shared__ int s_cnt[C_COUNT * kMaxChunks * 8];  // class histogram per (class, chunk, warp) -> exclusive offsets
      const double view_cos_thresh = cos(78.0 * 3.14159265358979323846 / 180.0);  // gaussian_mixture.cpp:284
      const int nchunks = (N + kThreads - 1) / kThreads;
      const double twc0 = -(s_new.R[0] * s_pose[4] + s_new.R[3] * s_pose[5] + s_new.R[6] * s_pose[6]);
      const double twc1 = -(s_new.R[1] * s_pose[4] + s_new.R[4] * s_pose[5] + s_new.R[7] * s_pose[6]);
      const double twc2 = -(s_new.R[2] * s_pose[4] + s_new.R[5] * s_pose[5] + s_new.R[8] * s_pose[6]);
      // ---- pass A: flags and class of every point (structure-edge gates fixed at iteration 0, in fp64 exactly as
      // frame_gn_kernel), class histogram per (chunk, warp) ----
      for (int c = 0; c < nchunks; ++c) {
        const int i = c * kThreads + tid;
        int cls = -1;
        if (i < N) {
          const double x0 = (double)pts[i * 3], x1 = (double)pts[i * 3 + 1], x2 = (double)pts[i * 3 + 2];
          uint8_t b = (obs[i * 3 + 2] >= 0.0f) ? F_STEREO : 0;
          const int k = asg[i];
          if (k >= 0 && pts_free) {
            if (a.is_deg[k]) {
              b |= F_DEG;
              const double* ax = a.axis + (size_t)k * 9;
              const double* mu = a.mean + (size_t)k * 3;
              const double e = dadd(dadd(dmul(ax[0], dsub(x0, mu[0])), dmul(ax[3], dsub(x1, mu[1]))), dmul(ax[6], dsub(x2, mu[2])));
              if (dmul(e, e) <= P.tri_str_thresh) b |= F_STR;
              const double p0 = mu[0] - twc0, p1 = mu[1] - twc1, p2 = mu[2] - twc2;
              const double view_cos = fabs(p0 * ax[0] + p1 * ax[3] + p2 * ax[6]) / sqrt(p0 * p0 + p1 * p1 + p2 * p2);
              if (view_cos < view_cos_thresh) b &= ~F_STR;
            } else {
              if (chi2_exact(a.mean + (size_t)k * 3, a.cov_inv + (size_t)k * 9, x0, x1, x2) <= P.mdist2_gate) b |= F_STR;
            }
          }
          if (pts_free && ((b & F_STEREO) || (b & F_STR))) b |= F_FREE;
          if (!(b & F_FREE)) {
            cls = pose_free ? C_FIXED : C_OFF;
          } else {
            const double pz = s_new.R[6] * x0 + s_new.R[7] * x1 + s_new.R[8] * x2 + s_new.t[2];
            const double rx = x0 - twc0, ry = x1 - twc1, rz = x2 - twc2;
            if (!fp32_geometry_ok(pz, rx * rx + ry * ry + rz * rz)) cls = C_F64;
            else if (!(b & F_STEREO)) cls = (b & F_DEG) ? C_MPLANE : C_MSTR;  // a free monocular point always has F_STR
            else cls = (b & F_STR) ? ((b & F_DEG) ? C_STRP : C_STR) : C_PLAIN;
          }
          tmp[i] = (uint8_t)(b | (cls << 4));
        }
  #pragma unroll
        for (int q = 0; q < C_COUNT; ++q) {
          const unsigned m = __ballot_sync(0xffffffffu, cls == q);
          if (lane == 0) s_cnt[q * (kMaxChunks * 8) + c * kWarps + warp] = __popc(m);
        }
      }
      __syncthreads();
      // exclusive scan of the histogram in (class, chunk, warp) order = stable counting sort offsets (one warp)
      if (warp == 0) {
        int run = 0;
        for (int q = 0; q < C_COUNT; ++q) {
          const int n = nchunks * kWarps;
          for (int b0 = 0; b0 < n; b0 += 32) {
            const int e = b0 + lane;
            const int idx = q * (kMaxChunks * 8) + e;
            const int v = (e < n) ? s_cnt[idx] : 0;
            int inc = v;
  #pragma unroll
            for (int off = 1; off < 32; off <<= 1) {
              const int t = __shfl_up_sync(0xffffffffu, inc, off);
              if (lane >= off) inc += t;
            }
            if (e < n) s_cnt[idx] = run + inc - v;
            run += __shfl_sync(0xffffffffu, inc, 31);
          }
        }
      }
      __syncthreads();
      // ---- pass B: scatter the frame into class-sorted shared memory ----
      for (int c = 0; c < nchunks; ++c) {
        const int i = c * kThreads + tid;
        int cls = -1;
        uint8_t b = 0;
        if (i < N) {
          b = tmp[i];
          cls = b >> 4;
        }
        int j = -1;
  #pragma unroll
        for (int q = 0; q < C_COUNT; ++q) {
          const unsigned m = __ballot_sync(0xffffffffu, cls == q);
          if (cls == q) j = s_cnt[q * (kMaxChunks * 8) + c * kWarps + warp] + __popc(m & ((1u << lane) - 1u));
        }
        if (i < N) {
          xs[j * 3] = (double)pts[i * 3];
          xs[j * 3 + 1] = (double)pts[i * 3 + 1];
          xs[j * 3 + 2] = (double)pts[i * 3 + 2];
          so[j * 3] = obs[i * 3];
          so[j * 3 + 1] = obs[i * 3 + 1];
          so[j * 3 + 2] = obs[i * 3 + 2];
          sk[j] = asg[i];
          perm[j] = (uint16_t)i;
          soct[j] = oct[i];
          fl[j] = b;
        }
      }
      __syncthreads();

    }

    const int iters = P.gn_iters;
    const int sweeps = (iters > 0) ? iters + 1 : 0;
    for (int sw = 0; sw < sweeps; ++sw) {
      const bool do_update = pts_free && sw > 0;
      const bool do_lin = sw < iters;           // state sw is linearised (the update of sweep sw + 1 needs it)
      const bool do_acc = pose_free && do_lin;  // ... and enters the pose system
      if (!do_update && !do_acc && !(do_lin && pts_free)) continue;
      const V3f om = {(float)s_dxi[0], (float)s_dxi[1], (float)s_dxi[2]}, up = {(float)s_dxi[3], (float)s_dxi[4], (float)s_dxi[5]};
      float acc[kAcc];
#pragma unroll
      for (int i = 0; i < kAcc; ++i) acc[i] = 0.f;
      int bad = 0;  // bit 0: singular block while accumulating, bit 1: while updating
      // Measured alternatives to this plain strided walk (C2, 4 096 frames, 3.51 ms): walking the class-sorted order
      // backwards 3.57 ms; dealing the 32-point chunks to the warps by estimated cost 3.56 ms (the second resident CTA
      // already absorbs the imbalance); staging the next round's stored linearisation with cp.async 4.17 ms.
      for (int j = tid; j < N; j += kThreads) {
        const uint8_t b = fl[j];
        const int cls = (b >> 4) == C_MSTR ? (int)C_F64 : (b >> 4);  // this kernel keeps mono + full Gaussian on the fp64 path
        const bool stereo = b & F_STEREO;
        const bool upd = do_update && cls >= C_PLAIN;
        const bool accp = do_acc && (cls == C_FIXED || cls == C_STRP || cls == C_STR || (cls == C_F64 && (b & F_STR)));
        const bool linp = do_lin && (cls == C_STRP || cls == C_STR);  // linearise + store (+ accumulate when accp)
        if (!upd && !accp && !linp) continue;
        double x0 = xs[j * 3], x1 = xs[j * 3 + 1], x2 = xs[j * 3 + 2];
        const float ob[3] = {so[j * 3], so[j * 3 + 1], so[j * 3 + 2]};
        const float w = P.s2inv[soct[j] & 7];
        if (cls == C_F64) {
          double x[3] = {x0, x1, x2};
          V3f pcf, vf;
          S3f Mf;
          point_fp64(a, s_old, s_new, s_dxi, b & 15, sk[j], upd, accp, ob, (double)w, x, pcf, Mf, vf, bad);
          if (upd) {
            xs[j * 3] = x[0];
            xs[j * 3 + 1] = x[1];
            xs[j * 3 + 2] = x[2];
          }
          if (accp) accumulate_posef(acc, pcf, Mf, vf);
          continue;
        }
        if (cls == C_PLAIN) {
          // free stereo point without structure edge: D = G = w P^T P with P square, so
          // dc = G^-1 (h - G A dxi) = P^-1 e - A dxi in closed form (rows u and u_r of P differ only in the z column)
          const Lin L = linearise(P, s_old.R, s_old.t, x0, x1, x2, ob, true);
          const float a_ = C.fx * L.iz, b_ = C.fy * L.iz;
          const float c0 = -a_ * L.pc.x * L.iz, c1 = -b_ * L.pc.y * L.iz;
          const float dz = (L.e.z - L.e.x) / (C.bf * L.iz * L.iz);
          const float dx = (L.e.x - c0 * dz) / a_;
          const float dy = (L.e.y - c1 * dz) / b_;
          const V3f ad = {om.y * L.pc.z - om.z * L.pc.y + up.x, om.z * L.pc.x - om.x * L.pc.z + up.y,
                          om.x * L.pc.y - om.y * L.pc.x + up.z};
          const V3f dw = rotf_t(s_old.Rf, {dx - ad.x, dy - ad.y, dz - ad.z});
          if (!(isfinite(dw.x) && isfinite(dw.y) && isfinite(dw.z))) bad |= 2;
          xs[j * 3] = x0 + (double)dw.x;
          xs[j * 3 + 1] = x1 + (double)dw.y;
          xs[j * 3 + 2] = x2 + (double)dw.z;
          continue;
        }
        StrF sc;
        if (cls == C_MPLANE || cls == C_STRP || cls == C_STR) load_structure(a, gn_tab, sk[j], sc);
        if (cls == C_MPLANE) {
          // rows u, v of P and the plane normal form a square system:  [P_uv ; nc^T] dc = [e - P_uv A dxi ; -es]
          // (= D^-1 (h + s - G A dxi) of the general path; weights and the Huber factor cancel row by row)
          const Lin L = linearise(P, s_old.R, s_old.t, x0, x1, x2, ob, false);
          const float a_ = C.fx * L.iz, b_ = C.fy * L.iz;
          const float c0 = -a_ * L.pc.x * L.iz, c1 = -b_ * L.pc.y * L.iz;
          const V3f d3 = {(float)(x0 - sc.mu[0]), (float)(x1 - sc.mu[1]), (float)(x2 - sc.mu[2])};
          const float es = sc.c[0] * d3.x + sc.c[1] * d3.y + sc.c[2] * d3.z;
          const V3f nc = rotf(s_old.Rf, {sc.c[0], sc.c[1], sc.c[2]});
          const V3f ad = {om.y * L.pc.z - om.z * L.pc.y + up.x, om.z * L.pc.x - om.x * L.pc.z + up.y,
                          om.x * L.pc.y - om.y * L.pc.x + up.z};
          const float r0 = L.e.x - (a_ * ad.x + c0 * ad.z), r1 = L.e.y - (b_ * ad.y + c1 * ad.z), r2 = -es;
          const float ia = 1.0f / a_, ib = 1.0f / b_;
          const float den = nc.z - nc.x * c0 * ia - nc.y * c1 * ib;  // = (nc . pc) / z: bounded away from 0 by the view-cos cull
          const float dz = (r2 - nc.x * r0 * ia - nc.y * r1 * ib) / den;
          const float dx = (r0 - c0 * dz) * ia, dy = (r1 - c1 * dz) * ib;
          const V3f dw = rotf_t(s_old.Rf, {dx, dy, dz});
          if (!(isfinite(dw.x) && isfinite(dw.y) && isfinite(dw.z))) bad |= 2;
          xs[j * 3] = x0 + (double)dw.x;
          xs[j * 3 + 1] = x1 + (double)dw.y;
          xs[j * 3 + 2] = x2 + (double)dw.z;
          continue;
        }
        if constexpr (!kStoreLin) {
          // no room in shared memory for the stored linearisation: update at the old state, accumulate at the new one
          if (cls == C_STRP || cls == C_STR) {
          if (cls == C_STRP) {
          const V3f n = {sc.c[0], sc.c[1], sc.c[2]};
          if (upd) {
            const PlaneLin Q = plane_lin(C, linearise(P, s_old.R, s_old.t, x0, x1, x2, ob, true), rotf(s_old.Rf, n), w);
            const float es = n.x * (float)(x0 - sc.mu[0]) + n.y * (float)(x1 - sc.mu[1]) + n.z * (float)(x2 - sc.mu[2]);
            const V3f ad = {om.y * Q.pc.z - om.z * Q.pc.y + up.x, om.z * Q.pc.x - om.x * Q.pc.z + up.y,
                            om.x * Q.pc.y - om.y * Q.pc.x + up.z};
            const float le = C.lam * es;
            const V3f u = {Q.pe.x - ad.x - le * Q.z.x, Q.pe.y - ad.y - le * Q.z.y, Q.pe.z - ad.z - le * Q.z.z};
            const float nu = Q.kappa * (Q.nc.x * u.x + Q.nc.y * u.y + Q.nc.z * u.z);
            const V3f dw = rotf_t(s_old.Rf, {u.x - nu * Q.z.x, u.y - nu * Q.z.y, u.z - nu * Q.z.z});
            if (!(isfinite(dw.x) && isfinite(dw.y) && isfinite(dw.z))) bad |= 2;
            x0 += (double)dw.x;
            x1 += (double)dw.y;
            x2 += (double)dw.z;
            xs[j * 3] = x0;
            xs[j * 3 + 1] = x1;
            xs[j * 3 + 2] = x2;
          }
          if (accp) {
            const PlaneLin Q = plane_lin(C, linearise(P, s_new.R, s_new.t, x0, x1, x2, ob, true), rotf(s_new.Rf, n), w);
            const float es = n.x * (float)(x0 - sc.mu[0]) + n.y * (float)(x1 - sc.mu[1]) + n.z * (float)(x2 - sc.mu[2]);
            const float rho = Q.kappa * (es + Q.nc.x * Q.pe.x + Q.nc.y * Q.pe.y + Q.nc.z * Q.pe.z);
            if (!isfinite(Q.kappa)) bad |= 1;
            // g = A^T nc = [pc x nc ; nc]:  H += kappa g g^T,  rhs += kappa rho' g   (rho = kappa rho')
            const float g[6] = {Q.pc.y * Q.nc.z - Q.pc.z * Q.nc.y, Q.pc.z * Q.nc.x - Q.pc.x * Q.nc.z, Q.pc.x * Q.nc.y - Q.pc.y * Q.nc.x,
                                Q.nc.x, Q.nc.y, Q.nc.z};
            int n_ = 0;
#pragma unroll
            for (int r = 0; r < 6; ++r) {
              const float kg = Q.kappa * g[r];
#pragma unroll
              for (int c = r; c < 6; ++c) acc[n_++] += kg * g[c];
            }
#pragma unroll
            for (int r = 0; r < 6; ++r) acc[21 + r] += rho * g[r];
          }
          continue;
        }
          {
        if (upd) {  // C_STR only
          const Lin L = linearise(P, s_old.R, s_old.t, x0, x1, x2, ob, stereo);
          S3f G, S;
          V3f h, s;
          reprojf(C, L, stereo, w, G, h);
          structuref(C, s_old.Rf, sc, b, x0, x1, x2, S, s);
          const S3f D = {G.xx + S.xx, G.xy + S.xy, G.xz + S.xz, G.yy + S.yy, G.yz + S.yz, G.zz + S.zz};
          const V3f bl = {h.x + s.x, h.y + s.y, h.z + s.z};
          bool ok;
          const S3f Di = inv_symf(D, ok);
          const V3f ad = {om.y * L.pc.z - om.z * L.pc.y + up.x, om.z * L.pc.x - om.x * L.pc.z + up.y,
                          om.x * L.pc.y - om.y * L.pc.x + up.z};
          const V3f ga = mulf(G, ad);
          const V3f dc = mulf(Di, {bl.x - ga.x, bl.y - ga.y, bl.z - ga.z});
          const V3f dw = rotf_t(s_old.Rf, dc);
          x0 += (double)dw.x;
          x1 += (double)dw.y;
          x2 += (double)dw.z;
          if (!(ok && isfinite(dw.x) && isfinite(dw.y) && isfinite(dw.z))) bad |= 2;
          xs[j * 3] = x0;
          xs[j * 3 + 1] = x1;
          xs[j * 3 + 2] = x2;
        }
        if (accp) {  // C_STR (two linearisations per sweep: no room for the stored one)
          const Lin L = linearise(P, s_new.R, s_new.t, x0, x1, x2, ob, stereo);
          S3f G;
          V3f h;
          reprojf(C, L, stereo, w, G, h);
          S3f M = G;
          V3f v = h;
          {
            S3f S;
            V3f s;
            structuref(C, s_new.Rf, sc, b, x0, x1, x2, S, s);
            const S3f D = {G.xx + S.xx, G.xy + S.xy, G.xz + S.xz, G.yy + S.yy, G.yz + S.yz, G.zz + S.zz};
            bool ok;
            const S3f Di = inv_symf(D, ok);
            if (!ok) bad |= 1;
            const float T00 = G.xx * Di.xx + G.xy * Di.xy + G.xz * Di.xz, T01 = G.xx * Di.xy + G.xy * Di.yy + G.xz * Di.yz,
                        T02 = G.xx * Di.xz + G.xy * Di.yz + G.xz * Di.zz;
            const float T10 = G.xy * Di.xx + G.yy * Di.xy + G.yz * Di.xz, T11 = G.xy * Di.xy + G.yy * Di.yy + G.yz * Di.yz,
                        T12 = G.xy * Di.xz + G.yy * Di.yz + G.yz * Di.zz;
            const float T20 = G.xz * Di.xx + G.yz * Di.xy + G.zz * Di.xz, T21 = G.xz * Di.xy + G.yz * Di.yy + G.zz * Di.yz,
                        T22 = G.xz * Di.xz + G.yz * Di.yz + G.zz * Di.zz;
            const V3f bl = {h.x + s.x, h.y + s.y, h.z + s.z};
            // v = h - T (h + s): equals h at the fixed point (h + s = 0) whatever the rounding of D^-1
            v = {h.x - (T00 * bl.x + T01 * bl.y + T02 * bl.z), h.y - (T10 * bl.x + T11 * bl.y + T12 * bl.z),
                 h.z - (T20 * bl.x + T21 * bl.y + T22 * bl.z)};
            M.xx = T00 * S.xx + T01 * S.xy + T02 * S.xz;
            M.xy = T00 * S.xy + T01 * S.yy + T02 * S.yz;
            M.xz = T00 * S.xz + T01 * S.yz + T02 * S.zz;
            M.yy = T10 * S.xy + T11 * S.yy + T12 * S.yz;
            M.yz = T10 * S.xz + T11 * S.yz + T12 * S.zz;
            M.zz = T20 * S.xz + T21 * S.yz + T22 * S.zz;
          }
          accumulate_posef(acc, L.pc, M, v);
        }
                }
          continue;
          }
        }
        if (cls == C_STRP || cls == C_STR) {
          float* sp = store + (size_t)j * kStore;
          if (upd) {  // step from the stored linearisation (state sw - 1): dc = u0 - W (omega x pc + upsilon)
            const V3f pc = rotf(s_old.Rf, {(float)x0, (float)x1, (float)x2});
            const V3f pcs = {pc.x + s_old.tf[0], pc.y + s_old.tf[1], pc.z + s_old.tf[2]};
            const V3f ad = {om.y * pcs.z - om.z * pcs.y + up.x, om.z * pcs.x - om.x * pcs.z + up.y,
                            om.x * pcs.y - om.y * pcs.x + up.z};
            const float4 s0 = *reinterpret_cast<const float4*>(sp), s1 = *reinterpret_cast<const float4*>(sp + 4),
                         s2 = *reinterpret_cast<const float4*>(sp + 8);
            const V3f dc = {s0.x - (s0.w * ad.x + s1.x * ad.y + s1.y * ad.z), s0.y - (s1.z * ad.x + s1.w * ad.y + s2.x * ad.z),
                            s0.z - (s2.y * ad.x + s2.z * ad.y + s2.w * ad.z)};
            const V3f dw = rotf_t(s_old.Rf, dc);
            if (!(isfinite(dw.x) && isfinite(dw.y) && isfinite(dw.z))) bad |= 2;
            x0 += (double)dw.x;
            x1 += (double)dw.y;
            x2 += (double)dw.z;
            xs[j * 3] = x0;
            xs[j * 3 + 1] = x1;
            xs[j * 3 + 2] = x2;
          }
          if (!linp) continue;
          const Lin L = linearise(P, s_new.R, s_new.t, x0, x1, x2, ob, true);
          V3f u0;
          float W[9];
          if (cls == C_STRP) {
            const V3f n = {sc.c[0], sc.c[1], sc.c[2]};
            const PlaneLin Q = plane_lin(C, L, rotf(s_new.Rf, n), w);
            const float es = n.x * (float)(x0 - sc.mu[0]) + n.y * (float)(x1 - sc.mu[1]) + n.z * (float)(x2 - sc.mu[2]);
            if (!isfinite(Q.kappa)) bad |= 1;
            // u0 = D^-1 (h + s) = u - kappa z (nc . u), u = pe - lam es z;   W = D^-1 G = I - kappa z nc^T
            const float le = C.lam * es;
            const V3f u = {Q.pe.x - le * Q.z.x, Q.pe.y - le * Q.z.y, Q.pe.z - le * Q.z.z};
            const float nu = Q.kappa * (Q.nc.x * u.x + Q.nc.y * u.y + Q.nc.z * u.z);
            u0 = {u.x - nu * Q.z.x, u.y - nu * Q.z.y, u.z - nu * Q.z.z};
            const V3f kz = {Q.kappa * Q.z.x, Q.kappa * Q.z.y, Q.kappa * Q.z.z};
            W[0] = 1.f - kz.x * Q.nc.x; W[1] = -kz.x * Q.nc.y; W[2] = -kz.x * Q.nc.z;
            W[3] = -kz.y * Q.nc.x; W[4] = 1.f - kz.y * Q.nc.y; W[5] = -kz.y * Q.nc.z;
            W[6] = -kz.z * Q.nc.x; W[7] = -kz.z * Q.nc.y; W[8] = 1.f - kz.z * Q.nc.z;
            if (accp) {
              const float rho = Q.kappa * (es + Q.nc.x * Q.pe.x + Q.nc.y * Q.pe.y + Q.nc.z * Q.pe.z);
              // g = A^T nc = [pc x nc ; nc]:  H += kappa g g^T,  rhs += kappa (es + nc . pe) g
              const float g[6] = {Q.pc.y * Q.nc.z - Q.pc.z * Q.nc.y, Q.pc.z * Q.nc.x - Q.pc.x * Q.nc.z,
                                  Q.pc.x * Q.nc.y - Q.pc.y * Q.nc.x, Q.nc.x, Q.nc.y, Q.nc.z};
              int n_ = 0;
#pragma unroll
              for (int r = 0; r < 6; ++r) {
                const float kg = Q.kappa * g[r];
#pragma unroll
                for (int c = r; c < 6; ++c) acc[n_++] += kg * g[c];
              }
#pragma unroll
              for (int r = 0; r < 6; ++r) acc[21 + r] += rho * g[r];
            }
          } else {
            S3f G, S;
            V3f h, s;
            reprojf(C, L, true, w, G, h);
            structuref(C, s_new.Rf, sc, b, x0, x1, x2, S, s);
            const S3f D = {G.xx + S.xx, G.xy + S.xy, G.xz + S.xz, G.yy + S.yy, G.yz + S.yz, G.zz + S.zz};
            bool ok;
            const S3f Di = inv_symf(D, ok);
            if (!ok) bad |= 1;
            // T = G D^-1;  W = D^-1 G = T^T
            const float T00 = G.xx * Di.xx + G.xy * Di.xy + G.xz * Di.xz, T01 = G.xx * Di.xy + G.xy * Di.yy + G.xz * Di.yz,
                        T02 = G.xx * Di.xz + G.xy * Di.yz + G.xz * Di.zz;
            const float T10 = G.xy * Di.xx + G.yy * Di.xy + G.yz * Di.xz, T11 = G.xy * Di.xy + G.yy * Di.yy + G.yz * Di.yz,
                        T12 = G.xy * Di.xz + G.yy * Di.yz + G.yz * Di.zz;
            const float T20 = G.xz * Di.xx + G.yz * Di.xy + G.zz * Di.xz, T21 = G.xz * Di.xy + G.yz * Di.yy + G.zz * Di.yz,
                        T22 = G.xz * Di.xz + G.yz * Di.yz + G.zz * Di.zz;
            const V3f bl = {h.x + s.x, h.y + s.y, h.z + s.z};
            u0 = mulf(Di, bl);
            W[0] = T00; W[1] = T10; W[2] = T20; W[3] = T01; W[4] = T11; W[5] = T21; W[6] = T02; W[7] = T12; W[8] = T22;
            if (accp) {
              // v = h - T (h + s): equals h at the fixed point (h + s = 0) whatever the rounding of D^-1
              const V3f v = {h.x - (T00 * bl.x + T01 * bl.y + T02 * bl.z), h.y - (T10 * bl.x + T11 * bl.y + T12 * bl.z),
                             h.z - (T20 * bl.x + T21 * bl.y + T22 * bl.z)};
              S3f M;
              M.xx = T00 * S.xx + T01 * S.xy + T02 * S.xz;
              M.xy = T00 * S.xy + T01 * S.yy + T02 * S.yz;
              M.xz = T00 * S.xz + T01 * S.yz + T02 * S.zz;
              M.yy = T10 * S.xy + T11 * S.yy + T12 * S.yz;
              M.yz = T10 * S.xz + T11 * S.yz + T12 * S.zz;
              M.zz = T20 * S.xz + T21 * S.yz + T22 * S.zz;
              accumulate_posef(acc, L.pc, M, v);
            }
          }
          *reinterpret_cast<float4*>(sp) = make_float4(u0.x, u0.y, u0.z, W[0]);
          *reinterpret_cast<float4*>(sp + 4) = make_float4(W[1], W[2], W[3], W[4]);
          *reinterpret_cast<float4*>(sp + 8) = make_float4(W[5], W[6], W[7], W[8]);
          continue;
        }
        if (accp) {  // C_FIXED: a held point contributes G, h
          const Lin L = linearise(P, s_new.R, s_new.t, x0, x1, x2, ob, stereo);
          S3f G;
          V3f h;
          reprojf(C, L, stereo, w, G, h);
          accumulate_posef(acc, L.pc, G, h);
        }
      }
      if (bad & 2) s_fail2 = 1;
      if (bad & 1) s_fail1 = 1;
      if (do_acc) {
#pragma unroll
        for (int i = 0; i < kAcc; ++i) {  // lanes in fp32 (fixed butterfly order), warps in fp64
          float s = acc[i];
#pragma unroll
          for (int off = 16; off > 0; off >>= 1) s += __shfl_xor_sync(0xffffffffu, s, off);
          if (lane == 0) s_red[warp][i] = (double)s;
        }
      }
      __syncthreads();
      if (s_fail2) break;  // uniform
      if (do_acc) {
        if (warp == 0) {  // fixed-order cross-warp sum, then the 6x6 solve and the SE3 update on lane 0
          if (lane < kAcc) {
            double s = s_red[0][lane];
#pragma unroll
            for (int w2 = 1; w2 < kWarps; ++w2) s += s_red[w2][lane];
            s_acc[lane] = s;
          }
          __syncwarp();
          if (lane == 0) {
            double dxi[6];
            bool okf = !s_fail1;
#pragma unroll
            for (int i = 0; i < kAcc; ++i) okf = okf && isfinite(s_acc[i]);
            if (!okf || !solve6_ldlt(s_acc, dxi)) {
              s_fail1 = 1;
            } else {
              double np[7];
#pragma unroll
              for (int i = 0; i < 9; ++i) {
                s_old.R[i] = s_new.R[i];
                s_old.Rf[i] = s_new.Rf[i];
              }
#pragma unroll
              for (int i = 0; i < 3; ++i) {
                s_old.t[i] = s_new.t[i];
                s_old.tf[i] = s_new.tf[i];
              }
              se3_oplus_fast(dxi, s_pose, np, s_new.R);
#pragma unroll
              for (int i = 0; i < 6; ++i) s_dxi[i] = dxi[i];
#pragma unroll
              for (int i = 0; i < 7; ++i) s_pose[i] = np[i];
#pragma unroll
              for (int i = 0; i < 9; ++i) s_new.Rf[i] = (float)s_new.R[i];
#pragma unroll
              for (int i = 0; i < 3; ++i) {
                s_new.t[i] = np[4 + i];
                s_new.tf[i] = (float)np[4 + i];
              }
            }
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
      quat_to_rot(q, s_new.R);
    }
    __syncthreads();
    if (tid < 7) a.pose_out[(size_t)f * 7 + tid] = s_pose[tid];
    if (tid == 0) a.status[f] = fail ? GMMLOC_FRAME_NUMERIC : 0;
    {
      // outputs in the ORIGINAL point order; outlier post-check at the final estimate in fp64:
      // chi2 > 5.991 / 7.815 or non-positive depth (localization_opt.cpp:857-881)
      double R[9], t[3];
#pragma unroll
      for (int i = 0; i < 9; ++i) R[i] = s_new.R[i];
#pragma unroll
      for (int i = 0; i < 3; ++i) t[i] = s_pose[4 + i];
      float* po = a.pts_out ? a.pts_out + (size_t)f * N * 3 : nullptr;
      uint8_t* out = a.outlier ? a.outlier + (size_t)f * N : nullptr;
      for (int j = tid; j < N; j += kThreads) {
        const int i = perm[j];
        Vec3 xw;
        if (fail) xw = {(double)pts[i * 3], (double)pts[i * 3 + 1], (double)pts[i * 3 + 2]};
        else xw = {xs[j * 3], xs[j * 3 + 1], xs[j * 3 + 2]};
        if (po != nullptr) {
          po[i * 3] = fail ? pts[i * 3] : (float)xw.x;
          po[i * 3 + 1] = fail ? pts[i * 3 + 1] : (float)xw.y;
          po[i * 3 + 2] = fail ? pts[i * 3 + 2] : (float)xw.z;
        }
        if (out != nullptr) {
          Vec3 pc = rot(R, xw);
          pc.x += t[0]; pc.y += t[1]; pc.z += t[2];
          const bool stereo = fl[j] & F_STEREO;
          const double invz = 1.0 / pc.z;
          const double u = pc.x * invz * P.fx + P.cx, v = pc.y * invz * P.fy + P.cy;
          const double e0 = (double)so[j * 3] - u, e1 = (double)so[j * 3 + 1] - v;
          double e2 = 0.0;
          if (stereo) e2 = (double)so[j * 3 + 2] - (u - P.bf * invz);
          const double chi = (double)P.s2inv[soct[j] & 7] * (e0 * e0 + e1 * e1 + e2 * e2);
          const double th = stereo ? P.chi2_stereo : P.chi2_mono;
          out[i] = (chi > th || !(pc.z > 0.0)) ? 1 : 0;
        }
      }
    }
  }
}

}  // namespace

// shared-memory bytes of one frame: base state (38 B / point), aux arrays (7 B), stored linearisation (48 B)
static size_t smem_base(int N) { return (((size_t)N * 38 + 15) & ~(size_t)15) + 16; }
static size_t smem_aux(int N) { return (((size_t)N * 7 + 15) & ~(size_t)15); }
static size_t smem_store(int N) { return (size_t)N * 4 * kStore; }

size_t frame_gn_mixed_limit_points() {
  const size_t by_smem = (size_t)(216 * 1024 - 32) / 38;
  return std::min(by_smem, (size_t)(8 * kMaxChunks - 12) * 32);
}

template <int kThreads, int kMinB, bool kStoreLin, bool kPrep>
static cudaError_t launch_mixed_t(const FrameArgs& a, const float* gn_tab, int* d_counters, size_t smem, float* store_ws,
                                  size_t store_ws_bytes, int aux_global, const unsigned char* prep, cudaStream_t stream, bool* fits) {
  cudaError_t e;
  // The opt-in is per function and per device, i.e. state shared by every host thread that launches this variant
  // (one gmmloc_ctx_t per thread): it is always set to the SAME value — the most this variant can use, 227 KB per CTA
  // minus its static shared memory, halved (minus the 1 KB system reserve) when two CTAs share an SM — so concurrent
  // callers with different N cannot shrink it under each other's launches.
  {
    cudaFuncAttributes fa;
    if ((e = cudaFuncGetAttributes(&fa, frame_gn_mixed_kernel<kThreads, kMinB, kStoreLin, kPrep>)) != cudaSuccess) return e;
    const size_t max_dyn = kMinB == 2 ? ((228 * 1024 / 2 - 1024 - fa.sharedSizeBytes) & ~(size_t)15)
                                      : ((227 * 1024 - fa.sharedSizeBytes) & ~(size_t)15);
    if (smem > max_dyn) return cudaErrorInvalidValue;
    if ((e = cudaFuncSetAttribute(frame_gn_mixed_kernel<kThreads, kMinB, kStoreLin, kPrep>, cudaFuncAttributeMaxDynamicSharedMemorySize,
                                  (int)max_dyn)) != cudaSuccess)
      return e;
  }
  int per_sm = 0;
  if ((e = cudaOccupancyMaxActiveBlocksPerMultiprocessor(&per_sm, frame_gn_mixed_kernel<kThreads, kMinB, kStoreLin, kPrep>, kThreads,
                                                         smem)) != cudaSuccess)
    return e;
  if (per_sm < 1) per_sm = 1;
  int dev = 0, sms = 148;
  cudaGetDevice(&dev);
  cudaDeviceGetAttribute(&sms, cudaDevAttrMultiProcessorCount, dev);
  const int grid = std::min(a.B, sms * per_sm);
  if (store_ws != nullptr && (size_t)grid * scratch_slice(a.N) > store_ws_bytes) {
    *fits = false;  // scratch too small for this grid: the caller picks another variant
    return cudaSuccess;
  }
  *fits = true;
  if ((e = cudaMemsetAsync(d_counters, 0, sizeof(int), stream)) != cudaSuccess) return e;
  frame_gn_mixed_kernel<kThreads, kMinB, kStoreLin, kPrep><<<grid, kThreads, smem, stream>>>(a, gn_tab, d_counters, store_ws, aux_global,
                                                                                         prep);
  return cudaGetLastError();
}

size_t frame_gn_mixed_scratch_bytes(int N) {
  int dev = 0, sms = 148;
  cudaGetDevice(&dev);
  cudaDeviceGetAttribute(&sms, cudaDevAttrMultiProcessorCount, dev);
  return (size_t)2 * sms * scratch_slice(N);
}

size_t frame_gn_mixed_prep_bytes(int B, int N, int mode) {
  // mode (gmmloc_ctx_set_option GMMLOC_OPT_GN_PREP): 0 / 1 = never / always (A/B runs, tests); -1 = batches of at least
  // 2 M points.  Measured on C2: 4 096-frame launches 3.51 -> 3.44 ms (the solver loses its two latency-bound load passes),
  // but the 888-frame chunks of the pipelined host path get slower (e2e 682 k -> 672 k frames/s): there the extra launch
  // and the round trip of the sorted frames through HBM cost more than the dependent loads they remove.
  if (const char* v = debug_env("GMMLOC_GN_PREP")) mode = (v[0] == '0') ? 0 : (v[0] == '1') ? 1 : mode;
  if (mode == 0 || (mode < 0 && (size_t)B * (size_t)N < ((size_t)2 << 20))) return 0;
  return (size_t)B * prep_stride(N) + 16;
}

template <bool kPrep>
static cudaError_t launch_variants(const FrameArgs& a, const float* gn_tab, int* d_counters, float* store_ws, size_t store_ws_bytes,
                                   const unsigned char* prep, cudaStream_t stream) {
  const size_t base = smem_base(a.N) + smem_aux(a.N), with_store = base + smem_store(a.N);
  // per-CTA dynamic budgets: an SM has 228 KB, every resident CTA costs its static shared memory + 1 KB of system
  // reserve, and one CTA may opt in to at most 227 KB (static included)
  static const size_t static192 = [] {
    cudaFuncAttributes fa;
    return cudaFuncGetAttributes(&fa, frame_gn_mixed_kernel<kThreads2, 2, true, kPrep>) == cudaSuccess ? fa.sharedSizeBytes : (size_t)8192;
  }();
  static const size_t static384 = [] {
    cudaFuncAttributes fa;
    return cudaFuncGetAttributes(&fa, frame_gn_mixed_kernel<384, 1, true, kPrep>) == cudaSuccess ? fa.sharedSizeBytes : (size_t)9728;
  }();
  const size_t two = ((228 * 1024 / 2 - 1024 - static192) & ~(size_t)15);  // when two CTAs share an SM
  const size_t one = std::min<size_t>(216 * 1024, (227 * 1024 - static384) & ~(size_t)15);  // a CTA with the SM to itself
  // Two frames per SM (192 threads each) hide each other's serial 6x6 solve and barriers: preferred whenever they fit.
  // The stored linearisation lives in shared memory if that fits too, else in the L2-resident global scratch.  Frames too
  // large for two residents run one CTA of 384 threads with the stored linearisation and the aux arrays (component id,
  // original index) in the global scratch (N <= 5 819).
  // 192 threads per CTA: 168 registers per thread without spills (256 threads cap at 128 registers and spill ~300 B in
  // the sweep loop; measured 4.25 ms vs 3.49 ms per 4 096 C2 frames; 160 / 128 threads: 3.78 / 4.12 ms)
  bool fits = true;
  cudaError_t e;
  if (with_store <= two)
    return launch_mixed_t<kThreads2, 2, true, kPrep>(a, gn_tab, d_counters, with_store, nullptr, 0, 0, prep, stream, &fits);
  if (base <= two) {
    if (store_ws != nullptr) {
      e = launch_mixed_t<kThreads2, 2, true, kPrep>(a, gn_tab, d_counters, base, store_ws, store_ws_bytes, 0, prep, stream, &fits);
      if (e != cudaSuccess || fits) return e;
    }
    return launch_mixed_t<kThreads2, 2, false, kPrep>(a, gn_tab, d_counters, base, nullptr, 0, 0, prep, stream, &fits);
  }
  if (with_store <= one) return launch_mixed_t<384, 1, true, kPrep>(a, gn_tab, d_counters, with_store, nullptr, 0, 0, prep, stream, &fits);
  if (store_ws != nullptr) {
    e = launch_mixed_t<384, 1, true, kPrep>(a, gn_tab, d_counters, smem_base(a.N), store_ws, store_ws_bytes, 1, prep, stream, &fits);
    if (e != cudaSuccess || fits) return e;
  }
  return cudaErrorInvalidValue;  // the caller checks frame_gn_mixed_limit_points() and passes the scratch: not reached
}

cudaError_t launch_frame_prep(const FrameArgs& a, unsigned char* prep, cudaStream_t stream) {
  if (a.B <= 0) return cudaSuccess;
  int dev = 0, sms = 148;
  cudaGetDevice(&dev);
  cudaDeviceGetAttribute(&sms, cudaDevAttrMultiProcessorCount, dev);
  const int grid = std::min(a.B, sms * 8);
  frame_prep_kernel<<<grid, kPrepThreads, (size_t)((a.N + 15) & ~15), stream>>>(a, prep);
  return cudaGetLastError();
}

cudaError_t launch_frame_gn_mixed(const FrameArgs& a, const float* gn_tab, int* d_counters, float* store_ws, size_t store_ws_bytes,
                                  unsigned char* prep_ws, size_t prep_ws_bytes, cudaStream_t stream) {
  const size_t need = (size_t)a.B * prep_stride(a.N);
  if (prep_ws != nullptr && prep_ws_bytes >= need && need > 0) {
    // K3p: classify + class-sort every frame into prep_ws, then the solver copies its frames from there
    unsigned char* prep = reinterpret_cast<unsigned char*>((reinterpret_cast<uintptr_t>(prep_ws) + 15) & ~(uintptr_t)15);
    if ((size_t)(prep - prep_ws) + need > prep_ws_bytes) return cudaErrorInvalidValue;
    cudaError_t e = launch_frame_prep(a, prep, stream);
    if (e != cudaSuccess) return e;
    return launch_variants<true>(a, gn_tab, d_counters, store_ws, store_ws_bytes, prep, stream);
  }
  return launch_variants<false>(a, gn_tab, d_counters, store_ws, store_ws_bytes, nullptr, stream);
}

}  // namespace gmmloc
