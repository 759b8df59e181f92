// frame_gn_point.cuh — fp64 per-point pieces of the frame Gauss-Newton (shared by frame_gn.cu, which runs every point
// through them, and frame_gn_mixed.cu, which uses them for the ill-conditioned point blocks only).
// Reference: EdgeStereoSE3ProjectXYZ / EdgeSE3ProjectXYZ, EdgePt2GaussianDeg / EdgePt2Gaussian
// (gmmloc/src/gmm/factors.cpp:5-17,55-64,84-107,133-168) inside Localization::jointOptimization
// (gmmloc/src/modules/localization_opt.cpp:456-925); formulation in the header of frame_gn.cu.
#pragma once
#include "common.cuh"
#include "gn_common.cuh"

namespace gmmloc {
namespace {

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
__device__ __forceinline__ bool pose_terms(const FrameArgs& a, const double* R, const double* t, const StrConst& sc, uint8_t b,
                                           const Vec3& xw, const float* ob, double w, Vec3& pc, Sym3& M, Vec3& v) {
  const FrameParamsDev& P = a.prm;
  pc = rot(R, xw);
  pc.x += t[0]; pc.y += t[1]; pc.z += t[2];
  Sym3 G;
  Vec3 h;
  reproj(P, pc, ob[0], ob[1], ob[2], b & F_STEREO, w, P.robust != 0, G, h);
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
  return ok;
}
__device__ __forceinline__ bool pose_contribution(const FrameArgs& a, const double* R, const double* t, const StrConst& sc,
                                                  uint8_t b, const Vec3& xw, const float* ob, double w, double* acc) {
  Vec3 pc, v;
  Sym3 M;
  const bool ok = pose_terms(a, R, t, sc, b, xw, ob, w, pc, M, v);
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

}  // namespace
}  // namespace gmmloc
