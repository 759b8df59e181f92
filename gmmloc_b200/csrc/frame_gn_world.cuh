// frame_gn_world.cuh — per-point algebra of the WORLD-FRAME mixed-precision frame Gauss-Newton (K3w, frame_gn_world.cu).
//
// Same system as frame_gn_kernel / frame_gn_mixed_kernel (Localization::jointOptimization with a single free pose,
// gmmloc/src/modules/localization_opt.cpp:456-925; edges EdgeStereoSE3ProjectXYZ / EdgeSE3ProjectXYZ + EdgePt2Gaussian /
// EdgePt2GaussianDeg, gmmloc/src/gmm/factors.cpp:5-17,55-64,84-107,133-168), same precision split as K3m (fp64: state,
// camera-frame point, residuals, the reduced 6x6 system; fp32: everything that multiplies a residual).  What changes is
// the FRAME the per-point blocks are written in.  g2o perturbs the pose on the left, T <- exp(xi_c) T; the same update
// is T <- T exp(xi_w) with xi_c = Ad_T xi_w, and in the right-hand (world) form
//
//      d pc = R (omega_w x x' + upsilon_w),       x' = x - c0   (c0: a fixed reference point, the initial camera centre)
//
// so every Jacobian factors through J = P R (P = d proj / d pc, 3x3 or 2x3) and A_w = [-[x']x  I], and the point block is
//
//      G_w = w J^T J,  h_w = w J^T e,  D_w = G_w + S_w,  S_w / b_w = the structure edge AS STORED IN THE MAP
//
// — the structure Hessian needs no rotation (K3m rotated it into the camera frame every sweep: 45 flop for a full
// Gaussian), the point step needs none either, and the step of a point does not need the pose of the linearisation:
//
//      dx = u0 - W a,   u0 = D_w^-1 (h_w + b_w),   W = D_w^-1 G_w,   a = omega_w x x' + upsilon_w
//
// (u0, W) is what the sweep stores per free point (48 B); the update pass is 15 FMAs per point whatever the class.
// The reduced pose system is accumulated in world-twist coordinates, H_w = sum A_w^T M_w A_w with M_w = W^T S_w
// (= G - G D^-1 G without the cancellation) and g_w = sum A_w^T (h_w - G_w u0); after the 6x6 solve the step is mapped
// back with the adjoint (omega_c = R omega_w, upsilon_c = R upsilon_w + t' x omega_c, t' = t + R c0) and applied
// exactly as the reference does, T <- exp(xi_c) T.  The identity of the two formulations is pinned on the CPU in
// tests/test_oracle_solve.py::test_world_frame_schur_formulation_equals_camera_frame.
//
// Everything here is plain C++ (`GNW_HD`), so that tests/cpp/gnw_host_check.cpp can run the very same functions on the
// host against the oracle before a GPU is involved.
#pragma once
#include <math.h>
#include <stdint.h>

#if defined(__CUDACC__)
#define GNW_HD __host__ __device__ __forceinline__
#else
#define GNW_HD inline
#endif

namespace gmmloc {
namespace gnw {

struct V3f {
  float x, y, z;
};
struct S3f {
  float xx, xy, xz, yy, yz, zz;
};

// camera constants in both precisions + solver switches
struct Cam {
  double fx, fy, cx, cy, bf;
  float fxf, fyf, bff, lam, d_mono, d_stereo;
  int robust;
};

// reciprocal / quotient of the fp32 Gauss-Newton MODEL (Jacobian scalings, D^-1, Sherman-Morrison factors): one MUFU.RCP
// (1 ulp) instead of the ~10-instruction IEEE sequence with its slow-path branch.  Every use either multiplies a residual
// that was formed in fp64 (the step it yields vanishes with the residual, so the fixed point is untouched) or shapes the
// Hessian model (convergence rate only); 1 / z of the projection gets a Newton step in fp64 on top (linearise()).
GNW_HD float rcpf(float x) {
#if defined(__CUDA_ARCH__)
  float r;
  asm("rcp.approx.ftz.f32 %0, %1;" : "=f"(r) : "f"(x));  // bare MUFU.RCP (__fdividef adds a range-scaling guard)
  return r;
#else
  return 1.0f / x;
#endif
}
GNW_HD float divf(float a, float b) { return a * rcpf(b); }
GNW_HD float rsqf(float x) {
#if defined(__CUDA_ARCH__)
  return rsqrtf(x);
#else
  return 1.0f / sqrtf(x);
#endif
}
GNW_HD bool finite3(const V3f& v) { return isfinite(v.x) && isfinite(v.y) && isfinite(v.z); }

GNW_HD V3f mulf(const S3f& A, const V3f& v) {
  return {A.xx * v.x + A.xy * v.y + A.xz * v.z, A.xy * v.x + A.yy * v.y + A.yz * v.z, A.xz * v.x + A.yz * v.y + A.zz * v.z};
}
GNW_HD float dotf(const V3f& a, const V3f& b) { return a.x * b.x + a.y * b.y + a.z * b.z; }
// R^T v for a row-major fp32 rotation
GNW_HD V3f rot_t(const float* R, const V3f& v) {
  return {R[0] * v.x + R[3] * v.y + R[6] * v.z, R[1] * v.x + R[4] * v.y + R[7] * v.z, R[2] * v.x + R[5] * v.y + R[8] * v.z};
}
GNW_HD V3f rot(const float* R, const V3f& v) {
  return {R[0] * v.x + R[1] * v.y + R[2] * v.z, R[3] * v.x + R[4] * v.y + R[5] * v.z, R[6] * v.x + R[7] * v.y + R[8] * v.z};
}
GNW_HD S3f inv_sym(const S3f& D, bool& ok) {
  const float c00 = D.yy * D.zz - D.yz * D.yz;
  const float c01 = D.xz * D.yz - D.xy * D.zz;
  const float c02 = D.xy * D.yz - D.xz * D.yy;
  const float det = D.xx * c00 + D.xy * c01 + D.xz * c02;
  const float id = rcpf(det);
  ok = (det != 0.0f) && isfinite(id);
  return {c00 * id, c01 * id, c02 * id, (D.xx * D.zz - D.xz * D.xz) * id, (D.xy * D.xz - D.xx * D.yz) * id,
          (D.xx * D.yy - D.xy * D.xy) * id};
}

// ---- fp64 part of one linearisation: camera-frame point and reprojection residual, rounded to fp32 at the end ----
struct Lin {
  V3f pc;    // camera-frame point
  float iz;  // 1 / z
  V3f e;     // residual (u, v, u_r); e.z = 0 for a monocular point
};
GNW_HD Lin linearise(const Cam& C, const double* R, const double* t, double x0, double x1, double x2, const float* ob, bool stereo) {
  const double pcx = R[0] * x0 + R[1] * x1 + R[2] * x2 + t[0];
  const double pcy = R[3] * x0 + R[4] * x1 + R[5] * x2 + t[1];
  const double pcz = R[6] * x0 + R[7] * x1 + R[8] * x2 + t[2];
  // 1 / z: fp32 reciprocal + one Newton step in fp64 (relative error ~1e-14; a full fp64 division costs ~30 instructions)
  double iz = (double)rcpf((float)pcz);
  iz = iz * (2.0 - pcz * iz);
  const double u = pcx * (C.fx * iz) + C.cx, v = pcy * (C.fy * iz) + C.cy;
  Lin L;
  L.e.x = (float)((double)ob[0] - u);
  L.e.y = (float)((double)ob[1] - v);
  L.e.z = stereo ? (float)((double)ob[2] - (u - C.bf * iz)) : 0.f;
  L.pc = {(float)pcx, (float)pcy, (float)pcz};
  L.iz = (float)iz;
  return L;
}

// entries of P = d proj / d pc:  rows  u: [a 0 c0]   v: [0 b c1]   u_r: [a 0 c0 + cz]
struct Proj {
  float a, b, c0, c1, cz;
};
GNW_HD Proj proj_jac(const Cam& C, const Lin& L) {
  Proj P;
  P.a = C.fxf * L.iz;
  P.b = C.fyf * L.iz;
  P.c0 = -P.a * L.pc.x * L.iz;
  P.c1 = -P.b * L.pc.y * L.iz;
  P.cz = C.bff * L.iz * L.iz;
  return P;
}
// g2o RobustKernelHuber on the reprojection edge: weight w delta / sqrt(chi2) beyond delta^2
GNW_HD float huber(const Cam& C, const Lin& L, bool stereo, float w) {
  if (C.robust) {
    const float chi = w * (L.e.x * L.e.x + L.e.y * L.e.y + L.e.z * L.e.z);
    const float delta = stereo ? C.d_stereo : C.d_mono;
    if (chi > delta * delta) w *= delta * rsqf(chi);
  }
  return w;
}

// world-frame Gauss-Newton model of the reprojection edge: rows of J = P R, G = w J^T J, h = w J^T e
GNW_HD void reproj_world(const Proj& P, const Lin& L, const float* Rf, bool stereo, float w, S3f& G, V3f& h) {
  const V3f r0 = {P.a * Rf[0] + P.c0 * Rf[6], P.a * Rf[1] + P.c0 * Rf[7], P.a * Rf[2] + P.c0 * Rf[8]};
  const V3f r1 = {P.b * Rf[3] + P.c1 * Rf[6], P.b * Rf[4] + P.c1 * Rf[7], P.b * Rf[5] + P.c1 * Rf[8]};
  const V3f q0 = {w * r0.x, w * r0.y, w * r0.z}, q1 = {w * r1.x, w * r1.y, w * r1.z};
  G.xx = q0.x * r0.x + q1.x * r1.x;
  G.xy = q0.x * r0.y + q1.x * r1.y;
  G.xz = q0.x * r0.z + q1.x * r1.z;
  G.yy = q0.y * r0.y + q1.y * r1.y;
  G.yz = q0.y * r0.z + q1.y * r1.z;
  G.zz = q0.z * r0.z + q1.z * r1.z;
  h.x = q0.x * L.e.x + q1.x * L.e.y;
  h.y = q0.y * L.e.x + q1.y * L.e.y;
  h.z = q0.z * L.e.x + q1.z * L.e.y;
  if (stereo) {
    const V3f r2 = {r0.x + P.cz * Rf[6], r0.y + P.cz * Rf[7], r0.z + P.cz * Rf[8]};
    const V3f q2 = {w * r2.x, w * r2.y, w * r2.z};
    G.xx += q2.x * r2.x;
    G.xy += q2.x * r2.y;
    G.xz += q2.x * r2.z;
    G.yy += q2.y * r2.y;
    G.yz += q2.y * r2.z;
    G.zz += q2.z * r2.z;
    h.x += q2.x * L.e.z;
    h.y += q2.y * L.e.z;
    h.z += q2.z * L.e.z;
  }
}

// acc[0..20] += upper triangle of A^T M A, acc[21..26] += A^T v, A = [-[p]x  I]
GNW_HD void accumulate_pose(float* acc, const V3f& p, const S3f& M, const V3f& v) {
  const float C00 = p.y * M.xz - p.z * M.xy, C01 = p.y * M.yz - p.z * M.yy, C02 = p.y * M.zz - p.z * M.yz;
  const float C10 = p.z * M.xx - p.x * M.xz, C11 = p.z * M.xy - p.x * M.yz, C12 = p.z * M.xz - p.x * M.zz;
  const float C20 = p.x * M.xy - p.y * M.xx, C21 = p.x * M.yy - p.y * M.xy, C22 = p.x * M.yz - p.y * M.xz;
  acc[0] += C02 * p.y - C01 * p.z;
  acc[1] += C00 * p.z - C02 * p.x;
  acc[2] += C01 * p.x - C00 * p.y;
  acc[3] += C00;
  acc[4] += C01;
  acc[5] += C02;
  acc[6] += C10 * p.z - C12 * p.x;
  acc[7] += C11 * p.x - C10 * p.y;
  acc[8] += C10;
  acc[9] += C11;
  acc[10] += C12;
  acc[11] += C21 * p.x - C20 * p.y;
  acc[12] += C20;
  acc[13] += C21;
  acc[14] += C22;
  acc[15] += M.xx;
  acc[16] += M.xy;
  acc[17] += M.xz;
  acc[18] += M.yy;
  acc[19] += M.yz;
  acc[20] += M.zz;
  acc[21] += p.y * v.z - p.z * v.y;
  acc[22] += p.z * v.x - p.x * v.z;
  acc[23] += p.x * v.y - p.y * v.x;
  acc[24] += v.x;
  acc[25] += v.y;
  acc[26] += v.z;
}
// rank-one twin: M = kappa n n^T, v = rho n  =>  with g = A^T n = [p x n ; n]:  H += kappa g g^T, rhs += rho g
GNW_HD void accumulate_pose_rank1(float* acc, const V3f& p, const V3f& n, float kappa, float rho) {
  const float g[6] = {p.y * n.z - p.z * n.y, p.z * n.x - p.x * n.z, p.x * n.y - p.y * n.x, n.x, n.y, n.z};
  int q = 0;
#pragma unroll
  for (int r = 0; r < 6; ++r) {
    const float kg = kappa * g[r];
#pragma unroll
    for (int c = r; c < 6; ++c) acc[q++] += kg * g[c];
  }
#pragma unroll
  for (int r = 0; r < 6; ++r) acc[21 + r] += rho * g[r];
}

// what a sweep stores per free point: the step of the point is  dx = u0 - W a
struct PointLin {
  V3f u0;
  float w0, w1, w2, w3, w4, w5, w6, w7, w8;  // W row-major (named scalars: an indexed array ends up in local memory)
};

// ---- class PLAIN: free stereo point without structure edge.  D = G, P square:  u0 = R^T P^-1 e,  W = I, no Schur term ----
GNW_HD V3f lin_plain(const Cam& C, const Lin& L, const float* Rf) {
  const Proj P = proj_jac(C, L);
  const float dz = divf(L.e.z - L.e.x, P.cz);
  const float dx = divf(L.e.x - P.c0 * dz, P.a);
  const float dy = divf(L.e.y - P.c1 * dz, P.b);
  return rot_t(Rf, {dx, dy, dz});
}

// ---- classes STR / MSTR: free stereo / monocular point with a full 3-D Gaussian edge (EdgePt2Gaussian: Hessian cov_inv,
// rhs -cov_inv d).  For a monocular point G has rank 2; D = G + S stays as well conditioned as S (eigenvalues of cov_inv
// between 1 / scale_max and 1e4 for a non-degenerate component), so the fp32 model serves it like the stereo case ----
// Sw: the six unique entries of cov_inv (world frame, as stored in the map), d = x - mu (formed in fp64, rounded).
// Returns false when D is singular.  M / v (the Schur terms) are produced only when `want_pose`.
GNW_HD bool lin_str(const Cam& C, const Lin& L, const float* Rf, bool stereo, float w, const S3f& Sw, const V3f& d, PointLin& out,
                    bool want_pose, S3f& M, V3f& v) {
  const Proj P = proj_jac(C, L);
  S3f G;
  V3f h;
  reproj_world(P, L, Rf, stereo, huber(C, L, stereo, w), G, h);
  const V3f sd = mulf(Sw, d);
  const V3f bl = {h.x - sd.x, h.y - sd.y, h.z - sd.z};
  const S3f D = {G.xx + Sw.xx, G.xy + Sw.xy, G.xz + Sw.xz, G.yy + Sw.yy, G.yz + Sw.yz, G.zz + Sw.zz};
  bool ok;
  const S3f Di = inv_sym(D, ok);
  out.u0 = mulf(Di, bl);
  // W = D^-1 G (general 3x3)
  out.w0 = Di.xx * G.xx + Di.xy * G.xy + Di.xz * G.xz;
  out.w1 = Di.xx * G.xy + Di.xy * G.yy + Di.xz * G.yz;
  out.w2 = Di.xx * G.xz + Di.xy * G.yz + Di.xz * G.zz;
  out.w3 = Di.xy * G.xx + Di.yy * G.xy + Di.yz * G.xz;
  out.w4 = Di.xy * G.xy + Di.yy * G.yy + Di.yz * G.yz;
  out.w5 = Di.xy * G.xz + Di.yy * G.yz + Di.yz * G.zz;
  out.w6 = Di.xz * G.xx + Di.yz * G.xy + Di.zz * G.xz;
  out.w7 = Di.xz * G.xy + Di.yz * G.yy + Di.zz * G.yz;
  out.w8 = Di.xz * G.xz + Di.yz * G.yz + Di.zz * G.zz;
  if (want_pose) {
    // v = h - G u0: equals h at the fixed point (h + b = 0 => u0 = 0) whatever the rounding of D^-1
    const V3f gu = mulf(G, out.u0);
    v = {h.x - gu.x, h.y - gu.y, h.z - gu.z};
    // M = W^T S (= G D^-1 S = G - G D^-1 G without the cancellation; symmetric in exact arithmetic, upper triangle taken)
    M.xx = out.w0 * Sw.xx + out.w3 * Sw.xy + out.w6 * Sw.xz;
    M.xy = out.w0 * Sw.xy + out.w3 * Sw.yy + out.w6 * Sw.yz;
    M.xz = out.w0 * Sw.xz + out.w3 * Sw.yz + out.w6 * Sw.zz;
    M.yy = out.w1 * Sw.xy + out.w4 * Sw.yy + out.w7 * Sw.yz;
    M.yz = out.w1 * Sw.xz + out.w4 * Sw.yz + out.w7 * Sw.zz;
    M.zz = out.w2 * Sw.xz + out.w5 * Sw.yz + out.w8 * Sw.zz;
    // M = (G^-1 + S^-1)^-1 is positive semidefinite; a diagonal entry that comes out clearly negative means D was too
    // ill-conditioned for the fp32 inverse: report it (the frame is then solved by the fp64 kernel)
    const float tol = 1e-3f * (fabsf(M.xx) + fabsf(M.yy) + fabsf(M.zz));
    ok = ok && (M.xx >= -tol) && (M.yy >= -tol) && (M.zz >= -tol);
  }
  return ok;
}

// ---- class MSTR in GAIN form (K3s): free MONOCULAR point with a full 3-D Gaussian edge.  G = w J^T J has rank 2 (J: rows u, v
// of P R), so D = G + S takes its third direction from the component alone: for a landmark within half a metre of the
// camera (every "monocular" point of the synthetic C2 frames is one: u_r < 0 <=> z < bf / u) G ~ w (f / z)^2 exceeds S by
// more than fp32 carries and the cofactor inverse of D returns noise in that direction (measured: pose gradients wrong by
// 1e3 x, the frame diverges — K3m keeps the class on the fp64 path for this reason).  The same block by the matrix inversion
// lemma needs the component's COVARIANCE Sig = S^-1 instead and only a 2 x 2 inverse of well-scaled positive terms:
//      C = J Sig J^T + I / w,   K = Sig J^T C^-1,   y = e + J d
//      u0 = K y - d,   W = K J,   M = J^T C^-1 J  (positive semidefinite by construction),   v = J^T C^-1 y
// (u0 = D^-1 (h - S d), W = D^-1 G, M = G - G D^-1 G, v = h - G u0 of the general path, identically.)
GNW_HD bool lin_mstr_gain(const Cam& C, const Lin& L, const float* Rf, float w, const S3f& Sig, const V3f& d, PointLin& out, S3f& M,
                          V3f& v) {
  const Proj P = proj_jac(C, L);
  const float iw = rcpf(huber(C, L, false, w));
  const V3f r0 = {P.a * Rf[0] + P.c0 * Rf[6], P.a * Rf[1] + P.c0 * Rf[7], P.a * Rf[2] + P.c0 * Rf[8]};
  const V3f r1 = {P.b * Rf[3] + P.c1 * Rf[6], P.b * Rf[4] + P.c1 * Rf[7], P.b * Rf[5] + P.c1 * Rf[8]};
  const V3f s0 = mulf(Sig, r0), s1 = mulf(Sig, r1);
  const float c00 = dotf(r0, s0) + iw, c01 = dotf(r0, s1), c11 = dotf(r1, s1) + iw;
  const float det = c00 * c11 - c01 * c01;
  const float idet = rcpf(det);
  const float i00 = c11 * idet, i01 = -c01 * idet, i11 = c00 * idet;
  const V3f k0 = {s0.x * i00 + s1.x * i01, s0.y * i00 + s1.y * i01, s0.z * i00 + s1.z * i01};
  const V3f k1 = {s0.x * i01 + s1.x * i11, s0.y * i01 + s1.y * i11, s0.z * i01 + s1.z * i11};
  const float y0 = L.e.x + dotf(r0, d), y1 = L.e.y + dotf(r1, d);
  out.u0 = {k0.x * y0 + k1.x * y1 - d.x, k0.y * y0 + k1.y * y1 - d.y, k0.z * y0 + k1.z * y1 - d.z};
  out.w0 = k0.x * r0.x + k1.x * r1.x; out.w1 = k0.x * r0.y + k1.x * r1.y; out.w2 = k0.x * r0.z + k1.x * r1.z;
  out.w3 = k0.y * r0.x + k1.y * r1.x; out.w4 = k0.y * r0.y + k1.y * r1.y; out.w5 = k0.y * r0.z + k1.y * r1.z;
  out.w6 = k0.z * r0.x + k1.z * r1.x; out.w7 = k0.z * r0.y + k1.z * r1.y; out.w8 = k0.z * r0.z + k1.z * r1.z;
  // M = J^T C^-1 J, v = J^T C^-1 y
  const V3f q0 = {i00 * r0.x + i01 * r1.x, i00 * r0.y + i01 * r1.y, i00 * r0.z + i01 * r1.z};  // row 0 of C^-1 J
  const V3f q1 = {i01 * r0.x + i11 * r1.x, i01 * r0.y + i11 * r1.y, i01 * r0.z + i11 * r1.z};
  M.xx = r0.x * q0.x + r1.x * q1.x;
  M.xy = r0.x * q0.y + r1.x * q1.y;
  M.xz = r0.x * q0.z + r1.x * q1.z;
  M.yy = r0.y * q0.y + r1.y * q1.y;
  M.yz = r0.y * q0.z + r1.y * q1.z;
  M.zz = r0.z * q0.z + r1.z * q1.z;
  v = {q0.x * y0 + q1.x * y1, q0.y * y0 + q1.y * y1, q0.z * y0 + q1.z * y1};
  return (det > 0.0f) && isfinite(idet) && finite3(out.u0);
}

// ---- fp64 evaluation of one free point block in the world frame (any class): the fall-back of K3s when an fp32 block
// reports a singular / non-finite result (a landmark that wandered to the camera plane or to infinity in a frame whose
// iteration does not contract).  S: six unique entries of the structure Hessian (cov_inv, or lam n n^T), bs: its right-hand
// side (-S d, or -lam es n); zero without structure edge.  Same outputs as lin_str.
GNW_HD bool lin_general64(const Cam& C, const double* R, const double* t, double x0, double x1, double x2, const float* ob, bool stereo,
                          double w, const double* S, const double* bs, PointLin& out, S3f& Mo, V3f& vo) {
  const double pc[3] = {R[0] * x0 + R[1] * x1 + R[2] * x2 + t[0], R[3] * x0 + R[4] * x1 + R[5] * x2 + t[1],
                        R[6] * x0 + R[7] * x1 + R[8] * x2 + t[2]};
  const double iz = 1.0 / pc[2];
  const double a = C.fx * iz, b = C.fy * iz, c0 = -a * pc[0] * iz, c1 = -b * pc[1] * iz, cz = C.bf * iz * iz;
  const double u = pc[0] * a + C.cx, vv = pc[1] * b + C.cy;
  const double e[3] = {(double)ob[0] - u, (double)ob[1] - vv, stereo ? (double)ob[2] - (u - C.bf * iz) : 0.0};
  if (C.robust) {
    const double chi = w * (e[0] * e[0] + e[1] * e[1] + e[2] * e[2]);
    const double dl = stereo ? (double)C.d_stereo : (double)C.d_mono;
    if (chi > dl * dl) w *= dl / sqrt(chi);
  }
  double J[9];
  for (int c = 0; c < 3; ++c) {
    J[c] = a * R[c] + c0 * R[6 + c];
    J[3 + c] = b * R[3 + c] + c1 * R[6 + c];
    J[6 + c] = stereo ? J[c] + cz * R[6 + c] : 0.0;
  }
  double G[9], h[3];
  for (int i = 0; i < 3; ++i) {
    for (int j = 0; j < 3; ++j) G[i * 3 + j] = w * (J[i] * J[j] + J[3 + i] * J[3 + j] + J[6 + i] * J[6 + j]);
    h[i] = w * (J[i] * e[0] + J[3 + i] * e[1] + J[6 + i] * e[2]);
  }
  const double Sf[9] = {S[0], S[1], S[2], S[1], S[3], S[4], S[2], S[4], S[5]};
  double D[9], Di[9];
  for (int i = 0; i < 9; ++i) D[i] = G[i] + Sf[i];
  const double det = D[0] * (D[4] * D[8] - D[5] * D[7]) - D[1] * (D[3] * D[8] - D[5] * D[6]) + D[2] * (D[3] * D[7] - D[4] * D[6]);
  const double id = 1.0 / det;
  Di[0] = (D[4] * D[8] - D[5] * D[7]) * id; Di[1] = (D[2] * D[7] - D[1] * D[8]) * id; Di[2] = (D[1] * D[5] - D[2] * D[4]) * id;
  Di[3] = (D[5] * D[6] - D[3] * D[8]) * id; Di[4] = (D[0] * D[8] - D[2] * D[6]) * id; Di[5] = (D[2] * D[3] - D[0] * D[5]) * id;
  Di[6] = (D[3] * D[7] - D[4] * D[6]) * id; Di[7] = (D[1] * D[6] - D[0] * D[7]) * id; Di[8] = (D[0] * D[4] - D[1] * D[3]) * id;
  const double bl[3] = {h[0] + bs[0], h[1] + bs[1], h[2] + bs[2]};
  double u0[3], W[9];
  for (int i = 0; i < 3; ++i) {
    u0[i] = Di[i * 3] * bl[0] + Di[i * 3 + 1] * bl[1] + Di[i * 3 + 2] * bl[2];
    for (int j = 0; j < 3; ++j) W[i * 3 + j] = Di[i * 3] * G[j] + Di[i * 3 + 1] * G[3 + j] + Di[i * 3 + 2] * G[6 + j];
  }
  double Mf[9], vf[3];
  for (int i = 0; i < 3; ++i) {
    vf[i] = h[i] - (G[i * 3] * u0[0] + G[i * 3 + 1] * u0[1] + G[i * 3 + 2] * u0[2]);
    for (int j = 0; j < 3; ++j) Mf[i * 3 + j] = W[i] * Sf[j] + W[3 + i] * Sf[3 + j] + W[6 + i] * Sf[6 + j];
  }
  out.u0 = {(float)u0[0], (float)u0[1], (float)u0[2]};
  out.w0 = (float)W[0]; out.w1 = (float)W[1]; out.w2 = (float)W[2];
  out.w3 = (float)W[3]; out.w4 = (float)W[4]; out.w5 = (float)W[5];
  out.w6 = (float)W[6]; out.w7 = (float)W[7]; out.w8 = (float)W[8];
  Mo = {(float)Mf[0], (float)Mf[1], (float)Mf[2], (float)Mf[4], (float)Mf[5], (float)Mf[8]};
  vo = {(float)vf[0], (float)vf[1], (float)vf[2]};
  bool ok = (det != 0.0) && isfinite(id) && finite3(out.u0) && finite3(vo);
  ok = ok && isfinite(out.w0 + out.w1 + out.w2 + out.w3 + out.w4 + out.w5 + out.w6 + out.w7 + out.w8) &&
       isfinite(Mo.xx + Mo.xy + Mo.xz + Mo.yy + Mo.yz + Mo.zz);
  return ok;
}

// ---- class STRP: free stereo point on a planar component (EdgePt2GaussianDeg: e_s = n.(x - mu), information lam) ----
// The stereo edge alone determines the point (P square), the plane adds one scalar equation: D = G + lam n n^T by
// Sherman-Morrison.  With pe = J^-1 e (the correction the stereo edge alone asks for), z = G^-1 n, kappa = lam/(1 + lam n.z):
//   u0 = u - kappa z (n.u),  u = pe - lam e_s z;   W = I - kappa z n^T;   M = kappa n n^T,  v = kappa (e_s + n.pe) n
// n: plane normal (world), es = n.(x - mu) (from the fp64 offset).  kappa / rho are returned for the rank-one accumulation.
// compact result of the two planar classes: W = I - kz n^T (n: the plane normal, known from the map), so that the step is
// dx = u0 - a + kz (n . a)
struct PlaneLin {
  V3f u0, kz;
};
GNW_HD bool lin_strp_c(const Cam& C, const Lin& L, const float* Rf, float w, const V3f& n, float es, PlaneLin& out, float& kappa,
                       float& rho) {
  const Proj P = proj_jac(C, L);
  w = huber(C, L, true, w);
  const float ia = rcpf(P.a), ib = rcpf(P.b), icz = rcpf(P.cz);
  // pe_c = P^-1 e
  V3f pe;
  pe.z = (L.e.z - L.e.x) * icz;
  pe.x = (L.e.x - P.c0 * pe.z) * ia;
  pe.y = (L.e.y - P.c1 * pe.z) * ib;
  // z_c = P^-1 P^-T nc / w
  const V3f nc = rot(Rf, n);
  const float q1 = nc.y * ib;
  const float q2 = (nc.z - P.c0 * nc.x * ia - P.c1 * q1) * icz;
  const float q0 = nc.x * ia - q2;
  const float iw = rcpf(w);
  const float zz = (q2 - q0) * icz;
  const V3f zc = {(q0 - P.c0 * zz) * ia * iw, (q1 - P.c1 * zz) * ib * iw, zz * iw};
  // into the world frame
  const V3f pw = rot_t(Rf, pe), zw = rot_t(Rf, zc);
  const float gamma = dotf(n, zw);
  kappa = divf(C.lam, 1.0f + C.lam * gamma);
  const float le = C.lam * es;
  const V3f u = {pw.x - le * zw.x, pw.y - le * zw.y, pw.z - le * zw.z};
  const float nu = kappa * dotf(n, u);
  out.u0 = {u.x - nu * zw.x, u.y - nu * zw.y, u.z - nu * zw.z};
  out.kz = {kappa * zw.x, kappa * zw.y, kappa * zw.z};
  rho = kappa * (es + dotf(n, pw));
  return isfinite(kappa);
}
GNW_HD void expand_plane(const PlaneLin& c, const V3f& n, PointLin& out) {
  out.u0 = c.u0;
  out.w0 = 1.f - c.kz.x * n.x; out.w1 = -c.kz.x * n.y; out.w2 = -c.kz.x * n.z;
  out.w3 = -c.kz.y * n.x; out.w4 = 1.f - c.kz.y * n.y; out.w5 = -c.kz.y * n.z;
  out.w6 = -c.kz.z * n.x; out.w7 = -c.kz.z * n.y; out.w8 = 1.f - c.kz.z * n.z;
}
GNW_HD bool lin_strp(const Cam& C, const Lin& L, const float* Rf, float w, const V3f& n, float es, PointLin& out, float& kappa, float& rho) {
  PlaneLin c;
  const bool ok = lin_strp_c(C, L, Rf, w, n, es, c, kappa, rho);
  expand_plane(c, n, out);
  return ok;
}

// ---- class MPLANE: free MONOCULAR point on a planar component.  Rows u, v of P and the plane normal form a square system
//   Q dc = [e - P_uv A dxi ; -e_s],  Q = [P_uv ; nc^T]   =>   dc = u0_c - W_c (A dxi),  W_c = I - q3 nc^T  (q3 = Q^-1 e_3)
// (= D^-1 (h + s - G A dxi) of the general path; weights and the Huber factor cancel row by row); zero Schur contribution.
GNW_HD bool lin_mplane_c(const Cam& C, const Lin& L, const float* Rf, const V3f& n, float es, PlaneLin& out) {
  const Proj P = proj_jac(C, L);
  const V3f nc = rot(Rf, n);
  const float ia = rcpf(P.a), ib = rcpf(P.b);
  const float den = nc.z - nc.x * P.c0 * ia - nc.y * P.c1 * ib;  // = (nc . pc) / z: bounded away from 0 by the view-cos cull
  const float iden = rcpf(den);
  // u0_c = Q^-1 [e.x, e.y, -es]
  const float uz = (-es - nc.x * L.e.x * ia - nc.y * L.e.y * ib) * iden;
  const V3f u0c = {(L.e.x - P.c0 * uz) * ia, (L.e.y - P.c1 * uz) * ib, uz};
  // q3 = Q^-1 [0, 0, 1]
  const V3f q3 = {-P.c0 * iden * ia, -P.c1 * iden * ib, iden};
  out.u0 = rot_t(Rf, u0c);
  out.kz = rot_t(Rf, q3);
  return finite3(out.u0) && isfinite(iden);
}
GNW_HD bool lin_mplane(const Cam& C, const Lin& L, const float* Rf, const V3f& n, float es, PointLin& out) {
  PlaneLin c;
  const bool ok = lin_mplane_c(C, L, Rf, n, es, c);
  expand_plane(c, n, out);
  return ok;
}

// ---- class FIXED: a held point contributes G, h of its reprojection edge to the pose system ----
GNW_HD void lin_fixed(const Cam& C, const Lin& L, const float* Rf, bool stereo, float w, S3f& G, V3f& h) {
  const Proj P = proj_jac(C, L);
  reproj_world(P, L, Rf, stereo, huber(C, L, stereo, w), G, h);
}

// the pose action at a point: a = omega_w x x' + upsilon_w
GNW_HD V3f twist_at(const V3f& om, const V3f& up, const V3f& xr) {
  return {om.y * xr.z - om.z * xr.y + up.x, om.z * xr.x - om.x * xr.z + up.y, om.x * xr.y - om.y * xr.x + up.z};
}
// dx = u0 - W a
GNW_HD V3f point_step(const V3f& u0, const float* W, const V3f& a) {
  return {u0.x - (W[0] * a.x + W[1] * a.y + W[2] * a.z), u0.y - (W[3] * a.x + W[4] * a.y + W[5] * a.z),
          u0.z - (W[6] * a.x + W[7] * a.y + W[8] * a.z)};
}

// planar classes: dx = u0 - (I - kz n^T) a
GNW_HD V3f plane_step(const V3f& u0, const V3f& kz, const V3f& n, const V3f& a) {
  const float na = dotf(n, a);
  return {u0.x - a.x + kz.x * na, u0.y - a.y + kz.y * na, u0.z - a.z + kz.z * na};
}

// camera-frame Schur terms (pc, M_c, v_c) of the fp64 per-point path -> world-twist coordinates: M_w = R^T M_c R, v_w = R^T v_c
GNW_HD void schur_to_world(const float* Rf, const S3f& Mc, const V3f& vc, S3f& Mw, V3f& vw) {
  float t[9];
#pragma unroll
  for (int c = 0; c < 3; ++c) {  // t = M_c R (columns of R)
    const V3f col = {Rf[c], Rf[3 + c], Rf[6 + c]};
    const V3f mc = mulf(Mc, col);
    t[c] = mc.x; t[3 + c] = mc.y; t[6 + c] = mc.z;
  }
  Mw.xx = Rf[0] * t[0] + Rf[3] * t[3] + Rf[6] * t[6];
  Mw.xy = Rf[0] * t[1] + Rf[3] * t[4] + Rf[6] * t[7];
  Mw.xz = Rf[0] * t[2] + Rf[3] * t[5] + Rf[6] * t[8];
  Mw.yy = Rf[1] * t[1] + Rf[4] * t[4] + Rf[7] * t[7];
  Mw.yz = Rf[1] * t[2] + Rf[4] * t[5] + Rf[7] * t[8];
  Mw.zz = Rf[2] * t[2] + Rf[5] * t[5] + Rf[8] * t[8];
  vw = rot_t(Rf, vc);
}

}  // namespace gnw
}  // namespace gmmloc
