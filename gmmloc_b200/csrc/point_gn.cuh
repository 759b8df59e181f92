// point_gn.cuh — device-side GMMLoc::optimizePoint shared by K4 (point_gn.cu) and the keyframe
// association kernel (view.cu).  INCLUDE ONLY FROM FILES COMPILED WITH -fmad=false: every product and sum
// below is meant to be a separate IEEE rn operation, in the CPU oracle's expression order (the oracle is
// built -ffp-contract=off), so results are bit-identical.
#pragma once
#include "common.cuh"

namespace gmmloc {

// g2o SE3Quat::normalizeRotation (flip to w >= 0, normalise), then Eigen toRotationMatrix
__device__ __forceinline__ void quat_normalize_d(const double* pose, double* q) {
  double x = pose[0], y = pose[1], z = pose[2], w = pose[3];
  if (w < 0) { x = -x; y = -y; z = -z; w = -w; }
  const double n = sqrt(x * x + y * y + z * z + w * w);
  q[0] = x / n; q[1] = y / n; q[2] = z / n; q[3] = w / n;
}
__device__ __forceinline__ void quat_to_rot_d(const double* q, double* R) {
  const double x = q[0], y = q[1], z = q[2], w = q[3];
  const double tx = 2 * x, ty = 2 * y, tz = 2 * z;
  const double twx = tx * w, twy = ty * w, twz = tz * w;
  const double txx = tx * x, txy = ty * x, txz = tz * x;
  const double tyy = ty * y, tyz = tz * y, tzz = tz * z;
  R[0] = 1 - (tyy + tzz); R[1] = txy - twz;       R[2] = txz + twy;
  R[3] = txy + twz;       R[4] = 1 - (txx + tzz); R[5] = tyz - twx;
  R[6] = txz - twy;       R[7] = tyz + twx;       R[8] = 1 - (txx + tyy);
}
__device__ __forceinline__ void quat_to_rot_n(const double* pose, double* R) {
  double q[4];
  quat_normalize_d(pose, q);
  quat_to_rot_d(q, R);
}

__device__ __forceinline__ double cof(const double* m, int i, int j) {
  const int i1 = (i + 1) % 3, i2 = (i + 2) % 3, j1 = (j + 1) % 3, j2 = (j + 2) % 3;
  return m[i1 * 3 + j1] * m[i2 * 3 + j2] - m[i1 * 3 + j2] * m[i2 * 3 + j1];
}

// 3x3 inverse by cofactors (same tree as the map build / oracle inverse3); false when singular
__device__ __forceinline__ bool inverse3_d(const double* H, double* Hi) {
  const double c00 = cof(H, 0, 0), c10 = cof(H, 1, 0), c20 = cof(H, 2, 0);
  const double det = (c00 * H[0] + c10 * H[3]) + c20 * H[6];
  const double invdet = 1.0 / det;
  Hi[0] = c00 * invdet;
  Hi[1] = c10 * invdet;
  Hi[2] = c20 * invdet;
  Hi[3] = cof(H, 0, 1) * invdet;
  Hi[4] = cof(H, 1, 1) * invdet;
  Hi[5] = cof(H, 2, 1) * invdet;
  Hi[6] = cof(H, 0, 2) * invdet;
  Hi[7] = cof(H, 1, 2) * invdet;
  Hi[8] = cof(H, 2, 2) * invdet;
  return isfinite(invdet) && det != 0.0;
}

struct PointSolve {
  double x[3];
  double chi2_proj, chi2_str;
  bool res;
};

// GMMLoc::optimizePoint (gmmloc_opt.cpp:260-342): one stereo reprojection edge (factors.cpp:116-168,
// information w_proj I3) + one point-to-plane edge (factors.cpp:55-64, information w_str = tri_lambda2 *
// proj_z2), Gauss-Newton, `iters` iterations without early exit; chi2 of the last linearisation point;
// res = !(chi2_proj > 7.815) && !(tri_check_str_chi2 && chi2_str > tri_str_thresh * tri_lambda2).
__device__ __forceinline__ PointSolve optimize_point_d(const FrameParamsDev& P, const double* R, const double* t,
                                                       const double* o, double w_proj, double w_str, const double* n,
                                                       const double* mu, const double* x0, int iters) {
  PointSolve s;
  double x[3] = {x0[0], x0[1], x0[2]};
  double c2p = 0, c2s = 0;
  bool ok = true;
  for (int it = 0; it < iters; ++it) {
    const double px = ((R[0] * x[0] + R[1] * x[1]) + R[2] * x[2]) + t[0];
    const double py = ((R[3] * x[0] + R[4] * x[1]) + R[5] * x[2]) + t[1];
    const double pz = ((R[6] * x[0] + R[7] * x[1]) + R[8] * x[2]) + t[2];
    const double z_2 = pz * pz;
    const double invz = 1.0 / pz;
    const double u = px * invz * P.fx + P.cx;
    const double v = py * invz * P.fy + P.cy;
    const double ur = u - P.bf * invz;
    const double e[3] = {o[0] - u, o[1] - v, o[2] - ur};
    double Jl[9];
    for (int j = 0; j < 3; ++j) {
      Jl[j] = -P.fx * R[j] / pz + P.fx * px * R[6 + j] / z_2;
      Jl[3 + j] = -P.fy * R[3 + j] / pz + P.fy * py * R[6 + j] / z_2;
      Jl[6 + j] = Jl[j] - P.bf * R[6 + j] / z_2;
    }
    const double es = (n[0] * (x[0] - mu[0]) + n[1] * (x[1] - mu[1])) + n[2] * (x[2] - mu[2]);
    c2p = w_proj * ((e[0] * e[0] + e[1] * e[1]) + e[2] * e[2]);
    c2s = w_str * (es * es);
    double H[9], b[3];
    for (int r = 0; r < 3; ++r) {
      for (int c = 0; c < 3; ++c)
        H[r * 3 + c] = w_proj * ((Jl[r] * Jl[c] + Jl[3 + r] * Jl[3 + c]) + Jl[6 + r] * Jl[6 + c]) + w_str * (n[r] * n[c]);
      b[r] = -(w_proj * ((Jl[r] * e[0] + Jl[3 + r] * e[1]) + Jl[6 + r] * e[2]) + w_str * (n[r] * es));
    }
    double Hi[9];
    if (!inverse3_d(H, Hi)) {
      ok = false;
      break;
    }
    for (int r = 0; r < 3; ++r) x[r] += (Hi[r * 3] * b[0] + Hi[r * 3 + 1] * b[1]) + Hi[r * 3 + 2] * b[2];
  }
  for (int r = 0; r < 3; ++r) s.x[r] = x[r];
  s.chi2_proj = c2p;
  s.chi2_str = c2s;
  bool res = ok;
  if (c2p > 7.815) res = false;                                                     // gmmloc_opt.cpp:332-335
  if (P.tri_check_str_chi2 && c2s > P.tri_str_thresh * P.tri_lambda2) res = false;  // gmmloc_opt.cpp:336-339
  s.res = res;
  return s;
}

}  // namespace gmmloc
