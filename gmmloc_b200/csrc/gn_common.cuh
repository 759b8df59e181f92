// gn_common.cuh — device helpers shared by the Gauss-Newton / Levenberg-Marquardt kernels (frame_gn.cu,
// pose_lm.cu): g2o SE3Quat semantics and the 6x6 SPD solve of the reduced pose system.
#pragma once
#include "common.cuh"

namespace gmmloc {
namespace {

struct Sym3 {
  double xx, xy, xz, yy, yz, zz;
};
struct Vec3 {
  double x, y, z;
};

__device__ __forceinline__ Vec3 rot(const double* R, const Vec3& v) {
  return {R[0] * v.x + R[1] * v.y + R[2] * v.z, R[3] * v.x + R[4] * v.y + R[5] * v.z, R[6] * v.x + R[7] * v.y + R[8] * v.z};
}
__device__ __forceinline__ Vec3 rot_t(const double* R, const Vec3& v) {
  return {R[0] * v.x + R[3] * v.y + R[6] * v.z, R[1] * v.x + R[4] * v.y + R[7] * v.z, R[2] * v.x + R[5] * v.y + R[8] * v.z};
}
__device__ __forceinline__ Vec3 mul(const Sym3& A, const Vec3& v) {
  return {A.xx * v.x + A.xy * v.y + A.xz * v.z, A.xy * v.x + A.yy * v.y + A.yz * v.z, A.xz * v.x + A.yz * v.y + A.zz * v.z};
}
__device__ __forceinline__ Vec3 cross(const Vec3& a, const Vec3& b) {
  return {a.y * b.z - a.z * b.y, a.z * b.x - a.x * b.z, a.x * b.y - a.y * b.x};
}
// inverse of a symmetric 3x3 by cofactors; ok=false when det is zero / non-finite (Hll singular)
__device__ __forceinline__ Sym3 inv_sym(const Sym3& D, bool& ok) {
  const double c00 = D.yy * D.zz - D.yz * D.yz;
  const double c01 = D.xz * D.yz - D.xy * D.zz;
  const double c02 = D.xy * D.yz - D.xz * D.yy;
  const double det = D.xx * c00 + D.xy * c01 + D.xz * c02;
  const double id = 1.0 / det;
  ok = (det != 0.0) && isfinite(id);
  return {c00 * id, c01 * id, c02 * id, (D.xx * D.zz - D.xz * D.xz) * id, (D.xy * D.xz - D.xx * D.yz) * id,
          (D.xx * D.yy - D.xy * D.xy) * id};
}

// reprojection edge at camera-frame position pc: G = w P^T P, h = w P^T e with P = d proj / d pc
// (EdgeSE3ProjectXYZ / EdgeStereoSE3ProjectXYZ[OnlyPose]; Huber weight folded into w when robust)
__device__ __forceinline__ void reproj(const FrameParamsDev& p, const Vec3& pc, float o0, float o1, float o2, bool stereo,
                                       double w, bool robust, Sym3& G, Vec3& h) {
  const double invz = 1.0 / pc.z;
  const double a = p.fx * invz, b = p.fy * invz;
  const double u = pc.x * a + p.cx, v = pc.y * b + p.cy;
  const double e0 = (double)o0 - u, e1 = (double)o1 - v;
  const double c0 = -a * pc.x * invz, c1 = -b * pc.y * invz;
  double e2 = 0.0, c2 = 0.0, a2 = 0.0;
  if (stereo) {
    const double bz = p.bf * invz;
    e2 = (double)o2 - (u - bz);
    c2 = c0 + bz * invz;
    a2 = a;
  }
  if (robust) {  // g2o RobustKernelHuber: weight rho1 = delta / sqrt(e2) beyond delta^2
    const double chi = w * (e0 * e0 + e1 * e1 + e2 * e2);
    const double delta = stereo ? p.delta_stereo : p.delta_mono;
    if (chi > delta * delta) w *= delta / sqrt(chi);
  }
  G.xx = w * (a * a + a2 * a2);
  G.xy = 0.0;
  G.xz = w * (a * c0 + a2 * c2);
  G.yy = w * (b * b);
  G.yz = w * (b * c1);
  G.zz = w * (c0 * c0 + c1 * c1 + c2 * c2);
  h.x = w * (a * e0 + a2 * e2);
  h.y = w * (b * e1);
  h.z = w * (c0 * e0 + c1 * e1 + c2 * e2);
}

// pose-system contribution of one point: acc[0..20] += upper triangle of A^T M A, acc[21..26] += A^T v,
// A = d pc / d xi = [-[pc]x  I]
__device__ __forceinline__ void accumulate_pose(double* acc, const Vec3& pc, const Sym3& M, const Vec3& v) {
  // A^T M A and A^T v with A = [-[pc]x  I]:  C = [pc]x M,  [pc]x = [[0,-z,y],[z,0,-x],[-y,x,0]]
  const double C00 = pc.y * M.xz - pc.z * M.xy, C01 = pc.y * M.yz - pc.z * M.yy, C02 = pc.y * M.zz - pc.z * M.yz;
  const double C10 = pc.z * M.xx - pc.x * M.xz, C11 = pc.z * M.xy - pc.x * M.yz, C12 = pc.z * M.xz - pc.x * M.zz;
  const double C20 = pc.x * M.xy - pc.y * M.xx, C21 = pc.x * M.yy - pc.y * M.xy, C22 = pc.x * M.yz - pc.y * M.xz;
  // rot-rot block = C [pc]x^T,  (C [pc]x^T)_{ab} = sum_c C_ac [pc]x_{bc}
  acc[0] += C02 * pc.y - C01 * pc.z;   // (0,0)
  acc[1] += C00 * pc.z - C02 * pc.x;   // (0,1)
  acc[2] += C01 * pc.x - C00 * pc.y;   // (0,2)
  acc[3] += C00;           // (0,3..5): rot-trans block = C
  acc[4] += C01;
  acc[5] += C02;
  acc[6] += C10 * pc.z - C12 * pc.x;   // (1,1)
  acc[7] += C11 * pc.x - C10 * pc.y;   // (1,2)
  acc[8] += C10;
  acc[9] += C11;
  acc[10] += C12;
  acc[11] += C21 * pc.x - C20 * pc.y;  // (2,2)
  acc[12] += C20;
  acc[13] += C21;
  acc[14] += C22;
  acc[15] += M.xx;         // trans-trans block = M
  acc[16] += M.xy;
  acc[17] += M.xz;
  acc[18] += M.yy;
  acc[19] += M.yz;
  acc[20] += M.zz;
  const Vec3 br = cross(pc, v);  // A^T v = [pc x v ; v]
  acc[21] += br.x;
  acc[22] += br.y;
  acc[23] += br.z;
  acc[24] += v.x;
  acc[25] += v.y;
  acc[26] += v.z;
}

// ---- SE3 (g2o SE3Quat semantics, SURVEY Appx B.1), thread-0 only ---------------------------------
__host__ __device__ inline void quat_normalize(double* q) {
  if (q[3] < 0)
    for (int i = 0; i < 4; ++i) q[i] = -q[i];
  const double n = sqrt(q[0] * q[0] + q[1] * q[1] + q[2] * q[2] + q[3] * q[3]);
  for (int i = 0; i < 4; ++i) q[i] /= n;
}
__host__ __device__ inline void quat_to_rot(const double* q, double* R) {
  const double x = q[0], y = q[1], z = q[2], w = q[3];
  const double tx = 2 * x, ty = 2 * y, tz = 2 * z;
  const double twx = tx * w, twy = ty * w, twz = tz * w;
  const double txx = tx * x, txy = ty * x, txz = tz * x;
  const double tyy = ty * y, tyz = tz * y, tzz = tz * z;
  R[0] = 1 - (tyy + tzz); R[1] = txy - twz;       R[2] = txz + twy;
  R[3] = txy + twz;       R[4] = 1 - (txx + tzz); R[5] = tyz - twx;
  R[6] = txz - twy;       R[7] = tyz + twx;       R[8] = 1 - (txx + tyy);
}
__device__ inline void rot_to_quat(const double* R, double* q) {
  double t = R[0] + R[4] + R[8];
  if (t > 0) {
    t = sqrt(t + 1.0);
    q[3] = 0.5 * t;
    t = 0.5 / t;
    q[0] = (R[7] - R[5]) * t;
    q[1] = (R[2] - R[6]) * t;
    q[2] = (R[3] - R[1]) * t;
  } else {
    int i = 0;
    if (R[4] > R[0]) i = 1;
    if (R[8] > R[i * 4]) i = 2;
    const int j = (i + 1) % 3, k = (j + 1) % 3;
    t = sqrt(R[i * 4] - R[j * 4] - R[k * 4] + 1.0);
    q[i] = 0.5 * t;
    t = 0.5 / t;
    q[3] = (R[k * 3 + j] - R[j * 3 + k]) * t;
    q[j] = (R[j * 3 + i] + R[i * 3 + j]) * t;
    q[k] = (R[k * 3 + i] + R[i * 3 + k]) * t;
  }
}
// pose <- exp(xi) * pose   (VertexSE3Expmap::oplusImpl)
__device__ inline void se3_oplus(const double* xi, const double* pose, double* out) {
  const double w0 = xi[0], w1 = xi[1], w2 = xi[2];
  const double theta = sqrt(w0 * w0 + w1 * w1 + w2 * w2);
  const double Om[9] = {0, -w2, w1, w2, 0, -w0, -w1, w0, 0};
  double Om2[9];
  for (int r = 0; r < 3; ++r)
    for (int c = 0; c < 3; ++c) Om2[r * 3 + c] = Om[r * 3] * Om[c] + Om[r * 3 + 1] * Om[3 + c] + Om[r * 3 + 2] * Om[6 + c];
  double a, b, d;
  if (theta < 1e-5) {
    a = 1.0; b = 0.5; d = 1.0 / 6.0;
  } else {
    const double st = sin(theta), ct = cos(theta);
    a = st / theta;
    b = (1 - ct) / (theta * theta);
    d = (theta - st) / (theta * theta * theta);
  }
  double Rm[9], V[9];
  for (int i = 0; i < 9; ++i) {
    const double I = (i % 4 == 0) ? 1.0 : 0.0;
    Rm[i] = I + a * Om[i] + b * Om2[i];
    V[i] = I + b * Om[i] + d * Om2[i];
  }
  const double t[3] = {V[0] * xi[3] + V[1] * xi[4] + V[2] * xi[5], V[3] * xi[3] + V[4] * xi[4] + V[5] * xi[5],
                       V[6] * xi[3] + V[7] * xi[4] + V[8] * xi[5]};
  double qd[4], Rd[9];
  rot_to_quat(Rm, qd);
  quat_normalize(qd);
  quat_to_rot(qd, Rd);
  const double* p = pose + 4;
  const double rt[3] = {Rd[0] * p[0] + Rd[1] * p[1] + Rd[2] * p[2], Rd[3] * p[0] + Rd[4] * p[1] + Rd[5] * p[2],
                        Rd[6] * p[0] + Rd[7] * p[1] + Rd[8] * p[2]};
  const double ax = qd[0], ay = qd[1], az = qd[2], aw = qd[3], bx = pose[0], by = pose[1], bz = pose[2], bw = pose[3];
  double q[4];
  q[3] = aw * bw - ax * bx - ay * by - az * bz;
  q[0] = aw * bx + ax * bw + ay * bz - az * by;
  q[1] = aw * by + ay * bw + az * bx - ax * bz;
  q[2] = aw * bz + az * bw + ax * by - ay * bx;
  quat_normalize(q);
  for (int i = 0; i < 4; ++i) out[i] = q[i];
  for (int i = 0; i < 3; ++i) out[4 + i] = t[i] + rt[i];
}

// 6x6 SPD solve from the 21 upper-triangular entries (row-major packing) — plain Cholesky.
__device__ inline bool solve6(const double* acc, double* x) {
  double H[36], L[36];
  int n = 0;
  for (int r = 0; r < 6; ++r)
    for (int c = r; c < 6; ++c) {
      H[r * 6 + c] = acc[n];
      H[c * 6 + r] = acc[n];
      ++n;
    }
  for (int i = 0; i < 36; ++i) L[i] = 0.0;
  for (int k = 0; k < 6; ++k) {
    double s = H[k * 6 + k];
    for (int j = 0; j < k; ++j) s -= L[k * 6 + j] * L[k * 6 + j];
    if (!(s > 0.0) || !isfinite(s)) return false;
    const double d = sqrt(s);
    L[k * 6 + k] = d;
    for (int i = k + 1; i < 6; ++i) {
      double t = H[i * 6 + k];
      for (int j = 0; j < k; ++j) t -= L[i * 6 + j] * L[k * 6 + j];
      L[i * 6 + k] = t / d;
    }
  }
  double y[6];
  for (int i = 0; i < 6; ++i) {
    double t = acc[21 + i];
    for (int j = 0; j < i; ++j) t -= L[i * 6 + j] * y[j];
    y[i] = t / L[i * 6 + i];
  }
  for (int i = 5; i >= 0; --i) {
    double t = y[i];
    for (int j = i + 1; j < 6; ++j) t -= L[j * 6 + i] * x[j];
    x[i] = t / L[i * 6 + i];
  }
  for (int i = 0; i < 6; ++i)
    if (!isfinite(x[i])) return false;
  return true;
}

// ---- latency-optimised twins for the serial section of the mixed-precision frame GN (one thread per frame and
// iteration; the other warps of the CTA wait for this dependent fp64 chain).  Same mathematics, different rounding:
// a fully unrolled LDL^T (six reciprocals, no square roots, everything in registers) and an SE3 update that goes
// straight from the rotation vector to the quaternion (one sincos, reciprocal square roots instead of divisions).
__host__ __device__ __forceinline__ bool solve6_ldlt(const double* acc, double* x) {
  double A[6][6];  // lower triangle; fully unrolled => registers
  {
    int n = 0;
#pragma unroll
    for (int r = 0; r < 6; ++r)
#pragma unroll
      for (int c = r; c < 6; ++c) A[c][r] = acc[n++];
  }
  double inv[6];
  bool ok = true;
#pragma unroll
  for (int k = 0; k < 6; ++k) {
    const double dk = A[k][k];
    ok = ok && (dk > 0.0) && isfinite(dk);  // same acceptance as the Cholesky pivot test of solve6
    inv[k] = 1.0 / dk;
    double col[6], l[6];
#pragma unroll
    for (int i = k + 1; i < 6; ++i) {
      col[i] = A[i][k];        // L_ik d_k
      l[i] = col[i] * inv[k];  // L_ik
    }
#pragma unroll
    for (int i = k + 1; i < 6; ++i) {
#pragma unroll
      for (int j = k + 1; j <= i; ++j) A[i][j] -= l[i] * col[j];
      A[i][k] = l[i];
    }
  }
  double y[6];
#pragma unroll
  for (int i = 0; i < 6; ++i) {
    double t = acc[21 + i];
#pragma unroll
    for (int j = 0; j < i; ++j) t -= A[i][j] * y[j];
    y[i] = t;
  }
#pragma unroll
  for (int i = 5; i >= 0; --i) {
    double t = y[i] * inv[i];
#pragma unroll
    for (int j = i + 1; j < 6; ++j) t -= A[j][i] * x[j];
    x[i] = t;
  }
#pragma unroll
  for (int i = 0; i < 6; ++i) ok = ok && isfinite(x[i]);
  return ok;
}

// pose <- exp(xi) * pose (VertexSE3Expmap::oplusImpl; SE3Quat::exp semantics as se3_oplus), R_out = rotation of the result
__host__ __device__ __forceinline__ void se3_oplus_fast(const double* xi, const double* pose, double* out, double* R_out) {
  const double w0 = xi[0], w1 = xi[1], w2 = xi[2];
  const double th2 = w0 * w0 + w1 * w1 + w2 * w2;
  double b, d, k, c2;  // b = (1 - cos t) / t^2, d = (t - sin t) / t^3, k = sin(t/2) / t, c2 = cos(t/2)
  if (th2 < 1e-10) {
    b = 0.5; d = 1.0 / 6.0;
    k = 0.5 - th2 * (1.0 / 48.0);
    c2 = 1.0 - th2 * 0.125;
  } else {
    const double it = rsqrt(th2), theta = th2 * it;
    double s2;
    sincos(0.5 * theta, &s2, &c2);
    const double st = 2.0 * s2 * c2;
    b = 2.0 * s2 * s2 * (it * it);
    d = (theta - st) * (it * it * it);
    k = s2 * it;
  }
  // V = I + b Om + d Om^2, t = V * upsilon  (Om^2 v = w (w.v) - |w|^2 v)
  const double u0 = xi[3], u1 = xi[4], u2 = xi[5];
  const double wu = w0 * u0 + w1 * u1 + w2 * u2;
  const double t0 = u0 + b * (w1 * u2 - w2 * u1) + d * (w0 * wu - th2 * u0);
  const double t1 = u1 + b * (w2 * u0 - w0 * u2) + d * (w1 * wu - th2 * u1);
  const double t2 = u2 + b * (w0 * u1 - w1 * u0) + d * (w2 * wu - th2 * u2);
  // rotation of exp(xi) as a unit quaternion (w >= 0 for |theta| < pi)
  const double ax = w0 * k, ay = w1 * k, az = w2 * k, aw = c2;
  double Rd[9];
  {
    const double qd[4] = {ax, ay, az, aw};
    quat_to_rot(qd, Rd);
  }
  const double* p = pose + 4;
  const double r0 = Rd[0] * p[0] + Rd[1] * p[1] + Rd[2] * p[2], r1 = Rd[3] * p[0] + Rd[4] * p[1] + Rd[5] * p[2],
               r2 = Rd[6] * p[0] + Rd[7] * p[1] + Rd[8] * p[2];
  const double bx = pose[0], by = pose[1], bz = pose[2], bw = pose[3];
  double q[4];
  q[3] = aw * bw - ax * bx - ay * by - az * bz;
  q[0] = aw * bx + ax * bw + ay * bz - az * by;
  q[1] = aw * by + ay * bw + az * bx - ax * bz;
  q[2] = aw * bz + az * bw + ax * by - ay * bx;
  double n = rsqrt(q[0] * q[0] + q[1] * q[1] + q[2] * q[2] + q[3] * q[3]);
  if (q[3] < 0) n = -n;
#pragma unroll
  for (int i = 0; i < 4; ++i) out[i] = q[i] * n;
  out[4] = t0 + r0;
  out[5] = t1 + r1;
  out[6] = t2 + r2;
  quat_to_rot(out, R_out);
}


}  // namespace
}  // namespace gmmloc
