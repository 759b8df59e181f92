// gmmloc_oracle.cpp — CPU ORACLE (test infrastructure, NOT product code; see gmmloc_oracle.h).
//
// Plain C++17, fp64, no dependencies.  Compile with -ffp-contract=off so every a*b+c below is two
// IEEE roundings, exactly as written (the GPU fp64 re-rank replays the same operation order).
// Citations are relative to /root/reference/.  g2o / Eigen arithmetic is restated from their
// published algorithms (SURVEY.md Appx B): parity unpinned.
#include "gmmloc_oracle.h"

#include <algorithm>
#include <cfloat>
#include <cmath>
#include <cstdio>
#include <cstring>
#include <limits>
#include <thread>
#include <vector>

namespace {

// ----------------------------------------------------------------------------------------------
// small dense helpers (row-major 3x3)
// ----------------------------------------------------------------------------------------------
inline double& M(double* a, int r, int c) { return a[r * 3 + c]; }
inline double M(const double* a, int r, int c) { return a[r * 3 + c]; }

// Eigen's fixed-size 3x3 determinant (Determinant.h, bruteforce_det3_helper):
//   det = h(0,1,2) - h(1,0,2) + h(2,0,1),  h(a,b,c) = m(0,a) * (m(1,b) m(2,c) - m(1,c) m(2,b))
inline double det3_helper(const double* m, int a, int b, int c) {
  return M(m, 0, a) * (M(m, 1, b) * M(m, 2, c) - M(m, 1, c) * M(m, 2, b));
}
inline double det3(const double* m) { return det3_helper(m, 0, 1, 2) - det3_helper(m, 1, 0, 2) + det3_helper(m, 2, 0, 1); }

// Eigen's fixed-size 3x3 inverse (InverseImpl.h, compute_inverse<.,.,3>): cofactors, det from the
// first cofactor column, one reciprocal, nine products.
inline double cofactor3(const double* m, int i, int j) {
  const int i1 = (i + 1) % 3, i2 = (i + 2) % 3, j1 = (j + 1) % 3, j2 = (j + 2) % 3;
  return M(m, i1, j1) * M(m, i2, j2) - M(m, i1, j2) * M(m, i2, j1);
}
inline bool inverse3(const double* m, double* out, double* det_out = nullptr) {
  const double c00 = cofactor3(m, 0, 0), c10 = cofactor3(m, 1, 0), c20 = cofactor3(m, 2, 0);
  const double det = (c00 * M(m, 0, 0) + c10 * M(m, 1, 0)) + c20 * M(m, 2, 0);
  if (det_out) *det_out = det;
  const double invdet = 1.0 / det;
  M(out, 0, 0) = c00 * invdet;
  M(out, 0, 1) = c10 * invdet;
  M(out, 0, 2) = c20 * invdet;
  M(out, 1, 0) = cofactor3(m, 0, 1) * invdet;
  M(out, 1, 1) = cofactor3(m, 1, 1) * invdet;
  M(out, 1, 2) = cofactor3(m, 2, 1) * invdet;
  M(out, 2, 0) = cofactor3(m, 0, 2) * invdet;
  M(out, 2, 1) = cofactor3(m, 1, 2) * invdet;
  M(out, 2, 2) = cofactor3(m, 2, 2) * invdet;
  return std::isfinite(invdet) && det != 0.0;
}

// Unblocked lower Cholesky of a 3x3 SPD matrix (Eigen LLT llt_inplace::unblocked order).
inline bool llt3(const double* a, double* L) {
  for (int i = 0; i < 9; ++i) L[i] = 0.0;
  bool ok = true;
  for (int k = 0; k < 3; ++k) {
    double x = M(a, k, k);
    for (int j = 0; j < k; ++j) x -= M(L, k, j) * M(L, k, j);
    if (!(x > 0.0)) ok = false;
    const double d = std::sqrt(x);
    M(L, k, k) = d;
    for (int i = k + 1; i < 3; ++i) {
      double s = M(a, i, k);
      for (int j = 0; j < k; ++j) s -= M(L, i, j) * M(L, k, j);
      M(L, i, k) = s / d;
    }
  }
  return ok;
}

// Cyclic Jacobi eigen-decomposition of a symmetric 3x3 (only + - * / sqrt so that a GPU replay is
// bit-identical).  Stand-in for Eigen::SelfAdjointEigenSolver<Matrix3d> (gaussian.cpp:39-42): the
// eigenvalues agree to rounding, eigenvector signs are arbitrary in both.  Output ascending.
void eig3_jacobi(const double* cov, double* evals, double* evecs /* evecs[r*3+c]: column c */) {
  double a[9], v[9] = {1, 0, 0, 0, 1, 0, 0, 0, 1};
  for (int i = 0; i < 9; ++i) a[i] = cov[i];
  // symmetrise from the upper triangle (the loader keeps Σ as stored; asymmetry ≤ 6e-17 in shipped maps)
  M(a, 1, 0) = M(a, 0, 1); M(a, 2, 0) = M(a, 0, 2); M(a, 2, 1) = M(a, 1, 2);
  const int P[3] = {0, 0, 1}, Q[3] = {1, 2, 2};
  for (int sweep = 0; sweep < 64; ++sweep) {
    const double off = std::fabs(M(a, 0, 1)) + std::fabs(M(a, 0, 2)) + std::fabs(M(a, 1, 2));
    if (off == 0.0) break;
    for (int r = 0; r < 3; ++r) {
      const int p = P[r], q = Q[r];
      const double apq = M(a, p, q);
      if (apq == 0.0) continue;
      const double app = M(a, p, p), aqq = M(a, q, q);
      const double g = 100.0 * std::fabs(apq);
      // after a few sweeps drop off-diagonals that no longer change the diagonal at fp64 precision
      if (sweep > 3 && std::fabs(app) + g == std::fabs(app) && std::fabs(aqq) + g == std::fabs(aqq)) {
        M(a, p, q) = 0.0; M(a, q, p) = 0.0;
        continue;
      }
      const double h = aqq - app;
      double t;
      if (std::fabs(h) + g == std::fabs(h)) {
        t = apq / h;
      } else {
        const double theta = 0.5 * h / apq;
        t = 1.0 / (std::fabs(theta) + std::sqrt(1.0 + theta * theta));
        if (theta < 0.0) t = -t;
      }
      const double c = 1.0 / std::sqrt(1.0 + t * t);
      const double s = t * c;
      const double tau = s / (1.0 + c);
      M(a, p, p) = app - t * apq;
      M(a, q, q) = aqq + t * apq;
      M(a, p, q) = 0.0; M(a, q, p) = 0.0;
      const int k = 3 - p - q;  // the remaining index
      const double akp = M(a, k, p), akq = M(a, k, q);
      const double nkp = akp - s * (akq + akp * tau);
      const double nkq = akq + s * (akp - akq * tau);
      M(a, k, p) = nkp; M(a, p, k) = nkp;
      M(a, k, q) = nkq; M(a, q, k) = nkq;
      for (int i = 0; i < 3; ++i) {
        const double vip = M(v, i, p), viq = M(v, i, q);
        M(v, i, p) = vip - s * (viq + vip * tau);
        M(v, i, q) = viq + s * (vip - viq * tau);
      }
    }
  }
  int idx[3] = {0, 1, 2};
  double d[3] = {M(a, 0, 0), M(a, 1, 1), M(a, 2, 2)};
  // stable 3-element insertion sort ascending
  for (int i = 1; i < 3; ++i)
    for (int j = i; j > 0 && d[idx[j]] < d[idx[j - 1]]; --j) std::swap(idx[j], idx[j - 1]);
  for (int c = 0; c < 3; ++c) {
    evals[c] = d[idx[c]];
    for (int r = 0; r < 3; ++r) M(evecs, r, c) = M(v, r, idx[c]);
  }
}

// ----------------------------------------------------------------------------------------------
// SE3 (g2o SE3Quat semantics, SURVEY Appx B.1).  pose = qx qy qz qw tx ty tz
// ----------------------------------------------------------------------------------------------
void quat_normalize(double q[4]) {
  // g2o SE3Quat::normalizeRotation: flip to w >= 0, then normalise
  if (q[3] < 0) for (int i = 0; i < 4; ++i) q[i] = -q[i];
  const double n = std::sqrt(q[0] * q[0] + q[1] * q[1] + q[2] * q[2] + q[3] * q[3]);
  for (int i = 0; i < 4; ++i) q[i] /= n;
}
void quat_to_rot(const double q[4], double R[9]) {
  // Eigen QuaternionBase::toRotationMatrix
  const double x = q[0], y = q[1], z = q[2], w = q[3];
  const double tx = 2 * x, ty = 2 * y, tz = 2 * z;
  const double twx = tx * w, twy = ty * w, twz = tz * w;
  const double txx = tx * x, txy = ty * x, txz = tz * x;
  const double tyy = ty * y, tyz = tz * y, tzz = tz * z;
  M(R, 0, 0) = 1 - (tyy + tzz); M(R, 0, 1) = txy - twz;       M(R, 0, 2) = txz + twy;
  M(R, 1, 0) = txy + twz;       M(R, 1, 1) = 1 - (txx + tzz); M(R, 1, 2) = tyz - twx;
  M(R, 2, 0) = txz - twy;       M(R, 2, 1) = tyz + twx;       M(R, 2, 2) = 1 - (txx + tyy);
}
void rot_to_quat(const double R[9], double q[4]) {
  // Eigen quaternionbase_assign_impl<Matrix3>: Shepperd's method
  double t = M(R, 0, 0) + M(R, 1, 1) + M(R, 2, 2);
  if (t > 0) {
    t = std::sqrt(t + 1.0);
    q[3] = 0.5 * t;
    t = 0.5 / t;
    q[0] = (M(R, 2, 1) - M(R, 1, 2)) * t;
    q[1] = (M(R, 0, 2) - M(R, 2, 0)) * t;
    q[2] = (M(R, 1, 0) - M(R, 0, 1)) * t;
  } else {
    int i = 0;
    if (M(R, 1, 1) > M(R, 0, 0)) i = 1;
    if (M(R, 2, 2) > M(R, i, i)) i = 2;
    const int j = (i + 1) % 3, k = (j + 1) % 3;
    t = std::sqrt(M(R, i, i) - M(R, j, j) - M(R, k, k) + 1.0);
    q[i] = 0.5 * t;
    t = 0.5 / t;
    q[3] = (M(R, k, j) - M(R, j, k)) * t;
    q[j] = (M(R, j, i) + M(R, i, j)) * t;
    q[k] = (M(R, k, i) + M(R, i, k)) * t;
  }
}
void quat_mul(const double a[4], const double b[4], double o[4]) {
  const double ax = a[0], ay = a[1], az = a[2], aw = a[3], bx = b[0], by = b[1], bz = b[2], bw = b[3];
  o[3] = aw * bw - ax * bx - ay * by - az * bz;
  o[0] = aw * bx + ax * bw + ay * bz - az * by;
  o[1] = aw * by + ay * bw + az * bx - ax * bz;
  o[2] = aw * bz + az * bw + ax * by - ay * bx;
}
inline void mat3_vec(const double* A, const double* x, double* y) {
  for (int r = 0; r < 3; ++r) y[r] = (M(A, r, 0) * x[0] + M(A, r, 1) * x[1]) + M(A, r, 2) * x[2];
}
void se3_exp_Rt(const double xi[6], double R[9], double t[3]) {
  const double w[3] = {xi[0], xi[1], xi[2]}, u[3] = {xi[3], xi[4], xi[5]};
  const double theta = std::sqrt(w[0] * w[0] + w[1] * w[1] + w[2] * w[2]);
  const double Om[9] = {0, -w[2], w[1], w[2], 0, -w[0], -w[1], w[0], 0};
  double Om2[9];
  for (int r = 0; r < 3; ++r)
    for (int c = 0; c < 3; ++c) M(Om2, r, c) = (M(Om, r, 0) * M(Om, 0, c) + M(Om, r, 1) * M(Om, 1, c)) + M(Om, r, 2) * M(Om, 2, c);
  double a, b, c2, d;
  if (theta < 1e-5) {  // series (newer g2o se3quat.h small-angle branch)
    a = 1.0; b = 0.5; c2 = 0.5; d = 1.0 / 6.0;
  } else {
    a = std::sin(theta) / theta;
    b = (1 - std::cos(theta)) / (theta * theta);
    c2 = b;
    d = (theta - std::sin(theta)) / (theta * theta * theta);
  }
  double V[9];
  for (int i = 0; i < 9; ++i) {
    const double I = (i % 4 == 0) ? 1.0 : 0.0;
    R[i] = I + a * Om[i] + b * Om2[i];
    V[i] = I + c2 * Om[i] + d * Om2[i];
  }
  mat3_vec(V, u, t);
}
void se3_oplus(const double xi[6], const double pose[7], double out[7]) {
  // VertexSE3Expmap::oplusImpl: estimate <- exp(update) * estimate;   SE3Quat::operator*:
  // t = t1 + r1 * t2 ; r = r1 * r2 ; normalizeRotation
  double R[9], t[3], qd[4], q[4];
  se3_exp_Rt(xi, R, t);
  rot_to_quat(R, qd);
  quat_normalize(qd);
  double Rd[9];
  quat_to_rot(qd, Rd);
  double rt[3];
  mat3_vec(Rd, pose + 4, rt);
  quat_mul(qd, pose, q);
  quat_normalize(q);
  for (int i = 0; i < 4; ++i) out[i] = q[i];
  for (int i = 0; i < 3; ++i) out[4 + i] = t[i] + rt[i];
}

// ----------------------------------------------------------------------------------------------
// frame::sigma2_inv in float (init_config.hpp:63-78, config.cpp:53-55)
// ----------------------------------------------------------------------------------------------
struct Sigma2Inv {
  float v[8];
  Sigma2Inv() {
    float sf = 1.0f;
    v[0] = 1.0f;
    const float scale_factor = 1.2f;
    for (int i = 1; i < 8; ++i) {
      sf = sf * scale_factor;
      const float s2 = sf * sf;
      v[i] = 1.0f / s2;
    }
  }
};
const Sigma2Inv g_sigma2inv;

// ----------------------------------------------------------------------------------------------
// edges
// ----------------------------------------------------------------------------------------------
struct Cam { double fx, fy, cx, cy, bf; };

// reprojection residual + Jacobians at camera-frame point pc = R x + t.
//  stereo: EdgeStereoSE3ProjectXYZ / EdgeProjectXYZOnlyStereo: the element-wise point Jacobian of factors.cpp:133-168
//  mono:   EdgeSE3ProjectXYZ / EdgeProjectXYZOnly: `_jacobianOplusXi = -1. / z * tmp * R` (factors.cpp:84-107), restated
//          in the order Eigen evaluates a fixed-size lazy product (the scaled 2x3 `-1./z * tmp` is materialised first,
//          every coefficient is then the left-to-right sum over k of T(i,k) * R(k,j)).  That evaluation order is
//          recalled from Eigen 3.3, not verifiable here (no Eigen in the box): parity unpinned, like everything g2o.
// Jl[m][3] point Jacobian, Jp[m][6] pose Jacobian ([omega; upsilon], SURVEY Appx B.2), e[m].
inline void reproj_edge(const Cam& cam, const double* R, const double pc[3], const double obs[3], bool stereo,
                        double e[3], double Jl[9], double Jp[18]) {
  const double x = pc[0], y = pc[1], z = pc[2];
  const double z_2 = z * z;
  if (stereo) {
    const double invz = 1.0 / z;  // factors.cpp:117 (float literal, double division)
    const double u = x * invz * cam.fx + cam.cx;
    const double v = y * invz * cam.fy + cam.cy;
    const double ur = u - cam.bf * invz;
    e[0] = obs[0] - u; e[1] = obs[1] - v; e[2] = obs[2] - ur;
  } else {
    const double u = x / z * cam.fx + cam.cx;  // project2d then scale (factors.cpp:66-74)
    const double v = y / z * cam.fy + cam.cy;
    e[0] = obs[0] - u; e[1] = obs[1] - v; e[2] = 0.0;
  }
  if (Jl && stereo) {
    for (int j = 0; j < 3; ++j) {
      Jl[0 * 3 + j] = -cam.fx * M(R, 0, j) / z + cam.fx * x * M(R, 2, j) / z_2;
      Jl[1 * 3 + j] = -cam.fy * M(R, 1, j) / z + cam.fy * y * M(R, 2, j) / z_2;
      Jl[2 * 3 + j] = Jl[0 * 3 + j] - cam.bf * M(R, 2, j) / z_2;
    }
  } else if (Jl) {
    const double s = -1. / z;
    const double T00 = s * cam.fx, T01 = s * 0.0, T02 = s * (-x / z * cam.fx);
    const double T10 = s * 0.0, T11 = s * cam.fy, T12 = s * (-y / z * cam.fy);
    for (int j = 0; j < 3; ++j) {
      Jl[0 * 3 + j] = (T00 * M(R, 0, j) + T01 * M(R, 1, j)) + T02 * M(R, 2, j);
      Jl[1 * 3 + j] = (T10 * M(R, 0, j) + T11 * M(R, 1, j)) + T12 * M(R, 2, j);
      Jl[2 * 3 + j] = 0.0;
    }
  }
  if (Jp) {
    Jp[0] = x * y / z_2 * cam.fx;
    Jp[1] = -(1 + (x * x / z_2)) * cam.fx;
    Jp[2] = y / z * cam.fx;
    Jp[3] = -1. / z * cam.fx;
    Jp[4] = 0;
    Jp[5] = x / z_2 * cam.fx;
    Jp[6] = (1 + y * y / z_2) * cam.fy;
    Jp[7] = -x * y / z_2 * cam.fy;
    Jp[8] = -x / z * cam.fy;
    Jp[9] = 0;
    Jp[10] = -1. / z * cam.fy;
    Jp[11] = y / z_2 * cam.fy;
    if (stereo) {
      Jp[12] = Jp[0] - cam.bf * y / z_2;
      Jp[13] = Jp[1] + cam.bf * x / z_2;
      Jp[14] = Jp[2];
      Jp[15] = Jp[3];
      Jp[16] = 0;
      Jp[17] = Jp[5] - cam.bf / z_2;
    } else {
      for (int j = 12; j < 18; ++j) Jp[j] = 0;
    }
  }
}

// g2o RobustKernelHuber::robustify -> rho1 (SURVEY Appx B.4); delta stored as float in the callers
inline double huber_rho1(double e2, double delta) {
  const double dsqr = delta * delta;
  if (e2 <= dsqr) return 1.0;
  return delta / std::sqrt(e2);
}

// 6x6 SPD solve, plain Cholesky (stand-in for g2o LinearSolverEigen / LinearSolverDense on the
// reduced pose system, SURVEY Appx B.8).
bool solve6(const double H[36], const double b[6], double x[6]) {
  double L[36] = {0};
  for (int k = 0; k < 6; ++k) {
    double s = H[k * 6 + k];
    for (int j = 0; j < k; ++j) s -= L[k * 6 + j] * L[k * 6 + j];
    if (!(s > 0.0) || !std::isfinite(s)) return false;
    const double d = std::sqrt(s);
    L[k * 6 + k] = d;
    for (int i = k + 1; i < 6; ++i) {
      double t = H[i * 6 + k];
      for (int j = 0; j < k; ++j) t -= L[i * 6 + j] * L[k * 6 + j];
      L[i * 6 + k] = t / d;
    }
  }
  double y[6];
  for (int i = 0; i < 6; ++i) {
    double t = b[i];
    for (int j = 0; j < i; ++j) t -= L[i * 6 + j] * y[j];
    y[i] = t / L[i * 6 + i];
  }
  for (int i = 5; i >= 0; --i) {
    double t = y[i];
    for (int j = i + 1; j < 6; ++j) t -= L[j * 6 + i] * x[j];
    x[i] = t / L[i * 6 + i];
  }
  for (int i = 0; i < 6; ++i)
    if (!std::isfinite(x[i])) return false;
  return true;
}

template <class F>
void parallel_for(int64_t n, int nthreads, F f) {
  if (nthreads <= 1 || n <= 1) {
    f(0, n, 0);
    return;
  }
  nthreads = (int)std::min<int64_t>(nthreads, n);
  std::vector<std::thread> th;
  const int64_t chunk = (n + nthreads - 1) / nthreads;
  for (int t = 0; t < nthreads; ++t) {
    const int64_t lo = t * chunk, hi = std::min<int64_t>(n, lo + chunk);
    if (lo >= hi) break;
    th.emplace_back([=] { f(lo, hi, t); });
  }
  for (auto& x : th) x.join();
}

}  // namespace

extern "C" {

void oracle_default_params(oracle_params_t* p) {
  p->tri_lambda2 = 400.0;
  p->ba_lambda2 = 400.0;
  p->tri_str_thresh = (double)0.0064f;  // loc::tri_str_thresh is a float (config.h / config.cpp:88)
  p->tri_check_str_chi2 = 1;
  p->mdist2_gate = 9.0;
  p->chi2_mono = 5.991;
  p->chi2_stereo = 7.815;
  p->gn_iters = 10;
  p->robust = 0;
  p->factor_mode = 0;
}

double oracle_sigma2_inv(int octave) {
  if (octave < 0) octave = 0;
  if (octave > 7) octave = 7;
  return (double)g_sigma2inv.v[octave];
}

void oracle_map_build(const double* mean, const double* cov, int K, double deg_eps, oracle_comp_t* out) {
  for (int k = 0; k < K; ++k) {
    oracle_comp_t& c = out[k];
    for (int i = 0; i < 3; ++i) c.mean[i] = mean[k * 3 + i];
    for (int i = 0; i < 9; ++i) c.cov[i] = cov[k * 9 + i];
    inverse3(c.cov, c.cov_inv);   // gaussian.h:31
    c.det = det3(c.cov);          // gaussian.h:33
    eig3_jacobi(c.cov, c.scale, c.axis);       // gaussian.cpp:39-42
    c.is_degenerated = c.scale[0] < deg_eps;   // gaussian.cpp:44
    llt3(c.cov_inv, c.sqrt_info);              // gaussian.cpp:49
    c.is_salient = (c.scale[1] > 0.2 && c.scale[2] > 0.2);  // gaussian.cpp:50-55
  }
}

double oracle_chi2(const oracle_comp_t* c, const double x[3]) {
  // gaussian.cpp:65-70: delta^T * cov_inv * delta, evaluated (delta^T A) then (.) delta, left to right
  const double d0 = x[0] - c->mean[0], d1 = x[1] - c->mean[1], d2 = x[2] - c->mean[2];
  const double* A = c->cov_inv;
  const double r0 = (d0 * A[0] + d1 * A[3]) + d2 * A[6];
  const double r1 = (d0 * A[1] + d1 * A[4]) + d2 * A[7];
  const double r2 = (d0 * A[2] + d1 * A[5]) + d2 * A[8];
  return (r0 * d0 + r1 * d1) + r2 * d2;
}

void oracle_associate(const oracle_comp_t* comps, int K, const float* pts, int64_t Mpts, double gate,
                      int32_t* assign, double* d2_out, int nthreads) {
  // SoA copy so the k loop vectorises; per-k operation order is exactly oracle_chi2's.
  std::vector<double> mu(3 * (size_t)K), A(9 * (size_t)K);
  for (int k = 0; k < K; ++k) {
    for (int i = 0; i < 3; ++i) mu[(size_t)i * K + k] = comps[k].mean[i];
    for (int i = 0; i < 9; ++i) A[(size_t)i * K + k] = comps[k].cov_inv[i];
  }
  const double* m0 = mu.data(); const double* m1 = m0 + K; const double* m2 = m1 + K;
  const double* a[9];
  for (int i = 0; i < 9; ++i) a[i] = A.data() + (size_t)i * K;
  parallel_for(Mpts, nthreads, [&](int64_t lo, int64_t hi, int) {
    std::vector<double> buf(K);
    double* s = buf.data();
    for (int64_t p = lo; p < hi; ++p) {
      const double x0 = (double)pts[p * 3 + 0], x1 = (double)pts[p * 3 + 1], x2 = (double)pts[p * 3 + 2];
      for (int k = 0; k < K; ++k) {
        const double d0 = x0 - m0[k], d1 = x1 - m1[k], d2 = x2 - m2[k];
        const double r0 = (d0 * a[0][k] + d1 * a[3][k]) + d2 * a[6][k];
        const double r1 = (d0 * a[1][k] + d1 * a[4][k]) + d2 * a[7][k];
        const double r2 = (d0 * a[2][k] + d1 * a[5][k]) + d2 * a[8][k];
        s[k] = (r0 * d0 + r1 * d1) + r2 * d2;
      }
      double best = std::numeric_limits<double>::max();  // gmmloc_opt.cpp:176
      int bi = -1;
      for (int k = 0; k < K; ++k)
        if (s[k] < best) { best = s[k]; bi = k; }  // strict '<', ascending index
      if (gate > 0 && bi >= 0 && best > gate) bi = -1;  // gmmloc_opt.cpp:230
      assign[p] = bi;
      if (d2_out) d2_out[p] = best;
    }
  });
}

double oracle_bh3(const oracle_comp_t* g0, const oracle_comp_t* g1) {
  // gmm_utils.h:30-52
  double cov[9], inv[9];
  for (int i = 0; i < 9; ++i) cov[i] = (g0->cov[i] + g1->cov[i]) / 2.0;
  const double d[3] = {g1->mean[0] - g0->mean[0], g1->mean[1] - g0->mean[1], g1->mean[2] - g0->mean[2]};
  inverse3(cov, inv);
  const double r0 = (d[0] * inv[0] + d[1] * inv[3]) + d[2] * inv[6];
  const double r1 = (d[0] * inv[1] + d[1] * inv[4]) + d[2] * inv[7];
  const double r2 = (d[0] * inv[2] + d[1] * inv[5]) + d[2] * inv[8];
  double d0 = (r0 * d[0] + r1 * d[1]) + r2 * d[2];
  d0 /= 8.0;
  const double d1 = std::log(det3(cov) / std::sqrt(g0->det * g1->det)) / 2.0;
  return d0 + d1;
}

int64_t oracle_neighbors(const oracle_comp_t* comps, int K, double thresh, int64_t* row_ptr, int32_t* col,
                         double* dist, int64_t cap) {
  int64_t n = 0;
  for (int i = 0; i < K; ++i) {
    row_ptr[i] = n;
    for (int j = 0; j < K; ++j) {
      if (i == j) continue;
      const double d = oracle_bh3(&comps[i], &comps[j]);
      if (d < thresh) {
        if (col && n < cap) { col[n] = j; if (dist) dist[n] = d; }
        ++n;
      }
    }
  }
  row_ptr[K] = n;
  return n;
}

void oracle_se3_exp(const double xi[6], double pose[7]) {
  const double id[7] = {0, 0, 0, 1, 0, 0, 0};
  se3_oplus(xi, id, pose);
}
void oracle_se3_oplus(const double xi[6], const double pose[7], double pose_out[7]) { se3_oplus(xi, pose, pose_out); }

int oracle_optimize_point(const double x0[3], const double obs[3], int octave, const double pose[7],
                          const oracle_comp_t* comp, double proj_z2, const oracle_camera_t* camp,
                          const oracle_params_t* prm, int iters, double pt_est[3], double* chi2_proj,
                          double* chi2_str) {
  // gmmloc_opt.cpp:260-342
  const Cam cam{camp->fx, camp->fy, camp->cx, camp->cy, camp->bf};
  double q[4] = {pose[0], pose[1], pose[2], pose[3]};
  quat_normalize(q);
  double R[9];
  quat_to_rot(q, R);
  const double* t = pose + 4;
  const double w_proj = oracle_sigma2_inv(octave);                 // :266,293
  const double w_str = prm->tri_lambda2 * proj_z2;                 // :311-312
  const double n[3] = {comp->axis[0], comp->axis[3], comp->axis[6]};  // axis_.col(0)  :314
  double x[3] = {x0[0], x0[1], x0[2]};
  double c2p = 0, c2s = 0;
  bool ok = true;
  for (int it = 0; it < iters; ++it) {
    double pc[3];
    mat3_vec(R, x, pc);
    for (int i = 0; i < 3; ++i) pc[i] += t[i];
    double e[3], Jl[9];
    reproj_edge(cam, R, pc, obs, true, e, Jl, nullptr);
    const double es = (n[0] * (x[0] - comp->mean[0]) + n[1] * (x[1] - comp->mean[1])) + n[2] * (x[2] - comp->mean[2]);
    c2p = w_proj * ((e[0] * e[0] + e[1] * e[1]) + e[2] * e[2]);
    c2s = w_str * (es * es);
    double H[9], b[3];
    for (int r = 0; r < 3; ++r) {
      for (int c = 0; c < 3; ++c)
        M(H, r, c) = w_proj * ((Jl[0 + r] * Jl[0 + c] + Jl[3 + r] * Jl[3 + c]) + Jl[6 + r] * Jl[6 + c]) + w_str * (n[r] * n[c]);
      b[r] = -(w_proj * ((Jl[0 + r] * e[0] + Jl[3 + r] * e[1]) + Jl[6 + r] * e[2]) + w_str * (n[r] * es));
    }
    double Hi[9], dx[3];
    if (!inverse3(H, Hi)) { ok = false; break; }
    mat3_vec(Hi, b, dx);
    for (int i = 0; i < 3; ++i) x[i] += dx[i];
  }
  for (int i = 0; i < 3; ++i) pt_est[i] = x[i];
  *chi2_proj = c2p;  // errors of the last linearisation point (SURVEY Appx B.6)
  *chi2_str = c2s;
  int res = ok ? 1 : 0;
  if (c2p > 7.815) res = 0;                                                            // :332-335
  if (prm->tri_check_str_chi2 && c2s > prm->tri_str_thresh * prm->tri_lambda2) res = 0;  // :336-339
  return res;
}

int oracle_optimize_triangulation(const double x0[3], const double obs1[3], const double obs2[3], int octave1,
                                  const double pose1[7], const double pose2[7], const oracle_comp_t* comp,
                                  const oracle_camera_t* camp, const oracle_params_t* prm, int iters, double pt_est[3],
                                  double* err1, double* err2, double* chi2_str) {
  // localization_opt.cpp:27-204, one candidate component
  const Cam cam{camp->fx, camp->fy, camp->cx, camp->cy, camp->bf};
  const double* poses[2] = {pose1, pose2};
  const double* obs[2] = {obs1, obs2};
  double R[2][9];
  bool stereo[2];
  for (int v = 0; v < 2; ++v) {
    double q[4] = {poses[v][0], poses[v][1], poses[v][2], poses[v][3]};
    quat_normalize(q);
    quat_to_rot(q, R[v]);
    stereo[v] = obs[v][2] >= 0.0;  // kp.depth > 0  <=>  a right-image match exists
  }
  const double w = oracle_sigma2_inv(octave1);  // sigma2_inv1 for BOTH views (:126-135)
  const double w_str = prm->tri_lambda2;        // :106-107
  const double n[3] = {comp->axis[0], comp->axis[3], comp->axis[6]};
  double x[3] = {x0[0], x0[1], x0[2]};
  double c2v[2] = {0, 0}, c2s = 0;
  bool ok = true;
  for (int it = 0; it < iters; ++it) {
    double e[2][3], Jl[2][9];
    for (int v = 0; v < 2; ++v) {
      double pc[3];
      mat3_vec(R[v], x, pc);
      for (int i = 0; i < 3; ++i) pc[i] += poses[v][4 + i];
      reproj_edge(cam, R[v], pc, obs[v], stereo[v], e[v], Jl[v], nullptr);
      c2v[v] = w * ((e[v][0] * e[v][0] + e[v][1] * e[v][1]) + e[v][2] * e[v][2]);
    }
    const double es = (n[0] * (x[0] - comp->mean[0]) + n[1] * (x[1] - comp->mean[1])) + n[2] * (x[2] - comp->mean[2]);
    c2s = w_str * (es * es);
    double H[9], b[3];
    for (int r = 0; r < 3; ++r) {
      for (int c = 0; c < 3; ++c) {
        const double h1 = (Jl[0][r] * Jl[0][c] + Jl[0][3 + r] * Jl[0][3 + c]) + Jl[0][6 + r] * Jl[0][6 + c];
        const double h2 = (Jl[1][r] * Jl[1][c] + Jl[1][3 + r] * Jl[1][3 + c]) + Jl[1][6 + r] * Jl[1][6 + c];
        M(H, r, c) = (w * h1 + w * h2) + w_str * (n[r] * n[c]);
      }
      const double g1 = (Jl[0][r] * e[0][0] + Jl[0][3 + r] * e[0][1]) + Jl[0][6 + r] * e[0][2];
      const double g2 = (Jl[1][r] * e[1][0] + Jl[1][3 + r] * e[1][1]) + Jl[1][6 + r] * e[1][2];
      b[r] = -((w * g1 + w * g2) + w_str * (n[r] * es));
    }
    double Hi[9], dx[3];
    if (!inverse3(H, Hi)) { ok = false; break; }
    mat3_vec(Hi, b, dx);
    for (int i = 0; i < 3; ++i) x[i] += dx[i];
  }
  for (int i = 0; i < 3; ++i) pt_est[i] = x[i];
  *err1 = c2v[0];
  *err2 = c2v[1];
  *chi2_str = c2s;
  int res = ok ? 1 : 0;
  if (prm->tri_check_str_chi2 && c2s > prm->tri_str_thresh * prm->tri_lambda2) res = 0;  // :176-179
  const double th1 = stereo[0] ? 7.8 : 5.991, th2 = stereo[1] ? 7.8 : 5.991;             // :118,127,133
  if (c2v[0] > th1 || c2v[1] > th2) res = 0;                                             // :184-186
  return res;
}

void oracle_frame_solve(const oracle_comp_t* comps, int K, const oracle_camera_t* camp,
                        const oracle_params_t* prm, const double pose0[7], const float* pts_w,
                        const float* obsf, const uint8_t* octave, int N, const int32_t* assign_in,
                        double pose_out[7], double* pts_out, int32_t* assign_o, uint8_t* outlier,
                        int32_t* status, double* trace, double* pt_last) {
  if (pt_last) for (int i = 0; i < N; ++i) pt_last[i] = 0.0;
  const Cam cam{camp->fx, camp->fy, camp->cx, camp->cy, camp->bf};
  std::vector<int32_t> asg(N);
  if (assign_in) {
    for (int i = 0; i < N; ++i) asg[i] = assign_in[i];
  } else {
    oracle_associate(comps, K, pts_w, N, 0.0, asg.data(), nullptr, 1);
  }
  std::vector<double> x(3 * (size_t)N);
  for (int i = 0; i < 3 * N; ++i) x[i] = (double)pts_w[i];
  double pose[7];
  for (int i = 0; i < 7; ++i) pose[i] = pose0[i];
  quat_normalize(pose);

  // structure-edge gates, fixed at iteration 0 (SURVEY §7 / Appx A11):
  //   degenerate: e^2 <= tri_str_thresh     (gmmloc_opt.cpp:336-338, localization_opt.cpp:781)
  //               and the component passes renderView's view-cos cull at the initial pose
  //   otherwise : d^2 <= mdist2_gate (9.0)  (gmmloc_opt.cpp:230)
  std::vector<uint8_t> str_on(N, 0), free_pt(N, 0), stereo(N, 0);
  const bool pose_free = prm->factor_mode != 2;
  const bool pts_free = prm->factor_mode != 1;
  // camera centre at the initial pose: t_w_c = -(rot_c_w.inverse() * t_c_w)   gaussian_mixture.cpp:289
  double R0[9], twc[3];
  quat_to_rot(pose, R0);
  for (int r = 0; r < 3; ++r) twc[r] = -((M(R0, 0, r) * pose[4] + M(R0, 1, r) * pose[5]) + M(R0, 2, r) * pose[6]);
  const double view_cos_thresh = std::cos(78.0 * M_PI / 180.0);  // gaussian_mixture.cpp:284
  for (int i = 0; i < N; ++i) {
    stereo[i] = obsf[i * 3 + 2] >= 0.0f;  // u_right < 0 => monocular (tracking_opt.cpp:67)
    const int k = asg[i];
    if (k >= 0 && pts_free) {
      const oracle_comp_t& c = comps[k];
      const double* xi = &x[3 * i];
      if (c.is_degenerated) {
        const double e = (c.axis[0] * (xi[0] - c.mean[0]) + c.axis[3] * (xi[1] - c.mean[1])) + c.axis[6] * (xi[2] - c.mean[2]);
        str_on[i] = (e * e <= prm->tri_str_thresh);
        // view-cos cull of planar components seen closer than 12 deg to edge-on (GMM::renderView,
        // gaussian_mixture.cpp:283-302): po = (mu - t_w_c).normalized(), |po . axis0| < cos 78 deg => culled
        const double po[3] = {c.mean[0] - twc[0], c.mean[1] - twc[1], c.mean[2] - twc[2]};
        const double pn = std::sqrt((po[0] * po[0] + po[1] * po[1]) + po[2] * po[2]);
        const double view_cos = std::fabs((po[0] * c.axis[0] + po[1] * c.axis[3]) + po[2] * c.axis[6]) / pn;
        if (view_cos < view_cos_thresh) str_on[i] = 0;
      } else {
        str_on[i] = (oracle_chi2(&c, xi) <= prm->mdist2_gate);
      }
    }
    // a monocular point without a structure edge has a rank-2 Hll: hold it fixed (== setFixed(true))
    free_pt[i] = pts_free && (stereo[i] || str_on[i]);
  }
  const double delta_mono = (double)(float)std::sqrt(5.991), delta_stereo = (double)(float)std::sqrt(7.815);

  int st = 0;
  for (int it = 0; it < prm->gn_iters && st == 0; ++it) {
    double q[4] = {pose[0], pose[1], pose[2], pose[3]}, R[9];
    quat_to_rot(q, R);
    const double* t = pose + 4;
    double Hpp[36] = {0}, bp[6] = {0};
    // pass 1: edges -> Hpp, bp, and per free point the Schur terms
    double S[36] = {0}, g[6] = {0};
    std::vector<double> Dinv(9 * (size_t)N), Hpl(18 * (size_t)N), bl(3 * (size_t)N);
    for (int i = 0; i < N; ++i) {
      const double* xi = &x[3 * i];
      double pc[3];
      mat3_vec(R, xi, pc);
      for (int r = 0; r < 3; ++r) pc[r] += t[r];
      const double obs[3] = {(double)obsf[i * 3], (double)obsf[i * 3 + 1], (double)obsf[i * 3 + 2]};
      double e[3], Jl[9], Jp[18];
      const int m = stereo[i] ? 3 : 2;
      reproj_edge(cam, R, pc, obs, stereo[i], e, Jl, Jp);
      double w = oracle_sigma2_inv(octave[i]);
      double e2 = 0;
      for (int r = 0; r < m; ++r) e2 += e[r] * e[r];
      e2 *= w;
      if (prm->robust) w *= huber_rho1(e2, stereo[i] ? delta_stereo : delta_mono);
      if (pose_free) {
        for (int a = 0; a < 6; ++a) {
          for (int b = 0; b < 6; ++b) {
            double s = 0;
            for (int r = 0; r < m; ++r) s += Jp[r * 6 + a] * Jp[r * 6 + b];
            Hpp[a * 6 + b] += w * s;
          }
          double s = 0;
          for (int r = 0; r < m; ++r) s += Jp[r * 6 + a] * e[r];
          bp[a] -= w * s;
        }
      }
      if (!free_pt[i]) continue;
      double Hll[9], bli[3], *Hpli = &Hpl[18 * (size_t)i];
      for (int a = 0; a < 3; ++a) {
        for (int b = 0; b < 3; ++b) {
          double s = 0;
          for (int r = 0; r < m; ++r) s += Jl[r * 3 + a] * Jl[r * 3 + b];
          M(Hll, a, b) = w * s;
        }
        double s = 0;
        for (int r = 0; r < m; ++r) s += Jl[r * 3 + a] * e[r];
        bli[a] = -(w * s);
      }
      for (int a = 0; a < 6; ++a)
        for (int b = 0; b < 3; ++b) {
          double s = 0;
          for (int r = 0; r < m; ++r) s += Jp[r * 6 + a] * Jl[r * 3 + b];
          Hpli[a * 3 + b] = w * s;
        }
      if (str_on[i]) {
        const oracle_comp_t& c = comps[asg[i]];
        const double d[3] = {xi[0] - c.mean[0], xi[1] - c.mean[1], xi[2] - c.mean[2]};
        if (c.is_degenerated) {  // EdgePt2GaussianDeg, Omega = ba_lambda2  (localization_opt.cpp:657-669)
          const double n[3] = {c.axis[0], c.axis[3], c.axis[6]};
          const double es = (n[0] * d[0] + n[1] * d[1]) + n[2] * d[2];
          for (int a = 0; a < 3; ++a) {
            for (int b = 0; b < 3; ++b) M(Hll, a, b) += prm->ba_lambda2 * (n[a] * n[b]);
            bli[a] -= prm->ba_lambda2 * (n[a] * es);
          }
        } else {  // EdgePt2Gaussian: e = L^T d, J = L^T, Omega = I (factors.cpp:5-17, localization_opt.cpp:670-681)
          const double* L = c.sqrt_info;
          double es[3];
          for (int r = 0; r < 3; ++r) es[r] = (M(L, 0, r) * d[0] + M(L, 1, r) * d[1]) + M(L, 2, r) * d[2];
          for (int a = 0; a < 3; ++a) {
            for (int b = 0; b < 3; ++b) {
              double s = 0;
              for (int r = 0; r < 3; ++r) s += M(L, a, r) * M(L, b, r);
              M(Hll, a, b) += s;
            }
            double s = 0;
            for (int r = 0; r < 3; ++r) s += M(L, a, r) * es[r];
            bli[a] -= s;
          }
        }
      }
      double* Di = &Dinv[9 * (size_t)i];
      if (!inverse3(Hll, Di)) { st |= 1; break; }
      for (int r = 0; r < 3; ++r) bl[3 * (size_t)i + r] = bli[r];
      if (pose_free) {
        // W = Hpl * Dinv (6x3);  S += W Hpl^T ; g += W bl       (g2o BlockSolver Schur, Appx B.8)
        double W[18];
        for (int a = 0; a < 6; ++a)
          for (int b = 0; b < 3; ++b)
            W[a * 3 + b] = (Hpli[a * 3 + 0] * M(Di, 0, b) + Hpli[a * 3 + 1] * M(Di, 1, b)) + Hpli[a * 3 + 2] * M(Di, 2, b);
        for (int a = 0; a < 6; ++a) {
          for (int b = 0; b < 6; ++b)
            S[a * 6 + b] += (W[a * 3 + 0] * Hpli[b * 3 + 0] + W[a * 3 + 1] * Hpli[b * 3 + 1]) + W[a * 3 + 2] * Hpli[b * 3 + 2];
          g[a] += (W[a * 3 + 0] * bli[0] + W[a * 3 + 1] * bli[1]) + W[a * 3 + 2] * bli[2];
        }
      }
    }
    if (st) break;
    double dxi[6] = {0, 0, 0, 0, 0, 0};
    if (pose_free) {
      double Hs[36], bs[6];
      for (int i = 0; i < 36; ++i) Hs[i] = Hpp[i] - S[i];
      for (int i = 0; i < 6; ++i) bs[i] = bp[i] - g[i];
      if (!solve6(Hs, bs, dxi)) { st |= 1; break; }
    }
    if (trace) for (int i = 0; i < 6; ++i) trace[it * 8 + i] = dxi[i];
    // back-substitution and update
    bool finite = true;
    std::vector<double> xn(x);
    double max_dx = 0.0, max_dx_idx = -1.0;
    for (int i = 0; i < N; ++i) {
      if (!free_pt[i]) continue;
      const double* Hpli = &Hpl[18 * (size_t)i];
      double c[3];
      for (int b = 0; b < 3; ++b) {
        double s = 0;
        for (int a = 0; a < 6; ++a) s += Hpli[a * 3 + b] * dxi[a];
        c[b] = bl[3 * (size_t)i + b] - s;
      }
      double dx[3];
      mat3_vec(&Dinv[9 * (size_t)i], c, dx);
      if (pt_last) pt_last[i] = 0.0;
      for (int r = 0; r < 3; ++r) {
        xn[3 * i + r] = x[3 * i + r] + dx[r];
        if (!std::isfinite(xn[3 * i + r])) finite = false;
        if (std::fabs(dx[r]) > max_dx) { max_dx = std::fabs(dx[r]); max_dx_idx = (double)i; }
        if (pt_last && !(std::fabs(dx[r]) <= pt_last[i])) pt_last[i] = std::fabs(dx[r]);  // NaN-propagating max
      }
    }
    if (trace) { trace[it * 8 + 6] = max_dx; trace[it * 8 + 7] = max_dx_idx; }
    if (!finite) { st |= 1; break; }
    x.swap(xn);
    if (pose_free) {
      double np[7];
      se3_oplus(dxi, pose, np);
      for (int i = 0; i < 7; ++i) pose[i] = np[i];
    }
  }

  if (st) {
    // numeric-failure contract of the build (g2o ignores solver failures): the frame is
    // returned UNSOLVED — normalised input pose, input landmarks — with the status bit set.
    for (int i = 0; i < 7; ++i) pose[i] = pose0[i];
    quat_normalize(pose);
    for (int i = 0; i < 3 * N; ++i) x[i] = (double)pts_w[i];
  }
  if (pose_out) for (int i = 0; i < 7; ++i) pose_out[i] = pose[i];
  if (pts_out) for (int i = 0; i < 3 * N; ++i) pts_out[i] = x[i];
  if (assign_o) for (int i = 0; i < N; ++i) assign_o[i] = asg[i];
  if (status) *status = st;
  if (outlier) {
    // post-check at the final estimate: chi2 > 5.991 / 7.815 or non-positive depth
    // (localization_opt.cpp:857-881)
    double q[4] = {pose[0], pose[1], pose[2], pose[3]}, R[9];
    quat_to_rot(q, R);
    for (int i = 0; i < N; ++i) {
      double pc[3];
      mat3_vec(R, &x[3 * i], pc);
      for (int r = 0; r < 3; ++r) pc[r] += pose[4 + r];
      const double obs[3] = {(double)obsf[i * 3], (double)obsf[i * 3 + 1], (double)obsf[i * 3 + 2]};
      double e[3];
      reproj_edge(cam, R, pc, obs, stereo[i], e, nullptr, nullptr);
      const int m = stereo[i] ? 3 : 2;
      double e2 = 0;
      for (int r = 0; r < m; ++r) e2 += e[r] * e[r];
      e2 *= oracle_sigma2_inv(octave[i]);
      const double th = stereo[i] ? prm->chi2_stereo : prm->chi2_mono;
      outlier[i] = (e2 > th || !(pc[2] > 0.0)) ? 1 : 0;
    }
  }
}

int oracle_optimize_pose_lm(const oracle_camera_t* camp, const double pose0[7], const float* pts_w, const float* obsf,
                            const uint8_t* octave, const uint8_t* valid, int N, double pose_out[7], uint8_t* outlier) {
  const Cam cam{camp->fx, camp->fy, camp->cx, camp->cy, camp->bf};
  std::vector<int> idx;  // features with a map point (tracking_opt.cpp:63-64)
  for (int i = 0; i < N; ++i) {
    if (outlier) outlier[i] = 0;
    if (!valid || valid[i]) idx.push_back(i);
  }
  const int E = (int)idx.size();
  for (int i = 0; i < 7; ++i) pose_out[i] = pose0[i];
  if (E < 3) return 0;  // :139-140
  const double delta_mono = (double)(float)std::sqrt(5.991), delta_stereo = (double)(float)std::sqrt(7.815);  // :57-58
  const float chi2_mono = 5.991f, chi2_stereo = 7.815f;                                                      // :145-146
  std::vector<uint8_t> level(E, 0), is_out(E, 0), stereo(E);
  std::vector<double> err(3 * (size_t)E, 0.0), w(E);
  for (int e = 0; e < E; ++e) {
    stereo[e] = obsf[idx[e] * 3 + 2] >= 0.0f;  // u_right < 0 => monocular (:67)
    w[e] = oracle_sigma2_inv(octave[idx[e]]);
  }
  bool robust = true;
  double pose[7];
  // errors of every ACTIVE edge at `T` (computeActiveErrors) and their robustified sum (activeRobustChi2)
  auto active_errors = [&](const double* T) {
    double R[9];
    quat_to_rot(T, R);
    double chi = 0.0;
    for (int e = 0; e < E; ++e) {
      if (level[e]) continue;
      const int i = idx[e];
      const double X[3] = {(double)pts_w[3 * i], (double)pts_w[3 * i + 1], (double)pts_w[3 * i + 2]};
      double pc[3];
      mat3_vec(R, X, pc);
      for (int r = 0; r < 3; ++r) pc[r] += T[4 + r];
      const double ob[3] = {(double)obsf[3 * i], (double)obsf[3 * i + 1], (double)obsf[3 * i + 2]};
      reproj_edge(cam, R, pc, ob, stereo[e], &err[3 * (size_t)e], nullptr, nullptr);
      const double* er = &err[3 * (size_t)e];
      const double c2 = w[e] * ((er[0] * er[0] + er[1] * er[1]) + er[2] * er[2]);
      if (robust) {
        const double d = stereo[e] ? delta_stereo : delta_mono, dsqr = d * d;
        chi += (c2 <= dsqr) ? c2 : 2 * std::sqrt(c2) * d - dsqr;  // RobustKernelHuber rho[0]
      } else {
        chi += c2;
      }
    }
    return chi;
  };
  int n_bad = 0;
  for (int round = 0; round < 4; ++round) {
    for (int i = 0; i < 7; ++i) pose[i] = pose0[i];  // vertex_se3->setEstimate(curr_frame_->getTcw())  :152
    quat_normalize(pose);
    double lambda = 0.0, ni = 2.0;
    bool ok = true;
    for (int it = 0; it < 10 && ok; ++it) {  // optimizer.optimize(its[it])  :154
      double cur_chi = active_errors(pose);
      // buildSystem at the current estimate
      double R[9], H[36] = {0}, b[6] = {0};
      quat_to_rot(pose, R);
      for (int e = 0; e < E; ++e) {
        if (level[e]) continue;
        const int i = idx[e];
        const double X[3] = {(double)pts_w[3 * i], (double)pts_w[3 * i + 1], (double)pts_w[3 * i + 2]};
        double pc[3], er[3], Jp[18];
        mat3_vec(R, X, pc);
        for (int r = 0; r < 3; ++r) pc[r] += pose[4 + r];
        const double ob[3] = {(double)obsf[3 * i], (double)obsf[3 * i + 1], (double)obsf[3 * i + 2]};
        reproj_edge(cam, R, pc, ob, stereo[e], er, nullptr, Jp);
        const int m = stereo[e] ? 3 : 2;
        double rho1 = 1.0;
        if (robust) {
          double c2 = 0;
          for (int r = 0; r < m; ++r) c2 += er[r] * er[r];
          c2 *= w[e];
          rho1 = huber_rho1(c2, stereo[e] ? delta_stereo : delta_mono);
        }
        for (int a = 0; a < 6; ++a) {
          for (int c = 0; c < 6; ++c) {
            double s = 0;
            for (int r = 0; r < m; ++r) s += Jp[r * 6 + a] * Jp[r * 6 + c];
            H[a * 6 + c] += rho1 * w[e] * s;
          }
          double s = 0;
          for (int r = 0; r < m; ++r) s += Jp[r * 6 + a] * er[r];
          b[a] -= rho1 * w[e] * s;
        }
      }
      if (it == 0) {  // computeLambdaInit: tau * max diagonal, tau = 1e-5
        double md = 0;
        for (int a = 0; a < 6; ++a) md = std::max(md, std::fabs(H[a * 6 + a]));
        lambda = 1e-5 * md;
        ni = 2.0;
      }
      double rho = 0;
      int qmax = 0;
      do {
        double Hl[36], dx[6], trial[7];
        for (int a = 0; a < 36; ++a) Hl[a] = H[a];
        for (int a = 0; a < 6; ++a) Hl[a * 6 + a] += lambda;
        const bool ok2 = solve6(Hl, b, dx);
        double tmp_chi;
        if (ok2) {
          se3_oplus(dx, pose, trial);
          tmp_chi = active_errors(trial);
        } else {
          for (int a = 0; a < 6; ++a) dx[a] = 0;
          tmp_chi = std::numeric_limits<double>::max();
        }
        rho = cur_chi - tmp_chi;
        double scale = 0;
        for (int a = 0; a < 6; ++a) scale += dx[a] * (lambda * dx[a] + b[a]);
        scale += 1e-3;
        rho /= scale;
        if (rho > 0 && std::isfinite(tmp_chi) && ok2) {  // good step
          double alpha = 1. - std::pow(2 * rho - 1, 3);
          alpha = std::min(alpha, 2. / 3.);
          lambda *= std::max(1. / 3., alpha);
          ni = 2;
          cur_chi = tmp_chi;
          for (int a = 0; a < 7; ++a) pose[a] = trial[a];
        } else {
          lambda *= ni;
          ni *= 2;
          if (!std::isfinite(lambda)) break;
        }
        ++qmax;
      } while (rho < 0 && qmax < 10);
      if (qmax == 10 || rho == 0 || !std::isfinite(lambda)) ok = false;  // OptimizationAlgorithm::Terminate
    }
    // classification (:157-203)
    double R[9];
    quat_to_rot(pose, R);
    n_bad = 0;
    for (int e = 0; e < E; ++e) {
      const int i = idx[e];
      if (is_out[e]) {  // e->computeError() at the current estimate
        const double X[3] = {(double)pts_w[3 * i], (double)pts_w[3 * i + 1], (double)pts_w[3 * i + 2]};
        double pc[3];
        mat3_vec(R, X, pc);
        for (int r = 0; r < 3; ++r) pc[r] += pose[4 + r];
        const double ob[3] = {(double)obsf[3 * i], (double)obsf[3 * i + 1], (double)obsf[3 * i + 2]};
        reproj_edge(cam, R, pc, ob, stereo[e], &err[3 * (size_t)e], nullptr, nullptr);
      }
      const double* er = &err[3 * (size_t)e];
      const float chi2 = (float)(w[e] * ((er[0] * er[0] + er[1] * er[1]) + er[2] * er[2]));
      if (chi2 > (stereo[e] ? chi2_stereo : chi2_mono)) {
        is_out[e] = 1;
        level[e] = 1;
        ++n_bad;
      } else {
        is_out[e] = 0;
        level[e] = 0;
      }
    }
    if (round == 2) robust = false;  // e->setRobustKernel(0)
    if (E < 10) break;               // optimizer.edges().size() < 10
  }
  for (int i = 0; i < 7; ++i) pose_out[i] = pose[i];
  if (outlier)
    for (int e = 0; e < E; ++e) outlier[idx[e]] = is_out[e];
  return E - n_bad;
}

void oracle_frames_solve(const oracle_comp_t* comps, int K, const oracle_camera_t* cam,
                         const oracle_params_t* prm, const double* pose0, const float* pts_w,
                         const float* obs, const uint8_t* octave, int B, int N, double* pose_out,
                         double* pts_out, int32_t* assign, uint8_t* outlier, int32_t* status, int nthreads) {
  oracle_frames_solve_diag(comps, K, cam, prm, pose0, pts_w, obs, octave, B, N, pose_out, pts_out, assign, outlier, status, nthreads,
                           nullptr, nullptr);
}

void oracle_frames_solve_diag(const oracle_comp_t* comps, int K, const oracle_camera_t* cam,
                              const oracle_params_t* prm, const double* pose0, const float* pts_w,
                              const float* obs, const uint8_t* octave, int B, int N, double* pose_out,
                              double* pts_out, int32_t* assign, uint8_t* outlier, int32_t* status, int nthreads,
                              double* conv, double* pt_last) {
  parallel_for(B, nthreads, [&](int64_t lo, int64_t hi, int) {
    std::vector<double> trace(conv ? 8 * (size_t)std::max(prm->gn_iters, 1) : 0);
    for (int64_t f = lo; f < hi; ++f) {
      std::fill(trace.begin(), trace.end(), 0.0);
      oracle_frame_solve(comps, K, cam, prm, pose0 + 7 * f, pts_w + 3 * (size_t)N * f, obs + 3 * (size_t)N * f,
                         octave + (size_t)N * f, N, nullptr, pose_out ? pose_out + 7 * f : nullptr,
                         pts_out ? pts_out + 3 * (size_t)N * f : nullptr, assign ? assign + (size_t)N * f : nullptr,
                         outlier ? outlier + (size_t)N * f : nullptr, status ? status + f : nullptr,
                         conv ? trace.data() : nullptr, pt_last ? pt_last + (size_t)N * f : nullptr);
      if (conv) {
        // convergence of the reference iteration itself: |dxi|_inf of the last iteration, the largest step that was
        // LARGER than its predecessor (0 when the pose steps shrink monotonically), the largest last point step
        double last = 0.0, grew = 0.0, prev = 0.0;
        for (int it = 0; it < prm->gn_iters; ++it) {
          double s = 0.0;
          for (int i = 0; i < 6; ++i) s = std::max(s, std::fabs(trace[it * 8 + i]));
          if (it >= 2 && s > prev) grew = std::max(grew, s);
          prev = s;
          last = s;
        }
        conv[3 * f] = last;
        conv[3 * f + 1] = grew;
        conv[3 * f + 2] = prm->gn_iters > 0 ? trace[(prm->gn_iters - 1) * 8 + 6] : 0.0;
      }
    }
  });
}

// ----------------------------------------------------------------------------------------------
// .gmm wire format: varint32 count, then per component varint32 size + ComponentProto bytes
// (protobuf_utils.cpp:12-101; GMM.proto:5-14: 1 bool is_degenerated, 2 bool is_salient,
//  3 packed double mean, 4 packed double covariance)
// ----------------------------------------------------------------------------------------------
static bool rd_varint(const std::vector<uint8_t>& b, size_t& pos, uint64_t& v) {
  v = 0;
  for (int shift = 0; shift < 64; shift += 7) {
    if (pos >= b.size()) return false;
    const uint8_t c = b[pos++];
    v |= (uint64_t)(c & 0x7f) << shift;
    if (!(c & 0x80)) return true;
  }
  return false;
}
static void wr_varint(std::vector<uint8_t>& b, uint64_t v) {
  while (v >= 0x80) { b.push_back((uint8_t)(v | 0x80)); v >>= 7; }
  b.push_back((uint8_t)v);
}

int oracle_gmm_read(const char* path, double* mean, double* cov, uint8_t* flags, int cap) {
  FILE* f = std::fopen(path, "rb");
  if (!f) return -1;
  std::vector<uint8_t> b;
  uint8_t tmp[65536];
  size_t n;
  while ((n = std::fread(tmp, 1, sizeof tmp, f)) > 0) b.insert(b.end(), tmp, tmp + n);
  std::fclose(f);
  size_t pos = 0;
  uint64_t cnt;
  if (!rd_varint(b, pos, cnt)) return -2;
  if (cnt == 0) return -3;  // "protobuf file empty!" gmm_utils.cpp:32-35
  const int K = (int)cnt;
  const bool fill = mean && cov && cap >= K;
  for (int k = 0; k < K; ++k) {
    uint64_t sz;
    if (!rd_varint(b, pos, sz) || sz == 0 || pos + sz > b.size()) return -4;
    const size_t end = pos + sz;
    int nm = 0, nc = 0;
    uint8_t fl[2] = {0, 0};
    double m[3] = {0, 0, 0}, c[9] = {0};
    while (pos < end) {
      uint64_t key;
      if (!rd_varint(b, pos, key)) return -5;
      const int field = (int)(key >> 3), wt = (int)(key & 7);
      if (wt == 0) {
        uint64_t v;
        if (!rd_varint(b, pos, v)) return -5;
        if (field == 1) fl[0] = v != 0;
        if (field == 2) fl[1] = v != 0;
      } else if (wt == 2) {
        uint64_t len;
        if (!rd_varint(b, pos, len) || pos + len > end) return -5;
        if ((field == 3 || field == 4) && len % 8 == 0) {
          for (uint64_t i = 0; i < len / 8; ++i) {
            double v;
            std::memcpy(&v, &b[pos + 8 * i], 8);
            if (field == 3) { if (nm < 3) m[nm] = v; ++nm; }
            else { if (nc < 9) c[nc] = v; ++nc; }
          }
        }
        pos += len;
      } else if (wt == 1) {  // non-packed double (proto2 parsers accept both encodings)
        if (pos + 8 > end) return -5;
        double v;
        std::memcpy(&v, &b[pos], 8);
        if (field == 3) { if (nm < 3) m[nm] = v; ++nm; }
        if (field == 4) { if (nc < 9) c[nc] = v; ++nc; }
        pos += 8;
      } else if (wt == 5) {
        pos += 4;
      } else {
        return -5;
      }
    }
    if (pos != end) return -5;
    if (nm != 3 || nc != 9) return -6;  // CHECK_EQ gmm_utils.cpp:50-51
    if (fill) {
      for (int i = 0; i < 3; ++i) mean[k * 3 + i] = m[i];
      for (int i = 0; i < 9; ++i) cov[k * 9 + i] = c[i];  // row-major fill, gmm_utils.cpp:55-59
      if (flags) { flags[2 * k] = fl[0]; flags[2 * k + 1] = fl[1]; }
    }
  }
  return K;
}

int oracle_gmm_write(const char* path, const double* mean, const double* cov, const uint8_t* flags, int K) {
  std::vector<uint8_t> b;
  wr_varint(b, (uint64_t)K);  // gmm_utils.cpp:87-92
  for (int k = 0; k < K; ++k) {
    std::vector<uint8_t> m;
    m.push_back(0x08); m.push_back(flags ? flags[2 * k] : 0);      // field 1 varint  (:97)
    m.push_back(0x10); m.push_back(flags ? flags[2 * k + 1] : 0);  // field 2 varint  (:99)
    m.push_back(0x1a); m.push_back(24);                            // field 3 packed 3 doubles (:103-105)
    for (int i = 0; i < 3; ++i) { uint8_t t[8]; std::memcpy(t, &mean[k * 3 + i], 8); m.insert(m.end(), t, t + 8); }
    m.push_back(0x22); m.push_back(72);                            // field 4 packed 9 doubles (:106-108)
    // writer emits cov(i) in Eigen's column-major linear order (:107)
    for (int i = 0; i < 9; ++i) {
      const int r = i % 3, c = i / 3;
      uint8_t t[8];
      std::memcpy(t, &cov[k * 9 + r * 3 + c], 8);
      m.insert(m.end(), t, t + 8);
    }
    wr_varint(b, m.size());
    b.insert(b.end(), m.begin(), m.end());
  }
  FILE* f = std::fopen(path, "wb");
  if (!f) return -1;
  const size_t w = std::fwrite(b.data(), 1, b.size(), f);
  std::fclose(f);
  return w == b.size() ? 0 : -2;
}

// ----------------------------------------------------------------------------------------------
// keyframe association pipeline: projectGaussian, renderView, searchCorrespondence, checkMapAssociation
// ----------------------------------------------------------------------------------------------
}  // extern "C"

namespace {

// Eigen QuaternionBase::_transformVector: uv = q.vec x v; uv += uv; v + w * uv + q.vec x uv
inline void quat_rotate(const double q[4], const double v[3], double out[3]) {
  double uv[3] = {q[1] * v[2] - q[2] * v[1], q[2] * v[0] - q[0] * v[2], q[0] * v[1] - q[1] * v[0]};
  for (int i = 0; i < 3; ++i) uv[i] += uv[i];
  const double c[3] = {q[1] * uv[2] - q[2] * uv[1], q[2] * uv[0] - q[0] * uv[2], q[0] * uv[1] - q[1] * uv[0]};
  for (int i = 0; i < 3; ++i) out[i] = (v[i] + q[3] * uv[i]) + c[i];
}

inline double det2(const double* m) { return m[0] * m[3] - m[2] * m[1]; }  // Eigen: m(0,0) m(1,1) - m(1,0) m(0,1)
inline void inverse2(const double* m, double* out) {
  // Eigen compute_inverse_size2_helper
  const double invdet = 1.0 / det2(m);
  out[0] = m[3] * invdet;
  out[2] = -m[2] * invdet;
  out[1] = -m[1] * invdet;
  out[3] = m[0] * invdet;
}
inline double quad2(const double* A, const double* d) {
  // delta^T * A * delta: (delta^T A) first, then . delta
  const double r0 = d[0] * A[0] + d[1] * A[2];
  const double r1 = d[0] * A[1] + d[1] * A[3];
  return r0 * d[0] + r1 * d[1];
}

}  // namespace

extern "C" {

int oracle_project_gaussian(const oracle_comp_t* c, int id, const oracle_camera_t* cam, const double pose[7],
                            oracle_comp2d_t* out) {
  // gmm_utils.cpp:121-146
  double q[4] = {pose[0], pose[1], pose[2], pose[3]};
  quat_normalize(q);
  const double* t = pose + 4;
  double pc[3];
  quat_rotate(q, c->mean, pc);  // rot_c_w * mu3d + t_c_w
  for (int i = 0; i < 3; ++i) pc[i] += t[i];
  // PinholeCamera::project3 with Jacobian (pinhole_camera.cpp:68-125)
  const double x = pc[0], y = pc[1], z = pc[2];
  const double rz = 1.0 / z;
  const double kx = x * rz, ky = y * rz;
  const double rz2 = rz * rz;
  const double J[6] = {cam->fx * rz, 0.0, -cam->fx * x * rz2, 0.0, cam->fy * rz, -cam->fy * y * rz2};
  const double u = cam->fx * kx + cam->cx;
  const double v = cam->fy * ky + cam->cy;
  // evaluateProjectionResult / isKeypointVisible (pinhole_camera.cpp:127-150), kMinimumDepth = 0
  const bool vis = u >= 0.0 && v >= 0.0 && u < (double)cam->width && v < (double)cam->height;
  if (!(vis && z > 0.0)) return 0;
  double R[9];
  quat_to_rot(q, R);
  // jacob_proj * rot * cov3d * rot^T * jacob_proj^T, left to right
  double T1[6], T2[6], T3[6];
  for (int i = 0; i < 2; ++i)
    for (int j = 0; j < 3; ++j) T1[i * 3 + j] = (J[i * 3] * M(R, 0, j) + J[i * 3 + 1] * M(R, 1, j)) + J[i * 3 + 2] * M(R, 2, j);
  for (int i = 0; i < 2; ++i)
    for (int j = 0; j < 3; ++j)
      T2[i * 3 + j] = (T1[i * 3] * M(c->cov, 0, j) + T1[i * 3 + 1] * M(c->cov, 1, j)) + T1[i * 3 + 2] * M(c->cov, 2, j);
  for (int i = 0; i < 2; ++i)
    for (int j = 0; j < 3; ++j) T3[i * 3 + j] = (T2[i * 3] * M(R, j, 0) + T2[i * 3 + 1] * M(R, j, 1)) + T2[i * 3 + 2] * M(R, j, 2);
  for (int i = 0; i < 2; ++i)
    for (int j = 0; j < 2; ++j) out->cov[i * 2 + j] = (T3[i * 3] * J[j * 3] + T3[i * 3 + 1] * J[j * 3 + 1]) + T3[i * 3 + 2] * J[j * 3 + 2];
  out->id = id;
  out->pad_ = 0;
  out->mean[0] = u;
  out->mean[1] = v;
  inverse2(out->cov, out->cov_inv);  // gaussian.h:109
  out->det = det2(out->cov);         // gaussian.h:111
  // SelfAdjointEigenSolver<Matrix2d> (lower triangle), closed form; ascending (gaussian.cpp:20-23)
  const double a = out->cov[0], b = out->cov[2], d = out->cov[3];
  const double tr = 0.5 * (a + d), h = 0.5 * (a - d);
  const double r = std::sqrt(h * h + b * b);
  out->scale[0] = tr - r;
  out->scale[1] = tr + r;
  out->depth = pc[2];
  return 1;
}

double oracle_bh2(const oracle_comp2d_t* g0, const oracle_comp2d_t* g1) {
  // gmm_utils.h:30-52 with 2-D types
  double cov[4], inv[4];
  for (int i = 0; i < 4; ++i) cov[i] = (g0->cov[i] + g1->cov[i]) / 2.0;
  const double d[2] = {g1->mean[0] - g0->mean[0], g1->mean[1] - g0->mean[1]};
  inverse2(cov, inv);
  double d0 = quad2(inv, d);
  d0 /= 8.0;
  const double d1 = std::log(det2(cov) / std::sqrt(g0->det * g1->det)) / 2.0;
  return d0 + d1;
}

double oracle_mdist2_2d(const oracle_comp2d_t* c, const double uv[2]) {
  const double d[2] = {uv[0] - c->mean[0], uv[1] - c->mean[1]};  // gaussian.h:123-126
  return quad2(c->cov_inv, d);
}

int oracle_render_view(const oracle_comp_t* comps, int K, const oracle_camera_t* cam, const double pose[7],
                       oracle_comp2d_t* out) {
  // gaussian_mixture.cpp:271-371
  double q[4] = {pose[0], pose[1], pose[2], pose[3]};
  quat_normalize(q);
  // t_w_c = -(rot_c_w.inverse() * t_c_w); Eigen inverse() = conjugate / squaredNorm
  const double n2 = ((q[0] * q[0] + q[1] * q[1]) + q[2] * q[2]) + q[3] * q[3];
  const double qi[4] = {-q[0] / n2, -q[1] / n2, -q[2] / n2, q[3] / n2};
  double twc[3];
  quat_rotate(qi, pose + 4, twc);
  for (int i = 0; i < 3; ++i) twc[i] = -twc[i];
  const double view_cos_thresh = std::cos(78.0 * M_PI / 180.0);  // :283
  int V = 0;
  for (int k = 0; k < K; ++k) {
    const oracle_comp_t& c = comps[k];
    if (c.is_degenerated) {  // STEP.0 check view cos (:285-302)
      double po[3] = {c.mean[0] - twc[0], c.mean[1] - twc[1], c.mean[2] - twc[2]};
      const double z = (po[0] * po[0] + po[1] * po[1]) + po[2] * po[2];
      if (z > 0.0) {  // Eigen normalized()
        const double nrm = std::sqrt(z);
        for (int i = 0; i < 3; ++i) po[i] /= nrm;
      }
      const double view_cos = std::fabs((po[0] * c.axis[0] + po[1] * c.axis[3]) + po[2] * c.axis[6]);
      if (view_cos < view_cos_thresh) continue;
    }
    oracle_comp2d_t g;
    if (!oracle_project_gaussian(&c, k, cam, pose, &g)) continue;
    if (g.scale[0] < 4.0 && g.scale[1] < 4.0) continue;  // STEP. check cov (:311-317)
    if (V > 0) {  // STEP. check near far (:327-355)
      int min_idx = 0;
      double min_dist = std::numeric_limits<double>::max();
      for (int j = 0; j < V; ++j) {
        const double dist = oracle_bh2(&out[j], &g);
        if (dist < min_dist) { min_dist = dist; min_idx = j; }
      }
      if (min_dist < 0.8) {
        if (g.depth < out[min_idx].depth) out[min_idx] = g;
      } else {
        out[V++] = g;
      }
    } else {
      out[V++] = g;
    }
  }
  std::stable_sort(out, out + V, [](const oracle_comp2d_t& a, const oracle_comp2d_t& b) { return a.depth > b.depth; });  // :362-364
  return V;
}

void oracle_search_correspondence(const oracle_comp2d_t* view, int V, const double* uv, int N, int num, int32_t* cand) {
  // gaussian_mixture.cpp:484-534
  std::vector<double> bd(num);
  std::vector<int> bi(num);
  for (int i = 0; i < N; ++i) {
    int n = 0;
    for (int j = 0; j < V; ++j) {
      const double dx = uv[i * 2] - view[j].mean[0], dy = uv[i * 2 + 1] - view[j].mean[1];
      const double d = dx * dx + dy * dy;  // nanoflann L2_Simple_Adaptor
      if (n == num && !(d < bd[num - 1])) continue;
      int pos = (n < num) ? n : num - 1;
      while (pos > 0 && d < bd[pos - 1]) {
        bd[pos] = bd[pos - 1];
        bi[pos] = bi[pos - 1];
        --pos;
      }
      bd[pos] = d;
      bi[pos] = j;
      if (n < num) ++n;
    }
    int m = 0;
    for (int r = 0; r < n; ++r)
      if (oracle_mdist2_2d(&view[bi[r]], &uv[i * 2]) < 9.0) cand[i * num + m++] = bi[r];  // :518-522
    for (; m < num; ++m) cand[i * num + m] = -1;
  }
}

int oracle_query_point(const oracle_comp_t* comps, int K, const double pt[3], int num, int32_t* knn, double* min_mdist2) {
  // GMM::queryPoint (gaussian_mixture.cpp:545-576): knnSearch(pt, 5) over the component means (ascending squared
  // Euclid distance; nanoflann kd-tree in the reference, brute force here: equal distances -> lower index first),
  // the minimum Mahalanobis distance over the neighbours is computed (:563-572) and DISCARDED: the function returns
  // ret_index[0], the Euclid-nearest mean (:574) — quirk preserved.  Returns -1 for an empty map (res stays empty).
  std::vector<double> bd(num);
  std::vector<int> bi(num);
  int n = 0;
  for (int k = 0; k < K; ++k) {
    const double d0 = pt[0] - comps[k].mean[0], d1 = pt[1] - comps[k].mean[1], d2 = pt[2] - comps[k].mean[2];
    const double d = d0 * d0 + d1 * d1 + d2 * d2;  // FLANNPoints3d::kdtree_distance, gaussian_mixture.h:33-39
    if (n == num && !(d < bd[num - 1])) continue;
    int pos = (n < num) ? n : num - 1;
    while (pos > 0 && d < bd[pos - 1]) {
      bd[pos] = bd[pos - 1];
      bi[pos] = bi[pos - 1];
      --pos;
    }
    bd[pos] = d;
    bi[pos] = k;
    if (n < num) ++n;
  }
  double min_dist = std::numeric_limits<double>::max();
  for (int r = 0; r < n; ++r) {
    const double dist = oracle_chi2(&comps[bi[r]], pt);
    if (dist < min_dist) min_dist = dist;
  }
  if (knn)
    for (int r = 0; r < num; ++r) knn[r] = r < n ? bi[r] : -1;
  if (min_mdist2) *min_mdist2 = min_dist;
  return n ? bi[0] : -1;
}

void oracle_unproject3(const oracle_camera_t* cam, const double pose[7], const double uv[2], double depth, double pt_w[3]) {
  // pinhole_camera.cpp:20-33 then getTwc().map (SE3Quat::inverse: r = conj, t = r * (t * -1))
  const double pc[3] = {depth * (uv[0] - cam->cx) / cam->fx, depth * (uv[1] - cam->cy) / cam->fy, depth};
  double q[4] = {pose[0], pose[1], pose[2], pose[3]};
  quat_normalize(q);
  const double qc[4] = {-q[0], -q[1], -q[2], q[3]};
  const double tn[3] = {pose[4] * -1.0, pose[5] * -1.0, pose[6] * -1.0};
  double twc[3], r[3];
  quat_rotate(qc, tn, twc);
  quat_rotate(qc, pc, r);
  for (int i = 0; i < 3; ++i) pt_w[i] = r[i] + twc[i];
}

int oracle_check_map_association(const oracle_comp_t* comps, int K, const int64_t* nbr_row_ptr, const int32_t* nbr_col,
                                 const oracle_camera_t* cam, const oracle_params_t* prm, const double pose[7],
                                 const double obs[3], int octave, const int32_t* cand, int ncand, double pt3d[3]) {
  // gmmloc_opt.cpp:156-258
  int nc = 0;
  for (int i = 0; i < ncand; ++i)
    if (cand[i] >= 0) ++nc;
  if (nc == 0) return -1;  // comps.empty()  :162-164
  const double pt_init[3] = {pt3d[0], pt3d[1], pt3d[2]};
  double q[4] = {pose[0], pose[1], pose[2], pose[3]};
  quat_normalize(q);
  double ptc[3];
  quat_rotate(q, pt3d, ptc);  // Tcw.map(pt3d)
  double proj_z = ptc[2] + pose[6];
  proj_z = proj_z > 1.0 ? 1.0 : proj_z;  // :171
  const double proj_z2 = proj_z * proj_z;
  int min_idx = -1;
  double min_value = std::numeric_limits<double>::max();
  double min_res[3] = {0, 0, 0};
  for (int i = 0; i < ncand; ++i) {
    if (cand[i] < 0) continue;
    double est[3], c2p, c2s;
    const int res = oracle_optimize_point(pt_init, obs, octave, pose, &comps[cand[i]], proj_z2, cam, prm, 5, est, &c2p, &c2s);
    if (res && c2p < min_value) {  // :192-196
      min_idx = i;
      min_value = c2p;
      for (int r = 0; r < 3; ++r) min_res[r] = est[r];
    }
  }
  if (min_idx != -1) {
    const int g3d = cand[min_idx];
    double ll = oracle_chi2(&comps[g3d], min_res);
    int str = g3d;
    for (int64_t e = nbr_row_ptr[g3d]; e < nbr_row_ptr[g3d + 1]; ++e) {  // :209-217
      const double ln = oracle_chi2(&comps[nbr_col[e]], min_res);
      if (ln < ll) { ll = ln; str = nbr_col[e]; }
    }
    if (str != g3d) {  // :219-228
      double est[3], c2p, c2s;
      const int res = oracle_optimize_point(pt_init, obs, octave, pose, &comps[str], proj_z2, cam, prm, 5, est, &c2p, &c2s);
      if (res) {
        for (int r = 0; r < 3; ++r) min_res[r] = est[r];
      } else {
        str = g3d;
        ll = oracle_chi2(&comps[g3d], min_res);
      }
    }
    if (ll > 9.0) return -1;  // :230-232
    for (int r = 0; r < 3; ++r) pt3d[r] = min_res[r];
    return str;
  }
  // use_nn3d (:237-256): GMM::queryPoint -> nearest mean (Euclid), result discarded except for pt3d
  int nn = -1;
  double best = std::numeric_limits<double>::max();
  for (int k = 0; k < K; ++k) {
    const double d0 = pt_init[0] - comps[k].mean[0], d1 = pt_init[1] - comps[k].mean[1], d2 = pt_init[2] - comps[k].mean[2];
    const double d = (d0 * d0 + d1 * d1) + d2 * d2;
    if (d < best) { best = d; nn = k; }
  }
  if (nn < 0 || !comps[nn].is_degenerated) return -1;
  double est[3], c2p, c2s;
  if (oracle_optimize_point(pt_init, obs, octave, pose, &comps[nn], proj_z2, cam, prm, 5, est, &c2p, &c2s))
    for (int r = 0; r < 3; ++r) pt3d[r] = est[r];
  return -1;
}

void oracle_keyframe_associate(const oracle_comp_t* comps, int K, const int64_t* nbr_row_ptr, const int32_t* nbr_col,
                               const oracle_camera_t* cam, const oracle_params_t* prm, const double pose[7],
                               const float* obs, const float* depth, const uint8_t* octave, int N, int num,
                               int32_t* assoc, double* pt3d, int32_t* cand_out, int32_t* nview, int nthreads) {
  std::vector<oracle_comp2d_t> view((size_t)K);
  const int V = oracle_render_view(comps, K, cam, pose, view.data());  // gmmloc_opt.cpp:122-126
  if (nview) *nview = V;
  std::vector<double> uv(2 * (size_t)N);
  for (int i = 0; i < N; ++i) {
    uv[2 * i] = (double)obs[3 * i];
    uv[2 * i + 1] = (double)obs[3 * i + 1];
  }
  std::vector<int32_t> cand((size_t)N * num);
  oracle_search_correspondence(view.data(), V, uv.data(), N, num, cand.data());  // :128-134
  for (size_t i = 0; i < cand.size(); ++i)
    if (cand[i] >= 0) cand[i] = view[cand[i]].id;  // gc2d->parent_
  if (cand_out) std::memcpy(cand_out, cand.data(), sizeof(int32_t) * cand.size());
  parallel_for(N, nthreads, [&](int64_t lo, int64_t hi, int) {
    for (int64_t i = lo; i < hi; ++i) {
      assoc[i] = -1;
      if (!(depth[i] > 0.0f)) {  // gmmloc_opt.cpp:42,68
        pt3d[3 * i] = pt3d[3 * i + 1] = pt3d[3 * i + 2] = std::numeric_limits<double>::quiet_NaN();
        continue;
      }
      oracle_unproject3(cam, pose, &uv[2 * i], (double)depth[i], &pt3d[3 * i]);  // :73
      const double o[3] = {(double)obs[3 * i], (double)obs[3 * i + 1], (double)obs[3 * i + 2]};
      assoc[i] = oracle_check_map_association(comps, K, nbr_row_ptr, nbr_col, cam, prm, pose, o, octave[i],
                                              &cand[(size_t)i * num], num, &pt3d[3 * i]);  // :76-77
    }
  });
}

// ----------------------------------------------------------------------------------------------
// local bundle adjustment: Localization::jointOptimization (localization_opt.cpp:456-925)
// ----------------------------------------------------------------------------------------------
}  // extern "C"

namespace {

// g2o SE3Quat::inverse: r = conj(r), t = r * (t * -1)
void se3_inverse(const double T[7], double out[7]) {
  const double qc[4] = {-T[0], -T[1], -T[2], T[3]};
  const double tn[3] = {T[4] * -1.0, T[5] * -1.0, T[6] * -1.0};
  quat_rotate(qc, tn, out + 4);
  for (int i = 0; i < 4; ++i) out[i] = qc[i];
}
// g2o SE3Quat::operator*: t = t1 + r1 * t2, r = r1 * r2, normalizeRotation
void se3_mul(const double A[7], const double B[7], double out[7]) {
  double rt[3], q[4];
  quat_rotate(A, B + 4, rt);
  quat_mul(A, B, q);
  quat_normalize(q);
  for (int i = 0; i < 4; ++i) out[i] = q[i];
  for (int i = 0; i < 3; ++i) out[4 + i] = A[4 + i] + rt[i];
}
// g2o SE3Quat::log (se3quat.h): [omega; upsilon]
void se3_log(const double T[7], double xi[6]) {
  double R[9];
  quat_to_rot(T, R);
  const double d = 0.5 * (M(R, 0, 0) + M(R, 1, 1) + M(R, 2, 2) - 1);
  const double dR[3] = {M(R, 2, 1) - M(R, 1, 2), M(R, 0, 2) - M(R, 2, 0), M(R, 1, 0) - M(R, 0, 1)};
  double om[3], Vi[9];
  double coef;
  if (std::fabs(d) > 0.99999) {
    for (int i = 0; i < 3; ++i) om[i] = 0.5 * dR[i];
    coef = 1. / 12.;
  } else {
    const double theta = std::acos(d);
    for (int i = 0; i < 3; ++i) om[i] = theta / (2 * std::sqrt(1 - d * d)) * dR[i];
    coef = (1 - theta / (2 * std::tan(theta / 2))) / (theta * theta);
  }
  const double Om[9] = {0, -om[2], om[1], om[2], 0, -om[0], -om[1], om[0], 0};
  for (int r = 0; r < 3; ++r)
    for (int c = 0; c < 3; ++c) {
      const double o2 = (M(Om, r, 0) * M(Om, 0, c) + M(Om, r, 1) * M(Om, 1, c)) + M(Om, r, 2) * M(Om, 2, c);
      M(Vi, r, c) = ((r == c) ? 1.0 : 0.0) - 0.5 * M(Om, r, c) + coef * o2;
    }
  mat3_vec(Vi, T + 4, xi + 3);
  for (int i = 0; i < 3; ++i) xi[i] = om[i];
}
// g2o SE3Quat::adj: [[R, 0], [skew(t) R, R]]  (row-major 6x6)
void se3_adj(const double T[7], double A[36]) {
  double R[9];
  quat_to_rot(T, R);
  const double* t = T + 4;
  const double St[9] = {0, -t[2], t[1], t[2], 0, -t[0], -t[1], t[0], 0};
  for (int i = 0; i < 36; ++i) A[i] = 0;
  for (int r = 0; r < 3; ++r)
    for (int c = 0; c < 3; ++c) {
      A[r * 6 + c] = M(R, r, c);
      A[(3 + r) * 6 + 3 + c] = M(R, r, c);
      A[(3 + r) * 6 + c] = (M(St, r, 0) * M(R, 0, c) + M(St, r, 1) * M(R, 1, c)) + M(St, r, 2) * M(R, 2, c);
    }
}

// dense SPD solve (in place Cholesky), n x n row-major
bool solve_dense_spd(std::vector<double>& A, std::vector<double>& b, int n) {
  for (int k = 0; k < n; ++k) {
    double s = A[(size_t)k * n + k];
    for (int j = 0; j < k; ++j) s -= A[(size_t)k * n + j] * A[(size_t)k * n + j];
    if (!(s > 0.0) || !std::isfinite(s)) return false;
    const double d = std::sqrt(s);
    A[(size_t)k * n + k] = d;
    for (int i = k + 1; i < n; ++i) {
      double t = A[(size_t)i * n + k];
      for (int j = 0; j < k; ++j) t -= A[(size_t)i * n + j] * A[(size_t)k * n + j];
      A[(size_t)i * n + k] = t / d;
    }
  }
  for (int i = 0; i < n; ++i) {
    double t = b[i];
    for (int j = 0; j < i; ++j) t -= A[(size_t)i * n + j] * b[j];
    b[i] = t / A[(size_t)i * n + i];
  }
  for (int i = n - 1; i >= 0; --i) {
    double t = b[i];
    for (int j = i + 1; j < n; ++j) t -= A[(size_t)j * n + i] * b[j];
    b[i] = t / A[(size_t)i * n + i];
  }
  for (int i = 0; i < n; ++i)
    if (!std::isfinite(b[i])) return false;
  return true;
}

struct LocalBA {
  const oracle_comp_t* comps;
  Cam cam;
  const oracle_params_t* prm;
  int P, L, E, nf;
  const uint8_t* fixed;
  const int32_t *assoc, *obs_pt, *obs_kf;
  const float* obs;
  std::vector<int> fidx;              // pose -> free index or -1
  std::vector<double> w;              // per observation information
  std::vector<uint8_t> stereo, level; // per observation
  std::vector<double> err;            // E x 3, errors of the last computeActiveErrors
  std::vector<uint8_t> str_type, str_level;  // per point: 0 none, 1 EdgePt2GaussianDeg, 2 EdgePt2Gaussian
  std::vector<double> str_err;        // L x 3
  std::vector<std::vector<int>> pt_obs;
  int prior_idx;
  double prior_inv[7], prior_info[6], prior_err[6];
  bool robust = true;
  double delta_mono, delta_stereo;
  // linear system of the current linearisation point
  std::vector<double> Hpp, bp, Hll, bl, Hpl;
  // last solution
  std::vector<double> dxp, dxl;

  void structure_error(const double* x, int l, double* e) const {
    const oracle_comp_t& c = comps[assoc[l]];
    const double d[3] = {x[0] - c.mean[0], x[1] - c.mean[1], x[2] - c.mean[2]};
    if (str_type[l] == 1) {  // factors.cpp:55-64
      e[0] = (c.axis[0] * d[0] + c.axis[3] * d[1]) + c.axis[6] * d[2];
      e[1] = e[2] = 0.0;
    } else {  // factors.cpp:5-17: sqrt_info^T * d
      const double* Lm = c.sqrt_info;
      for (int r = 0; r < 3; ++r) e[r] = (M(Lm, 0, r) * d[0] + M(Lm, 1, r) * d[1]) + M(Lm, 2, r) * d[2];
    }
  }
  double structure_chi2(int l, const double* e) const {
    if (str_type[l] == 1) return prm->ba_lambda2 * (e[0] * e[0]);
    return (e[0] * e[0] + e[1] * e[1]) + e[2] * e[2];
  }

  // computeActiveErrors + activeRobustChi2 at (pose, pts)
  double active_errors(const std::vector<double>& pose, const std::vector<double>& pts) {
    std::vector<double> R(9 * (size_t)P);
    for (int i = 0; i < P; ++i) quat_to_rot(&pose[7 * (size_t)i], &R[9 * (size_t)i]);
    double chi = 0.0;
    for (int e = 0; e < E; ++e) {
      if (level[e]) continue;
      const int i = obs_kf[e];
      double pc[3];
      mat3_vec(&R[9 * (size_t)i], &pts[3 * (size_t)obs_pt[e]], pc);
      for (int r = 0; r < 3; ++r) pc[r] += pose[7 * (size_t)i + 4 + r];
      const double ob[3] = {(double)obs[3 * e], (double)obs[3 * e + 1], (double)obs[3 * e + 2]};
      double* er = &err[3 * (size_t)e];
      reproj_edge(cam, &R[9 * (size_t)i], pc, ob, stereo[e], er, nullptr, nullptr);
      const double c2 = w[e] * ((er[0] * er[0] + er[1] * er[1]) + er[2] * er[2]);
      if (robust) {
        const double d = stereo[e] ? delta_stereo : delta_mono, dsqr = d * d;
        chi += (c2 <= dsqr) ? c2 : 2 * std::sqrt(c2) * d - dsqr;
      } else {
        chi += c2;
      }
    }
    for (int l = 0; l < L; ++l) {
      if (!str_type[l] || str_level[l]) continue;
      structure_error(&pts[3 * (size_t)l], l, &str_err[3 * (size_t)l]);
      chi += structure_chi2(l, &str_err[3 * (size_t)l]);
    }
    if (prior_idx >= 0) {  // factors.cpp:19-29
      double delta[7];
      se3_mul(prior_inv, &pose[7 * (size_t)prior_idx], delta);
      se3_log(delta, prior_err);
      for (int a = 0; a < 6; ++a) chi += prior_info[a] * (prior_err[a] * prior_err[a]);
    }
    return chi;
  }

  // buildSystem: linearizeOplus + constructQuadraticForm of every active edge
  void build(const std::vector<double>& pose, const std::vector<double>& pts) {
    Hpp.assign(36 * (size_t)nf, 0.0);
    bp.assign(6 * (size_t)nf, 0.0);
    Hll.assign(9 * (size_t)L, 0.0);
    bl.assign(3 * (size_t)L, 0.0);
    Hpl.assign(18 * (size_t)E, 0.0);
    std::vector<double> R(9 * (size_t)P);
    for (int i = 0; i < P; ++i) quat_to_rot(&pose[7 * (size_t)i], &R[9 * (size_t)i]);
    for (int e = 0; e < E; ++e) {
      if (level[e]) continue;
      const int i = obs_kf[e], l = obs_pt[e], f = fidx[i];
      const double* Ri = &R[9 * (size_t)i];
      double pc[3], er[3], Jl[9], Jp[18];
      mat3_vec(Ri, &pts[3 * (size_t)l], pc);
      for (int r = 0; r < 3; ++r) pc[r] += pose[7 * (size_t)i + 4 + r];
      const double ob[3] = {(double)obs[3 * e], (double)obs[3 * e + 1], (double)obs[3 * e + 2]};
      reproj_edge(cam, Ri, pc, ob, stereo[e], er, Jl, Jp);
      const int m = stereo[e] ? 3 : 2;
      double we = w[e];
      if (robust) {
        double c2 = 0;
        for (int r = 0; r < m; ++r) c2 += er[r] * er[r];
        c2 *= w[e];
        we *= huber_rho1(c2, stereo[e] ? delta_stereo : delta_mono);
      }
      for (int a = 0; a < 3; ++a) {
        for (int b = 0; b < 3; ++b) {
          double s = 0;
          for (int r = 0; r < m; ++r) s += Jl[r * 3 + a] * Jl[r * 3 + b];
          Hll[9 * (size_t)l + a * 3 + b] += we * s;
        }
        double s = 0;
        for (int r = 0; r < m; ++r) s += Jl[r * 3 + a] * er[r];
        bl[3 * (size_t)l + a] -= we * s;
      }
      if (f < 0) continue;
      for (int a = 0; a < 6; ++a) {
        for (int b = 0; b < 6; ++b) {
          double s = 0;
          for (int r = 0; r < m; ++r) s += Jp[r * 6 + a] * Jp[r * 6 + b];
          Hpp[36 * (size_t)f + a * 6 + b] += we * s;
        }
        double s = 0;
        for (int r = 0; r < m; ++r) s += Jp[r * 6 + a] * er[r];
        bp[6 * (size_t)f + a] -= we * s;
        for (int b = 0; b < 3; ++b) {
          double s2 = 0;
          for (int r = 0; r < m; ++r) s2 += Jp[r * 6 + a] * Jl[r * 3 + b];
          Hpl[18 * (size_t)e + a * 3 + b] = we * s2;
        }
      }
    }
    for (int l = 0; l < L; ++l) {
      if (!str_type[l] || str_level[l]) continue;
      const oracle_comp_t& c = comps[assoc[l]];
      double es[3];
      structure_error(&pts[3 * (size_t)l], l, es);
      double* H = &Hll[9 * (size_t)l];
      double* b = &bl[3 * (size_t)l];
      if (str_type[l] == 1) {  // Omega = ba_lambda2, J = n^T   (localization_opt.cpp:657-669)
        const double n[3] = {c.axis[0], c.axis[3], c.axis[6]};
        for (int a = 0; a < 3; ++a) {
          for (int bb = 0; bb < 3; ++bb) H[a * 3 + bb] += prm->ba_lambda2 * (n[a] * n[bb]);
          b[a] -= prm->ba_lambda2 * (n[a] * es[0]);
        }
      } else {  // Omega = I, J = sqrt_info^T   (:670-681)
        const double* Lm = c.sqrt_info;
        for (int a = 0; a < 3; ++a) {
          for (int bb = 0; bb < 3; ++bb) {
            double s = 0;
            for (int r = 0; r < 3; ++r) s += M(Lm, a, r) * M(Lm, bb, r);
            H[a * 3 + bb] += s;
          }
          double s = 0;
          for (int r = 0; r < 3; ++r) s += M(Lm, a, r) * es[r];
          b[a] -= s;
        }
      }
    }
    if (prior_idx >= 0 && fidx[prior_idx] >= 0) {  // factors.cpp:31-53
      const int f = fidx[prior_idx];
      double delta[7], dv[6], Jr[36], Tinv[7], Ad[36], J[36];
      se3_mul(prior_inv, &pose[7 * (size_t)prior_idx], delta);
      se3_log(delta, dv);
      for (int i = 0; i < 36; ++i) Jr[i] = 0;
      const double ph[9] = {0, -dv[2], dv[1], dv[2], 0, -dv[0], -dv[1], dv[0], 0};
      const double lu[9] = {0, -dv[5], dv[4], dv[5], 0, -dv[3], -dv[4], dv[3], 0};
      for (int r = 0; r < 3; ++r)
        for (int c = 0; c < 3; ++c) {
          Jr[r * 6 + c] = 0.5 * M(ph, r, c);
          Jr[(3 + r) * 6 + 3 + c] = 0.5 * M(ph, r, c);
          Jr[r * 6 + 3 + c] = 0.5 * M(lu, r, c);
        }
      for (int a = 0; a < 6; ++a) Jr[a * 6 + a] += 1.0;
      se3_inverse(&pose[7 * (size_t)prior_idx], Tinv);
      se3_adj(Tinv, Ad);
      for (int a = 0; a < 6; ++a)
        for (int b = 0; b < 6; ++b) {
          double s2 = 0;
          for (int k = 0; k < 6; ++k) s2 += Jr[a * 6 + k] * Ad[k * 6 + b];
          J[a * 6 + b] = s2;
        }
      for (int a = 0; a < 6; ++a) {
        for (int b = 0; b < 6; ++b) {
          double s2 = 0;
          for (int k = 0; k < 6; ++k) s2 += J[k * 6 + a] * prior_info[k] * J[k * 6 + b];
          Hpp[36 * (size_t)f + a * 6 + b] += s2;
        }
        double s2 = 0;
        for (int k = 0; k < 6; ++k) s2 += J[k * 6 + a] * prior_info[k] * dv[k];
        bp[6 * (size_t)f + a] -= s2;
      }
    }
  }

  double max_diag() const {
    double md = 0;
    for (int f = 0; f < nf; ++f)
      for (int a = 0; a < 6; ++a) md = std::max(md, std::fabs(Hpp[36 * (size_t)f + a * 7]));
    for (int l = 0; l < L; ++l)
      for (int a = 0; a < 3; ++a) md = std::max(md, std::fabs(Hll[9 * (size_t)l + a * 4]));
    return md;
  }

  // BlockSolver_6_3::solve with Schur complement at damping lambda
  bool solve(double lambda) {
    const int n = 6 * nf;
    std::vector<double> S((size_t)n * n, 0.0), rhs(bp);
    std::vector<double> Hinv(9 * (size_t)L);
    for (int f = 0; f < nf; ++f)
      for (int a = 0; a < 6; ++a)
        for (int b = 0; b < 6; ++b) S[(size_t)(6 * f + a) * n + 6 * f + b] = Hpp[36 * (size_t)f + a * 6 + b] + (a == b ? lambda : 0.0);
    bool ok = true;
    for (int l = 0; l < L; ++l) {
      double Hl[9];
      for (int i = 0; i < 9; ++i) Hl[i] = Hll[9 * (size_t)l + i];
      for (int a = 0; a < 3; ++a) Hl[a * 4] += lambda;
      if (!inverse3(Hl, &Hinv[9 * (size_t)l])) ok = false;
      const double* Hi = &Hinv[9 * (size_t)l];
      for (int ea : pt_obs[l]) {
        if (level[ea]) continue;
        const int fa = fidx[obs_kf[ea]];
        if (fa < 0) continue;
        const double* Wa = &Hpl[18 * (size_t)ea];
        double T[18];  // Hpl_a * Hll^-1
        for (int a = 0; a < 6; ++a)
          for (int b = 0; b < 3; ++b) T[a * 3 + b] = (Wa[a * 3] * M(Hi, 0, b) + Wa[a * 3 + 1] * M(Hi, 1, b)) + Wa[a * 3 + 2] * M(Hi, 2, b);
        for (int a = 0; a < 6; ++a)
          rhs[6 * (size_t)fa + a] -= (T[a * 3] * bl[3 * (size_t)l] + T[a * 3 + 1] * bl[3 * (size_t)l + 1]) + T[a * 3 + 2] * bl[3 * (size_t)l + 2];
        for (int eb : pt_obs[l]) {
          if (level[eb]) continue;
          const int fb = fidx[obs_kf[eb]];
          if (fb < 0) continue;
          const double* Wb = &Hpl[18 * (size_t)eb];
          for (int a = 0; a < 6; ++a)
            for (int b = 0; b < 6; ++b)
              S[(size_t)(6 * fa + a) * n + 6 * fb + b] -= (T[a * 3] * Wb[b * 3] + T[a * 3 + 1] * Wb[b * 3 + 1]) + T[a * 3 + 2] * Wb[b * 3 + 2];
        }
      }
    }
    if (!ok) return false;
    if (n > 0 && !solve_dense_spd(S, rhs, n)) return false;
    dxp = rhs;
    dxl.assign(3 * (size_t)L, 0.0);
    for (int l = 0; l < L; ++l) {
      double c[3] = {bl[3 * (size_t)l], bl[3 * (size_t)l + 1], bl[3 * (size_t)l + 2]};
      for (int ea : pt_obs[l]) {
        if (level[ea]) continue;
        const int fa = fidx[obs_kf[ea]];
        if (fa < 0) continue;
        const double* Wa = &Hpl[18 * (size_t)ea];
        for (int b = 0; b < 3; ++b) {
          double s2 = 0;
          for (int a = 0; a < 6; ++a) s2 += Wa[a * 3 + b] * dxp[6 * (size_t)fa + a];
          c[b] -= s2;
        }
      }
      mat3_vec(&Hinv[9 * (size_t)l], c, &dxl[3 * (size_t)l]);
      for (int r = 0; r < 3; ++r)
        if (!std::isfinite(dxl[3 * (size_t)l + r])) return false;
    }
    return true;
  }

  // g2o's force-stop flag (SparseOptimizer::setForceStopFlag(pbStopFlag), localization_opt.cpp:541-542): terminate()
  // reads *flag wherever g2o polls it.  stop > 0: raised; stop == 0 / no flag: never; stop == -n: the flag reads as
  // raised from the n-th poll on (deterministic replay of an asynchronous stop, the parity tests' hook).
  int32_t stop = 0;
  int polls = 0;
  bool terminate() {
    ++polls;
    return stop > 0 || (stop < 0 && polls >= -stop);
  }

  // OptimizationAlgorithmLevenberg, `iters` iterations (SURVEY Appx B.7); returns the iterations run.
  // terminate() is polled where g2o does: `for (i < iterations && !terminate() && ok)` (SparseOptimizer::optimize)
  // and `while (rho < 0 && qmax < maxTrials && !terminate())` (OptimizationAlgorithmLevenberg::solve).
  int optimize(int iters, std::vector<double>& pose, std::vector<double>& pts) {
    double lambda = 0.0, ni = 2.0;
    bool ok = true;
    int it = 0;
    for (; it < iters && !terminate() && ok; ++it) {
      double cur_chi = active_errors(pose, pts);
      build(pose, pts);
      if (it == 0) {
        lambda = 1e-5 * max_diag();
        ni = 2.0;
      }
      double rho = 0;
      int qmax = 0;
      do {
        std::vector<double> tp(pose), tx(pts);
        const bool ok2 = solve(lambda);
        double tmp_chi;
        if (ok2) {
          for (int i = 0; i < P; ++i)
            if (fidx[i] >= 0) se3_oplus(&dxp[6 * (size_t)fidx[i]], &pose[7 * (size_t)i], &tp[7 * (size_t)i]);
          for (size_t i = 0; i < tx.size(); ++i) tx[i] += dxl[i];
          tmp_chi = active_errors(tp, tx);
        } else {
          dxp.assign(6 * (size_t)nf, 0.0);
          dxl.assign(3 * (size_t)L, 0.0);
          tmp_chi = std::numeric_limits<double>::max();
        }
        rho = cur_chi - tmp_chi;
        double scale = 0;
        for (size_t a = 0; a < dxp.size(); ++a) scale += dxp[a] * (lambda * dxp[a] + bp[a]);
        for (size_t a = 0; a < dxl.size(); ++a) scale += dxl[a] * (lambda * dxl[a] + bl[a]);
        scale += 1e-3;
        rho /= scale;
        if (rho > 0 && std::isfinite(tmp_chi) && ok2) {
          double alpha = 1. - std::pow(2 * rho - 1, 3);
          alpha = std::min(alpha, 2. / 3.);
          lambda *= std::max(1. / 3., alpha);
          ni = 2;
          cur_chi = tmp_chi;
          pose.swap(tp);
          pts.swap(tx);
        } else {
          lambda *= ni;
          ni *= 2;
          if (!std::isfinite(lambda)) break;
        }
        ++qmax;
      } while (rho < 0 && qmax < 10 && !terminate());
      if (qmax == 10 || rho == 0 || !std::isfinite(lambda)) ok = false;
    }
    return it;
  }
};

}  // namespace

extern "C" {

void oracle_se3_log(const double pose[7], double xi[6]) {
  double q[7];
  for (int i = 0; i < 7; ++i) q[i] = pose[i];
  quat_normalize(q);
  se3_log(q, xi);
}

int oracle_local_ba(const oracle_comp_t* comps, int K, const oracle_camera_t* camp, const oracle_params_t* prm, int P,
                    double* pose_io, const uint8_t* fixed, int prior_idx, int L, double* pts_io, const int32_t* assoc, int E,
                    const int32_t* obs_pt, const int32_t* obs_kf, const float* obs, const uint8_t* octave,
                    uint8_t* obs_outlier, uint8_t* assoc_cleared, int32_t* iters) {
  return oracle_local_ba_stop(comps, K, camp, prm, P, pose_io, fixed, prior_idx, L, pts_io, assoc, E, obs_pt, obs_kf, obs, octave,
                              obs_outlier, assoc_cleared, iters, 0);
}

int oracle_local_ba_stop(const oracle_comp_t* comps, int K, const oracle_camera_t* camp, const oracle_params_t* prm, int P,
                         double* pose_io, const uint8_t* fixed, int prior_idx, int L, double* pts_io, const int32_t* assoc, int E,
                         const int32_t* obs_pt, const int32_t* obs_kf, const float* obs, const uint8_t* octave,
                         uint8_t* obs_outlier, uint8_t* assoc_cleared, int32_t* iters, int32_t stop) {
  for (int e = 0; e < E; ++e)
    if (obs_pt[e] < 0 || obs_pt[e] >= L || obs_kf[e] < 0 || obs_kf[e] >= P) return -1;
  for (int l = 0; l < L; ++l)
    if (assoc[l] >= K) return -1;
  if (prior_idx >= P) return -1;
  LocalBA ba;
  ba.stop = stop;
  // `if (pbStopFlag) if (*pbStopFlag) return;` before the first optimize (:765-767): nothing is touched
  if (ba.terminate()) {
    if (obs_outlier) for (int e = 0; e < E; ++e) obs_outlier[e] = 0;
    if (assoc_cleared) for (int l = 0; l < L; ++l) assoc_cleared[l] = 0;
    if (iters) iters[0] = iters[1] = iters[2] = 0;
    return 0;
  }
  ba.comps = comps;
  ba.cam = Cam{camp->fx, camp->fy, camp->cx, camp->cy, camp->bf};
  ba.prm = prm;
  ba.P = P; ba.L = L; ba.E = E;
  ba.fixed = fixed;
  ba.assoc = assoc; ba.obs_pt = obs_pt; ba.obs_kf = obs_kf; ba.obs = obs;
  ba.fidx.assign(P, -1);
  ba.nf = 0;
  for (int i = 0; i < P; ++i)
    if (!fixed[i]) ba.fidx[i] = ba.nf++;
  ba.w.resize(E); ba.stereo.resize(E); ba.level.assign(E, 0); ba.err.assign(3 * (size_t)E, 0.0);
  ba.pt_obs.assign(L, {});
  for (int e = 0; e < E; ++e) {
    ba.w[e] = oracle_sigma2_inv(octave[e]);
    ba.stereo[e] = !(obs[3 * e + 2] < 0.0f);  // u_right < 0 => monocular (:702)
    ba.pt_obs[obs_pt[e]].push_back(e);
  }
  ba.str_type.assign(L, 0); ba.str_level.assign(L, 0); ba.str_err.assign(3 * (size_t)L, 0.0);
  for (int l = 0; l < L; ++l)
    if (assoc[l] >= 0) ba.str_type[l] = comps[assoc[l]].is_degenerated ? 1 : 2;  // :655-682
  ba.delta_mono = (double)(float)std::sqrt(5.991);    // thHuberMono   :637
  ba.delta_stereo = (double)(float)std::sqrt(7.815);  // thHuberStereo :638
  std::vector<double> pose(pose_io, pose_io + 7 * (size_t)P), pts(pts_io, pts_io + 3 * (size_t)L);
  for (int i = 0; i < P; ++i) quat_normalize(&pose[7 * (size_t)i]);
  ba.prior_idx = prior_idx;
  if (prior_idx >= 0) {
    se3_inverse(&pose[7 * (size_t)prior_idx], ba.prior_inv);  // setMeasurement(kfi_ptr->getTcw())  :566
    const double sigma_rot = 2.0 * M_PI / 180.0;
    for (int a = 0; a < 3; ++a) {
      ba.prior_info[a] = 1.0 / (sigma_rot * sigma_rot);   // :568-575
      ba.prior_info[3 + a] = 1.0 / (0.01 * 0.01);
    }
  }
  int its[3] = {0, 0, 0};
  // stage 1: optimize(5)  (:770-771)
  ba.robust = true;
  its[0] = ba.optimize(5, pose, pts);
  // degenerate structure edges: computeError at the current estimate, reject above the threshold (:774-786)
  for (int l = 0; l < L; ++l) {
    if (ba.str_type[l] != 1) continue;
    double es[3];
    ba.structure_error(&pts[3 * (size_t)l], l, es);
    if (ba.structure_chi2(l, es) > prm->tri_str_thresh * prm->ba_lambda2) ba.str_level[l] = 1;
  }
  // stage 2: optimize(5)  (:788-789)
  its[1] = ba.optimize(5, pose, pts);
  // reprojection edges: chi2 of the last computeActiveErrors; isDepthPositive at the current estimate (:799-823)
  auto depth_positive = [&](int e) {
    const int i = obs_kf[e];
    double R[9], pc[3];
    quat_to_rot(&pose[7 * (size_t)i], R);
    mat3_vec(R, &pts[3 * (size_t)obs_pt[e]], pc);
    return pc[2] + pose[7 * (size_t)i + 6] > 0.0;
  };
  auto edge_chi2 = [&](int e) {
    const double* er = &ba.err[3 * (size_t)e];
    return ba.w[e] * ((er[0] * er[0] + er[1] * er[1]) + er[2] * er[2]);
  };
  // bDoMore = !*pbStopFlag (:791-795): a raised flag skips the reprojection gating, the kernel removal and stage 3
  if (!ba.terminate()) {
    for (int e = 0; e < E; ++e)
      if (edge_chi2(e) > (ba.stereo[e] ? 7.815 : 5.991) || !depth_positive(e)) ba.level[e] = 1;
    ba.robust = false;  // e->setRobustKernel(0)
    // stage 3: optimize(40)  (:826-828)
    its[2] = ba.optimize(40, pose, pts);
  }
  // results (:834-881)
  if (assoc_cleared)
    for (int l = 0; l < L; ++l) {
      assoc_cleared[l] = 0;
      if (ba.str_type[l] != 1) continue;
      double es[3];
      ba.structure_error(&pts[3 * (size_t)l], l, es);
      assoc_cleared[l] = ba.structure_chi2(l, es) > prm->tri_str_thresh * prm->ba_lambda2;
    }
  if (obs_outlier)
    for (int e = 0; e < E; ++e) obs_outlier[e] = (edge_chi2(e) > (ba.stereo[e] ? 7.815 : 5.991) || !depth_positive(e)) ? 1 : 0;
  for (size_t i = 0; i < pose.size(); ++i) pose_io[i] = pose[i];
  for (size_t i = 0; i < pts.size(); ++i) pts_io[i] = pts[i];
  if (iters) for (int s2 = 0; s2 < 3; ++s2) iters[s2] = its[s2];
  return 0;
}

}  // extern "C"
