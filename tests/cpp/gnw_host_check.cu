// gnw_host_check.cu — HOST replay of the world-frame frame Gauss-Newton (K3w) with the very functions the kernel runs
// (gmmloc_b200/csrc/frame_gn_world.cuh, solve6_ldlt / se3_oplus_fast of gn_common.cuh), same sweep schedule (update pass
// from the stored (u0, W), linearise pass at the new pose, world-twist 6x6 system, adjoint back to the camera twist).
// TEST INFRASTRUCTURE: built by tests/test_gnw_host.py with `nvcc -x cu` as host-only code and compared with the oracle, so the
// algebra of the kernel is checked on the CPU before (and independently of) any GPU run.  Points of class F64 (which the
// kernel hands to the fp64 camera-frame path of frame_gn_point.cuh) take a plain fp64 evaluation of the same block here.
#include <cmath>
#include <cstdint>
#include <cstring>
#include <vector>
#ifdef GNW_DEBUG
#include <cstdio>
#include <cstdlib>
#define GNW_TRACE(...) std::fprintf(stderr, __VA_ARGS__)
#else
#define GNW_TRACE(...) ((void)0)
#endif

#include "../../gmmloc_b200/csrc/common.cuh"
#include "../../gmmloc_b200/csrc/gn_common.cuh"
#include "../../gmmloc_b200/csrc/frame_gn_world.cuh"

using namespace gmmloc;
using gnw::Cam;
using gnw::Lin;
using gnw::PointLin;
using gnw::S3f;
using gnw::V3f;

namespace {
enum : uint8_t { F_STEREO = 1, F_STR = 2, F_FREE = 4, F_DEG = 8 };
enum : int { C_OFF = 0, C_FIXED = 1, C_PLAIN = 2, C_MPLANE = 3, C_STRP = 4, C_STR = 5, C_MSTR = 6, C_F64 = 7 };
constexpr double kFarZ = 6.0;

struct D3 { double x, y, z; };
// generic fp64 block (world frame): D = G + S, u0 = D^-1 (h + b), W = D^-1 G, M = W^T S, v = h - G u0
bool general_fp64(const Cam& C, const double* R, const double* t, const double* x, const float* ob, bool stereo, double w, bool has_str,
                  bool deg, const double* mu, const double* cst, double lam, bool robust, double u0[3], double W[9], double M[6], double v[3],
                  bool free_pt) {
  const double pc[3] = {R[0] * x[0] + R[1] * x[1] + R[2] * x[2] + t[0], R[3] * x[0] + R[4] * x[1] + R[5] * x[2] + t[1],
                        R[6] * x[0] + R[7] * x[1] + R[8] * x[2] + t[2]};
  const double iz = 1.0 / pc[2];
  const double a = C.fx * iz, b = C.fy * iz, c0 = -a * pc[0] * iz, c1 = -b * pc[1] * iz, cz = C.bf * iz * iz;
  const double u = pc[0] * a + C.cx, vv = pc[1] * b + C.cy;
  double e[3] = {ob[0] - u, ob[1] - vv, stereo ? ob[2] - (u - C.bf * iz) : 0.0};
  if (robust) {
    const double chi = w * (e[0] * e[0] + e[1] * e[1] + e[2] * e[2]);
    const double dl = stereo ? (double)C.d_stereo : (double)C.d_mono;
    if (chi > dl * dl) w *= dl / std::sqrt(chi);
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
  double S[9] = {0}, bs[3] = {0, 0, 0};
  if (has_str) {
    const double d[3] = {x[0] - mu[0], x[1] - mu[1], x[2] - mu[2]};
    if (deg) {
      const double es = cst[0] * d[0] + cst[1] * d[1] + cst[2] * d[2];
      for (int i = 0; i < 3; ++i) {
        for (int j = 0; j < 3; ++j) S[i * 3 + j] = lam * cst[i] * cst[j];
        bs[i] = -lam * es * cst[i];
      }
    } else {
      const double Sw[9] = {cst[0], cst[1], cst[2], cst[1], cst[3], cst[4], cst[2], cst[4], cst[5]};
      for (int i = 0; i < 3; ++i) {
        for (int j = 0; j < 3; ++j) S[i * 3 + j] = Sw[i * 3 + j];
        bs[i] = -(Sw[i * 3] * d[0] + Sw[i * 3 + 1] * d[1] + Sw[i * 3 + 2] * d[2]);
      }
    }
  }
  if (!free_pt) {  // held point: M = G, v = h
    M[0] = G[0]; M[1] = G[1]; M[2] = G[2]; M[3] = G[4]; M[4] = G[5]; M[5] = G[8];
    for (int i = 0; i < 3; ++i) v[i] = h[i];
    return true;
  }
  double D[9], Di[9];
  for (int i = 0; i < 9; ++i) D[i] = G[i] + S[i];
  const double det = D[0] * (D[4] * D[8] - D[5] * D[7]) - D[1] * (D[3] * D[8] - D[5] * D[6]) + D[2] * (D[3] * D[7] - D[4] * D[6]);
  if (det == 0.0 || !std::isfinite(1.0 / det)) return false;
  const double id = 1.0 / det;
  Di[0] = (D[4] * D[8] - D[5] * D[7]) * id; Di[1] = (D[2] * D[7] - D[1] * D[8]) * id; Di[2] = (D[1] * D[5] - D[2] * D[4]) * id;
  Di[3] = (D[5] * D[6] - D[3] * D[8]) * id; Di[4] = (D[0] * D[8] - D[2] * D[6]) * id; Di[5] = (D[2] * D[3] - D[0] * D[5]) * id;
  Di[6] = (D[3] * D[7] - D[4] * D[6]) * id; Di[7] = (D[1] * D[6] - D[0] * D[7]) * id; Di[8] = (D[0] * D[4] - D[1] * D[3]) * id;
  double bl[3] = {h[0] + bs[0], h[1] + bs[1], h[2] + bs[2]};
  for (int i = 0; i < 3; ++i) {
    u0[i] = Di[i * 3] * bl[0] + Di[i * 3 + 1] * bl[1] + Di[i * 3 + 2] * bl[2];
    for (int j = 0; j < 3; ++j) W[i * 3 + j] = Di[i * 3] * G[j] + Di[i * 3 + 1] * G[3 + j] + Di[i * 3 + 2] * G[6 + j];
  }
  double Mf[9];
  for (int i = 0; i < 3; ++i) {
    v[i] = h[i] - (G[i * 3] * u0[0] + G[i * 3 + 1] * u0[1] + G[i * 3 + 2] * u0[2]);
    for (int j = 0; j < 3; ++j) Mf[i * 3 + j] = W[i] * S[j] + W[3 + i] * S[3 + j] + W[6 + i] * S[6 + j];
  }
  M[0] = Mf[0]; M[1] = Mf[1]; M[2] = Mf[2]; M[3] = Mf[4]; M[4] = Mf[5]; M[5] = Mf[8];
  return true;
}
}  // namespace

// mode 0: K3w (monocular + full Gaussian on the fp64 path), mode 1: K3s (that class in gain form from the covariance)
extern "C" int gnw_host_frame_solve_mode(const double* mean, const double* cov, const double* cov_inv, const double* axis,
                                         const uint8_t* is_deg, const FrameParamsDev* Pp, const double* pose0, const float* pts,
                                         const float* obs, const uint8_t* octave, const int32_t* assign, int N, double* pose_out,
                                         double* pts_out, uint8_t* cls_out, int mode);
extern "C" int gnw_host_frame_solve(const double* mean, const double* cov_inv, const double* axis, const uint8_t* is_deg,
                                    const FrameParamsDev* Pp, const double* pose0, const float* pts, const float* obs,
                                    const uint8_t* octave, const int32_t* assign, int N, double* pose_out, double* pts_out,
                                    uint8_t* cls_out) {
  return gnw_host_frame_solve_mode(mean, nullptr, cov_inv, axis, is_deg, Pp, pose0, pts, obs, octave, assign, N, pose_out, pts_out, cls_out, 0);
}
extern "C" int gnw_host_frame_solve_mode(const double* mean, const double* cov, const double* cov_inv, const double* axis, const uint8_t* is_deg,
                                    const FrameParamsDev* Pp, const double* pose0, const float* pts, const float* obs,
                                    const uint8_t* octave, const int32_t* assign, int N, double* pose_out, double* pts_out,
                                    uint8_t* cls_out, int mode) {
  const FrameParamsDev& P = *Pp;
  const bool pose_free = P.factor_mode != GMMLOC_MODE_POINTS_ONLY, pts_free = P.factor_mode != GMMLOC_MODE_POSE_ONLY;
  Cam C;
  C.fx = P.fx; C.fy = P.fy; C.cx = P.cx; C.cy = P.cy; C.bf = P.bf;
  C.fxf = (float)P.fx; C.fyf = (float)P.fy; C.bff = (float)P.bf; C.lam = (float)P.ba_lambda2;
  C.d_mono = (float)P.delta_mono; C.d_stereo = (float)P.delta_stereo; C.robust = P.robust;
  double pose[7], R[9], t[3], Ro[9], to[3], c0[3];
  for (int i = 0; i < 7; ++i) pose[i] = pose0[i];
  quat_normalize(pose);
  quat_to_rot(pose, R);
  for (int i = 0; i < 3; ++i) t[i] = pose[4 + i];
  for (int c = 0; c < 3; ++c) c0[c] = -(R[c] * t[0] + R[3 + c] * t[1] + R[6 + c] * t[2]);
  const double view_cos_thresh = std::cos(78.0 * 3.14159265358979323846 / 180.0);
  std::vector<double> x(3 * (size_t)N);
  std::vector<uint8_t> fl(N);
  std::vector<float> store(12 * (size_t)N, 0.f);
  for (int i = 0; i < N; ++i) {  // classification: the arithmetic of frame_prep_kernel
    for (int r = 0; r < 3; ++r) x[3 * i + r] = (double)pts[3 * i + r];
    const double x0 = x[3 * i], x1 = x[3 * i + 1], x2 = x[3 * i + 2];
    uint8_t b = (obs[3 * i + 2] >= 0.0f) ? F_STEREO : 0;
    const int k = assign[i];
    if (k >= 0 && pts_free) {
      const double* mu = mean + 3 * (size_t)k;
      if (is_deg[k]) {
        b |= F_DEG;
        const double* ax = axis + 9 * (size_t)k;
        const double e = (ax[0] * (x0 - mu[0]) + ax[3] * (x1 - mu[1])) + ax[6] * (x2 - mu[2]);
        if (e * e <= P.tri_str_thresh) b |= F_STR;
        const double p0 = mu[0] - c0[0], p1 = mu[1] - c0[1], p2 = mu[2] - c0[2];
        const double vc = std::fabs(p0 * ax[0] + p1 * ax[3] + p2 * ax[6]) / std::sqrt(p0 * p0 + p1 * p1 + p2 * p2);
        if (vc < view_cos_thresh) b &= ~F_STR;
      } else {
        const double* A = cov_inv + 9 * (size_t)k;
        const double d0 = x0 - mu[0], d1 = x1 - mu[1], d2 = x2 - mu[2];
        const double r0 = (d0 * A[0] + d1 * A[3]) + d2 * A[6], r1 = (d0 * A[1] + d1 * A[4]) + d2 * A[7], r2 = (d0 * A[2] + d1 * A[5]) + d2 * A[8];
        if ((r0 * d0 + r1 * d1) + r2 * d2 <= P.mdist2_gate) b |= F_STR;
      }
    }
    if (pts_free && ((b & F_STEREO) || (b & F_STR))) b |= F_FREE;
    int cls;
    if (!(b & F_FREE)) cls = pose_free ? C_FIXED : C_OFF;
    else {
      const double pz = R[6] * x0 + R[7] * x1 + R[8] * x2 + t[2];
      const double rx = x0 - c0[0], ry = x1 - c0[1], rz = x2 - c0[2];
      if (!((pz > 0.1) && (pz < kFarZ) && (rx * rx + ry * ry + rz * rz < 2.0 * kFarZ * pz))) cls = C_F64;  // fp32_geometry_ok()
      else if (!(b & F_STEREO)) cls = (b & F_DEG) ? C_MPLANE : C_MSTR;
      else cls = (b & F_STR) ? ((b & F_DEG) ? C_STRP : C_STR) : C_PLAIN;
    }
    if (cls_out) cls_out[i] = (uint8_t)cls;
    if (mode == 0 && cls == C_MSTR) cls = C_F64;  // K3w: fp64 path
    fl[i] = (uint8_t)(b | (cls << 4));
  }
  std::vector<double> u0d(3 * (size_t)N, 0.0), Wd(9 * (size_t)N, 0.0);  // fp64 stored linearisation of the F64-class points
  float dxw[6] = {0, 0, 0, 0, 0, 0};
  double dww[6] = {0, 0, 0, 0, 0, 0};
  const int iters = P.gn_iters, sweeps = iters > 0 ? iters + 1 : 0;
  int fail = 0;
  for (int sw = 0; sw < sweeps && !fail; ++sw) {
    const bool do_update = pts_free && sw > 0, do_lin = sw < iters, do_acc = pose_free && do_lin;
    if (do_update) {
      const V3f om = {dxw[0], dxw[1], dxw[2]}, up = {dxw[3], dxw[4], dxw[5]};
      for (int j = 0; j < N; ++j) {
        const int cls = fl[j] >> 4;
        if (cls < C_PLAIN) continue;
        double* xp = &x[3 * j];
        if (cls == C_F64) {
          const double xr[3] = {xp[0] - c0[0], xp[1] - c0[1], xp[2] - c0[2]};
          const double a[3] = {dww[1] * xr[2] - dww[2] * xr[1] + dww[3], dww[2] * xr[0] - dww[0] * xr[2] + dww[4],
                               dww[0] * xr[1] - dww[1] * xr[0] + dww[5]};
          for (int r = 0; r < 3; ++r) {
            xp[r] += u0d[3 * j + r] - (Wd[9 * j + 3 * r] * a[0] + Wd[9 * j + 3 * r + 1] * a[1] + Wd[9 * j + 3 * r + 2] * a[2]);
            if (!std::isfinite(xp[r])) fail = 1;
          }
          continue;
        }
        const float* sp = &store[12 * (size_t)j];
        const V3f xr = {(float)(xp[0] - c0[0]), (float)(xp[1] - c0[1]), (float)(xp[2] - c0[2])};
        const V3f av = gnw::twist_at(om, up, xr);
        V3f dx;
        if (cls == C_PLAIN) dx = {sp[0] - av.x, sp[1] - av.y, sp[2] - av.z};
        else dx = gnw::point_step({sp[0], sp[1], sp[2]}, sp + 3, av);
#ifdef GNW_DEBUG
        if (getenv("GNW_TRACE_POINT") && j == atoi(getenv("GNW_TRACE_POINT")))
          GNW_TRACE("sweep %d point %d class %d: x %.6f %.6f %.6f dx %g %g %g u0 %g %g %g W %g %g %g / %g %g %g / %g %g %g av %g %g %g\n", sw, j, cls, xp[0], xp[1], xp[2], dx.x, dx.y, dx.z, sp[0], sp[1], sp[2], sp[3], sp[4], sp[5], sp[6], sp[7], sp[8], sp[9], sp[10], sp[11], av.x, av.y, av.z);
#endif
        if (!gnw::finite3(dx)) { fail = 1; GNW_TRACE("sweep %d: non-finite step, point %d class %d\n", sw, j, cls); }
        xp[0] += (double)dx.x; xp[1] += (double)dx.y; xp[2] += (double)dx.z;
      }
    }
    if (!do_lin || fail) continue;
    float Rl[9];
    for (int i = 0; i < 9; ++i) Rl[i] = (float)R[i];
    double accd[27];
    for (int i = 0; i < 27; ++i) accd[i] = 0.0;
    for (int j = 0; j < N; ++j) {
      const uint8_t b = fl[j];
      const int cls = b >> 4;
      if (cls == C_OFF) continue;
      if (cls == C_FIXED && !do_acc) continue;
      const double* xp = &x[3 * j];
      const float* ob = obs + 3 * (size_t)j;
      const float w = P.s2inv[octave[j] & 7];
      const bool stereo = b & F_STEREO;
      float acc[27];
      for (int i = 0; i < 27; ++i) acc[i] = 0.f;
      const V3f xr = {(float)(xp[0] - c0[0]), (float)(xp[1] - c0[1]), (float)(xp[2] - c0[2])};
      const int k = assign[j];
      if (cls == C_F64) {
        double cst[6] = {0, 0, 0, 0, 0, 0};
        if (b & F_STR) {
          if (b & F_DEG) { cst[0] = axis[9 * (size_t)k]; cst[1] = axis[9 * (size_t)k + 3]; cst[2] = axis[9 * (size_t)k + 6]; }
          else { const double* A = cov_inv + 9 * (size_t)k; cst[0] = A[0]; cst[1] = A[1]; cst[2] = A[2]; cst[3] = A[4]; cst[4] = A[5]; cst[5] = A[8]; }
        }
        double M[6], v[3];
        if (!general_fp64(C, R, t, xp, ob, stereo, (double)w, b & F_STR, b & F_DEG, k >= 0 ? mean + 3 * (size_t)k : nullptr, cst,
                          P.ba_lambda2, P.robust != 0, &u0d[3 * j], &Wd[9 * j], M, v, true)) { fail = 1; break; }
        if (do_acc && (b & F_STR)) {
          const S3f Mw = {(float)M[0], (float)M[1], (float)M[2], (float)M[3], (float)M[4], (float)M[5]};
          gnw::accumulate_pose(acc, xr, Mw, {(float)v[0], (float)v[1], (float)v[2]});
        }
      } else {
        const Lin L = gnw::linearise(C, R, t, xp[0], xp[1], xp[2], ob, stereo);
        float* sp = &store[12 * (size_t)j];
        if (cls == C_PLAIN) {
          const V3f u0 = gnw::lin_plain(C, L, Rl);
          sp[0] = u0.x; sp[1] = u0.y; sp[2] = u0.z;
        } else if (cls == C_FIXED) {
          S3f G; V3f h;
          gnw::lin_fixed(C, L, Rl, stereo, w, G, h);
          gnw::accumulate_pose(acc, xr, G, h);
        } else {
          const double* mu = mean + 3 * (size_t)k;
          const V3f d = {(float)(xp[0] - mu[0]), (float)(xp[1] - mu[1]), (float)(xp[2] - mu[2])};
          PointLin pl;
          if (cls == C_MSTR) {
            const double* Cv = cov + 9 * (size_t)k;
            const S3f Sig = {(float)Cv[0], (float)Cv[1], (float)Cv[2], (float)Cv[4], (float)Cv[5], (float)Cv[8]};
            S3f M; V3f v;
            if (!gnw::lin_mstr_gain(C, L, Rl, w, Sig, d, pl, M, v)) { fail = 1; GNW_TRACE("sweep %d: lin_mstr_gain failed, point %d\n", sw, j); break; }
            if (do_acc) gnw::accumulate_pose(acc, xr, M, v);
          } else if (cls == C_STR) {
            const double* A = cov_inv + 9 * (size_t)k;
            const S3f Sw = {(float)A[0], (float)A[1], (float)A[2], (float)A[4], (float)A[5], (float)A[8]};
            S3f M; V3f v;
            if (!gnw::lin_str(C, L, Rl, stereo, w, Sw, d, pl, do_acc, M, v)) { fail = 1; GNW_TRACE("sweep %d: lin_str failed, point %d class %d pc %g %g %g e %g %g %g\n", sw, j, cls, L.pc.x, L.pc.y, L.pc.z, L.e.x, L.e.y, L.e.z); break; }
            if (do_acc) gnw::accumulate_pose(acc, xr, M, v);
#ifdef GNW_DEBUG
            if (do_acc) {
              double cst[6] = {A[0], A[1], A[2], A[4], A[5], A[8]}, u0q[3], Wq[9], Mq[6], vq[3];
              general_fp64(C, R, t, xp, ob, stereo, (double)w, true, false, mu, cst, P.ba_lambda2, P.robust != 0, u0q, Wq, Mq, vq, true);
              const float Mf[6] = {M.xx, M.xy, M.xz, M.yy, M.yz, M.zz};
              double mx = 0, err = 0;
              for (int i = 0; i < 6; ++i) { mx = std::fmax(mx, std::fabs(Mq[i])); err = std::fmax(err, std::fabs(Mq[i] - Mf[i])); }
              const double verr = std::fmax(std::fabs(vq[0] - v.x), std::fmax(std::fabs(vq[1] - v.y), std::fabs(vq[2] - v.z)));
              const double vmx = std::fmax(std::fabs(vq[0]), std::fmax(std::fabs(vq[1]), std::fabs(vq[2])));
              if (err > 1e-2 * mx || verr > 1e-2 * vmx + 1e-3)
                GNW_TRACE("sweep %d point %d class %d: M err %g of %g, v err %g of %g | pc %g %g %g e %g %g %g\n", sw, j, cls, err, mx, verr, vmx, L.pc.x, L.pc.y, L.pc.z, L.e.x, L.e.y, L.e.z);
            }
            if (do_acc && (M.xx < 0 || M.yy < 0 || M.zz < 0)) {
              GNW_TRACE("  point %d: M diag %g %g %g | S %g %g %g %g %g %g | W %g %g %g / %g %g %g / %g %g %g | pc %g %g %g e %g %g %g w %g d %g %g %g\n", j,
                        M.xx, M.yy, M.zz, Sw.xx, Sw.xy, Sw.xz, Sw.yy, Sw.yz, Sw.zz, pl.w0, pl.w1, pl.w2, pl.w3, pl.w4, pl.w5, pl.w6, pl.w7,
                        pl.w8, L.pc.x, L.pc.y, L.pc.z, L.e.x, L.e.y, L.e.z, w, d.x, d.y, d.z);
            }
#endif
          } else {
            const V3f n = {(float)axis[9 * (size_t)k], (float)axis[9 * (size_t)k + 3], (float)axis[9 * (size_t)k + 6]};
            const float es = gnw::dotf(n, d);
            if (cls == C_STRP) {
              float kappa, rho;
              if (!gnw::lin_strp(C, L, Rl, w, n, es, pl, kappa, rho)) { fail = 1; GNW_TRACE("sweep %d: lin_strp failed, point %d\n", sw, j); break; }
              if (do_acc) gnw::accumulate_pose_rank1(acc, xr, n, kappa, rho);
            } else if (!gnw::lin_mplane(C, L, Rl, n, es, pl)) { fail = 1; GNW_TRACE("sweep %d: lin_mplane failed, point %d pc %g %g %g es %g\n", sw, j, L.pc.x, L.pc.y, L.pc.z, es); break; }
          }
          sp[0] = pl.u0.x; sp[1] = pl.u0.y; sp[2] = pl.u0.z;
          sp[3] = pl.w0; sp[4] = pl.w1; sp[5] = pl.w2; sp[6] = pl.w3; sp[7] = pl.w4; sp[8] = pl.w5; sp[9] = pl.w6; sp[10] = pl.w7; sp[11] = pl.w8;
        }
      }
      for (int i = 0; i < 27; ++i) accd[i] += (double)acc[i];
#ifdef GNW_DEBUG
      {
        // per-point contribution check: the 6x6 block of one point must be positive semidefinite
        const double dg[6] = {acc[0], acc[6], acc[11], acc[15], acc[18], acc[20]};
        for (int i = 0; i < 6; ++i)
          if (dg[i] < 0.0 || !std::isfinite(dg[i])) {
            GNW_TRACE("sweep %d point %d class %d: diagonal %d of its block = %g (x' %g %g %g)\n", sw, j, cls, i, dg[i], xr.x, xr.y, xr.z);
            break;
          }
      }
#endif
    }
    if (fail) break;
    if (do_acc) {
      double dw[6];
      bool ok = true;
      for (int i = 0; i < 27; ++i) ok = ok && std::isfinite(accd[i]);
      if (!ok || !solve6_ldlt(accd, dw)) { fail = 1; GNW_TRACE("sweep %d: pose system failed (finite %d)\n", sw, (int)ok); break; }
      double tp[3], dc[6];
      for (int r = 0; r < 3; ++r) {
        tp[r] = t[r] + (R[r * 3] * c0[0] + R[r * 3 + 1] * c0[1] + R[r * 3 + 2] * c0[2]);
        dc[r] = R[r * 3] * dw[0] + R[r * 3 + 1] * dw[1] + R[r * 3 + 2] * dw[2];
        dc[3 + r] = R[r * 3] * dw[3] + R[r * 3 + 1] * dw[4] + R[r * 3 + 2] * dw[5];
      }
      dc[3] += tp[1] * dc[2] - tp[2] * dc[1];
      dc[4] += tp[2] * dc[0] - tp[0] * dc[2];
      dc[5] += tp[0] * dc[1] - tp[1] * dc[0];
      double np[7];
      for (int i = 0; i < 9; ++i) Ro[i] = R[i];
      for (int i = 0; i < 3; ++i) to[i] = t[i];
      se3_oplus_fast(dc, pose, np, R);
      for (int i = 0; i < 7; ++i) pose[i] = np[i];
      for (int i = 0; i < 3; ++i) t[i] = np[4 + i];
      for (int i = 0; i < 6; ++i) { dxw[i] = (float)dw[i]; dww[i] = dw[i]; }
    }
  }
  (void)Ro; (void)to;
  if (fail) {
    for (int i = 0; i < 7; ++i) pose[i] = pose0[i];
    quat_normalize(pose);
    for (int i = 0; i < 3 * N; ++i) x[i] = (double)pts[i];
  }
  for (int i = 0; i < 7; ++i) pose_out[i] = pose[i];
  for (int i = 0; i < 3 * N; ++i) pts_out[i] = x[i];
  return fail;
}
