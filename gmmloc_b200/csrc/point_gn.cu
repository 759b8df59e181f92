// point_gn.cu — K4: batched 3-DoF structure-constrained point solves (one thread per problem, fp64).
//
// Replaces GMMLoc::optimizePoint (reference gmmloc/src/gmmloc_opt.cpp:260-342): vertex x in R^3, one
// stereo reprojection edge (EdgeProjectXYZOnlyStereo, gmmloc/src/gmm/factors.cpp:116-168, information
// sigma2_inv[octave] * I3, gmmloc_opt.cpp:287-304) and one point-to-plane edge (EdgePt2GaussianDeg,
// factors.cpp:55-64, information tri_lambda2 * proj_z2, gmmloc_opt.cpp:308-319); Gauss-Newton, `iters`
// iterations with no early exit (gmmloc_opt.cpp:324); chi2 values are those of the last linearisation
// point; res = !(chi2_proj > 7.815) && !(tri_check_str_chi2 && chi2_str > tri_str_thresh * tri_lambda2)
// (gmmloc_opt.cpp:332-339).  The reference builds a g2o graph (heap optimizer + solver + vertex + 2 edges)
// per call; here each problem is ~300 fp64 operations in registers.
//
// THIS FILE IS COMPILED WITH -fmad=false: every product and sum is a separate IEEE rn operation, in the
// same expression order as the CPU oracle (which is built -ffp-contract=off), so pt_est / chi2 come out
// bit-identical.
#include "common.cuh"
#include "point_gn.cuh"

namespace gmmloc {

namespace {

__global__ void __launch_bounds__(128) point_gn_kernel(const PointArgs a) {
  const int i = blockIdx.x * blockDim.x + threadIdx.x;
  if (i >= a.M) return;
  const FrameParamsDev& P = a.prm;
  const int k = a.comp[i];
  if (k < 0 || k >= a.K) {  // no candidate: rejected, estimate = input
    for (int r = 0; r < 3; ++r) a.pt_est[(size_t)i * 3 + r] = a.x0[(size_t)i * 3 + r];
    a.chi2_proj[i] = 0.0;
    a.chi2_str[i] = 0.0;
    a.res[i] = 0;
    return;
  }
  double R[9];
  quat_to_rot_n(a.pose + (size_t)i * 7, R);
  const double t0 = a.pose[(size_t)i * 7 + 4], t1 = a.pose[(size_t)i * 7 + 5], t2 = a.pose[(size_t)i * 7 + 6];
  const double o0 = (double)a.obs[(size_t)i * 3], o1 = (double)a.obs[(size_t)i * 3 + 1], o2 = (double)a.obs[(size_t)i * 3 + 2];
  int oc = a.octave[i];
  if (oc > 7) oc = 7;
  const double w_proj = (double)P.s2inv[oc];
  const double w_str = P.tri_lambda2 * a.proj_z2[i];
  const double* ax = a.axis + (size_t)k * 9;
  const double n[3] = {ax[0], ax[3], ax[6]};  // axis_.col(0), gmmloc_opt.cpp:314
  const double* mu = a.mean + (size_t)k * 3;
  const double x0[3] = {a.x0[(size_t)i * 3], a.x0[(size_t)i * 3 + 1], a.x0[(size_t)i * 3 + 2]};
  const double t[3] = {t0, t1, t2}, o[3] = {o0, o1, o2};
  const PointSolve s = optimize_point_d(P, R, t, o, w_proj, w_str, n, mu, x0, a.iters);
  for (int r = 0; r < 3; ++r) a.pt_est[(size_t)i * 3 + r] = s.x[r];
  a.chi2_proj[i] = s.chi2_proj;
  a.chi2_str[i] = s.chi2_str;
  a.res[i] = s.res ? 1 : 0;
}

// ---- K4b: one candidate of Localization::optimizeTriangulationVec (localization_opt.cpp:27-204): two view
// edges (stereo factors.cpp:116-168 or mono factors.cpp:70-107, both weighted with view 1's sigma2_inv,
// :126-135) + one point-to-plane edge (information tri_lambda2, :106-107), `iters` (20) GN iterations.
// residual + point Jacobian of one view at camera-frame point (px, py, pz); mono rows: e2 = 0, J row 2 = 0
__device__ __forceinline__ void view_edge(const FrameParamsDev& P, const double* R, double px, double py, double pz,
                                          const double* ob, bool stereo, double* e, double* Jl) {
  const double z_2 = pz * pz;
  if (stereo) {
    const double invz = 1.0 / pz;
    const double u = px * invz * P.fx + P.cx;
    const double v = py * invz * P.fy + P.cy;
    const double ur = u - P.bf * invz;
    e[0] = ob[0] - u; e[1] = ob[1] - v; e[2] = ob[2] - ur;
  } else {
    const double u = px / pz * P.fx + P.cx;
    const double v = py / pz * P.fy + P.cy;
    e[0] = ob[0] - u; e[1] = ob[1] - v; e[2] = 0.0;
  }
  if (stereo) {  // element-wise form of EdgeProjectXYZOnlyStereo::linearizeOplus (factors.cpp:133-168)
    for (int j = 0; j < 3; ++j) {
      Jl[j] = -P.fx * R[j] / pz + P.fx * px * R[6 + j] / z_2;
      Jl[3 + j] = -P.fy * R[3 + j] / pz + P.fy * py * R[6 + j] / z_2;
      Jl[6 + j] = Jl[j] - P.bf * R[6 + j] / z_2;
    }
  } else {  // EdgeProjectXYZOnly::linearizeOplus: -1. / z * tmp * R (factors.cpp:84-107), Eigen's lazy-product order
    const double s = -1. / pz;
    const double T00 = s * P.fx, T01 = s * 0.0, T02 = s * (-px / pz * P.fx);
    const double T10 = s * 0.0, T11 = s * P.fy, T12 = s * (-py / pz * P.fy);
    for (int j = 0; j < 3; ++j) {
      Jl[j] = (T00 * R[j] + T01 * R[3 + j]) + T02 * R[6 + j];
      Jl[3 + j] = (T10 * R[j] + T11 * R[3 + j]) + T12 * R[6 + j];
      Jl[6 + j] = 0.0;
    }
  }
}

__global__ void __launch_bounds__(128) tri_gn_kernel(const TriArgs a) {
  const int i = blockIdx.x * blockDim.x + threadIdx.x;
  if (i >= a.M) return;
  const FrameParamsDev& P = a.prm;
  const int k = a.comp[i];
  if (k < 0 || k >= a.K) {
    for (int r = 0; r < 3; ++r) a.pt_est[(size_t)i * 3 + r] = a.x0[(size_t)i * 3 + r];
    a.err1[i] = 0.0; a.err2[i] = 0.0; a.chi2_str[i] = 0.0; a.res[i] = 0;
    return;
  }
  double R1[9], R2[9];
  quat_to_rot_n(a.pose1 + (size_t)i * 7, R1);
  quat_to_rot_n(a.pose2 + (size_t)i * 7, R2);
  const double* t1 = a.pose1 + (size_t)i * 7 + 4;
  const double* t2 = a.pose2 + (size_t)i * 7 + 4;
  const double ob1[3] = {(double)a.obs1[(size_t)i * 3], (double)a.obs1[(size_t)i * 3 + 1], (double)a.obs1[(size_t)i * 3 + 2]};
  const double ob2[3] = {(double)a.obs2[(size_t)i * 3], (double)a.obs2[(size_t)i * 3 + 1], (double)a.obs2[(size_t)i * 3 + 2]};
  const bool st1 = ob1[2] >= 0.0, st2 = ob2[2] >= 0.0;
  int oc = a.octave1[i];
  if (oc > 7) oc = 7;
  const double w = (double)P.s2inv[oc];
  const double w_str = P.tri_lambda2;
  const double* ax = a.axis + (size_t)k * 9;
  const double n[3] = {ax[0], ax[3], ax[6]};
  const double* mu = a.mean + (size_t)k * 3;
  double x[3] = {a.x0[(size_t)i * 3], a.x0[(size_t)i * 3 + 1], a.x0[(size_t)i * 3 + 2]};
  double c1 = 0, c2 = 0, cs = 0;
  bool ok = true;
  for (int it = 0; it < a.iters; ++it) {
    double e1[3], e2[3], J1[9], J2[9];
    {
      const double px = ((R1[0] * x[0] + R1[1] * x[1]) + R1[2] * x[2]) + t1[0];
      const double py = ((R1[3] * x[0] + R1[4] * x[1]) + R1[5] * x[2]) + t1[1];
      const double pz = ((R1[6] * x[0] + R1[7] * x[1]) + R1[8] * x[2]) + t1[2];
      view_edge(P, R1, px, py, pz, ob1, st1, e1, J1);
    }
    {
      const double px = ((R2[0] * x[0] + R2[1] * x[1]) + R2[2] * x[2]) + t2[0];
      const double py = ((R2[3] * x[0] + R2[4] * x[1]) + R2[5] * x[2]) + t2[1];
      const double pz = ((R2[6] * x[0] + R2[7] * x[1]) + R2[8] * x[2]) + t2[2];
      view_edge(P, R2, px, py, pz, ob2, st2, e2, J2);
    }
    c1 = w * ((e1[0] * e1[0] + e1[1] * e1[1]) + e1[2] * e1[2]);
    c2 = w * ((e2[0] * e2[0] + e2[1] * e2[1]) + e2[2] * e2[2]);
    const double es = (n[0] * (x[0] - mu[0]) + n[1] * (x[1] - mu[1])) + n[2] * (x[2] - mu[2]);
    cs = w_str * (es * es);
    double H[9], b[3];
    for (int r = 0; r < 3; ++r) {
      for (int c = 0; c < 3; ++c) {
        const double h1 = (J1[r] * J1[c] + J1[3 + r] * J1[3 + c]) + J1[6 + r] * J1[6 + c];
        const double h2 = (J2[r] * J2[c] + J2[3 + r] * J2[3 + c]) + J2[6 + r] * J2[6 + c];
        H[r * 3 + c] = (w * h1 + w * h2) + w_str * (n[r] * n[c]);
      }
      const double g1 = (J1[r] * e1[0] + J1[3 + r] * e1[1]) + J1[6 + r] * e1[2];
      const double g2 = (J2[r] * e2[0] + J2[3 + r] * e2[1]) + J2[6 + r] * e2[2];
      b[r] = -((w * g1 + w * g2) + w_str * (n[r] * es));
    }
    double Hi[9];
    if (!inverse3_d(H, Hi)) {
      ok = false;
      break;
    }
    for (int r = 0; r < 3; ++r) x[r] += (Hi[r * 3] * b[0] + Hi[r * 3 + 1] * b[1]) + Hi[r * 3 + 2] * b[2];
  }
  for (int r = 0; r < 3; ++r) a.pt_est[(size_t)i * 3 + r] = x[r];
  a.err1[i] = c1;
  a.err2[i] = c2;
  a.chi2_str[i] = cs;
  uint8_t res = ok ? 1 : 0;
  if (P.tri_check_str_chi2 && cs > P.tri_str_thresh * P.tri_lambda2) res = 0;  // localization_opt.cpp:176-179
  const double th1 = st1 ? 7.8 : 5.991, th2 = st2 ? 7.8 : 5.991;               // :118,127,133
  if (c1 > th1 || c2 > th2) res = 0;                                           // :184-186
  a.res[i] = res;
}

}  // namespace

cudaError_t launch_tri_gn(const TriArgs& a, cudaStream_t stream) {
  if (a.M <= 0) return cudaSuccess;
  const int threads = 128;
  tri_gn_kernel<<<(a.M + threads - 1) / threads, threads, 0, stream>>>(a);
  return cudaGetLastError();
}

cudaError_t launch_point_gn(const PointArgs& a, cudaStream_t stream) {
  if (a.M <= 0) return cudaSuccess;
  const int threads = 128;
  const int blocks = (a.M + threads - 1) / threads;
  point_gn_kernel<<<blocks, threads, 0, stream>>>(a);
  return cudaGetLastError();
}

}  // namespace gmmloc
