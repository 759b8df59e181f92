// map_build.cu — K0: per-component constants on the device (one thread per component, fp64).
//
// Replaces the GaussianComponent constructor + decompose() (reference gmmloc/include/gmmloc/gmm/gaussian.h:30-39,
// gmmloc/src/gmm/gaussian.cpp:36-63): cov_inv, det, eigen-decomposition (scale ascending, axis),
// is_degenerated (scale.x < deg_eps), sqrt_info = LLT(cov_inv).L, is_salient; plus the fp32 whitening
// table the association scan reads and the scan's rigorous error bound alpha.
//
// Every product / sum is an explicit round-to-nearest intrinsic so that results are reproducible to
// the bit on any IEEE machine that evaluates the same expression trees without contraction.
#include <cuda_fp16.h>
#include <math_constants.h>

#include "common.cuh"

namespace gmmloc {

namespace {

__device__ __forceinline__ double& E(double* a, int r, int c) { return a[r * 3 + c]; }
__device__ __forceinline__ double E(const double* a, int r, int c) { return a[r * 3 + c]; }

__device__ double cof3(const double* m, int i, int j) {
  const int i1 = (i + 1) % 3, i2 = (i + 2) % 3, j1 = (j + 1) % 3, j2 = (j + 2) % 3;
  return dsub(dmul(E(m, i1, j1), E(m, i2, j2)), dmul(E(m, i1, j2), E(m, i2, j1)));
}

// 3x3 inverse by cofactors: determinant from the first cofactor column, one reciprocal (Eigen's
// fixed-size 3x3 inverse, which gaussian.h:31 calls).
__device__ void inv3(const double* m, double* out) {
  const double c00 = cof3(m, 0, 0), c10 = cof3(m, 1, 0), c20 = cof3(m, 2, 0);
  const double det = dadd(dadd(dmul(c00, E(m, 0, 0)), dmul(c10, E(m, 1, 0))), dmul(c20, E(m, 2, 0)));
  const double invdet = __ddiv_rn(1.0, det);
  E(out, 0, 0) = dmul(c00, invdet);
  E(out, 0, 1) = dmul(c10, invdet);
  E(out, 0, 2) = dmul(c20, invdet);
  E(out, 1, 0) = dmul(cof3(m, 0, 1), invdet);
  E(out, 1, 1) = dmul(cof3(m, 1, 1), invdet);
  E(out, 1, 2) = dmul(cof3(m, 2, 1), invdet);
  E(out, 2, 0) = dmul(cof3(m, 0, 2), invdet);
  E(out, 2, 1) = dmul(cof3(m, 1, 2), invdet);
  E(out, 2, 2) = dmul(cof3(m, 2, 2), invdet);
}

__device__ double det3_term(const double* m, int a, int b, int c) {
  return dmul(E(m, 0, a), dsub(dmul(E(m, 1, b), E(m, 2, c)), dmul(E(m, 1, c), E(m, 2, b))));
}
__device__ double det3(const double* m) {
  return dadd(dsub(det3_term(m, 0, 1, 2), det3_term(m, 1, 0, 2)), det3_term(m, 2, 0, 1));
}

// unblocked lower Cholesky; returns false when a pivot is not positive
__device__ bool chol3(const double* a, double* L) {
  for (int i = 0; i < 9; ++i) L[i] = 0.0;
  bool ok = true;
  for (int k = 0; k < 3; ++k) {
    double x = E(a, k, k);
    for (int j = 0; j < k; ++j) x = dsub(x, dmul(E(L, k, j), E(L, k, j)));
    if (!(x > 0.0)) ok = false;
    const double d = __dsqrt_rn(x);
    E(L, k, k) = d;
    for (int i = k + 1; i < 3; ++i) {
      double s = E(a, i, k);
      for (int j = 0; j < k; ++j) s = dsub(s, dmul(E(L, i, j), E(L, k, j)));
      E(L, i, k) = __ddiv_rn(s, d);
    }
  }
  return ok;
}

// cyclic Jacobi on a symmetric 3x3 using only + - * / sqrt; eigenvalues ascending, eigenvectors in
// the columns of `evecs`.  Stand-in for Eigen::SelfAdjointEigenSolver (gaussian.cpp:39-42).
__device__ void jacobi3(const double* cov, double* evals, double* evecs) {
  double a[9], v[9] = {1, 0, 0, 0, 1, 0, 0, 0, 1};
  for (int i = 0; i < 9; ++i) a[i] = cov[i];
  E(a, 1, 0) = E(a, 0, 1);
  E(a, 2, 0) = E(a, 0, 2);
  E(a, 2, 1) = E(a, 1, 2);
  for (int sweep = 0; sweep < 64; ++sweep) {
    const double off = dadd(dadd(fabs(E(a, 0, 1)), fabs(E(a, 0, 2))), fabs(E(a, 1, 2)));
    if (off == 0.0) break;
    for (int r = 0; r < 3; ++r) {
      const int p = (r == 2) ? 1 : 0, q = (r == 0) ? 1 : 2;
      const double apq = E(a, p, q);
      if (apq == 0.0) continue;
      const double app = E(a, p, p), aqq = E(a, q, q);
      const double g = dmul(100.0, fabs(apq));
      if (sweep > 3 && dadd(fabs(app), g) == fabs(app) && dadd(fabs(aqq), g) == fabs(aqq)) {
        E(a, p, q) = 0.0;
        E(a, q, p) = 0.0;
        continue;
      }
      const double h = dsub(aqq, app);
      double t;
      if (dadd(fabs(h), g) == fabs(h)) {
        t = __ddiv_rn(apq, h);
      } else {
        const double theta = __ddiv_rn(dmul(0.5, h), apq);
        t = __ddiv_rn(1.0, dadd(fabs(theta), __dsqrt_rn(dadd(1.0, dmul(theta, theta)))));
        if (theta < 0.0) t = -t;
      }
      const double c = __ddiv_rn(1.0, __dsqrt_rn(dadd(1.0, dmul(t, t))));
      const double s = dmul(t, c);
      const double tau = __ddiv_rn(s, dadd(1.0, c));
      E(a, p, p) = dsub(app, dmul(t, apq));
      E(a, q, q) = dadd(aqq, dmul(t, apq));
      E(a, p, q) = 0.0;
      E(a, q, p) = 0.0;
      const int k = 3 - p - q;
      const double akp = E(a, k, p), akq = E(a, k, q);
      const double nkp = dsub(akp, dmul(s, dadd(akq, dmul(akp, tau))));
      const double nkq = dadd(akq, dmul(s, dsub(akp, dmul(akq, tau))));
      E(a, k, p) = nkp;
      E(a, p, k) = nkp;
      E(a, k, q) = nkq;
      E(a, q, k) = nkq;
      for (int i = 0; i < 3; ++i) {
        const double vip = E(v, i, p), viq = E(v, i, q);
        E(v, i, p) = dsub(vip, dmul(s, dadd(viq, dmul(vip, tau))));
        E(v, i, q) = dadd(viq, dmul(s, dsub(vip, dmul(viq, tau))));
      }
    }
  }
  int idx[3] = {0, 1, 2};
  const double d[3] = {E(a, 0, 0), E(a, 1, 1), E(a, 2, 2)};
  for (int i = 1; i < 3; ++i)
    for (int j = i; j > 0 && d[idx[j]] < d[idx[j - 1]]; --j) {
      const int t = idx[j];
      idx[j] = idx[j - 1];
      idx[j - 1] = t;
    }
  for (int c = 0; c < 3; ++c) {
    evals[c] = d[idx[c]];
    for (int r = 0; r < 3; ++r) E(evecs, r, c) = E(v, r, idx[c]);
  }
}

// one whitening row (3 coefficients + constant) of component j of its tile into the UMMA B tile:
// K slots  0..2 w_hi | 3 c_hi | 4..6 w_hi (pairs with the landmark's lo part) | 7 c_lo | 8..10 w_lo | 11..15 zero
// against the landmark row  0..2 v_hi | 3 one | 4..6 v_lo | 7 one | 8..10 v_hi | 11..15 zero.
__device__ void tc_write_row(uint16_t* tile, int row, const double* w, double c, bool& fits) {
  __half h[kTcK];
  for (int i = 0; i < kTcK; ++i) h[i] = __float2half_rn(0.f);
  for (int i = 0; i < 3; ++i) {
    const __half hi = __double2half(w[i]);
    const __half lo = __double2half(w[i] - (double)__half2float(hi));
    h[i] = hi;
    h[4 + i] = hi;
    h[8 + i] = lo;
    if (!(fabs(w[i]) < 60000.0)) fits = false;
  }
  const __half chi = __double2half(c);
  h[3] = chi;
  h[7] = __double2half(c - (double)__half2float(chi));
  if (!(fabs(c) < 60000.0)) fits = false;
  for (int i = 0; i < kTcK; ++i) {
    const size_t off = (size_t)(i / 8) * (kTcRows * 8) + (size_t)row * 8 + (i % 8);
    tile[off] = __half_as_ushort(h[i]);
  }
}

__global__ void __launch_bounds__(128) map_build_kernel(MapDev m, double deg_eps, double xmax, int* err_flag,
                                                        unsigned int* alpha_bits) {
  const int k = blockIdx.x * blockDim.x + threadIdx.x;
  const int ktc = (m.Kpad + kTcTileComps - 1) / kTcTileComps * kTcTileComps;  // the tensor-core table is padded to whole tiles
  if (k >= ktc) return;
  float* sp = m.scan;
  uint16_t* tc_tile = m.tc_table ? m.tc_table + (size_t)(k / kTcTileComps) * (kTcTileBytes / 2) : nullptr;
  const int tj = k % kTcTileComps;
  if (k >= m.K) {  // padding rows score +inf and can never win
    if (k < m.Kpad) {
      for (int i = 0; i < kScanFloats; ++i) sp[scan_index(k, i)] = 0.f;
      sp[scan_index(k, 6)] = -CUDART_INF_F;
    }
    if (tc_tile) {  // tensor-core table: a constant whitened coordinate of 6e4 (score 3.6e9), finite in fp16
      bool fits = true;
      const double z[3] = {0.0, 0.0, 0.0};
      tc_write_row(tc_tile, 0 * kTcTileComps + tj, z, 59999.0, fits);
      tc_write_row(tc_tile, 1 * kTcTileComps + tj, z, 0.0, fits);
      tc_write_row(tc_tile, 2 * kTcTileComps + tj, z, 0.0, fits);
    }
    return;
  }
  double cov[9], ci[9], L[9], ev[3], ax[9];
  for (int i = 0; i < 9; ++i) cov[i] = m.cov[(size_t)k * 9 + i];
  const double mu[3] = {m.mean[(size_t)k * 3], m.mean[(size_t)k * 3 + 1], m.mean[(size_t)k * 3 + 2]};
  inv3(cov, ci);
  const double det = det3(cov);
  jacobi3(cov, ev, ax);
  const bool ok_l = chol3(ci, L);
  bool ok = ok_l && (det > 0.0) && (ev[0] > 0.0);
  for (int i = 0; i < 9; ++i) ok = ok && isfinite(ci[i]) && isfinite(L[i]);
  for (int i = 0; i < 3; ++i) ok = ok && isfinite(mu[i]);
  for (int i = 0; i < 9; ++i) {
    m.cov_inv[(size_t)k * 9 + i] = ci[i];
    m.axis[(size_t)k * 9 + i] = ax[i];
    m.sqrt_info[(size_t)k * 9 + i] = L[i];
  }
  for (int i = 0; i < 3; ++i) m.scale[(size_t)k * 3 + i] = ev[i];
  m.det[k] = det;
  m.is_deg[k] = ev[0] < deg_eps;                // gaussian.cpp:44
  if (m.gn_tab) {
    float* gt = m.gn_tab + (size_t)k * 8;
    if (ev[0] < deg_eps) {
      gt[0] = (float)ax[0]; gt[1] = (float)ax[3]; gt[2] = (float)ax[6];
      gt[3] = gt[4] = gt[5] = 0.f;
    } else {
      gt[0] = (float)ci[0]; gt[1] = (float)ci[1]; gt[2] = (float)ci[2]; gt[3] = (float)ci[4]; gt[4] = (float)ci[5]; gt[5] = (float)ci[8];
    }
    gt[6] = gt[7] = 0.f;
  }
  m.is_sal[k] = (ev[1] > 0.2 && ev[2] > 0.2);   // gaussian.cpp:50-55
  if (!ok) {
    atomicOr(err_flag, 1);
    for (int i = 0; i < kScanFloats; ++i) sp[scan_index(k, i)] = 0.f;
    sp[scan_index(k, 6)] = -CUDART_INF_F;
    if (m.rec) {
      for (int i = 0; i < 12; ++i) m.rec[(size_t)k * 12 + i] = 0.f;
      m.rec[(size_t)k * 12 + 6] = -CUDART_INF_F;
    }
    if (tc_tile) atomicOr(err_flag + 3, 1);
    return;
  }
  // fp32 whitening rows U = L^T, c = U mu
  const double U[6] = {E(L, 0, 0), E(L, 1, 0), E(L, 2, 0), E(L, 1, 1), E(L, 2, 1), E(L, 2, 2)};
  const double c0 = U[0] * mu[0] + U[1] * mu[1] + U[2] * mu[2];
  const double c1 = U[3] * mu[1] + U[4] * mu[2];
  const double c2 = U[5] * mu[2];
  for (int i = 0; i < 6; ++i) sp[scan_index(k, i)] = (float)U[i];
  sp[scan_index(k, 6)] = (float)(-c0);
  sp[scan_index(k, 7)] = (float)(-c1);
  sp[scan_index(k, 8)] = (float)(-c2);
  if (m.rec) {  // the same nine values, contiguous per component (association grid kernels)
    float* r = m.rec + (size_t)k * 12;
    for (int i = 0; i < 6; ++i) r[i] = (float)U[i];
    r[6] = (float)(-c0);
    r[7] = (float)(-c1);
    r[8] = (float)(-c2);
    r[9] = r[10] = r[11] = 0.f;
  }
  // error bound: |sqrt(score32) - d_exact| <= alpha_k for |x|_inf <= xmax,
  // alpha_k = (6 u32 + sqrt(64 u64)) * | |U| 1 (xmax + |mu|_inf) |_2
  const double mmax = fmax(fabs(mu[0]), fmax(fabs(mu[1]), fabs(mu[2])));
  const double ext = xmax + mmax;
  const double a0 = (fabs(U[0]) + fabs(U[1]) + fabs(U[2])) * ext;
  const double a1 = (fabs(U[3]) + fabs(U[4])) * ext;
  const double a2 = fabs(U[5]) * ext;
  const double A = sqrt(a0 * a0 + a1 * a1 + a2 * a2);
  const double u32 = 5.9604644775390625e-08, u64 = 1.1102230246251565e-16;
  const double alpha = (6.0 * u32 + sqrt(64.0 * u64)) * A;
  float af = (float)alpha;
  if ((double)af < alpha) af = nextafterf(af, CUDART_INF_F);
  atomicMax(alpha_bits, __float_as_uint(af));  // non-negative floats order like their bit patterns
  // ---- tensor-core pruning table (assoc_scan_tc_kernel): y = U (x - centre) - U (mu - centre) ----
  if (tc_tile) {
    const double cen[3] = {(double)m.tc_centre[0], (double)m.tc_centre[1], (double)m.tc_centre[2]};
    const double mc[3] = {mu[0] - cen[0], mu[1] - cen[1], mu[2] - cen[2]};
    const double w0[3] = {U[0], U[1], U[2]}, w1[3] = {0.0, U[3], U[4]}, w2[3] = {0.0, 0.0, U[5]};
    const double k0 = -(U[0] * mc[0] + U[1] * mc[1] + U[2] * mc[2]), k1 = -(U[3] * mc[1] + U[4] * mc[2]), k2 = -(U[5] * mc[2]);
    bool fits = true;
    tc_write_row(tc_tile, 0 * kTcTileComps + tj, w0, k0, fits);
    tc_write_row(tc_tile, 1 * kTcTileComps + tj, w1, k1, fits);
    tc_write_row(tc_tile, 2 * kTcTileComps + tj, w2, k2, fits);
    if (!fits) atomicOr(err_flag + 3, 1);
    // |y_tc - y_exact| <= eps_tc * (sum_i |w_i| vmax_i + |c|) + abs_tc per row, with
    //   eps_tc: hi/lo split residuals of w and v (2^-22 each, fp16 rounding of the lo parts), the dropped lo*lo term
    //           (2^-22), <= 11 fp32 accumulations in the tensor core (taken as 2^-22 each: covers truncating adders),
    //           the fp32 centring of the landmark (2^-24): 3 + 11 + 1/4 = 14.25 units of 2^-22, taken as 24 (6e-6);
    //   abs_tc: fp16 subnormal spacing of the lo parts (2^-24) times the partner magnitude, folded in below.
    const double vm[3] = {xmax + fabs(cen[0]), xmax + fabs(cen[1]), xmax + fabs(cen[2])};
    const double eps_tc = 24.0 * 2.384185791015625e-07, sub = 5.9604644775390625e-08;
    const double r0 = (fabs(U[0]) * vm[0] + fabs(U[1]) * vm[1] + fabs(U[2]) * vm[2]) + fabs(k0);
    const double r1 = (fabs(U[3]) * vm[1] + fabs(U[4]) * vm[2]) + fabs(k1);
    const double r2 = fabs(U[5]) * vm[2] + fabs(k2);
    const double e0 = eps_tc * r0 + sub * (fabs(U[0]) + fabs(U[1]) + fabs(U[2]) + vm[0] + vm[1] + vm[2] + 1.0);
    const double e1 = eps_tc * r1 + sub * (fabs(U[3]) + fabs(U[4]) + vm[1] + vm[2] + 1.0);
    const double e2 = eps_tc * r2 + sub * (fabs(U[5]) + vm[2] + 1.0);
    const double et = sqrt(e0 * e0 + e1 * e1 + e2 * e2);
    float ef = (float)et;
    if ((double)ef < et) ef = nextafterf(ef, CUDART_INF_F);
    atomicMax(alpha_bits + 1, __float_as_uint(ef));
  }
}

// ---- K5: Bhattacharyya neighbour lists (GMM::GMM, gaussian_mixture.cpp:61-78; BHCoefficient,
// gmm_utils.h:30-52).  One warp per row i, lanes stride over j in ascending order; a ballot-based
// ordered compaction keeps the reference's construction order (ascending component index).  Run twice:
// counting pass (col == nullptr), then the fill pass with the exclusive-scanned row offsets.
__device__ __forceinline__ double bh3(const double* cov0, const double* mu0, double det0, const double* cov1,
                                      const double* mu1, double det1) {
  double cov[9], inv[9];
  for (int i = 0; i < 9; ++i) cov[i] = dmul(dadd(cov0[i], cov1[i]), 0.5);  // (cov0 + cov1) / 2.0
  const double d[3] = {dsub(mu1[0], mu0[0]), dsub(mu1[1], mu0[1]), dsub(mu1[2], mu0[2])};
  inv3(cov, inv);
  const double r0 = dadd(dadd(dmul(d[0], inv[0]), dmul(d[1], inv[3])), dmul(d[2], inv[6]));
  const double r1 = dadd(dadd(dmul(d[0], inv[1]), dmul(d[1], inv[4])), dmul(d[2], inv[7]));
  const double r2 = dadd(dadd(dmul(d[0], inv[2]), dmul(d[1], inv[5])), dmul(d[2], inv[8]));
  const double d0 = dmul(dadd(dadd(dmul(r0, d[0]), dmul(r1, d[1])), dmul(r2, d[2])), 0.125);  // / 8.0
  const double d1 = dmul(log(__ddiv_rn(det3(cov), __dsqrt_rn(dmul(det0, det1)))), 0.5);          // / 2.0
  return dadd(d0, d1);
}

__global__ void __launch_bounds__(256) bh_neighbors_kernel(MapDev m, double thresh, int* counts, const long long* row_ptr,
                                                           int32_t* col, double* dist) {
  const int lane = threadIdx.x & 31;
  const int warp = (blockIdx.x * blockDim.x + threadIdx.x) >> 5;
  const int nwarps = (gridDim.x * blockDim.x) >> 5;
  for (int i = warp; i < m.K; i += nwarps) {
    double cov0[9], mu0[3];
    for (int q = 0; q < 9; ++q) cov0[q] = m.cov[(size_t)i * 9 + q];
    for (int q = 0; q < 3; ++q) mu0[q] = m.mean[(size_t)i * 3 + q];
    const double det0 = m.det[i];
    long long base = (col != nullptr) ? row_ptr[i] : 0;
    int cnt = 0;
    for (int j0 = 0; j0 < m.K; j0 += 32) {
      const int j = j0 + lane;
      bool hit = false;
      double dd = 0.0;
      if (j < m.K && j != i) {
        double cov1[9], mu1[3];
        for (int q = 0; q < 9; ++q) cov1[q] = m.cov[(size_t)j * 9 + q];
        for (int q = 0; q < 3; ++q) mu1[q] = m.mean[(size_t)j * 3 + q];
        dd = bh3(cov0, mu0, det0, cov1, mu1, m.det[j]);
        hit = dd < thresh;  // gaussian_mixture.cpp:68
      }
      const unsigned ballot = __ballot_sync(0xffffffffu, hit);
      if (col != nullptr && hit) {
        const long long pos = base + __popc(ballot & ((1u << lane) - 1u));
        col[pos] = j;
        if (dist != nullptr) dist[pos] = dd;
      }
      const int n = __popc(ballot);
      base += n;
      cnt += n;
    }
    if (col == nullptr && lane == 0) counts[i] = cnt;
  }
}

}  // namespace

cudaError_t launch_bh_neighbors(const MapDev& m, double thresh, int* d_counts, const long long* d_row_ptr, int32_t* d_col,
                                double* d_dist, cudaStream_t stream) {
  const int warps_needed = m.K;
  int blocks = (warps_needed + 7) / 8;
  if (blocks > 148 * 8) blocks = 148 * 8;
  bh_neighbors_kernel<<<blocks, 256, 0, stream>>>(m, thresh, d_counts, d_row_ptr, d_col, d_dist);
  return cudaGetLastError();
}

cudaError_t launch_map_build(const MapDev& m, double deg_eps, double xmax, int* d_err, unsigned int* d_alpha_bits,
                             cudaStream_t stream) {
  const int threads = 128;
  const int ktc = (m.Kpad + kTcTileComps - 1) / kTcTileComps * kTcTileComps;
  const int blocks = (ktc + threads - 1) / threads;
  map_build_kernel<<<blocks, threads, 0, stream>>>(m, deg_eps, xmax, d_err, d_alpha_bits);
  return cudaGetLastError();
}

}  // namespace gmmloc
