// view.cu — the keyframe association pipeline of the reference on the device (fp64, bit-for-bit replay of
// the CPU oracle's expression trees; THIS FILE IS COMPILED WITH -fmad=false):
//
//   K6 render_view_kernel   ≙ GMM::renderView(rot_c_w, t_c_w)            gaussian_mixture.cpp:271-371
//                             + GMMUtility::projectGaussian               gmm_utils.cpp:121-146
//                             + PinholeCamera::project3 (with Jacobian)   pinhole_camera.cpp:68-150
//                             + GaussianComponent2d ctor / decompose      gaussian.h:107-117, gaussian.cpp:17-32
//   K7 search_corr_kernel   ≙ GMM::searchCorrespondence(kpts, comps, num) gaussian_mixture.cpp:484-534
//   K8 check_assoc_kernel   ≙ GMMLoc::checkMapAssociation                 gmmloc_opt.cpp:156-258
//                             (+ Frame::unproject3, frame.cpp:27-36, when a depth array is given)
//
// renderView's occlusion dedupe is inherently sequential in the candidate order (each accepted 2-D component
// changes what later ones are compared with), so K6 runs ONE CTA PER VIEW: the cull / projection of 128
// components at a time is parallel, the surviving candidates are compacted in ascending component order
// (warp ballots), and each candidate's min-Bhattacharyya search over the kept list is a block-wide
// lexicographic (distance, position) argmin — the same winner as the reference's strict-'<' ascending loop.
// Many views (keyframes) are processed concurrently by as many CTAs.
#include <math_constants.h>

#include "common.cuh"
#include "point_gn.cuh"

namespace gmmloc {

namespace {

constexpr int kViewThreads = 128;
constexpr int kViewWarps = kViewThreads / 32;

// Eigen QuaternionBase::_transformVector: uv = q.vec x v; uv += uv; v + w * uv + q.vec x uv
__device__ __forceinline__ void quat_rotate_d(const double* q, const double* v, double* out) {
  double uv[3] = {q[1] * v[2] - q[2] * v[1], q[2] * v[0] - q[0] * v[2], q[0] * v[1] - q[1] * v[0]};
  for (int i = 0; i < 3; ++i) uv[i] += uv[i];
  const double c[3] = {q[1] * uv[2] - q[2] * uv[1], q[2] * uv[0] - q[0] * uv[2], q[0] * uv[1] - q[1] * uv[0]};
  for (int i = 0; i < 3; ++i) out[i] = (v[i] + q[3] * uv[i]) + c[i];
}
__device__ __forceinline__ double det2_d(const double* m) { return m[0] * m[3] - m[2] * m[1]; }
__device__ __forceinline__ void inverse2_d(const double* m, double* out) {
  const double invdet = 1.0 / det2_d(m);
  out[0] = m[3] * invdet;
  out[2] = -m[2] * invdet;
  out[1] = -m[1] * invdet;
  out[3] = m[0] * invdet;
}
__device__ __forceinline__ double quad2_d(const double* A, const double* d) {
  const double r0 = d[0] * A[0] + d[1] * A[2];
  const double r1 = d[0] * A[1] + d[1] * A[3];
  return r0 * d[0] + r1 * d[1];
}

// one 2-D component: mean(2) cov(4) det depth — the layout of the view records (kViewRec doubles)
struct Rec2 {
  double v[kViewRec];
};

// 2-D Bhattacharyya distance, gmm_utils.h:30-52 (g0 = kept entry, g1 = new component)
__device__ __forceinline__ double bh2_d(const double* g0, const double* g1) {
  double cov[4], inv[4];
  for (int i = 0; i < 4; ++i) cov[i] = (g0[2 + i] + g1[2 + i]) / 2.0;
  const double d[2] = {g1[0] - g0[0], g1[1] - g0[1]};
  inverse2_d(cov, inv);
  double d0 = quad2_d(inv, d);
  d0 /= 8.0;
  const double d1 = log(det2_d(cov) / sqrt(g0[6] * g1[6])) / 2.0;
  return d0 + d1;
}

// cull + projectGaussian of component k; true when it survives STEP.0 (view cos), the visibility test and
// the 2-D covariance cull (gaussian_mixture.cpp:283-317)
__device__ __forceinline__ bool project_component(const ViewArgs& a, const double* q, const double* R, const double* t,
                                                  const double* twc, int k, Rec2& g) {
  const double mu[3] = {a.mean[(size_t)k * 3], a.mean[(size_t)k * 3 + 1], a.mean[(size_t)k * 3 + 2]};
  if (a.is_deg[k]) {
    double po[3] = {mu[0] - twc[0], mu[1] - twc[1], mu[2] - twc[2]};
    const double z = (po[0] * po[0] + po[1] * po[1]) + po[2] * po[2];
    if (z > 0.0) {
      const double nrm = sqrt(z);
      for (int i = 0; i < 3; ++i) po[i] /= nrm;
    }
    const double* ax = a.axis + (size_t)k * 9;
    const double view_cos = fabs((po[0] * ax[0] + po[1] * ax[3]) + po[2] * ax[6]);
    if (view_cos < a.view_cos_thresh) return false;
  }
  double pc[3];
  quat_rotate_d(q, mu, pc);
  for (int i = 0; i < 3; ++i) pc[i] += t[i];
  const double x = pc[0], y = pc[1], z = pc[2];
  const double rz = 1.0 / z;
  const double kx = x * rz, ky = y * rz;
  const double rz2 = rz * rz;
  const double J[6] = {a.fx * rz, 0.0, -a.fx * x * rz2, 0.0, a.fy * rz, -a.fy * y * rz2};
  const double u = a.fx * kx + a.cx;
  const double v = a.fy * ky + a.cy;
  const bool vis = u >= 0.0 && v >= 0.0 && u < a.width && v < a.height;
  if (!(vis && z > 0.0)) return false;
  const double* C = a.cov + (size_t)k * 9;
  double T1[6], T2[6], T3[6];
  for (int i = 0; i < 2; ++i)
    for (int j = 0; j < 3; ++j) T1[i * 3 + j] = (J[i * 3] * R[j] + J[i * 3 + 1] * R[3 + j]) + J[i * 3 + 2] * R[6 + j];
  for (int i = 0; i < 2; ++i)
    for (int j = 0; j < 3; ++j) T2[i * 3 + j] = (T1[i * 3] * C[j] + T1[i * 3 + 1] * C[3 + j]) + T1[i * 3 + 2] * C[6 + j];
  for (int i = 0; i < 2; ++i)
    for (int j = 0; j < 3; ++j)
      T3[i * 3 + j] = (T2[i * 3] * R[j * 3] + T2[i * 3 + 1] * R[j * 3 + 1]) + T2[i * 3 + 2] * R[j * 3 + 2];
  double* cov = g.v + 2;
  for (int i = 0; i < 2; ++i)
    for (int j = 0; j < 2; ++j) cov[i * 2 + j] = (T3[i * 3] * J[j * 3] + T3[i * 3 + 1] * J[j * 3 + 1]) + T3[i * 3 + 2] * J[j * 3 + 2];
  g.v[0] = u;
  g.v[1] = v;
  g.v[6] = det2_d(cov);
  g.v[7] = pc[2];
  const double aa = cov[0], bb = cov[2], dd = cov[3];
  const double tr = 0.5 * (aa + dd), h = 0.5 * (aa - dd);
  const double r = sqrt(h * h + bb * bb);
  const double s0 = tr - r, s1 = tr + r;
  if (s0 < 4.0 && s1 < 4.0) return false;  // check_cov_2d, gaussian_mixture.cpp:311-317
  return true;
}

__global__ void __launch_bounds__(kViewThreads) render_view_kernel(const ViewArgs a) {
  const int b = blockIdx.x;
  const int tid = threadIdx.x, lane = tid & 31, warp = tid >> 5;
  __shared__ double s_q[4], s_R[9], s_t[3], s_twc[3];
  __shared__ Rec2 s_cand[kViewThreads];
  __shared__ int s_cand_id[kViewThreads];
  __shared__ int s_warp_cnt[kViewWarps];
  __shared__ double s_red_d[kViewWarps];
  __shared__ int s_red_i[kViewWarps];
  __shared__ int s_V, s_overflow;
  if (tid == 0) {
    const double* pose = a.pose + (size_t)b * 7;
    double q[4];
    quat_normalize_d(pose, q);
    quat_to_rot_d(q, s_R);
    // t_w_c = -(rot_c_w.inverse() * t_c_w), Eigen inverse() = conjugate / squaredNorm   (:288)
    const double n2 = ((q[0] * q[0] + q[1] * q[1]) + q[2] * q[2]) + q[3] * q[3];
    const double qi[4] = {-q[0] / n2, -q[1] / n2, -q[2] / n2, q[3] / n2};
    double twc[3];
    quat_rotate_d(qi, pose + 4, twc);
    for (int i = 0; i < 3; ++i) {
      s_twc[i] = -twc[i];
      s_t[i] = pose[4 + i];
    }
    for (int i = 0; i < 4; ++i) s_q[i] = q[i];
    s_V = 0;
    s_overflow = 0;
  }
  __syncthreads();
  double* wrec = a.wrec + (size_t)b * a.cap * kViewRec;
  int32_t* wids = a.wids + (size_t)b * a.cap;
  for (int k0 = 0; k0 < a.K; k0 += kViewThreads) {
    const int k = k0 + tid;
    Rec2 g;
    const bool pass = (k < a.K) && project_component(a, s_q, s_R, s_t, s_twc, k, g);
    // ordered compaction: candidates keep the ascending component order of the reference's loop
    const unsigned bal = __ballot_sync(0xffffffffu, pass);
    if (lane == 0) s_warp_cnt[warp] = __popc(bal);
    __syncthreads();
    int base = 0, total = 0;
    for (int w = 0; w < kViewWarps; ++w) {
      if (w < warp) base += s_warp_cnt[w];
      total += s_warp_cnt[w];
    }
    if (pass) {
      const int pos = base + __popc(bal & ((1u << lane) - 1u));
      s_cand[pos] = g;
      s_cand_id[pos] = k;
    }
    __syncthreads();
    // STEP. check near far (:327-355), one candidate at a time
    for (int c = 0; c < total; ++c) {
      const double* gc = s_cand[c].v;
      const int V = s_V;
      double md = 1.7976931348623157e308;  // numeric_limits<double>::max()
      int mi = 0x7fffffff;
      for (int j = tid; j < V; j += kViewThreads) {
        const double dist = bh2_d(wrec + (size_t)j * kViewRec, gc);
        if (dist < md) {
          md = dist;
          mi = j;
        }
      }
      for (int off = 16; off > 0; off >>= 1) {
        const double od = __shfl_xor_sync(0xffffffffu, md, off);
        const int oi = __shfl_xor_sync(0xffffffffu, mi, off);
        if (od < md || (od == md && oi < mi)) {
          md = od;
          mi = oi;
        }
      }
      if (lane == 0) {
        s_red_d[warp] = md;
        s_red_i[warp] = mi;
      }
      __syncthreads();
      if (tid == 0) {
        for (int w = 1; w < kViewWarps; ++w)
          if (s_red_d[w] < md || (s_red_d[w] == md && s_red_i[w] < mi)) {
            md = s_red_d[w];
            mi = s_red_i[w];
          }
        bool append = true;
        if (V > 0) {
          if (mi == 0x7fffffff) mi = 0;
          if (md < 0.8) {  // depth_thresh
            append = false;
            if (gc[7] < wrec[(size_t)mi * kViewRec + 7]) {  // the nearer one keeps the slot
              for (int i = 0; i < kViewRec; ++i) wrec[(size_t)mi * kViewRec + i] = gc[i];
              wids[mi] = s_cand_id[c];
            }
          }
        }
        if (append) {
          if (V < a.cap) {
            for (int i = 0; i < kViewRec; ++i) wrec[(size_t)V * kViewRec + i] = gc[i];
            wids[V] = s_cand_id[c];
            s_V = V + 1;
          } else {
            s_overflow = 1;
          }
        }
      }
      __syncthreads();
    }
  }
  // sort by depth descending (:362-364; stable: ties keep insertion order) by rank counting
  const int V = s_V;
  double* rec = a.rec + (size_t)b * a.cap * kViewRec;
  int32_t* ids = a.ids + (size_t)b * a.cap;
  for (int i = tid; i < V; i += kViewThreads) {
    const double di = wrec[(size_t)i * kViewRec + 7];
    int rank = 0;
    for (int j = 0; j < V; ++j) {
      const double dj = wrec[(size_t)j * kViewRec + 7];
      rank += (dj > di || (dj == di && j < i)) ? 1 : 0;
    }
    for (int r = 0; r < kViewRec; ++r) rec[(size_t)rank * kViewRec + r] = wrec[(size_t)i * kViewRec + r];
    ids[rank] = wids[i];
  }
  if (tid == 0) {
    a.count[b] = V;
    a.status[b] = s_overflow ? GMMLOC_VIEW_OVERFLOW : 0;
  }
}

// ---- K7: searchCorrespondence.  grid (B, ceil(N / 128)); one thread per keypoint, the view's 2-D means
// stream through shared memory; per keypoint the `num` nearest means in ascending (distance, position)
// order, filtered by the 2-D Mahalanobis gate 9.0.
constexpr int kSearchTile = 512;
constexpr int kMaxNum = 8;

__global__ void __launch_bounds__(kViewThreads) search_corr_kernel(const SearchArgs a) {
  const int b = blockIdx.x;
  const int i = blockIdx.y * kViewThreads + threadIdx.x;
  const int V = a.count[b];
  const double* rec = a.rec + (size_t)b * a.cap * kViewRec;
  __shared__ double s_u[kSearchTile], s_v[kSearchTile];
  double bd[kMaxNum];
  int bi[kMaxNum];
  int n = 0;
  const int num = a.num;
  double u = 0.0, v = 0.0;
  if (i < a.N) {
    u = (double)a.obs[((size_t)b * a.N + i) * 3];
    v = (double)a.obs[((size_t)b * a.N + i) * 3 + 1];
  }
  double worst = 0.0;
  for (int j0 = 0; j0 < V; j0 += kSearchTile) {
    const int nt = min(kSearchTile, V - j0);
    __syncthreads();
    for (int j = threadIdx.x; j < nt; j += kViewThreads) {
      s_u[j] = rec[(size_t)(j0 + j) * kViewRec];
      s_v[j] = rec[(size_t)(j0 + j) * kViewRec + 1];
    }
    __syncthreads();
    if (i < a.N) {
      for (int j = 0; j < nt; ++j) {
        const double dx = u - s_u[j], dy = v - s_v[j];
        const double d = dx * dx + dy * dy;
        if (n == num && !(d < worst)) continue;
        int pos = (n < num) ? n : num - 1;
        while (pos > 0 && d < bd[pos - 1]) {
          bd[pos] = bd[pos - 1];
          bi[pos] = bi[pos - 1];
          --pos;
        }
        bd[pos] = d;
        bi[pos] = j0 + j;
        if (n < num) ++n;
        worst = bd[n - 1];
      }
    }
  }
  if (i >= a.N) return;
  int32_t* out = a.cand + ((size_t)b * a.N + i) * num;
  int32_t* out_pos = a.cand_pos ? a.cand_pos + ((size_t)b * a.N + i) * num : nullptr;
  const int32_t* ids = a.ids + (size_t)b * a.cap;
  int m = 0;
  for (int r = 0; r < n; ++r) {
    const double* g = rec + (size_t)bi[r] * kViewRec;
    double inv[4];
    inverse2_d(g + 2, inv);  // GaussianComponent2d::cov_inv_, gaussian.h:109
    const double d[2] = {u - g[0], v - g[1]};
    if (quad2_d(inv, d) < 9.0) {  // mdist2_thresh, gaussian_mixture.cpp:518-522
      out[m] = ids[bi[r]];
      if (out_pos) out_pos[m] = bi[r];
      ++m;
    }
  }
  for (; m < num; ++m) {
    out[m] = -1;
    if (out_pos) out_pos[m] = -1;
  }
}

// ---- K8: checkMapAssociation, one thread per feature; the kd-tree fallback (GMM::queryPoint: nearest mean)
// is a warp-cooperative brute-force scan.
__global__ void __launch_bounds__(kViewThreads) check_assoc_kernel(const CheckArgs a) {
  const long long idx = (long long)blockIdx.x * kViewThreads + threadIdx.x;
  const long long total = (long long)a.B * a.N;
  const int lane = threadIdx.x & 31;
  const bool active = idx < total;
  const FrameParamsDev& P = a.prm;
  bool need_nn = false;
  int assoc = -1;
  double pt_init[3] = {0, 0, 0}, pt_out[3] = {0, 0, 0};
  double R[9], t[3] = {0, 0, 0}, o[3] = {0, 0, 0};
  double w_proj = 0, proj_z2 = 0;
  bool write_pt = false;
  if (active) {
    const int b = (int)(idx / a.N);
    const double* pose = a.pose + (size_t)b * 7;
    double q[4];
    quat_normalize_d(pose, q);
    quat_to_rot_d(q, R);
    for (int r = 0; r < 3; ++r) {
      t[r] = pose[4 + r];
      o[r] = (double)a.obs[(size_t)idx * 3 + r];
    }
    bool has_pt = true;
    if (a.depth != nullptr) {
      const float zf = a.depth[idx];
      if (!(zf > 0.0f)) {  // gmmloc_opt.cpp:42,68: no stereo depth, no new map point
        has_pt = false;
        pt_out[0] = pt_out[1] = pt_out[2] = CUDART_NAN;
        write_pt = true;
      } else {
        // Frame::unproject3: camera point, then getTwc().map (SE3Quat::inverse: r = conj, t = r * (t * -1))
        const double z = (double)zf;
        const double pc[3] = {z * (o[0] - P.cx) / P.fx, z * (o[1] - P.cy) / P.fy, z};
        const double qc[4] = {-q[0], -q[1], -q[2], q[3]};
        const double tn[3] = {t[0] * -1.0, t[1] * -1.0, t[2] * -1.0};
        double twc[3], r3[3];
        quat_rotate_d(qc, tn, twc);
        quat_rotate_d(qc, pc, r3);
        for (int r = 0; r < 3; ++r) pt_init[r] = r3[r] + twc[r];
        write_pt = true;
      }
    } else {
      for (int r = 0; r < 3; ++r) pt_init[r] = a.pt3d[(size_t)idx * 3 + r];
    }
    for (int r = 0; r < 3; ++r)
      if (has_pt) pt_out[r] = pt_init[r];
    const int32_t* cand = a.cand + (size_t)idx * a.num;
    int nc = 0;
    for (int c = 0; c < a.num; ++c) nc += (cand[c] >= 0 && cand[c] < a.K) ? 1 : 0;
    if (has_pt && nc > 0) {
      int oc = a.octave[idx];
      if (oc > 7) oc = 7;
      w_proj = (double)P.s2inv[oc];
      double ptc[3];
      quat_rotate_d(q, pt_init, ptc);  // Tcw.map(pt3d)
      double proj_z = ptc[2] + t[2];
      proj_z = proj_z > 1.0 ? 1.0 : proj_z;  // gmmloc_opt.cpp:171
      proj_z2 = proj_z * proj_z;
      const double w_str = P.tri_lambda2 * proj_z2;
      int min_idx = -1;
      double min_value = 1.7976931348623157e308;
      double min_res[3] = {0, 0, 0};
      for (int c = 0; c < a.num; ++c) {
        const int k = cand[c];
        if (k < 0 || k >= a.K) continue;
        const double* ax = a.axis + (size_t)k * 9;
        const double n[3] = {ax[0], ax[3], ax[6]};
        const PointSolve s = optimize_point_d(P, R, t, o, w_proj, w_str, n, a.mean + (size_t)k * 3, pt_init, 5);
        if (s.res && s.chi2_proj < min_value) {  // :192-196
          min_idx = c;
          min_value = s.chi2_proj;
          for (int r = 0; r < 3; ++r) min_res[r] = s.x[r];
        }
      }
      if (min_idx != -1) {
        const int g3d = cand[min_idx];
        double ll = chi2_exact(a.mean + (size_t)g3d * 3, a.cov_inv + (size_t)g3d * 9, min_res[0], min_res[1], min_res[2]);
        int str = g3d;
        for (long long e = a.nbr_row_ptr[g3d]; e < a.nbr_row_ptr[g3d + 1]; ++e) {  // :209-217
          const int nb = a.nbr_col[e];
          const double ln = chi2_exact(a.mean + (size_t)nb * 3, a.cov_inv + (size_t)nb * 9, min_res[0], min_res[1], min_res[2]);
          if (ln < ll) {
            ll = ln;
            str = nb;
          }
        }
        if (str != g3d) {  // :219-228
          const double* ax = a.axis + (size_t)str * 9;
          const double n[3] = {ax[0], ax[3], ax[6]};
          const PointSolve s = optimize_point_d(P, R, t, o, w_proj, w_str, n, a.mean + (size_t)str * 3, pt_init, 5);
          if (s.res) {
            for (int r = 0; r < 3; ++r) min_res[r] = s.x[r];
          } else {
            str = g3d;
            ll = chi2_exact(a.mean + (size_t)g3d * 3, a.cov_inv + (size_t)g3d * 9, min_res[0], min_res[1], min_res[2]);
          }
        }
        if (!(ll > 9.0)) {  // :230-232
          for (int r = 0; r < 3; ++r) pt_out[r] = min_res[r];
          write_pt = true;
          assoc = str;
        }
      } else {
        need_nn = true;  // use_nn3d, :237-256
      }
    }
  }
  // GMM::queryPoint (gaussian_mixture.cpp:545-576): nearest mean, strict '<' ascending => lowest index on ties
  unsigned todo = __ballot_sync(0xffffffffu, need_nn);
  int nn = -1;
  while (todo) {
    const int src = __ffs(todo) - 1;
    todo &= todo - 1;
    const double p0 = __shfl_sync(0xffffffffu, pt_init[0], src), p1 = __shfl_sync(0xffffffffu, pt_init[1], src),
                 p2 = __shfl_sync(0xffffffffu, pt_init[2], src);
    double md = 1.7976931348623157e308;
    int mi = 0x7fffffff;
    for (int k = lane; k < a.K; k += 32) {
      const double d0 = p0 - a.mean[(size_t)k * 3], d1 = p1 - a.mean[(size_t)k * 3 + 1], d2 = p2 - a.mean[(size_t)k * 3 + 2];
      const double d = (d0 * d0 + d1 * d1) + d2 * d2;
      if (d < md) {
        md = d;
        mi = k;
      }
    }
    for (int off = 16; off > 0; off >>= 1) {
      const double od = __shfl_xor_sync(0xffffffffu, md, off);
      const int oi = __shfl_xor_sync(0xffffffffu, mi, off);
      if (od < md || (od == md && oi < mi)) {
        md = od;
        mi = oi;
      }
    }
    if (lane == src) nn = (mi == 0x7fffffff) ? -1 : mi;
  }
  if (need_nn && nn >= 0 && a.is_deg[nn]) {
    const double w_str = P.tri_lambda2 * proj_z2;
    const double* ax = a.axis + (size_t)nn * 9;
    const double n[3] = {ax[0], ax[3], ax[6]};
    const PointSolve s = optimize_point_d(P, R, t, o, w_proj, w_str, n, a.mean + (size_t)nn * 3, pt_init, 5);
    if (s.res) {
      for (int r = 0; r < 3; ++r) pt_out[r] = s.x[r];
      write_pt = true;
    }
  }
  if (active) {
    a.assoc[idx] = assoc;
    if (write_pt)
      for (int r = 0; r < 3; ++r) a.pt3d[(size_t)idx * 3 + r] = pt_out[r];
  }
}

// ---- GMM::queryPoint (gaussian_mixture.cpp:545-576) ------------------------------------------------------------
// One warp per query point: every lane keeps the kQn nearest means of its strided share of the components (sorted
// insertion, (distance, index) lexicographic = lowest index first on equal distances), then kQn rounds of a warp-wide
// lexicographic argmin over the lanes' list heads merge them — the list nanoflann's knnSearch(pt, 5) returns (pinned
// against the reference's own kd-tree in tests/test_oracle_ref_nanoflann.py).  The reference then computes the minimum
// Mahalanobis distance over the list and DISCARDS it: it returns ret_index[0], the Euclid-nearest mean (:574).
constexpr int kQn = 5;
__global__ void __launch_bounds__(128) query_point_kernel(const double* __restrict__ mean, const double* __restrict__ cov_inv, int K,
                                                          const double* __restrict__ pts, int M, int32_t* __restrict__ nearest,
                                                          int32_t* __restrict__ knn, double* __restrict__ min_mdist2) {
  const int lane = threadIdx.x & 31;
  const int q = (int)((blockIdx.x * (size_t)blockDim.x + threadIdx.x) >> 5);
  if (q >= M) return;
  const double p0 = pts[(size_t)q * 3], p1 = pts[(size_t)q * 3 + 1], p2 = pts[(size_t)q * 3 + 2];
  double bd[kQn];
  int bi[kQn];
#pragma unroll
  for (int r = 0; r < kQn; ++r) {
    bd[r] = 1.7976931348623157e308;
    bi[r] = 0x7fffffff;
  }
  for (int k = lane; k < K; k += 32) {
    const double d0 = p0 - mean[(size_t)k * 3], d1 = p1 - mean[(size_t)k * 3 + 1], d2 = p2 - mean[(size_t)k * 3 + 2];
    double d = d0 * d0 + d1 * d1 + d2 * d2;  // FLANNPoints3d::kdtree_distance (gaussian_mixture.h:33-39)
    if (!(d < bd[kQn - 1])) continue;      // k ascends within a lane: an equal distance never displaces an earlier index
    int ki = k;
    bool ins = false;  // once inserted, the rest of the (sorted) list shifts down unconditionally
#pragma unroll
    for (int r = 0; r < kQn; ++r) {
      if (ins || d < bd[r]) {
        ins = true;
        const double td = bd[r];
        const int ti = bi[r];
        bd[r] = d;
        bi[r] = ki;
        d = td;
        ki = ti;
      }
    }
  }
  int out_i[kQn];
#pragma unroll
  for (int r = 0; r < kQn; ++r) {
    double md = bd[0];
    int mi = bi[0];
#pragma unroll
    for (int off = 16; off > 0; off >>= 1) {
      const double od = __shfl_xor_sync(0xffffffffu, md, off);
      const int oi = __shfl_xor_sync(0xffffffffu, mi, off);
      if (od < md || (od == md && oi < mi)) {
        md = od;
        mi = oi;
      }
    }
    out_i[r] = (mi == 0x7fffffff) ? -1 : mi;
    if (mi != 0x7fffffff && bi[0] == mi) {  // the owner pops its head
#pragma unroll
      for (int t = 0; t + 1 < kQn; ++t) {
        bd[t] = bd[t + 1];
        bi[t] = bi[t + 1];
      }
      bd[kQn - 1] = 1.7976931348623157e308;
      bi[kQn - 1] = 0x7fffffff;
    }
  }
  if (lane == 0) {
    nearest[q] = out_i[0];
    double mm = 1.7976931348623157e308;  // numeric_limits<double>::max(), :561
#pragma unroll
    for (int r = 0; r < kQn; ++r) {
      if (knn != nullptr) knn[(size_t)q * kQn + r] = out_i[r];
      if (out_i[r] >= 0) {
        const double dist = chi2_exact(mean + (size_t)out_i[r] * 3, cov_inv + (size_t)out_i[r] * 9, p0, p1, p2);
        if (dist < mm) mm = dist;
      }
    }
    if (min_mdist2 != nullptr) min_mdist2[q] = mm;
  }
}

}  // namespace

cudaError_t launch_render_views(const ViewArgs& a, cudaStream_t stream) {
  if (a.B <= 0) return cudaSuccess;
  render_view_kernel<<<a.B, kViewThreads, 0, stream>>>(a);
  return cudaGetLastError();
}

cudaError_t launch_search_corr(const SearchArgs& a, cudaStream_t stream) {
  if (a.B <= 0 || a.N <= 0) return cudaSuccess;
  const dim3 grid(a.B, (a.N + kViewThreads - 1) / kViewThreads);
  search_corr_kernel<<<grid, kViewThreads, 0, stream>>>(a);
  return cudaGetLastError();
}

int search_corr_max_num() { return kMaxNum; }

cudaError_t launch_query_point(const double* mean, const double* cov_inv, int K, const double* d_pts, int M, int32_t* d_nearest,
                               int32_t* d_knn, double* d_min_mdist2, cudaStream_t stream) {
  if (M <= 0) return cudaSuccess;
  const unsigned blocks = (unsigned)(((long long)M * 32 + 127) / 128);
  query_point_kernel<<<blocks, 128, 0, stream>>>(mean, cov_inv, K, d_pts, M, d_nearest, d_knn, d_min_mdist2);
  return cudaGetLastError();
}
int query_point_knn() { return kQn; }

cudaError_t launch_check_assoc(const CheckArgs& a, cudaStream_t stream) {
  const long long total = (long long)a.B * a.N;
  if (total <= 0) return cudaSuccess;
  const long long blocks = (total + kViewThreads - 1) / kViewThreads;
  check_assoc_kernel<<<(unsigned)blocks, kViewThreads, 0, stream>>>(a);
  return cudaGetLastError();
}

}  // namespace gmmloc
