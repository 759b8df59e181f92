// pose_lm.cu — K6: Tracking::optimizeCurrentPose, exactly as the reference schedules it (one CTA per
// frame, fp64).
//
// Replaces gmmloc/src/modules/tracking_opt.cpp:21-217: pose-only Levenberg-Marquardt on one
// VertexSE3Expmap with one EdgeSE3ProjectXYZOnlyPose (mono, 2-D) or EdgeStereoSE3ProjectXYZOnlyPose
// (stereo, 3-D) per matched feature, information sigma2_inv[octave] * I, Huber(sqrt(5.991) /
// sqrt(7.815)); 4 rounds x 10 LM iterations, every round restarted from the initial pose (:152); after
// each round edges are reclassified by (float)chi2 > 5.991f / 7.815f (current outliers re-evaluated at
// the round's estimate, inliers keep the error of the last LM trial), outliers leave the active set
// (setLevel(1)), the robust kernel is dropped at the end of round 3, early exit when fewer than 10
// edges exist; returns the inlier count.  LM follows g2o's OptimizationAlgorithmLevenberg (SURVEY Appx
// B.7): lambda0 = 1e-5 * max diag(H), trial loop with rho = (chi - chi_new) / (dx.(lambda dx + b) + 1e-3),
// lambda *= max(1/3, min(1 - (2 rho - 1)^3, 2/3)) on success, lambda *= nu, nu *= 2 on failure, at most 10
// trials, iteration loop ends when the trials are exhausted or rho == 0.
//
// The reference rebuilds a g2o graph (heap vertex + N heap edges + robust kernels) per call; here the
// per-edge state is 10 bytes in shared memory and every pass over the edges is one block-wide sweep
// with a fixed-shape (deterministic) reduction.
#include "gn_common.cuh"

namespace gmmloc {

namespace {

constexpr int kT = 128;
constexpr int kW = kT / 32;
constexpr int kAccLM = 28;  // 21 + 6 + robust chi2

__device__ __forceinline__ double block_sum(double v, double* s_red, int lane, int warp) {
#pragma unroll
  for (int off = 16; off > 0; off >>= 1) v += __shfl_xor_sync(0xffffffffu, v, off);
  __syncthreads();  // s_red free
  if (lane == 0) s_red[warp] = v;
  __syncthreads();
  double s = s_red[0];
#pragma unroll
  for (int w = 1; w < kW; ++w) s += s_red[w];
  return s;
}

// un-robustified chi2 = w e^T e of one edge at pose (R, t)
__device__ __forceinline__ double edge_chi2(const FrameParamsDev& P, const double* R, const double* t, const float* X,
                                            const float* ob, bool stereo, double w) {
  const Vec3 pc0 = rot(R, {(double)X[0], (double)X[1], (double)X[2]});
  const double x = pc0.x + t[0], y = pc0.y + t[1], z = pc0.z + t[2];
  const double invz = 1.0 / z;
  const double u = x * invz * P.fx + P.cx, v = y * invz * P.fy + P.cy;
  const double e0 = (double)ob[0] - u, e1 = (double)ob[1] - v;
  double e2 = 0.0;
  if (stereo) e2 = (double)ob[2] - (u - P.bf * invz);
  return w * (e0 * e0 + e1 * e1 + e2 * e2);
}

__device__ __forceinline__ double huber_rho0(double c2, double delta) {
  const double dsqr = delta * delta;
  return (c2 <= dsqr) ? c2 : 2.0 * sqrt(c2) * delta - dsqr;
}

__global__ void __launch_bounds__(kT) pose_lm_kernel(const PoseLmArgs a) {
  extern __shared__ __align__(16) unsigned char smem_raw[];
  __shared__ double s_red[kW][kAccLM];
  __shared__ double s_acc[kAccLM];
  __shared__ double s_sum[kW];
  __shared__ double s_pose[7], s_R[9], s_trial[7], s_Rt[9];
  __shared__ double s_lambda, s_ni, s_cur, s_dx[6];
  __shared__ int s_state;  // 0 = run the trial, 1 = iteration done (continue), 2 = terminate the round
  __shared__ int s_cnt[kW];

  const int f = blockIdx.x, tid = threadIdx.x, lane = tid & 31, warp = tid >> 5;
  const int N = a.N;
  const FrameParamsDev& P = a.prm;
  const float* pts = a.pts_w + (size_t)f * N * 3;
  const float* obs = a.obs + (size_t)f * N * 3;
  const uint8_t* oct = a.octave + (size_t)f * N;
  const uint8_t* valid = a.valid ? a.valid + (size_t)f * N : nullptr;
  double* c2_last;
  uint8_t* st;  // bit0 valid, bit1 stereo, bit2 level (inactive), bit3 is_outlier
  if (a.ws_c2 != nullptr) {
    c2_last = a.ws_c2 + (size_t)f * N;
    st = a.ws_st + (size_t)f * N;
  } else {
    c2_last = reinterpret_cast<double*>(smem_raw);
    st = reinterpret_cast<uint8_t*>(c2_last + N);
  }

  int cnt = 0;
  for (int i = tid; i < N; i += kT) {
    uint8_t b = (!valid || valid[i]) ? 1 : 0;
    if (obs[i * 3 + 2] >= 0.0f) b |= 2;  // u_right < 0 => monocular (:67)
    st[i] = b;
    c2_last[i] = 0.0;
    cnt += b & 1;
  }
  const int E = (int)(block_sum((double)cnt, s_sum, lane, warp) + 0.5);
  if (E < 3) {  // :139-140: nothing optimised, pose untouched
    if (tid < 7) a.pose_out[(size_t)f * 7 + tid] = a.pose0[(size_t)f * 7 + tid];
    if (tid == 0) a.inliers[f] = 0;
    if (a.outlier)
      for (int i = tid; i < N; i += kT) a.outlier[(size_t)f * N + i] = 0;
    return;
  }

  bool robust = true;
  int n_bad = 0;
  for (int round = 0; round < 4; ++round) {
    if (tid == 0) {
      double q[7];
      for (int i = 0; i < 7; ++i) q[i] = a.pose0[(size_t)f * 7 + i];  // setEstimate(getTcw())  :152
      quat_normalize(q);
      for (int i = 0; i < 7; ++i) s_pose[i] = q[i];
      quat_to_rot(q, s_R);
      s_lambda = 0.0;
      s_ni = 2.0;
    }
    __syncthreads();
    bool ok = true;
    for (int it = 0; it < 10 && ok; ++it) {
      // ---- computeActiveErrors + buildSystem at the current estimate ----
      double R[9], t[3];
#pragma unroll
      for (int i = 0; i < 9; ++i) R[i] = s_R[i];
#pragma unroll
      for (int i = 0; i < 3; ++i) t[i] = s_pose[4 + i];
      double acc[kAccLM];
#pragma unroll
      for (int i = 0; i < kAccLM; ++i) acc[i] = 0.0;
      for (int i = tid; i < N; i += kT) {
        const uint8_t b = st[i];
        if (!(b & 1) || (b & 4)) continue;
        const bool stereo = b & 2;
        const double w = (double)P.s2inv[oct[i] & 7];
        Vec3 pc = rot(R, {(double)pts[i * 3], (double)pts[i * 3 + 1], (double)pts[i * 3 + 2]});
        pc.x += t[0]; pc.y += t[1]; pc.z += t[2];
        const double c2 = edge_chi2(P, R, t, pts + i * 3, obs + i * 3, stereo, w);
        c2_last[i] = c2;
        acc[27] += robust ? huber_rho0(c2, stereo ? P.delta_stereo : P.delta_mono) : c2;
        Sym3 G;
        Vec3 h;
        reproj(P, pc, obs[i * 3], obs[i * 3 + 1], obs[i * 3 + 2], stereo, w, robust, G, h);
        accumulate_pose(acc, pc, G, h);
      }
#pragma unroll
      for (int i = 0; i < kAccLM; ++i) {
        double s = acc[i];
#pragma unroll
        for (int off = 16; off > 0; off >>= 1) s += __shfl_xor_sync(0xffffffffu, s, off);
        if (lane == 0) s_red[warp][i] = s;
      }
      __syncthreads();
      if (tid < kAccLM) {
        double s = s_red[0][tid];
#pragma unroll
        for (int w = 1; w < kW; ++w) s += s_red[w][tid];
        s_acc[tid] = s;
      }
      __syncthreads();
      if (tid == 0) {
        s_cur = s_acc[27];
        if (it == 0) {  // computeLambdaInit: tau * max diagonal
          const double md = fmax(fmax(fmax(fabs(s_acc[0]), fabs(s_acc[6])), fmax(fabs(s_acc[11]), fabs(s_acc[15]))),
                                 fmax(fabs(s_acc[18]), fabs(s_acc[20])));
          s_lambda = 1e-5 * md;
          s_ni = 2.0;
        }
      }
      // ---- trial loop ----
      int qmax = 0;
      double rho = 0.0;
      do {
        __syncthreads();
        if (tid == 0) {
          double Hl[27];
          for (int i = 0; i < 27; ++i) Hl[i] = s_acc[i];
          const double lam = s_lambda;
          Hl[0] += lam; Hl[6] += lam; Hl[11] += lam; Hl[15] += lam; Hl[18] += lam; Hl[20] += lam;
          double dx[6];
          const bool ok2 = solve6(Hl, dx);
          if (ok2) {
            se3_oplus(dx, s_pose, s_trial);
            quat_to_rot(s_trial, s_Rt);
            s_state = 0;
          } else {
            for (int i = 0; i < 6; ++i) dx[i] = 0.0;
            s_state = 3;  // linear solve failed: tempChi = max
          }
          for (int i = 0; i < 6; ++i) s_dx[i] = dx[i];
        }
        __syncthreads();
        double tmp = 1.7976931348623157e308;
        if (s_state == 0) {  // computeActiveErrors at the trial estimate
          double Rt[9], tt[3];
#pragma unroll
          for (int i = 0; i < 9; ++i) Rt[i] = s_Rt[i];
#pragma unroll
          for (int i = 0; i < 3; ++i) tt[i] = s_trial[4 + i];
          double chi = 0.0;
          for (int i = tid; i < N; i += kT) {
            const uint8_t b = st[i];
            if (!(b & 1) || (b & 4)) continue;
            const bool stereo = b & 2;
            const double c2 = edge_chi2(P, Rt, tt, pts + i * 3, obs + i * 3, stereo, (double)P.s2inv[oct[i] & 7]);
            c2_last[i] = c2;
            chi += robust ? huber_rho0(c2, stereo ? P.delta_stereo : P.delta_mono) : c2;
          }
          tmp = block_sum(chi, s_sum, lane, warp);
        }
        // every thread evaluates the (cheap) acceptance test on identical shared values
        const double lam = s_lambda;
        double scale = 0.0;
        for (int i = 0; i < 6; ++i) scale += s_dx[i] * (lam * s_dx[i] + s_acc[21 + i]);
        scale += 1e-3;
        rho = (s_cur - tmp) / scale;
        const bool good = rho > 0.0 && isfinite(tmp) && s_state == 0;
        __syncthreads();  // all reads of s_lambda / s_cur / s_state done
        if (tid == 0) {
          if (good) {
            double alpha = 1.0 - pow(2.0 * rho - 1.0, 3.0);
            alpha = fmin(alpha, 2.0 / 3.0);
            s_lambda = lam * fmax(1.0 / 3.0, alpha);
            s_ni = 2.0;
            s_cur = tmp;
            for (int i = 0; i < 7; ++i) s_pose[i] = s_trial[i];
            for (int i = 0; i < 9; ++i) s_R[i] = s_Rt[i];
          } else {
            s_lambda = lam * s_ni;
            s_ni *= 2.0;
          }
        }
        ++qmax;
        __syncthreads();
        if (!isfinite(s_lambda)) break;
      } while (rho < 0.0 && qmax < 10);
      if (qmax == 10 || rho == 0.0 || !isfinite(s_lambda)) ok = false;  // OptimizationAlgorithm::Terminate
      __syncthreads();
    }
    // ---- classification (:157-203) ----
    {
      double R[9], t[3];
#pragma unroll
      for (int i = 0; i < 9; ++i) R[i] = s_R[i];
#pragma unroll
      for (int i = 0; i < 3; ++i) t[i] = s_pose[4 + i];
      int bad = 0;
      for (int i = tid; i < N; i += kT) {
        uint8_t b = st[i];
        if (!(b & 1)) continue;
        const bool stereo = b & 2;
        double c2 = c2_last[i];
        if (b & 8) {  // current outlier: e->computeError() at the round's estimate
          c2 = edge_chi2(P, R, t, pts + i * 3, obs + i * 3, stereo, (double)P.s2inv[oct[i] & 7]);
          c2_last[i] = c2;
        }
        const float chi2 = (float)c2;
        if (chi2 > (stereo ? 7.815f : 5.991f)) {
          b |= 12;  // outlier, level 1
          ++bad;
        } else {
          b &= ~12;
        }
        st[i] = b;
      }
#pragma unroll
      for (int off = 16; off > 0; off >>= 1) bad += __shfl_xor_sync(0xffffffffu, bad, off);
      __syncthreads();
      if (lane == 0) s_cnt[warp] = bad;
      __syncthreads();
      n_bad = 0;
      for (int w = 0; w < kW; ++w) n_bad += s_cnt[w];
    }
    if (round == 2) robust = false;  // e->setRobustKernel(0)
    if (E < 10) break;               // optimizer.edges().size() < 10
    __syncthreads();
  }
  __syncthreads();
  if (tid < 7) a.pose_out[(size_t)f * 7 + tid] = s_pose[tid];
  if (tid == 0) a.inliers[f] = E - n_bad;
  if (a.outlier)
    for (int i = tid; i < N; i += kT) a.outlier[(size_t)f * N + i] = (st[i] & 8) ? 1 : 0;
}

}  // namespace

cudaError_t launch_pose_lm(const PoseLmArgs& a, cudaStream_t stream) {
  if (a.B <= 0) return cudaSuccess;
  size_t smem = 0;
  if (a.ws_c2 == nullptr) {
    smem = (size_t)a.N * 9 + 16;
    cudaError_t e = cudaFuncSetAttribute(pose_lm_kernel, cudaFuncAttributeMaxDynamicSharedMemorySize, 208 * 1024);
    if (e != cudaSuccess) return e;
  }
  pose_lm_kernel<<<a.B, kT, smem, stream>>>(a);
  return cudaGetLastError();
}

size_t pose_lm_smem_limit_points() { return (size_t)(200 * 1024) / 9; }

}  // namespace gmmloc
