// frame_gn_world.cu — K3w: the mixed-precision frame Gauss-Newton with the per-point blocks written in the WORLD frame.
//
// System, reference citations and precision split: see frame_gn.cu (formulation) and frame_gn_mixed.cu (K3m, the
// camera-frame mixed-precision kernel this one replaces as the default); the algebra is in frame_gn_world.cuh.
// What the world frame buys, besides ~15 % fewer instructions per structured point:
//
//   * the step of a point, dx = u0 - W a with a = omega_w x x' + upsilon_w, needs neither the pose of the linearisation
//     nor the class of the point, so every sweep splits into an UPDATE pass (15 FMAs per point on the stored (u0, W);
//     the loads of consecutive points are independent, the loop is unrolled and the L2 round trip of the stored
//     linearisation is overlapped instead of sitting at the head of every point as in K3m) and a LINEARISE pass;
//   * the linearise pass works at ONE pose (K3m kept two, old and new, in both precisions: 72 values the compiler tried to
//     keep in registers), free stereo points without structure edge and monocular plane points store their closed-form
//     (u0, W) like everybody else, and the kernel fits 128 registers: 2 CTAs x 256 threads = 16 warps per SM
//     (K3m: 168 registers, 12 warps);
//   * one class branch per point on class-sorted data (warp-uniform except at the six class boundaries of a frame).
//
// Frames arrive classified and class-sorted by frame_prep_kernel (K3p, frame_gn_mixed.cu); the kernel is persistent
// (frames handed out through an atomic counter) and deterministic (fixed-order reductions, no floating-point atomics).
#include <math_constants.h>

#include <algorithm>
#include <cstddef>
#include <cstdint>

#include "common.cuh"
#include "gn_common.cuh"
#include "frame_gn_point.cuh"
#include "frame_gn_classes.cuh"
#include "frame_gn_world.cuh"

namespace gmmloc {

namespace {

using gnw::Cam;
using gnw::Lin;
using gnw::PointLin;
using gnw::S3f;
using gnw::V3f;

// threads of the two-resident-frames variant: 256 (128 registers, 16 warps per SM), 224 (144, 14) or 192 (168, 12)
#ifndef GMMLOC_GNW_THREADS2
#define GMMLOC_GNW_THREADS2 256
#endif
constexpr int kWThreads2 = GMMLOC_GNW_THREADS2;
#ifndef GMMLOC_GNW_PREFETCH
#define GMMLOC_GNW_PREFETCH 0
#endif
#ifndef GMMLOC_GNW_CARVEOUT
#define GMMLOC_GNW_CARVEOUT 0  // preferred shared-memory carve-out in percent of the maximum (0: the driver's choice)
#endif
constexpr int kAccW = 27;
constexpr int kStoreW = 12;  // floats of the stored linearisation per point: u0 (3), W (9)
__host__ __device__ inline size_t world_slice(int N) { return ((size_t)N * (4 * kStoreW + 8) + 15) & ~(size_t)15; }

// pose of the linearisation point (fp64 + fp32 copy of R) and of the state before it (the fp64 per-point path updates
// at the old pose with the camera-frame step)
struct __align__(16) PoseW {
  double R[9], t[3];
  float Rf[12];  // rows padded to 16 B (read as three float4: the struct is 16-byte aligned, Rf sits at offset 96)
};
static_assert(offsetof(PoseW, Rf) % 16 == 0, "PoseW::Rf must be float4-aligned");

// fp64 treatment of one ill-conditioned point (class C_F64), camera frame, exactly as in K3m: update at the old pose,
// Schur terms at the new one.  Not inlined: a few warp-iterations per sweep that must not cost the fp32 path registers.
__device__ __noinline__ void point_fp64_w(const FrameArgs& a, const double* Ro, const double* to, const double* Rn, const double* tn,
                                          const double* dxi_c, uint8_t b, int k, bool upd, bool accp, const float* ob, double w,
                                          double* x, V3f& pcf, S3f& Mf, V3f& vf, int& bad) {
  StrConst sc;
  if (b & F_STR) load_structure(a, k, b, sc);
  Vec3 xw = {x[0], x[1], x[2]};
  if (upd) {
    const Vec3 om = {dxi_c[0], dxi_c[1], dxi_c[2]}, up = {dxi_c[3], dxi_c[4], dxi_c[5]};
    if (!point_update(a, Ro, to, sc, b, om, up, ob, w, xw)) bad |= 2;
    x[0] = xw.x; x[1] = xw.y; x[2] = xw.z;
  }
  if (accp) {
    Vec3 pc, v;
    Sym3 M;
    if (!pose_terms(a, Rn, tn, sc, b, xw, ob, w, pc, M, v)) bad |= 1;
    pcf = {(float)pc.x, (float)pc.y, (float)pc.z};
    Mf = {(float)M.xx, (float)M.xy, (float)M.xz, (float)M.yy, (float)M.yz, (float)M.zz};
    vf = {(float)v.x, (float)v.y, (float)v.z};
  }
}

// 27 per-lane partial sums -> lane l holds the warp total of value l (l < 27).  Recursive halving: at offset 16, 8, 4, 2, 1
// every lane keeps the half of its values that its partner does not and receives the partner's copy of that half:
// 16 + 8 + 4 + 2 + 1 = 31 shuffles instead of 27 x 5 for a butterfly per value; the order of the additions is fixed.
__device__ __forceinline__ float warp_reduce27(float* acc, int lane) {
  float a[32];
#pragma unroll
  for (int i = 0; i < 32; ++i) a[i] = i < kAccW ? acc[i] : 0.f;
#pragma unroll
  for (int h = 16; h >= 1; h >>= 1) {
    const bool up = (lane & h) != 0;
#pragma unroll
    for (int i = 0; i < h; ++i) {
      const float send = up ? a[i] : a[i + h];
      const float keep = up ? a[i + h] : a[i];
      a[i] = keep + __shfl_xor_sync(0xffffffffu, send, h);
    }
  }
  return a[0];
}

// the stored linearisation lives in shared memory or in this CTA's slice of the global scratch.  The scratch is written
// and read once per sweep (2000 x 48 B per frame: more than the L1 that is left next to two resident frames), so it
// bypasses L1 (ld.global.cg / st.global.cg): L1 keeps the component constants the linearise pass gathers every sweep.
__device__ __forceinline__ float4 ld_store(const float* p, bool in_smem) {
  return in_smem ? *reinterpret_cast<const float4*>(p) : __ldcg(reinterpret_cast<const float4*>(p));
}
__device__ __forceinline__ void st_store(float* p, const float4& v, bool in_smem) {
  if (in_smem) *reinterpret_cast<float4*>(p) = v;
  else __stcg(reinterpret_cast<float4*>(p), v);
}

template <int kThreads, int kMinB>
__global__ void __launch_bounds__(kThreads, kMinB) frame_gn_world_kernel(const FrameArgs a, const float* __restrict__ gn_tab,
                                                                          int* __restrict__ counters, float* __restrict__ store_ws,
                                                                          int aux_global, const unsigned char* __restrict__ prep) {
  constexpr int kWarps = kThreads / 32;
  extern __shared__ __align__(16) unsigned char smem_raw[];
  __shared__ double s_red[kWarps][kAccW];
  __shared__ double s_acc[kAccW];
  __shared__ double s_pose[7], s_dxc[6], s_c0[3];
  __shared__ float s_dxw[8];  // omega_w (3), upsilon_w (3): the pose step in world-twist coordinates about c0
  __shared__ PoseW s_new, s_old;
  __shared__ int s_fail1, s_fail2, s_frame, s_n64;
  __shared__ float s_w[8];  // frame::sigma2_inv per octave (a dynamically indexed kernel parameter would live in local memory)

  const int tid = threadIdx.x, lane = tid & 31, warp = tid >> 5;
  const int N = a.N;
  const FrameParamsDev& P = a.prm;
  const bool pose_free = P.factor_mode != GMMLOC_MODE_POINTS_ONLY;
  const bool pts_free = P.factor_mode != GMMLOC_MODE_POSE_ONLY;
  Cam C;
  C.fx = P.fx; C.fy = P.fy; C.cx = P.cx; C.cy = P.cy; C.bf = P.bf;
  C.fxf = (float)P.fx; C.fyf = (float)P.fy; C.bff = (float)P.bf; C.lam = (float)P.ba_lambda2;
  C.d_mono = (float)P.delta_mono; C.d_stereo = (float)P.delta_stereo; C.robust = P.robust;

  if (tid < 8) s_w[tid] = P.s2inv[tid];
  // per-frame state, class-sorted position j: landmark (fp64), observation, octave, flags; component id and original
  // index behind them, or — frames too large for that — behind the stored linearisation in this CTA's global slice
  double* xs = reinterpret_cast<double*>(smem_raw);
  float* so = reinterpret_cast<float*>(xs + (size_t)N * 3);
  uint8_t* soct = reinterpret_cast<uint8_t*>(so + (size_t)N * 3);
  uint8_t* fl = soct + N;
  const size_t slice = world_slice(N);
  const size_t base_bytes = ((size_t)N * 38 + 15) & ~(size_t)15;
  unsigned char* aux = aux_global ? reinterpret_cast<unsigned char*>(store_ws) + (size_t)blockIdx.x * slice + (size_t)N * 4 * kStoreW
                                  : smem_raw + base_bytes;
  int32_t* sk = reinterpret_cast<int32_t*>(aux);
  uint16_t* perm = reinterpret_cast<uint16_t*>(sk + N);
  const bool store_in_smem = store_ws == nullptr;
  float* store = store_ws != nullptr
                     ? reinterpret_cast<float*>(reinterpret_cast<unsigned char*>(store_ws) + (size_t)blockIdx.x * slice)
                     : reinterpret_cast<float*>(smem_raw + base_bytes + (((size_t)N * 6 + 15) & ~(size_t)15));

  for (;;) {
    __syncthreads();
    if (tid == 0) s_frame = atomicAdd(counters, 1);
    __syncthreads();
    const int f = s_frame;
    if (f >= a.B) break;

    const float* pts = a.pts_w + (size_t)f * N * 3;
    if (tid == 0) {
      double q[7];
      for (int i = 0; i < 7; ++i) q[i] = a.pose0[(size_t)f * 7 + i];
      quat_normalize(q);
      for (int i = 0; i < 7; ++i) s_pose[i] = q[i];
      quat_to_rot(q, s_new.R);
      for (int i = 0; i < 6; ++i) {
        s_dxc[i] = 0.0;
        s_dxw[i] = 0.f;
      }
      for (int r = 0; r < 3; ++r) {
        for (int c = 0; c < 3; ++c) {
          s_new.Rf[r * 4 + c] = (float)s_new.R[r * 3 + c];
          s_old.R[r * 3 + c] = s_new.R[r * 3 + c];
        }
        s_new.Rf[r * 4 + 3] = 0.f;
        s_new.t[r] = s_old.t[r] = q[4 + r];
      }
      // reference point of the world twist: the camera centre of the initial pose, c0 = -R^T t
      for (int c = 0; c < 3; ++c) s_c0[c] = -(s_new.R[c] * q[4] + s_new.R[3 + c] * q[5] + s_new.R[6 + c] * q[6]);
      s_fail1 = 0;
      s_fail2 = 0;
      s_n64 = 0;
    }
    __syncthreads();
    {
      // the frame as frame_prep_kernel left it: a plain coalesced copy, every load independent of the others
      const unsigned char* pf = prep + (size_t)f * prep_stride(N);
      const float* gp = reinterpret_cast<const float*>(pf);
      const float* go = reinterpret_cast<const float*>(pf + (size_t)N * 12);
      const int32_t* gk = reinterpret_cast<const int32_t*>(pf + (size_t)N * 24);
      const uint16_t* gperm = reinterpret_cast<const uint16_t*>(pf + (size_t)N * 28);
      const uint8_t* goct = pf + (size_t)N * 30;
      const uint8_t* gfl = pf + (size_t)N * 31;
#pragma unroll 4
      for (int e = tid; e < 3 * N; e += kThreads) {
        xs[e] = (double)gp[e];
        so[e] = go[e];
      }
      int n64 = 0;
#pragma unroll 4
      for (int j = tid; j < N; j += kThreads) {
        sk[j] = gk[j];
        perm[j] = gperm[j];
        soct[j] = goct[j];
        const uint8_t b = gfl[j];
        fl[j] = b;
        n64 += (b >> 4) >= C_MSTR;  // monocular + full Gaussian: fp64 path as in K3m (ill-conditioned in fp32 near the camera)
      }
#pragma unroll
      for (int off = 16; off > 0; off >>= 1) n64 += __shfl_xor_sync(0xffffffffu, n64, off);
      if (lane == 0 && n64) atomicAdd(&s_n64, n64);
    }
    __syncthreads();
    // C_MSTR and C_F64 are the last classes of the sorted frame: their points sit at [f64_begin, N).  They take the fp64 per-point path
    // (~8x the cost of a structured fp32 point), so they are dealt to the warps round robin — a warp that owned a whole
    // chunk of them kept the other seven waiting at the sweep barrier (21 % of the stall samples, profiles/frame_gn_world_r02.txt)
    const int n64 = s_n64, f64_begin = N - n64;

    const int iters = P.gn_iters;
    const int sweeps = (iters > 0) ? iters + 1 : 0;
    for (int sw = 0; sw < sweeps; ++sw) {
      const bool do_update = pts_free && sw > 0;
      const bool do_lin = sw < iters;           // state sw is linearised (the update of sweep sw + 1 needs it)
      const bool do_acc = pose_free && do_lin;  // ... and enters the pose system
      if (!do_update && !do_acc && !(do_lin && pts_free)) continue;
      int bad = 0;  // bit 0: singular block while linearising, bit 1: non-finite step while updating
      // read through volatile views inside the loops: the pose (12 doubles + 9 floats) and the reference point are then
      // loaded where they are used (broadcast LDS) instead of being carried in ~40 registers across every point
      const volatile double* vc0 = s_c0;
      const volatile double* vR = s_new.R;
      const volatile double* vt = s_new.t;

      // ---- update pass: x += u0 - W (omega_w x x' + upsilon_w) from the linearisation stored by the previous sweep ----
      if (do_update) {
        const V3f om = {s_dxw[0], s_dxw[1], s_dxw[2]}, up = {s_dxw[3], s_dxw[4], s_dxw[5]};
#pragma unroll 2
        for (int j = tid; j < f64_begin; j += kThreads) {
          const int cls = fl[j] >> 4;
          if (cls < C_PLAIN) continue;
          const float* sp = store + (size_t)j * kStoreW;
          const float4 s0 = ld_store(sp, store_in_smem);
          const double x0 = xs[j * 3], x1 = xs[j * 3 + 1], x2 = xs[j * 3 + 2];
          const V3f xr = {(float)(x0 - vc0[0]), (float)(x1 - vc0[1]), (float)(x2 - vc0[2])};
          const V3f av = gnw::twist_at(om, up, xr);
          V3f dx;
          if (cls == C_PLAIN) {
            dx = {s0.x - av.x, s0.y - av.y, s0.z - av.z};
          } else {
            const float4 s1 = ld_store(sp + 4, store_in_smem), s2 = ld_store(sp + 8, store_in_smem);
            dx = {s0.x - (s0.w * av.x + s1.x * av.y + s1.y * av.z), s0.y - (s1.z * av.x + s1.w * av.y + s2.x * av.z),
                  s0.z - (s2.y * av.x + s2.z * av.y + s2.w * av.z)};
          }
          if (!gnw::finite3(dx)) bad |= 2;
          xs[j * 3] = x0 + (double)dx.x;
          xs[j * 3 + 1] = x1 + (double)dx.y;
          xs[j * 3 + 2] = x2 + (double)dx.z;
        }
      }

      // ---- linearise pass at the current pose: store (u0, W) of every free point, accumulate the reduced pose system ----
      float acc[kAccW];
#pragma unroll
      for (int i = 0; i < kAccW; ++i) acc[i] = 0.f;
      // fp64 points, dealt round robin over the warps (point q of the class goes to lane q / kWarps of warp q % kWarps)
      for (int q = lane * kWarps + warp; q < n64; q += 32 * kWarps) {
        const int j = f64_begin + q;
        const uint8_t b = fl[j];
        const bool upd = do_update, accp = do_acc && (b & F_STR);
        if (!upd && !accp) continue;
        double x[3] = {xs[j * 3], xs[j * 3 + 1], xs[j * 3 + 2]};
        const float ob[3] = {so[j * 3], so[j * 3 + 1], so[j * 3 + 2]};
        V3f pcf, vf;
        S3f Mf;
        point_fp64_w(a, s_old.R, s_old.t, s_new.R, s_new.t, s_dxc, b & 15, sk[j], upd, accp, ob, (double)s_w[soct[j] & 7], x, pcf,
                     Mf, vf, bad);
        if (upd) {
          xs[j * 3] = x[0];
          xs[j * 3 + 1] = x[1];
          xs[j * 3 + 2] = x[2];
        }
        if (accp) {
          float Rl[9];
#pragma unroll
          for (int r = 0; r < 3; ++r)
#pragma unroll
            for (int c = 0; c < 3; ++c) Rl[r * 3 + c] = s_new.Rf[r * 4 + c];
          S3f Mw;
          V3f vw;
          gnw::schur_to_world(Rl, Mf, vf, Mw, vw);
          gnw::accumulate_pose(acc, {(float)(x[0] - vc0[0]), (float)(x[1] - vc0[1]), (float)(x[2] - vc0[2])}, Mw, vw);
        }
      }
      for (int j = tid; j < f64_begin && do_lin; j += kThreads) {
#ifndef GMMLOC_GNW_NO_RELOAD
        // compiler barrier: the pose (R, t in fp64 + R in fp32), the reference point and the camera constants are re-read
        // from shared / constant memory for every point instead of being kept in ~60 registers across the loop, which at
        // the 128-register cap (2 x 256 threads per SM) pushed the accumulators of the pose system into local memory
        asm volatile("" ::: "memory");
#endif
        const uint8_t b = fl[j];
        const int cls = b >> 4;
#if GMMLOC_GNW_PREFETCH
        // the component constants of this thread's NEXT point start their trip from L2 now (no register is tied up: the
        // lines land in L1, where the gathers of the next iteration find them)
        if (j + kThreads < f64_begin) {
          const int kn = sk[j + kThreads];
          if (kn >= 0) {
            asm volatile("prefetch.global.L1 [%0];" ::"l"(a.mean + (size_t)kn * 3));
            asm volatile("prefetch.global.L1 [%0];" ::"l"(gn_tab + (size_t)kn * 8));
          }
        }
#endif
        if (cls == C_OFF) continue;
        if (cls == C_FIXED && !do_acc) continue;
        const double x0 = xs[j * 3], x1 = xs[j * 3 + 1], x2 = xs[j * 3 + 2];
        const float ob[3] = {so[j * 3], so[j * 3 + 1], so[j * 3 + 2]};
        const float w = s_w[soct[j] & 7];
        const bool stereo = b & F_STEREO;
        const V3f xr = {(float)(x0 - vc0[0]), (float)(x1 - vc0[1]), (float)(x2 - vc0[2])};  // x' = x - c0
        Lin L;
        {
          double Rq[9], tq[3];
#pragma unroll
          for (int i = 0; i < 9; ++i) Rq[i] = vR[i];
#pragma unroll
          for (int i = 0; i < 3; ++i) tq[i] = vt[i];
          L = gnw::linearise(C, Rq, tq, x0, x1, x2, ob, stereo);
        }
        float Rl[9];
        {
          const volatile float* vRf = s_new.Rf;
#pragma unroll
          for (int r = 0; r < 3; ++r)
#pragma unroll
            for (int c = 0; c < 3; ++c) Rl[r * 3 + c] = vRf[r * 4 + c];
        }
        float* sp = store + (size_t)j * kStoreW;
        if (cls == C_PLAIN) {
          const V3f u0 = gnw::lin_plain(C, L, Rl);
          st_store(sp, make_float4(u0.x, u0.y, u0.z, 0.f), store_in_smem);
          continue;
        }
        if (cls == C_FIXED) {
          S3f G;
          V3f h;
          gnw::lin_fixed(C, L, Rl, stereo, w, G, h);
          gnw::accumulate_pose(acc, xr, G, h);
          continue;
        }
        // structured classes: the component constants (fp64 mean, fp32 structure table) are gathered from L2
        const int k = sk[j];
        const double* mu = a.mean + (size_t)k * 3;
        const V3f d = {(float)(x0 - mu[0]), (float)(x1 - mu[1]), (float)(x2 - mu[2])};
        const float4 t0 = __ldg(reinterpret_cast<const float4*>(gn_tab) + 2 * (size_t)k);
        PointLin pl;
        if (cls == C_STR || cls == C_MSTR) {
          const float2 t1 = __ldg(reinterpret_cast<const float2*>(gn_tab) + 4 * (size_t)k + 2);
          const S3f Sw = {t0.x, t0.y, t0.z, t0.w, t1.x, t1.y};
          S3f M;
          V3f v;
          if (!gnw::lin_str(C, L, Rl, stereo, w, Sw, d, pl, do_acc, M, v)) bad |= 1;
          if (do_acc) gnw::accumulate_pose(acc, xr, M, v);
        } else {
          const V3f n = {t0.x, t0.y, t0.z};
          const float es = gnw::dotf(n, d);
          if (cls == C_STRP) {
            float kappa, rho;
            if (!gnw::lin_strp(C, L, Rl, w, n, es, pl, kappa, rho)) bad |= 1;
            if (do_acc) gnw::accumulate_pose_rank1(acc, xr, n, kappa, rho);
          } else {  // C_MPLANE
            if (!gnw::lin_mplane(C, L, Rl, n, es, pl)) bad |= 2;
          }
        }
        st_store(sp, make_float4(pl.u0.x, pl.u0.y, pl.u0.z, pl.w0), store_in_smem);
        st_store(sp + 4, make_float4(pl.w1, pl.w2, pl.w3, pl.w4), store_in_smem);
        st_store(sp + 8, make_float4(pl.w5, pl.w6, pl.w7, pl.w8), store_in_smem);
      }
      if (bad & 2) s_fail2 = 1;
      if (bad & 1) s_fail1 = 1;
      if (do_acc) {  // lanes in fp32 (fixed halving order), warps in fp64
        const float tot = warp_reduce27(acc, lane);
        if (lane < kAccW) s_red[warp][lane] = (double)tot;
      }
      __syncthreads();
      if (s_fail2) break;  // uniform
      if (do_acc) {
        if (warp == 0) {  // fixed-order cross-warp sum, then the 6x6 solve and the SE3 update on lane 0
          if (lane < kAccW) {
            double s = s_red[0][lane];
#pragma unroll
            for (int w2 = 1; w2 < kWarps; ++w2) s += s_red[w2][lane];
            s_acc[lane] = s;
          }
          __syncwarp();
          if (lane == 0) {
            double dw[6];
            bool okf = !s_fail1;
#pragma unroll
            for (int i = 0; i < kAccW; ++i) okf = okf && isfinite(s_acc[i]);
            if (!okf || !solve6_ldlt(s_acc, dw)) {
              s_fail1 = 1;
            } else {
              // world twist about c0 -> camera twist at the linearisation pose (adjoint), then T <- exp(xi_c) T as the
              // reference applies it:  omega_c = R omega_w,  upsilon_c = R upsilon_w + t' x omega_c,  t' = t + R c0
              const double* R = s_new.R;
              double tp[3], dc[6];
#pragma unroll
              for (int r = 0; r < 3; ++r) {
                tp[r] = s_new.t[r] + (R[r * 3] * s_c0[0] + R[r * 3 + 1] * s_c0[1] + R[r * 3 + 2] * s_c0[2]);
                dc[r] = R[r * 3] * dw[0] + R[r * 3 + 1] * dw[1] + R[r * 3 + 2] * dw[2];
                dc[3 + r] = R[r * 3] * dw[3] + R[r * 3 + 1] * dw[4] + R[r * 3 + 2] * dw[5];
              }
              dc[3] += tp[1] * dc[2] - tp[2] * dc[1];
              dc[4] += tp[2] * dc[0] - tp[0] * dc[2];
              dc[5] += tp[0] * dc[1] - tp[1] * dc[0];
              double np[7];
#pragma unroll
              for (int i = 0; i < 9; ++i) s_old.R[i] = s_new.R[i];
#pragma unroll
              for (int i = 0; i < 3; ++i) s_old.t[i] = s_new.t[i];
              se3_oplus_fast(dc, s_pose, np, s_new.R);
#pragma unroll
              for (int i = 0; i < 6; ++i) {
                s_dxc[i] = dc[i];
                s_dxw[i] = (float)dw[i];
              }
#pragma unroll
              for (int i = 0; i < 7; ++i) s_pose[i] = np[i];
#pragma unroll
              for (int r = 0; r < 3; ++r) {
#pragma unroll
                for (int c = 0; c < 3; ++c) s_new.Rf[r * 4 + c] = (float)s_new.R[r * 3 + c];
                s_new.t[r] = np[4 + r];
              }
            }
          }
        }
        __syncthreads();
        if (s_fail1) break;  // uniform
      }
    }
    __syncthreads();
    const bool fail = s_fail1 || s_fail2;
    __syncthreads();
    if (fail && tid == 0) {  // unsolved frame: back to the normalised input pose
      double q[7];
      for (int i = 0; i < 7; ++i) q[i] = a.pose0[(size_t)f * 7 + i];
      quat_normalize(q);
      for (int i = 0; i < 7; ++i) s_pose[i] = q[i];
      quat_to_rot(q, s_new.R);
    }
    __syncthreads();
    if (tid < 7) a.pose_out[(size_t)f * 7 + tid] = s_pose[tid];
    if (tid == 0) a.status[f] = fail ? GMMLOC_FRAME_NUMERIC : 0;
    {
      // outputs in the ORIGINAL point order; outlier post-check at the final estimate in fp64:
      // chi2 > 5.991 / 7.815 or non-positive depth (localization_opt.cpp:857-881)
      double R[9], t[3];
#pragma unroll
      for (int i = 0; i < 9; ++i) R[i] = s_new.R[i];
#pragma unroll
      for (int i = 0; i < 3; ++i) t[i] = s_pose[4 + i];
      float* po = a.pts_out ? a.pts_out + (size_t)f * N * 3 : nullptr;
      uint8_t* out = a.outlier ? a.outlier + (size_t)f * N : nullptr;
      for (int j = tid; j < N; j += kThreads) {
        const int i = perm[j];
        Vec3 xw;
        if (fail) xw = {(double)pts[i * 3], (double)pts[i * 3 + 1], (double)pts[i * 3 + 2]};
        else xw = {xs[j * 3], xs[j * 3 + 1], xs[j * 3 + 2]};
        if (po != nullptr) {
          po[i * 3] = fail ? pts[i * 3] : (float)xw.x;
          po[i * 3 + 1] = fail ? pts[i * 3 + 1] : (float)xw.y;
          po[i * 3 + 2] = fail ? pts[i * 3 + 2] : (float)xw.z;
        }
        if (out != nullptr) {
          Vec3 pc = rot(R, xw);
          pc.x += t[0]; pc.y += t[1]; pc.z += t[2];
          const bool stereo = fl[j] & F_STEREO;
          const double invz = 1.0 / pc.z;
          const double u = pc.x * invz * P.fx + P.cx, v = pc.y * invz * P.fy + P.cy;
          const double e0 = (double)so[j * 3] - u, e1 = (double)so[j * 3 + 1] - v;
          double e2 = 0.0;
          if (stereo) e2 = (double)so[j * 3 + 2] - (u - P.bf * invz);
          const double chi = (double)s_w[soct[j] & 7] * (e0 * e0 + e1 * e1 + e2 * e2);
          const double th = stereo ? P.chi2_stereo : P.chi2_mono;
          out[i] = (chi > th || !(pc.z > 0.0)) ? 1 : 0;
        }
      }
    }
  }
}

// shared-memory bytes of one frame: base state (38 B / point), component id + original index (6 B), stored linearisation (48 B)
size_t w_base(int N) { return (((size_t)N * 38 + 15) & ~(size_t)15) + 16; }
size_t w_aux(int N) { return (((size_t)N * 6 + 15) & ~(size_t)15); }
size_t w_store(int N) { return (size_t)N * 4 * kStoreW; }

template <int kThreads, int kMinB>
cudaError_t launch_world_t(const FrameArgs& a, const float* gn_tab, int* d_counters, size_t smem, float* store_ws, size_t store_ws_bytes,
                           int aux_global, const unsigned char* prep, cudaStream_t stream, bool* fits) {
  cudaError_t e;
  // the opt-in is state shared by every host thread launching this variant: always the SAME value (see K3m's launcher)
  cudaFuncAttributes fa;
  if ((e = cudaFuncGetAttributes(&fa, frame_gn_world_kernel<kThreads, kMinB>)) != cudaSuccess) return e;
  const size_t max_dyn = kMinB == 2 ? ((228 * 1024 / 2 - 1024 - fa.sharedSizeBytes) & ~(size_t)15)
                                    : ((227 * 1024 - fa.sharedSizeBytes) & ~(size_t)15);
  if (smem > max_dyn) {
    *fits = false;
    return cudaSuccess;
  }
  if ((e = cudaFuncSetAttribute(frame_gn_world_kernel<kThreads, kMinB>, cudaFuncAttributeMaxDynamicSharedMemorySize, (int)max_dyn)) !=
      cudaSuccess)
    return e;
  if (GMMLOC_GNW_CARVEOUT > 0 &&
      (e = cudaFuncSetAttribute(frame_gn_world_kernel<kThreads, kMinB>, cudaFuncAttributePreferredSharedMemoryCarveout,
                                GMMLOC_GNW_CARVEOUT)) != cudaSuccess)
    return e;
  int per_sm = 0;
  if ((e = cudaOccupancyMaxActiveBlocksPerMultiprocessor(&per_sm, frame_gn_world_kernel<kThreads, kMinB>, kThreads, smem)) != cudaSuccess)
    return e;
  if (per_sm < 1) per_sm = 1;
  int dev = 0, sms = 148;
  cudaGetDevice(&dev);
  cudaDeviceGetAttribute(&sms, cudaDevAttrMultiProcessorCount, dev);
  const int grid = std::min(a.B, sms * per_sm);
  if (store_ws != nullptr && (size_t)grid * world_slice(a.N) > store_ws_bytes) {
    *fits = false;
    return cudaSuccess;
  }
  *fits = true;
  if ((e = cudaMemsetAsync(d_counters, 0, sizeof(int), stream)) != cudaSuccess) return e;
  frame_gn_world_kernel<kThreads, kMinB><<<grid, kThreads, smem, stream>>>(a, gn_tab, d_counters, store_ws, aux_global, prep);
  return cudaGetLastError();
}

}  // namespace

size_t frame_gn_world_limit_points() {
  const size_t by_smem = (size_t)(216 * 1024 - 32) / 38;
  return std::min(by_smem, (size_t)(8 * kMaxChunks - 12) * 32);  // the preparation kernel's histogram bounds N too
}

size_t frame_gn_world_scratch_bytes(int N) {
  int dev = 0, sms = 148;
  cudaGetDevice(&dev);
  cudaDeviceGetAttribute(&sms, cudaDevAttrMultiProcessorCount, dev);
  return (size_t)2 * sms * world_slice(N);
}

size_t frame_gn_world_prep_bytes(int B, int N) { return (size_t)B * prep_stride(N) + 16; }

cudaError_t launch_frame_gn_world(const FrameArgs& a, const float* gn_tab, int* d_counters, float* store_ws, size_t store_ws_bytes,
                                  unsigned char* prep_ws, size_t prep_ws_bytes, cudaStream_t stream) {
  const size_t need = (size_t)a.B * prep_stride(a.N);
  if (prep_ws == nullptr || need == 0) return cudaErrorInvalidValue;
  unsigned char* prep = reinterpret_cast<unsigned char*>((reinterpret_cast<uintptr_t>(prep_ws) + 15) & ~(uintptr_t)15);
  if ((size_t)(prep - prep_ws) + need > prep_ws_bytes) return cudaErrorInvalidValue;
  cudaError_t e = launch_frame_prep(a, prep, stream);
  if (e != cudaSuccess) return e;
  const size_t base = w_base(a.N) + w_aux(a.N), with_store = base + w_store(a.N);
  bool fits = false;
  // two resident frames of 256 threads (16 warps per SM) whenever they fit; the stored linearisation in shared memory if
  // that fits too, else in the L2-resident global scratch
  e = launch_world_t<kWThreads2, 2>(a, gn_tab, d_counters, with_store, nullptr, 0, 0, prep, stream, &fits);
  if (e != cudaSuccess || fits) return e;
  if (store_ws != nullptr) {
    e = launch_world_t<kWThreads2, 2>(a, gn_tab, d_counters, base, store_ws, store_ws_bytes, 0, prep, stream, &fits);
    if (e != cudaSuccess || fits) return e;
    // frames too large for two residents: one CTA of 512 threads; the aux arrays move behind the store in the scratch
    e = launch_world_t<512, 1>(a, gn_tab, d_counters, base, store_ws, store_ws_bytes, 0, prep, stream, &fits);
    if (e != cudaSuccess || fits) return e;
    e = launch_world_t<512, 1>(a, gn_tab, d_counters, w_base(a.N), store_ws, store_ws_bytes, 1, prep, stream, &fits);
    if (e != cudaSuccess || fits) return e;
  }
  return cudaErrorInvalidValue;  // the caller checks frame_gn_world_limit_points() and passes the scratch: not reached
}

}  // namespace gmmloc
