// frame_gn_stream.cu — K3s: the world-frame mixed-precision frame Gauss-Newton with every per-sweep operand STREAMED
// through shared memory by TMA bulk copies, plus its preparation kernel K3q.
//
// System, reference citations and precision split: frame_gn.cu (formulation), frame_gn_world.cuh (world-frame algebra,
// shared with K3w).  K3m / K3w are bound by latency, not by arithmetic or bandwidth (ncu, profiles/frame_gn_world_r02.txt:
// 33 % issue-slot utilisation at 12 warps per SM, 3.6 long-scoreboard stall cycles per issued instruction): every point
// and sweep waited for two gathers from L2 (component mean + structure table, through the component id) and for the
// linearisation it had stored in its L2-resident scratch slice.  Here no thread of the sweep loop touches global memory
// with a load:
//
//   * K3q (frame_pack_kernel) classifies and class-sorts a frame like K3p and also resolves the gathers ONCE: it writes,
//     per 32-point chunk of one class, the component constants the linearisation needs — already reduced against the
//     input landmark x0:  planar  (n, es0 = n.(x0 - mu)),  full Gaussian  (d0 = x0 - mu, Sigma^-1).  With the state kept
//     as x = x0 + delta (x0 fp32 exactly as given, delta fp32) the structure residual is  es0 + n.delta  /  d0 + delta:
//     no fp64 mean in the sweep.  K3q runs at full occupancy, where the dependent gathers are cheap;
//   * the solver's warps each own every kWarps-th chunk of the frame.  While a warp works on chunk i, the constants and
//     the stored linearisation of its chunk i + 1 are in flight into the other half of its private two-slot ring
//     (cp.async.bulk global -> shared, completion on an mbarrier: SASS UBLKCP + SYNCS), issued by lane 0.  Update and
//     linearisation of a point are fused again (one pass per sweep); the new linearisation goes straight from registers
//     to the CTA's L2-resident slice (st.global.cg, fire and forget) and comes back by TMA in the next sweep;
//   * chunks are single-class (a class that does not fill its last chunk leaves lanes idle), so the class dispatch is one
//     warp-uniform branch per 32 points and the record layouts are class-specific: 16 B of constants + 24 B of stored
//     linearisation for the planar classes (W = I - kz n^T), 36 + 48 B for a full Gaussian, 12 B for a plain stereo point.
//
// Shared memory per frame: x0, delta, observation (12 B each), octave, flags = 38 B per point + the rings (2 x 2688 B per
// warp): two frames of 2 000 points are resident per SM.  Deterministic (fixed-order reductions, no float atomics).
#include <math_constants.h>

#include <algorithm>
#include <cstddef>
#include <cstdint>
#include <type_traits>

#include "common.cuh"
#include "gn_common.cuh"
#include "frame_gn_point.cuh"
#include "frame_gn_classes.cuh"
#include "frame_gn_world.cuh"

namespace gmmloc {

namespace {

using gnw::Cam;
using gnw::Lin;
using gnw::PlaneLin;
using gnw::PointLin;
using gnw::S3f;
using gnw::V3f;

#ifndef GMMLOC_GNS_THREADS2
#define GMMLOC_GNS_THREADS2 192
#endif
constexpr int kSThreads2 = GMMLOC_GNS_THREADS2;
constexpr int kAccS = 27;
constexpr int kCstBytes = 1152;    // constants of one chunk:  2 float4 rows + 1 float row (full Gaussian); planar: 1 float4 row
constexpr int kStoreBytes = 1536;  // stored linearisation of one chunk: 3 float4 rows (full Gaussian)
constexpr int kSlotBytes = kCstBytes + kStoreBytes;
constexpr int kMaxNC = 168;        // chunks of the largest frame the kernel takes ((N + 31) / 32 + 8)

__host__ __device__ inline size_t al16(size_t x) { return (x + 15) & ~(size_t)15; }
__host__ __device__ inline int pack_max_chunks(int N) { return (N + 31) / 32 + 8; }

// Layout of frame f in the buffer frame_pack_kernel writes (all arrays in class-sorted order, every section 16-byte aligned):
//   hdr   int32 x 4: chunks, first position of class F64, points of class F64, 0
//   chunk uint32 per chunk: first position (bits 0..15) | points (16..23) | class (24..31)
//   x0 f32 x 3 | obs f32 x 3 | octave u8 | flags u8      <- one contiguous block = the solver's resident state
//   perm  u16 original index
//   cst   kCstBytes per chunk
struct PackLayout {
  size_t chunk, x0, obs, oct, fl, perm, cst, stride, state_bytes;
  int nc;
};
__host__ __device__ inline PackLayout pack_layout(int N) {
  PackLayout L;
  L.nc = pack_max_chunks(N);
  size_t o = 16;
  L.chunk = o; o += al16(4 * (size_t)L.nc);
  L.x0 = o; o += al16(12 * (size_t)N);
  L.obs = o; o += al16(12 * (size_t)N);
  L.oct = o; o += al16((size_t)N);
  L.fl = o; o += al16((size_t)N);
  L.state_bytes = o - L.x0;
  L.perm = o; o += al16(2 * (size_t)N);
  L.cst = o; o += (size_t)L.nc * kCstBytes;
  L.stride = o;
  return L;
}
// the CTA's scratch slice holds, per chunk, the constants (copied there in the first sweep) AND the stored linearisation back
// to back — [constants of the class | stored linearisation of the class] — so that every later sweep brings a chunk in with
// ONE bulk copy
__host__ __device__ inline size_t stream_slice(int N) { return (size_t)pack_max_chunks(N) * kSlotBytes; }
// bytes of constants / of stored linearisation of one chunk of class c, in units of 128 B, five bits per class from class 1
constexpr uint32_t lut5(int fixed, int plain, int mplane, int strp, int str, int mstr) {
  return (uint32_t)fixed | ((uint32_t)plain << 5) | ((uint32_t)mplane << 10) | ((uint32_t)strp << 15) | ((uint32_t)str << 20) |
         ((uint32_t)mstr << 25);
}
constexpr uint32_t kLutCst = lut5(0, 0, 4, 4, 9, 9);      // 0, 0, 512, 512, 1152, 1152 B
constexpr uint32_t kLutStore = lut5(0, 3, 6, 6, 12, 12);  // 0, 384, 768, 768, 1536, 1536 B
constexpr uint32_t kLutBoth = lut5(0, 3, 10, 10, 21, 21);
constexpr uint32_t kLutPlanarCst = lut5(0, 0, 4, 4, 0, 0);
__device__ __forceinline__ uint32_t lut_get(uint32_t lut, int cls) { return ((lut >> (5 * (cls - 1))) & 31u) << 7; }
template <int CLS>
struct ClsBytes {
  static constexpr int cst = (CLS == C_MPLANE || CLS == C_STRP) ? 512 : (CLS == C_STR || CLS == C_MSTR) ? kCstBytes : 0;
};

__device__ __forceinline__ uint32_t smem_u32(const void* p) { return (uint32_t)__cvta_generic_to_shared(p); }
__device__ __forceinline__ void mbar_init(uint64_t* bar, uint32_t count) {
  asm volatile("mbarrier.init.shared::cta.b64 [%0], %1;" ::"r"(smem_u32(bar)), "r"(count));
}
__device__ __forceinline__ void mbar_expect_tx(uint64_t* bar, uint32_t bytes) {
  asm volatile("mbarrier.arrive.expect_tx.shared::cta.b64 _, [%0], %1;" ::"r"(smem_u32(bar)), "r"(bytes) : "memory");
}
__device__ __forceinline__ void mbar_wait(uint64_t* bar, uint32_t parity) {
  uint32_t done;
  do {
    asm volatile(
        "{\n"
        ".reg .pred p;\n"
        "mbarrier.try_wait.parity.shared::cta.b64 p, [%1], %2;\n"
        "selp.u32 %0, 1, 0, p;\n"
        "}\n"
        : "=r"(done)
        : "r"(smem_u32(bar)), "r"(parity)
        : "memory");
  } while (!done);
}
__device__ __forceinline__ void mbar_wait_a(uint32_t bar, uint32_t parity) {  // `bar`: shared-window address
  uint32_t done;
  do {
    asm volatile(
        "{\n"
        ".reg .pred p;\n"
        "mbarrier.try_wait.parity.shared::cta.b64 p, [%1], %2;\n"
        "selp.u32 %0, 1, 0, p;\n"
        "}\n"
        : "=r"(done)
        : "r"(bar), "r"(parity)
        : "memory");
  } while (!done);
}
// TMA 1-D bulk copy global -> shared, completion signalled on an mbarrier (SASS: UBLKCP)
__device__ __forceinline__ void tma_bulk_g2s(void* dst_smem, const void* src_gmem, uint32_t bytes, uint64_t* bar) {
  asm volatile("cp.async.bulk.shared::cluster.global.mbarrier::complete_tx::bytes [%0], [%1], %2, [%3];" ::"r"(
                   smem_u32(dst_smem)),
               "l"(src_gmem), "r"(bytes), "r"(smem_u32(bar))
               : "memory");
}
// Volatile vector loads from shared memory.  The sweep re-reads the pose, the pose step and the reference point at every
// point (carrying them would cost ~40 registers across the whole class dispatch); read as 128-bit words they are 12 LDS per
// point instead of 30.  asm volatile: never hoisted out of the chunk loop or merged, exactly like the volatile views they replace.
#ifndef GMMLOC_GNS_VEC_LDS
#define GMMLOC_GNS_VEC_LDS 1
#endif
__device__ __forceinline__ void lds_v2f64(uint32_t addr, double& a, double& b) {
  asm volatile("ld.shared.v2.f64 {%0, %1}, [%2];" : "=d"(a), "=d"(b) : "r"(addr));
}
__device__ __forceinline__ float4 lds_v4f32(uint32_t addr) {
  float4 v;
  asm volatile("ld.shared.v4.f32 {%0, %1, %2, %3}, [%4];" : "=f"(v.x), "=f"(v.y), "=f"(v.z), "=f"(v.w) : "r"(addr));
  return v;
}
// generic-proxy writes to global memory (the stored linearisation) ordered before later async-proxy (TMA) reads of them
__device__ __forceinline__ void fence_proxy_async_global() { asm volatile("fence.proxy.async.global;" ::: "memory"); }

// ================================================================================================================
// K3q: classification + stable class sort (the arithmetic of frame_prep_kernel, frame_gn_mixed.cu) + chunk table +
// per-chunk component constants.  One CTA of 256 threads per frame.
// K3c: flags and class of every point of the batch, ONE THREAD PER POINT (structure-edge gates fixed at iteration 0, in fp64
// exactly as frame_gn_kernel).  The gate needs three dependent round trips per point (assignment -> degenerate flag ->
// component constants); inside the per-frame pack kernel, whose threads walk eight points one after the other, they held 17
// stall cycles per issued instruction.  Here every point is its own thread at full occupancy.
#ifndef GMMLOC_CLS_MINB
#define GMMLOC_CLS_MINB 4
#endif
// Classification as its own one-thread-per-point kernel (1) or inside the per-frame pack kernel (0, default).  Measured on
// C2 (4 096 frames): split 0.19 + 0.25 ms, fused 0.40 ms — the pack kernel's scattered stores, not the gate's dependent
// gathers, set its pace, and the split pays a second pass over the landmarks.
#ifndef GMMLOC_GNS_SPLIT_CLASSIFY
#define GMMLOC_GNS_SPLIT_CLASSIFY 0
#endif
#if GMMLOC_GNS_SPLIT_CLASSIFY
// per-frame constants of the classification (the normalised initial pose): third row of R, t_z, camera centre -R^T t
__global__ void __launch_bounds__(128) frame_pose_kernel(const FrameArgs a, double* __restrict__ fpose) {
  const int f = blockIdx.x * 128 + threadIdx.x;
  if (f >= a.B) return;
  double q[7], R[9];
  for (int i = 0; i < 7; ++i) q[i] = a.pose0[(size_t)f * 7 + i];
  quat_normalize(q);
  quat_to_rot(q, R);
  double* o = fpose + (size_t)f * 8;
  o[0] = R[6]; o[1] = R[7]; o[2] = R[8]; o[3] = q[6];
  o[4] = -(R[0] * q[4] + R[3] * q[5] + R[6] * q[6]);
  o[5] = -(R[1] * q[4] + R[4] * q[5] + R[7] * q[6]);
  o[6] = -(R[2] * q[4] + R[5] * q[5] + R[8] * q[6]);
  o[7] = 0.0;
}
#endif
// flag byte (low nibble) and class (bits 4..6) of one point; fp = the per-frame constants of frame_pose_kernel
// deg: the component's degenerate flag (a.is_deg[k]) when k >= 0 and the points are free, else unused — the callers load it so
// that they can have several of these first-level gathers in flight
__device__ __forceinline__ uint8_t classify_point(const FrameArgs& a, const double* fp, float px, float py, float pz, float our, int k,
                                                  int deg) {
  const FrameParamsDev& P = a.prm;
  const bool pose_free = P.factor_mode != GMMLOC_MODE_POINTS_ONLY;
  const bool pts_free = P.factor_mode != GMMLOC_MODE_POSE_ONLY;
  const double view_cos_thresh = 0.20791169081775934;  // cos(78 deg), gaussian_mixture.cpp:284
  const double x0 = (double)px, x1 = (double)py, x2 = (double)pz;
  uint8_t b = (our >= 0.0f) ? F_STEREO : 0;
  if (k >= 0 && pts_free) {
    if (deg) {
      b |= F_DEG;
      const double* ax = a.axis + (size_t)k * 9;
      const double* mu = a.mean + (size_t)k * 3;
      const double e = dadd(dadd(dmul(ax[0], dsub(x0, mu[0])), dmul(ax[3], dsub(x1, mu[1]))), dmul(ax[6], dsub(x2, mu[2])));
      if (dmul(e, e) <= P.tri_str_thresh) b |= F_STR;
      const double p0 = mu[0] - fp[4], p1 = mu[1] - fp[5], p2 = mu[2] - fp[6];  // from the camera centre -R^T t
      const double view_cos = fabs(p0 * ax[0] + p1 * ax[3] + p2 * ax[6]) / sqrt(p0 * p0 + p1 * p1 + p2 * p2);
      if (view_cos < view_cos_thresh) b &= ~F_STR;
    } else {
      if (chi2_exact(a.mean + (size_t)k * 3, a.cov_inv + (size_t)k * 9, x0, x1, x2) <= P.mdist2_gate) b |= F_STR;
    }
  }
  if (pts_free && ((b & F_STEREO) || (b & F_STR))) b |= F_FREE;
  int cls;
  if (!(b & F_FREE)) {
    cls = pose_free ? C_FIXED : C_OFF;
  } else {
    const double pzc = fp[0] * x0 + fp[1] * x1 + fp[2] * x2 + fp[3];
    const double rx = x0 - fp[4], ry = x1 - fp[5], rz = x2 - fp[6];
    if (!fp32_geometry_ok(pzc, rx * rx + ry * ry + rz * rz)) cls = C_F64;
    else if (!(b & F_STEREO)) cls = (b & F_DEG) ? C_MPLANE : C_MSTR;  // a free monocular point always has F_STR
    else cls = (b & F_STR) ? ((b & F_DEG) ? C_STRP : C_STR) : C_PLAIN;
  }
  return (uint8_t)(b | (cls << 4));
}
#if GMMLOC_GNS_SPLIT_CLASSIFY
__global__ void __launch_bounds__(256, GMMLOC_CLS_MINB) frame_classify_kernel(const FrameArgs a, const double* __restrict__ fpose,
                                                                               uint8_t* __restrict__ flags) {
  const int N = a.N;
  const long long total = (long long)a.B * N;
  for (long long gi = (long long)blockIdx.x * 256 + threadIdx.x; gi < total; gi += (long long)gridDim.x * 256) {
    const int f = (int)(gi / N);
    double fp[7];
#pragma unroll
    for (int i = 0; i < 7; ++i) fp[i] = __ldg(fpose + (size_t)f * 8 + i);
    const int k = __ldg(a.assign + gi);
    const int deg = (k >= 0 && a.prm.factor_mode != GMMLOC_MODE_POSE_ONLY) ? (int)a.is_deg[k] : 0;
    flags[gi] = classify_point(a, fp, __ldg(a.pts_w + gi * 3), __ldg(a.pts_w + gi * 3 + 1), __ldg(a.pts_w + gi * 3 + 2),
                               __ldg(a.obs + gi * 3 + 2), k, deg);
  }
}
#endif

constexpr int kPackThreads = 256;
#ifndef GMMLOC_PACK_MATCH
#define GMMLOC_PACK_MATCH 1
#endif
#ifndef GMMLOC_PACK_PSCAN
#define GMMLOC_PACK_PSCAN 1
#endif
#ifndef GMMLOC_PACK_GROUP
#define GMMLOC_PACK_GROUP 8
#endif
constexpr int kPackG = GMMLOC_PACK_GROUP;  // points per thread whose first-level gathers are in flight together
#ifndef GMMLOC_PACK_GRID
#define GMMLOC_PACK_GRID GMMLOC_PACK_MINB  // CTAs per SM the grid is sized for: the resident count (one wave, no tail of a partial one)
#endif
#ifndef GMMLOC_PACK_MINB
#define GMMLOC_PACK_MINB 5  // 51 registers: five CTAs per SM; measured 2.96 ms for the stage against 3.02 (four CTAs) and 3.00 (six, spills)
#endif
__global__ void __launch_bounds__(kPackThreads, GMMLOC_PACK_MINB) frame_pack_kernel(const FrameArgs a, const float* __restrict__ gn_tab,
                                                                  const uint8_t* __restrict__ flags, unsigned char* __restrict__ pack) {
  constexpr int kWarps = kPackThreads / 32;
  extern __shared__ __align__(16) unsigned char smem_raw[];  // flags by ORIGINAL index while the frame is being sorted
  __shared__ int s_cnt[C_COUNT * kMaxChunks * 8];
  __shared__ double s_fp[8];                                // fused classification: third row of R, t_z, camera centre
  __shared__ int s_start[C_COUNT + 1], s_cb[C_COUNT + 1];  // first sorted position / first chunk of every class
  __shared__ int s_tot[C_COUNT];                            // points per class
  (void)s_tot;
  uint8_t* tmp = smem_raw;
  const int tid = threadIdx.x, lane = tid & 31, warp = tid >> 5;
  const int N = a.N;
  const int nblocks = (N + kPackThreads - 1) / kPackThreads;
  const PackLayout L = pack_layout(N);
  int kk[kPackG];  // component ids of the thread's current group of points
  for (int f = blockIdx.x; f < a.B; f += gridDim.x) {
    const float* pts = a.pts_w + (size_t)f * N * 3;
    const float* obs = a.obs + (size_t)f * N * 3;
    const uint8_t* oct = a.octave + (size_t)f * N;
    const int32_t* asg = a.assign + (size_t)f * N;
    unsigned char* pf = pack + (size_t)f * L.stride;
    float* gp = reinterpret_cast<float*>(pf + L.x0);
    float* go = reinterpret_cast<float*>(pf + L.obs);
    uint16_t* gperm = reinterpret_cast<uint16_t*>(pf + L.perm);
    uint8_t* goct = pf + L.oct;
    uint8_t* gfl = pf + L.fl;
    unsigned char* gcst = pf + L.cst;
    const uint8_t* gflag = flags ? flags + (size_t)f * N : nullptr;
    __syncthreads();  // the previous frame is done with the shared arrays
#if GMMLOC_PACK_MATCH
    // the histogram below is written by the leader of every group of same-class lanes only: absent classes must read zero
    for (int e = tid; e < C_COUNT * nblocks * kWarps; e += kPackThreads)
      s_cnt[(e / (nblocks * kWarps)) * (kMaxChunks * 8) + e % (nblocks * kWarps)] = 0;
#endif
    if (flags == nullptr && tid == 0) {  // fused classification (the default): per-frame constants as frame_pose_kernel computes them
      double q[7], R[9];
      for (int i = 0; i < 7; ++i) q[i] = a.pose0[(size_t)f * 7 + i];
      quat_normalize(q);
      quat_to_rot(q, R);
      s_fp[0] = R[6]; s_fp[1] = R[7]; s_fp[2] = R[8]; s_fp[3] = q[6];
      s_fp[4] = -(R[0] * q[4] + R[3] * q[5] + R[6] * q[6]);
      s_fp[5] = -(R[1] * q[4] + R[4] * q[5] + R[7] * q[6]);
      s_fp[6] = -(R[2] * q[4] + R[5] * q[5] + R[8] * q[6]);
    }
    __syncthreads();
    // pass A: flags and class of every point — computed here (classify_point, the arithmetic of frame_classify_kernel) or
    // read from the flags that kernel wrote — and the class histogram per (block, warp)
    // The gate is a chain of dependent gathers (assignment -> degenerate flag -> component constants) and a thread walks its
    // points one after the other, so the first two levels are issued for a whole group of kPackG points at once (component ids
    // kept in registers for pass B): one serialised round trip per point instead of three.
    for (int c0 = 0; c0 < nblocks; c0 += kPackG) {
      unsigned dg = 0;
      if (gflag == nullptr) {
#pragma unroll
        for (int u = 0; u < kPackG; ++u) {
          const int i = (c0 + u) * kPackThreads + tid;
          kk[u] = (c0 + u < nblocks && i < N) ? asg[i] : -1;
        }
        if (a.prm.factor_mode != GMMLOC_MODE_POSE_ONLY) {
#pragma unroll
          for (int u = 0; u < kPackG; ++u)
            if (kk[u] >= 0) dg |= (unsigned)(a.is_deg[kk[u]] != 0) << u;
        }
      }
#pragma unroll
      for (int u = 0; u < kPackG; ++u) {
        const int c = c0 + u;
        if (c >= nblocks) break;  // uniform
        const int i = c * kPackThreads + tid;
        int cls = -1;
        if (i < N) {
          const uint8_t b = gflag ? gflag[i]
                                  : classify_point(a, s_fp, pts[i * 3], pts[i * 3 + 1], pts[i * 3 + 2], obs[i * 3 + 2], kk[u], (dg >> u) & 1u);
          tmp[i] = b;
          cls = b >> 4;
        }
#if GMMLOC_PACK_MATCH
        {  // one MATCH instead of a ballot per class: the lowest lane of every group of same-class lanes writes the group's size
          const unsigned mk = __match_any_sync(0xffffffffu, cls);
          if (cls >= 0 && lane == __ffs(mk) - 1) s_cnt[cls * (kMaxChunks * 8) + c * kWarps + warp] = __popc(mk);
        }
#else
#pragma unroll
        for (int q = 0; q < C_COUNT; ++q) {
          const unsigned m = __ballot_sync(0xffffffffu, cls == q);
          if (lane == 0) s_cnt[q * (kMaxChunks * 8) + c * kWarps + warp] = __popc(m);
        }
#endif
      }
    }
    __syncthreads();
#if GMMLOC_PACK_PSCAN
    {  // exclusive scan in (block, warp) order INSIDE every class, one warp per class (one warp walking all eight classes kept
       // the other seven waiting at the barrier: 6.5 % of the kernel's stall samples); the class offsets follow below
      static_assert(kWarps == C_COUNT, "one warp per class");
      const int q = warp, n = nblocks * kWarps;
      int run = 0;
      for (int b0 = 0; b0 < n; b0 += 32) {
        const int e = b0 + lane;
        const int idx = q * (kMaxChunks * 8) + e;
        const int v = (e < n) ? s_cnt[idx] : 0;
        int inc = v;
#pragma unroll
        for (int off = 1; off < 32; off <<= 1) {
          const int t = __shfl_up_sync(0xffffffffu, inc, off);
          if (lane >= off) inc += t;
        }
        if (e < n) s_cnt[idx] = run + inc - v;
        run += __shfl_sync(0xffffffffu, inc, 31);
      }
      if (lane == 0) s_tot[q] = run;
    }
    __syncthreads();
    if (tid == 0) {
      int run = 0, cb = 0;
      for (int q = 0; q < C_COUNT; ++q) {
        s_start[q] = run;
        s_cb[q] = cb;
        if (q >= C_FIXED && q < C_F64) cb += (s_tot[q] + 31) / 32;
        run += s_tot[q];
      }
      s_start[C_COUNT] = run;  // = N
      s_cb[C_COUNT] = cb;
    }
#else
    if (warp == 0) {  // exclusive scan in (class, block, warp) order = stable counting-sort offsets
      int run = 0;
      for (int q = 0; q < C_COUNT; ++q) {
        if (lane == 0) s_start[q] = run;
        const int n = nblocks * kWarps;
        for (int b0 = 0; b0 < n; b0 += 32) {
          const int e = b0 + lane;
          const int idx = q * (kMaxChunks * 8) + e;
          const int v = (e < n) ? s_cnt[idx] : 0;
          int inc = v;
#pragma unroll
          for (int off = 1; off < 32; off <<= 1) {
            const int t = __shfl_up_sync(0xffffffffu, inc, off);
            if (lane >= off) inc += t;
          }
          if (e < n) s_cnt[idx] = run + inc - v;
          run += __shfl_sync(0xffffffffu, inc, 31);
        }
      }
      if (lane == 0) {
        s_start[C_COUNT] = run;  // = N
        int cb = 0;
        for (int q = 0; q < C_COUNT; ++q) {
          s_cb[q] = cb;
          if (q >= C_FIXED && q < C_F64) cb += (s_start[q + 1] - s_start[q] + 31) / 32;
        }
        s_cb[C_COUNT] = cb;
      }
    }
#endif
    __syncthreads();
    {  // header + chunk table
      const int nch = s_cb[C_COUNT];
      if (tid == 0) {
        int32_t* hdr = reinterpret_cast<int32_t*>(pf);
        hdr[0] = nch;
        hdr[1] = s_start[C_F64];
        hdr[2] = N - s_start[C_F64];
        hdr[3] = 0;
      }
      uint32_t* gchunk = reinterpret_cast<uint32_t*>(pf + L.chunk);
      for (int g = tid; g < nch; g += kPackThreads) {
        int q = C_FIXED;
        while (g >= s_cb[q + 1]) ++q;
        const int j0 = s_start[q] + 32 * (g - s_cb[q]);
        const int cnt = min(32, s_start[q + 1] - j0);
        gchunk[g] = (uint32_t)j0 | ((uint32_t)cnt << 16) | ((uint32_t)q << 24);
      }
    }
    // pass B: scatter the frame into class-sorted order, resolve the component constants of the structured classes
    for (int c0 = 0; c0 < nblocks; c0 += kPackG) {
      if (gflag != nullptr || nblocks > kPackG) {  // (a frame of at most kPackG blocks still has pass A's ids in registers)
#pragma unroll
        for (int u = 0; u < kPackG; ++u) {
          const int i = (c0 + u) * kPackThreads + tid;
          kk[u] = (c0 + u < nblocks && i < N) ? asg[i] : -1;
        }
      }
#pragma unroll
      for (int u = 0; u < kPackG; ++u) {
        const int c = c0 + u;
        if (c >= nblocks) break;  // uniform
        const int i = c * kPackThreads + tid;
        int cls = -1;
        uint8_t b = 0;
        if (i < N) {
          b = tmp[i];
          cls = b >> 4;
        }
        int j = -1;
#if GMMLOC_PACK_MATCH
        {
          const unsigned mk = __match_any_sync(0xffffffffu, cls);
          if (cls >= 0) j = s_cnt[cls * (kMaxChunks * 8) + c * kWarps + warp] + __popc(mk & ((1u << lane) - 1u));
        }
#else
#pragma unroll
        for (int q = 0; q < C_COUNT; ++q) {
          const unsigned m = __ballot_sync(0xffffffffu, cls == q);
          if (cls == q) j = s_cnt[q * (kMaxChunks * 8) + c * kWarps + warp] + __popc(m & ((1u << lane) - 1u));
        }
#endif
#if GMMLOC_PACK_PSCAN
        if (cls >= 0) j += s_start[cls];  // the per-class scan left offsets relative to the class
#endif
        if (i < N) {
          const float p0 = pts[i * 3], p1 = pts[i * 3 + 1], p2 = pts[i * 3 + 2];
          gp[j * 3] = p0;
          gp[j * 3 + 1] = p1;
          gp[j * 3 + 2] = p2;
          go[j * 3] = obs[i * 3];
          go[j * 3 + 1] = obs[i * 3 + 1];
          go[j * 3 + 2] = obs[i * 3 + 2];
          gperm[j] = (uint16_t)i;
          goct[j] = oct[i];
          gfl[j] = b;
          if (cls >= C_MPLANE && cls <= C_MSTR) {
            const int k = kk[u];
            const double* mu = a.mean + (size_t)k * 3;
            const double d0 = (double)p0 - mu[0], d1 = (double)p1 - mu[1], d2 = (double)p2 - mu[2];
            const int r = j - s_start[cls];
            unsigned char* cp = gcst + (size_t)(s_cb[cls] + (r >> 5)) * kCstBytes;
            const int l = r & 31;
            const float4 t0 = __ldg(reinterpret_cast<const float4*>(gn_tab) + 2 * (size_t)k);
            if (cls == C_MPLANE || cls == C_STRP) {
              // es0 = n . (x0 - mu) with the fp32 normal the solver uses, evaluated in fp64 and rounded once
              const float es0 = (float)(((double)t0.x * d0 + (double)t0.y * d1) + (double)t0.z * d2);
              reinterpret_cast<float4*>(cp)[l] = make_float4(t0.x, t0.y, t0.z, es0);
            } else {
              // six unique entries of Sigma^-1 (C_STR) or — gain form of the monocular block, frame_gn_world.cuh — of Sigma (C_MSTR)
              float m0, m1, m2, m3, m4, m5;
              if (cls == C_STR) {
                const float2 t1 = __ldg(reinterpret_cast<const float2*>(gn_tab) + 4 * (size_t)k + 2);
                m0 = t0.x; m1 = t0.y; m2 = t0.z; m3 = t0.w; m4 = t1.x; m5 = t1.y;
              } else {
                const double* cv = a.cov + (size_t)k * 9;
                m0 = (float)cv[0]; m1 = (float)cv[1]; m2 = (float)cv[2]; m3 = (float)cv[4]; m4 = (float)cv[5]; m5 = (float)cv[8];
              }
              reinterpret_cast<float4*>(cp)[l] = make_float4((float)d0, (float)d1, (float)d2, m0);
              reinterpret_cast<float4*>(cp)[32 + l] = make_float4(m1, m2, m3, m4);
              reinterpret_cast<float*>(cp + 1024)[l] = m5;
            }
          }
        }
      }
    }
  }
}

// ================================================================================================================
struct __align__(16) PoseS {
  double R[9], t[3];
  float Rf[12];  // rows padded to 16 B
};

// fp64 treatment of one ill-conditioned point (class C_F64), camera frame, exactly as in K3m / K3w
__device__ __noinline__ void point_fp64_s(const FrameArgs& a, const double* Ro, const double* to, const double* Rn, const double* tn,
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

// 27 per-lane partial sums -> lane l holds the warp total of value l (recursive halving, 31 shuffles, fixed order)
__device__ __forceinline__ float warp_reduce27s(float* acc, int lane) {
  float a[32];
#pragma unroll
  for (int i = 0; i < 32; ++i) a[i] = i < kAccS ? acc[i] : 0.f;
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

template <int kThreads, int kMinB>
__global__ void __launch_bounds__(kThreads, kMinB) frame_gn_stream_kernel(const FrameArgs a, int* __restrict__ counters,
                                                                           unsigned char* __restrict__ store_ws,
                                                                           const unsigned char* __restrict__ pack) {
  constexpr int kWarps = kThreads / 32;
  extern __shared__ __align__(16) unsigned char smem_raw[];
  __shared__ double s_red[kWarps][kAccS];
  __shared__ double s_acc[kAccS];
  __shared__ double s_pose[7], s_dxc[6], s_c0[3];
  __shared__ __align__(16) float s_dxw[8];  // omega_w (3), upsilon_w (3): the pose step in world-twist coordinates about c0
  __shared__ __align__(16) float s_c0f[4];
  __shared__ PoseS s_new, s_old;
  __shared__ int s_fail1, s_fail2, s_frame;
  __shared__ int s_hdr[4];
  __shared__ float s_w[8];
  __shared__ uint32_t s_chunk[kMaxNC];
  __shared__ __align__(8) uint64_t s_bar[kWarps][2];
  __shared__ __align__(8) uint64_t s_fbar;

  const int tid = threadIdx.x, lane = tid & 31, warp = tid >> 5;
  const int N = a.N;
  const FrameParamsDev& P = a.prm;
  const bool pose_free = P.factor_mode != GMMLOC_MODE_POINTS_ONLY;
  const bool pts_free = P.factor_mode != GMMLOC_MODE_POSE_ONLY;
  Cam C;
  C.fx = P.fx; C.fy = P.fy; C.cx = P.cx; C.cy = P.cy; C.bf = P.bf;
  C.fxf = (float)P.fx; C.fyf = (float)P.fy; C.bff = (float)P.bf; C.lam = (float)P.ba_lambda2;
  C.d_mono = (float)P.delta_mono; C.d_stereo = (float)P.delta_stereo; C.robust = P.robust;

  const PackLayout L = pack_layout(N);
  // resident state (class-sorted position j): x0, observation, octave, flags = the block the pack kernel wrote; delta; rings
  const float* x0s = reinterpret_cast<const float*>(smem_raw);
  const float* so = reinterpret_cast<const float*>(smem_raw + (L.obs - L.x0));
  const uint8_t* soct = smem_raw + (L.oct - L.x0);
  const uint8_t* fl = smem_raw + (L.fl - L.x0);
  float* dl = reinterpret_cast<float*>(smem_raw + L.state_bytes);
  unsigned char* ring = smem_raw + L.state_bytes + al16(12 * (size_t)N) + (size_t)warp * 2 * kSlotBytes;
  unsigned char* gstore = store_ws + (size_t)blockIdx.x * stream_slice(N);
  const uint32_t ring_a = smem_u32(ring), bar_a = smem_u32(&s_bar[warp][0]);

  if (tid < 8) s_w[tid] = P.s2inv[tid];
  if (tid == 0) {
    for (int w = 0; w < kWarps; ++w) {
      mbar_init(&s_bar[w][0], 1);
      mbar_init(&s_bar[w][1], 1);
    }
    mbar_init(&s_fbar, 1);
    asm volatile("fence.mbarrier_init.release.cluster;" ::: "memory");
  }
  uint32_t n_iss = 0, n_con = 0;  // ring loads issued / consumed by this warp (warp-uniform)
  uint32_t fpar = 0;

  for (;;) {
    __syncthreads();
    if (tid == 0) s_frame = atomicAdd(counters, 1);
    __syncthreads();
    const int f = s_frame;
    if (f >= a.B) break;
    const unsigned char* pf = pack + (size_t)f * L.stride;
    const float* pts = a.pts_w + (size_t)f * N * 3;
    if (tid == 0) {
      // the resident block of the frame: ONE bulk copy
      mbar_expect_tx(&s_fbar, (uint32_t)L.state_bytes);
      tma_bulk_g2s(smem_raw, pf + L.x0, (uint32_t)L.state_bytes, &s_fbar);
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
      for (int c = 0; c < 3; ++c) {
        s_c0[c] = -(s_new.R[c] * q[4] + s_new.R[3 + c] * q[5] + s_new.R[6 + c] * q[6]);
        s_c0f[c] = (float)s_c0[c];
      }
      s_fail1 = 0;
      s_fail2 = 0;
    }
    if (tid < 4) s_hdr[tid] = reinterpret_cast<const int32_t*>(pf)[tid];
    {
      const uint32_t* gchunk = reinterpret_cast<const uint32_t*>(pf + L.chunk);
      for (int g = tid; g < L.nc; g += kThreads) s_chunk[g] = gchunk[g];
      for (int e = tid; e < 3 * N; e += kThreads) dl[e] = 0.f;
    }
    mbar_wait(&s_fbar, fpar);
    fpar ^= 1u;
    __syncthreads();
    const int nch = s_hdr[0], f64_begin = s_hdr[1], n64 = s_hdr[2];
    const unsigned char* gcst = pf + L.cst;
    const uint16_t* gperm = reinterpret_cast<const uint16_t*>(pf + L.perm);

    const int iters = P.gn_iters;
    const int sweeps = (iters > 0) ? iters + 1 : 0;
    for (int sw = 0; sw < sweeps; ++sw) {
      const bool do_update = pts_free && sw > 0;
      const bool do_lin = sw < iters;           // state sw is linearised (the update of sweep sw + 1 needs it)
      const bool do_acc = pose_free && do_lin;  // ... and enters the pose system
      if (!do_update && !do_acc && !(do_lin && pts_free)) continue;
      int bad = 0;  // bit 0: singular block while linearising, bit 1: non-finite step while updating
      // read through volatile views inside the loops: the pose and the reference point are then loaded where they are
      // used (broadcast LDS) instead of being carried in ~40 registers across every point
      const volatile double* vc0 = s_c0;
      const volatile double* vR = s_new.R;
      const volatile double* vt = s_new.t;
      const volatile float* vdx = s_dxw;
      const volatile float* vc0f = s_c0f;
#if GMMLOC_GNS_VEC_LDS
      const uint32_t a_pose = smem_u32(&s_new), a_rf = smem_u32(s_new.Rf), a_dxw = smem_u32(s_dxw), a_c0f = smem_u32(s_c0f);
      (void)vR; (void)vt; (void)vdx;
#endif

      float acc[kAccS];
#pragma unroll
      for (int i = 0; i < kAccS; ++i) acc[i] = 0.f;

      // ---- the chunks of this warp: g = warp, warp + kWarps, ...; chunk i + 1 is in flight while chunk i is worked on ----
      // bytes a chunk of class c brings in this sweep, in units of 128 B, one nibble per class (warp-uniform look-up):
      // constants when the chunk is linearised (the planar classes need their normal for the update too: W = I - kz n^T),
      // the stored linearisation when it is updated
      // three kinds of sweep:  first (linearise only): the constants come from the packed frame and are copied into the slice;
      // middle (update + linearise): constants + stored linearisation from the slice;  last (update only): the stored
      // linearisation, and the normal of the planar classes (W = I - kz n^T), from the slice
      const bool from_pack = !do_update;
      const uint32_t lut_n = from_pack ? ((do_lin && pts_free) ? kLutCst : 0u) : (do_lin ? kLutBoth : kLutStore + kLutPlanarCst);
      const uint32_t lut_o = (do_update && !do_lin) ? lut5(0, 0, 0, 0, 9, 9) : 0u;  // offset of the copy inside the chunk record
      auto issue = [&](int g) {  // warp-uniform bookkeeping, lane 0 issues
        const int cls = (int)(s_chunk[g] >> 24);
        const uint32_t nb = lut_get(lut_n, cls);
        if (nb == 0u) return;
        if (lane == 0) {
          const uint32_t slot = n_iss & 1u, off = lut_get(lut_o, cls);
          const uint32_t bar = bar_a + slot * 8u, dst = ring_a + slot * kSlotBytes + off;
          const unsigned char* src = from_pack ? gcst + (size_t)g * kCstBytes : gstore + (size_t)g * kSlotBytes + off;
          asm volatile("mbarrier.arrive.expect_tx.shared::cta.b64 _, [%0], %1;" ::"r"(bar), "r"(nb) : "memory");
          asm volatile("cp.async.bulk.shared::cluster.global.mbarrier::complete_tx::bytes [%0], [%1], %2, [%3];" ::"r"(dst), "l"(src),
                       "r"(nb), "r"(bar)
                       : "memory");
        }
        ++n_iss;
      };
      // one point of a chunk of class CLS (compile-time: every class gets its own straight-line code)
      auto point = [&](auto cls_tag, int j, const unsigned char* slot, unsigned char* gs) {
        constexpr int CLS = decltype(cls_tag)::value;
        constexpr bool kStereo = CLS == C_PLAIN || CLS == C_STRP || CLS == C_STR;
        const float4* cst4 = reinterpret_cast<const float4*>(slot);
        const float4* st4 = reinterpret_cast<const float4*>(slot + ClsBytes<CLS>::cst);
        unsigned char* gst = gs + ClsBytes<CLS>::cst;  // this chunk's stored linearisation in the slice
        float d0 = 0.f, d1 = 0.f, d2 = 0.f;
        if (CLS != C_FIXED) {
          d0 = dl[j * 3]; d1 = dl[j * 3 + 1]; d2 = dl[j * 3 + 2];
        }
        const double X0 = (double)x0s[j * 3], X1 = (double)x0s[j * 3 + 1], X2 = (double)x0s[j * 3 + 2];
        // x' = x - c0 for the Jacobian side, formed in fp32 (it only ever multiplies fp32 quantities)
#if GMMLOC_GNS_VEC_LDS
        const float4 c0q = lds_v4f32(a_c0f);
        const V3f xr0 = {x0s[j * 3] - c0q.x, x0s[j * 3 + 1] - c0q.y, x0s[j * 3 + 2] - c0q.z};
#else
        const V3f xr0 = {x0s[j * 3] - vc0f[0], x0s[j * 3 + 1] - vc0f[1], x0s[j * 3 + 2] - vc0f[2]};
#endif
        // ---- update: delta += u0 - W (omega_w x x' + upsilon_w) from the linearisation stored by the previous sweep ----
        if (CLS != C_FIXED && do_update) {
#if GMMLOC_GNS_VEC_LDS
          const float4 dq0 = lds_v4f32(a_dxw), dq1 = lds_v4f32(a_dxw + 16u);
          const V3f om = {dq0.x, dq0.y, dq0.z}, up = {dq0.w, dq1.x, dq1.y};
#else
          const V3f om = {vdx[0], vdx[1], vdx[2]}, up = {vdx[3], vdx[4], vdx[5]};
#endif
          const V3f av = gnw::twist_at(om, up, {xr0.x + d0, xr0.y + d1, xr0.z + d2});
          V3f dx;
          if (CLS == C_PLAIN) {
            const float* sf = reinterpret_cast<const float*>(st4);
            dx = {sf[lane] - av.x, sf[32 + lane] - av.y, sf[64 + lane] - av.z};
          } else if (CLS == C_MPLANE || CLS == C_STRP) {
            const float4 s0 = st4[lane];
            const float2 s1 = reinterpret_cast<const float2*>(st4 + 32)[lane];
            const float4 nn = cst4[lane];  // (n, es0)
            dx = gnw::plane_step({s0.x, s0.y, s0.z}, {s0.w, s1.x, s1.y}, {nn.x, nn.y, nn.z}, av);
          } else {
            const float4 s0 = st4[lane], s1 = st4[32 + lane], s2 = st4[64 + lane];
            dx = {s0.x - (s0.w * av.x + s1.x * av.y + s1.y * av.z), s0.y - (s1.z * av.x + s1.w * av.y + s2.x * av.z),
                  s0.z - (s2.y * av.x + s2.z * av.y + s2.w * av.z)};
          }
          if (!gnw::finite3(dx)) bad |= 2;
          d0 += dx.x; d1 += dx.y; d2 += dx.z;
          dl[j * 3] = d0; dl[j * 3 + 1] = d1; dl[j * 3 + 2] = d2;
        }
        // ---- linearise at the current pose: store (u0, W), accumulate the reduced pose system ----
        if (do_lin) {
          const float ob[3] = {so[j * 3], so[j * 3 + 1], so[j * 3 + 2]};
          const float w = s_w[soct[j] & 7];
          const bool stereo = CLS == C_FIXED ? ((fl[j] & F_STEREO) != 0) : kStereo;
          const V3f xr = {xr0.x + d0, xr0.y + d1, xr0.z + d2};
          Lin Ln;
          {
            double Rq[9], tq[3];
#if GMMLOC_GNS_VEC_LDS
            lds_v2f64(a_pose, Rq[0], Rq[1]);
            lds_v2f64(a_pose + 16u, Rq[2], Rq[3]);
            lds_v2f64(a_pose + 32u, Rq[4], Rq[5]);
            lds_v2f64(a_pose + 48u, Rq[6], Rq[7]);
            lds_v2f64(a_pose + 64u, Rq[8], tq[0]);
            lds_v2f64(a_pose + 80u, tq[1], tq[2]);
#else
#pragma unroll
            for (int i = 0; i < 9; ++i) Rq[i] = vR[i];
#pragma unroll
            for (int i = 0; i < 3; ++i) tq[i] = vt[i];
#endif
            Ln = gnw::linearise(C, Rq, tq, X0 + (double)d0, X1 + (double)d1, X2 + (double)d2, ob, stereo);
          }
          float Rl[9];
#if GMMLOC_GNS_VEC_LDS
          {
            const float4 r0 = lds_v4f32(a_rf), r1 = lds_v4f32(a_rf + 16u), r2 = lds_v4f32(a_rf + 32u);
            Rl[0] = r0.x; Rl[1] = r0.y; Rl[2] = r0.z; Rl[3] = r1.x; Rl[4] = r1.y; Rl[5] = r1.z; Rl[6] = r2.x; Rl[7] = r2.y; Rl[8] = r2.z;
          }
#else
          {
            const volatile float* vRf = s_new.Rf;
#pragma unroll
            for (int r = 0; r < 3; ++r)
#pragma unroll
              for (int c = 0; c < 3; ++c) Rl[r * 3 + c] = vRf[r * 4 + c];
          }
#endif
          if (CLS == C_FIXED) {
            S3f G;
            V3f h;
            gnw::lin_fixed(C, Ln, Rl, stereo, w, G, h);
            gnw::accumulate_pose(acc, xr, G, h);
          } else if (CLS == C_PLAIN) {
            const V3f u0 = gnw::lin_plain(C, Ln, Rl);
            float* gf = reinterpret_cast<float*>(gst);
            __stcg(gf + lane, u0.x);
            __stcg(gf + 32 + lane, u0.y);
            __stcg(gf + 64 + lane, u0.z);
          } else if (CLS == C_MPLANE || CLS == C_STRP) {
            const float4 c0v = cst4[lane];  // (n, es0)
            const V3f n = {c0v.x, c0v.y, c0v.z};
            const float es = c0v.w + (n.x * d0 + n.y * d1 + n.z * d2);
            PlaneLin pl;
            if (CLS == C_STRP) {
              float kappa, rho;
              if (!gnw::lin_strp_c(C, Ln, Rl, w, n, es, pl, kappa, rho)) bad |= 1;
              if (do_acc) gnw::accumulate_pose_rank1(acc, xr, n, kappa, rho);
            } else {
              if (!gnw::lin_mplane_c(C, Ln, Rl, n, es, pl)) bad |= 2;
            }
            if (from_pack) __stcg(reinterpret_cast<float4*>(gs) + lane, c0v);  // constants into the slice, once per frame
            __stcg(reinterpret_cast<float4*>(gst) + lane, make_float4(pl.u0.x, pl.u0.y, pl.u0.z, pl.kz.x));
            __stcg(reinterpret_cast<float2*>(gst + 512) + lane, make_float2(pl.kz.y, pl.kz.z));
          } else {  // C_STR / C_MSTR
            const float4 c0v = cst4[lane], c1v = cst4[32 + lane];
            const float c2v = reinterpret_cast<const float*>(slot + 1024)[lane];
            const V3f d = {c0v.x + d0, c0v.y + d1, c0v.z + d2};
            const S3f Sw = {c0v.w, c1v.x, c1v.y, c1v.z, c1v.w, c2v};
            PointLin pl;
            S3f M;
            V3f v;
            // C_STR: Sw = Sigma^-1, general block;  C_MSTR: Sw = Sigma, gain form (frame_gn_world.cuh)
            const bool ok32 = CLS == C_STR ? gnw::lin_str(C, Ln, Rl, true, w, Sw, d, pl, do_acc, M, v)
                                           : gnw::lin_mstr_gain(C, Ln, Rl, w, Sw, d, pl, M, v);
            if (!ok32) bad |= 1;
            if (do_acc) gnw::accumulate_pose(acc, xr, M, v);
            if (from_pack) {  // constants into the slice, once per frame
              __stcg(reinterpret_cast<float4*>(gs) + lane, c0v);
              __stcg(reinterpret_cast<float4*>(gs) + 32 + lane, c1v);
              __stcg(reinterpret_cast<float*>(gs + 1024) + lane, c2v);
            }
            __stcg(reinterpret_cast<float4*>(gst) + lane, make_float4(pl.u0.x, pl.u0.y, pl.u0.z, pl.w0));
            __stcg(reinterpret_cast<float4*>(gst) + 32 + lane, make_float4(pl.w1, pl.w2, pl.w3, pl.w4));
            __stcg(reinterpret_cast<float4*>(gst) + 64 + lane, make_float4(pl.w5, pl.w6, pl.w7, pl.w8));
          }
        }
      };
      if (warp < nch) issue(warp);
      for (int g = warp; g < nch; g += kWarps) {
        if (g + kWarps < nch) issue(g + kWarps);
        const uint32_t ent = s_chunk[g];
        const int cls = (int)(ent >> 24);
        const unsigned char* slot = ring + (n_con & 1u) * kSlotBytes;
        if (lut_get(lut_n, cls) != 0u) {
          mbar_wait_a(bar_a + (n_con & 1u) * 8u, (n_con >> 1) & 1u);
          ++n_con;
        }
        const bool work = (cls == C_FIXED) ? do_acc : (do_update || do_lin);
        if (work && lane < (int)((ent >> 16) & 0xffu)) {
          const int j = (int)(ent & 0xffffu) + lane;
          unsigned char* gs = gstore + (size_t)g * kSlotBytes;
          switch (cls) {
            case C_FIXED: point(std::integral_constant<int, C_FIXED>{}, j, slot, gs); break;
            case C_PLAIN: point(std::integral_constant<int, C_PLAIN>{}, j, slot, gs); break;
            case C_MPLANE: point(std::integral_constant<int, C_MPLANE>{}, j, slot, gs); break;
            case C_STRP: point(std::integral_constant<int, C_STRP>{}, j, slot, gs); break;
            case C_STR: point(std::integral_constant<int, C_STR>{}, j, slot, gs); break;
            default: point(std::integral_constant<int, C_MSTR>{}, j, slot, gs); break;
          }
        }
        __syncwarp();  // every lane is done with this slot before lane 0 lets the next copy overwrite it
      }

      // ---- fp64 points (last class of the sorted frame), dealt round robin over the warps ----
      // (from the LAST warp backwards: the chunks are dealt from warp 0, so the last warps are the ones with a chunk less)
      for (int q = lane * kWarps + (kWarps - 1 - warp); q < n64; q += 32 * kWarps) {
        const int j = f64_begin + q;
        const uint8_t b = fl[j];
        const bool upd = do_update, accp = do_acc && (b & F_STR);
        if (!upd && !accp) continue;
        const double X0 = (double)x0s[j * 3], X1 = (double)x0s[j * 3 + 1], X2 = (double)x0s[j * 3 + 2];
        double x[3] = {X0 + (double)dl[j * 3], X1 + (double)dl[j * 3 + 1], X2 + (double)dl[j * 3 + 2]};
        const float ob[3] = {so[j * 3], so[j * 3 + 1], so[j * 3 + 2]};
        const int k = (b & F_STR) ? a.assign[(size_t)f * N + gperm[j]] : -1;
        V3f pcf, vf;
        S3f Mf;
        point_fp64_s(a, s_old.R, s_old.t, s_new.R, s_new.t, s_dxc, b & 15, k, upd, accp, ob, (double)s_w[soct[j] & 7], x, pcf, Mf, vf,
                     bad);
        if (upd) {
          dl[j * 3] = (float)(x[0] - X0);
          dl[j * 3 + 1] = (float)(x[1] - X1);
          dl[j * 3 + 2] = (float)(x[2] - X2);
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
      if (bad & 2) s_fail2 = 1;
      if (bad & 1) s_fail1 = 1;
      if (do_acc) {  // lanes in fp32 (fixed halving order), warps in fp64
        const float tot = warp_reduce27s(acc, lane);
        if (lane < kAccS) s_red[warp][lane] = (double)tot;
      }
      if (do_lin && pts_free) fence_proxy_async_global();  // this sweep's stores before the next sweep's TMA reads
      __syncthreads();
      if (s_fail2) break;  // uniform
      if (do_acc) {
        if (warp == 0) {  // fixed-order cross-warp sum, then the 6x6 solve and the SE3 update on lane 0
          if (lane < kAccS) {
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
            for (int i = 0; i < kAccS; ++i) okf = okf && isfinite(s_acc[i]);
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
    if (tid == 0) {
      a.status[f] = fail ? GMMLOC_FRAME_NUMERIC : 0;
      // a frame the fp32 blocks could not carry (a landmark at the camera plane or at infinity in an iteration that does not
      // contract ...) is solved again, from its inputs, by the all-fp64 kernel after this launch: ITS verdict is the status
      if (fail) counters[2 + atomicAdd(counters + 1, 1)] = f;
    }
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
        const int i = gperm[j];
        Vec3 xw;
        if (fail) xw = {(double)pts[i * 3], (double)pts[i * 3 + 1], (double)pts[i * 3 + 2]};
        else xw = {(double)x0s[j * 3] + (double)dl[j * 3], (double)x0s[j * 3 + 1] + (double)dl[j * 3 + 1],
                   (double)x0s[j * 3 + 2] + (double)dl[j * 3 + 2]};
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

size_t s_dyn_bytes(int N, int threads) {
  const PackLayout L = pack_layout(N);
  return L.state_bytes + al16(12 * (size_t)N) + (size_t)(threads / 32) * 2 * kSlotBytes;
}

template <int kThreads, int kMinB>
cudaError_t launch_stream_t(const FrameArgs& a, int* d_counters, unsigned char* store_ws, size_t store_ws_bytes,
                            const unsigned char* pack, cudaStream_t stream, bool* fits) {
  cudaError_t e;
  const size_t smem = s_dyn_bytes(a.N, kThreads);
  // the opt-in is state shared by every host thread launching this variant: always the SAME value (see K3m's launcher)
  cudaFuncAttributes fa;
  if ((e = cudaFuncGetAttributes(&fa, frame_gn_stream_kernel<kThreads, kMinB>)) != cudaSuccess) return e;
  const size_t max_dyn = kMinB == 2 ? ((228 * 1024 / 2 - 1024 - fa.sharedSizeBytes) & ~(size_t)15)
                                    : ((227 * 1024 - fa.sharedSizeBytes) & ~(size_t)15);
  if (smem > max_dyn || pack_max_chunks(a.N) > kMaxNC) {
    *fits = false;
    return cudaSuccess;
  }
  if ((e = cudaFuncSetAttribute(frame_gn_stream_kernel<kThreads, kMinB>, cudaFuncAttributeMaxDynamicSharedMemorySize, (int)max_dyn)) !=
      cudaSuccess)
    return e;
  int per_sm = 0;
  if ((e = cudaOccupancyMaxActiveBlocksPerMultiprocessor(&per_sm, frame_gn_stream_kernel<kThreads, kMinB>, kThreads, smem)) != cudaSuccess)
    return e;
  if (per_sm < 1) per_sm = 1;
  int dev = 0, sms = 148;
  cudaGetDevice(&dev);
  cudaDeviceGetAttribute(&sms, cudaDevAttrMultiProcessorCount, dev);
  const int grid = std::min(a.B, sms * per_sm);
  if ((size_t)grid * stream_slice(a.N) > store_ws_bytes) {
    *fits = false;
    return cudaSuccess;
  }
  *fits = true;
  if ((e = cudaMemsetAsync(d_counters, 0, 2 * sizeof(int), stream)) != cudaSuccess) return e;
  frame_gn_stream_kernel<kThreads, kMinB><<<grid, kThreads, smem, stream>>>(a, d_counters, store_ws, pack);
  if ((e = cudaGetLastError()) != cudaSuccess) return e;
  // the frames the kernel gave up on (normally none: the CTAs of this launch return at once), re-solved in fp64
  FrameArgs r = a;
  r.retry_count = d_counters + 1;
  r.retry_list = d_counters + 2;
  return launch_frame_gn(r, stream);
}

}  // namespace

size_t frame_gn_stream_limit_points() {
  // one resident frame of 384 threads: state 38 B per point + 12 rings next to ~4 KB of static shared memory
  const size_t by_smem = (size_t)(227 * 1024 - 4096 - 12 * 2 * kSlotBytes - 64) / 38;
  const size_t by_chunks = (size_t)(kMaxNC - 8) * 32;
  return std::min(std::min(by_smem, by_chunks), (size_t)(8 * kMaxChunks - 12) * 32);  // the histogram of the pack kernel bounds N too
}

size_t frame_gn_stream_scratch_bytes(int N) {
  int dev = 0, sms = 148;
  cudaGetDevice(&dev);
  cudaDeviceGetAttribute(&sms, cudaDevAttrMultiProcessorCount, dev);
  return (size_t)2 * sms * stream_slice(N) + 16;
}

// Batches beyond kSubBatch frames run as consecutive sub-batches on the stream, so the packed-frame buffer is bounded
// (kSubBatch x ~69 B per landmark: 1.1 GB at N = 2 000) however large B is.
constexpr int kSubBatch = 8192;
size_t frame_gn_stream_prep_bytes(int B, int N) {
  const size_t b = (size_t)std::min(B, kSubBatch);
  return b * pack_layout(N).stride + b * N + b * 64 + 64;
}

static cudaError_t launch_stream_sub(const FrameArgs& a, const float* gn_tab, int* d_counters, unsigned char* st, size_t st_bytes,
                                     unsigned char* pack, int sms, cudaStream_t stream) {
  const size_t need = (size_t)a.B * pack_layout(a.N).stride;
  uint8_t* flags = pack + need;  // B x N class bytes behind the packed frames, then 8 doubles per frame
#if GMMLOC_GNS_SPLIT_CLASSIFY
  double* fpose = reinterpret_cast<double*>(flags + (((size_t)a.B * a.N + 15) & ~(size_t)15));
  frame_pose_kernel<<<(a.B + 127) / 128, 128, 0, stream>>>(a, fpose);
  {
    const long long total = (long long)a.B * a.N;
    const int grid = (int)std::min<long long>((total + 255) / 256, (long long)sms * 64);
    frame_classify_kernel<<<grid, 256, 0, stream>>>(a, fpose, flags);
  }
#else
  flags = nullptr;  // classification inside the pack kernel
#endif
  frame_pack_kernel<<<std::min(a.B, sms * GMMLOC_PACK_GRID), kPackThreads, (size_t)((a.N + 15) & ~15), stream>>>(a, gn_tab, flags, pack);
  cudaError_t e = cudaGetLastError();
  if (e != cudaSuccess) return e;
  bool fits = false;
  e = launch_stream_t<kSThreads2, 2>(a, d_counters, st, st_bytes, pack, stream, &fits);
  if (e != cudaSuccess || fits) return e;
  e = launch_stream_t<384, 1>(a, d_counters, st, st_bytes, pack, stream, &fits);
  if (e != cudaSuccess || fits) return e;
  return cudaErrorInvalidValue;  // the caller checks frame_gn_stream_limit_points() and sizes the scratch: not reached
}

// d_counters: min(a.B, kSubBatch) + 2 ints (frame counter, number of frames handed to the fp64 kernel, their indices)
cudaError_t launch_frame_gn_stream(const FrameArgs& a, const float* gn_tab, int* d_counters, float* store_ws, size_t store_ws_bytes,
                                   unsigned char* prep_ws, size_t prep_ws_bytes, cudaStream_t stream) {
  if (a.B <= 0) return cudaSuccess;
  if (prep_ws == nullptr || store_ws == nullptr || a.N <= 0) return cudaErrorInvalidValue;
  unsigned char* pack = reinterpret_cast<unsigned char*>((reinterpret_cast<uintptr_t>(prep_ws) + 15) & ~(uintptr_t)15);
  if ((size_t)(pack - prep_ws) + frame_gn_stream_prep_bytes(a.B, a.N) - 64 + 16 > prep_ws_bytes) return cudaErrorInvalidValue;
  unsigned char* st = reinterpret_cast<unsigned char*>((reinterpret_cast<uintptr_t>(store_ws) + 15) & ~(uintptr_t)15);
  const size_t st_bytes = store_ws_bytes - (size_t)(st - reinterpret_cast<unsigned char*>(store_ws));
  int dev = 0, sms = 148;
  cudaGetDevice(&dev);
  cudaDeviceGetAttribute(&sms, cudaDevAttrMultiProcessorCount, dev);
  for (int f0 = 0; f0 < a.B; f0 += kSubBatch) {
    FrameArgs s = a;
    s.B = std::min(kSubBatch, a.B - f0);
    const size_t m0 = (size_t)f0 * a.N;
    s.pose0 = a.pose0 + (size_t)f0 * 7;
    s.pts_w = a.pts_w + 3 * m0;
    s.obs = a.obs + 3 * m0;
    s.octave = a.octave + m0;
    s.assign = a.assign + m0;
    s.pose_out = a.pose_out + (size_t)f0 * 7;
    s.pts_out = a.pts_out ? a.pts_out + 3 * m0 : nullptr;
    s.outlier = a.outlier ? a.outlier + m0 : nullptr;
    s.status = a.status + f0;
    const cudaError_t e = launch_stream_sub(s, gn_tab, d_counters, st, st_bytes, pack, sms, stream);
    if (e != cudaSuccess) return e;
  }
  return cudaSuccess;
}

}  // namespace gmmloc
