// assoc_scan.cu — K1 (fp32 brute-force Mahalanobis scan) and K2 (exact fp64 re-rank).
//
// Replaces: GaussianComponent::chi2 / MDist2 (reference gmmloc/src/gmm/gaussian.cpp:65-70,
// gmmloc/include/gmmloc/gmm/gaussian.h:53-56) evaluated for every (landmark, component) pair with the
// reference's argmin convention (strict '<' on an ascending scan => lowest index wins ties,
// gmmloc/src/gmmloc_opt.cpp:175-197,209-217) and optional gate (gmmloc_opt.cpp:230).
//
// K1: point-stationary.  Each thread keeps P landmarks in registers (each coordinate duplicated into a
//     64-bit pair); the whitened component table (9 floats / component: U = L^T rows and -U mu, stored
//     4 components x 9 coefficients per group) streams through shared memory in TMA bulk-copy tiles
//     (cp.async.bulk + mbarrier, 3 stages) and is read with warp-broadcast LDS.128.  The score of TWO
//     components is evaluated per instruction with the packed fma.rn.f32x2 (SASS FFMA2): per (landmark,
//     component) pair 4.5 FFMA2 + 1/2 FMNMX3 issue slots (d2 = |U x - c|^2).  The running best / second
//     best are tracked per chunk of 16 components (5 ALU ops per chunk); the winning chunk is re-scanned
//     once at the end to recover the index and the in-chunk runner-up.
// Exactness: the scan certifies an assignment when sqrt(second) - sqrt(best) exceeds twice the
//     rigorous fp32 error bound alpha; everything else is pushed to a flag list.
// K2: one warp per flagged landmark replays the fp32 scores (bit-identical: same score32()), and
//     evaluates every candidate within the bound in fp64 with the oracle's operation order.
#include <math_constants.h>

#include "common.cuh"

namespace gmmloc {

namespace {

#ifndef GMMLOC_SCAN_P
#define GMMLOC_SCAN_P 8
#endif
#ifndef GMMLOC_SCAN_THREADS
#define GMMLOC_SCAN_THREADS 256
#endif
#ifndef GMMLOC_SCAN_MINB
#define GMMLOC_SCAN_MINB 2
#endif
constexpr int kStages = 3;
constexpr int kTileComps = 384;  // 384 * 36 B = 13.5 KB per stage
constexpr int kTileFloats = kTileComps * kScanFloats;
constexpr int kTileBytes = kTileFloats * 4;

__device__ __forceinline__ uint32_t smem_u32(const void* p) { return (uint32_t)__cvta_generic_to_shared(p); }

// ---- packed fp32x2 arithmetic (Blackwell FFMA2 / FMUL2): elementwise IEEE rn, so each half equals the
// scalar __fmaf_rn / __fmul_rn of score32() bit for bit ----
__device__ __forceinline__ uint64_t pack2(float lo, float hi) {
  uint64_t r;
  asm("mov.b64 %0, {%1, %2};" : "=l"(r) : "f"(lo), "f"(hi));
  return r;
}
__device__ __forceinline__ void unpack2(uint64_t v, float& lo, float& hi) {
  asm("mov.b64 {%0, %1}, %2;" : "=f"(lo), "=f"(hi) : "l"(v));
}
__device__ __forceinline__ uint64_t fma2(uint64_t a, uint64_t b, uint64_t c) {
  uint64_t d;
  asm("fma.rn.f32x2 %0, %1, %2, %3;" : "=l"(d) : "l"(a), "l"(b), "l"(c));
  return d;
}
__device__ __forceinline__ uint64_t mul2(uint64_t a, uint64_t b) {
  uint64_t d;
  asm("mul.rn.f32x2 %0, %1, %2;" : "=l"(d) : "l"(a), "l"(b));
  return d;
}
// scores of the two components held in the halves of the coefficient pairs, for one (duplicated) landmark
__device__ __forceinline__ uint64_t score2(const uint64_t* u, uint64_t X0, uint64_t X1, uint64_t X2) {
  const uint64_t y0 = fma2(u[0], X0, fma2(u[1], X1, fma2(u[2], X2, u[6])));
  const uint64_t y1 = fma2(u[3], X1, fma2(u[4], X2, u[7]));
  const uint64_t y2 = fma2(u[5], X2, u[8]);
  return fma2(y2, y2, fma2(y1, y1, mul2(y0, y0)));
}
// scalar score of component k read from the global table (epilogue / re-rank)
__device__ __forceinline__ float score_of(const float* __restrict__ scan, int k, float x0, float x1, float x2) {
  const float* f = scan + ((size_t)(k / kGroup) * kScanFloats) * kGroup + (k % kGroup);
  return score32(__ldg(f + 0), __ldg(f + 4), __ldg(f + 8), __ldg(f + 12), __ldg(f + 16), __ldg(f + 20), __ldg(f + 24),
                 __ldg(f + 28), __ldg(f + 32), x0, x1, x2);
}

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
// TMA 1-D bulk copy global -> shared, completion signalled on an mbarrier (SASS: UBLKCP)
__device__ __forceinline__ void tma_bulk_g2s(void* dst_smem, const void* src_gmem, uint32_t bytes, uint64_t* bar) {
  asm volatile("cp.async.bulk.shared::cluster.global.mbarrier::complete_tx::bytes [%0], [%1], %2, [%3];" ::"r"(
                   smem_u32(dst_smem)),
               "l"(src_gmem), "r"(bytes), "r"(smem_u32(bar))
               : "memory");
}

struct ScanArgs {
  const float* scan;  // Kpad x 9
  const double* mean;
  const double* cov_inv;
  int K, Kpad;
  float alpha, xmax;
  const float* pts;  // M x 3
  long long M;
  double gate;
  int32_t* assign;
  float* d2;
  int* flag_count;
  int32_t* flag_idx;
  float* flag_thr;
};

template <int P, int THREADS, int MINB>
__global__ void __launch_bounds__(THREADS, MINB) assoc_scan_kernel(const ScanArgs a) {
  extern __shared__ __align__(128) unsigned char smem_raw[];
  float* tiles = reinterpret_cast<float*>(smem_raw);
  uint64_t* full = reinterpret_cast<uint64_t*>(smem_raw + kStages * kTileBytes);

  const int tid = threadIdx.x;
  const long long base = (long long)blockIdx.x * (THREADS * P);
  const int ntiles = (a.Kpad + kTileComps - 1) / kTileComps;

  if (tid == 0) {
    for (int s = 0; s < kStages; ++s) mbar_init(&full[s], 1);
    asm volatile("fence.mbarrier_init.release.cluster;" ::: "memory");
    asm volatile("fence.proxy.async.shared::cta;" ::: "memory");
    for (int t = 0; t < kStages && t < ntiles; ++t) {
      const int ncomp = min(kTileComps, a.Kpad - t * kTileComps);
      const uint32_t bytes = (uint32_t)ncomp * kScanFloats * 4u;
      mbar_expect_tx(&full[t], bytes);
      tma_bulk_g2s(tiles + (size_t)t * kTileFloats, a.scan + (size_t)t * kTileFloats, bytes, &full[t]);
    }
  }

  uint64_t X0[P], X1[P], X2[P];  // landmark coordinates, duplicated into both halves
  float b1[P], b2[P];
  int cid[P];
#pragma unroll
  for (int p = 0; p < P; ++p) {
    const long long i = base + (long long)p * THREADS + tid;
    float x0 = 0.f, x1 = 0.f, x2 = 0.f;
    if (i < a.M) {
      x0 = __ldg(a.pts + i * 3 + 0);
      x1 = __ldg(a.pts + i * 3 + 1);
      x2 = __ldg(a.pts + i * 3 + 2);
    }
    X0[p] = pack2(x0, x0);
    X1[p] = pack2(x1, x1);
    X2[p] = pack2(x2, x2);
    b1[p] = CUDART_INF_F;
    b2[p] = CUDART_INF_F;
    cid[p] = 0;
  }
  __syncthreads();  // barrier inits visible before anyone waits

  for (int t = 0; t < ntiles; ++t) {
    const int stage = t % kStages;
    mbar_wait(&full[stage], (uint32_t)((t / kStages) & 1));
    const ulonglong2* sp = reinterpret_cast<const ulonglong2*>(tiles + (size_t)stage * kTileFloats);
    const int nchunks = min(kTileComps, a.Kpad - t * kTileComps) / kChunk;
    const int chunk0 = t * (kTileComps / kChunk);
    for (int c = 0; c < nchunks; ++c) {
      float mch[P];
#pragma unroll
      for (int p = 0; p < P; ++p) mch[p] = CUDART_INF_F;
#pragma unroll
      for (int q = 0; q < kChunk / kGroup; ++q) {
        // one group = 4 components = 9 broadcast LDS.128; .x = coefficient pair of components (0,1), .y = (2,3)
        uint64_t ua[9], ub[9];
#pragma unroll
        for (int v = 0; v < 9; ++v) {
          const ulonglong2 w = sp[(c * (kChunk / kGroup) + q) * 9 + v];
          ua[v] = w.x;
          ub[v] = w.y;
        }
#pragma unroll
        for (int p = 0; p < P; ++p) {
          float s0, s1, s2, s3;
          unpack2(score2(ua, X0[p], X1[p], X2[p]), s0, s1);
          unpack2(score2(ub, X0[p], X1[p], X2[p]), s2, s3);
          mch[p] = fminf(fminf(mch[p], s0), s1);  // FMNMX3
          mch[p] = fminf(fminf(mch[p], s2), s3);
        }
      }
      const int chunk = chunk0 + c;
#pragma unroll
      for (int p = 0; p < P; ++p) {
        const float mc = mch[p];
        b2[p] = fminf(b2[p], fmaxf(b1[p], mc));
        cid[p] = (mc < b1[p]) ? chunk : cid[p];
        b1[p] = fminf(b1[p], mc);
      }
    }
    __syncthreads();  // every warp is done with this stage
    if (tid == 0 && t + kStages < ntiles) {
      const int tn = t + kStages;
      const int ncomp = min(kTileComps, a.Kpad - tn * kTileComps);
      const uint32_t bytes = (uint32_t)ncomp * kScanFloats * 4u;
      mbar_expect_tx(&full[stage], bytes);
      tma_bulk_g2s(tiles + (size_t)stage * kTileFloats, a.scan + (size_t)tn * kTileFloats, bytes, &full[stage]);
    }
  }

  // ---- epilogue: recover index + in-chunk runner-up, certify or flag ----
#pragma unroll
  for (int p = 0; p < P; ++p) {
    const long long i = base + (long long)p * THREADS + tid;
    if (i >= a.M) continue;
    float px, py, pz, dup;
    unpack2(X0[p], px, dup);
    unpack2(X1[p], py, dup);
    unpack2(X2[p], pz, dup);
    const bool finite_in = (fabsf(px) <= a.xmax) && (fabsf(py) <= a.xmax) && (fabsf(pz) <= a.xmax);  // false for NaN
    int jbest = -1;
    float second = CUDART_INF_F;
    for (int j = 0; j < kChunk; ++j) {
      const float s = score_of(a.scan, cid[p] * kChunk + j, px, py, pz);
      if (jbest < 0 && s == b1[p]) {
        jbest = j;
      } else {
        second = fminf(second, s);
      }
    }
    const float sec = fminf(b2[p], second);
    bool certified = finite_in && (jbest >= 0);
    int idx = cid[p] * kChunk + (jbest < 0 ? 0 : jbest);
    double rt = CUDART_INF;
    if (certified) {
      const double r1 = sqrt((double)b1[p]), r2 = sqrt((double)sec);
      const double hi1 = r1 * (1.0 + kRelEps) + (double)a.alpha;  // upper bound of the winner's true distance
      const double lo2 = r2 * (1.0 - kRelEps) - (double)a.alpha;  // lower bound of every other component's
      certified = lo2 > hi1;
      rt = (hi1 + (double)a.alpha) / (1.0 - kRelEps);
    }
    if (certified) {
      if (a.d2 != nullptr || a.gate > 0.0) {
        const double d = chi2_exact(a.mean + (size_t)idx * 3, a.cov_inv + (size_t)idx * 9, (double)px, (double)py, (double)pz);
        if (a.gate > 0.0 && d > a.gate) idx = -1;
        if (a.d2 != nullptr) a.d2[i] = (float)d;
      }
      a.assign[i] = idx;
    } else {
      float thr = CUDART_INF_F;
      if (finite_in && jbest >= 0) {
        const double t2 = rt * rt;
        thr = (float)t2;
        if ((double)thr < t2) thr = nextafterf(thr, CUDART_INF_F);
      }
      const int pos = atomicAdd(a.flag_count, 1);
      a.flag_idx[pos] = (int32_t)i;  // flag list is sized M; callers keep M < 2^31
      a.flag_thr[pos] = thr;
    }
  }
}

// K2: exact re-rank of the flagged landmarks.  Same point-stationary structure as K1 (P flagged landmarks
// per thread, table streamed through the TMA tile pipeline, FFMA2 chunk minima), persistent over the
// flag list whose length is only known on the device.  A chunk whose fp32 minimum is within the
// landmark's threshold takes the slow path: every component of the chunk whose fp32 score (bit-identical
// to K1's) is within the threshold is evaluated in fp64 with the oracle's operation order; strict '<' on
// an ascending scan keeps the lowest index, exactly like the reference's loop (gmmloc_opt.cpp:209-217).
template <int P, int THREADS>
__global__ void __launch_bounds__(THREADS) assoc_rerank_kernel(const ScanArgs a) {
  extern __shared__ __align__(128) unsigned char smem_raw[];
  float* tiles = reinterpret_cast<float*>(smem_raw);
  uint64_t* full = reinterpret_cast<uint64_t*>(smem_raw + kStages * kTileBytes);

  const int tid = threadIdx.x;
  const int nflag = *a.flag_count;
  const int per = THREADS * P;
  if ((long long)blockIdx.x * per >= nflag) return;
  const int ntiles = (a.Kpad + kTileComps - 1) / kTileComps;
  if (tid == 0) {
    for (int s = 0; s < kStages; ++s) mbar_init(&full[s], 1);
    asm volatile("fence.mbarrier_init.release.cluster;" ::: "memory");
    asm volatile("fence.proxy.async.shared::cta;" ::: "memory");
  }
  __syncthreads();

  int g = 0;  // tiles consumed so far by this CTA (stage = g % kStages, parity = (g / kStages) & 1)
  for (long long base = (long long)blockIdx.x * per; base < nflag; base += (long long)gridDim.x * per) {
    if (tid == 0) {
      for (int t = 0; t < kStages && t < ntiles; ++t) {
        const int st = (g + t) % kStages;
        const int ncomp = min(kTileComps, a.Kpad - t * kTileComps);
        const uint32_t bytes = (uint32_t)ncomp * kScanFloats * 4u;
        mbar_expect_tx(&full[st], bytes);
        tma_bulk_g2s(tiles + (size_t)st * kTileFloats, a.scan + (size_t)t * kTileFloats, bytes, &full[st]);
      }
    }
    uint64_t X0[P], X1[P], X2[P];
    float thr[P];
    double best[P];
    int bi[P];
    long long pi[P];
#pragma unroll
    for (int p = 0; p < P; ++p) {
      const long long w = base + (long long)p * THREADS + tid;
      float x0 = 0.f, x1 = 0.f, x2 = 0.f;
      pi[p] = -1;
      thr[p] = -1.f;  // nothing passes: idle slot
      if (w < nflag) {
        pi[p] = a.flag_idx[w];
        thr[p] = a.flag_thr[w];
        x0 = __ldg(a.pts + pi[p] * 3 + 0);
        x1 = __ldg(a.pts + pi[p] * 3 + 1);
        x2 = __ldg(a.pts + pi[p] * 3 + 2);
      }
      X0[p] = pack2(x0, x0);
      X1[p] = pack2(x1, x1);
      X2[p] = pack2(x2, x2);
      best[p] = CUDART_INF;  // stands for numeric_limits<double>::max(): any finite d2 is smaller
      bi[p] = -1;
    }
    for (int t = 0; t < ntiles; ++t) {
      const int stage = (g + t) % kStages;
      mbar_wait(&full[stage], (uint32_t)(((g + t) / kStages) & 1));
      const float* tf = tiles + (size_t)stage * kTileFloats;
      const ulonglong2* sp = reinterpret_cast<const ulonglong2*>(tf);
      const int nchunks = min(kTileComps, a.Kpad - t * kTileComps) / kChunk;
      const int chunk0 = t * (kTileComps / kChunk);
      for (int c = 0; c < nchunks; ++c) {
        float mch[P];
#pragma unroll
        for (int p = 0; p < P; ++p) mch[p] = CUDART_INF_F;
#pragma unroll
        for (int q = 0; q < kChunk / kGroup; ++q) {
          uint64_t ua[9], ub[9];
#pragma unroll
          for (int v = 0; v < 9; ++v) {
            const ulonglong2 w = sp[(c * (kChunk / kGroup) + q) * 9 + v];
            ua[v] = w.x;
            ub[v] = w.y;
          }
#pragma unroll
          for (int p = 0; p < P; ++p) {
            float s0, s1, s2, s3;
            unpack2(score2(ua, X0[p], X1[p], X2[p]), s0, s1);
            unpack2(score2(ub, X0[p], X1[p], X2[p]), s2, s3);
            mch[p] = fminf(fminf(mch[p], s0), s1);
            mch[p] = fminf(fminf(mch[p], s2), s3);
          }
        }
#pragma unroll
        for (int p = 0; p < P; ++p) {
          if (pi[p] >= 0 && !(mch[p] > thr[p])) {  // rare: a candidate (or a non-finite landmark) in this chunk
            float px, py, pz, dup;
            unpack2(X0[p], px, dup);
            unpack2(X1[p], py, dup);
            unpack2(X2[p], pz, dup);
            for (int j = 0; j < kChunk; ++j) {
              const int k = (chunk0 + c) * kChunk + j;
              if (k >= a.K) break;
              const float* f = tf + ((size_t)(c * (kChunk / kGroup) + j / kGroup) * kScanFloats) * kGroup + (j % kGroup);
              const float s = score32(f[0], f[4], f[8], f[12], f[16], f[20], f[24], f[28], f[32], px, py, pz);
              if (!(s > thr[p])) {
                const double d = chi2_exact(a.mean + (size_t)k * 3, a.cov_inv + (size_t)k * 9, (double)px, (double)py, (double)pz);
                if (d < best[p]) {
                  best[p] = d;
                  bi[p] = k;
                }
              }
            }
          }
        }
      }
      __syncthreads();  // every warp is done with this stage
      if (tid == 0 && t + kStages < ntiles) {
        const int tn = t + kStages;
        const int ncomp = min(kTileComps, a.Kpad - tn * kTileComps);
        const uint32_t bytes = (uint32_t)ncomp * kScanFloats * 4u;
        mbar_expect_tx(&full[stage], bytes);
        tma_bulk_g2s(tiles + (size_t)stage * kTileFloats, a.scan + (size_t)tn * kTileFloats, bytes, &full[stage]);
      }
    }
    g += ntiles;
#pragma unroll
    for (int p = 0; p < P; ++p) {
      if (pi[p] < 0) continue;
      int idx = bi[p];
      if (a.gate > 0.0 && idx >= 0 && best[p] > a.gate) idx = -1;
      a.assign[pi[p]] = idx;
      if (a.d2 != nullptr) a.d2[pi[p]] = (bi[p] >= 0) ? (float)best[p] : CUDART_INF_F;
    }
  }
}

template <int P, int THREADS, int MINB>
cudaError_t launch_scan_t(const ScanArgs& a, cudaStream_t stream) {
  const size_t smem = (size_t)kStages * kTileBytes + kStages * sizeof(uint64_t);
  cudaError_t e = cudaFuncSetAttribute(assoc_scan_kernel<P, THREADS, MINB>, cudaFuncAttributeMaxDynamicSharedMemorySize, (int)smem);
  if (e != cudaSuccess) return e;
  const long long per_cta = (long long)THREADS * P;
  const long long grid = (a.M + per_cta - 1) / per_cta;
  assoc_scan_kernel<P, THREADS, MINB><<<(unsigned)grid, THREADS, smem, stream>>>(a);
  return cudaGetLastError();
}

}  // namespace

// Launches K1 + K2 on `stream`.  d_flag_count must be zeroed by the caller (memset on the same stream);
// flag buffers hold M entries.  Returns the number of kernels launched through *launches.
cudaError_t launch_associate(const MapDev& m, const float* d_pts, long long M, double gate, int32_t* d_assign,
                             float* d_d2, int* d_flag_count, int32_t* d_flag_idx, float* d_flag_thr, int variant,
                             cudaStream_t stream, int* launches, cudaEvent_t ev_between) {
  ScanArgs a;
  a.scan = m.scan;
  a.mean = m.mean;
  a.cov_inv = m.cov_inv;
  a.K = m.K;
  a.Kpad = m.Kpad;
  a.alpha = m.alpha;
  a.xmax = m.xmax;
  a.pts = d_pts;
  a.M = M;
  a.gate = gate;
  a.assign = d_assign;
  a.d2 = d_d2;
  a.flag_count = d_flag_count;
  a.flag_idx = d_flag_idx;
  a.flag_thr = d_flag_thr;
  cudaError_t e;
  // Fill the machine: 148 SMs x 2 CTAs; shrink the per-thread landmark count for small batches.
  int P = variant;
  if (P <= 0) {
    if (M >= 148LL * GMMLOC_SCAN_MINB * GMMLOC_SCAN_THREADS * GMMLOC_SCAN_P) P = 8;
    else if (M >= 148LL * 2 * 256 * 4) P = 4;
    else if (M >= 148LL * 2 * 128 * 2) P = 2;
    else P = 1;
  }
  switch (P) {
    case 8: e = launch_scan_t<GMMLOC_SCAN_P, GMMLOC_SCAN_THREADS, GMMLOC_SCAN_MINB>(a, stream); break;
    case 4: e = launch_scan_t<4, 256, 2>(a, stream); break;
    case 2: e = launch_scan_t<2, 128, 4>(a, stream); break;
    default: e = launch_scan_t<1, 128, 4>(a, stream); break;
  }
  if (e != cudaSuccess) return e;
  if (ev_between != nullptr && (e = cudaEventRecord(ev_between, stream)) != cudaSuccess) return e;
  {
    const size_t smem = (size_t)kStages * kTileBytes + kStages * sizeof(uint64_t);
    static bool attr_set = false;  // idempotent; a benign race between host threads
    if (!attr_set) {
      if ((e = cudaFuncSetAttribute(assoc_rerank_kernel<4, 128>, cudaFuncAttributeMaxDynamicSharedMemorySize, (int)smem)) != cudaSuccess) return e;
      attr_set = true;
    }
    // persistent over the flag list (length known only on the device); no more CTAs than the list can use
    const long long max_ctas = (M + 511) / 512;
    const unsigned grid = (unsigned)(max_ctas < 148 * 4 ? max_ctas : 148 * 4);
    assoc_rerank_kernel<4, 128><<<grid, 128, smem, stream>>>(a);
  }
  if (launches) *launches = 2;
  return cudaGetLastError();
}

}  // namespace gmmloc
