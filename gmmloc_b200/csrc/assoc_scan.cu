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
#include <cuda_fp16.h>
#include <math_constants.h>

#include <cstdio>
#include <cstdlib>
#include <cstring>

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
  // association grid kernels only: list A = landmarks whose exact re-rank can stay inside a grid cell list
  GridDev grid;
  const float4* rec;
  int* ga_count;
  int32_t* ga_idx;
  float* ga_thr;
  const float4* sorted;  // cell-sorted landmarks of the grid scan: (x, y, z, original index as bits), or null: natural order
  int rerank_two_pass;  // K2: first find the fp32 minimum over all components, then re-rank inside ITS radius (the
                        // thresholds of the grid path come from a cell-list candidate that may be far from the winner)
  // tensor-core pruning scan only
  const uint16_t* tc_table;
  float cen[3];
  float tc_err;    // bound of |sqrt(score_tc) - true distance|
  float tc_delta;  // candidate margin in sqrt(score) units
  float tc_slack;  // tc_delta - 2 (tc_err + alpha): guaranteed gap of every non-candidate component
};

// Certificate on exact fp32 scores: s1 = score32 of the candidate winner idx, sec = a lower bound of
// the score32 of every other component.  Certified => the assignment is written; otherwise the landmark goes to the
// exact re-rank list with the squared radius inside which K2 must evaluate in fp64.
__device__ __forceinline__ void finish_landmark(const ScanArgs& a, long long i, float px, float py, float pz, bool have_best,
                                                int idx, float s1, float sec) {
  bool certified = have_best;
  double rt = CUDART_INF;
  if (certified) {
    const double r1 = sqrt((double)s1), r2 = sqrt((double)sec);
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
    if (have_best) {
      const double t2 = rt * rt;
      thr = (float)t2;
      if ((double)thr < t2) thr = nextafterf(thr, CUDART_INF_F);
    }
    const int pos = atomicAdd(a.flag_count, 1);
    a.flag_idx[pos] = (int32_t)i;  // flag list is sized M; callers keep M < 2^31
    a.flag_thr[pos] = thr;
  }
}

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
    finish_landmark(a, i, px, py, pz, finite_in && (jbest >= 0), cid[p] * kChunk + (jbest < 0 ? 0 : jbest), b1[p],
                    fminf(b2[p], second));
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
  // Slot p of CTA b in round r takes the block of THREADS flagged landmarks number (r P + p) gridDim.x + b: a short flag
  // list is spread over ALL the CTAs of the launch with one active slot each (a CTA pays for its active slots only),
  // instead of filling the P slots of the first few CTAs while the other SMs idle.
  if ((long long)blockIdx.x * THREADS >= nflag) return;
  const int ntiles = (a.Kpad + kTileComps - 1) / kTileComps;
  if (tid == 0) {
    for (int s = 0; s < kStages; ++s) mbar_init(&full[s], 1);
    asm volatile("fence.mbarrier_init.release.cluster;" ::: "memory");
    asm volatile("fence.proxy.async.shared::cta;" ::: "memory");
  }
  __syncthreads();

  int g = 0;  // tiles consumed so far by this CTA (stage = g % kStages, parity = (g / kStages) & 1)
  for (long long round0 = 0; (round0 * gridDim.x + blockIdx.x) * THREADS < nflag; round0 += P) {
    int pact = 0;  // active slots of this CTA in this round (uniform over the CTA)
#pragma unroll
    for (int p = 0; p < P; ++p) pact += (((round0 + p) * gridDim.x + blockIdx.x) * THREADS < nflag) ? 1 : 0;
    if (tid == 0) {
      const int total0 = a.rerank_two_pass ? 2 * ntiles : ntiles;
      for (int q = 0; q < kStages && q < total0; ++q) {
        const int st = (g + q) % kStages;
        const int t = q % ntiles;
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
    bool finite_pt[P];
#pragma unroll
    for (int p = 0; p < P; ++p) {
      const long long w = ((round0 + p) * gridDim.x + blockIdx.x) * THREADS + tid;
      float x0 = 0.f, x1 = 0.f, x2 = 0.f;
      pi[p] = -1;
      thr[p] = -1.f;  // nothing passes: idle slot
      finite_pt[p] = false;
      if (w < nflag) {
        pi[p] = a.flag_idx[w];
        thr[p] = a.flag_thr[w];
        x0 = __ldg(a.pts + pi[p] * 3 + 0);
        x1 = __ldg(a.pts + pi[p] * 3 + 1);
        x2 = __ldg(a.pts + pi[p] * 3 + 2);
        finite_pt[p] = (fabsf(x0) <= a.xmax) && (fabsf(x1) <= a.xmax) && (fabsf(x2) <= a.xmax);  // alpha holds inside the box only
      }
      X0[p] = pack2(x0, x0);
      X1[p] = pack2(x1, x1);
      X2[p] = pack2(x2, x2);
      best[p] = CUDART_INF;  // stands for numeric_limits<double>::max(): any finite d2 is smaller
      bi[p] = -1;
    }
    // one or two passes over the table: tile tt of the sequence is table tile tt % ntiles
    const int total = a.rerank_two_pass ? 2 * ntiles : ntiles;
    float fmin32[P];
#pragma unroll
    for (int p = 0; p < P; ++p) fmin32[p] = CUDART_INF_F;
    for (int tt = 0; tt < total; ++tt) {
      const int t = tt % ntiles;
      const bool first_of_two = a.rerank_two_pass && tt < ntiles;
      if (a.rerank_two_pass && tt == ntiles) {
        // the fp32 minimum over ALL components bounds the winner's true distance: shrink the radius to it
#pragma unroll
        for (int p = 0; p < P; ++p) {
          if (pi[p] >= 0 && fmin32[p] < CUDART_INF_F && thr[p] == thr[p]) {
            const double rt = (sqrt((double)fmin32[p]) * (1.0 + kRelEps) + 2.0 * (double)a.alpha) / (1.0 - kRelEps);
            const double t2 = rt * rt;
            float tn = (float)t2;
            if ((double)tn < t2) tn = nextafterf(tn, CUDART_INF_F);
            if (finite_pt[p]) thr[p] = fminf(thr[p], tn);
          }
        }
      }
      const int stage = (g + tt) % kStages;
      mbar_wait(&full[stage], (uint32_t)(((g + tt) / kStages) & 1));
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
            if (p >= pact) continue;
            float s0, s1, s2, s3;
            unpack2(score2(ua, X0[p], X1[p], X2[p]), s0, s1);
            unpack2(score2(ub, X0[p], X1[p], X2[p]), s2, s3);
            mch[p] = fminf(fminf(mch[p], s0), s1);
            mch[p] = fminf(fminf(mch[p], s2), s3);
          }
        }
        if (first_of_two) {
#pragma unroll
          for (int p = 0; p < P; ++p) fmin32[p] = fminf(fmin32[p], mch[p]);
          continue;
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
      if (tid == 0 && tt + kStages < total) {
        const int tn = (tt + kStages) % ntiles;
        const int ncomp = min(kTileComps, a.Kpad - tn * kTileComps);
        const uint32_t bytes = (uint32_t)ncomp * kScanFloats * 4u;
        mbar_expect_tx(&full[stage], bytes);
        tma_bulk_g2s(tiles + (size_t)stage * kTileFloats, a.scan + (size_t)tn * kTileFloats, bytes, &full[stage]);
      }
    }
    g += total;
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


// ======================================================================================================
// K1-G / K2-G: association through the exact grid index (GridDev, grid_build.cu).
//
// One thread per landmark.  The landmark's cell list holds every component whose d2 can be <= tmax in that cell, so
// after scoring the list (plus the short `big` list) with the SAME score32() as the brute-force scan:
//   * winner's true distance provably <= sqrt(tmax)  =>  every unlisted component is farther than the winner: the
//     usual certificate (fp32 gap vs the error bound alpha) on best / second-best of the list decides, exactly as
//     in K1; near-ties go to list A and are re-ranked in fp64 over the same cell list (K2-G), lowest index first;
//   * otherwise the next (coarser, larger tmax) level is tried; what no level can decide — landmarks outside the
//     grid, non-finite ones, winners farther than the last tmax — goes to list B, the brute-force exact re-rank (K2).
// Results are therefore identical to the brute-force scan's; the work per landmark drops from K to ~30 scores.
__device__ __forceinline__ float score_rec(const float4* __restrict__ rec, int k, float x0, float x1, float x2) {
  const float4 r0 = __ldg(rec + 3 * (size_t)k), r1 = __ldg(rec + 3 * (size_t)k + 1);
  const float r2 = __ldg(reinterpret_cast<const float*>(rec + 3 * (size_t)k + 2));
  return score32(r0.x, r0.y, r0.z, r0.w, r1.x, r1.y, r1.z, r1.w, r2, x0, x1, x2);
}
__device__ __forceinline__ long long grid_cell(const GridLevel& G, float x0, float x1, float x2) {
  const float fx = __fmul_rn(__fsub_rn(x0, G.lo[0]), G.inv_h), fy = __fmul_rn(__fsub_rn(x1, G.lo[1]), G.inv_h),
              fz = __fmul_rn(__fsub_rn(x2, G.lo[2]), G.inv_h);
  if (!(fx >= 0.f && fy >= 0.f && fz >= 0.f && fx < (float)G.nx && fy < (float)G.ny && fz < (float)G.nz)) return -1;  // NaN too
  return ((long long)(int)fz * G.ny + (int)fy) * G.nx + (int)fx;
}

// ---- landmark sort by level-0 cell (counting sort: histogram -> exclusive scan -> scatter).  With the landmarks of
// a launch ordered by cell, a warp of the grid scan shares one cell list: list ids and component records are loaded
// once per warp (broadcast, one L1 wavefront per request) instead of once per lane (the unsorted kernel is
// L1-wavefront-bound at 8.6 wavefronts per request, profiles/assoc_grid_r01.txt). ----
#ifndef GMMLOC_SORT_RANK
#define GMMLOC_SORT_RANK 1
#endif
// The histogram pass keeps what its atomic returns — the landmark's rank inside its cell — in the (not yet written)
// assignment array, so that the scatter needs no second round of 8 M atomics: position = cell start + rank.
// A thread can carry kSortU independent landmarks (all loads before the first atomic, all atomics before the first store);
// more than one does not pay: the kernels wait on the L2 atomic units (8.2 M counter updates on 94 k addresses per launch).
#ifndef GMMLOC_SORT_UNROLL
#define GMMLOC_SORT_UNROLL 1  // measured on C2: 1 -> 0.698 ms (sort + scan), 4 -> 0.730, 8 -> 0.746: the L2 atomic units, not latency, set the pace
#endif
constexpr int kSortU = GMMLOC_SORT_UNROLL;
__device__ __forceinline__ long long sort_cell_of(const GridLevel& G, float px, float py, float pz) {
  const long long c = grid_cell(G, px, py, pz);
  return c >= 0 ? c : (long long)G.nx * G.ny * G.nz;  // last bucket: outside the grid / non-finite
}
__global__ void __launch_bounds__(256) grid_hist_kernel(const ScanArgs a, int* __restrict__ hist) {
  const long long i0 = (long long)blockIdx.x * (256 * kSortU) + threadIdx.x;
  const GridLevel& G = a.grid.lev[0];
  float p[kSortU][3];
#pragma unroll
  for (int u = 0; u < kSortU; ++u) {
    const long long i = i0 + u * 256;
    if (i < a.M) {
      p[u][0] = __ldg(a.pts + i * 3 + 0); p[u][1] = __ldg(a.pts + i * 3 + 1); p[u][2] = __ldg(a.pts + i * 3 + 2);
    }
  }
  int r[kSortU];
#pragma unroll
  for (int u = 0; u < kSortU; ++u)
    if (i0 + u * 256 < a.M) r[u] = atomicAdd(hist + sort_cell_of(G, p[u][0], p[u][1], p[u][2]), 1);
#if GMMLOC_SORT_RANK
#pragma unroll
  for (int u = 0; u < kSortU; ++u)
    if (i0 + u * 256 < a.M) a.assign[i0 + u * 256] = r[u];
#endif
}
// The scatter writes the landmark ITSELF next to its original index (one float4 per landmark) so that the scan streams
// its input coalesced: with an index-only permutation the scan gathered 12 B out of every 32-byte sector it fetched
// (510 MB of DRAM reads for 98 MB of landmarks, profiles/assoc_grid_r01.txt).
__global__ void __launch_bounds__(256) grid_scatter_kernel(const ScanArgs a, const int* __restrict__ cur, float4* __restrict__ sorted) {
  const long long i0 = (long long)blockIdx.x * (256 * kSortU) + threadIdx.x;
  const GridLevel& G = a.grid.lev[0];
  float p[kSortU][3];
  int r[kSortU];
#pragma unroll
  for (int u = 0; u < kSortU; ++u) {
    const long long i = i0 + u * 256;
    r[u] = 0;
    if (i < a.M) {
      p[u][0] = __ldg(a.pts + i * 3 + 0); p[u][1] = __ldg(a.pts + i * 3 + 1); p[u][2] = __ldg(a.pts + i * 3 + 2);
#if GMMLOC_SORT_RANK
      r[u] = a.assign[i];
#endif
    }
  }
  int pos[kSortU];
#pragma unroll
  for (int u = 0; u < kSortU; ++u) {
    if (i0 + u * 256 < a.M) {
      const long long c = sort_cell_of(G, p[u][0], p[u][1], p[u][2]);
#if GMMLOC_SORT_RANK
      pos[u] = __ldg(cur + c) + r[u];
#else
      pos[u] = atomicAdd(const_cast<int*>(cur) + c, 1);
#endif
    }
  }
#pragma unroll
  for (int u = 0; u < kSortU; ++u)
    if (i0 + u * 256 < a.M) sorted[pos[u]] = make_float4(p[u][0], p[u][1], p[u][2], __int_as_float((int)(i0 + u * 256)));
}
// In-place exclusive scan of n ints in two launches of one CTA per 4 096-entry tile: (1) every tile is scanned on its own
// and leaves its total in tile_tot; (2) every tile adds the sum of the totals before it.  (One CTA walking all the tiles
// took 30 us for the 94 k cells of the C2 grid — a fixed cost paid by every pipeline chunk of the host entry point.)
__global__ void __launch_bounds__(1024) excl_scan_tiles_kernel(int* __restrict__ d, int n, int* __restrict__ tile_tot) {
  __shared__ int wsum[32];
  const int tid = threadIdx.x, lane = tid & 31, warp = tid >> 5;
  const int idx = blockIdx.x * 4096 + tid * 4;
  int v[4];
#pragma unroll
  for (int q = 0; q < 4; ++q) v[q] = (idx + q < n) ? d[idx + q] : 0;
  const int t = v[0] + v[1] + v[2] + v[3];
  int inc = t;
#pragma unroll
  for (int off = 1; off < 32; off <<= 1) {
    const int u = __shfl_up_sync(0xffffffffu, inc, off);
    if (lane >= off) inc += u;
  }
  if (lane == 31) wsum[warp] = inc;
  __syncthreads();
  if (warp == 0) {
    const int wv = wsum[lane];
    int winc = wv;
#pragma unroll
    for (int off = 1; off < 32; off <<= 1) {
      const int u = __shfl_up_sync(0xffffffffu, winc, off);
      if (lane >= off) winc += u;
    }
    wsum[lane] = winc - wv;
    if (lane == 31) tile_tot[blockIdx.x] = winc;
  }
  __syncthreads();
  int ex = wsum[warp] + inc - t;
#pragma unroll
  for (int q = 0; q < 4; ++q) {
    if (idx + q < n) d[idx + q] = ex;
    ex += v[q];
  }
}
__global__ void __launch_bounds__(1024) excl_scan_add_kernel(int* __restrict__ d, int n, const int* __restrict__ tile_tot) {
  __shared__ int wsum[32];
  __shared__ int s_off;
  const int tid = threadIdx.x, lane = tid & 31, warp = tid >> 5;
  int part = 0;
  for (int u = tid; u < (int)blockIdx.x; u += 1024) part += tile_tot[u];
#pragma unroll
  for (int off = 16; off > 0; off >>= 1) part += __shfl_xor_sync(0xffffffffu, part, off);
  if (lane == 0) wsum[warp] = part;
  __syncthreads();
  if (warp == 0) {
    int w = wsum[lane];
#pragma unroll
    for (int off = 16; off > 0; off >>= 1) w += __shfl_xor_sync(0xffffffffu, w, off);
    if (lane == 0) s_off = w;
  }
  __syncthreads();
  const int off = s_off;
  if (off == 0) return;
  const int idx = blockIdx.x * 4096 + tid * 4;
#pragma unroll
  for (int q = 0; q < 4; ++q)
    if (idx + q < n) d[idx + q] += off;
}

// scores the n staged records (3 float4 each: U, -U mu, id) against one landmark; PEN: `pen` (0 or +inf) is added to every score
#ifndef GMMLOC_GRID_THREADS
#define GMMLOC_GRID_THREADS 128  // C2 sort + scan: 256 -> 0.665 ms, 128 -> 0.648, 512 -> 0.700
#endif
#ifndef GMMLOC_GRID_UNROLL
#define GMMLOC_GRID_UNROLL 4
#endif
constexpr int kGridUnroll = GMMLOC_GRID_UNROLL;
#ifndef GMMLOC_SORT_MIN_PER_CELL
#define GMMLOC_SORT_MIN_PER_CELL 2  // launches with fewer landmarks per level-0 cell keep the natural-order kernel
#endif
#ifndef GMMLOC_GRID_COOP
#define GMMLOC_GRID_COOP 1
#endif
#ifndef GMMLOC_GRID_F32_CERT
#define GMMLOC_GRID_F32_CERT 1
#endif
template <bool PEN>
__device__ __forceinline__ void score_staged(const float4* __restrict__ st, int n, float px, float py, float pz, float pen, float& b1,
                                             float& b2, int& idx) {
#pragma unroll kGridUnroll
  for (int q = 0; q < n; ++q) {
    const float4 r0 = st[3 * q], r1 = st[3 * q + 1], r2 = st[3 * q + 2];
    float s = score32(r0.x, r0.y, r0.z, r0.w, r1.x, r1.y, r1.z, r1.w, r2.x, px, py, pz);
    if (PEN) s = __fadd_rn(s, pen);  // s >= 0 or NaN: s + 0 == s exactly
    idx = (s < b1) ? __float_as_int(r2.y) : idx;
    b2 = fminf(b2, fmaxf(b1, s));  // s < b1: the old best becomes second; else min(b2, s)
    b1 = fminf(b1, s);
  }
}

template <int THREADS, bool SORTED>
__global__ void __launch_bounds__(THREADS) assoc_grid_kernel(const ScanArgs a) {
  __shared__ float4 stage[SORTED ? THREADS / 32 : 1][32][3];  // one staged record per lane and warp (sorted variant)
  __shared__ float4 memb[SORTED ? THREADS / 32 : 1][16];      // landmarks of the members of a sparsely filled round
  (void)memb;
  const int lane = threadIdx.x & 31, wrp = threadIdx.x >> 5;
  (void)lane; (void)wrp; (void)stage;
  const long long w = (long long)blockIdx.x * THREADS + threadIdx.x;
  const bool live = w < a.M;
  if (!SORTED && !live) return;  // the sorted variant keeps whole warps alive for the shuffles
  const float qnan = __int_as_float(0x7fc00000);
  long long i = live ? w : 0;
  float px = qnan, py = qnan, pz = qnan;
  if (SORTED) {
    if (live) {  // streamed once: evict-first, so that the scattered assignment stores keep their sectors in L2
      const float4 v = __ldcs(a.sorted + w);
      px = v.x; py = v.y; pz = v.z;
      i = (long long)__float_as_int(v.w);
    }
  } else if (live) {
    px = __ldg(a.pts + i * 3 + 0); py = __ldg(a.pts + i * 3 + 1); pz = __ldg(a.pts + i * 3 + 2);
  }
  const bool finite_in = (fabsf(px) <= a.xmax) && (fabsf(py) <= a.xmax) && (fabsf(pz) <= a.xmax);  // false for NaN
  float g1 = CUDART_INF_F, g2 = CUDART_INF_F;  // best / second best over the big list
  int gi = -1;
  for (int e = 0; e < a.grid.nbig; ++e) {
    const int k = __ldg(a.grid.big + e);
    const float s = score_rec(a.rec, k, px, py, pz);
    if (s < g1) {
      g2 = g1;
      g1 = s;
      gi = k;
    } else {
      g2 = fminf(g2, s);
    }
  }
  float b1 = g1, b2 = g2;
  int idx = gi, lev_ok = -1;
  double hi1 = CUDART_INF;
  int lev_first = 0;
  if (SORTED) {
    // Level by level, one cell list per round for all the lanes that sit in that cell.  Level 0 is the sort key, so a round
    // usually serves the whole warp; the lanes a level cannot decide (winner not provably inside its radius) meet again at the
    // next level, where the rounds run over THEIR cells — warp-cooperatively (below) instead of one lane walking a long list
    // with the other 31 idle, which is what held up the shipped maps (35 % of the C3 landmarks lie beyond level 0).
    bool fin = false;  // assignment written
    for (int lev = 0; lev < a.grid.nlev; ++lev) {
      const GridLevel& G = a.grid.lev[lev];
      const bool need = live && finite_in && lev_ok < 0 && !fin;
      if (!__any_sync(0xffffffffu, need)) break;
      const long long cell = need ? grid_cell(G, px, py, pz) : -1;
      if (lev > 0 && cell >= 0) {
        b1 = g1;
        b2 = g2;
        idx = gi;
      }
      unsigned todo = __ballot_sync(0xffffffffu, cell >= 0);
      while (todo) {
        const int leader = __ffs(todo) - 1;
        const long long c = __shfl_sync(0xffffffffu, cell, leader);
        const bool mine = (cell == c);
        const bool all_mine = __all_sync(0xffffffffu, mine);
        const float pen = mine ? 0.f : CUDART_INF_F;
        const int e0 = __ldg(G.start + c), e1 = __ldg(G.start + c + 1);
        // the list is staged through shared memory 32 components at a time: every lane fetches ONE record (32 gathers in
        // flight instead of a dependent id -> record chain per component), then all lanes read the staged records with
        // broadcast LDS.128 and score their own landmark
#if GMMLOC_GRID_COOP
        // Few lanes of the warp in this cell (the tail of a cell in a large launch; EVERY round of a small launch, where a cell
        // holds a handful of landmarks): instead of all 32 lanes walking the whole list for m members, the (member, record)
        // pairs are dealt over the lanes — m' = 2, 4, 8 or 16 member slots x 32 / m' records in parallel — and the per-slot
        // (best, second best, id) triples are merged with xor shuffles.  Same score32() values, so the same (b1, b2); which of two
        // EQUAL best scores names the id is immaterial: b2 == b1 then, and the certificate sends the landmark to the exact re-rank.
        const unsigned mm = __ballot_sync(0xffffffffu, mine);
        const int m = __popc(mm);
        if (m <= 16) {
          const int sh = m <= 2 ? 1 : m <= 4 ? 2 : m <= 8 ? 3 : 4;
          const int mp = 1 << sh, slot = lane & (mp - 1), sub = lane >> sh, nsub = 32 >> sh;
          const int rank = __popc(mm & ((1u << lane) - 1u));
          if (mine) memb[wrp][rank] = make_float4(px, py, pz, 0.f);
          __syncwarp();
          const float4 qv = memb[wrp][slot < m ? slot : 0];
          float c1 = CUDART_INF_F, c2 = CUDART_INF_F;
          int ci = -1;
          for (int base = e0; base < e1; base += 32) {
            const int n = min(32, e1 - base);
            if (lane < n) {
              const int k = __ldg(G.list + base + lane);
              const float4 r0 = __ldg(a.rec + 3 * (size_t)k), r1 = __ldg(a.rec + 3 * (size_t)k + 1);
              const float r2 = __ldg(reinterpret_cast<const float*>(a.rec + 3 * (size_t)k + 2));
              stage[wrp][lane][0] = r0;
              stage[wrp][lane][1] = r1;
              stage[wrp][lane][2] = make_float4(r2, __int_as_float(k), 0.f, 0.f);
            }
            __syncwarp();
#pragma unroll 2
            for (int q = sub; q < n; q += nsub) {
              const float4 r0 = stage[wrp][q][0], r1 = stage[wrp][q][1], r2 = stage[wrp][q][2];
              const float sc = score32(r0.x, r0.y, r0.z, r0.w, r1.x, r1.y, r1.z, r1.w, r2.x, qv.x, qv.y, qv.z);
              ci = (sc < c1) ? __float_as_int(r2.y) : ci;
              c2 = fminf(c2, fmaxf(c1, sc));
              c1 = fminf(c1, sc);
            }
            __syncwarp();
          }
          for (int off = mp; off < 32; off <<= 1) {  // lanes of one slot: lane bits sh..4
            const float o1 = __shfl_xor_sync(0xffffffffu, c1, off), o2 = __shfl_xor_sync(0xffffffffu, c2, off);
            const int oi = __shfl_xor_sync(0xffffffffu, ci, off);
            c2 = fminf(fmaxf(c1, o1), fminf(c2, o2));
            ci = (o1 < c1) ? oi : ci;
            c1 = fminf(c1, o1);
          }
          // the member of rank r finds its triple in lane r (slot r, sub 0)
          const int from = mine ? rank : 0;
          const float m1 = __shfl_sync(0xffffffffu, c1, from), m2 = __shfl_sync(0xffffffffu, c2, from);
          const int mi = __shfl_sync(0xffffffffu, ci, from);
          if (mine) {
            b2 = fminf(fmaxf(b1, m1), fminf(b2, m2));
            idx = (m1 < b1) ? mi : idx;
            b1 = fminf(b1, m1);
          }
          todo &= ~mm;
          __syncwarp();  // memb is rewritten by the next sparse round
          continue;
        }
#endif
        for (int base = e0; base < e1; base += 32) {
          const int n = min(32, e1 - base);
          if (lane < n) {
            const int k = __ldg(G.list + base + lane);
            const float4 r0 = __ldg(a.rec + 3 * (size_t)k), r1 = __ldg(a.rec + 3 * (size_t)k + 1);
            const float r2 = __ldg(reinterpret_cast<const float*>(a.rec + 3 * (size_t)k + 2));
            stage[wrp][lane][0] = r0;
            stage[wrp][lane][1] = r1;
            stage[wrp][lane][2] = make_float4(r2, __int_as_float(k), 0.f, 0.f);
          }
          __syncwarp();
          // branch-free best / second-best (the branchy form spent more instructions here than on the score itself:
          // profiles/assoc_grid_r02.txt).  Strict < keeps the first (lowest) id on equal scores, as the reference's ascending scan.
          // A lane that is not in this cell adds +inf to its score (nothing changes for it); written as `mine ? score : inf` the
          // compiler put a branch (BSSY / BRA / BSYNC, four instructions of 21 per record, and a scheduling fence between the
          // staged loads of consecutive records) around every score.  When the whole warp sits in the cell — the usual case —
          // the penalty is not even added.
          if (all_mine) score_staged<false>(&stage[wrp][0][0], n, px, py, pz, 0.f, b1, b2, idx);
          else score_staged<true>(&stage[wrp][0][0], n, px, py, pz, pen, b1, b2, idx);
          __syncwarp();
        }
        todo &= ~__ballot_sync(0xffffffffu, mine);
      }
#if GMMLOC_GRID_F32_CERT
      // The certificate below costs ~110 instructions per landmark in fp64 (two square roots).  A conservative fp32 form of the
      // same two tests decides almost every landmark: with correctly rounded sqrtf and the factors 1 +- 2e-6 (against kRelEps =
      // 2.4e-7 and three fp32 roundings of 6e-8 each) `hi` is an upper bound of the fp64 hi1 and `lo` a lower bound of the fp64
      // lo2, so  lo > hi  and  hi <= sqrt(tmax) (1 - 1e-6)  imply the fp64 tests: the SAME assignment is written, and what the
      // fp32 form cannot decide still goes through the fp64 form.
      if (cell >= 0 && idx >= 0 && a.d2 == nullptr && !(a.gate > 0.0)) {
        const float al = a.alpha * (1.0f + 1e-6f);
        const float hi = __fmaf_rn(__fsqrt_rn(b1), 1.0f + 2e-6f, al), lo = __fmaf_rn(__fsqrt_rn(b2), 1.0f - 2e-6f, -al);
        if (lo > hi && hi <= __fsqrt_rn(G.tmax) * (1.0f - 1e-6f)) {
          a.assign[i] = idx;
          fin = true;
        }
      }
#endif
      if (cell >= 0 && idx >= 0 && !fin) {
        hi1 = sqrt((double)b1) * (1.0 + kRelEps) + (double)a.alpha;
        if (hi1 * hi1 <= (double)G.tmax) lev_ok = lev;
      }
    }
    lev_first = a.grid.nlev;  // nothing left for the per-lane walk below
    if (!live || fin) return;
  }
  if (finite_in && lev_ok < 0) {
    for (int lev = lev_first; lev < a.grid.nlev; ++lev) {
      const GridLevel& G = a.grid.lev[lev];
      const long long cell = grid_cell(G, px, py, pz);
      if (cell < 0) continue;
      b1 = g1;
      b2 = g2;
      idx = gi;
      const int e0 = __ldg(G.start + cell), e1 = __ldg(G.start + cell + 1);
#pragma unroll 4
      for (int e = e0; e < e1; ++e) {
        const int k = __ldg(G.list + e);
        const float s = score_rec(a.rec, k, px, py, pz);
        if (s < b1) {
          b2 = b1;
          b1 = s;
          idx = k;
        } else {
          b2 = fminf(b2, s);
        }
      }
      if (idx >= 0) {
        hi1 = sqrt((double)b1) * (1.0 + kRelEps) + (double)a.alpha;  // upper bound of the winner's true distance
        if (hi1 * hi1 <= (double)G.tmax) {  // every unlisted component is farther than that
          lev_ok = lev;
          break;
        }
      }
    }
  }
  if (lev_ok >= 0) {
    const double lo2 = sqrt((double)b2) * (1.0 - kRelEps) - (double)a.alpha;  // lower bound of every other listed component's
    if (lo2 > hi1) {
      if (a.d2 != nullptr || a.gate > 0.0) {
        const double d = chi2_exact(a.mean + (size_t)idx * 3, a.cov_inv + (size_t)idx * 9, (double)px, (double)py, (double)pz);
        if (a.gate > 0.0 && d > a.gate) idx = -1;
        if (a.d2 != nullptr) a.d2[i] = (float)d;
      }
      a.assign[i] = idx;
      return;
    }
  }
  // not certified: exact re-rank with the squared radius inside which it must evaluate in fp64
  float thr = CUDART_INF_F;
  int lev_rr = -1;
  if (finite_in && idx >= 0) {
    const double rt = (sqrt((double)b1) * (1.0 + kRelEps) + 2.0 * (double)a.alpha) / (1.0 - kRelEps);
    const double t2 = rt * rt;
    thr = (float)t2;
    if ((double)thr < t2) thr = nextafterf(thr, CUDART_INF_F);
    // every component with score32 <= thr has a true distance <= hit: the grid re-rank needs them all listed
    const double hit = sqrt((double)thr) * (1.0 + kRelEps) + (double)a.alpha;
    for (int lev = 0; lev < a.grid.nlev && lev_rr < 0; ++lev)
      if (hit * hit <= (double)a.grid.lev[lev].tmax && grid_cell(a.grid.lev[lev], px, py, pz) >= 0) lev_rr = lev;
  }
  if (lev_rr >= 0) {
    const int pos = atomicAdd(a.ga_count, 1);
    a.ga_idx[pos] = (int32_t)i;
    a.ga_thr[pos] = thr;
  } else {
    const int pos = atomicAdd(a.flag_count, 1);
    a.flag_idx[pos] = (int32_t)i;
    a.flag_thr[pos] = thr;
  }
}

// K2-G: exact re-rank of list A over the landmark's cell list (ascending ids; the big list interleaves, hence the
// lexicographic (d2, id) minimum = the reference's strict '<' on an ascending scan, gmmloc_opt.cpp:209-217).
__global__ void __launch_bounds__(128) assoc_grid_rerank_kernel(const ScanArgs a) {
  // one WARP per flagged landmark: the lanes split the cell list (a thread per landmark leaves the kernel waiting for
  // the few landmarks with a long level-1 list: 500 dependent gathers in one thread)
  const int n = *a.ga_count;
  const int lane = threadIdx.x & 31;
  const int nwarps = (gridDim.x * blockDim.x) >> 5;
  for (int w = (blockIdx.x * blockDim.x + threadIdx.x) >> 5; w < n; w += nwarps) {
    const long long i = a.ga_idx[w];
    const float thr = a.ga_thr[w];
    const float px = __ldg(a.pts + i * 3 + 0), py = __ldg(a.pts + i * 3 + 1), pz = __ldg(a.pts + i * 3 + 2);
    // first level whose lists hold every component with score32 <= thr (K1-G made sure one exists)
    const double hit = sqrt((double)thr) * (1.0 + kRelEps) + (double)a.alpha;
    int lev = a.grid.nlev - 1;
    for (int l = a.grid.nlev - 1; l >= 0; --l)
      if (hit * hit <= (double)a.grid.lev[l].tmax && grid_cell(a.grid.lev[l], px, py, pz) >= 0) lev = l;
    const GridLevel& G = a.grid.lev[lev];
    const long long cell = grid_cell(G, px, py, pz);
    double best = CUDART_INF;
    int bi = -1;
    const int e0 = __ldg(G.start + cell), e1 = __ldg(G.start + cell + 1);
    const int nl = e1 - e0;
    // four entries per lane and round: the id -> record gathers of a round are independent and issued back to back (the
    // coarse lists of the shipped maps hold a thousand entries: one dependent gather chain at a time left the warp idle)
    const int ntot = nl + a.grid.nbig;
    for (int e = lane; e < ntot; e += 32 * 4) {
      int kk[4];
      float ss[4];
#pragma unroll
      for (int u = 0; u < 4; ++u) {
        const int ee = e + 32 * u;
        kk[u] = ee < ntot ? (ee < nl ? __ldg(G.list + e0 + ee) : __ldg(a.grid.big + (ee - nl))) : -1;
      }
#pragma unroll
      for (int u = 0; u < 4; ++u) ss[u] = kk[u] >= 0 ? score_rec(a.rec, kk[u], px, py, pz) : CUDART_INF_F;
#pragma unroll
      for (int u = 0; u < 4; ++u) {
        if (kk[u] >= 0 && !(ss[u] > thr)) {
          const int k = kk[u];
          const double d = chi2_exact(a.mean + (size_t)k * 3, a.cov_inv + (size_t)k * 9, (double)px, (double)py, (double)pz);
          if (d < best || (d == best && k < bi)) {
            best = d;
            bi = k;
          }
        }
      }
    }
    // lexicographic (d2, id) minimum over the lanes
#pragma unroll
    for (int off = 16; off > 0; off >>= 1) {
      const double ob = __shfl_xor_sync(0xffffffffu, best, off);
      const int ok = __shfl_xor_sync(0xffffffffu, bi, off);
      if (ok >= 0 && (bi < 0 || ob < best || (ob == best && ok < bi))) {
        best = ob;
        bi = ok;
      }
    }
    if (lane == 0) {
      int idx = bi;
      if (a.gate > 0.0 && idx >= 0 && best > a.gate) idx = -1;
      a.assign[i] = idx;
      if (a.d2 != nullptr) a.d2[i] = (bi >= 0) ? (float)best : CUDART_INF_F;
    }
  }
}

// sort_cnt: ncell0 + 2 ints + one int per 4 096 cells (tile totals of the scan), sort_perm: M x 16 bytes (one float4 per
// landmark) — or null for the natural-order kernel
cudaError_t launch_scan_grid(ScanArgs a, int* sort_cnt, int32_t* sort_perm, cudaStream_t stream, int* launches) {
  const unsigned grid = (unsigned)((a.M + 255) / 256);
  const GridLevel& G = a.grid.lev[0];
  const long long ncell = (long long)G.nx * G.ny * G.nz;
  if (sort_cnt != nullptr && sort_perm != nullptr && a.M >= GMMLOC_SORT_MIN_PER_CELL * ncell) {
    cudaError_t e;
    if ((e = cudaMemsetAsync(sort_cnt, 0, sizeof(int) * (size_t)(ncell + 1), stream)) != cudaSuccess) return e;
    const unsigned sgrid = (unsigned)((a.M + 256 * kSortU - 1) / (256 * kSortU));
    grid_hist_kernel<<<sgrid, 256, 0, stream>>>(a, sort_cnt);
    {
      const int n = (int)(ncell + 1), tiles = (n + 4095) / 4096;
      int* tile_tot = sort_cnt + ncell + 2;  // behind the counters (api.cu reserves the room)
      excl_scan_tiles_kernel<<<tiles, 1024, 0, stream>>>(sort_cnt, n, tile_tot);
      excl_scan_add_kernel<<<tiles, 1024, 0, stream>>>(sort_cnt, n, tile_tot);
    }
    grid_scatter_kernel<<<sgrid, 256, 0, stream>>>(a, sort_cnt, reinterpret_cast<float4*>(sort_perm));
    a.sorted = reinterpret_cast<const float4*>(sort_perm);
    assoc_grid_kernel<GMMLOC_GRID_THREADS, true><<<(unsigned)((a.M + GMMLOC_GRID_THREADS - 1) / GMMLOC_GRID_THREADS), GMMLOC_GRID_THREADS, 0,
                                                   stream>>>(a);
    if (launches) *launches += 4;
    return cudaGetLastError();
  }
  a.sorted = nullptr;
  assoc_grid_kernel<256, false><<<grid, 256, 0, stream>>>(a);
  return cudaGetLastError();
}
cudaError_t launch_rerank_grid(const ScanArgs& a, cudaStream_t stream) {
  const long long want = (a.M + 127) / 128;
  assoc_grid_rerank_kernel<<<(unsigned)(want < 148 * 8 ? want : 148 * 8), 128, 0, stream>>>(a);
  return cudaGetLastError();
}

// ======================================================================================================
// K1-TC: tensor-core pruning scan (tcgen05 / TMEM).
//
// The whitened coordinates y = U x - U mu of every (landmark, component) pair are a dense contraction with
// K = 4 ([x y z 1] against [U | -U mu]); with fp16 operands and a 3-way hi/lo split folded into the K dimension
// (hi*hi + lo*hi + hi*lo, 11 of 16 K slots) one tcgen05.mma kind::f16 M128 N48 K16 produces the three whitened
// coordinates of 128 landmarks x 16 components (one chunk) to ~2^-21 relative accuracy, fp32 accumulators in
// TMEM.  The CUDA cores then only square-and-add (3 packed FMA-pipe lane-ops per pair instead of 9) and fold the
// chunk minimum.
//
// The tensor-core scores are used for PRUNING only, so exactness is untouched: per landmark the kernel keeps the
// chunks whose minimum is within tc_delta (in sqrt-score units) of the running best; afterwards those candidate
// chunks (typically 1-3 of K/16) are re-scored with the exact fp32 score32() of the FFMA2 scan and go through the
// same certificate / exact re-rank path (finish_landmark, K2).  Every other component is farther than
// tc_delta - 2 (tc_err + alpha) from the winner in exact fp32 scores, where tc_err is the rigorous bound of the
// split / accumulation error computed at map build (map_build.cu) — that value enters the certificate as the
// lower bound of the non-candidates.
//
// CTA = 6 warps: warp 0 streams the B tiles (TMA bulk copies of the pre-laid-out UMMA table) — lane 0;
// warp 1 allocates TMEM and issues the MMAs (lane 0); warps 2-5 are the epilogue (thread = landmark row, TMEM lane
// quadrant = warp % 4).  kTcATiles landmark tiles of 128 stay in shared memory as A operands, so every B chunk
// is used kTcATiles times; kTcBufs accumulator buffers of 48 TMEM columns keep that many MMAs in flight (the
// MMA -> commit -> mbarrier round trip is ~600 cycles, far longer than one 128 x 48 x 16 MMA).
constexpr bool kTcAutoDefault = false;  // flipped once the kernel has been validated and measured on the GPU
constexpr int kTcATiles = 8;                 // landmark tiles of 128 per CTA (every B tile is used 8 times)
constexpr int kTcLm = 128 * kTcATiles;       // landmarks per CTA
constexpr int kTcBufs = 2;                   // 2 x 192 = 384 of the 512 TMEM columns (one CTA per SM)
constexpr int kTcTmemCols = 512;
constexpr int kTcStageTiles = 4;             // 64-component tiles per B stage
constexpr int kTcStageBytes = kTcStageTiles * kTcTileBytes;  // 24 KB
constexpr int kTcStages = 2;
constexpr int kTcCand = 4;                   // candidate chunks remembered per (landmark, tile half)
constexpr int kTcEpiWarps = 8;               // two per TMEM lane quadrant: each takes 32 of a tile's 64 components
constexpr int kTcThreads = 64 + 32 * kTcEpiWarps;
constexpr int kTcABytes = 128 * kTcK * 2;    // one A tile: 4 KB
// dynamic shared memory layout
constexpr int kTcOffA = 0;
constexpr int kTcOffB = kTcOffA + kTcATiles * kTcABytes;
constexpr int kTcOffBar = kTcOffB + kTcStages * kTcStageBytes;
constexpr int kTcOffCandId = kTcOffBar + 128;  // barriers: bfull[2] bempty[2] tfull[2] tempty[2]
constexpr int kTcOffCandM = kTcOffCandId + kTcLm * 2 * kTcCand * 2;
constexpr int kTcOffState = kTcOffCandM + kTcLm * 2 * kTcCand * 4;  // per (landmark, half): b1, evmin, ncand
constexpr int kTcSmem = kTcOffState + kTcLm * 2 * 12;

// kind::f16 (fp16 x fp16 -> fp32), A and B K-major, M = 128, N = 192 (cute::UMMA::InstrDescriptor bit layout)
constexpr uint32_t kTcIdesc = (1u << 4) | ((uint32_t)(kTcRows >> 3) << 17) | ((uint32_t)(128 >> 4) << 24);

// K-major, no-swizzle UMMA shared-memory descriptor: canonical layout ((8,n),2):((1,SBO),LBO) in 16-byte units
__device__ __forceinline__ uint64_t tc_desc(uint32_t saddr, uint32_t lbo_bytes, uint32_t sbo_bytes) {
  return (uint64_t)((saddr >> 4) & 0x3FFF) | ((uint64_t)((lbo_bytes >> 4) & 0x3FFF) << 16) |
         ((uint64_t)((sbo_bytes >> 4) & 0x3FFF) << 32) | (1ull << 46);
}
__device__ __forceinline__ void tc_mma(uint32_t tmem_d, uint64_t adesc, uint64_t bdesc) {
  asm volatile(
      "{\n.reg .pred p;\nsetp.ne.b32 p, %4, 0;\ntcgen05.mma.cta_group::1.kind::f16 [%0], %1, %2, %3, p;\n}\n" ::"r"(tmem_d),
      "l"(adesc), "l"(bdesc), "r"(kTcIdesc), "r"(0u)
      : "memory");
}
__device__ __forceinline__ void tc_commit(uint64_t* bar) {
  asm volatile("tcgen05.commit.cta_group::1.mbarrier::arrive::one.shared::cluster.b64 [%0];" ::"r"(smem_u32(bar)) : "memory");
}
__device__ __forceinline__ void tc_ld16(uint32_t taddr, uint32_t* v) {
  asm volatile(
      "tcgen05.ld.sync.aligned.32x32b.x16.b32 {%0, %1, %2, %3, %4, %5, %6, %7, %8, %9, %10, %11, %12, %13, %14, %15}, [%16];"
      : "=r"(v[0]), "=r"(v[1]), "=r"(v[2]), "=r"(v[3]), "=r"(v[4]), "=r"(v[5]), "=r"(v[6]), "=r"(v[7]), "=r"(v[8]), "=r"(v[9]),
        "=r"(v[10]), "=r"(v[11]), "=r"(v[12]), "=r"(v[13]), "=r"(v[14]), "=r"(v[15])
      : "r"(taddr));
}
__device__ __forceinline__ void mbar_arrive(uint64_t* bar) {
  asm volatile("mbarrier.arrive.shared::cta.b64 _, [%0];" ::"r"(smem_u32(bar)) : "memory");
}
__device__ __forceinline__ uint64_t pack2u(uint32_t lo, uint32_t hi) {
  uint64_t r;
  asm("mov.b64 %0, {%1, %2};" : "=l"(r) : "r"(lo), "r"(hi));
  return r;
}

// slow path of the epilogue (rare): chunk `c` with tensor-core minimum `mch` is within the margin of the running
// best of list `li` (one list per landmark and tile half).  The list stays in ascending chunk order; when it is full
// the entry with the largest minimum is dropped and the smallest dropped value is remembered (evmin): the landmark
// stays certifiable iff that value ends up outside the final margin.
__device__ __noinline__ void tc_candidate(float mch, int c, int li, float delta, uint16_t* cand_id, float* cand_m, float& b1,
                                          float& thr, int& ncand, float& evmin) {
  if (mch < b1) {
    b1 = mch;
    const float r = sqrtf(mch) * (1.f + 1e-6f) + delta;
    thr = r * r * (1.f + 1e-6f);
  }
  uint16_t* id = cand_id + li * kTcCand;
  float* cm = cand_m + li * kTcCand;
  if (ncand == kTcCand) {  // drop the entries the improved best has pushed out of the margin
    int n = 0;
    for (int q = 0; q < kTcCand; ++q)
      if (cm[q] <= thr) {
        cm[n] = cm[q];
        id[n] = id[q];
        ++n;
      }
    ncand = n;
  }
  if (ncand == kTcCand) {
    int w = 0;
    for (int q = 1; q < kTcCand; ++q)
      if (cm[q] > cm[w]) w = q;
    if (mch >= cm[w]) {
      evmin = fminf(evmin, mch);  // the new chunk itself is the worst: it is the one dropped
    } else {
      evmin = fminf(evmin, cm[w]);
      for (int q = w; q + 1 < kTcCand; ++q) {
        cm[q] = cm[q + 1];
        id[q] = id[q + 1];
      }
      cm[kTcCand - 1] = mch;
      id[kTcCand - 1] = (uint16_t)c;
    }
  } else {
    cm[ncand] = mch;
    id[ncand] = (uint16_t)c;
    ++ncand;
  }
}

#ifdef GMMLOC_TC_DEBUG
#define TC_T0() const long long t_dbg0 = clock64()
#define TC_ACC(var) var += clock64() - t_dbg0
#else
#define TC_T0()
#define TC_ACC(var)
#endif
__global__ void __launch_bounds__(kTcThreads, 1) assoc_scan_tc_kernel(const ScanArgs a) {
  long long dbg_wait_full = 0, dbg_wait_empty = 0, dbg_wait_ld = 0;
  const long long dbg_start = clock64();
  (void)dbg_wait_full; (void)dbg_wait_empty; (void)dbg_wait_ld; (void)dbg_start;
  extern __shared__ __align__(128) unsigned char smem_raw[];
  uint64_t* bars = reinterpret_cast<uint64_t*>(smem_raw + kTcOffBar);
  uint64_t* bfull = bars;
  uint64_t* bempty = bars + kTcStages;
  uint64_t* tfull = bars + 2 * kTcStages;
  uint64_t* tempty = tfull + kTcBufs;
  uint16_t* cand_id = reinterpret_cast<uint16_t*>(smem_raw + kTcOffCandId);
  float* cand_m = reinterpret_cast<float*>(smem_raw + kTcOffCandM);
  float* st_b1 = reinterpret_cast<float*>(smem_raw + kTcOffState);  // [kTcLm * 2]
  float* st_ev = st_b1 + kTcLm * 2;
  int* st_n = reinterpret_cast<int*>(st_ev + kTcLm * 2);
  __shared__ uint32_t s_tmem;

  const int tid = threadIdx.x, warp = tid >> 5, lane = tid & 31;
  const long long base = (long long)blockIdx.x * kTcLm;
  const int ntiles = (a.Kpad + kTcTileComps - 1) / kTcTileComps;
  const int nstages = (ntiles + kTcStageTiles - 1) / kTcStageTiles;

  if (tid == 0) {
    for (int i = 0; i < kTcStages; ++i) {
      mbar_init(&bfull[i], 1);
      mbar_init(&bempty[i], 1);
    }
    for (int i = 0; i < kTcBufs; ++i) {
      mbar_init(&tfull[i], 1);
      mbar_init(&tempty[i], kTcEpiWarps);  // one arrival per epilogue warp (per-thread arrivals serialise on the barrier)
    }
    asm volatile("fence.mbarrier_init.release.cluster;" ::: "memory");
  }
  if (warp == 1) {
    asm volatile("tcgen05.alloc.cta_group::1.sync.aligned.shared::cta.b32 [%0], %1;" ::"r"(smem_u32(&s_tmem)), "r"(kTcTmemCols));
    asm volatile("tcgen05.relinquish_alloc_permit.cta_group::1.sync.aligned;");
  }
  // ---- the 256 epilogue threads write the A operand rows: thread (row, hh) does the tiles t = hh, hh + 2, ... ----
  const int row = (warp & 3) * 32 + lane;            // TMEM lane = landmark row of a tile (epilogue warps only)
  const int hh = (warp >= 2) ? ((warp - 2) >> 2) : 0;  // which 32 components of every tile this warp folds
  if (warp >= 2) {
    for (int t = hh; t < kTcATiles; t += 2) {
      const long long i = base + t * 128 + row;
      float x = 0.f, y = 0.f, z = 0.f;
      if (i < a.M) {
        x = __ldg(a.pts + i * 3 + 0);
        y = __ldg(a.pts + i * 3 + 1);
        z = __ldg(a.pts + i * 3 + 2);
      }
      const float v[3] = {x - a.cen[0], y - a.cen[1], z - a.cen[2]};
      __half hi[3], lo[3];
#pragma unroll
      for (int c = 0; c < 3; ++c) {
        hi[c] = __float2half_rn(v[c]);
        lo[c] = __float2half_rn(v[c] - __half2float(hi[c]));
      }
      const __half one = __float2half_rn(1.f), zero = __float2half_rn(0.f);
      // K slots 0..7: v_hi(3) 1 v_lo(3) 1      K slots 8..15: v_hi(3) 0 0 0 0 0
      const __half k0[8] = {hi[0], hi[1], hi[2], one, lo[0], lo[1], lo[2], one};
      const __half k1[8] = {hi[0], hi[1], hi[2], zero, zero, zero, zero, zero};
      unsigned char* At = smem_raw + kTcOffA + t * kTcABytes;
      *reinterpret_cast<uint4*>(At + row * 16) = *reinterpret_cast<const uint4*>(k0);
      *reinterpret_cast<uint4*>(At + 128 * 16 + row * 16) = *reinterpret_cast<const uint4*>(k1);
    }
  }
  asm volatile("fence.proxy.async.shared::cta;" ::: "memory");  // A rows visible to the tensor core's async proxy
  asm volatile("tcgen05.fence::before_thread_sync;" ::: "memory");
  __syncthreads();
  asm volatile("tcgen05.fence::after_thread_sync;" ::: "memory");
  const uint32_t tmem = s_tmem;

  if (warp == 0) {
    // ===== B producer: one TMA bulk copy per stage of kTcStageTiles tiles =====
    if (lane == 0) {
      for (int st = 0; st < nstages; ++st) {
        const int slot = st % kTcStages;
        if (st >= kTcStages) mbar_wait(&bempty[slot], (uint32_t)(((st / kTcStages) - 1) & 1));
        const int nt = min(kTcStageTiles, ntiles - st * kTcStageTiles);
        const uint32_t bytes = (uint32_t)nt * kTcTileBytes;
        mbar_expect_tx(&bfull[slot], bytes);
        tma_bulk_g2s(smem_raw + kTcOffB + slot * kTcStageBytes,
                     reinterpret_cast<const unsigned char*>(a.tc_table) + (size_t)st * kTcStageBytes, bytes, &bfull[slot]);
      }
    }
  } else if (warp == 1) {
    // ===== MMA issuer: accumulator s = tile * kTcATiles + t lives in TMEM buffer s % 2 =====
    if (lane == 0) {
      const uint32_t a0 = smem_u32(smem_raw + kTcOffA), b0 = smem_u32(smem_raw + kTcOffB);
      uint32_t buf = 0, use = 0;  // use = s / kTcBufs
      for (int st = 0; st < nstages; ++st) {
        const int slot = st % kTcStages;
        mbar_wait(&bfull[slot], (uint32_t)((st / kTcStages) & 1));
        asm volatile("tcgen05.fence::after_thread_sync;" ::: "memory");
        const int nt = min(kTcStageTiles, ntiles - st * kTcStageTiles);
        for (int c = 0; c < nt; ++c) {
          const uint64_t bd = tc_desc(b0 + slot * kTcStageBytes + c * kTcTileBytes, kTcRows * 16, 128);
#pragma unroll
          for (int t = 0; t < kTcATiles; ++t) {
            if (use > 0) {
              TC_T0();
              mbar_wait(&tempty[buf], (use - 1) & 1);
              TC_ACC(dbg_wait_empty);
              asm volatile("tcgen05.fence::after_thread_sync;" ::: "memory");
            }
            tc_mma(tmem + buf * 256, tc_desc(a0 + t * kTcABytes, 128 * 16, 128), bd);
            tc_commit(&tfull[buf]);
            if (++buf == kTcBufs) {
              buf = 0;
              ++use;
            }
          }
        }
        tc_commit(&bempty[slot]);  // fires when every MMA reading this stage has completed
      }
#ifdef GMMLOC_TC_DEBUG
      if (blockIdx.x == 100) printf("issuer: total %lld wait_empty %lld\n", clock64() - dbg_start, dbg_wait_empty);
#endif
    }
  } else {
    // ===== epilogue: warp (q, hh) folds components [32 hh, 32 hh + 32) of every tile for the landmark rows of TMEM
    // lane quadrant q = warp % 4 =====
    const uint32_t lane_base = (uint32_t)((warp & 3) * 32) << 16;
    float b1[kTcATiles], thr[kTcATiles], evmin[kTcATiles];
    int ncand[kTcATiles];
#pragma unroll
    for (int t = 0; t < kTcATiles; ++t) {
      b1[t] = CUDART_INF_F;
      thr[t] = CUDART_INF_F;
      evmin[t] = CUDART_INF_F;  // smallest chunk minimum ever evicted from the candidate list
      ncand[t] = 0;
    }
    // Per accumulator: drain this warp's 96 TMEM columns into registers at once and hand the buffer straight back
    // to the MMA issuer (the MMA -> commit -> wake round trip bounds this kernel, so a buffer must not be held
    // while the arithmetic runs).
    uint32_t buf = 0, use = 0;
    for (int c = 0; c < ntiles; ++c) {
#pragma unroll
      for (int t = 0; t < kTcATiles; ++t) {
        uint32_t y[96];
        {
          TC_T0();
          mbar_wait(&tfull[buf], use & 1);
          TC_ACC(dbg_wait_full);
        }
        asm volatile("tcgen05.fence::after_thread_sync;" ::: "memory");
        const uint32_t tb = tmem + lane_base + buf * 256 + hh * 32;
        {
          TC_T0();
#pragma unroll
          for (int r = 0; r < 3; ++r) {
            tc_ld16(tb + r * kTcTileComps, y + r * 32);
            tc_ld16(tb + r * kTcTileComps + 16, y + r * 32 + 16);
          }
          asm volatile("tcgen05.wait::ld.sync.aligned;" ::: "memory");
          TC_ACC(dbg_wait_ld);
        }
        asm volatile("tcgen05.fence::before_thread_sync;" ::: "memory");
        __syncwarp();
        if (lane == 0) mbar_arrive(&tempty[buf]);
        if (++buf == kTcBufs) {
          buf = 0;
          ++use;
        }
        TC_T0();
#pragma unroll
        for (int h = 0; h < 2; ++h) {
          // registers [0,32) y0 | [32,64) y1 | [64,96) y2 of this warp's 32 components; chunk h = 16 of them
          uint64_t acc[8];
#pragma unroll
          for (int j = 0; j < 8; ++j) {
            const uint64_t u0 = pack2u(y[16 * h + 2 * j], y[16 * h + 2 * j + 1]);
            const uint64_t u1 = pack2u(y[32 + 16 * h + 2 * j], y[32 + 16 * h + 2 * j + 1]);
            const uint64_t u2 = pack2u(y[64 + 16 * h + 2 * j], y[64 + 16 * h + 2 * j + 1]);
            acc[j] = fma2(u2, u2, fma2(u1, u1, mul2(u0, u0)));
          }
          float mch = CUDART_INF_F;
#pragma unroll
          for (int j = 0; j < 8; ++j) {
            float s0, s1;
            unpack2(acc[j], s0, s1);
            mch = fminf(fminf(mch, s0), s1);
          }
          if (mch <= thr[t])  // rare: this chunk is within the margin of the running best
            tc_candidate(mch, 4 * c + 2 * hh + h, (t * 128 + row) * 2 + hh, a.tc_delta, cand_id, cand_m, b1[t], thr[t], ncand[t],
                         evmin[t]);
        }
        TC_ACC(dbg_wait_empty);
      }
    }
#ifdef GMMLOC_TC_DEBUG
    if (blockIdx.x == 100 && tid == 64)
      printf("epilogue main loop: %lld cycles, wait_full %lld ld %lld compute %lld\n", clock64() - dbg_start, dbg_wait_full, dbg_wait_ld, dbg_wait_empty);
#endif
#pragma unroll
    for (int t = 0; t < kTcATiles; ++t) {
      const int li = (t * 128 + row) * 2 + hh;
      st_b1[li] = b1[t];
      st_ev[li] = evmin[t];
      st_n[li] = ncand[t];
    }
  }
  asm volatile("tcgen05.fence::before_thread_sync;" ::: "memory");
  __syncthreads();
  if (warp == 1) {
    asm volatile("tcgen05.dealloc.cta_group::1.sync.aligned.b32 %0, %1;" ::"r"(tmem), "r"(kTcTmemCols));
  }
  if (warp < 2) return;
  // ---- merge the two halves of every landmark, exact fp32 re-score of the candidate chunks, usual certificate.
  // thread (row, hh) finishes the tiles t = hh, hh + 2, ... ----
  for (int t = hh; t < kTcATiles; t += 2) {
    const long long i = base + t * 128 + row;
    if (i >= a.M) continue;
    const int l2 = (t * 128 + row) * 2;
    const float px = __ldg(a.pts + i * 3 + 0), py = __ldg(a.pts + i * 3 + 1), pz = __ldg(a.pts + i * 3 + 2);
    const bool finite_in = (fabsf(px) <= a.xmax) && (fabsf(py) <= a.xmax) && (fabsf(pz) <= a.xmax);  // false for NaN
    const float bb = fminf(st_b1[l2], st_b1[l2 + 1]);  // tensor-core best over all components
    float thr = CUDART_INF_F;
    if (bb < CUDART_INF_F) {
      const float r = sqrtf(bb) * (1.f + 1e-6f) + a.tc_delta;
      thr = r * r * (1.f + 1e-6f);
    }
    const float ev = fminf(st_ev[l2], st_ev[l2 + 1]);
    float s1 = CUDART_INF_F, s2 = CUDART_INF_F;
    int idx = -1;
    if (finite_in && !(ev <= thr)) {  // an evicted chunk inside the final margin: not certifiable here
      // ascending component order over both lists (each list is ascending; chunks of the two halves interleave)
      int q0 = 0, q1 = 0;
      const int n0 = st_n[l2], n1 = st_n[l2 + 1];
      while (q0 < n0 || q1 < n1) {
        const int c0 = q0 < n0 ? (int)cand_id[l2 * kTcCand + q0] : 0x7fffffff;
        const int c1 = q1 < n1 ? (int)cand_id[(l2 + 1) * kTcCand + q1] : 0x7fffffff;
        int ch;
        float cm;
        if (c0 < c1) {
          ch = c0;
          cm = cand_m[l2 * kTcCand + q0++];
        } else {
          ch = c1;
          cm = cand_m[(l2 + 1) * kTcCand + q1++];
        }
        if (!(cm <= thr) || ch * kChunk >= a.Kpad) continue;  // outside the final margin / chunk of the table padding
        for (int j = 0; j < kChunk; ++j) {
          const float sc = score_of(a.scan, ch * kChunk + j, px, py, pz);
          if (sc < s1) {  // ascending component order, strict '<'
            s2 = s1;
            s1 = sc;
            idx = ch * kChunk + j;
          } else {
            s2 = fminf(s2, sc);
          }
        }
      }
    }
    if (idx >= 0) {
      // every non-candidate component is farther than tc_slack (sqrt-score units) from the winner in exact fp32 scores
      const float r = sqrtf(s1) * (1.f - 1e-6f) + a.tc_slack;
      finish_landmark(a, i, px, py, pz, true, idx, s1, fminf(s2, r * r * (1.f - 1e-6f)));
    } else {
      // not certifiable here (candidate list overflowed, or the landmark is outside the certified box): exact
      // re-rank.  Inside the box the tensor-core best still bounds the winner's true distance, which keeps the
      // radius K2 has to search finite.
      float thr2 = CUDART_INF_F;
      if (finite_in && bb < CUDART_INF_F) {
        const double rt = (sqrt((double)bb) * (1.0 + 1e-6) + (double)a.tc_err + 2.0 * (double)a.alpha) / (1.0 - kRelEps);
        const double t2 = rt * rt;
        thr2 = (float)t2;
        if ((double)thr2 < t2) thr2 = nextafterf(thr2, CUDART_INF_F);
      }
      const int pos = atomicAdd(a.flag_count, 1);
      a.flag_idx[pos] = (int32_t)i;
      a.flag_thr[pos] = thr2;
    }
  }
}

cudaError_t launch_scan_tc(const ScanArgs& a, cudaStream_t stream) {
  cudaError_t e = cudaFuncSetAttribute(assoc_scan_tc_kernel, cudaFuncAttributeMaxDynamicSharedMemorySize, kTcSmem);
  if (e != cudaSuccess) return e;
  const long long grid = (a.M + kTcLm - 1) / kTcLm;
  assoc_scan_tc_kernel<<<(unsigned)grid, kTcThreads, kTcSmem, stream>>>(a);
  return cudaGetLastError();
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
                             cudaStream_t stream, int* launches, cudaEvent_t ev_between, int* used_tc, int* d_sort_cnt,
                             int32_t* d_sort_perm) {
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
  // work lists: d_flag_count = {list A count, list B count}; entries [0, M) = list A (grid re-rank), [M, 2 M) = list B
  // (brute-force exact re-rank)
  a.flag_count = d_flag_count + 1;
  a.flag_idx = d_flag_idx + M;
  a.flag_thr = d_flag_thr + M;
  a.ga_count = d_flag_count;
  a.ga_idx = d_flag_idx;
  a.ga_thr = d_flag_thr;
  a.grid = m.grid;
  a.rec = reinterpret_cast<const float4*>(m.rec);
  a.sorted = nullptr;
  a.rerank_two_pass = 0;
  int extra_launches = 0;
  a.tc_table = m.tc_table;
  a.cen[0] = m.tc_centre[0]; a.cen[1] = m.tc_centre[1]; a.cen[2] = m.tc_centre[2];
  {
    // candidate margin: wide enough that the certificate's gap 2 alpha (1 + eps) fits inside the guaranteed slack
    const float need = 2.0f * (m.tc_err + m.alpha);
    a.tc_err = m.tc_err;
    a.tc_delta = fmaxf(1.0f, 1.25f * need + 8.0f * m.alpha);
    a.tc_slack = a.tc_delta - need;
  }
  cudaError_t e;
  // variant: 0 auto, -1 FFMA2 scan, 1/2/4/8 FFMA2 scan with that many landmarks per thread, 16 tensor-core scan,
  // 32 grid-indexed scan.  GMMLOC_ASSOC_KERNEL=grid|tc|ffma2 overrides the automatic choice (A/B runs of the bench).
  static const int forced = [] {
    const char* v = debug_env("GMMLOC_ASSOC_KERNEL");
    if (v == nullptr) return 0;
    if (std::strcmp(v, "tc") == 0) return 16;
    if (std::strcmp(v, "ffma2") == 0) return -1;
    if (std::strcmp(v, "grid") == 0) return 32;
    return 0;
  }();
  if (variant == 0) variant = forced;
  // automatic: the grid-indexed scan whenever the map has a grid (identical results, ~K/30 of the work)
  const bool use_grid = m.grid.nlev > 0 && (variant == 32 || variant == 0);
  const bool tc_possible = m.tc_ok && m.Kpad / kChunk < 65000;
  // the tensor-core scan needs enough landmarks to fill 148 SMs x 2 CTAs x 512 landmarks
  const bool use_tc = !use_grid && tc_possible && (variant == 16 || (variant == 0 && kTcAutoDefault && M >= 148LL * 4 * kTcLm));
  if (used_tc) *used_tc = use_grid ? 2 : (use_tc ? 1 : 0);
  if (use_grid) {
    if ((e = launch_scan_grid(a, d_sort_cnt, d_sort_perm, stream, &extra_launches)) != cudaSuccess) return e;
    if (ev_between != nullptr && (e = cudaEventRecord(ev_between, stream)) != cudaSuccess) return e;
    ev_between = nullptr;
    if ((e = launch_rerank_grid(a, stream)) != cudaSuccess) return e;
    a.rerank_two_pass = 1;
  }
  if (use_tc) {
    if ((e = launch_scan_tc(a, stream)) != cudaSuccess) return e;
  }
  // Fill the machine: 148 SMs x 2 CTAs; shrink the per-thread landmark count for small batches.
  int P = (variant == 16 || variant == 32 || variant < 0) ? 0 : variant;
  if (P <= 0) {
    if (M >= 148LL * GMMLOC_SCAN_MINB * GMMLOC_SCAN_THREADS * GMMLOC_SCAN_P) P = 8;
    else if (M >= 148LL * 2 * 256 * 4) P = 4;
    else if (M >= 148LL * 2 * 128 * 2) P = 2;
    else P = 1;
  }
  if (use_tc || use_grid) {
    e = cudaSuccess;
  } else {
    switch (P) {
      case 8: e = launch_scan_t<GMMLOC_SCAN_P, GMMLOC_SCAN_THREADS, GMMLOC_SCAN_MINB>(a, stream); break;
      case 4: e = launch_scan_t<4, 256, 2>(a, stream); break;
      case 2: e = launch_scan_t<2, 128, 4>(a, stream); break;
      default: e = launch_scan_t<1, 128, 4>(a, stream); break;
    }
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
    const long long max_ctas = (M + 127) / 128;  // one block of 128 landmarks per slot; a CTA has up to 4 slots per round
    const unsigned grid = (unsigned)(max_ctas < 148 * 4 ? max_ctas : 148 * 4);
    assoc_rerank_kernel<4, 128><<<grid, 128, smem, stream>>>(a);
  }
  if (launches) *launches = (use_grid ? 3 : 2) + extra_launches;
  return cudaGetLastError();
}

}  // namespace gmmloc
