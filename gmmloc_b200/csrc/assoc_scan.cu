// assoc_scan.cu — K1 (fp32 brute-force Mahalanobis scan) and K2 (exact fp64 re-rank).
//
// Replaces: GaussianComponent::chi2 / MDist2 (reference gmmloc/src/gmm/gaussian.cpp:65-70,
// gmmloc/include/gmmloc/gmm/gaussian.h:53-56) evaluated for every (landmark, component) pair with the
// reference's argmin convention (strict '<' on an ascending scan => lowest index wins ties,
// gmmloc/src/gmmloc_opt.cpp:175-197,209-217) and optional gate (gmmloc_opt.cpp:230).
//
// K1: point-stationary.  Each thread keeps P landmarks in registers; the whitened component table
//     (9 floats / component: U = L^T rows and -U mu) streams through shared memory in TMA bulk-copy
//     tiles (cp.async.bulk + mbarrier, 3 stages) and is read with warp-broadcast LDS.128.  Per pair:
//     9 FMA-pipe instructions (d2 = |U x - c|^2) + 1/2 FMNMX3.  The running best / second best are
//     tracked per chunk of 16 components (5 ALU ops per chunk), the winning chunk is re-scanned once at
//     the end to recover the index and the in-chunk runner-up.
// Exactness: the scan certifies an assignment when sqrt(second) - sqrt(best) exceeds twice the
//     rigorous fp32 error bound alpha; everything else is pushed to a flag list.
// K2: one warp per flagged landmark replays the fp32 scores (bit-identical: same score32()), and
//     evaluates every candidate within the bound in fp64 with the oracle's operation order.
#include <math_constants.h>

#include "common.cuh"

namespace gmmloc {

namespace {

constexpr int kStages = 3;
constexpr int kTileComps = 384;  // 384 * 36 B = 13.5 KB per stage
constexpr int kTileFloats = kTileComps * kScanFloats;
constexpr int kTileBytes = kTileFloats * 4;

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

  float x0[P], x1[P], x2[P], b1[P], b2[P];
  int cid[P];
#pragma unroll
  for (int p = 0; p < P; ++p) {
    const long long i = base + (long long)p * THREADS + tid;
    if (i < a.M) {
      x0[p] = __ldg(a.pts + i * 3 + 0);
      x1[p] = __ldg(a.pts + i * 3 + 1);
      x2[p] = __ldg(a.pts + i * 3 + 2);
    } else {
      x0[p] = x1[p] = x2[p] = 0.f;
    }
    b1[p] = CUDART_INF_F;
    b2[p] = CUDART_INF_F;
    cid[p] = 0;
  }
  __syncthreads();  // barrier inits visible before anyone waits

  for (int t = 0; t < ntiles; ++t) {
    const int stage = t % kStages;
    mbar_wait(&full[stage], (uint32_t)((t / kStages) & 1));
    const float4* sp = reinterpret_cast<const float4*>(tiles + (size_t)stage * kTileFloats);
    const int nchunks = min(kTileComps, a.Kpad - t * kTileComps) / kChunk;
    const int chunk0 = t * (kTileComps / kChunk);
    for (int c = 0; c < nchunks; ++c) {
      float mch[P];
#pragma unroll
      for (int p = 0; p < P; ++p) mch[p] = CUDART_INF_F;
#pragma unroll
      for (int q = 0; q < kChunk / 4; ++q) {
        // 4 components = 36 floats = 9 broadcast LDS.128
        float f[36];
#pragma unroll
        for (int v = 0; v < 9; ++v) {
          const float4 w = sp[(c * (kChunk / 4) + q) * 9 + v];
          f[4 * v + 0] = w.x;
          f[4 * v + 1] = w.y;
          f[4 * v + 2] = w.z;
          f[4 * v + 3] = w.w;
        }
#pragma unroll
        for (int j = 0; j < 4; j += 2) {
          const float* fa = f + 9 * j;
          const float* fb = f + 9 * (j + 1);
#pragma unroll
          for (int p = 0; p < P; ++p) {
            const float sa = score32(fa[0], fa[1], fa[2], fa[3], fa[4], fa[5], fa[6], fa[7], fa[8], x0[p], x1[p], x2[p]);
            const float sb = score32(fb[0], fb[1], fb[2], fb[3], fb[4], fb[5], fb[6], fb[7], fb[8], x0[p], x1[p], x2[p]);
            mch[p] = fminf(fminf(mch[p], sa), sb);  // FMNMX3
          }
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
    const float px = x0[p], py = x1[p], pz = x2[p];
    const bool finite_in = (fabsf(px) <= a.xmax) && (fabsf(py) <= a.xmax) && (fabsf(pz) <= a.xmax);  // false for NaN
    const float* cp = a.scan + (size_t)cid[p] * kChunk * kScanFloats;
    int jbest = -1;
    float second = CUDART_INF_F;
    for (int j = 0; j < kChunk; ++j) {
      const float* f = cp + j * kScanFloats;
      const float s = score32(__ldg(f + 0), __ldg(f + 1), __ldg(f + 2), __ldg(f + 3), __ldg(f + 4), __ldg(f + 5),
                              __ldg(f + 6), __ldg(f + 7), __ldg(f + 8), px, py, pz);
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

// K2: one warp per flagged landmark.
__global__ void __launch_bounds__(256) assoc_rerank_kernel(const ScanArgs a) {
  const int lane = threadIdx.x & 31;
  const int warp = (blockIdx.x * blockDim.x + threadIdx.x) >> 5;
  const int nwarps = (gridDim.x * blockDim.x) >> 5;
  const int nflag = *a.flag_count;
  for (int w = warp; w < nflag; w += nwarps) {
    const long long i = a.flag_idx[w];
    const float thr = a.flag_thr[w];
    const float px = __ldg(a.pts + i * 3 + 0), py = __ldg(a.pts + i * 3 + 1), pz = __ldg(a.pts + i * 3 + 2);
    const double dx = (double)px, dy = (double)py, dz = (double)pz;
    double best = CUDART_INF;  // stands for numeric_limits<double>::max(): any finite d2 is smaller
    int bi = -1;
    const bool full = !(thr < CUDART_INF_F);
    for (int k = lane; k < a.K; k += 32) {
      bool cand = full;
      if (!full) {
        const float* f = a.scan + (size_t)k * kScanFloats;
        const float s = score32(__ldg(f + 0), __ldg(f + 1), __ldg(f + 2), __ldg(f + 3), __ldg(f + 4), __ldg(f + 5),
                                __ldg(f + 6), __ldg(f + 7), __ldg(f + 8), px, py, pz);
        cand = s <= thr;
      }
      if (cand) {
        const double d = chi2_exact(a.mean + (size_t)k * 3, a.cov_inv + (size_t)k * 9, dx, dy, dz);
        if (d < best) {  // strict '<': ascending k within a lane keeps the lowest index
          best = d;
          bi = k;
        }
      }
    }
    // warp argmin, ties -> lowest index
#pragma unroll
    for (int off = 16; off > 0; off >>= 1) {
      const double ob = __shfl_xor_sync(0xffffffffu, best, off);
      const int oi = __shfl_xor_sync(0xffffffffu, bi, off);
      const bool take = (oi >= 0) && (bi < 0 || ob < best || (ob == best && oi < bi));
      if (take) {
        best = ob;
        bi = oi;
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
    if (M >= 148LL * 2 * 256 * 8) P = 8;
    else if (M >= 148LL * 2 * 256 * 4) P = 4;
    else if (M >= 148LL * 2 * 128 * 2) P = 2;
    else P = 1;
  }
  switch (P) {
    case 8: e = launch_scan_t<8, 256, 2>(a, stream); break;
    case 4: e = launch_scan_t<4, 256, 2>(a, stream); break;
    case 2: e = launch_scan_t<2, 128, 4>(a, stream); break;
    default: e = launch_scan_t<1, 128, 4>(a, stream); break;
  }
  if (e != cudaSuccess) return e;
  if (ev_between != nullptr && (e = cudaEventRecord(ev_between, stream)) != cudaSuccess) return e;
  assoc_rerank_kernel<<<148 * 4, 256, 0, stream>>>(a);
  if (launches) *launches = 2;
  return cudaGetLastError();
}

}  // namespace gmmloc
