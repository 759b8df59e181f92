// Micro-benchmark / bring-up of a tcgen05 (5th-gen tensor core) formulation of the association scan's hot
// loop — groundwork for the next round.
//
// Whitened scores as a dense contraction: y = [U | -U mu] [x; 1] is a GEMM with K = 4; with a 3-way TF32 split
// (hi*hi + hi*lo + lo*hi) folded into the K dimension it is K = 12 (padded to 16).  One CTA owns 128 landmarks
// (UMMA M = 128) and streams tiles of 64 components (3 output rows each => UMMA N = 192); the accumulator
// (128 x 192 fp32) lives in TMEM (double-buffered) and the CUDA cores only do  d2 = y0^2 + y1^2 + y2^2  and the
// running minimum (3 packed FMA-pipe lane-ops + 1/2 FMNMX3 per pair instead of 9 + 1/2).
//
// This tool checks the UMMA descriptors / TMEM round trip against a host reference and measures
//   (a) tcgen05.mma kind::tf32 M128 N192 K8 issue rate,   (b) the TMEM -> register -> d2 -> min epilogue rate,
//   (c) both together (double-buffered TMEM, MMA warp + 4 epilogue warps).
// Build: nvcc -gencode arch=compute_100a,code=sm_100a -O3 -o ubench_tc ubench_tc.cu
#include <cstdint>
#include <cstdio>
#include <cstdlib>
#include <cstring>
#include <vector>
#include <cuda_fp16.h>
#include <cuda_runtime.h>

constexpr int kM = 128;       // landmarks per CTA (TMEM lanes)
constexpr int kComps = 64;    // components per B tile
constexpr int kN = 3 * kComps;  // 192 accumulator columns: [0,64) y0, [64,128) y1, [128,192) y2
constexpr int kK = 16;        // 4 features x 3-way split, padded
constexpr int kTmemCols = 512;

__device__ __forceinline__ uint32_t smem_u32(const void* p) { return (uint32_t)__cvta_generic_to_shared(p); }

__device__ __forceinline__ void mbar_init(uint64_t* bar, uint32_t count) {
  asm volatile("mbarrier.init.shared::cta.b64 [%0], %1;" ::"r"(smem_u32(bar)), "r"(count));
}
__device__ __forceinline__ void mbar_arrive(uint64_t* bar) {
  asm volatile("mbarrier.arrive.shared::cta.b64 _, [%0];" ::"r"(smem_u32(bar)) : "memory");
}
__device__ __forceinline__ void mbar_wait(uint64_t* bar, uint32_t parity) {
  uint32_t done;
  do {
    asm volatile(
        "{\n.reg .pred p;\nmbarrier.try_wait.parity.shared::cta.b64 p, [%1], %2;\nselp.u32 %0, 1, 0, p;\n}\n"
        : "=r"(done)
        : "r"(smem_u32(bar)), "r"(parity)
        : "memory");
  } while (!done);
}

// K-major, no-swizzle UMMA shared-memory descriptor (cute::UMMA::SmemDescriptor, mma_sm100_desc.hpp):
// canonical layout ((8,n),2):((1,SBO),LBO) in 16-byte units; here element (row r, k-chunk kc of 4 tf32) sits at
// byte kc * (rows * 16) + r * 16  =>  SBO = 128 B (next 8-row core matrix), LBO = rows * 16 B (next k-chunk).
__device__ __forceinline__ uint64_t make_desc(uint32_t saddr, uint32_t lbo_bytes, uint32_t sbo_bytes) {
  uint64_t d = 0;
  d |= (uint64_t)((saddr >> 4) & 0x3FFF);
  d |= (uint64_t)((lbo_bytes >> 4) & 0x3FFF) << 16;
  d |= (uint64_t)((sbo_bytes >> 4) & 0x3FFF) << 32;
  d |= 1ull << 46;  // descriptor version 1 (Blackwell)
  return d;         // layout_type = SWIZZLE_NONE (0), base_offset 0
}
// cute::UMMA::InstrDescriptor for kind::tf32, fp32 accumulate, A and B K-major
constexpr uint32_t kIdescTf32 = (1u << 4) | (2u << 7) | (2u << 10) | ((uint32_t)(kN >> 3) << 17) | ((uint32_t)(kM >> 4) << 24);
// kind::f16 with fp16 operands (format 0), fp32 accumulate: K = 16 per instruction
constexpr uint32_t kIdescF16 = (1u << 4) | (0u << 7) | (0u << 10) | ((uint32_t)(kN >> 3) << 17) | ((uint32_t)(kM >> 4) << 24);

__device__ __forceinline__ void umma_tf32(uint32_t tmem_d, uint64_t adesc, uint64_t bdesc, uint32_t accumulate) {
  asm volatile(
      "{\n.reg .pred p;\nsetp.ne.b32 p, %4, 0;\ntcgen05.mma.cta_group::1.kind::tf32 [%0], %1, %2, %3, p;\n}\n" ::"r"(tmem_d),
      "l"(adesc), "l"(bdesc), "r"(kIdescTf32), "r"(accumulate)
      : "memory");
}
__device__ __forceinline__ void umma_f16(uint32_t tmem_d, uint64_t adesc, uint64_t bdesc, uint32_t accumulate) {
  asm volatile(
      "{\n.reg .pred p;\nsetp.ne.b32 p, %4, 0;\ntcgen05.mma.cta_group::1.kind::f16 [%0], %1, %2, %3, p;\n}\n" ::"r"(tmem_d),
      "l"(adesc), "l"(bdesc), "r"(kIdescF16), "r"(accumulate)
      : "memory");
}
__device__ __forceinline__ void umma_commit(uint64_t* bar) {
  asm volatile("tcgen05.commit.cta_group::1.mbarrier::arrive::one.shared::cluster.b64 [%0];" ::"r"(smem_u32(bar)) : "memory");
}
__device__ __forceinline__ void tmem_ld16(uint32_t taddr, uint32_t* v) {
  asm volatile(
      "tcgen05.ld.sync.aligned.32x32b.x16.b32 {%0, %1, %2, %3, %4, %5, %6, %7, %8, %9, %10, %11, %12, %13, %14, %15}, [%16];"
      : "=r"(v[0]), "=r"(v[1]), "=r"(v[2]), "=r"(v[3]), "=r"(v[4]), "=r"(v[5]), "=r"(v[6]), "=r"(v[7]), "=r"(v[8]), "=r"(v[9]),
        "=r"(v[10]), "=r"(v[11]), "=r"(v[12]), "=r"(v[13]), "=r"(v[14]), "=r"(v[15])
      : "r"(taddr));
}
__device__ __forceinline__ void tmem_wait_ld() { asm volatile("tcgen05.wait::ld.sync.aligned;" ::: "memory"); }
__device__ __forceinline__ void fence_before() { asm volatile("tcgen05.fence::before_thread_sync;" ::: "memory"); }
__device__ __forceinline__ void fence_after() { asm volatile("tcgen05.fence::after_thread_sync;" ::: "memory"); }

__device__ __forceinline__ uint64_t pack2(uint32_t lo, uint32_t hi) {
  uint64_t r;
  asm("mov.b64 %0, {%1, %2};" : "=l"(r) : "r"(lo), "r"(hi));
  return r;
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

// MODE 0: validate (one tile, dump D);  1: MMA only;  2: epilogue only (TMEM reads of whatever is there);
//      3: pipelined MMA + epilogue.   F16: fp16 operands, one K = 16 MMA per tile (else tf32, two K = 8 MMAs).
// NEPI epilogue warps (4: one per TMEM lane quadrant, all 64 components; 8: two per quadrant, 32 components each)
template <int MODE, bool F16, int NEPI, int SW = 0, int INFLIGHT = 2>
__global__ void __launch_bounds__(64 + 32 * NEPI, 1) tc_kernel(const float* gA, const float* gB, float* gD, float* gmin, int iters,
                                                                long long* cycles) {
  __shared__ __align__(128) unsigned char sA_raw[kM * kK * 4];
  __shared__ __align__(128) unsigned char sB_raw[kN * kK * 4];
  __shared__ __align__(8) uint64_t full[2], empty[2];
  __shared__ uint32_t s_tmem;
  const int tid = threadIdx.x, warp = tid >> 5, lane = tid & 31;
  constexpr int EPC = F16 ? 8 : 4;  // elements per 16-byte k-chunk

  // operand element (row r, k) lives at byte (k / EPC) * (rows * 16) + r * 16 + (k % EPC) * elem_size
  for (int i = tid; i < kM * kK; i += blockDim.x) {
    const int r = i / kK, k = i % kK;
    size_t e = (size_t)(k / EPC) * (kM * EPC) + (size_t)r * EPC + (k % EPC);
    if (SW == 1) e = (size_t)r * 16 + (size_t)(((k / 8) ^ ((r >> 2) & 1)) * 8) + (k % 8);  // fp16 only: 32-byte rows
    if (F16) reinterpret_cast<__half*>(sA_raw)[e] = __float2half_rn(gA[i]);
    else reinterpret_cast<float*>(sA_raw)[e] = gA[i];
  }
  for (int i = tid; i < kN * kK; i += blockDim.x) {
    const int r = i / kK, k = i % kK;
    size_t e = (size_t)(k / EPC) * (kN * EPC) + (size_t)r * EPC + (k % EPC);
    if (SW == 1) e = (size_t)r * 16 + (size_t)(((k / 8) ^ ((r >> 2) & 1)) * 8) + (k % 8);
    if (F16) reinterpret_cast<__half*>(sB_raw)[e] = __float2half_rn(gB[i]);
    else reinterpret_cast<float*>(sB_raw)[e] = gB[i];
  }
  if (tid == 0) {
    for (int b = 0; b < 2; ++b) {
      mbar_init(&full[b], 1);
      mbar_init(&empty[b], NEPI);  // one arrival per epilogue warp
    }
    asm volatile("fence.mbarrier_init.release.cluster;" ::: "memory");
  }
  if (warp == 0) {
    asm volatile("tcgen05.alloc.cta_group::1.sync.aligned.shared::cta.b32 [%0], %1;" ::"r"(smem_u32(&s_tmem)), "r"(kTmemCols));
    asm volatile("tcgen05.relinquish_alloc_permit.cta_group::1.sync.aligned;");
  }
  // generic-proxy writes of the operands must be visible to the async proxy (UMMA reads smem through it)
  asm volatile("fence.proxy.async.shared::cta;" ::: "memory");
  fence_before();
  __syncthreads();
  fence_after();
  const uint32_t tmem = s_tmem;

  const long long t0 = clock64();
  if (warp == 0) {
    // ===== MMA issuer =====
    if (MODE != 2) {
      const uint32_t a0 = smem_u32(sA_raw), b0 = smem_u32(sB_raw);
      for (int it = 0; it < iters; ++it) {
        const int buf = it & 1;
        if (MODE == 3 && it >= 2) mbar_wait(&empty[buf], ((it >> 1) - 1) & 1);
        fence_after();
        if (lane == 0) {
          if (F16 && SW == 1) {
            // SWIZZLE_32B K-major: 8-row groups of 256 B (SBO), LBO unused; layout_type = 6 in bits [61,64)
            umma_f16(tmem + buf * 256, make_desc(a0, 16, 256) | (6ull << 61), make_desc(b0, 16, 256) | (6ull << 61), 0);
          } else if (F16) {
            umma_f16(tmem + buf * 256, make_desc(a0, kM * 16, 128), make_desc(b0, kN * 16, 128), 0);
          } else {
#pragma unroll
            for (int ks = 0; ks < kK / 8; ++ks) {
              const uint64_t ad = make_desc(a0 + ks * 2 * (kM * 16), kM * 16, 128);
              const uint64_t bd = make_desc(b0 + ks * 2 * (kN * 16), kN * 16, 128);
              umma_tf32(tmem + buf * 256, ad, bd, ks > 0);
            }
          }
          umma_commit(&full[buf]);
        }
        __syncwarp();
        if (MODE == 0) mbar_wait(&full[buf], 0);
        if (MODE == 1 && INFLIGHT == 2 && it >= 1) mbar_wait(&full[buf ^ 1], ((it - 1) >> 1) & 1);  // keep two tiles in flight
      }
      if (MODE == 1) mbar_wait(&full[(iters - 1) & 1], ((iters - 1) >> 1) & 1);
    }
  } else if (warp >= 2) {
    // ===== epilogue: warp w may touch TMEM lanes 32 (w % 4) .. +31 =====
    const int q = warp & 3;
    const int half = (warp - 2) >> 2;  // 0 for the first four epilogue warps
    const uint32_t lane_base = (uint32_t)(q * 32) << 16;
    float best = 3e38f;
    if (MODE == 0) {
      if (half == 0) {
        mbar_wait(&full[0], 0);
        fence_after();
        for (int c0 = 0; c0 < kN; c0 += 16) {
          uint32_t v[16];
          tmem_ld16(tmem + lane_base + c0, v);
          tmem_wait_ld();
          for (int j = 0; j < 16; ++j) gD[(size_t)(q * 32 + lane) * kN + c0 + j] = __uint_as_float(v[j]);
        }
      }
    } else if (MODE == 2 || MODE == 3) {
      constexpr int CH = (NEPI == 8) ? 2 : 4;  // chunks of 16 components per warp and tile
      for (int it = 0; it < iters; ++it) {
        const int buf = it & 1;
        if (MODE == 3) {
          mbar_wait(&full[buf], (it >> 1) & 1);
          fence_after();
        }
        const uint32_t tb = tmem + lane_base + buf * 256 + half * (CH * 16);
        // y0 | y1 | y2 live 64 columns apart; software-pipelined: the loads of chunk c+1 are in flight while c computes
        uint32_t y0[2][16], y1[2][16], y2[2][16];
        tmem_ld16(tb, y0[0]);
        tmem_ld16(tb + 64, y1[0]);
        tmem_ld16(tb + 128, y2[0]);
#pragma unroll
        for (int c = 0; c < CH; ++c) {
          tmem_wait_ld();
          if (c + 1 < CH) {
            tmem_ld16(tb + (c + 1) * 16, y0[(c + 1) & 1]);
            tmem_ld16(tb + 64 + (c + 1) * 16, y1[(c + 1) & 1]);
            tmem_ld16(tb + 128 + (c + 1) * 16, y2[(c + 1) & 1]);
          }
          float mch = 3e38f;
#pragma unroll
          for (int j = 0; j < 16; j += 2) {
            const uint64_t a = pack2(y0[c & 1][j], y0[c & 1][j + 1]), b = pack2(y1[c & 1][j], y1[c & 1][j + 1]),
                           cc = pack2(y2[c & 1][j], y2[c & 1][j + 1]);
            const uint64_t s = fma2(cc, cc, fma2(b, b, mul2(a, a)));
            float s0, s1;
            asm("mov.b64 {%0, %1}, %2;" : "=f"(s0), "=f"(s1) : "l"(s));
            mch = fminf(fminf(mch, s0), s1);
          }
          best = fminf(best, mch);
        }
        if (MODE == 3) {
          fence_before();
          __syncwarp();
          if (lane == 0) mbar_arrive(&empty[buf]);
        }
      }
      gmin[(blockIdx.x * 2 + half) * kM + q * 32 + lane] = best;
    }
  }
  const long long t1 = clock64();
  fence_before();
  __syncthreads();
  if (warp == 0) {
    asm volatile("tcgen05.dealloc.cta_group::1.sync.aligned.b32 %0, %1;" ::"r"(tmem), "r"(kTmemCols));
  }
  if (tid == 64 && blockIdx.x == 0) *cycles = t1 - t0;
}

template <int MODE, bool F16, int NEPI, int SW = 0, int INFLIGHT = 2>
void run(const char* name, const float* dA, const float* dB, float* dD, float* dmin, int iters, int ctas) {
  long long* dc;
  cudaMalloc(&dc, 8);
  cudaEvent_t e0, e1;
  cudaEventCreate(&e0);
  cudaEventCreate(&e1);
  tc_kernel<MODE, F16, NEPI, SW, INFLIGHT><<<ctas, 64 + 32 * NEPI>>>(dA, dB, dD, dmin, iters, dc);
  cudaEventRecord(e0);
  tc_kernel<MODE, F16, NEPI, SW, INFLIGHT><<<ctas, 64 + 32 * NEPI>>>(dA, dB, dD, dmin, iters, dc);
  cudaEventRecord(e1);
  cudaError_t err = cudaEventSynchronize(e1);
  float ms = 0;
  cudaEventElapsedTime(&ms, e0, e1);
  long long hc = 0;
  cudaMemcpy(&hc, dc, 8, cudaMemcpyDeviceToHost);
  const double pairs = (double)ctas * iters * kM * kComps;
  printf("%-52s ms=%8.3f  cycles/tile=%7.1f  pairs/s=%.3e  tensor TFLOP/s=%6.1f  equiv frames/s (2000x10000)=%.0f  [%s]\n", name, ms,
         (double)hc / iters, pairs / (ms * 1e-3), pairs * 3 * kK * 2 / (ms * 1e-3) / 1e12, pairs / (ms * 1e-3) / 2e7,
         cudaGetErrorString(err));
  cudaFree(dc);
}

template <bool F16, int SW = 0>
bool validate(const std::vector<float>& A, const std::vector<float>& B, const float* dA, const float* dB, float* dD, float* dmin) {
  long long* dc;
  cudaMalloc(&dc, 8);
  cudaMemset(dD, 0, (size_t)kM * kN * 4);
  tc_kernel<0, F16, 4, SW><<<1, 192>>>(dA, dB, dD, dmin, 1, dc);
  cudaError_t err = cudaDeviceSynchronize();
  std::vector<float> D((size_t)kM * kN);
  cudaMemcpy(D.data(), dD, D.size() * 4, cudaMemcpyDeviceToHost);
  double maxerr = 0, maxref = 0;
  int bad = 0;
  for (int m = 0; m < kM; ++m)
    for (int n = 0; n < kN; ++n) {
      double ref = 0;
      for (int k = 0; k < kK; ++k) ref += (double)A[m * kK + k] * (double)B[n * kK + k];
      const double e = fabs(ref - (double)D[(size_t)m * kN + n]);
      if (e > maxerr) maxerr = e;
      if (fabs(ref) > maxref) maxref = fabs(ref);
      if (e > 1e-3 * (1 + fabs(ref)) && bad < 5) {
        printf("  mismatch D[%d][%d] = %g, expected %g\n", m, n, D[(size_t)m * kN + n], ref);
        ++bad;
      }
    }
  const bool ok = maxerr < 1e-3 * maxref;
  if (SW) printf("(SWIZZLE_32B) ");
  printf("validate %s: launch %s, max |D - ref| = %.3e (max |ref| = %.3e) -> %s\n", F16 ? "kind::f16 " : "kind::tf32", cudaGetErrorString(err),
         maxerr, maxref, ok ? "OK" : "MISMATCH");
  cudaFree(dc);
  return ok;
}

int main() {
  // values exactly representable in fp16 and tf32 (multiples of 1/64 below 8) so the products are exact in fp32
  std::vector<float> A(kM * kK), B(kN * kK);
  srand(1);
  for (auto& v : A) v = (float)(rand() % 513 - 256) / 64.0f;
  for (auto& v : B) v = (float)(rand() % 513 - 256) / 64.0f;
  float *dA, *dB, *dD, *dmin;
  const int ctas = 148;
  cudaMalloc(&dA, A.size() * 4);
  cudaMalloc(&dB, B.size() * 4);
  cudaMalloc(&dD, (size_t)kM * kN * 4);
  cudaMalloc(&dmin, (size_t)ctas * 2 * kM * 4);
  cudaMemcpy(dA, A.data(), A.size() * 4, cudaMemcpyHostToDevice);
  cudaMemcpy(dB, B.data(), B.size() * 4, cudaMemcpyHostToDevice);
  if (!validate<false>(A, B, dA, dB, dD, dmin)) return 1;
  if (!validate<true>(A, B, dA, dB, dD, dmin)) return 1;
  const bool sw_ok = validate<true, 1>(A, B, dA, dB, dD, dmin);

  const int iters = 20000;
  run<1, false, 4>("tf32  2 x (M128 N192 K8) MMA only", dA, dB, dD, dmin, iters, ctas);
  run<1, true, 4>("f16   1 x (M128 N192 K16) MMA only", dA, dB, dD, dmin, iters, ctas);
  run<1, true, 4, 0, 64>("f16   MMA only, no wait between MMAs (deep queue)", dA, dB, dD, dmin, iters, ctas);
  if (sw_ok) run<1, true, 4, 1, 2>("f16   SWIZZLE_32B MMA only, 2 in flight", dA, dB, dD, dmin, iters, ctas);
  if (sw_ok) run<1, true, 4, 1, 64>("f16   SWIZZLE_32B MMA only, deep queue", dA, dB, dD, dmin, iters, ctas);
  run<2, true, 4>("epilogue only, 4 warps (TMEM ld + d2 + min)", dA, dB, dD, dmin, iters, ctas);
  run<2, true, 8>("epilogue only, 8 warps", dA, dB, dD, dmin, iters, ctas);
  run<3, false, 4>("tf32 MMA + epilogue (4 warps), double-buffered TMEM", dA, dB, dD, dmin, iters, ctas);
  run<3, true, 4>("f16  MMA + epilogue (4 warps), double-buffered TMEM", dA, dB, dD, dmin, iters, ctas);
  run<3, true, 8>("f16  MMA + epilogue (8 warps), double-buffered TMEM", dA, dB, dD, dmin, iters, ctas);
  return 0;
}
