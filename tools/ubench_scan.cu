// Micro-benchmark of the association scan's inner loop (K1): isolates what limits the FFMA2 issue rate —
// instruction order / operand reuse, the FMNMX3 min-fold, the broadcast LDS.128 of the coefficient table.
// Build: nvcc -gencode arch=compute_100a,code=sm_100a -O3 -o ubench_scan ubench_scan.cu
#include <cstdint>
#include <cstdio>
#include <cuda_runtime.h>

#define ITERS 4096  // groups of 4 components per thread

__device__ __forceinline__ uint64_t pack2(float lo, float hi) {
  uint64_t r;
  asm("mov.b64 %0, {%1, %2};" : "=l"(r) : "f"(lo), "f"(hi));
  return r;
}
__device__ __forceinline__ void unpack2(uint64_t v, float& lo, float& hi) {
  asm("mov.b64 {%0, %1}, %2;" : "=f"(lo), "=f"(hi) : "l"(v));
}
template <bool VOL>
__device__ __forceinline__ uint64_t fma2(uint64_t a, uint64_t b, uint64_t c) {
  uint64_t d;
  if (VOL) asm volatile("fma.rn.f32x2 %0, %1, %2, %3;" : "=l"(d) : "l"(a), "l"(b), "l"(c));
  else asm("fma.rn.f32x2 %0, %1, %2, %3;" : "=l"(d) : "l"(a), "l"(b), "l"(c));
  return d;
}
template <bool VOL>
__device__ __forceinline__ uint64_t mul2(uint64_t a, uint64_t b) {
  uint64_t d;
  if (VOL) asm volatile("mul.rn.f32x2 %0, %1, %2;" : "=l"(d) : "l"(a), "l"(b));
  else asm("mul.rn.f32x2 %0, %1, %2;" : "=l"(d) : "l"(a), "l"(b));
  return d;
}

// ORDER 0: landmark-major, compiler-scheduled (K1 as written)   1: landmark-major, program order pinned
//       2: coefficient-major, program order pinned (the same coefficient pair feeds P consecutive FFMA2)
template <int ORDER, bool MIN, bool LDS, int P, int THREADS, int MINB>
__global__ void __launch_bounds__(THREADS, MINB) kern(float* out, const float* pts, int iters) {
  __shared__ __align__(16) float tab[256 * 36];  // 256 groups x 36 floats = 36 KB
  for (int i = threadIdx.x; i < 256 * 36; i += THREADS) tab[i] = 1e-3f * (float)(i % 97) - 0.04f;
  __syncthreads();
  uint64_t X0[P], X1[P], X2[P];
  float m[P];
#pragma unroll
  for (int p = 0; p < P; ++p) {
    const float x = pts[(blockIdx.x * THREADS + threadIdx.x) * 3 + 0] + p, y = pts[1] + p, z = pts[2] - p;
    X0[p] = pack2(x, x); X1[p] = pack2(y, y); X2[p] = pack2(z, z);
    m[p] = 3e38f;
  }
  const ulonglong2* sp = reinterpret_cast<const ulonglong2*>(tab);
  uint64_t ua[9], ub[9];
#pragma unroll
  for (int v = 0; v < 9; ++v) { ua[v] = pack2(0.5f + v, 0.25f * v); ub[v] = pack2(1.5f - v, 0.125f * v); }
  constexpr bool VOL = ORDER != 0;
  for (int it = 0; it < iters; ++it) {
    if (LDS) {
#pragma unroll
      for (int v = 0; v < 9; ++v) {
        const ulonglong2 w = sp[(it & 255) * 9 + v];
        ua[v] = w.x; ub[v] = w.y;
      }
    }
    uint64_t sa[P], sb[P];
    if (ORDER == 2) {
      uint64_t y0[P], y1[P], y2[P];
#define PASS(u)                                                                     \
      _Pragma("unroll") for (int p = 0; p < P; ++p) y0[p] = fma2<VOL>(u[2], X2[p], u[6]);   \
      _Pragma("unroll") for (int p = 0; p < P; ++p) y1[p] = fma2<VOL>(u[4], X2[p], u[7]);   \
      _Pragma("unroll") for (int p = 0; p < P; ++p) y2[p] = fma2<VOL>(u[5], X2[p], u[8]);   \
      _Pragma("unroll") for (int p = 0; p < P; ++p) y0[p] = fma2<VOL>(u[1], X1[p], y0[p]);  \
      _Pragma("unroll") for (int p = 0; p < P; ++p) y1[p] = fma2<VOL>(u[3], X1[p], y1[p]);  \
      _Pragma("unroll") for (int p = 0; p < P; ++p) y0[p] = fma2<VOL>(u[0], X0[p], y0[p]);  \
      _Pragma("unroll") for (int p = 0; p < P; ++p) y0[p] = mul2<VOL>(y0[p], y0[p]);        \
      _Pragma("unroll") for (int p = 0; p < P; ++p) y0[p] = fma2<VOL>(y1[p], y1[p], y0[p]); \
      _Pragma("unroll") for (int p = 0; p < P; ++p) y0[p] = fma2<VOL>(y2[p], y2[p], y0[p]);
      PASS(ua)
#pragma unroll
      for (int p = 0; p < P; ++p) sa[p] = y0[p];
      PASS(ub)
#pragma unroll
      for (int p = 0; p < P; ++p) sb[p] = y0[p];
    } else {
#pragma unroll
      for (int p = 0; p < P; ++p) {
        {
          const uint64_t y0 = fma2<VOL>(ua[0], X0[p], fma2<VOL>(ua[1], X1[p], fma2<VOL>(ua[2], X2[p], ua[6])));
          const uint64_t y1 = fma2<VOL>(ua[3], X1[p], fma2<VOL>(ua[4], X2[p], ua[7]));
          const uint64_t y2 = fma2<VOL>(ua[5], X2[p], ua[8]);
          sa[p] = fma2<VOL>(y2, y2, fma2<VOL>(y1, y1, mul2<VOL>(y0, y0)));
        }
        {
          const uint64_t y0 = fma2<VOL>(ub[0], X0[p], fma2<VOL>(ub[1], X1[p], fma2<VOL>(ub[2], X2[p], ub[6])));
          const uint64_t y1 = fma2<VOL>(ub[3], X1[p], fma2<VOL>(ub[4], X2[p], ub[7]));
          const uint64_t y2 = fma2<VOL>(ub[5], X2[p], ub[8]);
          sb[p] = fma2<VOL>(y2, y2, fma2<VOL>(y1, y1, mul2<VOL>(y0, y0)));
        }
      }
    }
#pragma unroll
    for (int p = 0; p < P; ++p) {
      if (MIN) {
        float s0, s1, s2, s3;
        unpack2(sa[p], s0, s1);
        unpack2(sb[p], s2, s3);
        m[p] = fminf(fminf(m[p], s0), s1);
        m[p] = fminf(fminf(m[p], s2), s3);
      } else {  // keep the results alive without ALU work: feed them back into the landmark operand
        X2[p] = fma2<VOL>(sa[p], sb[p], X2[p]);
      }
    }
  }
  float acc = 0.f;
#pragma unroll
  for (int p = 0; p < P; ++p) {
    float a, b;
    unpack2(X2[p], a, b);
    acc += m[p] + a + b;
  }
  out[blockIdx.x * THREADS + threadIdx.x] = acc;
}

template <int ORDER, bool MIN, bool LDS, int P, int THREADS, int MINB>
void run(const char* name) {
  const int ctas = 148 * MINB * 4;
  float *out, *pts;
  cudaMalloc(&out, (size_t)ctas * THREADS * 4);
  cudaMalloc(&pts, (size_t)ctas * THREADS * 12 + 64);
  cudaMemset(pts, 0, (size_t)ctas * THREADS * 12 + 64);
  cudaEvent_t e0, e1;
  cudaEventCreate(&e0); cudaEventCreate(&e1);
  for (int w = 0; w < 2; ++w) kern<ORDER, MIN, LDS, P, THREADS, MINB><<<ctas, THREADS>>>(out, pts, ITERS);
  cudaEventRecord(e0);
  const int reps = 5;
  for (int w = 0; w < reps; ++w) kern<ORDER, MIN, LDS, P, THREADS, MINB><<<ctas, THREADS>>>(out, pts, ITERS);
  cudaEventRecord(e1);
  cudaEventSynchronize(e1);
  float ms;
  cudaEventElapsedTime(&ms, e0, e1);
  ms /= reps;
  const double packed = (18.0 + (MIN ? 0 : 1)) * P * ITERS * (double)ctas * THREADS / 32.0;  // FFMA2/FMUL2 warp-instructions
  const double per_smsp_clk = packed / (ms * 1e-3 * 1.965e9 * 148 * 4);
  const double pairs = 4.0 * P * ITERS * (double)ctas * THREADS;  // (landmark, component) pairs
  printf("%-58s ms=%8.3f  FFMA2/clk/SMSP=%.3f (%.1f%% of 0.5)  pairs/s=%.3e  equiv frames/s(2000x10000)=%.0f  [%s]\n", name, ms,
         per_smsp_clk, 200.0 * per_smsp_clk, pairs / (ms * 1e-3), pairs / (ms * 1e-3) / 2e7, cudaGetErrorString(cudaGetLastError()));
  cudaFree(out); cudaFree(pts);
}

int main() {
  run<0, true, true, 8, 256, 2>("K1 as written: compiler order, FMNMX3, LDS   P8 T256 B2");
  run<1, true, true, 8, 256, 2>("landmark-major pinned, FMNMX3, LDS           P8 T256 B2");
  run<2, true, true, 8, 256, 2>("coefficient-major pinned, FMNMX3, LDS        P8 T256 B2");
  run<2, true, false, 8, 256, 2>("coefficient-major pinned, FMNMX3, no LDS     P8 T256 B2");
  run<2, false, false, 8, 256, 2>("coefficient-major pinned, no min, no LDS     P8 T256 B2");
  run<1, false, false, 8, 256, 2>("landmark-major pinned, no min, no LDS        P8 T256 B2");
  run<0, false, true, 8, 256, 2>("compiler order, no min, LDS                  P8 T256 B2");
  run<2, true, true, 8, 128, 4>("coefficient-major pinned, FMNMX3, LDS        P8 T128 B4");
  run<2, true, true, 4, 256, 4>("coefficient-major pinned, FMNMX3, LDS        P4 T256 B4");
  run<0, true, true, 4, 256, 4>("compiler order, FMNMX3, LDS                  P4 T256 B4");
  run<2, true, true, 12, 128, 3>("coefficient-major pinned, FMNMX3, LDS        P12 T128 B3");
  return 0;
}
