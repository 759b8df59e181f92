// Micro-benchmark: issue/pipe rates of the instructions the association scan is built from
// (FFMA, FFMA2, FMNMX3, DFMA, mixes).  Used to anchor the FP32-issue roofline in DESIGN.md.
// Build: nvcc -gencode arch=compute_100a,code=sm_100a -O3 -o ubench_pipes ubench_pipes.cu
#include <cstdio>
#include <cstdint>
#include <cuda_runtime.h>

#define CHAINS 16
#define ITERS 65536

template <int MODE>
__global__ void __launch_bounds__(256) kern(float* out, float a, float b, long long* cyc) {
  float r[CHAINS];
  double d[8];
  unsigned long long p[CHAINS / 2];
#pragma unroll
  for (int i = 0; i < CHAINS; ++i) r[i] = threadIdx.x * 1e-3f + i;
#pragma unroll
  for (int i = 0; i < 8; ++i) d[i] = threadIdx.x * 1e-3 + i;
#pragma unroll
  for (int i = 0; i < CHAINS / 2; ++i) {
    float2 t = make_float2(r[2 * i], r[2 * i + 1]);
    p[i] = *reinterpret_cast<unsigned long long*>(&t);
  }
  float2 ab2 = make_float2(a, a), bb2 = make_float2(b, b);
  unsigned long long pa = *reinterpret_cast<unsigned long long*>(&ab2);
  unsigned long long pb = *reinterpret_cast<unsigned long long*>(&bb2);
  float m = 1e30f, m2 = 1e30f;
  unsigned um = 0xffffffffu, um2 = 0xffffffffu;
  unsigned uacc[4] = {1u, 2u, 3u, 4u};
  unsigned uacc2[CHAINS];
#pragma unroll
  for (int i = 0; i < CHAINS; ++i) uacc2[i] = 0xffffff00u + i + threadIdx.x;
  __shared__ float sm[1024];
  for (int i = threadIdx.x; i < 1024; i += blockDim.x) sm[i] = i * 1e-3f;
  __syncthreads();
  long long t0 = clock64();
  for (int it = 0; it < ITERS; ++it) {
    if (MODE == 0) {  // FFMA 3-reg
#pragma unroll
      for (int i = 0; i < CHAINS; ++i) r[i] = fmaf(r[i], a, b);
    } else if (MODE == 1) {  // FFMA2
#pragma unroll
      for (int i = 0; i < CHAINS / 2; ++i)
        asm volatile("fma.rn.f32x2 %0, %0, %1, %2;" : "+l"(p[i]) : "l"(pa), "l"(pb));
    } else if (MODE == 2) {  // 8 FFMA : 1 FMNMX3 (two values folded per min3)
#pragma unroll
      for (int i = 0; i < CHAINS; ++i) r[i] = fmaf(r[i], a, b);
      m = fminf(fminf(m, r[0]), r[1]);
      m = fminf(fminf(m, r[2]), r[3]);
    } else if (MODE == 3) {  // FFMA2 + FMNMX3
#pragma unroll
      for (int i = 0; i < CHAINS / 2; ++i)
        asm volatile("fma.rn.f32x2 %0, %0, %1, %2;" : "+l"(p[i]) : "l"(pa), "l"(pb));
      float2 q0 = *reinterpret_cast<float2*>(&p[0]);
      float2 q1 = *reinterpret_cast<float2*>(&p[1]);
      m = fminf(fminf(m, q0.x), q0.y);
      m = fminf(fminf(m, q1.x), q1.y);
    } else if (MODE == 4) {  // DFMA
#pragma unroll
      for (int i = 0; i < 8; ++i) d[i] = fma(d[i], (double)a, (double)b);
    } else if (MODE == 5) {  // FMNMX3 only
#pragma unroll
      for (int i = 0; i < CHAINS; ++i) r[i] = fminf(fminf(r[i], a), r[(i + 1) % CHAINS]);
    } else if (MODE == 6) {  // 1:1 FFMA + FMNMX (alu pipe co-issue test)
#pragma unroll
      for (int i = 0; i < CHAINS / 2; ++i) r[i] = fmaf(r[i], a, b);
#pragma unroll
      for (int i = CHAINS / 2; i < CHAINS; ++i) r[i] = fminf(r[i], r[i - CHAINS / 2]);
    } else if (MODE == 7) {  // FMUL
#pragma unroll
      for (int i = 0; i < CHAINS; ++i) r[i] = r[i] * a;

    } else if (MODE == 8) {  // 16 FFMA + 2 VIMNMX3.U32 (integer min of the float bit patterns)
#pragma unroll
      for (int i = 0; i < CHAINS; ++i) r[i] = fmaf(r[i], a, b);
      um = min(min(um, __float_as_uint(r[0])), __float_as_uint(r[1]));
      um = min(min(um, __float_as_uint(r[2])), __float_as_uint(r[3]));
    } else if (MODE == 9) {  // 16 FFMA + 4 LOP3
#pragma unroll
      for (int i = 0; i < CHAINS; ++i) r[i] = fmaf(r[i], a, b);
#pragma unroll
      for (int i = 0; i < 4; ++i) uacc[i] = (uacc[i] & __float_as_uint(r[i])) ^ (unsigned)it;
    } else if (MODE == 10) {  // 8 FFMA2 + 2 VIMNMX3
#pragma unroll
      for (int i = 0; i < CHAINS / 2; ++i)
        asm volatile("fma.rn.f32x2 %0, %0, %1, %2;" : "+l"(p[i]) : "l"(pa), "l"(pb));
      um = min(min(um, (unsigned)(p[0] & 0xffffffffu)), (unsigned)(p[0] >> 32));
      um = min(min(um, (unsigned)(p[1] & 0xffffffffu)), (unsigned)(p[1] >> 32));
    } else if (MODE == 11) {  // 16 FFMA + 4 FMNMX3
#pragma unroll
      for (int i = 0; i < CHAINS; ++i) r[i] = fmaf(r[i], a, b);
      m = fminf(fminf(m, r[0]), r[1]);
      m = fminf(fminf(m, r[2]), r[3]);
      m2 = fminf(fminf(m2, r[4]), r[5]);
      m2 = fminf(fminf(m2, r[6]), r[7]);
    } else if (MODE == 12) {  // 16 FFMA + 4 VIMNMX3
#pragma unroll
      for (int i = 0; i < CHAINS; ++i) r[i] = fmaf(r[i], a, b);
      um = min(min(um, __float_as_uint(r[0])), __float_as_uint(r[1]));
      um = min(min(um, __float_as_uint(r[2])), __float_as_uint(r[3]));
      um2 = min(min(um2, __float_as_uint(r[4])), __float_as_uint(r[5]));
      um2 = min(min(um2, __float_as_uint(r[6])), __float_as_uint(r[7]));
    } else if (MODE == 13) {  // 16 FFMA + 4 broadcast LDS.128
#pragma unroll
      for (int i = 0; i < CHAINS; ++i) r[i] = fmaf(r[i], a, b);
#pragma unroll
      for (int i = 0; i < 4; ++i) {
        float4 v = reinterpret_cast<const float4*>(sm)[(it * 4 + i) & 255];
        m = fminf(m, v.x + v.y + v.z + v.w);
      }
    } else if (MODE == 14) {  // VIMNMX3 only
#pragma unroll
      for (int i = 0; i < CHAINS; ++i) uacc2[i] = min(min(uacc2[i], (unsigned)it), uacc2[(i + 1) % CHAINS]);
    }
  }
  long long t1 = clock64();
  float acc = m + m2 + __uint_as_float(um) + __uint_as_float(um2) + __uint_as_float(uacc[0] ^ uacc[1] ^ uacc[2] ^ uacc[3]);
#pragma unroll
  for (int i = 0; i < CHAINS; ++i) acc += __uint_as_float(uacc2[i]);
#pragma unroll
  for (int i = 0; i < CHAINS; ++i) acc += r[i];
#pragma unroll
  for (int i = 0; i < 8; ++i) acc += (float)d[i];
#pragma unroll
  for (int i = 0; i < CHAINS / 2; ++i) {
    float2 q = *reinterpret_cast<float2*>(&p[i]);
    acc += q.x + q.y;
  }
  out[blockIdx.x * blockDim.x + threadIdx.x] = acc;
  if (threadIdx.x == 0 && blockIdx.x == 0) *cyc = t1 - t0;
}

template <int MODE>
void run(const char* name, double inst_per_iter_per_thread, double flop_per_iter) {
  int sms = 148, ctas = sms * 4, thr = 256;
  float* out; long long* cyc;
  cudaMalloc(&out, ctas * thr * sizeof(float));
  cudaMalloc(&cyc, 8);
  cudaEvent_t e0, e1; cudaEventCreate(&e0); cudaEventCreate(&e1);
  for (int w = 0; w < 3; ++w) kern<MODE><<<ctas, thr>>>(out, 0.999f, 1e-3f, cyc);
  cudaEventRecord(e0);
  const int reps = 10;
  for (int w = 0; w < reps; ++w) kern<MODE><<<ctas, thr>>>(out, 0.999f, 1e-3f, cyc);
  cudaEventRecord(e1); cudaEventSynchronize(e1);
  float ms; cudaEventElapsedTime(&ms, e0, e1); ms /= reps;
  long long hc; cudaMemcpy(&hc, cyc, 8, cudaMemcpyDeviceToHost);
  double warps_per_sm = 4.0 * thr / 32;  // 4 CTAs/SM resident (assumed)
  double winst = inst_per_iter_per_thread * ITERS * warps_per_sm;  // warp-instr per SM
  double ipc_sm = winst / (double)hc;  // using CTA0's cycle count (all CTAs concurrent)
  double tflops = flop_per_iter * ITERS * (double)ctas * thr / (ms * 1e-3) / 1e12;
  printf("%-28s ms=%.4f cycles=%lld  eff_clk=%.0f MHz  warp-inst/clk/SM=%.3f  TFLOP/s=%.2f\n", name, ms, hc,
         hc / (ms * 1e-3) / 1e6, ipc_sm, tflops);
  cudaFree(out); cudaFree(cyc);
}

int main() {
  cudaDeviceProp p; cudaGetDeviceProperties(&p, 0);
  printf("device %s SMs=%d clock=%d kHz\n", p.name, p.multiProcessorCount, p.clockRate);
  run<0>("FFMA", CHAINS, 2.0 * CHAINS);
  run<1>("FFMA2", CHAINS / 2, 2.0 * CHAINS);
  run<2>("16 FFMA + 2 FMNMX3", CHAINS + 2, 2.0 * CHAINS);
  run<3>("8 FFMA2 + 2 FMNMX3", CHAINS / 2 + 2, 2.0 * CHAINS);
  run<4>("DFMA", 8, 16.0);
  run<5>("FMNMX3", CHAINS, 0);
  run<6>("8 FFMA + 8 FMNMX", CHAINS, CHAINS);
  run<7>("FMUL", CHAINS, CHAINS);
  run<8>("16 FFMA + 2 VIMNMX3", CHAINS + 2, 2.0 * CHAINS);
  run<9>("16 FFMA + 4 LOP3", CHAINS + 4, 2.0 * CHAINS);
  run<10>("8 FFMA2 + 2 VIMNMX3", CHAINS / 2 + 2, 2.0 * CHAINS);
  run<11>("16 FFMA + 4 FMNMX3", CHAINS + 4, 2.0 * CHAINS);
  run<12>("16 FFMA + 4 VIMNMX3", CHAINS + 4, 2.0 * CHAINS);
  run<13>("16 FFMA + 4 LDS.128 bcast", CHAINS + 4, 2.0 * CHAINS);
  run<14>("VIMNMX3", CHAINS, 0);
  return 0;
}
