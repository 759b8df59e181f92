// common.cuh — shared device/host declarations of the gmmloc_b200 CUDA library (sm_100a only).
#pragma once
#include <cuda_runtime.h>
#include <stdint.h>

#include "../../include/gmmloc_b200.h"

namespace gmmloc {

// Number of floats per component in the fp32 scan table:
//   [0..5] = rows of U = L^T (upper triangular, cov_inv = L L^T):  U00 U01 U02 U11 U12 U22
//   [6..8] = -c, c = U * mean       so that  y = U x - c  and  d2 = |y|^2
constexpr int kScanFloats = 9;
// Components are tracked in chunks of kChunk inside the scan (second-best bookkeeping granularity);
// the table is padded to a multiple of kChunk with +inf-scoring rows.
constexpr int kChunk = 16;

struct MapDev {
  int K = 0;      // real components
  int Kpad = 0;   // padded to a multiple of kChunk
  double* mean = nullptr;       // K x 3
  double* cov = nullptr;        // K x 9
  double* cov_inv = nullptr;    // K x 9
  double* det = nullptr;        // K
  double* scale = nullptr;      // K x 3
  double* axis = nullptr;       // K x 9
  double* sqrt_info = nullptr;  // K x 9
  uint8_t* is_deg = nullptr;    // K
  uint8_t* is_sal = nullptr;    // K
  float* scan = nullptr;        // Kpad x kScanFloats
  float alpha = 0.f;            // error bound of sqrt(score32) vs sqrt(exact d2), whitened units
  float xmax = 0.f;             // |x|_inf bound under which alpha holds
};

struct FrameParamsDev {
  double fx, fy, cx, cy, bf;
  double ba_lambda2, tri_str_thresh, mdist2_gate, chi2_mono, chi2_stereo;
  double delta_mono, delta_stereo;  // Huber deltas: (double)(float)sqrt(5.991 / 7.815)  tracking_opt.cpp:57-58
  int gn_iters, robust, factor_mode;
};

// ---- exact (non-contracted) fp64 helpers: the re-rank and the map build must replay the CPU
// oracle's operation order bit for bit, so every product and sum is an explicit IEEE rn op. ----
__device__ __forceinline__ double dmul(double a, double b) { return __dmul_rn(a, b); }
__device__ __forceinline__ double dadd(double a, double b) { return __dadd_rn(a, b); }
__device__ __forceinline__ double dsub(double a, double b) { return __dsub_rn(a, b); }

// d2 = delta^T * cov_inv * delta with the reference's evaluation order (gaussian.cpp:65-70):
// r = delta^T A (column sums, left to right), then r . delta.
__device__ __forceinline__ double chi2_exact(const double* __restrict__ mean3, const double* __restrict__ A,
                                             double x0, double x1, double x2) {
  const double d0 = dsub(x0, mean3[0]), d1 = dsub(x1, mean3[1]), d2 = dsub(x2, mean3[2]);
  const double r0 = dadd(dadd(dmul(d0, A[0]), dmul(d1, A[3])), dmul(d2, A[6]));
  const double r1 = dadd(dadd(dmul(d0, A[1]), dmul(d1, A[4])), dmul(d2, A[7]));
  const double r2 = dadd(dadd(dmul(d0, A[2]), dmul(d1, A[5])), dmul(d2, A[8]));
  return dadd(dadd(dmul(r0, d0), dmul(r1, d1)), dmul(r2, d2));
}

// fp32 whitened score; the ONE definition used by both the scan and the re-rank so that their values
// are bit-identical (explicit fma / mul intrinsics: no compiler re-association or contraction).
__device__ __forceinline__ float score32(float u00, float u01, float u02, float u11, float u12, float u22,
                                         float nc0, float nc1, float nc2, float x0, float x1, float x2) {
  const float y0 = __fmaf_rn(u00, x0, __fmaf_rn(u01, x1, __fmaf_rn(u02, x2, nc0)));
  const float y1 = __fmaf_rn(u11, x1, __fmaf_rn(u12, x2, nc1));
  const float y2 = __fmaf_rn(u22, x2, nc2);
  return __fmaf_rn(y2, y2, __fmaf_rn(y1, y1, __fmul_rn(y0, y0)));
}

// relative rounding slack of sqrt(score32) on top of alpha (3 roundings in the sum of squares -> 1.5u,
// plus the fp32->fp64 sqrt), taken as 4 * 2^-24
constexpr double kRelEps = 4.0 * 5.9604644775390625e-08;

}  // namespace gmmloc
