// frame_gn_classes.cuh — point classes and the prepared-frame layout shared by the mixed-precision frame Gauss-Newton
// kernels (frame_gn_mixed.cu: camera-frame K3m + the preparation kernel K3p; frame_gn_world.cu: world-frame K3w).
#pragma once
#include <stddef.h>
#include <stdint.h>

#include "common.cuh"

namespace gmmloc {
namespace {

constexpr double kFarZ = 6.0;  // metres; cond(G) ~ (z fx / bf)^2 ~ 3e3 at the shipped camera (v1.yaml:5-17)
constexpr double kNearZ = 0.1;
// Geometry the fp32 point blocks are trusted with.  The stereo projection Jacobian P has cond(P) ~ (1 + (x^2 + y^2) / z^2) z f / bf
// = (r^2 / z) f / bf (r: distance from the camera centre): it grows with depth AND with the off-axis angle.  A landmark 36 cm in
// front of the camera but 3 m to the side (frame 499 of the V1_03 replay: far outside the image, its association is a coincidence
// of the synthetic landmarks) has cond(G) ~ 1e5; its fp32 block was wrong by 15 % and derailed the pose.  Such points, points
// beyond kFarZ and points at or behind the camera plane take the fp64 path.
__host__ __device__ inline bool fp32_geometry_ok(double pz, double r2) { return (pz > kNearZ) && (pz < kFarZ) && (r2 < 2.0 * kFarZ * pz); }
// Point classes (bits 4..6 of the per-point flag byte).  The points of a frame are stably sorted by class when the
// frame is loaded, so that a warp runs ONE code path per sweep (the unsorted fp64 kernel executes with ~25 of 32
// lanes active) and the rare ill-conditioned blocks cost a few warp-iterations per sweep instead of stalling every warp.
enum : int {
  C_OFF = 0,     // takes no part (a held point when the pose is fixed)
  C_FIXED = 1,   // held point: contributes G, h to the pose system
  C_PLAIN = 2,   // free stereo point without structure edge: three equations, three unknowns — closed-form update,
                 // contributes exactly nothing to the pose system
  C_MPLANE = 3,  // free monocular point on a planar component: again three equations (u, v, plane distance) for three
                 // unknowns — closed-form update, zero Schur contribution.  (As a 3x3 normal-equation block this is
                 // the ill-conditioned case: G has rank 2 and the third direction comes from the plane alone.)
  C_STRP = 4,    // free stereo point on a planar component: the stereo edge alone determines the point (P is square),
                 // the plane adds ONE scalar equation, so D = G + lam n n^T is inverted by Sherman-Morrison in closed
                 // form: with pe = P^-1 e, z = G^-1 n, gamma = n.z, kappa = lam / (1 + lam gamma):
                 //   M = kappa n n^T,  v = kappa (es + n.pe) n,  dc = u - kappa z (n.u),  u = pe - A dxi - lam es z
  C_STR = 5,     // free stereo point with a full 3-D Gaussian edge: general 3x3 path, fp32 Hessian model
  C_MSTR = 6,    // free MONOCULAR point with a full 3-D Gaussian edge: G has rank 2, the depth direction comes from the
                 // component alone.  K3w runs it through the general fp32 path (D = G + S is as well conditioned as S);
                 // K3m treats it as C_F64
  C_F64 = 7,     // blocks the fp32 model is not trusted with (any free point beyond kFarZ): fp64 path of frame_gn_kernel.
                 // Must stay the LAST class: K3w deals the points at the end of the sorted frame round robin to its warps
  C_COUNT = 8
};
constexpr int kMaxChunks = 25;                        // histogram stride per class = 8 kMaxChunks >= N / 32 + warps per CTA

// Layout of frame f in the buffer frame_prep_kernel writes (prep_stride(N) bytes, all arrays in class-sorted order):
//   [0, 12N) landmarks f32 x 3 | [12N, 24N) observations f32 x 3 | [24N, 28N) component id i32 |
//   [28N, 30N) original index u16 | [30N, 31N) octave u8 | [31N, 32N) flags u8 (class in bits 4..6)
__host__ __device__ inline size_t prep_stride(int N) { return ((size_t)N * 32 + 15) & ~(size_t)15; }

}  // namespace

// K3p: classify + class-sort every frame of `a` into `prep` (B * prep_stride(N) bytes, 16-byte aligned)
cudaError_t launch_frame_prep(const FrameArgs& a, unsigned char* prep, cudaStream_t stream);

}  // namespace gmmloc
