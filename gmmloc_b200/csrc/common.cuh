// common.cuh — shared device/host declarations of the gmmloc_b200 CUDA library (sm_100a only).
#pragma once
#include <cuda_runtime.h>
#include <stdint.h>

#include <cstdlib>

#include "../../include/gmmloc_b200.h"

namespace gmmloc {

// Environment overrides (kernel selection, grid tuning, pipeline chunking) exist for A/B runs only: a production
// library must not change the kernel it runs because of a stray variable, so they are compiled in only with
// -DGMMLOC_DEBUG_ENV (python -m gmmloc_b200.build --debug-env).  The supported switches are the context / map
// setters of include/gmmloc_b200.h (gmmloc_ctx_set_assoc_kernel, _set_gn_precision, _set_option, gmmloc_map_create_ex).
inline const char* debug_env(const char* name) {
#ifdef GMMLOC_DEBUG_ENV
  return std::getenv(name);
#else
  (void)name;
  return nullptr;
#endif
}

// fp32 scan table: 9 floats per component,
//   c = 0..5 : rows of U = L^T (upper triangular, cov_inv = L L^T):  U00 U01 U02 U11 U12 U22
//   c = 6..8 : -(U * mean)            so that  y = U x - U mean  and  d2 = |y|^2
// stored in groups of 4 components, coefficient-major inside a group: scan[(g * 9 + c) * 4 + j] is
// coefficient c of component 4 g + j.  One LDS.128 then yields one coefficient of 4 components = two
// aligned 64-bit register pairs, the operand format of the packed fma.rn.f32x2 (FFMA2) the scan issues.
constexpr int kScanFloats = 9;
constexpr int kGroup = 4;
__host__ __device__ __forceinline__ size_t scan_index(int k, int c) {
  return ((size_t)(k / kGroup) * kScanFloats + c) * kGroup + (k % kGroup);
}
// Components are tracked in chunks of kChunk inside the scan (second-best bookkeeping granularity);
// the table is padded to a multiple of kChunk with +inf-scoring rows.
#ifndef GMMLOC_CHUNK
#define GMMLOC_CHUNK 16
#endif
constexpr int kChunk = GMMLOC_CHUNK;
constexpr int kTcK = 16;                         // fp16 K extent of one tcgen05.mma kind::f16
constexpr int kTcTileComps = 4 * kChunk;         // components per UMMA B tile (four chunks)
constexpr int kTcRows = 3 * kTcTileComps;        // UMMA N of one tile: [0,64) y0 | [64,128) y1 | [128,192) y2
constexpr int kTcTileBytes = kTcRows * kTcK * 2; // 6144 B for kChunk = 16

// ---- association grid (assoc_grid kernels) --------------------------------------------------
// Exact spatial index of the map for the association: cell c of level l lists every component k whose
// ellipsoid {x : d2_k(x) <= tmax_l} can touch the cell (conservative separating-axis test at map build), so a
// landmark that falls into c only has to score list(c): every other component has d2 > tmax_l.  The next levels
// have a larger tmax (and, the last one, a coarser grid) for the landmarks the level before cannot decide; what the
// last level cannot decide takes the brute-force exact path.  `big` lists the components too large to bin (scored
// for every landmark).
constexpr int kGridLevels = 3;
struct GridLevel {
  float lo[3] = {0.f, 0.f, 0.f};
  float inv_h = 0.f;
  int nx = 0, ny = 0, nz = 0;
  float tmax = 0.f;                 // squared Mahalanobis radius covered by the lists of this level
  int32_t* start = nullptr;         // nx * ny * nz + 1
  int32_t* list = nullptr;          // component ids, ascending inside a cell
  long long entries = 0;
};
struct GridDev {
  int nlev = 0;                     // 0: no grid (brute-force scan only)
  GridLevel lev[kGridLevels];
  int32_t* big = nullptr;
  int nbig = 0;
};

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
  float* rec = nullptr;         // K x 12: the same nine fp32 coefficients per component, contiguous (grid kernels gather them)
  float* gn_tab = nullptr;      // K x 8 fp32: structure-edge Hessian model of the mixed-precision frame GN — planar
                                // component: axis_.col(0) in [0..2]; otherwise the 6 unique entries of cov_inv in [0..5]
  GridDev grid;
  long long* nbr_row_ptr = nullptr;  // K + 1: CSR neighbour lists (gmmloc_map_build_neighbors), else null
  int32_t* nbr_col = nullptr;
  long long nbr_count = 0;
  double nbr_thresh = 0.0;
  // tensor-core pruning table (assoc_scan_tc_kernel): per kTcTileComps components one UMMA B tile of kTcRows rows
  // x kTcK fp16 (K-major, no swizzle: byte (k / 8) * (kTcRows * 16) + row * 16 + (k % 8) * 2),
  // row r * kTcTileComps + j = whitening row r of component j of the tile, 3-way hi/lo split folded into K
  uint16_t* tc_table = nullptr;  // ceil(Kpad / kTcTileComps) x kTcTileBytes
  float tc_centre[3] = {0.f, 0.f, 0.f};  // landmarks are centred with these fp32 values before the fp16 split
  float tc_err = 0.f;           // bound of |sqrt(score_tc) - sqrt(score32)| for |x|_inf <= xmax
  int tc_ok = 0;                // 0: some coefficient does not fit fp16 / chunk size unsupported -> FFMA2 scan only
  float alpha = 0.f;            // error bound of sqrt(score32) vs sqrt(exact d2), whitened units
  float xmax = 0.f;             // |x|_inf bound under which alpha holds
};

struct FrameParamsDev {
  double fx, fy, cx, cy, bf;
  double tri_lambda2, ba_lambda2, tri_str_thresh, mdist2_gate, chi2_mono, chi2_stereo;
  double delta_mono, delta_stereo;  // Huber deltas: (double)(float)sqrt(5.991 / 7.815)  tracking_opt.cpp:57-58
  float s2inv[8];                   // frame::sigma2_inv[octave], computed in float on the host (init_config.hpp:63-78)
  int gn_iters, robust, factor_mode, tri_check_str_chi2;
};

// Host-side launchers (one per kernel family); every launcher only enqueues on `stream`.
cudaError_t launch_map_build(const MapDev& m, double deg_eps, double xmax, int* d_err, unsigned int* d_alpha_bits,
                             cudaStream_t stream);  // d_err: [0] error flag, [1] alpha bits, [2] tc_err bits, [3] tc overflow flag
// builds m.grid (levels, lists, big list) from the component constants already on the device; lo / hi = bounding box
// of the means.  On any resource limit the grid is left empty (nlev = 0) and cudaSuccess is returned.
cudaError_t build_assoc_grid(MapDev& m, const double lo[3], const double hi[3], const gmmloc_grid_options_t* opt = nullptr);
void free_assoc_grid(MapDev& m);
cudaError_t launch_bh_neighbors(const MapDev& m, double thresh, int* d_counts, const long long* d_row_ptr, int32_t* d_col,
                                double* d_dist, cudaStream_t stream);
cudaError_t launch_associate(const MapDev& m, const float* d_pts, long long M, double gate, int32_t* d_assign,
                             float* d_d2, int* d_flag_count, int32_t* d_flag_idx, float* d_flag_thr, int variant,
                             cudaStream_t stream, int* launches, cudaEvent_t ev_between = nullptr, int* used_tc = nullptr,
                             int* d_sort_cnt = nullptr, int32_t* d_sort_perm = nullptr);  // scratch of the cell-sorted grid scan:
                                                                                          // cells(level 0) + 1 ints, M x 16 bytes
struct FrameArgs {
  const double* mean;       // map constants (fp64)
  const double* cov;
  const double* cov_inv;
  const double* axis;
  const uint8_t* is_deg;
  FrameParamsDev prm;
  const double* pose0;      // B x 7
  const float* pts_w;       // B x N x 3
  const float* obs;         // B x N x 3
  const uint8_t* octave;    // B x N
  const int32_t* assign;    // B x N
  int B, N;
  double* pose_out;         // B x 7
  float* pts_out;           // B x N x 3 or null
  uint8_t* outlier;         // B x N or null
  int32_t* status;          // B
  double* xws;              // B x N x 3 fp64 workspace when a frame does not fit in shared memory, else null
  uint8_t* fws;             // B x N flag workspace (same condition)
  // frame_gn_kernel only: solve the frames retry_list[0 .. *retry_count) instead of all B (CTAs beyond the count return)
  const int* retry_count = nullptr;
  const int* retry_list = nullptr;
};
cudaError_t launch_frame_gn(const FrameArgs& a, cudaStream_t stream);
// mixed-precision twin (frame_gn_mixed.cu): fp64 state / residuals, fp32 Jacobian algebra; persistent CTAs fetch frames
// through *d_next_frame (zeroed by the launcher on `stream`)
// store_ws: optional global scratch for the stored linearisation (frame_gn_mixed_scratch_bytes(N) bytes)
// prep_ws: optional global buffer (frame_gn_mixed_prep_bytes(B, N) bytes) for the class-sorted frames written by
// frame_prep_kernel; without it the solver kernel classifies and sorts its frames itself
cudaError_t launch_frame_gn_mixed(const FrameArgs& a, const float* gn_tab, int* d_next_frame, float* store_ws, size_t store_ws_bytes,
                                  unsigned char* prep_ws, size_t prep_ws_bytes, cudaStream_t stream);
size_t frame_gn_mixed_scratch_bytes(int N);
size_t frame_gn_mixed_prep_bytes(int B, int N, int mode = -1);  // 0: this batch is classified and sorted inside the solver
                                                                 // kernel; mode: GMMLOC_OPT_GN_PREP (-1 auto, 0 never, 1 always)
// world-frame twin (frame_gn_world.cu, K3w): always runs on frames prepared by frame_prep_kernel (prep_ws is mandatory:
// frame_gn_world_prep_bytes(B, N) bytes); store_ws as above (frame_gn_world_scratch_bytes(N) bytes)
cudaError_t launch_frame_gn_world(const FrameArgs& a, const float* gn_tab, int* d_next_frame, float* store_ws, size_t store_ws_bytes,
                                  unsigned char* prep_ws, size_t prep_ws_bytes, cudaStream_t stream);
size_t frame_gn_world_scratch_bytes(int N);
size_t frame_gn_world_prep_bytes(int B, int N);
size_t frame_gn_world_limit_points();
// streamed twin (frame_gn_stream.cu, K3q + K3s): frames packed by frame_pack_kernel (prep_ws: frame_gn_stream_prep_bytes(B, N)
// bytes), every per-sweep operand staged through shared memory by TMA bulk copies (store_ws: frame_gn_stream_scratch_bytes(N))
cudaError_t launch_frame_gn_stream(const FrameArgs& a, const float* gn_tab, int* d_next_frame, float* store_ws, size_t store_ws_bytes,
                                   unsigned char* prep_ws, size_t prep_ws_bytes, cudaStream_t stream);
size_t frame_gn_stream_scratch_bytes(int N);
size_t frame_gn_stream_prep_bytes(int B, int N);
size_t frame_gn_stream_limit_points();
size_t frame_gn_mixed_limit_points();  // largest N the mixed kernel takes (class-sorted frame state in shared memory)
size_t frame_gn_smem_limit_points();  // largest N whose per-frame state stays in shared memory

struct PoseLmArgs {
  FrameParamsDev prm;
  const double* pose0;    // B x 7
  const float* pts_w;     // B x N x 3
  const float* obs;       // B x N x 3
  const uint8_t* octave;  // B x N
  const uint8_t* valid;   // B x N or null
  int B, N;
  double* pose_out;       // B x 7
  uint8_t* outlier;       // B x N or null
  int32_t* inliers;       // B
  double* ws_c2;          // B x N workspace when a frame does not fit in shared memory, else null
  uint8_t* ws_st;
};
cudaError_t launch_pose_lm(const PoseLmArgs& a, cudaStream_t stream);
size_t pose_lm_smem_limit_points();

struct PointArgs {
  const double* mean;
  const double* axis;
  FrameParamsDev prm;
  const double* x0;       // M x 3
  const float* obs;       // M x 3
  const uint8_t* octave;  // M
  const double* pose;     // M x 7
  const int32_t* comp;    // M
  const double* proj_z2;  // M
  int M, iters, K;
  double* pt_est;     // M x 3
  double* chi2_proj;  // M
  double* chi2_str;   // M
  uint8_t* res;       // M
};
cudaError_t launch_point_gn(const PointArgs& a, cudaStream_t stream);

struct TriArgs {
  const double* mean;
  const double* axis;
  FrameParamsDev prm;
  const double* x0;        // M x 3
  const float* obs1;       // M x 3
  const float* obs2;       // M x 3
  const uint8_t* octave1;  // M
  const double* pose1;     // M x 7
  const double* pose2;     // M x 7
  const int32_t* comp;     // M
  int M, iters, K;
  double* pt_est;          // M x 3
  double* err1;            // M
  double* err2;            // M
  double* chi2_str;        // M
  uint8_t* res;            // M
};
cudaError_t launch_tri_gn(const TriArgs& a, cudaStream_t stream);

// ---- keyframe association pipeline (view.cu) ----------------------------------------------------------
// one rendered 2-D component = kViewRec doubles: mean(2) cov(4, row-major as computed) det depth
constexpr int kViewRec = 8;
struct ViewArgs {
  const double* mean;  // map constants
  const double* cov;
  const double* axis;
  const uint8_t* is_deg;
  int K;
  double fx, fy, cx, cy, width, height;
  double view_cos_thresh;  // cos(78 deg), computed on the host (gaussian_mixture.cpp:283)
  const double* pose;      // B x 7
  int B, cap;
  int32_t* count;          // B
  int32_t* ids;            // B x cap   depth-sorted parent component ids
  double* rec;             // B x cap x kViewRec, depth-sorted
  int32_t* wids;           // B x cap   working kept list (insertion order)
  double* wrec;            // B x cap x kViewRec
  int32_t* status;         // B   GMMLOC_VIEW_* bits
};
cudaError_t launch_render_views(const ViewArgs& a, cudaStream_t stream);

struct SearchArgs {
  const int32_t* count;  // rendered views (output of launch_render_views)
  const int32_t* ids;
  const double* rec;
  int cap;
  const float* obs;      // B x N x 3 (u v u_r; only u v are read)
  int B, N, num;
  int32_t* cand;         // B x N x num parent component ids, -1 padded
  int32_t* cand_pos;     // B x N x num positions in the view list (optional)
};
cudaError_t launch_search_corr(const SearchArgs& a, cudaStream_t stream);
// GMM::queryPoint for M device points: nearest M, knn M x query_point_knn() (optional), min_mdist2 M (optional)
cudaError_t launch_query_point(const double* mean, const double* cov_inv, int K, const double* d_pts, int M, int32_t* d_nearest,
                               int32_t* d_knn, double* d_min_mdist2, cudaStream_t stream);
int query_point_knn();
int search_corr_max_num();

struct CheckArgs {
  const double* mean;
  const double* cov_inv;
  const double* axis;
  const uint8_t* is_deg;
  int K;
  const long long* nbr_row_ptr;  // K + 1
  const int32_t* nbr_col;
  FrameParamsDev prm;
  const double* pose;     // B x 7
  const float* obs;       // B x N x 3
  const float* depth;     // B x N, or null when pt3d is an input
  const uint8_t* octave;  // B x N
  const int32_t* cand;    // B x N x num
  int B, N, num;
  double* pt3d;           // B x N x 3 (in/out)
  int32_t* assoc;         // B x N
};
cudaError_t launch_check_assoc(const CheckArgs& a, cudaStream_t stream);

// ---- local bundle adjustment (local_ba.cu) ------------------------------------------------------------
struct BaProb {
  int P, L, E, nf, n;  // poses, points, observations, free poses, n = 6 nf
  int prior;           // local index of the pose carrying the EdgeSE3QuatPrior, or -1
  long long pose_off, pt_off, obs_off, free_off;  // element offsets of this window in the flattened arrays
  long long kfo_off, blk_off, c_off, S_off;       // pose-observation lists, block pointers, contributions, S workspace
};
struct BaArgs {
  const BaProb* prob;
  int B;
  const double* mean;       // map constants
  const double* axis;
  const double* sqrt_info;
  const uint8_t* is_deg;
  FrameParamsDev prm;
  double* pose;             // sum P x 7 (in/out)
  const int32_t* fidx;      // sum P: free index or -1
  double* pts;              // sum L x 3 (in/out)
  const int32_t* assoc;     // sum L
  const int32_t* obs_pt;    // sum E (window-local point index, non-decreasing)
  const int32_t* obs_kf;    // sum E (window-local pose index)
  const float* obs;         // sum E x 3
  const uint8_t* octave;    // sum E
  const int32_t* pt_ptr;    // sum (L + 1): CSR of the observations of every point
  const int32_t* kf_ptr;    // sum (nf + 1): CSR of the observations of every free pose ...
  const int32_t* kf_obs;    // ... and their (window-local) observation indices
  const int32_t* blk_ptr;   // per window nf (nf + 1) / 2 + 1: CSR of the Schur contributions of every lower block
  const int2* contrib;      // (observation on pose fi, observation on pose fj) of one point, fi >= fj
  double* w_obs;            // workspaces
  uint8_t* level;
  double* w_pt;
  uint8_t* str_level;
  double* w_pose;
  double* w_free;
  double* S;
  uint8_t* obs_outlier;     // outputs (optional)
  uint8_t* assoc_cleared;
  int32_t* iters;           // B x 3
  int32_t* status;          // B
  const volatile int32_t* stop;  // force-stop flag in host-mapped memory (device pointer), or null
};
// cluster: CTAs per window (1, 2, 4, 8), or 0 = chosen from the number of windows
cudaError_t launch_local_ba(const BaArgs& a, cudaStream_t stream, int cluster = 0);
int local_ba_obs_ws();
int local_ba_pt_ws();
int local_ba_pose_ws();
int local_ba_free_ws();

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
