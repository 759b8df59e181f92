/*
 * gmmloc_b200.h — C ABI of the B200-native GMMLoc hot path (GMM map association + the
 * structure-constrained Gauss-Newton pose/point solve it feeds).
 *
 * The reference (HyHuang1995/gmmloc @ 08eb281) has no plugin / FFI interface: callers use C++ member
 * functions on heap objects.  Each entry point below names the reference member(s) it replaces
 * (paths relative to the reference root); include/gmmloc_b200.hpp is the header-only C++ shim that
 * restores the reference's own method names on top of this ABI.
 *
 * Conventions
 *  - plain pointers and sizes only; no exceptions cross the ABI; every call returns gmmloc_status_t.
 *  - pose = 7 doubles  qx qy qz qw tx ty tz  of T_cw (world -> camera), g2o::SE3Quat order.
 *  - landmarks / observations are fp32 (coalesced HBM reads), poses and map constants fp64.
 *  - "host" entry points take host pointers (pinned or pageable) and copy inside the call;
 *    "_async" entry points take DEVICE pointers and only enqueue work on the context's stream.
 *  - a gmmloc_map_t is immutable and may be shared by any number of contexts / host threads; a
 *    gmmloc_ctx_t (stream + workspaces) must be used by one host thread at a time (the reference's
 *    tracking thread and localisation thread each own one).
 *  - there is no CPU fallback: without a CUDA device every call fails with GMMLOC_ECUDA.
 */
#ifndef GMMLOC_B200_H_
#define GMMLOC_B200_H_

#include <stddef.h>
#include <stdint.h>

#if defined(__GNUC__)
#define GMMLOC_API __attribute__((visibility("default")))
#else
#define GMMLOC_API
#endif

#ifdef __cplusplus
extern "C" {
#endif

typedef struct gmmloc_map gmmloc_map_t;
typedef struct gmmloc_ctx gmmloc_ctx_t;

typedef enum {
  GMMLOC_OK = 0,
  GMMLOC_EINVAL = 1,   /* bad argument (NULL, negative size, non-SPD covariance ...)            */
  GMMLOC_ENOMEM = 2,   /* host or device allocation failed                                      */
  GMMLOC_ECUDA = 3,    /* CUDA runtime error / no device (see gmmloc_last_error)                */
  GMMLOC_ENUMERIC = 4, /* reserved: batch-level numeric failure (per-frame failures use status) */
  GMMLOC_EIO = 5       /* .gmm file could not be read / written / parsed                        */
} gmmloc_status_t;

/* gmmloc_ros/cfg/v1.yaml:5-17 (camera::fx ... camera::bf, width, height) */
typedef struct {
  double fx, fy, cx, cy, bf;
  int32_t width, height;
} gmmloc_camera_t;

/* Hot-path parameters; defaults (gmmloc_default_params) are the shipped yaml / literals:
 *  tri_lambda2 400, ba_lambda2 400, tri_str_thresh 0.0064f, tri_check_str_chi2 1 (v1.yaml:32-37),
 *  mdist2_gate 9.0 (gmmloc_opt.cpp:230), chi2_mono 5.991, chi2_stereo 7.815, gn_iters 10
 *  (tracking_opt.cpp:145-147), robust 0, factor_mode 0. */
typedef struct {
  double tri_lambda2;
  double ba_lambda2;
  double tri_str_thresh;
  int32_t tri_check_str_chi2;
  double mdist2_gate;
  double chi2_mono;
  double chi2_stereo;
  int32_t gn_iters;
  int32_t robust;      /* Huber on reprojection edges (localization_opt.cpp:716,746) */
  int32_t factor_mode; /* GMMLOC_MODE_* */
} gmmloc_params_t;

enum { GMMLOC_MODE_JOINT = 0, GMMLOC_MODE_POSE_ONLY = 1, GMMLOC_MODE_POINTS_ONLY = 2 };

/* per-frame status bits written by gmmloc_frames_solve */
enum { GMMLOC_FRAME_NUMERIC = 1 /* singular Hll / reduced system or non-finite value: frame frozen */ };

/* per-view status bits written by gmmloc_render_views */
enum { GMMLOC_VIEW_OVERFLOW = 1 /* more than `cap` 2-D components survived: the list was truncated */ };

/* counters of the last association on a context (evidence for the exact re-rank path) */
typedef struct {
  int64_t points;        /* points scanned                                              */
  int64_t flagged;       /* points whose fp32 top-2 gap was inside the error bound      */
  int64_t kernel_launches; /* kernels launched by this context since creation           */
  int64_t flagged_brute; /* of `flagged`: points the grid index could not decide (brute-force exact re-rank) */
} gmmloc_stats_t;

GMMLOC_API const char* gmmloc_version(void);
GMMLOC_API const char* gmmloc_status_string(gmmloc_status_t s);
/* last CUDA / IO error text of the calling thread ("" if none) */
GMMLOC_API const char* gmmloc_last_error(void);
GMMLOC_API void gmmloc_default_params(gmmloc_params_t* p);
/* frame::sigma2_inv[octave] exactly as the reference computes it in float (init_config.hpp:63-78) */
GMMLOC_API float gmmloc_sigma2_inv(int octave);

/* ---- map: replaces GMMUtility::loadGMMModel + GaussianComponent ctor/decompose + GMM ctor ------
 * gmm_utils.cpp:9-67, gaussian.h:30-39, gaussian.cpp:36-63.  mean K x 3, cov K x 9 row-major (fp64).
 * All per-component constants (cov_inv, det, eigen-decomposition, is_degenerated, sqrt_info and the
 * fp32 whitening rows used by the scan) are computed on the device. */
GMMLOC_API gmmloc_status_t gmmloc_map_create(const double* mean, const double* cov, int K, double deg_eps, int device,
                                  gmmloc_map_t** out);
/* .gmm wire format: protobuf_utils.cpp:12-80, GMM.proto:5-14 */
GMMLOC_API gmmloc_status_t gmmloc_map_load_gmm(const char* path, double deg_eps, int device, gmmloc_map_t** out);
/* GMMUtility::saveGMMModel, gmm_utils.cpp:69-119 (same byte stream) */
GMMLOC_API gmmloc_status_t gmmloc_map_save_gmm(const gmmloc_map_t* map, const char* path);
GMMLOC_API void gmmloc_map_destroy(gmmloc_map_t* map);
/* GMM::countComponents, gaussian_mixture.h:153 */
GMMLOC_API int gmmloc_map_size(const gmmloc_map_t* map);
GMMLOC_API int gmmloc_map_device(const gmmloc_map_t* map);
/* gmmloc_map_create with explicit tuning of the association grid built next to the map.  Every setting
 * returns identical associations; the defaults (grid == NULL, or a field <= 0) are what gmmloc_map_create uses.
 *  enable 0: no grid, the brute-force scan serves every association;  levels 2 or 3;  cell_size in metres (level 0);
 *  tmax[l]: squared Mahalanobis radius covered by the lists of level l (25 / 144 / 400). */
typedef struct {
  int32_t enable;
  int32_t levels;
  double cell_size;
  double tmax[3];
} gmmloc_grid_options_t;
GMMLOC_API gmmloc_status_t gmmloc_map_create_ex(const double* mean, const double* cov, int K, double deg_eps, int device,
                                                const gmmloc_grid_options_t* grid, gmmloc_map_t** out);
/* Device-built constants back to the host (any pointer may be NULL). Shapes: mean K x 3, cov K x 9,
 * cov_inv K x 9, det K, scale K x 3 (ascending), axis K x 9 (axis[r*3+c] = component r of eigenvector c),
 * sqrt_info K x 9 (lower L, cov_inv = L L^T), is_degenerated K, is_salient K.
 * ≙ GaussianComponent public fields, gaussian.h:69-90. */
GMMLOC_API gmmloc_status_t gmmloc_map_get_components(const gmmloc_map_t* map, double* mean, double* cov, double* cov_inv,
                                          double* det, double* scale, double* axis, double* sqrt_info,
                                          uint8_t* is_degenerated, uint8_t* is_salient);
/* Neighbour lists of GMM::GMM (gaussian_mixture.cpp:61-78): for every component i the components j != i
 * with Bhattacharyya distance BH(i, j) < thresh (GMMUtility::BHCoefficient, gmm_utils.h:30-52;
 * thresh = gmmmap::neighbor_dist_thresh, 2.5 in v1.yaml:27), in ascending j (the reference's construction
 * order), as CSR.  All K x K pairs are evaluated on the device.  row_ptr has K+1 entries; *count receives
 * the total; col / dist (dist optional) are filled only when col != NULL and cap >= *count, so callers
 * size them with a first call passing col = NULL. */
GMMLOC_API gmmloc_status_t gmmloc_map_neighbors(const gmmloc_map_t* map, double thresh, int64_t* row_ptr, int32_t* col,
                                                double* dist, int64_t cap, int64_t* count);
/* Builds the neighbour lists above once and keeps them on the device inside the map (the reference builds
 * GaussianComponent::nbs_ in the GMM constructor, gaussian_mixture.cpp:61-78); required by
 * gmmloc_check_map_association / gmmloc_keyframes_associate.  Not thread-safe: call it before the map is
 * shared.  *count (optional) receives the number of (i, j) pairs. */
GMMLOC_API gmmloc_status_t gmmloc_map_build_neighbors(gmmloc_map_t* map, double thresh, int64_t* count);
/* GMM::queryPoint(const Vector3d&, std::vector<int>&) (gaussian_mixture.h:169, gaussian_mixture.cpp:545-576) for M points:
 * the 5 nearest component MEANS by squared Euclid distance (the reference asks its nanoflann kd-tree; here all K means
 * are scanned, equal distances -> lower index first).  The reference then computes the minimum Mahalanobis distance
 * over those five and discards it: res = { ret_index[0] }, the Euclid-nearest mean (:574) — the quirk is kept:
 *  nearest M: what the reference returns in res[0];  knn M x 5 (optional): the list, ascending distance;
 *  min_mdist2 M (optional): the discarded Mahalanobis minimum.   pts M x 3 (fp64, host). */
GMMLOC_API gmmloc_status_t gmmloc_map_query_point(gmmloc_ctx_t* ctx, const gmmloc_map_t* map, const double* pts, int M,
                                                  int32_t* nearest, int32_t* knn, double* min_mdist2);
/* error bound of the fp32 scan in whitened units (|sqrt(s32) - sqrt(d2)| <= alpha for |x|_inf <= xmax) */
GMMLOC_API gmmloc_status_t gmmloc_map_scan_bound(const gmmloc_map_t* map, double* alpha, double* xmax);

/* ---- context ------------------------------------------------------------------------------- */
GMMLOC_API gmmloc_status_t gmmloc_ctx_create(int device, gmmloc_ctx_t** out);
GMMLOC_API void gmmloc_ctx_destroy(gmmloc_ctx_t* ctx);
/* cudaStream_t of the context (for event timing / interop) */
GMMLOC_API void* gmmloc_ctx_stream(gmmloc_ctx_t* ctx);
GMMLOC_API gmmloc_status_t gmmloc_ctx_synchronize(gmmloc_ctx_t* ctx);
GMMLOC_API gmmloc_status_t gmmloc_ctx_stats(gmmloc_ctx_t* ctx, gmmloc_stats_t* out);
/* device time (ms, CUDA events on the context's stream) of the kernels of the LAST association / frame
 * solve enqueued on ctx: ms[0] association scan (K1), ms[1] exact re-rank (K2), ms[2] frame GN (K3; 0 if
 * the last call was an association only).  Synchronises the stream.  The reference's counterpart is its
 * timing::Timer registry ("loc/render_view", "map/search_corr", "loc/ba": gmmloc_opt.cpp:123,130,
 * localization.cpp:101). */
GMMLOC_API gmmloc_status_t gmmloc_ctx_kernel_times(gmmloc_ctx_t* ctx, float ms[3]);

/* Association scan kernel of this context.  All of them return identical assignments (the exact argmin of the
 * reference's d2 over ALL components, lowest index on ties):
 *   GMMLOC_ASSOC_GRID   grid-indexed scan: an exact spatial index built at map creation lists, per cell, every
 *                       component whose d2 can be <= a radius anywhere in the cell; a landmark scores its cell list
 *                       only (~K/300 of the work) and falls back to the brute-force path when the index cannot decide;
 *   GMMLOC_ASSOC_FFMA2  brute-force FFMA2 scan over all K components (TMA-staged table);
 *   GMMLOC_ASSOC_TENSOR tensor-core pruning scan (tcgen05) over all K components, candidates re-scored exactly;
 *   GMMLOC_ASSOC_AUTO   GRID when the map has an index, else FFMA2.
 * A forced kind that is unavailable for the map falls back to FFMA2. */
enum { GMMLOC_ASSOC_AUTO = 0, GMMLOC_ASSOC_FFMA2 = 1, GMMLOC_ASSOC_TENSOR = 2, GMMLOC_ASSOC_GRID = 3 };
GMMLOC_API gmmloc_status_t gmmloc_ctx_set_assoc_kernel(gmmloc_ctx_t* ctx, int kind);
/* kernel the last association on ctx ran: 0 FFMA2 scan, 1 tensor-core pruning scan, 2 grid-indexed scan */
GMMLOC_API int gmmloc_ctx_last_assoc_kernel(gmmloc_ctx_t* ctx);
/* Arithmetic of the per-frame Gauss-Newton kernel behind gmmloc_frames_solve (Localization::jointOptimization with a
 * single free pose, localization_opt.cpp:456-925):
 *   GMMLOC_GN_MIXED            state, residuals and the 6x6 reduced system in fp64, Jacobian / Hessian-model algebra
 *                              in fp32 — same fixed point (poses agree with the fp64 solve to ~1e-9), ~3x faster;
 *   GMMLOC_GN_WORLD            the same precision split with the per-point blocks written in the world frame (the
 *                              structure edge needs no rotation, the point step no pose; 16 instead of 12 warps per SM);
 *   GMMLOC_GN_FP64             everything in fp64;
 *   GMMLOC_GN_STREAM (default) the world-frame kernel with every per-sweep operand (component constants resolved once per
 *                              frame, stored linearisation) staged through shared memory by TMA bulk copies; frames of more
 *                              than ~4 300 landmarks run GMMLOC_GN_MIXED. */
enum { GMMLOC_GN_FP64 = 0, GMMLOC_GN_MIXED = 1, GMMLOC_GN_WORLD = 2, GMMLOC_GN_STREAM = 3 };
GMMLOC_API gmmloc_status_t gmmloc_ctx_set_gn_precision(gmmloc_ctx_t* ctx, int precision);
/* GMMLOC_GN_* of the kernel the last gmmloc_frames_solve[_async] on ctx ran (frame size decides between the selected kernel
 * and its fall-backs), -1 before the first call */
GMMLOC_API int gmmloc_ctx_last_gn_kernel(gmmloc_ctx_t* ctx);

/* Tuning switches of a context for A/B measurements; every setting returns identical results.  (Environment variables
 * are NOT consulted by a production build: they are compiled in only with -DGMMLOC_DEBUG_ENV.)
 *   GMMLOC_OPT_GN_PREP          -1 automatic (default), 0 never, 1 always run the frame-preparation kernel ahead of the
 *                               mixed-precision frame GN;
 *   GMMLOC_OPT_E2E_CHUNK        frames per pipeline chunk of the host entry point gmmloc_frames_solve (0 = automatic:
 *                               6 frames per SM);
 *   GMMLOC_OPT_E2E_SHORT_FIRST  1: the first pipeline chunk is a third of the others (default 0);
 *   GMMLOC_OPT_BA_CLUSTER       CTAs (SMs) per window of gmmloc_local_ba: 1, 2, 4 or 8 (thread-block cluster), 0 = automatic
 *                               (8 for a single window ... 1 once the windows of a call fill the GPU);
 *   GMMLOC_OPT_E2E_TAPER        1 (default): large batches run a graded chunk schedule in gmmloc_frames_solve — the first
 *                               chunk is 2 frames per SM, every next one a frame per SM longer (up to twice
 *                               GMMLOC_OPT_E2E_CHUNK), the last takes the rest; 0: uniform chunks. */
enum { GMMLOC_OPT_GN_PREP = 0, GMMLOC_OPT_E2E_CHUNK = 1, GMMLOC_OPT_E2E_SHORT_FIRST = 2, GMMLOC_OPT_BA_CLUSTER = 3, GMMLOC_OPT_E2E_TAPER = 4 };
GMMLOC_API gmmloc_status_t gmmloc_ctx_set_option(gmmloc_ctx_t* ctx, int option, int64_t value);
/* The chunk schedule gmmloc_frames_solve would run for a batch of B frames on a device with `sms` multiprocessors under the
 * options above (chunk_opt = GMMLOC_OPT_E2E_CHUNK, taper = GMMLOC_OPT_E2E_TAPER, short_first = GMMLOC_OPT_E2E_SHORT_FIRST):
 * writes up to `cap` chunk sizes (frames, in pipeline order; they sum to B) and returns the number of chunks, -1 on a bad
 * argument.  Pure host arithmetic — no device, no context; results of gmmloc_frames_solve do not depend on the schedule. */
GMMLOC_API int32_t gmmloc_frames_solve_chunks(int32_t B, int32_t sms, int32_t chunk_opt, int32_t taper, int32_t short_first,
                                              int32_t* sizes, int32_t cap);

/* page-locked host memory for the host entry points (pageable buffers also work, via staging) */
GMMLOC_API gmmloc_status_t gmmloc_host_alloc(size_t bytes, void** out);
GMMLOC_API void gmmloc_host_free(void* p);

/* ---- association: replaces GaussianComponent::chi2 / MDist2 scanned over all components with the
 * reference's argmin convention (gaussian.cpp:65-70, gaussian.h:53-56; strict '<' ascending scan =>
 * lowest index wins ties, gmmloc_opt.cpp:175-197,209-217).  M = B*N points.
 *  assign[i] = argmin_k d2_k(x_i)   (exact: identical to an fp64 scan), or -1 iff gate > 0 and the
 *  best d2 > gate (gmmloc_opt.cpp:230).  d2 (optional) = (float) of the fp64 d2 of the winner. */
GMMLOC_API gmmloc_status_t gmmloc_associate(gmmloc_ctx_t* ctx, const gmmloc_map_t* map, const float* pts_w, int B, int N,
                                 double gate, int32_t* assign, float* d2);
GMMLOC_API gmmloc_status_t gmmloc_associate_async(gmmloc_ctx_t* ctx, const gmmloc_map_t* map, const float* d_pts_w,
                                       int64_t M, double gate, int32_t* d_assign, float* d_d2);

/* ---- frame solve: association + single-free-pose specialisation of Localization::jointOptimization
 * (localization_opt.cpp:456-925; edges EdgeStereoSE3ProjectXYZ / EdgeSE3ProjectXYZ + EdgePt2GaussianDeg /
 * EdgePt2Gaussian, factors.cpp:5-17,55-64) run with Gauss-Newton for params->gn_iters iterations;
 * factor_mode POSE_ONLY gives Tracking::optimizeCurrentPose's system (tracking_opt.cpp:21-217),
 * POINTS_ONLY the per-point structure-constrained solves.  B independent frames of N landmarks.
 *  pose0 B x 7, pts_w B x N x 3, obs B x N x 3 (u v u_r; u_r < 0 => monocular), octave B x N.
 *  outputs: pose_out B x 7, pts_out B x N x 3 (optional), assign B x N, outlier B x N (optional),
 *  status B (GMMLOC_FRAME_* bits). */
GMMLOC_API gmmloc_status_t gmmloc_frames_solve(gmmloc_ctx_t* ctx, const gmmloc_map_t* map, const gmmloc_camera_t* cam,
                                    const gmmloc_params_t* params, const double* pose0, const float* pts_w,
                                    const float* obs, const uint8_t* octave, int B, int N, double* pose_out,
                                    float* pts_out, int32_t* assign, uint8_t* outlier, int32_t* status);
GMMLOC_API gmmloc_status_t gmmloc_frames_solve_async(gmmloc_ctx_t* ctx, const gmmloc_map_t* map, const gmmloc_camera_t* cam,
                                          const gmmloc_params_t* params, const double* d_pose0,
                                          const float* d_pts_w, const float* d_obs, const uint8_t* d_octave,
                                          int B, int N, double* d_pose_out, float* d_pts_out, int32_t* d_assign,
                                          uint8_t* d_outlier, int32_t* d_status);

/* ---- pose-only solve with the reference's exact schedule: replaces Tracking::optimizeCurrentPose
 * (tracking_opt.cpp:21-217) for B independent frames: g2o Levenberg-Marquardt on one SE3 vertex, one mono
 * (u_r < 0) or stereo reprojection edge per feature that has a map point (valid[i] != 0; valid == NULL =>
 * all), Huber sqrt(5.991) / sqrt(7.815), 4 rounds x 10 iterations each restarted from pose0, inlier
 * reclassification by chi2 > 5.991 / 7.815 after every round, robust kernel dropped for the last round.
 *  outputs: pose_out B x 7, outlier B x N (Frame::is_outlier_, optional), inliers B (the function's return
 *  value: correspondences - outliers; 0 and pose_out = pose0 when fewer than 3 correspondences). */
GMMLOC_API gmmloc_status_t gmmloc_poses_optimize(gmmloc_ctx_t* ctx, const gmmloc_camera_t* cam, const double* pose0,
                                                 const float* pts_w, const float* obs, const uint8_t* octave,
                                                 const uint8_t* valid, int B, int N, double* pose_out, uint8_t* outlier,
                                                 int32_t* inliers);

/* ---- batched point solves: replaces GMMLoc::optimizePoint (gmmloc_opt.cpp:260-342), M independent
 * 3-DoF GN problems (one stereo reprojection edge + one point-to-plane edge), `iters` iterations (5).
 *  x0 M x 3 (fp64), obs M x 3 (u v u_r), octave M, pose M x 7, comp M (component index), proj_z2 M.
 *  outputs: pt_est M x 3, chi2_proj M, chi2_str M, res M (1 = accepted). */
GMMLOC_API gmmloc_status_t gmmloc_points_solve(gmmloc_ctx_t* ctx, const gmmloc_map_t* map, const gmmloc_camera_t* cam,
                                    const gmmloc_params_t* params, const double* x0, const float* obs,
                                    const uint8_t* octave, const double* pose, const int32_t* comp,
                                    const double* proj_z2, int M, int iters, double* pt_est, double* chi2_proj,
                                    double* chi2_str, uint8_t* res);

/* ---- batched two-view triangulation solves: replaces the per-candidate body of
 * Localization::optimizeTriangulationVec (localization_opt.cpp:27-204): M independent 3-DoF GN problems, two
 * view edges (u_r >= 0: stereo 3-D, factors.cpp:116-168; else mono 2-D, factors.cpp:70-107; both weighted
 * with sigma2_inv[octave1], :126-135) + one point-to-plane edge (information tri_lambda2), `iters` (20).
 *  x0 M x 3 (fp64), obs1 / obs2 M x 3 (u v u_r), octave1 M, pose1 / pose2 M x 7 (T_cw of the two
 *  keyframes), comp M (candidate component index; the reference only tries degenerate components).
 *  outputs: pt_est M x 3, err1 / err2 / chi2_str M (chi2 of the three edges), res M (1 = accepted:
 *  chi2_str <= tri_str_thresh * tri_lambda2 and err <= 7.8 (stereo view) / 5.991 (mono view)).
 *  The caller keeps, per match, the accepted candidate with the smallest err1 + err2 (:189-195). */
GMMLOC_API gmmloc_status_t gmmloc_triangulation_solve(gmmloc_ctx_t* ctx, const gmmloc_map_t* map, const gmmloc_camera_t* cam,
                                                      const gmmloc_params_t* params, const double* x0, const float* obs1,
                                                      const float* obs2, const uint8_t* octave1, const double* pose1,
                                                      const double* pose2, const int32_t* comp, int M, int iters,
                                                      double* pt_est, double* err1, double* err2, double* chi2_str,
                                                      uint8_t* res);


/* ---- keyframe association pipeline: GMMLoc::associateMapElements (gmmloc_opt.cpp:115-154) + the
 * per-feature association inside GMMLoc::createMapPointsFromStereo (:36-113), batched over B keyframes. ---- */

/* GMM::renderView(rot_c_w, t_c_w) (gaussian_mixture.cpp:271-371) for B views: view-cos cull of degenerate
 * components, GMMUtility::projectGaussian (gmm_utils.cpp:121-146), 2-D covariance cull, greedy occlusion
 * dedupe in ascending component order (2-D Bhattacharyya < 0.8 => the nearer one keeps the slot), sort by depth
 * descending.  The rendered views stay on the context (the reference keeps them in GMM::components2d_) for
 * gmmloc_search_correspondence; `cap` bounds the 2-D components kept per view (cap >= map size can never
 * overflow; a view that overflows is truncated and flagged GMMLOC_VIEW_OVERFLOW in status).
 *  pose B x 7 (T_cw).  Host outputs (each optional): count B, status B, ids B x cap (parent component index,
 *  GaussianComponent2d::id_), mean2d B x cap x 2, cov2d B x cap x 4 (row-major), depth B x cap (proj_d_). */
GMMLOC_API gmmloc_status_t gmmloc_render_views(gmmloc_ctx_t* ctx, const gmmloc_map_t* map, const gmmloc_camera_t* cam,
                                               const double* pose, int B, int cap, int32_t* count, int32_t* status,
                                               int32_t* ids, double* mean2d, double* cov2d, double* depth);

/* GMM::searchCorrespondence(kpts, comps, num) (gaussian_mixture.cpp:484-534) on the views last rendered on
 * ctx: per keypoint the `num` (<= 8; the reference uses 5) nearest 2-D means in ascending distance (brute
 * force instead of the nanoflann kd-tree; equal distances -> lower view position first), keeping those with
 * 2-D MDist2 < 9.0.   obs B x N x 3 (u v u_r; only u v are read).
 *  cand B x N x num: parent component indices (gc2d->parent_), -1 padded;  cand_pos (optional): positions in
 *  the rendered view list. */
GMMLOC_API gmmloc_status_t gmmloc_search_correspondence(gmmloc_ctx_t* ctx, const float* obs, int B, int N, int num,
                                                        int32_t* cand, int32_t* cand_pos);

/* GMMLoc::checkMapAssociation (gmmloc_opt.cpp:156-258), B x N features: optimizePoint against every candidate,
 * smallest accepted chi2_proj wins, Mahalanobis scan of the winner's neighbour list, optional re-solve, 9.0
 * gate; when no candidate is accepted: kd fallback (nearest mean, degenerate only), which may move the point
 * but never associates.   Needs gmmloc_map_build_neighbors.
 *  pose B x 7, obs B x N x 3 (u v u_r), octave B x N, cand B x N x num (component indices, -1 = none).
 *  pt3d B x N x 3 in/out (fp64): updated exactly where the reference mutates its argument.
 *  assoc B x N: associated component index or -1 (nullptr). */
GMMLOC_API gmmloc_status_t gmmloc_check_map_association(gmmloc_ctx_t* ctx, const gmmloc_map_t* map,
                                                        const gmmloc_camera_t* cam, const gmmloc_params_t* params,
                                                        const double* pose, const float* obs, const uint8_t* octave,
                                                        const int32_t* cand, int B, int N, int num, double* pt3d,
                                                        int32_t* assoc);

/* The three steps above chained on the device for B keyframes of N features: render, search (num candidates),
 * Frame::unproject3 (frame.cpp:27-36) of every feature with depth > 0, checkMapAssociation.
 *  depth B x N (Feature::depth; <= 0: no stereo depth => assoc -1, pt3d NaN).
 *  outputs: assoc B x N, pt3d B x N x 3 (fp64), cand B x N x num (optional), nview B (optional: rendered 2-D
 *  components per keyframe), vstatus B (optional: GMMLOC_VIEW_* bits). */
GMMLOC_API gmmloc_status_t gmmloc_keyframes_associate(gmmloc_ctx_t* ctx, const gmmloc_map_t* map, const gmmloc_camera_t* cam,
                                                      const gmmloc_params_t* params, const double* pose, const float* obs,
                                                      const float* depth, const uint8_t* octave, int B, int N, int num,
                                                      int cap, int32_t* assoc, double* pt3d, int32_t* cand,
                                                      int32_t* nview, int32_t* vstatus);

/* ---- local bundle adjustment: replaces Localization::jointOptimization (localization_opt.cpp:456-925) for B
 * independent covisibility windows (the reference solves one per new keyframe on the localisation thread).
 * g2o Levenberg-Marquardt with BlockSolver_6_3 (Schur complement over the map points), stages optimize(5) /
 * optimize(5) / optimize(40) with the reference's gating: degenerate structure edges with chi2 > tri_str_thresh *
 * ba_lambda2 leave after stage 1 (:774-786); reprojection edges with chi2 > 5.991 (mono) / 7.815 (stereo) or
 * non-positive depth leave after stage 2, when the Huber kernels are dropped (:797-823).
 * Window b owns poses [pose_ptr[b], pose_ptr[b+1]), points [pt_ptr[b], pt_ptr[b+1]) and observations
 * [obs_ptr[b], obs_ptr[b+1]); obs_pt / obs_kf are window-local indices and obs_pt must be non-decreasing inside a
 * window (observations grouped by point, the order of the reference's loop over local_mappts, :642-766).
 *  pose sum P x 7 in/out (T_cw; fixed[i] != 0: setFixed(true), i.e. the fixed_kfs and keyframe 0 without
 *  ba_first_as_prior; fixed poses are only normalised);  prior_idx B (optional): window-local pose that carries the
 *  EdgeSE3QuatPrior of :560-578 (measurement = its initial estimate, sigma_rot 2 deg, sigma_t 1 cm), -1 = none;
 *  pts sum L x 3 in/out (fp64);  assoc sum L: MapPoint::asscociations_ as a component index, -1 = none
 *  (EdgePt2GaussianDeg with information ba_lambda2 for degenerate components, EdgePt2Gaussian otherwise);
 *  obs sum E x 3 (u v u_r; u_r < 0 => EdgeSE3ProjectXYZ, else EdgeStereoSE3ProjectXYZ), octave sum E.
 * Outputs (optional): obs_outlier sum E (the vToErase test, :857-881), assoc_cleared sum L (associations the
 * reference clears, :842-855), iters B x 3 (LM iterations run per stage), status B (GMMLOC_FRAME_NUMERIC when a
 * linear solve failed inside the LM loop; g2o carries on, and so does this).
 *  stop (optional, host pointer): the reference's `bool* pbStopFlag` (optimizer.setForceStopFlag, :541-542), raised by
 *  another host thread to abort the BA (Localization::insertKeyFrame, localization.cpp:401-405).  It is read where the
 *  reference / g2o read it: before the first stage (raised => the window is returned untouched, iters 0 0 0, :765-767),
 *  ahead of every LM iteration and after every rejected LM trial (g2o's terminate()), and after stage 2 (raised => no
 *  reprojection gating, no stage 3, :791-795; results are still written).  The calling thread mirrors *stop into
 *  device-mapped memory while the kernel runs.  *stop > 0: raised; 0: not raised; a negative value -n is a deterministic
 *  replay hook: the flag reads as raised from its n-th poll on (used by the parity tests). */
GMMLOC_API gmmloc_status_t gmmloc_local_ba(gmmloc_ctx_t* ctx, const gmmloc_map_t* map, const gmmloc_camera_t* cam,
                                           const gmmloc_params_t* params, int B, const int32_t* pose_ptr, double* pose,
                                           const uint8_t* fixed, const int32_t* prior_idx, const int32_t* pt_ptr, double* pts,
                                           const int32_t* assoc, const int32_t* obs_ptr, const int32_t* obs_pt,
                                           const int32_t* obs_kf, const float* obs, const uint8_t* octave,
                                           uint8_t* obs_outlier, uint8_t* assoc_cleared, int32_t* iters, int32_t* status,
                                           const volatile int32_t* stop);

#ifdef __cplusplus
}
#endif
#endif /* GMMLOC_B200_H_ */
