/*
 * gmmloc_oracle.h — CPU ORACLE (test infrastructure, NOT product code).
 *
 * A dependency-free fp64 restatement of the reference's arithmetic for the GMMLoc hot path
 * (GMM map association + structure-constrained Gauss-Newton).  Only tests/, __graft_entry__.smoke()
 * and bench.py's cpu_baseline / --impl reference legs may load this library; the product library
 * (gmmloc_b200/csrc) never links, loads or calls it.
 *
 * PARITY STATUS: "parity unpinned" for everything that lives in g2o / Eigen (neither is vendored in
 * /root/reference nor installed here; g2o_catkin is not version-pinned, .gmmloc_https.install:16-18).
 * Pinned pieces: the is_degenerated flags stored in the shipped .gmm maps (8 395 / 8 395 agree with
 * oracle_map_build, tests/test_oracle_golden.py) and the .gmm wire format (byte-exact re-encode).
 *
 * All citations are relative to /root/reference/.
 */
#pragma once
#include <stdint.h>

#ifdef __cplusplus
extern "C" {
#endif

/* Per-component constants, one row per component (all row-major, fp64).
 * Follows GaussianComponent ctor + decompose(): gmmloc/include/gmmloc/gmm/gaussian.h:30-39,
 * gmmloc/src/gmm/gaussian.cpp:36-63. */
typedef struct {
  double mean[3];
  double cov[9];
  double cov_inv[9];   /* Eigen 3x3 inverse (cofactor form)                           gaussian.h:31 */
  double det;          /* cov.determinant()                                           gaussian.h:33 */
  double scale[3];     /* eigenvalues ascending                                       gaussian.cpp:41-42 */
  double axis[9];      /* axis[r*3+c] = component r of eigenvector c                  gaussian.cpp:41 */
  double sqrt_info[9]; /* lower L, cov_inv = L L^T                                    gaussian.cpp:49 */
  int32_t is_degenerated; /* scale[0] < deg_eps (1e-4)                                gaussian.cpp:44 */
  int32_t is_salient;     /* scale[1] > 0.2 && scale[2] > 0.2                         gaussian.cpp:51-55 */
} oracle_comp_t;

typedef struct {
  double fx, fy, cx, cy, bf;
  int32_t width, height;
} oracle_camera_t; /* gmmloc_ros/cfg/v1.yaml:5-17 */

typedef struct {
  double tri_lambda2;    /* loc/tri_lambda2   = 400     v1.yaml:32 */
  double ba_lambda2;     /* loc/ba_lambda2    = 400     v1.yaml:37 */
  double tri_str_thresh; /* loc/tri_str_thresh= 0.0064  v1.yaml:35 */
  int32_t tri_check_str_chi2; /* v1.yaml:34 */
  double mdist2_gate;    /* 9.0: gmmloc_opt.cpp:230 (non-degenerate structure-edge gate in the frame solve) */
  double chi2_mono;      /* 5.991  tracking_opt.cpp:145 */
  double chi2_stereo;    /* 7.815  tracking_opt.cpp:146 */
  int32_t gn_iters;      /* 10     tracking_opt.cpp:147 */
  int32_t robust;        /* Huber on reprojection edges (localization_opt.cpp:716,746) */
  int32_t factor_mode;   /* 0 = J joint pose+points, 1 = P pose only, 2 = S points only */
} oracle_params_t;

void oracle_default_params(oracle_params_t* p);

/* a1: component constants.  mean K x 3, cov K x 9 row-major. */
void oracle_map_build(const double* mean, const double* cov, int K, double deg_eps, oracle_comp_t* out);

/* a3: d2 = (x-mu)^T cov_inv (x-mu)     gaussian.cpp:65-70 */
double oracle_chi2(const oracle_comp_t* c, const double x[3]);

/* a3+a4: brute-force association of M points (fp32 landmarks promoted to fp64) against all K
 * components; running minimum with strict '<' while scanning ascending index (gmmloc_opt.cpp:209-217,
 * gaussian_mixture.cpp:464-475) => lowest index wins ties.  assign[i] = -1 iff gate > 0 and
 * best d2 > gate.  d2_out may be NULL. nthreads <= 1: single thread. */
void oracle_associate(const oracle_comp_t* comps, int K, const float* pts, int64_t M, double gate,
                      int32_t* assign, double* d2_out, int nthreads);

/* A3: Bhattacharyya distance between 3-D components (gmm_utils.h:30-52). */
double oracle_bh3(const oracle_comp_t* a, const oracle_comp_t* b);

/* a2: neighbour lists, ordered pairs i != j with BH < thresh (gaussian_mixture.cpp:61-78).
 * CSR output: row_ptr[K+1]; returns total count; col/dist written only if non-NULL and cap >= count. */
int64_t oracle_neighbors(const oracle_comp_t* comps, int K, double thresh, int64_t* row_ptr, int32_t* col,
                         double* dist, int64_t cap);

/* A7: optimizePoint (gmmloc_opt.cpp:260-342): 3-DoF GN, 5 iterations, one stereo reprojection edge
 * (factors.cpp:116-168) + one point-to-plane edge (factors.cpp:55-64).
 * pose = qx qy qz qw tx ty tz (T_cw).  Returns res (1/0). */
int oracle_optimize_point(const double x0[3], const double obs[3], int octave, const double pose[7],
                          const oracle_comp_t* comp, double proj_z2, const oracle_camera_t* cam,
                          const oracle_params_t* prm, int iters, double pt_est[3], double* chi2_proj,
                          double* chi2_str);

/* A9: one candidate of Localization::optimizeTriangulationVec (localization_opt.cpp:27-204): 3-DoF GN,
 * `iters` (20, :169) iterations from x0, two view edges (stereo 3-D when u_r >= 0, factors.cpp:116-168, else
 * mono 2-D, factors.cpp:70-107), BOTH weighted with sigma2_inv[octave of view 1] (:126-135), plus one
 * point-to-plane edge with information tri_lambda2 (:106-107).  Outputs the estimate, err1 / err2 / chi2_str at
 * the last linearisation point; res = !(tri_check_str_chi2 && chi2_str > tri_str_thresh * tri_lambda2) &&
 * !(err1 > th1 || err2 > th2), th = 7.8 for a stereo view, 5.991 for a mono view (:118,127,133,176-186). */
int oracle_optimize_triangulation(const double x0[3], const double obs1[3], const double obs2[3], int octave1,
                                  const double pose1[7], const double pose2[7], const oracle_comp_t* comp,
                                  const oracle_camera_t* cam, const oracle_params_t* prm, int iters, double pt_est[3],
                                  double* err1, double* err2, double* chi2_str);

/* A11: the frame solve = single-free-pose specialisation of Localization::jointOptimization
 * (localization_opt.cpp:456-925) run with Gauss-Newton; association by oracle_associate at
 * iteration 0, structure-edge gates fixed at iteration 0.
 *  pose0/pose_out: qx qy qz qw tx ty tz (T_cw) fp64;  pts_w N x 3 fp32;  obs N x 3 fp32 (u v u_r,
 *  u_r < 0 => mono);  octave N.  assign_in: if non-NULL use these assignments instead of
 *  associating.  Outputs (any may be NULL): pose_out[7], pts_out N x 3 (fp64), assign N,
 *  outlier N, status (bit0 numeric failure: singular Hll / reduced system or non-finite value; such a
 *  frame is returned unsolved = normalised pose0 + input landmarks).  trace (optional): gn_iters x 8 —
 *  per iteration the pose increment (6), the largest point step max |dx| over the free points and that point's index.
 *  pt_last (optional): N, |dx|_inf of every point in the last iteration that ran. */
void oracle_frame_solve(const oracle_comp_t* comps, int K, const oracle_camera_t* cam,
                        const oracle_params_t* prm, const double pose0[7], const float* pts_w,
                        const float* obs, const uint8_t* octave, int N, const int32_t* assign_in,
                        double pose_out[7], double* pts_out, int32_t* assign, uint8_t* outlier,
                        int32_t* status, double* trace, double* pt_last);

/* A10 / f1: Tracking::optimizeCurrentPose (tracking_opt.cpp:21-217) restated with g2o's Levenberg-
 * Marquardt (SURVEY Appx B.7): pose-only, 4 rounds x 10 iterations, each round restarted from pose0
 * (:152), Huber on all edges until the end of round 3 (:177-178,201-202), inlier reclassification by
 * (float)chi2 > 5.991f / 7.815f after every round (:157-203; current outliers are re-evaluated at the
 * round's final estimate, inliers keep the error of the last LM trial, accepted or not), early exit when
 * fewer than 10 edges exist (:205-206).  valid (optional) marks features that have a map point.
 * Returns num_init_correspondences - num_bad (0 when fewer than 3 correspondences, :139-140; pose_out =
 * pose0 then).  pts_w N x 3 fp32, obs N x 3 fp32 (u v u_r; u_r < 0 => mono). */
int oracle_optimize_pose_lm(const oracle_camera_t* cam, const double pose0[7], const float* pts_w, const float* obs,
                            const uint8_t* octave, const uint8_t* valid, int N, double pose_out[7], uint8_t* outlier);

/* Batched driver over B independent frames, sharded over nthreads std::threads (the reference runs
 * each solve on the caller's thread; sharding over frames is the CPU baseline's "all cores" mode). */
void oracle_frames_solve(const oracle_comp_t* comps, int K, const oracle_camera_t* cam,
                         const oracle_params_t* prm, const double* pose0, const float* pts_w,
                         const float* obs, const uint8_t* octave, int B, int N, double* pose_out,
                         double* pts_out, int32_t* assign, uint8_t* outlier, int32_t* status, int nthreads);
/* The same plus convergence diagnostics of the iteration itself (test / bench bookkeeping: a Gauss-Newton iteration that
 * does not contract is sensitive to the last bit of ANY implementation, so parity checks treat such frames / points
 * separately).  conv (optional) B x 3: |dxi|_inf of the last iteration | the largest pose step that exceeded its
 * predecessor (iterations >= 2; 0 = monotone) | the largest point step of the last iteration.  pt_last (optional) B x N. */
void oracle_frames_solve_diag(const oracle_comp_t* comps, int K, const oracle_camera_t* cam,
                              const oracle_params_t* prm, const double* pose0, const float* pts_w,
                              const float* obs, const uint8_t* octave, int B, int N, double* pose_out,
                              double* pts_out, int32_t* assign, uint8_t* outlier, int32_t* status, int nthreads,
                              double* conv, double* pt_last);

/* ---- keyframe association pipeline (a5, a6, a7, a9) ---------------------------------------------- */

/* a5: GaussianComponent2d as GMMUtility::projectGaussian builds it (gmm_utils.cpp:121-146, gaussian.h:107-117,
 * gaussian.cpp:17-32).  cov is the 2x2 product J R S R^T J^T as computed (row-major, not symmetrised). */
typedef struct {
  int32_t id;       /* g2d->id_: index of the parent 3-D component       gaussian_mixture.cpp:319 */
  int32_t pad_;
  double mean[2];
  double cov[4];
  double cov_inv[4];
  double det;
  double scale[2];  /* eigenvalues ascending (gaussian.cpp:20-23) */
  double depth;     /* proj_d_ = (R mu + t).z                             gaussian_mixture.cpp:322-325 */
} oracle_comp2d_t;

/* a5: returns 1 and fills *out iff the projected mean is KEYPOINT_VISIBLE (pinhole_camera.cpp:127-150). */
int oracle_project_gaussian(const oracle_comp_t* c, int id, const oracle_camera_t* cam, const double pose[7],
                            oracle_comp2d_t* out);
/* 2-D Bhattacharyya distance (gmm_utils.h:30-52 instantiated for GaussianComponent2d). */
double oracle_bh2(const oracle_comp2d_t* a, const oracle_comp2d_t* b);
/* 2-D Mahalanobis distance (gaussian.h:123-126). */
double oracle_mdist2_2d(const oracle_comp2d_t* c, const double uv[2]);

/* a6: GMM::renderView(rot_c_w, t_c_w) (gaussian_mixture.cpp:271-371): view-cos cull of degenerate components
 * (|n . ray| < cos 78 deg), projection, 2-D covariance cull (both eigenvalues < 4), greedy occlusion dedupe in
 * ascending component order (min-BH kept neighbour; < 0.8 => the nearer one keeps the slot, else append), then
 * sort by depth descending (std::sort in the reference; here stable, ties keep insertion order).
 * out must hold K entries; returns the number of 2-D components. */
int oracle_render_view(const oracle_comp_t* comps, int K, const oracle_camera_t* cam, const double pose[7],
                       oracle_comp2d_t* out);

/* a7: GMM::searchCorrespondence(kpts, comps, num) (gaussian_mixture.cpp:484-534): per keypoint the `num`
 * nearest 2-D means (Euclid, ascending; brute force stands in for the nanoflann kd-tree, ties -> lower view
 * position), keeping those with 2-D MDist2 < 9.0.  cand N x num: positions in the view list, -1 padded. */
void oracle_search_correspondence(const oracle_comp2d_t* view, int V, const double* uv, int N, int num, int32_t* cand);

/* GMM::queryPoint (gaussian_mixture.cpp:545-576): 5-NN of the component means, returns ret_index[0] (Euclid-nearest;
 * the Mahalanobis minimum the reference computes is returned through min_mdist2 only).  knn: num entries, -1 padded. */
int oracle_query_point(const oracle_comp_t* comps, int K, const double pt[3], int num, int32_t* knn, double* min_mdist2);

/* a13: Frame::unproject3 (frame.cpp:27-36, pinhole_camera.cpp:20-33): pt_w = T_wc * [z (u-cx)/fx, z (v-cy)/fy, z]. */
void oracle_unproject3(const oracle_camera_t* cam, const double pose[7], const double uv[2], double depth, double pt_w[3]);

/* a9: GMMLoc::checkMapAssociation (gmmloc_opt.cpp:156-258) for one feature.  cand: 3-D component ids of the
 * feature's 2-D candidates (KeyFrame::comps_[idx], in order), ncand of them.  nbr_row_ptr / nbr_col: CSR
 * neighbour lists (oracle_neighbors).  pt3d is updated exactly where the reference mutates its argument
 * (accepted association: the optimised point; kd fallback with an accepted solve: the optimised point).
 * Returns the associated component id or -1 (nullptr). */
int oracle_check_map_association(const oracle_comp_t* comps, int K, const int64_t* nbr_row_ptr, const int32_t* nbr_col,
                                 const oracle_camera_t* cam, const oracle_params_t* prm, const double pose[7],
                                 const double obs[3], int octave, const int32_t* cand, int ncand, double pt3d[3]);

/* GMMLoc::associateMapElements (gmmloc_opt.cpp:115-154) + the per-feature association of
 * createMapPointsFromStereo (:36-113) for one keyframe: render, search (num candidates), then for every
 * feature with depth > 0: unproject3 and, when it has candidates, checkMapAssociation.
 *  obs N x 3 (u v u_r) fp32, depth N fp32, octave N.  Outputs: assoc N (component id, -1 none), pt3d N x 3
 *  (fp64; NaN for features without depth), cand N x num (3-D component ids, -1 padded; optional),
 *  *nview = number of rendered 2-D components (optional).  nthreads shards the features. */
void oracle_keyframe_associate(const oracle_comp_t* comps, int K, const int64_t* nbr_row_ptr, const int32_t* nbr_col,
                               const oracle_camera_t* cam, const oracle_params_t* prm, const double pose[7],
                               const float* obs, const float* depth, const uint8_t* octave, int N, int num,
                               int32_t* assoc, double* pt3d, int32_t* cand, int32_t* nview, int nthreads);

/* a12 / f4: Localization::jointOptimization (localization_opt.cpp:456-925) — local bundle adjustment over P
 * keyframe poses (fixed[i] != 0: VertexSE3Expmap::setFixed, the "fixed_kfs" and keyframe 0 when
 * !ba_first_as_prior) and L marginalised points, g2o Levenberg-Marquardt with BlockSolver_6_3 (Schur
 * complement over the points, SURVEY Appx B.7-8; the reduced pose system is solved with a dense Cholesky in
 * place of LinearSolverEigen's sparse one), three stages:
 *   optimize(5): every reprojection edge Huber-robustified (sqrt 5.991 / sqrt 7.815), structure edges plain;
 *   then degenerate structure edges with chi2 > tri_str_thresh * ba_lambda2 move to level 1 (:774-786);
 *   optimize(5);  then reprojection edges with chi2 > 5.991 / 7.815 (errors of the last LM trial) or
 *   non-positive depth move to level 1 and all robust kernels are dropped (:797-823);  optimize(40).
 * Edges: one per observation (point obs_pt[e] seen by keyframe obs_kf[e]; obs u v u_r, u_r < 0 => mono
 * EdgeSE3ProjectXYZ, else EdgeStereoSE3ProjectXYZ; information sigma2_inv[octave]); per associated point
 * (assoc[l] >= 0) EdgePt2GaussianDeg (information ba_lambda2) or EdgePt2Gaussian (information I);
 * prior_idx >= 0: EdgeSE3QuatPrior on that pose, measurement = its initial estimate, information
 * diag(sigma_rot^-2 I, sigma_t^-2 I), sigma_rot = 2 deg, sigma_t = 1 cm (:560-578, factors.cpp:19-53).
 * In/out: pose P x 7 (fixed poses are only normalised), pts L x 3.  Outputs (optional): obs_outlier E (the
 * vToErase test, :857-881), assoc_cleared L (structure edge rejected at the end, :842-855), iters[3] = LM
 * iterations run per stage.  Returns 0, or -1 on invalid indices. */
int oracle_local_ba(const oracle_comp_t* comps, int K, const oracle_camera_t* cam, const oracle_params_t* prm, int P,
                    double* pose, const uint8_t* fixed, int prior_idx, int L, double* pts, const int32_t* assoc, int E,
                    const int32_t* obs_pt, const int32_t* obs_kf, const float* obs, const uint8_t* octave,
                    uint8_t* obs_outlier, uint8_t* assoc_cleared, int32_t* iters);
/* the same with g2o's force-stop flag (localization_opt.cpp:541-542, 765-767, 791-795): stop > 0 raised before the call,
 * 0 never, -n raised from the n-th terminate() poll on (deterministic replay of an asynchronous stop). */
int oracle_local_ba_stop(const oracle_comp_t* comps, int K, const oracle_camera_t* cam, const oracle_params_t* prm, int P,
                    double* pose, const uint8_t* fixed, int prior_idx, int L, double* pts, const int32_t* assoc, int E,
                    const int32_t* obs_pt, const int32_t* obs_kf, const float* obs, const uint8_t* octave,
                    uint8_t* obs_outlier, uint8_t* assoc_cleared, int32_t* iters, int32_t stop);

/* g2o SE3Quat::log (se3quat.h) of pose = qx qy qz qw tx ty tz -> [omega; upsilon] */
void oracle_se3_log(const double pose[7], double xi[6]);

/* frame::sigma2_inv[octave] as the reference computes it in float (init_config.hpp:63-78). */
double oracle_sigma2_inv(int octave);

/* SE3 exp (g2o SE3Quat::exp, SURVEY Appx B.1) -> qx qy qz qw tx ty tz. */
void oracle_se3_exp(const double xi[6], double pose[7]);
/* pose_out = exp(xi) * pose   (VertexSE3Expmap::oplusImpl) */
void oracle_se3_oplus(const double xi[6], const double pose[7], double pose_out[7]);

/* .gmm wire format (protobuf_utils.cpp:12-101, GMM.proto:5-14, gmm_utils.cpp:9-119).
 * read: returns K (<0 on error); if mean/cov non-NULL and cap >= K fills them (cov row-major as the
 * loader does, gmm_utils.cpp:55-59); flags (optional) K x 2 = stored is_degenerated,is_salient. */
int oracle_gmm_read(const char* path, double* mean, double* cov, uint8_t* flags, int cap);
/* write: same byte stream saveGMMModel emits (gmm_utils.cpp:69-119). returns 0 on success. */
int oracle_gmm_write(const char* path, const double* mean, const double* cov, const uint8_t* flags, int K);

#ifdef __cplusplus
}
#endif
