// gmmloc_b200.hpp — header-only C++17 shim that restores the reference's own names on top of the C ABI
// (include/gmmloc_b200.h).  Host code stays C++/ROS; this header needs neither Eigen nor g2o — vectors are
// std::array so that Eigen users pass `.data()` (see INTEGRATION.md for the exact call-site patches).
//
// Reference surface mirrored (paths relative to the reference root):
//   gmmloc::GMM                         gmmloc/include/gmmloc/gmm/gaussian_mixture.h:98-170
//   gmmloc::GMMUtility::load/saveGMMModel  gmmloc/include/gmmloc/gmm/gmm_utils.h:13-16
//   gmmloc::GaussianComponent (fields)  gmmloc/include/gmmloc/gmm/gaussian.h:60-92
//   gmmloc::StrOptStat                  gmmloc/include/gmmloc/types/map.h:30-35
//   GMMLoc::optimizePoint               gmmloc/include/gmmloc/gmmloc.h:43-46
//   Tracking::optimizeCurrentPose       gmmloc/include/gmmloc/modules/tracking.h:51
//   Localization::jointOptimization     gmmloc/include/gmmloc/modules/localization.h:61 (single free pose: optimizeFrames;
//                                       full covisibility window: jointOptimization(LocalWindow&))
//   GMM::renderView / getComponents2d / searchCorrespondence   gaussian_mixture.h:122,135-137,147
//   GMMLoc::checkMapAssociation / associateMapElements          gmmloc/include/gmmloc/gmmloc.h:48-51
//   Localization::optimizeTriangulationVec                      modules/localization.h:58-59
#pragma once
#include <algorithm>
#include <array>
#include <cstdint>
#include <memory>
#include <stdexcept>
#include <string>
#include <vector>

#include "gmmloc_b200.h"

namespace gmmloc_b200 {

using Vector3d = std::array<double, 3>;
using Matrix3d = std::array<double, 9>;  // row-major

// g2o::SE3Quat storage order: rotation quaternion (x y z w) then translation; T_cw maps world -> camera
struct SE3Quat {
  std::array<double, 7> v{0, 0, 0, 1, 0, 0, 0};
  const double* data() const { return v.data(); }
  double* data() { return v.data(); }
};

// the fields of gmmloc::Feature the hot path reads (feature.h:7-41): undistorted keypoint, right-image
// column (< 0 => monocular), stereo depth (<= 0: none) and pyramid octave
struct Feature {
  double u = 0, v = 0, u_right = -1;
  int octave = 0;
  float depth = -1.0f;
};

// gmmloc::GaussianComponent2d as renderView leaves it in GMM::components2d_ (gaussian.h:95-150): id_ is the
// index of the parent 3-D component (parent_ in the reference), proj_d_ its camera-frame depth
struct GaussianComponent2d {
  uint32_t id_ = 0;
  double proj_d_ = 0.0;
  std::array<double, 2> mean_{};
  std::array<double, 4> cov_{};  // row-major
  const std::array<double, 2>& mean() const { return mean_; }
  const std::array<double, 4>& cov() const { return cov_; }
};
using GaussianComponents2d = std::vector<GaussianComponent2d>;

// gmmloc::StrOptStat (map.h:30-35)
struct StrOptStat {
  bool res = false;
  double chi2_proj = 0, chi2_str = 0;
  Vector3d pt_est{0, 0, 0};
};

// host copy of one component's constants (public fields of gmmloc::GaussianComponent, gaussian.h:69-90)
struct GaussianComponent {
  Vector3d mean_{}, scale_{};
  Matrix3d cov_{}, cov_inv_{}, sqrt_info_{}, axis_{};
  double det_ = 0;
  bool is_degenerated = false, is_salient = false;
  int id = -1;
  const Vector3d& mean() const { return mean_; }
  const Matrix3d& cov_inv() const { return cov_inv_; }
  const double& det() const { return det_; }
};

class Error : public std::runtime_error {
 public:
  Error(gmmloc_status_t s, const std::string& what) : std::runtime_error(what), status(s) {}
  gmmloc_status_t status;
};
inline void check(gmmloc_status_t s) {
  if (s != GMMLOC_OK) throw Error(s, std::string(gmmloc_status_string(s)) + ": " + gmmloc_last_error());
}

// ---- gmmloc::GMM ------------------------------------------------------------------------------------
class GMM {
 public:
  using Ptr = std::shared_ptr<GMM>;
  using ConstPtr = std::shared_ptr<const GMM>;

  // ≙ explicit GMM(const GaussianComponents&): mean K x 3, cov K x 9 row-major
  GMM(const double* mean, const double* cov, int K, int device = 0, double deg_eps = 1e-4) {
    check(gmmloc_map_create(mean, cov, K, deg_eps, device, &h_));
  }
  explicit GMM(gmmloc_map_t* adopt) : h_(adopt) {}
  GMM(const GMM&) = delete;
  void operator=(const GMM&) = delete;
  ~GMM() {
    if (qctx_) gmmloc_ctx_destroy(qctx_);
    gmmloc_map_destroy(h_);
  }

  // ≙ void GMM::queryPoint(const Eigen::Vector3d& pt, std::vector<int>& res)  (gaussian_mixture.h:169,
  // gaussian_mixture.cpp:545-576): res = { index of the component whose mean is Euclid-nearest to pt } — the reference
  // computes the Mahalanobis minimum over its 5-NN and then returns ret_index[0]; the quirk is kept.  Uses a private
  // context created on first use: like renderView (which mutates components2d_), call it from one thread at a time.
  void queryPoint(const Vector3d& pt, std::vector<int>& res) const {
    if (!qctx_) check(gmmloc_ctx_create(gmmloc_map_device(h_), &qctx_));
    int32_t nn = -1;
    check(gmmloc_map_query_point(qctx_, h_, pt.data(), 1, &nn, nullptr, nullptr));
    res.clear();
    if (nn >= 0) res.push_back((int)nn);
  }

  size_t countComponents() const { return (size_t)gmmloc_map_size(h_); }

  // ≙ getComponents(): constants built on the device, fetched once and cached
  const std::vector<GaussianComponent>& getComponents() const {
    if (comps_.empty()) {
      const size_t K = countComponents();
      std::vector<double> mean(3 * K), cov(9 * K), ci(9 * K), det(K), sc(3 * K), ax(9 * K), sq(9 * K);
      std::vector<uint8_t> dg(K), sl(K);
      check(gmmloc_map_get_components(h_, mean.data(), cov.data(), ci.data(), det.data(), sc.data(), ax.data(), sq.data(),
                                      dg.data(), sl.data()));
      comps_.resize(K);
      for (size_t k = 0; k < K; ++k) {
        GaussianComponent& c = comps_[k];
        for (int i = 0; i < 3; ++i) { c.mean_[i] = mean[3 * k + i]; c.scale_[i] = sc[3 * k + i]; }
        for (int i = 0; i < 9; ++i) {
          c.cov_[i] = cov[9 * k + i]; c.cov_inv_[i] = ci[9 * k + i]; c.sqrt_info_[i] = sq[9 * k + i]; c.axis_[i] = ax[9 * k + i];
        }
        c.det_ = det[k];
        c.is_degenerated = dg[k] != 0;
        c.is_salient = sl[k] != 0;
        c.id = (int)k;
      }
    }
    return comps_;
  }
  const GaussianComponent& getComponent3d(size_t idx) const { return getComponents().at(idx); }

  // ≙ the nbs_ construction in GMM::GMM (gaussian_mixture.cpp:61-78, gmmmap::neighbor_dist_thresh): call once,
  // before checkMapAssociation / associateMapElements; returns the number of neighbour pairs
  int64_t buildNeighbors(double neighbor_dist_thresh = 2.5) {
    int64_t n = 0;
    check(gmmloc_map_build_neighbors(h_, neighbor_dist_thresh, &n));
    return n;
  }

  gmmloc_map_t* handle() const { return h_; }

 private:
  gmmloc_map_t* h_ = nullptr;
  mutable gmmloc_ctx_t* qctx_ = nullptr;
  mutable std::vector<GaussianComponent> comps_;
};

// ---- gmmloc::GMMUtility -----------------------------------------------------------------------------
struct GMMUtility {
  // returns false (message in gmmloc_last_error()) where the reference returns false + LOG(ERROR)
  static bool loadGMMModel(const std::string& file_path, GMM::Ptr& model, int device = 0) {
    gmmloc_map_t* h = nullptr;
    if (gmmloc_map_load_gmm(file_path.c_str(), 1e-4, device, &h) != GMMLOC_OK) return false;
    model = std::make_shared<GMM>(h);
    return true;
  }
  static bool saveGMMModel(const std::string& file_path, GMM::Ptr& model, bool /*clear_file*/ = true) {
    return model && gmmloc_map_save_gmm(model->handle(), file_path.c_str()) == GMMLOC_OK;
  }
};

// result of the batched keyframe association (one entry per feature)
struct KeyFrameAssociation {
  std::vector<int32_t> assoc;            // associated component index, -1 = nullptr
  std::vector<Vector3d> pt3d;            // the (possibly optimised) world point; NaN when the feature has no depth
  std::vector<std::vector<int>> comps;   // KeyFrame::comps_: candidate parent components per feature
  int num_components2d = 0;              // size of the rendered view
};

// one covisibility window of Localization::jointOptimization (localization_opt.cpp:456-925) in flat form:
// the local + fixed keyframes, the local map points and one entry per (map point, keyframe) observation
struct LocalWindow {
  std::vector<SE3Quat> poses;          // local_kfs then fixed_kfs (in/out)
  std::vector<uint8_t> fixed;          // VertexSE3Expmap::setFixed
  int prior_idx = -1;                  // keyframe 0 when loc::ba_first_as_prior (EdgeSE3QuatPrior), else -1
  std::vector<Vector3d> points;        // MapPoint::getPosition() (in/out)
  std::vector<int32_t> assoc;          // MapPoint::asscociations_[0] as a component index, -1 = none
  std::vector<int32_t> obs_pt, obs_kf; // observation -> point / keyframe index (grouped by point)
  std::vector<Feature> obs;            // the observing keyframe's feature (uv, u_right, octave)
  // results
  std::vector<uint8_t> erase;          // vToErase (:857-881): observation is an outlier
  std::vector<uint8_t> assoc_cleared;  // :842-855: the point loses its association
  std::array<int32_t, 3> iters{0, 0, 0};
};

struct FrameResult {
  std::vector<double> pose;      // B x 7
  std::vector<float> pts;        // B x N x 3
  std::vector<int32_t> assign;   // B x N
  std::vector<uint8_t> outlier;  // B x N
  std::vector<int32_t> status;   // B
};

// ---- per-thread solver context (the tracking thread and the localisation thread each own one) -------
class Optimizer {
 public:
  explicit Optimizer(GMM::Ptr model, const gmmloc_camera_t& cam, int device = 0) : model_(std::move(model)), cam_(cam) {
    check(gmmloc_ctx_create(device, &ctx_));
    gmmloc_default_params(&params_);
  }
  Optimizer(const Optimizer&) = delete;
  void operator=(const Optimizer&) = delete;
  ~Optimizer() { gmmloc_ctx_destroy(ctx_); }

  gmmloc_params_t& params() { return params_; }
  gmmloc_ctx_t* handle() const { return ctx_; }

  // brute-force association of N world points; -1 where gate > 0 and best d2 > gate (gmmloc_opt.cpp:230)
  std::vector<int32_t> associate(const float* pts_w, int N, double gate = 0.0, std::vector<float>* d2 = nullptr) {
    std::vector<int32_t> a((size_t)N);
    if (d2) d2->resize((size_t)N);
    check(gmmloc_associate(ctx_, model_->handle(), pts_w, 1, N, gate, a.data(), d2 ? d2->data() : nullptr));
    return a;
  }

  // batched GMM::queryPoint on this context: nearest[i] = what the reference returns in res[0] for pts[i]
  std::vector<int32_t> queryPoints(const std::vector<Vector3d>& pts, std::vector<std::array<int32_t, 5>>* knn = nullptr,
                                   std::vector<double>* min_mdist2 = nullptr) {
    const int M = (int)pts.size();
    std::vector<int32_t> nn((size_t)M);
    if (knn) knn->resize((size_t)M);
    if (min_mdist2) min_mdist2->resize((size_t)M);
    check(gmmloc_map_query_point(ctx_, model_->handle(), M ? pts[0].data() : nullptr, M, nn.data(),
                                 knn && M ? (*knn)[0].data() : nullptr, min_mdist2 && M ? min_mdist2->data() : nullptr));
    return nn;
  }

  // ≙ static StrOptStat GMMLoc::optimizePoint(pt3d, kp, Tcw, comp, proj_z)   (one problem)
  StrOptStat optimizePoint(const Vector3d& pt3d, const Feature& kp, const SE3Quat& Tcw, int comp, double proj_z2, int iters = 5) {
    const float obs[3] = {(float)kp.u, (float)kp.v, (float)kp.u_right};
    const uint8_t oct = (uint8_t)kp.octave;
    const int32_t c = comp;
    StrOptStat r;
    uint8_t res = 0;
    check(gmmloc_points_solve(ctx_, model_->handle(), &cam_, &params_, pt3d.data(), obs, &oct, Tcw.data(), &c, &proj_z2, 1, iters,
                              r.pt_est.data(), &r.chi2_proj, &r.chi2_str, &res));
    r.res = res != 0;
    return r;
  }
  // batched form: every (feature, candidate component) problem of a keyframe in one launch
  // (the loop of GMMLoc::checkMapAssociation, gmmloc_opt.cpp:179-197)
  std::vector<StrOptStat> optimizePoints(const std::vector<Vector3d>& pt3d, const std::vector<Feature>& kp,
                                         const std::vector<SE3Quat>& Tcw, const std::vector<int32_t>& comp,
                                         const std::vector<double>& proj_z2, int iters = 5) {
    const size_t M = pt3d.size();
    if (kp.size() != M || Tcw.size() != M || comp.size() != M || proj_z2.size() != M) throw Error(GMMLOC_EINVAL, "size mismatch");
    std::vector<float> obs(3 * M);
    std::vector<uint8_t> oct(M), res(M);
    std::vector<double> x0(3 * M), pose(7 * M), est(3 * M), c2p(M), c2s(M);
    for (size_t i = 0; i < M; ++i) {
      obs[3 * i] = (float)kp[i].u; obs[3 * i + 1] = (float)kp[i].v; obs[3 * i + 2] = (float)kp[i].u_right;
      oct[i] = (uint8_t)kp[i].octave;
      for (int j = 0; j < 3; ++j) x0[3 * i + j] = pt3d[i][j];
      for (int j = 0; j < 7; ++j) pose[7 * i + j] = Tcw[i].v[j];
    }
    check(gmmloc_points_solve(ctx_, model_->handle(), &cam_, &params_, x0.data(), obs.data(), oct.data(), pose.data(), comp.data(),
                              proj_z2.data(), (int)M, iters, est.data(), c2p.data(), c2s.data(), res.data()));
    std::vector<StrOptStat> out(M);
    for (size_t i = 0; i < M; ++i) {
      out[i].res = res[i] != 0;
      out[i].chi2_proj = c2p[i];
      out[i].chi2_str = c2s[i];
      out[i].pt_est = {est[3 * i], est[3 * i + 1], est[3 * i + 2]};
    }
    return out;
  }

  // association + structure-constrained GN for B independent frames of N landmarks
  // (single-free-pose Localization::jointOptimization; params().factor_mode selects P / S / J)
  FrameResult optimizeFrames(const double* pose0, const float* pts_w, const float* obs, const uint8_t* octave, int B, int N) {
    FrameResult r;
    r.pose.resize(7 * (size_t)B);
    r.pts.resize(3 * (size_t)B * N);
    r.assign.resize((size_t)B * N);
    r.outlier.resize((size_t)B * N);
    r.status.resize((size_t)B);
    check(gmmloc_frames_solve(ctx_, model_->handle(), &cam_, &params_, pose0, pts_w, obs, octave, B, N, r.pose.data(), r.pts.data(),
                              r.assign.data(), r.outlier.data(), r.status.data()));
    return r;
  }

  // ≙ int Tracking::optimizeCurrentPose() with the reference's exact schedule (4-round Levenberg-
  // Marquardt, Huber, inlier reclassification; tracking_opt.cpp:21-217): updates Tcw, fills is_outlier
  // (Frame::is_outlier_), returns the inlier count (0 when fewer than 3 correspondences).
  // valid[i] != 0 marks features that have a map point (nullptr: all).
  int optimizeCurrentPose(SE3Quat& Tcw, const float* pts_w, const float* obs, const uint8_t* octave, int N,
                          std::vector<uint8_t>* is_outlier = nullptr, const uint8_t* valid = nullptr) {
    std::vector<uint8_t> outl((size_t)N);
    int32_t inliers = 0;
    SE3Quat out;
    check(gmmloc_poses_optimize(ctx_, &cam_, Tcw.data(), pts_w, obs, octave, valid, 1, N, out.data(), outl.data(), &inliers));
    Tcw = out;
    if (is_outlier) *is_outlier = outl;
    return inliers;
  }

  // ---- keyframe association (main thread of the reference, gmmloc_opt.cpp:115-258) ----------------------
  // ≙ void GMM::renderView(const Quaterniond& rot_c_w, const Vector3d& t_c_w): the view stays on this context
  // (the reference mutates GMM::components2d_) and is mirrored on the host for getComponents2d()
  void renderView(const SE3Quat& Tcw) {
    const int cap = (int)model_->countComponents();
    int32_t count = 0, status = 0;
    std::vector<int32_t> ids((size_t)cap);
    std::vector<double> mean(2 * (size_t)cap), cov(4 * (size_t)cap), depth((size_t)cap);
    check(gmmloc_render_views(ctx_, model_->handle(), &cam_, Tcw.data(), 1, cap, &count, &status, ids.data(), mean.data(),
                              cov.data(), depth.data()));
    components2d_.resize((size_t)count);
    for (int j = 0; j < count; ++j) {
      GaussianComponent2d& g = components2d_[j];
      g.id_ = (uint32_t)ids[j];
      g.proj_d_ = depth[j];
      g.mean_ = {mean[2 * j], mean[2 * j + 1]};
      g.cov_ = {cov[4 * j], cov[4 * j + 1], cov[4 * j + 2], cov[4 * j + 3]};
    }
  }
  const GaussianComponents2d& getComponents2d() const { return components2d_; }

  // ≙ void GMM::searchCorrespondence(const std::vector<Feature>&, std::vector<GaussianComponents2d>&, int num):
  // comps[i] = parent component indices of feature i's candidates (KeyFrame::comps_), on the last rendered view
  void searchCorrespondence(const std::vector<Feature>& kpts, std::vector<std::vector<int>>& comps, int num = 5) {
    const size_t N = kpts.size();
    std::vector<float> obs(3 * N);
    for (size_t i = 0; i < N; ++i) { obs[3 * i] = (float)kpts[i].u; obs[3 * i + 1] = (float)kpts[i].v; obs[3 * i + 2] = (float)kpts[i].u_right; }
    std::vector<int32_t> cand(N * (size_t)num);
    check(gmmloc_search_correspondence(ctx_, obs.data(), 1, (int)N, num, cand.data(), nullptr));
    comps.assign(N, {});
    for (size_t i = 0; i < N; ++i)
      for (int r = 0; r < num && cand[i * num + r] >= 0; ++r) comps[i].push_back(cand[i * num + r]);
  }

  // ≙ GaussianComponent::ConstPtr GMMLoc::checkMapAssociation(Vector3d& pt3d, KeyFrame*, size_t idx): returns the
  // associated component index or -1 (nullptr); pt3d is updated where the reference mutates it
  int checkMapAssociation(Vector3d& pt3d, const Feature& kp, const SE3Quat& Tcw, const std::vector<int>& comps) {
    if (comps.empty()) return -1;
    const float obs[3] = {(float)kp.u, (float)kp.v, (float)kp.u_right};
    const uint8_t oct = (uint8_t)kp.octave;
    std::vector<int32_t> cand(comps.begin(), comps.end());
    int32_t assoc = -1;
    check(gmmloc_check_map_association(ctx_, model_->handle(), &cam_, &params_, Tcw.data(), obs, &oct, cand.data(), 1, 1,
                                       (int)cand.size(), pt3d.data(), &assoc));
    return assoc;
  }

  // ≙ GMMLoc::associateMapElements(kf) followed by the unproject3 + checkMapAssociation of every feature with
  // depth > 0 in GMMLoc::createMapPointsFromStereo (gmmloc_opt.cpp:36-154), one keyframe, chained on the device.
  // The caller replays createMapPointsFromStereo's bookkeeping loop (depth order, th_depth early exit) on the result.
  KeyFrameAssociation associateMapElements(const SE3Quat& Tcw, const std::vector<Feature>& kpts, int num = 5) {
    const size_t N = kpts.size();
    std::vector<float> obs(3 * N), depth(N);
    std::vector<uint8_t> oct(N);
    for (size_t i = 0; i < N; ++i) {
      obs[3 * i] = (float)kpts[i].u; obs[3 * i + 1] = (float)kpts[i].v; obs[3 * i + 2] = (float)kpts[i].u_right;
      depth[i] = kpts[i].depth;
      oct[i] = (uint8_t)kpts[i].octave;
    }
    KeyFrameAssociation r;
    r.assoc.resize(N);
    r.pt3d.resize(N);
    std::vector<int32_t> cand(N * (size_t)num);
    int32_t nview = 0;
    check(gmmloc_keyframes_associate(ctx_, model_->handle(), &cam_, &params_, Tcw.data(), obs.data(), depth.data(), oct.data(), 1,
                                     (int)N, num, (int)model_->countComponents(), r.assoc.data(),
                                     N ? r.pt3d[0].data() : nullptr, cand.data(), &nview, nullptr));
    r.num_components2d = nview;
    r.comps.assign(N, {});
    for (size_t i = 0; i < N; ++i)
      for (int c = 0; c < num && cand[i * num + c] >= 0; ++c) r.comps[i].push_back(cand[i * num + c]);
    return r;
  }

  // ≙ void Localization::jointOptimization(KeyFrame*, bool* stop, Map*): the caller flattens the window (the
  // reference's graph construction, :460-766) and applies the results (:842-925); the 3-stage LM runs on the GPU.
  // `stop` is the reference's pbStopFlag (may be raised by another thread while this call runs; see gmmloc_local_ba).
  void jointOptimization(LocalWindow& w, const volatile int32_t* stop = nullptr) {
    const int32_t P = (int32_t)w.poses.size(), L = (int32_t)w.points.size(), E = (int32_t)w.obs_pt.size();
    if ((int32_t)w.fixed.size() != P || (int32_t)w.assoc.size() != L || (int32_t)w.obs_kf.size() != E || (int32_t)w.obs.size() != E)
      throw Error(GMMLOC_EINVAL, "LocalWindow: size mismatch");
    std::vector<double> pose(7 * (size_t)P), pts(3 * (size_t)L);
    std::vector<float> obs(3 * (size_t)E);
    std::vector<uint8_t> oct((size_t)E);
    for (int32_t i = 0; i < P; ++i)
      for (int j = 0; j < 7; ++j) pose[7 * (size_t)i + j] = w.poses[i].v[j];
    for (int32_t l = 0; l < L; ++l)
      for (int j = 0; j < 3; ++j) pts[3 * (size_t)l + j] = w.points[l][j];
    for (int32_t e = 0; e < E; ++e) {
      obs[3 * (size_t)e] = (float)w.obs[e].u; obs[3 * (size_t)e + 1] = (float)w.obs[e].v; obs[3 * (size_t)e + 2] = (float)w.obs[e].u_right;
      oct[e] = (uint8_t)w.obs[e].octave;
    }
    const int32_t pp[2] = {0, P}, lp[2] = {0, L}, ep[2] = {0, E};
    w.erase.assign((size_t)E, 0);
    w.assoc_cleared.assign((size_t)L, 0);
    int32_t status = 0;
    check(gmmloc_local_ba(ctx_, model_->handle(), &cam_, &params_, 1, pp, pose.data(), w.fixed.data(), &w.prior_idx, lp, pts.data(),
                          w.assoc.data(), ep, w.obs_pt.data(), w.obs_kf.data(), obs.data(), oct.data(), w.erase.data(),
                          w.assoc_cleared.data(), w.iters.data(), &status, stop));
    for (int32_t i = 0; i < P; ++i)
      for (int j = 0; j < 7; ++j) w.poses[i].v[j] = pose[7 * (size_t)i + j];
    for (int32_t l = 0; l < L; ++l) w.points[l] = {pts[3 * (size_t)l], pts[3 * (size_t)l + 1], pts[3 * (size_t)l + 2]};
  }

  // ≙ GaussianComponent* Localization::optimizeTriangulationVec(Vector3d& x3d, KeyFrame* pKF1, idx1, idx2)
  // (localization_opt.cpp:27-204): every DEGENERATE candidate of the two views is solved (20 GN iterations) in one
  // launch; the accepted candidate with the smallest err1 + err2 wins (strict '<'; candidates in ascending index,
  // the reference iterates an unordered_set).  Returns its index or -1 and updates x3d.
  int optimizeTriangulationVec(Vector3d& x3d, const Feature& kp1, const Feature& kp2, const SE3Quat& Tcw1, const SE3Quat& Tcw2,
                               const std::vector<int>& candidates, int iters = 20) {
    const auto& comps = model_->getComponents();
    std::vector<int32_t> cand;
    for (int c : candidates)
      if (c >= 0 && (size_t)c < comps.size() && comps[(size_t)c].is_degenerated) cand.push_back(c);
    std::sort(cand.begin(), cand.end());
    cand.erase(std::unique(cand.begin(), cand.end()), cand.end());
    const size_t M = cand.size();
    if (M == 0) return -1;
    std::vector<double> x0(3 * M), p1(7 * M), p2(7 * M), est(3 * M), e1(M), e2(M), cs(M);
    std::vector<float> o1(3 * M), o2(3 * M);
    std::vector<uint8_t> oc(M, (uint8_t)kp1.octave), res(M);
    for (size_t i = 0; i < M; ++i) {
      for (int j = 0; j < 3; ++j) x0[3 * i + j] = x3d[j];
      for (int j = 0; j < 7; ++j) { p1[7 * i + j] = Tcw1.v[j]; p2[7 * i + j] = Tcw2.v[j]; }
      o1[3 * i] = (float)kp1.u; o1[3 * i + 1] = (float)kp1.v; o1[3 * i + 2] = (float)kp1.u_right;
      o2[3 * i] = (float)kp2.u; o2[3 * i + 1] = (float)kp2.v; o2[3 * i + 2] = (float)kp2.u_right;
    }
    check(gmmloc_triangulation_solve(ctx_, model_->handle(), &cam_, &params_, x0.data(), o1.data(), o2.data(), oc.data(), p1.data(),
                                     p2.data(), cand.data(), (int)M, iters, est.data(), e1.data(), e2.data(), cs.data(), res.data()));
    int best = -1;
    double best_v = 1.7976931348623157e308;
    for (size_t i = 0; i < M; ++i)
      if (res[i] && e1[i] + e2[i] < best_v) { best = (int)i; best_v = e1[i] + e2[i]; }
    if (best < 0) return -1;
    x3d = {est[3 * best], est[3 * best + 1], est[3 * best + 2]};
    return cand[(size_t)best];
  }

 private:
  GMM::Ptr model_;
  gmmloc_camera_t cam_;
  gmmloc_params_t params_;
  gmmloc_ctx_t* ctx_ = nullptr;
  GaussianComponents2d components2d_;
};

}  // namespace gmmloc_b200
