// Exercises include/gmmloc_b200.hpp the way a C++ caller of the reference would (known-answer checks).
// Built and run by tests/test_cpp_shim.py on the GPU box.
#include <cmath>
#include <cstdio>
#include <cstdlib>

#include "gmmloc_b200.hpp"

using namespace gmmloc_b200;

#define REQUIRE(c)                                                  \
  do {                                                              \
    if (!(c)) {                                                     \
      std::printf("FAILED %s:%d: %s\n", __FILE__, __LINE__, #c);    \
      return 1;                                                     \
    }                                                               \
  } while (0)

int main(int argc, char** argv) {
  // a room corner seen from the origin (camera looks along +z): 16 planar components on the wall z = 3
  // (normal z), 8 on the wall x = 2 (normal x), 8 on the floor y = 1.5 (normal y); sigma_n = 1 mm,
  // in-plane 10 cm.  Three plane orientations make the 6-DoF reduced system full rank.
  const int K = 32;
  std::vector<double> mean(3 * K), cov(9 * K, 0.0);
  for (int k = 0; k < K; ++k) {
    int nrm;
    if (k < 16) {
      mean[3 * k] = (k % 4) * 0.8 - 1.2; mean[3 * k + 1] = (k / 4) * 0.6 - 0.9; mean[3 * k + 2] = 3.0; nrm = 2;
    } else if (k < 24) {
      mean[3 * k] = 2.0; mean[3 * k + 1] = ((k - 16) % 4) * 0.5 - 0.75; mean[3 * k + 2] = 2.6 + ((k - 16) / 4) * 0.8; nrm = 0;
    } else {
      mean[3 * k] = ((k - 24) % 4) * 0.7 - 1.05; mean[3 * k + 1] = 1.5; mean[3 * k + 2] = 3.0 + ((k - 24) / 4) * 0.8; nrm = 1;
    }
    for (int a = 0; a < 3; ++a) cov[9 * k + 4 * a] = (a == nrm) ? 1e-6 : 1e-2;
  }
  auto gmm = std::make_shared<GMM>(mean.data(), cov.data(), K);
  REQUIRE(gmm->countComponents() == (size_t)K);
  const GaussianComponent& c5 = gmm->getComponent3d(5);
  REQUIRE(c5.is_degenerated && !c5.is_salient);
  REQUIRE(std::fabs(c5.cov_inv()[8] - 1e6) < 1e-3 && std::fabs(c5.scale_[0] - 1e-6) < 1e-12);
  REQUIRE(std::fabs(std::fabs(c5.axis_[6]) - 1.0) < 1e-12);  // axis_.col(0) = +-z
  REQUIRE(std::fabs(std::fabs(gmm->getComponent3d(20).axis_[0]) - 1.0) < 1e-12);  // +-x on the side wall

  gmmloc_camera_t cam{435.2046959714599, 435.2046959714599, 367.4517211914062, 252.2008514404297, 47.90639384423901, 752, 480};
  Optimizer opt(gmm, cam);

  // association: every mean belongs to its own component with d2 = 0; a point 2 mm off a plane: d2 = 4
  std::vector<float> pts(3 * K);
  for (int i = 0; i < 3 * K; ++i) pts[i] = (float)mean[i];
  std::vector<float> d2;
  auto a = opt.associate(pts.data(), K, 0.0, &d2);
  for (int k = 0; k < K; ++k) REQUIRE(a[k] == k && d2[k] < 1e-6f);  // the fp32 landmark is the mean rounded to float
  float off[3] = {(float)mean[15], (float)mean[16], (float)(mean[17] + 0.002)};
  auto a1 = opt.associate(off, 1, 0.0, &d2);
  REQUIRE(a1[0] == 5 && std::fabs(d2[0] - 4.0f) < 1e-2f);
  auto a2 = opt.associate(off, 1, 3.9);  // gate below d2 -> no association (nullptr in the reference)
  REQUIRE(a2[0] == -1);

  // optimizePoint: identity pose, point 5 cm in front of plane 5, observation consistent with the plane:
  // the structure edge pulls the estimate onto the plane, both chi2 end up tiny
  SE3Quat Tcw;
  const double X[3] = {mean[15] + 0.02, mean[16] - 0.03, mean[17]};
  Feature kp;
  kp.u = cam.fx * X[0] / X[2] + cam.cx;
  kp.v = cam.fy * X[1] / X[2] + cam.cy;
  kp.u_right = kp.u - cam.bf / X[2];
  kp.octave = 0;
  StrOptStat r = opt.optimizePoint({X[0], X[1], X[2] + 0.05}, kp, Tcw, 5, 1.0);
  REQUIRE(r.res);
  REQUIRE(std::fabs(r.pt_est[2] - X[2]) < 1e-3 && r.chi2_proj < 1.0 && r.chi2_str < 0.1);
  auto rb = opt.optimizePoints({{X[0], X[1], X[2] + 0.05}, {X[0], X[1], X[2] + 0.05}}, {kp, kp}, {Tcw, Tcw}, {5, -1}, {1.0, 1.0});
  REQUIRE(rb[0].res && !rb[1].res && rb[0].pt_est == r.pt_est);

  // frame solve: perfect observations of points on the planes, pose perturbed by 1 cm: GN recovers identity
  const int N = 96;
  std::vector<float> P(3 * N), O(3 * N);
  std::vector<uint8_t> oct(N, 0);
  for (int i = 0; i < N; ++i) {
    const int k = i % K;
    const GaussianComponent& c = gmm->getComponent3d(k);
    // offsets inside the plane of component k (its two large-variance axes)
    const double o1 = 0.05 * ((i * 7) % 5 - 2), o2 = 0.04 * ((i * 3) % 5 - 2);
    double x = c.mean()[0], y = c.mean()[1], z = c.mean()[2];
    if (k < 16) { x += o1; y += o2; } else if (k < 24) { y += o1; z += o2; } else { x += o1; z += o2; }
    P[3 * i] = (float)x; P[3 * i + 1] = (float)y; P[3 * i + 2] = (float)z;
    const double u = cam.fx * x / z + cam.cx, v = cam.fy * y / z + cam.cy;
    O[3 * i] = (float)u; O[3 * i + 1] = (float)v; O[3 * i + 2] = (float)(u - cam.bf / z);
  }
  SE3Quat T0;
  T0.v[4] = 0.01; T0.v[5] = -0.01;
  FrameResult fr = opt.optimizeFrames(T0.data(), P.data(), O.data(), oct.data(), 1, N);
  REQUIRE(fr.status[0] == 0);
  for (int i = 0; i < N; ++i) REQUIRE(fr.assign[i] == i % K && fr.outlier[i] == 0);
  REQUIRE(std::fabs(fr.pose[4]) < 2e-3 && std::fabs(fr.pose[5]) < 2e-3 && std::fabs(fr.pose[3] - 1.0) < 1e-5);

  // pose-only (Tracking::optimizeCurrentPose): all inliers, pose pulled back to identity
  SE3Quat T1 = T0;
  std::vector<uint8_t> outl;
  const int inl = opt.optimizeCurrentPose(T1, P.data(), O.data(), oct.data(), N, &outl);
  REQUIRE(inl == N && std::fabs(T1.v[4]) < 1e-4 && std::fabs(T1.v[5]) < 1e-4);
  REQUIRE(opt.optimizeCurrentPose(T1, P.data(), O.data(), oct.data(), 2) == 0);

  // keyframe association: render the room from the origin, search candidates for a keypoint on wall component 5,
  // associate it (the structure edge pulls the unprojected point onto the wall z = 3)
  REQUIRE(gmm->buildNeighbors(2.5) >= 0);
  opt.renderView(Tcw);
  const GaussianComponents2d& v2d = opt.getComponents2d();
  REQUIRE(!v2d.empty() && v2d.size() <= (size_t)K);
  for (size_t j = 1; j < v2d.size(); ++j) REQUIRE(v2d[j - 1].proj_d_ >= v2d[j].proj_d_);  // far -> near
  bool saw5 = false;
  for (const auto& g2 : v2d) {
    REQUIRE(g2.id_ < (uint32_t)K && g2.mean()[0] >= 0 && g2.mean()[0] < 752 && g2.mean()[1] >= 0 && g2.mean()[1] < 480);
    if (g2.id_ == 5) {
      saw5 = true;
      REQUIRE(std::fabs(g2.proj_d_ - 3.0) < 1e-12);
      REQUIRE(std::fabs(g2.mean()[0] - (cam.fx * mean[15] / 3.0 + cam.cx)) < 1e-9);
      const double jz = cam.fx * mean[15] / 9.0;  // J S J^T of a fronto-parallel patch: (fx/z)^2 s_x + (fx x/z^2)^2 s_z
      REQUIRE(std::fabs(g2.cov()[0] - (cam.fx * cam.fx * 1e-2 / 9.0 + jz * jz * 1e-6)) < 1e-6);
    }
  }
  REQUIRE(saw5);
  Feature kf_kp = kp;
  kf_kp.depth = (float)(X[2] + 0.04);  // stereo depth 4 cm too deep
  kf_kp.u_right = kf_kp.u - cam.bf / (X[2] + 0.04);
  std::vector<std::vector<int>> comps2;
  opt.searchCorrespondence({kf_kp}, comps2);
  REQUIRE(comps2.size() == 1 && !comps2[0].empty() && comps2[0][0] == 5);  // nearest 2-D mean first
  Vector3d p3{X[0] * (X[2] + 0.04) / X[2], X[1] * (X[2] + 0.04) / X[2], X[2] + 0.04};  // unproject3 at the identity pose
  const int as = opt.checkMapAssociation(p3, kf_kp, Tcw, comps2[0]);
  REQUIRE(as == 5 && std::fabs(p3[2] - 3.0) < 5e-3);  // pulled from z = 3.04 onto the wall
  REQUIRE(opt.checkMapAssociation(p3, kf_kp, Tcw, {}) == -1);
  Feature nodepth = kf_kp;
  nodepth.depth = -1.0f;
  KeyFrameAssociation ka = opt.associateMapElements(Tcw, {kf_kp, nodepth});
  REQUIRE(ka.num_components2d == (int)v2d.size() && ka.assoc[0] == 5 && ka.assoc[1] == -1);
  REQUIRE(std::fabs(ka.pt3d[0][2] - 3.0) < 5e-3 && std::isnan(ka.pt3d[1][0]) && ka.comps[0] == comps2[0]);

  // two-view triangulation against candidate planes: the right wall component wins, the point lands on z = 3
  SE3Quat Tcw2;
  Tcw2.v[4] = -0.2;  // second keyframe 20 cm to the right
  Feature kp2;
  kp2.u = cam.fx * (X[0] - 0.2) / X[2] + cam.cx;
  kp2.v = kp.v;
  kp2.u_right = -1;  // monocular in view 2
  Vector3d x3{X[0], X[1], X[2] + 0.03};
  const int tri = opt.optimizeTriangulationVec(x3, kp, kp2, Tcw, Tcw2, {5, 20, 5, -1});
  REQUIRE(tri == 5 && std::fabs(x3[2] - 3.0) < 2e-3);

  // local BA over two keyframes (second one fixed): perfect observations, first pose perturbed by 1 cm -> recovered
  {
    LocalWindow w;
    w.poses = {T0, Tcw2};
    w.fixed = {0, 1};
    for (int i = 0; i < N; ++i) {
      const double x = P[3 * i], y = P[3 * i + 1], z = P[3 * i + 2];
      w.points.push_back({x, y, z});
      w.assoc.push_back(i % K);
      Feature f0, f1;
      f0.u = cam.fx * x / z + cam.cx; f0.v = cam.fy * y / z + cam.cy; f0.u_right = f0.u - cam.bf / z;
      f1.u = cam.fx * (x - 0.2) / z + cam.cx; f1.v = f0.v; f1.u_right = f1.u - cam.bf / z;
      w.obs_pt.push_back(i); w.obs_kf.push_back(0); w.obs.push_back(f0);
      w.obs_pt.push_back(i); w.obs_kf.push_back(1); w.obs.push_back(f1);
    }
    {  // pbStopFlag raised before the call (localization_opt.cpp:765-767): the window comes back untouched
      LocalWindow ws = w;
      const volatile int32_t stop = 1;
      opt.jointOptimization(ws, &stop);
      REQUIRE(ws.iters[0] == 0 && ws.iters[1] == 0 && ws.iters[2] == 0);
      REQUIRE(ws.poses[0].v == w.poses[0].v && ws.points[3] == w.points[3]);
      const volatile int32_t go = 0;  // a flag that stays down changes nothing
      LocalWindow wa = w, wb = w;
      opt.jointOptimization(wa, &go);
      opt.jointOptimization(wb);
      REQUIRE(wa.poses[0].v == wb.poses[0].v && wa.iters == wb.iters);
    }
    opt.jointOptimization(w);
    REQUIRE(w.iters[0] == 5 && w.iters[1] == 5 && w.iters[2] >= 1);
    REQUIRE(std::fabs(w.poses[0].v[4]) < 1e-4 && std::fabs(w.poses[0].v[5]) < 1e-4 && w.poses[1].v[4] == -0.2);
    int erased = 0;
    for (uint8_t e : w.erase) erased += e;
    REQUIRE(erased == 0);
  }

  // GMM::queryPoint: the Euclid-nearest mean (the reference's ret_index[0] quirk), also where another component is
  // nearer in the Mahalanobis sense: 1 cm in front of the wall z = 3 but closer to the MEAN of a floor component
  {
    std::vector<int> res;
    gmm->queryPoint({mean[15], mean[16], mean[17] + 0.01}, res);
    REQUIRE(res.size() == 1 && res[0] == 5);
    std::vector<std::array<int32_t, 5>> knn;
    std::vector<double> md;
    auto nn = opt.queryPoints({{mean[15], mean[16], mean[17] + 0.01}, {-1.05 + 0.05, 1.5 - 0.2, 3.0 + 0.05}}, &knn, &md);
    REQUIRE(nn[0] == 5 && knn[0][0] == 5 && std::fabs(md[0] - 100.0) < 1e-6);  // 1 cm / 1 mm sigma
    REQUIRE(nn[1] == 24 && knn[1][0] == 24);  // 20 cm above the floor component 24 (d2 = 4e4 there) — still returned
  }

  // gmmloc_render_views through the raw ABI with cap > K and B = 2: the host rows keep the caller's stride `cap`
  {
    const int cap = K + 7, B = 2;
    std::vector<double> poses = {0, 0, 0, 1, 0, 0, 0, 0, 0, 0, 1, 0.1, 0, 0};
    std::vector<int32_t> cnt(B), ids((size_t)B * cap, -2);
    std::vector<double> depth((size_t)B * cap, -2.0);
    REQUIRE(gmmloc_render_views(opt.handle(), gmm->handle(), &cam, poses.data(), B, cap, cnt.data(), nullptr, ids.data(), nullptr,
                                nullptr, depth.data()) == GMMLOC_OK);
    for (int b = 0; b < B; ++b) {
      REQUIRE(cnt[b] > 0 && cnt[b] <= K);
      for (int j = 0; j < cap; ++j) {
        if (j < cnt[b]) REQUIRE(ids[(size_t)b * cap + j] >= 0 && ids[(size_t)b * cap + j] < K && depth[(size_t)b * cap + j] > 0.0);
        if (j >= K) REQUIRE(ids[(size_t)b * cap + j] == -1 && depth[(size_t)b * cap + j] == 0.0);
      }
      for (int j = 1; j < cnt[b]; ++j) REQUIRE(depth[(size_t)b * cap + j] <= depth[(size_t)b * cap + j - 1]);  // depth descending
    }
  }

  // .gmm round trip through GMMUtility
  if (argc > 1) {
    GMM::Ptr g2;
    REQUIRE(GMMUtility::saveGMMModel(argv[1], gmm));
    REQUIRE(GMMUtility::loadGMMModel(argv[1], g2));
    REQUIRE(g2->countComponents() == (size_t)K && g2->getComponent3d(7).mean() == gmm->getComponent3d(7).mean());
    REQUIRE(!GMMUtility::loadGMMModel("/nonexistent.gmm", g2));
  }
  std::printf("OK\n");
  return 0;
}
