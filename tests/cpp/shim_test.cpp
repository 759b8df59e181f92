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
  // a 4 x 4 x 2 lattice of planar components (normal = z, sigma_z = 1 mm, in-plane 10 cm)
  const int K = 32;
  std::vector<double> mean(3 * K), cov(9 * K, 0.0);
  for (int k = 0; k < K; ++k) {
    mean[3 * k] = (k % 4) * 1.0 - 1.5;
    mean[3 * k + 1] = ((k / 4) % 4) * 1.0 - 1.5;
    mean[3 * k + 2] = 3.0 + (k / 16) * 1.0;
    cov[9 * k] = 1e-2; cov[9 * k + 4] = 1e-2; cov[9 * k + 8] = 1e-6;
  }
  auto gmm = std::make_shared<GMM>(mean.data(), cov.data(), K);
  REQUIRE(gmm->countComponents() == (size_t)K);
  const GaussianComponent& c5 = gmm->getComponent3d(5);
  REQUIRE(c5.is_degenerated && !c5.is_salient);
  REQUIRE(std::fabs(c5.cov_inv()[8] - 1e6) < 1e-3 && std::fabs(c5.scale_[0] - 1e-6) < 1e-12);
  REQUIRE(std::fabs(std::fabs(c5.axis_[6]) - 1.0) < 1e-12);  // axis_.col(0) = +-z

  gmmloc_camera_t cam{435.2046959714599, 435.2046959714599, 367.4517211914062, 252.2008514404297, 47.90639384423901, 752, 480};
  Optimizer opt(gmm, cam);

  // association: every mean belongs to its own component with d2 = 0; a point 2 mm off a plane: d2 = 4
  std::vector<float> pts(3 * K);
  for (int i = 0; i < 3 * K; ++i) pts[i] = (float)mean[i];
  std::vector<float> d2;
  auto a = opt.associate(pts.data(), K, 0.0, &d2);
  for (int k = 0; k < K; ++k) REQUIRE(a[k] == k && d2[k] == 0.0f);
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
  const int N = 64;
  std::vector<float> P(3 * N), O(3 * N);
  std::vector<uint8_t> oct(N, 0);
  for (int i = 0; i < N; ++i) {
    const int k = i % 16;  // front layer
    const double x = mean[3 * k] + 0.05 * ((i * 7) % 5 - 2), y = mean[3 * k + 1] + 0.04 * ((i * 3) % 5 - 2), z = mean[3 * k + 2];
    P[3 * i] = (float)x; P[3 * i + 1] = (float)y; P[3 * i + 2] = (float)z;
    const double u = cam.fx * x / z + cam.cx, v = cam.fy * y / z + cam.cy;
    O[3 * i] = (float)u; O[3 * i + 1] = (float)v; O[3 * i + 2] = (float)(u - cam.bf / z);
  }
  SE3Quat T0;
  T0.v[4] = 0.01; T0.v[5] = -0.01;
  FrameResult fr = opt.optimizeFrames(T0.data(), P.data(), O.data(), oct.data(), 1, N);
  REQUIRE(fr.status[0] == 0);
  for (int i = 0; i < N; ++i) REQUIRE(fr.assign[i] == i % 16 && fr.outlier[i] == 0);
  REQUIRE(std::fabs(fr.pose[4]) < 2e-3 && std::fabs(fr.pose[5]) < 2e-3 && std::fabs(fr.pose[3] - 1.0) < 1e-5);

  // pose-only (Tracking::optimizeCurrentPose): all inliers, pose pulled back to identity
  SE3Quat T1 = T0;
  std::vector<uint8_t> outl;
  const int inl = opt.optimizeCurrentPose(T1, P.data(), O.data(), oct.data(), N, &outl);
  REQUIRE(inl == N && std::fabs(T1.v[4]) < 1e-4 && std::fabs(T1.v[5]) < 1e-4);
  REQUIRE(opt.optimizeCurrentPose(T1, P.data(), O.data(), oct.data(), 2) == 0);

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
