// ref_nanoflann.cpp — kNN harness over the REFERENCE's own vendored nanoflann (test infrastructure).
//
// The one third-party piece of the hot path that is vendored in the reference tree is nanoflann 1.3.2
// (gmmloc/include/gmmloc/utils/nanoflann.hpp).  This file is compiled AGAINST that header where it lies under
// /root/reference (never copied; recipe: oracle/Makefile target `ref`, output oracle/_ref/libref_nanoflann.so) and
// pins what the oracle restates with a brute-force scan:
//   * GMM::searchCorrespondence's  KDTreePoints2d(2, pts, KDTreeSingleIndexAdaptorParams(5)) + knnSearch(q, num)
//     (gmmloc/src/gmm/gaussian_mixture.cpp:484-534; dataset adaptor FLANNPoints2d, gaussian_mixture.h:55-91)
//   * GMM::queryPoint's            KDTreePoints3d(3, pts, ...Params(5)) + knnSearch(q, 5)
//     (gaussian_mixture.cpp:537-576; adaptor FLANNPoints3d, gaussian_mixture.h:15-53)
// The adaptors below restate the reference's (same distance expression d0*d0 + d1*d1 (+ d2*d2), same accessors, no
// bounding box) because the originals pull in Eigen through gaussian.h.
#include <cstddef>
#include <cstdint>
#include <vector>

#include "gmmloc/utils/nanoflann.hpp"

namespace {

template <int DIM>
struct Points {
  const double* p;
  size_t n;
  inline size_t kdtree_get_point_count() const { return n; }
  inline double kdtree_distance(const double* q, const size_t idx, size_t) const {
    const double d0 = q[0] - p[idx * DIM], d1 = q[1] - p[idx * DIM + 1];
    if (DIM == 2) return d0 * d0 + d1 * d1;
    const double d2 = q[2] - p[idx * DIM + 2];
    return d0 * d0 + d1 * d1 + d2 * d2;
  }
  inline double kdtree_get_pt(const size_t idx, const size_t dim) const { return p[idx * DIM + dim]; }
  template <class BBOX>
  bool kdtree_get_bbox(BBOX&) const { return false; }
};

template <int DIM>
void knn(const double* pts, int64_t n, const double* q, int64_t nq, int k, int64_t* idx, double* dist, int32_t* count) {
  using P = Points<DIM>;
  using Tree = nanoflann::KDTreeSingleIndexAdaptor<nanoflann::L2_Simple_Adaptor<double, P>, P, DIM>;
  P cloud{pts, (size_t)n};
  Tree tree(DIM, cloud, nanoflann::KDTreeSingleIndexAdaptorParams(5));
  tree.buildIndex();
  std::vector<size_t> ret((size_t)k);
  std::vector<double> d2((size_t)k);
  for (int64_t i = 0; i < nq; ++i) {
    const size_t m = n > 0 ? tree.knnSearch(q + i * DIM, (size_t)k, ret.data(), d2.data()) : 0;
    count[i] = (int32_t)m;
    for (int j = 0; j < k; ++j) {
      idx[i * k + j] = j < (int)m ? (int64_t)ret[j] : -1;
      dist[i * k + j] = j < (int)m ? d2[j] : -1.0;
    }
  }
}

}  // namespace

extern "C" {
// kNN of nq queries among n points: idx / dist nq x k (ascending squared distance, -1 padded), count nq
void ref_knn2d(const double* pts, int64_t n, const double* q, int64_t nq, int k, int64_t* idx, double* dist, int32_t* count) {
  knn<2>(pts, n, q, nq, k, idx, dist, count);
}
void ref_knn3d(const double* pts, int64_t n, const double* q, int64_t nq, int k, int64_t* idx, double* dist, int32_t* count) {
  knn<3>(pts, n, q, nq, k, idx, dist, count);
}
}
