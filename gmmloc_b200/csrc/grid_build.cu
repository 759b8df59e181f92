// grid_build.cu — builds the association grid (GridDev in common.cuh) on the device at map creation.
//
// The reference narrows the association with a 2-D kd-tree over the projected components
// (GMM::searchCorrespondence, gmmloc/src/gmm/gaussian_mixture.cpp:484-534) and the Bhattacharyya neighbour
// lists (gmmloc_opt.cpp:203-217).  The brute-force argmin of this library gets the same kind of
// index, but an EXACT one: cell c of a uniform grid lists every component k whose ellipsoid
//     E_k(T) = { x : (x - mu_k)^T cov_inv_k (x - mu_k) <= T }
// can touch the (slightly inflated) cell box.  A component is left out of a cell only when a separating axis
// exists: for ANY direction a,  max_{x in E_k(T)} a.(x - mu_k) = sqrt(T a^T M_k a)  with M_k = sym(cov_inv_k)^-1,
// so "the projection of the cell box on a is farther from mu_k than that" proves that no point of the cell has
// d2_k <= T.  Six directions are tested (the world axes = the ellipsoid's bounding box, and the component's
// eigen-axes = its oriented bounding box; the eigen-axes only need to be approximate, the bound holds for any a).
// T is inflated by 1e-3 and the cell box by 1e-3 h + 1e-4 m, which covers the fp32 rounding of the landmark's cell
// index on the device and the fp64 rounding of cov_inv / d2 in the reference (relative 1e-16 cond(cov)).
//
// Passes (per level): count (one warp per component walks the cells of its bounding box) -> host prefix sum ->
// fill -> per-cell sort (ascending component id: deterministic lists).  Components whose bounding box covers more
// than kBigCells cells of level 0 are not binned; they go to the `big` list every landmark scores.
#include <algorithm>
#include <cmath>
#include <cstdlib>
#include <vector>

#include "common.cuh"

namespace gmmloc {

namespace {

constexpr int kBigCells = 1 << 15;
constexpr int kMaxBig = 256;
constexpr long long kMaxCells = 4LL << 20;
constexpr long long kMaxEntries = 96LL << 20;

struct LevelBuild {
  double lo[3];
  double h;
  int nx, ny, nz;
  double T;  // inflated squared radius
};

// symmetric 3x3 inverse (row-major in, row-major out); plain fp64
__device__ void sym_inv3(const double* a, double* o) {
  const double a00 = a[0], a01 = 0.5 * (a[1] + a[3]), a02 = 0.5 * (a[2] + a[6]);
  const double a11 = a[4], a12 = 0.5 * (a[5] + a[7]), a22 = a[8];
  const double c00 = a11 * a22 - a12 * a12, c01 = a02 * a12 - a01 * a22, c02 = a01 * a12 - a02 * a11;
  const double det = a00 * c00 + a01 * c01 + a02 * c02;
  const double id = 1.0 / det;
  o[0] = c00 * id;
  o[1] = o[3] = c01 * id;
  o[2] = o[6] = c02 * id;
  o[4] = (a00 * a22 - a02 * a02) * id;
  o[5] = o[7] = (a01 * a02 - a00 * a12) * id;
  o[8] = (a00 * a11 - a01 * a01) * id;
}

struct CompGeom {
  double mu[3];
  double hw[3];      // half widths of the ellipsoid's axis-aligned bounding box
  double ax[3][3];   // three further test directions
  double r[3];       // support radius along them
  int c0[3], c1[3];  // cell range (clamped); empty if c0 > c1 somewhere
  long long ncell;
  bool finite;
};

__device__ void comp_geom(const MapDev& m, int k, const LevelBuild& L, CompGeom& g) {
  double M[9];
  sym_inv3(m.cov_inv + (size_t)k * 9, M);
  const double eps = 1e-3 * L.h + 1e-4;
  g.finite = true;
  g.ncell = 1;
  const int dims[3] = {L.nx, L.ny, L.nz};
  for (int i = 0; i < 3; ++i) {
    g.mu[i] = m.mean[(size_t)k * 3 + i];
    g.hw[i] = sqrt(L.T * fmax(M[i * 3 + i], 0.0)) + eps;
    g.finite = g.finite && isfinite(g.hw[i]) && isfinite(g.mu[i]);
    const double f0 = floor((g.mu[i] - g.hw[i] - L.lo[i]) / L.h), f1 = floor((g.mu[i] + g.hw[i] - L.lo[i]) / L.h);
    g.c0[i] = (int)fmax(f0, 0.0);
    g.c1[i] = (int)fmin(f1, (double)(dims[i] - 1));
    if (!(f1 >= 0.0) || !(f0 <= (double)(dims[i] - 1))) {
      g.c0[i] = 0;
      g.c1[i] = -1;
    }
    g.ncell *= (long long)max(g.c1[i] - g.c0[i] + 1, 0);
  }
  for (int j = 0; j < 3; ++j) {
    double a[3] = {m.axis[(size_t)k * 9 + 0 * 3 + j], m.axis[(size_t)k * 9 + 1 * 3 + j], m.axis[(size_t)k * 9 + 2 * 3 + j]};
    const double n = sqrt(a[0] * a[0] + a[1] * a[1] + a[2] * a[2]);
    if (!(n > 0.5 && n < 2.0)) {  // not a usable direction: fall back to a world axis (always valid)
      a[0] = (j == 0);
      a[1] = (j == 1);
      a[2] = (j == 2);
    }
    double q = 0.0;
    for (int r = 0; r < 3; ++r)
      for (int c = 0; c < 3; ++c) q += a[r] * M[r * 3 + c] * a[c];
    for (int i = 0; i < 3; ++i) g.ax[j][i] = a[i];
    g.r[j] = sqrt(L.T * fmax(q, 0.0)) * (1.0 + 1e-9);
    g.finite = g.finite && isfinite(g.r[j]);
  }
}

// does the (inflated) box of cell (ix, iy, iz) possibly touch the ellipsoid?  false only with a separating axis
__device__ bool cell_touches(const CompGeom& g, const LevelBuild& L, int ix, int iy, int iz) {
  const double half = 0.5 * L.h + 1e-3 * L.h + 1e-4;
  const double c[3] = {L.lo[0] + (ix + 0.5) * L.h - g.mu[0], L.lo[1] + (iy + 0.5) * L.h - g.mu[1],
                       L.lo[2] + (iz + 0.5) * L.h - g.mu[2]};
  for (int j = 0; j < 3; ++j) {
    const double p = g.ax[j][0] * c[0] + g.ax[j][1] * c[1] + g.ax[j][2] * c[2];
    const double ext = (fabs(g.ax[j][0]) + fabs(g.ax[j][1]) + fabs(g.ax[j][2])) * half;
    if (fabs(p) - ext > g.r[j]) return false;
  }
  return true;
}

// mode 0: count into cnt[cell] (and, when big_list != null, divert oversized components to the big list);
// mode 1: fill list[start[cell] + cursor[cell]++].  One warp per component.
__global__ void __launch_bounds__(256) grid_bin_kernel(MapDev m, LevelBuild L, int mode, const uint8_t* is_big, uint8_t* mark_big,
                                                       int* nbig, int32_t* big_list, int* cnt, const int32_t* start,
                                                       int32_t* list, int* bad) {
  const int k = (blockIdx.x * blockDim.x + threadIdx.x) >> 5;
  const int lane = threadIdx.x & 31;
  if (k >= m.K) return;
  if (is_big != nullptr && is_big[k]) return;
  CompGeom g;
  comp_geom(m, k, L, g);
  if (!g.finite) {
    if (lane == 0) atomicOr(bad, 1);
    return;
  }
  if (mark_big != nullptr && g.ncell > kBigCells) {
    if (lane == 0) {
      mark_big[k] = 1;
      const int pos = atomicAdd(nbig, 1);
      if (pos < kMaxBig) big_list[pos] = k;
    }
    return;
  }
  const int wx = g.c1[0] - g.c0[0] + 1, wy = g.c1[1] - g.c0[1] + 1;
  for (long long t = lane; t < g.ncell; t += 32) {
    const int ix = g.c0[0] + (int)(t % wx), iy = g.c0[1] + (int)((t / wx) % wy), iz = g.c0[2] + (int)(t / ((long long)wx * wy));
    if (!cell_touches(g, L, ix, iy, iz)) continue;
    const long long cell = ((long long)iz * L.ny + iy) * L.nx + ix;
    const int pos = atomicAdd(cnt + cell, 1);
    if (mode == 1) list[(long long)start[cell] + pos] = k;
  }
}

__global__ void grid_sort_kernel(const int32_t* start, int32_t* list, long long ncell) {
  const long long c = (long long)blockIdx.x * blockDim.x + threadIdx.x;
  if (c >= ncell) return;
  const int b = start[c], e = start[c + 1];
  for (int i = b + 1; i < e; ++i) {
    const int v = list[i];
    int j = i - 1;
    while (j >= b && list[j] > v) {
      list[j + 1] = list[j];
      --j;
    }
    list[j + 1] = v;
  }
}

double env_double(const char* name, double dflt) {
  const char* v = debug_env(name);
  if (v == nullptr || *v == 0) return dflt;
  char* end = nullptr;
  const double x = std::strtod(v, &end);
  return (end != v && std::isfinite(x) && x > 0.0) ? x : dflt;
}

}  // namespace

void free_assoc_grid(MapDev& m) {
  for (int l = 0; l < kGridLevels; ++l) {
    cudaFree(m.grid.lev[l].start);
    cudaFree(m.grid.lev[l].list);
    m.grid.lev[l] = GridLevel();
  }
  cudaFree(m.grid.big);
  m.grid.big = nullptr;
  m.grid.nbig = 0;
  m.grid.nlev = 0;
}

cudaError_t build_assoc_grid(MapDev& m, const double lo[3], const double hi[3], const gmmloc_grid_options_t* opt) {
  free_assoc_grid(m);
  if (m.rec == nullptr || m.K <= 0) return cudaSuccess;
  if (opt != nullptr && opt->enable == 0) return cudaSuccess;
  if (const char* off = debug_env("GMMLOC_GRID")) {
    if (off[0] == '0') return cudaSuccess;
  }
  // explicit options (gmmloc_map_create_ex) first, then the debug-build environment, then the defaults
  auto pick = [&](double from_opt, const char* name, double dflt) {
    return (opt != nullptr && from_opt > 0.0 && std::isfinite(from_opt)) ? from_opt : env_double(name, dflt);
  };
  cudaError_t e;
  // cell size from the component density: 0.6 (V / K)^(1/3), about one component centre per 4.6 cells at level 0.  (Round 1
  // used 0.75: with the warp-cooperative rounds of the sorted scan, which no longer need a cell full of landmarks to be
  // efficient, finer cells with shorter lists win — C2 sort + scan 0.615 -> 0.576 ms, C3 0.408 -> 0.357 ms; at 0.5 the
  // C5 map (50 000 components) outgrows the entry budget.)
  double ext[3], vol = 1.0;
  for (int i = 0; i < 3; ++i) {
    ext[i] = std::fmax(hi[i] - lo[i], 1.0);
    vol *= ext[i];
  }
  const double h0 = pick(opt ? opt->cell_size : 0.0, "GMMLOC_GRID_H",
                         std::fmin(std::fmax(env_double("GMMLOC_GRID_HSCALE", 0.6) * std::cbrt(vol / m.K), 0.05), 4.0));
  // Three levels: T = 25 and T = 144 on the fine grid, T = 400 on a grid of twice the cell size.  (T0 was 36 while the
  // lanes level 0 could not decide walked the next list one by one; now that they meet warp-cooperatively at level 1 the
  // shorter T = 25 lists pay although 1.4 % of the C2 landmarks move on: C2 0.652 -> 0.615 ms; 20 / 30: 0.626 / 0.628.)  The middle level is
  // what the shipped maps need: their planes are millimetres thin, so a stereo depth error of one or two centimetres
  // puts a third of the landmarks beyond 6 sigma of their component (35 % of the C3 landmarks have a winner with
  // d2 > 36, 7 % with d2 > 144, 1.3 % with d2 > 400); without it they all walk the long coarse lists.
  // GMMLOC_GRID_LEVELS=2 builds the fine T = 25 and the coarse T = 400 level only.
  const int want_levels = (int)pick(opt ? (double)opt->levels : 0.0, "GMMLOC_GRID_LEVELS", (double)kGridLevels) >= kGridLevels ? kGridLevels : 2;
  double tmax[kGridLevels] = {pick(opt ? opt->tmax[0] : 0.0, "GMMLOC_GRID_TMAX0", 25.0), pick(opt ? opt->tmax[1] : 0.0, "GMMLOC_GRID_TMAX1", 144.0),
                              pick(opt ? opt->tmax[2] : 0.0, "GMMLOC_GRID_TMAX2", 400.0)};
  double margin[kGridLevels] = {1.0, 1.0, 2.0};
  int scale[kGridLevels] = {1, 1, 2};
  if (want_levels == 2) {
    tmax[1] = tmax[2];
    margin[1] = margin[2];
    scale[1] = scale[2];
  }

  uint8_t* d_isbig = nullptr;
  int* d_misc = nullptr;  // [0] nbig, [1] bad flag
  int32_t* d_big = nullptr;
  int* d_cnt = nullptr;
  auto cleanup = [&](cudaError_t rc, bool keep) {
    cudaFree(d_isbig);
    cudaFree(d_misc);
    cudaFree(d_cnt);
    if (!keep) {
      cudaFree(d_big);
      free_assoc_grid(m);
    }
    return rc;
  };
  if ((e = cudaMalloc((void**)&d_isbig, (size_t)m.K)) != cudaSuccess) return cleanup(e, false);
  if ((e = cudaMemset(d_isbig, 0, (size_t)m.K)) != cudaSuccess) return cleanup(e, false);
  if ((e = cudaMalloc((void**)&d_misc, 2 * sizeof(int))) != cudaSuccess) return cleanup(e, false);
  if ((e = cudaMemset(d_misc, 0, 2 * sizeof(int))) != cudaSuccess) return cleanup(e, false);
  if ((e = cudaMalloc((void**)&d_big, kMaxBig * sizeof(int32_t))) != cudaSuccess) return cleanup(e, false);

  const unsigned blocks = (unsigned)(((long long)m.K * 32 + 255) / 256);
  int nlev = 0;
  for (int l = 0; l < want_levels; ++l) {
    LevelBuild L;
    L.h = h0 * (double)scale[l];
    long long ncell;
    for (;;) {
      // the device computes the cell index in fp32 from (float)lo and (float)(1/h): build with exactly those values
      const float inv_hf = (float)(1.0 / L.h);
      L.h = 1.0 / (double)inv_hf;
      int dims[3];
      for (int i = 0; i < 3; ++i) {
        const float lof = (float)(lo[i] - margin[l]);
        L.lo[i] = (double)lof;
        dims[i] = (int)std::ceil((hi[i] + margin[l] - L.lo[i]) / L.h) + 1;
      }
      L.nx = dims[0];
      L.ny = dims[1];
      L.nz = dims[2];
      ncell = (long long)L.nx * L.ny * L.nz;
      if (ncell <= kMaxCells && L.nx < 2000 && L.ny < 2000 && L.nz < 2000) break;
      L.h *= 1.26;
    }
    L.T = tmax[l] * (1.0 + 1e-3);
    cudaFree(d_cnt);
    d_cnt = nullptr;
    if ((e = cudaMalloc((void**)&d_cnt, sizeof(int) * (size_t)ncell)) != cudaSuccess) return cleanup(e, false);
    if ((e = cudaMemset(d_cnt, 0, sizeof(int) * (size_t)ncell)) != cudaSuccess) return cleanup(e, false);
    // pass 1: count (level 0 also diverts the oversized components)
    grid_bin_kernel<<<blocks, 256>>>(m, L, 0, l == 0 ? nullptr : d_isbig, l == 0 ? d_isbig : nullptr, d_misc, d_big, d_cnt, nullptr,
                                     nullptr, d_misc + 1);
    if ((e = cudaGetLastError()) != cudaSuccess) return cleanup(e, false);
    int h_misc[2];
    if ((e = cudaMemcpy(h_misc, d_misc, sizeof h_misc, cudaMemcpyDeviceToHost)) != cudaSuccess) return cleanup(e, false);
    if (h_misc[1] != 0 || h_misc[0] > kMaxBig) {  // non-finite geometry or too many oversized components: no grid
      cudaFree(d_big);
      d_big = nullptr;
      free_assoc_grid(m);
      return cleanup(cudaSuccess, true);
    }
    std::vector<int> h_cnt((size_t)ncell);
    if ((e = cudaMemcpy(h_cnt.data(), d_cnt, sizeof(int) * (size_t)ncell, cudaMemcpyDeviceToHost)) != cudaSuccess)
      return cleanup(e, false);
    std::vector<int32_t> h_start((size_t)ncell + 1);
    long long total = 0;
    for (long long c = 0; c < ncell; ++c) {
      h_start[(size_t)c] = (int32_t)total;
      total += h_cnt[(size_t)c];
      if (total > kMaxEntries) break;
    }
    if (total > kMaxEntries) {  // this level's lists would be too large: skip it (without level 0 there is no grid)
      if (l == 0) break;
      continue;
    }
    h_start[(size_t)ncell] = (int32_t)total;
    GridLevel& G = m.grid.lev[nlev];
    if ((e = cudaMalloc((void**)&G.start, sizeof(int32_t) * ((size_t)ncell + 1))) != cudaSuccess) return cleanup(e, false);
    if ((e = cudaMalloc((void**)&G.list, sizeof(int32_t) * (size_t)std::max<long long>(total, 1))) != cudaSuccess)
      return cleanup(e, false);
    if ((e = cudaMemcpy(G.start, h_start.data(), sizeof(int32_t) * ((size_t)ncell + 1), cudaMemcpyHostToDevice)) != cudaSuccess)
      return cleanup(e, false);
    if ((e = cudaMemset(d_cnt, 0, sizeof(int) * (size_t)ncell)) != cudaSuccess) return cleanup(e, false);
    grid_bin_kernel<<<blocks, 256>>>(m, L, 1, d_isbig, nullptr, d_misc, d_big, d_cnt, G.start, G.list, d_misc + 1);
    if ((e = cudaGetLastError()) != cudaSuccess) return cleanup(e, false);
    grid_sort_kernel<<<(unsigned)((ncell + 127) / 128), 128>>>(G.start, G.list, ncell);
    if ((e = cudaGetLastError()) != cudaSuccess) return cleanup(e, false);
    if ((e = cudaDeviceSynchronize()) != cudaSuccess) return cleanup(e, false);
    for (int i = 0; i < 3; ++i) G.lo[i] = (float)L.lo[i];
    G.inv_h = (float)(1.0 / L.h);
    G.nx = L.nx;
    G.ny = L.ny;
    G.nz = L.nz;
    G.tmax = (float)tmax[l];
    if ((double)G.tmax > tmax[l]) G.tmax = nextafterf(G.tmax, 0.f);
    G.entries = total;
    if (l == 0) m.grid.nbig = h_misc[0];
    ++nlev;
  }
  if (nlev == 0) {
    cudaFree(d_big);
    d_big = nullptr;
    free_assoc_grid(m);
    return cleanup(cudaSuccess, true);
  }
  // sort the big list (ascending ids) on the host: it is tiny
  if (m.grid.nbig > 0) {
    std::vector<int32_t> hb((size_t)m.grid.nbig);
    if ((e = cudaMemcpy(hb.data(), d_big, sizeof(int32_t) * hb.size(), cudaMemcpyDeviceToHost)) != cudaSuccess) return cleanup(e, false);
    std::sort(hb.begin(), hb.end());
    if ((e = cudaMemcpy(d_big, hb.data(), sizeof(int32_t) * hb.size(), cudaMemcpyHostToDevice)) != cudaSuccess) return cleanup(e, false);
  }
  m.grid.big = d_big;
  m.grid.nlev = nlev;
  return cleanup(cudaSuccess, true);
}

}  // namespace gmmloc
