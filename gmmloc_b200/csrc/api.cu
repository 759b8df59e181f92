// api.cu — the C ABI of include/gmmloc_b200.h: handles, workspaces, host<->device staging, launches.
//
// No CPU fallback anywhere: every entry point that computes needs a CUDA device and fails with
// GMMLOC_ECUDA otherwise.  Nothing here links, loads or calls oracle/.
#include <algorithm>
#include <cmath>
#include <cstdio>
#include <cstdlib>
#include <cstring>
#include <new>
#include <string>
#include <vector>

#include "common.cuh"

using namespace gmmloc;

namespace {

thread_local std::string g_err;

gmmloc_status_t fail_cuda(cudaError_t e, const char* where) {
  g_err = std::string(where) + ": " + cudaGetErrorString(e);
  cudaGetLastError();  // clear the sticky-less error state
  return e == cudaErrorMemoryAllocation ? GMMLOC_ENOMEM : GMMLOC_ECUDA;
}
gmmloc_status_t fail(gmmloc_status_t s, const std::string& msg) {
  g_err = msg;
  return s;
}

#define CU(call)                                            \
  do {                                                      \
    cudaError_t e__ = (call);                               \
    if (e__ != cudaSuccess) return fail_cuda(e__, #call);   \
  } while (0)

// No exception may cross the C ABI: entry points that allocate host memory (std::vector / std::string) are
// function-try-blocks closed by this handler.
#define GMMLOC_ABI_CATCH                                                                                  \
  catch (const std::bad_alloc&) { return fail(GMMLOC_ENOMEM, "host allocation failed"); }                 \
  catch (const std::exception& ex__) { return fail(GMMLOC_EINVAL, std::string("exception: ") + ex__.what()); } \
  catch (...) { return fail(GMMLOC_EINVAL, "unknown exception"); }

// grow-only device buffer
struct DevBuf {
  void* p = nullptr;
  size_t cap = 0;
  cudaError_t reserve(size_t bytes) {
    if (bytes <= cap) return cudaSuccess;
    if (p) cudaFree(p);
    p = nullptr;
    cap = 0;
    const size_t want = bytes + bytes / 8 + 256;
    cudaError_t e = cudaMalloc(&p, want);
    if (e == cudaSuccess) cap = want;
    return e;
  }
  void release() {
    if (p) cudaFree(p);
    p = nullptr;
    cap = 0;
  }
  template <class T>
  T* as() const { return static_cast<T*>(p); }
};

void fill_params(const gmmloc_camera_t* cam, const gmmloc_params_t* prm, FrameParamsDev& d) {
  d.fx = cam->fx; d.fy = cam->fy; d.cx = cam->cx; d.cy = cam->cy; d.bf = cam->bf;
  d.tri_lambda2 = prm->tri_lambda2;
  d.ba_lambda2 = prm->ba_lambda2;
  d.tri_str_thresh = prm->tri_str_thresh;
  d.mdist2_gate = prm->mdist2_gate;
  d.chi2_mono = prm->chi2_mono;
  d.chi2_stereo = prm->chi2_stereo;
  d.delta_mono = (double)(float)std::sqrt(5.991);    // const float deltaMono = sqrt(5.991)  tracking_opt.cpp:57
  d.delta_stereo = (double)(float)std::sqrt(7.815);  // tracking_opt.cpp:58
  for (int i = 0; i < 8; ++i) d.s2inv[i] = gmmloc_sigma2_inv(i);
  d.gn_iters = prm->gn_iters;
  d.robust = prm->robust;
  d.factor_mode = prm->factor_mode;
  d.tri_check_str_chi2 = prm->tri_check_str_chi2;
}

// ---- .gmm wire format (protobuf_utils.cpp:12-101, GMM.proto:5-14) ------------------------------
bool get_varint(const std::vector<unsigned char>& buf, size_t& pos, uint64_t& out) {
  out = 0;
  int shift = 0;
  while (pos < buf.size() && shift < 64) {
    const unsigned char c = buf[pos++];
    out |= (uint64_t)(c & 0x7fu) << shift;
    if ((c & 0x80u) == 0) return true;
    shift += 7;
  }
  return false;
}
void put_varint(std::string& s, uint64_t v) {
  while (v > 0x7f) {
    s.push_back((char)((v & 0x7f) | 0x80));
    v >>= 7;
  }
  s.push_back((char)v);
}
void put_double(std::string& s, double v) {
  char t[8];
  std::memcpy(t, &v, 8);
  s.append(t, 8);
}

struct GmmFile {
  std::vector<double> mean, cov;
  std::vector<uint8_t> flags;  // K x 2 as stored
};

// parses one ComponentProto message occupying buf[pos, end)
bool parse_component(const std::vector<unsigned char>& buf, size_t pos, size_t end, double* mean, double* cov,
                     uint8_t* flags) {
  int nmean = 0, ncov = 0;
  flags[0] = flags[1] = 0;
  auto take = [&](int field, double v) {
    if (field == 3) {
      if (nmean < 3) mean[nmean] = v;
      ++nmean;
    } else if (field == 4) {
      if (ncov < 9) cov[ncov] = v;  // row-major fill like the loader, gmm_utils.cpp:55-59
      ++ncov;
    }
  };
  while (pos < end) {
    uint64_t tag;
    if (!get_varint(buf, pos, tag)) return false;
    const int field = (int)(tag >> 3);
    switch (tag & 7u) {
      case 0: {
        uint64_t v;
        if (!get_varint(buf, pos, v)) return false;
        if (field == 1) flags[0] = v ? 1 : 0;
        if (field == 2) flags[1] = v ? 1 : 0;
        break;
      }
      case 1: {
        if (end - pos < 8) return false;
        double v;
        std::memcpy(&v, &buf[pos], 8);
        take(field, v);
        pos += 8;
        break;
      }
      case 2: {
        uint64_t len;
        if (!get_varint(buf, pos, len) || len > (uint64_t)(end - pos)) return false;
        if ((field == 3 || field == 4) && (len & 7u) == 0) {
          for (uint64_t off = 0; off < len; off += 8) {
            double v;
            std::memcpy(&v, &buf[pos + off], 8);
            take(field, v);
          }
        }
        pos += len;
        break;
      }
      case 5:
        if (end - pos < 4) return false;
        pos += 4;
        break;
      default:
        return false;
    }
  }
  return pos == end && nmean == 3 && ncov == 9;  // CHECK_EQ sizes, gmm_utils.cpp:50-51
}

gmmloc_status_t read_gmm(const char* path, GmmFile& out) {
  FILE* fp = std::fopen(path, "rb");
  if (!fp) return fail(GMMLOC_EIO, std::string("cannot open ") + path);
  std::vector<unsigned char> buf;
  unsigned char chunk[1 << 16];
  size_t got;
  while ((got = std::fread(chunk, 1, sizeof chunk, fp)) > 0) buf.insert(buf.end(), chunk, chunk + got);
  std::fclose(fp);
  size_t pos = 0;
  uint64_t count;
  if (!get_varint(buf, pos, count)) return fail(GMMLOC_EIO, "could not read the message count");
  if (count == 0) return fail(GMMLOC_EIO, "protobuf file empty");  // gmm_utils.cpp:32-35
  // a component message holds at least 3 + 9 doubles: a count the file cannot hold is a corrupt header (the reference
  // reads through protobuf's CodedInputStream, which fails on the first missing message instead of allocating)
  if (count > buf.size() / 96) return fail(GMMLOC_EIO, "message count exceeds the file size");
  out.mean.resize(count * 3);
  out.cov.resize(count * 9);
  out.flags.resize(count * 2);
  for (uint64_t k = 0; k < count; ++k) {
    uint64_t len;
    if (!get_varint(buf, pos, len) || len == 0 || len > (uint64_t)(buf.size() - pos))
      return fail(GMMLOC_EIO, "truncated component message " + std::to_string(k));
    if (!parse_component(buf, pos, pos + len, &out.mean[k * 3], &out.cov[k * 9], &out.flags[k * 2]))
      return fail(GMMLOC_EIO, "malformed component message " + std::to_string(k));
    pos += len;
  }
  return GMMLOC_OK;
}

}  // namespace

struct gmmloc_map {
  MapDev d;
  int device = 0;
  std::vector<double> h_mean, h_cov;  // kept for save / get_components
};

// one compute stream with the work lists / workspaces / timing events its kernels use
struct Lane {
  cudaStream_t stream = nullptr;
  DevBuf flag_count, flag_idx, flag_thr;  // association work lists
  DevBuf next_frame;                      // frame counter of the persistent frame-GN kernel
  DevBuf sort_cnt, sort_perm;             // cell histogram / landmark order of the cell-sorted grid scan
  DevBuf gn_store;                        // stored linearisations of the mixed-precision frame GN (per resident CTA)
  DevBuf gn_prep;                         // class-sorted frames written by frame_prep_kernel (32 B per point)
  DevBuf xws, fws;                        // frame-GN global workspaces (N too large for shared memory)
  int* h_flag_count = nullptr;            // pinned
  cudaEvent_t ev[4] = {nullptr, nullptr, nullptr, nullptr};  // scan begin | scan end | re-rank end | GN end
  bool ev_assoc = false, ev_gn = false;
};

struct gmmloc_ctx {
  int device = 0;
  // lane 0 serves the _async entry points; the pipelined host entry point alternates chunks between the
  // two lanes so that the tail wave of one chunk's kernels overlaps the head of the next chunk's
  Lane lanes[2];
  int active = 0;
  Lane& L() { return lanes[active]; }
  DevBuf in_pts, in_obs, in_oct, in_pose, in_misc;  // device copies of host inputs
  DevBuf out_assign, out_d2, out_pose, out_pts, out_outlier, out_status, out_misc;
  // rendered views of the last gmmloc_render_views (≙ GMM::components2d_) + keyframe-association staging
  DevBuf view_pose, view_count, view_status, view_ids, view_rec, view_wids, view_wrec;
  DevBuf kf_obs, kf_depth, kf_oct, kf_cand, kf_cand_pos, kf_pt3d, kf_assoc;
  int view_B = 0, view_cap = 0;
  cudaStream_t h2d = nullptr, d2h = nullptr;        // copy streams of the pipelined host entry point
  std::vector<cudaEvent_t> ev_in, ev_done;          // per-chunk events (inputs landed / results ready)
  gmmloc_stats_t stats{};
  int assoc_kind = GMMLOC_ASSOC_AUTO;
  int gn_precision = GMMLOC_GN_STREAM;
  int last_assoc_tc = 0;
  int last_gn = -1;              // GMMLOC_GN_* of the frame-GN kernel the last gmmloc_frames_solve[_async] ran
  int32_t* h_stop = nullptr;     // page-locked, device-mapped mirror of the caller's BA stop flag (gmmloc_local_ba)
  cudaEvent_t ev_ba = nullptr;   // end of the BA kernel (the host polls it while it mirrors the flag)
  int opt_gn_prep = -1;          // GMMLOC_OPT_GN_PREP
  int opt_e2e_chunk = 0;         // GMMLOC_OPT_E2E_CHUNK
  int opt_e2e_short_first = 0;   // GMMLOC_OPT_E2E_SHORT_FIRST
  int opt_e2e_taper = 1;         // GMMLOC_OPT_E2E_TAPER
  int opt_ba_cluster = 0;        // GMMLOC_OPT_BA_CLUSTER
};

extern "C" {

const char* gmmloc_version(void) { return "gmmloc_b200 0.1 (sm_100a)"; }

const char* gmmloc_status_string(gmmloc_status_t s) {
  switch (s) {
    case GMMLOC_OK: return "ok";
    case GMMLOC_EINVAL: return "invalid argument";
    case GMMLOC_ENOMEM: return "out of memory";
    case GMMLOC_ECUDA: return "CUDA error / no device";
    case GMMLOC_ENUMERIC: return "numeric failure";
    case GMMLOC_EIO: return "I/O error";
  }
  return "unknown";
}

const char* gmmloc_last_error(void) { return g_err.c_str(); }

void gmmloc_default_params(gmmloc_params_t* p) {
  if (!p) return;
  p->tri_lambda2 = 400.0;                 // v1.yaml:32
  p->ba_lambda2 = 400.0;                  // v1.yaml:37
  p->tri_str_thresh = (double)0.0064f;    // loc::tri_str_thresh is a float in the reference (config.h)
  p->tri_check_str_chi2 = 1;              // v1.yaml:34
  p->mdist2_gate = 9.0;                   // gmmloc_opt.cpp:230
  p->chi2_mono = 5.991;                   // tracking_opt.cpp:145
  p->chi2_stereo = 7.815;                 // tracking_opt.cpp:146
  p->gn_iters = 10;                       // tracking_opt.cpp:147
  p->robust = 0;
  p->factor_mode = GMMLOC_MODE_JOINT;
}

float gmmloc_sigma2_inv(int octave) {
  // scale_factors[i] = scale_factors[i-1] * 1.2f; sigma2 = sf*sf; sigma2_inv = 1.0f / sigma2   (float)
  if (octave < 0) octave = 0;
  if (octave > 7) octave = 7;
  volatile float sf = 1.0f;
  float inv = 1.0f;
  for (int i = 1; i <= octave; ++i) {
    sf = sf * 1.2f;
    const float s2 = sf * sf;
    inv = 1.0f / s2;
  }
  return inv;
}

// ---- map -----------------------------------------------------------------------------------------
void gmmloc_map_destroy(gmmloc_map_t* m) {
  if (!m) return;
  int prev = 0;
  cudaGetDevice(&prev);
  cudaSetDevice(m->device);
  cudaFree(m->d.mean); cudaFree(m->d.cov); cudaFree(m->d.cov_inv); cudaFree(m->d.det); cudaFree(m->d.scale);
  cudaFree(m->d.axis); cudaFree(m->d.sqrt_info); cudaFree(m->d.is_deg); cudaFree(m->d.is_sal); cudaFree(m->d.scan);
  cudaFree(m->d.nbr_row_ptr); cudaFree(m->d.nbr_col); cudaFree(m->d.tc_table); cudaFree(m->d.rec); cudaFree(m->d.gn_tab);
  free_assoc_grid(m->d);
  cudaSetDevice(prev);
  delete m;
}

gmmloc_status_t gmmloc_map_create(const double* mean, const double* cov, int K, double deg_eps, int device,
                                  gmmloc_map_t** out) {
  return gmmloc_map_create_ex(mean, cov, K, deg_eps, device, nullptr, out);
}

gmmloc_status_t gmmloc_map_create_ex(const double* mean, const double* cov, int K, double deg_eps, int device,
                                     const gmmloc_grid_options_t* grid_opt, gmmloc_map_t** out) try {
  if (!mean || !cov || !out || K <= 0) return fail(GMMLOC_EINVAL, "gmmloc_map_create: null pointer or K <= 0");
  *out = nullptr;
  CU(cudaSetDevice(device));
  gmmloc_map* m = new (std::nothrow) gmmloc_map();
  if (!m) return fail(GMMLOC_ENOMEM, "host allocation failed");
  m->device = device;
  try {
    m->h_mean.assign(mean, mean + (size_t)K * 3);
    m->h_cov.assign(cov, cov + (size_t)K * 9);
  } catch (const std::bad_alloc&) {
    delete m;
    return fail(GMMLOC_ENOMEM, "host allocation failed");
  }
  MapDev& d = m->d;
  d.K = K;
  d.Kpad = (K + kChunk - 1) / kChunk * kChunk;
  double mmax = 0.0, lo[3] = {1e300, 1e300, 1e300}, hi[3] = {-1e300, -1e300, -1e300};
  for (size_t i = 0; i < (size_t)K * 3; ++i) {
    if (!std::isfinite(mean[i])) {
      delete m;
      return fail(GMMLOC_EINVAL, "non-finite mean");
    }
    mmax = std::fmax(mmax, std::fabs(mean[i]));
    lo[i % 3] = std::fmin(lo[i % 3], mean[i]);
    hi[i % 3] = std::fmax(hi[i % 3], mean[i]);
  }
  for (int i = 0; i < 3; ++i) d.tc_centre[i] = (float)(0.5 * (lo[i] + hi[i]));  // tensor-core scan works on centred landmarks
  // landmarks are certified by the fp32 scan only inside |x|_inf <= xmax; anything farther out takes the
  // exact fp64 path (still correct, just slower)
  const double xmax = 1.5 * mmax + 1.0;
  cudaError_t e = cudaSuccess;
  auto alloc = [&](void** p, size_t bytes) {
    if (e == cudaSuccess) e = cudaMalloc(p, bytes);
  };
  alloc((void**)&d.mean, sizeof(double) * 3 * K);
  alloc((void**)&d.cov, sizeof(double) * 9 * K);
  alloc((void**)&d.cov_inv, sizeof(double) * 9 * K);
  alloc((void**)&d.det, sizeof(double) * K);
  alloc((void**)&d.scale, sizeof(double) * 3 * K);
  alloc((void**)&d.axis, sizeof(double) * 9 * K);
  alloc((void**)&d.sqrt_info, sizeof(double) * 9 * K);
  alloc((void**)&d.is_deg, K);
  alloc((void**)&d.is_sal, K);
  alloc((void**)&d.scan, sizeof(float) * kScanFloats * d.Kpad);
  alloc((void**)&d.rec, sizeof(float) * 12 * (size_t)K);
  alloc((void**)&d.gn_tab, sizeof(float) * 8 * (size_t)K);
  if (kChunk == 16) alloc((void**)&d.tc_table, (size_t)((d.Kpad + kTcTileComps - 1) / kTcTileComps) * kTcTileBytes);
  int* d_err = nullptr;
  alloc((void**)&d_err, 4 * sizeof(int));
  if (e != cudaSuccess) {
    cudaFree(d_err);
    gmmloc_map_destroy(m);
    return fail_cuda(e, "cudaMalloc(map)");
  }
  auto bail = [&](cudaError_t ce, const char* where) {
    cudaFree(d_err);
    gmmloc_map_destroy(m);
    return fail_cuda(ce, where);
  };
  if ((e = cudaMemcpy(d.mean, mean, sizeof(double) * 3 * K, cudaMemcpyHostToDevice)) != cudaSuccess) return bail(e, "H2D mean");
  if ((e = cudaMemcpy(d.cov, cov, sizeof(double) * 9 * K, cudaMemcpyHostToDevice)) != cudaSuccess) return bail(e, "H2D cov");
  if ((e = cudaMemset(d_err, 0, 4 * sizeof(int))) != cudaSuccess) return bail(e, "memset");
  if ((e = launch_map_build(d, deg_eps, xmax, d_err, reinterpret_cast<unsigned int*>(d_err + 1), nullptr)) != cudaSuccess)
    return bail(e, "map_build launch");
  int h[4] = {0, 0, 0, 0};
  if ((e = cudaMemcpy(h, d_err, sizeof h, cudaMemcpyDeviceToHost)) != cudaSuccess) return bail(e, "map_build");
  cudaFree(d_err);
  if (h[0] != 0) {
    gmmloc_map_destroy(m);
    return fail(GMMLOC_EINVAL, "covariance not symmetric positive definite (or non-finite)");
  }
  float alpha;
  std::memcpy(&alpha, &h[1], 4);
  d.alpha = alpha;
  d.xmax = (float)xmax;
  std::memcpy(&d.tc_err, &h[2], 4);
  d.tc_ok = (d.tc_table != nullptr && h[3] == 0) ? 1 : 0;
  // exact spatial index for the association (grid_build.cu); a map it cannot index keeps the brute-force scan
  if ((e = build_assoc_grid(d, lo, hi, grid_opt)) != cudaSuccess) {
    gmmloc_map_destroy(m);
    return fail_cuda(e, "build_assoc_grid");
  }
  *out = m;
  return GMMLOC_OK;
} GMMLOC_ABI_CATCH

gmmloc_status_t gmmloc_map_load_gmm(const char* path, double deg_eps, int device, gmmloc_map_t** out) try {
  if (!path || !out) return fail(GMMLOC_EINVAL, "gmmloc_map_load_gmm: null pointer");
  GmmFile f;
  gmmloc_status_t s = read_gmm(path, f);
  if (s != GMMLOC_OK) return s;  // bool false + LOG(ERROR) in the reference, gmm_utils.cpp:19-48
  return gmmloc_map_create(f.mean.data(), f.cov.data(), (int)(f.mean.size() / 3), deg_eps, device, out);
} GMMLOC_ABI_CATCH

gmmloc_status_t gmmloc_map_save_gmm(const gmmloc_map_t* map, const char* path) try {
  if (!map || !path) return fail(GMMLOC_EINVAL, "gmmloc_map_save_gmm: null pointer");
  const int K = map->d.K;
  std::vector<uint8_t> deg(K), sal(K);
  int prev = 0;
  cudaGetDevice(&prev);
  CU(cudaSetDevice(map->device));
  CU(cudaMemcpy(deg.data(), map->d.is_deg, K, cudaMemcpyDeviceToHost));
  CU(cudaMemcpy(sal.data(), map->d.is_sal, K, cudaMemcpyDeviceToHost));
  cudaSetDevice(prev);
  std::string s;
  put_varint(s, (uint64_t)K);  // writeProtoMsgCountToStream, gmm_utils.cpp:87-92
  for (int k = 0; k < K; ++k) {
    std::string msg;
    msg.push_back(0x08); msg.push_back((char)deg[k]);  // is_degenerated  gmm_utils.cpp:97
    msg.push_back(0x10); msg.push_back((char)sal[k]);  // is_salient      gmm_utils.cpp:99
    msg.push_back(0x1a); msg.push_back(24);
    for (int i = 0; i < 3; ++i) put_double(msg, map->h_mean[(size_t)k * 3 + i]);
    msg.push_back(0x22); msg.push_back(72);
    for (int i = 0; i < 9; ++i) put_double(msg, map->h_cov[(size_t)k * 9 + (i % 3) * 3 + i / 3]);  // cov(i): column-major, :107
    put_varint(s, msg.size());
    s += msg;
  }
  FILE* fp = std::fopen(path, "wb");
  if (!fp) return fail(GMMLOC_EIO, std::string("cannot open ") + path + " for writing");
  const size_t w = std::fwrite(s.data(), 1, s.size(), fp);
  std::fclose(fp);
  if (w != s.size()) return fail(GMMLOC_EIO, "short write");
  return GMMLOC_OK;
} GMMLOC_ABI_CATCH

int gmmloc_map_size(const gmmloc_map_t* map) { return map ? map->d.K : 0; }
int gmmloc_map_device(const gmmloc_map_t* map) { return map ? map->device : -1; }

gmmloc_status_t gmmloc_map_get_components(const gmmloc_map_t* map, double* mean, double* cov, double* cov_inv,
                                          double* det, double* scale, double* axis, double* sqrt_info,
                                          uint8_t* is_degenerated, uint8_t* is_salient) {
  if (!map) return fail(GMMLOC_EINVAL, "null map");
  const size_t K = map->d.K;
  int prev = 0;
  cudaGetDevice(&prev);
  CU(cudaSetDevice(map->device));
  if (mean) CU(cudaMemcpy(mean, map->d.mean, 24 * K, cudaMemcpyDeviceToHost));
  if (cov) CU(cudaMemcpy(cov, map->d.cov, 72 * K, cudaMemcpyDeviceToHost));
  if (cov_inv) CU(cudaMemcpy(cov_inv, map->d.cov_inv, 72 * K, cudaMemcpyDeviceToHost));
  if (det) CU(cudaMemcpy(det, map->d.det, 8 * K, cudaMemcpyDeviceToHost));
  if (scale) CU(cudaMemcpy(scale, map->d.scale, 24 * K, cudaMemcpyDeviceToHost));
  if (axis) CU(cudaMemcpy(axis, map->d.axis, 72 * K, cudaMemcpyDeviceToHost));
  if (sqrt_info) CU(cudaMemcpy(sqrt_info, map->d.sqrt_info, 72 * K, cudaMemcpyDeviceToHost));
  if (is_degenerated) CU(cudaMemcpy(is_degenerated, map->d.is_deg, K, cudaMemcpyDeviceToHost));
  if (is_salient) CU(cudaMemcpy(is_salient, map->d.is_sal, K, cudaMemcpyDeviceToHost));
  cudaSetDevice(prev);
  return GMMLOC_OK;
}

// counting pass + host prefix sum + fill pass; on success rp / dcol (/ ddist) hold the CSR lists on the device
static gmmloc_status_t neighbors_device(const gmmloc_map_t* map, double thresh, bool want_dist, DevBuf& rp, DevBuf& dcol,
                                        DevBuf& ddist, std::vector<long long>& h_rp) try {
  const int K = map->d.K;
  DevBuf counts;
  cudaError_t e;
  if ((e = counts.reserve(sizeof(int) * (size_t)K)) != cudaSuccess) return fail_cuda(e, "cudaMalloc(counts)");
  auto done = [&](gmmloc_status_t s) {
    counts.release();
    return s;
  };
  if ((e = launch_bh_neighbors(map->d, thresh, counts.as<int>(), nullptr, nullptr, nullptr, nullptr)) != cudaSuccess)
    return done(fail_cuda(e, "bh_neighbors (count)"));
  std::vector<int> h_counts((size_t)K);
  if ((e = cudaMemcpy(h_counts.data(), counts.p, sizeof(int) * (size_t)K, cudaMemcpyDeviceToHost)) != cudaSuccess)
    return done(fail_cuda(e, "bh_neighbors (count)"));
  h_rp.assign((size_t)K + 1, 0);
  for (int i = 0; i < K; ++i) h_rp[i + 1] = h_rp[i] + h_counts[i];
  const size_t n = (size_t)h_rp[K];
  if ((e = rp.reserve(sizeof(long long) * ((size_t)K + 1))) != cudaSuccess) return done(fail_cuda(e, "cudaMalloc(row_ptr)"));
  if ((e = cudaMemcpy(rp.p, h_rp.data(), sizeof(long long) * ((size_t)K + 1), cudaMemcpyHostToDevice)) != cudaSuccess)
    return done(fail_cuda(e, "H2D row_ptr"));
  if ((e = dcol.reserve(sizeof(int32_t) * (n ? n : 1))) != cudaSuccess) return done(fail_cuda(e, "cudaMalloc(col)"));
  if (want_dist && (e = ddist.reserve(sizeof(double) * (n ? n : 1))) != cudaSuccess) return done(fail_cuda(e, "cudaMalloc(dist)"));
  if (n > 0) {
    if ((e = launch_bh_neighbors(map->d, thresh, nullptr, rp.as<long long>(), dcol.as<int32_t>(),
                                 want_dist ? ddist.as<double>() : nullptr, nullptr)) != cudaSuccess)
      return done(fail_cuda(e, "bh_neighbors (fill)"));
    if ((e = cudaDeviceSynchronize()) != cudaSuccess) return done(fail_cuda(e, "bh_neighbors (fill)"));
  }
  return done(GMMLOC_OK);
} GMMLOC_ABI_CATCH

gmmloc_status_t gmmloc_map_neighbors(const gmmloc_map_t* map, double thresh, int64_t* row_ptr, int32_t* col, double* dist,
                                     int64_t cap, int64_t* count) try {
  if (!map || !row_ptr || !count) return fail(GMMLOC_EINVAL, "gmmloc_map_neighbors: null pointer");
  const int K = map->d.K;
  int prev = 0;
  cudaGetDevice(&prev);
  CU(cudaSetDevice(map->device));
  DevBuf rp, dcol, ddist;
  std::vector<long long> h_rp;
  auto done = [&](gmmloc_status_t s) {
    rp.release(); dcol.release(); ddist.release();
    cudaSetDevice(prev);
    return s;
  };
  gmmloc_status_t st = neighbors_device(map, thresh, dist != nullptr, rp, dcol, ddist, h_rp);
  if (st != GMMLOC_OK) return done(st);
  for (int i = 0; i <= K; ++i) row_ptr[i] = h_rp[i];
  *count = h_rp[K];
  if (col == nullptr || cap < h_rp[K] || h_rp[K] == 0) return done(GMMLOC_OK);
  const size_t n = (size_t)h_rp[K];
  cudaError_t e;
  if ((e = cudaMemcpy(col, dcol.p, sizeof(int32_t) * n, cudaMemcpyDeviceToHost)) != cudaSuccess) return done(fail_cuda(e, "D2H col"));
  if (dist && (e = cudaMemcpy(dist, ddist.p, sizeof(double) * n, cudaMemcpyDeviceToHost)) != cudaSuccess)
    return done(fail_cuda(e, "D2H dist"));
  return done(GMMLOC_OK);
} GMMLOC_ABI_CATCH

gmmloc_status_t gmmloc_map_build_neighbors(gmmloc_map_t* map, double thresh, int64_t* count) try {
  if (!map) return fail(GMMLOC_EINVAL, "gmmloc_map_build_neighbors: null map");
  int prev = 0;
  cudaGetDevice(&prev);
  CU(cudaSetDevice(map->device));
  DevBuf rp, dcol, ddist;
  std::vector<long long> h_rp;
  gmmloc_status_t st = neighbors_device(map, thresh, false, rp, dcol, ddist, h_rp);
  if (st != GMMLOC_OK) {
    rp.release(); dcol.release();
    cudaSetDevice(prev);
    return st;
  }
  cudaFree(map->d.nbr_row_ptr);
  cudaFree(map->d.nbr_col);
  map->d.nbr_row_ptr = rp.as<long long>();  // ownership moves into the map
  map->d.nbr_col = dcol.as<int32_t>();
  map->d.nbr_count = h_rp[map->d.K];
  map->d.nbr_thresh = thresh;
  if (count) *count = map->d.nbr_count;
  cudaSetDevice(prev);
  return GMMLOC_OK;
} GMMLOC_ABI_CATCH

gmmloc_status_t gmmloc_map_query_point(gmmloc_ctx_t* ctx, const gmmloc_map_t* map, const double* pts, int M, int32_t* nearest,
                                       int32_t* knn, double* min_mdist2) {
  if (!ctx || !map || M < 0) return fail(GMMLOC_EINVAL, "gmmloc_map_query_point: bad argument");
  if (M == 0) return GMMLOC_OK;
  if (!pts || !nearest) return fail(GMMLOC_EINVAL, "gmmloc_map_query_point: null pointer");
  if (map->device != ctx->device) return fail(GMMLOC_EINVAL, "map and context live on different devices");
  CU(cudaSetDevice(ctx->device));
  ctx->active = 0;
  cudaStream_t st = ctx->L().stream;
  const size_t m = (size_t)M, nk = (size_t)query_point_knn();
  // slabs: pts (24) | min_mdist2 (8) | nearest (4) | knn (4 nk)
  const size_t o_md = 24 * m, o_nn = o_md + 8 * m, o_knn = o_nn + 4 * m, total = o_knn + 4 * nk * m;
  CU(ctx->in_misc.reserve(total));
  char* base = ctx->in_misc.as<char>();
  CU(cudaMemcpyAsync(base, pts, 24 * m, cudaMemcpyHostToDevice, st));
  CU(launch_query_point(map->d.mean, map->d.cov_inv, map->d.K, reinterpret_cast<const double*>(base), M,
                        reinterpret_cast<int32_t*>(base + o_nn), reinterpret_cast<int32_t*>(base + o_knn),
                        reinterpret_cast<double*>(base + o_md), st));
  ctx->stats.kernel_launches += 1;
  CU(cudaMemcpyAsync(nearest, base + o_nn, 4 * m, cudaMemcpyDeviceToHost, st));
  if (knn) CU(cudaMemcpyAsync(knn, base + o_knn, 4 * nk * m, cudaMemcpyDeviceToHost, st));
  if (min_mdist2) CU(cudaMemcpyAsync(min_mdist2, base + o_md, 8 * m, cudaMemcpyDeviceToHost, st));
  CU(cudaStreamSynchronize(st));
  return GMMLOC_OK;
}

gmmloc_status_t gmmloc_map_scan_bound(const gmmloc_map_t* map, double* alpha, double* xmax) {
  if (!map) return fail(GMMLOC_EINVAL, "null map");
  if (alpha) *alpha = map->d.alpha;
  if (xmax) *xmax = map->d.xmax;
  return GMMLOC_OK;
}

// ---- context -------------------------------------------------------------------------------------
gmmloc_status_t gmmloc_ctx_create(int device, gmmloc_ctx_t** out) {
  if (!out) return fail(GMMLOC_EINVAL, "null out");
  *out = nullptr;
  CU(cudaSetDevice(device));
  gmmloc_ctx* c = new (std::nothrow) gmmloc_ctx();
  if (!c) return fail(GMMLOC_ENOMEM, "host allocation failed");
  c->device = device;
  cudaError_t e = cudaSuccess;
  for (Lane& l : c->lanes) {
    if (e == cudaSuccess) e = cudaStreamCreateWithFlags(&l.stream, cudaStreamNonBlocking);
    if (e == cudaSuccess) e = cudaMallocHost((void**)&l.h_flag_count, 2 * sizeof(int));
    if (e == cudaSuccess) e = l.flag_count.reserve(2 * sizeof(int));
    if (e == cudaSuccess) e = l.next_frame.reserve(sizeof(int));
    if (e == cudaSuccess) e = cudaMemset(l.flag_count.p, 0, 2 * sizeof(int));
    for (int i = 0; i < 4 && e == cudaSuccess; ++i) e = cudaEventCreate(&l.ev[i]);
  }
  if (e == cudaSuccess) e = cudaHostAlloc((void**)&c->h_stop, sizeof(int32_t), cudaHostAllocMapped);
  if (e == cudaSuccess) {
    *c->h_stop = 0;
    e = cudaEventCreateWithFlags(&c->ev_ba, cudaEventDisableTiming);
  }
  if (e == cudaSuccess) e = cudaStreamCreateWithFlags(&c->h2d, cudaStreamNonBlocking);
  if (e == cudaSuccess) e = cudaStreamCreateWithFlags(&c->d2h, cudaStreamNonBlocking);
  if (e != cudaSuccess) {
    gmmloc_ctx_destroy(c);
    return fail_cuda(e, "gmmloc_ctx_create");
  }
  *out = c;
  return GMMLOC_OK;
}

void gmmloc_ctx_destroy(gmmloc_ctx_t* c) {
  if (!c) return;
  int prev = 0;
  cudaGetDevice(&prev);
  cudaSetDevice(c->device);
  for (Lane& l : c->lanes) {
    if (l.stream) cudaStreamSynchronize(l.stream);
    DevBuf* lb[] = {&l.flag_count, &l.flag_idx, &l.flag_thr, &l.xws, &l.fws, &l.next_frame, &l.sort_cnt, &l.sort_perm, &l.gn_store, &l.gn_prep};
    for (DevBuf* b : lb) b->release();
    if (l.h_flag_count) cudaFreeHost(l.h_flag_count);
    for (cudaEvent_t e : l.ev)
      if (e) cudaEventDestroy(e);
    if (l.stream) cudaStreamDestroy(l.stream);
  }
  DevBuf* bufs[] = {&c->in_pts, &c->in_obs, &c->in_oct, &c->in_pose, &c->in_misc, &c->out_assign, &c->out_d2, &c->out_pose,
                    &c->out_pts, &c->out_outlier, &c->out_status, &c->out_misc, &c->view_pose, &c->view_count, &c->view_status,
                    &c->view_ids, &c->view_rec, &c->view_wids, &c->view_wrec, &c->kf_obs, &c->kf_depth, &c->kf_oct, &c->kf_cand,
                    &c->kf_cand_pos, &c->kf_pt3d, &c->kf_assoc};
  for (DevBuf* b : bufs) b->release();
  for (cudaEvent_t e : c->ev_in) cudaEventDestroy(e);
  for (cudaEvent_t e : c->ev_done) cudaEventDestroy(e);
  if (c->h2d) cudaStreamDestroy(c->h2d);
  if (c->d2h) cudaStreamDestroy(c->d2h);
  if (c->h_stop) cudaFreeHost(c->h_stop);
  if (c->ev_ba) cudaEventDestroy(c->ev_ba);
  cudaSetDevice(prev);
  delete c;
}

void* gmmloc_ctx_stream(gmmloc_ctx_t* ctx) { return ctx ? (void*)ctx->lanes[0].stream : nullptr; }

gmmloc_status_t gmmloc_ctx_synchronize(gmmloc_ctx_t* ctx) {
  if (!ctx) return fail(GMMLOC_EINVAL, "null ctx");
  CU(cudaStreamSynchronize(ctx->L().stream));
  return GMMLOC_OK;
}

gmmloc_status_t gmmloc_ctx_stats(gmmloc_ctx_t* ctx, gmmloc_stats_t* out) {
  if (!ctx || !out) return fail(GMMLOC_EINVAL, "null pointer");
  // the flagged count of the last association is read back lazily (stream must be idle for it to be final)
  CU(cudaSetDevice(ctx->device));
  CU(cudaMemcpyAsync(ctx->L().h_flag_count, ctx->L().flag_count.p, 2 * sizeof(int), cudaMemcpyDeviceToHost, ctx->L().stream));
  CU(cudaStreamSynchronize(ctx->L().stream));
  ctx->stats.flagged = (int64_t)ctx->L().h_flag_count[0] + ctx->L().h_flag_count[1];  // grid re-rank list + brute-force list
  ctx->stats.flagged_brute = ctx->L().h_flag_count[1];
  *out = ctx->stats;
  return GMMLOC_OK;
}

gmmloc_status_t gmmloc_ctx_kernel_times(gmmloc_ctx_t* ctx, float ms[3]) {
  if (!ctx || !ms) return fail(GMMLOC_EINVAL, "null pointer");
  ms[0] = ms[1] = ms[2] = 0.f;
  CU(cudaSetDevice(ctx->device));
  CU(cudaStreamSynchronize(ctx->L().stream));
  if (ctx->L().ev_assoc) {
    CU(cudaEventElapsedTime(&ms[0], ctx->L().ev[0], ctx->L().ev[1]));
    CU(cudaEventElapsedTime(&ms[1], ctx->L().ev[1], ctx->L().ev[2]));
  }
  if (ctx->L().ev_gn) CU(cudaEventElapsedTime(&ms[2], ctx->L().ev[2], ctx->L().ev[3]));
  return GMMLOC_OK;
}

gmmloc_status_t gmmloc_ctx_set_assoc_kernel(gmmloc_ctx_t* ctx, int kind) {
  if (!ctx || kind < GMMLOC_ASSOC_AUTO || kind > GMMLOC_ASSOC_GRID) return fail(GMMLOC_EINVAL, "gmmloc_ctx_set_assoc_kernel: bad argument");
  ctx->assoc_kind = kind;
  return GMMLOC_OK;
}
int gmmloc_ctx_last_assoc_kernel(gmmloc_ctx_t* ctx) { return ctx ? ctx->last_assoc_tc : 0; }
int gmmloc_ctx_last_gn_kernel(gmmloc_ctx_t* ctx) { return ctx ? ctx->last_gn : -1; }

gmmloc_status_t gmmloc_ctx_set_gn_precision(gmmloc_ctx_t* ctx, int precision) {
  if (!ctx || (precision != GMMLOC_GN_FP64 && precision != GMMLOC_GN_MIXED && precision != GMMLOC_GN_WORLD &&
               precision != GMMLOC_GN_STREAM))
    return fail(GMMLOC_EINVAL, "gmmloc_ctx_set_gn_precision: bad argument");
  ctx->gn_precision = precision;
  return GMMLOC_OK;
}

gmmloc_status_t gmmloc_ctx_set_option(gmmloc_ctx_t* ctx, int option, int64_t value) {
  if (!ctx) return fail(GMMLOC_EINVAL, "gmmloc_ctx_set_option: null ctx");
  switch (option) {
    case GMMLOC_OPT_GN_PREP:
      if (value < -1 || value > 1) return fail(GMMLOC_EINVAL, "GMMLOC_OPT_GN_PREP: -1, 0 or 1");
      ctx->opt_gn_prep = (int)value;
      return GMMLOC_OK;
    case GMMLOC_OPT_E2E_CHUNK:
      if (value < 0 || value > (1 << 24)) return fail(GMMLOC_EINVAL, "GMMLOC_OPT_E2E_CHUNK: 0 .. 2^24");
      ctx->opt_e2e_chunk = (int)value;
      return GMMLOC_OK;
    case GMMLOC_OPT_E2E_SHORT_FIRST:
      ctx->opt_e2e_short_first = value ? 1 : 0;
      return GMMLOC_OK;
    case GMMLOC_OPT_E2E_TAPER:
      ctx->opt_e2e_taper = value ? 1 : 0;
      return GMMLOC_OK;
    case GMMLOC_OPT_BA_CLUSTER:
      if (value != 0 && value != 1 && value != 2 && value != 4 && value != 8) return fail(GMMLOC_EINVAL, "GMMLOC_OPT_BA_CLUSTER: 0, 1, 2, 4 or 8");
      ctx->opt_ba_cluster = (int)value;
      return GMMLOC_OK;
    default:
      return fail(GMMLOC_EINVAL, "gmmloc_ctx_set_option: unknown option");
  }
}

gmmloc_status_t gmmloc_host_alloc(size_t bytes, void** out) {
  if (!out) return fail(GMMLOC_EINVAL, "null out");
  *out = nullptr;
  CU(cudaMallocHost(out, bytes ? bytes : 1));
  return GMMLOC_OK;
}
void gmmloc_host_free(void* p) {
  if (p) cudaFreeHost(p);
}

// ---- association ---------------------------------------------------------------------------------
gmmloc_status_t gmmloc_associate_async(gmmloc_ctx_t* ctx, const gmmloc_map_t* map, const float* d_pts_w, int64_t M,
                                       double gate, int32_t* d_assign, float* d_d2) {
  if (!ctx || !map || !d_assign || M < 0 || (M > 0 && !d_pts_w)) return fail(GMMLOC_EINVAL, "gmmloc_associate: bad argument");
  if (M >= (int64_t)1 << 31) return fail(GMMLOC_EINVAL, "gmmloc_associate: M must be < 2^31 per call");
  if (map->device != ctx->device) return fail(GMMLOC_EINVAL, "map and context live on different devices");
  ctx->stats.points = M;
  if (M == 0) return GMMLOC_OK;
  CU(cudaSetDevice(ctx->device));
  CU(ctx->L().flag_idx.reserve(sizeof(int32_t) * 2 * (size_t)M));  // list A | list B (launch_associate)
  CU(ctx->L().flag_thr.reserve(sizeof(float) * 2 * (size_t)M));
  CU(cudaMemsetAsync(ctx->L().flag_count.p, 0, 2 * sizeof(int), ctx->L().stream));
  int* sort_cnt = nullptr;
  int32_t* sort_perm = nullptr;
  if (map->d.grid.nlev > 0) {  // scratch of the cell-sorted grid scan
    const GridLevel& G0 = map->d.grid.lev[0];
    CU(ctx->L().sort_cnt.reserve(sizeof(int) * ((size_t)G0.nx * G0.ny * G0.nz + 2 + ((size_t)G0.nx * G0.ny * G0.nz + 1) / 4096 + 4)));
    CU(ctx->L().sort_perm.reserve(16 * (size_t)M));  // one float4 per landmark
    sort_cnt = ctx->L().sort_cnt.as<int>();
    sort_perm = ctx->L().sort_perm.as<int32_t>();
  }
  int launches = 0;
  CU(cudaEventRecord(ctx->L().ev[0], ctx->L().stream));
  // kernel choice: variant 16 = tensor-core pruning scan, -1 = FFMA2 scan, 0 = launcher default
  const int variant = ctx->assoc_kind == GMMLOC_ASSOC_TENSOR ? 16
                      : ctx->assoc_kind == GMMLOC_ASSOC_FFMA2 ? -1
                      : ctx->assoc_kind == GMMLOC_ASSOC_GRID  ? 32
                                                              : 0;
  int used_tc = 0;
  CU(launch_associate(map->d, d_pts_w, M, gate, d_assign, d_d2, ctx->L().flag_count.as<int>(), ctx->L().flag_idx.as<int32_t>(),
                      ctx->L().flag_thr.as<float>(), variant, ctx->L().stream, &launches, ctx->L().ev[1], &used_tc, sort_cnt,
                      sort_perm));
  ctx->last_assoc_tc = used_tc;
  CU(cudaEventRecord(ctx->L().ev[2], ctx->L().stream));
  ctx->L().ev_assoc = true;
  ctx->L().ev_gn = false;
  ctx->stats.kernel_launches += launches;
  return GMMLOC_OK;
}

gmmloc_status_t gmmloc_associate(gmmloc_ctx_t* ctx, const gmmloc_map_t* map, const float* pts_w, int B, int N, double gate,
                                 int32_t* assign, float* d2) {
  if (!ctx || !map || !assign || B < 0 || N < 0) return fail(GMMLOC_EINVAL, "gmmloc_associate: bad argument");
  const int64_t M = (int64_t)B * N;
  if (M == 0) {
    ctx->stats.points = 0;
    return GMMLOC_OK;
  }
  if (!pts_w) return fail(GMMLOC_EINVAL, "gmmloc_associate: null landmarks");
  CU(cudaSetDevice(ctx->device));
  CU(ctx->in_pts.reserve(12 * (size_t)M));
  CU(ctx->out_assign.reserve(4 * (size_t)M));
  if (d2) CU(ctx->out_d2.reserve(4 * (size_t)M));
  CU(cudaMemcpyAsync(ctx->in_pts.p, pts_w, 12 * (size_t)M, cudaMemcpyHostToDevice, ctx->L().stream));
  gmmloc_status_t s = gmmloc_associate_async(ctx, map, ctx->in_pts.as<float>(), M, gate, ctx->out_assign.as<int32_t>(),
                                             d2 ? ctx->out_d2.as<float>() : nullptr);
  if (s != GMMLOC_OK) return s;
  CU(cudaMemcpyAsync(assign, ctx->out_assign.p, 4 * (size_t)M, cudaMemcpyDeviceToHost, ctx->L().stream));
  if (d2) CU(cudaMemcpyAsync(d2, ctx->out_d2.p, 4 * (size_t)M, cudaMemcpyDeviceToHost, ctx->L().stream));
  CU(cudaStreamSynchronize(ctx->L().stream));
  return GMMLOC_OK;
}

// ---- frame solve ---------------------------------------------------------------------------------
gmmloc_status_t gmmloc_frames_solve_async(gmmloc_ctx_t* ctx, const gmmloc_map_t* map, const gmmloc_camera_t* cam,
                                          const gmmloc_params_t* params, const double* d_pose0, const float* d_pts_w,
                                          const float* d_obs, const uint8_t* d_octave, int B, int N, double* d_pose_out,
                                          float* d_pts_out, int32_t* d_assign, uint8_t* d_outlier, int32_t* d_status) {
  if (!ctx || !map || !cam || !params || B < 0 || N < 0) return fail(GMMLOC_EINVAL, "gmmloc_frames_solve: bad argument");
  if (B == 0) return GMMLOC_OK;
  if (!d_pose0 || !d_pose_out || !d_status || (N > 0 && (!d_pts_w || !d_obs || !d_octave || !d_assign)))
    return fail(GMMLOC_EINVAL, "gmmloc_frames_solve: null pointer");
  if (params->gn_iters < 0 || params->factor_mode < 0 || params->factor_mode > 2)
    return fail(GMMLOC_EINVAL, "gmmloc_frames_solve: bad params");
  if (map->device != ctx->device) return fail(GMMLOC_EINVAL, "map and context live on different devices");
  CU(cudaSetDevice(ctx->device));
  const int64_t M = (int64_t)B * N;
  // association: ungated argmin (the 9.0 gate of gmmloc_opt.cpp:230 applies after refinement;
  // here it is the structure-edge gate evaluated inside the solve)
  if (params->factor_mode != GMMLOC_MODE_POSE_ONLY) {
    gmmloc_status_t s = gmmloc_associate_async(ctx, map, d_pts_w, M, 0.0, d_assign, nullptr);
    if (s != GMMLOC_OK) return s;
  } else {
    ctx->L().ev_assoc = false;
    if (M > 0) CU(cudaMemsetAsync(d_assign, 0xff, 4 * (size_t)M, ctx->L().stream));  // -1: no structure edges in pose-only mode
  }
  FrameArgs a;
  a.mean = map->d.mean;
  a.cov = map->d.cov;
  a.cov_inv = map->d.cov_inv;
  a.axis = map->d.axis;
  a.is_deg = map->d.is_deg;
  fill_params(cam, params, a.prm);
  a.pose0 = d_pose0;
  a.pts_w = d_pts_w;
  a.obs = d_obs;
  a.octave = d_octave;
  a.assign = d_assign;
  a.B = B;
  a.N = N;
  a.pose_out = d_pose_out;
  a.pts_out = d_pts_out;
  a.outlier = d_outlier;
  a.status = d_status;
  a.xws = nullptr;
  a.fws = nullptr;
  if ((size_t)N > frame_gn_smem_limit_points()) {
    CU(ctx->L().xws.reserve(24 * (size_t)M));
    CU(ctx->L().fws.reserve((size_t)M));
    a.xws = ctx->L().xws.as<double>();
    a.fws = ctx->L().fws.as<uint8_t>();
  }
  if (!ctx->L().ev_assoc) CU(cudaEventRecord(ctx->L().ev[2], ctx->L().stream));
  static const int gn_forced = [] {  // GMMLOC_GN_PRECISION=fp64|mixed overrides the context setting (A/B runs)
    const char* v = debug_env("GMMLOC_GN_PRECISION");
    if (v == nullptr) return -1;
    if (std::strcmp(v, "fp64") == 0) return (int)GMMLOC_GN_FP64;
    if (std::strcmp(v, "mixed") == 0) return (int)GMMLOC_GN_MIXED;
    if (std::strcmp(v, "world") == 0) return (int)GMMLOC_GN_WORLD;
    if (std::strcmp(v, "stream") == 0) return (int)GMMLOC_GN_STREAM;
    return -1;
  }();
  const int gn_prec = gn_forced >= 0 ? gn_forced : ctx->gn_precision;
  if (gn_prec == GMMLOC_GN_STREAM && map->d.gn_tab != nullptr && N > 0 && (size_t)N <= frame_gn_stream_limit_points()) {
    CU(ctx->L().next_frame.reserve(sizeof(int) * ((size_t)B + 2)));
    CU(ctx->L().gn_store.reserve(frame_gn_stream_scratch_bytes(N)));
    CU(ctx->L().gn_prep.reserve(frame_gn_stream_prep_bytes(B, N)));
    CU(launch_frame_gn_stream(a, map->d.gn_tab, ctx->L().next_frame.as<int>(), ctx->L().gn_store.as<float>(), ctx->L().gn_store.cap,
                              ctx->L().gn_prep.as<unsigned char>(), ctx->L().gn_prep.cap, ctx->L().stream));
    ctx->stats.kernel_launches += 2;  // + the pack kernel and the (normally empty) fp64 retry launch
    ctx->last_gn = GMMLOC_GN_STREAM;
  } else if (gn_prec == GMMLOC_GN_WORLD && map->d.gn_tab != nullptr && N > 0 && (size_t)N <= frame_gn_world_limit_points()) {
    CU(ctx->L().next_frame.reserve(sizeof(int)));
    CU(ctx->L().gn_store.reserve(frame_gn_world_scratch_bytes(N)));
    CU(ctx->L().gn_prep.reserve(frame_gn_world_prep_bytes(B, N)));
    CU(launch_frame_gn_world(a, map->d.gn_tab, ctx->L().next_frame.as<int>(), ctx->L().gn_store.as<float>(), ctx->L().gn_store.cap,
                             ctx->L().gn_prep.as<unsigned char>(), ctx->L().gn_prep.cap, ctx->L().stream));
    ctx->last_gn = GMMLOC_GN_WORLD;
    ctx->stats.kernel_launches += 1;
  } else if (gn_prec != GMMLOC_GN_FP64 && map->d.gn_tab != nullptr && (size_t)N <= frame_gn_mixed_limit_points()) {
    CU(ctx->L().next_frame.reserve(sizeof(int)));
    CU(ctx->L().gn_store.reserve(frame_gn_mixed_scratch_bytes(N)));
    const size_t prep_bytes = frame_gn_mixed_prep_bytes(B, N, ctx->opt_gn_prep);
    if (prep_bytes) CU(ctx->L().gn_prep.reserve(prep_bytes));
    CU(launch_frame_gn_mixed(a, map->d.gn_tab, ctx->L().next_frame.as<int>(), ctx->L().gn_store.as<float>(), ctx->L().gn_store.cap,
                             prep_bytes ? ctx->L().gn_prep.as<unsigned char>() : nullptr, prep_bytes ? ctx->L().gn_prep.cap : 0,
                             ctx->L().stream));
    ctx->last_gn = GMMLOC_GN_MIXED;
    if (prep_bytes) ctx->stats.kernel_launches += 1;
  } else {
    CU(launch_frame_gn(a, ctx->L().stream));
    ctx->last_gn = GMMLOC_GN_FP64;
  }
  CU(cudaEventRecord(ctx->L().ev[3], ctx->L().stream));
  ctx->L().ev_gn = true;
  ctx->stats.kernel_launches += 1;
  return GMMLOC_OK;
}

// Chunk boundaries of the host pipeline of gmmloc_frames_solve: a pure function of the batch size, the SM count and the
// context options (exported as gmmloc_frames_solve_chunks so that the schedule is testable without a device).
// chunk_opt: GMMLOC_OPT_E2E_CHUNK (0 = automatic), taper: GMMLOC_OPT_E2E_TAPER, short_first: GMMLOC_OPT_E2E_SHORT_FIRST.
static std::vector<int> frames_solve_chunk_starts(int B, int sms, int chunk_env, int taper, int short_first) {
  // Chunk = 3 frames per resident CTA of the persistent frame-GN kernel (2 CTAs x 148 SMs -> 888 frames on B200): a
  // chunk that is not a multiple of the resident CTA count leaves most CTAs idle for a whole extra frame.  Measured on
  // the C2 frame size (tools/probe_e2e.py): 256 -> 463 k, 512 -> 537 k, 888 -> 574 k, 1024 -> 548 k, 2048 -> 506 k
  // frames/s; PCIe moves 55 GB/s each way on this box.
  const int chunk = chunk_env > 0 ? std::min(chunk_env, B) : ((B >= 1024) ? 6 * sms : B);
  // chunk boundaries: uniform chunks; a remainder smaller than half a chunk is merged into the last chunk instead of
  // running as a tiny one.  GMMLOC_E2E_SHORT_FIRST=1 makes the first chunk a third of the others (kernels start after a
  // shorter H2D) — measured no faster (652 k vs 658 k frames/s on C2), so it is off by default.
  std::vector<int> cstart;
  if (taper && !short_first && B >= 6 * 2 * sms + 2 * chunk) {
    // Graded schedule (default for large batches).  The pipeline is bound by the kernels, not by the copy-in: a chunk of n
    // frames costs ~0.22 ms + 0.84 us n on the device (the sort + grid scan loses its cell sharing on small launches, the
    // persistent solver runs whole waves of 2 x SMs frames) against 0.99 us n of copy-in at 50 GB/s, so chunk i + 1 has
    // arrived long before chunk i is solved whenever it is not much larger.  The chunks therefore start at one frame per
    // resident solver CTA (the kernels start after 0.3 ms of copy-in) and GROW by half a wave each — fewer chunks, less fixed
    // cost — and the last one takes the rest.  Measured on C2 (4 096 frames, tools/gpu_sched.sh): 296, 444, 592, 740, 888, 1 136
    // -> 775 k frames/s; the round-1 taper 296, 592, 544, 888, 888, 592, 296 -> 765 k; 296, 592, 1 184, 2 024 -> 667 k.
    const int unit = 2 * sms;
    std::vector<int> sizes;
    int left = B, next = unit;
    while (left > 0) {
      if (left < 2 * next + sms) {  // the rest would be a chunk smaller than this one: merge
        sizes.push_back(left);
        break;
      }
      sizes.push_back(next);
      left -= next;
      // growth capped at 12 frames per SM (GMMLOC_OPT_E2E_CHUNK: at twice that option): beyond ~1 800 frames a chunk's copy-in
      // takes longer than the kernels of the chunk before it and the pipeline would wait on PCIe with a long last chunk
      next = std::min(next + sms, std::max(2 * chunk, unit));
    }
    int f = 0;
    for (int sz : sizes) {
      cstart.push_back(f);
      f += sz;
    }
    cstart.push_back(B);
  } else {
    int f = 0;
    const int first = (short_first && chunk_env <= 0 && B >= 4 * chunk) ? chunk / 3 : chunk;
    while (f < B) {
      cstart.push_back(f);
      f += (f == 0) ? first : chunk;
    }
    if (cstart.size() >= 2 && B - cstart.back() < chunk / 2) cstart.pop_back();
    cstart.push_back(B);
  }
  return cstart;
}

int32_t gmmloc_frames_solve_chunks(int32_t B, int32_t sms, int32_t chunk_opt, int32_t taper, int32_t short_first, int32_t* sizes,
                                   int32_t cap) try {
  if (B <= 0 || sms <= 0 || cap < 0 || (cap > 0 && sizes == nullptr)) return -1;
  const std::vector<int> cs = frames_solve_chunk_starts(B, sms, chunk_opt, taper, short_first);
  const int n = (int)cs.size() - 1;
  for (int i = 0; i < n && i < cap; ++i) sizes[i] = cs[i + 1] - cs[i];
  return n;
} catch (...) {
  return -1;
}

gmmloc_status_t gmmloc_frames_solve(gmmloc_ctx_t* ctx, const gmmloc_map_t* map, const gmmloc_camera_t* cam,
                                    const gmmloc_params_t* params, const double* pose0, const float* pts_w,
                                    const float* obs, const uint8_t* octave, int B, int N, double* pose_out, float* pts_out,
                                    int32_t* assign, uint8_t* outlier, int32_t* status) try {
  if (!ctx || !map || !cam || !params || B < 0 || N < 0) return fail(GMMLOC_EINVAL, "gmmloc_frames_solve: bad argument");
  if (B == 0) return GMMLOC_OK;
  if (!pose0 || !pose_out || !status || (N > 0 && (!pts_w || !obs || !octave)))
    return fail(GMMLOC_EINVAL, "gmmloc_frames_solve: null pointer");
  CU(cudaSetDevice(ctx->device));
  const size_t M = (size_t)B * N;
  CU(ctx->in_pose.reserve(56 * (size_t)B));
  CU(ctx->in_pts.reserve(12 * M + 16));
  CU(ctx->in_obs.reserve(12 * M + 16));
  CU(ctx->in_oct.reserve(M + 16));
  CU(ctx->out_pose.reserve(56 * (size_t)B));
  CU(ctx->out_assign.reserve(4 * M + 16));
  CU(ctx->out_status.reserve(4 * (size_t)B));
  if (pts_out) CU(ctx->out_pts.reserve(12 * M + 16));
  if (outlier) CU(ctx->out_outlier.reserve(M + 16));
  // Pipelined over chunks of frames: H2D of chunk c+1 (copy stream) overlaps the kernels of chunk c
  // (compute stream) and the D2H of chunk c-1 (second copy stream); PCIe is full duplex.  With
  // page-locked caller buffers the copies are truly asynchronous; pageable buffers still work.
  const int chunk_env = [&] {
    const char* v = debug_env("GMMLOC_E2E_CHUNK");
    return v ? std::atoi(v) : ctx->opt_e2e_chunk;
  }();
  // (chunk boundaries: frames_solve_chunk_starts above)
  int sms = 148;
  cudaDeviceGetAttribute(&sms, cudaDevAttrMultiProcessorCount, ctx->device);
  const int short_first = [&] {
    const char* v = debug_env("GMMLOC_E2E_SHORT_FIRST");
    return v ? std::atoi(v) : ctx->opt_e2e_short_first;
  }();
  std::vector<int> cstart;
  // -DGMMLOC_DEBUG_ENV builds only: GMMLOC_E2E_SCHEDULE="296,444,592,..." sets the chunk sizes (the last one takes the rest)
  if (const char* sch = debug_env("GMMLOC_E2E_SCHEDULE")) {
    int f = 0;
    for (const char* q = sch; *q && f < B;) {
      const int sz = std::atoi(q);
      if (sz <= 0) break;
      cstart.push_back(f);
      f += sz;
      while (*q && *q != ',') ++q;
      if (*q == ',') ++q;
    }
    if (cstart.empty()) cstart.push_back(0);
    cstart.push_back(B);
  } else {
    cstart = frames_solve_chunk_starts(B, sms, chunk_env, ctx->opt_e2e_taper, short_first);
  }
  const int nchunks = (int)cstart.size() - 1;
  while ((int)ctx->ev_in.size() < nchunks) {
    cudaEvent_t a = nullptr, b = nullptr;
    CU(cudaEventCreateWithFlags(&a, cudaEventDisableTiming));
    ctx->ev_in.push_back(a);
    CU(cudaEventCreateWithFlags(&b, cudaEventDisableTiming));
    ctx->ev_done.push_back(b);
  }
  // work lists of both lanes sized up front (a growing reserve() frees, which would synchronise mid-pipeline)
  for (int l = 0; l < (nchunks > 1 ? 2 : 1); ++l) {
    Lane& ln = ctx->lanes[l];
    int maxc = 0;
    for (size_t q = 0; q + 1 < cstart.size(); ++q) maxc = std::max(maxc, cstart[q + 1] - cstart[q]);
    const size_t mc = (size_t)maxc * N;
    CU(ln.flag_idx.reserve(sizeof(int32_t) * 2 * mc));
    CU(ln.flag_thr.reserve(sizeof(float) * 2 * mc));
    CU(ln.next_frame.reserve(sizeof(int) * (maxc + 2)));
    CU(ln.gn_store.reserve(std::max(std::max(frame_gn_mixed_scratch_bytes(N), frame_gn_world_scratch_bytes(N)),
                                    frame_gn_stream_scratch_bytes(N))));
    if (ctx->gn_precision == GMMLOC_GN_WORLD) CU(ln.gn_prep.reserve(frame_gn_world_prep_bytes(maxc, N)));
    if (ctx->gn_precision == GMMLOC_GN_STREAM && (size_t)N <= frame_gn_stream_limit_points())
      CU(ln.gn_prep.reserve(frame_gn_stream_prep_bytes(maxc, N)));
    if (frame_gn_mixed_prep_bytes(maxc, N, ctx->opt_gn_prep)) CU(ln.gn_prep.reserve(frame_gn_mixed_prep_bytes(maxc, N, ctx->opt_gn_prep)));
    if (map->d.grid.nlev > 0) {
      const GridLevel& G0 = map->d.grid.lev[0];
      CU(ln.sort_cnt.reserve(sizeof(int) * ((size_t)G0.nx * G0.ny * G0.nz + 2 + ((size_t)G0.nx * G0.ny * G0.nz + 1) / 4096 + 4)));
      CU(ln.sort_perm.reserve(16 * mc));
    }
    if ((size_t)N > frame_gn_smem_limit_points()) {
      CU(ln.xws.reserve(24 * mc));
      CU(ln.fws.reserve(mc));
    }
  }
  gmmloc_status_t rc = GMMLOC_OK;
  // an error inside the loop must not return while copies into the caller's buffers are in flight: CUB records the
  // status and leaves the loop; every exit path then drains the four streams below and resets ctx->active
#define CUB(call)                                       \
  {                                                     \
    const cudaError_t e__ = (call);                     \
    if (e__ != cudaSuccess) {                           \
      rc = fail_cuda(e__, #call);                       \
      break;                                            \
    }                                                   \
  }
  // -DGMMLOC_DEBUG_ENV builds only: GMMLOC_E2E_TRACE=1 prints the timeline of the pipeline (ms since the first copy was issued)
  const bool trace = debug_env("GMMLOC_E2E_TRACE") != nullptr;
  std::vector<cudaEvent_t> tev;  // per chunk: copy-in done, kernels started, kernels done, copy-out done; [4 n]: start
  if (trace) {
    tev.resize(4 * (size_t)nchunks + 1);
    for (auto& e : tev) cudaEventCreate(&e);
    cudaEventRecord(tev[4 * (size_t)nchunks], ctx->h2d);
  }
  for (int c = 0; c < nchunks && rc == GMMLOC_OK; ++c) {
    ctx->active = (nchunks > 1) ? (c & 1) : 0;
    cudaStream_t st = ctx->L().stream;
    const size_t f0 = (size_t)cstart[c], nb = (size_t)(cstart[c + 1] - cstart[c]), m0 = f0 * N, nm = nb * N;
    CUB(cudaMemcpyAsync(ctx->in_pose.as<double>() + 7 * f0, pose0 + 7 * f0, 56 * nb, cudaMemcpyHostToDevice, ctx->h2d));
    if (nm) {
      CUB(cudaMemcpyAsync(ctx->in_pts.as<float>() + 3 * m0, pts_w + 3 * m0, 12 * nm, cudaMemcpyHostToDevice, ctx->h2d));
      CUB(cudaMemcpyAsync(ctx->in_obs.as<float>() + 3 * m0, obs + 3 * m0, 12 * nm, cudaMemcpyHostToDevice, ctx->h2d));
      CUB(cudaMemcpyAsync(ctx->in_oct.as<uint8_t>() + m0, octave + m0, nm, cudaMemcpyHostToDevice, ctx->h2d));
    }
    CUB(cudaEventRecord(ctx->ev_in[c], ctx->h2d));
    if (trace) cudaEventRecord(tev[4 * c], ctx->h2d);
    CUB(cudaStreamWaitEvent(st, ctx->ev_in[c], 0));
    if (trace) cudaEventRecord(tev[4 * c + 1], st);
    rc = gmmloc_frames_solve_async(
        ctx, map, cam, params, ctx->in_pose.as<double>() + 7 * f0, ctx->in_pts.as<float>() + 3 * m0,
        ctx->in_obs.as<float>() + 3 * m0, ctx->in_oct.as<uint8_t>() + m0, (int)nb, N, ctx->out_pose.as<double>() + 7 * f0,
        pts_out ? ctx->out_pts.as<float>() + 3 * m0 : nullptr, ctx->out_assign.as<int32_t>() + m0,
        outlier ? ctx->out_outlier.as<uint8_t>() + m0 : nullptr, ctx->out_status.as<int32_t>() + f0);
    if (rc != GMMLOC_OK) break;
    CUB(cudaEventRecord(ctx->ev_done[c], st));
    if (trace) cudaEventRecord(tev[4 * c + 2], st);
    CUB(cudaStreamWaitEvent(ctx->d2h, ctx->ev_done[c], 0));
    CUB(cudaMemcpyAsync(pose_out + 7 * f0, ctx->out_pose.as<double>() + 7 * f0, 56 * nb, cudaMemcpyDeviceToHost, ctx->d2h));
    CUB(cudaMemcpyAsync(status + f0, ctx->out_status.as<int32_t>() + f0, 4 * nb, cudaMemcpyDeviceToHost, ctx->d2h));
    if (nm) {
      if (assign) CUB(cudaMemcpyAsync(assign + m0, ctx->out_assign.as<int32_t>() + m0, 4 * nm, cudaMemcpyDeviceToHost, ctx->d2h));
      if (pts_out) CUB(cudaMemcpyAsync(pts_out + 3 * m0, ctx->out_pts.as<float>() + 3 * m0, 12 * nm, cudaMemcpyDeviceToHost, ctx->d2h));
      if (outlier) CUB(cudaMemcpyAsync(outlier + m0, ctx->out_outlier.as<uint8_t>() + m0, nm, cudaMemcpyDeviceToHost, ctx->d2h));
    }
    if (trace) cudaEventRecord(tev[4 * c + 3], ctx->d2h);
  }
  ctx->active = 0;
  const cudaError_t e0 = cudaStreamSynchronize(ctx->h2d), e1 = cudaStreamSynchronize(ctx->lanes[0].stream),
                    e2 = cudaStreamSynchronize(ctx->lanes[1].stream), e3 = cudaStreamSynchronize(ctx->d2h);
  if (trace) {
    for (int c = 0; c < nchunks; ++c) {
      float t[4] = {0, 0, 0, 0};
      for (int i = 0; i < 4; ++i) cudaEventElapsedTime(&t[i], tev[4 * (size_t)nchunks], tev[4 * c + i]);
      std::fprintf(stderr, "chunk %d: %d frames | in done %.3f | kernels %.3f .. %.3f | out done %.3f ms\n", c, cstart[c + 1] - cstart[c],
                   t[0], t[1], t[2], t[3]);
    }
    for (auto& e : tev) cudaEventDestroy(e);
  }
#undef CUB
  if (rc != GMMLOC_OK) return rc;
  CU(e0);
  CU(e1);
  CU(e2);
  CU(e3);
  return GMMLOC_OK;
} GMMLOC_ABI_CATCH

// ---- pose-only LM (Tracking::optimizeCurrentPose) --------------------------------------------------
gmmloc_status_t gmmloc_poses_optimize(gmmloc_ctx_t* ctx, const gmmloc_camera_t* cam, const double* pose0, const float* pts_w,
                                      const float* obs, const uint8_t* octave, const uint8_t* valid, int B, int N,
                                      double* pose_out, uint8_t* outlier, int32_t* inliers) {
  if (!ctx || !cam || B < 0 || N < 0) return fail(GMMLOC_EINVAL, "gmmloc_poses_optimize: bad argument");
  if (B == 0) return GMMLOC_OK;
  if (!pose0 || !pose_out || !inliers || (N > 0 && (!pts_w || !obs || !octave)))
    return fail(GMMLOC_EINVAL, "gmmloc_poses_optimize: null pointer");
  CU(cudaSetDevice(ctx->device));
  ctx->active = 0;
  const size_t M = (size_t)B * N;
  CU(ctx->in_pose.reserve(56 * (size_t)B));
  CU(ctx->in_pts.reserve(12 * M + 16));
  CU(ctx->in_obs.reserve(12 * M + 16));
  CU(ctx->in_oct.reserve(M + 16));
  if (valid) CU(ctx->in_misc.reserve(M + 16));
  CU(ctx->out_pose.reserve(56 * (size_t)B));
  CU(ctx->out_status.reserve(4 * (size_t)B));
  if (outlier) CU(ctx->out_outlier.reserve(M + 16));
  cudaStream_t st = ctx->L().stream;
  CU(cudaMemcpyAsync(ctx->in_pose.p, pose0, 56 * (size_t)B, cudaMemcpyHostToDevice, st));
  if (M) {
    CU(cudaMemcpyAsync(ctx->in_pts.p, pts_w, 12 * M, cudaMemcpyHostToDevice, st));
    CU(cudaMemcpyAsync(ctx->in_obs.p, obs, 12 * M, cudaMemcpyHostToDevice, st));
    CU(cudaMemcpyAsync(ctx->in_oct.p, octave, M, cudaMemcpyHostToDevice, st));
    if (valid) CU(cudaMemcpyAsync(ctx->in_misc.p, valid, M, cudaMemcpyHostToDevice, st));
  }
  PoseLmArgs a;
  gmmloc_params_t prm;
  gmmloc_default_params(&prm);
  fill_params(cam, &prm, a.prm);
  a.pose0 = ctx->in_pose.as<double>();
  a.pts_w = ctx->in_pts.as<float>();
  a.obs = ctx->in_obs.as<float>();
  a.octave = ctx->in_oct.as<uint8_t>();
  a.valid = valid ? ctx->in_misc.as<uint8_t>() : nullptr;
  a.B = B;
  a.N = N;
  a.pose_out = ctx->out_pose.as<double>();
  a.outlier = outlier ? ctx->out_outlier.as<uint8_t>() : nullptr;
  a.inliers = ctx->out_status.as<int32_t>();
  a.ws_c2 = nullptr;
  a.ws_st = nullptr;
  if ((size_t)N > pose_lm_smem_limit_points()) {
    CU(ctx->L().xws.reserve(8 * M));
    CU(ctx->L().fws.reserve(M));
    a.ws_c2 = ctx->L().xws.as<double>();
    a.ws_st = ctx->L().fws.as<uint8_t>();
  }
  CU(launch_pose_lm(a, st));
  ctx->stats.kernel_launches += 1;
  CU(cudaMemcpyAsync(pose_out, ctx->out_pose.p, 56 * (size_t)B, cudaMemcpyDeviceToHost, st));
  CU(cudaMemcpyAsync(inliers, ctx->out_status.p, 4 * (size_t)B, cudaMemcpyDeviceToHost, st));
  if (outlier && M) CU(cudaMemcpyAsync(outlier, ctx->out_outlier.p, M, cudaMemcpyDeviceToHost, st));
  CU(cudaStreamSynchronize(st));
  return GMMLOC_OK;
}

// ---- batched point solves ------------------------------------------------------------------------
gmmloc_status_t gmmloc_points_solve(gmmloc_ctx_t* ctx, const gmmloc_map_t* map, const gmmloc_camera_t* cam,
                                    const gmmloc_params_t* params, const double* x0, const float* obs,
                                    const uint8_t* octave, const double* pose, const int32_t* comp, const double* proj_z2,
                                    int M, int iters, double* pt_est, double* chi2_proj, double* chi2_str, uint8_t* res) {
  if (!ctx || !map || !cam || !params || M < 0 || iters < 0) return fail(GMMLOC_EINVAL, "gmmloc_points_solve: bad argument");
  if (M == 0) return GMMLOC_OK;
  if (!x0 || !obs || !octave || !pose || !comp || !proj_z2 || !pt_est || !chi2_proj || !chi2_str || !res)
    return fail(GMMLOC_EINVAL, "gmmloc_points_solve: null pointer");
  if (map->device != ctx->device) return fail(GMMLOC_EINVAL, "map and context live on different devices");
  CU(cudaSetDevice(ctx->device));
  const size_t m = (size_t)M;
  // one input slab: x0 (24) | pose (56) | proj_z2 (8) | obs (12) | comp (4) | octave (1)
  const size_t o_x0 = 0, o_pose = o_x0 + 24 * m, o_pz = o_pose + 56 * m, o_obs = o_pz + 8 * m, o_comp = o_obs + 12 * m,
               o_oct = o_comp + 4 * m, in_bytes = o_oct + m;
  // one output slab: pt_est (24) | chi2_proj (8) | chi2_str (8) | res (1)
  const size_t q_pt = 0, q_cp = 24 * m, q_cs = q_cp + 8 * m, q_res = q_cs + 8 * m, out_bytes = q_res + m;
  CU(ctx->in_misc.reserve(in_bytes));
  CU(ctx->out_misc.reserve(out_bytes));
  char* in = ctx->in_misc.as<char>();
  char* outp = ctx->out_misc.as<char>();
  cudaStream_t st = ctx->L().stream;
  CU(cudaMemcpyAsync(in + o_x0, x0, 24 * m, cudaMemcpyHostToDevice, st));
  CU(cudaMemcpyAsync(in + o_pose, pose, 56 * m, cudaMemcpyHostToDevice, st));
  CU(cudaMemcpyAsync(in + o_pz, proj_z2, 8 * m, cudaMemcpyHostToDevice, st));
  CU(cudaMemcpyAsync(in + o_obs, obs, 12 * m, cudaMemcpyHostToDevice, st));
  CU(cudaMemcpyAsync(in + o_comp, comp, 4 * m, cudaMemcpyHostToDevice, st));
  CU(cudaMemcpyAsync(in + o_oct, octave, m, cudaMemcpyHostToDevice, st));
  PointArgs a;
  a.mean = map->d.mean;
  a.axis = map->d.axis;
  fill_params(cam, params, a.prm);
  a.x0 = reinterpret_cast<const double*>(in + o_x0);
  a.pose = reinterpret_cast<const double*>(in + o_pose);
  a.proj_z2 = reinterpret_cast<const double*>(in + o_pz);
  a.obs = reinterpret_cast<const float*>(in + o_obs);
  a.comp = reinterpret_cast<const int32_t*>(in + o_comp);
  a.octave = reinterpret_cast<const uint8_t*>(in + o_oct);
  a.M = M;
  a.iters = iters;
  a.K = map->d.K;
  a.pt_est = reinterpret_cast<double*>(outp + q_pt);
  a.chi2_proj = reinterpret_cast<double*>(outp + q_cp);
  a.chi2_str = reinterpret_cast<double*>(outp + q_cs);
  a.res = reinterpret_cast<uint8_t*>(outp + q_res);
  CU(launch_point_gn(a, st));
  ctx->stats.kernel_launches += 1;
  CU(cudaMemcpyAsync(pt_est, outp + q_pt, 24 * m, cudaMemcpyDeviceToHost, st));
  CU(cudaMemcpyAsync(chi2_proj, outp + q_cp, 8 * m, cudaMemcpyDeviceToHost, st));
  CU(cudaMemcpyAsync(chi2_str, outp + q_cs, 8 * m, cudaMemcpyDeviceToHost, st));
  CU(cudaMemcpyAsync(res, outp + q_res, m, cudaMemcpyDeviceToHost, st));
  CU(cudaStreamSynchronize(st));
  return GMMLOC_OK;
}

// ---- batched triangulation solves -----------------------------------------------------------------
gmmloc_status_t gmmloc_triangulation_solve(gmmloc_ctx_t* ctx, const gmmloc_map_t* map, const gmmloc_camera_t* cam,
                                           const gmmloc_params_t* params, const double* x0, const float* obs1, const float* obs2,
                                           const uint8_t* octave1, const double* pose1, const double* pose2, const int32_t* comp,
                                           int M, int iters, double* pt_est, double* err1, double* err2, double* chi2_str,
                                           uint8_t* res) {
  if (!ctx || !map || !cam || !params || M < 0 || iters < 0) return fail(GMMLOC_EINVAL, "gmmloc_triangulation_solve: bad argument");
  if (M == 0) return GMMLOC_OK;
  if (!x0 || !obs1 || !obs2 || !octave1 || !pose1 || !pose2 || !comp || !pt_est || !err1 || !err2 || !chi2_str || !res)
    return fail(GMMLOC_EINVAL, "gmmloc_triangulation_solve: null pointer");
  if (map->device != ctx->device) return fail(GMMLOC_EINVAL, "map and context live on different devices");
  CU(cudaSetDevice(ctx->device));
  ctx->active = 0;
  const size_t m = (size_t)M;
  // input slab: x0 (24) | pose1 (56) | pose2 (56) | obs1 (12) | obs2 (12) | comp (4) | octave1 (1)
  const size_t o_x0 = 0, o_p1 = 24 * m, o_p2 = o_p1 + 56 * m, o_o1 = o_p2 + 56 * m, o_o2 = o_o1 + 12 * m, o_c = o_o2 + 12 * m,
               o_oc = o_c + 4 * m, in_bytes = o_oc + m;
  // output slab: pt_est (24) | err1 (8) | err2 (8) | chi2_str (8) | res (1)
  const size_t q_pt = 0, q_e1 = 24 * m, q_e2 = q_e1 + 8 * m, q_cs = q_e2 + 8 * m, q_r = q_cs + 8 * m, out_bytes = q_r + m;
  CU(ctx->in_misc.reserve(in_bytes));
  CU(ctx->out_misc.reserve(out_bytes));
  char* in = ctx->in_misc.as<char>();
  char* outp = ctx->out_misc.as<char>();
  cudaStream_t st = ctx->L().stream;
  CU(cudaMemcpyAsync(in + o_x0, x0, 24 * m, cudaMemcpyHostToDevice, st));
  CU(cudaMemcpyAsync(in + o_p1, pose1, 56 * m, cudaMemcpyHostToDevice, st));
  CU(cudaMemcpyAsync(in + o_p2, pose2, 56 * m, cudaMemcpyHostToDevice, st));
  CU(cudaMemcpyAsync(in + o_o1, obs1, 12 * m, cudaMemcpyHostToDevice, st));
  CU(cudaMemcpyAsync(in + o_o2, obs2, 12 * m, cudaMemcpyHostToDevice, st));
  CU(cudaMemcpyAsync(in + o_c, comp, 4 * m, cudaMemcpyHostToDevice, st));
  CU(cudaMemcpyAsync(in + o_oc, octave1, m, cudaMemcpyHostToDevice, st));
  TriArgs a;
  a.mean = map->d.mean;
  a.axis = map->d.axis;
  fill_params(cam, params, a.prm);
  a.x0 = reinterpret_cast<const double*>(in + o_x0);
  a.pose1 = reinterpret_cast<const double*>(in + o_p1);
  a.pose2 = reinterpret_cast<const double*>(in + o_p2);
  a.obs1 = reinterpret_cast<const float*>(in + o_o1);
  a.obs2 = reinterpret_cast<const float*>(in + o_o2);
  a.comp = reinterpret_cast<const int32_t*>(in + o_c);
  a.octave1 = reinterpret_cast<const uint8_t*>(in + o_oc);
  a.M = M;
  a.iters = iters;
  a.K = map->d.K;
  a.pt_est = reinterpret_cast<double*>(outp + q_pt);
  a.err1 = reinterpret_cast<double*>(outp + q_e1);
  a.err2 = reinterpret_cast<double*>(outp + q_e2);
  a.chi2_str = reinterpret_cast<double*>(outp + q_cs);
  a.res = reinterpret_cast<uint8_t*>(outp + q_r);
  CU(launch_tri_gn(a, st));
  ctx->stats.kernel_launches += 1;
  CU(cudaMemcpyAsync(pt_est, outp + q_pt, 24 * m, cudaMemcpyDeviceToHost, st));
  CU(cudaMemcpyAsync(err1, outp + q_e1, 8 * m, cudaMemcpyDeviceToHost, st));
  CU(cudaMemcpyAsync(err2, outp + q_e2, 8 * m, cudaMemcpyDeviceToHost, st));
  CU(cudaMemcpyAsync(chi2_str, outp + q_cs, 8 * m, cudaMemcpyDeviceToHost, st));
  CU(cudaMemcpyAsync(res, outp + q_r, m, cudaMemcpyDeviceToHost, st));
  CU(cudaStreamSynchronize(st));
  return GMMLOC_OK;
}

// ---- keyframe association pipeline ------------------------------------------------------------------
namespace {

// enqueue K6 on the context's stream for poses already on the device (ctx->view_pose)
gmmloc_status_t render_enqueue(gmmloc_ctx_t* ctx, const gmmloc_map_t* map, const gmmloc_camera_t* cam, int B, int cap) {
  const size_t n = (size_t)B * cap;
  CU(ctx->view_count.reserve(4 * (size_t)B));
  CU(ctx->view_status.reserve(4 * (size_t)B));
  CU(ctx->view_ids.reserve(4 * n));
  CU(ctx->view_wids.reserve(4 * n));
  CU(ctx->view_rec.reserve(8 * kViewRec * n));
  CU(ctx->view_wrec.reserve(8 * kViewRec * n));
  ViewArgs a;
  a.mean = map->d.mean;
  a.cov = map->d.cov;
  a.axis = map->d.axis;
  a.is_deg = map->d.is_deg;
  a.K = map->d.K;
  a.fx = cam->fx; a.fy = cam->fy; a.cx = cam->cx; a.cy = cam->cy;
  a.width = (double)cam->width;
  a.height = (double)cam->height;
  a.view_cos_thresh = std::cos(78.0 * M_PI / 180.0);  // gaussian_mixture.cpp:283
  a.pose = ctx->view_pose.as<double>();
  a.B = B;
  a.cap = cap;
  a.count = ctx->view_count.as<int32_t>();
  a.ids = ctx->view_ids.as<int32_t>();
  a.rec = ctx->view_rec.as<double>();
  a.wids = ctx->view_wids.as<int32_t>();
  a.wrec = ctx->view_wrec.as<double>();
  a.status = ctx->view_status.as<int32_t>();
  CU(launch_render_views(a, ctx->L().stream));
  ctx->stats.kernel_launches += 1;
  ctx->view_B = B;
  ctx->view_cap = cap;
  return GMMLOC_OK;
}

gmmloc_status_t search_enqueue(gmmloc_ctx_t* ctx, const float* d_obs, int B, int N, int num, bool want_pos) {
  CU(ctx->kf_cand.reserve(4 * (size_t)B * N * num));
  if (want_pos) CU(ctx->kf_cand_pos.reserve(4 * (size_t)B * N * num));
  SearchArgs a;
  a.count = ctx->view_count.as<int32_t>();
  a.ids = ctx->view_ids.as<int32_t>();
  a.rec = ctx->view_rec.as<double>();
  a.cap = ctx->view_cap;
  a.obs = d_obs;
  a.B = B;
  a.N = N;
  a.num = num;
  a.cand = ctx->kf_cand.as<int32_t>();
  a.cand_pos = want_pos ? ctx->kf_cand_pos.as<int32_t>() : nullptr;
  CU(launch_search_corr(a, ctx->L().stream));
  ctx->stats.kernel_launches += 1;
  return GMMLOC_OK;
}

gmmloc_status_t check_enqueue(gmmloc_ctx_t* ctx, const gmmloc_map_t* map, const gmmloc_camera_t* cam,
                              const gmmloc_params_t* params, const double* d_pose, const float* d_obs, const float* d_depth,
                              const uint8_t* d_oct, const int32_t* d_cand, int B, int N, int num, double* d_pt3d,
                              int32_t* d_assoc) {
  CheckArgs a;
  a.mean = map->d.mean;
  a.cov_inv = map->d.cov_inv;
  a.axis = map->d.axis;
  a.is_deg = map->d.is_deg;
  a.K = map->d.K;
  a.nbr_row_ptr = map->d.nbr_row_ptr;
  a.nbr_col = map->d.nbr_col;
  fill_params(cam, params, a.prm);
  a.pose = d_pose;
  a.obs = d_obs;
  a.depth = d_depth;
  a.octave = d_oct;
  a.cand = d_cand;
  a.B = B;
  a.N = N;
  a.num = num;
  a.pt3d = d_pt3d;
  a.assoc = d_assoc;
  CU(launch_check_assoc(a, ctx->L().stream));
  ctx->stats.kernel_launches += 1;
  return GMMLOC_OK;
}

}  // namespace

gmmloc_status_t gmmloc_render_views(gmmloc_ctx_t* ctx, const gmmloc_map_t* map, const gmmloc_camera_t* cam, const double* pose,
                                    int B, int cap, int32_t* count, int32_t* status, int32_t* ids, double* mean2d,
                                    double* cov2d, double* depth) try {
  if (!ctx || !map || !cam || B < 0 || cap <= 0) return fail(GMMLOC_EINVAL, "gmmloc_render_views: bad argument");
  if (map->device != ctx->device) return fail(GMMLOC_EINVAL, "map and context live on different devices");
  if (B == 0) {
    ctx->view_B = 0;
    return GMMLOC_OK;
  }
  if (!pose) return fail(GMMLOC_EINVAL, "gmmloc_render_views: null pose");
  CU(cudaSetDevice(ctx->device));
  ctx->active = 0;
  // The device keeps at most K components per view; the HOST outputs keep the caller's `cap` as their row stride
  // (B x cap as documented), columns beyond the device capacity are padding (ids -1, records 0).
  const int dcap = std::min(cap, map->d.K);
  cudaStream_t st = ctx->L().stream;
  CU(ctx->view_pose.reserve(56 * (size_t)B));
  CU(cudaMemcpyAsync(ctx->view_pose.p, pose, 56 * (size_t)B, cudaMemcpyHostToDevice, st));
  gmmloc_status_t rc = render_enqueue(ctx, map, cam, B, dcap);
  if (rc != GMMLOC_OK) return rc;
  if (count) CU(cudaMemcpyAsync(count, ctx->view_count.p, 4 * (size_t)B, cudaMemcpyDeviceToHost, st));
  if (status) CU(cudaMemcpyAsync(status, ctx->view_status.p, 4 * (size_t)B, cudaMemcpyDeviceToHost, st));
  if (ids)
    CU(cudaMemcpy2DAsync(ids, 4 * (size_t)cap, ctx->view_ids.p, 4 * (size_t)dcap, 4 * (size_t)dcap, (size_t)B, cudaMemcpyDeviceToHost, st));
  std::vector<double> rec;
  if (mean2d || cov2d || depth) {
    rec.resize((size_t)B * dcap * kViewRec);
    CU(cudaMemcpyAsync(rec.data(), ctx->view_rec.p, 8 * rec.size(), cudaMemcpyDeviceToHost, st));
  }
  CU(cudaStreamSynchronize(st));
  if (ids && cap > dcap)
    for (size_t b = 0; b < (size_t)B; ++b)
      for (int j = dcap; j < cap; ++j) ids[b * cap + j] = -1;
  if (!rec.empty()) {
    // entries past a view's count are unspecified on the device: zero them on the way out
    std::vector<int32_t> cnt((size_t)B);
    CU(cudaMemcpy(cnt.data(), ctx->view_count.p, 4 * (size_t)B, cudaMemcpyDeviceToHost));
    for (size_t b = 0; b < (size_t)B; ++b)
      for (size_t j = 0; j < (size_t)cap; ++j) {
        const size_t e = b * cap + j;
        const bool live = (int)j < cnt[b] && (int)j < dcap;
        const double* r = &rec[(b * dcap + (live ? j : 0)) * kViewRec];
        if (mean2d) { mean2d[e * 2] = live ? r[0] : 0.0; mean2d[e * 2 + 1] = live ? r[1] : 0.0; }
        if (cov2d) for (int q = 0; q < 4; ++q) cov2d[e * 4 + q] = live ? r[2 + q] : 0.0;
        if (depth) depth[e] = live ? r[7] : 0.0;
      }
  }
  return GMMLOC_OK;
} GMMLOC_ABI_CATCH

gmmloc_status_t gmmloc_search_correspondence(gmmloc_ctx_t* ctx, const float* obs, int B, int N, int num, int32_t* cand,
                                             int32_t* cand_pos) {
  if (!ctx || B < 0 || N < 0) return fail(GMMLOC_EINVAL, "gmmloc_search_correspondence: bad argument");
  if (num < 1 || num > search_corr_max_num()) return fail(GMMLOC_EINVAL, "gmmloc_search_correspondence: num must be in [1, 8]");
  if (B != ctx->view_B) return fail(GMMLOC_EINVAL, "gmmloc_search_correspondence: B differs from the views rendered on this context");
  if (B == 0 || N == 0) return GMMLOC_OK;
  if (!obs || !cand) return fail(GMMLOC_EINVAL, "gmmloc_search_correspondence: null pointer");
  CU(cudaSetDevice(ctx->device));
  ctx->active = 0;
  cudaStream_t st = ctx->L().stream;
  const size_t M = (size_t)B * N;
  CU(ctx->kf_obs.reserve(12 * M));
  CU(cudaMemcpyAsync(ctx->kf_obs.p, obs, 12 * M, cudaMemcpyHostToDevice, st));
  gmmloc_status_t rc = search_enqueue(ctx, ctx->kf_obs.as<float>(), B, N, num, cand_pos != nullptr);
  if (rc != GMMLOC_OK) return rc;
  CU(cudaMemcpyAsync(cand, ctx->kf_cand.p, 4 * M * num, cudaMemcpyDeviceToHost, st));
  if (cand_pos) CU(cudaMemcpyAsync(cand_pos, ctx->kf_cand_pos.p, 4 * M * num, cudaMemcpyDeviceToHost, st));
  CU(cudaStreamSynchronize(st));
  return GMMLOC_OK;
}

gmmloc_status_t gmmloc_check_map_association(gmmloc_ctx_t* ctx, const gmmloc_map_t* map, const gmmloc_camera_t* cam,
                                             const gmmloc_params_t* params, const double* pose, const float* obs,
                                             const uint8_t* octave, const int32_t* cand, int B, int N, int num, double* pt3d,
                                             int32_t* assoc) {
  if (!ctx || !map || !cam || !params || B < 0 || N < 0 || num < 1)
    return fail(GMMLOC_EINVAL, "gmmloc_check_map_association: bad argument");
  if (map->device != ctx->device) return fail(GMMLOC_EINVAL, "map and context live on different devices");
  if (!map->d.nbr_row_ptr) return fail(GMMLOC_EINVAL, "gmmloc_check_map_association: call gmmloc_map_build_neighbors first");
  if (B == 0 || N == 0) return GMMLOC_OK;
  if (!pose || !obs || !octave || !cand || !pt3d || !assoc) return fail(GMMLOC_EINVAL, "gmmloc_check_map_association: null pointer");
  CU(cudaSetDevice(ctx->device));
  ctx->active = 0;
  cudaStream_t st = ctx->L().stream;
  const size_t M = (size_t)B * N;
  CU(ctx->in_pose.reserve(56 * (size_t)B));
  CU(ctx->kf_obs.reserve(12 * M));
  CU(ctx->kf_oct.reserve(M));
  CU(ctx->kf_cand.reserve(4 * M * num));
  CU(ctx->kf_pt3d.reserve(24 * M));
  CU(ctx->kf_assoc.reserve(4 * M));
  CU(cudaMemcpyAsync(ctx->in_pose.p, pose, 56 * (size_t)B, cudaMemcpyHostToDevice, st));
  CU(cudaMemcpyAsync(ctx->kf_obs.p, obs, 12 * M, cudaMemcpyHostToDevice, st));
  CU(cudaMemcpyAsync(ctx->kf_oct.p, octave, M, cudaMemcpyHostToDevice, st));
  CU(cudaMemcpyAsync(ctx->kf_cand.p, cand, 4 * M * num, cudaMemcpyHostToDevice, st));
  CU(cudaMemcpyAsync(ctx->kf_pt3d.p, pt3d, 24 * M, cudaMemcpyHostToDevice, st));
  gmmloc_status_t rc = check_enqueue(ctx, map, cam, params, ctx->in_pose.as<double>(), ctx->kf_obs.as<float>(), nullptr,
                                     ctx->kf_oct.as<uint8_t>(), ctx->kf_cand.as<int32_t>(), B, N, num, ctx->kf_pt3d.as<double>(),
                                     ctx->kf_assoc.as<int32_t>());
  if (rc != GMMLOC_OK) return rc;
  CU(cudaMemcpyAsync(pt3d, ctx->kf_pt3d.p, 24 * M, cudaMemcpyDeviceToHost, st));
  CU(cudaMemcpyAsync(assoc, ctx->kf_assoc.p, 4 * M, cudaMemcpyDeviceToHost, st));
  CU(cudaStreamSynchronize(st));
  return GMMLOC_OK;
}

gmmloc_status_t gmmloc_keyframes_associate(gmmloc_ctx_t* ctx, const gmmloc_map_t* map, const gmmloc_camera_t* cam,
                                           const gmmloc_params_t* params, const double* pose, const float* obs,
                                           const float* depth, const uint8_t* octave, int B, int N, int num, int cap,
                                           int32_t* assoc, double* pt3d, int32_t* cand, int32_t* nview, int32_t* vstatus) {
  if (!ctx || !map || !cam || !params || B < 0 || N < 0 || cap <= 0)
    return fail(GMMLOC_EINVAL, "gmmloc_keyframes_associate: bad argument");
  if (num < 1 || num > search_corr_max_num()) return fail(GMMLOC_EINVAL, "gmmloc_keyframes_associate: num must be in [1, 8]");
  if (map->device != ctx->device) return fail(GMMLOC_EINVAL, "map and context live on different devices");
  if (!map->d.nbr_row_ptr) return fail(GMMLOC_EINVAL, "gmmloc_keyframes_associate: call gmmloc_map_build_neighbors first");
  if (B == 0) return GMMLOC_OK;
  if (!pose || (N > 0 && (!obs || !depth || !octave || !assoc || !pt3d)))
    return fail(GMMLOC_EINVAL, "gmmloc_keyframes_associate: null pointer");
  CU(cudaSetDevice(ctx->device));
  ctx->active = 0;
  if (cap > map->d.K) cap = map->d.K;
  cudaStream_t st = ctx->L().stream;
  const size_t M = (size_t)B * N;
  CU(ctx->view_pose.reserve(56 * (size_t)B));
  CU(cudaMemcpyAsync(ctx->view_pose.p, pose, 56 * (size_t)B, cudaMemcpyHostToDevice, st));
  if (M) {
    CU(ctx->kf_obs.reserve(12 * M));
    CU(ctx->kf_depth.reserve(4 * M));
    CU(ctx->kf_oct.reserve(M));
    CU(ctx->kf_pt3d.reserve(24 * M));
    CU(ctx->kf_assoc.reserve(4 * M));
    CU(cudaMemcpyAsync(ctx->kf_obs.p, obs, 12 * M, cudaMemcpyHostToDevice, st));
    CU(cudaMemcpyAsync(ctx->kf_depth.p, depth, 4 * M, cudaMemcpyHostToDevice, st));
    CU(cudaMemcpyAsync(ctx->kf_oct.p, octave, M, cudaMemcpyHostToDevice, st));
  }
  gmmloc_status_t rc = render_enqueue(ctx, map, cam, B, cap);  // gmmloc_opt.cpp:122-126
  if (rc != GMMLOC_OK) return rc;
  if (M) {
    rc = search_enqueue(ctx, ctx->kf_obs.as<float>(), B, N, num, false);  // :128-134
    if (rc != GMMLOC_OK) return rc;
    rc = check_enqueue(ctx, map, cam, params, ctx->view_pose.as<double>(), ctx->kf_obs.as<float>(), ctx->kf_depth.as<float>(),
                       ctx->kf_oct.as<uint8_t>(), ctx->kf_cand.as<int32_t>(), B, N, num, ctx->kf_pt3d.as<double>(),
                       ctx->kf_assoc.as<int32_t>());  // :73-77
    if (rc != GMMLOC_OK) return rc;
    CU(cudaMemcpyAsync(assoc, ctx->kf_assoc.p, 4 * M, cudaMemcpyDeviceToHost, st));
    CU(cudaMemcpyAsync(pt3d, ctx->kf_pt3d.p, 24 * M, cudaMemcpyDeviceToHost, st));
    if (cand) CU(cudaMemcpyAsync(cand, ctx->kf_cand.p, 4 * M * num, cudaMemcpyDeviceToHost, st));
  }
  if (nview) CU(cudaMemcpyAsync(nview, ctx->view_count.p, 4 * (size_t)B, cudaMemcpyDeviceToHost, st));
  if (vstatus) CU(cudaMemcpyAsync(vstatus, ctx->view_status.p, 4 * (size_t)B, cudaMemcpyDeviceToHost, st));
  CU(cudaStreamSynchronize(st));
  return GMMLOC_OK;
}

// ---- local bundle adjustment ------------------------------------------------------------------------
gmmloc_status_t gmmloc_local_ba(gmmloc_ctx_t* ctx, const gmmloc_map_t* map, const gmmloc_camera_t* cam,
                                const gmmloc_params_t* params, int B, const int32_t* pose_ptr, double* pose, const uint8_t* fixed,
                                const int32_t* prior_idx, const int32_t* pt_ptr, double* pts, const int32_t* assoc,
                                const int32_t* obs_ptr, const int32_t* obs_pt, const int32_t* obs_kf, const float* obs,
                                const uint8_t* octave, uint8_t* obs_outlier, uint8_t* assoc_cleared, int32_t* iters,
                                int32_t* status, const volatile int32_t* stop) try {
  if (!ctx || !map || !cam || !params || B < 0) return fail(GMMLOC_EINVAL, "gmmloc_local_ba: bad argument");
  if (B == 0) return GMMLOC_OK;
  if (!pose_ptr || !pose || !fixed || !pt_ptr || !pts || !assoc || !obs_ptr || !obs_pt || !obs_kf || !obs || !octave)
    return fail(GMMLOC_EINVAL, "gmmloc_local_ba: null pointer");
  if (map->device != ctx->device) return fail(GMMLOC_EINVAL, "map and context live on different devices");
  const long long sumP = pose_ptr[B] - pose_ptr[0], sumL = pt_ptr[B] - pt_ptr[0], sumE = obs_ptr[B] - obs_ptr[0];
  if (pose_ptr[0] != 0 || pt_ptr[0] != 0 || obs_ptr[0] != 0) return fail(GMMLOC_EINVAL, "gmmloc_local_ba: *_ptr[0] must be 0");
  // ---- host-side graph preparation (the reference builds a g2o hyper-graph here, :528-766) ----
  std::vector<BaProb> prob((size_t)B);
  std::vector<int32_t> fidx((size_t)sumP), h_pt_ptr((size_t)(sumL + B)), h_kf_ptr, h_kf_obs, h_blk_ptr;
  std::vector<int2> h_contrib;
  long long free_off = 0, S_off = 0;
  for (int b = 0; b < B; ++b) {
    BaProb& d = prob[b];
    const long long p0 = pose_ptr[b], l0 = pt_ptr[b], e0 = obs_ptr[b];
    d.P = (int)(pose_ptr[b + 1] - p0);
    d.L = (int)(pt_ptr[b + 1] - l0);
    d.E = (int)(obs_ptr[b + 1] - e0);
    if (d.P < 0 || d.L < 0 || d.E < 0) return fail(GMMLOC_EINVAL, "gmmloc_local_ba: *_ptr must be non-decreasing");
    d.pose_off = p0; d.pt_off = l0; d.obs_off = e0;
    d.prior = prior_idx ? prior_idx[b] : -1;
    if (d.prior >= d.P || d.prior < -1) return fail(GMMLOC_EINVAL, "gmmloc_local_ba: prior_idx out of range");
    d.nf = 0;
    for (int i = 0; i < d.P; ++i) fidx[p0 + i] = fixed[p0 + i] ? -1 : d.nf++;
    d.n = 6 * d.nf;
    d.free_off = free_off;
    d.S_off = S_off;
    for (int l = 0; l < d.L; ++l)
      if (assoc[l0 + l] >= map->d.K) return fail(GMMLOC_EINVAL, "gmmloc_local_ba: assoc out of range");
    // CSR of the observations of every point (obs_pt non-decreasing)
    int32_t* pp = &h_pt_ptr[(size_t)(l0 + b)];
    int e = 0;
    for (int l = 0; l <= d.L; ++l) {
      while (e < d.E && obs_pt[e0 + e] < l) ++e;
      pp[l] = e;
    }
    for (int q = 0; q < d.E; ++q) {
      const int l = obs_pt[e0 + q], i = obs_kf[e0 + q];
      if (l < 0 || l >= d.L || i < 0 || i >= d.P) return fail(GMMLOC_EINVAL, "gmmloc_local_ba: observation index out of range");
      if (q > 0 && l < obs_pt[e0 + q - 1]) return fail(GMMLOC_EINVAL, "gmmloc_local_ba: obs_pt must be non-decreasing inside a window");
    }
    // observations of every free pose
    d.kfo_off = (long long)h_kf_obs.size();
    std::vector<int32_t> cnt((size_t)d.nf + 1, 0);
    for (int q = 0; q < d.E; ++q) {
      const int f = fidx[p0 + obs_kf[e0 + q]];
      if (f >= 0) ++cnt[(size_t)f + 1];
    }
    for (int f = 0; f < d.nf; ++f) cnt[(size_t)f + 1] += cnt[f];
    h_kf_ptr.insert(h_kf_ptr.end(), cnt.begin(), cnt.end());
    const size_t kbase = h_kf_obs.size();
    h_kf_obs.resize(kbase + (size_t)cnt[d.nf]);
    {
      std::vector<int32_t> cur(cnt.begin(), cnt.end() - 1);
      for (int q = 0; q < d.E; ++q) {
        const int f = fidx[p0 + obs_kf[e0 + q]];
        if (f >= 0) h_kf_obs[kbase + (size_t)cur[f]++] = q;
      }
    }
    // Schur contributions, grouped by lower block (fi >= fj), point order kept inside a block
    const int nblk = d.nf * (d.nf + 1) / 2;
    d.blk_off = (long long)h_blk_ptr.size();
    d.c_off = (long long)h_contrib.size();
    std::vector<int32_t> bc((size_t)nblk + 1, 0);
    for (int pass = 0; pass < 2; ++pass) {
      std::vector<int32_t> cur;
      if (pass == 1) {
        for (int k = 0; k < nblk; ++k) bc[(size_t)k + 1] += bc[k];
        h_blk_ptr.insert(h_blk_ptr.end(), bc.begin(), bc.end());
        h_contrib.resize((size_t)d.c_off + (size_t)bc[nblk]);
        cur.assign(bc.begin(), bc.end() - 1);
      }
      for (int l = 0; l < d.L; ++l)
        for (int qa = pp[l]; qa < pp[l + 1]; ++qa) {
          const int fa = fidx[p0 + obs_kf[e0 + qa]];
          if (fa < 0) continue;
          for (int qb = pp[l]; qb < pp[l + 1]; ++qb) {
            const int fb = fidx[p0 + obs_kf[e0 + qb]];
            if (fb < 0 || fb > fa) continue;
            const int blk = fa * (fa + 1) / 2 + fb;
            if (pass == 0) ++bc[(size_t)blk + 1];
            else h_contrib[(size_t)d.c_off + (size_t)cur[blk]++] = make_int2(qa, qb);
          }
        }
    }
    free_off += d.nf;
    S_off += (long long)d.n * d.n + 3LL * d.n;
  }
  // ---- device staging ----
  CU(cudaSetDevice(ctx->device));
  ctx->active = 0;
  cudaStream_t st = ctx->L().stream;
  const int ows = local_ba_obs_ws(), pws = local_ba_pt_ws(), qws = local_ba_pose_ws(), fws = local_ba_free_ws();
  struct Slab { size_t off, bytes; };
  size_t total = 0;
  auto take = [&](size_t bytes) { Slab s{total, bytes}; total += (bytes + 255) / 256 * 256; return s; };
  const Slab s_prob = take(sizeof(BaProb) * (size_t)B), s_pose = take(56 * (size_t)sumP), s_fidx = take(4 * (size_t)sumP),
             s_pts = take(24 * (size_t)sumL), s_assoc = take(4 * (size_t)sumL), s_opt = take(4 * (size_t)sumE),
             s_okf = take(4 * (size_t)sumE), s_obs = take(12 * (size_t)sumE), s_oct = take((size_t)sumE),
             s_ptp = take(4 * h_pt_ptr.size()), s_kfp = take(4 * h_kf_ptr.size()), s_kfo = take(4 * h_kf_obs.size() + 4),
             s_blk = take(4 * h_blk_ptr.size()), s_con = take(8 * h_contrib.size() + 8),
             s_wo = take(8 * (size_t)ows * (size_t)sumE), s_lvl = take((size_t)sumE), s_wp = take(8 * (size_t)pws * (size_t)sumL),
             s_slv = take((size_t)sumL), s_wq = take(8 * (size_t)qws * (size_t)sumP), s_wf = take(8 * (size_t)fws * (size_t)free_off + 8),
             s_S = take(8 * (size_t)S_off + 8), s_out = take((size_t)sumE), s_clr = take((size_t)sumL), s_it = take(12 * (size_t)B),
             s_st = take(4 * (size_t)B);
  CU(ctx->in_misc.reserve(total));
  char* base = ctx->in_misc.as<char>();
  auto up = [&](const Slab& s, const void* src, size_t bytes) {
    return bytes ? cudaMemcpyAsync(base + s.off, src, bytes, cudaMemcpyHostToDevice, st) : cudaSuccess;
  };
  CU(up(s_prob, prob.data(), sizeof(BaProb) * (size_t)B));
  CU(up(s_pose, pose, 56 * (size_t)sumP));
  CU(up(s_fidx, fidx.data(), 4 * (size_t)sumP));
  CU(up(s_pts, pts, 24 * (size_t)sumL));
  CU(up(s_assoc, assoc, 4 * (size_t)sumL));
  CU(up(s_opt, obs_pt, 4 * (size_t)sumE));
  CU(up(s_okf, obs_kf, 4 * (size_t)sumE));
  CU(up(s_obs, obs, 12 * (size_t)sumE));
  CU(up(s_oct, octave, (size_t)sumE));
  CU(up(s_ptp, h_pt_ptr.data(), 4 * h_pt_ptr.size()));
  CU(up(s_kfp, h_kf_ptr.data(), 4 * h_kf_ptr.size()));
  CU(up(s_kfo, h_kf_obs.data(), 4 * h_kf_obs.size()));
  CU(up(s_blk, h_blk_ptr.data(), 4 * h_blk_ptr.size()));
  CU(up(s_con, h_contrib.data(), 8 * h_contrib.size()));
  BaArgs a;
  a.prob = reinterpret_cast<const BaProb*>(base + s_prob.off);
  a.B = B;
  a.mean = map->d.mean;
  a.axis = map->d.axis;
  a.sqrt_info = map->d.sqrt_info;
  a.is_deg = map->d.is_deg;
  fill_params(cam, params, a.prm);
  a.pose = reinterpret_cast<double*>(base + s_pose.off);
  a.fidx = reinterpret_cast<const int32_t*>(base + s_fidx.off);
  a.pts = reinterpret_cast<double*>(base + s_pts.off);
  a.assoc = reinterpret_cast<const int32_t*>(base + s_assoc.off);
  a.obs_pt = reinterpret_cast<const int32_t*>(base + s_opt.off);
  a.obs_kf = reinterpret_cast<const int32_t*>(base + s_okf.off);
  a.obs = reinterpret_cast<const float*>(base + s_obs.off);
  a.octave = reinterpret_cast<const uint8_t*>(base + s_oct.off);
  a.pt_ptr = reinterpret_cast<const int32_t*>(base + s_ptp.off);
  a.kf_ptr = reinterpret_cast<const int32_t*>(base + s_kfp.off);
  a.kf_obs = reinterpret_cast<const int32_t*>(base + s_kfo.off);
  a.blk_ptr = reinterpret_cast<const int32_t*>(base + s_blk.off);
  a.contrib = reinterpret_cast<const int2*>(base + s_con.off);
  a.w_obs = reinterpret_cast<double*>(base + s_wo.off);
  a.level = reinterpret_cast<uint8_t*>(base + s_lvl.off);
  a.w_pt = reinterpret_cast<double*>(base + s_wp.off);
  a.str_level = reinterpret_cast<uint8_t*>(base + s_slv.off);
  a.w_pose = reinterpret_cast<double*>(base + s_wq.off);
  a.w_free = reinterpret_cast<double*>(base + s_wf.off);
  a.S = reinterpret_cast<double*>(base + s_S.off);
  a.obs_outlier = reinterpret_cast<uint8_t*>(base + s_out.off);
  a.assoc_cleared = reinterpret_cast<uint8_t*>(base + s_clr.off);
  a.iters = reinterpret_cast<int32_t*>(base + s_it.off);
  a.status = reinterpret_cast<int32_t*>(base + s_st.off);
  a.stop = nullptr;
  if (stop != nullptr) {
    // Localization::jointOptimization(KeyFrame*, bool* pbStopFlag, Map*): another host thread raises the flag
    // (Localization::insertKeyFrame -> interruptBA, localization.cpp:401-405) and g2o polls it.  The kernel polls a
    // device-mapped mirror; this (blocked) calling thread keeps the mirror in step with *stop until the kernel ends.
    int32_t* d_stop = nullptr;
    CU(cudaHostGetDevicePointer((void**)&d_stop, ctx->h_stop, 0));
    *reinterpret_cast<volatile int32_t*>(ctx->h_stop) = *stop;
    a.stop = d_stop;
  }
  CU(launch_local_ba(a, st, ctx->opt_ba_cluster));
  ctx->stats.kernel_launches += 1;
  if (stop != nullptr) {
    CU(cudaEventRecord(ctx->ev_ba, st));
    cudaError_t q;
    while ((q = cudaEventQuery(ctx->ev_ba)) == cudaErrorNotReady) {
      const int32_t v = *stop;
      volatile int32_t* m = reinterpret_cast<volatile int32_t*>(ctx->h_stop);
      if (*m != v) *m = v;
    }
    if (q != cudaSuccess) return fail_cuda(q, "local_ba_kernel");
  }
  CU(cudaMemcpyAsync(pose, base + s_pose.off, 56 * (size_t)sumP, cudaMemcpyDeviceToHost, st));
  if (sumL) CU(cudaMemcpyAsync(pts, base + s_pts.off, 24 * (size_t)sumL, cudaMemcpyDeviceToHost, st));
  if (obs_outlier && sumE) CU(cudaMemcpyAsync(obs_outlier, base + s_out.off, (size_t)sumE, cudaMemcpyDeviceToHost, st));
  if (assoc_cleared && sumL) CU(cudaMemcpyAsync(assoc_cleared, base + s_clr.off, (size_t)sumL, cudaMemcpyDeviceToHost, st));
  if (iters) CU(cudaMemcpyAsync(iters, base + s_it.off, 12 * (size_t)B, cudaMemcpyDeviceToHost, st));
  if (status) CU(cudaMemcpyAsync(status, base + s_st.off, 4 * (size_t)B, cudaMemcpyDeviceToHost, st));
  CU(cudaStreamSynchronize(st));
  return GMMLOC_OK;
} GMMLOC_ABI_CATCH

}  // extern "C"
