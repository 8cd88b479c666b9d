// oracle/mlo_oracle.hpp — TEST INFRASTRUCTURE, not product code.
//
// CPU restatement (C++17, no dependencies) of the per-scan registration hot path that
// mola::LidarOdometry::onLidarImpl drives (module/src/LidarOdometry.cpp:732-735, 961-962,
// 1161-1206) with the parameter values of pipelines/lidar3d-default.yaml:162-242,278-368 and
// pipelines/lidar3d-ndt.yaml:162-254.
//
// *** parity unpinned *** The arithmetic of this path lives in third-party packages that are not
// vendored under /root/reference and carry no pinned version (package.xml:18-25):
//   mp2p_icp (+ mp2p_icp_filters)   ~1.5-1.6   ICP::align, Matcher_*, Solver_*, FilterDecimateVoxels
//   mola_metric_maps                ~1.1-1.2   mola::HashedVoxelPointCloud, mola::NDT
//   MRPT                            ~2.13-2.14 poses / Lie groups
// and the reference's own tests hold no vector for it (its two end-to-end goldens need datasets
// that are absent, test/CMakeLists.txt:3,30,45).  This file restates the published algorithms of
// those packages as recorded in SURVEY.md Appendix A; each function names the upstream routine and
// the in-tree call site / YAML block that anchors it.  Choices that affect discrete results and
// could not be verified sit behind the single functions voxel_index_map(), voxel_index_filter(),
// geman_mcclure_weight(), nn_cell_order and cull_keep().
//
// Only tests/, __graft_entry__.smoke() and bench.py's cpu_baseline / --impl reference legs may
// use this code, and only as the checker or the timed CPU baseline.
#pragma once
#include <algorithm>
#include <atomic>
#include <chrono>
#include <cmath>
#include <condition_variable>
#include <cstdint>
#include <cstring>
#include <functional>
#include <limits>
#include <mutex>
#include <thread>
#include <vector>

#include "se3.hpp"

namespace orc {

// ---------------------------------------------------------------- switchable discrete choices
// mola::HashedVoxelPointCloud::coordToGlobalIdx: static_cast<int32_t>(coord * voxel_size_inv),
// truncation toward zero (SURVEY.md A.3 [VERIFY trunc vs floor]).
// The [VERIFY] choices are switchable at run time (orc_set_conventions) and mirrored by the product
// (mlo_set_option "convention_*"): whoever checks upstream's source flips a flag on both sides; the parity tests run
// under both settings of each.  Defaults = the readings recorded in SURVEY.md Appendix A.
struct Conventions {
  int index_floor = 0;  // 0: static_cast<int32_t>(x) (truncation toward zero), 1: floor
  int gm_form = 0;      // 0: c^4/(c^2+e^2)^2, 1: c^2/(c^2+e^2)^2
  int cull_metric = 0;  // 0: max-norm in cells, 1: L1 in cells, 2: Euclidean in cells
};
inline Conventions& conv() {
  static Conventions c;
  return c;
}
inline int32_t voxel_index_map(float coord, float inv_voxel) {
  const float v = coord * inv_voxel;
  return conv().index_floor ? static_cast<int32_t>(std::floor(v)) : static_cast<int32_t>(v);
}
// mp2p_icp_filters::PointCloudToVoxelGridSingle: static_cast<int32_t>(coord / resolution)
// (SURVEY.md A.6 [VERIFY]).
inline int32_t voxel_index_filter(float coord, float resolution) {
  const float v = coord / resolution;
  return conv().index_floor ? static_cast<int32_t>(std::floor(v)) : static_cast<int32_t>(v);
}
// mp2p_icp robust kernel GemanMcClure: w(e^2) = c^4 / (c^2 + e^2)^2   (SURVEY.md A.4 [VERIFY])
inline double geman_mcclure_weight(double err_sqr, double c) {
  const double c2 = c * c;
  const double d = err_sqr + c2;
  return conv().gm_form ? c2 / (d * d) : (c2 * c2) / (d * d);
}
inline double cauchy_weight(double err_sqr, double c) { return 1.0 / (1.0 + err_sqr / (c * c)); }

// ---------------------------------------------------------------- tiny thread pool (CPU baseline)
// Mirrors upstream's optional oneTBB parallel_reduce in the matcher / GN accumulation
// (SURVEY.md 2.3 iii): fixed chunking, partial results summed in chunk order => deterministic.
class Pool {
 public:
  explicit Pool(int n) : n_(std::max(1, n)) {
    for (int i = 1; i < n_; i++) th_.emplace_back([this, i] { worker(i); });
  }
  ~Pool() {
    {
      std::unique_lock<std::mutex> l(m_);
      stop_ = true;
      gen_++;
    }
    cv_.notify_all();
    for (auto& t : th_) t.join();
  }
  int size() const { return n_; }
  void run(const std::function<void(int)>& fn) {
    if (n_ == 1) {
      fn(0);
      return;
    }
    {
      std::unique_lock<std::mutex> l(m_);
      fn_ = &fn;
      pending_ = n_ - 1;
      gen_++;
    }
    cv_.notify_all();
    fn(0);
    std::unique_lock<std::mutex> l(m_);
    done_.wait(l, [this] { return pending_ == 0; });
  }

 private:
  void worker(int id) {
    uint64_t seen = 0;
    for (;;) {
      const std::function<void(int)>* f;
      {
        std::unique_lock<std::mutex> l(m_);
        cv_.wait(l, [&] { return gen_ != seen; });
        seen = gen_;
        if (stop_) return;
        f = fn_;
      }
      (*f)(id);
      {
        std::unique_lock<std::mutex> l(m_);
        if (--pending_ == 0) done_.notify_one();
      }
    }
  }
  int n_;
  std::vector<std::thread> th_;
  std::mutex m_;
  std::condition_variable cv_, done_;
  const std::function<void(int)>* fn_ = nullptr;
  int pending_ = 0;
  uint64_t gen_ = 0;
  bool stop_ = false;
};

// ---------------------------------------------------------------- hash-voxel point map
// Restates mola::HashedVoxelPointCloud (pipelines/lidar3d-default.yaml:228-242)
// and, with kind==1, the per-voxel statistics of mola::NDT (pipelines/lidar3d-ndt.yaml:234-254;
// SURVEY.md A.5).
struct VoxelStats {  // NDT only
  float mean[3];
  float normal[3];
  uint8_t is_plane;
};

class VoxelMap {
 public:
  static constexpr uint32_t HARD_LIMIT = 32;  // upstream HARDLIMIT_MAX_POINTS_PER_VOXEL
  int kind = 0;
  float voxel_size = 1.f, inv = 1.f;
  uint32_t cap = 20;
  float min_dist = 0.f;
  float max_eigen_ratio = 0.05f;
  uint32_t min_pts_plane = 5;

  VoxelMap(int kind_, float vs, uint32_t cap_, float min_dist_, float eig_ratio, uint32_t min_pts_plane_)
      : kind(kind_), voxel_size(vs), inv(1.0f / vs), min_dist(min_dist_), max_eigen_ratio(eig_ratio),
        min_pts_plane(min_pts_plane_ ? min_pts_plane_ : 5) {
    cap = (cap_ == 0 || cap_ > HARD_LIMIT) ? HARD_LIMIT : cap_;
    rehash(1 << 12);
  }

  size_t n_voxels() const { return cnt_.size(); }
  size_t n_points() const {
    size_t s = 0;
    for (auto c : cnt_) s += c;
    return s;
  }
  void clear() {
    cnt_.clear();
    vkeys_.clear();
    pts_.clear();
    stats_.clear();
    dirty_.clear();
    rehash(1 << 12);
  }

  // insertPoint loop of FilterMerge -> map insert, input order (SURVEY.md A.3 "Insert").
  void insert(const float* p, uint32_t stride, size_t n, const Pose& T) {
    const float md2 = min_dist * min_dist;
    for (size_t i = 0; i < n; i++) {
      float gx, gy, gz;
      compose_point_f(T, p[i * stride], p[i * stride + 1], p[i * stride + 2], gx, gy, gz);
      const int32_t kx = voxel_index_map(gx, inv), ky = voxel_index_map(gy, inv), kz = voxel_index_map(gz, inv);
      const int32_t v = find_or_create(kx, ky, kz);
      uint32_t& c = cnt_[v];
      if (c >= cap) continue;
      float* vp = &pts_[size_t(v) * cap * 3];
      if (min_dist > 0.f) {
        bool too_close = false;
        for (uint32_t j = 0; j < c; j++) {
          const float dx = vp[3 * j] - gx, dy = vp[3 * j + 1] - gy, dz = vp[3 * j + 2] - gz;
          const float d2 = dx * dx + dy * dy + dz * dz;
          if (d2 < md2) {
            too_close = true;
            break;
          }
        }
        if (too_close) continue;
      }
      vp[3 * c] = gx;
      vp[3 * c + 1] = gy;
      vp[3 * c + 2] = gz;
      c++;
      if (kind == 1) dirty_[v] = 1;
    }
  }

  // insertOpts.remove_voxels_farther_than (default.yaml:238; SURVEY.md A.3 [VERIFY metric]):
  // max-norm in cells against ceil(dist * voxel_size_inv).
  static bool cull_keep(int32_t kx, int32_t ky, int32_t kz, int32_t sx, int32_t sy, int32_t sz, int32_t d) {
    const int64_t ax = std::abs(kx - sx), ay = std::abs(ky - sy), az = std::abs(kz - sz);
    if (conv().cull_metric == 1) return ax + ay + az <= d;
    if (conv().cull_metric == 2) return ax * ax + ay * ay + az * az <= int64_t(d) * d;
    return ax <= d && ay <= d && az <= d;
  }
  void cull(const double sensor[3], float dist) {
    if (!(dist > 0.f)) return;
    const int32_t sx = voxel_index_map(float(sensor[0]), inv), sy = voxel_index_map(float(sensor[1]), inv),
                  sz = voxel_index_map(float(sensor[2]), inv);
    const int32_t d = static_cast<int32_t>(std::ceil(dist * inv));
    std::vector<uint32_t> cnt2;
    std::vector<int32_t> vk2;
    std::vector<float> pts2;
    std::vector<VoxelStats> st2;
    std::vector<uint8_t> dirty2;
    for (size_t v = 0; v < cnt_.size(); v++) {
      if (!cull_keep(vkeys_[3 * v], vkeys_[3 * v + 1], vkeys_[3 * v + 2], sx, sy, sz, d)) continue;
      cnt2.push_back(cnt_[v]);
      vk2.insert(vk2.end(), &vkeys_[3 * v], &vkeys_[3 * v] + 3);
      pts2.insert(pts2.end(), &pts_[v * cap * 3], &pts_[v * cap * 3] + cap * 3);
      if (kind == 1) {
        st2.push_back(stats_[v]);
        dirty2.push_back(dirty_[v]);
      }
    }
    cnt_.swap(cnt2);
    vkeys_.swap(vk2);
    pts_.swap(pts2);
    stats_.swap(st2);
    dirty_.swap(dirty2);
    size_t want = 1 << 12;
    while (want < cnt_.size() * 4) want <<= 1;
    rebuild_table(want);
  }

  int32_t find(int32_t kx, int32_t ky, int32_t kz) const {
    size_t h = hash(kx, ky, kz) & mask_;
    for (;;) {
      const int32_t v = slot_vid_[h];
      if (v < 0) return -1;
      if (vkeys_[3 * v] == kx && vkeys_[3 * v + 1] == ky && vkeys_[3 * v + 2] == kz) return v;
      h = (h + 1) & mask_;
    }
  }

  // NearestNeighborsCapable::nn_single_search: the 27 cells key(q)+{-1,0,1}^3 in
  // cx, cy, cz nested order, stored slot order inside a cell, strict '<' keeps the first minimum.
  bool nn_single(float qx, float qy, float qz, float out[3], float& out_d2, uint64_t* n_candidates = nullptr) const {
    const int32_t kx = voxel_index_map(qx, inv), ky = voxel_index_map(qy, inv), kz = voxel_index_map(qz, inv);
    float best = std::numeric_limits<float>::infinity();
    bool found = false;
    uint64_t nc = 0;
    for (int dx = -1; dx <= 1; dx++)
      for (int dy = -1; dy <= 1; dy++)
        for (int dz = -1; dz <= 1; dz++) {
          const int32_t v = find(kx + dx, ky + dy, kz + dz);
          if (v < 0) continue;
          const float* vp = &pts_[size_t(v) * cap * 3];
          const uint32_t c = cnt_[v];
          nc += c;
          for (uint32_t j = 0; j < c; j++) {
            const float ex = vp[3 * j] - qx, ey = vp[3 * j + 1] - qy, ez = vp[3 * j + 2] - qz;
            const float d2 = ex * ex + ey * ey + ez * ez;
            if (d2 < best) {
              best = d2;
              out[0] = vp[3 * j];
              out[1] = vp[3 * j + 1];
              out[2] = vp[3 * j + 2];
              found = true;
            }
          }
        }
    out_d2 = best;
    if (n_candidates) *n_candidates += nc;
    return found;
  }

  // mola::NDT nearest-plane query: among the 27 cells around key(q), the planar
  // voxel with the smallest |n.(q - mean)|; same cell order / strict '<' tie rule as nn_single.
  bool nn_plane(float qx, float qy, float qz, float mean[3], float normal[3], float& out_dist,
                uint64_t* n_candidates = nullptr) {
    const int32_t kx = voxel_index_map(qx, inv), ky = voxel_index_map(qy, inv), kz = voxel_index_map(qz, inv);
    float best = std::numeric_limits<float>::infinity();
    bool found = false;
    for (int dx = -1; dx <= 1; dx++)
      for (int dy = -1; dy <= 1; dy++)
        for (int dz = -1; dz <= 1; dz++) {
          const int32_t v = find(kx + dx, ky + dy, kz + dz);
          if (v < 0) continue;
          if (n_candidates) *n_candidates += 2;  // mean + normal records
          const VoxelStats& s = stats(v);
          if (!s.is_plane) continue;
          const float ex = qx - s.mean[0], ey = qy - s.mean[1], ez = qz - s.mean[2];
          const float d = std::fabs(s.normal[0] * ex + s.normal[1] * ey + s.normal[2] * ez);
          if (d < best) {
            best = d;
            for (int k = 0; k < 3; k++) {
              mean[k] = s.mean[k];
              normal[k] = s.normal[k];
            }
            found = true;
          }
        }
    out_dist = best;
    return found;
  }

  void update_all_stats() {
    if (kind != 1) return;
    for (size_t v = 0; v < cnt_.size(); v++) (void)stats(int32_t(v));
  }
  // Lazy per-voxel mean / covariance / eigen (mola::NDT).  Double accumulation in
  // stored order; symmetric 3x3 eigen by cyclic Jacobi (fixed 12 sweeps) so the device can run the
  // identical sequence; plane iff n >= min_pts_plane and l_min < ratio * l_max.
  const VoxelStats& stats(int32_t v) {
    if (!dirty_[v]) return stats_[v];
    dirty_[v] = 0;
    VoxelStats& s = stats_[v];
    compute_stats(&pts_[size_t(v) * cap * 3], cnt_[v], max_eigen_ratio, min_pts_plane, s);
    return s;
  }
  static void jacobi3(double A[3][3], double V[3][3]) {
    for (int i = 0; i < 3; i++)
      for (int j = 0; j < 3; j++) V[i][j] = (i == j) ? 1.0 : 0.0;
    for (int sweep = 0; sweep < 12; sweep++) {
      for (int p = 0; p < 2; p++)
        for (int q = p + 1; q < 3; q++) {
          const double apq = A[p][q];
          if (std::fabs(apq) < 1e-300) continue;
          const double theta = (A[q][q] - A[p][p]) / (2.0 * apq);
          const double t = (theta >= 0 ? 1.0 : -1.0) / (std::fabs(theta) + std::sqrt(theta * theta + 1.0));
          const double c = 1.0 / std::sqrt(t * t + 1.0), s = t * c;
          for (int k = 0; k < 3; k++) {
            const double akp = A[k][p], akq = A[k][q];
            A[k][p] = c * akp - s * akq;
            A[k][q] = s * akp + c * akq;
          }
          for (int k = 0; k < 3; k++) {
            const double apk = A[p][k], aqk = A[q][k];
            A[p][k] = c * apk - s * aqk;
            A[q][k] = s * apk + c * aqk;
          }
          for (int k = 0; k < 3; k++) {
            const double vkp = V[k][p], vkq = V[k][q];
            V[k][p] = c * vkp - s * vkq;
            V[k][q] = s * vkp + c * vkq;
          }
        }
    }
  }
  static void compute_stats(const float* vp, uint32_t n, float ratio, uint32_t min_pts, VoxelStats& s) {
    std::memset(&s, 0, sizeof(s));
    if (n == 0) return;
    double m[3] = {0, 0, 0};
    for (uint32_t j = 0; j < n; j++)
      for (int k = 0; k < 3; k++) m[k] += double(vp[3 * j + k]);
    const double invn = 1.0 / double(n);
    for (int k = 0; k < 3; k++) m[k] *= invn;
    for (int k = 0; k < 3; k++) s.mean[k] = float(m[k]);
    if (n < min_pts) return;
    double C[3][3] = {{0}};
    for (uint32_t j = 0; j < n; j++) {
      const double d[3] = {double(vp[3 * j]) - m[0], double(vp[3 * j + 1]) - m[1], double(vp[3 * j + 2]) - m[2]};
      for (int a = 0; a < 3; a++)
        for (int b = a; b < 3; b++) C[a][b] += d[a] * d[b];
    }
    const double invn1 = 1.0 / double(n - 1);
    for (int a = 0; a < 3; a++)
      for (int b = a; b < 3; b++) {
        C[a][b] *= invn1;
        C[b][a] = C[a][b];
      }
    double V[3][3];
    jacobi3(C, V);
    int imin = 0, imax = 0;
    for (int k = 1; k < 3; k++) {
      if (C[k][k] < C[imin][imin]) imin = k;
      if (C[k][k] > C[imax][imax]) imax = k;
    }
    const double lmin = C[imin][imin], lmax = C[imax][imax];
    if (!(lmax > 0.0)) return;
    if (lmin < double(ratio) * lmax) {
      s.is_plane = 1;
      for (int k = 0; k < 3; k++) s.normal[k] = float(V[k][imin]);
    }
  }

  // flat export sorted by key (tests compare against mlo_map_export)
  void export_sorted(std::vector<int32_t>& keys, std::vector<uint32_t>& counts, std::vector<float>& xyz) const {
    std::vector<uint32_t> order(cnt_.size());
    for (size_t i = 0; i < order.size(); i++) order[i] = uint32_t(i);
    std::sort(order.begin(), order.end(), [&](uint32_t a, uint32_t b) {
      for (int k = 0; k < 3; k++)
        if (vkeys_[3 * a + k] != vkeys_[3 * b + k]) return vkeys_[3 * a + k] < vkeys_[3 * b + k];
      return false;
    });
    keys.clear();
    counts.clear();
    xyz.clear();
    for (uint32_t v : order) {
      keys.insert(keys.end(), &vkeys_[3 * v], &vkeys_[3 * v] + 3);
      counts.push_back(cnt_[v]);
      xyz.insert(xyz.end(), &pts_[size_t(v) * cap * 3], &pts_[size_t(v) * cap * 3] + 3 * cnt_[v]);
    }
  }

 private:
  static size_t hash(int32_t x, int32_t y, int32_t z) {
    return size_t(uint32_t(x) * 73856093u ^ uint32_t(y) * 19349663u ^ uint32_t(z) * 83492791u);
  }
  void rehash(size_t n) {
    slot_vid_.assign(n, -1);
    mask_ = n - 1;
  }
  void rebuild_table(size_t n) {
    rehash(n);
    for (size_t v = 0; v < cnt_.size(); v++) {
      size_t h = hash(vkeys_[3 * v], vkeys_[3 * v + 1], vkeys_[3 * v + 2]) & mask_;
      while (slot_vid_[h] >= 0) h = (h + 1) & mask_;
      slot_vid_[h] = int32_t(v);
    }
  }
  int32_t find_or_create(int32_t kx, int32_t ky, int32_t kz) {
    if ((cnt_.size() + 1) * 4 > slot_vid_.size()) rebuild_table(slot_vid_.size() * 2);
    size_t h = hash(kx, ky, kz) & mask_;
    for (;;) {
      const int32_t v = slot_vid_[h];
      if (v < 0) break;
      if (vkeys_[3 * v] == kx && vkeys_[3 * v + 1] == ky && vkeys_[3 * v + 2] == kz) return v;
      h = (h + 1) & mask_;
    }
    const int32_t v = int32_t(cnt_.size());
    slot_vid_[h] = v;
    cnt_.push_back(0);
    vkeys_.push_back(kx);
    vkeys_.push_back(ky);
    vkeys_.push_back(kz);
    pts_.resize(pts_.size() + size_t(cap) * 3, 0.f);
    if (kind == 1) {
      stats_.push_back(VoxelStats{});
      dirty_.push_back(1);
    }
    return v;
  }
  std::vector<int32_t> slot_vid_;
  size_t mask_ = 0;
  std::vector<uint32_t> cnt_;
  std::vector<int32_t> vkeys_;
  std::vector<float> pts_;
  std::vector<VoxelStats> stats_;
  std::vector<uint8_t> dirty_;
};

// ---------------------------------------------------------------- FilterDecimateVoxels (FirstPoint)
// mp2p_icp_filters::FilterDecimateVoxels with DecimateMethod::FirstPoint
// (pipelines/lidar3d-default.yaml:285-292,312-319).  Optional predicates restate
// FilterByRange (:297-302) and FilterBoundingBox "outside" (:305-310), applied before decimation.
// Output: kept input indices in ascending order (upstream's order is hash-iteration order, i.e.
// implementation-defined; we fix it to input order).
struct DecimateParams {
  float resolution = 0.5f;
  uint32_t min_input_points = 2000;
  bool use_range = false;
  float range_min = 0, range_max = 0;
  bool use_bbox_outside = false;
  float bbox_min[3] = {0, 0, 0}, bbox_max[3] = {0, 0, 0};
};

inline bool predicate_keep(const DecimateParams& p, float x, float y, float z) {
  if (p.use_range) {
    const float n2 = x * x + y * y + z * z;
    const float lo = p.range_min * p.range_min, hi = p.range_max * p.range_max;
    if (!(n2 >= lo && n2 <= hi)) return false;
  }
  if (p.use_bbox_outside) {
    const bool inside = x >= p.bbox_min[0] && y >= p.bbox_min[1] && z >= p.bbox_min[2] && x <= p.bbox_max[0] &&
                        y <= p.bbox_max[1] && z <= p.bbox_max[2];
    if (inside) return false;
  }
  return true;
}

inline void decimate_first(const float* p, uint32_t stride, size_t n, const DecimateParams& prm,
                           std::vector<uint32_t>& kept) {
  kept.clear();
  std::vector<uint32_t> cand;
  cand.reserve(n);
  for (size_t i = 0; i < n; i++)
    if (predicate_keep(prm, p[i * stride], p[i * stride + 1], p[i * stride + 2])) cand.push_back(uint32_t(i));
  if (cand.size() < prm.min_input_points) {
    kept = cand;
    return;
  }
  size_t cap = 1024;
  while (cap < cand.size() * 2) cap <<= 1;
  struct Slot {
    int32_t k[3];
    int32_t used;
  };
  std::vector<Slot> tab(cap, Slot{{0, 0, 0}, 0});
  const size_t mask = cap - 1;
  for (uint32_t i : cand) {
    const int32_t kx = voxel_index_filter(p[size_t(i) * stride], prm.resolution),
                  ky = voxel_index_filter(p[size_t(i) * stride + 1], prm.resolution),
                  kz = voxel_index_filter(p[size_t(i) * stride + 2], prm.resolution);
    size_t h = size_t(uint32_t(kx) * 73856093u ^ uint32_t(ky) * 19349663u ^ uint32_t(kz) * 83492791u) & mask;
    bool isnew = true;
    while (tab[h].used) {
      if (tab[h].k[0] == kx && tab[h].k[1] == ky && tab[h].k[2] == kz) {
        isnew = false;
        break;
      }
      h = (h + 1) & mask;
    }
    if (isnew) {
      tab[h] = Slot{{kx, ky, kz}, 1};
      kept.push_back(i);
    }
  }
}

// ---------------------------------------------------------------- FilterDeskew
// mp2p_icp_filters::FilterDeskew (pipelines/lidar3d-default.yaml:328-350): per point with relative time
// t, p' = exp_SO3(w t) p + v t.  Rotation coefficients by the Taylor series to t^8 for |w t| < 0.05 rad (exact to
// < 1e-18 there) so that the device reproduces them bit for bit; libm beyond.  Double arithmetic, float result.
inline void deskew_coeffs(double th2, double& A, double& B) {
  if (th2 < 2.5e-3) {
    A = 1.0 - th2 * (1.0 / 6.0) * (1.0 - th2 * (1.0 / 20.0) * (1.0 - th2 * (1.0 / 42.0) * (1.0 - th2 * (1.0 / 72.0))));
    B = 0.5 * (1.0 - th2 * (1.0 / 12.0) * (1.0 - th2 * (1.0 / 30.0) * (1.0 - th2 * (1.0 / 56.0) * (1.0 - th2 * (1.0 / 90.0)))));
  } else {
    const double th = std::sqrt(th2);
    A = std::sin(th) / th;
    B = (1.0 - std::cos(th)) / th2;
  }
}
inline void deskew(const float* xyzt, size_t n, const double twist[6], float* out_xyz) {
  for (size_t i = 0; i < n; i++) {
    const double x = xyzt[4 * i], y = xyzt[4 * i + 1], z = xyzt[4 * i + 2], t = xyzt[4 * i + 3];
    const double wx = twist[3] * t, wy = twist[4] * t, wz = twist[5] * t;
    double A, B;
    deskew_coeffs(wx * wx + wy * wy + wz * wz, A, B);
    // R p = p + A (w x p) + B (w x (w x p))
    const double cx = wy * z - wz * y, cy = wz * x - wx * z, cz = wx * y - wy * x;
    const double dx = wy * cz - wz * cy, dy = wz * cx - wx * cz, dz = wx * cy - wy * cx;
    out_xyz[3 * i] = static_cast<float>(x + A * cx + B * dx + twist[0] * t);
    out_xyz[3 * i + 1] = static_cast<float>(y + A * cy + B * dy + twist[1] * t);
    out_xyz[3 * i + 2] = static_cast<float>(z + A * cz + B * dz + twist[2] * t);
  }
}

// ---------------------------------------------------------------- ICP
struct IcpParams {
  uint32_t max_iterations = 300;
  double min_abs_step_trans = 1e-4, min_abs_step_rot = 5e-5;
  int solver = 0;  // 0 GN, 1 Horn
  uint32_t gn_max_iterations = 2;
  double gn_min_delta = 1e-7;
  int robust_kernel = 1;
  uint32_t matcher_mask = 1;
  std::vector<double> thr_pt2pt, thr_pt2pl, kernel_param;  // per-iteration tables (last entry repeats)
  double threshold_angular_deg = 0;
  double w_pt2pt = 1.0, w_pt2pl = 1.0;
  bool has_prior = false;
  Pose prior_pose = Pose::identity();
  double prior_info[6][6] = {{0}};
  bool hook_enabled = false;
  double hook_min_trans = 0, hook_min_rot = 0;
  Pose hook_checkpoint = Pose::identity();
};

enum Term { T_UNDEF = 0, T_NO_PAIRINGS = 1, T_SOLVER_ERROR = 2, T_MAX_ITER = 3, T_STALLED = 4, T_HOOK = 5 };

struct IcpResult {
  Pose pose = Pose::identity();
  double cov[6][6] = {{0}};
  double quality = 0;
  uint32_t n_iterations = 0;
  int termination = T_UNDEF;
  uint64_t n_pairings = 0, n_potential = 0;
  uint64_t n_query_iterations = 0, n_candidate_points = 0;
  // per-iteration trace (tests): pose after each executed iteration, pair counts
  std::vector<Pose> trace_pose;
  std::vector<uint32_t> trace_pairs;
};

struct PairPt {  // mp2p_icp::point_pair: global + untransformed local
  float g[3], l[3];
};
struct PairPl {  // mp2p_icp::point_plane_pair
  float c[3], n[3], l[3];
};

inline double tab(const std::vector<double>& t, uint32_t it) {
  if (t.empty()) return 0.0;
  return t[std::min<size_t>(it, t.size() - 1)];
}

struct HG {
  double H[6][6];
  double g[6];
  void zero() { std::memset(this, 0, sizeof(*this)); }
  void add(const HG& o) {
    for (int i = 0; i < 6; i++) {
      for (int j = 0; j < 6; j++) H[i][j] += o.H[i][j];
      g[i] += o.g[i];
    }
  }
};

// One pair's contribution to the normal equations, written out the way upstream's
// optimal_tf_gauss_newton does it: r = T l - g, J = [R | -R [l]x],
// w = pair weight * robust(|r|^2), H += w J^T J, g += w J^T r.
inline void accumulate_pt2pt(const Pose& T, const PairPt& p, double weight, int kernel, double c, HG& a) {
  const double l[3] = {p.l[0], p.l[1], p.l[2]};
  double r[3];
  for (int i = 0; i < 3; i++) r[i] = T.R[i][0] * l[0] + T.R[i][1] * l[1] + T.R[i][2] * l[2] + T.t[i] - double(p.g[i]);
  double J[3][6];
  // -R [l]x : column k of [l]x is l x e_k ... ([l]x)_{ij}: [[0,-lz,ly],[lz,0,-lx],[-ly,lx,0]]
  const double Lx[3][3] = {{0, -l[2], l[1]}, {l[2], 0, -l[0]}, {-l[1], l[0], 0}};
  for (int i = 0; i < 3; i++) {
    for (int j = 0; j < 3; j++) J[i][j] = T.R[i][j];
    for (int j = 0; j < 3; j++) J[i][3 + j] = -(T.R[i][0] * Lx[0][j] + T.R[i][1] * Lx[1][j] + T.R[i][2] * Lx[2][j]);
  }
  const double e2 = r[0] * r[0] + r[1] * r[1] + r[2] * r[2];
  double w = weight;
  if (kernel == 1)
    w *= geman_mcclure_weight(e2, c);
  else if (kernel == 2)
    w *= cauchy_weight(e2, c);
  for (int i = 0; i < 6; i++) {
    const double jtr = J[0][i] * r[0] + J[1][i] * r[1] + J[2][i] * r[2];
    a.g[i] += w * jtr;
    for (int j = 0; j < 6; j++) a.H[i][j] += w * (J[0][i] * J[0][j] + J[1][i] * J[1][j] + J[2][i] * J[2][j]);
  }
}

inline void accumulate_pt2pl(const Pose& T, const PairPl& p, double weight, int kernel, double c, HG& a) {
  const double l[3] = {p.l[0], p.l[1], p.l[2]};
  const double n[3] = {p.n[0], p.n[1], p.n[2]};
  double g[3];
  for (int i = 0; i < 3; i++) g[i] = T.R[i][0] * l[0] + T.R[i][1] * l[1] + T.R[i][2] * l[2] + T.t[i];
  const double r = n[0] * (g[0] - double(p.c[0])) + n[1] * (g[1] - double(p.c[1])) + n[2] * (g[2] - double(p.c[2]));
  const double Lx[3][3] = {{0, -l[2], l[1]}, {l[2], 0, -l[0]}, {-l[1], l[0], 0}};
  double J3[3][6];
  for (int i = 0; i < 3; i++) {
    for (int j = 0; j < 3; j++) J3[i][j] = T.R[i][j];
    for (int j = 0; j < 3; j++) J3[i][3 + j] = -(T.R[i][0] * Lx[0][j] + T.R[i][1] * Lx[1][j] + T.R[i][2] * Lx[2][j]);
  }
  double J[6];
  for (int j = 0; j < 6; j++) J[j] = n[0] * J3[0][j] + n[1] * J3[1][j] + n[2] * J3[2][j];
  double w = weight;
  if (kernel == 1)
    w *= geman_mcclure_weight(r * r, c);
  else if (kernel == 2)
    w *= cauchy_weight(r * r, c);
  for (int i = 0; i < 6; i++) {
    a.g[i] += w * J[i] * r;
    for (int j = 0; j < 6; j++) a.H[i][j] += w * J[i] * J[j];
  }
}

// d log(D exp(eps)) / d eps at eps = 0 (upstream Lie::SE<3>::jacob_dDinvP1invP2_de1e2, the e2 half),
// evaluated here by central differences: an implementation independent of the device's closed form.
inline void prior_jacobian(const Pose& D, double J[6][6]) {
  const double h = 1e-6;
  for (int k = 0; k < 6; k++) {
    double e[6] = {0, 0, 0, 0, 0, 0};
    e[k] = h;
    double lp[6], lm[6];
    se3_log(compose(D, se3_exp(e)), lp);
    e[k] = -h;
    se3_log(compose(D, se3_exp(e)), lm);
    for (int r = 0; r < 6; r++) J[r][k] = (lp[r] - lm[r]) / (2 * h);
  }
}

inline void add_prior(const Pose& T, const IcpParams& p, HG& a) {
  const Pose D = minus(T, p.prior_pose);  // prior^-1 * T
  double e[6];
  se3_log(D, e);
  double J[6][6];
  prior_jacobian(D, J);
  // g += J^T L e ; H += J^T L J
  double LJ[6][6], Le[6];
  for (int i = 0; i < 6; i++) {
    Le[i] = 0;
    for (int k = 0; k < 6; k++) Le[i] += p.prior_info[i][k] * e[k];
    for (int j = 0; j < 6; j++) {
      LJ[i][j] = 0;
      for (int k = 0; k < 6; k++) LJ[i][j] += p.prior_info[i][k] * J[k][j];
    }
  }
  for (int i = 0; i < 6; i++) {
    for (int k = 0; k < 6; k++) a.g[i] += J[k][i] * Le[k];
    for (int j = 0; j < 6; j++)
      for (int k = 0; k < 6; k++) a.H[i][j] += J[k][i] * LJ[k][j];
  }
}

// mp2p_icp::Solver_Horn / olae-free closed form: weighted centroids, 3x3
// cross-covariance, Horn's 4x4 symmetric N matrix, dominant eigenvector by Jacobi -> quaternion.
inline void jacobi4(double A[4][4], double V[4][4]) {
  for (int i = 0; i < 4; i++)
    for (int j = 0; j < 4; j++) V[i][j] = (i == j) ? 1.0 : 0.0;
  for (int sweep = 0; sweep < 16; sweep++)
    for (int p = 0; p < 3; p++)
      for (int q = p + 1; q < 4; q++) {
        const double apq = A[p][q];
        if (std::fabs(apq) < 1e-300) continue;
        const double theta = (A[q][q] - A[p][p]) / (2.0 * apq);
        const double t = (theta >= 0 ? 1.0 : -1.0) / (std::fabs(theta) + std::sqrt(theta * theta + 1.0));
        const double c = 1.0 / std::sqrt(t * t + 1.0), s = t * c;
        for (int k = 0; k < 4; k++) {
          const double akp = A[k][p], akq = A[k][q];
          A[k][p] = c * akp - s * akq;
          A[k][q] = s * akp + c * akq;
        }
        for (int k = 0; k < 4; k++) {
          const double apk = A[p][k], aqk = A[q][k];
          A[p][k] = c * apk - s * aqk;
          A[q][k] = s * apk + c * aqk;
        }
        for (int k = 0; k < 4; k++) {
          const double vkp = V[k][p], vkq = V[k][q];
          V[k][p] = c * vkp - s * vkq;
          V[k][q] = s * vkp + c * vkq;
        }
      }
}

inline bool horn_solve(const std::vector<PairPt>& pairs, Pose& out) {
  if (pairs.size() < 3) return false;
  double cg[3] = {0, 0, 0}, cl[3] = {0, 0, 0};
  for (const auto& p : pairs)
    for (int k = 0; k < 3; k++) {
      cg[k] += p.g[k];
      cl[k] += p.l[k];
    }
  const double invn = 1.0 / double(pairs.size());
  for (int k = 0; k < 3; k++) {
    cg[k] *= invn;
    cl[k] *= invn;
  }
  double S[3][3] = {{0}};  // sum (l - cl)(g - cg)^T
  for (const auto& p : pairs) {
    const double a[3] = {p.l[0] - cl[0], p.l[1] - cl[1], p.l[2] - cl[2]};
    const double b[3] = {p.g[0] - cg[0], p.g[1] - cg[1], p.g[2] - cg[2]};
    for (int i = 0; i < 3; i++)
      for (int j = 0; j < 3; j++) S[i][j] += a[i] * b[j];
  }
  double N[4][4] = {
      {S[0][0] + S[1][1] + S[2][2], S[1][2] - S[2][1], S[2][0] - S[0][2], S[0][1] - S[1][0]},
      {S[1][2] - S[2][1], S[0][0] - S[1][1] - S[2][2], S[0][1] + S[1][0], S[2][0] + S[0][2]},
      {S[2][0] - S[0][2], S[0][1] + S[1][0], -S[0][0] + S[1][1] - S[2][2], S[1][2] + S[2][1]},
      {S[0][1] - S[1][0], S[2][0] + S[0][2], S[1][2] + S[2][1], -S[0][0] - S[1][1] + S[2][2]}};
  double V[4][4];
  jacobi4(N, V);
  int im = 0;
  for (int k = 1; k < 4; k++)
    if (N[k][k] > N[im][im]) im = k;
  double q[4] = {V[0][im], V[1][im], V[2][im], V[3][im]};
  const double qn = std::sqrt(q[0] * q[0] + q[1] * q[1] + q[2] * q[2] + q[3] * q[3]);
  if (!(qn > 0)) return false;
  for (double& v : q) v /= qn;
  if (q[0] < 0)
    for (double& v : q) v = -v;
  const double w = q[0], x = q[1], y = q[2], z = q[3];
  out.R[0][0] = 1 - 2 * (y * y + z * z);
  out.R[0][1] = 2 * (x * y - w * z);
  out.R[0][2] = 2 * (x * z + w * y);
  out.R[1][0] = 2 * (x * y + w * z);
  out.R[1][1] = 1 - 2 * (x * x + z * z);
  out.R[1][2] = 2 * (y * z - w * x);
  out.R[2][0] = 2 * (x * z - w * y);
  out.R[2][1] = 2 * (y * z + w * x);
  out.R[2][2] = 1 - 2 * (x * x + y * y);
  for (int i = 0; i < 3; i++) out.t[i] = cg[i] - (out.R[i][0] * cl[0] + out.R[i][1] * cl[1] + out.R[i][2] * cl[2]);
  return true;
}

// mp2p_icp::ICP::align (call site LidarOdometry.cpp:961-962; loop structure SURVEY.md A.1) with
// Matcher_Points_DistanceThreshold (A.2), Matcher_Point2Plane over mola::NDT (A.5),
// Solver_GaussNewton (A.4) or Solver_Horn, QualityEvaluator_PairedRatio (A.8).
inline void icp_align(const float* lp, uint32_t stride, size_t n, VoxelMap& map, const Pose& init,
                      const IcpParams& prm, IcpResult& res, Pool* pool = nullptr, bool trace = false) {
  res = IcpResult{};
  Pose T = init, prev = init, prev2 = init;
  bool has_prev2 = false;
  std::vector<PairPt> pt;
  std::vector<PairPl> pl;
  HG last_hg;
  last_hg.zero();
  bool have_hg = false;
  const int nth = pool ? pool->size() : 1;
  std::vector<std::vector<PairPt>> cpt(nth);
  std::vector<std::vector<PairPl>> cpl(nth);
  std::vector<HG> chg(nth);
  std::vector<uint64_t> ccand(nth);
  if (prm.matcher_mask & 2u) map.update_all_stats();  // lazy stats are not thread-safe: realise up front

  const double ang = prm.threshold_angular_deg * M_PI / 180.0;
  const float ang2 = float(ang * ang);

  for (; res.n_iterations < prm.max_iterations; res.n_iterations++) {
    const uint32_t it = res.n_iterations;
    const double thr = tab(prm.thr_pt2pt, it), thr_pl = tab(prm.thr_pt2pl, it), kc = tab(prm.kernel_param, it);
    const float thr2 = float(thr * thr);
    const float thr_plf = float(thr_pl);
    // ---- matchers
    uint64_t potential = 0;
    if (prm.matcher_mask & 2u) potential += n;
    if (prm.matcher_mask & 1u) potential += n;
    auto match_chunk = [&](int c) {
      const size_t lo = n * size_t(c) / nth, hi = n * size_t(c + 1) / nth;
      cpt[c].clear();
      cpl[c].clear();
      ccand[c] = 0;
      for (size_t i = lo; i < hi; i++) {
        const float lx = lp[i * stride], ly = lp[i * stride + 1], lz = lp[i * stride + 2];
        float gx, gy, gz;
        compose_point_f(T, lx, ly, lz, gx, gy, gz);
        bool paired = false;
        if (prm.matcher_mask & 2u) {
          float mean[3] = {0, 0, 0}, nrm[3] = {0, 0, 0}, d;
          if (map.nn_plane(gx, gy, gz, mean, nrm, d, &ccand[c]) && d < thr_plf) {
            cpl[c].push_back(PairPl{{mean[0], mean[1], mean[2]}, {nrm[0], nrm[1], nrm[2]}, {lx, ly, lz}});
            paired = true;
          }
        }
        // Matcher base rule: a local point already paired by an earlier matcher is skipped
        if ((prm.matcher_mask & 1u) && !paired) {
          float q[3] = {0, 0, 0}, d2;
          if (map.nn_single(gx, gy, gz, q, d2, &ccand[c])) {
            const float lim = thr2 + ang2 * (gx * gx + gy * gy + gz * gz);
            if (d2 < lim) cpt[c].push_back(PairPt{{q[0], q[1], q[2]}, {lx, ly, lz}});
          }
        }
      }
    };
    if (pool)
      pool->run(match_chunk);
    else
      match_chunk(0);
    pt.clear();
    pl.clear();
    for (int c = 0; c < nth; c++) {
      pt.insert(pt.end(), cpt[c].begin(), cpt[c].end());
      pl.insert(pl.end(), cpl[c].begin(), cpl[c].end());
      res.n_candidate_points += ccand[c];
    }
    res.n_query_iterations += n;
    res.n_potential = potential;
    res.n_pairings = pt.size() + pl.size();
    if (pt.empty() && pl.empty()) {
      res.termination = T_NO_PAIRINGS;
      break;
    }
    // ---- solver
    bool ok = true;
    if (prm.solver == 1) {
      ok = horn_solve(pt, T);
    } else {
      for (uint32_t inner = 0; inner < prm.gn_max_iterations; inner++) {
        auto acc_chunk = [&](int c) {
          chg[c].zero();
          const size_t lo = pt.size() * size_t(c) / nth, hi = pt.size() * size_t(c + 1) / nth;
          for (size_t i = lo; i < hi; i++) accumulate_pt2pt(T, pt[i], prm.w_pt2pt, prm.robust_kernel, kc, chg[c]);
          const size_t lo2 = pl.size() * size_t(c) / nth, hi2 = pl.size() * size_t(c + 1) / nth;
          for (size_t i = lo2; i < hi2; i++) accumulate_pt2pl(T, pl[i], prm.w_pt2pl, prm.robust_kernel, kc, chg[c]);
        };
        if (pool)
          pool->run(acc_chunk);
        else
          acc_chunk(0);
        HG a;
        a.zero();
        for (int c = 0; c < nth; c++) a.add(chg[c]);
        if (prm.has_prior) add_prior(T, prm, a);
        last_hg = a;
        have_hg = true;
        double mg[6], delta[6];
        for (int i = 0; i < 6; i++) mg[i] = -a.g[i];
        if (!ldlt6_solve(a.H, mg, delta)) {
          ok = false;
          break;
        }
        T = compose(T, se3_exp(delta));
        double dn = 0;
        for (int i = 0; i < 6; i++) dn += delta[i] * delta[i];
        if (std::sqrt(dn) < prm.gn_min_delta) break;
      }
    }
    if (!ok) {
      res.termination = T_SOLVER_ERROR;
      break;
    }
    // ---- convergence measure (min of step vs prev and vs prev-prev: catches 2-cycles)
    double d[6];
    se3_log(minus(T, prev), d);
    double dt = norm3(d), dr = norm3(d + 3);
    if (has_prev2) {
      double d2[6];
      se3_log(minus(T, prev2), d2);
      dt = std::min(dt, norm3(d2));
      dr = std::min(dr, norm3(d2 + 3));
    }
    prev2 = prev;
    has_prev2 = true;
    prev = T;
    if (trace) {
      res.trace_pose.push_back(T);
      res.trace_pairs.push_back(uint32_t(pt.size() + pl.size()));
    }
    // ---- iteration hook as data (LidarOdometry.cpp:923-952)
    if (prm.hook_enabled) {
      const Pose dd = minus(T, prm.hook_checkpoint);
      double w[3];
      so3_log(dd.R, w);
      if (norm3(dd.t) > prm.hook_min_trans || norm3(w) > prm.hook_min_rot) {
        res.termination = T_HOOK;
        break;
      }
    }
    if (std::fabs(dt) < prm.min_abs_step_trans && std::fabs(dr) < prm.min_abs_step_rot) {
      res.termination = T_STALLED;
      break;
    }
  }
  if (res.n_iterations >= prm.max_iterations) res.termination = T_MAX_ITER;
  res.pose = T;
  res.quality = res.n_potential ? double(res.n_pairings) / double(res.n_potential) : 0.0;
  if (have_hg) spd6_inverse(last_hg.H, res.cov);
}

}  // namespace orc
