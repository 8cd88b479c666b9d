// oracle/backend_oracle.hpp — TEST INFRASTRUCTURE: runs the product's host orchestrator
// (mola_lidar_odometry_b200/host/pipeline.hpp) over the CPU oracle, so tests can compare whole trajectories
// (caller contract + hot path) GPU vs CPU with identical host logic.  parity unpinned — see mlo_oracle.hpp.
#pragma once
#include <algorithm>
#include <vector>

#include "mlo_oracle.hpp"

extern "C" {
// from oracle_capi.cpp
void orc_filter_1st_pass(const float* pts, uint32_t stride, uint64_t n, const mlo_filter1_params* p, float* out_map_xyz,
                         uint64_t* out_map_n, float* out_icp_xyz, uint64_t* out_icp_n);
void orc_icp_align(void* map, const float* local, uint32_t stride, uint64_t n, const double* init_pose, const mlo_icp_params* p,
                   mlo_icp_result* out, void* pool, double* trace_poses, uint32_t* trace_pairs, uint32_t trace_cap);
void orc_filter_1st_pass_xyzt(const float* pts, uint32_t stride, const float* t, uint64_t n, const mlo_filter1_params* p,
                              float* out_map_xyzt, uint64_t* out_map_n, float* out_icp_xyzt, uint64_t* out_icp_n);
void orc_deskew(const float* xyzt, uint64_t n, const double* twist, float* out_xyz);
void* orc_map_create(const mlo_map_params* p);
void orc_map_destroy(void* m);
void orc_map_clear(void* m);
void orc_map_insert(void* m, const float* pts, uint32_t stride, uint64_t n, const double* pose);
void orc_map_cull(void* m, const double* sensor, float dist);
void orc_map_stats(void* m, uint64_t* nv, uint64_t* np);
void orc_se3_exp(const double* xi, double* pose);
void orc_se3_log(const double* pose, double* xi);
}

struct BackendOracle {
  void* create_map(const mlo_map_params& p) { return orc_map_create(&p); }
  void destroy_map(void* m) { orc_map_destroy(m); }
  void map_clear(void* m) { orc_map_clear(m); }
  void map_insert(void* m, const float* xyz, uint64_t n, const double* pose) { orc_map_insert(m, xyz, 3, n, pose); }
  void map_cull(void* m, const double* sensor, float dist) { orc_map_cull(m, sensor, dist); }
  void map_stats(void* m, uint64_t& nv, uint64_t& np) { orc_map_stats(m, &nv, &np); }
  // scan sets on the CPU: plain vectors per slot, one oracle call per job
  struct ScanSet {
    struct Slot {
      std::vector<float> map_skewed, icp_skewed, map_layer, icp_layer;  // xyzt x2, xyz x2
      bool skewed = false;
    };
    std::vector<Slot> slots;
  };
  ScanSet* scanset_create(uint32_t n_slots) {
    auto* s = new ScanSet();
    s->slots.resize(n_slots);
    return s;
  }
  void scanset_destroy(ScanSet* s) { delete s; }
  static void fill_info(const ScanSet::Slot& sl, mlo_scan_info& info) {
    info = mlo_scan_info{};
    info.n_map = sl.map_layer.size() / 3;
    info.n_icp = sl.icp_layer.size() / 3;
    const uint64_t n = info.n_icp;
    if (!n) return;
    for (int k = 0; k < 3; k++) info.icp_min[k] = info.icp_max[k] = sl.icp_layer[k];
    for (uint64_t i = 1; i < n; i++)
      for (int k = 0; k < 3; k++) {
        info.icp_min[k] = std::min(info.icp_min[k], sl.icp_layer[3 * i + k]);
        info.icp_max[k] = std::max(info.icp_max[k], sl.icp_layer[3 * i + k]);
      }
  }
  void scanset_filter(ScanSet* s, uint32_t n, const mlo_scan_job* jobs, uint32_t stride, mlo_scan_info* info) {
    for (auto& sl : s->slots) sl = ScanSet::Slot{};
    for (uint32_t j = 0; j < n; j++) {
      auto& sl = s->slots.at(jobs[j].slot);
      uint64_t na = 0, nb = 0;
      if (jobs[j].t) {
        sl.skewed = true;
        sl.map_skewed.resize(4 * jobs[j].n);
        sl.icp_skewed.resize(4 * jobs[j].n);
        orc_filter_1st_pass_xyzt(jobs[j].pts, stride, jobs[j].t, jobs[j].n, &jobs[j].fp, sl.map_skewed.data(), &na,
                                 sl.icp_skewed.data(), &nb);
        sl.map_skewed.resize(4 * na);
        sl.icp_skewed.resize(4 * nb);
        info[j] = mlo_scan_info{};
        info[j].n_map = na;
        info[j].n_icp = nb;
      } else {
        sl.map_layer.resize(3 * jobs[j].n);
        sl.icp_layer.resize(3 * jobs[j].n);
        orc_filter_1st_pass(jobs[j].pts, stride, jobs[j].n, &jobs[j].fp, sl.map_layer.data(), &na, sl.icp_layer.data(), &nb);
        sl.map_layer.resize(3 * na);
        sl.icp_layer.resize(3 * nb);
        fill_info(sl, info[j]);
      }
    }
  }
  void scanset_prefetch(ScanSet*, uint32_t, const float* const*, const uint64_t*, uint32_t) {}  // nothing to overlap on the CPU
  void scanset_deskew(ScanSet* s, uint32_t n, const uint32_t* slots, const double* twists6, mlo_scan_info* info) {
    for (uint32_t i = 0; i < n; i++) {
      auto& sl = s->slots.at(slots[i]);
      sl.map_layer.resize(3 * (sl.map_skewed.size() / 4));
      sl.icp_layer.resize(3 * (sl.icp_skewed.size() / 4));
      orc_deskew(sl.map_skewed.data(), sl.map_skewed.size() / 4, twists6 + 6 * size_t(i), sl.map_layer.data());
      orc_deskew(sl.icp_skewed.data(), sl.icp_skewed.size() / 4, twists6 + 6 * size_t(i), sl.icp_layer.data());
      fill_info(sl, info[i]);
    }
  }
  void scanset_align(ScanSet* s, uint32_t n, const mlo_align_job* jobs, mlo_icp_result* out) {
    for (uint32_t j = 0; j < n; j++) {
      const auto& sl = s->slots.at(jobs[j].slot);
      orc_icp_align(const_cast<void*>(reinterpret_cast<const void*>(jobs[j].map)), sl.icp_layer.data(), 3, sl.icp_layer.size() / 3,
                    jobs[j].init_pose_3x4, &jobs[j].params, &out[j], nullptr, nullptr, nullptr, 0);
    }
  }
  void scanset_insert(ScanSet* s, uint32_t n, const mlo_insert_job* jobs, mlo_map_counts* out) {
    for (uint32_t j = 0; j < n; j++) {
      const auto& sl = s->slots.at(jobs[j].slot);
      void* m = reinterpret_cast<void*>(jobs[j].map);
      orc_map_insert(m, sl.map_layer.data(), 3, sl.map_layer.size() / 3, jobs[j].pose_3x4);
      if (jobs[j].cull_farther_than > 0.f) {
        const double sxyz[3] = {jobs[j].pose_3x4[3], jobs[j].pose_3x4[7], jobs[j].pose_3x4[11]};
        orc_map_cull(m, sxyz, jobs[j].cull_farther_than);
      }
      orc_map_stats(m, &out[j].n_voxels, &out[j].n_points);
    }
  }
  void se3_exp(const double* xi, double* pose) { orc_se3_exp(xi, pose); }
  void se3_log(const double* pose, double* xi) { orc_se3_log(pose, xi); }
};
