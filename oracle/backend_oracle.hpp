// oracle/backend_oracle.hpp — TEST INFRASTRUCTURE: runs the product's host orchestrator
// (mola_lidar_odometry_b200/host/pipeline.hpp) over the CPU oracle, so tests can compare whole trajectories
// (caller contract + hot path) GPU vs CPU with identical host logic.  parity unpinned — see mlo_oracle.hpp.
#pragma once
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
  void filter_1st_pass(const float* pts, uint32_t stride, uint64_t n, const mlo_filter1_params& f, std::vector<float>& a,
                       std::vector<float>& b) {
    a.resize(3 * n);
    b.resize(3 * n);
    uint64_t na = 0, nb = 0;
    orc_filter_1st_pass(pts, stride, n, &f, a.data(), &na, b.data(), &nb);
    a.resize(3 * na);
    b.resize(3 * nb);
  }
  void filter_1st_pass_xyzt(const float* pts, uint32_t stride, const float* t, uint64_t n, const mlo_filter1_params& f,
                            std::vector<float>& a, std::vector<float>& b) {
    a.resize(4 * n);
    b.resize(4 * n);
    uint64_t na = 0, nb = 0;
    orc_filter_1st_pass_xyzt(pts, stride, t, n, &f, a.data(), &na, b.data(), &nb);
    a.resize(4 * na);
    b.resize(4 * nb);
  }
  void deskew(const float* xyzt, uint64_t n, const double* twist, std::vector<float>& out_xyz) {
    out_xyz.resize(3 * n);
    orc_deskew(xyzt, n, twist, out_xyz.data());
  }
  void icp_align(const float* xyz, uint64_t n, void* map, const double* init, const mlo_icp_params& p, mlo_icp_result& r) {
    orc_icp_align(map, xyz, 3, n, init, &p, &r, nullptr, nullptr, nullptr, 0);
  }
  void se3_exp(const double* xi, double* pose) { orc_se3_exp(xi, pose); }
  void se3_log(const double* pose, double* xi) { orc_se3_log(pose, xi); }
};
