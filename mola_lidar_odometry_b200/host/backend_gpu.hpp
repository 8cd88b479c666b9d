// backend_gpu.hpp — the product backend of the host layer: every operation is a call into the CUDA C ABI
// (include/mlo_b200.h).  Negative status codes become exceptions, as the reference's MRPT assertions would.
#pragma once
#include <stdexcept>
#include <string>
#include <vector>

#include "mlo_b200.h"

namespace mlo_host {

struct BackendGpu {
  mlo_ctx* ctx = nullptr;
  explicit BackendGpu(mlo_ctx* c) : ctx(c) {}
  void check(int rc) const {
    if (rc != MLO_OK) throw std::runtime_error(std::string("mlo_b200: ") + mlo_last_error(ctx));
  }
  void* create_map(const mlo_map_params& p) {
    mlo_map* m = nullptr;
    check(mlo_map_create(ctx, &p, &m));
    return m;
  }
  void destroy_map(void* m) { mlo_map_destroy(static_cast<mlo_map*>(m)); }
  void map_clear(void* m) { check(mlo_map_clear(static_cast<mlo_map*>(m))); }
  void map_insert(void* m, const float* xyz, uint64_t n, const double* pose) {
    check(mlo_map_insert(static_cast<mlo_map*>(m), xyz, 3, n, pose));
  }
  void map_cull(void* m, const double* sensor, float dist) { check(mlo_map_cull(static_cast<mlo_map*>(m), sensor, dist)); }
  void map_stats(void* m, uint64_t& nv, uint64_t& np) { check(mlo_map_stats(static_cast<mlo_map*>(m), &nv, &np)); }
  void filter_1st_pass(const float* pts, uint32_t stride, uint64_t n, const mlo_filter1_params& f, std::vector<float>& map_xyz,
                       std::vector<float>& icp_xyz) {
    map_xyz.resize(3 * n);
    icp_xyz.resize(3 * n);
    uint64_t na = 0, nb = 0;
    check(mlo_filter_1st_pass(ctx, pts, stride, n, &f, map_xyz.data(), &na, icp_xyz.data(), &nb));
    map_xyz.resize(3 * na);
    icp_xyz.resize(3 * nb);
  }
  void filter_1st_pass_xyzt(const float* pts, uint32_t stride, const float* t, uint64_t n, const mlo_filter1_params& f,
                            std::vector<float>& map_xyzt, std::vector<float>& icp_xyzt) {
    map_xyzt.resize(4 * n);
    icp_xyzt.resize(4 * n);
    uint64_t na = 0, nb = 0;
    check(mlo_filter_1st_pass_xyzt(ctx, pts, stride, t, n, &f, map_xyzt.data(), &na, icp_xyzt.data(), &nb));
    map_xyzt.resize(4 * na);
    icp_xyzt.resize(4 * nb);
  }
  void deskew(const float* xyzt, uint64_t n, const double* twist, std::vector<float>& out_xyz) {
    out_xyz.resize(3 * n);
    check(mlo_deskew(ctx, xyzt, n, twist, out_xyz.data()));
  }
  void icp_align(const float* xyz, uint64_t n, void* map, const double* init, const mlo_icp_params& p, mlo_icp_result& r) {
    check(mlo_icp_align(ctx, xyz, 3, n, static_cast<mlo_map*>(map), init, &p, &r));
  }
  void se3_exp(const double* xi, double* pose) { mlo_se3_exp(xi, pose); }
  void se3_log(const double* pose, double* xi) { mlo_se3_log(pose, xi); }
};

}  // namespace mlo_host
