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
  // scan sets: the layers of an observation stay in HBM between the filter, align and insert calls
  using ScanSet = mlo_scanset;
  ScanSet* scanset_create(uint32_t n_slots) {
    mlo_scanset* s = nullptr;
    check(mlo_scanset_create(ctx, n_slots, &s));
    return s;
  }
  void scanset_destroy(ScanSet* s) { mlo_scanset_destroy(s); }
  void scanset_filter(ScanSet* s, uint32_t n, const mlo_scan_job* jobs, uint32_t stride, mlo_scan_info* info) {
    check(mlo_scanset_filter(s, n, jobs, stride, info));
  }
  void scanset_prefetch(ScanSet* s, uint32_t n, const float* const* pts, const uint64_t* np, uint32_t stride) {
    check(mlo_scanset_prefetch(s, n, pts, np, stride));
  }
  void scanset_deskew(ScanSet* s, uint32_t n, const uint32_t* slots, const double* twists6, mlo_scan_info* info) {
    check(mlo_scanset_deskew(s, n, slots, twists6, info));
  }
  void scanset_align(ScanSet* s, uint32_t n, const mlo_align_job* jobs, mlo_icp_result* out) {
    check(mlo_scanset_align(s, n, jobs, out));
  }
  void scanset_insert(ScanSet* s, uint32_t n, const mlo_insert_job* jobs, mlo_map_counts* out) {
    check(mlo_scanset_insert(s, n, jobs, out));
  }
  void se3_exp(const double* xi, double* pose) { mlo_se3_exp(xi, pose); }
  void se3_log(const double* pose, double* xi) { mlo_se3_log(pose, xi); }
};

}  // namespace mlo_host
