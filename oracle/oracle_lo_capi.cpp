// oracle/oracle_lo_capi.cpp — TEST INFRASTRUCTURE: the host orchestrator instantiated over the CPU oracle
// (orc_lo_* mirrors mlo_lo_* of include/mlo_b200_host.h).  Built into oracle/liboracle.so.
#include <cstring>
#include <memory>
#include <string>
#include <vector>

#include "../include/mlo_b200_host.h"
#include "../mola_lidar_odometry_b200/host/pipeline.hpp"
#include "backend_oracle.hpp"

using namespace mlo_host;

struct orc_lo {
  BackendOracle be;
  LidarOdometryT<BackendOracle> lo;
  std::string err;
  orc_lo() : lo(be) {}
};

struct orc_fleet {
  BackendOracle be;
  LidarOdometryFleetT<BackendOracle> fleet;
  std::vector<ScanOutput> tmp;
  explicit orc_fleet(uint32_t n) : fleet(be, n), tmp(n) {}
};

static void put_output(const ScanOutput& s, mlo_lo_scan_output* out) {
  out->processed = s.processed;
  out->icp_ran = s.icp_ran;
  out->icp_good = s.icp_good;
  out->map_updated = s.map_updated;
  std::memcpy(out->pose_3x4, s.pose.data(), sizeof(out->pose_3x4));
  out->quality = s.quality;
  out->sigma = s.sigma;
  out->est_max_range = s.est_max_range;
  out->icp_iterations = s.icp_iterations;
  out->icp_runs = s.icp_runs;
  out->termination = s.termination;
  out->n_map_layer = s.n_map_layer;
  out->n_icp_layer = s.n_icp_layer;
  out->icp_had_prior = s.icp_had_prior;
  out->has_motion_model = s.has_motion_model;
  out->prior_info_trace = s.prior_info_trace;
}

extern "C" {
// the fleet orchestrator over the oracle backend (mirrors mlo_fleet_* of include/mlo_b200_host.h)
void* orc_fleet_create(const char* yaml, int is_text, uint32_t n) {
  try {
    auto o = std::make_unique<orc_fleet>(n);
    o->fleet.initialize(is_text ? yaml_parse(yaml) : yaml_load_file(yaml));
    return o.release();
  } catch (const std::exception& e) {
    std::fprintf(stderr, "orc_fleet_create: %s\n", e.what());
    return nullptr;
  }
}
void orc_fleet_destroy(void* h) { delete static_cast<orc_fleet*>(h); }
int orc_fleet_on_lidar(void* h, const float* const* pts, uint32_t stride, const uint64_t* n, const double* stamps,
                       const float* const* t, mlo_lo_scan_output* out) {
  auto* o = static_cast<orc_fleet*>(h);
  try {
    o->fleet.onLidarBatch(pts, stride, n, stamps, t, o->tmp.data());
    for (uint32_t i = 0; i < o->fleet.size(); i++) put_output(o->tmp[i], &out[i]);
    return 0;
  } catch (const std::exception& e) {
    std::fprintf(stderr, "orc_fleet_on_lidar: %s\n", e.what());
    return -1;
  }
}

void* orc_lo_create(const char* yaml, int is_text) {
  try {
    auto o = std::make_unique<orc_lo>();
    o->lo.initialize(is_text ? yaml_parse(yaml) : yaml_load_file(yaml));
    return o.release();
  } catch (const std::exception& e) {
    std::fprintf(stderr, "orc_lo_create: %s\n", e.what());
    return nullptr;
  }
}
void orc_lo_destroy(void* h) { delete static_cast<orc_lo*>(h); }
int orc_lo_on_lidar_t(void* h, const float* pts, uint32_t stride, const float* t, uint64_t n, double stamp, mlo_lo_scan_output* out);
int orc_lo_on_lidar(void* h, const float* pts, uint32_t stride, uint64_t n, double stamp, mlo_lo_scan_output* out) {
  return orc_lo_on_lidar_t(h, pts, stride, nullptr, n, stamp, out);
}
int orc_lo_on_lidar_t(void* h, const float* pts, uint32_t stride, const float* t, uint64_t n, double stamp, mlo_lo_scan_output* out) {
  auto* o = static_cast<orc_lo*>(h);
  try {
    put_output(o->lo.onLidar(pts, stride, n, stamp, t), out);
    return 0;
  } catch (const std::exception& e) {
    std::fprintf(stderr, "orc_lo_on_lidar: %s\n", e.what());
    return -1;
  }
}
}
