// host_capi.cpp — extern "C" surface (include/mlo_b200_host.h) of the C++ host layer over the GPU backend.
#include <cstring>
#include <memory>
#include <string>
#include <vector>

#include "backend_gpu.hpp"
#include "mlo_b200_host.h"
#include "pipeline.hpp"

using namespace mlo_host;

struct mlo_lo {
  BackendGpu be;
  LidarOdometryT<BackendGpu> lo;
  std::string err;
  YamlNode cfg;
  explicit mlo_lo(mlo_ctx* c) : be(c), lo(be) {}
};

struct mlo_fleet {
  BackendGpu be;
  LidarOdometryFleetT<BackendGpu> fleet;
  std::string err;
  std::vector<ScanOutput> tmp;
  mlo_fleet(mlo_ctx* c, uint32_t n) : be(c), fleet(be, n), tmp(n) {}
};

namespace {
void put_output(const ScanOutput& s, mlo_lo_scan_output* out) {
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
thread_local std::string g_err;
YamlNode load_cfg(const char* yaml, int is_text) { return is_text ? yaml_parse(yaml) : yaml_load_file(yaml); }
}  // namespace

extern "C" {

int mlo_lo_create(mlo_ctx* ctx, const char* yaml, int is_text, mlo_lo** out) {
  if (!ctx || !yaml || !out) return MLO_ERR_INVALID_ARG;
  *out = nullptr;
  try {
    auto lo = std::make_unique<mlo_lo>(ctx);
    lo->cfg = load_cfg(yaml, is_text);
    lo->lo.initialize(lo->cfg);
    *out = lo.release();
    return MLO_OK;
  } catch (const std::exception& e) {
    g_err = e.what();
    return MLO_ERR_INVALID_ARG;
  }
}
void mlo_lo_destroy(mlo_lo* lo) { delete lo; }
const char* mlo_lo_last_error(const mlo_lo* lo) { return lo ? lo->err.c_str() : g_err.c_str(); }

int mlo_lo_on_lidar(mlo_lo* lo, const float* pts, uint32_t stride, uint64_t n, double stamp, mlo_lo_scan_output* out) {
  return mlo_lo_on_lidar_t(lo, pts, stride, nullptr, n, stamp, out);
}

int mlo_lo_on_lidar_t(mlo_lo* lo, const float* pts, uint32_t stride, const float* t, uint64_t n, double stamp,
                      mlo_lo_scan_output* out) {
  if (!lo || !out || (n && !pts) || (stride != 3 && stride != 4)) return MLO_ERR_INVALID_ARG;
  try {
    put_output(lo->lo.onLidar(pts, stride, n, stamp, t), out);
    return MLO_OK;
  } catch (const std::exception& e) {
    lo->err = e.what();  // the reference latches this as fatal_error (LidarOdometry.cpp:614-619)
    return MLO_ERR_CUDA;
  }
}

int mlo_lo_trajectory(const mlo_lo* lo, double* stamps, double* poses, uint64_t max_n, uint64_t* n) {
  if (!lo || !n) return MLO_ERR_INVALID_ARG;
  const auto& t = lo->lo.estimatedTrajectory();
  *n = t.size();
  if (!stamps || !poses) return MLO_OK;
  for (uint64_t i = 0; i < t.size() && i < max_n; i++) {
    stamps[i] = t[i].first;
    std::memcpy(poses + 12 * i, t[i].second.data(), 12 * sizeof(double));
  }
  return MLO_OK;
}

int mlo_lo_reset(mlo_lo* lo) {
  if (!lo) return MLO_ERR_INVALID_ARG;
  try {
    lo->lo.initialize(lo->cfg);
    return MLO_OK;
  } catch (const std::exception& e) {
    lo->err = e.what();
    return MLO_ERR_INVALID_ARG;
  }
}

int mlo_fleet_create(mlo_ctx* ctx, const char* yaml, int is_text, uint32_t n_sequences, mlo_fleet** out) {
  if (!ctx || !yaml || !out || n_sequences == 0) return MLO_ERR_INVALID_ARG;
  *out = nullptr;
  try {
    auto f = std::make_unique<mlo_fleet>(ctx, n_sequences);
    f->fleet.initialize(load_cfg(yaml, is_text));
    *out = f.release();
    return MLO_OK;
  } catch (const std::exception& e) {
    g_err = e.what();
    return MLO_ERR_INVALID_ARG;
  }
}
void mlo_fleet_destroy(mlo_fleet* f) { delete f; }
const char* mlo_fleet_last_error(const mlo_fleet* f) { return f ? f->err.c_str() : g_err.c_str(); }

int mlo_fleet_on_lidar(mlo_fleet* f, const float* const* pts, uint32_t stride, const uint64_t* n, const double* stamps,
                       const float* const* t, mlo_lo_scan_output* out) {
  if (!f || !pts || !n || !stamps || !out || (stride != 3 && stride != 4)) return MLO_ERR_INVALID_ARG;
  try {
    f->fleet.onLidarBatch(pts, stride, n, stamps, t, f->tmp.data());
    for (uint32_t i = 0; i < f->fleet.size(); i++) put_output(f->tmp[i], &out[i]);
    return MLO_OK;
  } catch (const std::exception& e) {
    f->err = e.what();
    return MLO_ERR_CUDA;
  }
}

int mlo_fleet_prefetch(mlo_fleet* f, const float* const* pts, uint32_t stride, const uint64_t* n) {
  if (!f || !pts || !n || (stride != 3 && stride != 4)) return MLO_ERR_INVALID_ARG;
  try {
    f->fleet.prefetch(pts, stride, n);
    return MLO_OK;
  } catch (const std::exception& e) {
    f->err = e.what();
    return MLO_ERR_CUDA;
  }
}

int mlo_fleet_phase_times(mlo_fleet* f, double out_ms[8], int reset) {
  if (!f || !out_ms) return MLO_ERR_INVALID_ARG;
  for (int k = 0; k < 8; k++) out_ms[k] = f->fleet.phase_ms[k];
  if (reset) f->fleet.phase_ms.fill(0.0);
  return MLO_OK;
}

int mlo_fleet_trajectory(const mlo_fleet* f, uint32_t sequence, double* stamps, double* poses, uint64_t max_n, uint64_t* n) {
  if (!f || !n || sequence >= const_cast<mlo_fleet*>(f)->fleet.size()) return MLO_ERR_INVALID_ARG;
  const auto& t = const_cast<mlo_fleet*>(f)->fleet.sequence(sequence).estimatedTrajectory();
  *n = t.size();
  if (!stamps || !poses) return MLO_OK;
  for (uint64_t i = 0; i < t.size() && i < max_n; i++) {
    stamps[i] = t[i].first;
    std::memcpy(poses + 12 * i, t[i].second.data(), 12 * sizeof(double));
  }
  return MLO_OK;
}

const char* mlo_host_last_error(void) { return g_err.c_str(); }

int mlo_host_icp_tables(const char* yaml_text, double sigma, uint32_t n_it, double* t1, double* t2, double* t3, mlo_icp_params* sc) {
  try {
    const YamlNode cfg = yaml_parse(yaml_text);
    const YamlNode& icp = cfg.has("icp_settings_with_vel") ? cfg["icp_settings_with_vel"] : cfg;
    struct Dummy {};
    ICP<Dummy> o;  // only the YAML -> formula part is used here
    o.initialize(icp);
    ParameterSource ps;
    ps.updateVariable("ADAPTIVE_THRESHOLD_SIGMA", sigma);
    for (uint32_t it = 0; it < n_it; it++) {
      ps.updateVariable("ICP_ITERATION", it);
      if (t1) t1[it] = o.m_pt2pt ? o.m_pt2pt->threshold.eval(ps) : 0.0;
      if (t2) t2[it] = o.m_pt2pl ? o.m_pt2pl->distanceThreshold.eval(ps) : 0.0;
      if (t3) t3[it] = o.gn ? o.gn->robustKernelParam.eval(ps) : 0.0;
    }
    if (sc) {
      std::memset(sc, 0, sizeof(*sc));
      sc->max_iterations = o.params.maxIterations;
      sc->min_abs_step_trans = o.params.minAbsStep_trans;
      sc->min_abs_step_rot = o.params.minAbsStep_rot;
      sc->solver = o.gn ? MLO_SOLVER_GAUSS_NEWTON : MLO_SOLVER_HORN;
      sc->gn_max_iterations = o.gn ? o.gn->maxIterations : 0;
      sc->robust_kernel = o.gn ? o.gn->robustKernel : 0;
      sc->matcher_mask = (o.m_pt2pt ? MLO_MATCHER_PT2PT : 0u) | (o.m_pt2pl ? MLO_MATCHER_PT2PL : 0u);
      sc->threshold_angular_deg = o.m_pt2pt ? o.m_pt2pt->thresholdAngularDeg : 0;
    }
    return MLO_OK;
  } catch (const std::exception& e) {
    g_err = e.what();
    return MLO_ERR_INVALID_ARG;
  }
}

int mlo_host_filter1(const char* yaml_text, double est, double inst, mlo_filter1_params* out) {
  try {
    const YamlNode cfg = yaml_parse(yaml_text);
    FilterPipeline1st f;
    f.initialize(cfg.at("observations_filter_1st_pass"));
    ParameterSource ps;
    ps.updateVariable("ESTIMATED_SENSOR_MAX_RANGE", est);
    ps.updateVariable("INSTANTANEOUS_SENSOR_MAX_RANGE", inst);
    *out = f.realize(ps);
    return MLO_OK;
  } catch (const std::exception& e) {
    g_err = e.what();
    return MLO_ERR_INVALID_ARG;
  }
}

int mlo_host_mapdef(const char* yaml_text, double est, mlo_map_params* out, float* cull) {
  try {
    const YamlNode cfg = yaml_parse(yaml_text);
    LocalMapDefinition d;
    bool found = false;
    for (const YamlNode& g : cfg.at("localmap_generator").seq)
      if (g["params"].has("metric_map_definition")) {
        d.initialize(g["params"]["metric_map_definition"]);
        found = true;
      }
    if (!found) throw std::runtime_error("no metric_map_definition");
    ParameterSource ps;
    ps.updateVariable("ESTIMATED_SENSOR_MAX_RANGE", est);
    std::memset(out, 0, sizeof(*out));
    out->kind = d.kind;
    out->voxel_size = float(d.voxel_size.eval(ps));
    out->max_points_per_voxel = d.max_points_per_voxel;
    out->min_distance_between_points = float(d.min_distance_between_points);
    out->max_eigen_ratio_for_planes = float(d.max_eigen_ratio_for_planes);
    out->min_points_for_plane = 5;
    out->capacity_voxels = d.capacity_voxels;
    if (cull) *cull = float(d.remove_voxels_farther_than.eval(ps));
    return MLO_OK;
  } catch (const std::exception& e) {
    g_err = e.what();
    return MLO_ERR_INVALID_ARG;
  }
}

int mlo_host_eval_formula(const char* expr, const char* const* names, const double* values, uint32_t n, double* out) {
  try {
    ParameterSource ps;
    for (uint32_t i = 0; i < n; i++) ps.updateVariable(names[i], values[i]);
    *out = Formula(expr).eval(ps);
    return MLO_OK;
  } catch (const std::exception& e) {
    g_err = e.what();
    return MLO_ERR_INVALID_ARG;
  }
}

}  // extern "C"
