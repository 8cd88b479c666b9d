// pipeline.hpp — C++ host mirror of the reference's plugin surface for the hot path, over a backend.
//
// Mirrors (same names / YAML keys / argument meaning / error behaviour, re-implemented from the call sites):
//   mp2p_icp::ICP + Parameters + Results          pipelines/lidar3d-default.yaml:169-182, LidarOdometry.cpp:961-962
//   mp2p_icp::Matcher_Points_DistanceThreshold    default.yaml:196-204
//   mp2p_icp::Matcher_Point2Plane                 lidar3d-ndt.yaml:195-200
//   mp2p_icp::Solver_GaussNewton / Solver_Horn    default.yaml:185-190, extras/icp-pipeline_no_motion_model.yaml:24-29
//   mp2p_icp::QualityEvaluator_PairedRatio        default.yaml:206-209
//   mp2p_icp_filters::FilterDecimateVoxels / FilterByRange / FilterBoundingBox   default.yaml:285-319
//   mola::HashedVoxelPointCloud / mola::NDT       default.yaml:228-242, ndt.yaml:234-254
//   mola::LidarOdometry::onLidarImpl (caller contract of the hot path)           LidarOdometry.cpp:627-1206
// The Backend supplies the arithmetic: BackendGpu (backend_gpu.hpp, the CUDA C ABI) in the product;
// oracle/backend_oracle.hpp exists only so tests can run the SAME orchestrator over the CPU oracle.
// Errors: std::runtime_error, like the reference's MRPT exceptions that its worker latches (LidarOdometry.cpp:614-619).
#pragma once
#include <algorithm>
#include <array>
#include <atomic>
#include <condition_variable>
#include <functional>
#include <mutex>
#include <thread>
#include <chrono>
#include <cmath>
#include <cstring>
#include <memory>
#include <optional>
#include <string>
#include <vector>

#include <cstdlib>

#include "formula.hpp"
#include "navstate_fuse.hpp"
#include "mlo_b200.h"
#include "yaml_lite.hpp"

namespace mlo_host {

using Pose = std::array<double, 12>;  // 3x4 row-major [R|t]

inline Pose pose_identity() { return {1, 0, 0, 0, 0, 1, 0, 0, 0, 0, 1, 0}; }
inline Pose pose_compose(const Pose& a, const Pose& b) {
  Pose c{};
  for (int r = 0; r < 3; r++) {
    for (int k = 0; k < 3; k++) c[4 * r + k] = a[4 * r] * b[k] + a[4 * r + 1] * b[4 + k] + a[4 * r + 2] * b[8 + k];
    c[4 * r + 3] = a[4 * r] * b[3] + a[4 * r + 1] * b[7] + a[4 * r + 2] * b[11] + a[4 * r + 3];
  }
  return c;
}
inline Pose pose_inverse(const Pose& a) {
  Pose c{};
  for (int r = 0; r < 3; r++)
    for (int k = 0; k < 3; k++) c[4 * r + k] = a[4 * k + r];
  for (int r = 0; r < 3; r++) c[4 * r + 3] = -(c[4 * r] * a[3] + c[4 * r + 1] * a[7] + c[4 * r + 2] * a[11]);
  return c;
}
inline Pose pose_minus(const Pose& a, const Pose& b) { return pose_compose(pose_inverse(b), a); }  // MRPT "a - b"
inline double norm3(const double* v) { return std::sqrt(v[0] * v[0] + v[1] * v[1] + v[2] * v[2]); }
// yaw/pitch/roll of the rotation block (MRPT CPose3D convention), for the robot_yaw/pitch/roll variables
inline void pose_ypr(const Pose& p, double& yaw, double& pitch, double& roll) {
  pitch = std::atan2(-p[8], std::hypot(p[0], p[4]));
  yaw = std::atan2(p[4], p[0]);
  roll = std::atan2(p[9], p[10]);
}

// ------------------------------------------------------------------------------------------------ plugin params
struct Matcher_Points_DistanceThreshold {
  Formula threshold{"1.0"};
  double thresholdAngularDeg = 0;
  int pairingsPerPoint = 1;
  bool allowMatchAlreadyMatchedGlobalPoints = true;
  double weight = 1.0;
  std::string global_layer = "localmap", local_layer = "decimated_for_icp";
  void initialize(const YamlNode& p) {
    threshold = Formula(p.at("threshold").str());
    thresholdAngularDeg = p["thresholdAngularDeg"].num(0);
    pairingsPerPoint = int(p["pairingsPerPoint"].num(1));
    allowMatchAlreadyMatchedGlobalPoints = p["allowMatchAlreadyMatchedGlobalPoints"].boolean(true);
    if (pairingsPerPoint != 1) throw std::runtime_error("Matcher_Points_DistanceThreshold: only pairingsPerPoint=1 is supported");
    const YamlNode& lm = p["pointLayerMatches"];
    if (lm.isSeq() && !lm.seq.empty()) {
      global_layer = lm.seq[0]["global"].str_or(global_layer);
      local_layer = lm.seq[0]["local"].str_or(local_layer);
      weight = lm.seq[0]["weight"].num(1.0);
    }
  }
};
struct Matcher_Point2Plane {
  Formula distanceThreshold{"1.0"};
  double weight = 1.0;
  void initialize(const YamlNode& p) {
    distanceThreshold = Formula(p.at("distanceThreshold").str());
    const YamlNode& lm = p["pointLayerMatches"];
    if (lm.isSeq() && !lm.seq.empty()) weight = lm.seq[0]["weight"].num(1.0);
  }
};
struct Solver_GaussNewton {
  uint32_t maxIterations = 2;
  int robustKernel = MLO_KERNEL_NONE;
  Formula robustKernelParam{"1.0"};
  double minDelta = 1e-7;
  void initialize(const YamlNode& p) {
    maxIterations = uint32_t(p["maxIterations"].num(2));
    const std::string k = p["robustKernel"].str_or("RobustKernel::None");
    if (k == "RobustKernel::GemanMcClure") robustKernel = MLO_KERNEL_GEMAN_MCCLURE;
    else if (k == "RobustKernel::Cauchy") robustKernel = MLO_KERNEL_CAUCHY;
    else if (k == "RobustKernel::None") robustKernel = MLO_KERNEL_NONE;
    else throw std::runtime_error("Solver_GaussNewton: unknown robustKernel '" + k + "'");
    if (p.has("robustKernelParam")) robustKernelParam = Formula(p["robustKernelParam"].str());
    minDelta = p["minDelta"].num(1e-7);
  }
};
struct Solver_Horn {
  double runUntilTranslationCorrectionSmallerThan = 0;
  void initialize(const YamlNode& p) {
    runUntilTranslationCorrectionSmallerThan = p["runUntilTranslationCorrectionSmallerThan"].num(0);
  }
};
struct QualityEvaluator_PairedRatio {
  void initialize(const YamlNode&) {}  // no parameters required (default.yaml:208-209): reuses the ICP pairings
};

struct Parameters {  // mp2p_icp::Parameters (default.yaml:172-182)
  uint32_t maxIterations = 300;
  double minAbsStep_trans = 1e-4, minAbsStep_rot = 5e-5;
};
struct Results {  // mp2p_icp::Results as consumed at LidarOdometry.cpp:964-1011
  Pose optimal_tf_mean = pose_identity();
  std::array<double, 36> optimal_tf_cov{};
  double quality = 0;
  uint32_t nIterations = 0;
  int terminationReason = MLO_TERM_UNDEFINED;
  uint64_t nPairings = 0;
};
struct IterationHook {  // the lambda of LidarOdometry.cpp:923-952 expressed as data
  bool enabled = false;
  Pose checkpoint = pose_identity();
  double min_trans = 0.15, min_rot_rad = 0.75 * M_PI / 180.0;
};
struct Prior {
  Pose mean = pose_identity();
  std::array<double, 36> cov_inv{};
};

template <class Backend>
class ICP {
 public:
  Parameters params;
  std::optional<Solver_GaussNewton> gn;
  std::optional<Solver_Horn> horn;
  std::optional<Matcher_Points_DistanceThreshold> m_pt2pt;
  std::optional<Matcher_Point2Plane> m_pt2pl;
  QualityEvaluator_PairedRatio quality;
  IterationHook hook;
  ParameterSource* source = nullptr;

  // mp2p_icp::icp_pipeline_from_yaml (LidarOdometry.cpp:115-122)
  void initialize(const YamlNode& n) {
    const std::string cn = n["class_name"].str_or("mp2p_icp::ICP");
    if (cn != "mp2p_icp::ICP" && cn != "mlo_b200::ICP") throw std::runtime_error("unsupported ICP class_name '" + cn + "'");
    const YamlNode& p = n["params"];
    params.maxIterations = uint32_t(p["maxIterations"].num(300));
    params.minAbsStep_trans = p["minAbsStep_trans"].num(1e-4);
    params.minAbsStep_rot = p["minAbsStep_rot"].num(5e-5);
    for (const YamlNode& s : n.at("solvers").seq) {
      const std::string c = s.at("class").str();
      if (c == "mp2p_icp::Solver_GaussNewton") { gn.emplace(); gn->initialize(s["params"]); }
      else if (c == "mp2p_icp::Solver_Horn") { horn.emplace(); horn->initialize(s["params"]); }
      else throw std::runtime_error("unsupported solver class '" + c + "'");
    }
    for (const YamlNode& m : n.at("matchers").seq) {
      const std::string c = m.at("class").str();
      if (c == "mp2p_icp::Matcher_Points_DistanceThreshold") { m_pt2pt.emplace(); m_pt2pt->initialize(m["params"]); }
      else if (c == "mp2p_icp::Matcher_Point2Plane") { m_pt2pl.emplace(); m_pt2pl->initialize(m["params"]); }
      else throw std::runtime_error("unsupported matcher class '" + c + "'");
    }
    for (const YamlNode& q : n["quality"].seq) {
      if (q.at("class").str() != "mp2p_icp::QualityEvaluator_PairedRatio")
        throw std::runtime_error("unsupported quality evaluator '" + q["class"].str() + "'");
      quality.initialize(q["params"]);
    }
    if (!gn && !horn) throw std::runtime_error("ICP: no solver defined");
    if (!m_pt2pt && !m_pt2pl) throw std::runtime_error("ICP: no matcher defined");
  }
  void attachToParameterSource(ParameterSource& ps) { source = &ps; }
  void setIterationHook(const IterationHook& h) { hook = h; }

  // mp2p_icp::ICP::align(local, global, init, params, result, prior) — call site LidarOdometry.cpp:961-962 — is split in
  // two so that a fleet of sequences can put the aligns of one lock step into ONE device pass: make_params() realises
  // the plugin objects into the POD the C ABI takes (tables live in this object until the call returns),
  // read_result() turns the POD result into mp2p_icp::Results.
  void make_params(const Parameters& p, const std::optional<Prior>& prior, mlo_icp_params& q) {
    if (!source) throw std::runtime_error("ICP::align: not attached to a ParameterSource");
    std::memset(&q, 0, sizeof(q));
    q.max_iterations = p.maxIterations;
    q.min_abs_step_trans = p.minAbsStep_trans;
    q.min_abs_step_rot = p.minAbsStep_rot;
    q.solver = (gn ? MLO_SOLVER_GAUSS_NEWTON : MLO_SOLVER_HORN);
    q.gn_max_iterations = gn ? gn->maxIterations : 1;
    q.gn_min_delta = gn ? gn->minDelta : 1e-7;
    q.robust_kernel = gn ? gn->robustKernel : MLO_KERNEL_NONE;
    q.matcher_mask = (m_pt2pt ? MLO_MATCHER_PT2PT : 0u) | (m_pt2pl ? MLO_MATCHER_PT2PL : 0u);
    q.pt2pt_weight = m_pt2pt ? m_pt2pt->weight : 1.0;
    q.pt2pl_weight = m_pt2pl ? m_pt2pl->weight : 1.0;
    q.threshold_angular_deg = m_pt2pt ? m_pt2pt->thresholdAngularDeg : 0.0;
    // formulas are re-realised per ICP_ITERATION: tabulate them
    // (the device reads entry min(it, len - 1): 64 entries cover every formula that has settled by then - the shipped ones
    // are constant from iteration 20 on; one that still moves at entry 63 is tabulated up to maxIterations)
    const uint32_t full = std::max<uint32_t>(1, std::min<uint32_t>(p.maxIterations, 300));
    uint32_t len = std::min<uint32_t>(full, 64);
    t1_.assign(full, 0.0);
    t2_.assign(full, 0.0);
    t3_.assign(full, 1.0);
    if (!source->has("ICP_ITERATION")) source->updateVariable("ICP_ITERATION", 0.0);
    const int it_slot = source->slot("ICP_ITERATION");
    const double saved_it = source->at(it_slot);
    for (uint32_t it = 0; it < len; it++) {
      source->set(it_slot, double(it));
      if (m_pt2pt) t1_[it] = m_pt2pt->threshold.eval(*source);
      if (m_pt2pl) t2_[it] = m_pt2pl->distanceThreshold.eval(*source);
      if (gn) t3_[it] = gn->robustKernelParam.eval(*source);
      if (it + 1 == len && len < full && len >= 2 &&
          (t1_[it] != t1_[it - 1] || t2_[it] != t2_[it - 1] || t3_[it] != t3_[it - 1]))
        len = full;  // still changing: no truncation
    }
    t1_.resize(len);
    t2_.resize(len);
    t3_.resize(len);
    source->set(it_slot, saved_it);
    q.table_len = len;
    q.pt2pt_threshold_by_iter = t1_.data();
    q.pt2pl_threshold_by_iter = t2_.data();
    q.kernel_param_by_iter = t3_.data();
    const Pose I = pose_identity();
    std::memcpy(q.prior_pose_3x4, I.data(), sizeof(q.prior_pose_3x4));
    if (prior) {
      q.has_prior = 1;
      std::memcpy(q.prior_pose_3x4, prior->mean.data(), sizeof(q.prior_pose_3x4));
      std::memcpy(q.prior_info_6x6, prior->cov_inv.data(), sizeof(q.prior_info_6x6));
    }
    q.hook_enabled = hook.enabled ? 1 : 0;
    q.hook_min_trans = hook.min_trans;
    q.hook_min_rot_rad = hook.min_rot_rad;
    std::memcpy(q.hook_checkpoint_pose_3x4, hook.checkpoint.data(), sizeof(q.hook_checkpoint_pose_3x4));
  }
  static void read_result(const mlo_icp_result& r, Results& out) {
    std::memcpy(out.optimal_tf_mean.data(), r.pose_3x4, sizeof(r.pose_3x4));
    std::memcpy(out.optimal_tf_cov.data(), r.cov_6x6, sizeof(r.cov_6x6));
    out.quality = r.quality;
    out.nIterations = r.n_iterations;
    out.terminationReason = r.termination;
    out.nPairings = r.n_pairings;
  }

 private:
  std::vector<double> t1_, t2_, t3_;
};

// observations_filter_1st_pass (default.yaml:278-319): the four filters of the default pipelines, recognised by class
// and realised into the fused device filter (decimate -> by-range -> bbox-outside -> decimate).
struct FilterPipeline1st {
  Formula res_map{"0.2"}, res_icp{"0.6"}, range_min{"0"}, range_max{"1e9"};
  Formula bb_min[3], bb_max[3];
  uint32_t min_pts_map = 2000, min_pts_icp = 2000;
  bool has_range = false, has_bbox = false;
  void initialize(const YamlNode& seq) {
    int decim = 0;
    for (const YamlNode& f : seq.seq) {
      const std::string c = f.at("class_name").str();
      const YamlNode& p = f["params"];
      if (c == "mp2p_icp_filters::FilterDecimateVoxels") {
        const std::string m = p["decimate_method"].str_or("DecimateMethod::FirstPoint");
        if (m != "DecimateMethod::FirstPoint") throw std::runtime_error("FilterDecimateVoxels: only DecimateMethod::FirstPoint is supported");
        (decim == 0 ? res_map : res_icp) = Formula(p.at("voxel_filter_resolution").str());
        (decim == 0 ? min_pts_map : min_pts_icp) = uint32_t(p["minimum_input_points_to_filter"].num(0));
        decim++;
      } else if (c == "mp2p_icp_filters::FilterByRange") {
        has_range = true;
        range_min = Formula(p.at("range_min").str());
        range_max = Formula(p.at("range_max").str());
      } else if (c == "mp2p_icp_filters::FilterBoundingBox") {
        has_bbox = true;
        for (int k = 0; k < 3; k++) {
          bb_min[k] = Formula(p.at("bounding_box_min").seq.at(k).str());
          bb_max[k] = Formula(p.at("bounding_box_max").seq.at(k).str());
        }
      } else {
        throw std::runtime_error("observations_filter_1st_pass: unsupported filter '" + c + "'");
      }
    }
    if (decim != 2) throw std::runtime_error("observations_filter_1st_pass: expected two FilterDecimateVoxels stages");
  }
  mlo_filter1_params realize(const ParameterSource& ps) const {
    mlo_filter1_params f;
    std::memset(&f, 0, sizeof(f));
    f.for_map.voxel_filter_resolution = float(res_map.eval(ps));
    f.for_map.minimum_input_points_to_filter = min_pts_map;
    f.for_icp.voxel_filter_resolution = float(res_icp.eval(ps));
    f.for_icp.minimum_input_points_to_filter = min_pts_icp;
    if (has_range) {
      f.for_icp.use_range = 1;
      f.for_icp.range_min = float(range_min.eval(ps));
      f.for_icp.range_max = float(range_max.eval(ps));
    }
    if (has_bbox) {
      f.for_icp.use_bbox_outside = 1;
      for (int k = 0; k < 3; k++) {
        f.for_icp.bbox_min[k] = float(bb_min[k].eval(ps));
        f.for_icp.bbox_max[k] = float(bb_max[k].eval(ps));
      }
    }
    return f;
  }
};

// metric_map_definition (default.yaml:228-242 / ndt.yaml:234-254)
struct LocalMapDefinition {
  int kind = MLO_MAP_HASHED_VOXEL_POINTS;
  Formula voxel_size{"1.0"}, remove_voxels_farther_than{"0"};
  uint32_t max_points_per_voxel = 20;
  double min_distance_between_points = 0, max_eigen_ratio_for_planes = 0.05;
  uint64_t capacity_voxels = 1u << 17;  // INITIAL device capacity: the map grows on demand (mlo_map_insert), as upstream's is unbounded
  void initialize(const YamlNode& def) {
    const std::string c = def.at("class").str();
    if (c == "mola::HashedVoxelPointCloud") kind = MLO_MAP_HASHED_VOXEL_POINTS;
    else if (c == "mola::NDT") kind = MLO_MAP_NDT;
    else throw std::runtime_error("unsupported metric map class '" + c + "'");
    voxel_size = Formula(def.at("creationOpts").at("voxel_size").str());
    const YamlNode& io = def["insertOpts"];
    max_points_per_voxel = uint32_t(io["max_points_per_voxel"].num(20));
    min_distance_between_points = io["min_distance_between_points"].num(0);
    if (io.has("remove_voxels_farther_than")) remove_voxels_farther_than = Formula(io["remove_voxels_farther_than"].str());
    max_eigen_ratio_for_planes = io["max_eigen_ratio_for_planes"].num(0.05);
    if (def.has("capacity_voxels")) capacity_voxels = uint64_t(def["capacity_voxels"].num());  // extension key (device budget)
  }
};

// ------------------------------------------------------------------------------------------------ orchestrator
struct LidarOdometryParams {  // the subset of LidarOdometry::Parameters around the hot path (LidarOdometry.cpp:125-321)
  double min_time_between_scans = 1e-3;
  double max_sensor_range_filter_coefficient = 0.95, absolute_minimum_sensor_range = 5.0;
  bool optimize_twist = true;
  double optimize_twist_rerun_min_trans = 0.15, optimize_twist_rerun_min_rot_deg = 0.75;
  bool local_map_updates_enabled = true;
  Formula min_translation_between_keyframes{"1.0"}, min_rotation_between_keyframes{"30"}, max_distance_to_keep_keyframes{"0"};
  uint32_t check_for_removal_every_n = 100;
  bool measure_from_last_kf_only = false;  // local_map_updates.measure_from_last_kf_only (LidarOdometry.cpp:188,1066-1068)
  // observation_validity_checks (default.yaml:118-121, LidarOdometry.cpp:749-755,1548-1569); only the 'raw' layer is known here
  bool obs_validity_enabled = false;
  uint64_t obs_validity_min_points = 1000;
  double min_icp_goodness = 0.25;
  bool adaptive_enabled = true;
  double initial_sigma = 2.0, min_motion = 0.1, maximum_sigma = 3.0, kp = 2.0, alpha = 0.9;
  NavStateFuseParams navstate;  // navstate_fuse_params (default.yaml:126-144)
  bool icp_prior_enabled = true;  // MLO_ICP_PRIOR=0: never send the motion-model information to the solver (A/B, tests)
  // observations_filter_2nd_pass FilterDeskew (default.yaml:328-350) + FilterAdjustTimestamps (:267-275)
  bool skip_deskew = false, silently_ignore_no_timestamps = true;
  bool timestamps_middle_is_zero = true;  // TimestampAdjustMethod::MiddleIsZero (else EarliestIsZero)
};

struct ScanOutput {
  bool processed = false, icp_ran = false, icp_good = false, map_updated = false;
  Pose pose = pose_identity();
  double quality = 0, sigma = 0, est_max_range = 0;
  uint32_t icp_iterations = 0, icp_runs = 0;
  int termination = MLO_TERM_UNDEFINED;
  uint64_t n_map_layer = 0, n_icp_layer = 0;
  bool icp_had_prior = false, has_motion_model = false;
  double prior_info_trace = 0;
};

template <class Backend>
class LidarOdometryT {
 public:
  explicit LidarOdometryT(Backend& be) : be_(be) {}
  ~LidarOdometryT() {
    if (map_) be_.destroy_map(map_);
    if (set_) be_.scanset_destroy(set_);
  }
  LidarOdometryT(const LidarOdometryT&) = delete;
  LidarOdometryT& operator=(const LidarOdometryT&) = delete;
  LidarOdometryParams params_;
  ParameterSource parameter_source;

  // mola::LidarOdometry::initialize_frontend (LidarOdometry.cpp:246-476), hot-path subset
  void initialize(const YamlNode& cfg) {
    const YamlNode& p = cfg.at("params");
    params_.min_time_between_scans = p["min_time_between_scans"].num(1e-3);
    params_.max_sensor_range_filter_coefficient = p["max_sensor_range_filter_coefficient"].num(0.95);
    params_.absolute_minimum_sensor_range = p["absolute_minimum_sensor_range"].num(5.0);
    params_.optimize_twist = p["optimize_twist"].boolean(true);
    params_.optimize_twist_rerun_min_trans = p["optimize_twist_rerun_min_trans"].num(0.15);
    params_.optimize_twist_rerun_min_rot_deg = p["optimize_twist_rerun_min_rot_deg"].num(0.75);
    const YamlNode& lm = p["local_map_updates"];
    params_.local_map_updates_enabled = lm["enabled"].boolean(true);
    params_.min_translation_between_keyframes = Formula(lm.at("min_translation_between_keyframes").str());
    params_.min_rotation_between_keyframes = Formula(lm.at("min_rotation_between_keyframes").str());
    if (lm.has("max_distance_to_keep_keyframes")) params_.max_distance_to_keep_keyframes = Formula(lm["max_distance_to_keep_keyframes"].str());
    params_.check_for_removal_every_n = uint32_t(lm["check_for_removal_every_n"].num(100));
    params_.measure_from_last_kf_only = lm["measure_from_last_kf_only"].boolean(false);
    if (p.has("observation_validity_checks")) {
      const YamlNode& ov = p["observation_validity_checks"];
      params_.obs_validity_enabled = ov["enabled"].boolean(false);
      params_.obs_validity_min_points = uint64_t(ov["minimum_point_count"].num(1000));
      const std::string layer = ov["check_layer_name"].str_or("raw");
      if (params_.obs_validity_enabled && layer != "raw")
        throw std::runtime_error("observation_validity_checks: only check_layer_name 'raw' is supported, got '" + layer + "'");
    }
    params_.min_icp_goodness = p["min_icp_goodness"].num(0.25);
    const YamlNode& at = p["adaptive_threshold"];
    params_.adaptive_enabled = at["enabled"].boolean(true);
    params_.initial_sigma = at["initial_sigma"].num(2.0);
    params_.min_motion = at["min_motion"].num(0.1);
    params_.maximum_sigma = at["maximum_sigma"].num(3.0);
    params_.kp = at["kp"].num(2.0);
    if (params_.adaptive_enabled && !(params_.kp > 1.0)) throw std::runtime_error("adaptive_threshold.kp must be > 1 (LidarOdometry.cpp:1468)");
    params_.alpha = at["alpha"].num(0.9);
    if (cfg.has("navstate_fuse_params"))
    {
      const YamlNode& nf = cfg["navstate_fuse_params"];
      NavStateFuseParams& np = params_.navstate;
      np.max_time_to_use_velocity_model = nf["max_time_to_use_velocity_model"].num(0.75);
      np.sliding_window_length = nf["sliding_window_length"].num(0.5);
      np.sigma_random_walk_acceleration_linear = nf["sigma_random_walk_acceleration_linear"].num(1.0);
      np.sigma_random_walk_acceleration_angular = nf["sigma_random_walk_acceleration_angular"].num(10.0);
      np.sigma_integrator_position = nf["sigma_integrator_position"].num(1.0);
      np.sigma_integrator_orientation = nf["sigma_integrator_orientation"].num(1.0);
      np.initial_twist_sigma_lin = nf["initial_twist_sigma_lin"].num(20.0);
      np.initial_twist_sigma_ang = nf["initial_twist_sigma_ang"].num(3.0);
      if (nf["initial_twist"].isSeq())
        for (size_t k = 0; k < 6 && k < nf["initial_twist"].seq.size(); k++) np.initial_twist[k] = nf["initial_twist"].seq[k].num(0.0);
    }
    if (cfg.has("observations_filter_2nd_pass"))
      for (const YamlNode& f : cfg["observations_filter_2nd_pass"].seq)
        if (f["class_name"].str_or("") == "mp2p_icp_filters::FilterDeskew") {
          params_.skip_deskew = f["params"]["skip_deskew"].boolean(false);
          params_.silently_ignore_no_timestamps = f["params"]["silently_ignore_no_timestamps"].boolean(true);
        }
    if (cfg.has("observations_filter_adjust_timestamps"))
      for (const YamlNode& f : cfg["observations_filter_adjust_timestamps"].seq) {
        const std::string m = f["params"]["method"].str_or("TimestampAdjustMethod::MiddleIsZero");
        if (m == "TimestampAdjustMethod::MiddleIsZero") params_.timestamps_middle_is_zero = true;
        else if (m == "TimestampAdjustMethod::EarliestIsZero") params_.timestamps_middle_is_zero = false;
        else throw std::runtime_error("FilterAdjustTimestamps: unsupported method '" + m + "'");
      }
    icp_.initialize(cfg.at("icp_settings_with_vel"));  // AlignKind::RegularOdometry (LidarOdometry.cpp:340-341)
    icp_.attachToParameterSource(parameter_source);
    // AlignKind::NoMotionModel: optional, defaults to the regular set (LidarOdometry.cpp:343-349, default.yaml:154)
    has_icp_no_vel_ = cfg.has("icp_settings_without_vel");
    icp_no_vel_ = ICP<Backend>();
    if (has_icp_no_vel_) {
      icp_no_vel_.initialize(cfg["icp_settings_without_vel"]);
      icp_no_vel_.attachToParameterSource(parameter_source);
    }
    filter1_.initialize(cfg.at("observations_filter_1st_pass"));
    bool found = false;
    for (const YamlNode& g : cfg.at("localmap_generator").seq)
      if (g["params"].has("metric_map_definition")) {
        mapdef_.initialize(g["params"]["metric_map_definition"]);
        found = true;
      }
    if (!found) throw std::runtime_error("localmap_generator: no metric_map_definition");
    if (const char* e = std::getenv("MLO_ICP_PRIOR")) params_.icp_prior_enabled = std::atoi(e) != 0;
    navstate_.params = params_.navstate;
    navstate_.set_lie([this](const double* xi, double* T) { be_.se3_exp(xi, T); },
                      [this](const double* T, double* xi) { be_.se3_log(T, xi); });
    reset_state();
  }

  void reset_state() {
    if (map_) be_.destroy_map(map_);
    map_ = nullptr;
    map_points_ = 0;
    trajectory_.clear();
    keyframes_.clear();
    last_lidar_pose_ = pose_identity();
    navstate_.reset();
    sigma_ = 0;
    est_max_range_.reset();
    inst_max_range_.reset();
    last_obs_time_.reset();
    last_icp_time_.reset();
    first_ever_time_.reset();
    last_motion_model_output_.reset();
    last_icp_was_good_ = true;
    last_icp_quality_ = 0.0;
    removal_counter_ = 0;
  }

  const std::vector<std::pair<double, Pose>>& estimatedTrajectory() const { return trajectory_; }
  double adaptiveSigma() const { return sigma_; }
  void* localMap() const { return map_; }

  // ---------------------------------------------------------------------------------------------------------------
  // mola::LidarOdometry::onLidarImpl for one point cloud (LidarOdometry.cpp:627-1206), cut into phases at the points
  // where the reference calls into its plugins (filter pipelines :732-741, ICP::align :961, merge pipeline :1197).
  // Between phases the scan's layers stay in the backend's scan set (device memory for BackendGpu); a fleet
  // (LidarOdometryFleetT below) runs the same phase of many sequences with ONE backend call.
  //   begin_scan -> [filter] -> after_filter -> ([deskew] -> after_deskew)
  //     -> while icp_pending(): make_align_job -> [align] -> on_align_result -> ([deskew])
  //     -> after_icp -> ([insert] -> after_insert) -> finish_scan
  // `t` = optional per-point timestamps [s] relative to the scan stamp (CPointsMapXYZIRT "t" channel).

  // Phase A.  Returns false when the observation is dropped (:643-657); otherwise `job` describes the filter call.
  bool begin_scan(const float* pts, uint32_t stride, uint64_t n, double stamp, const float* t, mlo_scan_job& job) {
    out_ = ScanOutput{};
    icp_pending_ = false;
    needs_deskew_ = false;
    insert_pending_ = false;
    dropped_ = false;
    if (last_obs_time_ && stamp - *last_obs_time_ < params_.min_time_between_scans) return false;  // :643-657
    prev_obs_time_ = last_obs_time_;
    last_obs_time_ = stamp;  // (:763; taken back in on_layers if the observation turns out to be invalid, :749-755)
    n_raw_ = n;
    stamp_ = stamp;
    out_.processed = true;
    if (!est_max_range_) {  // doInitializeEstimatedMaxSensorRange, :1487-1513
      const double r = std::max(bbox_radius(pts, stride, n), params_.absolute_minimum_sensor_range);
      if (n) est_max_range_ = r;
    }
    // :692 runs BEFORE the motion model is queried for this scan (:808-815): the twist variables seen by the filters (and,
    // unless the hook re-estimates them, by the keyframe formulas) are those of the PREVIOUS scan's motion-model output
    updatePipelineDynamicVariables(last_motion_model_output_);
    motion_ = estimated_navstate(stamp);      // :808-815 (pure query; stored as last_motion_model_output once the
                                              // observation has passed the validity check, on_layers)
    // 1st-pass filter (:732-735); the 2nd pass is the identity with deskew skipped (:737-741)
    std::memset(&job, 0, sizeof(job));
    job.fp = filter1_.realize(parameter_source);
    if (!t && !params_.skip_deskew && !params_.silently_ignore_no_timestamps)
      throw std::runtime_error("FilterDeskew: the point cloud has no per-point timestamps (silently_ignore_no_timestamps is false)");
    do_deskew_ = t && !params_.skip_deskew && n > 0;
    job.pts = pts;
    job.n = n;
    job.t = nullptr;
    if (do_deskew_) {
      // FilterAdjustTimestamps (:267-275) on the raw layer, then the 1st pass keeps t through both decimations and
      // the 2nd pass deskews the two '_skewed' layers with the current twist variables (:328-350)
      adj_t_.assign(t, t + n);
      float tmin = adj_t_[0], tmax = adj_t_[0];
      for (float v : adj_t_) {
        tmin = std::min(tmin, v);
        tmax = std::max(tmax, v);
      }
      const float shift = (params_.timestamps_middle_is_zero ? 0.5f * (tmin + tmax) : tmin) -
                          float(parameter_source.has("SENSOR_TIME_OFFSET") ? parameter_source.get("SENSOR_TIME_OFFSET") : 0.0);
      for (float& v : adj_t_) v -= shift;
      job.t = adj_t_.data();
    }
    return true;
  }
  // Phase C.  With deskew on, the layers only exist after the deskew call: ask for it.
  void after_filter(const mlo_scan_info& info) {
    if (do_deskew_) {
      needs_deskew_ = true;
      first_deskew_ = true;
      return;
    }
    on_layers(info);
  }
  bool needs_deskew() const { return needs_deskew_; }
  const std::array<double, 6>& deskew_twist() const { return twist_; }
  void after_deskew(const mlo_scan_info& info) {
    needs_deskew_ = false;
    if (first_deskew_) {
      first_deskew_ = false;
      on_layers(info);
    }
  }
  bool icp_pending() const { return icp_pending_; }
  // Phase D: one ICP::align call of the do/while at :954-1007
  void make_align_job(mlo_align_job& job) {
    ICP<Backend>& icp = *icp_case_;
    ip_.maxIterations = remaining_;
    hook_.checkpoint = current_solution_;
    icp.setIterationHook(hook_);
    std::memset(&job, 0, sizeof(job));
    job.map = static_cast<const mlo_map*>(map_);
    std::memcpy(job.init_pose_3x4, current_solution_.data(), sizeof(job.init_pose_3x4));
    icp.make_params(ip_, prior_, job.params);  // the same prior on every align call of the do/while (:961-962)
  }
  void on_align_result(const mlo_icp_result& res) {
    Results& r = result_;
    ICP<Backend>::read_result(res, r);
    out_.icp_runs++;
    total_iterations_ += r.nIterations;
    remaining_ = r.nIterations <= remaining_ ? remaining_ - r.nIterations : 0;
    if (r.terminationReason == MLO_TERM_HOOK_REQUEST) {
      current_solution_ = r.optimal_tf_mean;  // the hook stored the new checkpoint (:949)
      if (since_last_ > 0) {  // re-estimate the twist (:973-992), then re-apply the 2nd pass with it (:996-1001)
        const Pose incr = pose_minus(r.optimal_tf_mean, last_keyframe_pose_);
        double w[3];
        rot_log(incr, w);
        twist_ = {incr[3] / since_last_, incr[7] / since_last_, incr[11] / since_last_, w[0] / since_last_, w[1] / since_last_,
                  w[2] / since_last_};
        updatePipelineTwistVariables();
        if (do_deskew_) needs_deskew_ = true;
      }
      return;  // still pending: the loop runs again with the remaining budget
    }
    icp_pending_ = false;
  }
  // Phase E: gating, adaptive sigma, keyframe decision (:1011-1158).  Returns true when a map insert must follow.
  bool after_icp(mlo_insert_job& job) {
    if (dropped_) return false;
    bool updateLocalMap = first_scan_;
    const bool hasMotionModel = motion_.has_value();
    if (!first_scan_) {
      const Results& r = result_;
      out_.icp_ran = true;
      out_.quality = r.quality;
      out_.icp_iterations = r.nIterations;
      out_.termination = r.terminationReason;
      const bool icpIsGood = r.quality >= params_.min_icp_goodness;  // :1026
      last_icp_was_good_ = icpIsGood;
      last_icp_quality_ = r.quality;
      out_.icp_good = icpIsGood;
      if (icpIsGood) {
        last_lidar_pose_ = r.optimal_tf_mean;
        fuse_pose(stamp_, r.optimal_tf_mean, r.optimal_tf_cov.data());  // :1035-1036 (covariance of the ICP result)
        trajectory_.emplace_back(stamp_, last_lidar_pose_);
      } else {
        navstate_.reset();  // navstate_fuse.reset() (:1039)
      }
      parameter_source.updateVariable("icp_iterations", double(r.nIterations));  // :1046-1048
      // (the reference's local counter is never incremented - :926,939 bump the PARAMETER instead - so the variable stays
      // 0 and optimize_twist_max_corrections never limits the re-runs; reproduced as is)
      parameter_source.updateVariable("twistCorrectionCount", 0.0);
      if (params_.adaptive_enabled) doUpdateAdaptiveThreshold(pose_minus(r.optimal_tf_mean, init_guess_), motion_);  // :1052-1063
      // keyframe decision (:1066-1115)
      double dist = 0, rot = 0;
      const bool isFirst = closest_keyframe(last_lidar_pose_, dist, rot);
      const double min_t = params_.min_translation_between_keyframes.eval(parameter_source);
      const double min_r = params_.min_rotation_between_keyframes.eval(parameter_source) * M_PI / 180.0;
      updateLocalMap = icpIsGood && params_.local_map_updates_enabled && hasMotionModel && (isFirst || dist > min_t || rot > min_r);
      if (updateLocalMap) {
        keyframes_.push_back(last_lidar_pose_);
        const double keep = params_.max_distance_to_keep_keyframes.eval(parameter_source);
        if (keep > 0 && removal_counter_++ >= params_.check_for_removal_every_n) {
          removal_counter_ = 0;
          std::vector<Pose> kept;
          for (const Pose& k : keyframes_) {
            const double d[3] = {k[3] - last_lidar_pose_[3], k[7] - last_lidar_pose_[7], k[11] - last_lidar_pose_[11]};
            if (norm3(d) <= keep) kept.push_back(k);
          }
          keyframes_.swap(kept);
        }
      }
    }
    // bad first ICP: restart from scratch (:1150-1158)
    if (!last_icp_was_good_ && trajectory_.size() == 1) {
      if (map_) be_.map_clear(map_);
      map_points_ = 0;
      trajectory_.clear();
      keyframes_.clear();
      updateLocalMap = false;
      last_icp_was_good_ = true;
    }
    if (updateLocalMap) {  // :1161-1206
      updatePipelineDynamicVariables(motion_);  // robot_x.. for FilterMerge
      if (!map_) {
        mlo_map_params mp;
        std::memset(&mp, 0, sizeof(mp));
        mp.kind = mapdef_.kind;
        mp.voxel_size = float(mapdef_.voxel_size.eval(parameter_source));
        mp.max_points_per_voxel = mapdef_.max_points_per_voxel;
        mp.min_distance_between_points = float(mapdef_.min_distance_between_points);
        mp.max_eigen_ratio_for_planes = float(mapdef_.max_eigen_ratio_for_planes);
        mp.min_points_for_plane = 5;
        mp.capacity_voxels = mapdef_.capacity_voxels;
        map_ = be_.create_map(mp);
        cull_dist_ = float(mapdef_.remove_voxels_farther_than.eval(parameter_source));
      }
      std::memset(&job, 0, sizeof(job));
      job.map = static_cast<mlo_map*>(map_);
      std::memcpy(job.pose_3x4, last_lidar_pose_.data(), sizeof(job.pose_3x4));
      job.cull_farther_than = cull_dist_ > 0 ? cull_dist_ : 0.f;
      insert_pending_ = true;
    }
    return updateLocalMap;
  }
  // after_icp() creates the map on the first scan and clears it on a bad first ICP: those calls reach the device
  bool after_icp_touches_backend() const { return !map_ || trajectory_.size() <= 1; }
  void after_insert(const mlo_map_counts& cnt) {
    map_points_ = cnt.n_points;
    out_.map_updated = true;
    insert_pending_ = false;
  }
  ScanOutput finish_scan() {
    if (dropped_) return ScanOutput{};
    out_.pose = last_lidar_pose_;
    out_.sigma = sigma_;
    return out_;
  }

  // One sequence on its own: the phases back to back on a one-slot scan set.
  ScanOutput onLidar(const float* pts, uint32_t stride, uint64_t n, double stamp, const float* t = nullptr) {
    if (!set_) set_ = be_.scanset_create(1);
    mlo_scan_job fj;
    if (!begin_scan(pts, stride, n, stamp, t, fj)) return finish_dropped();
    fj.slot = 0;
    mlo_scan_info info;
    be_.scanset_filter(set_, 1, &fj, stride, &info);
    after_filter(info);
    const uint32_t slot0 = 0;
    auto deskew = [&] {
      be_.scanset_deskew(set_, 1, &slot0, twist_.data(), &info);
      after_deskew(info);
    };
    if (needs_deskew()) deskew();
    while (icp_pending()) {
      mlo_align_job aj;
      make_align_job(aj);
      aj.slot = 0;
      mlo_icp_result res;
      be_.scanset_align(set_, 1, &aj, &res);
      on_align_result(res);
      if (needs_deskew()) deskew();
    }
    mlo_insert_job ij;
    if (after_icp(ij)) {
      ij.slot = 0;
      mlo_map_counts cnt;
      be_.scanset_insert(set_, 1, &ij, &cnt);
      after_insert(cnt);
    }
    return finish_scan();
  }
  ScanOutput finish_dropped() { return ScanOutput{}; }

 private:
  // layers are known (sizes + ICP-layer bounding box): max-range estimate, then first-scan seeding or ICP set-up
  void on_layers(const mlo_scan_info& info) {
    out_.n_map_layer = info.n_map;
    out_.n_icp_layer = info.n_icp;
    doUpdateEstimatedMaxSensorRange(info);  // :744-769 (first points layer of the observation = decimated_for_icp)
    out_.est_max_range = est_max_range_.value_or(0.0);
    if (params_.obs_validity_enabled && !(n_raw_ > params_.obs_validity_min_points)) {  // :749-755: discarded, nothing stored
      last_obs_time_ = prev_obs_time_;
      dropped_ = true;
      out_ = ScanOutput{};
      return;
    }
    if (!first_ever_time_) first_ever_time_ = stamp_;  // :766-767
    last_motion_model_output_ = motion_;                // :811
    first_scan_ = !map_ || map_points_ == 0;
    if (first_scan_) {
      // first point cloud: no ICP, seed the map at the origin (:817-839)
      trajectory_.emplace_back(stamp_, last_lidar_pose_);
      fuse_pose(stamp_, pose_identity());
      return;
    }
    const bool hasMotionModel = motion_.has_value();
    init_guess_ = hasMotionModel ? motion_->pose : last_lidar_pose_;  // :852-897
    prior_.reset();
    if (hasMotionModel && params_.icp_prior_enabled) {  // ICP prior term: any information != 0? (:859-861)
      bool any = false;
      for (double v : motion_->cov_inv) any = any || v != 0.0;
      if (any) prior_ = Prior{motion_->pose, motion_->cov_inv};
    }
    out_.has_motion_model = hasMotionModel;
    out_.icp_had_prior = prior_.has_value();
    if (prior_) for (int k = 0; k < 6; k++) out_.prior_info_trace += prior_->cov_inv[7 * k];
    last_keyframe_pose_ = last_lidar_pose_;                           // :904
    since_last_ = last_icp_time_ ? stamp_ - *last_icp_time_ : 0.0;
    last_icp_time_ = stamp_;
    // without a valid twist estimate the NoMotionModel ICP set applies, if the pipeline defines one (:899-903)
    icp_case_ = (hasMotionModel || !has_icp_no_vel_) ? &icp_ : &icp_no_vel_;
    ip_ = icp_case_->params;
    remaining_ = ip_.maxIterations;
    total_iterations_ = 0;
    hook_ = IterationHook{};
    hook_.enabled = params_.optimize_twist;
    hook_.min_trans = params_.optimize_twist_rerun_min_trans;
    hook_.min_rot_rad = params_.optimize_twist_rerun_min_rot_deg * M_PI / 180.0;
    current_solution_ = init_guess_;
    icp_pending_ = true;
  }

  // mola::NavStateFuse (host/navstate_fuse.hpp): sliding-window constant-velocity model; pose = ICP initial guess,
  // cov_inv = information of the GN prior (:854-877), twist = the vx..wz pipeline variables (:1581-1600)
  using NavState = NavStateFuse::NavState;
  std::optional<NavState> estimated_navstate(double stamp) { return navstate_.estimated_navstate(stamp); }
  void fuse_pose(double stamp, const Pose& p, const double* cov = nullptr) { navstate_.fuse_pose(stamp, p, cov); }
  void rot_log(const Pose& p, double w[3]) {
    Pose r = p;
    r[3] = r[7] = r[11] = 0.0;
    double xi[6];
    be_.se3_log(r.data(), xi);
    w[0] = xi[3];
    w[1] = xi[4];
    w[2] = xi[5];
  }
  static double bbox_radius(const float* p, uint32_t stride, uint64_t n) {
    if (!n) return 0.0;
    float mn[3] = {p[0], p[1], p[2]}, mx[3] = {p[0], p[1], p[2]};
    for (uint64_t i = 1; i < n; i++)
      for (int k = 0; k < 3; k++) {
        mn[k] = std::min(mn[k], p[i * stride + k]);
        mx[k] = std::max(mx[k], p[i * stride + k]);
      }
    const double a[3] = {mn[0], mn[1], mn[2]}, b[3] = {mx[0], mx[1], mx[2]};
    return std::max(norm3(a), norm3(b));
  }
  void doUpdateEstimatedMaxSensorRange(const mlo_scan_info& info) {  // :1515-1546
    if (!est_max_range_ || info.n_icp == 0) return;
    const double a3[3] = {info.icp_min[0], info.icp_min[1], info.icp_min[2]}, b3[3] = {info.icp_max[0], info.icp_max[1], info.icp_max[2]};
    const double radius = std::max(std::max(norm3(a3), norm3(b3)), params_.absolute_minimum_sensor_range);
    inst_max_range_ = radius;
    const double a = params_.max_sensor_range_filter_coefficient;
    est_max_range_ = *est_max_range_ * a + radius * (1.0 - a);
  }
  void doUpdateAdaptiveThreshold(const Pose& err, const std::optional<NavState>& motion) {  // :1449-1485
    if (!est_max_range_) return;
    const double max_range = *est_max_range_;
    double w[3];
    rot_log(err, w);
    const double theta = norm3(w);
    const double t[3] = {err[3], err[7], err[11]};
    const double model_error = norm3(t) + 2.0 * max_range * std::sin(theta / 2.0);
    double rot_error = 0;
    if (motion) rot_error = 0.1 * norm3(&motion->twist[3]) * max_range;
    const double gain = std::min(std::max(params_.kp * (1.0 - last_icp_quality_), 0.1), params_.kp);
    const double new_sigma = (model_error + rot_error) * gain;
    if (sigma_ == 0) sigma_ = params_.initial_sigma;
    sigma_ = params_.alpha * sigma_ + (1.0 - params_.alpha) * new_sigma;
    sigma_ = std::min(std::max(sigma_, params_.min_motion), params_.maximum_sigma);
  }
  void updatePipelineTwistVariables() {  // :1571-1579
    static const char* names[6] = {"vx", "vy", "vz", "wx", "wy", "wz"};
    for (int k = 0; k < 6; k++) parameter_source.updateVariable(names[k], twist_[k]);
  }
  void updatePipelineDynamicVariables(const std::optional<NavState>& motion) {  // :1581-1635
    twist_ = motion ? motion->twist : std::array<double, 6>{0, 0, 0, 0, 0, 0};
    updatePipelineTwistVariables();
    double yaw, pitch, roll;
    pose_ypr(last_lidar_pose_, yaw, pitch, roll);
    parameter_source.updateVariable("robot_x", last_lidar_pose_[3]);
    parameter_source.updateVariable("robot_y", last_lidar_pose_[7]);
    parameter_source.updateVariable("robot_z", last_lidar_pose_[11]);
    parameter_source.updateVariable("robot_yaw", yaw);
    parameter_source.updateVariable("robot_pitch", pitch);
    parameter_source.updateVariable("robot_roll", roll);
    parameter_source.updateVariable("ADAPTIVE_THRESHOLD_SIGMA", sigma_ != 0 ? sigma_ : params_.initial_sigma);
    parameter_source.updateVariable("ICP_ITERATION", 0);
    for (const char* v : {"icp_iterations", "SENSOR_TIME_OFFSET", "twistCorrectionCount"})
      if (!parameter_source.has(v)) parameter_source.updateVariable(v, 0);
    if (est_max_range_) parameter_source.updateVariable("ESTIMATED_SENSOR_MAX_RANGE", *est_max_range_);
    parameter_source.updateVariable("INSTANTANEOUS_SENSOR_MAX_RANGE", inst_max_range_ ? *inst_max_range_ : 20.0);
    if (last_obs_time_ && first_ever_time_)  // :1626-1630
      parameter_source.updateVariable("current_relative_timestamp", *last_obs_time_ - *first_ever_time_);
  }
  // mola::SearchablePoseList::check: distance to the closest stored keyframe; true when the list is empty
  bool closest_keyframe(const Pose& p, double& dist, double& rot) {
    if (keyframes_.empty()) {
      dist = rot = 0;
      return true;
    }
    double best = 1e300;
    const Pose* bk = nullptr;
    if (params_.measure_from_last_kf_only) bk = &keyframes_.back();
    for (const Pose& k : keyframes_) {
      if (params_.measure_from_last_kf_only) break;
      const double d[3] = {k[3] - p[3], k[7] - p[7], k[11] - p[11]};
      const double n = norm3(d);
      if (n < best) {
        best = n;
        bk = &k;
      }
    }
    const Pose rel = pose_minus(p, *bk);
    const double t[3] = {rel[3], rel[7], rel[11]};
    dist = norm3(t);
    double w[3];
    rot_log(rel, w);
    rot = norm3(w);
    return false;
  }

  Backend& be_;
  ICP<Backend> icp_, icp_no_vel_;
  ICP<Backend>* icp_case_ = &icp_;
  bool has_icp_no_vel_ = false;
  FilterPipeline1st filter1_;
  LocalMapDefinition mapdef_;
  void* map_ = nullptr;
  uint64_t map_points_ = 0;
  float cull_dist_ = 0;
  typename Backend::ScanSet* set_ = nullptr;  // one-slot set of the stand-alone onLidar() path
  std::vector<float> adj_t_;
  // per-scan transient state between the phases
  ScanOutput out_;
  std::optional<NavState> motion_;
  std::optional<Prior> prior_;  // in.prior of :859-877
  double stamp_ = 0, since_last_ = 0;
  bool do_deskew_ = false, needs_deskew_ = false, first_deskew_ = false, first_scan_ = false, icp_pending_ = false,
       insert_pending_ = false;
  Pose init_guess_ = pose_identity(), last_keyframe_pose_ = pose_identity(), current_solution_ = pose_identity();
  Parameters ip_;
  IterationHook hook_;
  Results result_;
  uint32_t remaining_ = 0, total_iterations_ = 0;
  std::vector<std::pair<double, Pose>> trajectory_;
  NavStateFuse navstate_;
  std::vector<Pose> keyframes_;
  Pose last_lidar_pose_ = pose_identity();
  std::array<double, 6> twist_{};
  double sigma_ = 0;
  std::optional<double> est_max_range_, inst_max_range_, last_obs_time_, prev_obs_time_, first_ever_time_, last_icp_time_;
  std::optional<NavState> last_motion_model_output_;  // state_.last_motion_model_output: assigned at :811, read at :692
  uint64_t n_raw_ = 0;
  bool dropped_ = false;
  bool last_icp_was_good_ = true;
  double last_icp_quality_ = 0;
  uint32_t removal_counter_ = 0;
};

// A fleet of independent LidarOdometry instances (one sequence each: the reference runs them as separate processes,
// eval/cli_kitti.sh:23) advanced in lock step on one device: every phase of onLidar() is issued ONCE for all
// sequences (one filter pass, one align pass over per-sequence local maps, one insert pass), which is what fills
// a B200 when a single 64-beam scan cannot.  Results are those of the stand-alone instances.
// A small persistent pool for the per-sequence HOST logic of a fleet step (formula realisation, motion model, gating:
// pure functions of one sequence's state).  Device passes stay on the caller's thread: one context = one stream = one
// caller thread (include/mlo_b200.h).  MLO_HOST_THREADS overrides the size (default: min(8, cores / ranks on this node)).
class HostPool {
 public:
  HostPool() {
    unsigned n = std::max(1u, std::thread::hardware_concurrency());
    if (const char* lw = std::getenv("LOCAL_WORLD_SIZE")) n = std::max(1u, n / unsigned(std::max(1, std::atoi(lw))));
    n = std::min(8u, n);
    if (const char* e = std::getenv("MLO_HOST_THREADS")) n = unsigned(std::max(1, std::atoi(e)));
    for (unsigned i = 1; i < n; i++) th_.emplace_back([this] { worker(); });
  }
  ~HostPool() {
    {
      std::lock_guard<std::mutex> l(m_);
      stop_ = true;
      gen_++;
    }
    cv_.notify_all();
    for (auto& t : th_) t.join();
  }
  // fn(i) for i in [0, count); returns when all are done.  Small counts run inline.
  void parallel_for(size_t count, const std::function<void(size_t)>& fn) {
    if (th_.empty() || count < 8) {
      for (size_t i = 0; i < count; i++) fn(i);
      return;
    }
    {
      std::lock_guard<std::mutex> l(m_);
      fn_ = &fn;
      count_ = count;
      next_.store(0);
      pending_ = th_.size();
      gen_++;
    }
    cv_.notify_all();
    run();
    std::unique_lock<std::mutex> l(m_);
    done_.wait(l, [this] { return pending_ == 0; });
    fn_ = nullptr;
    if (err_) {
      auto e = err_;
      err_ = nullptr;
      std::rethrow_exception(e);
    }
  }

 private:
  void run() {
    for (;;) {
      const size_t i = next_.fetch_add(1);
      if (i >= count_) return;
      try {
        (*fn_)(i);
      } catch (...) {
        std::lock_guard<std::mutex> l(m_);
        if (!err_) err_ = std::current_exception();
      }
    }
  }
  void worker() {
    uint64_t seen = 0;
    for (;;) {
      {
        std::unique_lock<std::mutex> l(m_);
        cv_.wait(l, [&] { return gen_ != seen; });
        seen = gen_;
        if (stop_) return;
      }
      run();
      {
        std::lock_guard<std::mutex> l(m_);
        if (--pending_ == 0) done_.notify_one();
      }
    }
  }
  std::vector<std::thread> th_;
  std::mutex m_;
  std::condition_variable cv_, done_;
  const std::function<void(size_t)>* fn_ = nullptr;
  size_t count_ = 0, pending_ = 0;
  std::atomic<size_t> next_{0};
  uint64_t gen_ = 0;
  bool stop_ = false;
  std::exception_ptr err_;
};

template <class Backend>
class LidarOdometryFleetT {
 public:
  LidarOdometryFleetT(Backend& be, uint32_t n_sequences) : be_(be) {
    for (uint32_t i = 0; i < n_sequences; i++) seq_.push_back(std::make_unique<LidarOdometryT<Backend>>(be));
    set_ = be_.scanset_create(n_sequences);
  }
  ~LidarOdometryFleetT() {
    seq_.clear();
    if (set_) be_.scanset_destroy(set_);
  }
  uint32_t size() const { return uint32_t(seq_.size()); }
  LidarOdometryT<Backend>& sequence(uint32_t i) { return *seq_.at(i); }
  void initialize(const YamlNode& cfg) {
    for (auto& s : seq_) s->initialize(cfg);
  }
  // Optional: the clouds of the NEXT onLidarBatch call; their upload overlaps the ICP of the call made in between.
  void prefetch(const float* const* pts, uint32_t stride, const uint64_t* n) { be_.scanset_prefetch(set_, size(), pts, n, stride); }
  // One lock step: cloud i (pts[i] with n[i] points, stamp stamps[i], optional per-point times t[i]) goes to
  // sequence i; pts[i] == nullptr leaves sequence i idle.  out[i] is what onLidar() would have returned.
  // host wall time [ms] spent per phase since the last reset: 0 begin (host), 1 filter, 2 deskew, 3 align, 4 host logic
  // after ICP, 5 insert, 6 lock steps counted
  std::array<double, 8> phase_ms{};
  void onLidarBatch(const float* const* pts, uint32_t stride, const uint64_t* n, const double* stamps, const float* const* t,
                    ScanOutput* out) {
    using clk = std::chrono::steady_clock;
    auto t_last = clk::now();
    auto lap = [&](int k) {
      const auto now = clk::now();
      phase_ms[k] += std::chrono::duration<double, std::milli>(now - t_last).count();
      t_last = now;
    };
    phase_ms[6] += 1.0;
    const uint32_t S = size();
    std::vector<uint32_t> live;  // sequences with an accepted observation this step
    std::vector<mlo_scan_job> fjobs;
    {
      std::vector<mlo_scan_job> all(S);
      std::vector<uint8_t> ok(S, 0);
      pool_.parallel_for(S, [&](size_t i) {  // per-sequence host logic: independent states
        out[i] = ScanOutput{};
        if (!pts[i]) return;
        ok[i] = seq_[i]->begin_scan(pts[i], stride, n[i], stamps[i], t ? t[i] : nullptr, all[i]) ? 1 : 0;
      });
      for (uint32_t i = 0; i < S; i++)
        if (ok[i]) {
          all[i].slot = i;
          fjobs.push_back(all[i]);
          live.push_back(i);
        }
    }
    if (live.empty()) return;
    lap(0);
    std::vector<mlo_scan_info> info(live.size());
    be_.scanset_filter(set_, uint32_t(fjobs.size()), fjobs.data(), stride, info.data());
    lap(1);
    for (size_t k = 0; k < live.size(); k++) seq_[live[k]]->after_filter(info[k]);
    run_deskews(live);
    lap(2);
    for (;;) {
      std::vector<uint32_t> who;
      for (uint32_t i : live)
        if (seq_[i]->icp_pending()) who.push_back(i);
      if (who.empty()) break;
      std::vector<mlo_align_job> ajobs(who.size());
      pool_.parallel_for(who.size(), [&](size_t k) {  // realises the per-iteration formula tables of each sequence
        seq_[who[k]]->make_align_job(ajobs[k]);
        ajobs[k].slot = who[k];
      });
      std::vector<mlo_icp_result> res(who.size());
      be_.scanset_align(set_, uint32_t(ajobs.size()), ajobs.data(), res.data());
      lap(3);
      pool_.parallel_for(who.size(), [&](size_t k) { seq_[who[k]]->on_align_result(res[k]); });
      run_deskews(who);
      lap(2);
    }
    std::vector<uint32_t> ins;
    std::vector<mlo_insert_job> ijobs;
    {
      std::vector<mlo_insert_job> all(live.size());
      std::vector<uint8_t> want(live.size(), 0);
      bool serial = false;  // (device calls stay on the caller's thread)
      for (uint32_t i : live) serial = serial || seq_[i]->after_icp_touches_backend();
      if (serial)
        for (size_t k = 0; k < live.size(); k++) want[k] = seq_[live[k]]->after_icp(all[k]) ? 1 : 0;
      else
        pool_.parallel_for(live.size(), [&](size_t k) { want[k] = seq_[live[k]]->after_icp(all[k]) ? 1 : 0; });
      for (size_t k = 0; k < live.size(); k++)
        if (want[k]) {
          all[k].slot = live[k];
          ijobs.push_back(all[k]);
          ins.push_back(live[k]);
        }
    }
    lap(4);
    if (!ins.empty()) {
      std::vector<mlo_map_counts> cnt(ins.size());
      be_.scanset_insert(set_, uint32_t(ijobs.size()), ijobs.data(), cnt.data());
      for (size_t k = 0; k < ins.size(); k++) seq_[ins[k]]->after_insert(cnt[k]);
    }
    lap(5);
    for (uint32_t i : live) out[i] = seq_[i]->finish_scan();
  }

 private:
  void run_deskews(const std::vector<uint32_t>& among) {
    std::vector<uint32_t> slots;
    std::vector<double> tw;
    for (uint32_t i : among)
      if (seq_[i]->needs_deskew()) {
        slots.push_back(i);
        const auto& w = seq_[i]->deskew_twist();
        tw.insert(tw.end(), w.begin(), w.end());
      }
    if (slots.empty()) return;
    std::vector<mlo_scan_info> info(slots.size());
    be_.scanset_deskew(set_, uint32_t(slots.size()), slots.data(), tw.data(), info.data());
    for (size_t k = 0; k < slots.size(); k++) seq_[slots[k]]->after_deskew(info[k]);
  }
  Backend& be_;
  std::vector<std::unique_ptr<LidarOdometryT<Backend>>> seq_;
  typename Backend::ScanSet* set_ = nullptr;
  HostPool pool_;
};

}  // namespace mlo_host
