// oracle/oracle_capi.cpp — TEST INFRASTRUCTURE: extern "C" surface of the CPU oracle for ctypes.
// Mirrors the shapes of include/mlo_b200.h (the POD parameter/result structs are shared so a test
// passes the very same bytes to both sides).  parity unpinned — see mlo_oracle.hpp header.
#include <cstdio>
#include <memory>

#include "../include/mlo_b200.h"
#include "mlo_oracle.hpp"

using namespace orc;

namespace {
IcpParams convert(const mlo_icp_params* p) {
  IcpParams q;
  q.max_iterations = p->max_iterations;
  q.min_abs_step_trans = p->min_abs_step_trans;
  q.min_abs_step_rot = p->min_abs_step_rot;
  q.solver = p->solver;
  q.gn_max_iterations = p->gn_max_iterations;
  q.gn_min_delta = p->gn_min_delta;
  q.robust_kernel = p->robust_kernel;
  q.matcher_mask = p->matcher_mask;
  if (p->pt2pt_threshold_by_iter) q.thr_pt2pt.assign(p->pt2pt_threshold_by_iter, p->pt2pt_threshold_by_iter + p->table_len);
  if (p->pt2pl_threshold_by_iter) q.thr_pt2pl.assign(p->pt2pl_threshold_by_iter, p->pt2pl_threshold_by_iter + p->table_len);
  if (p->kernel_param_by_iter) q.kernel_param.assign(p->kernel_param_by_iter, p->kernel_param_by_iter + p->table_len);
  q.threshold_angular_deg = p->threshold_angular_deg;
  q.w_pt2pt = p->pt2pt_weight;
  q.w_pt2pl = p->pt2pl_weight;
  q.has_prior = p->has_prior != 0;
  q.prior_pose = Pose::from3x4(p->prior_pose_3x4);
  for (int i = 0; i < 6; i++)
    for (int j = 0; j < 6; j++) q.prior_info[i][j] = p->prior_info_6x6[i * 6 + j];
  q.hook_enabled = p->hook_enabled != 0;
  q.hook_min_trans = p->hook_min_trans;
  q.hook_min_rot = p->hook_min_rot_rad;
  q.hook_checkpoint = Pose::from3x4(p->hook_checkpoint_pose_3x4);
  return q;
}
void convert(const IcpResult& r, mlo_icp_result* o) {
  r.pose.to3x4(o->pose_3x4);
  for (int i = 0; i < 6; i++)
    for (int j = 0; j < 6; j++) o->cov_6x6[i * 6 + j] = r.cov[i][j];
  o->quality = r.quality;
  o->n_iterations = r.n_iterations;
  o->termination = r.termination;
  o->n_pairings = r.n_pairings;
  o->n_potential_pairings = r.n_potential;
  o->n_query_iterations = r.n_query_iterations;
  o->n_candidate_points = r.n_candidate_points;
}
DecimateParams convert(const mlo_decimate_params* p) {
  DecimateParams d;
  d.resolution = p->voxel_filter_resolution;
  d.min_input_points = p->minimum_input_points_to_filter;
  d.use_range = p->use_range != 0;
  d.range_min = p->range_min;
  d.range_max = p->range_max;
  d.use_bbox_outside = p->use_bbox_outside != 0;
  for (int k = 0; k < 3; k++) {
    d.bbox_min[k] = p->bbox_min[k];
    d.bbox_max[k] = p->bbox_max[k];
  }
  return d;
}
}  // namespace

extern "C" {

int32_t orc_voxel_index_map(float coord, float voxel_size) { return voxel_index_map(coord, 1.0f / voxel_size); }
int32_t orc_voxel_index_filter(float coord, float resolution) { return voxel_index_filter(coord, resolution); }
double orc_geman_mcclure(double e2, double c) { return geman_mcclure_weight(e2, c); }

void* orc_pool_create(int n) { return new Pool(n); }
void orc_pool_destroy(void* p) { delete static_cast<Pool*>(p); }

void* orc_map_create(const mlo_map_params* p) {
  return new VoxelMap(p->kind, p->voxel_size, p->max_points_per_voxel, p->min_distance_between_points,
                      p->max_eigen_ratio_for_planes, p->min_points_for_plane);
}
void orc_map_destroy(void* m) { delete static_cast<VoxelMap*>(m); }
void orc_map_clear(void* m) { static_cast<VoxelMap*>(m)->clear(); }
void orc_map_insert(void* m, const float* pts, uint32_t stride, uint64_t n, const double* pose) {
  static_cast<VoxelMap*>(m)->insert(pts, stride, n, Pose::from3x4(pose));
}
void orc_map_cull(void* m, const double* sensor, float dist) { static_cast<VoxelMap*>(m)->cull(sensor, dist); }
void orc_map_stats(void* m, uint64_t* nv, uint64_t* np) {
  *nv = static_cast<VoxelMap*>(m)->n_voxels();
  *np = static_cast<VoxelMap*>(m)->n_points();
}
void orc_map_nn_single(void* m, const float* q, uint32_t stride, uint64_t n, float* out_xyz, float* out_d2,
                       uint8_t* out_found, uint64_t* n_candidates) {
  auto* M = static_cast<VoxelMap*>(m);
  uint64_t nc = 0;
  for (uint64_t i = 0; i < n; i++) {
    float o[3] = {0, 0, 0}, d2;
    const bool f = M->nn_single(q[i * stride], q[i * stride + 1], q[i * stride + 2], o, d2, &nc);
    out_xyz[3 * i] = o[0];
    out_xyz[3 * i + 1] = o[1];
    out_xyz[3 * i + 2] = o[2];
    out_d2[i] = d2;
    out_found[i] = f ? 1 : 0;
  }
  if (n_candidates) *n_candidates = nc;
}
void orc_map_nn_plane(void* m, const float* q, uint32_t stride, uint64_t n, float* out_mean, float* out_normal,
                      float* out_dist, uint8_t* out_found) {
  auto* M = static_cast<VoxelMap*>(m);
  for (uint64_t i = 0; i < n; i++) {
    float mean[3] = {0, 0, 0}, nr[3] = {0, 0, 0}, d;
    const bool f = M->nn_plane(q[i * stride], q[i * stride + 1], q[i * stride + 2], mean, nr, d);
    for (int k = 0; k < 3; k++) {
      out_mean[3 * i + k] = mean[k];
      out_normal[3 * i + k] = nr[k];
    }
    out_dist[i] = d;
    out_found[i] = f ? 1 : 0;
  }
}
int orc_map_export(void* m, int32_t* keys, uint32_t* counts, float* xyz, uint64_t max_voxels, uint64_t max_points,
                   uint64_t* n_voxels, uint64_t* n_points) {
  std::vector<int32_t> k;
  std::vector<uint32_t> c;
  std::vector<float> p;
  static_cast<VoxelMap*>(m)->export_sorted(k, c, p);
  *n_voxels = c.size();
  *n_points = p.size() / 3;
  if (keys && counts && xyz) {
    if (c.size() > max_voxels || p.size() / 3 > max_points) return -1;
    std::memcpy(keys, k.data(), k.size() * sizeof(int32_t));
    std::memcpy(counts, c.data(), c.size() * sizeof(uint32_t));
    std::memcpy(xyz, p.data(), p.size() * sizeof(float));
  }
  return 0;
}

void orc_voxel_decimate_first(const float* pts, uint32_t stride, uint64_t n, const mlo_decimate_params* p,
                              uint32_t* out_idx, uint64_t* out_n) {
  std::vector<uint32_t> kept;
  decimate_first(pts, stride, n, convert(p), kept);
  *out_n = kept.size();
  if (out_idx) std::memcpy(out_idx, kept.data(), kept.size() * sizeof(uint32_t));
}

// observations_filter_1st_pass (default.yaml:278-319): outputs xyz packed.
void orc_filter_1st_pass(const float* pts, uint32_t stride, uint64_t n, const mlo_filter1_params* p,
                         float* out_map_xyz, uint64_t* out_map_n, float* out_icp_xyz, uint64_t* out_icp_n) {
  std::vector<uint32_t> k1, k2;
  DecimateParams d1 = convert(&p->for_map);
  d1.use_range = d1.use_bbox_outside = false;
  decimate_first(pts, stride, n, d1, k1);
  std::vector<float> a(k1.size() * 3);
  for (size_t i = 0; i < k1.size(); i++)
    for (int k = 0; k < 3; k++) a[3 * i + k] = pts[size_t(k1[i]) * stride + k];
  // range + bbox predicates produce the map layer ("decimated_for_map_skewed"), then 2nd decimation
  DecimateParams pred = convert(&p->for_icp);
  std::vector<float> b;
  b.reserve(a.size());
  for (size_t i = 0; i < k1.size(); i++)
    if (predicate_keep(pred, a[3 * i], a[3 * i + 1], a[3 * i + 2])) b.insert(b.end(), &a[3 * i], &a[3 * i] + 3);
  *out_map_n = b.size() / 3;
  if (out_map_xyz) std::memcpy(out_map_xyz, b.data(), b.size() * sizeof(float));
  DecimateParams d2 = pred;
  d2.use_range = d2.use_bbox_outside = false;
  decimate_first(b.data(), 3, b.size() / 3, d2, k2);
  *out_icp_n = k2.size();
  if (out_icp_xyz)
    for (size_t i = 0; i < k2.size(); i++)
      for (int k = 0; k < 3; k++) out_icp_xyz[3 * i + k] = b[size_t(k2[i]) * 3 + k];
}

// the same 1st pass carrying one extra per-point channel (timestamps): outputs are x, y, z, t
void orc_filter_1st_pass_xyzt(const float* pts, uint32_t stride, const float* t, uint64_t n, const mlo_filter1_params* p,
                              float* out_map_xyzt, uint64_t* out_map_n, float* out_icp_xyzt, uint64_t* out_icp_n) {
  std::vector<uint32_t> k1, k2;
  DecimateParams d1 = convert(&p->for_map);
  d1.use_range = d1.use_bbox_outside = false;
  decimate_first(pts, stride, n, d1, k1);
  DecimateParams pred = convert(&p->for_icp);
  std::vector<float> b;
  for (uint32_t i : k1) {
    const float* q = pts + size_t(i) * stride;
    if (predicate_keep(pred, q[0], q[1], q[2])) {
      b.insert(b.end(), q, q + 3);
      b.push_back(t ? t[i] : 0.f);
    }
  }
  *out_map_n = b.size() / 4;
  if (out_map_xyzt) std::memcpy(out_map_xyzt, b.data(), b.size() * sizeof(float));
  DecimateParams d2 = pred;
  d2.use_range = d2.use_bbox_outside = false;
  decimate_first(b.data(), 4, b.size() / 4, d2, k2);
  *out_icp_n = k2.size();
  if (out_icp_xyzt)
    for (size_t i = 0; i < k2.size(); i++)
      for (int k = 0; k < 4; k++) out_icp_xyzt[4 * i + k] = b[size_t(k2[i]) * 4 + k];
}
void orc_deskew(const float* xyzt, uint64_t n, const double* twist, float* out_xyz) { deskew(xyzt, n, twist, out_xyz); }

void orc_icp_align(void* map, const float* local, uint32_t stride, uint64_t n, const double* init_pose,
                   const mlo_icp_params* p, mlo_icp_result* out, void* pool, double* trace_poses,
                   uint32_t* trace_pairs, uint32_t trace_cap) {
  IcpResult r;
  icp_align(local, stride, n, *static_cast<VoxelMap*>(map), Pose::from3x4(init_pose), convert(p), r,
            static_cast<Pool*>(pool), trace_poses != nullptr);
  convert(r, out);
  if (trace_poses)
    for (size_t i = 0; i < r.trace_pose.size() && i < trace_cap; i++) {
      r.trace_pose[i].to3x4(trace_poses + 12 * i);
      if (trace_pairs) trace_pairs[i] = r.trace_pairs[i];
    }
}

// SE(3) helpers exposed for known-answer tests
void orc_set_conventions(int index_floor, int gm_form, int cull_metric) {
  conv().index_floor = index_floor;
  conv().gm_form = gm_form;
  conv().cull_metric = cull_metric;
}
void orc_se3_exp(const double* xi, double* pose) { se3_exp(xi).to3x4(pose); }
void orc_se3_log(const double* pose, double* xi) { se3_log(Pose::from3x4(pose), xi); }
// Covariance of mp2p_icp::Results::optimal_tf in MRPT's (x y z yaw pitch roll) chart from the tangent-space covariance
// (consumed at LidarOdometry.cpp:1035-1036): J by central differences of ypr(T exp(eps)), independent of
// the product's closed form.
void orc_cov_tangent_to_ypr(const double* pose, const double* cov, double* out) {
  const Pose T = Pose::from3x4(pose);
  auto chart = [](const Pose& P, double* v) {
    v[0] = P.t[0]; v[1] = P.t[1]; v[2] = P.t[2];
    v[3] = std::atan2(P.R[1][0], P.R[0][0]);
    v[4] = std::atan2(-P.R[2][0], std::hypot(P.R[0][0], P.R[1][0]));
    v[5] = std::atan2(P.R[2][1], P.R[2][2]);
  };
  double J[36];
  const double h = 1e-6;
  for (int k = 0; k < 6; k++) {
    double e[6] = {0, 0, 0, 0, 0, 0}, a[6], b[6];
    e[k] = h;
    chart(compose(T, se3_exp(e)), a);
    e[k] = -h;
    chart(compose(T, se3_exp(e)), b);
    for (int i = 0; i < 6; i++) {
      double d = a[i] - b[i];
      if (i >= 3) d = std::remainder(d, 2.0 * M_PI);
      J[6 * i + k] = d / (2.0 * h);
    }
  }
  for (int i = 0; i < 6; i++)
    for (int j = 0; j < 6; j++) {
      double s = 0;
      for (int m = 0; m < 6; m++)
        for (int n = 0; n < 6; n++) s += J[6 * i + m] * cov[6 * m + n] * J[6 * j + n];
      out[6 * i + j] = s;
    }
}
void orc_pose_minus(const double* a, const double* b, double* out) {
  minus(Pose::from3x4(a), Pose::from3x4(b)).to3x4(out);
}
void orc_pose_compose(const double* a, const double* b, double* out) {
  compose(Pose::from3x4(a), Pose::from3x4(b)).to3x4(out);
}
int orc_horn(const float* g, const float* l, uint64_t n, double* pose) {
  std::vector<PairPt> pr(n);
  for (uint64_t i = 0; i < n; i++)
    for (int k = 0; k < 3; k++) {
      pr[i].g[k] = g[3 * i + k];
      pr[i].l[k] = l[3 * i + k];
    }
  Pose T = Pose::identity();
  if (!horn_solve(pr, T)) return -1;
  T.to3x4(pose);
  return 0;
}

// Timed CPU baseline of one scan step: filter_1st_pass -> align (-> optional insert), returning the
// three profiler buckets of the reference (LidarOdometry.cpp:732,916,1162) in milliseconds.
void orc_scan_register(void* map, const float* raw, uint32_t stride, uint64_t n, const mlo_filter1_params* fp,
                       const double* init_pose, const mlo_icp_params* ip, int insert_into_map, float cull_dist,
                       void* pool, mlo_icp_result* out, double* ms3) {
  using clk = std::chrono::steady_clock;
  auto t0 = clk::now();
  std::vector<float> a(n * 3), b(n * 3);
  uint64_t na = 0, nb = 0;
  orc_filter_1st_pass(raw, stride, n, fp, a.data(), &na, b.data(), &nb);
  auto t1 = clk::now();
  IcpResult r;
  icp_align(b.data(), 3, nb, *static_cast<VoxelMap*>(map), Pose::from3x4(init_pose), convert(ip), r,
            static_cast<Pool*>(pool), false);
  convert(r, out);
  auto t2 = clk::now();
  if (insert_into_map) {
    auto* M = static_cast<VoxelMap*>(map);
    M->insert(a.data(), 3, na, r.pose);
    if (cull_dist > 0) M->cull(r.pose.t, cull_dist);
  }
  auto t3 = clk::now();
  if (ms3) {
    ms3[0] = std::chrono::duration<double, std::milli>(t1 - t0).count();
    ms3[1] = std::chrono::duration<double, std::milli>(t2 - t1).count();
    ms3[2] = std::chrono::duration<double, std::milli>(t3 - t2).count();
  }
}

}  // extern "C"
