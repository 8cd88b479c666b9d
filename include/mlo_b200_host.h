/* mlo_b200_host.h — C surface of the C++ host layer (mola_lidar_odometry_b200/host/pipeline.hpp): the
 * mola::LidarOdometry caller contract around the hot path (module/src/LidarOdometry.cpp:627-1206) driven by the
 * reference's pipeline YAML (pipelines/lidar3d-default.yaml, pipelines/lidar3d-ndt.yaml), with all arithmetic
 * done by the CUDA C ABI of mlo_b200.h.  Exported by libmlo_b200.so. */
#ifndef MLO_B200_HOST_H
#define MLO_B200_HOST_H
#include "mlo_b200.h"
#ifdef __cplusplus
extern "C" {
#endif

typedef struct mlo_lo mlo_lo; /* one mola::LidarOdometry instance (one sequence, one worker thread) */

typedef struct mlo_lo_scan_output {
  int32_t processed, icp_ran, icp_good, map_updated;
  double pose_3x4[12];
  double quality, sigma, est_max_range;
  uint32_t icp_iterations, icp_runs;
  int32_t termination;
  uint64_t n_map_layer, n_icp_layer;
  int32_t icp_had_prior;     /* the motion model's information went to the solver as the prior (LidarOdometry.cpp:859-861) */
  int32_t has_motion_model;  /* estimated_navstate() produced an estimate for this scan (LidarOdometry.cpp:808-815) */
  double prior_info_trace;   /* trace of that 6x6 information matrix */
} mlo_lo_scan_output;

/* LidarOdometry::initialize(cfg) (LidarOdometry.cpp:246): yaml = file path (is_text == 0) or YAML text. */
int mlo_lo_create(mlo_ctx* ctx, const char* yaml, int is_text, mlo_lo** out);
void mlo_lo_destroy(mlo_lo* lo);
const char* mlo_lo_last_error(const mlo_lo* lo);
/* LidarOdometry::onNewObservation + spin until !isBusy (apps/mola-lidar-odometry-cli.cpp:494-521): one cloud. */
int mlo_lo_on_lidar(mlo_lo* lo, const float* pts, uint32_t stride_floats, uint64_t n, double stamp_s, mlo_lo_scan_output* out);
/* The same with the per-point time channel `t` [s, relative] of a CPointsMapXYZIRT cloud: enables FilterAdjustTimestamps +
 * FilterDeskew + the twist re-estimation loop of LidarOdometry.cpp:923-1005 (unless MOLA_SKIP_DESKEW / optimize_twist off). */
int mlo_lo_on_lidar_t(mlo_lo* lo, const float* pts, uint32_t stride_floats, const float* t, uint64_t n, double stamp_s,
                      mlo_lo_scan_output* out);
/* LidarOdometry::estimatedTrajectory (LidarOdometry.cpp:1425) */
int mlo_lo_trajectory(const mlo_lo* lo, double* stamps, double* poses_3x4, uint64_t max_n, uint64_t* n);
int mlo_lo_reset(mlo_lo* lo); /* LidarOdometry::reset (LidarOdometry.cpp:495) */

/* A fleet of n_sequences independent LidarOdometry instances advanced in lock step on one context: every phase of the
 * per-scan path (filter, ICP::align, map merge) is ONE device pass over all sequences, each with its own local map
 * (mlo_scanset_* in mlo_b200.h).  The reference runs independent sequences as separate processes
 * (eval/cli_kitti.sh:23); results per sequence are those of mlo_lo_on_lidar. */
typedef struct mlo_fleet mlo_fleet;
int mlo_fleet_create(mlo_ctx* ctx, const char* yaml, int is_text, uint32_t n_sequences, mlo_fleet** out);
void mlo_fleet_destroy(mlo_fleet* f);
const char* mlo_fleet_last_error(const mlo_fleet* f);
/* One lock step: cloud i -> sequence i (pts[i] == NULL leaves sequence i idle; t may be NULL, or hold NULL entries). */
int mlo_fleet_on_lidar(mlo_fleet* f, const float* const* pts, uint32_t stride_floats, const uint64_t* n, const double* stamps_s,
                       const float* const* t, mlo_lo_scan_output* out);
/* Optional: announce the clouds of the NEXT mlo_fleet_on_lidar call (same pointers / sizes); their host-to-device
 * transfer then overlaps the ICP of the call made in between (mlo_scanset_prefetch).  Buffers must stay valid. */
int mlo_fleet_prefetch(mlo_fleet* f, const float* const* pts, uint32_t stride_floats, const uint64_t* n);
/* Host wall time [ms] per phase accumulated since the last reset: [0] per-scan host logic before the filter,
 * [1] filter pass, [2] deskew passes, [3] align passes, [4] host logic after ICP, [5] insert pass, [6] lock steps. */
int mlo_fleet_phase_times(mlo_fleet* f, double out_ms[8], int reset);
int mlo_fleet_trajectory(const mlo_fleet* f, uint32_t sequence, double* stamps, double* poses_3x4, uint64_t max_n, uint64_t* n);

/* Pure host helpers (no device): parse the pipeline YAML and realise its formulas for given variable values. */
const char* mlo_host_last_error(void);
int mlo_host_icp_tables(const char* yaml_text, double sigma, uint32_t n_iterations, double* thr_pt2pt, double* thr_pt2pl,
                        double* kernel_param, mlo_icp_params* scalars);
int mlo_host_filter1(const char* yaml_text, double est_max_range, double inst_max_range, mlo_filter1_params* out);
int mlo_host_mapdef(const char* yaml_text, double est_max_range, mlo_map_params* out, float* remove_voxels_farther_than);
int mlo_host_eval_formula(const char* expr, const char* const* names, const double* values, uint32_t n, double* out);

#ifdef __cplusplus
}
#endif
#endif
