/* mlo_b200.h — C ABI of libmlo_b200.so: B200-native (sm_100a) per-scan ICP registration.
 *
 * This is the drop-in boundary for ONE hot path of MOLAorg/mola_lidar_odometry:
 *   mola::LidarOdometry::onLidarImpl  (module/src/LidarOdometry.cpp:627-1314)
 *     -> mp2p_icp_filters::apply_filter_pipeline (LidarOdometry.cpp:732-735)   [FilterDecimateVoxels]
 *     -> mp2p_icp::ICP::align               (LidarOdometry.cpp:961-962)        [matcher + solver + quality]
 *     -> FilterMerge -> HashedVoxelPointCloud insert (LidarOdometry.cpp:1161-1206)
 *
 * Each entry point cites the reference interface it replaces.  The arithmetic of those
 * interfaces lives in un-vendored dependencies (mp2p_icp, mola_metric_maps); the citations
 * are therefore the reference's own call sites and YAML parameter blocks.
 *
 * Conventions
 *   - POD only, host pointers in / host pointers out, no exceptions cross this boundary.
 *   - Every function returns MLO_OK (0) or a negative mlo_status; the message is kept per
 *     context (mlo_last_error).  The reference reports errors by C++ exceptions that its
 *     worker thread latches into `fatal_error` (LidarOdometry.cpp:614-619); the adapter
 *     shim re-throws on a negative code.
 *   - One mlo_ctx = one CUDA device + one stream + one caller thread, mirroring the
 *     reference's one-worker-thread-per-LidarOdometry rule (LidarOdometry.h:546-549).
 *   - Poses are 3x4 row-major doubles [R | t] mapping local (sensor/vehicle) -> global (map).
 *   - Point clouds are float32, array-of-structs with a stride in floats (3 = xyz packed,
 *     4 = KITTI .bin x,y,z,intensity) or struct-of-arrays (the layout of mrpt::maps::CPointsMap).
 *   - There is NO CPU fallback: every compute entry point runs sm_100a kernels or fails.
 */
#ifndef MLO_B200_H
#define MLO_B200_H

#include <stdint.h>

#ifdef __cplusplus
extern "C" {
#endif

#define MLO_ABI_VERSION 1

typedef enum mlo_status {
  MLO_OK = 0,
  MLO_ERR_INVALID_ARG = -1,
  MLO_ERR_CUDA = -2,        /* a CUDA runtime call failed; see mlo_last_error */
  MLO_ERR_NO_DEVICE = -3,   /* no sm_100 device: the product path refuses to run elsewhere */
  MLO_ERR_CAPACITY = -4,    /* the map would exceed 2^26 voxels, or device memory for its growth is exhausted */
  MLO_ERR_KEY_RANGE = -5,   /* a voxel index does not fit the packed 21-bit-per-axis key */
  MLO_ERR_UNSUPPORTED = -6
} mlo_status;

/* mp2p_icp::IterTermReason as consumed at LidarOdometry.cpp:970,1007,1019 */
typedef enum mlo_term_reason {
  MLO_TERM_UNDEFINED = 0,
  MLO_TERM_NO_PAIRINGS = 1,
  MLO_TERM_SOLVER_ERROR = 2,
  MLO_TERM_MAX_ITERATIONS = 3,
  MLO_TERM_STALLED = 4,
  MLO_TERM_HOOK_REQUEST = 5
} mlo_term_reason;

/* mp2p_icp::RobustKernel (pipelines/lidar3d-default.yaml:188) */
typedef enum mlo_robust_kernel {
  MLO_KERNEL_NONE = 0,
  MLO_KERNEL_GEMAN_MCCLURE = 1,
  MLO_KERNEL_CAUCHY = 2
} mlo_robust_kernel;

/* metric map kinds (pipelines/lidar3d-default.yaml:230, pipelines/lidar3d-ndt.yaml:236) */
typedef enum mlo_map_kind { MLO_MAP_HASHED_VOXEL_POINTS = 0, MLO_MAP_NDT = 1 } mlo_map_kind;

#define MLO_MATCHER_PT2PT 1u /* mp2p_icp::Matcher_Points_DistanceThreshold */
#define MLO_MATCHER_PT2PL 2u /* mp2p_icp::Matcher_Point2Plane */

#define MLO_SOLVER_GAUSS_NEWTON 0 /* mp2p_icp::Solver_GaussNewton */
#define MLO_SOLVER_HORN 1         /* mp2p_icp::Solver_Horn */

typedef struct mlo_ctx mlo_ctx; /* opaque: device, stream, scratch */
typedef struct mlo_map mlo_map; /* opaque: hash-voxel local map resident in HBM */

/* ------------------------------------------------------------------ context */
int mlo_abi_version(void);
int mlo_create(int cuda_device, mlo_ctx** out);
void mlo_destroy(mlo_ctx* ctx);
const char* mlo_last_error(const mlo_ctx* ctx);
/* Name + SM count + compute capability of the device this context is bound to. */
int mlo_device_info(const mlo_ctx* ctx, char* name, uint32_t name_len, int* sm_count, int* cc_major, int* cc_minor);
/* The CUDA stream (cudaStream_t) every kernel of this context is launched on; lets a
 * harness record CUDA events around calls without the library knowing about it. */
void* mlo_stream(const mlo_ctx* ctx);
/* Count of kernels this context has launched since creation (bench.py "gpu_launches"). */
uint64_t mlo_launch_count(const mlo_ctx* ctx);
/* Launch-policy knobs of this context (never change results beyond summation order; for A/B runs and for tests that must
 * exercise one specific device path).  Names: "align_path" (0 auto, 1 one kernel per phase = the large-batch launch
 * sequence, 2 queue-driven persistent kernel, 3 one thread block per problem), "large_batch_queries" (total queries at
 * which auto picks the launch sequence; 0 = SM count x 1024), "tail_handover", "tail_path", "stream_groups",
 * "fuse_inner", "prior_ahead" (a second warp linearises the prior term for the next solve while the first finishes the
 * current one), "tail_queries_per_sm", "check_every", "block_threads" (256 / 512), "block_cluster" (thread blocks per problem: 1 / 2 / 4 / 8), "filter_group_mb",
 * "filter_ppt", "filter_kernel" (0 by batch size, 1 global scratch tables with the blocks of a cloud spread over the
 * device, 2 one thread block per cloud with its scratch in shared memory), "filter_cta_min_clouds" (batch size at which 0
 * picks 2), "force_kernel", "wl_variant" (drain loop of the work-list kernel: 0-3 segment-wise merge at 8 / 6 blocks per
 * SM, plain or software-pipelined; 4-5 cp.async.bulk staging; 6-9 contiguous ranges per lane group; 10-11 one partial per
 * warp; 12-13 split 32-bit keys; 14-15 ballot winner; 16-17 eight segments per round; 18-19 run merging), "wl_warps", "pers_minb"; read-only "last_align_path", "last_stream_groups", "last_tail_handover",
 * "last_block_cluster", "last_block_threads" describe the last align call, "last_filter_kernel" the last filter batch.
 * Unknown name: MLO_ERR_INVALID_ARG. */
int mlo_set_option(mlo_ctx* ctx, const char* name, int64_t value);
int mlo_get_option(const mlo_ctx* ctx, const char* name, int64_t* value);

/* ------------------------------------------------------------------ local map
 * Replaces mola::HashedVoxelPointCloud / mola::NDT behind
 *   metric_map_definition.{class,creationOpts.voxel_size,insertOpts.*}
 *   (pipelines/lidar3d-default.yaml:228-242, pipelines/lidar3d-ndt.yaml:234-254). */
typedef struct mlo_map_params {
  int32_t kind;                      /* mlo_map_kind */
  float voxel_size;                  /* creationOpts.voxel_size [m] */
  uint32_t max_points_per_voxel;     /* insertOpts.max_points_per_voxel (1..32) */
  float min_distance_between_points; /* insertOpts.min_distance_between_points [m], 0 = off */
  float max_eigen_ratio_for_planes;  /* NDT insertOpts.max_eigen_ratio_for_planes */
  uint32_t min_points_for_plane;     /* NDT: voxel needs >= this many points to be a plane (default 5) */
  uint64_t capacity_voxels;          /* INITIAL capacity in occupied voxels: inserts re-hash the map into larger buffers
                                        on demand (upstream's map is unbounded); hard limit 2^26 */
} mlo_map_params;

int mlo_map_create(mlo_ctx* ctx, const mlo_map_params* p, mlo_map** out);
void mlo_map_destroy(mlo_map* map);
int mlo_map_clear(mlo_map* map); /* local_map->clear() at LidarOdometry.cpp:1152 */

/* FilterMerge -> insertPoint loop (pipelines/lidar3d-default.yaml:362-368; LidarOdometry.cpp:1197):
 * g = pose * p for every input point in input order; append to voxel key(g) unless the voxel is
 * full or (min_distance_between_points > 0 and) a stored point of that voxel is closer than that. */
int mlo_map_insert(mlo_map* map, const float* pts, uint32_t stride_floats, uint64_t n, const double pose_3x4[12]);
int mlo_map_insert_soa(mlo_map* map, const float* x, const float* y, const float* z, uint64_t n,
                       const double pose_3x4[12]);
/* insertOpts.remove_voxels_farther_than (pipelines/lidar3d-default.yaml:238): erase voxels whose
 * per-axis cell distance to the sensor's cell exceeds ceil(dist / voxel_size). */
int mlo_map_cull(mlo_map* map, const double sensor_xyz[3], float remove_farther_than);
/* NearestNeighborsCapable::nn_single_search over the 3x3x3 cells around key(q).
 * out_xyz: n*3 floats, out_d2: n floats (+inf when nothing found), out_found: n bytes. */
int mlo_map_nn_single(const mlo_map* map, const float* q, uint32_t stride_floats, uint64_t n, float* out_xyz,
                      float* out_d2, uint8_t* out_found);
/* mola::NDT nearest-plane query (Matcher_Point2Plane's map interface, pipelines/lidar3d-ndt.yaml:195-200): among
 * the 3x3x3 cells around key(q), the planar voxel with the smallest |n.(q - mean)|.  NDT maps only.
 * out_mean / out_normal: n*3 floats, out_dist: n floats (+inf when none), out_found: n bytes. */
int mlo_map_nn_plane(const mlo_map* map, const float* q, uint32_t stride_floats, uint64_t n, float* out_mean,
                     float* out_normal, float* out_dist, uint8_t* out_found);
int mlo_map_stats(const mlo_map* map, uint64_t* n_voxels, uint64_t* n_points);
/* Flat export, sorted by (kx,ky,kz): keys 3*i32 per voxel, counts u32 per voxel, then points
 * (x,y,z f32) in stored slot order.  Pass NULL buffers to query sizes only. */
int mlo_map_export(const mlo_map* map, int32_t* keys, uint32_t* counts, float* xyz, uint64_t max_voxels,
                   uint64_t max_points, uint64_t* n_voxels, uint64_t* n_points);
/* The voxel index of one coordinate exactly as the device computes it (bit-exact parity probe). */
int32_t mlo_voxel_index(float coord, float voxel_size);
/* Host evaluations of the SE(3) routines the solve kernel uses (same source, compiled for the host):
 * exp/log with tangent order (x y z rx ry rz), and d log(D exp(e))/de at e = 0 (6x6 row-major) used by the
 * prior term of Solver_GaussNewton (LidarOdometry.cpp:854-877).  Pure functions, no device needed. */
void mlo_se3_exp(const double xi[6], double pose_3x4[12]);
void mlo_se3_log(const double pose_3x4[12], double xi[6]);
void mlo_se3_right_jacobian_inv(const double xi[6], double J_6x6[36]);
/* mp2p_icp::Results::optimal_tf.cov lives in MRPT's (x y z yaw pitch roll) chart (it feeds NavStateFuse::fuse_pose at
 * LidarOdometry.cpp:1035-1036); mlo_icp_result.cov_6x6 is in the tangent space of T <- T exp(eps).  First-order change of
 * chart: cov_ypr = J cov J^T with J = d(x y z yaw pitch roll)/d eps at `pose` (R = Rz(yaw) Ry(pitch) Rx(roll)). */
void mlo_cov_tangent_to_ypr(const double pose_3x4[12], const double cov_tangent_6x6[36], double cov_ypr_6x6[36]);

/* ------------------------------------------------------------------ filters
 * Replaces mp2p_icp_filters::FilterDecimateVoxels, DecimateMethod::FirstPoint
 * (pipelines/lidar3d-default.yaml:285-292,312-319), optionally fused with the FilterByRange
 * (:297-302) and FilterBoundingBox "outside" (:305-310) predicates that sit between the two
 * decimations.  Output = input indices of the kept points, ascending. */
typedef struct mlo_decimate_params {
  float voxel_filter_resolution;           /* [m] */
  uint32_t minimum_input_points_to_filter; /* pass-through below this size */
  int32_t use_range;                       /* FilterByRange: keep range_min <= |p| <= range_max */
  float range_min, range_max;
  int32_t use_bbox_outside; /* FilterBoundingBox: keep points OUTSIDE [bbox_min,bbox_max] */
  float bbox_min[3], bbox_max[3];
} mlo_decimate_params;

int mlo_voxel_decimate_first(mlo_ctx* ctx, const float* pts, uint32_t stride_floats, uint64_t n,
                             const mlo_decimate_params* p, uint32_t* out_kept_idx, uint64_t* out_n);

/* The whole observations_filter_1st_pass (pipelines/lidar3d-default.yaml:278-319) on device:
 * decimate(res_map) -> by-range -> bbox-outside -> [map layer] -> decimate(res_icp) -> [icp layer].
 * Outputs are xyz packed (3 floats per point). */
typedef struct mlo_filter1_params {
  mlo_decimate_params for_map; /* first FilterDecimateVoxels; use_range/use_bbox ignored here */
  mlo_decimate_params for_icp; /* range + bbox predicates applied BEFORE this decimation */
} mlo_filter1_params;
int mlo_filter_1st_pass(mlo_ctx* ctx, const float* pts, uint32_t stride_floats, uint64_t n,
                        const mlo_filter1_params* p, float* out_map_xyz, uint64_t* out_map_n, float* out_icp_xyz,
                        uint64_t* out_icp_n);

/* The same 1st pass carrying one extra per-point channel `t` (per-point timestamps of CPointsMapXYZIRT; NULL = zeros):
 * outputs are x, y, z, t (4 floats per point), the "..._skewed" layers that FilterDeskew consumes. */
int mlo_filter_1st_pass_xyzt(mlo_ctx* ctx, const float* pts, uint32_t stride_floats, const float* t, uint64_t n,
                             const mlo_filter1_params* p, float* out_map_xyzt, uint64_t* out_map_n, float* out_icp_xyzt,
                             uint64_t* out_icp_n);
/* mp2p_icp_filters::FilterDeskew (pipelines/lidar3d-default.yaml:328-350; re-run inside the ICP loop at
 * LidarOdometry.cpp:999): p' = exp_SO3(w t) p + v t with twist = (vx vy vz wx wy wz) and t the 4th input float. */
int mlo_deskew(mlo_ctx* ctx, const float* xyzt, uint64_t n, const double twist[6], float* out_xyz);

/* ------------------------------------------------------------------ ICP
 * Replaces mp2p_icp::ICP::align as called at LidarOdometry.cpp:961-962 with the object graph of
 * pipelines/lidar3d-default.yaml:162-209 (ndt: pipelines/lidar3d-ndt.yaml:162-216):
 *   Matcher_Points_DistanceThreshold / Matcher_Point2Plane -> Solver_GaussNewton (or Solver_Horn)
 *   -> QualityEvaluator_PairedRatio.
 * Runtime formulas (threshold, robustKernelParam are expressions over ICP_ITERATION and
 * ADAPTIVE_THRESHOLD_SIGMA, default.yaml:190,198) are evaluated by the host per iteration and
 * passed as tables; the iteration hook of LidarOdometry.cpp:923-952 is passed as data. */
typedef struct mlo_icp_params {
  uint32_t max_iterations;      /* params.maxIterations */
  double min_abs_step_trans;    /* params.minAbsStep_trans */
  double min_abs_step_rot;      /* params.minAbsStep_rot */
  int32_t solver;               /* MLO_SOLVER_* */
  uint32_t gn_max_iterations;   /* Solver_GaussNewton.maxIterations (inner) */
  double gn_min_delta;          /* Solver_GaussNewton minDelta (upstream default 1e-7) */
  int32_t robust_kernel;        /* mlo_robust_kernel */
  uint32_t matcher_mask;        /* MLO_MATCHER_* bits */
  /* per-iteration tables, each of length table_len; iteration i uses entry min(i, table_len-1) */
  uint32_t table_len;
  const double* pt2pt_threshold_by_iter;  /* Matcher_Points_DistanceThreshold.threshold */
  const double* pt2pl_threshold_by_iter;  /* Matcher_Point2Plane.distanceThreshold (may be NULL) */
  const double* kernel_param_by_iter;     /* Solver_GaussNewton.robustKernelParam */
  double threshold_angular_deg;           /* Matcher_Points_DistanceThreshold.thresholdAngularDeg */
  double pt2pt_weight, pt2pl_weight;      /* pair-type weights (1.0) */
  /* prior term (LidarOdometry.cpp:854-877): mean pose + 6x6 information, tangent order (x y z rx ry rz) */
  int32_t has_prior;
  double prior_pose_3x4[12];
  double prior_info_6x6[36];
  /* iteration hook as data (LidarOdometry.cpp:923-952) */
  int32_t hook_enabled;
  double hook_min_trans;
  double hook_min_rot_rad;
  double hook_checkpoint_pose_3x4[12];
} mlo_icp_params;

typedef struct mlo_icp_result {
  double pose_3x4[12]; /* Results::optimal_tf.mean */
  double cov_6x6[36];  /* inverse of the final Gauss-Newton Hessian, tangent order */
  double quality;      /* QualityEvaluator_PairedRatio */
  uint32_t n_iterations;
  int32_t termination; /* mlo_term_reason */
  uint64_t n_pairings; /* size of the final pairing set */
  uint64_t n_potential_pairings;
  /* oracle-countable traffic terms of SURVEY.md §8(d) (sum over executed iterations) */
  uint64_t n_query_iterations;  /* sum over iterations of N_q */
  uint64_t n_candidate_points;  /* P: candidate points visited */
} mlo_icp_result;

void mlo_icp_params_default(mlo_icp_params* p); /* values of pipelines/lidar3d-default.yaml:169-209 minus tables */

int mlo_icp_align(mlo_ctx* ctx, const float* local_pts, uint32_t stride_floats, uint64_t n_local,
                  const mlo_map* global, const double init_pose_3x4[12], const mlo_icp_params* p,
                  mlo_icp_result* out);
int mlo_icp_align_soa(mlo_ctx* ctx, const float* x, const float* y, const float* z, uint64_t n_local,
                      const mlo_map* global, const double init_pose_3x4[12], const mlo_icp_params* p,
                      mlo_icp_result* out);

/* Per-iteration record of an align call: the content of mp2p_icp's ICP log files (params.generateDebugFiles /
 * saveIterationDetails, pipelines/lidar3d-default.yaml:177-182; docs/mola_lo_pipelines.rst:239-260) that can be had
 * without the MRPT serialisation: one record per executed ICP iteration, written by the device as the loop runs.
 * mlo_icp_log_enable(ctx, n) keeps up to n records per problem for every following align call of the context (0 = off);
 * mlo_icp_log_read returns those of problem `problem` of the LAST align call, in iteration order. */
typedef struct mlo_icp_iteration_record {
  uint32_t iteration;        /* ICP_ITERATION of this record */
  uint32_t n_pairings;       /* pairings found by the matchers in this iteration */
  double pose_3x4[12];       /* solution after this iteration's solver step(s) */
  double threshold_pt2pt;    /* Matcher_Points_DistanceThreshold.threshold realised for this iteration */
  double threshold_pt2pl;    /* Matcher_Point2Plane.distanceThreshold */
  double kernel_param;       /* Solver_GaussNewton.robustKernelParam */
  double step_trans, step_rot; /* the step measure of the stall test (min over prev / prev-prev) */
  int32_t termination;       /* mlo_term_reason decided after this iteration, MLO_TERM_UNDEFINED = the loop goes on */
  int32_t pad;
} mlo_icp_iteration_record;
int mlo_icp_log_enable(mlo_ctx* ctx, uint32_t max_records_per_problem);
int mlo_icp_log_read(mlo_ctx* ctx, uint32_t problem, mlo_icp_iteration_record* out, uint32_t max_records, uint32_t* n);

/* B independent aligns against ONE read-only map in one device pass (Monte-Carlo
 * initial poses / independent scans).  Problem b uses local points [offsets[b], offsets[b+1]).
 * `params` has one entry per problem (tables may be shared). */
int mlo_icp_align_batch(mlo_ctx* ctx, uint32_t n_problems, const float* local_pts, uint32_t stride_floats,
                        const uint64_t* offsets, const mlo_map* global, const double* init_poses_3x4,
                        const mlo_icp_params* params, mlo_icp_result* out);

/* One scan step of the hot path with host buffers in and out (bench.py "e2e"):
 *   raw cloud --H2D--> filter_1st_pass --> align(icp layer, map) --D2H--> result
 * and, when insert_into_map != 0, insert of the map layer at the resulting pose (keyframe update). */
int mlo_scan_register(mlo_ctx* ctx, mlo_map* map, const float* raw_pts, uint32_t stride_floats, uint64_t n,
                      const mlo_filter1_params* fp, const double init_pose_3x4[12], const mlo_icp_params* ip,
                      int insert_into_map, float cull_farther_than, mlo_icp_result* out);

/* Batched variant: B raw scans against one read-only map (no insert). */
int mlo_scan_register_batch(mlo_ctx* ctx, const mlo_map* map, uint32_t n_scans, const float* raw_pts,
                            uint32_t stride_floats, const uint64_t* offsets, const mlo_filter1_params* fps,
                            const double* init_poses_3x4, const mlo_icp_params* ips, mlo_icp_result* out);

/* Pipelined host-buffer path: two staging slots on a dedicated copy stream.  mlo_stage_upload_async registers
 * the H2D of one batch of raw clouds (pinned host memory recommended; the buffer must stay valid until the slot
 * is consumed) into `slot` (0/1) and returns at once; the transfer itself is enqueued by the next compute call
 * of this context right after that call's own small parameter uploads, so it overlaps that call's ICP loop
 * (the H2D copy engine serves transfers in submission order).  mlo_scan_register_batch_staged runs
 * filter -> align on the batch staged in `slot` once its copy has landed.  Registering batch k+1 before
 * consuming batch k therefore hides its transfer behind the compute of batch k. */
int mlo_stage_upload_async(mlo_ctx* ctx, int slot, const float* raw_pts, uint32_t stride_floats, uint32_t n_scans,
                           const uint64_t* offsets);
int mlo_scan_register_batch_staged(mlo_ctx* ctx, const mlo_map* map, int slot, const mlo_filter1_params* fps,
                                   const double* init_poses_3x4, const mlo_icp_params* ips, mlo_icp_result* out);

/* ------------------------------------------------------------------ device-resident handles
 * Same operations with inputs already resident in HBM (bench.py "value"; pipelines that keep
 * scans on the device).  A mlo_dcloud is a device float4 array owned by the library. */
typedef struct mlo_dcloud mlo_dcloud;
int mlo_dcloud_upload(mlo_ctx* ctx, const float* pts, uint32_t stride_floats, uint64_t n, mlo_dcloud** out);
int mlo_dcloud_upload_batch(mlo_ctx* ctx, const float* pts, uint32_t stride_floats, uint32_t n_clouds,
                            const uint64_t* offsets, mlo_dcloud** out);
void mlo_dcloud_destroy(mlo_dcloud* c);
uint64_t mlo_dcloud_size(const mlo_dcloud* c);
int mlo_scan_register_batch_resident(mlo_ctx* ctx, const mlo_map* map, const mlo_dcloud* raw_batch,
                                     const mlo_filter1_params* fps, const double* init_poses_3x4,
                                     const mlo_icp_params* ips, mlo_icp_result* out);
int mlo_icp_align_batch_resident(mlo_ctx* ctx, const mlo_dcloud* local_batch, const mlo_map* global,
                                 const double* init_poses_3x4, const mlo_icp_params* params, mlo_icp_result* out);

/* ------------------------------------------------------------------ scan sets: device-resident scan layers
 * The layers an observation is reduced to by the 1st/2nd pass filters ("decimated_for_map", "decimated_for_icp" and
 * their "_skewed" inputs, pipelines/lidar3d-default.yaml:278-350) stay in HBM between the calls that the reference makes
 * on them (apply_filter_pipeline at LidarOdometry.cpp:732-741, ICP::align at :961, the merge pipeline at :1197), and
 * every call works on MANY scans at once: a set of n_slots scans belonging to independent LidarOdometry instances
 * (one sequence each) that advance in lock step.  Only the raw clouds go up and only counts,
 * bounding boxes and ICP results come back.  One set belongs to one context; calls are stream-ordered. */
typedef struct mlo_scanset mlo_scanset;
typedef struct mlo_scan_job {
  uint32_t slot;          /* which scan of the set this cloud becomes */
  const float* pts;       /* host, array-of-structs */
  const float* t;         /* host per-point times (CPointsMapXYZIRT "t") or NULL */
  uint64_t n;
  mlo_filter1_params fp;  /* observations_filter_1st_pass realised for this scan */
} mlo_scan_job;
typedef struct mlo_scan_info {
  uint64_t n_map, n_icp;          /* sizes of the two layers */
  float icp_min[3], icp_max[3];   /* bounding box of the ICP layer (doUpdateEstimatedMaxSensorRange, :1515-1546) */
} mlo_scan_info;
typedef struct mlo_align_job {
  uint32_t slot;
  const mlo_map* map;             /* the local map of the sequence this scan belongs to */
  double init_pose_3x4[12];
  mlo_icp_params params;
} mlo_align_job;
typedef struct mlo_insert_job {
  uint32_t slot;
  mlo_map* map;
  double pose_3x4[12];
  float cull_farther_than;        /* insertOpts.remove_voxels_farther_than, 0 = off */
} mlo_insert_job;
typedef struct mlo_map_counts { uint64_t n_voxels, n_points; } mlo_map_counts;

int mlo_scanset_create(mlo_ctx* ctx, uint32_t n_slots, mlo_scanset** out);
void mlo_scanset_destroy(mlo_scanset* set);
/* 1st-pass filter of n_jobs raw clouds (all with the same stride); replaces the previous contents of the WHOLE set
 * (slots without a job become empty).  With any job carrying `t`, the layers are the "_skewed" ones and
 * mlo_scanset_deskew must run before align/insert.  info[j] describes job j. */
int mlo_scanset_filter(mlo_scanset* set, uint32_t n_jobs, const mlo_scan_job* jobs, uint32_t stride_floats,
                       mlo_scan_info* info);
/* Optional: announce the raw clouds of the NEXT mlo_scanset_filter call (same host pointers, sizes and order of the
 * non-empty clouds).  Their host-to-device transfer is enqueued on the context's copy stream from inside the next
 * compute call, so it overlaps that call's ICP loop; the buffers must stay valid and unchanged until that filter call
 * (pinned memory recommended).  A filter call with other clouds simply ignores the announcement. */
int mlo_scanset_prefetch(mlo_scanset* set, uint32_t n_clouds, const float* const* pts, const uint64_t* n,
                         uint32_t stride_floats);
/* FilterDeskew of the listed slots with one twist (vx vy vz wx wy wz) each; info[i] is refreshed for slots[i]. */
int mlo_scanset_deskew(mlo_scanset* set, uint32_t n, const uint32_t* slots, const double* twists6, mlo_scan_info* info);
/* ICP::align of the listed scans, each against its own map, in one device pass. */
int mlo_scanset_align(mlo_scanset* set, uint32_t n_jobs, const mlo_align_job* jobs, mlo_icp_result* out);
/* Map-layer insert (+ cull) of the listed scans into their maps; one synchronisation for all of them. */
int mlo_scanset_insert(mlo_scanset* set, uint32_t n_jobs, const mlo_insert_job* jobs, mlo_map_counts* out);
/* Copy one layer of one slot back (tests): layer 0 = map layer, 1 = ICP layer; xyz packed. */
int mlo_scanset_download(mlo_scanset* set, uint32_t slot, int layer, float* out_xyz, uint64_t max_points, uint64_t* n);

/* Per-kernel device time (ms, CUDA events on the context stream) accumulated since the last reset,
 * for the three profiler buckets of the reference (LidarOdometry.cpp:732,916,1162) and the
 * dominant kernel.  Timing is only collected when enabled (it adds event records). */
typedef struct mlo_profile {
  double filter_1st_ms, run_icp_ms, update_local_map_ms;
  double nn_kernel_ms;       /* sum of fused match+accumulate kernel durations */
  uint64_t nn_kernel_launches;
  uint64_t nn_query_iterations, nn_candidate_points, nn_blocks; /* terms of the §8(d) byte formula */
} mlo_profile;
int mlo_profile_enable(mlo_ctx* ctx, int enabled);
int mlo_profile_get(mlo_ctx* ctx, mlo_profile* out, int reset);

#ifdef __cplusplus
}
#endif
#endif /* MLO_B200_H */
