// mlo_b200.cu — C ABI (include/mlo_b200.h) over the sm_100a kernels in map.cuh / filter.cuh / icp.cuh.
// Host side only orchestrates: it owns device memory, the context stream and the launch sequence.
// There is no CPU implementation of any compute entry point in this library.
#include <algorithm>
#include <functional>
#include <map>
#include <mutex>
#include <vector>

#include "filter.cuh"
#include "icp.cuh"
#include "icp_block.cuh"
#include "map.cuh"
#include "se3.cuh"

using namespace mlo;

// ------------------------------------------------------------------ small host utilities
struct DBuf {  // grow-only device buffer, backed by the device's stream-ordered pool (release threshold unlimited, see
               // mlo_create): a scan set or context that is destroyed and re-created gets its buffers back without a
               // driver allocation.  Allocation and release are ordered on the null stream and completed before use.
  void* p = nullptr;
  size_t cap = 0;
  static cudaError_t pool_alloc(void** out, size_t bytes) {
    cudaError_t e = cudaMallocAsync(out, bytes, nullptr);
    if (e == cudaSuccess) e = cudaStreamSynchronize(nullptr);
    return e;
  }
  cudaError_t ensure(size_t bytes) {
    if (bytes <= cap) return cudaSuccess;
    release();
    // geometric growth: batch sizes creep up by a few points from step to step, and every re-allocation is a
    // device-wide synchronisation
    size_t want = std::max(bytes + bytes / 4, size_t(1) << 16);
    want = (want + 255) & ~size_t(255);
    cudaError_t e = pool_alloc(&p, want);
    if (e != cudaSuccess && want > bytes) {  // tight on memory: retry with the exact size
      cudaGetLastError();
      want = (bytes + 255) & ~size_t(255);
      e = pool_alloc(&p, want);
    }
    if (e == cudaSuccess) cap = want;
    else p = nullptr;
    return e;
  }
  void release() {
    if (p) {
      cudaDeviceSynchronize();  // (work on any stream may still read the buffer; growth and teardown are rare)
      cudaFreeAsync(p, nullptr);
    }
    p = nullptr;
    cap = 0;
  }
  template <class T>
  T* as() const {
    return static_cast<T*>(p);
  }
};
struct HBuf {  // grow-only pinned host buffer
  void* p = nullptr;
  size_t cap = 0;
  cudaError_t ensure(size_t bytes) {
    if (bytes <= cap) return cudaSuccess;
    if (p) cudaFreeHost(p);
    p = nullptr;
    cap = 0;
    size_t want = std::max(bytes, size_t(1) << 12);
    cudaError_t e = cudaMallocHost(&p, want);
    if (e == cudaSuccess) cap = want;
    return e;
  }
  void release() {
    if (p) cudaFreeHost(p);
    p = nullptr;
    cap = 0;
  }
  template <class T>
  T* as() const {
    return static_cast<T*>(p);
  }
};

struct mlo_ctx {
  int device = 0;
  cudaStream_t stream = nullptr;
  std::string err;
  uint64_t launches = 0;
  int sm_count = 0, cc_major = 0, cc_minor = 0;
  int force_kernel = 0;  // 0 auto, 1 thread-per-query, 2 warp-per-query (MLO_FORCE_KERNEL, experiments only)
  // two staging slots for the pipelined host-buffer path
  cudaStream_t copy_stream = nullptr;
  // Large batches run their launch sequence as `stream_groups` independent halves on separate streams so that the
  // latency-bound one-warp-per-problem solve of one group overlaps the bandwidth-bound match kernel of the other.
  static constexpr int MAX_GROUPS = 4;
  cudaStream_t aux_stream[MAX_GROUPS - 1] = {nullptr, nullptr, nullptr};
  cudaEvent_t ev_fork = nullptr, ev_join[MAX_GROUPS - 1] = {nullptr, nullptr, nullptr};
  int stream_groups = 3;  // MLO_STREAM_GROUPS (A/B at B=512: 1: 21.5k, 2: 22.5k, 3: 22.9k, 4: 22.8k scans/s)
  struct Stage {
    DBuf buf;
    cudaEvent_t ready = nullptr;
    std::vector<uint64_t> offsets;
    uint32_t stride = 0;
    bool pending = false;
    const float* deferred_src = nullptr;  // upload requested but not yet enqueued (see issue_stage_upload)
  } stage[2];
  bool use_persistent = true;  // MLO_PERSISTENT=0 selects the one-kernel-per-phase launch sequence
  bool persistent_forced = false;
  // mlo_set_option("align_path") / MLO_ALIGN_PATH: 0 = auto (block kernel for small batches, launch sequence for large
  // ones), 1 = launch sequence, 2 = queue-driven persistent kernel, 3 = one thread block per problem (k_icp_block)
  int align_path = 0;
  int tail_path = 2;      // kernel that finishes the stragglers of a launch sequence: 2 = queue (measured faster at B = 512), 3 = block
  int block_threads = 0;  // threads per block of k_icp_block: 0 = auto, else 256 / 512
  int block_cluster = 0;  // thread blocks per problem (cluster size) of k_icp_block: 0 = auto, else 1 / 2 / 4 / 8
  int last_block_cluster = 0, last_block_threads = 0;
  bool block_attr_set[6] = {false, false, false, false, false, false};
  // mlo_set_option("filter_group_mb"): scratch bytes per group of clouds in the 1st-pass filter.  Groups small enough to
  // keep their tables in L2 (72 MB) measured SLOWER than one group for the whole batch (more, shorter launches:
  // profiles/README.md), so the default is "everything in one group".
  int filter_group_mb = 1 << 20;
  int filter_ppt = 4;  // mlo_set_option("filter_ppt"): input points per thread of the decimation kernels (1 / 2 / 4)
  // mlo_set_option("filter_kernel"): 0 = by batch size (default), 1 = k_decim_claim / k_decim_finalize (global scratch
  // tables, blocks of a cloud spread over the device), 2 = k_decim_cta (one thread block per cloud, scratch in shared
  // memory; falls back to 1 for a batch whose clouds do not fit its 32-bit keys / its table).
  int filter_kernel = 0;
  // mlo_set_option("filter_cta_min_clouds"): smallest batch that takes k_decim_cta under 0.  Filter wall per lock step of
  // a fleet, global tables vs block per cloud: 1 cloud 0.098 / 0.195 ms, 8: 0.157 / 0.233, 32: 0.297 / 0.298,
  // 512 (config[1]): 2.5 / 0.65 ms (profiles/README.md)
  int filter_cta_min_clouds = 32;
  int filter_cta_backoff = 0;      // batches left before k_decim_cta is tried again after a fallback
  int last_filter_kernel = 0;      // what the last filter batch ran (tests)
  bool cta_attr_set = false;
  int conv_index_floor = 0, conv_gm_form = 0, conv_cull_metric = 0;  // [VERIFY] conventions (common.cuh), captured by maps at creation
  uint32_t log_cap = 0;  // mlo_icp_log_enable: records kept per problem (0 = off)
  DBuf d_log;
  std::vector<mlo_icp_iteration_record> h_log;
  std::vector<unsigned long long> h_queue;  // host image of the work queue's slot words
  uint32_t h_log_problems = 0;
  int last_align_path = 0, last_stream_groups = 0, last_tail_handover = 0;  // what the last align call did (tests)
  uint64_t large_batch_queries = 0;  // 0 = auto (sm_count * 1024): batches at or above it take the launch sequence
  int tpq_min_queries_per_sm = 512;  // MLO_TPQ_MIN: below this many queries per SM the warp-per-query chunks win
                                     // (S=8 fleet, 56 k queries: align 1.40 ms warp vs 2.15 ms thread-per-query)
  int wl_min_blocks = 32;  // MLO_WL_MIN_BLOCKS (one-warp blocks per SM)
  int wl_variant = 3;      // MLO_WL_VARIANT (A/B at B=512, scans/s): 0 = 8 blocks/SM, 64 registers: 25.7k; 1 = drain loop software-
                           // pipelined at 8 blocks/SM: 23.8k; 2 = pipelined at 6 blocks/SM: 25.3k; 3 = 6 blocks/SM, 80 registers,
                           // fewer spills: 26.1k (default)
  int wl_warps = 4;        // MLO_WL_WARPS: 4 = four-warp blocks (default), 1 = one-warp blocks (A/B: slower)
  int table_factor = 8;    // MLO_TABLE_FACTOR: hash buckets per voxel of capacity (load factor ~0.08: fewer re-probes)
  int pers_minb = 0;         // MLO_PERS_MINB: resident blocks per SM the persistent kernel is compiled for (4: 128 registers,
                             // 2: 255 registers - the solve step keeps more of its 6x6 arrays in registers); 0 = auto: 2 for a
                             // single problem (align 0.926 vs 0.978 ms per scan), 4 otherwise (32 sequences: 1.72 vs 2.01 ms)
  int pers_minb_now = 4;     // the choice for the align call in progress
  int qpw_floor = 4;         // MLO_QPW_FLOOR: fewest queries a warp handles per chunk in the warp-per-query kernels
  bool fuse_inner = true;    // MLO_FUSE_INNER=0: inner GN iterations as separate accumulate + solve launches (A/B)
  bool prior_ahead = true;   // mlo_set_option("prior_ahead"): a second warp prepares the prior's linearisation for the next solve
                             // while the first finishes the current one (icp.cuh block_solve)
  bool tail_handover = true;  // MLO_TAIL_HANDOVER=0 disables the launch-sequence -> persistent hand-over
  int tail_queries_per_sm = 512;  // mlo_set_option("tail_queries_per_sm"): the hand-over happens once the still-active problems
                                  // hold fewer queries than this per SM
  int check_every = 4;            // mlo_set_option("check_every"): ICP iterations between two looks at the active-problem counter
  int consuming_slot = -1;  // staging slot read by the compute call in progress
  // transfers registered by a prefetch call and enqueued from inside the next compute call, right after that call's
  // own small parameter uploads (the H2D copy engine serves transfers in submission order)
  struct Deferred {
    void* owner;
    std::function<int()> fn;
  };
  std::vector<Deferred> deferred;
  void drop_deferred(void* owner) {
    deferred.erase(std::remove_if(deferred.begin(), deferred.end(), [owner](const Deferred& d) { return d.owner == owner; }),
                   deferred.end());
  }
  int persistent_blocks = 0;
  std::string dev_name;
  // scratch
  DBuf d_in, d_local, d_pairA, d_pairB, d_partials, d_partcnt, d_probs, d_states, d_tables, d_init, d_misc;
  DBuf d_f_tab, d_f_pslot, d_f_flags, d_f_blk, d_f_jobs, d_f_cnt, d_f_map, d_f_icp;
  DBuf d_ins_g, d_ins_slot, d_ins_next, d_ins_jobs, d_queue, d_tchan, d_maps;
  HBuf h_misc, h_states, h_stage;
  // profiling
  bool prof_on = false;
  mlo_profile prof{};
  std::vector<cudaEvent_t> ev_pool;
  size_t ev_used = 0;
  struct Span {
    int bucket;
    size_t e0, e1;
  };
  std::vector<Span> spans;
};

struct mlo_map {
  mlo_ctx* ctx = nullptr;
  mlo_map_params prm{};
  MapDev dev{};
  MapDev alt{};  // second buffer set for filtered rebuilds (allocated on first cull)
  bool alt_ready = false;
  uint64_t table_size = 0;  // number of 32-byte buckets
  int32_t* head = nullptr;  // per-cell (4 per bucket) scratch list heads for insert
  uint64_t n_voxels = 0, n_points = 0;  // as of the last digest of the device counters
  uint64_t hwm = 0;                     // voxel ids handed out so far (counters[0]) as of the last digest
  uint64_t n_free = 0;                  // reusable voxel ids on the free stack (counters[3]) as of the last digest
  uint32_t n_grown = 0;                 // how many times the map was re-hashed into larger buffers
};

struct mlo_dcloud {
  mlo_ctx* ctx = nullptr;
  float4* pts = nullptr;
  uint64_t n = 0;
  std::vector<uint64_t> offsets;  // n_clouds + 1
};

// Device-resident layers of n_slots scans (include/mlo_b200.h "scan sets").  A filter call packs its raw clouds
// back to back; slot s keeps (offset, sizes) of where its layers sit in the shared buffers.
struct mlo_scanset {
  mlo_ctx* ctx = nullptr;
  struct Slot {
    bool valid = false;
    uint64_t off = 0;  // in points, common to every buffer
    uint32_t n_raw = 0, n_map = 0, n_icp = 0;
  };
  std::vector<Slot> slots;
  bool skewed = false;     // the last filter call produced "_skewed" layers (x, y, z, t)
  bool deskewed = false;   // ... and mlo_scanset_deskew has produced the final layers since
  DBuf raw, tchan, mapS, icpS, mapL, icpL, cnt, bbox_jobs, bbox_out;
  // prefetch of the NEXT step's raw clouds (mlo_scanset_prefetch): second raw buffer filled on the copy stream
  struct Staged {
    const float* src;
    uint64_t n;
  };
  struct Announce {
    std::vector<Staged> clouds;
    uint32_t stride = 0;
    bool valid = false;
    void clear() {
      clouds.clear();
      valid = false;
    }
  };
  DBuf raw_next;
  Announce pend;  // announced, transfer not yet enqueued (waits for the next compute call)
  Announce fly;   // transfer enqueued on the copy stream into raw_next
  cudaEvent_t staged_ready = nullptr;
  const float4* map_layer() const { return mapL.as<float4>(); }
  const float4* icp_layer() const { return icpL.as<float4>(); }
};

namespace {

int fail(mlo_ctx* c, int code, const std::string& msg) {
  if (c) c->err = msg;
  return code;
}
#define CU(ctx, call)                                                                                          \
  do {                                                                                                         \
    cudaError_t e__ = (call);                                                                                  \
    if (e__ != cudaSuccess)                                                                                    \
      return fail(ctx, MLO_ERR_CUDA, std::string(#call) + ": " + cudaGetErrorString(e__) + " @" + std::to_string(__LINE__)); \
  } while (0)
#define LAUNCH(ctx, kern, grid, block, ...)            \
  do {                                                 \
    kern<<<grid, block, 0, (ctx)->stream>>>(__VA_ARGS__); \
    (ctx)->launches++;                                 \
  } while (0)
#define LAUNCH_SMEM(ctx, kern, grid, block, smem, ...)          \
  do {                                                          \
    kern<<<grid, block, smem, (ctx)->stream>>>(__VA_ARGS__);    \
    (ctx)->launches++;                                          \
  } while (0)
#define LAUNCH_ON(ctx, strm, kern, grid, block, ...) \
  do {                                               \
    kern<<<grid, block, 0, strm>>>(__VA_ARGS__);     \
    (ctx)->launches++;                               \
  } while (0)

struct DeviceGuard {
  int prev = 0;
  explicit DeviceGuard(int d) {
    cudaGetDevice(&prev);
    if (prev != d) cudaSetDevice(d);
    cur = d;
  }
  ~DeviceGuard() {
    if (prev != cur) cudaSetDevice(prev);
  }
  int cur;
};

uint64_t next_pow2(uint64_t v) {
  uint64_t p = 1;
  while (p < v) p <<= 1;
  return p;
}

size_t prof_begin(mlo_ctx* c) {
  if (!c->prof_on) return 0;
  if (c->ev_used >= c->ev_pool.size()) {
    cudaEvent_t e;
    cudaEventCreate(&e);
    c->ev_pool.push_back(e);
  }
  cudaEventRecord(c->ev_pool[c->ev_used], c->stream);
  return c->ev_used++;
}
void prof_end(mlo_ctx* c, int bucket, size_t e0) {
  if (!c->prof_on) return;
  const size_t e1 = prof_begin(c);
  c->spans.push_back({bucket, e0, e1});
}
void prof_collect(mlo_ctx* c) {  // call after a stream synchronize
  if (!c->prof_on) return;
  for (auto& s : c->spans) {
    float ms = 0.f;
    cudaEventElapsedTime(&ms, c->ev_pool[s.e0], c->ev_pool[s.e1]);
    switch (s.bucket) {
      case 0: c->prof.filter_1st_ms += ms; break;
      case 1: c->prof.run_icp_ms += ms; break;
      case 2: c->prof.update_local_map_ms += ms; break;
      case 3: c->prof.nn_kernel_ms += ms; c->prof.nn_kernel_launches++; break;
    }
  }
  c->spans.clear();
  c->ev_used = 0;
}

int alloc_map_buffers(mlo_ctx* c, const mlo_map_params& p, uint64_t table_size, MapDev& d) {
  uint32_t cap = p.max_points_per_voxel;
  if (cap == 0 || cap > HARD_LIMIT_PTS) cap = HARD_LIMIT_PTS;
  d.cap = cap;
  d.row = (cap + 1u) & ~1u;
  d.capacity_voxels = uint32_t(p.capacity_voxels);
  d.inv_voxel = 1.0f / p.voxel_size;
  d.voxel_size = p.voxel_size;
  d.min_dist2 = p.min_distance_between_points * p.min_distance_between_points;
  d.eig_ratio = p.max_eigen_ratio_for_planes;
  d.min_pts_plane = p.min_points_for_plane ? p.min_points_for_plane : 5;
  d.kind = p.kind;
  d.index_floor = c->conv_index_floor;
  d.cull_metric = c->conv_cull_metric;
  d.mask = table_size - 1;
  CU(c, cudaMallocAsync(&d.buckets, table_size * sizeof(Bucket), c->stream));
  CU(c, cudaMallocAsync(&d.pts, size_t(p.capacity_voxels) * d.row * sizeof(float4), c->stream));
  CU(c, cudaMallocAsync(&d.counters, MAP_COUNTERS * sizeof(uint32_t), c->stream));
  CU(c, cudaMallocAsync(&d.vkey, size_t(p.capacity_voxels) * sizeof(unsigned long long), c->stream));
  CU(c, cudaMallocAsync(&d.free_ids, size_t(p.capacity_voxels) * sizeof(uint32_t), c->stream));
  d.mean = d.normal = nullptr;
  if (p.kind == MLO_MAP_NDT) {
    CU(c, cudaMallocAsync(&d.mean, size_t(p.capacity_voxels) * sizeof(float4), c->stream));
    CU(c, cudaMallocAsync(&d.normal, size_t(p.capacity_voxels) * sizeof(float4), c->stream));
  }
  return MLO_OK;
}
int clear_map_buffers(mlo_ctx* c, MapDev& d, uint64_t table_size) {
  CU(c, cudaMemsetAsync(d.buckets, 0xFF, table_size * sizeof(Bucket), c->stream));
  CU(c, cudaMemsetAsync(d.counters, 0, MAP_COUNTERS * sizeof(uint32_t), c->stream));
  return MLO_OK;
}
// Map buffers come from the device's stream-ordered memory pool (release threshold = unlimited, set in mlo_create):
// a LidarOdometry instance that is destroyed hands its ~0.7 GB back to the pool and the next one gets it without
// a driver allocation or a device synchronisation.
void free_map_buffers(mlo_ctx* c, MapDev& d) {
  if (d.buckets) cudaFreeAsync(d.buckets, c->stream);
  if (d.vkey) cudaFreeAsync(d.vkey, c->stream);
  if (d.free_ids) cudaFreeAsync(d.free_ids, c->stream);
  if (d.pts) cudaFreeAsync(d.pts, c->stream);
  if (d.counters) cudaFreeAsync(d.counters, c->stream);
  if (d.mean) cudaFreeAsync(d.mean, c->stream);
  if (d.normal) cudaFreeAsync(d.normal, c->stream);
  d = MapDev{};
}

int map_rebuild(mlo_map* m, bool use_filter, int32_t sx, int32_t sy, int32_t sz, int32_t d);

// Interpret one host copy of a map's counters (after a stream synchronize): error bits, statistics and the
// compaction trigger (claimed column buckets above half of the table -> rebuild into the second buffer set).
int digest_map_counters(mlo_map* m, const uint32_t* h) {
  mlo_ctx* c = m->ctx;
  if (h[2] & ERR_KEY_RANGE) return fail(c, MLO_ERR_KEY_RANGE, "voxel index outside the packed 21-bit range");
  if ((h[2] & ERR_CAPACITY) || h[0] > m->dev.capacity_voxels)
    return fail(c, MLO_ERR_CAPACITY, "map voxel capacity exhausted (" + std::to_string(h[0]) + " > " +
                                         std::to_string(m->dev.capacity_voxels) + ")");
  m->n_voxels = uint64_t(h[0]) - uint64_t(h[3]);
  m->n_points = h[1];
  m->hwm = h[0];
  m->n_free = h[3];
  if (uint64_t(h[4]) * 2 > m->table_size) return map_rebuild(m, false, 0, 0, 0, 0);
  return MLO_OK;
}

int check_map_errors(mlo_map* m) {
  mlo_ctx* c = m->ctx;
  CU(c, c->h_misc.ensure(64));
  uint32_t* h = c->h_misc.as<uint32_t>();
  CU(c, cudaMemcpyAsync(h, m->dev.counters, MAP_COUNTERS * sizeof(uint32_t), cudaMemcpyDeviceToHost, c->stream));
  CU(c, cudaStreamSynchronize(c->stream));
  return digest_map_counters(m, h);
}

// device-side insert of n points already on the device (float array with stride)
int map_insert_device(mlo_map* m, const float* d_pts, uint32_t stride, uint64_t n, const double pose[12]) {
  mlo_ctx* c = m->ctx;
  if (n == 0) return MLO_OK;
  CU(c, c->d_ins_g.ensure(n * sizeof(float4)));
  CU(c, c->d_ins_slot.ensure(n * sizeof(uint32_t)));
  CU(c, c->d_ins_next.ensure(n * sizeof(int32_t)));
  Pose34 T;
  std::memcpy(T.m, pose, sizeof(T.m));
  const uint32_t nb = uint32_t((n + 255) / 256);
  LAUNCH(c, k_insert_link, nb, 256, m->dev, d_pts, stride, uint32_t(n), T, c->d_ins_g.as<float4>(),
         c->d_ins_slot.as<uint32_t>(), m->head, c->d_ins_next.as<int32_t>());
  LAUNCH(c, k_insert_commit, nb, 256, m->dev, uint32_t(n), c->d_ins_g.as<float4>(), c->d_ins_slot.as<uint32_t>(),
         m->head, c->d_ins_next.as<int32_t>());
  CU(c, cudaGetLastError());
  return MLO_OK;
}

// in-place cull (no host synchronisation): thread per voxel id, early exit above the device-side high-water mark
// `new_since_digest` = upper bound of the voxel ids handed out since the counters were last read by the host
int map_cull_device(mlo_map* m, int32_t sx, int32_t sy, int32_t sz, int32_t d, uint64_t new_since_digest) {
  mlo_ctx* c = m->ctx;
  CU(c, cudaMemsetAsync(m->dev.counters + 5, 0, sizeof(uint32_t), c->stream));
  const uint64_t bound = std::min<uint64_t>(m->dev.capacity_voxels, m->hwm + new_since_digest);
  if (bound == 0) return MLO_OK;
  LAUNCH(c, k_cull_inplace, uint32_t((bound + 255) / 256), 256, m->dev, sx, sy, sz, d);
  CU(c, cudaGetLastError());
  return MLO_OK;
}

int map_rebuild(mlo_map* m, bool use_filter, int32_t sx, int32_t sy, int32_t sz, int32_t d) {
  mlo_ctx* c = m->ctx;
  if (!m->alt_ready) {
    int rc = alloc_map_buffers(c, m->prm, m->table_size, m->alt);
    if (rc != MLO_OK) return rc;
    m->alt.index_floor = m->dev.index_floor;  // (a map keeps the conventions it was created with)
    m->alt.cull_metric = m->dev.cull_metric;
    m->alt_ready = true;
  }
  int rc = clear_map_buffers(c, m->alt, m->table_size);
  if (rc != MLO_OK) return rc;
  const uint64_t threads = m->table_size;
  LAUNCH(c, k_rebuild, uint32_t((threads + 255) / 256), 256, m->dev, m->alt, m->table_size, sx, sy, sz, d,
         use_filter ? 1 : 0);
  CU(c, cudaGetLastError());
  std::swap(m->dev, m->alt);
  return MLO_OK;
}

// upstream's HashedVoxelPointCloud is unbounded; here `capacity_voxels` is only the INITIAL size of the voxel payload
// and of the hash table.  Before an insert of n points (each can open at most one voxel) the map is re-hashed into
// buffers of twice the size (or more) if the worst case would not fit: inserts never fail for lack of capacity, and a
// map that stays small keeps its table small (fewer pages and cache lines under the random probes of the NN search).
int map_ensure_capacity(mlo_map* m, uint64_t n_new_points) {
  mlo_ctx* c = m->ctx;
  const uint64_t worst = m->hwm - std::min(m->hwm, m->n_free) + n_new_points;
  if (worst <= m->dev.capacity_voxels) return MLO_OK;
  uint64_t cap = m->dev.capacity_voxels;
  while (cap < worst + worst / 4) cap *= 2;
  if (cap > (1ull << 26)) cap = 1ull << 26;
  if (cap < worst) return fail(c, MLO_ERR_CAPACITY, "map would exceed 2^26 voxels");
  mlo_map_params np = m->prm;
  np.capacity_voxels = cap;
  const uint64_t new_table = next_pow2(std::max<uint64_t>(uint64_t(c->table_factor) * cap, 1024));
  if (m->alt_ready) {
    free_map_buffers(c, m->alt);
    m->alt_ready = false;
  }
  MapDev nd{};
  int rc = alloc_map_buffers(c, np, new_table, nd);
  nd.index_floor = m->dev.index_floor;  // (a map keeps the conventions it was created with)
  nd.cull_metric = m->dev.cull_metric;
  if (rc == MLO_OK) rc = clear_map_buffers(c, nd, new_table);
  if (rc != MLO_OK) return rc;
  LAUNCH(c, k_rebuild, uint32_t((m->table_size + 255) / 256), 256, m->dev, nd, m->table_size, 0, 0, 0, 0, 0);
  CU(c, cudaGetLastError());
  free_map_buffers(c, m->dev);  // (stream-ordered: after the rebuild has read them)
  if (m->head) cudaFreeAsync(m->head, c->stream);
  m->head = nullptr;
  CU(c, cudaMallocAsync(&m->head, 4 * new_table * sizeof(int32_t), c->stream));
  CU(c, cudaMemsetAsync(m->head, 0xFF, 4 * new_table * sizeof(int32_t), c->stream));
  m->dev = nd;
  m->prm = np;
  m->table_size = new_table;
  m->hwm = m->hwm - std::min(m->hwm, m->n_free);  // ids are dense again: an upper bound of the new high-water mark
  m->n_free = 0;
  m->n_grown++;
  return MLO_OK;
}

void fill_pred(PointPred& q, const mlo_decimate_params& p, bool on) {
  q.use_range = on && p.use_range;
  q.rmin2 = p.range_min * p.range_min;
  q.rmax2 = p.range_max * p.range_max;
  q.use_bbox = on && p.use_bbox_outside;
  for (int k = 0; k < 3; k++) {
    q.bmin[k] = p.bbox_min[k];
    q.bmax[k] = p.bbox_max[k];
  }
}

// Device filter pipeline over a batch of raw clouds resident on the device (filter.cuh):
//   stage 1: decimate(raw, for_map.resolution), then the by-range / bbox predicates   -> map layer
//   stage 2: decimate(map layer, for_icp.resolution)                                   -> icp layer
// (the predicates sit between the two FilterDecimateVoxels of pipelines/lidar3d-default.yaml:285-319: applying them
// to the winners of stage 1 and decimating the survivors is the same computation).
// Outputs per cloud b live at [out_off[b], ...) of d_f_map / d_f_icp with device counts in d_f_cnt
// (layout per cloud: [n_single, n_map, n_icp, npred1, npred2, err]).
// Clouds are processed in GROUPS whose scratch (hash tables + candidate slices) fits the L2 with room to spare: the
// memset that clears a group's tables, the claims, and the winner tests all hit L2-resident lines, and the next group
// reuses the same addresses.  DRAM sees the raw clouds once and the two layers once.
struct FilterBatch {
  uint32_t n_clouds = 0;
  std::vector<uint64_t> out_off;  // per-cloud output offset (== raw offset: outputs never exceed inputs)
  uint32_t max_n = 0;
  // where the layers and the counters go; null = the context's scratch (d_f_map / d_f_icp / d_f_cnt)
  DBuf* out_map = nullptr;
  DBuf* out_icp = nullptr;
  DBuf* out_cnt = nullptr;
  // what was run (a retry with conservative table sizes repeats exactly this)
  const float* d_raw = nullptr;
  uint32_t stride = 0;
  const mlo_filter1_params* fps = nullptr;
  bool single = false;
  uint32_t* d_idx_out = nullptr;
  const float* d_t = nullptr;
  bool used_cta = false;  // the batch ran k_decim_cta (a raised ERR_CTA_FALLBACK repeats it with the global-table kernels)
};
constexpr uint32_t CNT_STRIDE = 8;
constexpr size_t CTA_DECIM_SMEM = 220 * 1024;  // dynamic shared memory of k_decim_cta: table + bitmap

// The parts of a cloud's two decimation jobs that do not depend on the scratch layout (inputs, predicates, outputs).
// j1 / j2 must be zero-initialised by the caller or carry only scratch pointers.
void fill_decim_jobs(mlo_ctx* c, DecimJob& j1, DecimJob& j2, uint32_t b, const float* d_raw, uint32_t stride, const uint64_t* offsets,
                     const mlo_filter1_params* fps, bool single_decimate_idx, uint32_t* d_idx_out, const float* d_t, DBuf& b_map,
                     DBuf& b_icp, uint32_t* cnt) {
  const uint32_t n = uint32_t(offsets[b + 1] - offsets[b]);
  j1.err = j2.err = cnt + b * CNT_STRIDE + 5;
  j1.in = d_raw + offsets[b] * stride;
  j1.in_t = d_t ? d_t + offsets[b] : nullptr;
  j1.in_stride = stride;
  j1.n_in_static = n;
  j1.index_floor = j2.index_floor = c->conv_index_floor;
  j1.resolution = fps[b].for_map.voxel_filter_resolution;
  j1.min_pts = fps[b].for_map.minimum_input_points_to_filter;
  j1.npred = cnt + b * CNT_STRIDE + 3;
  if (single_decimate_idx) {  // mlo_voxel_decimate_first: predicates in front of ONE decimation, indices out
    fill_pred(j1.pre, fps[b].for_map, true);
    j1.out = b_map.as<float4>() + offsets[b];
    j1.n_out = cnt + b * CNT_STRIDE + 0;
    j1.out_idx = d_idx_out + offsets[b];
  } else {
    fill_pred(j1.post, fps[b].for_icp, true);
    j1.out = b_map.as<float4>() + offsets[b];
    j1.n_out = cnt + b * CNT_STRIDE + 1;
    j2.in = reinterpret_cast<const float*>(b_map.as<float4>() + offsets[b]);
    j2.in_stride = 4;
    j2.keep_w = d_t ? 1 : 0;
    j2.n_in_dev = cnt + b * CNT_STRIDE + 1;
    j2.resolution = fps[b].for_icp.voxel_filter_resolution;
    j2.min_pts = fps[b].for_icp.minimum_input_points_to_filter;
    j2.npred = cnt + b * CNT_STRIDE + 4;
    j2.out = b_icp.as<float4>() + offsets[b];
    j2.n_out = cnt + b * CNT_STRIDE + 2;
  }
}

int run_filter_batch(mlo_ctx* c, const float* d_raw, uint32_t stride, uint32_t n_clouds, const uint64_t* offsets,
                     const mlo_filter1_params* fps, bool single_decimate_idx, uint32_t* d_idx_out, FilterBatch& fb,
                     const float* d_t = nullptr, bool conservative = false, bool force_global = false) {
  fb.n_clouds = n_clouds;
  fb.out_off.assign(offsets, offsets + n_clouds + 1);
  fb.d_raw = d_raw;
  fb.stride = stride;
  fb.fps = fps;
  fb.single = single_decimate_idx;
  fb.d_idx_out = d_idx_out;
  fb.d_t = d_t;
  const uint64_t total = offsets[n_clouds];
  uint32_t max_n = 0;
  for (uint32_t b = 0; b < n_clouds; b++) max_n = std::max<uint32_t>(max_n, uint32_t(offsets[b + 1] - offsets[b]));
  fb.max_n = max_n;
  DBuf& b_map = fb.out_map ? *fb.out_map : c->d_f_map;
  DBuf& b_icp = fb.out_icp ? *fb.out_icp : c->d_f_icp;
  DBuf& b_cnt = fb.out_cnt ? *fb.out_cnt : c->d_f_cnt;
  CU(c, b_cnt.ensure(std::max<size_t>(size_t(n_clouds), 1) * CNT_STRIDE * sizeof(uint32_t)));
  uint32_t* cnt = b_cnt.as<uint32_t>();
  CU(c, cudaMemsetAsync(cnt, 0, std::max<size_t>(size_t(n_clouds), 1) * CNT_STRIDE * sizeof(uint32_t), c->stream));
  if (total == 0 || max_n == 0) return MLO_OK;
  // ---- which kernels: one block per cloud with shared-memory scratch (large batches), or global scratch tables
  const size_t bitmap_words = (size_t(max_n) + 31) / 32;
  const bool cta_fits = bitmap_words * 4 + 4096 * 8 <= CTA_DECIM_SMEM;
  bool use_cta = false;
  if (!force_global && !conservative && cta_fits) {
    if (c->filter_kernel == 2) use_cta = true;
    else if (c->filter_kernel == 0 && int(n_clouds) >= c->filter_cta_min_clouds) {
      if (c->filter_cta_backoff > 0) c->filter_cta_backoff--;
      else use_cta = true;
    }
  }
  fb.used_cta = use_cta;
  c->last_filter_kernel = use_cta ? 2 : 1;
  if (use_cta) {
    const uint32_t tab_cap = uint32_t((CTA_DECIM_SMEM - bitmap_words * 4) / 8);
    if (!c->cta_attr_set) {
      CU(c, cudaFuncSetAttribute(k_decim_cta<4>, cudaFuncAttributeMaxDynamicSharedMemorySize, int(CTA_DECIM_SMEM)));
      c->cta_attr_set = true;
    }
    CU(c, b_map.ensure(total * sizeof(float4)));
    CU(c, b_icp.ensure(total * sizeof(float4)));
    CU(c, c->d_f_jobs.ensure(2 * size_t(n_clouds) * sizeof(DecimJob)));
    CU(c, c->h_stage.ensure(2 * size_t(n_clouds) * sizeof(DecimJob)));
    DecimJob* hj = c->h_stage.as<DecimJob>();
    std::memset(hj, 0, 2 * size_t(n_clouds) * sizeof(DecimJob));
    for (uint32_t b = 0; b < n_clouds; b++)
      fill_decim_jobs(c, hj[b], hj[n_clouds + b], b, d_raw, stride, offsets, fps, single_decimate_idx, d_idx_out, d_t, b_map, b_icp, cnt);
    CU(c, cudaMemcpyAsync(c->d_f_jobs.p, hj, 2 * size_t(n_clouds) * sizeof(DecimJob), cudaMemcpyHostToDevice, c->stream));
    const DecimJob* dj1 = c->d_f_jobs.as<DecimJob>();
    for (int stage = 0; stage < (single_decimate_idx ? 1 : 2); stage++)
      LAUNCH_SMEM(c, k_decim_cta<4>, n_clouds, CTA_DECIM_THREADS, CTA_DECIM_SMEM, dj1 + size_t(stage) * n_clouds, tab_cap,
                  uint32_t(bitmap_words));
    CU(c, cudaGetLastError());
    return MLO_OK;
  }
  // ---- per-cloud scratch geometry and the groups
  // stage-1 table: half an entry per input point (a 0.5 m grid keeps a third of a 64-beam sweep at most: load factor
  // <= 0.6); stage-2 table: an eighth (its input is the map layer, its output a few thousand points).  A cloud that
  // overflows either (hardly decimated at all) raises ERR_CAPACITY and the batch is run again with `conservative`
  // tables of one entry per input point (filter_counts).
  const int ppt = (c->filter_ppt == 1 || c->filter_ppt == 2) ? c->filter_ppt : 4;
  const uint64_t tile = uint64_t(DECIM_BLOCK) * ppt;
  struct Geo {
    uint64_t tab1, tab2, nblk, pts;
  };
  std::vector<Geo> geo(n_clouds);
  for (uint32_t b = 0; b < n_clouds; b++) {
    const uint64_t n = offsets[b + 1] - offsets[b];
    geo[b].pts = n;
    geo[b].nblk = (n + tile - 1) / tile;
    geo[b].tab1 = next_pow2(std::max<uint64_t>(conservative ? n : n / 2, 1024));
    geo[b].tab2 = single_decimate_idx ? 0 : next_pow2(std::max<uint64_t>(conservative ? n : n / 8, 1024));
  }
  const uint64_t budget = uint64_t(std::max(1, c->filter_group_mb)) << 20;
  std::vector<uint32_t> group_begin{0};
  uint64_t g_tab = 0, g_pts = 0, g_blk = 0, max_tab = 0, max_pts = 0, max_blk = 0, acc = 0;
  for (uint32_t b = 0; b < n_clouds; b++) {
    const uint64_t bytes = (geo[b].tab1 + geo[b].tab2) * sizeof(DecimEntry) + geo[b].pts * (sizeof(float4) + sizeof(uint2));
    if (b > group_begin.back() && acc + bytes > budget) {
      group_begin.push_back(b);
      acc = g_tab = g_pts = g_blk = 0;
    }
    acc += bytes;
    g_tab += geo[b].tab1 + geo[b].tab2;
    g_pts += geo[b].pts;
    g_blk += geo[b].nblk;
    max_tab = std::max(max_tab, g_tab);
    max_pts = std::max(max_pts, g_pts);
    max_blk = std::max(max_blk, g_blk);
  }
  group_begin.push_back(n_clouds);
  CU(c, c->d_f_tab.ensure(max_tab * sizeof(DecimEntry)));
  CU(c, c->d_f_pslot.ensure(max_pts * sizeof(float4)));   // candidate points
  CU(c, c->d_f_flags.ensure(max_pts * sizeof(uint2)));    // candidate (slot, index)
  CU(c, c->d_f_blk.ensure(max_blk * (2 * sizeof(unsigned long long) + sizeof(uint32_t))));  // status x2 stages + blockcnt
  CU(c, b_map.ensure(total * sizeof(float4)));
  CU(c, b_icp.ensure(total * sizeof(float4)));
  CU(c, c->d_f_jobs.ensure(2 * size_t(n_clouds) * sizeof(DecimJob)));
  CU(c, c->h_stage.ensure(2 * size_t(n_clouds) * sizeof(DecimJob)));

  DecimEntry* tab = c->d_f_tab.as<DecimEntry>();
  float4* cand_pt = c->d_f_pslot.as<float4>();
  uint2* cand_meta = c->d_f_flags.as<uint2>();
  unsigned long long* status = c->d_f_blk.as<unsigned long long>();
  uint32_t* blockcnt = reinterpret_cast<uint32_t*>(status + 2 * max_blk);
  DecimJob* hj = c->h_stage.as<DecimJob>();
  struct GroupPlan {
    uint32_t b0, b1, nblk_max;
    uint64_t tab_entries, blocks;
  };
  std::vector<GroupPlan> plan;
  for (size_t g = 0; g + 1 < group_begin.size(); g++) {
    const uint32_t b0 = group_begin[g], b1 = group_begin[g + 1];
    uint64_t t_off = 0, p_off = 0, k_off = 0;
    uint32_t nblk_max = 0;
    for (uint32_t b = b0; b < b1; b++) {
      const uint32_t n = uint32_t(geo[b].pts);
      DecimJob& j1 = hj[b];
      DecimJob& j2 = hj[n_clouds + b];
      std::memset(&j1, 0, sizeof(j1));
      std::memset(&j2, 0, sizeof(j2));
      // scratch shared by the two stages of a cloud (stage 2 starts when stage 1 has finished): candidate slices and
      // block counts; own tables and own look-back words (both are cleared once per group)
      j1.cand_pt = j2.cand_pt = cand_pt + p_off;
      j1.cand_meta = j2.cand_meta = cand_meta + p_off;
      j1.blockcnt = j2.blockcnt = blockcnt + k_off;
      j1.status = status + k_off;
      j2.status = status + max_blk + k_off;
      j1.tab = tab + t_off;
      j1.tab_mask = uint32_t(geo[b].tab1 - 1);
      j2.tab = tab + t_off + geo[b].tab1;
      j2.tab_mask = geo[b].tab2 ? uint32_t(geo[b].tab2 - 1) : 0u;
      fill_decim_jobs(c, j1, j2, b, d_raw, stride, offsets, fps, single_decimate_idx, d_idx_out, d_t, b_map, b_icp, cnt);
      t_off += geo[b].tab1 + geo[b].tab2;
      p_off += geo[b].pts;
      k_off += geo[b].nblk;
      nblk_max = std::max<uint32_t>(nblk_max, uint32_t(geo[b].nblk));
    }
    plan.push_back({b0, b1, nblk_max, t_off, k_off});
  }
  CU(c, cudaMemcpyAsync(c->d_f_jobs.p, hj, 2 * size_t(n_clouds) * sizeof(DecimJob), cudaMemcpyHostToDevice, c->stream));
  const DecimJob* dj1 = c->d_f_jobs.as<DecimJob>();
  const DecimJob* dj2 = dj1 + n_clouds;
  for (const GroupPlan& g : plan) {
    if (g.nblk_max == 0) continue;
    CU(c, cudaMemsetAsync(tab, 0xFF, g.tab_entries * sizeof(DecimEntry), c->stream));
    CU(c, cudaMemsetAsync(status, 0, g.blocks * sizeof(unsigned long long), c->stream));
    if (!single_decimate_idx) CU(c, cudaMemsetAsync(status + max_blk, 0, g.blocks * sizeof(unsigned long long), c->stream));
    const dim3 grid(g.nblk_max, g.b1 - g.b0);
#define MLO_DECIM_STAGE(PPT, JOBS)                                   \
  do {                                                               \
    LAUNCH(c, k_decim_claim<PPT>, grid, DECIM_BLOCK, (JOBS) + g.b0);    \
    LAUNCH(c, k_decim_finalize<PPT>, grid, DECIM_BLOCK, (JOBS) + g.b0); \
  } while (0)
    for (int stage = 0; stage < (single_decimate_idx ? 1 : 2); stage++) {
      const DecimJob* js = stage == 0 ? dj1 : dj2;
      if (ppt == 1) MLO_DECIM_STAGE(1, js);
      else if (ppt == 2) MLO_DECIM_STAGE(2, js);
      else MLO_DECIM_STAGE(4, js);
    }
#undef MLO_DECIM_STAGE
  }
  CU(c, cudaGetLastError());
  return MLO_OK;
}

// Copy the per-cloud counters of a filter batch back (one synchronisation) and check its error bits.  A scratch hash
// table that ran full (a cloud whose first decimation keeps more than a quarter of its points) is not an error of the
// caller: the batch is run again with tables sized for the worst case.
int filter_counts(mlo_ctx* c, FilterBatch& fb, std::vector<uint32_t>& h) {
  const DBuf& b_cnt = fb.out_cnt ? *fb.out_cnt : c->d_f_cnt;
  h.assign(std::max<size_t>(fb.n_clouds, 1) * CNT_STRIDE, 0u);
  bool conservative = false;
  for (int attempt = 0; attempt < 3; attempt++) {
    CU(c, cudaMemcpyAsync(h.data(), b_cnt.p, h.size() * sizeof(uint32_t), cudaMemcpyDeviceToHost, c->stream));
    CU(c, cudaStreamSynchronize(c->stream));
    bool full = false, fallback = false;
    for (uint32_t b = 0; b < fb.n_clouds; b++) {
      if (h[b * CNT_STRIDE + 5] & ERR_KEY_RANGE) return fail(c, MLO_ERR_KEY_RANGE, "voxel index outside the packed 21-bit range");
      full = full || (h[b * CNT_STRIDE + 5] & ERR_CAPACITY);
      fallback = fallback || (h[b * CNT_STRIDE + 5] & ERR_CTA_FALLBACK);
    }
    if (!full && !fallback) return MLO_OK;
    if (fallback && fb.used_cta) {
      // a cloud outside k_decim_cta's 32-bit key box, or more voxels than its shared-memory table holds: the global-table
      // kernels take the batch, and the next batches of this context go to them directly for a while
      if (c->filter_kernel == 0) c->filter_cta_backoff = 64;
    } else {
      if (conservative) return fail(c, MLO_ERR_CAPACITY, "decimation scratch table exhausted");
      conservative = true;
    }
    std::vector<uint64_t> offs = fb.out_off;
    int rc = run_filter_batch(c, fb.d_raw, fb.stride, fb.n_clouds, offs.data(), fb.fps, fb.single, fb.d_idx_out, fb, fb.d_t, conservative, true);
    if (rc != MLO_OK) return rc;
  }
  return fail(c, MLO_ERR_CAPACITY, "decimation scratch table exhausted");
}

__global__ void k_to_float4(const float* __restrict__ src, uint32_t stride, uint64_t n, float4* __restrict__ dst) {
  const uint64_t i = uint64_t(blockIdx.x) * blockDim.x + threadIdx.x;
  if (i >= n) return;
  const float* p = src + i * stride;
  dst[i] = make_float4(p[0], p[1], p[2], 0.f);
}
__global__ void k_soa_to_float4(const float* __restrict__ x, const float* __restrict__ y, const float* __restrict__ z, uint64_t n,
                                float4* __restrict__ dst) {
  const uint64_t i = uint64_t(blockIdx.x) * blockDim.x + threadIdx.x;
  if (i >= n) return;
  dst[i] = make_float4(x[i], y[i], z[i], 0.f);
}

// axis-aligned bounding box of each listed cloud (one block per cloud): out[6*j] = min xyz, max xyz
struct BBoxJob {
  const float4* p;
  uint32_t n;
};
__global__ void __launch_bounds__(256) k_bbox(const BBoxJob* __restrict__ jobs, float* __restrict__ out) {
  const BBoxJob j = jobs[blockIdx.x];
  float mn[3] = {INFINITY, INFINITY, INFINITY}, mx[3] = {-INFINITY, -INFINITY, -INFINITY};
  for (uint32_t i = threadIdx.x; i < j.n; i += blockDim.x) {
    const float4 p = __ldg(&j.p[i]);
    mn[0] = fminf(mn[0], p.x); mn[1] = fminf(mn[1], p.y); mn[2] = fminf(mn[2], p.z);
    mx[0] = fmaxf(mx[0], p.x); mx[1] = fmaxf(mx[1], p.y); mx[2] = fmaxf(mx[2], p.z);
  }
  __shared__ float s[8][6];
  for (int k = 0; k < 3; k++)
    for (int o = 16; o > 0; o >>= 1) {
      mn[k] = fminf(mn[k], __shfl_xor_sync(0xFFFFFFFFu, mn[k], o));
      mx[k] = fmaxf(mx[k], __shfl_xor_sync(0xFFFFFFFFu, mx[k], o));
    }
  const uint32_t w = threadIdx.x >> 5;
  if ((threadIdx.x & 31u) == 0)
    for (int k = 0; k < 3; k++) {
      s[w][k] = mn[k];
      s[w][3 + k] = mx[k];
    }
  __syncthreads();
  if (threadIdx.x < 6) {
    float v = s[0][threadIdx.x];
    for (uint32_t q = 1; q < blockDim.x / 32; q++) v = threadIdx.x < 3 ? fminf(v, s[q][threadIdx.x]) : fmaxf(v, s[q][threadIdx.x]);
    out[6 * blockIdx.x + threadIdx.x] = v;
  }
}

int upload_strided(mlo_ctx* c, const float* pts, uint32_t stride, uint64_t n, DBuf& dst) {
  if (stride != 3 && stride != 4) return fail(c, MLO_ERR_INVALID_ARG, "stride_floats must be 3 or 4");
  CU(c, dst.ensure(std::max<size_t>(n * stride * sizeof(float), 16)));
  if (n) CU(c, cudaMemcpyAsync(dst.p, pts, n * stride * sizeof(float), cudaMemcpyHostToDevice, c->stream));
  return MLO_OK;
}

int issue_stage_upload(mlo_ctx* c, int slot) {
  auto& st = c->stage[slot];
  if (!st.deferred_src) return MLO_OK;
  const float* raw = st.deferred_src;
  st.deferred_src = nullptr;
  const uint64_t total = st.offsets.back();
  const size_t bytes = std::max<size_t>(total * st.stride * sizeof(float), 16);
  // No ordering against the context stream: the last reader of this slot was the filter of an EARLIER
  // mlo_scan_register_batch_staged call, and every such call returns only after a stream synchronisation.  (Waiting on
  // the context stream here would put the copy behind the ICP kernel of the call in progress - the one it is meant to
  // overlap - on the single-launch align paths.)
  if (st.buf.cap < bytes) {
    CU(c, cudaStreamSynchronize(c->stream));
    CU(c, cudaStreamSynchronize(c->copy_stream));
    CU(c, st.buf.ensure(bytes));
  }
  if (total)
    CU(c, cudaMemcpyAsync(st.buf.p, raw, total * st.stride * sizeof(float), cudaMemcpyHostToDevice, c->copy_stream));
  CU(c, cudaEventRecord(st.ready, c->copy_stream));
  return MLO_OK;
}
int issue_deferred_uploads(mlo_ctx* c, int except_slot) {
  if (!c->deferred.empty()) {
    std::vector<mlo_ctx::Deferred> fns;
    fns.swap(c->deferred);
    for (auto& d : fns) {
      int rc = d.fn();
      if (rc != MLO_OK) return rc;
    }
  }
  for (int s = 0; s < 2; s++)
    if (s != except_slot) {
      int rc = issue_stage_upload(c, s);
      if (rc != MLO_OK) return rc;
    }
  return MLO_OK;
}

void launch_persistent(mlo_ctx* c, bool tpq, bool multi, uint32_t nblk, const MapDev& map, const MapDev* d_maps,
                       const IcpProblem* dP, IcpState* dS, const float4* d_local, const IcpQueue& q, uint32_t qpw, bool planes = true) {
#define MLO_PERS(T, M, PL, QPW)                                                                                          \
  do {                                                                                                                   \
    if (c->pers_minb_now == 2)                                                                                           \
      LAUNCH(c, (k_icp_persistent<T, M, 2, PL>), nblk, ICP_BLOCK, map, d_maps, dP, dS, d_local, c->d_pairA.as<float4>(), \
             c->d_pairB.as<float4>(), c->d_partials.as<double>(), c->d_partcnt.as<uint32_t>(), q, QPW,                   \
             (c->fuse_inner ? 1 : 0) | (c->prior_ahead ? 2 : 0));                                                                                     \
    else                                                                                                                 \
      LAUNCH(c, (k_icp_persistent<T, M, 4, PL>), nblk, ICP_BLOCK, map, d_maps, dP, dS, d_local, c->d_pairA.as<float4>(), \
             c->d_pairB.as<float4>(), c->d_partials.as<double>(), c->d_partcnt.as<uint32_t>(), q, QPW,                   \
             (c->fuse_inner ? 1 : 0) | (c->prior_ahead ? 2 : 0));                                                                                     \
  } while (0)
  // (the warp-per-query form exists with and without the point-to-plane matcher's code: point-to-point pipelines run
  // the lean instantiation)
  if (tpq && multi) MLO_PERS(true, true, true, 0u);
  else if (tpq) MLO_PERS(true, false, true, 0u);
  else if (multi && planes) MLO_PERS(false, true, true, qpw);
  else if (multi) MLO_PERS(false, true, false, qpw);
  else if (planes) MLO_PERS(false, false, true, qpw);
  else MLO_PERS(false, false, false, qpw);
#undef MLO_PERS
}

template <int NT, bool PLANES>
int launch_block_nt(mlo_ctx* c, int slot, uint32_t B, uint32_t cl, const MapDev* d_maps, const IcpProblem* dP, IcpState* dS,
                    const float4* d_local) {
  const size_t smem = sizeof(BlockShared<NT>);
  if (!c->block_attr_set[slot]) {
    CU(c, cudaFuncSetAttribute(k_icp_block<NT, PLANES>, cudaFuncAttributeMaxDynamicSharedMemorySize, int(smem)));
    c->block_attr_set[slot] = true;
  }
  cudaLaunchConfig_t cfg = {};
  cfg.gridDim = dim3(B * cl);
  cfg.blockDim = dim3(NT);
  cfg.dynamicSmemBytes = smem;
  cfg.stream = c->stream;
  cudaLaunchAttribute at[1];
  at[0].id = cudaLaunchAttributeClusterDimension;
  at[0].val.clusterDim.x = cl;
  at[0].val.clusterDim.y = 1;
  at[0].val.clusterDim.z = 1;
  cfg.attrs = at;
  cfg.numAttrs = 1;
  float4* pa = c->d_pairA.as<float4>();
  float4* pb = c->d_pairB.as<float4>();
  CU(c, cudaLaunchKernelEx(&cfg, k_icp_block<NT, PLANES>, d_maps, dP, dS, d_local, pa, pb));
  c->launches++;
  return MLO_OK;
}
// One thread-block cluster per problem (icp_block.cuh).  Blocks per cluster: as many as fit one block per SM for the
// whole batch (a single sequence gets 8 SMs, a fleet of 32 gets 4 each), but at least 64 queries per block; 512 threads
// per block (one block per SM at 128 registers) while the batch fits the machine, 256 (two per SM) beyond that.
int launch_block(mlo_ctx* c, uint32_t B, uint32_t max_nq, bool planes, const MapDev* d_maps, const IcpProblem* dP, IcpState* dS,
                 const float4* d_local) {
  uint32_t cl = uint32_t(c->block_cluster);
  if (cl != 1 && cl != 2 && cl != 4 && cl != 8) {
    cl = 8;
    while (cl > 1 && B * cl > uint32_t(c->sm_count)) cl >>= 1;
    while (cl > 1 && max_nq / cl < 32) cl >>= 1;
  }
  int nt = c->block_threads;
  if (nt != 256 && nt != 512) nt = (B * cl <= uint32_t(c->sm_count) && max_nq > 32 * cl) ? 512 : 256;  // (8 lanes per query)
  c->last_block_cluster = int(cl);
  c->last_block_threads = nt;
  if (planes) {  // any problem of the batch runs Matcher_Point2Plane
    if (nt == 512) return launch_block_nt<512, true>(c, 0, B, cl, d_maps, dP, dS, d_local);
    return launch_block_nt<256, true>(c, 1, B, cl, d_maps, dP, dS, d_local);
  }
  if (nt == 512) return launch_block_nt<512, false>(c, 2, B, cl, d_maps, dP, dS, d_local);
  return launch_block_nt<256, false>(c, 3, B, cl, d_maps, dP, dS, d_local);
}

// The batched align driver over device-resident float4 local points.  Problem b reads local points
// [q_begin[b], q_begin[b] + n_q[b]) of d_local and matches against maps[b] (all equal: the read-only-map batch of
// config[1]; different: a fleet of independent sequences advanced in lock step, one local map each).
int align_batch_core(mlo_ctx* c, uint32_t B, const float4* d_local, const uint64_t* q_begin, const uint32_t* n_q,
                     const mlo_map* const* maps, const double* init_poses, const mlo_icp_params* params, mlo_icp_result* out) {
  if (B == 0) return MLO_OK;
  // ---- map table
  std::vector<const mlo_map*> uniq;
  std::vector<uint32_t> map_idx(B, 0);
  for (uint32_t b = 0; b < B; b++) {
    if (!maps[b]) return fail(c, MLO_ERR_INVALID_ARG, "null map");
    size_t k = 0;
    while (k < uniq.size() && uniq[k] != maps[b]) k++;
    if (k == uniq.size()) uniq.push_back(maps[b]);
    map_idx[b] = uint32_t(k);
  }
  const bool multi = uniq.size() > 1;
  const mlo_map* map = uniq[0];
  const MapDev* d_maps = nullptr;
  {  // (always: the block kernel stages its map descriptor from this table, single map or not)
    std::vector<MapDev> hm(uniq.size());
    for (size_t k = 0; k < uniq.size(); k++) hm[k] = uniq[k]->dev;
    CU(c, c->d_maps.ensure(hm.size() * sizeof(MapDev)));
    CU(c, cudaMemcpyAsync(c->d_maps.p, hm.data(), hm.size() * sizeof(MapDev), cudaMemcpyHostToDevice, c->stream));
    d_maps = c->d_maps.as<MapDev>();
  }
  // ---- problems + tables
  std::vector<IcpProblem> probs(B);
  std::vector<double> tables;
  std::map<const double*, size_t> seen;
  auto put = [&](const double* t, uint32_t len) -> size_t {
    if (!t || len == 0) return size_t(-1);
    auto it = seen.find(t);
    if (it != seen.end()) return it->second;
    const size_t off = tables.size();
    tables.insert(tables.end(), t, t + len);
    seen[t] = off;
    return off;
  };
  std::vector<size_t> toff(3 * size_t(B));
  uint32_t part_total = 0, max_blocks = 0, max_blocks_acc = 0, max_it = 0, max_inner = 1, max_nq = 0;
  bool any_planes = false;
  // queries per warp: 32 when the batch alone fills the machine, fewer for latency-bound small batches
  uint64_t total_queries = 0, q_end = 0;
  for (uint32_t b = 0; b < B; b++) {
    total_queries += n_q[b];
    q_end = std::max(q_end, q_begin[b] + n_q[b]);
  }
  uint32_t qpw = 32;
  const uint32_t qpw_floor = total_queries >= 256 ? uint32_t(c->qpw_floor) : 1u;  // (MLO_QPW_FLOOR)
  while (qpw > qpw_floor && total_queries / qpw < uint64_t(c->sm_count) * 32) qpw >>= 1;
  // large batches: thread-per-query kernel (hundreds of queries in flight per SM); small: warp-per-query
  const bool use_tpq = c->force_kernel == 1 || c->force_kernel == 3 ||
                       (c->force_kernel == 0 && total_queries >= uint64_t(c->sm_count) * c->tpq_min_queries_per_sm);
  for (uint32_t b = 0; b < B; b++) {
    const mlo_icp_params& p = params[b];
    IcpProblem& P = probs[b];
    std::memset(&P, 0, sizeof(P));
    if ((p.matcher_mask & MLO_MATCHER_PT2PL) && maps[b]->prm.kind != MLO_MAP_NDT)
      return fail(c, MLO_ERR_UNSUPPORTED, "Matcher_Point2Plane needs an NDT map");
    if (p.solver == MLO_SOLVER_HORN && (p.matcher_mask & MLO_MATCHER_PT2PL))
      return fail(c, MLO_ERR_UNSUPPORTED, "Solver_Horn handles point-to-point pairings only");
    if (p.table_len == 0 || !p.pt2pt_threshold_by_iter || !p.kernel_param_by_iter)
      if (p.matcher_mask & MLO_MATCHER_PT2PT) return fail(c, MLO_ERR_INVALID_ARG, "missing per-iteration tables");
    P.q_begin = q_begin[b];
    P.n_q = n_q[b];
    P.map_idx = map_idx[b];
    P.log = nullptr;
    P.log_cap = c->log_cap;
    P.max_iterations = p.max_iterations;
    P.min_abs_step_trans = p.min_abs_step_trans;
    P.min_abs_step_rot = p.min_abs_step_rot;
    P.solver = p.solver;
    P.gn_max_iterations = std::max<uint32_t>(1, p.gn_max_iterations);
    P.gn_min_delta = p.gn_min_delta;
    P.robust_kernel = p.robust_kernel | (c->conv_gm_form ? KERNEL_GM_FORM_BIT : 0);
    P.matcher_mask = p.matcher_mask;
    any_planes = any_planes || (p.matcher_mask & MLO_MATCHER_PT2PL);
    P.table_len = p.table_len;
    toff[3 * b] = put(p.pt2pt_threshold_by_iter, p.table_len);
    toff[3 * b + 1] = put(p.pt2pl_threshold_by_iter, p.table_len);
    toff[3 * b + 2] = put(p.kernel_param_by_iter, p.table_len);
    const double ang = p.threshold_angular_deg * M_PI / 180.0;
    P.ang2 = float(ang * ang);
    P.w_pt2pt = p.pt2pt_weight;
    P.w_pt2pl = p.pt2pl_weight;
    P.has_prior = p.has_prior;
    std::memcpy(P.prior_pose, p.prior_pose_3x4, sizeof(P.prior_pose));
    std::memcpy(P.prior_info, p.prior_info_6x6, sizeof(P.prior_info));
    P.hook_enabled = p.hook_enabled;
    P.hook_min_trans = p.hook_min_trans;
    P.hook_min_rot = p.hook_min_rot_rad;
    std::memcpy(P.hook_checkpoint, p.hook_checkpoint_pose_3x4, sizeof(P.hook_checkpoint));
    const bool use_wl = use_tpq && (multi || c->force_kernel != 1);
    const uint32_t qpb_pers = use_tpq ? ICP_BLOCK : (ICP_BLOCK / 32) * qpw;  // persistent kernel: tpq or warp chunks
    // launch sequence: work-list chunks of 128 queries (one partial per block) or of 32 (one partial per warp: one-warp
    // blocks, or four-warp blocks without the closing block barrier - wl_variant 10 / 11)
    const bool warp_partials = !multi && (c->wl_warps != 4 || c->wl_variant == 10 || c->wl_variant == 11);
    const uint32_t qpb = use_wl ? (warp_partials ? WL_BLOCK : ICP_BLOCK) : qpb_pers;
    P.n_blocks = (P.n_q + qpb - 1) / qpb;
    P.n_blocks_pers = (P.n_q + qpb_pers - 1) / qpb_pers;
    P.n_blocks_acc = (P.n_q + ICP_BLOCK - 1) / ICP_BLOCK;
    P.part_begin = part_total;
    part_total += std::max(std::max(P.n_blocks, P.n_blocks_pers), P.n_blocks_acc);
    max_blocks = std::max(max_blocks, P.n_blocks);
    max_nq = std::max(max_nq, P.n_q);
    max_blocks_acc = std::max(max_blocks_acc, P.n_blocks_acc);
    max_it = std::max(max_it, P.max_iterations);
    if (P.solver == MLO_SOLVER_GAUSS_NEWTON) max_inner = std::max(max_inner, P.gn_max_iterations);
  }
  CU(c, c->d_tables.ensure(std::max<size_t>(tables.size(), 1) * sizeof(double)));
  if (!tables.empty())
    CU(c, cudaMemcpyAsync(c->d_tables.p, tables.data(), tables.size() * sizeof(double), cudaMemcpyHostToDevice, c->stream));
  for (uint32_t b = 0; b < B; b++) {
    const double* base = c->d_tables.as<double>();
    probs[b].thr_pt2pt = toff[3 * b] == size_t(-1) ? nullptr : base + toff[3 * b];
    probs[b].thr_pt2pl = toff[3 * b + 1] == size_t(-1) ? nullptr : base + toff[3 * b + 1];
    probs[b].kparam = toff[3 * b + 2] == size_t(-1) ? nullptr : base + toff[3 * b + 2];
  }
  const uint64_t total_q = q_end;
  if (c->log_cap) {
    CU(c, c->d_log.ensure(size_t(B) * c->log_cap * sizeof(mlo_icp_iteration_record)));
    CU(c, cudaMemsetAsync(c->d_log.p, 0xFF, size_t(B) * c->log_cap * sizeof(mlo_icp_iteration_record), c->stream));
    for (uint32_t b = 0; b < B; b++) probs[b].log = c->d_log.as<mlo_icp_iteration_record>() + size_t(b) * c->log_cap;
  }
  CU(c, c->d_probs.ensure(B * sizeof(IcpProblem)));
  CU(c, c->d_states.ensure(B * sizeof(IcpState)));
  CU(c, c->d_init.ensure(B * 12 * sizeof(double)));
  CU(c, c->d_pairA.ensure(std::max<size_t>(total_q, 1) * sizeof(float4)));
  CU(c, c->d_pairB.ensure(std::max<size_t>(total_q, 1) * sizeof(float4)));
  CU(c, c->d_partials.ensure(std::max<size_t>(part_total, 1) * NACC * sizeof(double)));
  CU(c, c->d_partcnt.ensure(std::max<size_t>(part_total, 1) * 2 * sizeof(uint32_t)));
  CU(c, c->d_misc.ensure(256));
  CU(c, c->h_misc.ensure(256));
  CU(c, c->h_states.ensure(B * sizeof(IcpState)));
  CU(c, cudaMemcpyAsync(c->d_probs.p, probs.data(), B * sizeof(IcpProblem), cudaMemcpyHostToDevice, c->stream));
  CU(c, cudaMemcpyAsync(c->d_init.p, init_poses, B * 12 * sizeof(double), cudaMemcpyHostToDevice, c->stream));
  uint32_t* d_active = c->d_misc.as<uint32_t>();
  uint32_t* h_active = c->h_misc.as<uint32_t>() + 8;
  CU(c, cudaMemsetAsync(d_active, 0, sizeof(uint32_t), c->stream));
  const IcpProblem* dP = c->d_probs.as<IcpProblem>();
  IcpState* dS = c->d_states.as<IcpState>();
  LAUNCH(c, k_init_states, (B + 127) / 128, 128, dP, dS, c->d_init.as<double>(), B, d_active);

  // The deferred transfer of the next staged batch is enqueued once ALL host-to-device uploads of this call are in
  // the queue (the work-queue upload of the persistent path included): the copy engine serves transfers in submission
  // order, and a 64 MB batch in front of a 4 KB queue upload would hold the ICP kernel back by milliseconds.
  bool deferred_issued = false;
  auto issue_deferred_once = [&]() -> int {
    if (deferred_issued) return MLO_OK;
    deferred_issued = true;
    return issue_deferred_uploads(c, c->consuming_slot);
  };
  const size_t e_icp = prof_begin(c);
  const dim3 grid(std::max(max_blocks, 1u), B);
  const dim3 grid_acc(std::max(max_blocks_acc, 1u), B);
  const uint32_t check_every = uint32_t(std::max(1, c->check_every));
  // the queue-driven kernel wins while a launch sequence would be latency/launch-bound (small batches);
  // for large batches one kernel per phase streams better (profiles/README.md)
  const uint64_t large_at = c->large_batch_queries ? c->large_batch_queries : uint64_t(c->sm_count) * 1024;
  const bool queue_ok = B < 65536 && max_blocks < 32768 && max_blocks_acc < 32768;  // (chunk counts of the queue geometry)
  int path = c->align_path;
  if (path == 0) {
    if (!c->use_persistent) path = 1;
    else if (c->persistent_forced) path = 2;
    else path = total_queries < large_at ? 2 : 1;  // (the block kernel, path 3, measured slower than the queue kernel: profiles/README.md)
  }
  if (path == 2 && !queue_ok) path = 1;
  const bool persistent = path == 2;
  c->last_align_path = path;
  c->last_stream_groups = 1;
  c->last_tail_handover = 0;
  if (path == 3) {
    // ---- one thread block per problem runs the whole align loop (icp_block.cuh)
    const size_t e_nn = prof_begin(c);
    int rc = launch_block(c, B, max_nq, any_planes, d_maps, dP, dS, d_local);
    prof_end(c, 3, e_nn);
    if (rc != MLO_OK) return rc;
    rc = issue_deferred_once();
    if (rc != MLO_OK) return rc;
  }
  if (persistent) {
    // ---- one launch for the whole align loop: queue of (problem, phase, chunk) items
    std::vector<uint32_t> items;
    uint32_t n_act = 0;
    for (uint32_t b = 0; b < B; b++) {
      if (probs[b].max_iterations == 0 || probs[b].n_q == 0) continue;
      n_act++;
      for (uint32_t ch = 0; ch < probs[b].n_blocks_pers; ch++) items.push_back(item_make(b, ch, 0u));
    }
    if (n_act) {
      const uint32_t qcap = uint32_t(next_pow2(std::max<uint64_t>(2ull * part_total, 1024)));
      const uint32_t n_items = uint32_t(std::min<size_t>(items.size(), qcap));
      c->h_queue.resize(qcap);  // (a member: the asynchronous upload below must not outlive a local)
      for (uint32_t i = 0; i < qcap; i++) c->h_queue[i] = i < n_items ? slot_word(i + 1, items[i]) : slot_word(i, 0u);
      CU(c, c->d_queue.ensure((2ull * qcap + 8 + B) * sizeof(uint32_t)));
      uint32_t* dq = c->d_queue.as<uint32_t>();
      IcpQueue q;
      q.slots = reinterpret_cast<unsigned long long*>(dq);
      q.ctrl = dq + 2ull * qcap;
      q.phase_cnt = dq + 2ull * qcap + 8;
      q.mask = qcap - 1;
      const uint32_t h_ctrl[8] = {0u, n_items, n_act, 0u, 0u, 0u, 0u, 0u};
      CU(c, cudaMemcpyAsync(q.slots, c->h_queue.data(), qcap * sizeof(unsigned long long), cudaMemcpyHostToDevice, c->stream));
      CU(c, cudaMemcpyAsync(q.ctrl, h_ctrl, sizeof(h_ctrl), cudaMemcpyHostToDevice, c->stream));
      CU(c, cudaMemsetAsync(q.phase_cnt, 0, B * sizeof(uint32_t), c->stream));
      if (c->persistent_blocks == 0) {
        int per_sm = 0;
        CU(c, cudaOccupancyMaxActiveBlocksPerMultiprocessor(&per_sm, k_icp_persistent<true, true>, ICP_BLOCK, 0));
        c->persistent_blocks = std::max(1, per_sm) * c->sm_count;
      }
      c->pers_minb_now = c->pers_minb ? c->pers_minb : (B == 1 ? 2 : 4);
      const uint32_t resident = c->pers_minb_now == 2 ? uint32_t(2 * c->sm_count) : uint32_t(c->persistent_blocks);
      const uint32_t nblk = std::min<uint32_t>(resident, std::max<uint32_t>(part_total, 1u));
      const size_t e_nn = prof_begin(c);
      launch_persistent(c, use_tpq, multi, nblk, map->dev, d_maps, dP, dS, d_local, q, qpw, any_planes);
      prof_end(c, 3, e_nn);
      {
        int rc = issue_deferred_once();
        if (rc != MLO_OK) return rc;
      }
      CU(c, cudaMemcpyAsync(h_active, q.ctrl + 4, sizeof(uint32_t), cudaMemcpyDeviceToHost, c->stream));
      CU(c, cudaStreamSynchronize(c->stream));
      if (*h_active != 0) return fail(c, MLO_ERR_CUDA, "persistent ICP kernel timed out waiting on its work queue");
    }
  }
  {  // launch-sequence path (or nothing to run): no further uploads follow
    int rc = issue_deferred_once();
    if (rc != MLO_OK) return rc;
  }
  // stream groups: contiguous slices of the batch, each with its own launch sequence (group 0 on the context stream)
  // (per-kernel timing needs the kernel alone on the device: one group while profiling)
  const uint32_t n_groups = (path == 1 && use_tpq && B >= 64 && !c->prof_on) ? uint32_t(c->stream_groups) : 1u;
  if (path == 1) c->last_stream_groups = int(n_groups);
  auto group_stream = [&](uint32_t g) { return g == 0 ? c->stream : c->aux_stream[g - 1]; };
  if (n_groups > 1) {
    CU(c, cudaEventRecord(c->ev_fork, c->stream));
    for (uint32_t g = 1; g < n_groups; g++) CU(c, cudaStreamWaitEvent(c->aux_stream[g - 1], c->ev_fork, 0));
  }
  for (uint32_t it = 0; path == 1 && it < max_it; it++) {
    for (uint32_t g = 0; g < n_groups; g++) {
      const uint32_t g0 = uint32_t(uint64_t(B) * g / n_groups), g1 = uint32_t(uint64_t(B) * (g + 1) / n_groups);
      const uint32_t Bg = g1 - g0;
      if (Bg == 0) continue;
      cudaStream_t sg = group_stream(g);
      const IcpProblem* gP = dP + g0;
      IcpState* gS = dS + g0;
      const dim3 grid_g(grid.x, Bg), grid_acc_g(grid_acc.x, Bg);
      const size_t e_nn = g == 0 ? prof_begin(c) : 0;
      if (multi && use_tpq) {
        LAUNCH_ON(c, sg, k_match_accumulate_wl4<true>, grid_g, ICP_BLOCK, map->dev, d_maps, gP, gS, d_local,
                  c->d_pairA.as<float4>(), c->d_pairB.as<float4>(), c->d_partials.as<double>(), c->d_partcnt.as<uint32_t>());
      } else if (multi) {
        LAUNCH_ON(c, sg, k_match_accumulate<true>, grid_g, ICP_BLOCK, map->dev, d_maps, gP, gS, d_local,
                  c->d_pairA.as<float4>(), c->d_pairB.as<float4>(), c->d_partials.as<double>(), c->d_partcnt.as<uint32_t>(), qpw);
      } else if (use_tpq && c->force_kernel != 1) {
#define MLO_WL_LAUNCH(MB)                                                                                              \
  LAUNCH_ON(c, sg, k_match_accumulate_wl<MB>, grid_g, WL_BLOCK, map->dev, gP, gS, d_local, c->d_pairA.as<float4>(),   \
            c->d_pairB.as<float4>(), c->d_partials.as<double>(), c->d_partcnt.as<uint32_t>())
#define MLO_WL4_BULK_LAUNCH(MB)                                                                                             \
  LAUNCH_ON(c, sg, (k_match_accumulate_wl4<false, false, MB, true>), grid_g, ICP_BLOCK, map->dev, d_maps, gP, gS, d_local,    \
            c->d_pairA.as<float4>(), c->d_pairB.as<float4>(), c->d_partials.as<double>(), c->d_partcnt.as<uint32_t>())
#define MLO_WL4_LAUNCH(PIPE, MB)                                                                                            \
  LAUNCH_ON(c, sg, (k_match_accumulate_wl4<false, PIPE, MB>), grid_g, ICP_BLOCK, map->dev, d_maps, gP, gS, d_local,           \
            c->d_pairA.as<float4>(), c->d_pairB.as<float4>(), c->d_partials.as<double>(), c->d_partcnt.as<uint32_t>())
#define MLO_WL4_OCT_LAUNCH(DEPTH, MB)                                                                                       \
  LAUNCH_ON(c, sg, (k_match_accumulate_wl4<false, false, MB, false, DEPTH>), grid_g, ICP_BLOCK, map->dev, d_maps, gP, gS,   \
            d_local, c->d_pairA.as<float4>(), c->d_pairB.as<float4>(), c->d_partials.as<double>(), c->d_partcnt.as<uint32_t>())
#define MLO_WL4_A32_LAUNCH(MB, MODE)                                                                                        \
  LAUNCH_ON(c, sg, (k_match_accumulate_wl4<false, false, MB, false, 0, false, MODE>), grid_g, ICP_BLOCK, map->dev, d_maps, gP, \
            gS, d_local, c->d_pairA.as<float4>(), c->d_pairB.as<float4>(), c->d_partials.as<double>(), c->d_partcnt.as<uint32_t>())
#define MLO_WL4_WPART_LAUNCH(MB)                                                                                            \
  LAUNCH_ON(c, sg, (k_match_accumulate_wl4<false, false, MB, false, 0, true>), dim3((grid_g.x + 3) / 4, grid_g.y), ICP_BLOCK,  \
            map->dev, d_maps, gP, gS, d_local, c->d_pairA.as<float4>(), c->d_pairB.as<float4>(), c->d_partials.as<double>(), \
            c->d_partcnt.as<uint32_t>())
        if (c->wl_warps == 4) {
          switch (c->wl_variant) {  // MLO_WL_VARIANT: A/B of the drain loop (profiles/README.md)
            case 6: MLO_WL4_OCT_LAUNCH(4, 6); break;  // contiguous-range drain, register-resident running best
            case 7: MLO_WL4_OCT_LAUNCH(4, 8); break;
            case 8: MLO_WL4_OCT_LAUNCH(8, 6); break;
            case 9: MLO_WL4_OCT_LAUNCH(8, 5); break;
            case 12: MLO_WL4_A32_LAUNCH(6, 1); break;  // split 32-bit keys: native shared-memory atomics
            case 13: MLO_WL4_A32_LAUNCH(8, 1); break;
            case 14: MLO_WL4_A32_LAUNCH(6, 2); break;  // 32-bit segment minimum + ballot, 64-bit best
            case 15: MLO_WL4_A32_LAUNCH(8, 2); break;
            case 16: MLO_WL4_A32_LAUNCH(5, 3); break;  // eight segments per lane group and round
            case 17: MLO_WL4_A32_LAUNCH(6, 3); break;
            case 18: MLO_WL4_A32_LAUNCH(6, 4); break;  // segment minima merged along runs of equal queries before the atomic
            case 19: MLO_WL4_A32_LAUNCH(8, 4); break;
            case 10: MLO_WL4_WPART_LAUNCH(6); break;  // one partial per warp, no block barrier
            case 11: MLO_WL4_WPART_LAUNCH(8); break;
            case 1: MLO_WL4_LAUNCH(true, 8); break;
            case 2: MLO_WL4_LAUNCH(true, 6); break;
            case 3: MLO_WL4_LAUNCH(false, 6); break;
            case 4: MLO_WL4_BULK_LAUNCH(5); break;  // cp.async.bulk staging of the row segments (A/B)
            case 5: MLO_WL4_BULK_LAUNCH(4); break;
            default: MLO_WL4_LAUNCH(false, 8); break;
          }
        }
        else switch (c->wl_min_blocks) {
          case 16: MLO_WL_LAUNCH(16); break;
          case 24: MLO_WL_LAUNCH(24); break;
          default: MLO_WL_LAUNCH(32); break;
        }
      } else if (use_tpq)
        LAUNCH_ON(c, sg, k_match_accumulate_tpq, grid_g, ICP_BLOCK, map->dev, gP, gS, d_local, c->d_pairA.as<float4>(),
                  c->d_pairB.as<float4>(), c->d_partials.as<double>(), c->d_partcnt.as<uint32_t>());
      else
        LAUNCH_ON(c, sg, k_match_accumulate<false>, grid_g, ICP_BLOCK, map->dev, d_maps, gP, gS, d_local,
                  c->d_pairA.as<float4>(), c->d_pairB.as<float4>(), c->d_partials.as<double>(), c->d_partcnt.as<uint32_t>(), qpw);
      if (g == 0) prof_end(c, 3, e_nn);
      // problems up to FUSE_MAX_Q queries run their inner Gauss-Newton iterations inside the solve block
      const int fuse = ((c->fuse_inner && max_nq <= FUSE_MAX_Q) ? 1 : 0) | (c->prior_ahead ? 2 : 0);
      LAUNCH_ON(c, sg, k_solve, Bg, ICP_BLOCK, gP, gS, c->d_partials.as<double>(), c->d_partcnt.as<uint32_t>(), 1, d_active, fuse,
                d_local, c->d_pairA.as<float4>(), c->d_pairB.as<float4>());
      for (uint32_t inner = 1; inner < max_inner && !(fuse & 1); inner++) {
        LAUNCH_ON(c, sg, k_accumulate, grid_acc_g, ICP_BLOCK, gP, gS, d_local, c->d_pairA.as<float4>(), c->d_pairB.as<float4>(),
                  c->d_partials.as<double>(), c->d_partcnt.as<uint32_t>());
        LAUNCH_ON(c, sg, k_solve, Bg, ICP_BLOCK, gP, gS, c->d_partials.as<double>(), c->d_partcnt.as<uint32_t>(), 0, d_active, fuse & 2,
                  d_local, c->d_pairA.as<float4>(), c->d_pairB.as<float4>());
      }
    }
    if (((it % check_every) == check_every - 1 || it + 1 == max_it) && n_groups > 1)
      for (uint32_t g = 1; g < n_groups; g++) {  // join: the context stream waits for every group
        CU(c, cudaEventRecord(c->ev_join[g - 1], c->aux_stream[g - 1]));
        CU(c, cudaStreamWaitEvent(c->stream, c->ev_join[g - 1], 0));
      }
    if ((it % check_every) == check_every - 1 || it + 1 == max_it) {
      CU(c, cudaMemcpyAsync(h_active, d_active, sizeof(uint32_t), cudaMemcpyDeviceToHost, c->stream));
      CU(c, cudaStreamSynchronize(c->stream));
      if (*h_active == 0) break;
      // Tail hand-over: once few problems remain, a launch sequence is latency-bound (each iteration still costs
      // four launches); the queue-driven kernel finishes the stragglers in one launch.
      const uint64_t avg_q = total_queries / std::max<uint32_t>(B, 1);
      if (c->tail_handover && c->tail_path == 3 && it + 1 < max_it && uint64_t(*h_active) * avg_q < uint64_t(c->sm_count) * uint64_t(std::max(1, c->tail_queries_per_sm))) {
        const size_t e_nn = prof_begin(c);
        int rc = launch_block(c, B, max_nq, any_planes, d_maps, dP, dS, d_local);  // (finished problems leave at once)
        c->last_tail_handover = 3;
        prof_end(c, 3, e_nn);
        if (rc != MLO_OK) return rc;
        break;
      }
      if (c->use_persistent && c->tail_handover && it + 1 < max_it && uint64_t(*h_active) * avg_q < uint64_t(c->sm_count) * uint64_t(std::max(1, c->tail_queries_per_sm)) &&
          queue_ok) {
        const uint32_t qcap = uint32_t(next_pow2(std::max<uint64_t>(2ull * part_total, 1024)));
        CU(c, c->d_queue.ensure((2ull * qcap + 8 + B) * sizeof(uint32_t)));
        uint32_t* dq = c->d_queue.as<uint32_t>();
        IcpQueue q;
        q.slots = reinterpret_cast<unsigned long long*>(dq);
        q.ctrl = dq + 2ull * qcap;
        q.phase_cnt = dq + 2ull * qcap + 8;
        q.mask = qcap - 1;
        c->last_tail_handover = 2;
        LAUNCH(c, k_queue_reset, (std::max(qcap, B) + 255) / 256, 256, q, B);
        LAUNCH(c, k_queue_build, (B + 127) / 128, 128, dP, dS, q, B);
        if (c->persistent_blocks == 0) {
          int per_sm = 0;
          CU(c, cudaOccupancyMaxActiveBlocksPerMultiprocessor(&per_sm, k_icp_persistent<true, true>, ICP_BLOCK, 0));
          c->persistent_blocks = std::max(1, per_sm) * c->sm_count;
        }
        c->pers_minb_now = c->pers_minb == 2 ? 2 : 4;  // (a tail of a large batch: several problems)
        const uint32_t resident = c->pers_minb_now == 2 ? uint32_t(2 * c->sm_count) : uint32_t(c->persistent_blocks);
        const uint32_t nblk = std::min<uint32_t>(resident, std::max<uint32_t>(part_total, 1u));
        const size_t e_nn = prof_begin(c);
        launch_persistent(c, use_tpq, multi, nblk, map->dev, d_maps, dP, dS, d_local, q, qpw, any_planes);
        prof_end(c, 3, e_nn);
        CU(c, cudaMemcpyAsync(h_active, q.ctrl + 4, sizeof(uint32_t), cudaMemcpyDeviceToHost, c->stream));
        CU(c, cudaStreamSynchronize(c->stream));
        if (*h_active != 0) return fail(c, MLO_ERR_CUDA, "persistent ICP kernel timed out waiting on its work queue");
        break;
      }
    }
  }
  prof_end(c, 1, e_icp);
  CU(c, cudaMemcpyAsync(c->h_states.p, dS, B * sizeof(IcpState), cudaMemcpyDeviceToHost, c->stream));
  c->h_log_problems = 0;
  if (c->log_cap) {
    c->h_log.resize(size_t(B) * c->log_cap);
    CU(c, cudaMemcpyAsync(c->h_log.data(), c->d_log.p, c->h_log.size() * sizeof(mlo_icp_iteration_record), cudaMemcpyDeviceToHost,
                          c->stream));
    c->h_log_problems = B;
  }
  CU(c, cudaStreamSynchronize(c->stream));
  CU(c, cudaGetLastError());
  const IcpState* hs = c->h_states.as<IcpState>();
  for (uint32_t b = 0; b < B; b++) {
    const IcpState& S = hs[b];
    mlo_icp_result& r = out[b];
    std::memset(&r, 0, sizeof(r));
    std::memcpy(r.pose_3x4, S.T, sizeof(r.pose_3x4));
    if (S.have_H) spd6_inverse(S.H, r.cov_6x6);
    r.n_iterations = S.it;
    r.termination = S.term;
    r.n_pairings = S.n_pairs;
    r.n_potential_pairings = S.n_potential;
    r.quality = S.n_potential ? double(S.n_pairs) / double(S.n_potential) : 0.0;
    r.n_query_iterations = S.n_query_it;
    r.n_candidate_points = S.n_cand;
    if (c->prof_on) {
      c->prof.nn_query_iterations += S.n_query_it;
      c->prof.nn_candidate_points += S.n_cand;
      c->prof.nn_blocks += uint64_t(probs[b].n_blocks) * (S.n_query_it / std::max<uint32_t>(1, probs[b].n_q));
    }
  }
  prof_collect(c);
  return MLO_OK;
}

int align_batch_device(mlo_ctx* c, uint32_t B, const float4* d_local, const uint64_t* offsets, const mlo_map* map,
                       const double* init_poses, const mlo_icp_params* params, mlo_icp_result* out) {
  std::vector<uint64_t> qb(offsets, offsets + B);
  std::vector<uint32_t> nq(B);
  for (uint32_t b = 0; b < B; b++) nq[b] = uint32_t(offsets[b + 1] - offsets[b]);
  std::vector<const mlo_map*> maps(B, map);
  return align_batch_core(c, B, d_local, qb.data(), nq.data(), maps.data(), init_poses, params, out);
}

}  // namespace

// =================================================================== C ABI
extern "C" {

int mlo_abi_version(void) { return MLO_ABI_VERSION; }

int mlo_create(int cuda_device, mlo_ctx** out) {
  if (!out) return MLO_ERR_INVALID_ARG;
  *out = nullptr;
  int n = 0;
  if (cudaGetDeviceCount(&n) != cudaSuccess || n == 0 || cuda_device < 0 || cuda_device >= n) return MLO_ERR_NO_DEVICE;
  cudaDeviceProp prop;
  if (cudaGetDeviceProperties(&prop, cuda_device) != cudaSuccess) return MLO_ERR_CUDA;
  if (prop.major != 10) return MLO_ERR_NO_DEVICE;  // sm_100a cubin only: no other architecture, no CPU path
  auto* c = new mlo_ctx;
  c->device = cuda_device;
  c->sm_count = prop.multiProcessorCount;
  c->cc_major = prop.major;
  c->cc_minor = prop.minor;
  c->dev_name = prop.name;
  if (const char* fk = getenv("MLO_FORCE_KERNEL")) c->force_kernel = atoi(fk);
  if (const char* th = getenv("MLO_TAIL_HANDOVER")) c->tail_handover = atoi(th) != 0;
  if (const char* tq = getenv("MLO_TAIL_QUERIES_PER_SM")) c->tail_queries_per_sm = std::max(1, atoi(tq));
  if (const char* ce = getenv("MLO_CHECK_EVERY")) c->check_every = std::max(1, atoi(ce));
  if (const char* fi = getenv("MLO_FUSE_INNER")) c->fuse_inner = atoi(fi) != 0;
  if (const char* pa = getenv("MLO_PRIOR_AHEAD")) c->prior_ahead = atoi(pa) != 0;
  if (const char* pm = getenv("MLO_PERS_MINB")) c->pers_minb = atoi(pm) == 2 ? 2 : (atoi(pm) == 4 ? 4 : 0);
  if (const char* qf = getenv("MLO_QPW_FLOOR")) c->qpw_floor = std::min(32, std::max(1, atoi(qf)));
  if (const char* wb = getenv("MLO_WL_MIN_BLOCKS")) c->wl_min_blocks = atoi(wb);
  if (const char* tq = getenv("MLO_TPQ_MIN")) c->tpq_min_queries_per_sm = std::max(1, atoi(tq));
  if (const char* ww = getenv("MLO_WL_WARPS")) c->wl_warps = atoi(ww);
  if (const char* wv = getenv("MLO_WL_VARIANT")) c->wl_variant = atoi(wv);
  if (const char* tf = getenv("MLO_TABLE_FACTOR")) c->table_factor = std::max(1, atoi(tf));
  if (const char* sg = getenv("MLO_STREAM_GROUPS")) c->stream_groups = std::min(int(mlo_ctx::MAX_GROUPS), std::max(1, atoi(sg)));
  if (const char* ap = getenv("MLO_ALIGN_PATH")) c->align_path = std::min(3, std::max(0, atoi(ap)));
  if (const char* lb = getenv("MLO_LARGE_BATCH_QUERIES")) c->large_batch_queries = uint64_t(std::max(0ll, atoll(lb)));
  if (const char* tp = getenv("MLO_TAIL_PATH")) c->tail_path = atoi(tp) == 2 ? 2 : 3;
  if (const char* bt = getenv("MLO_BLOCK_THREADS")) c->block_threads = atoi(bt);
  if (const char* bc = getenv("MLO_BLOCK_CLUSTER")) c->block_cluster = atoi(bc);
  if (const char* fg = getenv("MLO_FILTER_GROUP_MB")) c->filter_group_mb = std::max(1, atoi(fg));
  if (const char* fp = getenv("MLO_FILTER_PPT")) c->filter_ppt = atoi(fp);
  if (const char* fk = getenv("MLO_FILTER_KERNEL")) c->filter_kernel = atoi(fk);
  if (const char* fm = getenv("MLO_FILTER_CTA_MIN_CLOUDS")) c->filter_cta_min_clouds = std::max(1, atoi(fm));
  if (const char* pk = getenv("MLO_PERSISTENT")) {
    c->use_persistent = atoi(pk) != 0;
    c->persistent_forced = atoi(pk) == 2;  // 2 = always, regardless of batch size (experiments)
  }
  cudaSetDevice(cuda_device);
  if (cudaStreamCreateWithFlags(&c->stream, cudaStreamNonBlocking) != cudaSuccess ||
      cudaStreamCreateWithFlags(&c->copy_stream, cudaStreamNonBlocking) != cudaSuccess ||
      cudaEventCreateWithFlags(&c->stage[0].ready, cudaEventDisableTiming) != cudaSuccess ||
      cudaEventCreateWithFlags(&c->stage[1].ready, cudaEventDisableTiming) != cudaSuccess) {
    delete c;
    return MLO_ERR_CUDA;
  }
  {  // keep freed map buffers in the stream-ordered pool instead of returning them to the driver (free_map_buffers)
    cudaMemPool_t pool = nullptr;
    if (cudaDeviceGetDefaultMemPool(&pool, cuda_device) == cudaSuccess) {
      uint64_t keep = ~0ull;
      cudaMemPoolSetAttribute(pool, cudaMemPoolAttrReleaseThreshold, &keep);
    }
  }
  bool ok = cudaEventCreateWithFlags(&c->ev_fork, cudaEventDisableTiming) == cudaSuccess;
  for (int g = 0; ok && g < mlo_ctx::MAX_GROUPS - 1; g++)
    ok = cudaStreamCreateWithFlags(&c->aux_stream[g], cudaStreamNonBlocking) == cudaSuccess &&
         cudaEventCreateWithFlags(&c->ev_join[g], cudaEventDisableTiming) == cudaSuccess;
  if (!ok) {
    delete c;
    return MLO_ERR_CUDA;
  }
  *out = c;
  return MLO_OK;
}

void mlo_destroy(mlo_ctx* c) {
  if (!c) return;
  DeviceGuard g(c->device);
  cudaStreamSynchronize(c->stream);
  for (DBuf* b : {&c->d_in, &c->d_local, &c->d_pairA, &c->d_pairB, &c->d_partials, &c->d_partcnt, &c->d_probs, &c->d_states,
                  &c->d_tables, &c->d_init, &c->d_misc, &c->d_f_tab, &c->d_f_pslot, &c->d_f_flags, &c->d_f_blk, &c->d_f_jobs,
                  &c->d_f_cnt, &c->d_f_map, &c->d_f_icp, &c->d_ins_g, &c->d_ins_slot, &c->d_ins_next, &c->d_queue, &c->d_tchan, &c->d_maps, &c->d_ins_jobs, &c->d_log})
    b->release();
  c->h_misc.release();
  c->h_states.release();
  c->h_stage.release();
  for (auto e : c->ev_pool) cudaEventDestroy(e);
  cudaStreamSynchronize(c->copy_stream);
  for (auto& st : c->stage) {
    st.buf.release();
    if (st.ready) cudaEventDestroy(st.ready);
  }
  for (int g = 0; g < mlo_ctx::MAX_GROUPS - 1; g++) {
    if (c->aux_stream[g]) cudaStreamSynchronize(c->aux_stream[g]), cudaStreamDestroy(c->aux_stream[g]);
    if (c->ev_join[g]) cudaEventDestroy(c->ev_join[g]);
  }
  if (c->ev_fork) cudaEventDestroy(c->ev_fork);
  cudaStreamDestroy(c->copy_stream);
  cudaStreamDestroy(c->stream);
  delete c;
}

const char* mlo_last_error(const mlo_ctx* c) { return c ? c->err.c_str() : "null context"; }

int mlo_device_info(const mlo_ctx* c, char* name, uint32_t name_len, int* sm_count, int* cc_major, int* cc_minor) {
  if (!c) return MLO_ERR_INVALID_ARG;
  if (name && name_len) {
    std::strncpy(name, c->dev_name.c_str(), name_len - 1);
    name[name_len - 1] = 0;
  }
  if (sm_count) *sm_count = c->sm_count;
  if (cc_major) *cc_major = c->cc_major;
  if (cc_minor) *cc_minor = c->cc_minor;
  return MLO_OK;
}
void* mlo_stream(const mlo_ctx* c) { return c ? (void*)c->stream : nullptr; }

int mlo_set_option(mlo_ctx* c, const char* name, int64_t v) {
  if (!c || !name) return MLO_ERR_INVALID_ARG;
  const std::string n(name);
  if (n == "align_path") c->align_path = int(std::min<int64_t>(3, std::max<int64_t>(0, v)));
  else if (n == "tail_path") c->tail_path = v == 2 ? 2 : 3;
  else if (n == "tail_handover") c->tail_handover = v != 0;
  else if (n == "tail_queries_per_sm") c->tail_queries_per_sm = int(std::max<int64_t>(1, v));
  else if (n == "check_every") c->check_every = int(std::max<int64_t>(1, v));
  else if (n == "block_threads") c->block_threads = int(v);
  else if (n == "block_cluster") c->block_cluster = int(v);
  else if (n == "stream_groups") c->stream_groups = int(std::min<int64_t>(mlo_ctx::MAX_GROUPS, std::max<int64_t>(1, v)));
  else if (n == "fuse_inner") c->fuse_inner = v != 0;
  else if (n == "prior_ahead") c->prior_ahead = v != 0;
  else if (n == "force_kernel") c->force_kernel = int(v);
  else if (n == "wl_variant") c->wl_variant = int(v);
  else if (n == "wl_warps") c->wl_warps = int(v);
  else if (n == "large_batch_queries") c->large_batch_queries = uint64_t(std::max<int64_t>(0, v));
  else if (n == "pers_minb") c->pers_minb = v == 2 ? 2 : (v == 4 ? 4 : 0);
  else if (n == "filter_group_mb") c->filter_group_mb = int(std::max<int64_t>(1, v));
  else if (n == "filter_ppt") c->filter_ppt = int(v);
  else if (n == "filter_kernel") {
    c->filter_kernel = int(v);
    c->filter_cta_backoff = 0;
  }
  else if (n == "filter_cta_min_clouds") c->filter_cta_min_clouds = int(std::max<int64_t>(1, v));
  else if (n == "convention_index_floor") c->conv_index_floor = v != 0;
  else if (n == "convention_gm_form") c->conv_gm_form = v != 0;
  else if (n == "convention_cull_metric") c->conv_cull_metric = int(std::min<int64_t>(2, std::max<int64_t>(0, v)));
  else return fail(c, MLO_ERR_INVALID_ARG, "unknown option: " + n);
  return MLO_OK;
}
int mlo_get_option(const mlo_ctx* c, const char* name, int64_t* out) {
  if (!c || !name || !out) return MLO_ERR_INVALID_ARG;
  const std::string n(name);
  if (n == "align_path") *out = c->align_path;
  else if (n == "tail_path") *out = c->tail_path;
  else if (n == "tail_handover") *out = c->tail_handover;
  else if (n == "tail_queries_per_sm") *out = c->tail_queries_per_sm;
  else if (n == "check_every") *out = c->check_every;
  else if (n == "block_threads") *out = c->block_threads;
  else if (n == "block_cluster") *out = c->block_cluster;
  else if (n == "last_block_cluster") *out = c->last_block_cluster;
  else if (n == "last_block_threads") *out = c->last_block_threads;
  else if (n == "stream_groups") *out = c->stream_groups;
  else if (n == "fuse_inner") *out = c->fuse_inner;
  else if (n == "prior_ahead") *out = c->prior_ahead;
  else if (n == "force_kernel") *out = c->force_kernel;
  else if (n == "wl_variant") *out = c->wl_variant;
  else if (n == "wl_warps") *out = c->wl_warps;
  else if (n == "large_batch_queries") *out = int64_t(c->large_batch_queries);
  else if (n == "pers_minb") *out = c->pers_minb;
  else if (n == "filter_group_mb") *out = c->filter_group_mb;
  else if (n == "filter_ppt") *out = c->filter_ppt;
  else if (n == "filter_kernel") *out = c->filter_kernel;
  else if (n == "filter_cta_min_clouds") *out = c->filter_cta_min_clouds;
  else if (n == "last_filter_kernel") *out = c->last_filter_kernel;
  else if (n == "convention_index_floor") *out = c->conv_index_floor;
  else if (n == "convention_gm_form") *out = c->conv_gm_form;
  else if (n == "convention_cull_metric") *out = c->conv_cull_metric;
  else if (n == "last_align_path") *out = c->last_align_path;
  else if (n == "last_stream_groups") *out = c->last_stream_groups;
  else if (n == "last_tail_handover") *out = c->last_tail_handover;
  else return MLO_ERR_INVALID_ARG;
  return MLO_OK;
}
uint64_t mlo_launch_count(const mlo_ctx* c) { return c ? c->launches : 0; }

int32_t mlo_voxel_index(float coord, float voxel_size) { return voxel_index_map(coord, 1.0f / voxel_size); }
void mlo_se3_exp(const double xi[6], double pose[12]) { se3_exp(xi, pose); }
void mlo_se3_log(const double pose[12], double xi[6]) { se3_log(pose, xi); }
void mlo_se3_right_jacobian_inv(const double xi[6], double J[36]) { se3_right_jacobian_inv(xi, J); }
void mlo_cov_tangent_to_ypr(const double T[12], const double cov[36], double out[36]) {
  // T' = T exp(eps): t' = t + R v, R' = R exp(w^).  Euler rates of R = Rz(yaw) Ry(pitch) Rx(roll) from the body rate w:
  //   yaw' = (wy sr + wz cr) / cp,  pitch' = wy cr - wz sr,  roll' = wx + (wy sr + wz cr) tp
  const double pitch = std::atan2(-T[8], std::hypot(T[0], T[4])), roll = std::atan2(T[9], T[10]);
  const double sr = std::sin(roll), cr = std::cos(roll);
  double cp = std::cos(pitch);
  if (std::fabs(cp) < 1e-9) cp = cp < 0 ? -1e-9 : 1e-9;  // gimbal lock: the chart itself is singular there
  const double tp = std::sin(pitch) / cp;
  double J[36] = {0};
  for (int r = 0; r < 3; r++)
    for (int k = 0; k < 3; k++) J[6 * r + k] = T[4 * r + k];
  J[6 * 3 + 4] = sr / cp; J[6 * 3 + 5] = cr / cp;
  J[6 * 4 + 4] = cr;      J[6 * 4 + 5] = -sr;
  J[6 * 5 + 3] = 1.0;     J[6 * 5 + 4] = sr * tp;  J[6 * 5 + 5] = cr * tp;
  double JC[36];
  for (int i = 0; i < 6; i++)
    for (int j = 0; j < 6; j++) {
      double s = 0;
      for (int m = 0; m < 6; m++) s += J[6 * i + m] * cov[6 * m + j];
      JC[6 * i + j] = s;
    }
  for (int i = 0; i < 6; i++)
    for (int j = 0; j < 6; j++) {
      double s = 0;
      for (int m = 0; m < 6; m++) s += JC[6 * i + m] * J[6 * j + m];
      out[6 * i + j] = s;
    }
}

// ------------------------------------------------------------------ map
int mlo_map_create(mlo_ctx* c, const mlo_map_params* p, mlo_map** out) {
  if (!c || !p || !out) return MLO_ERR_INVALID_ARG;
  *out = nullptr;
  if (!(p->voxel_size > 0.f) || p->capacity_voxels == 0 || p->capacity_voxels > (1ull << 26))
    return fail(c, MLO_ERR_INVALID_ARG, "voxel_size must be > 0 and capacity_voxels in [1, 2^26]");
  if (p->kind != MLO_MAP_HASHED_VOXEL_POINTS && p->kind != MLO_MAP_NDT) return fail(c, MLO_ERR_INVALID_ARG, "bad map kind");
  DeviceGuard g(c->device);
  auto* m = new mlo_map;
  m->ctx = c;
  m->prm = *p;
  m->table_size = next_pow2(std::max<uint64_t>(uint64_t(c->table_factor) * p->capacity_voxels, 1024));
  int rc = alloc_map_buffers(c, *p, m->table_size, m->dev);
  if (rc == MLO_OK) rc = clear_map_buffers(c, m->dev, m->table_size);
  if (rc == MLO_OK) {
    if (cudaMallocAsync(&m->head, 4 * m->table_size * sizeof(int32_t), c->stream) != cudaSuccess ||
        cudaMemsetAsync(m->head, 0xFF, 4 * m->table_size * sizeof(int32_t), c->stream) != cudaSuccess)
      rc = fail(c, MLO_ERR_CUDA, "cudaMalloc(head) failed");
  }
  if (rc != MLO_OK) {
    free_map_buffers(m->ctx, m->dev);
    delete m;
    return rc;
  }
  *out = m;
  return MLO_OK;
}

void mlo_map_destroy(mlo_map* m) {
  if (!m) return;
  DeviceGuard g(m->ctx->device);
  cudaStreamSynchronize(m->ctx->stream);
  free_map_buffers(m->ctx, m->dev);
  if (m->alt_ready) free_map_buffers(m->ctx, m->alt);
  if (m->head) cudaFreeAsync(m->head, m->ctx->stream);
  delete m;
}

int mlo_map_clear(mlo_map* m) {
  if (!m) return MLO_ERR_INVALID_ARG;
  DeviceGuard g(m->ctx->device);
  m->hwm = m->n_voxels = m->n_points = m->n_free = 0;
  return clear_map_buffers(m->ctx, m->dev, m->table_size);
}

int mlo_map_insert(mlo_map* m, const float* pts, uint32_t stride, uint64_t n, const double pose[12]) {
  if (!m || (!pts && n) || !pose) return MLO_ERR_INVALID_ARG;
  mlo_ctx* c = m->ctx;
  DeviceGuard g(c->device);
  const size_t e0 = prof_begin(c);
  int rc = upload_strided(c, pts, stride, n, c->d_in);
  if (rc != MLO_OK) return rc;
  rc = map_ensure_capacity(m, n);
  if (rc != MLO_OK) return rc;
  rc = map_insert_device(m, c->d_in.as<float>(), stride, n, pose);
  prof_end(c, 2, e0);
  if (rc != MLO_OK) return rc;
  rc = check_map_errors(m);
  prof_collect(c);
  return rc;
}

int mlo_map_insert_soa(mlo_map* m, const float* x, const float* y, const float* z, uint64_t n, const double pose[12]) {
  if (!m || !pose || (n && (!x || !y || !z))) return MLO_ERR_INVALID_ARG;
  mlo_ctx* c = m->ctx;
  DeviceGuard g(c->device);
  CU(c, c->d_in.ensure(std::max<size_t>(3 * n * sizeof(float), 16)));
  CU(c, c->d_local.ensure(std::max<size_t>(n, 1) * sizeof(float4)));
  float* d = c->d_in.as<float>();
  if (n) {
    CU(c, cudaMemcpyAsync(d, x, n * sizeof(float), cudaMemcpyHostToDevice, c->stream));
    CU(c, cudaMemcpyAsync(d + n, y, n * sizeof(float), cudaMemcpyHostToDevice, c->stream));
    CU(c, cudaMemcpyAsync(d + 2 * n, z, n * sizeof(float), cudaMemcpyHostToDevice, c->stream));
    LAUNCH(c, k_soa_to_float4, uint32_t((n + 255) / 256), 256, d, d + n, d + 2 * n, n, c->d_local.as<float4>());
  }
  int rc = map_ensure_capacity(m, n);
  if (rc != MLO_OK) return rc;
  rc = map_insert_device(m, c->d_local.as<float>(), 4, n, pose);
  if (rc != MLO_OK) return rc;
  return check_map_errors(m);
}

int mlo_map_cull(mlo_map* m, const double sensor[3], float dist) {
  if (!m || !sensor) return MLO_ERR_INVALID_ARG;
  if (!(dist > 0.f)) return MLO_OK;
  mlo_ctx* c = m->ctx;
  DeviceGuard g(c->device);
  const float inv = m->dev.inv_voxel;
  const int32_t sx = voxel_index_map(float(sensor[0]), inv, m->dev.index_floor), sy = voxel_index_map(float(sensor[1]), inv, m->dev.index_floor),
                sz = voxel_index_map(float(sensor[2]), inv, m->dev.index_floor);
  const int32_t d = int32_t(std::ceil(dist * inv));
  const size_t e0 = prof_begin(c);
  int rc = map_cull_device(m, sx, sy, sz, d, 0);
  prof_end(c, 2, e0);
  if (rc != MLO_OK) return rc;
  rc = check_map_errors(m);
  prof_collect(c);
  return rc;
}

int mlo_map_nn_single(const mlo_map* m, const float* q, uint32_t stride, uint64_t n, float* out_xyz, float* out_d2,
                      uint8_t* out_found) {
  if (!m || (n && (!q || !out_xyz || !out_d2 || !out_found))) return MLO_ERR_INVALID_ARG;
  mlo_ctx* c = m->ctx;
  DeviceGuard g(c->device);
  if (n == 0) return MLO_OK;
  int rc = upload_strided(c, q, stride, n, c->d_in);
  if (rc != MLO_OK) return rc;
  CU(c, c->d_local.ensure(n * (3 * sizeof(float) + sizeof(float) + 1) + 64));
  float* dxyz = c->d_local.as<float>();
  float* dd2 = dxyz + 3 * n;
  uint8_t* df = reinterpret_cast<uint8_t*>(dd2 + n);
  LAUNCH(c, k_nn_single, uint32_t((n * 32 + 127) / 128), 128, m->dev, c->d_in.as<float>(), stride, uint32_t(n), dxyz, dd2, df);
  CU(c, cudaMemcpyAsync(out_xyz, dxyz, 3 * n * sizeof(float), cudaMemcpyDeviceToHost, c->stream));
  CU(c, cudaMemcpyAsync(out_d2, dd2, n * sizeof(float), cudaMemcpyDeviceToHost, c->stream));
  CU(c, cudaMemcpyAsync(out_found, df, n, cudaMemcpyDeviceToHost, c->stream));
  CU(c, cudaStreamSynchronize(c->stream));
  CU(c, cudaGetLastError());
  return MLO_OK;
}

int mlo_map_nn_plane(const mlo_map* m, const float* q, uint32_t stride, uint64_t n, float* out_mean, float* out_normal,
                     float* out_dist, uint8_t* out_found) {
  if (!m || (n && (!q || !out_mean || !out_normal || !out_dist || !out_found))) return MLO_ERR_INVALID_ARG;
  mlo_ctx* c = m->ctx;
  if (m->prm.kind != MLO_MAP_NDT) return fail(c, MLO_ERR_UNSUPPORTED, "nearest-plane queries need an NDT map");
  DeviceGuard g(c->device);
  if (n == 0) return MLO_OK;
  int rc = upload_strided(c, q, stride, n, c->d_in);
  if (rc != MLO_OK) return rc;
  CU(c, c->d_local.ensure(n * (7 * sizeof(float) + 1) + 64));
  float* dm = c->d_local.as<float>();
  float* dn = dm + 3 * n;
  float* dd = dn + 3 * n;
  uint8_t* df = reinterpret_cast<uint8_t*>(dd + n);
  LAUNCH(c, k_nn_plane, uint32_t((n + 127) / 128), 128, m->dev, c->d_in.as<float>(), stride, uint32_t(n), dm, dn, dd, df);
  CU(c, cudaMemcpyAsync(out_mean, dm, 3 * n * sizeof(float), cudaMemcpyDeviceToHost, c->stream));
  CU(c, cudaMemcpyAsync(out_normal, dn, 3 * n * sizeof(float), cudaMemcpyDeviceToHost, c->stream));
  CU(c, cudaMemcpyAsync(out_dist, dd, n * sizeof(float), cudaMemcpyDeviceToHost, c->stream));
  CU(c, cudaMemcpyAsync(out_found, df, n, cudaMemcpyDeviceToHost, c->stream));
  CU(c, cudaStreamSynchronize(c->stream));
  CU(c, cudaGetLastError());
  return MLO_OK;
}

int mlo_map_stats(const mlo_map* m, uint64_t* n_voxels, uint64_t* n_points) {
  if (!m) return MLO_ERR_INVALID_ARG;
  mlo_ctx* c = m->ctx;
  DeviceGuard g(c->device);
  uint32_t h[MAP_COUNTERS];
  CU(c, cudaMemcpyAsync(h, m->dev.counters, sizeof(h), cudaMemcpyDeviceToHost, c->stream));
  CU(c, cudaStreamSynchronize(c->stream));
  if (n_voxels) *n_voxels = uint64_t(std::min<uint32_t>(h[0], m->dev.capacity_voxels)) - h[3];
  if (n_points) *n_points = h[1];
  return MLO_OK;
}

int mlo_map_export(const mlo_map* m, int32_t* keys, uint32_t* counts, float* xyz, uint64_t max_voxels, uint64_t max_points,
                   uint64_t* n_voxels, uint64_t* n_points) {
  if (!m || !n_voxels || !n_points) return MLO_ERR_INVALID_ARG;
  mlo_ctx* c = m->ctx;
  DeviceGuard g(c->device);
  uint64_t nv = 0, np = 0;
  int rc = mlo_map_stats(m, &nv, &np);
  if (rc != MLO_OK) return rc;
  *n_voxels = nv;
  *n_points = np;
  if (!keys || !counts || !xyz) return MLO_OK;
  if (nv > max_voxels || np > max_points) return fail(c, MLO_ERR_INVALID_ARG, "export buffers too small");
  if (nv == 0) return MLO_OK;
  CU(c, c->d_local.ensure(nv * 5 * sizeof(uint32_t) + 64));
  int32_t* dk = c->d_local.as<int32_t>();
  uint32_t* dc = reinterpret_cast<uint32_t*>(dk + 3 * nv);
  uint32_t* dv = dc + nv;
  CU(c, c->d_misc.ensure(256));
  uint32_t* cursor = c->d_misc.as<uint32_t>() + 16;
  CU(c, cudaMemsetAsync(cursor, 0, sizeof(uint32_t), c->stream));
  LAUNCH(c, k_export_list, uint32_t((m->table_size * 4 + 255) / 256), 256, m->dev, m->table_size, cursor, dk, dc, dv);
  std::vector<int32_t> k3(3 * nv);
  std::vector<uint32_t> hc(nv), hv(nv);
  CU(c, cudaMemcpyAsync(k3.data(), dk, 3 * nv * sizeof(int32_t), cudaMemcpyDeviceToHost, c->stream));
  CU(c, cudaMemcpyAsync(hc.data(), dc, nv * sizeof(uint32_t), cudaMemcpyDeviceToHost, c->stream));
  CU(c, cudaMemcpyAsync(hv.data(), dv, nv * sizeof(uint32_t), cudaMemcpyDeviceToHost, c->stream));
  CU(c, cudaStreamSynchronize(c->stream));
  // payload of the allocated voxels (ids are dense in [0, nv) right after a rebuild, sparse otherwise)
  uint32_t max_vid = 0;
  for (auto v : hv) max_vid = std::max(max_vid, v);
  std::vector<float4> hp(size_t(max_vid + 1) * m->dev.row);
  CU(c, cudaMemcpyAsync(hp.data(), m->dev.pts, hp.size() * sizeof(float4), cudaMemcpyDeviceToHost, c->stream));
  CU(c, cudaStreamSynchronize(c->stream));
  std::vector<uint32_t> order(nv);
  for (uint32_t i = 0; i < nv; i++) order[i] = i;
  std::sort(order.begin(), order.end(), [&](uint32_t a, uint32_t b) {
    for (int k = 0; k < 3; k++)
      if (k3[3 * a + k] != k3[3 * b + k]) return k3[3 * a + k] < k3[3 * b + k];
    return false;
  });
  uint64_t o = 0;
  for (uint32_t i = 0; i < nv; i++) {
    const uint32_t s = order[i];
    for (int k = 0; k < 3; k++) keys[3 * i + k] = k3[3 * s + k];
    counts[i] = hc[s];
    for (uint32_t j = 0; j < hc[s]; j++) {
      if (o >= max_points) return fail(c, MLO_ERR_INVALID_ARG, "export buffers too small");
      const float4 p = hp[size_t(hv[s]) * m->dev.row + j];
      xyz[3 * o] = p.x;
      xyz[3 * o + 1] = p.y;
      xyz[3 * o + 2] = p.z;
      o++;
    }
  }
  *n_points = o;
  return MLO_OK;
}

// ------------------------------------------------------------------ filters
int mlo_voxel_decimate_first(mlo_ctx* c, const float* pts, uint32_t stride, uint64_t n, const mlo_decimate_params* p,
                             uint32_t* out_idx, uint64_t* out_n) {
  if (!c || !p || !out_n || (n && (!pts || !out_idx))) return MLO_ERR_INVALID_ARG;
  DeviceGuard g(c->device);
  *out_n = 0;
  if (n == 0) return MLO_OK;
  if (!(p->voxel_filter_resolution > 0.f)) return fail(c, MLO_ERR_INVALID_ARG, "voxel_filter_resolution must be > 0");
  int rc = upload_strided(c, pts, stride, n, c->d_in);
  if (rc != MLO_OK) return rc;
  CU(c, c->d_local.ensure(n * sizeof(uint32_t)));
  mlo_filter1_params fp;
  std::memset(&fp, 0, sizeof(fp));
  fp.for_map = *p;
  fp.for_icp = *p;
  const uint64_t offs[2] = {0, n};
  FilterBatch fb;
  const size_t e0 = prof_begin(c);
  rc = run_filter_batch(c, c->d_in.as<float>(), stride, 1, offs, &fp, true, c->d_local.as<uint32_t>(), fb);
  prof_end(c, 0, e0);
  if (rc != MLO_OK) return rc;
  std::vector<uint32_t> h;
  rc = filter_counts(c, fb, h);
  if (rc != MLO_OK) return rc;
  *out_n = h[0];
  CU(c, cudaMemcpyAsync(out_idx, c->d_local.p, size_t(h[0]) * sizeof(uint32_t), cudaMemcpyDeviceToHost, c->stream));
  CU(c, cudaStreamSynchronize(c->stream));
  prof_collect(c);
  return MLO_OK;
}

static int download_xyz(mlo_ctx* c, const float4* d, uint64_t n, float* out) {
  if (n == 0) return MLO_OK;
  std::vector<float4> tmp(n);
  CU(c, cudaMemcpyAsync(tmp.data(), d, n * sizeof(float4), cudaMemcpyDeviceToHost, c->stream));
  CU(c, cudaStreamSynchronize(c->stream));
  for (uint64_t i = 0; i < n; i++) {
    out[3 * i] = tmp[i].x;
    out[3 * i + 1] = tmp[i].y;
    out[3 * i + 2] = tmp[i].z;
  }
  return MLO_OK;
}

int mlo_filter_1st_pass(mlo_ctx* c, const float* pts, uint32_t stride, uint64_t n, const mlo_filter1_params* p,
                        float* out_map_xyz, uint64_t* out_map_n, float* out_icp_xyz, uint64_t* out_icp_n) {
  if (!c || !p || !out_map_n || !out_icp_n || (n && !pts)) return MLO_ERR_INVALID_ARG;
  DeviceGuard g(c->device);
  *out_map_n = *out_icp_n = 0;
  if (n == 0) return MLO_OK;
  int rc = upload_strided(c, pts, stride, n, c->d_in);
  if (rc != MLO_OK) return rc;
  const uint64_t offs[2] = {0, n};
  FilterBatch fb;
  const size_t e0 = prof_begin(c);
  rc = run_filter_batch(c, c->d_in.as<float>(), stride, 1, offs, p, false, nullptr, fb);
  prof_end(c, 0, e0);
  if (rc != MLO_OK) return rc;
  std::vector<uint32_t> h;
  rc = filter_counts(c, fb, h);
  if (rc != MLO_OK) return rc;
  *out_map_n = h[1];
  *out_icp_n = h[2];
  if (out_map_xyz) {
    rc = download_xyz(c, c->d_f_map.as<float4>(), h[1], out_map_xyz);
    if (rc != MLO_OK) return rc;
  }
  if (out_icp_xyz) {
    rc = download_xyz(c, c->d_f_icp.as<float4>(), h[2], out_icp_xyz);
    if (rc != MLO_OK) return rc;
  }
  prof_collect(c);
  return MLO_OK;
}

int mlo_filter_1st_pass_xyzt(mlo_ctx* c, const float* pts, uint32_t stride, const float* t, uint64_t n,
                             const mlo_filter1_params* p, float* out_map_xyzt, uint64_t* out_map_n, float* out_icp_xyzt,
                             uint64_t* out_icp_n) {
  if (!c || !p || !out_map_n || !out_icp_n || (n && !pts)) return MLO_ERR_INVALID_ARG;
  DeviceGuard g(c->device);
  *out_map_n = *out_icp_n = 0;
  if (n == 0) return MLO_OK;
  int rc = upload_strided(c, pts, stride, n, c->d_in);
  if (rc != MLO_OK) return rc;
  const float* d_t = nullptr;
  if (t) {
    CU(c, c->d_tchan.ensure(n * sizeof(float)));
    CU(c, cudaMemcpyAsync(c->d_tchan.p, t, n * sizeof(float), cudaMemcpyHostToDevice, c->stream));
    d_t = c->d_tchan.as<float>();
  }
  const uint64_t offs[2] = {0, n};
  FilterBatch fb;
  const size_t e0 = prof_begin(c);
  rc = run_filter_batch(c, c->d_in.as<float>(), stride, 1, offs, p, false, nullptr, fb, d_t);
  prof_end(c, 0, e0);
  if (rc != MLO_OK) return rc;
  std::vector<uint32_t> h;
  rc = filter_counts(c, fb, h);
  if (rc != MLO_OK) return rc;
  *out_map_n = h[1];
  *out_icp_n = h[2];
  if (out_map_xyzt && h[1])
    CU(c, cudaMemcpyAsync(out_map_xyzt, c->d_f_map.p, size_t(h[1]) * sizeof(float4), cudaMemcpyDeviceToHost, c->stream));
  if (out_icp_xyzt && h[2])
    CU(c, cudaMemcpyAsync(out_icp_xyzt, c->d_f_icp.p, size_t(h[2]) * sizeof(float4), cudaMemcpyDeviceToHost, c->stream));
  CU(c, cudaStreamSynchronize(c->stream));
  prof_collect(c);
  return MLO_OK;
}

int mlo_deskew(mlo_ctx* c, const float* xyzt, uint64_t n, const double twist[6], float* out_xyz) {
  if (!c || !twist || (n && (!xyzt || !out_xyz))) return MLO_ERR_INVALID_ARG;
  DeviceGuard g(c->device);
  if (n == 0) return MLO_OK;
  CU(c, c->d_in.ensure(n * sizeof(float4)));
  CU(c, c->d_local.ensure(n * sizeof(float4)));
  CU(c, cudaMemcpyAsync(c->d_in.p, xyzt, n * sizeof(float4), cudaMemcpyHostToDevice, c->stream));
  Twist6 tw;
  std::memcpy(tw.v, twist, sizeof(tw.v));
  const size_t e0 = prof_begin(c);
  LAUNCH(c, k_deskew, uint32_t((n + 255) / 256), 256, c->d_in.as<float4>(), uint32_t(n), tw, c->d_local.as<float4>());
  prof_end(c, 0, e0);
  CU(c, cudaGetLastError());
  int rc = download_xyz(c, c->d_local.as<float4>(), n, out_xyz);
  prof_collect(c);
  return rc;
}

// ------------------------------------------------------------------ ICP
void mlo_icp_params_default(mlo_icp_params* p) {
  if (!p) return;
  std::memset(p, 0, sizeof(*p));
  p->max_iterations = 300;       // default.yaml:173
  p->min_abs_step_trans = 1e-4;  // :174
  p->min_abs_step_rot = 5e-5;    // :175
  p->solver = MLO_SOLVER_GAUSS_NEWTON;
  p->gn_max_iterations = 2;  // :187
  p->gn_min_delta = 1e-7;
  p->robust_kernel = MLO_KERNEL_GEMAN_MCCLURE;  // :188
  p->matcher_mask = MLO_MATCHER_PT2PT;          // :196
  p->pt2pt_weight = p->pt2pl_weight = 1.0;
  p->prior_pose_3x4[0] = p->prior_pose_3x4[5] = p->prior_pose_3x4[10] = 1.0;
  p->hook_checkpoint_pose_3x4[0] = p->hook_checkpoint_pose_3x4[5] = p->hook_checkpoint_pose_3x4[10] = 1.0;
}

int mlo_icp_log_enable(mlo_ctx* c, uint32_t max_records) {
  if (!c) return MLO_ERR_INVALID_ARG;
  c->log_cap = std::min<uint32_t>(max_records, 1024u);
  c->h_log_problems = 0;
  return MLO_OK;
}
int mlo_icp_log_read(mlo_ctx* c, uint32_t problem, mlo_icp_iteration_record* out, uint32_t max_records, uint32_t* n) {
  if (!c || !n) return MLO_ERR_INVALID_ARG;
  *n = 0;
  if (!c->log_cap || problem >= c->h_log_problems) return fail(c, MLO_ERR_INVALID_ARG, "no ICP log for this problem (mlo_icp_log_enable first)");
  const mlo_icp_iteration_record* r = c->h_log.data() + size_t(problem) * c->log_cap;
  uint32_t k = 0;
  while (k < c->log_cap && r[k].iteration == k) k++;  // (unwritten records are all-ones)
  *n = k;
  if (out)
    for (uint32_t i = 0; i < k && i < max_records; i++) out[i] = r[i];
  return MLO_OK;
}

int mlo_icp_align_batch(mlo_ctx* c, uint32_t B, const float* local, uint32_t stride, const uint64_t* offsets,
                        const mlo_map* map, const double* init_poses, const mlo_icp_params* params, mlo_icp_result* out) {
  if (!c || !map || !offsets || !init_poses || !params || !out) return MLO_ERR_INVALID_ARG;
  if (map->ctx != c) return fail(c, MLO_ERR_INVALID_ARG, "map belongs to another context");
  DeviceGuard g(c->device);
  const uint64_t total = offsets[B];
  int rc = upload_strided(c, local, stride, total, c->d_in);
  if (rc != MLO_OK) return rc;
  CU(c, c->d_local.ensure(std::max<size_t>(total, 1) * sizeof(float4)));
  if (total) LAUNCH(c, k_to_float4, uint32_t((total + 255) / 256), 256, c->d_in.as<float>(), stride, total, c->d_local.as<float4>());
  return align_batch_device(c, B, c->d_local.as<float4>(), offsets, map, init_poses, params, out);
}

int mlo_icp_align(mlo_ctx* c, const float* local, uint32_t stride, uint64_t n, const mlo_map* map, const double init[12],
                  const mlo_icp_params* p, mlo_icp_result* out) {
  const uint64_t offs[2] = {0, n};
  return mlo_icp_align_batch(c, 1, local, stride, offs, map, init, p, out);
}

int mlo_icp_align_soa(mlo_ctx* c, const float* x, const float* y, const float* z, uint64_t n, const mlo_map* map,
                      const double init[12], const mlo_icp_params* p, mlo_icp_result* out) {
  if (!c || !map || !init || !p || !out || (n && (!x || !y || !z))) return MLO_ERR_INVALID_ARG;
  DeviceGuard g(c->device);
  CU(c, c->d_in.ensure(std::max<size_t>(3 * n * sizeof(float), 16)));
  CU(c, c->d_local.ensure(std::max<size_t>(n, 1) * sizeof(float4)));
  float* d = c->d_in.as<float>();
  if (n) {
    CU(c, cudaMemcpyAsync(d, x, n * sizeof(float), cudaMemcpyHostToDevice, c->stream));
    CU(c, cudaMemcpyAsync(d + n, y, n * sizeof(float), cudaMemcpyHostToDevice, c->stream));
    CU(c, cudaMemcpyAsync(d + 2 * n, z, n * sizeof(float), cudaMemcpyHostToDevice, c->stream));
    LAUNCH(c, k_soa_to_float4, uint32_t((n + 255) / 256), 256, d, d + n, d + 2 * n, n, c->d_local.as<float4>());
  }
  const uint64_t offs[2] = {0, n};
  return align_batch_device(c, 1, c->d_local.as<float4>(), offs, map, init, p, out);
}

// ------------------------------------------------------------------ device-resident clouds
int mlo_dcloud_upload_batch(mlo_ctx* c, const float* pts, uint32_t stride, uint32_t n_clouds, const uint64_t* offsets,
                            mlo_dcloud** out) {
  if (!c || !out || !offsets || n_clouds == 0) return MLO_ERR_INVALID_ARG;
  DeviceGuard g(c->device);
  *out = nullptr;
  const uint64_t total = offsets[n_clouds];
  int rc = upload_strided(c, pts, stride, total, c->d_in);
  if (rc != MLO_OK) return rc;
  auto* dc = new mlo_dcloud;
  dc->ctx = c;
  dc->n = total;
  dc->offsets.assign(offsets, offsets + n_clouds + 1);
  if (cudaMalloc(&dc->pts, std::max<size_t>(total, 1) * sizeof(float4)) != cudaSuccess) {
    delete dc;
    return fail(c, MLO_ERR_CUDA, "cudaMalloc(dcloud) failed");
  }
  if (total) LAUNCH(c, k_to_float4, uint32_t((total + 255) / 256), 256, c->d_in.as<float>(), stride, total, dc->pts);
  CU(c, cudaStreamSynchronize(c->stream));
  *out = dc;
  return MLO_OK;
}
int mlo_dcloud_upload(mlo_ctx* c, const float* pts, uint32_t stride, uint64_t n, mlo_dcloud** out) {
  const uint64_t offs[2] = {0, n};
  return mlo_dcloud_upload_batch(c, pts, stride, 1, offs, out);
}
void mlo_dcloud_destroy(mlo_dcloud* d) {
  if (!d) return;
  DeviceGuard g(d->ctx->device);
  cudaStreamSynchronize(d->ctx->stream);
  if (d->pts) cudaFree(d->pts);
  delete d;
}
uint64_t mlo_dcloud_size(const mlo_dcloud* d) { return d ? d->n : 0; }

int mlo_icp_align_batch_resident(mlo_ctx* c, const mlo_dcloud* local, const mlo_map* map, const double* init_poses,
                                 const mlo_icp_params* params, mlo_icp_result* out) {
  if (!c || !local || !map || !init_poses || !params || !out) return MLO_ERR_INVALID_ARG;
  if (map->ctx != c || local->ctx != c) return fail(c, MLO_ERR_INVALID_ARG, "handle belongs to another context");
  DeviceGuard g(c->device);
  return align_batch_device(c, uint32_t(local->offsets.size() - 1), local->pts, local->offsets.data(), map, init_poses, params, out);
}

// filter -> align over device-resident raw clouds (float pointer + stride)
static int scan_register_device(mlo_ctx* c, const mlo_map* map, uint32_t B, const float* d_raw, uint32_t stride,
                                const uint64_t* offsets, const mlo_filter1_params* fps, const double* init_poses,
                                const mlo_icp_params* ips, mlo_icp_result* out, std::vector<uint64_t>* map_layer_n) {
  FilterBatch fb;
  const size_t e0 = prof_begin(c);
  int rc = run_filter_batch(c, d_raw, stride, B, offsets, fps, false, nullptr, fb);
  prof_end(c, 0, e0);
  if (rc != MLO_OK) return rc;
  // the ICP grid depends on the decimated sizes: one small D2H of the device counters
  std::vector<uint32_t> h;
  rc = filter_counts(c, fb, h);
  if (rc != MLO_OK) return rc;
  // the ICP layers stay where the filter wrote them (at the raw offsets): problem b = [offsets[b], offsets[b] + n_icp)
  std::vector<uint64_t> qb(offsets, offsets + B);
  std::vector<uint32_t> nq(B);
  std::vector<const mlo_map*> maps(B, map);
  for (uint32_t b = 0; b < B; b++) {
    nq[b] = h[b * CNT_STRIDE + 2];
    if (map_layer_n) (*map_layer_n)[b] = h[b * CNT_STRIDE + 1];
  }
  return align_batch_core(c, B, c->d_f_icp.as<float4>(), qb.data(), nq.data(), maps.data(), init_poses, ips, out);
}

int mlo_scan_register_batch_resident(mlo_ctx* c, const mlo_map* map, const mlo_dcloud* raw, const mlo_filter1_params* fps,
                                     const double* init_poses, const mlo_icp_params* ips, mlo_icp_result* out) {
  if (!c || !map || !raw || !fps || !init_poses || !ips || !out) return MLO_ERR_INVALID_ARG;
  if (map->ctx != c || raw->ctx != c) return fail(c, MLO_ERR_INVALID_ARG, "handle belongs to another context");
  DeviceGuard g(c->device);
  return scan_register_device(c, map, uint32_t(raw->offsets.size() - 1), reinterpret_cast<const float*>(raw->pts), 4,
                              raw->offsets.data(), fps, init_poses, ips, out, nullptr);
}

int mlo_scan_register_batch(mlo_ctx* c, const mlo_map* map, uint32_t B, const float* raw, uint32_t stride,
                            const uint64_t* offsets, const mlo_filter1_params* fps, const double* init_poses,
                            const mlo_icp_params* ips, mlo_icp_result* out) {
  if (!c || !map || !raw || !offsets || !fps || !init_poses || !ips || !out || B == 0) return MLO_ERR_INVALID_ARG;
  if (map->ctx != c) return fail(c, MLO_ERR_INVALID_ARG, "map belongs to another context");
  DeviceGuard g(c->device);
  int rc = upload_strided(c, raw, stride, offsets[B], c->d_in);
  if (rc != MLO_OK) return rc;
  return scan_register_device(c, map, B, c->d_in.as<float>(), stride, offsets, fps, init_poses, ips, out, nullptr);
}

int mlo_stage_upload_async(mlo_ctx* c, int slot, const float* raw, uint32_t stride, uint32_t B, const uint64_t* offsets) {
  if (!c || !raw || !offsets || B == 0 || slot < 0 || slot > 1) return MLO_ERR_INVALID_ARG;
  if (stride != 3 && stride != 4) return fail(c, MLO_ERR_INVALID_ARG, "stride_floats must be 3 or 4");
  auto& st = c->stage[slot];
  st.offsets.assign(offsets, offsets + B + 1);
  st.stride = stride;
  st.pending = true;
  // Deferred: the transfer is enqueued by the next compute call of this context right after that call's own
  // small parameter uploads (the H2D copy engine serves transfers in submission order, so a 1 GB transfer
  // submitted first would stall them), or at the latest when this slot is consumed.
  st.deferred_src = raw;
  return MLO_OK;
}

int mlo_scan_register_batch_staged(mlo_ctx* c, const mlo_map* map, int slot, const mlo_filter1_params* fps,
                                   const double* init_poses, const mlo_icp_params* ips, mlo_icp_result* out) {
  if (!c || !map || !fps || !init_poses || !ips || !out || slot < 0 || slot > 1) return MLO_ERR_INVALID_ARG;
  if (map->ctx != c) return fail(c, MLO_ERR_INVALID_ARG, "map belongs to another context");
  auto& st = c->stage[slot];
  if (!st.pending) return fail(c, MLO_ERR_INVALID_ARG, "nothing staged in this slot");
  DeviceGuard g(c->device);
  int rc = issue_stage_upload(c, slot);  // not yet enqueued (no compute call ran in between): do it now
  if (rc != MLO_OK) return rc;
  CU(c, cudaStreamWaitEvent(c->stream, st.ready, 0));
  st.pending = false;
  c->consuming_slot = slot;
  rc = scan_register_device(c, map, uint32_t(st.offsets.size() - 1), st.buf.as<float>(), st.stride, st.offsets.data(), fps,
                            init_poses, ips, out, nullptr);
  c->consuming_slot = -1;
  return rc;
}

int mlo_scan_register(mlo_ctx* c, mlo_map* map, const float* raw, uint32_t stride, uint64_t n, const mlo_filter1_params* fp,
                      const double init[12], const mlo_icp_params* ip, int insert_into_map, float cull_farther_than,
                      mlo_icp_result* out) {
  if (!c || !map || !raw || !fp || !init || !ip || !out) return MLO_ERR_INVALID_ARG;
  if (map->ctx != c) return fail(c, MLO_ERR_INVALID_ARG, "map belongs to another context");
  DeviceGuard g(c->device);
  int rc = upload_strided(c, raw, stride, n, c->d_in);
  if (rc != MLO_OK) return rc;
  const uint64_t offs[2] = {0, n};
  std::vector<uint64_t> nmap(1, 0);
  rc = scan_register_device(c, map, 1, c->d_in.as<float>(), stride, offs, fp, init, ip, out, &nmap);
  if (rc != MLO_OK) return rc;
  if (insert_into_map) {
    const size_t e0 = prof_begin(c);
    rc = map_ensure_capacity(map, nmap[0]);
    if (rc != MLO_OK) return rc;
    rc = map_insert_device(map, reinterpret_cast<const float*>(c->d_f_map.as<float4>()), 4, nmap[0], out->pose_3x4);
    if (rc != MLO_OK) return rc;
    if (cull_farther_than > 0.f) {
      const double s[3] = {out->pose_3x4[3], out->pose_3x4[7], out->pose_3x4[11]};
      const float inv = map->dev.inv_voxel;
      rc = map_cull_device(map, voxel_index_map(float(s[0]), inv, map->dev.index_floor), voxel_index_map(float(s[1]), inv, map->dev.index_floor),
                           voxel_index_map(float(s[2]), inv, map->dev.index_floor), int32_t(std::ceil(cull_farther_than * inv)), nmap[0]);
      if (rc != MLO_OK) return rc;
    }
    prof_end(c, 2, e0);
    rc = check_map_errors(map);
    prof_collect(c);
  }
  return rc;
}

// ------------------------------------------------------------------ scan sets (device-resident layers, fleets)
static int scanset_bbox(mlo_scanset* set, uint32_t n, const uint32_t* slots, mlo_scan_info* info) {
  mlo_ctx* c = set->ctx;
  if (n == 0) return MLO_OK;
  std::vector<BBoxJob> jobs(n);
  for (uint32_t i = 0; i < n; i++) {
    const auto& sl = set->slots[slots[i]];
    jobs[i].p = set->icp_layer() + sl.off;
    jobs[i].n = sl.valid ? sl.n_icp : 0;
  }
  CU(c, set->bbox_jobs.ensure(n * sizeof(BBoxJob)));
  CU(c, set->bbox_out.ensure(n * 6 * sizeof(float)));
  CU(c, cudaMemcpyAsync(set->bbox_jobs.p, jobs.data(), n * sizeof(BBoxJob), cudaMemcpyHostToDevice, c->stream));
  LAUNCH(c, k_bbox, n, 256, set->bbox_jobs.as<BBoxJob>(), set->bbox_out.as<float>());
  std::vector<float> h(6 * size_t(n));
  CU(c, cudaMemcpyAsync(h.data(), set->bbox_out.p, h.size() * sizeof(float), cudaMemcpyDeviceToHost, c->stream));
  CU(c, cudaStreamSynchronize(c->stream));
  CU(c, cudaGetLastError());
  for (uint32_t i = 0; i < n; i++) {
    const auto& sl = set->slots[slots[i]];
    info[i].n_map = sl.n_map;
    info[i].n_icp = sl.n_icp;
    for (int k = 0; k < 3; k++) {
      info[i].icp_min[k] = sl.n_icp ? h[6 * i + k] : 0.f;
      info[i].icp_max[k] = sl.n_icp ? h[6 * i + 3 + k] : 0.f;
    }
  }
  return MLO_OK;
}

int mlo_scanset_create(mlo_ctx* c, uint32_t n_slots, mlo_scanset** out) {
  if (!c || !out || n_slots == 0) return MLO_ERR_INVALID_ARG;
  DeviceGuard g(c->device);
  auto* s = new mlo_scanset();
  s->ctx = c;
  s->slots.resize(n_slots);
  if (cudaEventCreateWithFlags(&s->staged_ready, cudaEventDisableTiming) != cudaSuccess) {
    delete s;
    return fail(c, MLO_ERR_CUDA, "cudaEventCreate failed");
  }
  *out = s;
  return MLO_OK;
}

// enqueue the announced transfer on the copy stream (called from inside a compute call, or by the filter itself)
static int scanset_issue_prefetch(mlo_scanset* set) {
  mlo_ctx* c = set->ctx;
  if (!set->pend.valid) return MLO_OK;
  uint64_t total = 0;
  for (const auto& st : set->pend.clouds) total += st.n;
  const uint32_t stride = set->pend.stride;
  const size_t bytes = std::max<size_t>(total * stride * sizeof(float), 16);
  if (set->raw_next.cap < bytes) {
    CU(c, cudaStreamSynchronize(c->copy_stream));
    CU(c, cudaStreamSynchronize(c->stream));
    CU(c, set->raw_next.ensure(bytes));
  }
  // raw_next was the raw buffer of an EARLIER filter call, and every filter call ends with a synchronisation of the
  // context stream (it returns the layer sizes): nothing in flight reads raw_next, so the copy needs no ordering
  // against the context stream - in particular it must not wait for the ICP kernel it is meant to overlap.
  uint64_t off = 0;
  for (const auto& st : set->pend.clouds) {
    CU(c, cudaMemcpyAsync(set->raw_next.as<float>() + off * stride, st.src, st.n * stride * sizeof(float), cudaMemcpyHostToDevice,
                          c->copy_stream));
    off += st.n;
  }
  CU(c, cudaEventRecord(set->staged_ready, c->copy_stream));
  set->fly = std::move(set->pend);  // (an older, never consumed transfer is simply overwritten)
  set->pend.clear();
  return MLO_OK;
}

int mlo_scanset_prefetch(mlo_scanset* set, uint32_t n_clouds, const float* const* pts, const uint64_t* n, uint32_t stride) {
  if (!set || (n_clouds && (!pts || !n))) return MLO_ERR_INVALID_ARG;
  mlo_ctx* c = set->ctx;
  if (stride != 3 && stride != 4) return fail(c, MLO_ERR_INVALID_ARG, "stride_floats must be 3 or 4");
  set->pend.clear();
  for (uint32_t j = 0; j < n_clouds; j++)
    if (pts[j] && n[j]) set->pend.clouds.push_back({pts[j], n[j]});
  set->pend.stride = stride;
  set->pend.valid = !set->pend.clouds.empty();
  c->drop_deferred(set);
  if (set->pend.valid) c->deferred.push_back({set, [set]() { return scanset_issue_prefetch(set); }});
  return MLO_OK;
}

void mlo_scanset_destroy(mlo_scanset* s) {
  if (!s) return;
  DeviceGuard g(s->ctx->device);
  cudaStreamSynchronize(s->ctx->stream);
  cudaStreamSynchronize(s->ctx->copy_stream);
  s->ctx->drop_deferred(s);  // (a registered prefetch of this set must not outlive it)
  for (DBuf* b : {&s->raw, &s->raw_next, &s->tchan, &s->mapS, &s->icpS, &s->mapL, &s->icpL, &s->cnt, &s->bbox_jobs, &s->bbox_out})
    b->release();
  if (s->staged_ready) cudaEventDestroy(s->staged_ready);
  delete s;
}

int mlo_scanset_filter(mlo_scanset* set, uint32_t n_jobs, const mlo_scan_job* jobs, uint32_t stride, mlo_scan_info* info) {
  if (!set || (n_jobs && (!jobs || !info))) return MLO_ERR_INVALID_ARG;
  mlo_ctx* c = set->ctx;
  if (stride != 3 && stride != 4) return fail(c, MLO_ERR_INVALID_ARG, "stride_floats must be 3 or 4");
  DeviceGuard g(c->device);
  for (auto& sl : set->slots) sl = mlo_scanset::Slot{};
  set->skewed = set->deskewed = false;
  if (n_jobs == 0) return MLO_OK;
  std::vector<uint64_t> off(n_jobs + 1, 0);
  std::vector<mlo_filter1_params> fps(n_jobs);
  bool any_t = false;
  for (uint32_t j = 0; j < n_jobs; j++) {
    if (jobs[j].slot >= set->slots.size() || (jobs[j].n && !jobs[j].pts)) return fail(c, MLO_ERR_INVALID_ARG, "bad scan job");
    if (set->slots[jobs[j].slot].valid) return fail(c, MLO_ERR_INVALID_ARG, "two jobs for one slot");
    set->slots[jobs[j].slot].valid = true;
    off[j + 1] = off[j] + jobs[j].n;
    fps[j] = jobs[j].fp;
    any_t = any_t || (jobs[j].t && jobs[j].n);
  }
  const uint64_t total = off[n_jobs];
  // were exactly these clouds announced (same host pointers, sizes, order)?  Then the raw data is already on its way.
  auto matches = [&](const mlo_scanset::Announce& a) {
    if (!a.valid || a.stride != stride) return false;
    size_t k = 0;
    for (uint32_t j = 0; j < n_jobs; j++) {
      if (!jobs[j].n) continue;
      if (k >= a.clouds.size() || a.clouds[k].src != jobs[j].pts || a.clouds[k].n != jobs[j].n) return false;
      k++;
    }
    return k == a.clouds.size();
  };
  bool prefetched = matches(set->fly);
  if (!prefetched && matches(set->pend)) {  // announced, but no compute call ran in between: enqueue it now
    int rc = scanset_issue_prefetch(set);
    if (rc != MLO_OK) return rc;
    c->drop_deferred(set);
    prefetched = true;
  }
  if (prefetched) {
    CU(c, cudaStreamWaitEvent(c->stream, set->staged_ready, 0));
    std::swap(set->raw, set->raw_next);
    set->fly.clear();
  } else {
    // other clouds than the announced ones (e.g. the announcement is for the call after this one): plain upload;
    // a pending announcement stays registered and is enqueued by this step's align
    CU(c, set->raw.ensure(std::max<size_t>(total * stride * sizeof(float), 16)));
    for (uint32_t j = 0; j < n_jobs; j++)
      if (jobs[j].n)
        CU(c, cudaMemcpyAsync(set->raw.as<float>() + off[j] * stride, jobs[j].pts, jobs[j].n * stride * sizeof(float),
                              cudaMemcpyHostToDevice, c->stream));
  }
  const float* d_t = nullptr;
  if (any_t) {
    CU(c, set->tchan.ensure(std::max<size_t>(total * sizeof(float), 16)));
    for (uint32_t j = 0; j < n_jobs; j++) {
      if (!jobs[j].n) continue;
      if (jobs[j].t)
        CU(c, cudaMemcpyAsync(set->tchan.as<float>() + off[j], jobs[j].t, jobs[j].n * sizeof(float), cudaMemcpyHostToDevice, c->stream));
      else
        CU(c, cudaMemsetAsync(set->tchan.as<float>() + off[j], 0, jobs[j].n * sizeof(float), c->stream));
    }
    d_t = set->tchan.as<float>();
  }
  FilterBatch fb;
  fb.out_map = any_t ? &set->mapS : &set->mapL;
  fb.out_icp = any_t ? &set->icpS : &set->icpL;
  fb.out_cnt = &set->cnt;
  const size_t e0 = prof_begin(c);
  int rc = run_filter_batch(c, set->raw.as<float>(), stride, n_jobs, off.data(), fps.data(), false, nullptr, fb, d_t);
  prof_end(c, 0, e0);
  if (rc != MLO_OK) return rc;
  std::vector<uint32_t> h;
  rc = filter_counts(c, fb, h);
  if (rc != MLO_OK) return rc;
  std::vector<uint32_t> sl_idx(n_jobs);
  for (uint32_t j = 0; j < n_jobs; j++) {
    auto& sl = set->slots[jobs[j].slot];
    sl.off = off[j];
    sl.n_raw = uint32_t(jobs[j].n);
    sl.n_map = h[j * CNT_STRIDE + 1];
    sl.n_icp = h[j * CNT_STRIDE + 2];
    sl_idx[j] = jobs[j].slot;
  }
  set->skewed = any_t;
  if (any_t) {
    // the final layers do not exist before mlo_scanset_deskew: report sizes only
    CU(c, set->mapL.ensure(std::max<size_t>(total, 1) * sizeof(float4)));
    CU(c, set->icpL.ensure(std::max<size_t>(total, 1) * sizeof(float4)));
    for (uint32_t j = 0; j < n_jobs; j++) {
      info[j] = mlo_scan_info{};
      info[j].n_map = set->slots[jobs[j].slot].n_map;
      info[j].n_icp = set->slots[jobs[j].slot].n_icp;
    }
    prof_collect(c);
    return MLO_OK;
  }
  rc = scanset_bbox(set, n_jobs, sl_idx.data(), info);
  prof_collect(c);
  return rc;
}

int mlo_scanset_deskew(mlo_scanset* set, uint32_t n, const uint32_t* slots, const double* twists6, mlo_scan_info* info) {
  if (!set || (n && (!slots || !twists6 || !info))) return MLO_ERR_INVALID_ARG;
  mlo_ctx* c = set->ctx;
  DeviceGuard g(c->device);
  if (!set->skewed) return fail(c, MLO_ERR_INVALID_ARG, "the set holds no skewed layers (filter ran without timestamps)");
  const size_t e0 = prof_begin(c);
  for (uint32_t i = 0; i < n; i++) {
    if (slots[i] >= set->slots.size() || !set->slots[slots[i]].valid) return fail(c, MLO_ERR_INVALID_ARG, "bad slot");
    const auto& sl = set->slots[slots[i]];
    Twist6 tw;
    std::memcpy(tw.v, twists6 + 6 * size_t(i), sizeof(tw.v));
    if (sl.n_map)
      LAUNCH(c, k_deskew, (sl.n_map + 255) / 256, 256, set->mapS.as<float4>() + sl.off, sl.n_map, tw, set->mapL.as<float4>() + sl.off);
    if (sl.n_icp)
      LAUNCH(c, k_deskew, (sl.n_icp + 255) / 256, 256, set->icpS.as<float4>() + sl.off, sl.n_icp, tw, set->icpL.as<float4>() + sl.off);
  }
  prof_end(c, 0, e0);
  set->deskewed = true;
  int rc = scanset_bbox(set, n, slots, info);
  prof_collect(c);
  return rc;
}

int mlo_scanset_align(mlo_scanset* set, uint32_t n_jobs, const mlo_align_job* jobs, mlo_icp_result* out) {
  if (!set || (n_jobs && (!jobs || !out))) return MLO_ERR_INVALID_ARG;
  mlo_ctx* c = set->ctx;
  DeviceGuard g(c->device);
  if (set->skewed && !set->deskewed) return fail(c, MLO_ERR_INVALID_ARG, "skewed layers: call mlo_scanset_deskew first");
  std::vector<uint64_t> qb(n_jobs);
  std::vector<uint32_t> nq(n_jobs);
  std::vector<const mlo_map*> maps(n_jobs);
  std::vector<double> init(12 * size_t(n_jobs));
  std::vector<mlo_icp_params> prm(n_jobs);
  for (uint32_t j = 0; j < n_jobs; j++) {
    if (jobs[j].slot >= set->slots.size() || !set->slots[jobs[j].slot].valid || !jobs[j].map)
      return fail(c, MLO_ERR_INVALID_ARG, "bad align job");
    if (jobs[j].map->ctx != c) return fail(c, MLO_ERR_INVALID_ARG, "map belongs to another context");
    const auto& sl = set->slots[jobs[j].slot];
    qb[j] = sl.off;
    nq[j] = sl.n_icp;
    maps[j] = jobs[j].map;
    std::memcpy(&init[12 * size_t(j)], jobs[j].init_pose_3x4, 12 * sizeof(double));
    prm[j] = jobs[j].params;
  }
  return align_batch_core(c, n_jobs, set->icp_layer(), qb.data(), nq.data(), maps.data(), init.data(), prm.data(), out);
}

int mlo_scanset_insert(mlo_scanset* set, uint32_t n_jobs, const mlo_insert_job* jobs, mlo_map_counts* out) {
  if (!set || (n_jobs && !jobs)) return MLO_ERR_INVALID_ARG;
  mlo_ctx* c = set->ctx;
  DeviceGuard g(c->device);
  if (set->skewed && !set->deskewed) return fail(c, MLO_ERR_INVALID_ARG, "skewed layers: call mlo_scanset_deskew first");
  if (n_jobs == 0) return MLO_OK;
  CU(c, c->h_misc.ensure(std::max<size_t>(256, size_t(n_jobs) * MAP_COUNTERS * sizeof(uint32_t))));
  uint32_t* h = c->h_misc.as<uint32_t>();
  // one InsertJobDev per (map, scan): the link, commit and cull passes of ALL jobs run as three launches
  std::vector<InsertJobDev> dj(n_jobs);
  uint64_t total = 0;
  uint32_t max_n = 0, max_cap = 0;
  bool any_cull = false;
  for (uint32_t j = 0; j < n_jobs; j++) {
    if (jobs[j].slot >= set->slots.size() || !set->slots[jobs[j].slot].valid || !jobs[j].map)
      return fail(c, MLO_ERR_INVALID_ARG, "bad insert job");
    if (jobs[j].map->ctx != c) return fail(c, MLO_ERR_INVALID_ARG, "map belongs to another context");
    for (uint32_t k = 0; k < j; k++)
      if (jobs[k].map == jobs[j].map) return fail(c, MLO_ERR_INVALID_ARG, "two insert jobs for one map in one pass");
    total += set->slots[jobs[j].slot].n_map;
  }
  CU(c, c->d_ins_g.ensure(std::max<uint64_t>(total, 1) * sizeof(float4)));
  CU(c, c->d_ins_slot.ensure(std::max<uint64_t>(total, 1) * sizeof(uint32_t)));
  CU(c, c->d_ins_next.ensure(std::max<uint64_t>(total, 1) * sizeof(int32_t)));
  CU(c, c->d_ins_jobs.ensure(n_jobs * sizeof(InsertJobDev)));
  uint64_t off = 0;
  for (uint32_t j = 0; j < n_jobs; j++) {
    mlo_map* m = jobs[j].map;
    const auto& sl = set->slots[jobs[j].slot];
    {
      int rc = map_ensure_capacity(m, sl.n_map);
      if (rc != MLO_OK) return rc;
    }
    InsertJobDev& d = dj[j];
    d.m = m->dev;
    d.src = reinterpret_cast<const float*>(set->map_layer() + sl.off);
    d.n = sl.n_map;
    std::memcpy(d.T.m, jobs[j].pose_3x4, sizeof(d.T.m));
    d.g = c->d_ins_g.as<float4>() + off;
    d.pslot = c->d_ins_slot.as<uint32_t>() + off;
    d.next = c->d_ins_next.as<int32_t>() + off;
    d.head = m->head;
    d.d = -1;
    d.sx = d.sy = d.sz = 0;
    if (jobs[j].cull_farther_than > 0.f) {
      const float inv = m->dev.inv_voxel;
      d.sx = voxel_index_map(float(jobs[j].pose_3x4[3]), inv, m->dev.index_floor);
      d.sy = voxel_index_map(float(jobs[j].pose_3x4[7]), inv, m->dev.index_floor);
      d.sz = voxel_index_map(float(jobs[j].pose_3x4[11]), inv, m->dev.index_floor);
      d.d = int32_t(std::ceil(jobs[j].cull_farther_than * inv));
      any_cull = true;  // the cull pass only has to visit voxel ids below the high-water mark after this insert
      max_cap = std::max<uint32_t>(max_cap, uint32_t(std::min<uint64_t>(m->dev.capacity_voxels, m->hwm + sl.n_map)));
    }
    off += sl.n_map;
    max_n = std::max(max_n, sl.n_map);
  }
  const size_t e0 = prof_begin(c);
  CU(c, cudaMemcpyAsync(c->d_ins_jobs.p, dj.data(), n_jobs * sizeof(InsertJobDev), cudaMemcpyHostToDevice, c->stream));
  const InsertJobDev* ddj = c->d_ins_jobs.as<InsertJobDev>();
  if (max_n) {
    const dim3 grid((max_n + 255) / 256, n_jobs);
    LAUNCH(c, k_insert_link_batch, grid, 256, ddj);
    LAUNCH(c, k_insert_commit_batch, grid, 256, ddj);
  }
  if (any_cull) LAUNCH(c, k_cull_inplace_batch, dim3((max_cap + 255) / 256, n_jobs), 256, ddj);
  CU(c, cudaGetLastError());
  // (one gather kernel + ONE device-to-host copy: a copy per map cost ~10 us each, 1.3 ms of a 128-sequence lock step)
  CU(c, c->d_misc.ensure(std::max<size_t>(256, size_t(n_jobs) * MAP_COUNTERS * sizeof(uint32_t))));
  LAUNCH(c, k_gather_counters, n_jobs, 32, ddj, c->d_misc.as<uint32_t>());
  CU(c, cudaMemcpyAsync(h, c->d_misc.p, size_t(n_jobs) * MAP_COUNTERS * sizeof(uint32_t), cudaMemcpyDeviceToHost, c->stream));
  prof_end(c, 2, e0);
  CU(c, cudaStreamSynchronize(c->stream));
  CU(c, cudaGetLastError());
  std::vector<uint32_t> hc(h, h + size_t(n_jobs) * MAP_COUNTERS);  // digest may reuse h_misc (rebuild path)
  for (uint32_t j = 0; j < n_jobs; j++) {
    int rc = digest_map_counters(jobs[j].map, &hc[size_t(j) * MAP_COUNTERS]);
    if (rc != MLO_OK) return rc;
    if (out) {
      out[j].n_voxels = jobs[j].map->n_voxels;
      out[j].n_points = jobs[j].map->n_points;
    }
  }
  prof_collect(c);
  return MLO_OK;
}

int mlo_scanset_download(mlo_scanset* set, uint32_t slot, int layer, float* out_xyz, uint64_t max_points, uint64_t* n) {
  if (!set || !n || slot >= set->slots.size() || (layer != 0 && layer != 1)) return MLO_ERR_INVALID_ARG;
  mlo_ctx* c = set->ctx;
  DeviceGuard g(c->device);
  const auto& sl = set->slots[slot];
  *n = sl.valid ? (layer == 0 ? sl.n_map : sl.n_icp) : 0;
  if (!out_xyz || *n == 0) return MLO_OK;
  if (*n > max_points) return fail(c, MLO_ERR_INVALID_ARG, "download buffer too small");
  if (set->skewed && !set->deskewed) return fail(c, MLO_ERR_INVALID_ARG, "skewed layers: call mlo_scanset_deskew first");
  return download_xyz(c, (layer == 0 ? set->map_layer() : set->icp_layer()) + sl.off, *n, out_xyz);
}

#ifdef MLO_TRACE
// scratch builds only: copy out and reset the persistent-kernel timeline of problem 0
int mlo_debug_trace_read(unsigned long long* out, unsigned int max_n, unsigned int* n) {
  unsigned int cnt = 0;
  cudaDeviceSynchronize();
  cudaMemcpyFromSymbol(&cnt, g_trace_n, sizeof(cnt));
  *n = cnt < 16384 ? cnt : 16384;
  if (*n > max_n) *n = max_n;
  if (*n) cudaMemcpyFromSymbol(out, g_trace, size_t(*n) * sizeof(unsigned long long));
  cnt = 0;
  cudaMemcpyToSymbol(g_trace_n, &cnt, sizeof(cnt));
  return MLO_OK;
}
#endif

// ------------------------------------------------------------------ profiling
int mlo_profile_enable(mlo_ctx* c, int enabled) {
  if (!c) return MLO_ERR_INVALID_ARG;
  c->prof_on = enabled != 0;
  return MLO_OK;
}
int mlo_profile_get(mlo_ctx* c, mlo_profile* out, int reset) {
  if (!c || !out) return MLO_ERR_INVALID_ARG;
  *out = c->prof;
  if (reset) c->prof = mlo_profile{};
  return MLO_OK;
}

}  // extern "C"
