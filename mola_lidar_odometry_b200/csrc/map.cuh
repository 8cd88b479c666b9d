// map.cuh — hash-voxel local map resident in HBM (replaces mola::HashedVoxelPointCloud and the
// per-voxel statistics of mola::NDT; pipelines/lidar3d-default.yaml:228-242, lidar3d-ndt.yaml:234-254).
//
// Layout in HBM (designed around the 32-byte DRAM sector and the 3x3x3 neighbourhood query)
//   buckets[n_buckets]  32-byte records {u64 key(kx, ky, kz>>2); u32 cell[4]; u64 pad}.  One bucket is a
//                       z-column of 4 consecutive voxels; cell[kz&3] = (voxel_id << 6) | count, or ABSENT.
//                       Open addressing with linear probing over buckets.  The 27 cells around a query
//                       live in 9 columns x (1 or 2) buckets = 13.5 sector reads on average (not 27), and
//                       each read returns key, payload ids and fill counts at once.
//   pts[capacity*cap]   float4 (x, y, z, 0): the <= cap points of voxel v at [v*cap, v*cap+count),
//                       64-byte aligned rows, read as one coalesced warp load per cell.
//   mean/normal[capacity] float4, NDT only: (mean xyz, is_plane) and (unit normal xyz, 0).
//   vkey[capacity]      u64 packed (kx, ky, kz) of voxel v, or KEY_EMPTY once v was culled; free_ids[capacity] is the
//                       stack of culled voxel ids that the next inserts reuse before bumping counters[0].
// n_buckets = table_factor * capacity_voxels (power of two).  Culling (insertOpts.remove_voxels_farther_than) is in
// place: one thread per voxel id tests its key, clears the cell word of an out-of-range voxel and pushes the id on
// the free stack; the emptied column buckets stay claimed (they keep probe chains intact) and are only compacted
// by a rebuild into the second buffer set once claimed columns exceed half of the table.
#pragma once
#include "common.cuh"

namespace mlo {

struct __align__(32) Bucket {
  unsigned long long key;
  uint32_t cell[4];
  unsigned long long pad;
};
static_assert(sizeof(Bucket) == 32, "bucket must be one DRAM sector");

constexpr uint32_t CELL_ABSENT = 0xFFFFFFFFu;
constexpr uint32_t CELL_PENDING = 0xFFFFFFFEu;
MLO_HD uint32_t cell_vid(uint32_t w) { return w >> 6; }
MLO_HD uint32_t cell_cnt(uint32_t w) { return w & 63u; }
MLO_HD uint32_t cell_make(uint32_t vid, uint32_t cnt) { return (vid << 6) | cnt; }
MLO_HD uint64_t column_key(int32_t kx, int32_t ky, int32_t kz) { return pack_key(kx, ky, kz >> 2); }

struct MapDev {
  Bucket* buckets;
  uint64_t mask;  // n_buckets - 1
  float4* pts;
  float4* mean;
  float4* normal;
  unsigned long long* vkey;  // per voxel id: packed (kx,ky,kz), KEY_EMPTY = culled
  uint32_t* free_ids;        // stack of reusable voxel ids (counters[3] entries)
  uint32_t* counters;  // [0] voxel ids handed out (high-water mark), [1] points stored, [2] error bits,
                       // [3] free-stack size (int), [4] column buckets claimed, [5] voxels removed by the last cull
  uint32_t cap;  // max points per voxel (logical)
  uint32_t row;  // physical row length in float4 (cap rounded up to even: rows are 32-byte aligned)
  uint32_t capacity_voxels;
  float inv_voxel;
  float voxel_size;
  float min_dist2;
  float eig_ratio;
  uint32_t min_pts_plane;
  int32_t kind;
  int32_t index_floor;  // [VERIFY] convention: voxel index = floor instead of truncation (common.cuh)
  int32_t cull_metric;  // [VERIFY] convention: metric of remove_voxels_farther_than
};

// Read-only bucket fetch: two 16-byte loads of one sector (ld.global.nc).
struct BucketRO {
  uint64_t key;
  uint32_t cell[4];
};
MLO_D BucketRO load_bucket(const MapDev& m, uint64_t b) {
  const uint4* p = reinterpret_cast<const uint4*>(&m.buckets[b]);
  const uint4 a = __ldg(p);
  const uint4 c = __ldg(p + 1);
  BucketRO r;
  r.key = (uint64_t(a.y) << 32) | uint64_t(a.x);
  r.cell[0] = a.z;
  r.cell[1] = a.w;
  r.cell[2] = c.x;
  r.cell[3] = c.y;
  return r;
}

// Find the bucket of a column key; returns false if the column does not exist.
MLO_D bool find_column(const MapDev& m, uint64_t key, BucketRO& out) {
  uint64_t h = uint64_t(hash_packed(key)) & m.mask;
  for (;;) {
    out = load_bucket(m, h);
    if (out.key == key) return true;
    if (out.key == KEY_EMPTY) return false;
    h = (h + 1) & m.mask;
  }
}

// Look up one cell: (vid, cnt) packed word or CELL_ABSENT.
MLO_D uint32_t map_find_cell(const MapDev& m, int32_t kx, int32_t ky, int32_t kz) {
  BucketRO b;
  if (!find_column(m, column_key(kx, ky, kz), b)) return CELL_ABSENT;
  return b.cell[kz & 3];
}

// Writers: find-or-claim the bucket of `key`. Returns the bucket index or ~0 on table exhaustion.
MLO_D uint64_t find_or_insert_column(const MapDev& m, uint64_t key) {
  uint64_t h = uint64_t(hash_packed(key)) & m.mask;
  for (uint64_t probes = 0; probes <= m.mask; probes++) {
    unsigned long long* kp = &m.buckets[h].key;
    unsigned long long cur = *reinterpret_cast<volatile unsigned long long*>(kp);
    if (cur == KEY_EMPTY) {
      cur = atomicCAS(kp, (unsigned long long)KEY_EMPTY, (unsigned long long)key);
      if (cur == KEY_EMPTY) atomicAdd(&m.counters[4], 1u);
    }
    if (cur == KEY_EMPTY || cur == key) return h;
    h = (h + 1) & m.mask;
  }
  atomicOr(&m.counters[2], ERR_CAPACITY);
  return ~0ull;
}

// Writers: make sure cell (bucket, sub) exists; the creator allocates the payload id. Returns the
// flat cell index bucket*4+sub (used for the per-voxel pending lists), or ~0 on error.
MLO_D uint64_t find_or_insert_cell(const MapDev& m, int32_t kx, int32_t ky, int32_t kz) {
  const uint64_t b = find_or_insert_column(m, column_key(kx, ky, kz));
  if (b == ~0ull) return ~0ull;
  const uint32_t sub = uint32_t(kz & 3);
  uint32_t* cp = &m.buckets[b].cell[sub];
  if (*reinterpret_cast<volatile uint32_t*>(cp) == CELL_ABSENT) {
    if (atomicCAS(cp, CELL_ABSENT, CELL_PENDING) == CELL_ABSENT) {
      // payload id: reuse a culled voxel's row if the free stack has one (no pushes run concurrently with inserts)
      uint32_t v;
      int* fn = reinterpret_cast<int*>(&m.counters[3]);
      int f = -1;
      if (*reinterpret_cast<volatile int*>(fn) > 0) {
        f = atomicSub(fn, 1) - 1;
        if (f < 0) atomicAdd(fn, 1);  // lost the race for the last entry: undo
      }
      if (f >= 0) {
        v = m.free_ids[f];
      } else {
        v = atomicAdd(&m.counters[0], 1u);
        if (v >= m.capacity_voxels) atomicOr(&m.counters[2], ERR_CAPACITY);
      }
      if (v < m.capacity_voxels) m.vkey[v] = pack_key(kx, ky, kz);
      atomicExch(cp, cell_make(v, 0u));
    }
  }
  return b * 4 + sub;
}

// ------------------------------------------------------------------ NDT voxel statistics
// mean, covariance (double, stored order) and symmetric 3x3 eigen by cyclic Jacobi with a fixed 12
// sweeps; plane iff n >= min_pts and l_min < ratio * l_max; normal = eigenvector of l_min.
MLO_D void jacobi3(double A[3][3], double V[3][3]) {
  for (int i = 0; i < 3; i++)
    for (int j = 0; j < 3; j++) V[i][j] = (i == j) ? 1.0 : 0.0;
  for (int sweep = 0; sweep < 12; sweep++)
    for (int p = 0; p < 2; p++)
      for (int q = p + 1; q < 3; q++) {
        const double apq = A[p][q];
        if (fabs(apq) < 1e-300) continue;
        const double theta = (A[q][q] - A[p][p]) / (2.0 * apq);
        const double t = (theta >= 0 ? 1.0 : -1.0) / (fabs(theta) + sqrt(theta * theta + 1.0));
        const double c = 1.0 / sqrt(t * t + 1.0), s = t * c;
        for (int k = 0; k < 3; k++) {
          const double akp = A[k][p], akq = A[k][q];
          A[k][p] = c * akp - s * akq;
          A[k][q] = s * akp + c * akq;
        }
        for (int k = 0; k < 3; k++) {
          const double apk = A[p][k], aqk = A[q][k];
          A[p][k] = c * apk - s * aqk;
          A[q][k] = s * apk + c * aqk;
        }
        for (int k = 0; k < 3; k++) {
          const double vkp = V[k][p], vkq = V[k][q];
          V[k][p] = c * vkp - s * vkq;
          V[k][q] = s * vkp + c * vkq;
        }
      }
}

MLO_D void voxel_stats(const MapDev& m, uint32_t vid, uint32_t n) {
  const float4* vp = m.pts + size_t(vid) * m.row;
  double mu[3] = {0, 0, 0};
  for (uint32_t j = 0; j < n; j++) {
    const float4 p = vp[j];
    mu[0] += double(p.x);
    mu[1] += double(p.y);
    mu[2] += double(p.z);
  }
  const double invn = 1.0 / double(n);
  for (int k = 0; k < 3; k++) mu[k] *= invn;
  float4 mean = make_float4(float(mu[0]), float(mu[1]), float(mu[2]), 0.f);
  float4 nrm = make_float4(0.f, 0.f, 0.f, 0.f);
  if (n >= m.min_pts_plane) {
    double C[3][3] = {{0, 0, 0}, {0, 0, 0}, {0, 0, 0}};
    for (uint32_t j = 0; j < n; j++) {
      const float4 p = vp[j];
      const double d[3] = {double(p.x) - mu[0], double(p.y) - mu[1], double(p.z) - mu[2]};
      for (int a = 0; a < 3; a++)
        for (int b = a; b < 3; b++) C[a][b] += d[a] * d[b];
    }
    const double invn1 = 1.0 / double(n - 1);
    for (int a = 0; a < 3; a++)
      for (int b = a; b < 3; b++) {
        C[a][b] *= invn1;
        C[b][a] = C[a][b];
      }
    double V[3][3];
    jacobi3(C, V);
    int imin = 0, imax = 0;
    for (int k = 1; k < 3; k++) {
      if (C[k][k] < C[imin][imin]) imin = k;
      if (C[k][k] > C[imax][imax]) imax = k;
    }
    const double lmin = C[imin][imin], lmax = C[imax][imax];
    if (lmax > 0.0 && lmin < double(m.eig_ratio) * lmax) {
      mean.w = 1.f;
      nrm = make_float4(float(V[0][imin]), float(V[1][imin]), float(V[2][imin]), 0.f);
    }
  }
  m.mean[vid] = mean;
  m.normal[vid] = nrm;
}

// ------------------------------------------------------------------ insert
// Deterministic parallel form of the sequential insertPoint loop (FilterMerge -> map insert,
// default.yaml:362-368): the points that land in one voxel are appended in ascending input index.
//   pass 1 (thread per point): g = pose*p, key, find-or-claim cell, push the point on the cell's list.
//   pass 2 (thread per point): the thread holding the smallest index of a list owns that voxel and
//           appends the pending points in index order (cap and min-distance tests as upstream).
constexpr uint32_t PSLOT_NONE = 0xFFFFFFFFu;

MLO_D void insert_link_point(const MapDev& m, const float* __restrict__ src, uint32_t stride, uint32_t i, const Pose34& T,
                             float4* __restrict__ g_out, uint32_t* __restrict__ pslot, int32_t* head,
                             int32_t* __restrict__ next) {
  const float* p = src + size_t(i) * stride;
  float gx, gy, gz;
  compose_point_f(T.m, p[0], p[1], p[2], gx, gy, gz);
  const int32_t kx = voxel_index_map(gx, m.inv_voxel, m.index_floor), ky = voxel_index_map(gy, m.inv_voxel, m.index_floor),
                kz = voxel_index_map(gz, m.inv_voxel, m.index_floor);
  g_out[i] = make_float4(gx, gy, gz, 0.f);
  if (!(key_in_range(kx) && key_in_range(ky) && key_in_range(kz))) {
    atomicOr(&m.counters[2], ERR_KEY_RANGE);
    pslot[i] = PSLOT_NONE;
    return;
  }
  const uint64_t c = find_or_insert_cell(m, kx, ky, kz);
  if (c == ~0ull) {
    pslot[i] = PSLOT_NONE;
    return;
  }
  pslot[i] = uint32_t(c);
  next[i] = atomicExch(&head[c], int32_t(i));
}

__global__ void k_insert_link(MapDev m, const float* __restrict__ src, uint32_t stride, uint32_t n, Pose34 T,
                              float4* __restrict__ g_out, uint32_t* __restrict__ pslot, int32_t* head,
                              int32_t* __restrict__ next) {
  const uint32_t i = blockIdx.x * blockDim.x + threadIdx.x;
  if (i >= n) return;
  insert_link_point(m, src, stride, i, T, g_out, pslot, head, next);
}

MLO_D void insert_commit_point(const MapDev& m, uint32_t i, const float4* __restrict__ g, const uint32_t* __restrict__ pslot,
                               int32_t* head, const int32_t* __restrict__ next) {
  const uint32_t c = pslot[i];
  if (c == PSLOT_NONE) return;
  // owner = smallest index on the list
  const int32_t first = *reinterpret_cast<volatile int32_t*>(&head[c]);
  for (int32_t j = first; j >= 0; j = next[j])
    if (uint32_t(j) < i) return;
  uint32_t* cp = &m.buckets[c >> 2].cell[c & 3u];
  const uint32_t w = *cp;
  const uint32_t vid = cell_vid(w);
  uint32_t cnt = cell_cnt(w);
  if (vid >= m.capacity_voxels) return;  // capacity error already flagged
  float4* vp = m.pts + size_t(vid) * m.row;
  const uint32_t c0 = cnt;
  int32_t last = -1;
  while (cnt < m.cap) {
    int32_t best = 0x7FFFFFFF;  // next pending index in ascending order
    for (int32_t j = first; j >= 0; j = next[j])
      if (j > last && j < best) best = j;
    if (best == 0x7FFFFFFF) break;
    last = best;
    const float4 q = g[best];
    bool ok = true;
    if (m.min_dist2 > 0.f) {
      for (uint32_t k = 0; k < cnt; k++) {
        const float4 e = vp[k];
        if (sqr_dist(e.x, e.y, e.z, q.x, q.y, q.z) < m.min_dist2) {
          ok = false;
          break;
        }
      }
    }
    if (ok) vp[cnt++] = q;
  }
  if (cnt != c0) {
    *cp = cell_make(vid, cnt);
    atomicAdd(&m.counters[1], cnt - c0);
    if (m.kind == MLO_MAP_NDT) voxel_stats(m, vid, cnt);
  }
  head[c] = -1;  // leave the scratch list heads clean for the next insert
}

__global__ void k_insert_commit(MapDev m, uint32_t n, const float4* __restrict__ g, const uint32_t* __restrict__ pslot,
                                int32_t* head, const int32_t* __restrict__ next) {
  const uint32_t i = blockIdx.x * blockDim.x + threadIdx.x;
  if (i >= n) return;
  insert_commit_point(m, i, g, pslot, head, next);
}

// Batched forms for a lock step of a fleet: job blockIdx.y = one (map, cloud, pose) triple; the three passes of all
// jobs run as three launches instead of three per map.
struct InsertJobDev {
  MapDev m;
  const float* src;  // float4 points of the map layer
  uint32_t n;
  Pose34 T;
  float4* g;         // scratch, n entries each
  uint32_t* pslot;
  int32_t* next;
  int32_t* head;     // the map's per-cell list heads
  int32_t sx, sy, sz, d;  // cull: sensor cell and distance in cells (d < 0: no cull)
};
__global__ void k_insert_link_batch(const InsertJobDev* __restrict__ jobs) {
  const InsertJobDev& j = jobs[blockIdx.y];
  const uint32_t i = blockIdx.x * blockDim.x + threadIdx.x;
  if (i >= j.n) return;
  insert_link_point(j.m, j.src, 4, i, j.T, j.g, j.pslot, j.head, j.next);
}
__global__ void k_insert_commit_batch(const InsertJobDev* __restrict__ jobs) {
  const InsertJobDev& j = jobs[blockIdx.y];
  const uint32_t i = blockIdx.x * blockDim.x + threadIdx.x;
  if (i >= j.n) return;
  insert_commit_point(j.m, i, j.g, j.pslot, j.head, j.next);
}

// ------------------------------------------------------------------ cull (filtered rebuild)
// insertOpts.remove_voxels_farther_than (default.yaml:238): keep voxels whose per-axis cell distance
// to the sensor's cell is <= ceil(dist * voxel_size_inv); survivors are re-hashed into `dst`.
// One thread per source bucket (most are empty: the table is sized for a low load factor); the thread of a live
// bucket re-inserts its cells and copies their payload rows.
__global__ void k_rebuild(MapDev src, MapDev dst, uint64_t n_buckets, int32_t sx, int32_t sy, int32_t sz, int32_t d,
                          int32_t use_filter) {
  const uint64_t i = uint64_t(blockIdx.x) * blockDim.x + threadIdx.x;
  if (i >= n_buckets) return;
  const unsigned long long key = src.buckets[i].key;
  if (key == KEY_EMPTY) return;
  const Bucket b = src.buckets[i];
  int32_t kx, ky, kzq;
  unpack_key(key, kx, ky, kzq);
  for (int sub = 0; sub < 4; sub++) {
    const uint32_t w = b.cell[sub];
    if (w == CELL_ABSENT || w == CELL_PENDING) continue;
    const int32_t kz = kzq * 4 + sub;
    if (use_filter && cull_out_of_range(kx - sx, ky - sy, kz - sz, d, src.cull_metric)) continue;
    const uint32_t cnt = cell_cnt(w), ov = cell_vid(w);
    const uint64_t c = find_or_insert_cell(dst, kx, ky, kz);
    if (c == ~0ull) continue;
    uint32_t* cp = &dst.buckets[c >> 2].cell[c & 3u];
    const uint32_t nv = cell_vid(*reinterpret_cast<volatile uint32_t*>(cp));
    if (nv >= dst.capacity_voxels) continue;
    *cp = cell_make(nv, cnt);
    atomicAdd(&dst.counters[1], cnt);
    const float4* sp = src.pts + size_t(ov) * src.row;
    float4* dp = dst.pts + size_t(nv) * dst.row;
    for (uint32_t j = 0; j < cnt; j++) dp[j] = sp[j];
    if (src.kind == MLO_MAP_NDT) {
      dst.mean[nv] = src.mean[ov];
      dst.normal[nv] = src.normal[ov];
    }
  }
}

// In-place cull: one thread per voxel id below the high-water mark.  An out-of-range voxel loses its cell word
// (the column bucket stays claimed), its id goes on the free stack and its points leave the statistics.
MLO_D void cull_voxel(const MapDev& m, uint32_t v, int32_t sx, int32_t sy, int32_t sz, int32_t d) {
  const uint32_t hwm = min(*reinterpret_cast<volatile uint32_t*>(&m.counters[0]), m.capacity_voxels);
  if (v >= hwm) return;
  const unsigned long long key = m.vkey[v];
  if (key == KEY_EMPTY) return;
  int32_t kx, ky, kz;
  unpack_key(key, kx, ky, kz);
  if (!cull_out_of_range(kx - sx, ky - sy, kz - sz, d, m.cull_metric)) return;
  const unsigned long long ck = column_key(kx, ky, kz);
  uint64_t h = uint64_t(hash_packed(ck)) & m.mask;
  for (uint64_t probes = 0; probes <= m.mask; probes++) {
    const unsigned long long cur = m.buckets[h].key;
    if (cur == ck) break;
    if (cur == KEY_EMPTY) return;  // cannot happen: the voxel's column exists
    h = (h + 1) & m.mask;
  }
  uint32_t* cp = &m.buckets[h].cell[kz & 3];
  const uint32_t w = *cp;
  if (w == CELL_ABSENT || cell_vid(w) != v) return;
  *cp = CELL_ABSENT;
  m.vkey[v] = KEY_EMPTY;
  atomicSub(&m.counters[1], cell_cnt(w));
  atomicAdd(&m.counters[5], 1u);
  const int slot = atomicAdd(reinterpret_cast<int*>(&m.counters[3]), 1);
  m.free_ids[slot] = v;
}
__global__ void k_cull_inplace(MapDev m, int32_t sx, int32_t sy, int32_t sz, int32_t d) {
  cull_voxel(m, blockIdx.x * blockDim.x + threadIdx.x, sx, sy, sz, d);
}
// the counters of every job's map into ONE contiguous buffer (a fleet step then needs one device-to-host copy, not one per map)
__global__ void k_gather_counters(const InsertJobDev* __restrict__ jobs, uint32_t* __restrict__ out) {
  if (threadIdx.x < MAP_COUNTERS) out[blockIdx.x * MAP_COUNTERS + threadIdx.x] = jobs[blockIdx.x].m.counters[threadIdx.x];
}
__global__ void k_cull_inplace_batch(const InsertJobDev* __restrict__ jobs) {
  const InsertJobDev& j = jobs[blockIdx.y];
  if (j.d < 0) return;
  cull_voxel(j.m, blockIdx.x * blockDim.x + threadIdx.x, j.sx, j.sy, j.sz, j.d);
}

// ------------------------------------------------------------------ nearest neighbour
// NearestNeighborsCapable::nn_single_search over the 27 cells key(q)+{-1,0,1}^3, visited in cx, cy, cz
// nested order, stored slot order inside a cell, strict '<' so the first minimum wins.
struct NNHit {
  float x, y, z, d2;
  uint32_t found;
  uint32_t ncand;
};

// 256-bit read-only loads (LDG.E.256 on sm_100a): one instruction per 32-byte bucket / per point pair.
struct __align__(32) F8 {
  float v[8];
};
MLO_D F8 ldg256_f(const void* p) {
  F8 r;
  asm volatile("ld.global.nc.L1::no_allocate.v8.f32 {%0,%1,%2,%3,%4,%5,%6,%7}, [%8];"
               : "=f"(r.v[0]), "=f"(r.v[1]), "=f"(r.v[2]), "=f"(r.v[3]), "=f"(r.v[4]), "=f"(r.v[5]), "=f"(r.v[6]),
                 "=f"(r.v[7])
               : "l"(p));
  return r;
}
MLO_D BucketRO load_bucket256(const MapDev& m, uint64_t b) {
  uint32_t x[8];
  asm volatile("ld.global.nc.v8.u32 {%0,%1,%2,%3,%4,%5,%6,%7}, [%8];"
               : "=r"(x[0]), "=r"(x[1]), "=r"(x[2]), "=r"(x[3]), "=r"(x[4]), "=r"(x[5]), "=r"(x[6]), "=r"(x[7])
               : "l"(&m.buckets[b]));
  BucketRO r;
  r.key = (uint64_t(x[1]) << 32) | uint64_t(x[0]);
  r.cell[0] = x[2];
  r.cell[1] = x[3];
  r.cell[2] = x[4];
  r.cell[3] = x[5];
  return r;
}

// Per-axis squared gaps from q to the neighbour cells at -1 / 0 / +1 (conservative, see below): the lower
// bound of cell (dx,dy,dz) is then g2[0][dx+1] + g2[1][dy+1] + g2[2][dz+1] — three adds per cell.
struct AxisGaps {
  float g2[3][3];
};
MLO_D AxisGaps axis_gaps(float vs, const float qv[3], const int32_t kq[3], int floor_mode = 0) {
  AxisGaps g;
#pragma unroll
  for (int a = 0; a < 3; a++) {
    g.g2[a][1] = 0.f;
    {  // dd = +1: lower face of cell kq+1
      const int32_t cc = kq[a] + 1;
      const float edge = floor_mode ? float(cc) * vs : (cc > 0 ? float(cc) * vs : (cc == 0 ? -vs : float(cc - 1) * vs));
      float gap = edge - qv[a];
      gap -= 4e-6f * (fabsf(edge) + vs);
      g.g2[a][2] = gap > 0.f ? gap * gap : 0.f;
    }
    {  // dd = -1: upper face of cell kq-1
      const int32_t cc = kq[a] - 1;
      const float edge = floor_mode ? float(cc + 1) * vs : (cc > 0 ? float(cc + 1) * vs : (cc == 0 ? vs : float(cc) * vs));
      float gap = qv[a] - edge;
      gap -= 4e-6f * (fabsf(edge) + vs);
      g.g2[a][0] = gap > 0.f ? gap * gap : 0.f;
    }
  }
  return g;
}
#define MLO_LB2(g, e) (((g).g2[0][(e) / 9] + (g).g2[1][((e) / 3) % 3] + (g).g2[2][(e) % 3]) * 0.99999f)

// Squared distance from q to the box of neighbour cell (kq + dd), a conservative lower bound for every
// point stored in that cell.  Cell boxes follow the truncation-toward-zero index: cell 0 spans (-vs, vs),
// cell c > 0 spans [c vs, (c+1) vs), cell c < 0 spans ((c-1) vs, c vs]; faces are pulled in by a few ulps
// so a coordinate that rounds across a face is never excluded.
MLO_D float cell_lower_bound2(float vs, const float qv[3], const int32_t kq[3], const int32_t dd[3], int floor_mode = 0) {
  float lb2 = 0.f;
#pragma unroll
  for (int a = 0; a < 3; a++) {
    if (dd[a] == 0) continue;
    const int32_t cc = kq[a] + dd[a];
    float edge, gap;
    if (dd[a] > 0) {  // lower face of cell cc
      edge = floor_mode ? float(cc) * vs : (cc > 0 ? float(cc) * vs : (cc == 0 ? -vs : float(cc - 1) * vs));
      gap = edge - qv[a];
    } else {  // upper face of cell cc
      edge = floor_mode ? float(cc + 1) * vs : (cc > 0 ? float(cc + 1) * vs : (cc == 0 ? vs : float(cc) * vs));
      gap = qv[a] - edge;
    }
    gap -= 4e-6f * (fabsf(edge) + vs);
    if (gap > 0.f) lb2 += gap * gap;
  }
  return lb2 * 0.99999f;
}

// Probe the 3x3x3 neighbourhood of cell kq (thread per query): 9 columns x (1..2) buckets as 256-bit loads,
// issued three columns at a time; writes the 27 packed cell words (canonical order) to ws[e * wstride] and
// returns the number of points stored in those cells (the algorithmic candidate count).
MLO_D uint32_t probe_words(const MapDev& m, const int32_t kq[3], uint32_t* ws, uint32_t wstride) {
  const int32_t kz = kq[2];
  const int32_t zq0 = (kz - 1) >> 2, zq1 = (kz + 1) >> 2;
  const bool needB = zq1 != zq0;
  uint32_t ncand = 0;
  // partial hash products shared by the 9 columns
  uint32_t hy[3];
#pragma unroll
  for (int cy = 0; cy < 3; cy++) hy[cy] = uint32_t(kq[1] + cy - 1) * HASH_PY;
  const uint32_t hz0 = uint32_t(zq0) * HASH_PZ, hz1 = uint32_t(zq1) * HASH_PZ;
#pragma unroll
  for (int cx = 0; cx < 3; cx++) {
    BucketRO a[3], b[3];
    uint64_t ka[3], kb[3], ha[3], hb[3];
    const uint32_t hx = uint32_t(kq[0] + cx - 1) * HASH_PX;
#pragma unroll
    for (int cy = 0; cy < 3; cy++) {
      ka[cy] = pack_key(kq[0] + cx - 1, kq[1] + cy - 1, zq0);
      ha[cy] = uint64_t(hash_mix(hx ^ hy[cy] ^ hz0)) & m.mask;
      a[cy] = load_bucket256(m, ha[cy]);
    }
    if (needB) {
#pragma unroll
      for (int cy = 0; cy < 3; cy++) {
        kb[cy] = pack_key(kq[0] + cx - 1, kq[1] + cy - 1, zq1);
        hb[cy] = uint64_t(hash_mix(hx ^ hy[cy] ^ hz1)) & m.mask;
        b[cy] = load_bucket256(m, hb[cy]);
      }
    }
#pragma unroll
    for (int cy = 0; cy < 3; cy++) {
      while (a[cy].key != ka[cy] && a[cy].key != KEY_EMPTY) {
        ha[cy] = (ha[cy] + 1) & m.mask;
        a[cy] = load_bucket256(m, ha[cy]);
      }
      if (needB) {
        while (b[cy].key != kb[cy] && b[cy].key != KEY_EMPTY) {
          hb[cy] = (hb[cy] + 1) & m.mask;
          b[cy] = load_bucket256(m, hb[cy]);
        }
      }
#pragma unroll
      for (int t = 0; t < 3; t++) {
        const int32_t z = kz - 1 + t;
        const bool inA = (z >> 2) == zq0;
        const bool have = inA ? (a[cy].key == ka[cy]) : (needB && b[cy].key == kb[cy]);
        const uint32_t s = uint32_t(z & 3);
        const uint32_t wa = s == 0 ? a[cy].cell[0] : s == 1 ? a[cy].cell[1] : s == 2 ? a[cy].cell[2] : a[cy].cell[3];
        uint32_t wb = CELL_ABSENT;
        if (needB) wb = s == 0 ? b[cy].cell[0] : s == 1 ? b[cy].cell[1] : s == 2 ? b[cy].cell[2] : b[cy].cell[3];
        uint32_t ww = have ? (inA ? wa : wb) : CELL_ABSENT;
        if (ww == CELL_PENDING) ww = CELL_ABSENT;
        ws[((cx * 3 + cy) * 3 + t) * wstride] = ww;
        if (ww != CELL_ABSENT) ncand += cell_cnt(ww);
      }
    }
  }
  return ncand;
}

// Thread-per-query form: the high-MLP path of the fused ICP kernel for large batches (hundreds of
// queries in flight per SM).  Probes: 9 columns x (1..2) buckets as 256-bit loads, issued three columns
// at a time.  Exact pruning: the query's own cell is scanned first, a neighbour cell is read only if its
// box can still hold a point at distance <= the running best; candidates are compared on
// (d2, canonical order) so the result equals the sequential first-minimum scan bit for bit.
// `ws` points at this thread's column of a shared-memory scratch [27][wstride] holding the 27 packed cell
// words (dynamic indexing without local memory; unaffected by the L1 invalidation of device-scope fences).
// Second half of nn_single_thread: the pruned scan over the 27 packed cell words already in ws[e * wstride].
MLO_D NNHit nn_scan_words(const MapDev& m, float qx, float qy, float qz, const int32_t kq[3], const uint32_t* ws, uint32_t wstride) {
  NNHit r;
  r.x = r.y = r.z = 0.f;
  r.d2 = __int_as_float(0x7f800000);
  r.found = 0;
  r.ncand = 0;
  uint32_t border = 0xFFFFFFFFu;
  const float qv[3] = {qx, qy, qz};
  auto scan_cell = [&](uint32_t ww, uint32_t e) {
    const float4* row = m.pts + size_t(cell_vid(ww)) * m.row;
    const uint32_t c = cell_cnt(ww);
    for (uint32_t j = 0; j < c; j += 8) {
      F8 v[4];
#pragma unroll
      for (int u = 0; u < 4; u++)
        if (j + 2 * u < c) v[u] = ldg256_f(row + j + 2 * u);
#pragma unroll
      for (int u = 0; u < 4; u++) {
#pragma unroll
        for (int hlf = 0; hlf < 2; hlf++) {
          const uint32_t jj = j + 2 * u + hlf;
          if (jj < c) {
            const float px = v[u].v[4 * hlf], py = v[u].v[4 * hlf + 1], pz = v[u].v[4 * hlf + 2];
            const float d2 = sqr_dist(px, py, pz, qx, qy, qz);
            const uint32_t ord = e * 32u + jj;
            if (d2 < r.d2 || (d2 == r.d2 && ord < border)) {
              r.d2 = d2;
              r.x = px;
              r.y = py;
              r.z = pz;
              border = ord;
            }
          }
        }
      }
    }
  };
  {
    const uint32_t wh = ws[13 * wstride];
    if (wh != CELL_ABSENT) scan_cell(wh, 13u);
  }
  // Each lane walks ITS OWN compacted list of candidate cells (bit e of `todo`), so one warp iteration
  // serves every lane's k-th visited cell: the trip count is max-over-lanes of cells visited, not the
  // union of cells any lane visits.
  const AxisGaps gaps = axis_gaps(m.voxel_size, qv, kq, m.index_floor);
  uint32_t todo = 0;
#pragma unroll
  for (int e = 0; e < 27; e++) {
    if (e == 13) continue;
    const uint32_t we = ws[e * wstride];
    if (we == CELL_ABSENT || cell_cnt(we) == 0) continue;
    if (MLO_LB2(gaps, e) <= r.d2) todo |= 1u << e;
  }
  while (todo) {
    const int e = __ffs(todo) - 1;
    todo &= todo - 1;
    scan_cell(ws[e * wstride], uint32_t(e));
  }
  r.found = border != 0xFFFFFFFFu;
  return r;
}

MLO_D NNHit nn_single_thread(const MapDev& m, float qx, float qy, float qz, uint32_t* ws, uint32_t wstride) {
  const int32_t kq[3] = {voxel_index_map(qx, m.inv_voxel, m.index_floor), voxel_index_map(qy, m.inv_voxel, m.index_floor),
                         voxel_index_map(qz, m.inv_voxel, m.index_floor)};
  if (!(key_in_range(kq[0]) && key_in_range(kq[1]) && key_in_range(kq[2]))) {
    NNHit r;
    r.x = r.y = r.z = 0.f;
    r.d2 = __int_as_float(0x7f800000);
    r.found = 0;
    r.ncand = 0;
    return r;
  }
  const uint32_t ncand = probe_words(m, kq, ws, wstride);
  NNHit r = nn_scan_words(m, qx, qy, qz, kq, ws, wstride);
  r.ncand = ncand;
  return r;
}

// ---- warp-per-query form: the hot loop of the fused ICP kernel.
// Phase 1 (probe): lane = column*2 + half, 18 lanes fetch the one or two buckets of each of the 9
//   (dx,dy) columns: one 32-byte sector per lane, all in flight at once.
// Phase 2 (gather): the 27 packed cell words are redistributed so lane e owns cell e in canonical
//   order; occupied cells are streamed four at a time, each as ONE coalesced warp load of <= cap float4
//   (lane j reads slot j).  Per-lane running minima see candidates in canonical order; the final warp
//   arg-min breaks ties by (cell, slot) order, reproducing the sequential first-minimum rule bit for bit.
struct WarpProbe {
  BucketRO b;       // this lane's bucket (valid for lanes < 18 that probe)
  uint64_t key;     // the column key this lane looks for
  uint64_t h;       // current bucket index
  int32_t kz;       // query cell z
  bool active;      // lane takes part in the probe
  bool in_range;
};

MLO_D WarpProbe warp_probe_issue(const MapDev& m, float qx, float qy, float qz) {
  const uint32_t lane = threadIdx.x & 31u;
  WarpProbe p;
  const int32_t kx = voxel_index_map(qx, m.inv_voxel, m.index_floor), ky = voxel_index_map(qy, m.inv_voxel, m.index_floor),
                kz = voxel_index_map(qz, m.inv_voxel, m.index_floor);
  p.kz = kz;
  p.in_range = key_in_range(kx) && key_in_range(ky) && key_in_range(kz);
  const uint32_t col = lane >> 1, half = lane & 1u;
  const int32_t dx = int32_t(col / 3) - 1, dy = int32_t(col % 3) - 1;
  const int32_t zq0 = (kz - 1) >> 2, zq1 = (kz + 1) >> 2;
  p.active = p.in_range && lane < 18 && (half == 0 || zq1 != zq0);
  p.key = pack_key(kx + dx, ky + dy, half ? zq1 : zq0);
  p.h = uint64_t(hash_packed(p.key)) & m.mask;
  p.b.key = KEY_EMPTY;
  p.b.cell[0] = p.b.cell[1] = p.b.cell[2] = p.b.cell[3] = CELL_ABSENT;
  if (p.active) p.b = load_bucket(m, p.h);
  return p;
}

MLO_D NNHit warp_nn_finish(const MapDev& m, WarpProbe& p, float qx, float qy, float qz) {
  const uint32_t lane = threadIdx.x & 31u;
  const uint32_t FULL = 0xFFFFFFFFu;
  // resolve hash collisions (linear probing), warp-uniform loop
  bool pending = p.active && p.b.key != p.key && p.b.key != KEY_EMPTY;
  while (__any_sync(FULL, pending)) {
    if (pending) {
      p.h = (p.h + 1) & m.mask;
      p.b = load_bucket(m, p.h);
      pending = p.b.key != p.key && p.b.key != KEY_EMPTY;
    }
  }
  const bool have = p.active && p.b.key == p.key;
  // per probing lane: the packed words of the column's cells dz = -1, 0, +1 that live in ITS bucket
  uint32_t w3[3];
  const int32_t myzq = (lane & 1u) ? ((p.kz + 1) >> 2) : ((p.kz - 1) >> 2);
#pragma unroll
  for (int t = 0; t < 3; t++) {
    const int32_t z = p.kz - 1 + t;
    uint32_t w = CELL_ABSENT;
    if (have && (z >> 2) == myzq) {
      const uint32_t s = uint32_t(z & 3);
      w = s == 0 ? p.b.cell[0] : s == 1 ? p.b.cell[1] : s == 2 ? p.b.cell[2] : p.b.cell[3];
    }
    w3[t] = w;
  }
  // lane e (< 27) owns cell e = col*3 + t in canonical (dx, dy, dz) order
  const uint32_t e_col = lane / 3, e_t = lane % 3;
  const int32_t ez = p.kz - 1 + int32_t(e_t);
  const uint32_t src = (lane < 27) ? (e_col * 2 + (((ez >> 2) != ((p.kz - 1) >> 2)) ? 1u : 0u)) : 0u;
  const uint32_t v0 = __shfl_sync(FULL, w3[0], src), v1 = __shfl_sync(FULL, w3[1], src), v2 = __shfl_sync(FULL, w3[2], src);
  uint32_t mine = e_t == 0 ? v0 : e_t == 1 ? v1 : v2;
  if (lane >= 27 || mine == CELL_PENDING) mine = CELL_ABSENT;
  const uint32_t my_cnt = (mine == CELL_ABSENT) ? 0u : cell_cnt(mine);
  // algorithmic candidate count = all points of the 27 cells (what a plain scan would visit)
  uint32_t ncand = my_cnt;
#pragma unroll
  for (int o = 16; o > 0; o >>= 1) ncand += __shfl_xor_sync(FULL, ncand, o);

  float best = __int_as_float(0x7f800000);
  float bx = 0.f, by = 0.f, bz = 0.f;
  uint32_t border = 0xFFFFFFFFu;  // canonical order of this lane's best = cell*32 + slot

  // Exact pruning: scan the query's own cell first (e = 13); a neighbour cell is then read only if
  // the distance from q to its box can still be <= that best.  The cell boxes follow the
  // truncation-toward-zero index (cell 0 spans (-vs, vs)); bounds are shrunk by a few ulps so that a
  // point that rounds across a face is never missed.  Result (incl. tie order) equals the full scan.
  const uint32_t w_home = __shfl_sync(FULL, mine, 13);
  float bound = __int_as_float(0x7f800000);
  if (w_home != CELL_ABSENT) {
    const uint32_t c = cell_cnt(w_home);
    if (lane < c) {
      const float4 q = __ldg(m.pts + size_t(cell_vid(w_home)) * m.row + lane);
      best = sqr_dist(q.x, q.y, q.z, qx, qy, qz);
      bx = q.x;
      by = q.y;
      bz = q.z;
      border = 13u * 32u + lane;
    }
    bound = best;
#pragma unroll
    for (int o = 16; o > 0; o >>= 1) bound = fminf(bound, __shfl_xor_sync(FULL, bound, o));
  }
  bool visit = my_cnt > 0 && lane != 13;
  if (visit && bound < __int_as_float(0x7f800000)) {
    const int32_t kq[3] = {voxel_index_map(qx, m.inv_voxel, m.index_floor), voxel_index_map(qy, m.inv_voxel, m.index_floor), p.kz};
    const float qv[3] = {qx, qy, qz};
    const int32_t dd[3] = {int32_t(lane / 9) - 1, int32_t((lane / 3) % 3) - 1, int32_t(lane % 3) - 1};
    visit = cell_lower_bound2(m.voxel_size, qv, kq, dd, m.index_floor) <= bound;  // one cell per lane: the direct form
  }
  uint32_t occ = __ballot_sync(FULL, visit);
  while (occ) {
    // up to four cells per round: loads issued together, consumed in canonical order
    uint32_t e[4], c[4];
    float4 q[4];
#pragma unroll
    for (int k = 0; k < 4; k++) {
      if (occ) {
        e[k] = __ffs(occ) - 1;
        occ &= occ - 1;
        const uint32_t w = __shfl_sync(FULL, mine, e[k]);
        c[k] = cell_cnt(w);
        if (lane < c[k]) q[k] = __ldg(m.pts + size_t(cell_vid(w)) * m.row + lane);
      } else {
        c[k] = 0;
        e[k] = 0;
      }
    }
#pragma unroll
    for (int k = 0; k < 4; k++) {
      if (lane < c[k]) {
        const float d2 = sqr_dist(q[k].x, q[k].y, q[k].z, qx, qy, qz);
        if (d2 < best) {
          best = d2;
          bx = q[k].x;
          by = q[k].y;
          bz = q[k].z;
          border = e[k] * 32u + lane;
        }
      }
    }
  }
  // warp arg-min on (d2, canonical order)
#pragma unroll
  for (int o = 16; o > 0; o >>= 1) {
    const float od = __shfl_xor_sync(FULL, best, o);
    const uint32_t oo = __shfl_xor_sync(FULL, border, o);
    const float ox = __shfl_xor_sync(FULL, bx, o), oy = __shfl_xor_sync(FULL, by, o), oz = __shfl_xor_sync(FULL, bz, o);
    if (od < best || (od == best && oo < border)) {
      best = od;
      border = oo;
      bx = ox;
      by = oy;
      bz = oz;
    }
  }
  NNHit r;
  r.x = bx;
  r.y = by;
  r.z = bz;
  r.d2 = best;
  r.found = border != 0xFFFFFFFFu;
  r.ncand = ncand;  // identical on every lane (counts are warp-uniform)
  return r;
}

// mola::NDT nearest-plane query: among the 27 cells, the planar voxel with the smallest |n.(q - mean)|.
struct PlaneHit {
  float cx, cy, cz, nx, ny, nz, dist;
  uint32_t found;
  uint32_t ncand;
};

MLO_D PlaneHit nn_plane_thread(const MapDev& m, float qx, float qy, float qz) {
  PlaneHit r;
  r.cx = r.cy = r.cz = r.nx = r.ny = r.nz = 0.f;
  r.dist = __int_as_float(0x7f800000);
  r.found = 0;
  r.ncand = 0;
  const int32_t kx = voxel_index_map(qx, m.inv_voxel, m.index_floor), ky = voxel_index_map(qy, m.inv_voxel, m.index_floor),
                kz = voxel_index_map(qz, m.inv_voxel, m.index_floor);
  if (!(key_in_range(kx) && key_in_range(ky) && key_in_range(kz))) return r;
#pragma unroll 1
  for (int dx = -1; dx <= 1; dx++)
#pragma unroll 1
    for (int dy = -1; dy <= 1; dy++)
#pragma unroll 1
      for (int dz = -1; dz <= 1; dz++) {
        const uint32_t w = map_find_cell(m, kx + dx, ky + dy, kz + dz);
        if (w == CELL_ABSENT || w == CELL_PENDING) continue;
        const uint32_t v = cell_vid(w);
        r.ncand += 2;
        const float4 mu = __ldg(&m.mean[v]);
        if (mu.w == 0.f) continue;
        const float4 nr = __ldg(&m.normal[v]);
        const float ex = qx - mu.x, ey = qy - mu.y, ez = qz - mu.z;
        const float d = fabsf(nr.x * ex + nr.y * ey + nr.z * ez);
        if (d < r.dist) {
          r.dist = d;
          r.cx = mu.x; r.cy = mu.y; r.cz = mu.z;
          r.nx = nr.x; r.ny = nr.y; r.nz = nr.z;
          r.found = 1;
        }
      }
  return r;
}

// Plain query API: one warp per query through the same probe/gather path as the ICP kernel.
__global__ void k_nn_single(MapDev m, const float* __restrict__ q, uint32_t stride, uint32_t n, float* __restrict__ out_xyz,
                            float* __restrict__ out_d2, uint8_t* __restrict__ out_found) {
  const uint32_t i = (blockIdx.x * blockDim.x + threadIdx.x) >> 5;
  if (i >= n) return;
  const float* p = q + size_t(i) * stride;
  const float qx = __ldg(p), qy = __ldg(p + 1), qz = __ldg(p + 2);
  WarpProbe pr = warp_probe_issue(m, qx, qy, qz);
  const NNHit r = warp_nn_finish(m, pr, qx, qy, qz);
  if ((threadIdx.x & 31u) == 0) {
    out_xyz[3 * size_t(i)] = r.x;
    out_xyz[3 * size_t(i) + 1] = r.y;
    out_xyz[3 * size_t(i) + 2] = r.z;
    out_d2[i] = r.d2;
    out_found[i] = uint8_t(r.found);
  }
}

__global__ void k_nn_plane(MapDev m, const float* __restrict__ q, uint32_t stride, uint32_t n, float* __restrict__ out_mean,
                           float* __restrict__ out_normal, float* __restrict__ out_dist, uint8_t* __restrict__ out_found) {
  const uint32_t i = blockIdx.x * blockDim.x + threadIdx.x;
  if (i >= n) return;
  const float* p = q + size_t(i) * stride;
  const PlaneHit h = nn_plane_thread(m, __ldg(p), __ldg(p + 1), __ldg(p + 2));
  out_mean[3 * size_t(i)] = h.cx;
  out_mean[3 * size_t(i) + 1] = h.cy;
  out_mean[3 * size_t(i) + 2] = h.cz;
  out_normal[3 * size_t(i)] = h.nx;
  out_normal[3 * size_t(i) + 1] = h.ny;
  out_normal[3 * size_t(i) + 2] = h.nz;
  out_dist[i] = h.dist;
  out_found[i] = uint8_t(h.found);
}

// ------------------------------------------------------------------ export
// writes (kx, ky, kz, count, vid) of every live cell, unordered; the host sorts by key.
__global__ void k_export_list(MapDev m, uint64_t n_buckets, uint32_t* cursor, int32_t* keys3, uint32_t* counts,
                              uint32_t* vids) {
  const uint64_t i = uint64_t(blockIdx.x) * blockDim.x + threadIdx.x;
  if (i >= n_buckets * 4) return;
  const Bucket& b = m.buckets[i >> 2];
  if (b.key == KEY_EMPTY) return;
  const uint32_t w = b.cell[i & 3u];
  if (w == CELL_ABSENT || w == CELL_PENDING) return;
  int32_t kx, ky, kzq;
  unpack_key(b.key, kx, ky, kzq);
  const uint32_t o = atomicAdd(cursor, 1u);
  keys3[3 * size_t(o)] = kx;
  keys3[3 * size_t(o) + 1] = ky;
  keys3[3 * size_t(o) + 2] = kzq * 4 + int32_t(i & 3u);
  counts[o] = cell_cnt(w);
  vids[o] = cell_vid(w);
}

}  // namespace mlo
