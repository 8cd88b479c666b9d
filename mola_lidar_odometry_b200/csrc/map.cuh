// map.cuh — hash-voxel local map resident in HBM (replaces mola::HashedVoxelPointCloud and the
// per-voxel statistics of mola::NDT; pipelines/lidar3d-default.yaml:228-242, lidar3d-ndt.yaml:234-254).
//
// Layout in HBM
//   slots[table_size]  uint4 {key_lo, key_hi, voxel_id, count}: open addressing, linear probing,
//                      packed 3x21-bit key; one 16-byte probe yields key, payload id and fill count.
//   pts[capacity*cap]  float4 (x, y, z, 0): the <= cap points of voxel v at [v*cap, v*cap+count).
//   mean/normal[capacity] float4, NDT only: (mean xyz, is_plane) and (unit normal xyz, 0).
// table_size = next power of two >= 4 * capacity_voxels (load factor <= 0.25, so a probe for an
// absent neighbour cell terminates after ~1.2 slots on average).
// There are no tombstones: culling is a filtered rebuild into a second set of buffers.
#pragma once
#include "common.cuh"

namespace mlo {

struct MapDev {
  uint4* slots;
  uint64_t mask;
  float4* pts;
  float4* mean;
  float4* normal;
  uint32_t* counters;  // [0] voxels allocated, [1] points stored, [2] error bits
  uint32_t cap;
  uint32_t capacity_voxels;
  float inv_voxel;
  float min_dist2;
  float eig_ratio;
  uint32_t min_pts_plane;
  int32_t kind;
};


MLO_D uint64_t slot_key(const uint4& s) { return (uint64_t(s.y) << 32) | uint64_t(s.x); }

// Look up one cell. Returns true and (vid,count) if present. Read-only path (ld.global.nc).
MLO_D bool map_find(const MapDev& m, uint64_t key, uint32_t& vid, uint32_t& cnt) {
  uint64_t h = hash_key(key) & m.mask;
  for (;;) {
    const uint4 s = __ldg(&m.slots[h]);
    const uint64_t k = slot_key(s);
    if (k == key) {
      vid = s.z;
      cnt = s.w;
      return true;
    }
    if (k == KEY_EMPTY) return false;
    h = (h + 1) & m.mask;
  }
}

// Find-or-claim the slot of `key` (writers). Returns the slot index, or ~0 on table exhaustion.
MLO_D uint64_t map_find_or_insert(const MapDev& m, uint64_t key) {
  uint64_t h = hash_key(key) & m.mask;
  for (uint64_t probes = 0; probes <= m.mask; probes++) {
    unsigned long long* kp = reinterpret_cast<unsigned long long*>(&m.slots[h]);
    unsigned long long cur = *reinterpret_cast<volatile unsigned long long*>(kp);
    if (cur == KEY_EMPTY) {
      cur = atomicCAS(kp, (unsigned long long)KEY_EMPTY, (unsigned long long)key);
      if (cur == KEY_EMPTY) {
        // we created the voxel: allocate its payload id, count starts at 0
        const uint32_t v = atomicAdd(&m.counters[0], 1u);
        if (v >= m.capacity_voxels) atomicOr(&m.counters[2], ERR_CAPACITY);
        m.slots[h].z = v;
        m.slots[h].w = 0u;
        return h;
      }
    }
    if (cur == key) return h;
    h = (h + 1) & m.mask;
  }
  atomicOr(&m.counters[2], ERR_CAPACITY);
  return ~0ull;
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
  const float4* vp = m.pts + size_t(vid) * m.cap;
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
//   pass 1 (thread per point): g = pose*p, key, find-or-claim slot, push the point on the slot's list.
//   pass 2 (thread per point): the thread holding the smallest index of a list owns that voxel and
//           appends the pending points in index order (cap and min-distance tests as upstream).
__global__ void k_insert_link(MapDev m, const float* __restrict__ src, uint32_t stride, uint32_t n, Pose34 T,
                              float4* __restrict__ g_out, uint32_t* __restrict__ pslot, int32_t* head,
                              int32_t* __restrict__ next) {
  const uint32_t i = blockIdx.x * blockDim.x + threadIdx.x;
  if (i >= n) return;
  const float* p = src + size_t(i) * stride;
  float gx, gy, gz;
  compose_point_f(T.m, p[0], p[1], p[2], gx, gy, gz);
  const int32_t kx = voxel_index_map(gx, m.inv_voxel), ky = voxel_index_map(gy, m.inv_voxel),
                kz = voxel_index_map(gz, m.inv_voxel);
  g_out[i] = make_float4(gx, gy, gz, 0.f);
  if (!(key_in_range(kx) && key_in_range(ky) && key_in_range(kz))) {
    atomicOr(&m.counters[2], ERR_KEY_RANGE);
    pslot[i] = 0xFFFFFFFFu;
    return;
  }
  const uint64_t h = map_find_or_insert(m, pack_key(kx, ky, kz));
  if (h == ~0ull) {
    pslot[i] = 0xFFFFFFFFu;
    return;
  }
  pslot[i] = uint32_t(h);
  next[i] = atomicExch(&head[h], int32_t(i));
}

__global__ void k_insert_commit(MapDev m, uint32_t n, const float4* __restrict__ g, const uint32_t* __restrict__ pslot,
                                int32_t* head, const int32_t* __restrict__ next) {
  const uint32_t i = blockIdx.x * blockDim.x + threadIdx.x;
  if (i >= n) return;
  const uint32_t h = pslot[i];
  if (h == 0xFFFFFFFFu) return;
  // owner = smallest index on the list
  const int32_t first = *reinterpret_cast<volatile int32_t*>(&head[h]);
  for (int32_t j = first; j >= 0; j = next[j])
    if (uint32_t(j) < i) return;
  uint4 s = m.slots[h];
  const uint32_t vid = s.z;
  uint32_t c = s.w;
  if (vid >= m.capacity_voxels) return;  // capacity error already flagged
  float4* vp = m.pts + size_t(vid) * m.cap;
  const uint32_t c0 = c;
  int32_t last = -1;
  while (c < m.cap) {
    // next pending index in ascending order
    int32_t best = 0x7FFFFFFF;
    for (int32_t j = first; j >= 0; j = next[j])
      if (j > last && j < best) best = j;
    if (best == 0x7FFFFFFF) break;
    last = best;
    const float4 q = g[best];
    bool ok = true;
    if (m.min_dist2 > 0.f) {
      for (uint32_t k = 0; k < c; k++) {
        const float4 e = vp[k];
        if (sqr_dist(e.x, e.y, e.z, q.x, q.y, q.z) < m.min_dist2) {
          ok = false;
          break;
        }
      }
    }
    if (ok) vp[c++] = q;
  }
  if (c != c0) {
    m.slots[h].w = c;
    atomicAdd(&m.counters[1], c - c0);
    if (m.kind == MLO_MAP_NDT) voxel_stats(m, vid, c);
  }
  head[h] = -1;  // leave the scratch list heads clean for the next insert
}

// ------------------------------------------------------------------ cull (filtered rebuild)
// insertOpts.remove_voxels_farther_than (default.yaml:238): keep voxels whose per-axis cell distance
// to the sensor's cell is <= ceil(dist * voxel_size_inv); survivors are re-hashed into `dst`.
__global__ void k_rebuild(MapDev src, MapDev dst, uint64_t n_slots, int32_t sx, int32_t sy, int32_t sz, int32_t d,
                          int32_t use_filter) {
  // one warp per slot group: lanes cooperate on the payload copy
  const uint64_t warp = (uint64_t(blockIdx.x) * blockDim.x + threadIdx.x) >> 5;
  const uint32_t lane = threadIdx.x & 31u;
  if (warp >= n_slots) return;
  const uint4 s = src.slots[warp];
  const uint64_t key = slot_key(s);
  if (key == KEY_EMPTY) return;
  if (use_filter) {
    int32_t kx, ky, kz;
    unpack_key(key, kx, ky, kz);
    if (abs(kx - sx) > d || abs(ky - sy) > d || abs(kz - sz) > d) return;
  }
  uint32_t nv = 0;
  uint64_t h = 0;
  if (lane == 0) {
    h = map_find_or_insert(dst, key);
    nv = dst.slots[h].z;
    dst.slots[h].w = s.w;
    atomicAdd(&dst.counters[1], s.w);
  }
  nv = __shfl_sync(0xFFFFFFFFu, nv, 0);
  if (nv >= dst.capacity_voxels) return;
  if (lane < s.w) dst.pts[size_t(nv) * dst.cap + lane] = src.pts[size_t(s.z) * src.cap + lane];
  if (src.kind == MLO_MAP_NDT && lane == 0) {
    dst.mean[nv] = src.mean[s.z];
    dst.normal[nv] = src.normal[s.z];
  }
}

// ------------------------------------------------------------------ nearest neighbour, thread per query
// NearestNeighborsCapable::nn_single_search over the 27 cells key(q)+{-1,0,1}^3, visited in cx, cy, cz
// nested order, stored slot order inside a cell, strict '<' so the first minimum wins.
struct NNHit {
  float x, y, z, d2;
  uint32_t found;
  uint32_t ncand;
};

MLO_D NNHit nn_single_thread(const MapDev& m, float qx, float qy, float qz) {
  NNHit r;
  r.x = r.y = r.z = 0.f;
  r.d2 = __int_as_float(0x7f800000);
  r.found = 0;
  r.ncand = 0;
  const int32_t kx = voxel_index_map(qx, m.inv_voxel), ky = voxel_index_map(qy, m.inv_voxel),
                kz = voxel_index_map(qz, m.inv_voxel);
  if (!(key_in_range(kx) && key_in_range(ky) && key_in_range(kz))) return r;
#pragma unroll 1
  for (int dx = -1; dx <= 1; dx++) {
#pragma unroll 1
    for (int dy = -1; dy <= 1; dy++) {
      uint32_t vid[3], cnt[3];
#pragma unroll
      for (int dz = -1; dz <= 1; dz++) {
        cnt[dz + 1] = 0;
        vid[dz + 1] = 0;
        uint32_t v, c;
        if (map_find(m, pack_key(kx + dx, ky + dy, kz + dz), v, c)) {
          vid[dz + 1] = v;
          cnt[dz + 1] = c;
        }
      }
#pragma unroll
      for (int t = 0; t < 3; t++) {
        const float4* vp = m.pts + size_t(vid[t]) * m.cap;
        const uint32_t c = cnt[t];
        r.ncand += c;
        for (uint32_t j = 0; j < c; j++) {
          const float4 p = __ldg(&vp[j]);
          const float d2 = sqr_dist(p.x, p.y, p.z, qx, qy, qz);
          if (d2 < r.d2) {
            r.d2 = d2;
            r.x = p.x;
            r.y = p.y;
            r.z = p.z;
            r.found = 1;
          }
        }
      }
    }
  }
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
  const int32_t kx = voxel_index_map(qx, m.inv_voxel), ky = voxel_index_map(qy, m.inv_voxel),
                kz = voxel_index_map(qz, m.inv_voxel);
  if (!(key_in_range(kx) && key_in_range(ky) && key_in_range(kz))) return r;
#pragma unroll 1
  for (int dx = -1; dx <= 1; dx++)
#pragma unroll 1
    for (int dy = -1; dy <= 1; dy++)
#pragma unroll
      for (int dz = -1; dz <= 1; dz++) {
        uint32_t v, c;
        if (!map_find(m, pack_key(kx + dx, ky + dy, kz + dz), v, c)) continue;
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

__global__ void k_nn_single(MapDev m, const float* __restrict__ q, uint32_t stride, uint32_t n, float* __restrict__ out_xyz,
                            float* __restrict__ out_d2, uint8_t* __restrict__ out_found) {
  const uint32_t i = blockIdx.x * blockDim.x + threadIdx.x;
  if (i >= n) return;
  const float* p = q + size_t(i) * stride;
  const NNHit r = nn_single_thread(m, p[0], p[1], p[2]);
  out_xyz[3 * size_t(i)] = r.x;
  out_xyz[3 * size_t(i) + 1] = r.y;
  out_xyz[3 * size_t(i) + 2] = r.z;
  out_d2[i] = r.d2;
  out_found[i] = uint8_t(r.found);
}

// ------------------------------------------------------------------ export
__global__ void k_export_count(MapDev m, uint64_t n_slots, uint32_t* n_vox) {
  const uint64_t i = uint64_t(blockIdx.x) * blockDim.x + threadIdx.x;
  if (i >= n_slots) return;
  if (slot_key(m.slots[i]) != KEY_EMPTY) atomicAdd(n_vox, 1u);
}
// writes (key, count, vid) of every live slot, unordered; the host sorts by key.
__global__ void k_export_list(MapDev m, uint64_t n_slots, uint32_t* cursor, uint64_t* keys, uint32_t* counts,
                              uint32_t* vids) {
  const uint64_t i = uint64_t(blockIdx.x) * blockDim.x + threadIdx.x;
  if (i >= n_slots) return;
  const uint4 s = m.slots[i];
  const uint64_t k = slot_key(s);
  if (k == KEY_EMPTY) return;
  const uint32_t o = atomicAdd(cursor, 1u);
  keys[o] = k;
  counts[o] = s.w;
  vids[o] = s.z;
}

}  // namespace mlo
