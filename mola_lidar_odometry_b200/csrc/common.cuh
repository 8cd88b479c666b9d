// common.cuh — shared device/host helpers of libmlo_b200 (sm_100a).
//
// Bit-exactness rules (DESIGN.md "Numerics"): the translation unit is compiled with --fmad=false so
// every float/double expression below rounds exactly as written, in the same operation order as the
// CPU statement of the algorithm; voxel indices, NN argmin and threshold tests are therefore
// bit-identical to the reference arithmetic, while the double-precision normal-equation sums differ
// only by summation order.
#pragma once
#include <cuda_runtime.h>
#include <stdint.h>

#include <cmath>
#include <cstdio>
#include <cstring>
#include <string>

#include "mlo_b200.h"

#define MLO_HD __host__ __device__ __forceinline__
#define MLO_D __device__ __forceinline__

namespace mlo {

constexpr uint64_t KEY_EMPTY = 0xFFFFFFFFFFFFFFFFull;
constexpr int32_t KEY_BIAS = 1 << 20;  // 21 bits per axis
constexpr uint32_t HARD_LIMIT_PTS = 32; // upstream HARDLIMIT_MAX_POINTS_PER_VOXEL
enum : uint32_t { ERR_CAPACITY = 1u, ERR_KEY_RANGE = 2u, ERR_CTA_FALLBACK = 4u };  // device-side error bits (the last: filter.cuh k_decim_cta)
constexpr uint32_t MAP_COUNTERS = 8;  // u32 words of MapDev::counters (map.cuh)

// mola::HashedVoxelPointCloud::coordToGlobalIdx (pipelines/lidar3d-default.yaml:233 voxel_size):
// static_cast<int32_t>(coord * voxel_size_inv), truncation toward zero.
// [VERIFY] conventions that decide discrete results and could not be checked against upstream's source (SURVEY.md
// Appendix A) are switchable per context - mlo_set_option("convention_*") - and mirrored by the oracle
// (oracle/mlo_oracle.hpp conv()): index rounding 0 = truncation toward zero (default), 1 = floor; Geman-McClure weight
// 0 = c^4/(c^2+e^2)^2 (default), 1 = c^2/(c^2+e^2)^2; cull metric 0 = max-norm in cells (default), 1 = L1 in cells,
// 2 = Euclidean in cells.  The parity tests run under both settings of each.
MLO_HD int32_t voxel_index_map(float coord, float inv_voxel, int floor_mode = 0) {
  const float v = coord * inv_voxel;
  return floor_mode ? static_cast<int32_t>(floorf(v)) : static_cast<int32_t>(v);
}
MLO_HD bool cull_out_of_range(int32_t dx, int32_t dy, int32_t dz, int32_t d, int metric) {
  const int32_t ax = dx < 0 ? -dx : dx, ay = dy < 0 ? -dy : dy, az = dz < 0 ? -dz : dz;
  if (metric == 1) return int64_t(ax) + ay + az > d;
  if (metric == 2) return int64_t(ax) * ax + int64_t(ay) * ay + int64_t(az) * az > int64_t(d) * d;
  return ax > d || ay > d || az > d;
}
// mp2p_icp_filters::FilterDecimateVoxels grid index (default.yaml:289,316): static_cast<int32_t>(coord / resolution)
MLO_HD int32_t voxel_index_filter(float coord, float resolution, int floor_mode = 0) {
  const float v = coord / resolution;
  return floor_mode ? static_cast<int32_t>(floorf(v)) : static_cast<int32_t>(v);
}

MLO_HD bool key_in_range(int32_t k) { return k > -KEY_BIAS && k < KEY_BIAS - 1; }
MLO_HD uint64_t pack_key(int32_t kx, int32_t ky, int32_t kz) {
  return (uint64_t(uint32_t(kx + KEY_BIAS) & 0x1FFFFFu) << 42) | (uint64_t(uint32_t(ky + KEY_BIAS) & 0x1FFFFFu) << 21) |
         uint64_t(uint32_t(kz + KEY_BIAS) & 0x1FFFFFu);
}
MLO_HD void unpack_key(uint64_t k, int32_t& kx, int32_t& ky, int32_t& kz) {
  kx = int32_t((k >> 42) & 0x1FFFFFu) - KEY_BIAS;
  ky = int32_t((k >> 21) & 0x1FFFFFu) - KEY_BIAS;
  kz = int32_t(k & 0x1FFFFFu) - KEY_BIAS;
}
// 32-bit cell hash: one multiply per axis (shared between the 9 columns of a neighbourhood probe) and a
// short avalanche; cheap on the integer pipe (the 64-bit finaliser below costs ~4x the instructions).
MLO_HD uint32_t hash_mix(uint32_t h) {
  h ^= h >> 15;
  h *= 0x2C1B3C6Du;
  h ^= h >> 12;
  h *= 0x297A2D39u;
  h ^= h >> 15;
  return h;
}
constexpr uint32_t HASH_PX = 0x9E3779B1u, HASH_PY = 0x85EBCA77u, HASH_PZ = 0xC2B2AE3Du;
MLO_HD uint32_t hash_cell(int32_t kx, int32_t ky, int32_t kz) {
  return hash_mix(uint32_t(kx) * HASH_PX ^ uint32_t(ky) * HASH_PY ^ uint32_t(kz) * HASH_PZ);
}
// 64-bit finaliser (splitmix / murmur3 style): spreads neighbouring cells over the table.
MLO_HD uint64_t hash_key(uint64_t k) {
  k ^= k >> 33;
  k *= 0xff51afd7ed558ccdull;
  k ^= k >> 33;
  k *= 0xc4ceb9fe1a85ec53ull;
  k ^= k >> 33;
  return k;
}

// Row-major 3x4 pose in double.
struct Pose34 {
  double m[12];
};

// CPose3D::composePoint evaluated in double and stored as float, operation order
// R0*x + R1*y + R2*z + t (left to right), matching the CPU statement bit for bit.
MLO_HD void compose_point_f(const double* T, float lx, float ly, float lz, float& gx, float& gy, float& gz) {
  const double x = lx, y = ly, z = lz;
  gx = static_cast<float>(T[0] * x + T[1] * y + T[2] * z + T[3]);
  gy = static_cast<float>(T[4] * x + T[5] * y + T[6] * z + T[7]);
  gz = static_cast<float>(T[8] * x + T[9] * y + T[10] * z + T[11]);
}

MLO_HD uint32_t hash_packed(uint64_t k) {
  int32_t kx, ky, kz;
  unpack_key(k, kx, ky, kz);
  return hash_cell(kx, ky, kz);
}

MLO_HD float sqr_dist(float ax, float ay, float az, float bx, float by, float bz) {
  const float dx = ax - bx, dy = ay - by, dz = az - bz;
  return dx * dx + dy * dy + dz * dz;
}

}  // namespace mlo
