// octet.cuh — eight lanes of a warp share one query (a warp serves four queries at once): neighbourhood probe and the
// mola::NDT nearest-plane search in that shape.  Used by the block-per-problem kernel (icp_block.cuh) and by the
// warp-per-query chunks of the queue-driven kernel (icp.cuh chunk_match_warp, point-to-plane matcher).
#pragma once
#include "map.cuh"

namespace mlo {

MLO_D unsigned long long octet_min_u64(unsigned long long k) {
#pragma unroll
  for (int o = 1; o < 8; o <<= 1) {
    const unsigned long long other = __shfl_xor_sync(0xFFFFFFFFu, k, o);
    k = other < k ? other : k;
  }
  return k;
}
MLO_D uint32_t octet_sum_u32(uint32_t v) {
#pragma unroll
  for (int o = 1; o < 8; o <<= 1) v += __shfl_xor_sync(0xFFFFFFFFu, v, o);
  return v;
}

// Probe the neighbourhood of cell kq for this octet's query; the 27 packed words land in `ow` (shared, per octet).
// Returns (on every lane of the octet) the number of points stored in the 27 cells.  Warp-uniform control flow except the
// per-lane collision loops.
MLO_D uint32_t octet_probe(const MapDev& m, const int32_t kq[3], bool in_range, uint32_t* ow) {
  const uint32_t sub = threadIdx.x & 7u;
  const int32_t kz = kq[2];
  const int32_t zq0 = (kz - 1) >> 2, zq1 = (kz + 1) >> 2;
  BucketRO b[3];
  uint64_t key[3], h[3];
  bool act[3];
#pragma unroll
  for (int r = 0; r < 3; r++) {
    const uint32_t idx = sub + 8u * r;  // probe index = column * 2 + half
    const uint32_t col = idx >> 1, half = idx & 1u;
    act[r] = in_range && idx < 18u && (half == 0u || zq1 != zq0);
    key[r] = pack_key(kq[0] + int32_t(col / 3) - 1, kq[1] + int32_t(col % 3) - 1, half ? zq1 : zq0);
    h[r] = uint64_t(hash_packed(key[r])) & m.mask;
    b[r].key = KEY_EMPTY;
    if (act[r]) b[r] = load_bucket256(m, h[r]);
  }
#pragma unroll
  for (int r = 0; r < 3; r++) {
    if (act[r]) {
      while (b[r].key != key[r] && b[r].key != KEY_EMPTY) {  // hash collision: linear probing
        h[r] = (h[r] + 1) & m.mask;
        b[r] = load_bucket256(m, h[r]);
      }
    }
  }
  if (!in_range) {
#pragma unroll
    for (int r = 0; r < 4; r++)
      if (sub + 8u * r < 27u) ow[sub + 8u * r] = CELL_ABSENT;
  } else {
#pragma unroll
    for (int r = 0; r < 3; r++) {
      const uint32_t idx = sub + 8u * r;
      if (idx >= 18u) continue;
      const uint32_t col = idx >> 1, half = idx & 1u;
      const int32_t myzq = half ? zq1 : zq0;
      if (half && zq1 == zq0) continue;  // (the column's single bucket was taken by the half-0 probe)
      const bool have = b[r].key == key[r];
#pragma unroll
      for (int t = 0; t < 3; t++) {
        const int32_t z = kz - 1 + t;
        if ((z >> 2) != myzq) continue;
        const uint32_t sidx = uint32_t(z & 3);
        uint32_t w = sidx == 0 ? b[r].cell[0] : sidx == 1 ? b[r].cell[1] : sidx == 2 ? b[r].cell[2] : b[r].cell[3];
        if (!have || w == CELL_PENDING) w = CELL_ABSENT;
        ow[col * 3 + t] = w;
      }
    }
  }
  __syncwarp();
  uint32_t n = 0;
#pragma unroll
  for (int r = 0; r < 4; r++) {
    const uint32_t e = sub + 8u * r;
    if (e < 27u) {
      const uint32_t w = ow[e];
      if (w != CELL_ABSENT) n += cell_cnt(w);
    }
  }
  return octet_sum_u32(n);
}

// mola::NDT nearest-plane query for the octet's query: lane s takes cells s, s+8, s+16, s+24; the winner is the smallest
// (|n.(q - mean)|, canonical order) = the first minimum of the sequential (dx, dy, dz) scan.
MLO_D PlaneHit octet_plane(const MapDev& m, float qx, float qy, float qz, bool want, const uint32_t* ow) {
  const uint32_t FULL = 0xFFFFFFFFu;
  const uint32_t lane = threadIdx.x & 31u, sub = lane & 7u, oshift = lane & 24u;
  float bd = __int_as_float(0x7f800000);
  uint32_t be = 0xFFFFFFFFu, n = 0;
  float4 bmu = make_float4(0.f, 0.f, 0.f, 0.f), bn = make_float4(0.f, 0.f, 0.f, 0.f);
  float4 mu[4];
  uint32_t vid[4];
#pragma unroll
  for (int r = 0; r < 4; r++) {
    const uint32_t e = sub + 8u * r;
    vid[r] = CELL_ABSENT;
    if (want && e < 27u) {
      const uint32_t w = ow[e];
      if (w != CELL_ABSENT) {
        vid[r] = cell_vid(w);
        mu[r] = __ldg(&m.mean[vid[r]]);
      }
    }
  }
#pragma unroll
  for (int r = 0; r < 4; r++) {
    if (vid[r] == CELL_ABSENT) continue;
    n += 2;
    if (mu[r].w == 0.f) continue;
    const float4 nr = __ldg(&m.normal[vid[r]]);
    const float ex = qx - mu[r].x, ey = qy - mu[r].y, ez = qz - mu[r].z;
    const float d = fabsf(nr.x * ex + nr.y * ey + nr.z * ez);
    if (d < bd) {  // (this lane's cells come in ascending canonical order)
      bd = d;
      be = sub + 8u * r;
      bmu = mu[r];
      bn = nr;
    }
  }
  const unsigned long long mykey = be == 0xFFFFFFFFu ? ~0ull : ((uint64_t(__float_as_uint(bd)) << 32) | uint64_t(be));
  const unsigned long long best = octet_min_u64(mykey);
  const uint32_t winners = (__ballot_sync(FULL, mykey == best && best != ~0ull) >> oshift) & 0xFFu;
  const uint32_t src = oshift + (winners ? uint32_t(__ffs(winners) - 1) : 0u);
  PlaneHit r;
  r.cx = __shfl_sync(FULL, bmu.x, src);
  r.cy = __shfl_sync(FULL, bmu.y, src);
  r.cz = __shfl_sync(FULL, bmu.z, src);
  r.nx = __shfl_sync(FULL, bn.x, src);
  r.ny = __shfl_sync(FULL, bn.y, src);
  r.nz = __shfl_sync(FULL, bn.z, src);
  r.dist = __uint_as_float(uint32_t(best >> 32));
  r.found = best != ~0ull;
  r.ncand = octet_sum_u32(n);
  return r;
}

}  // namespace mlo
