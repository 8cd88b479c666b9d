// filter.cuh — voxel-grid decimation on device (replaces mp2p_icp_filters::FilterDecimateVoxels with
// DecimateMethod::FirstPoint, pipelines/lidar3d-default.yaml:285-292,312-319) with the FilterByRange
// (:297-302) and FilterBoundingBox "outside" (:305-310) predicates fused in front of the 2nd decimation.
//
// FirstPoint on a GPU: every point hashes its voxel into a scratch table and does atomicMin(first, i);
// the survivors are the points with first[voxel] == i, compacted in input order by a block scan.
// One launch handles a batch: blockIdx.y selects the job (one cloud each).
#pragma once
#include "common.cuh"

namespace mlo {

struct DecimJob {
  const float* in;         // first input point
  const float* in_t;       // optional per-point channel carried in .w (timestamps for FilterDeskew), else nullptr
  int32_t keep_w;          // carry .w of the input (stride 4) through to the outputs
  uint32_t in_stride;      // floats per point (3 or 4)
  uint32_t n_in_static;    // input size when n_in_dev == nullptr
  const uint32_t* n_in_dev;  // input size produced on device by the previous stage
  float resolution;
  uint32_t min_pts;
  int32_t use_range;
  float rmin2, rmax2;
  int32_t use_bbox;
  float bmin[3], bmax[3];
  uint64_t* tab_keys;  // scratch hash: keys (0xFF.. = empty) and first index (0xFFFFFFFF)
  uint32_t* tab_first;
  uint32_t tab_mask;
  uint32_t* pslot;     // per input point: its table slot (or NONE)
  uint8_t* flags;      // per input point: bit0 keepA (predicate), bit1 keepB (predicate && first-in-voxel)
  uint32_t* blockcnt;  // [nblocks][2]
  uint32_t* blockoff;  // [nblocks][2]
  uint32_t* npred;     // number of predicate survivors
  uint32_t* err;       // error bits (ERR_KEY_RANGE)
  float4* outA;        // predicate survivors ("decimated_for_map_skewed"), may be null
  uint32_t* nA;
  float4* outB;        // decimated survivors
  uint32_t* nB;
  uint32_t* outB_idx;  // optional: input indices of outB
};

constexpr uint32_t DECIM_BLOCK = 256;
constexpr uint32_t SLOT_NONE = 0xFFFFFFFFu;

MLO_D uint32_t job_n(const DecimJob& j) { return j.n_in_dev ? *j.n_in_dev : j.n_in_static; }

MLO_D float4 load_point(const float* base, uint32_t stride, uint32_t i) {
  if (stride == 4) return __ldg(reinterpret_cast<const float4*>(base) + i);
  const float* p = base + size_t(i) * stride;
  return make_float4(__ldg(p), __ldg(p + 1), __ldg(p + 2), 0.f);
}

MLO_D bool predicate_keep(const DecimJob& j, float x, float y, float z) {
  if (j.use_range) {
    const float n2 = x * x + y * y + z * z;
    if (!(n2 >= j.rmin2 && n2 <= j.rmax2)) return false;
  }
  if (j.use_bbox) {
    const bool inside = x >= j.bmin[0] && y >= j.bmin[1] && z >= j.bmin[2] && x <= j.bmax[0] && y <= j.bmax[1] &&
                        z <= j.bmax[2];
    if (inside) return false;
  }
  return true;
}

// Neighbouring returns of a sweep mostly fall into the same voxel, so the lanes of a warp first agree on their distinct
// keys (MATCH.ANY): only the lowest lane of each group — the smallest input index of the group — probes the table and does
// the atomicMin; the others take its slot.  Same table contents as one atomic per point, several times fewer atomics.
__global__ void __launch_bounds__(DECIM_BLOCK) k_decim_hash(const DecimJob* __restrict__ jobs) {
  const DecimJob& j = jobs[blockIdx.y];
  const uint32_t n = job_n(j);
  if (blockIdx.x * DECIM_BLOCK >= n) return;  // whole block beyond the cloud (uniform)
  const uint32_t i = blockIdx.x * DECIM_BLOCK + threadIdx.x;
  const uint32_t lane = threadIdx.x & 31u;
  bool pred = false, valid = false;
  uint64_t key = 0;
  uint32_t h = 0;
  if (i < n) {
    const float4 p = load_point(j.in, j.in_stride, i);
    pred = predicate_keep(j, p.x, p.y, p.z);
    if (pred) {
      const int32_t kx = voxel_index_filter(p.x, j.resolution), ky = voxel_index_filter(p.y, j.resolution),
                    kz = voxel_index_filter(p.z, j.resolution);
      if (key_in_range(kx) && key_in_range(ky) && key_in_range(kz)) {
        key = pack_key(kx, ky, kz);
        h = hash_cell(kx, ky, kz) & j.tab_mask;
        valid = true;
      } else {
        atomicOr(j.err, ERR_KEY_RANGE);
        pred = false;
      }
    }
  }
  // packed keys use 63 bits: the top bit marks lanes without a key (each its own group)
  const uint64_t mkey = valid ? key : (0x8000000000000000ull | lane);
  const uint32_t peers = __match_any_sync(0xFFFFFFFFu, mkey);
  const int leader = __ffs(peers) - 1;
  uint32_t slot = SLOT_NONE;
  if (valid && int(lane) == leader) {
    for (;;) {
      unsigned long long* kp = reinterpret_cast<unsigned long long*>(&j.tab_keys[h]);
      unsigned long long cur = *reinterpret_cast<volatile unsigned long long*>(kp);
      if (cur == KEY_EMPTY) cur = atomicCAS(kp, (unsigned long long)KEY_EMPTY, (unsigned long long)key);
      if (cur == KEY_EMPTY || cur == key) break;
      h = (h + 1) & j.tab_mask;
    }
    atomicMin(&j.tab_first[h], i);
    slot = h;
  }
  slot = __shfl_sync(0xFFFFFFFFu, slot, leader);
  if (i < n) j.pslot[i] = valid ? slot : SLOT_NONE;
  const uint32_t c = __syncthreads_count(pred);
  if (threadIdx.x == 0 && c) atomicAdd(j.npred, c);
}

__global__ void __launch_bounds__(DECIM_BLOCK) k_decim_flag(const DecimJob* __restrict__ jobs) {
  const DecimJob& j = jobs[blockIdx.y];
  const uint32_t n = job_n(j);
  const uint32_t i = blockIdx.x * DECIM_BLOCK + threadIdx.x;
  if (blockIdx.x * DECIM_BLOCK >= n) return;  // whole block beyond the cloud (uniform)
  bool a = false, b = false;
  if (i < n) {
    const uint32_t s = j.pslot[i];
    a = (s != SLOT_NONE);
    // minimum_input_points_to_filter: below it the layer passes through undecimated
    b = a && (*j.npred < j.min_pts || j.tab_first[s] == i);
    j.flags[i] = uint8_t((a ? 1 : 0) | (b ? 2 : 0));
  }
  const uint32_t ca = __syncthreads_count(a), cb = __syncthreads_count(b);
  if (threadIdx.x == 0) {
    j.blockcnt[2 * blockIdx.x] = ca;
    j.blockcnt[2 * blockIdx.x + 1] = cb;
  }
}

// one block per job: exclusive scan of the per-block counts -> block offsets and totals
__global__ void __launch_bounds__(512) k_decim_scan(const DecimJob* __restrict__ jobs) {
  const DecimJob& j = jobs[blockIdx.x];
  const uint32_t n = job_n(j);
  const uint32_t nblk = (n + DECIM_BLOCK - 1) / DECIM_BLOCK;
  __shared__ uint32_t wsum[2][16];
  __shared__ uint32_t carry[2];
  if (threadIdx.x < 2) carry[threadIdx.x] = 0;
  __syncthreads();
  const uint32_t lane = threadIdx.x & 31, warp = threadIdx.x >> 5;
  for (uint32_t base = 0; base < nblk; base += 512) {
    const uint32_t b = base + threadIdx.x;
    uint32_t v[2] = {0, 0}, inc[2];
    if (b < nblk) {
      v[0] = j.blockcnt[2 * b];
      v[1] = j.blockcnt[2 * b + 1];
    }
#pragma unroll
    for (int k = 0; k < 2; k++) {
      uint32_t x = v[k];
#pragma unroll
      for (int o = 1; o < 32; o <<= 1) {
        const uint32_t y = __shfl_up_sync(0xFFFFFFFFu, x, o);
        if (lane >= o) x += y;
      }
      inc[k] = x;
      if (lane == 31) wsum[k][warp] = x;
    }
    __syncthreads();
    if (warp == 0) {
#pragma unroll
      for (int k = 0; k < 2; k++) {
        uint32_t x = (lane < 16) ? wsum[k][lane] : 0;
#pragma unroll
        for (int o = 1; o < 16; o <<= 1) {
          const uint32_t y = __shfl_up_sync(0xFFFFFFFFu, x, o);
          if (lane >= o) x += y;
        }
        if (lane < 16) wsum[k][lane] = x;  // inclusive over warps
      }
    }
    __syncthreads();
#pragma unroll
    for (int k = 0; k < 2; k++) {
      const uint32_t wprev = warp ? wsum[k][warp - 1] : 0;
      if (b < nblk) j.blockoff[2 * b + k] = carry[k] + wprev + inc[k] - v[k];
    }
    __syncthreads();
    if (threadIdx.x == 0) {
      carry[0] += wsum[0][15];
      carry[1] += wsum[1][15];
    }
    __syncthreads();
  }
  if (threadIdx.x == 0) {
    if (j.nA) *j.nA = carry[0];
    *j.nB = carry[1];
  }
}

__global__ void __launch_bounds__(DECIM_BLOCK) k_decim_scatter(const DecimJob* __restrict__ jobs) {
  const DecimJob& j = jobs[blockIdx.y];
  const uint32_t n = job_n(j);
  if (blockIdx.x * DECIM_BLOCK >= n) return;
  const uint32_t i = blockIdx.x * DECIM_BLOCK + threadIdx.x;
  const uint32_t lane = threadIdx.x & 31, warp = threadIdx.x >> 5;
  uint32_t f = 0;
  float4 p = make_float4(0, 0, 0, 0);
  if (i < n) {
    f = j.flags[i];
    if (f) {
      p = load_point(j.in, j.in_stride, i);
      if (j.in_t) p.w = __ldg(j.in_t + i);
      else if (!j.keep_w) p.w = 0.f;
    }
  }
  __shared__ uint32_t wcnt[2][DECIM_BLOCK / 32];
  const uint32_t ba = __ballot_sync(0xFFFFFFFFu, f & 1u), bb = __ballot_sync(0xFFFFFFFFu, f & 2u);
  if (lane == 0) {
    wcnt[0][warp] = __popc(ba);
    wcnt[1][warp] = __popc(bb);
  }
  __syncthreads();
  uint32_t offA = j.blockoff[2 * blockIdx.x], offB = j.blockoff[2 * blockIdx.x + 1];
  for (uint32_t w = 0; w < warp; w++) {
    offA += wcnt[0][w];
    offB += wcnt[1][w];
  }
  const uint32_t lt = (1u << lane) - 1u;
  if ((f & 1u) && j.outA) j.outA[offA + __popc(ba & lt)] = p;
  if (f & 2u) {
    const uint32_t o = offB + __popc(bb & lt);
    j.outB[o] = p;
    if (j.outB_idx) j.outB_idx[o] = i;
  }
}

// mp2p_icp_filters::FilterDeskew (pipelines/lidar3d-default.yaml:328-350): p' = exp_SO3(w t) p + v t, t = in.w.
// Same operation order and the same small-angle series as the CPU statement (bit-exact for |w t| < 0.05 rad).
MLO_D void deskew_coeffs(double th2, double& A, double& B) {
  if (th2 < 2.5e-3) {
    A = 1.0 - th2 * (1.0 / 6.0) * (1.0 - th2 * (1.0 / 20.0) * (1.0 - th2 * (1.0 / 42.0) * (1.0 - th2 * (1.0 / 72.0))));
    B = 0.5 * (1.0 - th2 * (1.0 / 12.0) * (1.0 - th2 * (1.0 / 30.0) * (1.0 - th2 * (1.0 / 56.0) * (1.0 - th2 * (1.0 / 90.0)))));
  } else {
    const double th = sqrt(th2);
    A = sin(th) / th;
    B = (1.0 - cos(th)) / th2;
  }
}
struct Twist6 {
  double v[6];
};
__global__ void k_deskew(const float4* __restrict__ in, uint32_t n, Twist6 tw, float4* __restrict__ out) {
  const uint32_t i = blockIdx.x * blockDim.x + threadIdx.x;
  if (i >= n) return;
  const float4 p = __ldg(&in[i]);
  const double x = p.x, y = p.y, z = p.z, t = p.w;
  const double wx = tw.v[3] * t, wy = tw.v[4] * t, wz = tw.v[5] * t;
  double A, B;
  deskew_coeffs(wx * wx + wy * wy + wz * wz, A, B);
  const double cx = wy * z - wz * y, cy = wz * x - wx * z, cz = wx * y - wy * x;
  const double dx = wy * cz - wz * cy, dy = wz * cx - wx * cz, dz = wx * cy - wy * cx;
  out[i] = make_float4(static_cast<float>(x + A * cx + B * dx + tw.v[0] * t), static_cast<float>(y + A * cy + B * dy + tw.v[1] * t),
                       static_cast<float>(z + A * cz + B * dz + tw.v[2] * t), p.w);
}

}  // namespace mlo
