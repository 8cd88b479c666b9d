// filter.cuh — voxel-grid decimation on device (replaces mp2p_icp_filters::FilterDecimateVoxels with
// DecimateMethod::FirstPoint, pipelines/lidar3d-default.yaml:285-292,312-319) with the FilterByRange
// (:297-302) and FilterBoundingBox "outside" (:305-310) predicates fused in front of the 2nd decimation.
//
// FirstPoint on a GPU: every point hashes its voxel into a scratch table and does atomicMin(first, i);
// the survivors are the points with first[voxel] == i, compacted in input order (k_decim_claim / k_decim_finalize).
// One launch handles a group of clouds: blockIdx.y selects the job (one cloud each).
#pragma once
#include "common.cuh"

namespace mlo {

// Predicates fused with a decimation: FilterByRange (keep rmin <= |p| <= rmax) and FilterBoundingBox "outside".
struct PointPred {
  int32_t use_range;
  float rmin2, rmax2;
  int32_t use_bbox;
  float bmin[3], bmax[3];
};

// One 16-byte table entry = half a DRAM sector: voxel key and the smallest input index seen in that voxel sit in the
// same sector, so the claim (CAS on the key) and the atomicMin that follows touch one L2 line.
struct __align__(16) DecimEntry {
  unsigned long long key;  // KEY_EMPTY = free
  uint32_t first;          // smallest input index that fell into this voxel
  uint32_t pad;
};
static_assert(sizeof(DecimEntry) == 16, "entry layout");

struct DecimJob {
  const float* in;         // first input point
  const float* in_t;       // optional per-point channel carried in .w (timestamps for FilterDeskew), else nullptr
  int32_t keep_w;          // carry .w of the input (stride 4) through to the output
  uint32_t in_stride;      // floats per point (3 or 4)
  uint32_t n_in_static;    // input size when n_in_dev == nullptr
  const uint32_t* n_in_dev;  // input size produced on device by the previous stage
  float resolution;
  int32_t index_floor;     // [VERIFY] convention: grid index = floor instead of truncation (common.cuh)
  uint32_t min_pts;        // minimum_input_points_to_filter: below it the cloud passes through undecimated
  PointPred pre;           // applied BEFORE the decimation (mlo_voxel_decimate_first with range / bbox)
  PointPred post;          // applied to the decimated points (the 1st-pass pipeline: by-range and bbox sit AFTER the
                           // first FilterDecimateVoxels, pipelines/lidar3d-default.yaml:285-310)
  DecimEntry* tab;         // scratch hash (all bytes 0xFF = empty)
  uint32_t tab_mask;
  float4* cand_pt;         // per block of DECIM_BLOCK inputs: the block's candidate points, in input order ...
  uint2* cand_meta;        // ... with (table slot, input index)
  uint32_t* blockcnt;      // candidates per block
  unsigned long long* status;  // decoupled look-back: (flag << 32) | count per block; 0 = not yet published
  uint32_t* npred;         // number of points that passed `pre`
  uint32_t* err;           // error bits (ERR_KEY_RANGE, ERR_CAPACITY = scratch table exhausted)
  float4* out;             // decimated (and post-filtered) points, input order
  uint32_t* n_out;
  uint32_t* out_idx;       // optional: input indices of `out`
};

constexpr uint32_t DECIM_BLOCK = 256;
constexpr uint32_t SLOT_NONE = 0xFFFFFFFFu;

MLO_D uint32_t job_n(const DecimJob& j) { return j.n_in_dev ? *j.n_in_dev : j.n_in_static; }

MLO_D float4 load_point(const float* base, uint32_t stride, uint32_t i) {
  if (stride == 4) return __ldg(reinterpret_cast<const float4*>(base) + i);
  const float* p = base + size_t(i) * stride;
  return make_float4(__ldg(p), __ldg(p + 1), __ldg(p + 2), 0.f);
}

MLO_D bool predicate_keep(const PointPred& j, float x, float y, float z) {
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

// FirstPoint decimation in two passes that read the cloud ONCE:
//
//  k_decim_claim     PPT input points per thread (all loads of a thread issued before the first use): predicate, voxel
//                    key, then the lanes of a warp agree on their distinct keys (MATCH.ANY: neighbouring returns of a
//                    sweep share voxels) and the lowest lane of each group - the smallest input index of the group -
//                    claims the voxel's table entry (one optimistic CAS) and does atomicMin(first, i); the PPT claims
//                    and the PPT atomicMins of a thread are in flight together.  A point whose atomicMin did not lower
//                    the entry is already beaten and is dropped here; the others are CANDIDATES (about one per output
//                    point): the block writes them, in input order, with their coordinates into its own slice of the
//                    candidate buffers.
//  k_decim_finalize  PPT candidates per thread: winner iff first[slot] is still its index; `post` predicates; ordered
//                    compaction across the blocks of a cloud by a decoupled look-back (each block publishes its count,
//                    a warp sums its predecessors' counts 32 at a time); winners go out in input order.
//
// The raw cloud, the per-point slot array and the flag array of a hash / flag / scan / scatter chain are never re-read:
// the second pass touches candidates only.  One launch handles a group of clouds (blockIdx.y).
template <int PPT>
__global__ void __launch_bounds__(DECIM_BLOCK) k_decim_claim(const DecimJob* __restrict__ jobs) {
  static_assert(PPT * (DECIM_BLOCK / 32) <= 32, "one warp scans the per-(pass, warp) counts");
  const DecimJob& j = jobs[blockIdx.y];
  const uint32_t n = job_n(j);
  constexpr uint32_t TILE = DECIM_BLOCK * PPT;
  const uint32_t base = blockIdx.x * TILE;
  if (base >= n) return;  // whole block beyond the cloud (uniform)
  const uint32_t lane = threadIdx.x & 31u, warp = threadIdx.x >> 5;
  const bool has_pre = j.pre.use_range || j.pre.use_bbox;
  const bool pass_all = n < j.min_pts;  // fewer inputs than minimum_input_points_to_filter: certain pass-through
  float4 p[PPT];
  bool pred[PPT], valid[PPT], cand[PPT];
  uint64_t key[PPT];
  uint32_t h[PPT], slot[PPT];
#pragma unroll
  for (int u = 0; u < PPT; u++) {
    const uint32_t i = base + u * DECIM_BLOCK + threadIdx.x;
    p[u] = make_float4(0.f, 0.f, 0.f, 0.f);
    if (i < n) {
      p[u] = load_point(j.in, j.in_stride, i);
      if (j.in_t) p[u].w = __ldg(j.in_t + i);
      else if (!j.keep_w) p[u].w = 0.f;
    }
  }
#pragma unroll
  for (int u = 0; u < PPT; u++) {
    const uint32_t i = base + u * DECIM_BLOCK + threadIdx.x;
    pred[u] = valid[u] = cand[u] = false;
    key[u] = 0;
    h[u] = 0;
    slot[u] = SLOT_NONE;
    if (i < n) {
      pred[u] = predicate_keep(j.pre, p[u].x, p[u].y, p[u].z);
      if (pred[u]) {
        const int32_t kx = voxel_index_filter(p[u].x, j.resolution, j.index_floor), ky = voxel_index_filter(p[u].y, j.resolution, j.index_floor),
                      kz = voxel_index_filter(p[u].z, j.resolution, j.index_floor);
        if (key_in_range(kx) && key_in_range(ky) && key_in_range(kz)) {
          key[u] = pack_key(kx, ky, kz);
          h[u] = hash_cell(kx, ky, kz) & j.tab_mask;
          valid[u] = true;
        } else {
          atomicOr(j.err, ERR_KEY_RANGE);
          pred[u] = false;
        }
      }
    }
  }
  if (pass_all) {
#pragma unroll
    for (int u = 0; u < PPT; u++) cand[u] = valid[u];
  } else {
    int leader[PPT];
    bool lead[PPT];
    unsigned long long cur[PPT];
#pragma unroll
    for (int u = 0; u < PPT; u++) {
      // packed keys use 63 bits: the top bit marks lanes without a key (each its own group)
      const uint64_t mkey = valid[u] ? key[u] : (0x8000000000000000ull | lane);
      const uint32_t peers = __match_any_sync(0xFFFFFFFFu, mkey);
      leader[u] = __ffs(peers) - 1;
      lead[u] = valid[u] && int(lane) == leader[u];
    }
    // optimistic claim: one CAS per leader, all of a thread's in flight together
#pragma unroll
    for (int u = 0; u < PPT; u++) {
      cur[u] = KEY_EMPTY;
      if (lead[u]) cur[u] = atomicCAS(&j.tab[h[u]].key, (unsigned long long)KEY_EMPTY, (unsigned long long)key[u]);
    }
    // collisions (another voxel sits in the slot): linear probing
#pragma unroll
    for (int u = 0; u < PPT; u++) {
      if (lead[u] && cur[u] != KEY_EMPTY && cur[u] != key[u]) {
        uint32_t probes = 0;
        for (;;) {
          h[u] = (h[u] + 1) & j.tab_mask;
          if (++probes > j.tab_mask) {  // scratch table exhausted (the caller retries with a larger one)
            atomicOr(j.err, ERR_CAPACITY);
            h[u] = SLOT_NONE;
            break;
          }
          const unsigned long long c2 = atomicCAS(&j.tab[h[u]].key, (unsigned long long)KEY_EMPTY, (unsigned long long)key[u]);
          if (c2 == KEY_EMPTY || c2 == key[u]) break;
        }
      }
    }
    uint32_t old[PPT];
#pragma unroll
    for (int u = 0; u < PPT; u++) {
      old[u] = 0;
      if (lead[u] && h[u] != SLOT_NONE) old[u] = atomicMin(&j.tab[h[u]].first, base + u * DECIM_BLOCK + threadIdx.x);
    }
#pragma unroll
    for (int u = 0; u < PPT; u++) {
      const uint32_t i = base + u * DECIM_BLOCK + threadIdx.x;
      const bool lowered = lead[u] && h[u] != SLOT_NONE && old[u] > i;
      slot[u] = __shfl_sync(0xFFFFFFFFu, lead[u] ? h[u] : SLOT_NONE, leader[u]);
      // with `pre` predicates the pass-through rule depends on the number of survivors, known only after this pass:
      // every survivor stays a candidate and k_decim_finalize applies whichever rule holds
      cand[u] = valid[u] && (lowered || has_pre);
      if (!valid[u]) slot[u] = SLOT_NONE;
    }
  }
  // ---- block-ordered compaction of the candidates into this block's slice (input order = pass-major)
  constexpr int NW = DECIM_BLOCK / 32;
  __shared__ uint32_t wcnt[PPT * NW];
  __shared__ uint32_t s_total, s_pred;
  uint32_t bc[PPT];
  uint32_t np = 0;
#pragma unroll
  for (int u = 0; u < PPT; u++) {
    bc[u] = __ballot_sync(0xFFFFFFFFu, cand[u]);
    np += __popc(__ballot_sync(0xFFFFFFFFu, pred[u]));
    if (lane == 0) wcnt[u * NW + warp] = __popc(bc[u]);
  }
  if (threadIdx.x == 0) s_pred = 0;
  __syncthreads();
  if (lane == 0 && np) atomicAdd(&s_pred, np);
  if (warp == 0) {
    const uint32_t v = lane < PPT * NW ? wcnt[lane] : 0u;
    uint32_t incl = v;
#pragma unroll
    for (int o = 1; o < 32; o <<= 1) {
      const uint32_t y = __shfl_up_sync(0xFFFFFFFFu, incl, o);
      if (lane >= uint32_t(o)) incl += y;
    }
    if (lane < PPT * NW) wcnt[lane] = incl - v;
    if (lane == 31) s_total = incl;
  }
  __syncthreads();
#pragma unroll
  for (int u = 0; u < PPT; u++) {
    if (cand[u]) {
      const size_t o = size_t(base) + wcnt[u * NW + warp] + __popc(bc[u] & ((1u << lane) - 1u));
      j.cand_pt[o] = p[u];
      j.cand_meta[o] = make_uint2(slot[u], base + u * DECIM_BLOCK + threadIdx.x);
    }
  }
  if (threadIdx.x == 0) {
    j.blockcnt[blockIdx.x] = s_total;
    if (s_pred) atomicAdd(j.npred, s_pred);
  }
}

// Exclusive prefix of `count` over the blocks [0, b) of one cloud (decoupled look-back), called by warp 0 of block b.
// status[k] = (flag << 32) | value with flag 1 = the block's own count, 2 = inclusive prefix up to and including it.
// Blocks of a cloud are dispatched in ascending blockIdx.x, so every predecessor is running or done: no deadlock.
MLO_D uint32_t lookback_exclusive(unsigned long long* status, uint32_t b, uint32_t count) {
  const uint32_t FULL = 0xFFFFFFFFu, lane = threadIdx.x & 31u;
  volatile unsigned long long* st = status;
  if (lane == 0) st[b] = ((b == 0 ? 2ull : 1ull) << 32) | count;
  if (b == 0) return 0u;
  uint32_t excl = 0;
  int32_t idx = int32_t(b) - 1 - int32_t(lane);
  for (;;) {
    unsigned long long s = idx >= 0 ? st[idx] : (2ull << 32);
    while (__any_sync(FULL, (s >> 32) == 0ull)) {
      if ((s >> 32) == 0ull) s = st[idx];
    }
    const uint32_t pm = __ballot_sync(FULL, (s >> 32) == 2ull);
    const uint32_t upto = pm ? uint32_t(__ffs(pm) - 1) : 31u;  // nearest predecessor holding an inclusive prefix
    uint32_t v = lane <= upto ? uint32_t(s & 0xFFFFFFFFull) : 0u;
#pragma unroll
    for (int o = 16; o > 0; o >>= 1) v += __shfl_xor_sync(FULL, v, o);
    excl += v;
    if (pm) break;
    idx -= 32;
  }
  if (lane == 0) st[b] = (2ull << 32) | (excl + count);
  return excl;
}

template <int PPT>
__global__ void __launch_bounds__(DECIM_BLOCK) k_decim_finalize(const DecimJob* __restrict__ jobs) {
  const DecimJob& j = jobs[blockIdx.y];
  const uint32_t n = job_n(j);
  constexpr uint32_t TILE = DECIM_BLOCK * PPT;
  const uint32_t base = blockIdx.x * TILE;
  if (base >= n) return;
  const uint32_t nblk = (n + TILE - 1) / TILE;
  const uint32_t lane = threadIdx.x & 31u, warp = threadIdx.x >> 5;
  const uint32_t c = j.blockcnt[blockIdx.x];
  const bool has_pre = j.pre.use_range || j.pre.use_bbox;
  const bool pass_all = n < j.min_pts || (has_pre && *j.npred < j.min_pts);
  bool keep[PPT];
  float4 p[PPT];
  uint2 m[PPT];
#pragma unroll
  for (int u = 0; u < PPT; u++) {
    const uint32_t k = u * DECIM_BLOCK + threadIdx.x;
    keep[u] = false;
    m[u] = make_uint2(SLOT_NONE, 0u);
    p[u] = make_float4(0.f, 0.f, 0.f, 0.f);
    if (k < c) {
      m[u] = j.cand_meta[size_t(base) + k];
      p[u] = j.cand_pt[size_t(base) + k];
    }
  }
  uint32_t first[PPT];
#pragma unroll
  for (int u = 0; u < PPT; u++) {
    first[u] = 0xFFFFFFFFu;
    if (!pass_all && m[u].x != SLOT_NONE) first[u] = j.tab[m[u].x].first;
  }
  constexpr int NW = DECIM_BLOCK / 32;
  __shared__ uint32_t wcnt[PPT * NW];
  __shared__ uint32_t s_total, s_excl;
  uint32_t bk[PPT];
#pragma unroll
  for (int u = 0; u < PPT; u++) {
    const uint32_t k = u * DECIM_BLOCK + threadIdx.x;
    if (k < c) {
      const bool win = pass_all || (m[u].x != SLOT_NONE && first[u] == m[u].y);
      keep[u] = win && predicate_keep(j.post, p[u].x, p[u].y, p[u].z);
    }
    bk[u] = __ballot_sync(0xFFFFFFFFu, keep[u]);
    if (lane == 0) wcnt[u * NW + warp] = __popc(bk[u]);
  }
  __syncthreads();
  if (warp == 0) {
    const uint32_t v = lane < PPT * NW ? wcnt[lane] : 0u;
    uint32_t incl = v;
#pragma unroll
    for (int o = 1; o < 32; o <<= 1) {
      const uint32_t y = __shfl_up_sync(0xFFFFFFFFu, incl, o);
      if (lane >= uint32_t(o)) incl += y;
    }
    if (lane < PPT * NW) wcnt[lane] = incl - v;
    const uint32_t total = __shfl_sync(0xFFFFFFFFu, incl, 31);
    const uint32_t e = lookback_exclusive(j.status, blockIdx.x, total);
    if (lane == 0) {
      s_excl = e;
      s_total = total;
      if (blockIdx.x == nblk - 1) *j.n_out = e + total;
    }
  }
  __syncthreads();
#pragma unroll
  for (int u = 0; u < PPT; u++) {
    if (keep[u]) {
      const uint32_t o = s_excl + wcnt[u * NW + warp] + __popc(bk[u] & ((1u << lane) - 1u));
      j.out[o] = p[u];
      if (j.out_idx) j.out_idx[o] = m[u].y;
    }
  }
}

// ---------------------------------------------------------------------------------------------------------------
// FirstPoint decimation of one cloud by ONE thread block, scratch in SHARED memory (large batches: one block per SM,
// a cloud per block, the hardware block scheduler hands the next cloud to whichever SM finishes first).
//
//   table   one 64-bit word per voxel: (packed voxel key << 32) | smallest input index.  Equal keys share the high
//           word, so atomicMin on the whole word IS "smallest index of that voxel"; an entry never changes its key once
//           claimed (CAS from EMPTY), so a plain read that shows the voxel with a smaller index already settles a point
//           without any atomic - the common case, since a block walks its cloud in ascending index order.
//   key     11 + 11 + 10 bits around grid index 0 (sensor-frame clouds: +-1024 cells in x / y, +-512 in z).  A point
//           outside that box, or a table that runs full, raises ERR_CTA_FALLBACK and the host repeats the batch with
//           k_decim_claim / k_decim_finalize (global tables, 3 x 21-bit keys): exact whenever it succeeds.
//   bitmap  one bit per input point; winners that pass `post` set theirs, then a block-wide prefix sum over the bitmap
//           words gives every winner its rank: the output is in input order, as FirstPoint over a sequential walk.
//
// The raw cloud is read once (coalesced, PPT loads in flight per thread); the few thousand winners are fetched again by
// index.  No global scratch, no memset, no candidate buffers, no inter-block look-back: DRAM sees N_in * 16 bytes in and
// N_out * 16 bytes out, which is SURVEY.md §8(d)'s figure for this operator.
constexpr uint32_t CTA_DECIM_THREADS = 1024;
constexpr unsigned long long CTA_EMPTY = ~0ull;
constexpr uint32_t CTA_MAX_PROBES = 256;

MLO_D bool pack_key32(int32_t kx, int32_t ky, int32_t kz, uint32_t& pk) {
  const uint32_t ux = uint32_t(kx + 1024), uy = uint32_t(ky + 1024), uz = uint32_t(kz + 512);
  pk = (ux << 21) | (uy << 10) | uz;
  return ux < 2048u && uy < 2048u && uz < 1024u;
}

// false = no room within CTA_MAX_PROBES slots
MLO_D bool cta_table_put(unsigned long long* tab, uint32_t T, uint32_t pk, uint32_t i) {
  const unsigned long long want = (uint64_t(pk) << 32) | i;
  uint32_t slot = __umulhi(hash_mix(pk * HASH_PX), T);
  for (uint32_t probes = 0; probes < CTA_MAX_PROBES; probes++) {
    unsigned long long cur = *reinterpret_cast<volatile unsigned long long*>(&tab[slot]);
    if (cur == CTA_EMPTY) {
      cur = atomicCAS(&tab[slot], CTA_EMPTY, want);
      if (cur == CTA_EMPTY) return true;
    }
    if (uint32_t(cur >> 32) == pk) {
      if (cur > want) atomicMin(&tab[slot], want);
      return true;
    }
    slot = slot + 1 == T ? 0u : slot + 1;
  }
  return false;
}

template <int PPT>
__global__ void __launch_bounds__(CTA_DECIM_THREADS, 1) k_decim_cta(const DecimJob* __restrict__ jobs, uint32_t tab_cap, uint32_t bitmap_cap_words) {
  extern __shared__ __align__(16) unsigned char cta_smem[];
  unsigned long long* tab = reinterpret_cast<unsigned long long*>(cta_smem);
  uint32_t* bitmap = reinterpret_cast<uint32_t*>(tab + tab_cap);
  __shared__ uint32_t s_wsum[CTA_DECIM_THREADS / 32];
  __shared__ uint32_t s_npred, s_fail;
  const DecimJob& j = jobs[blockIdx.x];
  const uint32_t n = job_n(j);
  if (n == 0) return;  // (n_out / npred were cleared by the host)
  const uint32_t FULL = 0xFFFFFFFFu;
  const uint32_t tid = threadIdx.x, lane = tid & 31u, warp = tid >> 5;
  const uint32_t words = (n + 31u) >> 5;
  if (words > bitmap_cap_words) {  // (the host sizes the bitmap for the largest cloud of the batch)
    if (tid == 0) atomicOr(j.err, ERR_CTA_FALLBACK);
    return;
  }
  const uint32_t T = min(tab_cap, max(2048u, n * 4u));
  const bool has_pre = j.pre.use_range || j.pre.use_bbox;
  const bool has_post = j.post.use_range || j.post.use_bbox;
  const bool pass_static = n < j.min_pts;  // fewer inputs than minimum_input_points_to_filter: certain pass-through
  if (!pass_static)
    for (uint32_t s = tid; s < T; s += CTA_DECIM_THREADS) tab[s] = CTA_EMPTY;
  for (uint32_t s = tid; s < words; s += CTA_DECIM_THREADS) bitmap[s] = 0u;
  if (tid == 0) {
    s_npred = 0;
    s_fail = 0;
  }
  __syncthreads();

  // ---- pass A: the cloud, once.  Straight-line per point (flags, no early exits) and a __syncwarp after every table
  // insert: the lanes of a warp walk the PPT points of a round together instead of drifting apart after the first
  // divergent probe (measured: 585 M vs 346 M warp instructions for the same work with `continue`-style exits).
  uint32_t my_pred = 0;
  bool fail = false;
  const float res = j.resolution;
  const int fmode = j.index_floor;
  for (uint32_t base = 0; base < n; base += CTA_DECIM_THREADS * PPT) {
    float4 p[PPT];
#pragma unroll
    for (int u = 0; u < PPT; u++) {
      const uint32_t i = base + u * CTA_DECIM_THREADS + tid;
      p[u] = make_float4(0.f, 0.f, 0.f, 0.f);
      if (i < n) p[u] = load_point(j.in, j.in_stride, i);
    }
    uint32_t pk[PPT];
    bool valid[PPT];
#pragma unroll
    for (int u = 0; u < PPT; u++) {
      const uint32_t i = base + u * CTA_DECIM_THREADS + tid;
      const bool pre = (i < n) && (!has_pre || predicate_keep(j.pre, p[u].x, p[u].y, p[u].z));
      const int32_t kx = voxel_index_filter(p[u].x, res, fmode), ky = voxel_index_filter(p[u].y, res, fmode),
                    kz = voxel_index_filter(p[u].z, res, fmode);
      const bool packs = pack_key32(kx, ky, kz, pk[u]);
      fail = fail || (pre && !packs);
      valid[u] = pre && packs;
      my_pred += valid[u] ? 1u : 0u;
    }
    if (pass_static) {
#pragma unroll
      for (int u = 0; u < PPT; u++) {
        const uint32_t i = base + u * CTA_DECIM_THREADS + tid;
        if (valid[u] && (!has_post || predicate_keep(j.post, p[u].x, p[u].y, p[u].z))) atomicOr(&bitmap[i >> 5], 1u << (i & 31u));
      }
    } else {
#pragma unroll
      for (int u = 0; u < PPT; u++) {
        const uint32_t i = base + u * CTA_DECIM_THREADS + tid;
        if (valid[u] && !cta_table_put(tab, T, pk[u], i)) fail = true;
        __syncwarp();
      }
    }
  }
#pragma unroll
  for (int o = 16; o > 0; o >>= 1) my_pred += __shfl_xor_sync(FULL, my_pred, o);
  if (lane == 0 && my_pred) atomicAdd(&s_npred, my_pred);
  if (fail) atomicOr(&s_fail, 1u);
  __syncthreads();
  if (s_fail) {
    if (tid == 0) atomicOr(j.err, ERR_CTA_FALLBACK);
    return;
  }
  const uint32_t npred = s_npred;
  if (!pass_static) {
    if (has_pre && npred < j.min_pts) {
      // pass-through decided by the number of survivors of `pre` (known only now): every survivor that passes `post`
      for (uint32_t i = tid; i < n; i += CTA_DECIM_THREADS) {
        const float4 q = load_point(j.in, j.in_stride, i);
        if (!predicate_keep(j.pre, q.x, q.y, q.z)) continue;
        if (!has_post || predicate_keep(j.post, q.x, q.y, q.z)) atomicOr(&bitmap[i >> 5], 1u << (i & 31u));
      }
    } else {
      // ---- pass B: winners = the table's entries
      for (uint32_t s = tid; s < T; s += CTA_DECIM_THREADS) {
        const unsigned long long e = tab[s];
        if (e == CTA_EMPTY) continue;
        const uint32_t i = uint32_t(e);
        bool keep = true;
        if (has_post) {
          const float4 q = load_point(j.in, j.in_stride, i);
          keep = predicate_keep(j.post, q.x, q.y, q.z);
        }
        if (keep) atomicOr(&bitmap[i >> 5], 1u << (i & 31u));
      }
    }
  }
  __syncthreads();

  // ---- pass C: ranks from the bitmap (thread t owns words [t * wpt, (t + 1) * wpt)), winners out in input order
  const uint32_t wpt = (words + CTA_DECIM_THREADS - 1) / CTA_DECIM_THREADS;
  const uint32_t w0 = min(words, tid * wpt), w1 = min(words, w0 + wpt);
  uint32_t cnt = 0;
  for (uint32_t w = w0; w < w1; w++) cnt += __popc(bitmap[w]);
  uint32_t incl = cnt;
#pragma unroll
  for (int o = 1; o < 32; o <<= 1) {
    const uint32_t y = __shfl_up_sync(FULL, incl, o);
    if (lane >= uint32_t(o)) incl += y;
  }
  if (lane == 31) s_wsum[warp] = incl;
  __syncthreads();
  if (warp == 0) {
    const uint32_t v = s_wsum[lane];
    uint32_t wi = v;
#pragma unroll
    for (int o = 1; o < 32; o <<= 1) {
      const uint32_t y = __shfl_up_sync(FULL, wi, o);
      if (lane >= uint32_t(o)) wi += y;
    }
    s_wsum[lane] = wi - v;
    if (lane == 31) {
      *j.n_out = wi;
      *j.npred = npred;
    }
  }
  __syncthreads();
  uint32_t rank = s_wsum[warp] + incl - cnt;
  for (uint32_t w = w0; w < w1; w++) {
    uint32_t bits = bitmap[w];
    while (bits) {
      const uint32_t i = (w << 5) + uint32_t(__ffs(bits) - 1);
      bits &= bits - 1;
      float4 q = load_point(j.in, j.in_stride, i);
      if (j.in_t) q.w = __ldg(j.in_t + i);
      else if (!j.keep_w) q.w = 0.f;
      j.out[rank] = q;
      if (j.out_idx) j.out_idx[rank] = i;
      rank++;
    }
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
