// icp_block.cuh — the whole mp2p_icp::ICP::align loop of ONE problem inside ONE thread-block cluster
// (call site module/src/LidarOdometry.cpp:961-962; object graph pipelines/lidar3d-default.yaml:162-209).
//
// Why: a single sequence (and every sequence of a lock-step fleet) is a serial chain
//   match -> reduce -> solve -> re-linearise -> reduce -> solve -> stall test          (per ICP iteration)
// of 20-35 iterations.  With work items spread over the grid through a global queue (k_icp_persistent) every arrow of
// that chain is a trip through L2 (partials, queue tickets, fences, problem state): ~41 us per iteration.  Here a
// cluster of 1..8 thread blocks owns the problem from the first iteration to the last:
//   * problem, state and map descriptor live in shared memory;
//   * match = thread per query for the hash probes (9-18 independent 256-bit bucket loads in flight per thread), then
//     ONE block-wide work list of 8-point row segments drained by all warps of the block, four loads in flight per
//     lane: a phase costs one memory round trip per 4 x (threads / 8) segments, however unevenly the candidates are
//     spread over the queries (measured: a thread-per-query scan, one dependent round trip per 8 candidates of the
//     slowest lane, took 57 us per iteration for 653 queries - profiles/README.md);
//   * the 27 normal-equation sums are reduced by a transposing warp reduction (31 shuffles instead of 135), one
//     shared-memory pass, and - across the blocks of the cluster - distributed shared memory: every block stores its
//     partial into block 0's memory, cluster barrier, block 0 sums and solves, stores the new pose into every block's
//     memory, cluster barrier.  No partials, no tickets, no fences in global memory;
//   * the solve runs on block 0's first warp straight out of shared memory (solve_core, icp.cuh).
// A fleet of S sequences uses S clusters, each advancing at its own pace: no queue, no grid barrier, no host round trip.
// The arithmetic per query is that of the other kernels (same candidates, same first-minimum rule, same pruning).
#pragma once
#include <cooperative_groups.h>

#include "icp.cuh"

namespace mlo {
namespace cg = cooperative_groups;

// Transposing warp reduction: on return lane k holds the warp-wide sum of v[k] (k = 0..31).
// Step h halves the live entries: lanes with bit h set keep the upper half and hand the lower half to their
// partner, so the whole reduction costs 16+8+4+2+1 = 31 double shuffles.  Fixed order: reproducible.
MLO_D double warp_reduce32_transpose(double (&v)[32]) {
  const uint32_t lane = threadIdx.x & 31u;
#pragma unroll
  for (int h = 16; h >= 1; h >>= 1) {
    const bool up = (lane & uint32_t(h)) != 0;
#pragma unroll
    for (int k = 0; k < h; k++) {
      const double send = up ? v[k] : v[k + h];
      const double keep = up ? v[k + h] : v[k];
      v[k] = keep + __shfl_xor_sync(0xFFFFFFFFu, send, h);
    }
  }
  return v[0];
}

__device__ __noinline__ int solve_core_ool(const IcpProblem& P, IcpState& S, SolveScratch& sc, int after_match) {
  return solve_core(P, S, sc, after_match);
}

constexpr uint32_t BLK_LIST_CAP = 8192;  // segments per drain window
constexpr uint32_t BLK_MAX_CLUSTER = 8;

template <int NT>
struct BlockShared {
  IcpProblem P;
  IcpState S;       // authoritative copy in block 0 of the cluster
  MapDev map;
  SolveScratch sc;
  double wpart[NT / 32][32];
  double cpart[BLK_MAX_CLUSTER][32];  // block 0: the partials of every block of the cluster
  double T[12];     // this block's copy of the current pose (written by block 0 after every solve)
  int next;         // ... and of the solve's verdict
  uint32_t it;
  // match scratch
  uint32_t words[27][NT];  // packed cell words of each thread's 3x3x3 neighbourhood
  float q[3][NT];
  unsigned long long best[NT];
  uint32_t wsum[NT / 32];
  uint32_t scan_total;
  uint16_t list[BLK_LIST_CAP];
};

// mola::NDT nearest-plane query from the 27 packed cell words already probed into shared memory: same visiting
// order and the same strict '<' as nn_plane_thread, but the 27 hash probes are the batched 256-bit loads of
// probe_words and the per-voxel means are fetched nine at a time instead of one dependent chain per cell.
MLO_D PlaneHit nn_plane_words(const MapDev& m, float qx, float qy, float qz, const uint32_t* ws, uint32_t wstride) {
  PlaneHit r;
  r.cx = r.cy = r.cz = r.nx = r.ny = r.nz = 0.f;
  r.dist = __int_as_float(0x7f800000);
  r.found = 0;
  r.ncand = 0;
#pragma unroll 1
  for (int g = 0; g < 3; g++) {
    float4 mu[9];
    uint32_t vid[9];
#pragma unroll
    for (int u = 0; u < 9; u++) {
      const uint32_t w = ws[(g * 9 + u) * wstride];
      vid[u] = CELL_ABSENT;
      mu[u] = make_float4(0.f, 0.f, 0.f, 0.f);
      if (w != CELL_ABSENT) {
        vid[u] = cell_vid(w);
        mu[u] = __ldg(&m.mean[vid[u]]);
      }
    }
#pragma unroll
    for (int u = 0; u < 9; u++) {
      if (vid[u] == CELL_ABSENT) continue;
      r.ncand += 2;
      if (mu[u].w == 0.f) continue;
      const float4 nr = __ldg(&m.normal[vid[u]]);
      const float ex = qx - mu[u].x, ey = qy - mu[u].y, ez = qz - mu[u].z;
      const float d = fabsf(nr.x * ex + nr.y * ey + nr.z * ez);
      if (d < r.dist) {
        r.dist = d;
        r.cx = mu[u].x; r.cy = mu[u].y; r.cz = mu[u].z;
        r.nx = nr.x; r.ny = nr.y; r.nz = nr.z;
        r.found = 1;
      }
    }
  }
  return r;
}

// Block-wide exclusive scan of one value per thread (two barriers); `total` is returned on every thread.
template <int NT>
MLO_D uint32_t block_exclusive_scan(uint32_t v, BlockShared<NT>& sh, uint32_t& total) {
  const uint32_t lane = threadIdx.x & 31u, warp = threadIdx.x >> 5;
  uint32_t incl = v;
#pragma unroll
  for (int o = 1; o < 32; o <<= 1) {
    const uint32_t y = __shfl_up_sync(0xFFFFFFFFu, incl, o);
    if (lane >= uint32_t(o)) incl += y;
  }
  if (lane == 31) sh.wsum[warp] = incl;
  __syncthreads();
  if (warp == 0) {
    uint32_t w = lane < NT / 32 ? sh.wsum[lane] : 0u;
    uint32_t wi = w;
#pragma unroll
    for (int o = 1; o < 32; o <<= 1) {
      const uint32_t y = __shfl_up_sync(0xFFFFFFFFu, wi, o);
      if (lane >= uint32_t(o)) wi += y;
    }
    if (lane < NT / 32) sh.wsum[lane] = wi - w;  // exclusive over warps
    if (lane == 31) sh.scan_total = wi;
  }
  __syncthreads();
  total = sh.scan_total;
  return sh.wsum[warp] + incl - v;
}

// Drain n_items segments of the block's work list: 8 lanes per segment (one coalesced 128-byte read of 8 stored
// points), NT/8 segments per instruction, four instructions in flight.  Each query's running best is one 64-bit
// (d2 bits << 32 | canonical order) word updated by a shared-memory atomicMin: exactly the sequential first-minimum rule.
// item = seg << 14 | e << 9 | q  (q < 512, e < 27, seg < 4)
template <int NT>
MLO_D void block_list_drain(const MapDev& map, BlockShared<NT>& sh, uint32_t n_items) {
  const uint32_t tid = threadIdx.x, grp = tid >> 3, sub = tid & 7u;
  constexpr uint32_t GROUPS = NT / 8;
  for (uint32_t base = 0; base < n_items; base += GROUPS * 4) {
    float4 p[4];
    uint32_t meta[4];
#pragma unroll
    for (int u = 0; u < 4; u++) {
      const uint32_t idx = base + u * GROUPS + grp;
      meta[u] = 0;
      if (idx < n_items) {
        const uint32_t it = sh.list[idx];
        const uint32_t q = it & 511u, e = (it >> 9) & 31u, slot = (it >> 14) * 8u + sub;
        const uint32_t w = sh.words[e][q];
        meta[u] = q;
        if (slot < cell_cnt(w)) {
          p[u] = __ldg(map.pts + size_t(cell_vid(w)) * map.row + slot);
          meta[u] = 0x80000000u | ((e * 32u + slot) << 9) | q;
        }
      }
    }
#pragma unroll
    for (int u = 0; u < 4; u++) {
      unsigned long long key = ~0ull;
      const uint32_t q = meta[u] & 511u;
      if (meta[u] & 0x80000000u) {
        const float d2 = sqr_dist(p[u].x, p[u].y, p[u].z, sh.q[0][q], sh.q[1][q], sh.q[2][q]);
        key = (uint64_t(__float_as_uint(d2)) << 32) | uint64_t((meta[u] >> 9) & 0x3FFu);
      }
#pragma unroll
      for (int o = 1; o < 8; o <<= 1) {  // segmented min over the 8 lanes of the segment (uniform control flow)
        const unsigned long long other = __shfl_xor_sync(0xFFFFFFFFu, key, o);
        key = other < key ? other : key;
      }
      if (sub == 0 && key != ~0ull) atomicMin(&sh.best[q], key);
    }
  }
}

// Publish this thread's segments (cells in `visit`, canonical order) at block-wide offset `off` and drain the list,
// window by window when the block has more than BLK_LIST_CAP segments.
template <int NT>
MLO_D void block_publish_and_drain(const MapDev& map, BlockShared<NT>& sh, uint32_t visit, uint32_t my_items) {
  uint32_t total;
  const uint32_t off = block_exclusive_scan<NT>(my_items, sh, total);
  const uint32_t tid = threadIdx.x;
  for (uint32_t win = 0; win < total; win += BLK_LIST_CAP) {
    if (off < win + BLK_LIST_CAP && off + my_items > win) {
      uint32_t o = off, m = visit;
      while (m) {
        const uint32_t e = __ffs(m) - 1;
        m &= m - 1;
        const uint32_t np = (cell_cnt(sh.words[e][tid]) + 7u) >> 3;
        for (uint32_t k = 0; k < np; k++, o++)
          if (o >= win && o < win + BLK_LIST_CAP) sh.list[o - win] = uint16_t((k << 14) | (e << 9) | tid);
      }
    }
    __syncthreads();
    block_list_drain<NT>(map, sh, min(BLK_LIST_CAP, total - win));
    __syncthreads();
  }
}

// The match phase of one ICP iteration for the queries of this block: query (pass, thread) = q0 + pass * stride + tid.
// Out of line so that the probe (18 buckets in flight) and the rest of the kernel each get their own register allocation.
template <int NT, bool PLANES>
__device__ __noinline__ uint32_t block_match(BlockShared<NT>& sh, const double* sT, float thr2, float thr_pl, uint32_t q0,
                                             uint32_t qstride, const float4* __restrict__ local, float4* pairA, float4* pairB) {
  const IcpProblem& P = sh.P;
  const MapDev& map = sh.map;
  const uint32_t tid = threadIdx.x;
  const uint64_t qb = P.q_begin;
  const uint32_t nq = P.n_q;
  uint32_t ncand = 0;
  for (uint32_t qbase = q0; qbase < nq; qbase += qstride) {  // (block-uniform trip count)
    const uint32_t q = qbase + tid;
    const bool mine = q < nq;
    float gx = 0.f, gy = 0.f, gz = 0.f;
    float4 pa = make_float4(0.f, 0.f, 0.f, 0.f);
    int32_t kq[3] = {0, 0, 0};
    bool active = false, want = false;
    if (mine) {
      const float4 l = __ldg(&local[qb + q]);
      compose_point_f(sT, l.x, l.y, l.z, gx, gy, gz);
      kq[0] = voxel_index_map(gx, map.inv_voxel, map.index_floor);
      kq[1] = voxel_index_map(gy, map.inv_voxel, map.index_floor);
      kq[2] = voxel_index_map(gz, map.inv_voxel, map.index_floor);
      active = key_in_range(kq[0]) && key_in_range(kq[1]) && key_in_range(kq[2]);
    }
    sh.q[0][tid] = gx;
    sh.q[1][tid] = gy;
    sh.q[2][tid] = gz;
    sh.best[tid] = ~0ull;
    // ---- phase 1: probe (thread per query), the 27 packed cell words land in shared memory
    uint32_t npts = 0;
    if (active) {
      npts = probe_words(map, kq, &sh.words[0][tid], NT);
      bool paired = false;
      if (PLANES && (P.matcher_mask & MLO_MATCHER_PT2PL)) {
        const PlaneHit h = nn_plane_words(map, gx, gy, gz, &sh.words[0][tid], NT);
        ncand += h.ncand;
        if (h.found && h.dist < thr_pl) {
          paired = true;  // Matcher base rule: the point-to-point matcher skips local points already paired
          pa = make_float4(h.cx, h.cy, h.cz, 2.f);
          pairB[qb + q] = make_float4(h.nx, h.ny, h.nz, 0.f);
        }
      }
      want = (P.matcher_mask & MLO_MATCHER_PT2PT) && !paired;
      if (want) ncand += npts;
    }
    if (blockIdx.x == 0) MLO_TRACE_EVENT(0u, 21);  // probes done (this thread)
    // ---- phase 2: own cells
    {
      const uint32_t wh = want ? sh.words[13][tid] : CELL_ABSENT;
      const uint32_t np = (wh == CELL_ABSENT) ? 0u : (cell_cnt(wh) + 7u) >> 3;
      block_publish_and_drain<NT>(map, sh, np ? (1u << 13) : 0u, np);
    }
    if (blockIdx.x == 0) MLO_TRACE_EVENT(0u, 22);  // own cells drained
    // ---- phase 3: per query, the neighbour cells whose box can still beat the bound from the own cell (exact pruning)
    uint32_t visit = 0, my_items = 0;
    if (want) {
      const unsigned long long b0 = sh.best[tid];
      const float bound = (b0 == ~0ull) ? __int_as_float(0x7f800000) : __uint_as_float(uint32_t(b0 >> 32));
      const float qv[3] = {gx, gy, gz};
      const AxisGaps gaps = axis_gaps(map.voxel_size, qv, kq, map.index_floor);
#pragma unroll
      for (int e = 0; e < 27; e++) {
        if (e == 13) continue;
        const uint32_t we = sh.words[e][tid];
        if (we == CELL_ABSENT || cell_cnt(we) == 0) continue;
        if (MLO_LB2(gaps, e) <= bound) {
          visit |= 1u << e;
          my_items += (cell_cnt(we) + 7u) >> 3;
        }
      }
    }
    if (blockIdx.x == 0) MLO_TRACE_EVENT(0u, 23);  // neighbour cells selected
    block_publish_and_drain<NT>(map, sh, visit, my_items);
    if (blockIdx.x == 0) MLO_TRACE_EVENT(0u, 24);  // neighbour cells drained
    // ---- result per query
    if (want) {
      const unsigned long long b = sh.best[tid];
      if (b != ~0ull) {
        const float d2 = __uint_as_float(uint32_t(b >> 32));
        const uint32_t ord = uint32_t(b & 0xFFFFFFFFu), e = ord >> 5, slot = ord & 31u;
        const float lim = thr2 + P.ang2 * (gx * gx + gy * gy + gz * gz);
        if (d2 < lim) {
          const float4 g = __ldg(map.pts + size_t(cell_vid(sh.words[e][tid])) * map.row + slot);
          pa = make_float4(g.x, g.y, g.z, 1.f);
        }
      }
    }
    if (mine) pairA[qb + q] = pa;
    __syncthreads();  // words / best / q are rewritten by the next pass
  }
  return ncand;
}

template <int NT, bool PLANES>
__global__ void __launch_bounds__(NT, 512 / NT)
    k_icp_block(const MapDev* __restrict__ maps, const IcpProblem* __restrict__ probs, IcpState* states,
                const float4* __restrict__ local, float4* pairA, float4* pairB) {
  extern __shared__ __align__(16) unsigned char smem_raw[];
  BlockShared<NT>& sh = *reinterpret_cast<BlockShared<NT>*>(smem_raw);
  cg::cluster_group cluster = cg::this_cluster();
  const uint32_t CL = cluster.num_blocks(), rank = cluster.block_rank();
  const uint32_t tid = threadIdx.x, lane = tid & 31u, warp = tid >> 5;
  const uint32_t prob = blockIdx.x / CL;
  static_assert(sizeof(IcpProblem) % 4 == 0 && sizeof(IcpState) % 4 == 0 && sizeof(MapDev) % 4 == 0, "copied word by word");
  {
    const uint32_t* gp = reinterpret_cast<const uint32_t*>(&probs[prob]);
    const uint32_t* gs = reinterpret_cast<const uint32_t*>(&states[prob]);
    uint32_t* dp = reinterpret_cast<uint32_t*>(&sh.P);
    uint32_t* ds = reinterpret_cast<uint32_t*>(&sh.S);
    for (uint32_t i = tid; i < sizeof(IcpProblem) / 4; i += NT) dp[i] = __ldg(gp + i);
    for (uint32_t i = tid; i < sizeof(IcpState) / 4; i += NT) ds[i] = __ldcg(gs + i);
  }
  __syncthreads();
  if (sh.S.done) return;  // (uniform over the cluster: every block read the same state)
  {
    const uint32_t* gm = reinterpret_cast<const uint32_t*>(&maps[sh.P.map_idx]);
    uint32_t* dm = reinterpret_cast<uint32_t*>(&sh.map);
    for (uint32_t i = tid; i < sizeof(MapDev) / 4; i += NT) dm[i] = __ldg(gm + i);
    if (tid < 12) sh.T[tid] = sh.S.T[tid];
    if (tid == 12) sh.it = sh.S.it;
  }
  __syncthreads();
  if (CL > 1) cluster.sync();  // every block of the cluster is resident before anyone stores into a neighbour's memory
  const IcpProblem& P = sh.P;
  const uint64_t qb = P.q_begin;
  const uint32_t nq = P.n_q;
  const uint32_t q0 = rank * NT, qstride = CL * NT;
  BlockShared<NT>* sh0 = CL > 1 ? cluster.map_shared_rank(&sh, 0) : &sh;
  for (;;) {
    const uint32_t it = sh.it;
    const double* sT = sh.T;
    if (blockIdx.x == 0) MLO_TRACE_EVENT(prob, 11);  // iteration starts
    // ---------------- match: Matcher_Point2Plane, then Matcher_Points_DistanceThreshold on the still unpaired points
    const double thr = table_at(P.thr_pt2pt, P.table_len, it);
    const float thr2 = float(thr * thr);
    const float thr_pl = float(table_at(P.thr_pt2pl, P.table_len, it));
    const double kc = table_at(P.kparam, P.table_len, it);
    uint32_t ncand = block_match<NT, PLANES>(sh, sT, thr2, thr_pl, q0, qstride, local, pairA, pairB);
    if (blockIdx.x == 0) MLO_TRACE_EVENT(prob, 12);  // matches done
    // ---------------- Solver_GaussNewton inner iterations (or the one Horn step) over the stored pairings.
    // Every thread re-reads the records it wrote itself: no barrier between match and accumulate.
    int next;
    int after_match = 1;
    for (;;) {
      double a[32];
#pragma unroll
      for (int k = 0; k < 32; k++) a[k] = 0.0;
      uint32_t npairs = 0;
      for (uint32_t q = q0 + tid; q < nq; q += qstride) {
        const float4 pa = pairA[qb + q];
        if (pa.w == 0.f) continue;
        const float4 l = __ldg(&local[qb + q]);
        if (pa.w == 1.f) {
          if (P.solver == MLO_SOLVER_GAUSS_NEWTON) contrib_pt2pt(sT, l.x, l.y, l.z, pa.x, pa.y, pa.z, P.w_pt2pt, P.robust_kernel, kc, a);
          else contrib_horn(l.x, l.y, l.z, pa.x, pa.y, pa.z, a);
        } else {
          const float4 nb = pairB[qb + q];
          contrib_pt2pl(sT, l.x, l.y, l.z, pa.x, pa.y, pa.z, nb.x, nb.y, nb.z, P.w_pt2pl, P.robust_kernel, kc, a);
        }
        npairs++;
      }
      a[NACC] = double(npairs);  // (exact: counts are far below 2^53)
      a[NACC + 1] = double(ncand);
      const double mine = warp_reduce32_transpose(a);
      sh.wpart[warp][lane] = mine;
      __syncthreads();
      if (blockIdx.x == 0) MLO_TRACE_EVENT(prob, 15);  // accumulated + warp-reduced
      if (warp == 0) {
        double t = sh.wpart[0][lane];
#pragma unroll
        for (int w = 1; w < NT / 32; w++) t += sh.wpart[w][lane];
        sh0->cpart[rank][lane] = t;  // (distributed shared memory when the cluster has several blocks)
      }
      if (CL > 1) cluster.sync();
      else __syncthreads();
      if (blockIdx.x == 0) MLO_TRACE_EVENT(prob, 13);  // linearisation reduced to one partial per block
      if (rank == 0 && warp == 0) {
        double t = sh.cpart[0][lane];
        for (uint32_t r = 1; r < CL; r++) t += sh.cpart[r][lane];
        if (lane < NACC) sh.sc.tot[lane] = t;
        else if (lane < NACC + 2) sh.sc.cnt[lane - NACC] = uint32_t(t);
        __syncwarp();
        const int n = solve_core_ool(sh.P, sh.S, sh.sc, after_match);
        if (blockIdx.x == 0 && lane == 0) trace_event_any(16);  // solve_core returned
        // hand the verdict, the iteration index and the new pose to every block of the cluster
        for (uint32_t r = 0; r < CL; r++) {
          BlockShared<NT>* dst = (CL > 1 && r > 0) ? cluster.map_shared_rank(&sh, r) : &sh;
          if (lane < 12) dst->T[lane] = sh.S.T[lane];
          if (lane == 12) dst->next = n;
          if (lane == 13) dst->it = sh.S.it;
        }
      }
      if (CL > 1) cluster.sync();
      else __syncthreads();
      if (blockIdx.x == 0) MLO_TRACE_EVENT(prob, 14);  // solved
      next = sh.next;
      after_match = 0;
      ncand = 0;
      if (next != 1) break;
    }
    if (next == 0) break;
  }
  if (rank == 0) {
    uint32_t* gs = reinterpret_cast<uint32_t*>(&states[prob]);
    const uint32_t* ds = reinterpret_cast<const uint32_t*>(&sh.S);
    for (uint32_t i = tid; i < sizeof(IcpState) / 4; i += NT) gs[i] = ds[i];
  }
}

}  // namespace mlo
