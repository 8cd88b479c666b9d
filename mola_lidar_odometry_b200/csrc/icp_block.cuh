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

__device__ __noinline__ int solve_core_ool(const IcpProblem& P, IcpState& S, SolveScratch& sc, int after_match) {
  return solve_core(P, S, sc, after_match);
}

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
  uint32_t owords[NT / 8][28];  // match scratch: the 27 packed cell words of each octet's query (+1 pad: no bank conflicts)
};

// ---- octet-per-query match ------------------------------------------------------------------------------------------
// Eight lanes share one query, a warp runs four queries at once.  Why this shape (measured, profiles/README.md): with
// one SM per problem a thread-per-query scan is a chain of dependent round trips (57 us per iteration for 653 queries),
// and a block-wide work list of 8-point segments costs ~50 instructions per candidate slot (ncu: 3.0 M warp
// instructions per align, the SM's issue slots are the limit).  Here
//   probe     the 18 column buckets of the 3x3x3 neighbourhood are spread over the 8 lanes (2-3 independent 256-bit
//             loads per lane, ~50 instructions per lane instead of ~2000 for one thread doing all 27 cells);
//   own cell  lane s reads slots s, s+8, s+16, s+24 of the query's own voxel together (one round trip), the octet's
//             best distance is the pruning bound;
//   cells     every lane tests the box of up to four neighbour cells against the bound (exact pruning, as everywhere);
//             the surviving cells are read two at a time, lane s again taking slots s, s+8, ...: ~15 instructions per
//             candidate, up to 8 loads in flight per lane;
//   result    arg-min over the octet on (d2, canonical order) = the sequential first-minimum rule, bit for bit.
// Lane 0 of the octet then owns the pairing: threshold test, record, and - fused - its Gauss-Newton / Horn contribution.
struct OctetHit {
  float x, y, z, d2;
  uint32_t found, ncand;
};

// lane-local scan of one cell: slots sub, sub+8, sub+16, sub+24; (d2, order) strict first-minimum
MLO_D void octet_scan_cell(const MapDev& m, uint32_t w, uint32_t e, float qx, float qy, float qz, float& bd2, uint32_t& bord,
                           float& bx, float& by, float& bz) {
  const uint32_t sub = threadIdx.x & 7u;
  const uint32_t c = cell_cnt(w);
  const float4* row = m.pts + size_t(cell_vid(w)) * m.row;
  float4 p[4];
#pragma unroll
  for (int r = 0; r < 4; r++)
    if (sub + 8u * r < c) p[r] = __ldg(row + sub + 8u * r);
#pragma unroll
  for (int r = 0; r < 4; r++) {
    const uint32_t slot = sub + 8u * r;
    if (slot < c) {
      const float d2 = sqr_dist(p[r].x, p[r].y, p[r].z, qx, qy, qz);
      const uint32_t ord = e * 32u + slot;
      if (d2 < bd2 || (d2 == bd2 && ord < bord)) {
        bd2 = d2;
        bord = ord;
        bx = p[r].x;
        by = p[r].y;
        bz = p[r].z;
      }
    }
  }
}

// NearestNeighborsCapable::nn_single_search for the octet's query (words already probed).  `want` = this octet has a
// query to search (warp-uniform loops run for the longest octet of the warp).
MLO_D OctetHit octet_nn(const MapDev& m, float qx, float qy, float qz, const int32_t kq[3], bool want, const uint32_t* ow) {
  const uint32_t FULL = 0xFFFFFFFFu;
  const uint32_t lane = threadIdx.x & 31u, sub = lane & 7u, oshift = lane & 24u;  // first lane of this octet in the warp
  float bd2 = __int_as_float(0x7f800000), bx = 0.f, by = 0.f, bz = 0.f;
  uint32_t bord = 0xFFFFFFFFu;
  // own cell first: its best distance is the pruning bound
  const uint32_t wh = want ? ow[13] : CELL_ABSENT;
  if (wh != CELL_ABSENT) octet_scan_cell(m, wh, 13u, qx, qy, qz, bd2, bord, bx, by, bz);
  float bound = bd2;
#pragma unroll
  for (int o = 1; o < 8; o <<= 1) bound = fminf(bound, __shfl_xor_sync(FULL, bound, o));
  // neighbour cells whose box can still hold a point at distance <= bound
  uint32_t visit = 0;
  {
    const float qv[3] = {qx, qy, qz};
    const AxisGaps gaps = axis_gaps(m.voxel_size, qv, kq, m.index_floor);
#pragma unroll
    for (int r = 0; r < 4; r++) {
      const uint32_t e = sub + 8u * r;
      bool pass = false;
      if (want && e < 27u && e != 13u) {
        const uint32_t we = ow[e];
        if (we != CELL_ABSENT && cell_cnt(we) != 0) {
          const float lb = (gaps.g2[0][e / 9] + gaps.g2[1][(e / 3) % 3] + gaps.g2[2][e % 3]) * 0.99999f;
          pass = lb <= bound;
        }
      }
      const uint32_t bal = __ballot_sync(FULL, pass);
      visit |= ((bal >> oshift) & 0xFFu) << (8 * r);
    }
  }
  while (__any_sync(FULL, visit != 0u)) {  // two cells per round: up to 8 row loads in flight per lane
    uint32_t e0 = 0xFFu, e1 = 0xFFu;
    if (visit) {
      e0 = __ffs(visit) - 1;
      visit &= visit - 1;
    }
    if (visit) {
      e1 = __ffs(visit) - 1;
      visit &= visit - 1;
    }
    if (e0 != 0xFFu) octet_scan_cell(m, ow[e0], e0, qx, qy, qz, bd2, bord, bx, by, bz);
    if (e1 != 0xFFu) octet_scan_cell(m, ow[e1], e1, qx, qy, qz, bd2, bord, bx, by, bz);
  }
  // arg-min over the octet on (d2, canonical order)
  const unsigned long long mykey = bord == 0xFFFFFFFFu ? ~0ull : ((uint64_t(__float_as_uint(bd2)) << 32) | uint64_t(bord));
  const unsigned long long best = octet_min_u64(mykey);
  const uint32_t winners = (__ballot_sync(FULL, mykey == best && best != ~0ull) >> oshift) & 0xFFu;
  const uint32_t src = oshift + (winners ? uint32_t(__ffs(winners) - 1) : 0u);
  OctetHit r;
  r.x = __shfl_sync(FULL, bx, src);
  r.y = __shfl_sync(FULL, by, src);
  r.z = __shfl_sync(FULL, bz, src);
  r.d2 = __uint_as_float(uint32_t(best >> 32));
  r.found = best != ~0ull;
  r.ncand = 0;
  return r;
}


// The match phase of one ICP iteration for this block's share of the queries, fused with the first linearisation:
// octet o of block `rank` of a cluster of CL blocks takes queries (o * CL + rank) + k * (NT / 8 * CL).  Lane 0 of the
// octet adds the pairing's normal-equation terms to a[] (27 sums + pairings count in a[27]); returns the candidate count.
template <int NT, bool PLANES>
__device__ __noinline__ uint32_t block_match(BlockShared<NT>& sh, const double* sT, float thr2, float thr_pl, double kc,
                                             uint32_t rank, uint32_t CL, const float4* __restrict__ local, float4* pairA,
                                             float4* pairB, double* a_out) {
  const IcpProblem& P = sh.P;
  double a[NACC];  // (registers; handed to the caller once at the end)
#pragma unroll
  for (int k = 0; k < int(NACC); k++) a[k] = 0.0;
  const MapDev& map = sh.map;
  const uint32_t tid = threadIdx.x, sub = tid & 7u, oct = tid >> 3;
  const uint64_t qb = P.q_begin;
  const uint32_t nq = P.n_q;
  uint32_t* ow = sh.owords[oct];
  uint32_t ncand = 0, npairs = 0;
  for (uint32_t qbase = 0; qbase < nq; qbase += (NT / 8) * CL) {  // (block-uniform trip count)
    const uint32_t q = qbase + oct * CL + rank;
    const bool mine = q < nq;
    float4 l = make_float4(0.f, 0.f, 0.f, 0.f);
    float gx = 0.f, gy = 0.f, gz = 0.f;
    int32_t kq[3] = {0, 0, 0};
    bool in_range = false;
    if (mine) {
      l = __ldg(&local[qb + q]);
      compose_point_f(sT, l.x, l.y, l.z, gx, gy, gz);
      kq[0] = voxel_index_map(gx, map.inv_voxel, map.index_floor);
      kq[1] = voxel_index_map(gy, map.inv_voxel, map.index_floor);
      kq[2] = voxel_index_map(gz, map.inv_voxel, map.index_floor);
      in_range = key_in_range(kq[0]) && key_in_range(kq[1]) && key_in_range(kq[2]);
    }
    __syncwarp();  // (the previous pass has finished reading this octet's words)
    const uint32_t npts = octet_probe(map, kq, in_range, ow);
    float4 pa = make_float4(0.f, 0.f, 0.f, 0.f), pb = make_float4(0.f, 0.f, 0.f, 0.f);
    bool paired = false;
    if (PLANES && (P.matcher_mask & MLO_MATCHER_PT2PL)) {  // (block-uniform)
      const bool wantp = in_range && (P.matcher_mask & MLO_MATCHER_PT2PL);
      const PlaneHit h = octet_plane(map, gx, gy, gz, wantp, ow);
      if (wantp) {
        if (sub == 0) ncand += h.ncand;
        if (h.found && h.dist < thr_pl) {
          paired = true;  // Matcher base rule: the point-to-point matcher skips local points already paired
          pa = make_float4(h.cx, h.cy, h.cz, 2.f);
          pb = make_float4(h.nx, h.ny, h.nz, 0.f);
        }
      }
    }
    const bool want = in_range && (P.matcher_mask & MLO_MATCHER_PT2PT) && !paired;
    const OctetHit h = octet_nn(map, gx, gy, gz, kq, want, ow);
    if (want) {
      if (sub == 0) ncand += npts;
      const float lim = thr2 + P.ang2 * (gx * gx + gy * gy + gz * gz);
      if (h.found && h.d2 < lim) pa = make_float4(h.x, h.y, h.z, 1.f);
    }
    if (mine && sub == 0) {  // lane 0 of the octet owns the pairing: record + first linearisation
      pairA[qb + q] = pa;
      if (pa.w == 1.f) {
        if (P.solver == MLO_SOLVER_GAUSS_NEWTON) contrib_pt2pt(sT, l.x, l.y, l.z, pa.x, pa.y, pa.z, P.w_pt2pt, P.robust_kernel, kc, a);
        else contrib_horn(l.x, l.y, l.z, pa.x, pa.y, pa.z, a);
        npairs++;
      } else if (pa.w == 2.f) {
        pairB[qb + q] = pb;
        contrib_pt2pl(sT, l.x, l.y, l.z, pa.x, pa.y, pa.z, pb.x, pb.y, pb.z, P.w_pt2pl, P.robust_kernel, kc, a);
        npairs++;
      }
    }
  }
#pragma unroll
  for (int k = 0; k < int(NACC); k++) a_out[k] = a[k];
  a_out[NACC] = double(npairs);  // (exact: counts are far below 2^53)
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
    int next;
    int after_match = 1;
    for (;;) {
      double a[32];
#pragma unroll
      for (int k = 0; k < 32; k++) a[k] = 0.0;
      uint32_t ncand = 0;
      if (after_match) {
        // match fused with the first linearisation (lane 0 of every octet contributes its pairing)
        ncand = block_match<NT, PLANES>(sh, sT, thr2, thr_pl, kc, rank, CL, local, pairA, pairB, a);
        if (blockIdx.x == 0) MLO_TRACE_EVENT(prob, 12);  // matches done
      } else {
        // inner Gauss-Newton iterations >= 1: the lane that recorded a pairing re-reads it (its own write)
        uint32_t npairs = 0;
        const uint32_t sub = tid & 7u, oct = tid >> 3;
        if (sub == 0) {
          for (uint32_t q = oct * CL + rank; q < nq; q += (NT / 8) * CL) {
            const float4 pa = pairA[qb + q];
            if (pa.w == 0.f) continue;
            const float4 l = __ldg(&local[qb + q]);
            if (pa.w == 1.f) {
              contrib_pt2pt(sT, l.x, l.y, l.z, pa.x, pa.y, pa.z, P.w_pt2pt, P.robust_kernel, kc, a);
            } else {
              const float4 nb = pairB[qb + q];
              contrib_pt2pl(sT, l.x, l.y, l.z, pa.x, pa.y, pa.z, nb.x, nb.y, nb.z, P.w_pt2pl, P.robust_kernel, kc, a);
            }
            npairs++;
          }
        }
        a[NACC] = double(npairs);
      }
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
