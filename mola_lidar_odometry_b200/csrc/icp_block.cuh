// icp_block.cuh — the whole mp2p_icp::ICP::align loop of ONE problem inside ONE thread block
// (call site module/src/LidarOdometry.cpp:961-962; object graph pipelines/lidar3d-default.yaml:162-209).
//
// Why: a single sequence (and every sequence of a lock-step fleet) is a serial chain
//   match -> reduce -> solve -> re-linearise -> reduce -> solve -> stall test          (per ICP iteration)
// of 20-35 iterations.  With work items spread over the grid (k_icp_persistent) every arrow of that chain is a
// trip through L2 (partials, queue tickets, fences, problem state): ~41 us per iteration.  Here one block owns
// the problem from the first iteration to the last: problem, state and map descriptor live in shared memory,
// the 27-double reduction is a transposing warp reduction (31 shuffles instead of 135) plus one shared-memory
// pass, the solve runs on the block's first warp straight out of shared memory, and the only global traffic
// is the map itself and the pairing records.  A fleet of S sequences occupies S SMs, each advancing at its own
// pace: no queue, no grid barrier, no host round trip.
//
//   match      thread per query, nn_single_thread / nn_plane_words of map.cuh (exactly the arithmetic of the
//              other kernels: same candidates, same first-minimum rule)
//   accumulate Solver_GaussNewton linearisation over the stored pairings, every inner iteration alike
//   solve      solve_core (icp.cuh): prior, 6x6 LDL^T in registers, retraction, stall / hook tests on three
//              lanes in parallel, termination bookkeeping of ICP::align
#pragma once
#include "icp.cuh"

namespace mlo {

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

template <int NT>
struct BlockShared {
  IcpProblem P;
  IcpState S;
  MapDev map;
  SolveScratch sc;
  double wpart[NT / 32][32];
  int next;
  uint32_t words[27][NT];  // packed cell words of each thread's 3x3x3 neighbourhood (map.cuh nn_single_thread)
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

// The match phase of one ICP iteration for the calling thread's queries (tid, tid + NT, ...): out of line so that the
// probe (18 buckets in flight) and the solve each get their own register allocation under the kernel's 128-register cap.
template <int NT, bool PLANES>
__device__ __noinline__ uint32_t block_match(const IcpProblem& P, const MapDev& map, const double* sT, float thr2, float thr_pl,
                                             uint32_t* ws, const float4* __restrict__ local, float4* pairA, float4* pairB) {
  const uint32_t tid = threadIdx.x;
  const uint64_t qb = P.q_begin;
  const uint32_t nq = P.n_q;
  uint32_t ncand = 0;
  for (uint32_t q = tid; q < nq; q += NT) {
    const float4 l = __ldg(&local[qb + q]);
    float gx, gy, gz;
    compose_point_f(sT, l.x, l.y, l.z, gx, gy, gz);
    float4 pa = make_float4(0.f, 0.f, 0.f, 0.f);
    const int32_t kq[3] = {voxel_index_map(gx, map.inv_voxel), voxel_index_map(gy, map.inv_voxel),
                           voxel_index_map(gz, map.inv_voxel)};
    if (key_in_range(kq[0]) && key_in_range(kq[1]) && key_in_range(kq[2])) {
      // one probe of the 3x3x3 neighbourhood serves both matchers
      const uint32_t npts = probe_words(map, kq, ws, NT);
      bool paired = false;
      if (PLANES && (P.matcher_mask & MLO_MATCHER_PT2PL)) {
        const PlaneHit h = nn_plane_words(map, gx, gy, gz, ws, NT);
        ncand += h.ncand;
        if (h.found && h.dist < thr_pl) {
          paired = true;
          pa = make_float4(h.cx, h.cy, h.cz, 2.f);
          pairB[qb + q] = make_float4(h.nx, h.ny, h.nz, 0.f);
        }
      }
      if ((P.matcher_mask & MLO_MATCHER_PT2PT) && !paired) {
        const NNHit h = nn_scan_words(map, gx, gy, gz, kq, ws, NT);
        ncand += npts;
        const float lim = thr2 + P.ang2 * (gx * gx + gy * gy + gz * gz);
        if (h.found && h.d2 < lim) pa = make_float4(h.x, h.y, h.z, 1.f);
      }
    }
    pairA[qb + q] = pa;
  }
  return ncand;
}

template <int NT, bool PLANES, int MINB = (512 / NT)>
__global__ void __launch_bounds__(NT, MINB)
    k_icp_block(const MapDev* __restrict__ maps, const IcpProblem* __restrict__ probs, IcpState* states,
                const float4* __restrict__ local, float4* pairA, float4* pairB) {
  extern __shared__ __align__(16) unsigned char smem_raw[];
  BlockShared<NT>& sh = *reinterpret_cast<BlockShared<NT>*>(smem_raw);
  const uint32_t tid = threadIdx.x, lane = tid & 31u, warp = tid >> 5;
  const uint32_t prob = blockIdx.x;
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
  if (sh.S.done) return;  // (block-uniform)
  {
    const uint32_t* gm = reinterpret_cast<const uint32_t*>(&maps[sh.P.map_idx]);
    uint32_t* dm = reinterpret_cast<uint32_t*>(&sh.map);
    for (uint32_t i = tid; i < sizeof(MapDev) / 4; i += NT) dm[i] = __ldg(gm + i);
  }
  __syncthreads();
  const IcpProblem& P = sh.P;
  const MapDev& map = sh.map;
  const uint64_t qb = P.q_begin;
  const uint32_t nq = P.n_q;
  for (;;) {
    const uint32_t it = sh.S.it;
    const double* sT = sh.S.T;
    MLO_TRACE_EVENT(prob, 11);  // iteration starts
    // ---------------- match: Matcher_Point2Plane, then Matcher_Points_DistanceThreshold on the still unpaired points
    const double thr = table_at(P.thr_pt2pt, P.table_len, it);
    const float thr2 = float(thr * thr);
    const float thr_pl = float(table_at(P.thr_pt2pl, P.table_len, it));
    const double kc = table_at(P.kparam, P.table_len, it);
    uint32_t ncand = block_match<NT, PLANES>(sh.P, sh.map, sT, thr2, thr_pl, &sh.words[0][tid], local, pairA, pairB);
    MLO_TRACE_EVENT(prob, 12);  // this thread's matches done
    // ---------------- Solver_GaussNewton inner iterations (or the one Horn step) over the stored pairings.
    // Every thread re-reads the records it wrote itself: no barrier between match and accumulate.
    int next;
    int after_match = 1;
    for (;;) {
      double a[32];
#pragma unroll
      for (int k = 0; k < 32; k++) a[k] = 0.0;
      uint32_t npairs = 0;
      for (uint32_t q = tid; q < nq; q += NT) {
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
      MLO_TRACE_EVENT(prob, 13);  // linearisation reduced to one partial per warp
      if (warp == 0) {
        double t = sh.wpart[0][lane];
#pragma unroll
        for (int w = 1; w < NT / 32; w++) t += sh.wpart[w][lane];
        if (lane < NACC) sh.sc.tot[lane] = t;
        else if (lane < NACC + 2) sh.sc.cnt[lane - NACC] = uint32_t(t);
        __syncwarp();
        const int n = solve_core_ool(sh.P, sh.S, sh.sc, after_match);
        if (lane == 0) sh.next = n;
      }
      __syncthreads();
      MLO_TRACE_EVENT(prob, 14);  // solved
      next = sh.next;
      after_match = 0;
      ncand = 0;
      if (next != 1) break;
    }
    if (next == 0) break;
  }
  __syncthreads();
  {
    uint32_t* gs = reinterpret_cast<uint32_t*>(&states[prob]);
    const uint32_t* ds = reinterpret_cast<const uint32_t*>(&sh.S);
    for (uint32_t i = tid; i < sizeof(IcpState) / 4; i += NT) gs[i] = ds[i];
  }
}

}  // namespace mlo
