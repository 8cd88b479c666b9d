// icp.cuh — the ICP iteration on device (replaces mp2p_icp::ICP::align as called at
// module/src/LidarOdometry.cpp:961-962 with the object graph of pipelines/lidar3d-default.yaml:162-209):
//
//   k_match_accumulate  Matcher_Points_DistanceThreshold / Matcher_Point2Plane fused with the first
//                       Gauss-Newton linearisation of Solver_GaussNewton: transform, 27-cell NN against
//                       the HBM-resident hash-voxel map, threshold, robust weight, J^T J / J^T r,
//                       warp-shuffle + block reduction -> 27 doubles per block.  Pairings are kept as one
//                       float4 (+ one for planes) per query for the later inner iterations only.
//   k_accumulate        inner Gauss-Newton iterations >= 1 over the stored pairings (no NN).
//   k_solve             one warp per problem: ordered sum of block partials, prior term, 6x6 LDL^T,
//                       retraction T <- T exp(delta), stall / oscillation test, hook-as-data,
//                       termination bookkeeping of ICP::align.
//
// A launch covers a batch of independent problems (blockIdx.y), all against one read-only map.
#pragma once
#include "map.cuh"
#include "octet.cuh"
#include "se3.cuh"

namespace mlo {

constexpr uint32_t ICP_BLOCK = 128;
constexpr uint32_t NACC = 27;  // 21 upper-triangular H entries + 6 gradient entries

struct IcpProblem {
  uint64_t q_begin;  // first local point of this problem in the batch arrays
  uint32_t n_q;
  uint32_t max_iterations;
  double min_abs_step_trans, min_abs_step_rot;
  int32_t solver;
  uint32_t gn_max_iterations;
  double gn_min_delta;
  int32_t robust_kernel;
  uint32_t matcher_mask;
  uint32_t table_len;
  const double* thr_pt2pt;  // device tables
  const double* thr_pt2pl;
  const double* kparam;
  float ang2;  // (thresholdAngularDeg in rad)^2
  double w_pt2pt, w_pt2pl;
  int32_t has_prior;
  double prior_pose[12];
  double prior_info[36];
  int32_t hook_enabled;
  double hook_min_trans, hook_min_rot;
  double hook_checkpoint[12];
  uint32_t part_begin;    // first partial block of this problem
  uint32_t n_blocks;      // blocks of k_match_accumulate (4 warps x qpw queries each)
  uint32_t n_blocks_acc;  // blocks of k_accumulate (ICP_BLOCK queries each)
  uint32_t n_blocks_pers; // match chunks of the persistent kernel (its own chunk geometry)
  uint32_t map_idx;       // index into the launch's map table (fleet launches: one local map per sequence)
  mlo_icp_iteration_record* log;  // per-iteration records of this problem (mlo_icp_log_enable), or nullptr
  uint32_t log_cap;
  uint32_t log_pad;
};

// Fleet launches (independent sequences advanced in lock step) give every problem its own local map: the
// descriptor of problem P is staged from maps[P.map_idx] into shared memory once per block / work item.
MLO_D void stage_map(MapDev& dst, const MapDev* __restrict__ maps, uint32_t idx) {
  static_assert(sizeof(MapDev) % 4 == 0, "MapDev is copied word by word");
  const uint32_t* src = reinterpret_cast<const uint32_t*>(&maps[idx]);
  uint32_t* d = reinterpret_cast<uint32_t*>(&dst);
  if (threadIdx.x < sizeof(MapDev) / 4) d[threadIdx.x] = __ldg(&src[threadIdx.x]);
}

struct IcpState {
  double T[12], prev[12], prev2[12];
  double H[36];
  int32_t has_prev2, done, term, have_H;
  uint32_t it;
  int32_t inner_pending;  // 1: another inner GN iteration must run for the current ICP iteration
  uint32_t inner;         // index of the next inner iteration
  uint64_t n_pairs, n_potential, n_query_it, n_cand;
  // linearisation of the prior term at the CURRENT pose T, prepared ahead of the solve that will use it (solve_phase:
  // a second warp computes it while the first finishes the previous solve): error, right-Jacobian inverse, L e, L J.
  // prior_valid = 0 whenever T has moved since.
  double pe[6], pJ[36], pLe[6], pLJ[36];
  int32_t prior_valid, pad_;
};

MLO_D double table_at(const double* t, uint32_t len, uint32_t it) {
  if (!t || len == 0) return 0.0;
  return t[it < len ? it : len - 1];
}

// `kernel` carries the [VERIFY] form of the Geman-McClure weight in bit 8 (IcpProblem::robust_kernel, set by the host
// from mlo_set_option("convention_gm_form")): 0 = c^4/(c^2+e^2)^2, 1 = c^2/(c^2+e^2)^2.
constexpr int KERNEL_GM_FORM_BIT = 0x100;
MLO_D double robust_weight(int kernel, double e2, double c) {
  if ((kernel & 0xFF) == MLO_KERNEL_GEMAN_MCCLURE) {
    const double c2 = c * c, d = e2 + c2;
    return (kernel & KERNEL_GM_FORM_BIT) ? c2 / (d * d) : (c2 * c2) / (d * d);
  }
  kernel &= 0xFF;
  if (kernel == MLO_KERNEL_CAUCHY) return 1.0 / (1.0 + e2 / (c * c));
  return 1.0;
}

// Normal-equation contribution of one point-to-point pair, in the local frame:
//   r = R l + t - g,  r' = R^T r,  J^T J = [[I, -[l]x], [[l]x, |l|^2 I - l l^T]],  J^T r = [r'; l x r']
// (identical to J = [R | -R [l]x] of Solver_GaussNewton, with R^T R = I applied analytically).
MLO_D void contrib_pt2pt(const double* T, float lxf, float lyf, float lzf, float gxf, float gyf, float gzf, double weight,
                         int kernel, double c, double* a) {
  const double lx = lxf, ly = lyf, lz = lzf;
  const double rx = T[0] * lx + T[1] * ly + T[2] * lz + T[3] - double(gxf);
  const double ry = T[4] * lx + T[5] * ly + T[6] * lz + T[7] - double(gyf);
  const double rz = T[8] * lx + T[9] * ly + T[10] * lz + T[11] - double(gzf);
  const double e2 = rx * rx + ry * ry + rz * rz;
  const double w = weight * robust_weight(kernel, e2, c);
  const double px = T[0] * rx + T[4] * ry + T[8] * rz;  // r' = R^T r
  const double py = T[1] * rx + T[5] * ry + T[9] * rz;
  const double pz = T[2] * rx + T[6] * ry + T[10] * rz;
  const double ll = lx * lx + ly * ly + lz * lz;
  // upper triangle, row-major (i <= j): rows 0..2 translation, 3..5 rotation
  a[0] += w;           a[1] += 0.0;         a[2] += 0.0;         a[3] += 0.0;          a[4] += w * lz;       a[5] += -w * ly;
  /* row 1 */          a[6] += w;           a[7] += 0.0;         a[8] += -w * lz;      a[9] += 0.0;          a[10] += w * lx;
  /* row 2 */                               a[11] += w;          a[12] += w * ly;      a[13] += -w * lx;     a[14] += 0.0;
  /* row 3 */                                                    a[15] += w * (ll - lx * lx); a[16] += -w * lx * ly; a[17] += -w * lx * lz;
  /* row 4 */                                                                          a[18] += w * (ll - ly * ly); a[19] += -w * ly * lz;
  /* row 5 */                                                                                                a[20] += w * (ll - lz * lz);
  a[21] += w * px;
  a[22] += w * py;
  a[23] += w * pz;
  a[24] += w * (ly * pz - lz * py);
  a[25] += w * (lz * px - lx * pz);
  a[26] += w * (lx * py - ly * px);
}

// Point-to-plane pair: r = n.(R l + t - c),  J = [n'; l x n'] with n' = R^T n.
MLO_D void contrib_pt2pl(const double* T, float lxf, float lyf, float lzf, float cxf, float cyf, float czf, float nxf,
                         float nyf, float nzf, double weight, int kernel, double c, double* a) {
  const double lx = lxf, ly = lyf, lz = lzf, nx = nxf, ny = nyf, nz = nzf;
  const double gx = T[0] * lx + T[1] * ly + T[2] * lz + T[3];
  const double gy = T[4] * lx + T[5] * ly + T[6] * lz + T[7];
  const double gz = T[8] * lx + T[9] * ly + T[10] * lz + T[11];
  const double r = nx * (gx - double(cxf)) + ny * (gy - double(cyf)) + nz * (gz - double(czf));
  const double w = weight * robust_weight(kernel, r * r, c);
  double J[6];
  J[0] = T[0] * nx + T[4] * ny + T[8] * nz;
  J[1] = T[1] * nx + T[5] * ny + T[9] * nz;
  J[2] = T[2] * nx + T[6] * ny + T[10] * nz;
  J[3] = ly * J[2] - lz * J[1];
  J[4] = lz * J[0] - lx * J[2];
  J[5] = lx * J[1] - ly * J[0];
  int k = 0;
#pragma unroll
  for (int i = 0; i < 6; i++)
#pragma unroll
    for (int j = i; j < 6; j++) a[k++] += w * J[i] * J[j];
#pragma unroll
  for (int i = 0; i < 6; i++) a[21 + i] += w * J[i] * r;
}

// Solver_Horn sums reuse the same 27-slot vector: [sum l (3), sum g (3), sum l g^T (9)]
MLO_D void contrib_horn(float lx, float ly, float lz, float gx, float gy, float gz, double* a) {
  const double l[3] = {lx, ly, lz}, g[3] = {gx, gy, gz};
  for (int i = 0; i < 3; i++) {
    a[i] += l[i];
    a[3 + i] += g[i];
    for (int j = 0; j < 3; j++) a[6 + 3 * i + j] += l[i] * g[j];
  }
}

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

// The same block reduction with the transposing warp step (31 double shuffles instead of 27 x 5): for the latency-bound
// callers (warp-per-query chunks, re-linearisation chunks of the queue-driven kernel).  Lane 27 / 28 of the array carry
// the pairing / candidate counts (exact in double).  The summation order differs from block_reduce_store's butterfly,
// so a caller must stay with one of the two.
MLO_D void block_reduce_store_t(const double* a, uint32_t npairs, uint32_t ncand, double* part, uint32_t* part_cnt) {
  __shared__ double sm[ICP_BLOCK / 32][32];
  const uint32_t lane = threadIdx.x & 31, warp = threadIdx.x >> 5;
  double v[32];
#pragma unroll
  for (int k = 0; k < int(NACC); k++) v[k] = a[k];
  v[NACC] = double(npairs);
  v[NACC + 1] = double(ncand);
#pragma unroll
  for (int k = int(NACC) + 2; k < 32; k++) v[k] = 0.0;
  sm[warp][lane] = warp_reduce32_transpose(v);
  __syncthreads();
  if (threadIdx.x < NACC + 2) {
    double t = 0.0;
#pragma unroll
    for (int w = 0; w < int(ICP_BLOCK / 32); w++) t += sm[w][threadIdx.x];
    if (threadIdx.x < NACC) part[threadIdx.x] = t;
    else part_cnt[threadIdx.x - NACC] = uint32_t(t);
  }
}

// warp butterfly + ordered cross-warp sum in shared memory; thread k < NACC of the block ends with element k.
MLO_D void block_reduce_store(double* a, uint32_t npairs, uint32_t ncand, double* part, uint32_t* part_cnt) {
  __shared__ double sm[ICP_BLOCK / 32][NACC];
  __shared__ uint32_t smc[ICP_BLOCK / 32][2];
  const uint32_t lane = threadIdx.x & 31, warp = threadIdx.x >> 5;
#pragma unroll
  for (int k = 0; k < int(NACC); k++) {
    double v = a[k];
#pragma unroll
    for (int o = 16; o > 0; o >>= 1) v += __shfl_xor_sync(0xFFFFFFFFu, v, o);
    if (lane == 0) sm[warp][k] = v;
  }
#pragma unroll
  for (int o = 16; o > 0; o >>= 1) {
    npairs += __shfl_xor_sync(0xFFFFFFFFu, npairs, o);
    ncand += __shfl_xor_sync(0xFFFFFFFFu, ncand, o);
  }
  if (lane == 0) {
    smc[warp][0] = npairs;
    smc[warp][1] = ncand;
  }
  __syncthreads();
  if (threadIdx.x < NACC) {
    double s = 0.0;
#pragma unroll
    for (int w = 0; w < int(ICP_BLOCK / 32); w++) s += sm[w][threadIdx.x];
    part[threadIdx.x] = s;
  }
  if (threadIdx.x == 32) {
    uint32_t p = 0, c = 0;
#pragma unroll
    for (int w = 0; w < int(ICP_BLOCK / 32); w++) {
      p += smc[w][0];
      c += smc[w][1];
    }
    part_cnt[0] = p;
    part_cnt[1] = c;
  }
}

// one-warp variant (blocks of 32 threads: no block barrier, every warp writes its own partial)
MLO_D void warp_reduce_store(double* a, uint32_t npairs, uint32_t ncand, double* part, uint32_t* part_cnt) {
  const uint32_t lane = threadIdx.x & 31;
#pragma unroll
  for (int k = 0; k < int(NACC); k++) {
    double v = a[k];
#pragma unroll
    for (int o = 16; o > 0; o >>= 1) v += __shfl_xor_sync(0xFFFFFFFFu, v, o);
    if (lane == uint32_t(k)) part[k] = v;
  }
#pragma unroll
  for (int o = 16; o > 0; o >>= 1) {
    npairs += __shfl_xor_sync(0xFFFFFFFFu, npairs, o);
    ncand += __shfl_xor_sync(0xFFFFFFFFu, ncand, o);
  }
  if (lane == 0) {
    part_cnt[0] = npairs;
    part_cnt[1] = ncand;
  }
}

// pair record: A = (gx, gy, gz | cx, cy, cz, kind) with kind 0 none, 1 pt2pt, 2 pt2pl; B = plane normal.
//
// A "chunk" is the unit of work of one thread block: ICP_BLOCK queries (thread-per-query form) or
// 4 warps x qpw queries (warp-per-query form).  The chunk functions below are shared by the
// one-kernel-per-phase launch sequence and by the persistent queue-driven kernel.

// mola::NDT nearest-plane search for the (up to 32) queries a warp holds one per lane, four queries at a time, eight
// lanes per query (octet.cuh): the 18 column buckets of a neighbourhood are probed by the eight lanes together, then each
// lane reads the mean / normal of up to four cells - three round trips for four queries, where one thread walking the
// 27 cells of its query (nn_plane_thread) is a chain of ~80 dependent loads with four lanes of the warp active.
// Compiled only into the PLANES instantiations of the callers: the point-to-point-only pipelines must not pay for this
// code in the registers of their hot path (present but not executed, it cost the 32-sequence fleet of the default
// pipeline 8 %; as a __noinline__ function it crashes ptxas 12.9).
struct PlanePairing {
  float4 a, b;
  uint32_t ncand;
  bool paired;
};
MLO_D void warp_plane_match(const MapDev& map, float gx, float gy, float gz, uint32_t nq_warp, float thr_pl, uint32_t* ow,
                                              PlanePairing& out) {
  const uint32_t FULL = 0xFFFFFFFFu;
  const uint32_t lane = threadIdx.x & 31u;
  const bool mine = lane < nq_warp;
  out.a = make_float4(0.f, 0.f, 0.f, 0.f);
  out.b = make_float4(0.f, 0.f, 0.f, 0.f);
  out.ncand = 0;
  out.paired = false;
  const uint32_t og = lane >> 3;
  for (uint32_t base = 0; base < nq_warp; base += 4) {  // (nq_warp is warp-uniform)
    const uint32_t src = base + og;  // the query this octet serves
    const bool have = src < nq_warp;
    const float ox = __shfl_sync(FULL, gx, src & 31u), oy = __shfl_sync(FULL, gy, src & 31u), oz = __shfl_sync(FULL, gz, src & 31u);
    int32_t kq[3] = {0, 0, 0};
    bool in_range = false;
    if (have) {
      kq[0] = voxel_index_map(ox, map.inv_voxel, map.index_floor);
      kq[1] = voxel_index_map(oy, map.inv_voxel, map.index_floor);
      kq[2] = voxel_index_map(oz, map.inv_voxel, map.index_floor);
      in_range = key_in_range(kq[0]) && key_in_range(kq[1]) && key_in_range(kq[2]);
    }
    octet_probe(map, kq, in_range, ow);
    const PlaneHit h = octet_plane(map, ox, oy, oz, in_range, ow);
    __syncwarp();
    // hand each result to the lane that owns the query: lane base + j takes it from lane 8 j
    const uint32_t from = lane >= base && lane < base + 4 ? 8u * (lane - base) : 0u;
    const float hcx = __shfl_sync(FULL, h.cx, from), hcy = __shfl_sync(FULL, h.cy, from), hcz = __shfl_sync(FULL, h.cz, from);
    const float hnx = __shfl_sync(FULL, h.nx, from), hny = __shfl_sync(FULL, h.ny, from), hnz = __shfl_sync(FULL, h.nz, from);
    const float hd = __shfl_sync(FULL, h.dist, from);
    const uint32_t hf = __shfl_sync(FULL, h.found, from), hn = __shfl_sync(FULL, h.ncand, from);
    if (mine && lane >= base && lane < base + 4) {
      out.ncand += hn;
      if (hf && hd < thr_pl) {
        out.paired = true;
        out.a = make_float4(hcx, hcy, hcz, 2.f);
        out.b = make_float4(hnx, hny, hnz, 0.f);
      }
    }
  }
}

// Warp-per-query chunk: a warp owns `qpw` consecutive queries.  Lane t holds query t: its local point, its
// transformed point and, after the warp-cooperative NN of that query (map.cuh: probe prefetched one query
// ahead), its pairing.  The normal-equation terms are computed once per query by the owning lane.
template <bool PLANES = true>
MLO_D void chunk_match_warp(const MapDev& map, const IcpProblem& P, const double* sT, uint32_t it, uint32_t chunk,
                            const float4* __restrict__ local, float4* __restrict__ pairA, float4* __restrict__ pairB,
                            double* __restrict__ partials, uint32_t* __restrict__ part_cnt, uint32_t qpw) {
  const uint32_t FULL = 0xFFFFFFFFu;
  const double thr = table_at(P.thr_pt2pt, P.table_len, it);
  const float thr2 = float(thr * thr);
  const float thr_pl = float(table_at(P.thr_pt2pl, P.table_len, it));
  const double kc = table_at(P.kparam, P.table_len, it);
  const uint32_t lane = threadIdx.x & 31u, warp = threadIdx.x >> 5;

  const uint32_t qbase = (chunk * (ICP_BLOCK / 32) + warp) * qpw;  // first query of this warp
  const uint32_t nq_warp = qbase < P.n_q ? min(qpw, P.n_q - qbase) : 0u;
  const bool mine = lane < nq_warp;
  float4 l = make_float4(0.f, 0.f, 0.f, 0.f);
  float gx = 0.f, gy = 0.f, gz = 0.f;
  if (mine) {
    l = __ldg(&local[P.q_begin + qbase + lane]);
    compose_point_f(sT, l.x, l.y, l.z, gx, gy, gz);
  }
  float4 pa = make_float4(0.f, 0.f, 0.f, 0.f);
  float4 pb = make_float4(0.f, 0.f, 0.f, 0.f);
  uint32_t ncand = 0;
  bool paired = false;
  if (PLANES && (P.matcher_mask & MLO_MATCHER_PT2PL)) {  // (warp-uniform)
    __shared__ uint32_t s_ow[ICP_BLOCK / 8][28];  // the 27 cell words of each octet's query
    PlanePairing pp;
    warp_plane_match(map, gx, gy, gz, nq_warp, thr_pl, s_ow[threadIdx.x >> 3], pp);
    ncand += pp.ncand;
    if (mine && pp.paired) {
      paired = true;
      pa = pp.a;
      pb = pp.b;
    }
  }
  if (P.matcher_mask & MLO_MATCHER_PT2PT) {
    // Matcher base rule: local points already paired by an earlier matcher are skipped
    uint32_t todo = __ballot_sync(FULL, mine && !paired);
    WarpProbe cur;
    float cx = 0.f, cy = 0.f, cz = 0.f;
    int t = -1;
    if (todo) {
      t = __ffs(todo) - 1;
      todo &= todo - 1;
      cx = __shfl_sync(FULL, gx, t);
      cy = __shfl_sync(FULL, gy, t);
      cz = __shfl_sync(FULL, gz, t);
      cur = warp_probe_issue(map, cx, cy, cz);
    }
    while (t >= 0) {
      // prefetch the probe of the next query before consuming this one
      int tn = -1;
      float nx = 0.f, ny = 0.f, nz = 0.f;
      WarpProbe nxt;
      if (todo) {
        tn = __ffs(todo) - 1;
        todo &= todo - 1;
        nx = __shfl_sync(FULL, gx, tn);
        ny = __shfl_sync(FULL, gy, tn);
        nz = __shfl_sync(FULL, gz, tn);
        nxt = warp_probe_issue(map, nx, ny, nz);
      }
      const NNHit h = warp_nn_finish(map, cur, cx, cy, cz);
      if (int(lane) == t) {
        ncand += h.ncand;
        const float lim = thr2 + P.ang2 * (gx * gx + gy * gy + gz * gz);
        if (h.found && h.d2 < lim) pa = make_float4(h.x, h.y, h.z, 1.f);
      }
      t = tn;
      if (tn >= 0) {
        cur = nxt;
        cx = nx;
        cy = ny;
        cz = nz;
      }
    }
  }
  double a[NACC];
#pragma unroll
  for (int k = 0; k < int(NACC); k++) a[k] = 0.0;
  uint32_t npairs = 0;
  if (mine) {
    if (pa.w == 1.f) {
      if (P.solver == MLO_SOLVER_GAUSS_NEWTON)
        contrib_pt2pt(sT, l.x, l.y, l.z, pa.x, pa.y, pa.z, P.w_pt2pt, P.robust_kernel, kc, a);
      else
        contrib_horn(l.x, l.y, l.z, pa.x, pa.y, pa.z, a);
      npairs = 1;
    } else if (pa.w == 2.f) {
      contrib_pt2pl(sT, l.x, l.y, l.z, pa.x, pa.y, pa.z, pb.x, pb.y, pb.z, P.w_pt2pl, P.robust_kernel, kc, a);
      npairs = 1;
      pairB[P.q_begin + qbase + lane] = pb;
    }
    pairA[P.q_begin + qbase + lane] = pa;
  }
  const uint32_t pbi = P.part_begin + chunk;
  block_reduce_store_t(a, npairs, ncand, partials + size_t(pbi) * NACC, part_cnt + 2 * size_t(pbi));
}

// Thread-per-query chunk for large batches: one query per thread, map.cuh nn_single_thread (pruned,
// 256-bit loads).  Same outputs and the same block-partial layout as the warp form.
MLO_D void chunk_match_tpq(const MapDev& map, const IcpProblem& P, const double* sT, uint32_t it, uint32_t chunk,
                           const float4* __restrict__ local, float4* __restrict__ pairA, float4* __restrict__ pairB,
                           double* __restrict__ partials, uint32_t* __restrict__ part_cnt) {
  const double thr = table_at(P.thr_pt2pt, P.table_len, it);
  const float thr2 = float(thr * thr);
  const float thr_pl = float(table_at(P.thr_pt2pl, P.table_len, it));
  const double kc = table_at(P.kparam, P.table_len, it);
  __shared__ uint32_t s_words[27][ICP_BLOCK];  // per-thread packed cell words of the 3x3x3 neighbourhood
  double a[NACC];
#pragma unroll
  for (int k = 0; k < int(NACC); k++) a[k] = 0.0;
  uint32_t npairs = 0, ncand = 0;
  const uint32_t q = chunk * ICP_BLOCK + threadIdx.x;
  if (q < P.n_q) {
    const float4 l = __ldg(&local[P.q_begin + q]);
    float gx, gy, gz;
    compose_point_f(sT, l.x, l.y, l.z, gx, gy, gz);
    float4 pa = make_float4(0.f, 0.f, 0.f, 0.f);
    bool paired = false;
    if (P.matcher_mask & MLO_MATCHER_PT2PL) {
      const PlaneHit h = nn_plane_thread(map, gx, gy, gz);
      ncand += h.ncand;
      if (h.found && h.dist < thr_pl) {
        paired = true;
        pa = make_float4(h.cx, h.cy, h.cz, 2.f);
        pairB[P.q_begin + q] = make_float4(h.nx, h.ny, h.nz, 0.f);
        contrib_pt2pl(sT, l.x, l.y, l.z, h.cx, h.cy, h.cz, h.nx, h.ny, h.nz, P.w_pt2pl, P.robust_kernel, kc, a);
        npairs++;
      }
    }
    if ((P.matcher_mask & MLO_MATCHER_PT2PT) && !paired) {
      const NNHit h = nn_single_thread(map, gx, gy, gz, &s_words[0][threadIdx.x], ICP_BLOCK);
      ncand += h.ncand;
      const float lim = thr2 + P.ang2 * (gx * gx + gy * gy + gz * gz);
      if (h.found && h.d2 < lim) {
        pa = make_float4(h.x, h.y, h.z, 1.f);
        if (P.solver == MLO_SOLVER_GAUSS_NEWTON)
          contrib_pt2pt(sT, l.x, l.y, l.z, h.x, h.y, h.z, P.w_pt2pt, P.robust_kernel, kc, a);
        else
          contrib_horn(l.x, l.y, l.z, h.x, h.y, h.z, a);
        npairs++;
      }
    }
    pairA[P.q_begin + q] = pa;
  }
  const uint32_t pbi = P.part_begin + chunk;
  block_reduce_store(a, npairs, ncand, partials + size_t(pbi) * NACC, part_cnt + 2 * size_t(pbi));
}

// Work-list chunk (large batches): probes are thread-per-query (each lane keeps 9-18 independent 256-bit
// bucket loads in flight), the 27 packed cell words of every query land in shared memory, and the candidate
// gather is drained COOPERATIVELY from a flattened per-warp list of 8-point row segments: 8 lanes per segment
// (one coalesced 128-byte read), 4 segments per warp instruction, 4 instructions in flight.  Each query's
// running best is one 64-bit (d2 bits << 32 | canonical order) word updated by a shared-memory atomicMin, which
// is exactly the sequential first-minimum rule.  Exact pruning as in map.cuh: own cell first, then only the
// neighbour cells whose box can still beat that bound.
constexpr uint32_t WL_CAP = 768;
struct WarpScratch {
  uint32_t words[27][32];
  float q[3][32];
  unsigned long long best[32];
  uint16_t list[WL_CAP];
  uint32_t bd2[32], bord[32];  // the split form of `best` used by wl_process_a32: distance bits / visiting order
};

// one batch of the drain: 4 segments per lane group (16 per warp); issue the loads of [base, base + 16)
MLO_D void wl_issue(const MapDev& map, const WarpScratch& ws, uint32_t n, uint32_t base, uint32_t grp, uint32_t sub, float4 (&p)[4],
                    uint32_t (&meta)[4]) {
#pragma unroll
  for (int u = 0; u < 4; u++) {
    const uint32_t idx = base + u * 4 + grp;
    meta[u] = 0;  // (valid << 31) | (order << 5) | q, order = e*32 + slot; q is shared by the 8 lanes of a segment
    if (idx < n) {
      const uint32_t it = ws.list[idx];
      const uint32_t q = it & 31u, e = (it >> 5) & 31u, slot = (it >> 10) * 8u + sub;
      const uint32_t w = ws.words[e][q];
      meta[u] = q;
      if (slot < cell_cnt(w)) {
        p[u] = __ldg(map.pts + size_t(cell_vid(w)) * map.row + slot);
        meta[u] = 0x80000000u | ((e * 32u + slot) << 5) | q;
      }
    }
  }
}
MLO_D void wl_consume(WarpScratch& ws, uint32_t sub, const float4 (&p)[4], const uint32_t (&meta)[4]) {
#pragma unroll
  for (int u = 0; u < 4; u++) {
    unsigned long long key = ~0ull;
    const uint32_t q = meta[u] & 31u;
    if (meta[u] & 0x80000000u) {
      const float d2 = sqr_dist(p[u].x, p[u].y, p[u].z, ws.q[0][q], ws.q[1][q], ws.q[2][q]);
      key = (uint64_t(__float_as_uint(d2)) << 32) | uint64_t((meta[u] >> 5) & 0x3FFu);
    }
#pragma unroll
    for (int o = 1; o < 8; o <<= 1) {  // segmented min over the 8 lanes of the segment (uniform control flow)
      const unsigned long long other = __shfl_xor_sync(0xFFFFFFFFu, key, o);
      key = other < key ? other : key;
    }
    if (sub == 0 && key != ~0ull) atomicMin(&ws.best[q], key);
  }
}

// PIPE = false: issue 16 segments, wait, reduce, repeat.  PIPE = true: the loads of the next 16 segments are issued
// before the current 16 are reduced (two batches of registers), so a warp always has 4-8 row loads in flight.
template <bool PIPE>
MLO_D void wl_process(const MapDev& map, WarpScratch& ws, uint32_t n) {
  const uint32_t lane = threadIdx.x & 31u, grp = lane >> 3, sub = lane & 7u;
  if constexpr (!PIPE) {
    for (uint32_t base = 0; base < n; base += 16) {
      float4 p[4];
      uint32_t meta[4];
      wl_issue(map, ws, n, base, grp, sub, p, meta);
      wl_consume(ws, sub, p, meta);
    }
  } else {
    float4 p[4], pn[4];
    uint32_t meta[4], mn[4];
    if (n) wl_issue(map, ws, n, 0, grp, sub, p, meta);
    for (uint32_t base = 0; base < n; base += 16) {
      const bool more = base + 16 < n;
      if (more) wl_issue(map, ws, n, base + 16, grp, sub, pn, mn);
      wl_consume(ws, sub, p, meta);
      if (more) {
#pragma unroll
        for (int u = 0; u < 4; u++) {
          p[u] = pn[u];
          meta[u] = mn[u];
        }
      }
    }
  }
  __syncwarp();
}

// ---- run-merging drain (MLO_WL_VARIANT 18 / 19): consecutive work-list items mostly belong to the SAME query (a query
// brings ~13 row segments), so the four lane groups of one instruction usually update the same 64-bit best and their
// shared-memory atomicMin CAS loops collide.  Here the groups first merge their segment minima along the run of equal
// queries (two conditional shuffle steps towards the run's first group) and only that group touches shared memory.
MLO_D void wl_consume_runs(WarpScratch& ws, uint32_t lane, const float4 (&p)[4], const uint32_t (&meta)[4]) {
  const uint32_t FULL = 0xFFFFFFFFu;
  const uint32_t sub = lane & 7u, grp = lane >> 3;
#pragma unroll
  for (int u = 0; u < 4; u++) {
    unsigned long long key = ~0ull;
    const uint32_t q = meta[u] & 31u;
    if (meta[u] & 0x80000000u) {
      const float d2 = sqr_dist(p[u].x, p[u].y, p[u].z, ws.q[0][q], ws.q[1][q], ws.q[2][q]);
      key = (uint64_t(__float_as_uint(d2)) << 32) | uint64_t((meta[u] >> 5) & 0x3FFu);
    }
#pragma unroll
    for (int o = 1; o < 8; o <<= 1) {
      const unsigned long long other = __shfl_xor_sync(FULL, key, o);
      key = other < key ? other : key;
    }
    // merge along the run of groups that hold the same query (items are sorted by query: runs are contiguous)
    const uint32_t q_prev = __shfl_up_sync(FULL, q, 8);
#pragma unroll
    for (int o = 8; o <= 16; o <<= 1) {
      const unsigned long long kd = __shfl_down_sync(FULL, key, o);
      const uint32_t qd = __shfl_down_sync(FULL, q, o);
      if (qd == q && kd < key) key = kd;  // (beyond the warp's edge a lane reads itself: a no-op)
    }
    const bool first_of_run = grp == 0 || q_prev != q;
    if (sub == 0 && first_of_run && key != ~0ull) atomicMin(&ws.best[q], key);
  }
}
MLO_D void wl_process_runs(const MapDev& map, WarpScratch& ws, uint32_t n) {
  const uint32_t lane = threadIdx.x & 31u, grp = lane >> 3, sub = lane & 7u;
  for (uint32_t base = 0; base < n; base += 16) {
    float4 p[4];
    uint32_t meta[4];
    wl_issue(map, ws, n, base, grp, sub, p, meta);
    wl_consume_runs(ws, lane, p, meta);
  }
  __syncwarp();
}

// ---- deep drain (MLO_WL_VARIANT 16 / 17): the default's issue / consume pair with EIGHT row segments per lane group and
// round (32 per warp) instead of four: half as many dependent rounds per work list, twice the row loads in flight, at
// the price of 16 more registers (96 at 5 blocks/SM, or 80 with spills at 6).
MLO_D void wl_process_deep(const MapDev& map, WarpScratch& ws, uint32_t n) {
  const uint32_t lane = threadIdx.x & 31u, grp = lane >> 3, sub = lane & 7u;
  for (uint32_t base = 0; base < n; base += 32) {
    float4 p[8];
    uint32_t meta[8];
#pragma unroll
    for (int u = 0; u < 8; u++) {
      const uint32_t idx = base + u * 4 + grp;
      meta[u] = 0;
      if (idx < n) {
        const uint32_t it = ws.list[idx];
        const uint32_t q = it & 31u, e = (it >> 5) & 31u, slot = (it >> 10) * 8u + sub;
        const uint32_t w = ws.words[e][q];
        meta[u] = q;
        if (slot < cell_cnt(w)) {
          p[u] = __ldg(map.pts + size_t(cell_vid(w)) * map.row + slot);
          meta[u] = 0x80000000u | ((e * 32u + slot) << 5) | q;
        }
      }
    }
#pragma unroll
    for (int u = 0; u < 8; u++) {
      unsigned long long key = ~0ull;
      const uint32_t q = meta[u] & 31u;
      if (meta[u] & 0x80000000u) {
        const float d2 = sqr_dist(p[u].x, p[u].y, p[u].z, ws.q[0][q], ws.q[1][q], ws.q[2][q]);
        key = (uint64_t(__float_as_uint(d2)) << 32) | uint64_t((meta[u] >> 5) & 0x3FFu);
      }
#pragma unroll
      for (int o = 1; o < 8; o <<= 1) {
        const unsigned long long other = __shfl_xor_sync(0xFFFFFFFFu, key, o);
        key = other < key ? other : key;
      }
      if (sub == 0 && key != ~0ull) atomicMin(&ws.best[q], key);
    }
  }
  __syncwarp();
}

// ---- ballot drain (MLO_WL_VARIANT 14 / 15): as wl_consume, but the minimum over the 8 lanes of a segment is taken on the
// 32-bit distance alone (3 x SHFL + min instead of 3 x (2 SHFL + 64-bit compare + 2 selects)) and the lane that holds it -
// the lowest such lane: lowest slot = lowest order within a segment - is found by a ballot and updates the query's 64-bit
// best itself; no key travels between lanes.  Each segment is still consumed as soon as its own load has arrived.
MLO_D void wl_consume_ballot(WarpScratch& ws, uint32_t lane, const float4 (&p)[4], const uint32_t (&meta)[4]) {
  const uint32_t FULL = 0xFFFFFFFFu;
  const uint32_t sub = lane & 7u, oshift = lane & 24u;
#pragma unroll
  for (int u = 0; u < 4; u++) {
    const uint32_t q = meta[u] & 31u;
    const bool valid = (meta[u] & 0x80000000u) != 0;
    uint32_t d2b = 0xFFFFFFFFu;
    if (valid) d2b = __float_as_uint(sqr_dist(p[u].x, p[u].y, p[u].z, ws.q[0][q], ws.q[1][q], ws.q[2][q]));
    uint32_t m = d2b;
#pragma unroll
    for (int o = 1; o < 8; o <<= 1) m = min(m, __shfl_xor_sync(FULL, m, o));
    const bool is_min = valid && d2b == m;
    const uint32_t b = (__ballot_sync(FULL, is_min) >> oshift) & 0xFFu;
    if (is_min && sub == uint32_t(__ffs(b) - 1))
      atomicMin(&ws.best[q], (uint64_t(d2b) << 32) | uint64_t((meta[u] >> 5) & 0x3FFu));
  }
}
MLO_D void wl_process_ballot(const MapDev& map, WarpScratch& ws, uint32_t n) {
  const uint32_t lane = threadIdx.x & 31u, grp = lane >> 3, sub = lane & 7u;
  for (uint32_t base = 0; base < n; base += 16) {
    float4 p[4];
    uint32_t meta[4];
    wl_issue(map, ws, n, base, grp, sub, p, meta);
    wl_consume_ballot(ws, lane, p, meta);
  }
  __syncwarp();
}

// ---- split-key drain (MLO_WL_VARIANT 12 / 13): the running best of a query is kept as TWO 32-bit words, distance bits
// and visiting order, so that every update is a native 32-bit shared-memory atomicMin (ATOMS.MIN) instead of the CAS
// loop a 64-bit shared atomicMin compiles to (ATOMS.CAST.SPIN: ~20 instructions, 15 % of the kernel's instruction
// stream), and the minimum over the 8 lanes of a segment is taken on the 32-bit distance alone (lowest lane among equal
// distances = lowest slot = lowest order within a segment).  Exactly the 64-bit rule:
//   1  d2 goes to bd2[q] by atomicMin, which returns the previous value;
//   2  after a __syncwarp, a candidate whose d2 equals bd2[q] is a current minimum; if it strictly lowered the word it is
//      the FIRST to reach that value and resets bord[q] (orders recorded for a larger distance are obsolete);
//   3  after another __syncwarp every current minimum offers its order to bord[q] by atomicMin.
// Candidates that tie on the distance - in the same round or rounds apart - therefore end with the smallest order.
MLO_D void wl_consume_a32(WarpScratch& ws, uint32_t grp, uint32_t sub, const float4 (&p)[4], const uint32_t (&meta)[4]) {
  const uint32_t FULL = 0xFFFFFFFFu;
  uint32_t d2b[4], old[4];
  bool win[4];
#pragma unroll
  for (int u = 0; u < 4; u++) {
    const uint32_t q = meta[u] & 31u;
    const bool valid = (meta[u] & 0x80000000u) != 0;
    d2b[u] = 0xFFFFFFFFu;
    if (valid) d2b[u] = __float_as_uint(sqr_dist(p[u].x, p[u].y, p[u].z, ws.q[0][q], ws.q[1][q], ws.q[2][q]));
    uint32_t m = d2b[u];
#pragma unroll
    for (int o = 1; o < 8; o <<= 1) m = min(m, __shfl_xor_sync(FULL, m, o));
    const bool is_min = valid && d2b[u] == m;
    const uint32_t b = (__ballot_sync(FULL, is_min) >> (8u * grp)) & 0xFFu;
    win[u] = is_min && sub == uint32_t(__ffs(b) - 1);
  }
#pragma unroll
  for (int u = 0; u < 4; u++) {
    old[u] = 0;
    if (win[u]) old[u] = atomicMin(&ws.bd2[meta[u] & 31u], d2b[u]);
  }
  __syncwarp();
#pragma unroll
  for (int u = 0; u < 4; u++) {
    if (win[u]) {
      const uint32_t q = meta[u] & 31u;
      win[u] = ws.bd2[q] == d2b[u];
      if (win[u] && old[u] > d2b[u]) ws.bord[q] = 0xFFFFFFFFu;
    }
  }
  __syncwarp();
#pragma unroll
  for (int u = 0; u < 4; u++)
    if (win[u]) atomicMin(&ws.bord[meta[u] & 31u], (meta[u] >> 5) & 0x3FFu);
}
MLO_D void wl_process_a32(const MapDev& map, WarpScratch& ws, uint32_t n) {
  const uint32_t lane = threadIdx.x & 31u, grp = lane >> 3, sub = lane & 7u;
  for (uint32_t base = 0; base < n; base += 16) {
    float4 p[4];
    uint32_t meta[4];
    wl_issue(map, ws, n, base, grp, sub, p, meta);
    wl_consume_a32(ws, grp, sub, p, meta);
  }
  __syncwarp();
  // fold the split words back into the 64-bit form the rest of the chunk reads
  ws.best[lane] = ws.bd2[lane] == 0xFFFFFFFFu ? ~0ull : ((uint64_t(ws.bd2[lane]) << 32) | uint64_t(ws.bord[lane]));
  __syncwarp();
}

// ---- contiguous-range drain (MLO_WL_VARIANT 6-8): each group of 8 lanes owns a CONTIGUOUS quarter of the list, so
// consecutive items of a group mostly belong to the same query: the running best of that query stays in a register of
// each lane (one 64-bit compare per candidate point) and the 8 lanes are merged (3 shuffle steps + one shared-memory
// atomicMin) only when a group moves on to the next query - not after every segment as in wl_consume.  The merge runs
// under a warp-uniform vote, i.e. for all four groups whenever any of them crosses a query boundary.  Same keys, same
// minimum: bit-identical pairings.  DEPTH = row loads in flight per lane.
MLO_D void wl_oct_flush(WarpScratch& ws, uint32_t sub, bool flush, uint32_t cur_q, unsigned long long& best) {
  unsigned long long key = best;
#pragma unroll
  for (int o = 1; o < 8; o <<= 1) {
    const unsigned long long other = __shfl_xor_sync(0xFFFFFFFFu, key, o);
    key = other < key ? other : key;
  }
  if (flush) {
    if (sub == 0 && key != ~0ull) atomicMin(&ws.best[cur_q], key);
    best = ~0ull;
  }
}
template <int DEPTH>
MLO_D void wl_process_oct(const MapDev& map, WarpScratch& ws, uint32_t n) {
  const uint32_t FULL = 0xFFFFFFFFu;
  const uint32_t lane = threadIdx.x & 31u, grp = lane >> 3, sub = lane & 7u;
  const uint32_t per = (n + 3u) >> 2;  // items per group
  const uint32_t g0 = grp * per, g1 = min(n, g0 + per);
  uint32_t cur_q = 32u;  // none yet
  float qx = 0.f, qy = 0.f, qz = 0.f;
  unsigned long long best = ~0ull;
  for (uint32_t r = 0; r < per; r += DEPTH) {
    float4 p[DEPTH];
    uint32_t meta[DEPTH];  // (item valid << 31) | (point valid << 30) | (order << 5) | q
#pragma unroll
    for (int u = 0; u < DEPTH; u++) {
      const uint32_t idx = g0 + r + u;
      meta[u] = 0;
      if (idx < g1) {
        const uint32_t it = ws.list[idx];
        const uint32_t q = it & 31u, e = (it >> 5) & 31u, slot = (it >> 10) * 8u + sub;
        const uint32_t w = ws.words[e][q];
        meta[u] = 0x80000000u | q;
        if (slot < cell_cnt(w)) {
          p[u] = __ldg(map.pts + size_t(cell_vid(w)) * map.row + slot);
          meta[u] = 0xC0000000u | ((e * 32u + slot) << 5) | q;
        }
      }
    }
#pragma unroll
    for (int u = 0; u < DEPTH; u++) {
      const uint32_t q = meta[u] & 31u;
      const bool moved = (meta[u] & 0x80000000u) && q != cur_q;
      if (__any_sync(FULL, moved)) {  // (warp-uniform)
        wl_oct_flush(ws, sub, moved && cur_q < 32u, cur_q, best);
        if (moved) {
          cur_q = q;
          qx = ws.q[0][q];
          qy = ws.q[1][q];
          qz = ws.q[2][q];
        }
      }
      if (meta[u] & 0x40000000u) {
        const float d2 = sqr_dist(p[u].x, p[u].y, p[u].z, qx, qy, qz);
        const unsigned long long key = (uint64_t(__float_as_uint(d2)) << 32) | uint64_t((meta[u] >> 5) & 0x3FFu);
        best = key < best ? key : best;
      }
    }
  }
  wl_oct_flush(ws, sub, cur_q < 32u, cur_q, best);
  __syncwarp();
}

// ---- bulk-async variant of the drain (A/B, MLO_WL_VARIANT=4): the 8-point row segments of a round are fetched by
// cp.async.bulk (the TMA unit, SASS UBLKCP) into a per-warp shared-memory stage, completion on an mbarrier, and the next
// round's copies are in flight while the current round is reduced - the software pipeline of PIPE without its register
// cost (rows never sit in registers while in flight).  16 segments of <= 128 bytes per round and warp.
struct WarpStage {
  float4 pts[2][16][8];
  unsigned long long bar[2];
};
MLO_D uint32_t smem_u32(const void* p) { return uint32_t(__cvta_generic_to_shared(p)); }
MLO_D void mbar_init(unsigned long long* bar, uint32_t count) {
  asm volatile("mbarrier.init.shared::cta.b64 [%0], %1;" ::"r"(smem_u32(bar)), "r"(count));
}
MLO_D void mbar_expect_tx(unsigned long long* bar, uint32_t bytes) {
  asm volatile("mbarrier.arrive.expect_tx.shared::cta.b64 _, [%0], %1;" ::"r"(smem_u32(bar)), "r"(bytes) : "memory");
}
MLO_D bool mbar_try_wait(unsigned long long* bar, uint32_t parity) {
  uint32_t ok;
  asm volatile(
      "{\n\t.reg .pred p;\n\tmbarrier.try_wait.parity.shared::cta.b64 p, [%1], %2;\n\tselp.u32 %0, 1, 0, p;\n\t}"
      : "=r"(ok)
      : "r"(smem_u32(bar)), "r"(parity)
      : "memory");
  return ok != 0;
}
MLO_D void bulk_g2s(void* dst_smem, const void* src_gmem, uint32_t bytes, unsigned long long* bar) {
  asm volatile("cp.async.bulk.shared::cluster.global.mbarrier::complete_tx::bytes [%0], [%1], %2, [%3];" ::"r"(smem_u32(dst_smem)),
               "l"(src_gmem), "r"(bytes), "r"(smem_u32(bar))
               : "memory");
}
MLO_D void wl_bulk_issue(const MapDev& map, const WarpScratch& ws, WarpStage& st, uint32_t n, uint32_t base, int s) {
  const uint32_t lane = threadIdx.x & 31u;
  uint32_t bytes = 0;
  const float4* src = nullptr;
  if (lane < 16 && base + lane < n) {
    const uint32_t it = ws.list[base + lane];
    const uint32_t q = it & 31u, e = (it >> 5) & 31u, k = it >> 10;
    const uint32_t w = ws.words[e][q];
    const uint32_t c = cell_cnt(w);
    const uint32_t npts = c > k * 8u ? min(8u, c - k * 8u) : 0u;
    bytes = npts * 16u;
    src = map.pts + size_t(cell_vid(w)) * map.row + k * 8u;
  }
  uint32_t total = bytes;
#pragma unroll
  for (int o = 16; o > 0; o >>= 1) total += __shfl_xor_sync(0xFFFFFFFFu, total, o);
  if (lane == 0) mbar_expect_tx(&st.bar[s], total);  // (one arrival per phase; the copies complete its byte count)
  __syncwarp();
  if (bytes) bulk_g2s(&st.pts[s][lane][0], src, bytes, &st.bar[s]);
}
MLO_D void wl_bulk_consume(WarpScratch& ws, WarpStage& st, uint32_t n, uint32_t base, int s, uint32_t parity) {
  const uint32_t lane = threadIdx.x & 31u, grp = lane >> 3, sub = lane & 7u;
  uint32_t spins = 0;
  while (!mbar_try_wait(&st.bar[s], parity)) {
    if (++spins > (1u << 22)) break;  // safety net: never hang the device
  }
#pragma unroll
  for (int u = 0; u < 4; u++) {
    const uint32_t j = u * 4 + grp, idx = base + j;
    unsigned long long key = ~0ull;
    uint32_t q = 0;
    if (idx < n) {
      const uint32_t it = ws.list[idx];
      q = it & 31u;
      const uint32_t e = (it >> 5) & 31u, slot = (it >> 10) * 8u + sub;
      if (slot < cell_cnt(ws.words[e][q])) {
        const float4 p = st.pts[s][j][sub];
        const float d2 = sqr_dist(p.x, p.y, p.z, ws.q[0][q], ws.q[1][q], ws.q[2][q]);
        key = (uint64_t(__float_as_uint(d2)) << 32) | uint64_t(e * 32u + slot);
      }
    }
#pragma unroll
    for (int o = 1; o < 8; o <<= 1) {
      const unsigned long long other = __shfl_xor_sync(0xFFFFFFFFu, key, o);
      key = other < key ? other : key;
    }
    if (sub == 0 && key != ~0ull) atomicMin(&ws.best[q], key);
  }
}
// `par` = the phase parities of the two stage barriers of this warp (bit s), carried across calls
MLO_D void wl_process_bulk(const MapDev& map, WarpScratch& ws, WarpStage& st, uint32_t n, uint32_t& par) {
  if (n) wl_bulk_issue(map, ws, st, n, 0, 0);
  int s = 0;
  for (uint32_t base = 0; base < n; base += 16, s ^= 1) {
    if (base + 16 < n) {
      // the other stage was read (generic proxy) two rounds ago: order those reads before the async-proxy writes
      asm volatile("fence.proxy.async.shared::cta;" ::: "memory");
      __syncwarp();
      wl_bulk_issue(map, ws, st, n, base + 16, s ^ 1);
    }
    wl_bulk_consume(ws, st, n, base, s, (par >> s) & 1u);
    par ^= 1u << s;
    __syncwarp();
  }
  // an odd number of rounds leaves the next call starting on stage 0 again: parities are per stage, so nothing to fix up
  __syncwarp();
}

constexpr uint32_t WL_BLOCK = 32;  // the work-list kernel runs one warp per block: a chunk is 32 queries
template <int NWARPS, bool PIPE = false, bool BULK = false, int OCT = 0, bool WPART = false, int A32 = 0>
MLO_D void chunk_match_wl(const MapDev& map, const IcpProblem& P, const double* sT, uint32_t it, uint32_t chunk,
                          const float4* __restrict__ local, float4* __restrict__ pairA, float4* __restrict__ pairB,
                          double* __restrict__ partials, uint32_t* __restrict__ part_cnt) {
  __shared__ WarpScratch s_ws[NWARPS];
  __shared__ __align__(128) WarpStage s_stage[BULK ? NWARPS : 1];
  uint32_t bulk_par = 0;
  if constexpr (BULK) {
    if ((threadIdx.x & 31u) == 0) {
      mbar_init(&s_stage[threadIdx.x >> 5].bar[0], 1);
      mbar_init(&s_stage[threadIdx.x >> 5].bar[1], 1);
      asm volatile("fence.mbarrier_init.release.cluster;" ::: "memory");
    }
    __syncwarp();
  }
  const uint32_t FULL = 0xFFFFFFFFu;
  const double thr = table_at(P.thr_pt2pt, P.table_len, it);
  const float thr2 = float(thr * thr);
  const float thr_pl = float(table_at(P.thr_pt2pl, P.table_len, it));
  const double kc = table_at(P.kparam, P.table_len, it);
  const uint32_t lane = threadIdx.x & 31u, warp = threadIdx.x >> 5;
  WarpScratch& ws = s_ws[warp];
  const uint32_t q = chunk * (32u * NWARPS) + threadIdx.x;
  const bool mine = q < P.n_q;
  float4 l = make_float4(0.f, 0.f, 0.f, 0.f);
  float gx = 0.f, gy = 0.f, gz = 0.f;
  if (mine) {
    l = __ldg(&local[P.q_begin + q]);
    compose_point_f(sT, l.x, l.y, l.z, gx, gy, gz);
  }
  float4 pa = make_float4(0.f, 0.f, 0.f, 0.f);
  float4 pb = make_float4(0.f, 0.f, 0.f, 0.f);
  uint32_t ncand = 0;
  bool paired = false;
  if ((P.matcher_mask & MLO_MATCHER_PT2PL) && mine) {
    const PlaneHit h = nn_plane_thread(map, gx, gy, gz);
    ncand += h.ncand;
    if (h.found && h.dist < thr_pl) {
      paired = true;
      pa = make_float4(h.cx, h.cy, h.cz, 2.f);
      pb = make_float4(h.nx, h.ny, h.nz, 0.f);
    }
  }
  if (P.matcher_mask & MLO_MATCHER_PT2PT) {
    const bool want = mine && !paired;  // Matcher base rule: already-paired local points are skipped
    ws.q[0][lane] = gx;
    ws.q[1][lane] = gy;
    ws.q[2][lane] = gz;
    ws.best[lane] = ~0ull;
    if constexpr (A32 == 1) {
      ws.bd2[lane] = 0xFFFFFFFFu;
      ws.bord[lane] = 0xFFFFFFFFu;
    }
    int32_t kq[3] = {0, 0, 0};
    bool active = false;
    if (want) {
      kq[0] = voxel_index_map(gx, map.inv_voxel, map.index_floor);
      kq[1] = voxel_index_map(gy, map.inv_voxel, map.index_floor);
      kq[2] = voxel_index_map(gz, map.inv_voxel, map.index_floor);
      active = key_in_range(kq[0]) && key_in_range(kq[1]) && key_in_range(kq[2]);
    }
    // ---- phase 1: probes (thread per query), words -> shared memory
    if (active) {
      ncand += probe_words(map, kq, &ws.words[0][lane], 32);
    } else {
#pragma unroll
      for (int e = 0; e < 27; e++) ws.words[e][lane] = CELL_ABSENT;
    }
    __syncwarp();
    // ---- phase 2: own cells
    {
      const uint32_t wh = ws.words[13][lane];
      const uint32_t np = (wh == CELL_ABSENT) ? 0u : (cell_cnt(wh) + 7u) >> 3;
      uint32_t incl = np;
#pragma unroll
      for (int o = 1; o < 32; o <<= 1) {
        const uint32_t y = __shfl_up_sync(FULL, incl, o);
        if (lane >= uint32_t(o)) incl += y;
      }
      const uint32_t total = __shfl_sync(FULL, incl, 31);
      uint32_t off = incl - np;
      for (uint32_t k = 0; k < np; k++) ws.list[off + k] = uint16_t((k << 10) | (13u << 5) | lane);
      __syncwarp();
      if constexpr (BULK) wl_process_bulk(map, ws, s_stage[warp], total, bulk_par);
      else if constexpr (OCT > 0) wl_process_oct<OCT>(map, ws, total);
      else if constexpr (A32 == 1) wl_process_a32(map, ws, total);
      else if constexpr (A32 == 2) wl_process_ballot(map, ws, total);
      else if constexpr (A32 == 3) wl_process_deep(map, ws, total);
      else if constexpr (A32 == 4) wl_process_runs(map, ws, total);
      else wl_process<PIPE>(map, ws, total);
    }
    // ---- phase 3: per query, the neighbour cells that can still beat the bound from the own cell
    uint32_t visit = 0, my_items = 0;
    if (active) {
      const unsigned long long b0 = ws.best[lane];
      const float bound = (b0 == ~0ull) ? __int_as_float(0x7f800000) : __uint_as_float(uint32_t(b0 >> 32));
      const float qv[3] = {gx, gy, gz};
      const AxisGaps gaps = axis_gaps(map.voxel_size, qv, kq, map.index_floor);
#pragma unroll
      for (int e = 0; e < 27; e++) {
        if (e == 13) continue;
        const uint32_t we = ws.words[e][lane];
        if (we == CELL_ABSENT || cell_cnt(we) == 0) continue;
        if (MLO_LB2(gaps, e) <= bound) {
          visit |= 1u << e;
          my_items += (cell_cnt(we) + 7u) >> 3;
        }
      }
    }
    // ---- phase 4: drain the neighbour segments, in lane ranges that fit the list
    uint32_t start = 0;
    while (start < 32) {
      const uint32_t c = lane >= start ? my_items : 0u;
      uint32_t incl = c;
#pragma unroll
      for (int o = 1; o < 32; o <<= 1) {
        const uint32_t y = __shfl_up_sync(FULL, incl, o);
        if (lane >= uint32_t(o)) incl += y;
      }
      const uint32_t end = __popc(__ballot_sync(FULL, incl <= WL_CAP));  // fits is a prefix of the lanes
      const uint32_t total = __shfl_sync(FULL, incl, end - 1);
      if (lane >= start && lane < end) {
        uint32_t off = incl - c;
        uint32_t m = visit;
        while (m) {
          const uint32_t e = __ffs(m) - 1;
          m &= m - 1;
          const uint32_t np = (cell_cnt(ws.words[e][lane]) + 7u) >> 3;
          for (uint32_t k = 0; k < np; k++) ws.list[off++] = uint16_t((k << 10) | (e << 5) | lane);
        }
      }
      __syncwarp();
      if constexpr (BULK) wl_process_bulk(map, ws, s_stage[warp], total, bulk_par);
      else if constexpr (OCT > 0) wl_process_oct<OCT>(map, ws, total);
      else if constexpr (A32 == 1) wl_process_a32(map, ws, total);
      else if constexpr (A32 == 2) wl_process_ballot(map, ws, total);
      else if constexpr (A32 == 3) wl_process_deep(map, ws, total);
      else if constexpr (A32 == 4) wl_process_runs(map, ws, total);
      else wl_process<PIPE>(map, ws, total);
      start = end;
    }
    // ---- result per query
    if (active) {
      const unsigned long long b = ws.best[lane];
      if (b != ~0ull) {
        const float d2 = __uint_as_float(uint32_t(b >> 32));
        const uint32_t ord = uint32_t(b & 0xFFFFFFFFu), e = ord >> 5, slot = ord & 31u;
        const float lim = thr2 + P.ang2 * (gx * gx + gy * gy + gz * gz);
        if (d2 < lim) {
          const float4 g = __ldg(map.pts + size_t(cell_vid(ws.words[e][lane])) * map.row + slot);
          pa = make_float4(g.x, g.y, g.z, 1.f);
        }
      }
    }
    __syncwarp();
  }
  double a[NACC];
#pragma unroll
  for (int k = 0; k < int(NACC); k++) a[k] = 0.0;
  uint32_t npairs = 0;
  if (mine) {
    if (pa.w == 1.f) {
      if (P.solver == MLO_SOLVER_GAUSS_NEWTON)
        contrib_pt2pt(sT, l.x, l.y, l.z, pa.x, pa.y, pa.z, P.w_pt2pt, P.robust_kernel, kc, a);
      else
        contrib_horn(l.x, l.y, l.z, pa.x, pa.y, pa.z, a);
      npairs = 1;
    } else if (pa.w == 2.f) {
      contrib_pt2pl(sT, l.x, l.y, l.z, pa.x, pa.y, pa.z, pb.x, pb.y, pb.z, P.w_pt2pl, P.robust_kernel, kc, a);
      npairs = 1;
      pairB[P.q_begin + q] = pb;
    }
    pairA[P.q_begin + q] = pa;
  }
  if constexpr (WPART && NWARPS > 1) {
    // every warp of the block publishes its own partial (P.n_blocks counts 32-query chunks): no block barrier, a warp
    // that finishes its work list early retires without waiting for its neighbours
    const uint32_t wchunk = chunk * NWARPS + warp;
    if (wchunk < P.n_blocks) {
      const uint32_t pbi = P.part_begin + wchunk;
      warp_reduce_store(a, npairs, ncand, partials + size_t(pbi) * NACC, part_cnt + 2 * size_t(pbi));
    }
    return;
  }
  const uint32_t pbi = P.part_begin + chunk;
  if (NWARPS == 1)
    warp_reduce_store(a, npairs, ncand, partials + size_t(pbi) * NACC, part_cnt + 2 * size_t(pbi));
  else
    block_reduce_store(a, npairs, ncand, partials + size_t(pbi) * NACC, part_cnt + 2 * size_t(pbi));
}

// Inner Gauss-Newton iterations >= 1: re-linearise over the stored pairings (no NN). Pairings were written
// by other blocks, possibly within the same launch: read them through L2 (ld.global.cg).
MLO_D void chunk_accumulate(const IcpProblem& P, const double* sT, uint32_t it, uint32_t chunk,
                            const float4* __restrict__ local, const float4* pairA, const float4* pairB,
                            double* __restrict__ partials, uint32_t* __restrict__ part_cnt) {
  const double kc = table_at(P.kparam, P.table_len, it);
  double a[NACC];
#pragma unroll
  for (int k = 0; k < int(NACC); k++) a[k] = 0.0;
  uint32_t npairs = 0;
  const uint32_t q = chunk * ICP_BLOCK + threadIdx.x;
  if (q < P.n_q) {
    const float4 pa = __ldcg(&pairA[P.q_begin + q]);
    if (pa.w != 0.f) {
      const float4 l = __ldg(&local[P.q_begin + q]);
      if (pa.w == 1.f) {
        contrib_pt2pt(sT, l.x, l.y, l.z, pa.x, pa.y, pa.z, P.w_pt2pt, P.robust_kernel, kc, a);
      } else {
        const float4 nb = __ldcg(&pairB[P.q_begin + q]);
        contrib_pt2pl(sT, l.x, l.y, l.z, pa.x, pa.y, pa.z, nb.x, nb.y, nb.z, P.w_pt2pl, P.robust_kernel, kc, a);
      }
      npairs++;
    }
  }
  const uint32_t pbi = P.part_begin + chunk;
  block_reduce_store_t(a, npairs, 0u, partials + size_t(pbi) * NACC, part_cnt + 2 * size_t(pbi));
}

// Horn's closed form from the reduced sums (Solver_Horn): dominant eigenvector of the 4x4 N matrix.
MLO_D bool horn_from_sums(const double* a, double n, double* T) {
  if (n < 3.0) return false;
  const double invn = 1.0 / n;
  double cl[3], cg[3], S[3][3];
  for (int i = 0; i < 3; i++) {
    cl[i] = a[i] * invn;
    cg[i] = a[3 + i] * invn;
  }
  for (int i = 0; i < 3; i++)
    for (int j = 0; j < 3; j++) S[i][j] = a[6 + 3 * i + j] - n * cl[i] * cg[j];
  double N[4][4] = {{S[0][0] + S[1][1] + S[2][2], S[1][2] - S[2][1], S[2][0] - S[0][2], S[0][1] - S[1][0]},
                    {S[1][2] - S[2][1], S[0][0] - S[1][1] - S[2][2], S[0][1] + S[1][0], S[2][0] + S[0][2]},
                    {S[2][0] - S[0][2], S[0][1] + S[1][0], -S[0][0] + S[1][1] - S[2][2], S[1][2] + S[2][1]},
                    {S[0][1] - S[1][0], S[2][0] + S[0][2], S[1][2] + S[2][1], -S[0][0] - S[1][1] + S[2][2]}};
  double V[4][4];
  for (int i = 0; i < 4; i++)
    for (int j = 0; j < 4; j++) V[i][j] = (i == j) ? 1.0 : 0.0;
  for (int sweep = 0; sweep < 16; sweep++)
    for (int p = 0; p < 3; p++)
      for (int q = p + 1; q < 4; q++) {
        const double apq = N[p][q];
        if (fabs(apq) < 1e-300) continue;
        const double theta = (N[q][q] - N[p][p]) / (2.0 * apq);
        const double t = (theta >= 0 ? 1.0 : -1.0) / (fabs(theta) + sqrt(theta * theta + 1.0));
        const double c = 1.0 / sqrt(t * t + 1.0), s = t * c;
        for (int k = 0; k < 4; k++) {
          const double akp = N[k][p], akq = N[k][q];
          N[k][p] = c * akp - s * akq;
          N[k][q] = s * akp + c * akq;
        }
        for (int k = 0; k < 4; k++) {
          const double apk = N[p][k], aqk = N[q][k];
          N[p][k] = c * apk - s * aqk;
          N[q][k] = s * apk + c * aqk;
        }
        for (int k = 0; k < 4; k++) {
          const double vkp = V[k][p], vkq = V[k][q];
          V[k][p] = c * vkp - s * vkq;
          V[k][q] = s * vkp + c * vkq;
        }
      }
  int im = 0;
  for (int k = 1; k < 4; k++)
    if (N[k][k] > N[im][im]) im = k;
  double qw = V[0][im], qx = V[1][im], qy = V[2][im], qz = V[3][im];
  const double qn = sqrt(qw * qw + qx * qx + qy * qy + qz * qz);
  if (!(qn > 0.0)) return false;
  qw /= qn; qx /= qn; qy /= qn; qz /= qn;
  if (qw < 0) { qw = -qw; qx = -qx; qy = -qy; qz = -qz; }
  T[0] = 1 - 2 * (qy * qy + qz * qz); T[1] = 2 * (qx * qy - qw * qz); T[2] = 2 * (qx * qz + qw * qy);
  T[4] = 2 * (qx * qy + qw * qz); T[5] = 1 - 2 * (qx * qx + qz * qz); T[6] = 2 * (qy * qz - qw * qx);
  T[8] = 2 * (qx * qz - qw * qy); T[9] = 2 * (qy * qz + qw * qx); T[10] = 1 - 2 * (qx * qx + qy * qy);
  for (int i = 0; i < 3; i++) T[4 * i + 3] = cg[i] - (T[4 * i] * cl[0] + T[4 * i + 1] * cl[1] + T[4 * i + 2] * cl[2]);
  return true;
}

#ifdef MLO_TRACE
// Timeline of problem 0 inside the persistent kernel (scratch builds only: scratch/trace_persistent.py).
__device__ unsigned long long g_trace[16384];
__device__ unsigned int g_trace_n;
MLO_D void trace_event(uint32_t prob, uint32_t code) {
  if (prob != 0) return;
  unsigned long long t;
  asm volatile("mov.u64 %0, %%globaltimer;" : "=l"(t));
  const unsigned int i = atomicAdd(&g_trace_n, 1u);
  if (i < 16384) g_trace[i] = (t << 8) | code;
}
#define MLO_TRACE_EVENT(prob, code) do { if (threadIdx.x == 0) trace_event(prob, code); } while (0)
#define MLO_TRACE_SOLVE(code) do { if ((threadIdx.x & 31u) == 0) trace_event(0u, code); } while (0)
MLO_D void trace_event_any(uint32_t code) { trace_event(0u, code); }
#else
MLO_D void trace_event_any(uint32_t) {}
#define MLO_TRACE_EVENT(prob, code) do { } while (0)
#define MLO_TRACE_SOLVE(code) do { } while (0)
#endif

// End-of-iteration bookkeeping of mp2p_icp::ICP::align — step measure against prev and prev-prev,
// hook-as-data, stall test, iteration counter, MaxIterations — lives in solve_core below.

// Sum of a problem's block partials by a whole thread block (SOLVE_WARPS warps): warp w adds the partials
// b = w, w + SOLVE_WARPS, ... in ascending order (lane k owns element k, eight L2 loads in flight), then the warp
// sums are combined in warp order.  The order is fixed by (nblk, SOLVE_WARPS) alone, so a run is reproducible.
// Every thread of the block must call this (it contains a barrier); the totals land in s_tot / s_cnt.
constexpr uint32_t SOLVE_WARPS = ICP_BLOCK / 32;
struct SolveScratch {
  double tot[NACC];
  uint32_t cnt[2];
  double part[SOLVE_WARPS][NACC];
  uint32_t pcnt[SOLVE_WARPS][2];
  // prior term, spread over the warp (prior_add_warp)
  double g[6], e[6], Le[6], J[36], LJ[36];
  double m[10][9];  // 3x3 intermediates of the warp-parallel right-Jacobian inverse (jr_inv_warp)
};
// The solving block's shared-memory copy of a problem and its state (solve_phase below).
struct SolveStage {
  IcpProblem P;
  IcpState S;
};
static_assert(sizeof(IcpProblem) % 4 == 0 && sizeof(IcpState) % 4 == 0, "copied word by word");
constexpr uint32_t STAGE_P_WORDS = sizeof(IcpProblem) / 4, STAGE_S_WORDS = sizeof(IcpState) / 4;
constexpr uint32_t STAGE_P_REGS = (STAGE_P_WORDS + 31) / 32, STAGE_S_REGS = (STAGE_S_WORDS + 31) / 32;

// `st` != nullptr: the last warp also stages the problem and its state for the solve that follows - its loads are issued
// BEFORE the partial loads and stored after them, so the two round trips to L2 overlap instead of following each other.
MLO_D void sum_partials_block(const IcpProblem& P, const double* partials, const uint32_t* part_cnt, uint32_t nblk,
                              SolveScratch& sc, const IcpState* Sg = nullptr, SolveStage* st = nullptr) {
  const uint32_t lane = threadIdx.x & 31u, warp = threadIdx.x >> 5;
  const bool stager = st != nullptr && warp == SOLVE_WARPS - 1;
  uint32_t rp[STAGE_P_REGS], rs[STAGE_S_REGS];
  if (stager) {
    const uint32_t* gp = reinterpret_cast<const uint32_t*>(&P);
    const uint32_t* gs = reinterpret_cast<const uint32_t*>(Sg);
#pragma unroll
    for (uint32_t u = 0; u < STAGE_P_REGS; u++) rp[u] = lane + 32 * u < STAGE_P_WORDS ? __ldg(gp + lane + 32 * u) : 0u;
#pragma unroll
    for (uint32_t u = 0; u < STAGE_S_REGS; u++) rs[u] = lane + 32 * u < STAGE_S_WORDS ? __ldcg(gs + lane + 32 * u) : 0u;
  }
  if (warp < SOLVE_WARPS) {
    if (lane < NACC) {
      double acc = 0.0;
      const double* base = partials + size_t(P.part_begin) * NACC + lane;
      uint32_t b = warp;
      for (; b + 7 * SOLVE_WARPS < nblk; b += 8 * SOLVE_WARPS) {
        double v[8];
#pragma unroll
        for (int u = 0; u < 8; u++) v[u] = __ldcg(base + size_t(b + u * SOLVE_WARPS) * NACC);
#pragma unroll
        for (int u = 0; u < 8; u++) acc += v[u];
      }
      if (b < nblk) {
        // the last (up to eight) partials of this warp: again all loads in flight before the first add - a plain
        // `acc += load` loop is one L2 round trip per partial (41 chunks: three of them in a row on warp 0).  Added in
        // the same ascending order; the absent ones contribute +0.0.
        double v[8];
#pragma unroll
        for (int u = 0; u < 8; u++) {
          const uint32_t bb = b + u * SOLVE_WARPS;
          v[u] = bb < nblk ? __ldcg(base + size_t(bb) * NACC) : 0.0;
        }
#pragma unroll
        for (int u = 0; u < 8; u++) acc += v[u];
      }
      sc.part[warp][lane] = acc;
    } else if (lane < NACC + 2) {
      uint32_t cnt = 0;
      for (uint32_t b0 = warp; b0 < nblk; b0 += 8 * SOLVE_WARPS) {  // (eight loads in flight, as above)
        uint32_t v[8];
#pragma unroll
        for (int u = 0; u < 8; u++) {
          const uint32_t bb = b0 + u * SOLVE_WARPS;
          v[u] = bb < nblk ? __ldcg(&part_cnt[2 * size_t(P.part_begin + bb) + (lane - NACC)]) : 0u;
        }
#pragma unroll
        for (int u = 0; u < 8; u++) cnt += v[u];
      }
      sc.pcnt[warp][lane - NACC] = cnt;
    }
  }
  if (stager) {
    uint32_t* dp = reinterpret_cast<uint32_t*>(&st->P);
    uint32_t* ds = reinterpret_cast<uint32_t*>(&st->S);
#pragma unroll
    for (uint32_t u = 0; u < STAGE_P_REGS; u++)
      if (lane + 32 * u < STAGE_P_WORDS) dp[lane + 32 * u] = rp[u];
#pragma unroll
    for (uint32_t u = 0; u < STAGE_S_REGS; u++)
      if (lane + 32 * u < STAGE_S_WORDS) ds[lane + 32 * u] = rs[u];
  }
  __syncthreads();
  if (threadIdx.x < NACC) {
    double t = sc.part[0][threadIdx.x];
#pragma unroll
    for (uint32_t w = 1; w < SOLVE_WARPS; w++) t += sc.part[w][threadIdx.x];
    sc.tot[threadIdx.x] = t;
  } else if (threadIdx.x < NACC + 2) {
    uint32_t t = 0;
#pragma unroll
    for (uint32_t w = 0; w < SOLVE_WARPS; w++) t += sc.pcnt[w][threadIdx.x - NACC];
    sc.cnt[threadIdx.x - NACC] = t;
  }
  __syncthreads();
}

// ---- the solve step (one warp; problem and state in SHARED memory) --------------------------------------------
// Prior term of Solver_GaussNewton (LidarOdometry.cpp:854-877): e = log(prior^-1 T), J = d log(D exp(eps))/d eps;
// g += J^T L e ; H += J^T L J.  Lane 0 evaluates e and J (out of line: loop-heavy, keeps the common path's arrays in
// registers); the 6x6 products are spread over the warp, one output entry per lane, each entry summed in the same order
// (m = 0..5 onto the running value) as a single-thread loop would: same bits, a third of the time.
__device__ __noinline__ void prior_e(const IcpProblem& P, const double* T, double* e_out) {
  double D[12], e[6];
  pose_minus(T, P.prior_pose, D);
  se3_log(D, e);
  for (int i = 0; i < 6; i++) e_out[i] = e[i];
}
// se3_right_jacobian_inv (se3.cuh) by nine lanes of a warp, one 3x3 entry each: the ten 3x3 products of the closed form
// are five dependent levels of 3-term dot products instead of ~550 serial double operations on one lane.  Every entry is
// evaluated by the SAME expression, in the same order, as in the single-thread function: identical bits.
// xi and J live in shared memory; the whole warp must call this.
MLO_D double m3e(const double* A, const double* B, uint32_t i, uint32_t j) {
  return A[3 * i] * B[j] + A[3 * i + 1] * B[3 + j] + A[3 * i + 2] * B[6 + j];
}
MLO_D double hat_entry(const double* w, uint32_t idx) {
  switch (idx) {
    case 1: return -w[2];
    case 2: return w[1];
    case 3: return w[2];
    case 5: return -w[0];
    case 6: return -w[1];
    case 7: return w[0];
    default: return 0.0;
  }
}
__device__ __noinline__ void jr_inv_warp(const double* xi, double* J, double (*m)[9]) {
  const uint32_t lane = threadIdx.x & 31u;
  const bool on = lane < 9;
  const uint32_t idx = on ? lane : 0u, i = idx / 3, j = idx % 3;
  const double rho[3] = {-xi[0], -xi[1], -xi[2]};
  const double phi[3] = {-xi[3], -xi[4], -xi[5]};
  const double t2 = phi[0] * phi[0] + phi[1] * phi[1] + phi[2] * phi[2];
  double k, c1, c2, c3;
  jr_inv_coeffs(t2, k, c1, c2, c3);
  double* Pm = m[0]; double* F = m[1]; double* FP = m[2]; double* PF = m[3]; double* FPF = m[4];
  double* A = m[5]; double* Q = m[6]; double* AQ = m[7];
  // level 0: P = hat(rho), F = hat(phi)
  const double p_e = hat_entry(rho, idx), f_e = hat_entry(phi, idx);
  if (on) {
    Pm[idx] = p_e;
    F[idx] = f_e;
  }
  __syncwarp();
  // level 1: FF, FP, PF
  const double ff_e = m3e(F, F, i, j), fp_e = m3e(F, Pm, i, j), pf_e = m3e(Pm, F, i, j);
  const double a_e = ((idx % 4 == 0) ? 1.0 : 0.0) - 0.5 * f_e + k * ff_e;
  if (on) {
    FP[idx] = fp_e;
    PF[idx] = pf_e;
    A[idx] = a_e;
  }
  __syncwarp();
  // level 2: FPF, FFP, PFF
  const double fpf_e = m3e(FP, F, i, j), ffp_e = m3e(F, FP, i, j), pff_e = m3e(PF, F, i, j);
  if (on) FPF[idx] = fpf_e;
  __syncwarp();
  // level 3: FPFF, FFPF, then Q
  const double fpff_e = m3e(FPF, F, i, j), ffpf_e = m3e(F, FPF, i, j);
  const double q_e = 0.5 * p_e + c1 * (fp_e + pf_e + fpf_e) + c2 * (ffp_e + pff_e - 3.0 * fpf_e) + c3 * (fpff_e + ffpf_e);
  if (on) Q[idx] = q_e;
  __syncwarp();
  // level 4: AQ
  const double aq_e = m3e(A, Q, i, j);
  if (on) AQ[idx] = aq_e;
  __syncwarp();
  // level 5: AQA and the 6x6 blocks
  const double aqa_e = m3e(AQ, A, i, j);
  if (on) {
    J[6 * i + j] = a_e;
    J[6 * i + 3 + j] = -aqa_e;
    J[6 * (i + 3) + j] = 0.0;
    J[6 * (i + 3) + 3 + j] = a_e;
  }
  __syncwarp();
}
// (one warp) linearisation of the prior term at pose T: e, J, L e, L J into the given arrays; m = 3x3 scratch
MLO_D void prior_compute_warp(const IcpProblem& P, const double* T, double* e, double* J, double* Le, double* LJ, double (*m)[9]) {
  const uint32_t lane = threadIdx.x & 31u;
  if (lane == 0) prior_e(P, T, e);
  __syncwarp();
  MLO_TRACE_SOLVE(47);  // prior: error vector done
  jr_inv_warp(e, J, m);
  MLO_TRACE_SOLVE(48);  // prior: Jacobian done
  if (lane < 6) {
    double s = 0;
    for (int k = 0; k < 6; k++) s += P.prior_info[6 * lane + k] * e[k];
    Le[lane] = s;
  }
  for (uint32_t idx = lane; idx < 36; idx += 32) {
    const uint32_t i = idx / 6, j = idx % 6;
    double t = 0;
    for (int k = 0; k < 6; k++) t += P.prior_info[6 * i + k] * J[6 * k + j];
    LJ[idx] = t;
  }
  __syncwarp();
}
// (one warp) g += J^T L e ; H += J^T L J, each entry summed term by term onto the running value
MLO_D void prior_accumulate_warp(const double* J, const double* Le, const double* LJ, double* g, double* H) {
  const uint32_t lane = threadIdx.x & 31u;
  if (lane < 6) {
    double gi = g[lane];
    for (int k = 0; k < 6; k++) gi += J[6 * k + lane] * Le[k];
    g[lane] = gi;
  }
  for (uint32_t idx = lane; idx < 36; idx += 32) {
    const uint32_t i = idx / 6, j = idx % 6;
    double h = H[idx];
    for (int k = 0; k < 6; k++) h += J[6 * k + i] * LJ[6 * k + j];
    H[idx] = h;
  }
  __syncwarp();
}
MLO_D void prior_add_warp(const IcpProblem& P, IcpState& S, SolveScratch& sc) {
  if (S.prior_valid) {  // prepared ahead by another warp for exactly this pose (solve_phase)
    prior_accumulate_warp(S.pJ, S.pLe, S.pLJ, sc.g, S.H);
  } else {
    prior_compute_warp(P, S.T, sc.e, sc.J, sc.Le, sc.LJ, sc.m);
    prior_accumulate_warp(sc.J, sc.Le, sc.LJ, sc.g, S.H);
  }
}
// (one warp, while the others are busy elsewhere) prepare the prior's linearisation at the pose the next solve starts from
MLO_D void prior_precompute_warp(const IcpProblem& P, IcpState& S, SolveScratch& sc) {
  prior_compute_warp(P, S.T, S.pe, S.pJ, S.pLe, S.pLJ, sc.m);
  if ((threadIdx.x & 31u) == 0) S.prior_valid = 1;
  __syncwarp();
}
__device__ __noinline__ bool horn_from_sums_ool(const double* a, double n, double* T) { return horn_from_sums(a, n, T); }

// Measures of T against a reference pose, one lane each: D = ref^-1 T; dt, dr = norms of the two halves of
// log_SE3(D) (the stall / oscillation test of ICP::align); tD = |D.t|, wD = |log_SO3(D.R)| (the twist hook of
// LidarOdometry.cpp:930-940).  The three references of an iteration (prev, prev-prev, hook checkpoint) run on
// three lanes through this ONE code path, i.e. concurrently.
MLO_D void pose_measures(const double* T, const double* ref, double& dt, double& dr, double& tD, double& wD) {
  double D[12], d[6];
  pose_minus(T, ref, D);
  se3_log(D, d);
  dt = nrm3(d);
  dr = nrm3(d + 3);
  const double tt[3] = {D[3], D[7], D[11]};
  tD = nrm3(tt);
  wD = dr;  // se3_log's rotation half IS so3_log_of_pose(D)
}

// Part A (sums -> system -> prior -> LDL^T -> retraction) returns 0 = finished, 1 = inner iteration pending, 3 = the last
// inner iteration retracted: part B (step measures, termination, log) follows and returns 0 or 2.
MLO_D int solve_part_a(const IcpProblem& P, IcpState& S, SolveScratch& sc, int after_match) {
  const uint32_t FULL = 0xFFFFFFFFu;
  const uint32_t lane = threadIdx.x & 31u;
  const uint32_t npairs = sc.cnt[0];
  const uint32_t ncand = sc.cnt[1];
  double* const sT = S.T;
  int next = 0;
  bool finished = false;
  if (after_match) {
    if (lane == 0) {
      S.n_pairs = npairs;
      uint64_t pot = 0;
      if (P.matcher_mask & MLO_MATCHER_PT2PL) pot += P.n_q;
      if (P.matcher_mask & MLO_MATCHER_PT2PT) pot += P.n_q;
      S.n_potential = pot;
      S.n_query_it += P.n_q;
      S.n_cand += ncand;
      S.inner = 0;
      if (npairs == 0) {
        S.term = MLO_TERM_NO_PAIRINGS;
        S.done = 1;
      }
    }
    finished = npairs == 0;  // (warp-uniform)
  }
  if (!finished) {
    if (P.solver == MLO_SOLVER_HORN) {
      if (lane == 0) {
        const bool ok = horn_from_sums_ool(sc.tot, double(npairs), sT);
        if (!ok) {
          S.term = MLO_TERM_SOLVER_ERROR;
          S.done = 1;
        } else {
          next = 3;
        }
      }
    } else {
      // the symmetric 6x6 system is laid out in S.H by the whole warp (it is also the Hessian reported as covariance)
      for (uint32_t idx = lane; idx < 36; idx += 32) {
        const uint32_t i = idx / 6, j = idx % 6, r = i < j ? i : j, c = i < j ? j : i;
        S.H[idx] = sc.tot[r * 6 - (r * (r - 1)) / 2 + (c - r)];
      }
      if (lane < 6) sc.g[lane] = sc.tot[21 + lane];
      __syncwarp();
      MLO_TRACE_SOLVE(40);  // system laid out
      if (P.has_prior) prior_add_warp(P, S, sc);  // (warp-uniform branch)
      MLO_TRACE_SOLVE(41);  // prior term added
      if (lane == 0) {
        double g[6];
#pragma unroll
        for (int i = 0; i < 6; i++) g[i] = sc.g[i];
        S.have_H = 1;
        double H[36], mg[6], delta[6];
#pragma unroll
        for (int i = 0; i < 36; i++) H[i] = S.H[i];
#pragma unroll
        for (int i = 0; i < 6; i++) mg[i] = -g[i];
        const bool ok = ldlt6(H, mg, delta);
        MLO_TRACE_SOLVE(42);  // 6x6 solved
        if (!ok) {
          S.term = MLO_TERM_SOLVER_ERROR;
          S.done = 1;
        } else {
          double E[12], Tn[12];
          se3_exp(delta, E);
          pose_mul(sT, E, Tn);
#pragma unroll
          for (int i = 0; i < 12; i++) sT[i] = Tn[i];
          S.prior_valid = 0;  // (a prepared prior linearisation belongs to the pose just left)
          MLO_TRACE_SOLVE(43);  // retraction done
          double dn = 0;
#pragma unroll
          for (int i = 0; i < 6; i++) dn += delta[i] * delta[i];
          S.inner++;
          const bool last_inner = (sqrt(dn) < P.gn_min_delta) || (S.inner >= P.gn_max_iterations);
          if (!last_inner) {
            S.inner_pending = 1;
            next = 1;
          } else {
            next = 3;  // retraction done, this was the last inner iteration: end-of-iteration bookkeeping follows
          }
        }
      }
    }
  }
  __syncwarp();
  return __shfl_sync(FULL, next, 0);
}
MLO_D int solve_part_b(const IcpProblem& P, IcpState& S) {
  const uint32_t FULL = 0xFFFFFFFFu;
  const uint32_t lane = threadIdx.x & 31u;
  double* const sT = S.T;
  int next = 0;
  {
    __syncwarp();  // the new pose (written by lane 0) is visible to lanes 1 and 2
    const int has2 = S.has_prev2;
    double dt = 1e300, dr = 1e300, tD = 0.0, wD = 0.0;
    const bool measure = lane == 0 || (lane == 1 && has2) || (lane == 2 && P.hook_enabled);
    if (measure) {
      const double* ref = lane == 0 ? S.prev : (lane == 1 ? S.prev2 : P.hook_checkpoint);
      pose_measures(sT, ref, dt, dr, tD, wD);
    }
    MLO_TRACE_SOLVE(44);  // step measures done
    const double dt1 = __shfl_sync(FULL, dt, 1), dr1 = __shfl_sync(FULL, dr, 1);
    const double tH = __shfl_sync(FULL, tD, 2), wH = __shfl_sync(FULL, wD, 2);
    __syncwarp();  // lane 1 has read prev2 before lane 0 overwrites it below
    if (lane == 0) {
      double* prev = S.prev;
      double* prev2 = S.prev2;
#pragma unroll
      for (int k = 0; k < 12; k++) {
        prev2[k] = prev[k];
        prev[k] = sT[k];
      }
      S.has_prev2 = 1;
      S.inner_pending = 0;
      dt = fmin(dt, dt1);
      dr = fmin(dr, dr1);
      const uint32_t it_now = S.it;
      if (P.hook_enabled && (tH > P.hook_min_trans || wH > P.hook_min_rot)) {
        S.term = MLO_TERM_HOOK_REQUEST;
        S.done = 1;
      } else if (fabs(dt) < P.min_abs_step_trans && fabs(dr) < P.min_abs_step_rot) {
        S.term = MLO_TERM_STALLED;
        S.done = 1;
      } else {
        S.it++;
        if (S.it >= P.max_iterations) {
          S.term = MLO_TERM_MAX_ITERATIONS;
          S.done = 1;
        }
      }
      next = S.done ? 0 : 2;
      if (P.log && it_now < P.log_cap) {  // the ICP log (mlo_icp_log_enable)
        mlo_icp_iteration_record& r = P.log[it_now];
        r.iteration = it_now;
        r.n_pairings = uint32_t(S.n_pairs);
#pragma unroll
        for (int k = 0; k < 12; k++) r.pose_3x4[k] = sT[k];
        r.threshold_pt2pt = table_at(P.thr_pt2pt, P.table_len, it_now);
        r.threshold_pt2pl = table_at(P.thr_pt2pl, P.table_len, it_now);
        r.kernel_param = table_at(P.kparam, P.table_len, it_now);
        r.step_trans = dt;
        r.step_rot = dr;
        r.termination = S.done ? S.term : MLO_TERM_UNDEFINED;
        r.pad = 0;
      }
    }
  }
  __syncwarp();
  return __shfl_sync(FULL, next, 0);
}
// Executed by ONE warp once the 27 sums (+ counts) of the current linearisation sit in sc.tot / sc.cnt.
// P and S live in shared memory.  Returns (on every lane) 0 = problem finished, 1 = another inner GN iteration is
// pending, 2 = next ICP iteration.
MLO_D int solve_core(const IcpProblem& P, IcpState& S, SolveScratch& sc, int after_match) {
  const int a = solve_part_a(P, S, sc, after_match);
  return a == 3 ? solve_part_b(P, S) : a;
}

// solve_core for callers whose problem and state live in GLOBAL memory (launch sequence, queue-driven kernel): the
// solving block stages both in shared memory ONCE per solve phase (two coalesced reads by a whole warp instead of ~150
// serial, mostly dependent global accesses by lane 0); the first solve and every fused inner iteration that follows work
// on the staged copy (the re-linearisation reads the new pose from there), and the state goes back to global memory in
// one coalesced write when the phase is over.
// (one warp)
MLO_D void solve_stage_in(const IcpProblem& Pg, const IcpState& Sg, SolveStage& st) {
  const uint32_t lane = threadIdx.x & 31u;
  const uint32_t* gp = reinterpret_cast<const uint32_t*>(&Pg);
  const uint32_t* gs = reinterpret_cast<const uint32_t*>(&Sg);
  uint32_t* dp = reinterpret_cast<uint32_t*>(&st.P);
  uint32_t* ds = reinterpret_cast<uint32_t*>(&st.S);
  for (uint32_t i = lane; i < sizeof(IcpProblem) / 4; i += 32) dp[i] = __ldg(gp + i);
  for (uint32_t i = lane; i < sizeof(IcpState) / 4; i += 32) ds[i] = __ldcg(gs + i);
  __syncwarp();
}
// (one warp) the state back to global memory, fenced: whoever is told about this problem next sees it
MLO_D void solve_stage_out(IcpState& Sg, const SolveStage& st) {
  const uint32_t lane = threadIdx.x & 31u;
  __syncwarp();
  uint32_t* gs = reinterpret_cast<uint32_t*>(&Sg);
  const uint32_t* ds = reinterpret_cast<const uint32_t*>(&st.S);
  for (uint32_t i = lane; i < sizeof(IcpState) / 4; i += 32) gs[i] = ds[i];
  __threadfence();
  __syncwarp();
}
__device__ __noinline__ int solve_part_a_staged(SolveStage& st, SolveScratch& sc, int after_match) {
  MLO_TRACE_SOLVE(45);  // problem + state staged
  const int code = solve_part_a(st.P, st.S, sc, after_match);
  MLO_TRACE_SOLVE(46);  // part A returned
  return code;
}
__device__ __noinline__ int solve_part_b_staged(SolveStage& st) { return solve_part_b(st.P, st.S); }
__device__ __noinline__ void prior_precompute_staged(SolveStage& st, SolveScratch& sc) { prior_precompute_warp(st.P, st.S, sc); }

// One solve by a block of ICP_BLOCK threads on the staged copy, every thread calls it: the first warp runs part A; once
// the new pose stands, the first warp finishes the iteration's bookkeeping (part B: three SE(3) logs, termination, log
// record) WHILE the second warp prepares the prior's linearisation at that pose for the next solve of the problem -
// ~2 us of serial double-precision work taken off the next solve's critical path.  Returns the block-uniform verdict
// (0 finished, 1 inner iteration pending, 2 next ICP iteration).
MLO_D int block_solve(SolveStage& st, SolveScratch& sc, int after_match, bool prior_ahead) {
  // (two words: a thread still reading part A's verdict never races with part B's; between two calls of a block there
  // is always a barrier - the fused loop's reduction, solve_phase's closing one)
  __shared__ int b_code_a, b_code_b;
  const uint32_t warp = threadIdx.x >> 5;
  if (warp == 0) {
    const int code = solve_part_a_staged(st, sc, after_match);
    if (threadIdx.x == 0) b_code_a = code;
  }
  __syncthreads();  // the new pose is visible to every warp
  int code = b_code_a;
  if (code == 3) {  // (block-uniform)
    if (warp == 0) {
      const int n = solve_part_b_staged(st);
      if (threadIdx.x == 0) b_code_b = n;
    } else if (warp == 1 && prior_ahead) {
      prior_precompute_staged(st, sc);
    }
    __syncthreads();
    code = b_code_b;
  }
  return code;
}

// Inner Gauss-Newton iterations >= 1 inside the block that just solved (small problems): the block re-linearises ALL
// stored pairings of the problem at the updated pose (thread-strided, then the warp butterfly + ordered cross-warp sum
// of block_reduce_store, into shared memory), and its first warp solves again - no queue round trip, no partials in
// global memory, no other block involved.  Returns the same codes as solve_step (never 1).
constexpr uint32_t FUSE_MAX_Q = 16384;  // <= 128 pairings per thread
MLO_D void block_reduce_to(double* a, uint32_t npairs, SolveScratch& sc) {
  const uint32_t lane = threadIdx.x & 31, warp = threadIdx.x >> 5;
  double v[32];
#pragma unroll
  for (int k = 0; k < int(NACC); k++) v[k] = a[k];
  v[NACC] = double(npairs);
#pragma unroll
  for (int k = int(NACC) + 1; k < 32; k++) v[k] = 0.0;
  const double mine = warp_reduce32_transpose(v);  // lane k: the warp's sum of element k
  if (lane < NACC) sc.part[warp][lane] = mine;
  else if (lane == NACC) sc.pcnt[warp][0] = uint32_t(mine);
  __syncthreads();
  if (threadIdx.x < NACC) {
    double t = sc.part[0][threadIdx.x];
#pragma unroll
    for (uint32_t w = 1; w < SOLVE_WARPS; w++) t += sc.part[w][threadIdx.x];
    sc.tot[threadIdx.x] = t;
  } else if (threadIdx.x == NACC) {
    uint32_t t = 0;
#pragma unroll
    for (uint32_t w = 0; w < SOLVE_WARPS; w++) t += sc.pcnt[w][0];
    sc.cnt[0] = t;
    sc.cnt[1] = 0;
  }
  __syncthreads();
}
// Pairings of the problem being solved, copied into shared memory by the three warps that would otherwise idle while
// the first warp solves (solve_phase): the re-linearisation of a fused inner iteration then reads them from there instead
// of waiting on two dependent global loads per pairing.  Point-to-point pairings only (no normals); problems with more
// than PAIR_STAGE_MAX queries read the rest from global memory.
constexpr uint32_t PAIR_STAGE_MAX = 768;
struct PairStage {
  float4 pa[PAIR_STAGE_MAX];
  float4 l[PAIR_STAGE_MAX];
};

// every thread of the (ICP_BLOCK-wide) block calls this with the block-uniform `next` of the preceding solve, whose
// problem and state sit in `st` (written by the block's first warp; the caller's barrier made them visible)
MLO_D int fused_inner_iterations(SolveStage& st, SolveScratch& sc, int next, uint32_t it, const float4* __restrict__ local,
                                 const float4* pairA, const float4* pairB, const PairStage* ps, bool prior_ahead) {
  const IcpProblem& P = st.P;
  const uint32_t warp = threadIdx.x >> 5;
  // with a prior to prepare, the second warp does that while the other three re-linearise (thread rank 0..95)
  const uint32_t rank = prior_ahead ? (warp == 0 ? threadIdx.x : threadIdx.x - 32u) : threadIdx.x;
  const uint32_t nthr = prior_ahead ? ICP_BLOCK - 32u : ICP_BLOCK;
  while (next == 1) {
    const double* f_T = st.S.T;  // the pose the first warp just retracted
    const double kc = table_at(P.kparam, P.table_len, it);
    double a[NACC];
#pragma unroll
    for (int k = 0; k < int(NACC); k++) a[k] = 0.0;
    uint32_t npairs = 0;
    if (prior_ahead && warp == 1) {
      prior_precompute_staged(st, sc);
    } else {
      for (uint32_t q = rank; q < P.n_q; q += nthr) {
        const bool staged = ps != nullptr && q < PAIR_STAGE_MAX;  // (block-uniform pointer)
        const float4 pa = staged ? ps->pa[q] : __ldcg(&pairA[P.q_begin + q]);
        if (pa.w == 0.f) continue;
        const float4 l = staged ? ps->l[q] : __ldg(&local[P.q_begin + q]);
        if (pa.w == 1.f) {
          contrib_pt2pt(f_T, l.x, l.y, l.z, pa.x, pa.y, pa.z, P.w_pt2pt, P.robust_kernel, kc, a);
        } else {
          const float4 nb = __ldcg(&pairB[P.q_begin + q]);
          contrib_pt2pl(f_T, l.x, l.y, l.z, pa.x, pa.y, pa.z, nb.x, nb.y, nb.z, P.w_pt2pl, P.robust_kernel, kc, a);
        }
        npairs++;
      }
    }
    block_reduce_to(a, npairs, sc);  // (its barriers: every thread has read the pose; the prepared prior is visible)
    next = block_solve(st, sc, 0, prior_ahead);
  }
  return next;
}

// One solve phase of a problem by the block that owns it at this moment (ICP_BLOCK threads, all of them call this; the
// sums of the current linearisation sit in sc.tot / sc.cnt): first solve, fused inner iterations if allowed, state back
// to global memory.  Returns the block-uniform verdict (0 finished, 1 inner iteration pending, 2 next ICP iteration).
MLO_D int solve_phase(const IcpProblem& Pg, IcpState& Sg, SolveStage& st, SolveScratch& sc, int after_match, int fuse, uint32_t it,
                      const float4* __restrict__ local, const float4* pairA, const float4* pairB, PairStage* ps) {
  // (the staged problem is visible to every warp: sum_partials_block brought it in behind its barrier)
  const bool gn = st.P.solver == MLO_SOLVER_GAUSS_NEWTON;
  // `fuse`: bit 0 = run the inner iterations inside this block, bit 1 = prepare the prior's linearisation ahead (host options
  // "fuse_inner", "prior_ahead")
  const bool fused = (fuse & 1) && st.P.n_q <= FUSE_MAX_Q;
  const bool prefetch = fused && gn && st.P.gn_max_iterations > 1;
  const bool prior_ahead = (fuse & 2) && gn && st.P.has_prior;
  if (threadIdx.x >= 32 && prefetch) {
    const uint32_t nq = min(st.P.n_q, PAIR_STAGE_MAX), qb = st.P.q_begin;
    for (uint32_t q = threadIdx.x - 32; q < nq; q += ICP_BLOCK - 32) {
      ps->pa[q] = __ldcg(&pairA[qb + q]);
      ps->l[q] = __ldg(&local[qb + q]);
    }
  }
  int next = block_solve(st, sc, after_match, prior_ahead);  // (its first barrier also publishes the prefetched pairings)
  if (fused) next = fused_inner_iterations(st, sc, next, it, local, pairA, pairB, prefetch ? ps : nullptr, prior_ahead);
  if (threadIdx.x < 32) solve_stage_out(Sg, st);
  __syncthreads();
  return next;
}

// ------------------------------------------------------------------ one kernel per phase (launch sequence)
template <bool MULTI>
__global__ void __launch_bounds__(ICP_BLOCK)
    k_match_accumulate(MapDev map, const MapDev* __restrict__ maps, const IcpProblem* __restrict__ probs,
                       const IcpState* __restrict__ states, const float4* __restrict__ local, float4* __restrict__ pairA,
                       float4* __restrict__ pairB, double* __restrict__ partials, uint32_t* __restrict__ part_cnt, uint32_t qpw) {
  const IcpProblem& P = probs[blockIdx.y];
  if (blockIdx.x >= P.n_blocks) return;
  const IcpState& S = states[blockIdx.y];
  if (S.done) return;
  __shared__ double sT[12];
  __shared__ MapDev sMap;
  if (threadIdx.x < 12) sT[threadIdx.x] = S.T[threadIdx.x];
  if (MULTI) stage_map(sMap, maps, P.map_idx);
  __syncthreads();
  if constexpr (MULTI) chunk_match_warp(sMap, P, sT, S.it, blockIdx.x, local, pairA, pairB, partials, part_cnt, qpw);
  else chunk_match_warp(map, P, sT, S.it, blockIdx.x, local, pairA, pairB, partials, part_cnt, qpw);
}

__global__ void __launch_bounds__(ICP_BLOCK, 4)
    k_match_accumulate_tpq(MapDev map, const IcpProblem* __restrict__ probs, const IcpState* __restrict__ states,
                           const float4* __restrict__ local, float4* __restrict__ pairA, float4* __restrict__ pairB,
                           double* __restrict__ partials, uint32_t* __restrict__ part_cnt) {
  const IcpProblem& P = probs[blockIdx.y];
  if (blockIdx.x >= P.n_blocks) return;
  const IcpState& S = states[blockIdx.y];
  if (S.done) return;
  __shared__ double sT[12];
  if (threadIdx.x < 12) sT[threadIdx.x] = S.T[threadIdx.x];
  __syncthreads();
  chunk_match_tpq(map, P, sT, S.it, blockIdx.x, local, pairA, pairB, partials, part_cnt);
}

// four-warp variant of the work-list kernel (chunk = ICP_BLOCK queries), kept for A/B runs
template <bool MULTI, bool PIPE = false, int MINB = 8, bool BULK = false, int OCT = 0, bool WPART = false, int A32 = 0>
__global__ void __launch_bounds__(ICP_BLOCK, MINB)
    k_match_accumulate_wl4(MapDev map, const MapDev* __restrict__ maps, const IcpProblem* __restrict__ probs,
                           const IcpState* __restrict__ states, const float4* __restrict__ local, float4* __restrict__ pairA,
                           float4* __restrict__ pairB, double* __restrict__ partials, uint32_t* __restrict__ part_cnt) {
  const IcpProblem& P = probs[blockIdx.y];
  if (blockIdx.x * (WPART ? 4u : 1u) >= P.n_blocks) return;
  const IcpState& S = states[blockIdx.y];
  if (S.done) return;
  __shared__ double sT[12];
  __shared__ MapDev sMap;
  if (threadIdx.x < 12) sT[threadIdx.x] = S.T[threadIdx.x];
  if (MULTI) stage_map(sMap, maps, P.map_idx);
  __syncthreads();
  if constexpr (MULTI) chunk_match_wl<4, PIPE, BULK, OCT, WPART, A32>(sMap, P, sT, S.it, blockIdx.x, local, pairA, pairB, partials, part_cnt);
  else chunk_match_wl<4, PIPE, BULK, OCT, WPART, A32>(map, P, sT, S.it, blockIdx.x, local, pairA, pairB, partials, part_cnt);
}

template <int MIN_BLOCKS>
__global__ void __launch_bounds__(WL_BLOCK, MIN_BLOCKS)
    k_match_accumulate_wl(MapDev map, const IcpProblem* __restrict__ probs, const IcpState* __restrict__ states,
                          const float4* __restrict__ local, float4* __restrict__ pairA, float4* __restrict__ pairB,
                          double* __restrict__ partials, uint32_t* __restrict__ part_cnt) {
  const IcpProblem& P = probs[blockIdx.y];
  if (blockIdx.x >= P.n_blocks) return;
  const IcpState& S = states[blockIdx.y];
  if (S.done) return;
  __shared__ double sT[12];
  if (threadIdx.x < 12) sT[threadIdx.x] = S.T[threadIdx.x];
  __syncthreads();
  chunk_match_wl<1>(map, P, sT, S.it, blockIdx.x, local, pairA, pairB, partials, part_cnt);
}

__global__ void __launch_bounds__(ICP_BLOCK)
    k_accumulate(const IcpProblem* __restrict__ probs, const IcpState* __restrict__ states, const float4* __restrict__ local,
                 const float4* pairA, const float4* pairB, double* __restrict__ partials, uint32_t* __restrict__ part_cnt) {
  const IcpProblem& P = probs[blockIdx.y];
  if (blockIdx.x >= P.n_blocks_acc) return;
  const IcpState& S = states[blockIdx.y];
  if (S.done || !S.inner_pending) return;
  __shared__ double sT[12];
  if (threadIdx.x < 12) sT[threadIdx.x] = S.T[threadIdx.x];
  __syncthreads();
  chunk_accumulate(P, sT, S.it, blockIdx.x, local, pairA, pairB, partials, part_cnt);
}

// one block per problem. `after_match` = 1 when the partials come from the match phase (inner 0).  With `fuse` the
// block also runs the remaining inner Gauss-Newton iterations of the problem itself (problems up to FUSE_MAX_Q
// queries), so an ICP iteration of the launch sequence is two launches: match, solve.
__global__ void __launch_bounds__(ICP_BLOCK)
    k_solve(const IcpProblem* __restrict__ probs, IcpState* __restrict__ states, const double* partials,
            const uint32_t* part_cnt, int after_match, uint32_t* __restrict__ n_active, int fuse,
            const float4* __restrict__ local, const float4* pairA, const float4* pairB) {
  const IcpProblem& P = probs[blockIdx.x];
  IcpState& S = states[blockIdx.x];
  if (S.done) return;
  if (!after_match && !S.inner_pending) return;
  __shared__ SolveScratch sc;
  __shared__ SolveStage st;
  const uint32_t it = S.it;  // (the match phase of this iteration used the same index; read before the solve bumps it)
  sum_partials_block(P, partials, part_cnt, after_match ? P.n_blocks : P.n_blocks_acc, sc, &S, &st);
  __shared__ PairStage ps;
  const int next = solve_phase(P, S, st, sc, after_match, fuse, it, local, pairA, pairB, &ps);
  if (next == 0 && threadIdx.x == 0) atomicSub(n_active, 1u);
}

__global__ void k_init_states(const IcpProblem* __restrict__ probs, IcpState* __restrict__ states, const double* __restrict__ init_poses,
                              uint32_t n, uint32_t* n_active) {
  const uint32_t b = blockIdx.x * blockDim.x + threadIdx.x;
  if (b >= n) return;
  IcpState& S = states[b];
  for (int k = 0; k < 12; k++) {
    S.T[k] = init_poses[12 * size_t(b) + k];
    S.prev[k] = S.T[k];
    S.prev2[k] = S.T[k];
  }
  for (int k = 0; k < 36; k++) S.H[k] = 0.0;
  S.has_prev2 = 0;
  S.prior_valid = 0;
  S.have_H = 0;
  S.it = 0;
  S.inner_pending = 0;
  S.inner = 0;
  S.n_pairs = S.n_potential = S.n_query_it = S.n_cand = 0;
  S.term = MLO_TERM_UNDEFINED;
  S.done = 0;
  if (probs[b].max_iterations == 0) {  // ICP::align with an exhausted budget (LidarOdometry.cpp:956-967)
    S.term = MLO_TERM_MAX_ITERATIONS;
    S.done = 1;
  } else if (probs[b].n_q == 0) {      // empty local cloud: the matcher produces nothing
    S.term = MLO_TERM_NO_PAIRINGS;
    S.done = 1;
  } else {
    atomicAdd(n_active, 1u);
  }
}


// ------------------------------------------------------------------ persistent, queue-driven align
// One launch runs the whole ICP::align loop of a batch.  Work items = (problem, phase, chunk); a bounded
// MPMC ring in global memory (ticket + per-slot sequence numbers) feeds resident thread blocks.  The block
// that completes the last chunk of a phase runs solve_step for that problem and publishes the chunks of the
// next phase (inner GN re-linearisation, or the match phase of the next ICP iteration), so problems advance
// independently: no grid-wide barrier, no host round trip, no idle tail while other problems still iterate.
struct IcpQueue {
  // one 64-bit word per slot: (sequence number << 32) | item.  Sequence = ticket of the producer that may fill the slot
  // (free), ticket + 1 once filled: a consumer polls ONE word and has the item with it (it used to be a sequence array and
  // an item array: one more dependent L2 round trip per pop, one more fence per push).
  unsigned long long* slots;
  uint32_t mask;       // ring capacity - 1
  uint32_t* ctrl;      // [0] head, [1] tail, [2] problems active, [3] all done, [4] error/timeout
  uint32_t* phase_cnt; // per problem: chunks of the current phase completed
};
constexpr uint32_t ITEM_EXIT = 0xFFFFFFFFu;
MLO_HD uint32_t item_make(uint32_t prob, uint32_t chunk, uint32_t phase) { return (prob << 16) | (chunk << 1) | phase; }

MLO_D uint32_t ld_volatile_u32(const uint32_t* p) { return *reinterpret_cast<const volatile uint32_t*>(p); }
MLO_D unsigned long long ld_volatile_u64(const unsigned long long* p) { return *reinterpret_cast<const volatile unsigned long long*>(p); }
MLO_D void st_volatile_u64(unsigned long long* p, unsigned long long v) { *reinterpret_cast<volatile unsigned long long*>(p) = v; }
MLO_HD unsigned long long slot_word(uint32_t seq, uint32_t item) { return (static_cast<unsigned long long>(seq) << 32) | item; }

// lanes of one warp publish n consecutive items (prob, phase, chunk 0..n-1)
MLO_D void queue_push(const IcpQueue& q, uint32_t prob, uint32_t phase, uint32_t n) {
  const uint32_t lane = threadIdx.x & 31u;
  uint32_t pos = 0;
  if (lane == 0) pos = atomicAdd(&q.ctrl[1], n);
  pos = __shfl_sync(0xFFFFFFFFu, pos, 0);
  for (uint32_t i = lane; i < n; i += 32) {
    const uint32_t t = pos + i, slot = t & q.mask;
    uint32_t spins = 0;
    while (uint32_t(ld_volatile_u64(&q.slots[slot]) >> 32) != t) {  // slot still holds an unconsumed older item (ring full): rare
      __nanosleep(64);
      if (++spins > (1u << 24)) {
        atomicExch(&q.ctrl[4], 1u);
        atomicExch(&q.ctrl[3], 1u);
        break;
      }
    }
    st_volatile_u64(&q.slots[slot], slot_word(t + 1, item_make(prob, i, phase)));  // (item and sequence in one store)
  }
}

// out-of-line copies for the persistent kernel: each phase keeps its own register allocation
template <int TAG>
__device__ __noinline__ void chunk_match_tpq_ool(const MapDev& map, const IcpProblem& P, const double* sT, uint32_t it,
                                                 uint32_t chunk, const float4* local, float4* pairA, float4* pairB,
                                                 double* partials, uint32_t* part_cnt) {
  chunk_match_tpq(map, P, sT, it, chunk, local, pairA, pairB, partials, part_cnt);
}
template <int TAG, bool PLANES>
__device__ __noinline__ void chunk_match_warp_ool(const MapDev& map, const IcpProblem& P, const double* sT, uint32_t it,
                                                  uint32_t chunk, const float4* local, float4* pairA, float4* pairB,
                                                  double* partials, uint32_t* part_cnt, uint32_t qpw) {
  chunk_match_warp<PLANES>(map, P, sT, it, chunk, local, pairA, pairB, partials, part_cnt, qpw);
}
template <int TAG>
__device__ __noinline__ void chunk_accumulate_ool(const IcpProblem& P, const double* sT, uint32_t it, uint32_t chunk,
                                                  const float4* local, const float4* pairA, const float4* pairB,
                                                  double* partials, uint32_t* part_cnt) {
  chunk_accumulate(P, sT, it, chunk, local, pairA, pairB, partials, part_cnt);
}

// Build the work queue on the device from the current problem states (used when a launch sequence hands its
// tail of still-active problems over to the persistent kernel): every slot first gets seq = index (free), then
// each active problem appends the chunks of its next match phase.
__global__ void k_queue_reset(IcpQueue q, uint32_t n_problems) {
  const uint32_t i = blockIdx.x * blockDim.x + threadIdx.x;
  if (i <= q.mask) q.slots[i] = slot_word(i, 0u);
  if (i < n_problems) q.phase_cnt[i] = 0;
  if (i < 8) q.ctrl[i] = 0;
}
__global__ void k_queue_build(const IcpProblem* __restrict__ probs, const IcpState* __restrict__ states, IcpQueue q,
                              uint32_t n_problems) {
  const uint32_t b = blockIdx.x * blockDim.x + threadIdx.x;
  if (b >= n_problems || states[b].done) return;
  const uint32_t n = probs[b].n_blocks_pers;
  const uint32_t pos = atomicAdd(&q.ctrl[1], n);
  for (uint32_t i = 0; i < n; i++) {
    q.slots[(pos + i) & q.mask] = slot_word(pos + i + 1, item_make(b, i, 0u));
  }
  atomicAdd(&q.ctrl[2], 1u);
}

template <bool TPQ, bool MULTI, int MINB = 4, bool PLANES = true>
__global__ void __launch_bounds__(ICP_BLOCK, MINB)
    k_icp_persistent(MapDev map, const MapDev* __restrict__ maps, const IcpProblem* __restrict__ probs, IcpState* states,
                     const float4* __restrict__ local, float4* pairA, float4* pairB, double* partials, uint32_t* part_cnt,
                     IcpQueue q, uint32_t qpw, int fuse) {
  __shared__ MapDev sMap;
  __shared__ SolveScratch s_solve;
  __shared__ SolveStage s_stage;
  __shared__ PairStage s_pairs;
  __shared__ uint32_t s_item;
  __shared__ int s_last;
  __shared__ double sT[12];
  __shared__ uint32_t s_it;
  for (;;) {
    if (threadIdx.x == 0) {
      const uint32_t t = atomicAdd(&q.ctrl[0], 1u);
      const uint32_t slot = t & q.mask;
      uint32_t item = ITEM_EXIT, spins = 0;
      for (;;) {
        const unsigned long long v = ld_volatile_u64(&q.slots[slot]);
        if (uint32_t(v >> 32) == t + 1) {
          item = uint32_t(v);
          st_volatile_u64(&q.slots[slot], slot_word(t + q.mask + 1, 0u));  // free the slot for ticket t + capacity
          __threadfence();  // acquire: what the producer wrote before publishing (problem state) is visible below
          break;
        }
        if (ld_volatile_u32(&q.ctrl[3])) break;  // every problem finished
        __nanosleep(128);
        if (++spins > (1u << 24)) {  // safety net (~seconds): never hang the device
          atomicExch(&q.ctrl[4], 1u);
          atomicExch(&q.ctrl[3], 1u);
          break;
        }
      }
      s_item = item;
    }
    __syncthreads();
    const uint32_t item = s_item;
    if (item == ITEM_EXIT) return;
    const uint32_t prob = item >> 16, chunk = (item >> 1) & 0x7FFFu, phase = item & 1u;
    if (chunk == 0) MLO_TRACE_EVENT(prob, 1 + phase);  // 1: first match chunk popped, 2: first accumulate chunk popped
    const IcpProblem& P = probs[prob];
    IcpState& S = states[prob];
    if (threadIdx.x < 12) sT[threadIdx.x] = __ldcg(&S.T[threadIdx.x]);
    if (threadIdx.x == 12) s_it = __ldcg(&S.it);
    if (MULTI && phase == 0) stage_map(sMap, maps, P.map_idx);
    __syncthreads();
    if (phase == 0) {
      constexpr int TAG = (TPQ ? 2 : 0) + (MULTI ? 1 : 0) + (MINB == 4 ? 0 : 4) + (PLANES ? 0 : 8);  // one out-of-line copy per kernel instance
      if constexpr (TPQ && MULTI) chunk_match_tpq_ool<TAG>(sMap, P, sT, s_it, chunk, local, pairA, pairB, partials, part_cnt);
      else if constexpr (TPQ) chunk_match_tpq_ool<TAG>(map, P, sT, s_it, chunk, local, pairA, pairB, partials, part_cnt);
      else if constexpr (MULTI) chunk_match_warp_ool<TAG, PLANES>(sMap, P, sT, s_it, chunk, local, pairA, pairB, partials, part_cnt, qpw);
      else chunk_match_warp_ool<TAG, PLANES>(map, P, sT, s_it, chunk, local, pairA, pairB, partials, part_cnt, qpw);
    } else {
      chunk_accumulate_ool<(TPQ ? 2 : 0) + (MULTI ? 1 : 0) + (MINB == 4 ? 0 : 4) + (PLANES ? 0 : 8)>(P, sT, s_it, chunk, local, pairA, pairB, partials, part_cnt);
    }
    __syncthreads();
    if (threadIdx.x == 0) {
      __threadfence();  // partials / pairings of this chunk (ordered by the barrier) visible before the count
      const uint32_t nblk = phase == 0 ? P.n_blocks_pers : P.n_blocks_acc;
      const uint32_t old = atomicAdd(&q.phase_cnt[prob], 1u);
      s_last = (old + 1 == nblk);
      if (s_last) atomicExch(&q.phase_cnt[prob], 0u);
    }
    __syncthreads();
    if (chunk == 0) MLO_TRACE_EVENT(prob, 3);  // chunk 0 of the phase finished
    if (s_last) {
      MLO_TRACE_EVENT(prob, 4);  // last chunk of the phase finished
      __threadfence();  // acquire: the other blocks' partials (read through L2) and the problem state
      sum_partials_block(P, partials, part_cnt, phase == 0 ? P.n_blocks_pers : P.n_blocks_acc, s_solve, &S, &s_stage);
      MLO_TRACE_EVENT(prob, 5);  // partials summed
    }
    int nx = 0;  // (block-uniform)
    if (s_last) {
      nx = solve_phase(P, S, s_stage, s_solve, phase == 0, fuse, s_it, local, pairA, pairB, &s_pairs);
      MLO_TRACE_EVENT(prob, 7);  // solve phase (with its fused inner iterations) done, state written back
    }
    if (s_last && threadIdx.x < 32) {
      const int next = nx;
      if (next == 1) {
        queue_push(q, prob, 1u, P.n_blocks_acc);
      } else if (next == 2) {
        queue_push(q, prob, 0u, P.n_blocks_pers);
        MLO_TRACE_EVENT(prob, 8);  // next match phase published
      } else if (threadIdx.x == 0) {
        if (atomicSub(&q.ctrl[2], 1u) == 1u) atomicExch(&q.ctrl[3], 1u);
      }
    }
    __syncthreads();
  }
}

}  // namespace mlo
