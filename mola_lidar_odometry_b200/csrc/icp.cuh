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
};

struct IcpState {
  double T[12], prev[12], prev2[12];
  double H[36];
  int32_t has_prev2, done, term, have_H;
  uint32_t it;
  int32_t inner_pending;  // 1: another inner GN iteration must run for the current ICP iteration
  uint32_t inner;         // index of the next inner iteration
  uint64_t n_pairs, n_potential, n_query_it, n_cand;
};

MLO_D double table_at(const double* t, uint32_t len, uint32_t it) {
  if (!t || len == 0) return 0.0;
  return t[it < len ? it : len - 1];
}

MLO_D double robust_weight(int kernel, double e2, double c) {
  if (kernel == MLO_KERNEL_GEMAN_MCCLURE) {
    const double c2 = c * c, d = e2 + c2;
    return (c2 * c2) / (d * d);
  }
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

// pair record: A = (gx, gy, gz | cx, cy, cz, kind) with kind 0 none, 1 pt2pt, 2 pt2pl; B = plane normal.
//
// Work decomposition: a warp owns `qpw` consecutive queries (qpw = 32 for throughput, smaller for
// latency-bound small batches).  Lane t holds query t: its local point, its transformed point and, after
// the warp-cooperative NN of that query (map.cuh: probe prefetched one query ahead), its pairing.  The
// normal-equation terms are then computed once per query by the owning lane and butterfly-reduced.
__global__ void __launch_bounds__(ICP_BLOCK)
    k_match_accumulate(MapDev map, const IcpProblem* __restrict__ probs, const IcpState* __restrict__ states,
                       const float4* __restrict__ local, float4* __restrict__ pairA, float4* __restrict__ pairB,
                       double* __restrict__ partials, uint32_t* __restrict__ part_cnt, uint32_t qpw) {
  const IcpProblem& P = probs[blockIdx.y];
  if (blockIdx.x >= P.n_blocks) return;
  const IcpState& S = states[blockIdx.y];
  if (S.done) return;
  __shared__ double sT[12];
  if (threadIdx.x < 12) sT[threadIdx.x] = S.T[threadIdx.x];
  __syncthreads();
  const uint32_t FULL = 0xFFFFFFFFu;
  const uint32_t it = S.it;
  const double thr = table_at(P.thr_pt2pt, P.table_len, it);
  const float thr2 = float(thr * thr);
  const float thr_pl = float(table_at(P.thr_pt2pl, P.table_len, it));
  const double kc = table_at(P.kparam, P.table_len, it);
  const uint32_t lane = threadIdx.x & 31u, warp = threadIdx.x >> 5;

  const uint32_t qbase = (blockIdx.x * (ICP_BLOCK / 32) + warp) * qpw;  // first query of this warp
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
  const uint32_t pbi = P.part_begin + blockIdx.x;
  block_reduce_store(a, npairs, ncand, partials + size_t(pbi) * NACC, part_cnt + 2 * size_t(pbi));
}

// Thread-per-query variant for large batches: one query per thread, map.cuh nn_single_thread (pruned,
// 256-bit loads).  Same outputs and the same block-partial layout as k_match_accumulate.
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
  const uint32_t it = S.it;
  const double thr = table_at(P.thr_pt2pt, P.table_len, it);
  const float thr2 = float(thr * thr);
  const float thr_pl = float(table_at(P.thr_pt2pl, P.table_len, it));
  const double kc = table_at(P.kparam, P.table_len, it);
  double a[NACC];
#pragma unroll
  for (int k = 0; k < int(NACC); k++) a[k] = 0.0;
  uint32_t npairs = 0, ncand = 0;
  const uint32_t q = blockIdx.x * ICP_BLOCK + threadIdx.x;
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
      const NNHit h = nn_single_thread(map, gx, gy, gz);
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
  const uint32_t pbi = P.part_begin + blockIdx.x;
  block_reduce_store(a, npairs, ncand, partials + size_t(pbi) * NACC, part_cnt + 2 * size_t(pbi));
}

__global__ void __launch_bounds__(ICP_BLOCK)
    k_accumulate(const IcpProblem* __restrict__ probs, const IcpState* __restrict__ states, const float4* __restrict__ local,
                 const float4* __restrict__ pairA, const float4* __restrict__ pairB, double* __restrict__ partials,
                 uint32_t* __restrict__ part_cnt) {
  const IcpProblem& P = probs[blockIdx.y];
  if (blockIdx.x >= P.n_blocks_acc) return;
  const IcpState& S = states[blockIdx.y];
  if (S.done || !S.inner_pending) return;
  __shared__ double sT[12];
  if (threadIdx.x < 12) sT[threadIdx.x] = S.T[threadIdx.x];
  __syncthreads();
  const double kc = table_at(P.kparam, P.table_len, S.it);
  double a[NACC];
#pragma unroll
  for (int k = 0; k < int(NACC); k++) a[k] = 0.0;
  uint32_t npairs = 0;
  const uint32_t q = blockIdx.x * ICP_BLOCK + threadIdx.x;
  if (q < P.n_q) {
    const float4 pa = pairA[P.q_begin + q];
    if (pa.w != 0.f) {
      const float4 l = __ldg(&local[P.q_begin + q]);
      if (pa.w == 1.f) {
        contrib_pt2pt(sT, l.x, l.y, l.z, pa.x, pa.y, pa.z, P.w_pt2pt, P.robust_kernel, kc, a);
      } else {
        const float4 nb = pairB[P.q_begin + q];
        contrib_pt2pl(sT, l.x, l.y, l.z, pa.x, pa.y, pa.z, nb.x, nb.y, nb.z, P.w_pt2pl, P.robust_kernel, kc, a);
      }
      npairs++;
    }
  }
  const uint32_t pb = P.part_begin + blockIdx.x;
  block_reduce_store(a, npairs, 0u, partials + size_t(pb) * NACC, part_cnt + 2 * size_t(pb));
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

// End-of-iteration bookkeeping of mp2p_icp::ICP::align: step measure against prev and
// prev-prev, hook-as-data, stall test, iteration counter, MaxIterations.
MLO_D void finish_iteration(const IcpProblem& P, IcpState& S) {
  double D[12], d[6];
  pose_minus(S.T, S.prev, D);
  se3_log(D, d);
  double dt = nrm3(d), dr = nrm3(d + 3);
  if (S.has_prev2) {
    double d2[6];
    pose_minus(S.T, S.prev2, D);
    se3_log(D, d2);
    dt = fmin(dt, nrm3(d2));
    dr = fmin(dr, nrm3(d2 + 3));
  }
  for (int k = 0; k < 12; k++) {
    S.prev2[k] = S.prev[k];
    S.prev[k] = S.T[k];
  }
  S.has_prev2 = 1;
  S.inner_pending = 0;
  if (P.hook_enabled) {
    double w[3];
    pose_minus(S.T, P.hook_checkpoint, D);
    so3_log_of_pose(D, w);
    const double tt[3] = {D[3], D[7], D[11]};
    if (nrm3(tt) > P.hook_min_trans || nrm3(w) > P.hook_min_rot) {
      S.term = MLO_TERM_HOOK_REQUEST;
      S.done = 1;
      return;
    }
  }
  if (fabs(dt) < P.min_abs_step_trans && fabs(dr) < P.min_abs_step_rot) {
    S.term = MLO_TERM_STALLED;
    S.done = 1;
    return;
  }
  S.it++;
  if (S.it >= P.max_iterations) {
    S.term = MLO_TERM_MAX_ITERATIONS;
    S.done = 1;
  }
}

// one warp per problem. `after_match` = 1 when the partials come from k_match_accumulate (inner 0).
__global__ void __launch_bounds__(32)
    k_solve(const IcpProblem* __restrict__ probs, IcpState* __restrict__ states, const double* __restrict__ partials,
            const uint32_t* __restrict__ part_cnt, int after_match, uint32_t* __restrict__ n_active) {
  const IcpProblem& P = probs[blockIdx.x];
  IcpState& S = states[blockIdx.x];
  if (S.done) return;
  if (!after_match && !S.inner_pending) return;
  const uint32_t lane = threadIdx.x;
  // ordered sum over this problem's block partials: lane k owns element k
  double acc = 0.0;
  uint32_t cnt = 0;
  const uint32_t nblk = after_match ? P.n_blocks : P.n_blocks_acc;
  if (lane < NACC) {
    for (uint32_t b = 0; b < nblk; b++) acc += partials[size_t(P.part_begin + b) * NACC + lane];
  } else if (lane < NACC + 2) {
    for (uint32_t b = 0; b < nblk; b++) cnt += part_cnt[2 * size_t(P.part_begin + b) + (lane - NACC)];
  }
  double a[NACC];
#pragma unroll
  for (int k = 0; k < int(NACC); k++) a[k] = __shfl_sync(0xFFFFFFFFu, acc, k);
  const uint32_t npairs = __shfl_sync(0xFFFFFFFFu, cnt, NACC);
  const uint32_t ncand = __shfl_sync(0xFFFFFFFFu, cnt, NACC + 1);
  if (lane != 0) return;

  if (after_match) {
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
      atomicSub(n_active, 1u);
      return;
    }
  }
  bool ok = true;
  bool last_inner = true;
  if (P.solver == MLO_SOLVER_HORN) {
    ok = horn_from_sums(a, double(npairs), S.T);
  } else {
    double H[36], g[6];
    int k = 0;
    for (int i = 0; i < 6; i++)
      for (int j = i; j < 6; j++) {
        H[6 * i + j] = a[k];
        H[6 * j + i] = a[k];
        k++;
      }
    for (int i = 0; i < 6; i++) g[i] = a[21 + i];
    if (P.has_prior) {
      // e = log(prior^-1 T), J = d log(D exp(eps))/d eps ; g += J^T L e ; H += J^T L J
      double D[12], e[6], J[36], LJ[36], Le[6];
      pose_minus(S.T, P.prior_pose, D);
      se3_log(D, e);
      se3_right_jacobian_inv(e, J);
      for (int i = 0; i < 6; i++) {
        double s = 0;
        for (int m = 0; m < 6; m++) s += P.prior_info[6 * i + m] * e[m];
        Le[i] = s;
        for (int j = 0; j < 6; j++) {
          double t = 0;
          for (int m = 0; m < 6; m++) t += P.prior_info[6 * i + m] * J[6 * m + j];
          LJ[6 * i + j] = t;
        }
      }
      for (int i = 0; i < 6; i++) {
        for (int m = 0; m < 6; m++) g[i] += J[6 * m + i] * Le[m];
        for (int j = 0; j < 6; j++)
          for (int m = 0; m < 6; m++) H[6 * i + j] += J[6 * m + i] * LJ[6 * m + j];
      }
    }
    for (int i = 0; i < 36; i++) S.H[i] = H[i];
    S.have_H = 1;
    double mg[6], delta[6];
    for (int i = 0; i < 6; i++) mg[i] = -g[i];
    ok = ldlt6(H, mg, delta);
    if (ok) {
      double E[12], Tn[12];
      se3_exp(delta, E);
      pose_mul(S.T, E, Tn);
      for (int i = 0; i < 12; i++) S.T[i] = Tn[i];
      double dn = 0;
      for (int i = 0; i < 6; i++) dn += delta[i] * delta[i];
      S.inner++;
      last_inner = (sqrt(dn) < P.gn_min_delta) || (S.inner >= P.gn_max_iterations);
    }
  }
  if (!ok) {
    S.term = MLO_TERM_SOLVER_ERROR;
    S.done = 1;
    atomicSub(n_active, 1u);
    return;
  }
  if (!last_inner) {
    S.inner_pending = 1;
    return;
  }
  finish_iteration(P, S);
  if (S.done) atomicSub(n_active, 1u);
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
  } else {
    atomicAdd(n_active, 1u);
  }
}

}  // namespace mlo
