// synth_gpu.cu — the ray caster of synth.cpp as a CUDA kernel (INPUT GENERATOR ONLY, no part of the registration path).
//
// bench.py's whole-sequence workloads (BASELINE.json configs[2]-[4]: 32 sequences x 1000 scans per GPU, 230 k-point
// O128 sweeps) need tens of thousands of synthetic scans; ray casting them on the host cores (39 ms per scan) would take
// longer than the benchmark.  Same scene (exported from synth.cpp), same sensors, same noise hash; the returns of a
// sweep come out in the same azimuth-major order.  Results are NOT bit-identical to the CPU generator (libm vs CUDA
// sin / cos / log): the two arms of a benchmark must both read the clouds generated here.
//
//   synth_gpu_scan_batch  n_scans sweeps in one launch: block = (scan, 256 rays); every ray tests the primitives within
//                         max_range of its sensor origin (culled per scan by k_cull into a compact list).
//   outputs               dense float4 [n_scans][n_rays] (x, y, z, intensity; w < 0 marks "no return") - the caller
//                         compacts (order preserving).
#include <cuda_runtime.h>
#include <math.h>
#include <stdint.h>

namespace {

struct Box { float x0, y0, x1, y1, z0, z1; };
struct Cyl { float cx, cy, r, z0, z1; };
struct Sph { float cx, cy, cz, r; };

__host__ __device__ inline uint64_t splitmix(uint64_t x) {
  x += 0x9E3779B97F4A7C15ull;
  x = (x ^ (x >> 30)) * 0xBF58476D1CE4E5B9ull;
  x = (x ^ (x >> 27)) * 0x94D049BB133111EBull;
  return x ^ (x >> 31);
}
__device__ inline double u01(uint64_t h) { return (double(h >> 11) + 0.5) * (1.0 / 9007199254740992.0); }

__device__ inline bool hit_box(const Box& b, const double o[3], const double d[3], double& t_out) {
  double t0 = 0.0, t1 = 1e30;
  const double lo[3] = {b.x0, b.y0, b.z0}, hi[3] = {b.x1, b.y1, b.z1};
#pragma unroll
  for (int k = 0; k < 3; k++) {
    if (fabs(d[k]) < 1e-12) {
      if (o[k] < lo[k] || o[k] > hi[k]) return false;
    } else {
      double a = (lo[k] - o[k]) / d[k], c = (hi[k] - o[k]) / d[k];
      if (a > c) { const double s = a; a = c; c = s; }
      if (a > t0) t0 = a;
      if (c < t1) t1 = c;
      if (t0 > t1) return false;
    }
  }
  if (t0 <= 1e-6) return false;
  t_out = t0;
  return true;
}
__device__ inline bool hit_cyl(const Cyl& c, const double o[3], const double d[3], double& t_out) {
  const double ox = o[0] - c.cx, oy = o[1] - c.cy;
  const double a = d[0] * d[0] + d[1] * d[1];
  if (a < 1e-12) return false;
  const double b = ox * d[0] + oy * d[1];
  const double cc = ox * ox + oy * oy - double(c.r) * c.r;
  const double disc = b * b - a * cc;
  if (disc < 0) return false;
  const double t = (-b - sqrt(disc)) / a;
  if (t <= 1e-6) return false;
  const double z = o[2] + t * d[2];
  if (z < c.z0 || z > c.z1) return false;
  t_out = t;
  return true;
}
__device__ inline bool hit_sph(const Sph& s, const double o[3], const double d[3], double& t_out) {
  const double ox = o[0] - s.cx, oy = o[1] - s.cy, oz = o[2] - s.cz;
  const double b = ox * d[0] + oy * d[1] + oz * d[2];
  const double c = ox * ox + oy * oy + oz * oz - double(s.r) * s.r;
  const double disc = b * b - c;
  if (disc < 0) return false;
  const double t = -b - sqrt(disc);
  if (t <= 1e-6) return false;
  t_out = t;
  return true;
}

// per scan: indices of the primitives whose footprint comes within max_range of the sensor origin
__global__ void k_cull(const Box* boxes, int nb, const Cyl* cyls, int nc, const Sph* sphs, int ns, const double* poses, int n_scans,
                       double max_range, int cap, int* lists, int* counts) {
  const int s = blockIdx.y;
  const int i = blockIdx.x * blockDim.x + threadIdx.x;
  const int total = nb + nc + ns;
  if (i >= total) return;
  const double ox = poses[12 * s + 3], oy = poses[12 * s + 7];
  double cx, cy, rad;
  if (i < nb) {
    const Box b = boxes[i];
    cx = 0.5 * (double(b.x0) + b.x1); cy = 0.5 * (double(b.y0) + b.y1); rad = 0.5 * hypot(double(b.x1) - b.x0, double(b.y1) - b.y0);
  } else if (i < nb + nc) {
    const Cyl c = cyls[i - nb];
    cx = c.cx; cy = c.cy; rad = c.r;
  } else {
    const Sph p = sphs[i - nb - nc];
    cx = p.cx; cy = p.cy; rad = p.r;
  }
  if (hypot(cx - ox, cy - oy) - rad > max_range) return;
  const int k = atomicAdd(&counts[s], 1);
  if (k < cap) lists[size_t(s) * cap + k] = i;
}

__global__ void __launch_bounds__(256) k_scan(const Box* boxes, int nb, const Cyl* cyls, int nc, const Sph* sphs, const double* poses,
                                              const unsigned long long* seeds, int n_beams, int n_az, double el_top_deg,
                                              double el_bot_deg, double max_range, double noise_sigma, int cap, const int* lists,
                                              const int* counts, float4* out) {
  const int s = blockIdx.y;
  const int n_rays = n_beams * n_az;
  const int ray = blockIdx.x * blockDim.x + threadIdx.x;
  extern __shared__ int s_list[];  // this scan's culled primitive ids, staged in chunks
  const int n_prim = min(counts[s], cap);
  const double* pose = poses + 12 * size_t(s);
  const double o[3] = {pose[3], pose[7], pose[11]};
  const int a = ray / n_beams, b = ray % n_beams;  // azimuth-major, like synth.cpp
  const double az = -M_PI + 2.0 * M_PI * (double(a) + 0.5) / n_az;
  const double el = (el_top_deg + (el_bot_deg - el_top_deg) * (n_beams > 1 ? double(b) / (n_beams - 1) : 0.0)) * M_PI / 180.0;
  const double ds[3] = {cos(el) * cos(az), cos(el) * sin(az), sin(el)};
  double d[3];
#pragma unroll
  for (int k = 0; k < 3; k++) d[k] = pose[4 * k] * ds[0] + pose[4 * k + 1] * ds[1] + pose[4 * k + 2] * ds[2];
  double best = 1e30;
  if (d[2] < -1e-9) {
    const double t = -o[2] / d[2];
    if (t > 0) best = t;
  }
  constexpr int CHUNK = 1024;
  for (int base = 0; base < n_prim; base += CHUNK) {
    const int m = min(CHUNK, n_prim - base);
    __syncthreads();
    for (int k = threadIdx.x; k < m; k += blockDim.x) s_list[k] = lists[size_t(s) * cap + base + k];
    __syncthreads();
    if (ray < n_rays) {
      for (int k = 0; k < m; k++) {
        const int id = s_list[k];
        double t;
        bool h;
        if (id < nb) h = hit_box(boxes[id], o, d, t);
        else if (id < nb + nc) h = hit_cyl(cyls[id - nb], o, d, t);
        else h = hit_sph(sphs[id - nb - nc], o, d, t);
        if (h && t < best) best = t;
      }
    }
  }
  if (ray >= n_rays) return;
  float4 r4 = make_float4(0.f, 0.f, 0.f, -1.f);
  if (best <= max_range) {
    const uint64_t scan_seed = seeds[s];
    const uint64_t h1 = splitmix(scan_seed * 0x100000001B3ull + uint64_t(ray) * 2 + 1), h2 = splitmix(h1 + 0x632BE59BD9B4E019ull);
    const double g = sqrt(-2.0 * log(u01(h1))) * cos(2.0 * M_PI * u01(h2));
    const double r = best + noise_sigma * g;
    if (r >= 0.5) r4 = make_float4(float(r * ds[0]), float(r * ds[1]), float(r * ds[2]), float(u01(splitmix(h2 + 7))));
  }
  out[size_t(s) * n_rays + ray] = r4;
}

}  // namespace

extern "C" {

// All pointers are DEVICE pointers (the caller owns them, e.g. torch tensors); `stream` is a cudaStream_t (0 = default).
// lists: int [n_scans * cap], counts: int [n_scans] scratch.  Returns 0 or a cudaError_t.
int synth_gpu_scan_batch(const void* boxes, int nb, const void* cyls, int nc, const void* sphs, int ns, const double* poses,
                         const unsigned long long* seeds, int n_scans, int n_beams, int n_az, double el_top_deg, double el_bot_deg,
                         double max_range, double noise_sigma, int cap, int* lists, int* counts, void* out_xyzi, void* stream) {
  cudaStream_t st = static_cast<cudaStream_t>(stream);
  cudaMemsetAsync(counts, 0, sizeof(int) * n_scans, st);
  const int total = nb + nc + ns;
  k_cull<<<dim3((total + 255) / 256, n_scans), 256, 0, st>>>(static_cast<const Box*>(boxes), nb, static_cast<const Cyl*>(cyls), nc,
                                                             static_cast<const Sph*>(sphs), ns, poses, n_scans, max_range, cap, lists,
                                                             counts);
  const int n_rays = n_beams * n_az;
  k_scan<<<dim3((n_rays + 255) / 256, n_scans), 256, 1024 * sizeof(int), st>>>(
      static_cast<const Box*>(boxes), nb, static_cast<const Cyl*>(cyls), nc, static_cast<const Sph*>(sphs), poses, seeds, n_beams, n_az,
      el_top_deg, el_bot_deg, max_range, noise_sigma, cap, lists, counts, static_cast<float4*>(out_xyzi));
  return int(cudaGetLastError());
}

}  // extern "C"
