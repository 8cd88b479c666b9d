// synth.cpp — seeded synthetic KITTI-shaped LiDAR data (SURVEY.md §8(d) "Configs restated as
// concrete synthetic inputs").  Input generator only: no part of the registration path.
//
//   Scene S(seed): ground plane z = 0 (sensor rides at z = 1.73 m), a 1.5 km x 1.5 km street grid of
//   40 m pitch with 12 m wide streets, 3-6 axis-aligned "building" boxes per block (stepped facades),
//   kerb-side trees (trunk cylinder + spherical crown), vertical poles on the street edges and
//   parked-car boxes along the kerbs.
//   Sensor K64: 64 beams, elevation linspace(+2.0, -24.8) deg, 2048 azimuth steps, 120 m.
//   Sensor O128: 128 beams, +-22.5 deg, 1800 azimuth steps, 120 m.
//   Range noise N(0, sigma) from a counter-based hash of (scan seed, ray index): order independent.
//
// Output layout = KITTI velodyne .bin: x, y, z, intensity (float32 x4) in the SENSOR frame.
#include <cmath>
#include <cstdint>
#include <cstring>
#include <vector>

namespace {

struct Box {
  float x0, y0, x1, y1, z0, z1;
};
struct Cyl {
  float cx, cy, r, z0, z1;
};
struct Sph {
  float cx, cy, cz, r;
};

inline uint64_t splitmix(uint64_t x) {
  x += 0x9E3779B97F4A7C15ull;
  x = (x ^ (x >> 30)) * 0xBF58476D1CE4E5B9ull;
  x = (x ^ (x >> 27)) * 0x94D049BB133111EBull;
  return x ^ (x >> 31);
}
struct Rng {
  uint64_t s;
  explicit Rng(uint64_t seed) : s(splitmix(seed)) {}
  uint64_t next() { return s = splitmix(s); }
  double uni() { return double(next() >> 11) * (1.0 / 9007199254740992.0); }
  double uni(double a, double b) { return a + (b - a) * uni(); }
};
inline double u01(uint64_t h) { return (double(h >> 11) + 0.5) * (1.0 / 9007199254740992.0); }

struct Scene {
  std::vector<Box> boxes;
  std::vector<Cyl> cyls;
  std::vector<Sph> sphs;
  float extent;
};

constexpr float GRID = 40.f, STREET = 12.f;

Scene* make_scene(uint64_t seed, float extent, int n_poles, int n_cars) {
  auto* s = new Scene;
  s->extent = extent;
  Rng r(seed);
  const int nb = int(extent / GRID);
  const float lo = STREET / 2, hi = GRID - STREET / 2;  // block spans [lo, hi] inside each cell
  for (int i = 0; i < nb; i++)
    for (int j = 0; j < nb; j++) {
      const int nbld = 3 + int(r.uni() * 4);
      for (int b = 0; b < nbld; b++) {
        const float w = float(r.uni(6, 16)), d = float(r.uni(6, 16)), h = float(r.uni(4, 20));
        const float ox = float(r.uni(lo, hi - w)), oy = float(r.uni(lo, hi - d));
        s->boxes.push_back(Box{i * GRID + ox, j * GRID + oy, i * GRID + ox + w, j * GRID + oy + d, 0.f, h});
      }
      // kerb-side trees: 2-3 per block side, 1 m outside the block edge
      for (int side = 0; side < 4; side++) {
        const int nt = 2 + int(r.uni() * 2);
        for (int t = 0; t < nt; t++) {
          const float u = float(r.uni(lo, hi)), off = 1.0f;
          float x, y;
          if (side == 0) { x = i * GRID + u; y = j * GRID + lo - off; }
          else if (side == 1) { x = i * GRID + u; y = j * GRID + hi + off; }
          else if (side == 2) { x = i * GRID + lo - off; y = j * GRID + u; }
          else { x = i * GRID + hi + off; y = j * GRID + u; }
          const float th = float(r.uni(3, 5)), cr = float(r.uni(1.2, 2.2));
          s->cyls.push_back(Cyl{x, y, float(r.uni(0.15, 0.3)), 0.f, th});
          s->sphs.push_back(Sph{x, y, th + cr * 0.6f, cr});
        }
      }
    }
  for (int k = 0; k < n_poles; k++) {
    const int i = int(r.uni() * nb), j = int(r.uni() * nb), side = int(r.uni() * 4);
    const float u = float(r.uni(lo, hi));
    float x, y;
    const float off = 0.6f;
    if (side == 0) { x = i * GRID + u; y = j * GRID + lo - off; }
    else if (side == 1) { x = i * GRID + u; y = j * GRID + hi + off; }
    else if (side == 2) { x = i * GRID + lo - off; y = j * GRID + u; }
    else { x = i * GRID + hi + off; y = j * GRID + u; }
    s->cyls.push_back(Cyl{x, y, 0.15f, 0.f, 6.f});
  }
  for (int k = 0; k < n_cars; k++) {
    const int i = int(r.uni() * nb), j = int(r.uni() * nb), side = int(r.uni() * 4);
    const float u = float(r.uni(lo + 3, hi - 3));
    const float L = 4.5f, W = 1.8f, H = 1.5f, off = 2.2f;
    float cx, cy, hx, hy;
    if (side == 0) { cx = i * GRID + u; cy = j * GRID + lo - off; hx = L / 2; hy = W / 2; }
    else if (side == 1) { cx = i * GRID + u; cy = j * GRID + hi + off; hx = L / 2; hy = W / 2; }
    else if (side == 2) { cx = i * GRID + lo - off; cy = j * GRID + u; hx = W / 2; hy = L / 2; }
    else { cx = i * GRID + hi + off; cy = j * GRID + u; hx = W / 2; hy = L / 2; }
    s->boxes.push_back(Box{cx - hx, cy - hy, cx + hx, cy + hy, 0.f, H});
  }
  return s;
}

inline bool hit_box(const Box& b, const double o[3], const double d[3], double& tmin_out) {
  double t0 = 0.0, t1 = 1e30;
  const double lo[3] = {b.x0, b.y0, b.z0}, hi[3] = {b.x1, b.y1, b.z1};
  for (int k = 0; k < 3; k++) {
    if (std::fabs(d[k]) < 1e-12) {
      if (o[k] < lo[k] || o[k] > hi[k]) return false;
    } else {
      double a = (lo[k] - o[k]) / d[k], c = (hi[k] - o[k]) / d[k];
      if (a > c) std::swap(a, c);
      if (a > t0) t0 = a;
      if (c < t1) t1 = c;
      if (t0 > t1) return false;
    }
  }
  if (t0 <= 1e-6) return false;  // origin inside or behind
  tmin_out = t0;
  return true;
}
inline bool hit_cyl(const Cyl& c, const double o[3], const double d[3], double& t_out) {
  const double ox = o[0] - c.cx, oy = o[1] - c.cy;
  const double a = d[0] * d[0] + d[1] * d[1];
  if (a < 1e-12) return false;
  const double b = ox * d[0] + oy * d[1];
  const double cc = ox * ox + oy * oy - double(c.r) * c.r;
  const double disc = b * b - a * cc;
  if (disc < 0) return false;
  const double t = (-b - std::sqrt(disc)) / a;
  if (t <= 1e-6) return false;
  const double z = o[2] + t * d[2];
  if (z < c.z0 || z > c.z1) return false;
  t_out = t;
  return true;
}

inline bool hit_sph(const Sph& s, const double o[3], const double d[3], double& t_out) {
  const double ox = o[0] - s.cx, oy = o[1] - s.cy, oz = o[2] - s.cz;
  const double b = ox * d[0] + oy * d[1] + oz * d[2];
  const double c = ox * ox + oy * oy + oz * oz - double(s.r) * s.r;
  const double disc = b * b - c;
  if (disc < 0) return false;
  const double t = -b - std::sqrt(disc);
  if (t <= 1e-6) return false;
  t_out = t;
  return true;
}

}  // namespace

extern "C" {

void* synth_scene_create(uint64_t seed, float extent_m, int n_poles, int n_cars) {
  return make_scene(seed, extent_m, n_poles, n_cars);
}
void synth_scene_destroy(void* s) { delete static_cast<Scene*>(s); }
// flat copies of the primitives (for the CUDA ray caster, synth_gpu.cu): boxes 6 floats, cylinders 5, spheres 4 each
void synth_scene_export(void* s, int* n_boxes, float* boxes, int* n_cyls, float* cyls, int* n_sphs, float* sphs) {
  const Scene& S = *static_cast<Scene*>(s);
  *n_boxes = int(S.boxes.size());
  *n_cyls = int(S.cyls.size());
  *n_sphs = int(S.sphs.size());
  if (boxes) std::memcpy(boxes, S.boxes.data(), S.boxes.size() * sizeof(Box));
  if (cyls) std::memcpy(cyls, S.cyls.data(), S.cyls.size() * sizeof(Cyl));
  if (sphs) std::memcpy(sphs, S.sphs.data(), S.sphs.size() * sizeof(Sph));
}
void synth_scene_counts(void* s, int* n_boxes, int* n_cyls) {
  *n_boxes = int(static_cast<Scene*>(s)->boxes.size());
  *n_cyls = int(static_cast<Scene*>(s)->cyls.size());
}

// pose: 3x4 row-major sensor->world.  Returns the number of returns written (<= max_pts).
// out_t (optional): per-point relative time (az/2pi - 0.5) * 0.1 s.
static uint64_t scan_impl(void* scene, const double* pose, const double* twist, double sweep_s, int n_beams, int n_az,
                          double el_top_deg, double el_bot_deg, double max_range, double noise_sigma, uint64_t scan_seed,
                          float* out_xyzi, float* out_t, uint64_t max_pts);

uint64_t synth_scan(void* scene, const double* pose, int n_beams, int n_az, double el_top_deg, double el_bot_deg,
                    double max_range, double noise_sigma, uint64_t scan_seed, float* out_xyzi, float* out_t,
                    uint64_t max_pts) {
  return scan_impl(scene, pose, nullptr, 0.1, n_beams, n_az, el_top_deg, el_bot_deg, max_range, noise_sigma, scan_seed,
                   out_xyzi, out_t, max_pts);
}

// A spinning sensor moving with body-frame twist (vx vy vz wx wy wz) during the sweep: the ray of azimuth step a is
// cast at time t = (az/2pi) * sweep_s (t = 0 at the middle of the sweep) from pose(t) = pose * exp(twist * t), and the
// return is reported in the sensor frame of THAT instant, which is what FilterDeskew undoes.
uint64_t synth_scan_skewed(void* scene, const double* pose, const double* twist, double sweep_s, int n_beams, int n_az,
                           double el_top_deg, double el_bot_deg, double max_range, double noise_sigma, uint64_t scan_seed,
                           float* out_xyzi, float* out_t, uint64_t max_pts) {
  return scan_impl(scene, pose, twist, sweep_s, n_beams, n_az, el_top_deg, el_bot_deg, max_range, noise_sigma, scan_seed,
                   out_xyzi, out_t, max_pts);
}

static void se3_exp_small(const double* xi, double t, double* P) {  // P = exp(xi * t) as 3x4
  const double v[3] = {xi[0] * t, xi[1] * t, xi[2] * t}, w[3] = {xi[3] * t, xi[4] * t, xi[5] * t};
  const double th2 = w[0] * w[0] + w[1] * w[1] + w[2] * w[2], th = std::sqrt(th2);
  double A, B, C;
  if (th < 1e-6) { A = 1 - th2 / 6; B = 0.5 - th2 / 24; C = 1.0 / 6 - th2 / 120; }
  else { A = std::sin(th) / th; B = (1 - std::cos(th)) / th2; C = (th - std::sin(th)) / (th2 * th); }
  const double W[9] = {0, -w[2], w[1], w[2], 0, -w[0], -w[1], w[0], 0};
  double W2[9];
  for (int i = 0; i < 3; i++)
    for (int j = 0; j < 3; j++) W2[3 * i + j] = W[3 * i] * W[j] + W[3 * i + 1] * W[3 + j] + W[3 * i + 2] * W[6 + j];
  for (int i = 0; i < 3; i++) {
    double tr = v[i];
    for (int j = 0; j < 3; j++) {
      P[4 * i + j] = (i == j) + A * W[3 * i + j] + B * W2[3 * i + j];
      tr += (B * W[3 * i + j] + C * W2[3 * i + j]) * v[j];
    }
    P[4 * i + 3] = tr;
  }
}

static uint64_t scan_impl(void* scene, const double* pose_mid, const double* twist, double sweep_s, int n_beams, int n_az,
                          double el_top_deg, double el_bot_deg, double max_range, double noise_sigma, uint64_t scan_seed,
                          float* out_xyzi, float* out_t, uint64_t max_pts) {
  const Scene& S = *static_cast<Scene*>(scene);
  const double o_mid[3] = {pose_mid[3], pose_mid[7], pose_mid[11]};
  const double* pose = pose_mid;
  const double* o = o_mid;
  const double bin_margin = twist ? 1.5 : 0.0;  // the origin moves during the sweep: widen the bearing bins
  // cull primitives by distance and bucket by bearing
  constexpr int NB = 360;
  std::vector<std::vector<int>> bb(NB), cb(NB), sb(NB);
  auto add = [&](std::vector<std::vector<int>>& bins, int id, double cx, double cy, double rad) {
    (void)0;
    rad += bin_margin;
    const double dx = cx - o[0], dy = cy - o[1];
    const double dist = std::sqrt(dx * dx + dy * dy);
    if (dist - rad > max_range) return;
    if (dist <= rad * 1.05) {
      for (int k = 0; k < NB; k++) bins[k].push_back(id);
      return;
    }
    const double c = std::atan2(dy, dx), h = std::asin(std::min(1.0, rad / dist)) + 2.0 * M_PI / NB;
    const int k0 = int(std::floor((c - h + M_PI) / (2 * M_PI) * NB)), k1 = int(std::floor((c + h + M_PI) / (2 * M_PI) * NB));
    for (int k = k0; k <= k1; k++) bins[((k % NB) + NB) % NB].push_back(id);
  };
  for (size_t i = 0; i < S.boxes.size(); i++) {
    const Box& b = S.boxes[i];
    add(bb, int(i), 0.5 * (b.x0 + b.x1), 0.5 * (b.y0 + b.y1), 0.5 * std::hypot(b.x1 - b.x0, b.y1 - b.y0));
  }
  for (size_t i = 0; i < S.cyls.size(); i++) add(cb, int(i), S.cyls[i].cx, S.cyls[i].cy, S.cyls[i].r);
  for (size_t i = 0; i < S.sphs.size(); i++) add(sb, int(i), S.sphs[i].cx, S.sphs[i].cy, S.sphs[i].r);

  uint64_t n = 0;
  double pose_t[12], o_t[3];
  for (int a = 0; a < n_az; a++) {
    const double az = -M_PI + 2.0 * M_PI * (double(a) + 0.5) / n_az;
    if (twist) {  // pose at the instant this azimuth column is fired
      double E[12];
      se3_exp_small(twist, (az / (2.0 * M_PI)) * sweep_s, E);
      for (int r = 0; r < 3; r++) {
        for (int c = 0; c < 3; c++)
          pose_t[4 * r + c] = pose_mid[4 * r] * E[c] + pose_mid[4 * r + 1] * E[4 + c] + pose_mid[4 * r + 2] * E[8 + c];
        pose_t[4 * r + 3] = pose_mid[4 * r] * E[3] + pose_mid[4 * r + 1] * E[7] + pose_mid[4 * r + 2] * E[11] + pose_mid[4 * r + 3];
      }
      o_t[0] = pose_t[3]; o_t[1] = pose_t[7]; o_t[2] = pose_t[11];
      pose = pose_t;
      o = o_t;
    }
    for (int b = 0; b < n_beams; b++) {
      const double el = (el_top_deg + (el_bot_deg - el_top_deg) * (n_beams > 1 ? double(b) / (n_beams - 1) : 0.0)) * M_PI / 180.0;
      const double ds[3] = {std::cos(el) * std::cos(az), std::cos(el) * std::sin(az), std::sin(el)};
      double d[3];
      for (int k = 0; k < 3; k++) d[k] = pose[4 * k] * ds[0] + pose[4 * k + 1] * ds[1] + pose[4 * k + 2] * ds[2];
      double best = 1e30;
      if (d[2] < -1e-9) {
        const double t = -o[2] / d[2];
        if (t > 0) best = t;
      }
      const double phi = std::atan2(d[1], d[0]);
      int bin = int(std::floor((phi + M_PI) / (2 * M_PI) * NB));
      bin = std::min(NB - 1, std::max(0, bin));
      double t;
      for (int id : bb[bin])
        if (hit_box(S.boxes[id], o, d, t) && t < best) best = t;
      for (int id : cb[bin])
        if (hit_cyl(S.cyls[id], o, d, t) && t < best) best = t;
      for (int id : sb[bin])
        if (hit_sph(S.sphs[id], o, d, t) && t < best) best = t;
      if (best > max_range) continue;
      const uint64_t ray = uint64_t(a) * uint64_t(n_beams) + uint64_t(b);
      const uint64_t h1 = splitmix(scan_seed * 0x100000001B3ull + ray * 2 + 1), h2 = splitmix(h1 + 0x632BE59BD9B4E019ull);
      const double g = std::sqrt(-2.0 * std::log(u01(h1))) * std::cos(2.0 * M_PI * u01(h2));
      const double r = best + noise_sigma * g;
      if (r < 0.5 || n >= max_pts) continue;
      out_xyzi[4 * n] = float(r * ds[0]);
      out_xyzi[4 * n + 1] = float(r * ds[1]);
      out_xyzi[4 * n + 2] = float(r * ds[2]);
      out_xyzi[4 * n + 3] = float(u01(splitmix(h2 + 7)));
      if (out_t) out_t[n] = float((az / (2.0 * M_PI)) * sweep_s);
      n++;
    }
  }
  return n;
}

}  // extern "C"
