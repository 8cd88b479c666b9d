// mlo-lidar-odometry-cli — offline driver over the B200 host layer, with the flag names of the reference's
// apps/mola-lidar-odometry-cli.cpp:84-161 for the subset that concerns the hot path:
//   -c/--config <pipeline.yaml>            (required)  pipeline file (pipelines/lidar3d-default.yaml surface)
//   --input-kitti-seq <00|01|...>          KITTI odometry sequence; the velodyne directory is
//                                          $KITTI_BASE_DIR/sequences/<seq>/velodyne (as mola_input_kitti_dataset does)
//   --input-bin-dir <dir>                  (extension) any directory of KITTI-layout *.bin clouds (x,y,z,i float32)
//   --kitti-correction-angle-deg <deg>     vertical angle correction of Deschaud 2018 (default 0.205, KITTI only)
//   --output-tum-path <file>               estimated trajectory in TUM format (apps/...cli.cpp:524-531)
//   --only-first-n <N> / --skip-first-n <N>
//   --lidar-hz <Hz>                        (extension) scan rate used to stamp .bin files (default 10)
// Several inputs separated by commas (--input-kitti-seq 00,02,05 or --input-bin-dir a,b) run as a FLEET: independent
// LidarOdometry instances advancing in lock step on one GPU (mlo_fleet_*), the one-GPU counterpart of the reference's
// `parallel -j` over sequences (eval/cli_kitti.sh:23); the trajectory of input <name> goes to <output>_<name>.<ext>.
// Loop shape = cli.cpp:469-522: read observation i, onNewObservation, wait until processed.  rawlog / rosbag2 /
// MulRan / KITTI-360 / Paris-Luco readers, simplemap output and plugin loading are out of scope.
#include <algorithm>
#include <chrono>
#include <cmath>
#include <cstdio>
#include <cstdlib>
#include <cstring>
#include <filesystem>
#include <fstream>
#include <string>
#include <vector>

#include "mlo_b200_host.h"

namespace fs = std::filesystem;

static void usage() {
  std::fprintf(stderr,
               "USAGE: mlo-lidar-odometry-cli -c <pipeline.yaml> (--input-kitti-seq <NN>[,<NN>...] | --input-bin-dir <dir>[,<dir>...])\n"
               "         [--output-tum-path <file>] [--only-first-n N] [--skip-first-n N] [--kitti-correction-angle-deg D]\n"
               "         [--lidar-hz HZ] [--cuda-device ID]\n");
}

// quaternion (x y z w) of the rotation block of a 3x4 pose
static void quat_of(const double* T, double q[4]) {
  const double r00 = T[0], r01 = T[1], r02 = T[2], r10 = T[4], r11 = T[5], r12 = T[6], r20 = T[8], r21 = T[9], r22 = T[10];
  const double tr = r00 + r11 + r22;
  double w, x, y, z;
  if (tr > 0) { const double s = 2 * std::sqrt(tr + 1); w = 0.25 * s; x = (r21 - r12) / s; y = (r02 - r20) / s; z = (r10 - r01) / s; }
  else if (r00 > r11 && r00 > r22) { const double s = 2 * std::sqrt(1 + r00 - r11 - r22); w = (r21 - r12) / s; x = 0.25 * s; y = (r01 + r10) / s; z = (r02 + r20) / s; }
  else if (r11 > r22) { const double s = 2 * std::sqrt(1 + r11 - r00 - r22); w = (r02 - r20) / s; x = (r01 + r10) / s; y = 0.25 * s; z = (r12 + r21) / s; }
  else { const double s = 2 * std::sqrt(1 + r22 - r00 - r11); w = (r10 - r01) / s; x = (r02 + r20) / s; y = (r12 + r21) / s; z = 0.25 * s; }
  q[0] = x; q[1] = y; q[2] = z; q[3] = w;
}

int main(int argc, char** argv) {
  std::string yaml, kitti_seq, bin_dir, out_tum;
  long first_n = 0, skip_n = 0;
  double angle_deg = 0.205, hz = 10.0;
  int device = 0;
  bool angle_given = false;
  for (int i = 1; i < argc; i++) {
    const std::string a = argv[i];
    auto val = [&](const char* name) -> std::string {
      if (i + 1 >= argc) { std::fprintf(stderr, "missing value for %s\n", name); usage(); std::exit(2); }
      return argv[++i];
    };
    if (a == "-c" || a == "--config") yaml = val("--config");
    else if (a == "--input-kitti-seq") kitti_seq = val("--input-kitti-seq");
    else if (a == "--input-bin-dir") bin_dir = val("--input-bin-dir");
    else if (a == "--output-tum-path") out_tum = val("--output-tum-path");
    else if (a == "--only-first-n") first_n = std::atol(val("--only-first-n").c_str());
    else if (a == "--skip-first-n") skip_n = std::atol(val("--skip-first-n").c_str());
    else if (a == "--kitti-correction-angle-deg") { angle_deg = std::atof(val("--kitti-correction-angle-deg").c_str()); angle_given = true; }
    else if (a == "--lidar-hz") hz = std::atof(val("--lidar-hz").c_str());
    else if (a == "--cuda-device") device = std::atoi(val("--cuda-device").c_str());
    else if (a == "-h" || a == "--help") { usage(); return 0; }
    else { std::fprintf(stderr, "unknown argument '%s'\n", a.c_str()); usage(); return 2; }
  }
  if (yaml.empty() || (kitti_seq.empty() && bin_dir.empty())) { usage(); return 2; }
  auto split = [](const std::string& v) {
    std::vector<std::string> out;
    size_t b = 0;
    while (b <= v.size()) {
      const size_t e = v.find(',', b);
      const std::string tok = v.substr(b, e == std::string::npos ? std::string::npos : e - b);
      if (!tok.empty()) out.push_back(tok);
      if (e == std::string::npos) break;
      b = e + 1;
    }
    return out;
  };
  bool is_kitti = false;
  std::vector<std::string> dirs, names;
  if (bin_dir.empty()) {
    const char* base = std::getenv("KITTI_BASE_DIR");
    if (!base) { std::fprintf(stderr, "KITTI_BASE_DIR is not set (needed by --input-kitti-seq)\n"); return 2; }
    for (const std::string& sq : split(kitti_seq)) {
      dirs.push_back(std::string(base) + "/sequences/" + sq + "/velodyne");
      names.push_back(sq);
    }
    is_kitti = true;
  } else {
    dirs = split(bin_dir);
    for (size_t i = 0; i < dirs.size(); i++) names.push_back(fs::path(dirs[i]).filename().string().empty() ? std::to_string(i) : fs::path(dirs[i]).filename().string());
  }
  const uint32_t S = uint32_t(dirs.size());
  std::vector<std::vector<fs::path>> files(S);
  size_t longest = 0;
  for (uint32_t q = 0; q < S; q++) {
    std::error_code ec;
    for (auto& e : fs::directory_iterator(dirs[q], ec))
      if (e.path().extension() == ".bin") files[q].push_back(e.path());
    if (ec || files[q].empty()) { std::fprintf(stderr, "no *.bin clouds under '%s'\n", dirs[q].c_str()); return 2; }
    std::sort(files[q].begin(), files[q].end());
    longest = std::max(longest, files[q].size());
  }

  mlo_ctx* ctx = nullptr;
  if (mlo_create(device, &ctx) != MLO_OK) { std::fprintf(stderr, "mlo_create failed: no sm_100 device (there is no CPU fallback)\n"); return 3; }
  mlo_fleet* fleet = nullptr;
  if (mlo_fleet_create(ctx, yaml.c_str(), 0, S, &fleet) != MLO_OK) {
    std::fprintf(stderr, "cannot initialise from '%s': %s\n", yaml.c_str(), mlo_fleet_last_error(nullptr));
    mlo_destroy(ctx);
    return 2;
  }
  const double corr = (is_kitti || angle_given) ? angle_deg * M_PI / 180.0 : 0.0;
  std::vector<std::vector<float>> cloud(S);
  std::vector<const float*> pts(S);
  std::vector<uint64_t> npts(S);
  std::vector<double> stamps(S);
  std::vector<mlo_lo_scan_output> outs(S);
  size_t n_done = 0, n_steps = 0;
  const auto t0 = std::chrono::steady_clock::now();
  for (size_t i = size_t(std::max(0L, skip_n)); i < longest; i++) {
    if (first_n > 0 && long(n_steps) >= first_n) break;
    for (uint32_t q = 0; q < S; q++) {
      pts[q] = nullptr;
      npts[q] = 0;
      stamps[q] = double(i) / hz;
      if (i >= files[q].size()) continue;  // this sequence has ended: its slot stays idle
      std::ifstream f(files[q][i], std::ios::binary | std::ios::ate);
      const std::streamsize bytes = f.tellg();
      f.seekg(0);
      cloud[q].resize(size_t(bytes) / sizeof(float));
      f.read(reinterpret_cast<char*>(cloud[q].data()), bytes);
      const uint64_t n = cloud[q].size() / 4;
      if (corr != 0.0) {  // Deschaud 2018: rotate every point by `corr` about the axis (p x z)
        for (uint64_t k = 0; k < n; k++) {
          float* p = &cloud[q][4 * k];
          const double ax = p[1], ay = -p[0];  // p x (0,0,1)
          const double an = std::sqrt(ax * ax + ay * ay);
          if (an < 1e-9) continue;
          const double ux = ax / an, uy = ay / an, c = std::cos(corr), s = std::sin(corr);
          const double x = p[0], y = p[1], z = p[2], d = ux * x + uy * y;
          p[0] = float(x * c + (uy * z) * s + ux * d * (1 - c));
          p[1] = float(y * c + (-ux * z) * s + uy * d * (1 - c));
          p[2] = float(z * c + (ux * y - uy * x) * s);
        }
      }
      pts[q] = cloud[q].data();
      npts[q] = n;
      n_done++;
    }
    if (mlo_fleet_on_lidar(fleet, pts.data(), 4, npts.data(), stamps.data(), nullptr, outs.data()) != MLO_OK) {
      std::fprintf(stderr, "fatal error at scan %zu: %s\n", i, mlo_fleet_last_error(fleet));  // LidarOdometry.cpp:614-619
      mlo_fleet_destroy(fleet);
      mlo_destroy(ctx);
      return 1;
    }
    n_steps++;
    if (n_steps % 100 == 0) std::fprintf(stderr, "[cli] %zu steps, quality %.2f, sigma %.2f\n", n_steps, outs[0].quality, outs[0].sigma);
  }
  const double secs = std::chrono::duration<double>(std::chrono::steady_clock::now() - t0).count();
  std::fprintf(stderr, "[cli] %zu scans of %u sequence(s) in %.2f s (%.1f scans/s)\n", n_done, S, secs, n_done / std::max(secs, 1e-9));
  if (!out_tum.empty()) {
    for (uint32_t q = 0; q < S; q++) {
      std::string path = out_tum;
      if (S > 1) {  // <stem>_<name><ext>
        const fs::path p(out_tum);
        path = (p.parent_path() / (p.stem().string() + "_" + names[q] + p.extension().string())).string();
      }
      uint64_t n = 0;
      mlo_fleet_trajectory(fleet, q, nullptr, nullptr, 0, &n);
      std::vector<double> st(n), ps(12 * n);
      mlo_fleet_trajectory(fleet, q, st.data(), ps.data(), n, &n);
      std::FILE* fo = std::fopen(path.c_str(), "w");
      if (!fo) { std::fprintf(stderr, "cannot write '%s'\n", path.c_str()); return 1; }
      for (uint64_t k = 0; k < n; k++) {
        double qt[4];
        quat_of(&ps[12 * k], qt);
        std::fprintf(fo, "%.6f %.6f %.6f %.6f %.6f %.6f %.6f %.6f\n", st[k], ps[12 * k + 3], ps[12 * k + 7], ps[12 * k + 11], qt[0], qt[1], qt[2], qt[3]);
      }
      std::fclose(fo);
    }
  }
  mlo_fleet_destroy(fleet);
  mlo_destroy(ctx);
  return 0;
}
