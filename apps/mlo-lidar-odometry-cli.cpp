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
               "USAGE: mlo-lidar-odometry-cli -c <pipeline.yaml> (--input-kitti-seq <NN> | --input-bin-dir <dir>)\n"
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
  bool is_kitti = false;
  if (bin_dir.empty()) {
    const char* base = std::getenv("KITTI_BASE_DIR");
    if (!base) { std::fprintf(stderr, "KITTI_BASE_DIR is not set (needed by --input-kitti-seq)\n"); return 2; }
    bin_dir = std::string(base) + "/sequences/" + kitti_seq + "/velodyne";
    is_kitti = true;
  }
  std::vector<fs::path> files;
  std::error_code ec;
  for (auto& e : fs::directory_iterator(bin_dir, ec))
    if (e.path().extension() == ".bin") files.push_back(e.path());
  if (ec || files.empty()) { std::fprintf(stderr, "no *.bin clouds under '%s'\n", bin_dir.c_str()); return 2; }
  std::sort(files.begin(), files.end());

  mlo_ctx* ctx = nullptr;
  if (mlo_create(device, &ctx) != MLO_OK) { std::fprintf(stderr, "mlo_create failed: no sm_100 device (there is no CPU fallback)\n"); return 3; }
  mlo_lo* lo = nullptr;
  if (mlo_lo_create(ctx, yaml.c_str(), 0, &lo) != MLO_OK) {
    std::fprintf(stderr, "cannot initialise from '%s': %s\n", yaml.c_str(), mlo_lo_last_error(nullptr));
    mlo_destroy(ctx);
    return 2;
  }
  const double corr = (is_kitti || angle_given) ? angle_deg * M_PI / 180.0 : 0.0;
  std::vector<float> cloud;
  size_t n_done = 0;
  const auto t0 = std::chrono::steady_clock::now();
  for (size_t i = size_t(std::max(0L, skip_n)); i < files.size(); i++) {
    if (first_n > 0 && long(n_done) >= first_n) break;
    std::ifstream f(files[i], std::ios::binary | std::ios::ate);
    const std::streamsize bytes = f.tellg();
    f.seekg(0);
    cloud.resize(size_t(bytes) / sizeof(float));
    f.read(reinterpret_cast<char*>(cloud.data()), bytes);
    const uint64_t n = cloud.size() / 4;
    if (corr != 0.0) {  // Deschaud 2018: rotate every point by `corr` about the axis (p x z)
      for (uint64_t k = 0; k < n; k++) {
        float* p = &cloud[4 * k];
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
    mlo_lo_scan_output out;
    if (mlo_lo_on_lidar(lo, cloud.data(), 4, n, double(i) / hz, &out) != MLO_OK) {
      std::fprintf(stderr, "fatal error at scan %zu: %s\n", i, mlo_lo_last_error(lo));  // LidarOdometry.cpp:614-619
      mlo_lo_destroy(lo);
      mlo_destroy(ctx);
      return 1;
    }
    n_done++;
    if (n_done % 100 == 0) std::fprintf(stderr, "[cli] %zu scans, quality %.2f, sigma %.2f\n", n_done, out.quality, out.sigma);
  }
  const double secs = std::chrono::duration<double>(std::chrono::steady_clock::now() - t0).count();
  std::fprintf(stderr, "[cli] %zu scans in %.2f s (%.1f scans/s)\n", n_done, secs, n_done / std::max(secs, 1e-9));
  if (!out_tum.empty()) {
    uint64_t n = 0;
    mlo_lo_trajectory(lo, nullptr, nullptr, 0, &n);
    std::vector<double> st(n), ps(12 * n);
    mlo_lo_trajectory(lo, st.data(), ps.data(), n, &n);
    std::FILE* fo = std::fopen(out_tum.c_str(), "w");
    if (!fo) { std::fprintf(stderr, "cannot write '%s'\n", out_tum.c_str()); return 1; }
    for (uint64_t k = 0; k < n; k++) {
      double q[4];
      quat_of(&ps[12 * k], q);
      std::fprintf(fo, "%.6f %.6f %.6f %.6f %.6f %.6f %.6f %.6f\n", st[k], ps[12 * k + 3], ps[12 * k + 7], ps[12 * k + 11], q[0], q[1], q[2], q[3]);
    }
    std::fclose(fo);
  }
  mlo_lo_destroy(lo);
  mlo_destroy(ctx);
  return 0;
}
