// navstate_fuse.hpp — motion model of the LidarOdometry caller contract.
//
// Reference: mola::NavStateFuse (package mola_navstate_fuse, NOT in /root/reference) as used at
//   module/src/LidarOdometry.cpp:338        initialize(cfg["navstate_fuse_params"])
//   module/src/LidarOdometry.cpp:808-815    estimated_navstate(stamp)  -> pose (mean + cov_inv) + twist
//   module/src/LidarOdometry.cpp:854-877    pose.mean = ICP initial guess; pose.cov_inv != 0 => prior of Solver_GaussNewton
//   module/src/LidarOdometry.cpp:1035-1039  fuse_pose(stamp, ICP result with covariance) / reset() on a rejected ICP
// with the parameter block of pipelines/lidar3d-default.yaml:126-144.
//
// parity unpinned: upstream solves a sliding-window factor graph (constant-velocity factors with random-walk
// acceleration noise, integrator factors, one pose factor per fused observation, a twist prior) with an external
// non-linear least-squares library.  This header restates that MODEL in closed form, in the tangent space of the newest
// fused pose (where it is linear-Gaussian), honouring every parameter of the YAML block:
//   window      fused poses younger than sliding_window_length (relative to the newest) take part
//   twist       Kalman filter over the pose increments of the window, per tangent component: process noise
//               (sigma_random_walk_acceleration dt)^2 per step, measurement = increment / dt with the fused poses'
//               variances; a single fused pose falls back on initial_twist (sigma initial_twist_sigma_lin / _ang)
//   prediction  T(t) = T_n exp(w dt), valid for dt <= max_time_to_use_velocity_model
//   covariance  Sigma_n + dt^2 Var[w] + sigma_a^2 dt^3 / 3 + sigma_integrator^2 (YAML units: m, rad), inverted into
//               cov_inv: with the shipped values (1 m, 1 rad) the prior is weak next to a few hundred pairings, i.e. it
//               regularises an under-constrained ICP and otherwise leaves the solution to the data
// Tangent order (x y z rx ry rz), right-multiplicative, the convention of the GN prior term (csrc/icp.cuh prior_add).
#pragma once
#include <array>
#include <cmath>
#include <cstring>
#include <functional>
#include <optional>
#include <vector>

namespace mlo_host {

struct NavStateFuseParams {  // navstate_fuse_params (default.yaml:126-144)
  double max_time_to_use_velocity_model = 0.75;
  double sliding_window_length = 0.50;
  double sigma_random_walk_acceleration_linear = 1.0;
  double sigma_random_walk_acceleration_angular = 10.0;
  double sigma_integrator_position = 1.0;
  double sigma_integrator_orientation = 1.0;
  std::array<double, 6> initial_twist{};
  double initial_twist_sigma_lin = 20.0, initial_twist_sigma_ang = 3.0;
};

class NavStateFuse {
 public:
  using Pose = std::array<double, 12>;
  using Mat66 = std::array<double, 36>;
  struct NavState {
    Pose pose;
    Mat66 cov_inv{};  // information of `pose` (all zero = no information)
    std::array<double, 6> twist{};
  };
  using ExpFn = std::function<void(const double*, double*)>;
  using LogFn = std::function<void(const double*, double*)>;

  NavStateFuseParams params;
  void set_lie(ExpFn e, LogFn l) {
    exp_ = std::move(e);
    log_ = std::move(l);
  }
  void reset() { obs_.clear(); }
  size_t size() const { return obs_.size(); }

  // fuse_pose(stamp, CPose3DPDFGaussian): `cov` = 6x6 covariance in tangent order (row-major); null = 1e-12 I (:832)
  void fuse_pose(double stamp, const Pose& p, const double* cov) {
    Obs o;
    o.t = stamp;
    o.T = p;
    for (int k = 0; k < 6; k++) o.var[k] = cov ? std::max(cov[6 * k + k], 0.0) : 1e-12;
    if (cov) std::memcpy(o.cov.data(), cov, sizeof(double) * 36);
    else {
      o.cov.fill(0.0);
      for (int k = 0; k < 6; k++) o.cov[6 * k + k] = 1e-12;
    }
    obs_.push_back(o);
    while (obs_.size() > 1 && obs_.front().t < stamp - params.sliding_window_length) obs_.erase(obs_.begin());
  }

  std::optional<NavState> estimated_navstate(double stamp) const {
    if (obs_.empty()) return std::nullopt;
    const Obs& n = obs_.back();
    const double dt = stamp - n.t;
    if (dt > params.max_time_to_use_velocity_model) return std::nullopt;
    const double sa[6] = {params.sigma_random_walk_acceleration_linear, params.sigma_random_walk_acceleration_linear,
                          params.sigma_random_walk_acceleration_linear, params.sigma_random_walk_acceleration_angular,
                          params.sigma_random_walk_acceleration_angular, params.sigma_random_walk_acceleration_angular};
    const double si[6] = {params.sigma_integrator_position, params.sigma_integrator_position, params.sigma_integrator_position,
                          params.sigma_integrator_orientation, params.sigma_integrator_orientation,
                          params.sigma_integrator_orientation};
    NavState ns;
    double var_w[6];
    if (obs_.size() >= 2) {
      // Kalman filter over the increments of the window, per tangent component: state = twist, process noise
      // (sigma_a dt)^2 per step (random-walk acceleration), measurement z_k = log(T_{k-1}^-1 T_k) / dt_k with variance
      // (var_k + var_{k-1}) / dt_k^2.  With millimetre pose observations and sigma_a ~ 1 m/s^2 the estimate follows the
      // newest increment closely (gain ~0.85), as the smoother of the factor graph does.
      double v[6] = {0, 0, 0, 0, 0, 0}, Pv[6];
      bool have = false;
      for (size_t i = 1; i < obs_.size(); i++) {
        const Obs& o0 = obs_[i - 1];
        const Obs& o1 = obs_[i];
        const double dtk = o1.t - o0.t;
        if (!(dtk > 0)) continue;
        const Pose rel = minus(o1.T, o0.T);  // T_{k-1}^-1 T_k
        double xi[6];
        log_(rel.data(), xi);
        for (int k = 0; k < 6; k++) {
          const double z = xi[k] / dtk, R = (o0.var[k] + o1.var[k]) / (dtk * dtk);
          if (!have) {
            v[k] = z;
            Pv[k] = R;
          } else {
            const double Pp = Pv[k] + sa[k] * sa[k] * dtk * dtk;
            const double K = Pp / (Pp + R);
            v[k] += K * (z - v[k]);
            Pv[k] = (1.0 - K) * Pp;
          }
        }
        have = true;
      }
      if (!have) return std::nullopt;
      for (int k = 0; k < 6; k++) {
        ns.twist[k] = v[k];
        var_w[k] = Pv[k];
      }
    } else {
      // only the initial pose is known: the configured initial twist, if any (its sigma is the YAML's)
      bool any = false;
      for (double v : params.initial_twist) any = any || v != 0.0;
      if (!any) return std::nullopt;
      ns.twist = params.initial_twist;
      for (int k = 0; k < 6; k++) {
        const double s = k < 3 ? params.initial_twist_sigma_lin : params.initial_twist_sigma_ang;
        var_w[k] = s * s;
      }
    }
    double step[6];
    for (int k = 0; k < 6; k++) step[k] = ns.twist[k] * dt;
    Pose d;
    exp_(step, d.data());
    ns.pose = compose(n.T, d);
    // covariance of the prediction (tangent space of the newest pose), then its inverse
    Mat66 S = n.cov;
    const double adt = std::fabs(dt);
    for (int k = 0; k < 6; k++)
      S[6 * k + k] += dt * dt * var_w[k] + sa[k] * sa[k] * adt * adt * adt / 3.0 + si[k] * si[k];  // (sigma_integrator_*: [m], [rad])
    if (!spd_inverse(S, ns.cov_inv)) ns.cov_inv.fill(0.0);
    return ns;
  }

  static Pose compose(const Pose& A, const Pose& B) {
    Pose C;
    for (int r = 0; r < 3; r++) {
      for (int c = 0; c < 3; c++) C[4 * r + c] = A[4 * r] * B[c] + A[4 * r + 1] * B[4 + c] + A[4 * r + 2] * B[8 + c];
      C[4 * r + 3] = A[4 * r] * B[3] + A[4 * r + 1] * B[7] + A[4 * r + 2] * B[11] + A[4 * r + 3];
    }
    return C;
  }
  static Pose minus(const Pose& A, const Pose& B) {  // B^-1 A
    Pose Bi;
    for (int r = 0; r < 3; r++)
      for (int c = 0; c < 3; c++) Bi[4 * r + c] = B[4 * c + r];
    for (int r = 0; r < 3; r++) Bi[4 * r + 3] = -(Bi[4 * r] * B[3] + Bi[4 * r + 1] * B[7] + Bi[4 * r + 2] * B[11]);
    return compose(Bi, A);
  }
  // inverse of a symmetric positive-definite 6x6 (Cholesky); false when not SPD
  static bool spd_inverse(const Mat66& A, Mat66& inv) {
    double L[36] = {0};
    for (int j = 0; j < 6; j++) {
      double d = A[6 * j + j];
      for (int k = 0; k < j; k++) d -= L[6 * j + k] * L[6 * j + k];
      if (!(d > 0.0) || !std::isfinite(d)) return false;
      L[6 * j + j] = std::sqrt(d);
      for (int i = j + 1; i < 6; i++) {
        double s = A[6 * i + j];
        for (int k = 0; k < j; k++) s -= L[6 * i + k] * L[6 * j + k];
        L[6 * i + j] = s / L[6 * j + j];
      }
    }
    for (int c = 0; c < 6; c++) {
      double y[6], x[6];
      for (int i = 0; i < 6; i++) {
        double s = (i == c) ? 1.0 : 0.0;
        for (int k = 0; k < i; k++) s -= L[6 * i + k] * y[k];
        y[i] = s / L[6 * i + i];
      }
      for (int i = 5; i >= 0; i--) {
        double s = y[i];
        for (int k = i + 1; k < 6; k++) s -= L[6 * k + i] * x[k];
        x[i] = s / L[6 * i + i];
      }
      for (int r = 0; r < 6; r++) inv[6 * r + c] = x[r];
    }
    return true;
  }

 private:
  struct Obs {
    double t;
    Pose T;
    Mat66 cov;
    double var[6];
  };
  std::vector<Obs> obs_;
  ExpFn exp_;
  LogFn log_;
};

}  // namespace mlo_host
