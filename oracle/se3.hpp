// oracle/se3.hpp — TEST INFRASTRUCTURE (CPU oracle), not product code.
//
// Double-precision SE(3) helpers restating the MRPT calls the reference makes around the hot
// path: CPose3D compose / inverse-compose ("a - b", LidarOdometry.cpp:930-931,973,1056),
// Lie::SO<3>::log (LidarOdometry.cpp:936,981,1080), Lie::SE<3>::exp/log (used by
// mp2p_icp::ICP::align's stall test and Solver_GaussNewton's retraction).
// Tangent order is (v, w) = translation first, rotation second, right-multiplicative update.
// parity unpinned: MRPT is absent from /root/reference.
#pragma once
#include <cmath>
#include <cstring>

namespace orc {

struct Pose {
  double R[3][3];
  double t[3];
  static Pose identity() {
    Pose p;
    std::memset(&p, 0, sizeof(p));
    p.R[0][0] = p.R[1][1] = p.R[2][2] = 1.0;
    return p;
  }
  static Pose from3x4(const double* m) {
    Pose p;
    for (int r = 0; r < 3; r++) {
      for (int c = 0; c < 3; c++) p.R[r][c] = m[r * 4 + c];
      p.t[r] = m[r * 4 + 3];
    }
    return p;
  }
  void to3x4(double* m) const {
    for (int r = 0; r < 3; r++) {
      for (int c = 0; c < 3; c++) m[r * 4 + c] = R[r][c];
      m[r * 4 + 3] = t[r];
    }
  }
};

// a (+) b : CPose3D::composeFrom
inline Pose compose(const Pose& a, const Pose& b) {
  Pose o;
  for (int r = 0; r < 3; r++) {
    for (int c = 0; c < 3; c++) o.R[r][c] = a.R[r][0] * b.R[0][c] + a.R[r][1] * b.R[1][c] + a.R[r][2] * b.R[2][c];
    o.t[r] = a.R[r][0] * b.t[0] + a.R[r][1] * b.t[1] + a.R[r][2] * b.t[2] + a.t[r];
  }
  return o;
}

inline Pose inverse(const Pose& a) {
  Pose o;
  for (int r = 0; r < 3; r++)
    for (int c = 0; c < 3; c++) o.R[r][c] = a.R[c][r];
  for (int r = 0; r < 3; r++) o.t[r] = -(o.R[r][0] * a.t[0] + o.R[r][1] * a.t[1] + o.R[r][2] * a.t[2]);
  return o;
}

// "a - b" in MRPT = b^{-1} (+) a  (pose of a as seen from b)
inline Pose minus(const Pose& a, const Pose& b) { return compose(inverse(b), a); }

inline void so3_exp(const double w[3], double R[3][3]) {
  const double th2 = w[0] * w[0] + w[1] * w[1] + w[2] * w[2];
  const double th = std::sqrt(th2);
  double A, B;  // A = sin(th)/th, B = (1-cos(th))/th^2
  if (th < 0.05) {  // series (see se3_exp): (1 - cos th) / th^2 cancels catastrophically for small th
    A = 1.0 - th2 / 6.0 + th2 * th2 / 120.0 - th2 * th2 * th2 / 5040.0 + th2 * th2 * th2 * th2 / 362880.0;
    B = 0.5 - th2 / 24.0 + th2 * th2 / 720.0 - th2 * th2 * th2 / 40320.0 + th2 * th2 * th2 * th2 / 3628800.0;
  } else {
    A = std::sin(th) / th;
    B = (1.0 - std::cos(th)) / th2;
  }
  const double wx = w[0], wy = w[1], wz = w[2];
  R[0][0] = 1.0 - B * (wy * wy + wz * wz);
  R[0][1] = -A * wz + B * wx * wy;
  R[0][2] = A * wy + B * wx * wz;
  R[1][0] = A * wz + B * wx * wy;
  R[1][1] = 1.0 - B * (wx * wx + wz * wz);
  R[1][2] = -A * wx + B * wy * wz;
  R[2][0] = -A * wy + B * wx * wz;
  R[2][1] = A * wx + B * wy * wz;
  R[2][2] = 1.0 - B * (wx * wx + wy * wy);
}

// Rotation log through the unit quaternion: stable for all angles in [0, pi].
inline void so3_log(const double R[3][3], double w[3]) {
  double q[4];  // (w, x, y, z)
  const double tr = R[0][0] + R[1][1] + R[2][2];
  if (tr > 0.0) {
    const double s = std::sqrt(tr + 1.0) * 2.0;
    q[0] = 0.25 * s;
    q[1] = (R[2][1] - R[1][2]) / s;
    q[2] = (R[0][2] - R[2][0]) / s;
    q[3] = (R[1][0] - R[0][1]) / s;
  } else if (R[0][0] > R[1][1] && R[0][0] > R[2][2]) {
    const double s = std::sqrt(1.0 + R[0][0] - R[1][1] - R[2][2]) * 2.0;
    q[0] = (R[2][1] - R[1][2]) / s;
    q[1] = 0.25 * s;
    q[2] = (R[0][1] + R[1][0]) / s;
    q[3] = (R[0][2] + R[2][0]) / s;
  } else if (R[1][1] > R[2][2]) {
    const double s = std::sqrt(1.0 + R[1][1] - R[0][0] - R[2][2]) * 2.0;
    q[0] = (R[0][2] - R[2][0]) / s;
    q[1] = (R[0][1] + R[1][0]) / s;
    q[2] = 0.25 * s;
    q[3] = (R[1][2] + R[2][1]) / s;
  } else {
    const double s = std::sqrt(1.0 + R[2][2] - R[0][0] - R[1][1]) * 2.0;
    q[0] = (R[1][0] - R[0][1]) / s;
    q[1] = (R[0][2] + R[2][0]) / s;
    q[2] = (R[1][2] + R[2][1]) / s;
    q[3] = 0.25 * s;
  }
  if (q[0] < 0) {
    for (double& v : q) v = -v;
  }
  const double vn = std::sqrt(q[1] * q[1] + q[2] * q[2] + q[3] * q[3]);
  double k;  // w = k * q_vec
  if (vn < 1e-10)
    k = 2.0 / q[0];  // 2*atan2(vn,w)/vn -> 2/w
  else
    k = 2.0 * std::atan2(vn, q[0]) / vn;
  w[0] = k * q[1];
  w[1] = k * q[2];
  w[2] = k * q[3];
}

// xi = (v, w)
inline Pose se3_exp(const double xi[6]) {
  Pose p;
  const double* v = xi;
  const double* w = xi + 3;
  so3_exp(w, p.R);
  const double th2 = w[0] * w[0] + w[1] * w[1] + w[2] * w[2];
  const double th = std::sqrt(th2);
  double B, C;  // B=(1-cos)/th^2, C=(th-sin)/th^3
  if (th < 0.05) {
    // Taylor series: below ~0.05 rad the closed forms cancel catastrophically ((1 - cos th) carries an absolute error
    // of 1e-16 against a value of th^2 / 2), which a finite-difference Jacobian of log(D exp(e)) then amplifies by 1/h
    B = 0.5 - th2 / 24.0 + th2 * th2 / 720.0 - th2 * th2 * th2 / 40320.0 + th2 * th2 * th2 * th2 / 3628800.0;
    C = 1.0 / 6.0 - th2 / 120.0 + th2 * th2 / 5040.0 - th2 * th2 * th2 / 362880.0 + th2 * th2 * th2 * th2 / 39916800.0;
  } else {
    B = (1.0 - std::cos(th)) / th2;
    C = (th - std::sin(th)) / (th2 * th);
  }
  // V = I + B [w]x + C [w]x^2 ;  t = V v
  const double wxv[3] = {w[1] * v[2] - w[2] * v[1], w[2] * v[0] - w[0] * v[2], w[0] * v[1] - w[1] * v[0]};
  const double wxwxv[3] = {w[1] * wxv[2] - w[2] * wxv[1], w[2] * wxv[0] - w[0] * wxv[2],
                           w[0] * wxv[1] - w[1] * wxv[0]};
  for (int i = 0; i < 3; i++) p.t[i] = v[i] + B * wxv[i] + C * wxwxv[i];
  return p;
}

inline void se3_log(const Pose& p, double xi[6]) {
  double* v = xi;
  double* w = xi + 3;
  so3_log(p.R, w);
  const double th2 = w[0] * w[0] + w[1] * w[1] + w[2] * w[2];
  const double th = std::sqrt(th2);
  double D;  // V^-1 = I - 1/2 [w]x + D [w]x^2
  if (th < 0.05)  // series of (1 - (th/2) cot(th/2)) / th^2: the closed form below loses all digits for small th
    D = 1.0 / 12.0 + th2 / 720.0 + th2 * th2 / 30240.0 + th2 * th2 * th2 / 1209600.0 + th2 * th2 * th2 * th2 / 47900160.0;
  else
    D = (1.0 - (th * std::sin(th)) / (2.0 * (1.0 - std::cos(th)))) / th2;
  const double* t = p.t;
  const double wxt[3] = {w[1] * t[2] - w[2] * t[1], w[2] * t[0] - w[0] * t[2], w[0] * t[1] - w[1] * t[0]};
  const double wxwxt[3] = {w[1] * wxt[2] - w[2] * wxt[1], w[2] * wxt[0] - w[0] * wxt[2],
                           w[0] * wxt[1] - w[1] * wxt[0]};
  for (int i = 0; i < 3; i++) v[i] = t[i] - 0.5 * wxt[i] + D * wxwxt[i];
}

inline double norm3(const double* a) { return std::sqrt(a[0] * a[0] + a[1] * a[1] + a[2] * a[2]); }

// CPose3D::composePoint in double, result stored as float: the operation order
//   g = R00*lx + R01*ly + R02*lz + tx   (left to right, no fused multiply-add)
// is what the device kernel reproduces bit-for-bit.
inline void compose_point_f(const Pose& T, float lx, float ly, float lz, float& gx, float& gy, float& gz) {
  const double x = lx, y = ly, z = lz;
  gx = static_cast<float>(T.R[0][0] * x + T.R[0][1] * y + T.R[0][2] * z + T.t[0]);
  gy = static_cast<float>(T.R[1][0] * x + T.R[1][1] * y + T.R[1][2] * z + T.t[1]);
  gz = static_cast<float>(T.R[2][0] * x + T.R[2][1] * y + T.R[2][2] * z + T.t[2]);
}

// Solve A x = b for symmetric positive-definite 6x6 A by LDL^T (no pivoting). Returns false when
// a pivot is not finite-positive (reported by the caller as IterTermReason::SolverError).
inline bool ldlt6_solve(const double A[6][6], const double b[6], double x[6]) {
  double L[6][6] = {{0}};
  double D[6];
  for (int j = 0; j < 6; j++) {
    double d = A[j][j];
    for (int k = 0; k < j; k++) d -= L[j][k] * L[j][k] * D[k];
    if (!(d > 0.0) || !std::isfinite(d)) return false;
    D[j] = d;
    L[j][j] = 1.0;
    for (int i = j + 1; i < 6; i++) {
      double s = A[i][j];
      for (int k = 0; k < j; k++) s -= L[i][k] * L[j][k] * D[k];
      L[i][j] = s / d;
    }
  }
  double y[6];
  for (int i = 0; i < 6; i++) {
    double s = b[i];
    for (int k = 0; k < i; k++) s -= L[i][k] * y[k];
    y[i] = s;
  }
  for (int i = 0; i < 6; i++) y[i] /= D[i];
  for (int i = 5; i >= 0; i--) {
    double s = y[i];
    for (int k = i + 1; k < 6; k++) s -= L[k][i] * x[k];
    x[i] = s;
  }
  return true;
}

// Inverse of a symmetric positive-definite 6x6 through six LDL^T solves.
inline bool spd6_inverse(const double A[6][6], double inv[6][6]) {
  for (int c = 0; c < 6; c++) {
    double e[6] = {0, 0, 0, 0, 0, 0}, x[6];
    e[c] = 1.0;
    if (!ldlt6_solve(A, e, x)) return false;
    for (int r = 0; r < 6; r++) inv[r][c] = x[r];
  }
  return true;
}

}  // namespace orc
