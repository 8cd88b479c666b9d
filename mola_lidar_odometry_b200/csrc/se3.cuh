// se3.cuh — double-precision SE(3) algebra used by the single-warp solve step (host + device).
// Tangent vectors are (rho, phi) = translation first, rotation second; updates are right-multiplicative
// (T <- T * exp(delta)), the convention of mp2p_icp::Solver_GaussNewton (pipelines/lidar3d-default.yaml:185-190)
// and of the stall test in mp2p_icp::ICP::align (default.yaml:174-175).
#pragma once
#include "common.cuh"

namespace mlo {

MLO_HD void mat3_mul(const double* A, const double* B, double* C) {
  #pragma unroll
  for (int i = 0; i < 3; i++)
    #pragma unroll
    for (int j = 0; j < 3; j++) C[3 * i + j] = A[3 * i] * B[j] + A[3 * i + 1] * B[3 + j] + A[3 * i + 2] * B[6 + j];
}
MLO_HD void hat(const double* w, double* W) {
  W[0] = 0; W[1] = -w[2]; W[2] = w[1];
  W[3] = w[2]; W[4] = 0; W[5] = -w[0];
  W[6] = -w[1]; W[7] = w[0]; W[8] = 0;
}
MLO_HD void cross3(const double* a, const double* b, double* c) {
  c[0] = a[1] * b[2] - a[2] * b[1];
  c[1] = a[2] * b[0] - a[0] * b[2];
  c[2] = a[0] * b[1] - a[1] * b[0];
}
MLO_HD double nrm3(const double* a) { return sqrt(a[0] * a[0] + a[1] * a[1] + a[2] * a[2]); }

// C = A * B for 3x4 poses
MLO_HD void pose_mul(const double* A, const double* B, double* C) {
  #pragma unroll
  for (int r = 0; r < 3; r++) {
    #pragma unroll
    for (int c = 0; c < 3; c++) C[4 * r + c] = A[4 * r] * B[c] + A[4 * r + 1] * B[4 + c] + A[4 * r + 2] * B[8 + c];
    C[4 * r + 3] = A[4 * r] * B[3] + A[4 * r + 1] * B[7] + A[4 * r + 2] * B[11] + A[4 * r + 3];
  }
}
MLO_HD void pose_inv(const double* A, double* C) {
  #pragma unroll
  for (int r = 0; r < 3; r++)
    #pragma unroll
    for (int c = 0; c < 3; c++) C[4 * r + c] = A[4 * c + r];
  #pragma unroll
  for (int r = 0; r < 3; r++) C[4 * r + 3] = -(C[4 * r] * A[3] + C[4 * r + 1] * A[7] + C[4 * r + 2] * A[11]);
}
// B^-1 * A  ("A - B" in MRPT notation, LidarOdometry.cpp:930-931)
MLO_HD void pose_minus(const double* A, const double* B, double* C) {
  double Bi[12];
  pose_inv(B, Bi);
  pose_mul(Bi, A, C);
}

// Rodrigues coefficients with series near zero: a = sin(t)/t, b = (1-cos t)/t^2, c = (t - sin t)/t^3
MLO_HD void rodrigues_coeffs(double t2, double& a, double& b, double& c) {
  if (t2 < 2.5e-3) {  // |phi| < 0.05 rad: Taylor series to t^8, truncation error < 1e-18 (no sin/cos calls)
    a = 1.0 - t2 * (1.0 / 6.0) * (1.0 - t2 * (1.0 / 20.0) * (1.0 - t2 * (1.0 / 42.0) * (1.0 - t2 * (1.0 / 72.0))));
    b = 0.5 * (1.0 - t2 * (1.0 / 12.0) * (1.0 - t2 * (1.0 / 30.0) * (1.0 - t2 * (1.0 / 56.0) * (1.0 - t2 * (1.0 / 90.0)))));
    c = (1.0 / 6.0) * (1.0 - t2 * (1.0 / 20.0) * (1.0 - t2 * (1.0 / 42.0) * (1.0 - t2 * (1.0 / 72.0) * (1.0 - t2 * (1.0 / 110.0)))));
  } else {
    const double t = sqrt(t2);
    a = sin(t) / t;
    b = (1.0 - cos(t)) / t2;
    c = (t - sin(t)) / (t2 * t);
  }
}

MLO_HD void se3_exp(const double* xi, double* T) {
  const double* rho = xi;
  const double* phi = xi + 3;
  const double t2 = phi[0] * phi[0] + phi[1] * phi[1] + phi[2] * phi[2];
  double a, b, c;
  rodrigues_coeffs(t2, a, b, c);
  double W[9], W2[9];
  hat(phi, W);
  mat3_mul(W, W, W2);
  #pragma unroll
  for (int i = 0; i < 3; i++)
    #pragma unroll
    for (int j = 0; j < 3; j++) T[4 * i + j] = (i == j ? 1.0 : 0.0) + a * W[3 * i + j] + b * W2[3 * i + j];
  // t = (I + b W + c W^2) rho
  #pragma unroll
  for (int i = 0; i < 3; i++) {
    double s = rho[i];
    #pragma unroll
    for (int j = 0; j < 3; j++) s += (b * W[3 * i + j] + c * W2[3 * i + j]) * rho[j];
    T[4 * i + 3] = s;
  }
}

// rotation vector of the 3x3 block of a 3x4 pose, via the unit quaternion (valid on [0, pi])
MLO_HD void so3_log_of_pose(const double* T, double* w) {
  const double r00 = T[0], r01 = T[1], r02 = T[2], r10 = T[4], r11 = T[5], r12 = T[6], r20 = T[8], r21 = T[9], r22 = T[10];
  double qw, qx, qy, qz;
  const double tr = r00 + r11 + r22;
  if (tr > 0.0) {
    const double s = 2.0 * sqrt(tr + 1.0);
    qw = 0.25 * s; qx = (r21 - r12) / s; qy = (r02 - r20) / s; qz = (r10 - r01) / s;
  } else if (r00 > r11 && r00 > r22) {
    const double s = 2.0 * sqrt(1.0 + r00 - r11 - r22);
    qw = (r21 - r12) / s; qx = 0.25 * s; qy = (r01 + r10) / s; qz = (r02 + r20) / s;
  } else if (r11 > r22) {
    const double s = 2.0 * sqrt(1.0 + r11 - r00 - r22);
    qw = (r02 - r20) / s; qx = (r01 + r10) / s; qy = 0.25 * s; qz = (r12 + r21) / s;
  } else {
    const double s = 2.0 * sqrt(1.0 + r22 - r00 - r11);
    qw = (r10 - r01) / s; qx = (r02 + r20) / s; qy = (r12 + r21) / s; qz = 0.25 * s;
  }
  if (qw < 0) { qw = -qw; qx = -qx; qy = -qy; qz = -qz; }
  const double vn = sqrt(qx * qx + qy * qy + qz * qz);
  double k;
  if (vn < 0.03 * qw) {  // small angle: 2 atan(x)/vn with x = vn/qw, alternating series to x^14 (error < 1e-20)
    const double x2 = (vn / qw) * (vn / qw);
    const double ser = 1.0 - x2 * ((1.0 / 3.0) - x2 * ((1.0 / 5.0) - x2 * ((1.0 / 7.0) - x2 * ((1.0 / 9.0) - x2 * ((1.0 / 11.0) - x2 * (1.0 / 13.0))))));
    k = 2.0 * ser / qw;
  } else {
    k = 2.0 * atan2(vn, qw) / vn;
  }
  w[0] = k * qx; w[1] = k * qy; w[2] = k * qz;
}

MLO_HD void se3_log(const double* T, double* xi) {
  double* rho = xi;
  double* phi = xi + 3;
  so3_log_of_pose(T, phi);
  const double t2 = phi[0] * phi[0] + phi[1] * phi[1] + phi[2] * phi[2];
  // V^-1 = I - W/2 + d W^2,  d = (1 - t sin t / (2 (1 - cos t))) / t^2
  double d;
  if (t2 < 2.5e-3)  // series of (1 - (t/2) cot(t/2)) / t^2 (Bernoulli numbers), truncation error < 1e-18 for |phi| < 0.05
    d = (1.0 / 12.0) + t2 * ((1.0 / 720.0) + t2 * ((1.0 / 30240.0) + t2 * ((1.0 / 1209600.0) + t2 * (1.0 / 47900160.0))));
  else { const double t = sqrt(t2); d = (1.0 - (t * sin(t)) / (2.0 * (1.0 - cos(t)))) / t2; }
  const double tt[3] = {T[3], T[7], T[11]};
  double a[3], b[3];
  cross3(phi, tt, a);
  cross3(phi, a, b);
  #pragma unroll
  for (int i = 0; i < 3; i++) rho[i] = tt[i] - 0.5 * a[i] + d * b[i];
}

// Closed-form d log(D exp(eps)) / d eps at eps = 0, i.e. the inverse right Jacobian of SE(3) at
// xi = log(D) (the e2 half of MRPT's jacob_dDinvP1invP2_de1e2 used for the prior term,
// LidarOdometry.cpp:854-877 -> Solver_GaussNewton prior).  J is 6x6 row-major.
// Scalar coefficients of Jl_so3^-1 and of the Q block (se3_right_jacobian_inv below; csrc/icp.cuh jr_inv_warp)
MLO_HD void jr_inv_coeffs(double t2, double& k, double& c1, double& c2, double& c3) {
  if (t2 < 2.5e-3) {
    // Taylor series for |phi| < 0.05 rad (truncation < 1e-16).  The closed forms below cancel catastrophically for small
    // angles: c2's numerator t^2 + 2 cos t - 2 ~ t^4 / 12 is pure round-off below t ~ 1e-2 (found against the BCH series,
    // tests/test_oracle_independent.py::test_se3_small_angle_accuracy_against_series).
    k = (1.0 / 12.0) + t2 * ((1.0 / 720.0) + t2 * ((1.0 / 30240.0) + t2 * (1.0 / 1209600.0)));
    c1 = (1.0 / 6.0) - t2 * ((1.0 / 120.0) - t2 * ((1.0 / 5040.0) - t2 * (1.0 / 362880.0)));
    c2 = (1.0 / 24.0) - t2 * ((1.0 / 720.0) - t2 * ((1.0 / 40320.0) - t2 * (1.0 / 3628800.0)));
    c3 = (1.0 / 120.0) - t2 * ((1.0 / 2520.0) - t2 * ((1.0 / 120960.0) - t2 * (1.0 / 9979200.0)));
  } else {
    const double t = sqrt(t2), s = sin(t), c = cos(t);
    k = 1.0 / t2 - (1.0 + c) / (2.0 * t * s);
    c1 = (t - s) / (t2 * t);
    c2 = (t2 + 2.0 * c - 2.0) / (2.0 * t2 * t2);
    c3 = (2.0 * t - 3.0 * s + t * c) / (2.0 * t2 * t2 * t);
  }
}

MLO_HD void se3_right_jacobian_inv(const double* xi, double* J) {
  // Jr^-1(xi) = Jl^-1(-xi); with Jl^-1 = [[A, -A Q A], [0, A]], A = Jl_so3^-1(phi)
  const double rho[3] = {-xi[0], -xi[1], -xi[2]};
  const double phi[3] = {-xi[3], -xi[4], -xi[5]};
  const double t2 = phi[0] * phi[0] + phi[1] * phi[1] + phi[2] * phi[2];
  double P[9], F[9], FF[9];
  hat(rho, P);
  hat(phi, F);
  mat3_mul(F, F, FF);
  // A = I - F/2 + k FF,  k = 1/t^2 - (1 + cos t) / (2 t sin t)
  double k, c1, c2, c3;
  jr_inv_coeffs(t2, k, c1, c2, c3);
  double A[9];
  #pragma unroll
  for (int i = 0; i < 9; i++) A[i] = ((i % 4 == 0) ? 1.0 : 0.0) - 0.5 * F[i] + k * FF[i];
  // Q = P/2 + c1 (FP + PF + FPF) + c2 (FFP + PFF - 3 FPF) + c3 (FPFF + FFPF)
  double FP[9], PF[9], FPF[9], FFP[9], PFF[9], FPFF[9], FFPF[9];
  mat3_mul(F, P, FP);
  mat3_mul(P, F, PF);
  mat3_mul(FP, F, FPF);
  mat3_mul(F, FP, FFP);
  mat3_mul(PF, F, PFF);
  mat3_mul(FPF, F, FPFF);
  mat3_mul(F, FPF, FFPF);
  double Q[9];
  #pragma unroll
  for (int i = 0; i < 9; i++)
    Q[i] = 0.5 * P[i] + c1 * (FP[i] + PF[i] + FPF[i]) + c2 * (FFP[i] + PFF[i] - 3.0 * FPF[i]) + c3 * (FPFF[i] + FFPF[i]);
  double AQ[9], AQA[9];
  mat3_mul(A, Q, AQ);
  mat3_mul(AQ, A, AQA);
  #pragma unroll
  for (int i = 0; i < 3; i++)
    #pragma unroll
    for (int j = 0; j < 3; j++) {
      J[6 * i + j] = A[3 * i + j];
      J[6 * i + 3 + j] = -AQA[3 * i + j];
      J[6 * (i + 3) + j] = 0.0;
      J[6 * (i + 3) + 3 + j] = A[3 * i + j];
    }
}

// LDL^T solve of a symmetric positive-definite 6x6 system (row-major H). Returns false on a
// non-positive / non-finite pivot (reported as IterTermReason::SolverError).
MLO_HD bool ldlt6(const double* H, const double* b, double* x) {
  double L[36];
  double D[6];
#pragma unroll
  for (int i = 0; i < 36; i++) L[i] = 0.0;
#pragma unroll
  for (int j = 0; j < 6; j++) {
    double d = H[6 * j + j];
#pragma unroll
    for (int k = 0; k < j; k++) d -= L[6 * j + k] * L[6 * j + k] * D[k];
    if (!(d > 0.0) || !(d < 1e300)) return false;
    D[j] = d;
    const double inv_d = 1.0 / d;
    L[6 * j + j] = 1.0;
#pragma unroll
    for (int i = j + 1; i < 6; i++) {
      double s = H[6 * i + j];
#pragma unroll
      for (int k = 0; k < j; k++) s -= L[6 * i + k] * L[6 * j + k] * D[k];
      L[6 * i + j] = s * inv_d;
    }
  }
  double y[6];
#pragma unroll
  for (int i = 0; i < 6; i++) {
    double s = b[i];
#pragma unroll
    for (int k = 0; k < i; k++) s -= L[6 * i + k] * y[k];
    y[i] = s;
  }
#pragma unroll
  for (int i = 0; i < 6; i++) y[i] /= D[i];
#pragma unroll
  for (int i = 5; i >= 0; i--) {
    double s = y[i];
#pragma unroll
    for (int k = i + 1; k < 6; k++) s -= L[6 * k + i] * x[k];
    x[i] = s;
  }
  return true;
}

MLO_HD bool spd6_inverse(const double* H, double* inv) {
  #pragma unroll
  for (int c = 0; c < 6; c++) {
    double e[6] = {0, 0, 0, 0, 0, 0}, x[6];
    e[c] = 1.0;
    if (!ldlt6(H, e, x)) return false;
    #pragma unroll
    for (int r = 0; r < 6; r++) inv[6 * r + c] = x[r];
  }
  return true;
}

}  // namespace mlo
