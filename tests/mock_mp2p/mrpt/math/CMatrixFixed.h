#pragma once
namespace mrpt::math {
template <typename T, int R, int C>
struct CMatrixFixed {
  T v[R * C] = {};
  T& operator()(int r, int c) { return v[r * C + c]; }
  const T& operator()(int r, int c) const { return v[r * C + c]; }
};
using CMatrixDouble66 = CMatrixFixed<double, 6, 6>;
using CMatrixDouble33 = CMatrixFixed<double, 3, 3>;
struct TPose3D { double x = 0, y = 0, z = 0, yaw = 0, pitch = 0, roll = 0; };
struct TPoint3Df { float x, y, z; };
}  // namespace mrpt::math
