#pragma once
#include <mrpt/math/CMatrixFixed.h>
namespace mrpt::poses {
class CPose3D {
 public:
  CPose3D() = default;
  explicit CPose3D(const mrpt::math::TPose3D&);
  static CPose3D FromRotationAndTranslation(const mrpt::math::CMatrixDouble33& R, const double (&t)[3]);
  const mrpt::math::CMatrixDouble33& getRotationMatrix() const;
  double x() const; double y() const; double z() const;
  mrpt::math::TPose3D asTPose() const;
};
struct CPose3DPDFGaussian { CPose3D mean; mrpt::math::CMatrixDouble66 cov; };
struct CPose3DPDFGaussianInf { CPose3D mean; mrpt::math::CMatrixDouble66 cov_inv; };
}  // namespace mrpt::poses
