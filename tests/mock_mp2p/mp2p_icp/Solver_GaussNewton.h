#pragma once
#include <mp2p_icp/ICP.h>
namespace mp2p_icp {
// parameters of pipelines/lidar3d-default.yaml:185-190
class Solver_GaussNewton : public Solver {
 public:
  uint32_t maxIterations = 6;
  RobustKernel robustKernel = RobustKernel::None;
  double robustKernelParam = 1.0, minDelta = 1e-7;
};
}  // namespace mp2p_icp
