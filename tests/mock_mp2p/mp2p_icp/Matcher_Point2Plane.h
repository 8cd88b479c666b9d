#pragma once
#include <mp2p_icp/ICP.h>
namespace mp2p_icp {
// parameters of pipelines/lidar3d-ndt.yaml:195-200
class Matcher_Point2Plane : public Matcher {
 public:
  double distanceThreshold = 0.5;
  std::map<std::string, std::map<std::string, double>> weight_pc2pc_layers;
};
}  // namespace mp2p_icp
