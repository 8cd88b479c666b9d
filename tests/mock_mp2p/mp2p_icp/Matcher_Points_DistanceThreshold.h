#pragma once
#include <mp2p_icp/ICP.h>
namespace mp2p_icp {
// parameters of pipelines/lidar3d-default.yaml:196-204 (formula-valued ones are re-realised per ICP_ITERATION)
class Matcher_Points_DistanceThreshold : public Matcher {
 public:
  double threshold = 0.5, thresholdAngularDeg = 0;
  uint32_t pairingsPerPoint = 1;
  bool allowMatchAlreadyMatchedGlobalPoints = true;
  std::map<std::string, std::map<std::string, double>> weight_pc2pc_layers;  // global layer -> local layer -> weight
};
}  // namespace mp2p_icp
