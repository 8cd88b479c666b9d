#pragma once
#include <mp2p_icp/ICP.h>
namespace mp2p_icp {
// pipelines/extras/icp-pipeline_no_motion_model.yaml:24-29
class Solver_Horn : public Solver {};
}  // namespace mp2p_icp
