#pragma once
// mock of the mp2p_icp interface the reference uses: ICP::align call site module/src/LidarOdometry.cpp:961-962, iteration
// hook :923-952, Parameters::maxIterations :956-959, Results fields :964-1011, YAML blocks default.yaml:162-209.
#include <mrpt/maps/CPointsMap.h>
#include <mrpt/poses/CPose3D.h>
#include <mrpt/rtti/CObject.h>

#include <cstdint>
#include <functional>
#include <map>
#include <memory>
#include <optional>
#include <string>
#include <vector>

namespace mp2p_icp {
enum class IterTermReason : uint8_t { Undefined = 0, NoPairings, SolverError, MaxIterations, Stalled, HookRequest };
enum class RobustKernel : uint8_t { None = 0, GemanMcClure, Cauchy };

struct metric_map_t {
  using Ptr = std::shared_ptr<metric_map_t>;
  std::map<std::string, mrpt::maps::CMetricMap::Ptr> layers;
  mrpt::maps::CPointsMap::Ptr point_layer(const std::string& name) const;
};
struct Parameters {
  uint32_t maxIterations = 40;
  double minAbsStep_trans = 5e-4, minAbsStep_rot = 1e-4;
};
struct Pairings { std::size_t size() const; };
struct Results {
  mrpt::poses::CPose3DPDFGaussian optimal_tf;
  std::size_t nIterations = 0;
  IterTermReason terminationReason = IterTermReason::Undefined;
  double quality = 0;
  Pairings finalPairings;
};
struct LogRecord;
class ParameterSource {
 public:
  void updateVariable(const std::string& name, double value);
  void realize();
};
class Parameterizable {
 public:
  ParameterSource* attachedSource();
};
class Matcher : public mrpt::rtti::CObject, public Parameterizable { public: using Ptr = std::shared_ptr<Matcher>; };
class Solver : public mrpt::rtti::CObject, public Parameterizable { public: using Ptr = std::shared_ptr<Solver>; };
using matcher_list_t = std::vector<Matcher::Ptr>;
using solver_list_t = std::vector<Solver::Ptr>;

class ICP : public mrpt::rtti::CObject, public Parameterizable {
 public:
  using Ptr = std::shared_ptr<ICP>;
  struct IterationHook_Input {
    uint32_t currentIteration = 0;
    struct Solution { mrpt::poses::CPose3D optimalPose; };
    const Solution* currentSolution = nullptr;
  };
  struct IterationHook_Output { bool request_stop = false; };
  using iteration_hook_t = std::function<IterationHook_Output(const IterationHook_Input&)>;
  virtual void align(const metric_map_t& pcLocal, const metric_map_t& pcGlobal, const mrpt::math::TPose3D& initialGuessLocalWrtGlobal,
                     const Parameters& p, Results& result,
                     const std::optional<mrpt::poses::CPose3DPDFGaussianInf>& prior = std::nullopt,
                     LogRecord* outputDebugInfo = nullptr);
  void setIterationHook(const iteration_hook_t& hook) { iteration_hook_ = hook; }
  const matcher_list_t& matchers() const;
  const solver_list_t& solvers() const;
 protected:
  iteration_hook_t iteration_hook_;
};
}  // namespace mp2p_icp
