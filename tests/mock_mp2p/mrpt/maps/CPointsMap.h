#pragma once
#include <cstddef>
#include <memory>
#include <vector>
namespace mrpt::maps {
// SoA float32 point vectors: the in-memory layout the matcher reads
class CMetricMap { public: virtual ~CMetricMap() = default; using Ptr = std::shared_ptr<CMetricMap>; };
class CPointsMap : public CMetricMap {
 public:
  using Ptr = std::shared_ptr<CPointsMap>;
  std::size_t size() const;
  const std::vector<float>& getPointsBufferRef_x() const;
  const std::vector<float>& getPointsBufferRef_y() const;
  const std::vector<float>& getPointsBufferRef_z() const;
};
}  // namespace mrpt::maps
