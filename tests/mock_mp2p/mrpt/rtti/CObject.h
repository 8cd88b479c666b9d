#pragma once
// mock of the MRPT runtime class registry (module/src/LidarOdometry.cpp:2122-2123, module/src/register.cpp:40-46)
#include <memory>
namespace mrpt::rtti {
struct TRuntimeClassId { const char* className; };
class CObject { public: virtual ~CObject() = default; using Ptr = std::shared_ptr<CObject>; };
void registerClass(const TRuntimeClassId* id);
}  // namespace mrpt::rtti
#define DEFINE_MRPT_OBJECT(cls, ns) public: static const mrpt::rtti::TRuntimeClassId runtimeClassId; using Ptr = std::shared_ptr<cls>;
#define IMPLEMENTS_MRPT_OBJECT(cls, base, ns) const mrpt::rtti::TRuntimeClassId cls::runtimeClassId = {#ns "::" #cls};
#define CLASS_ID(T) (&T::runtimeClassId)
