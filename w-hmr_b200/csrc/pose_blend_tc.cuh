// placeholder until the tcgen05 kernel lands (next commit)
#pragma once
#include <vector>
#include "common.cuh"
namespace whmr {
struct DeviceArena;
constexpr int kTcBodyTile = 256;
struct TcPlan { int ready = 0; };
static inline int tc_plan_create(const SmplDevice&, const std::vector<float>&, DeviceArena&, TcPlan*) { return WHMR_OK; }
static inline int tc_pose_blend_launch(const TcPlan&, const SmplDevice&, int, const void*, int, int, int, float*, cudaStream_t) {
  return set_error(WHMR_E_INVALID, "tensor-core pose-blend kernel not built");
}
}
