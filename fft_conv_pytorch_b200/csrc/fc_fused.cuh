// fc_fused.cuh — fused / specialised sm_100a kernels for the benchmark shapes (checked against fc_kernels.cuh).
#pragma once
#include "fc_kernels.cuh"
#include "fc_plan.h"

static inline void fc_fused_init() {}
static inline void fc_fused_plan(fc_plan* pl) { pl->fused.enabled = 0; }
static inline int fc_fused_conv(const fc_plan*, const float2*, const float*, const float2*, const float*, float*, void*, cudaStream_t) {
  return FC_EUNSUPPORTED;
}
static inline int fc_fused_launch_info(const fc_plan*, int, std::string*, int64_t*) { return FC_EINVAL; }
