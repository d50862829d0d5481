// Draft/verify engine: buffer layout, per-step kernel schedule.
#pragma once
#include "fused_ops.cuh"

namespace dfl {
int cuda_fail(cudaError_t e, const char* what);  // api.cu
}  // namespace dfl
