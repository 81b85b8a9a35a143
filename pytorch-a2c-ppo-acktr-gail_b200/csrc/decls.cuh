// Internal C++ entry points implemented by the .cu files and wrapped by capi.cu.
#pragma once
#include <cuda_runtime.h>
#include <stdint.h>

namespace a2cb {

// returns_scan.cu
int returns_scan(const float* rewards, float* value_preds, float* returns, const float* masks,
                 const float* bad_masks, const float* next_value, int T, long N, double gamma,
                 double gae_lambda, int use_gae, int use_ptl, int schedule, cudaStream_t st);

// gather.cu
struct GatherField { const void* src; void* dst; long row_bytes; };
int gather_rows(const GatherField* fields, int n_fields, const int64_t* idx, long n_rows, int mode,
                long E, long N, cudaStream_t st);

}  // namespace a2cb
