// Shared helpers for the sm_100a kernels of the rollout-update path.
// Internal header: nothing here is part of the C ABI (see include/a2cb.h).
#pragma once
#include <cuda_runtime.h>
#include <stdint.h>
#include <stdio.h>

#define A2CB_OK 0
#define A2CB_ERR_INVALID (-1)
#define A2CB_ERR_CUDA (-2)
#define A2CB_ERR_UNSUPPORTED (-3)
#define A2CB_ERR_NCCL (-4)

namespace a2cb {

// Thread-local last-error string, readable through a2cb_last_error().
void set_error(const char* fmt, ...);
int check_launch(const char* what);

constexpr int kNumSM = 148;  // B200: 2 dies x 74 SMs

__device__ __forceinline__ float warp_sum(float v) {
#pragma unroll
    for (int o = 16; o > 0; o >>= 1) v += __shfl_xor_sync(0xffffffffu, v, o);
    return v;
}
__device__ __forceinline__ double warp_sum(double v) {
#pragma unroll
    for (int o = 16; o > 0; o >>= 1) v += __shfl_xor_sync(0xffffffffu, v, o);
    return v;
}
__device__ __forceinline__ float warp_max(float v) {
#pragma unroll
    for (int o = 16; o > 0; o >>= 1) v = fmaxf(v, __shfl_xor_sync(0xffffffffu, v, o));
    return v;
}

// Block-wide sum for blockDim.x <= 1024 (multiple of 32). `red` needs 32 slots.
template <typename T>
__device__ __forceinline__ T block_sum(T v, T* red) {
    v = warp_sum(v);
    const int lane = threadIdx.x & 31, w = threadIdx.x >> 5;
    __syncthreads();
    if (lane == 0) red[w] = v;
    __syncthreads();
    const int nw = (blockDim.x + 31) >> 5;
    T r = (threadIdx.x < nw) ? red[threadIdx.x] : T(0);
    if (w == 0) r = warp_sum(r);
    if (threadIdx.x == 0) red[0] = r;
    __syncthreads();
    r = red[0];
    return r;
}

}  // namespace a2cb

#define A2CB_REQUIRE(cond, ...)                 \
    do {                                        \
        if (!(cond)) {                          \
            a2cb::set_error(__VA_ARGS__);       \
            return A2CB_ERR_INVALID;            \
        }                                       \
    } while (0)
