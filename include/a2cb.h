/*
 * a2cb.h -- C ABI of liba2cb.so, the sm_100a implementation of the rollout-batch
 * policy-update hot path of ikostrikov/pytorch-a2c-ppo-acktr-gail.
 *
 * The reference has no FFI of its own: its seam is the Python class API of the
 * `a2c_ppo_acktr` package.  Each entry point below names the
 * reference code (file:line under the reference tree) whose arithmetic it replaces;
 * the Python classes in pytorch-a2c-ppo-acktr-gail_b200/a2c_ppo_acktr/ bind these
 * through ctypes (see INTEGRATION.md for the binding a maintainer would add).
 *
 * Conventions
 *  - plain C types only: device pointers, sizes, a CUstream; no torch types;
 *  - every function returns 0 on success, a negative A2CB_ERR_* otherwise;
 *    a2cb_last_error() returns a thread-local description of the last failure;
 *  - caller owns all memory and workspaces; no hidden allocation and no host
 *    synchronisation unless a function's comment says so;
 *  - all floating tensors are fp32, row-major, in the reference's layouts.
 */
#ifndef A2CB_H_
#define A2CB_H_

#include <stdint.h>

#ifdef __cplusplus
extern "C" {
#endif

#define A2CB_ABI_VERSION 1

#define A2CB_OK 0
#define A2CB_ERR_INVALID (-1)     /* bad argument */
#define A2CB_ERR_CUDA (-2)        /* CUDA runtime / launch failure */
#define A2CB_ERR_UNSUPPORTED (-3) /* shape or mode not implemented */
#define A2CB_ERR_NCCL (-4)

typedef void* a2cb_stream_t; /* cudaStream_t */

int a2cb_abi_version(void);
const char* a2cb_last_error(void);
/* Number of kernels this library has launched in the calling process (bench.py's
 * "gpu_launches" evidence). */
int64_t a2cb_launch_count(void);

/* ---------------------------------------------------------------------------
 * (1) RolloutStorage.compute_returns              a2c_ppo_acktr/storage.py:66-105
 *
 * rewards [T,N,1]; value_preds, returns, masks, bad_masks [T+1,N,1]; next_value [N,1].
 * Writes returns[0..T) and, as the reference does, value_preds[T] = next_value
 * (use_gae) or returns[T] = next_value (otherwise).
 * gamma / gae_lambda are doubles because the reference multiplies them in Python
 * double precision before rounding to fp32 (storage.py:79).
 * schedule: 0 auto, 1 env-parallel (bit-exact to the reference's operation order),
 *           2 time-parallel warp scan (<= 1e-5 relative).
 * ------------------------------------------------------------------------- */
int a2cb_returns_scan(const float* rewards, float* value_preds, float* returns, const float* masks,
                      const float* bad_masks, const float* next_value, int32_t T, int64_t N,
                      double gamma, double gae_lambda, int32_t use_gae,
                      int32_t use_proper_time_limits, int32_t schedule, a2cb_stream_t stream);

/* ---------------------------------------------------------------------------
 * (2) minibatch shuffle-gather                    a2c_ppo_acktr/storage.py:126-140
 *                                                 (recurrent: storage.py:154-199)
 * Copies rows bit-exactly.  Field f is a [rows, row_bytes] byte matrix.
 *   mode 0: out row j <- src row idx[j]                     (feed_forward_generator)
 *   mode 1: out row j <- src row (j / E) * N + idx[j % E]   (recurrent_generator,
 *           rows ordered t*E + e; idx holds E environment numbers)
 * idx is a device pointer to int64.
 * ------------------------------------------------------------------------- */
typedef struct {
    const void* src;
    void* dst;
    int64_t row_bytes;
} a2cb_gather_field_t;

int a2cb_gather_rows(const a2cb_gather_field_t* fields, int32_t n_fields, const int64_t* idx,
                     int64_t n_out_rows, int32_t mode, int64_t E, int64_t N, a2cb_stream_t stream);

#ifdef __cplusplus
}
#endif
#endif /* A2CB_H_ */
