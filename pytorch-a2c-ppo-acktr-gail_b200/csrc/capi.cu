// extern "C" surface of liba2cb.so -- see include/a2cb.h for the contract.
#include "../../include/a2cb.h"
#include "common.cuh"
#include "decls.cuh"

#include <atomic>
#include <stdarg.h>

namespace a2cb {

static thread_local char g_err[512] = "";
static std::atomic<int64_t> g_launches{0};

void set_error(const char* fmt, ...) {
    va_list ap;
    va_start(ap, fmt);
    vsnprintf(g_err, sizeof(g_err), fmt, ap);
    va_end(ap);
}

int check_launch(const char* what) {
    g_launches.fetch_add(1, std::memory_order_relaxed);
    cudaError_t e = cudaPeekAtLastError();
    if (e != cudaSuccess) {
        cudaGetLastError();
        set_error("%s: %s", what, cudaGetErrorString(e));
        return A2CB_ERR_CUDA;
    }
    return A2CB_OK;
}

}  // namespace a2cb

using namespace a2cb;

extern "C" {

int a2cb_abi_version(void) { return A2CB_ABI_VERSION; }
const char* a2cb_last_error(void) { return g_err; }
int64_t a2cb_launch_count(void) { return g_launches.load(); }

int a2cb_returns_scan(const float* rewards, float* value_preds, float* returns, const float* masks,
                      const float* bad_masks, const float* next_value, int32_t T, int64_t N,
                      double gamma, double gae_lambda, int32_t use_gae,
                      int32_t use_proper_time_limits, int32_t schedule, a2cb_stream_t stream) {
    A2CB_REQUIRE((T == 0 || rewards) && value_preds && returns && (T == 0 || masks) && next_value, "returns_scan: null pointer");
    A2CB_REQUIRE(T == 0 || !use_proper_time_limits || bad_masks, "returns_scan: bad_masks required");
    A2CB_REQUIRE(T >= 0 && N >= 0 && schedule >= 0 && schedule <= 2, "returns_scan: bad T/N/schedule");
    if (N == 0) return A2CB_OK;
    return returns_scan(rewards, value_preds, returns, masks, bad_masks, next_value, T, (long)N, gamma,
                        gae_lambda, use_gae, use_proper_time_limits, schedule, (cudaStream_t)stream);
}

int a2cb_gather_rows(const a2cb_gather_field_t* fields, int32_t n_fields, const int64_t* idx,
                     int64_t n_out_rows, int32_t mode, int64_t E, int64_t N, a2cb_stream_t stream) {
    A2CB_REQUIRE(n_fields >= 0 && n_out_rows >= 0, "gather_rows: negative size");
    A2CB_REQUIRE(n_fields == 0 || n_out_rows == 0 || (fields && idx), "gather_rows: null pointer");
    static_assert(sizeof(a2cb_gather_field_t) == sizeof(GatherField), "layout");
    return gather_rows(reinterpret_cast<const GatherField*>(fields), n_fields, idx, (long)n_out_rows, mode,
                       (long)E, (long)N, (cudaStream_t)stream);
}

}  // extern "C"
