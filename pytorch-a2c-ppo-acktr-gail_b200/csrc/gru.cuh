// Descriptor of the GRU gate kernels (mirrors a2cb_gru_t in include/a2cb.h).
#pragma once
#include <stdint.h>

namespace a2cb {

struct GruDesc {
    const float* gi;       // [T*N, 3H]  x W_ih^T + b_ih
    float* gh;             // [T*N, 3H]  hm_t W_hh^T + b_hh (written per step by the GEMM engine, or by the
                           //            forward gate kernel when it folds gh_part)
    float* hm;             // [T*N, H]   masked previous state of every step (input of the per-step GEMM)
    const float* h0;       // [N, H]     initial state (init op)
    const float* masks;    // flat rollout masks; row (t, n) -> masks[idx ? idx[t*N+n] : t*N+n]
    const int64_t* idx;    // optional minibatch row list (time-major t*N + n)
    float* out;            // [T*N, H]   h_t
    float* r; float* z; float* n;   // [T*N, H] saved gates (NULL in no-grad forward)
    const float* dout;     // [T*N, H]   d loss / d h_t from the layers above
    float* dgi; float* dgh;          // [T*N, 3H]
    float* dcarry;         // [N, H]
    // split-K partials of the per-step GEMMs, folded HERE instead of by a separate launch:
    const float* gh_part;  // [gh_splits][N, 3H] partial sums of hm_t W_hh^T (no bias); NULL -> gh is final
    const float* gh_bias;  // [3H] b_hh, added when folding gh_part
    const float* dc_part;  // [dc_splits][N, H] partial sums of dgh_{t+1} W_hh; NULL -> dcarry is final
    int32_t T, N, H, t;
    int32_t gh_splits, dc_splits;
};

}  // namespace a2cb
