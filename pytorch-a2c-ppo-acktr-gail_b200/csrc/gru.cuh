// Descriptor of the GRU gate kernels (mirrors a2cb_gru_t in include/a2cb.h).
#pragma once
#include <stdint.h>

namespace a2cb {

struct GruDesc {
    const float* gi;       // [T*N, 3H]  x W_ih^T + b_ih
    const float* gh;       // [T*N, 3H]  hm_t W_hh^T + b_hh (written per step by the GEMM engine)
    float* hm;             // [T*N, H]   masked previous state of every step (input of the per-step GEMM)
    const float* h0;       // [N, H]     initial state (init op)
    const float* masks;    // flat rollout masks; row (t, n) -> masks[idx ? idx[t*N+n] : t*N+n]
    const int64_t* idx;    // optional minibatch row list (time-major t*N + n)
    float* out;            // [T*N, H]   h_t
    float* r; float* z; float* n;   // [T*N, H] saved gates (NULL in no-grad forward)
    const float* dout;     // [T*N, H]   d loss / d h_t from the layers above
    float* dgi; float* dgh;          // [T*N, 3H]
    float* dcarry;         // [N, H]
    int32_t T, N, H, t;
};

}  // namespace a2cb
