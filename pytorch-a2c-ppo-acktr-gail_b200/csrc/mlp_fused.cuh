// Descriptor of the fused MLPBase kernels (mlp_fused.cu); mirrors a2cb_mlp_t in include/a2cb.h.
#pragma once
#include <stdint.h>

namespace a2cb {

struct MlpDesc {
    const float* P;        // flat parameters
    float* G;              // flat gradients (backward, batches of <= 64 rows write here directly)
    const float* obs;      // [rows, d_in] fp32 observations
    const int64_t* idx;    // optional minibatch row list
    float* acts;           // [4][B][64] saved post-tanh activations: actor l1, actor l2, critic l1, critic l2
    float* z;              // [B][zs] head outputs (value, distribution parameters)
    const float* dz;       // [B][zs] upstream gradient (backward)
    float* partial;        // [ceil(B / 64)][span] per-CTA gradient slabs in the flat layout (backward, B > 64)
    int64_t off_a0w, off_a0b, off_a2w, off_a2b, off_c0w, off_c0b, off_c2w, off_c2b, off_hw, off_hb;
    int64_t span;          // end of the MLP parameters in the flat layout (slab size)
    int32_t B, d_in, H, zs;
};

}  // namespace a2cb
