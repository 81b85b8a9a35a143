// Descriptor of the fused loss epilogue (mirrors a2cb_loss_t in include/a2cb.h).
#pragma once
#include <stdint.h>

namespace a2cb {

constexpr int kMaxActions = 32;

struct LossDesc {
    const float* z;            // [B, z_stride]: z[b,0] = value, z[b,1..A] = logits | action mean
    const float* logstd;       // [A] (DiagGaussian)
    const int64_t* idx;        // optional minibatch rows into the flat [T*N] rollout fields
    const void* actions;       // int64 [rows, act_stride] (Categorical) | fp32 [rows, act_stride]
    const float* old_logp;     // [rows]
    const float* old_value;    // [rows]
    const float* returns;      // [rows]
    const float* adv;          // [rows] normalised advantages (PPO)
    const float* noise;        // [rows] value noise (Fisher mode)
    float* dz;                 // [B, dz_stride] gradient of the total loss w.r.t. z (may be NULL)
    float* dlogstd;            // [A] (written, not accumulated)
    float* logp_out;           // [B] optional
    float* ent_out;            // [B] optional per-row entropy
    float* loss_out;           // [3] value_loss, action_loss, entropy of this minibatch
    float* loss_accum;         // [3] += the same (running sums over minibatches)
    float* scratch;            // a2cb_loss_scratch_floats(B) floats, zero-initialised once
    int32_t B, A;
    int32_t z_stride, dz_stride;
    int32_t act_stride, dlogstd_stride;
    int32_t dist;              // 0 Categorical, 1 DiagGaussian
    int32_t mode;              // 0 PPO, 1 A2C, 2 Fisher sample, 3 forward only, 4 VJP (upstream grads in old_value/old_logp/noise[0])
    int32_t use_clipped_value_loss;
    float clip, c_v, c_e;
    float inv_B;               // 1 / (global minibatch size)
    int32_t n_valid;           // rows [n_valid, B) are padding (dz = 0, no loss); < 0: all B rows valid
    int32_t valid_mod;         // > 0: row b is valid iff (b % valid_mod) < n_valid (time-major recurrent batches)
    int32_t pad2_;
    float* dlogstd_rows;       // optional [B, A] per-row gradient w.r.t. the broadcast log-std (KFAC g-factor of the AddBias)
};

}  // namespace a2cb
