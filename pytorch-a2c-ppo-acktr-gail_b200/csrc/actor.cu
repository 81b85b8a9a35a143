// Actor-side head of Policy.act (reference model.py:54-66, distributions.py:18-44): from the head outputs
// z[b] = (value, logits[A] | mean[A]) of the no-grad forward pass, ONE launch draws the action (or takes the mode),
// evaluates its log-probability with the arithmetic of the loss epilogue (loss.cu) and writes value, action and
// log-prob to their destinations -- which may be the [step] slots of the rollout buffer (storage.py:46-58), so the
// per-step softmax / multinomial / gather / copy launches of the reference disappear.
//
// Randomness: Philox4x32-10, key = seed, counter = (row_base + row, draw, offset_lo, offset_hi).  The stream state
// {seed, offset, ticket, row_base} lives in device memory (row_base = first GLOBAL environment of this rank's shard, so
// that ranks seeded identically still draw different numbers for different environments); the last block of a launch advances offset by one, so replaying the
// same op list (or a CUDA graph holding it) keeps drawing fresh numbers without any host involvement.
// The reference samples with torch.multinomial / torch.normal on the device generator; the draws are a different
// stream of the SAME distribution (tests: exact against oracle/actor.py on the Philox stream, chi-square against the
// softmax probabilities).
#include "common.cuh"
#include "decls.cuh"
#include "loss.cuh"

namespace a2cb {

struct SampleDesc {
    const float* z;            // [B, z_stride]
    const float* logstd;       // [A] (DiagGaussian)
    float* value_out;          // [B]
    void* action_out;          // int64 [B] (Categorical) | float [B, A] (DiagGaussian)
    float* logp_out;           // [B]
    unsigned long long* rng;   // device {seed, offset, ticket, row_base}
    int32_t B, A, z_stride, dist, deterministic, pad_;
};

constexpr float kActLogSqrt2Pi = 0.918938533204672742f;

__device__ __forceinline__ void philox_round(uint32_t (&c)[4], uint32_t k0, uint32_t k1) {
    const uint32_t hi0 = __umulhi(0xD2511F53u, c[0]), lo0 = 0xD2511F53u * c[0];
    const uint32_t hi1 = __umulhi(0xCD9E8D57u, c[2]), lo1 = 0xCD9E8D57u * c[2];
    c[0] = hi1 ^ c[1] ^ k0; c[1] = lo1; c[2] = hi0 ^ c[3] ^ k1; c[3] = lo0;
}
__device__ __forceinline__ void philox4x32_10(uint32_t (&c)[4], uint32_t k0, uint32_t k1) {
#pragma unroll
    for (int r = 0; r < 10; ++r) {
        philox_round(c, k0, k1);
        k0 += 0x9E3779B9u; k1 += 0xBB67AE85u;
    }
}
// 24 random bits -> (0, 1): never 0, never 1
__device__ __forceinline__ float u01(uint32_t x) { return ((float)(x >> 8) + 0.5f) * (1.0f / 16777216.0f); }

__global__ void __launch_bounds__(128)
sample_actions_kernel(const SampleDesc d) {
    __shared__ unsigned long long s_seed, s_off, s_base;
    if (threadIdx.x == 0) {
        s_seed = d.rng ? d.rng[0] : 0ull;
        s_off = d.rng ? d.rng[1] : 0ull;
        s_base = d.rng ? d.rng[3] : 0ull;
    }
    __syncthreads();
    const int b = blockIdx.x * blockDim.x + threadIdx.x;
    if (b < d.B) {
        const float* z = d.z + (long)b * d.z_stride;
        const int A = d.A;
        const uint32_t k0 = (uint32_t)s_seed, k1 = (uint32_t)(s_seed >> 32);
        if (d.value_out) d.value_out[b] = z[0];
        if (d.dist == 0) {
            float mx = -INFINITY;
            int amax = 0;
            for (int j = 0; j < A; ++j)
                if (z[1 + j] > mx) { mx = z[1 + j]; amax = j; }       // first maximum, like argmax
            float se = 0.f;
            for (int j = 0; j < A; ++j) se += expf(z[1 + j] - mx);
            const float lse = mx + logf(se);
            int act = amax;
            if (!d.deterministic) {
                uint32_t c[4] = {(uint32_t)s_base + (uint32_t)b, 0u, (uint32_t)s_off, (uint32_t)(s_off >> 32)};
                philox4x32_10(c, k0, k1);
                const float u = u01(c[0]);
                // inverse CDF over p_j = exp(z_j - lse); the last category absorbs the rounding of the sum
                float cum = 0.f;
                act = A - 1;
                for (int j = 0; j < A; ++j) {
                    cum += expf(z[1 + j] - lse);
                    if (u < cum) { act = j; break; }
                }
            }
            reinterpret_cast<int64_t*>(d.action_out)[b] = (int64_t)act;
            if (d.logp_out) d.logp_out[b] = z[1 + act] - lse;
        } else {
            float* x = reinterpret_cast<float*>(d.action_out) + (long)b * A;
            float logp = 0.f;
            for (int a0 = 0; a0 < A; a0 += 4) {
                float n[4] = {0.f, 0.f, 0.f, 0.f};
                if (!d.deterministic) {
                    uint32_t c[4] = {(uint32_t)s_base + (uint32_t)b, (uint32_t)(a0 >> 2), (uint32_t)s_off, (uint32_t)(s_off >> 32)};
                    philox4x32_10(c, k0, k1);
                    // Box-Muller on two pairs of uniforms
                    const float r0 = sqrtf(-2.f * logf(u01(c[0]))), r1 = sqrtf(-2.f * logf(u01(c[2])));
                    float s0, c0, s1, c1;
                    sincospif(2.f * u01(c[1]), &s0, &c0);
                    sincospif(2.f * u01(c[3]), &s1, &c1);
                    n[0] = r0 * c0; n[1] = r0 * s0; n[2] = r1 * c1; n[3] = r1 * s1;
                }
                for (int i = 0; i < 4 && a0 + i < A; ++i) {
                    const int a = a0 + i;
                    const float ls = d.logstd[a], s = expf(ls), var = s * s;
                    const float mu = z[1 + a];
                    const float xa = mu + s * n[i];
                    x[a] = xa;
                    const float diff = xa - mu;
                    logp += -(diff * diff) / (2.f * var) - logf(s) - kActLogSqrt2Pi;
                }
            }
            if (d.logp_out) d.logp_out[b] = logp;
        }
    }
    // the last block to finish advances the stream (every block has read seed / offset before taking a ticket)
    if (d.rng && !d.deterministic) {
        __syncthreads();
        if (threadIdx.x == 0) {
            const unsigned long long t = atomicAdd(d.rng + 2, 1ull);
            if (t == gridDim.x - 1) {
                d.rng[2] = 0ull;
                d.rng[1] = s_off + 1ull;
            }
        }
    }
}

int sample_actions(const void* desc, cudaStream_t st) {
    const SampleDesc& d = *reinterpret_cast<const SampleDesc*>(desc);
    A2CB_REQUIRE(d.z && d.action_out && d.B > 0 && d.A >= 1 && d.A <= kMaxActions && d.z_stride >= 1 + d.A,
                 "sample_actions: bad arguments");
    A2CB_REQUIRE(d.dist == 0 || d.logstd, "sample_actions: DiagGaussian needs logstd");
    A2CB_REQUIRE(d.deterministic || d.rng, "sample_actions: sampling needs the device stream state");
    sample_actions_kernel<<<(unsigned)((d.B + 127) / 128), 128, 0, st>>>(d);
    return check_launch("sample_actions");
}

}  // namespace a2cb
