// Fused loss epilogue: action-distribution log-prob + entropy, PPO clipped surrogate and
// clipped value loss (or the A2C / Fisher-sample losses), their analytic gradients with
// respect to the head outputs, and the advantage statistics -- one kernel per minibatch
// instead of ~40 elementwise/reduction launches.
//
// Replaces reference algo/ppo.py:35-37,61-77,86-88, algo/a2c_acktr.py:48-51,55-64,
// distributions.py:18-44 (FixedCategorical / FixedNormal) and model.py:76-77.
//
// One thread per row: the head outputs z[b] = (value, logits[A] | mean[A]) are read from
// the GEMM engine's output, the per-row rollout fields are read straight from the rollout
// buffer through the minibatch index list (the narrow-field gather of storage.py:131-140 is
// fused here), gradients dz are written for the backward GEMMs, and the three loss terms
// are reduced with warp shuffles -> per-block partials -> a fixed-order final sum by the
// last block (deterministic; no float atomics).  Loss means are accumulated on the device,
// so update() needs a single device->host copy instead of three .item() syncs per minibatch.
#include "common.cuh"
#include "decls.cuh"
#include "loss.cuh"

namespace a2cb {

constexpr int kLossThreads = 128;
constexpr float kLogSqrt2Pi = 0.918938533204672742f;   // log(sqrt(2*pi))
constexpr float kHalfLog2PiE = 1.418938533204672742f;  // 0.5 + 0.5*log(2*pi)

// scratch layout: [grid][4 + A] floats of partials, then one unsigned counter.
__global__ void __launch_bounds__(kLossThreads)
loss_kernel(const LossDesc d) {
    __shared__ float red[4 + kMaxActions][kLossThreads / 32];
    __shared__ bool is_last;
    const int b = blockIdx.x * blockDim.x + threadIdx.x;
    const int A = d.A;
    float vloss = 0.f, aloss = 0.f, ent = 0.f;
    float dls[kMaxActions];
#pragma unroll
    for (int a = 0; a < kMaxActions; ++a) dls[a] = 0.f;

    // padding rows of a sharded minibatch: rows [n_valid, B), or -- time-major recurrent batches padded
    // per environment -- rows whose environment slot (b % valid_mod) is >= n_valid
    const bool row_valid = d.n_valid < 0 ? true : (d.valid_mod > 0 ? (b % d.valid_mod) < d.n_valid : b < d.n_valid);
    if (b < d.B && !row_valid && d.dz) {
        // padding row of a sharded minibatch: contributes nothing
        float* dz = d.dz + (long)b * d.dz_stride;
        for (int j = 0; j <= A; ++j) dz[j] = 0.f;
        if (d.dlogstd_rows)
            for (int a = 0; a < A; ++a) d.dlogstd_rows[(long)b * A + a] = 0.f;
    }
    if (b < d.B && row_valid) {
        const long src = d.idx ? (long)d.idx[b] : (long)b;
        const float* z = d.z + (long)b * d.z_stride;
        const float v = z[0];
        float logp, H;
        float g_logp = 0.f, g_v = 0.f;

        // ---- per-row losses (mode) -> g_logp, g_v ----
        // (needs logp first)
        if (d.dist == 0) {
            // Categorical: l = z - logsumexp(z); logp = l[a]; H = -sum p l
            float mx = -INFINITY;
            for (int j = 0; j < A; ++j) mx = fmaxf(mx, z[1 + j]);
            float se = 0.f;
            for (int j = 0; j < A; ++j) se += expf(z[1 + j] - mx);
            const float lse = mx + logf(se);
            const int act = (int)reinterpret_cast<const int64_t*>(d.actions)[src * d.act_stride];
            logp = z[1 + act] - lse;
            H = 0.f;
            for (int j = 0; j < A; ++j) { const float l = z[1 + j] - lse; H -= expf(l) * l; }
        } else {
            // DiagGaussian: logp = sum_a -(x-mu)^2/(2 s^2) - log s - log sqrt(2 pi)
            const float* x = reinterpret_cast<const float*>(d.actions) + src * d.act_stride;
            logp = 0.f; H = 0.f;
            for (int a = 0; a < A; ++a) {
                const float ls = d.logstd[a], s = expf(ls), var = s * s;
                const float diff = x[a] - z[1 + a];
                logp += -(diff * diff) / (2.f * var) - logf(s) - kLogSqrt2Pi;
                H += kHalfLog2PiE + logf(s);
            }
        }
        if (d.logp_out) d.logp_out[b] = logp;
        if (d.ent_out) d.ent_out[b] = H;
        ent = H;

        const float invB = d.inv_B;
        if (d.mode == 0) {                      // PPO  (ppo.py:61-77)
            const float adv = d.adv[src];
            const float ratio = expf(logp - d.old_logp[src]);
            const float s1 = ratio * adv;
            const float rc = fminf(fmaxf(ratio, 1.f - d.clip), 1.f + d.clip);
            const float s2 = rc * adv;
            aloss = -fminf(s1, s2);
            const bool inside = (ratio >= 1.f - d.clip) && (ratio <= 1.f + d.clip);
            // d(-min(s1,s2))/dlogp: through s1 when it is the min (or tied), through s2 only inside the clamp
            float g = 0.f;
            if (s1 < s2) g = s1;                                   // ratio * adv
            else if (s1 == s2) g = 0.5f * s1 + (inside ? 0.5f * s1 : 0.f);
            else g = inside ? s1 : 0.f;                            // s2 is the min; d s2/d ratio = adv inside
            g_logp = -g * invB;
            const float R = d.returns[src];
            if (d.use_clipped_value_loss) {
                const float vo = d.old_value[src];
                const float dv = v - vo;
                const float vc = vo + fminf(fmaxf(dv, -d.clip), d.clip);
                const float l1 = (v - R) * (v - R), l2 = (vc - R) * (vc - R);
                vloss = 0.5f * fmaxf(l1, l2);
                const bool in_v = (dv >= -d.clip) && (dv <= d.clip);
                float gv;
                if (l1 > l2) gv = (v - R);
                else if (l1 < l2) gv = in_v ? (vc - R) : 0.f;
                else gv = 0.5f * (v - R) + (in_v ? 0.5f * (vc - R) : 0.f);
                g_v = d.c_v * gv * invB;
            } else {
                vloss = 0.5f * (R - v) * (R - v);
                g_v = d.c_v * (v - R) * invB;
            }
        } else if (d.mode == 1) {               // A2C / ACKTR main loss  (a2c_acktr.py:48-51)
            const float adv = d.returns[src] - v;
            vloss = adv * adv;
            aloss = -adv * logp;
            g_v = -2.f * d.c_v * adv * invB;
            g_logp = -adv * invB;
        } else if (d.mode == 2) {               // Fisher sampling loss  (a2c_acktr.py:55-64)
            const float nz = d.noise[src];       // sample = v + noise (detached)
            vloss = -(nz * nz);                  // -(v - sample)^2
            aloss = -logp;
            g_logp = -invB;
            g_v = 2.f * nz * invB;               // d/dv [-(v - s)^2] = -2 (v - s) = 2 noise
        } else if (d.mode == 4) {               // VJP of evaluate_actions for caller-supplied upstream grads
            g_v = d.old_value[b];                // dL/dvalue[b]
            g_logp = d.old_logp[b];              // dL/dlogp[b]
        }

        // ---- gradients w.r.t. head outputs ----
        if (d.dz) {
            float* dz = d.dz + (long)b * d.dz_stride;
            dz[0] = g_v;
            // total = ... - c_e * mean(H); mode 4: upstream dL/d(mean H) is read from noise[0]
            const float ce = (d.mode == 2) ? 0.f : (d.mode == 4 ? -d.noise[0] * invB : d.c_e * invB);
            if (d.dist == 0) {
                float mx = -INFINITY;
                for (int j = 0; j < A; ++j) mx = fmaxf(mx, z[1 + j]);
                float se = 0.f;
                for (int j = 0; j < A; ++j) se += expf(z[1 + j] - mx);
                const float lse = mx + logf(se);
                const int act = (int)reinterpret_cast<const int64_t*>(d.actions)[src * d.act_stride];
                for (int j = 0; j < A; ++j) {
                    const float l = z[1 + j] - lse, p = expf(l);
                    dz[1 + j] = g_logp * ((j == act ? 1.f : 0.f) - p) + ce * p * (l + H);
                }
            } else {
                const float* x = reinterpret_cast<const float*>(d.actions) + src * d.act_stride;
                for (int a = 0; a < A; ++a) {
                    const float s = expf(d.logstd[a]), var = s * s;
                    const float diff = x[a] - z[1 + a];
                    dz[1 + a] = g_logp * diff / var;
                    dls[a] = g_logp * (diff * diff / var - 1.f) - ce;   // entropy: -c_e per action dim per row mean
                    if (d.dlogstd_rows) d.dlogstd_rows[(long)b * A + a] = dls[a];
                }
            }
        }
    }

    // ---- block reduction ----
    const int lane = threadIdx.x & 31, w = threadIdx.x >> 5;
    vloss = warp_sum(vloss); aloss = warp_sum(aloss); ent = warp_sum(ent);
    if (lane == 0) { red[0][w] = vloss; red[1][w] = aloss; red[2][w] = ent; }
    if (d.dist == 1) {
        for (int a = 0; a < A; ++a) {
            const float s = warp_sum(dls[a]);
            if (lane == 0) red[4 + a][w] = s;
        }
    }
    __syncthreads();
    const int nslots = 4 + ((d.dist == 1) ? A : 0);
    float* part = d.scratch + (long)blockIdx.x * (4 + kMaxActions);
    if (threadIdx.x < nslots && threadIdx.x != 3) {
        float s = 0.f;
        for (int k = 0; k < kLossThreads / 32; ++k) s += red[threadIdx.x][k];
        part[threadIdx.x] = s;
    }
    __threadfence();
    __syncthreads();
    unsigned* counter = reinterpret_cast<unsigned*>(d.scratch + (long)gridDim.x * (4 + kMaxActions));
    if (threadIdx.x == 0) {
        const unsigned t = atomicAdd(counter, 1u);
        is_last = (t == gridDim.x - 1);
    }
    __syncthreads();
    if (!is_last) return;
    __threadfence();
    if (threadIdx.x < nslots && threadIdx.x != 3) {
        float s = 0.f;
        for (unsigned k = 0; k < gridDim.x; ++k) s += d.scratch[(long)k * (4 + kMaxActions) + threadIdx.x];
        if (threadIdx.x < 3) {
            // minibatch means; the entropy bonus is mean(H) over the (global) minibatch
            const float m = s * d.inv_B;
            if (d.loss_out) d.loss_out[threadIdx.x] = m;
            if (d.loss_accum) d.loss_accum[threadIdx.x] += m;
        } else if (d.dlogstd) {
            const int a = threadIdx.x - 4;
            d.dlogstd[(long)a * d.dlogstd_stride] = s;
        }
    }
    if (threadIdx.x == 0) *counter = 0;   // re-arm for the next launch
}

int loss_epilogue(const LossDesc& d, cudaStream_t st) {
    A2CB_REQUIRE(d.z && d.actions && d.scratch, "loss: null pointer");
    A2CB_REQUIRE(d.A >= 1 && d.A <= kMaxActions, "loss: 1 <= A <= %d", kMaxActions);
    A2CB_REQUIRE(d.dist == 0 || d.logstd, "loss: logstd required for DiagGaussian");
    A2CB_REQUIRE(d.mode >= 0 && d.mode <= 4, "loss: mode");
    A2CB_REQUIRE(d.mode != 4 || (d.old_value && d.old_logp && d.noise && d.dz), "loss: VJP mode needs upstream grads");
    A2CB_REQUIRE(d.mode != 0 || (d.adv && d.old_logp && d.returns && (!d.use_clipped_value_loss || d.old_value)),
                 "loss: PPO needs adv / old_logp / returns / old_value");
    A2CB_REQUIRE(d.mode != 1 || d.returns, "loss: A2C needs returns");
    A2CB_REQUIRE(d.mode != 2 || d.noise, "loss: Fisher mode needs noise");
    if (d.B == 0) return A2CB_OK;
    const int grid = (d.B + kLossThreads - 1) / kLossThreads;
    loss_kernel<<<grid, kLossThreads, 0, st>>>(d);
    return check_launch("loss_epilogue");
}

long loss_scratch_floats(int B) {
    const int grid = (B + kLossThreads - 1) / kLossThreads;
    return (long)grid * (4 + kMaxActions) + 4;
}

// ---------------------------------------------------------------------------
// advantage statistics + normalisation (ppo.py:35-37): mean and UNBIASED std over all T*N
// ---------------------------------------------------------------------------
__global__ void __launch_bounds__(256)
adv_moments_kernel(const float* __restrict__ returns, const float* __restrict__ values, long n,
                   double* __restrict__ partials, double* __restrict__ moments, unsigned* counter) {
    __shared__ double red[32];
    __shared__ bool is_last;
    double s = 0.0, ss = 0.0;
    for (long e = (long)blockIdx.x * blockDim.x + threadIdx.x; e < n; e += (long)gridDim.x * blockDim.x) {
        const double a = (double)(returns[e] - values[e]);     // fp32 subtraction, as the reference
        s += a; ss += a * a;
    }
    s = block_sum(s, red);
    ss = block_sum(ss, red);
    if (threadIdx.x == 0) {
        partials[2 * blockIdx.x] = s; partials[2 * blockIdx.x + 1] = ss;
        __threadfence();
        is_last = (atomicAdd(counter, 1u) == gridDim.x - 1);
    }
    __syncthreads();
    if (!is_last) return;
    __threadfence();
    if (threadIdx.x == 0) {
        double S = 0.0, SS = 0.0;
        for (unsigned k = 0; k < gridDim.x; ++k) { S += partials[2 * k]; SS += partials[2 * k + 1]; }
        moments[0] = S; moments[1] = SS; moments[2] = (double)n;
        *counter = 0;
    }
}

// moments = (sum, sum of squares, count) -- possibly all-reduced over ranks by the caller.
__global__ void __launch_bounds__(256)
adv_normalize_kernel(const float* __restrict__ returns, const float* __restrict__ values, long n,
                     const double* __restrict__ moments, float* __restrict__ adv) {
    const double cnt = moments[2];
    const double mean = moments[0] / cnt;
    double var = (moments[1] - cnt * mean * mean) / (cnt - 1.0);
    if (var < 0.0) var = 0.0;
    const float fmean = (float)mean, denom = (float)sqrt(var) + 1e-5f;
    const long e = (long)blockIdx.x * blockDim.x + threadIdx.x;
    if (e < n) adv[e] = ((returns[e] - values[e]) - fmean) / denom;
}

int adv_moments(const float* returns, const float* values, long n, double* scratch, double* moments,
                cudaStream_t st) {
    A2CB_REQUIRE(returns && values && scratch && moments && n >= 0, "adv_moments: bad argument");
    int grid = (int)((n + 255) / 256);
    if (grid > 128) grid = 128;
    if (grid < 1) grid = 1;
    unsigned* counter = reinterpret_cast<unsigned*>(scratch + 2 * 128);
    adv_moments_kernel<<<grid, 256, 0, st>>>(returns, values, n, scratch, moments, counter);
    return check_launch("adv_moments");
}

int adv_normalize(const float* returns, const float* values, long n, const double* moments, float* adv,
                  cudaStream_t st) {
    A2CB_REQUIRE(returns && values && moments && adv && n >= 0, "adv_normalize: bad argument");
    if (n == 0) return A2CB_OK;
    adv_normalize_kernel<<<(unsigned)((n + 255) / 256), 256, 0, st>>>(returns, values, n, moments, adv);
    return check_launch("adv_normalize");
}

}  // namespace a2cb
