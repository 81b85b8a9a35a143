// GAIL discriminator (reference a2c_ppo_acktr/algo/gail.py:11-110): trunk Linear(D, H) - Tanh - Linear(H, H) - Tanh -
// Linear(H, 1), exact fp32.
//   gail_disc_step  : ONE launch per minibatch pair = the three forward passes (policy rows gathered from the rollout
//                     buffer through the index list, expert rows, mixup rows alpha e + (1 - alpha) p), both
//                     BCE-with-logits terms, the gradient penalty lambda mean((||d D / d mixup||_2 - 1)^2) and ALL
//                     parameter gradients -- including the double-backward terms of the penalty, in closed form:
//                       g2 = w3 * s2, u = W2^T g2, g1 = u * s1, gx = W1^T g1            (s = 1 - tanh^2)
//                       gx^ = 2 lambda (n - 1) / (B n) gx,   dW1 += g1 gx^T,  dg1 = W1 gx^,  du = dg1 * s1,  ds1 = dg1 * u
//                       dW2 += g2 du^T,  dg2 = W2 du,  dw3 += sum dg2 * s2,  da2 = -2 h2 (dg2 * w3) * s2
//                       dW2 += da2 h1^T, db2 += da2,  dh1 = W2^T da2 - 2 h1 ds1,  da1 = dh1 * s1,  dW1 += da1 x^T, db1 += da1
//                     A CTA takes 16 rows of all three streams; per-CTA gradient slabs (+ the loss terms in the last
//                     slot) are folded by reduce_splits in a fixed order.  The reference runs ~60 autograd kernels and
//                     a host sync (loss.item()) per step.
//   gail_reward     : d -> r = log s - log(1 - s), s = sigmoid(d), for any number of rows at once (the per-step Python
//                     loop of reference main.py:152-155 batched over [T * N]).
//   gail_normalize  : the sequential part of predict_reward over T steps in one launch: returns = returns * masks * gamma
//                     + r, RunningMeanStd merge of the batch moments (stable_baselines3's parallel-variance update,
//                     float64), r / sqrt(var + 1e-8) with the variance AFTER that step's update.
#include "common.cuh"
#include "decls.cuh"

namespace a2cb {

struct GailDesc {
    const float* P;            // flat parameters [W1 (H x D) | b1 (H) | W2 (H x H) | b2 (H) | w3 (H) | b3 (1)]
    float* slabs;              // [n_cta][n_params + 2] per-CTA gradients, then {bce loss sum, penalty sum}
    const float* pol_state;    // policy stream: rows of the rollout buffer, [rows, Ds], picked through idx
    const float* pol_action;   //   [rows, Da]
    const int64_t* idx;        //   [B] row list (NULL: rows 0..B-1)
    const float* exp_state;    // expert stream [B, Ds] (already filtered), [B, Da]
    const float* exp_action;
    const float* alpha;        // [B] mixup coefficients
    float* raw_out;            // reward mode: [B] raw rewards
    int32_t B, Ds, Da, H;
    float lambda;
    int32_t pad_;
};

namespace {

constexpr int GL_R = 16;                   // rows per CTA
constexpr int GL_NT = 256;
constexpr int GL_MAX_H = 128, GL_MAX_D = 128;

// thread = (row r = tid & 15, output group tid >> 4): a warp covers 2 output columns x 16 rows, so weight loads are
// 2-address broadcasts and the activation walks are conflict-free with odd pitches
// out[r][j] = act(bias[j] + sum_k in[r][k] W[j][k])
__device__ __forceinline__ void gl_fwd(float* out, int ldo, const float* in, int ldi, const float* __restrict__ W, int ldw,
                                       const float* __restrict__ bias, int J, int K, bool tanh_act) {
    const int r = threadIdx.x & (GL_R - 1);
    for (int j = threadIdx.x >> 4; j < J; j += GL_NT / GL_R) {
        float acc = bias ? __ldg(bias + j) : 0.f;
        const float* w = W + (long)j * ldw;
        for (int k = 0; k < K; ++k) acc = fmaf(in[r * ldi + k], __ldg(w + k), acc);
        out[r * ldo + j] = tanh_act ? tanhf(acc) : acc;
    }
}
// out[r][k] = sum_j in[r][j] W[j][k]
__device__ __forceinline__ void gl_bwd(float* out, int ldo, const float* in, int ldi, const float* __restrict__ W, int ldw,
                                       int J, int K) {
    const int r = threadIdx.x & (GL_R - 1);
    for (int k = threadIdx.x >> 4; k < K; k += GL_NT / GL_R) {
        float acc = 0.f;
        for (int j = 0; j < J; ++j) acc = fmaf(in[r * ldi + j], __ldg(W + (long)j * ldw + k), acc);
        out[r * ldo + k] = acc;
    }
}
// dW[j][k] (+)= sum_r g[r][j] in[r][k];  db[j] (+)= sum_r g[r][j]   -- the CTA's own slab: the same thread owns the same
// element in every call, so accumulation needs no atomics
__device__ __forceinline__ void gl_wgrad(float* dW, int ldw, float* db, const float* g, int ldg, const float* in, int ldi,
                                         int J, int K, bool accumulate) {
    for (int e = threadIdx.x; e < J * K; e += GL_NT) {
        const int j = e / K, k = e - j * K;
        float acc = 0.f;
#pragma unroll
        for (int r = 0; r < GL_R; ++r) acc = fmaf(g[r * ldg + j], in[r * ldi + k], acc);
        dW[(long)j * ldw + k] = accumulate ? dW[(long)j * ldw + k] + acc : acc;
    }
    if (db)
        for (int j = threadIdx.x; j < J; j += GL_NT) {
            float acc = 0.f;
#pragma unroll
            for (int r = 0; r < GL_R; ++r) acc += g[r * ldg + j];
            db[j] = accumulate ? db[j] + acc : acc;
        }
}

struct GlOffsets { int w1, b1, w2, b2, w3, b3, n; };
__device__ __host__ inline GlOffsets gl_offsets(int D, int H) {
    GlOffsets o;
    o.w1 = 0; o.b1 = H * D; o.w2 = o.b1 + H; o.b2 = o.w2 + H * H; o.w3 = o.b2 + H; o.b3 = o.w3 + H; o.n = o.b3 + 1;
    return o;
}

// REWARD: forward only on `pol_*` rows (no index list needed, but honoured), writes raw rewards
template <bool REWARD>
__global__ void __launch_bounds__(GL_NT)
gail_kernel(const GailDesc d) {
    extern __shared__ __align__(16) float gs[];
    const int D = d.Ds + d.Da, H = d.H, DP = D | 1, HP = H | 1;           // odd pitches
    const GlOffsets o = gl_offsets(D, H);
    const float* W1 = d.P + o.w1; const float* b1 = d.P + o.b1; const float* W2 = d.P + o.w2;
    const float* b2 = d.P + o.b2; const float* w3 = d.P + o.w3; const float b3 = __ldg(d.P + o.b3);
    const int row0 = blockIdx.x * GL_R;
    const int tid = threadIdx.x;
    const int NS = REWARD ? 1 : 3;                                        // streams: policy | + expert + mixup
    float* X = gs;                                                        // [NS][16][DP]
    float* H1 = X + NS * GL_R * DP;                                       // [NS][16][HP]
    float* H2 = H1 + NS * GL_R * HP;
    float* dv = H2 + NS * GL_R * HP;                                      // [NS][16] logits, then d loss / d logit
    // ---- inputs ----
    for (int e = tid; e < GL_R * D; e += GL_NT) {
        const int r = e / D, k = e - r * D;
        const int b = row0 + r;
        float p = 0.f, x = 0.f, a = 0.f;
        if (b < d.B) {
            const long src = d.idx ? (long)d.idx[b] : (long)b;
            p = k < d.Ds ? d.pol_state[src * d.Ds + k] : d.pol_action[src * d.Da + (k - d.Ds)];
            if (!REWARD) {
                x = k < d.Ds ? d.exp_state[(long)b * d.Ds + k] : d.exp_action[(long)b * d.Da + (k - d.Ds)];
                a = d.alpha[b];
            }
        }
        X[r * DP + k] = p;
        if (!REWARD) {
            X[(GL_R + r) * DP + k] = x;
            X[(2 * GL_R + r) * DP + k] = a * x + (1.f - a) * p;           // reference gail.py:43
        }
    }
    __syncthreads();
    for (int s = 0; s < NS; ++s) gl_fwd(H1 + s * GL_R * HP, HP, X + s * GL_R * DP, DP, W1, D, b1, H, D, true);
    __syncthreads();
    for (int s = 0; s < NS; ++s) gl_fwd(H2 + s * GL_R * HP, HP, H1 + s * GL_R * HP, HP, W2, H, b2, H, H, true);
    __syncthreads();
    if (tid < NS * GL_R) {                                                // logits
        const float* h2 = H2 + tid * HP;
        float acc = b3;
        for (int k = 0; k < H; ++k) acc = fmaf(h2[k], __ldg(w3 + k), acc);
        dv[tid] = acc;
    }
    __syncthreads();
    if (REWARD) {
        if (tid < GL_R && row0 + tid < d.B) {
            const float s = 1.f / (1.f + expf(-dv[tid]));                 // torch.sigmoid
            d.raw_out[row0 + tid] = logf(s) - logf(1.f - s);              // reference gail.py:102-103
        }
        return;
    }
    // ================= training step =================
    float* slab = d.slabs + (long)blockIdx.x * (o.n + 2);
    float* T1 = dv + 3 * GL_R;                                            // six [16][HP] temporaries + one [16][DP]
    float* T2 = T1 + GL_R * HP; float* T3 = T2 + GL_R * HP; float* T4 = T3 + GL_R * HP; float* T5 = T4 + GL_R * HP;
    float* T6 = T5 + GL_R * HP; float* GX = T6 + GL_R * HP;
    float* red = GX + GL_R * DP;                                          // [16] row coefficients, [2] loss terms
    const float invB = 1.f / (float)d.B;
    for (int j = tid; j < H; j += GL_NT) { slab[o.b1 + j] = 0.f; slab[o.b2 + j] = 0.f; }   // first touched by accumulating calls
    const float* Xm = X + 2 * GL_R * DP; const float* H1m = H1 + 2 * GL_R * HP; const float* H2m = H2 + 2 * GL_R * HP;
    // ---- gradient penalty, first-order part: gx = d logit / d mixup ----
    for (int e = tid; e < GL_R * H; e += GL_NT) {
        const int r = e / H, k = e - r * H;
        const float h = H2m[r * HP + k];
        T1[r * HP + k] = __ldg(w3 + k) * (1.f - h * h);                   // g2
    }
    __syncthreads();
    gl_bwd(T2, HP, T1, HP, W2, H, H, H);                                  // u = g2 W2
    __syncthreads();
    for (int e = tid; e < GL_R * H; e += GL_NT) {
        const int r = e / H, k = e - r * H;
        const float h = H1m[r * HP + k];
        T3[r * HP + k] = T2[r * HP + k] * (1.f - h * h);                  // g1
    }
    __syncthreads();
    gl_bwd(GX, DP, T3, HP, W1, D, H, D);                                  // gx = g1 W1
    __syncthreads();
    if (tid < GL_R) {
        float ss = 0.f;
        for (int k = 0; k < D; ++k) ss = fmaf(GX[tid * DP + k], GX[tid * DP + k], ss);
        const float n = sqrtf(ss);
        const bool live = row0 + tid < d.B;
        red[tid] = (live && n > 0.f) ? d.lambda * 2.f * (n - 1.f) * invB / n : 0.f;
        dv[2 * GL_R + tid] = live ? (n - 1.f) * (n - 1.f) : 0.f;          // penalty term of this row
    }
    __syncthreads();
    for (int e = tid; e < GL_R * D; e += GL_NT) GX[(e / D) * DP + e % D] *= red[e / D];      // gx^
    __syncthreads();
    // ---- second-order part ----
    gl_wgrad(slab + o.w1, D, nullptr, T3, HP, GX, DP, H, D, false);       // dW1  = g1^T gx^
    gl_fwd(T4, HP, GX, DP, W1, D, nullptr, H, D, false);                  // dg1  = gx^ W1^T
    __syncthreads();
    for (int e = tid; e < GL_R * H; e += GL_NT) {
        const int r = e / H, k = e - r * H;
        const float h = H1m[r * HP + k], dg1 = T4[r * HP + k], u = T2[r * HP + k];
        T5[r * HP + k] = dg1 * (1.f - h * h);                             // du
        T2[r * HP + k] = dg1 * u;                                         // ds1
    }
    __syncthreads();
    gl_wgrad(slab + o.w2, H, nullptr, T1, HP, T5, HP, H, H, false);       // dW2  = g2^T du
    gl_fwd(T6, HP, T5, HP, W2, H, nullptr, H, H, false);                  // dg2  = du W2^T
    __syncthreads();
    for (int k = tid; k < H; k += GL_NT) {                                // dw3 = sum_r dg2 * s2
        float acc = 0.f;
        for (int r = 0; r < GL_R; ++r) { const float h = H2m[r * HP + k]; acc = fmaf(T6[r * HP + k], 1.f - h * h, acc); }
        slab[o.w3 + k] = acc;
    }
    __syncthreads();
    for (int e = tid; e < GL_R * H; e += GL_NT) {
        const int r = e / H, k = e - r * H;
        const float h = H2m[r * HP + k];
        T6[r * HP + k] = -2.f * h * (T6[r * HP + k] * __ldg(w3 + k)) * (1.f - h * h);      // da2
    }
    __syncthreads();
    gl_wgrad(slab + o.w2, H, slab + o.b2, T6, HP, H1m, HP, H, H, true);   // dW2 += da2^T h1, db2 = sum da2
    gl_bwd(T4, HP, T6, HP, W2, H, H, H);                                  // dh1 = da2 W2
    __syncthreads();
    for (int e = tid; e < GL_R * H; e += GL_NT) {
        const int r = e / H, k = e - r * H;
        const float h = H1m[r * HP + k];
        T4[r * HP + k] = (T4[r * HP + k] - 2.f * h * T2[r * HP + k]) * (1.f - h * h);       // da1
    }
    __syncthreads();
    gl_wgrad(slab + o.w1, D, slab + o.b1, T4, HP, Xm, DP, H, D, true);    // dW1 += da1^T x, db1 = sum da1
    __syncthreads();
    // ---- the two BCE-with-logits streams: policy (target 0), expert (target 1) ----
    if (tid < 2 * GL_R) {
        const int s = tid / GL_R, r = tid - s * GL_R;
        const float x = dv[tid], y = (float)s;
        const bool live = row0 + r < d.B;
        // max(x, 0) - x y + log(1 + exp(-|x|)): torch's binary_cross_entropy_with_logits
        T5[tid] = live ? fmaxf(x, 0.f) - x * y + log1pf(expf(-fabsf(x))) : 0.f;
        dv[tid] = live ? (1.f / (1.f + expf(-x)) - y) * invB : 0.f;       // d loss / d logit
    }
    __syncthreads();
    if (tid == 0) {
        float bce = 0.f, pen = 0.f;
        for (int i = 0; i < 2 * GL_R; ++i) bce += T5[i];
        for (int r = 0; r < GL_R; ++r) pen += dv[2 * GL_R + r];
        slab[o.n] = bce * invB;
        slab[o.n + 1] = d.lambda * pen * invB;
        slab[o.b3] = 0.f;
    }
    __syncthreads();
    for (int s = 0; s < 2; ++s) {
        const float* h1 = H1 + s * GL_R * HP; const float* h2 = H2 + s * GL_R * HP; const float* dd = dv + s * GL_R;
        for (int k = tid; k < H; k += GL_NT) {                            // dw3 += sum_r dd h2
            float acc = 0.f;
            for (int r = 0; r < GL_R; ++r) acc = fmaf(dd[r], h2[r * HP + k], acc);
            slab[o.w3 + k] += acc;
        }
        if (tid == 0) { float acc = 0.f; for (int r = 0; r < GL_R; ++r) acc += dd[r]; slab[o.b3] += acc; }
        for (int e = tid; e < GL_R * H; e += GL_NT) {
            const int r = e / H, k = e - r * H;
            const float h = h2[r * HP + k];
            T6[r * HP + k] = dd[r] * __ldg(w3 + k) * (1.f - h * h);       // da2
        }
        __syncthreads();
        gl_wgrad(slab + o.w2, H, slab + o.b2, T6, HP, h1, HP, H, H, true);
        gl_bwd(T4, HP, T6, HP, W2, H, H, H);
        __syncthreads();
        for (int e = tid; e < GL_R * H; e += GL_NT) {
            const int r = e / H, k = e - r * H;
            const float h = h1[r * HP + k];
            T4[r * HP + k] *= 1.f - h * h;                                // da1
        }
        __syncthreads();
        gl_wgrad(slab + o.w1, D, slab + o.b1, T4, HP, X + s * GL_R * DP, DP, H, D, true);
        __syncthreads();
    }
}

// T sequential steps of predict_reward's running statistics (reference gail.py:104-110), one CTA.
// state = {returns [N] float, flag}; rms = double {mean, var, count}
__global__ void __launch_bounds__(GL_NT)
gail_normalize_kernel(float* __restrict__ out, const float* __restrict__ raw, const float* __restrict__ masks, int T, int N,
                      float gamma, float* __restrict__ returns, int* __restrict__ returns_init, double* __restrict__ rms,
                      int update_rms) {
    __shared__ double red[GL_NT / 32 + 1];
    __shared__ double s_mean, s_var, s_count;
    const int tid = threadIdx.x;
    if (tid == 0) { s_mean = rms[0]; s_var = rms[1]; s_count = rms[2]; }
    const bool init = *returns_init != 0;
    __syncthreads();
    for (int t = 0; t < T; ++t) {
        const float* r_t = raw + (long)t * N;
        if (update_rms) {
            double sum = 0.0;
            for (int n = tid; n < N; n += GL_NT) {
                const float prev = (init || t > 0) ? returns[n] : r_t[n];          // `returns is None` -> reward.clone()
                const float v = prev * masks[(long)t * N + n] * gamma + r_t[n];
                returns[n] = v;
                sum += (double)v;
            }
            sum = block_sum(sum, red);
            const double bmean = sum / N;
            double sq = 0.0;
            for (int n = tid; n < N; n += GL_NT) { const double dlt = (double)returns[n] - bmean; sq += dlt * dlt; }
            sq = block_sum(sq, red);
            if (tid == 0) {                                                         // RunningMeanStd.update_from_moments
                const double bvar = sq / N, bcount = (double)N;
                const double delta = bmean - s_mean, tot = s_count + bcount;
                const double m2 = s_var * s_count + bvar * bcount + delta * delta * s_count * bcount / tot;
                s_mean = s_mean + delta * bcount / tot;
                s_var = m2 / tot;
                s_count = tot;
            }
            __syncthreads();
        }
        const float div = (float)sqrt(s_var + 1e-8);
        for (int n = tid; n < N; n += GL_NT) out[(long)t * N + n] = r_t[n] / div;
        __syncthreads();
    }
    if (tid == 0) {
        rms[0] = s_mean; rms[1] = s_var; rms[2] = s_count;
        if (update_rms || !init) *returns_init = 1;
    }
    if (!update_rms && !init)                                                       // returns = reward.clone() only
        for (int n = tid; n < N; n += GL_NT) returns[n] = raw[n];
}

size_t gl_smem(int D, int H, bool reward) {
    const int DP = D | 1, HP = H | 1, NS = reward ? 1 : 3;
    size_t f = (size_t)NS * GL_R * DP + 2 * (size_t)NS * GL_R * HP + NS * GL_R;
    if (!reward) f += 6 * (size_t)GL_R * HP + (size_t)GL_R * DP + GL_R + 2;
    return f * sizeof(float);
}

}  // namespace

int gail_slab_floats(int D, int H) { return gl_offsets(D, H).n + 2; }
int gail_ctas(int B) { return (B + GL_R - 1) / GL_R; }

int gail_disc(int reward, const void* desc, cudaStream_t st) {
    const GailDesc& d = *reinterpret_cast<const GailDesc*>(desc);
    const int D = d.Ds + d.Da;
    A2CB_REQUIRE(d.P && d.pol_state && d.pol_action && d.B > 0 && d.Ds > 0 && d.Da > 0, "gail: bad arguments");
    A2CB_REQUIRE(d.H >= 1 && d.H <= GL_MAX_H && D <= GL_MAX_D, "gail: hidden width <= %d and state + action width <= %d",
                 GL_MAX_H, GL_MAX_D);
    A2CB_REQUIRE(reward ? d.raw_out != nullptr : (d.slabs && d.exp_state && d.exp_action && d.alpha), "gail: null pointer");
    const size_t smem = gl_smem(D, d.H, reward != 0);
    static bool attr = false;
    if (!attr) {
        if (cudaFuncSetAttribute(gail_kernel<false>, cudaFuncAttributeMaxDynamicSharedMemorySize, 200 * 1024) != cudaSuccess ||
            cudaFuncSetAttribute(gail_kernel<true>, cudaFuncAttributeMaxDynamicSharedMemorySize, 200 * 1024) != cudaSuccess) {
            set_error("gail: shared memory attribute");
            cudaGetLastError();
            return A2CB_ERR_CUDA;
        }
        attr = true;
    }
    A2CB_REQUIRE(smem <= 200 * 1024, "gail: shared memory");
    if (reward) gail_kernel<true><<<(unsigned)gail_ctas(d.B), GL_NT, smem, st>>>(d);
    else gail_kernel<false><<<(unsigned)gail_ctas(d.B), GL_NT, smem, st>>>(d);
    return check_launch("gail_disc");
}

int gail_normalize(float* out, const float* raw, const float* masks, int T, int N, float gamma, float* returns,
                   int* returns_init, double* rms, int update_rms, cudaStream_t st) {
    A2CB_REQUIRE(out && raw && masks && returns && returns_init && rms && T > 0 && N > 0, "gail_normalize: bad arguments");
    gail_normalize_kernel<<<1, GL_NT, 0, st>>>(out, raw, masks, T, N, gamma, returns, returns_init, rms, update_rms);
    return check_launch("gail_normalize");
}

}  // namespace a2cb
