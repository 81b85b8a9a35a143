// MLPBase (reference a2c_ppo_acktr/model.py:198-229) fused: both 2-layer tanh trunks (actor, critic), critic_linear
// and the distribution head in ONE launch per direction.  The layers are 64 x 64: as separate GEMM launches each
// costs more in launch and set-up than in arithmetic (C1: 16 launches of ~13 us per 64-row minibatch step).  Here
// a CTA takes 64 rows through the whole network with every weight matrix (10 K parameters, 40 KiB) and every
// activation in shared memory; exact fp32 FMA, so it is also the reference-grade path.
//   forward : z[b] = (value, head outputs), saves the four post-tanh activations for the backward pass
//   backward: dz -> all weight / bias gradients (per-CTA slabs in the flat gradient layout when the batch
//             spans several CTAs, folded by reduce_splits) -- tanh' = 1 - h^2 from the saved activations.
#include "common.cuh"
#include "decls.cuh"
#include "mlp_fused.cuh"

namespace a2cb {

namespace {

constexpr int MR = 64;          // rows per CTA
constexpr int MH = 64;          // hidden width (the reference's MLPBase default, model.py:199)
constexpr int MLD = MH + 1;     // padded row pitch of activation tiles in shared memory (conflict-free column walks)
constexpr int MLH = MH + 4;     // 4-aligned pitch of the tiles the weight-gradient products read with 16-byte broadcasts
constexpr int MNT = 256;

// out[r][c] = act( bias[c] + sum_k in[r][k] * Wt[k][c] ),  r < 64, c < 64; thread = (16-column block, row): a warp is 32
// consecutive rows of ONE column block, so the weight loads are warp-wide broadcasts (one wavefront per LDS.128)
// and the input column walk is conflict-free with the odd row pitch
template <bool TANH>
__device__ __forceinline__ void mlp_layer(const float* __restrict__ in, int ld_in, int K, const float* __restrict__ Wt,
                                          const float* __restrict__ bias, float* __restrict__ out, int tid) {
    const int r = tid & 63, cb = (tid >> 6) * 16;
    float acc[16];
#pragma unroll
    for (int j = 0; j < 16; ++j) acc[j] = bias[cb + j];
    for (int k = 0; k < K; ++k) {
        const float a = in[r * ld_in + k];
        const float4* w = reinterpret_cast<const float4*>(Wt + k * MH + cb);
#pragma unroll
        for (int q = 0; q < 4; ++q) {
            const float4 wv = w[q];
            acc[4 * q + 0] = fmaf(a, wv.x, acc[4 * q + 0]); acc[4 * q + 1] = fmaf(a, wv.y, acc[4 * q + 1]);
            acc[4 * q + 2] = fmaf(a, wv.z, acc[4 * q + 2]); acc[4 * q + 3] = fmaf(a, wv.w, acc[4 * q + 3]);
        }
    }
#pragma unroll
    for (int j = 0; j < 16; ++j) out[r * MLD + cb + j] = TANH ? tanhf(acc[j]) : acc[j];
}

// Wt[k][c] = W[c][k]  (W is [64][K] row-major in the flat parameter buffer)
__device__ __forceinline__ void mlp_load_wt(float* Wt, const float* __restrict__ W, int K, int tid) {
    // consecutive threads take consecutive OUTPUT units c: the shared-memory stores are conflict-free (the global
    // reads are strided by K floats and served by L1 / L2)
    for (int e = tid; e < MH * K; e += MNT) {
        const int k = e / MH, c = e - k * MH;
        Wt[e] = __ldg(W + (long)c * K + k);
    }
}

// 512 threads: threads 0..255 run the actor trunk, 256..511 the critic trunk, each on its own shared-memory tiles;
// the observation tile is shared.
__global__ void __launch_bounds__(2 * MNT)
mlp_fwd_kernel(const MlpDesc d) {
    extern __shared__ __align__(16) float ms[];
    const int trunk = threadIdx.x >> 8, tid = threadIdx.x & 255;
    const int K0 = d.d_in, ldx = K0 + 1;
    const int per_trunk = K0 * MH + MH * MH + 2 * MH + 2 * MR * MLD;
    float* xs = ms;                          // [64][K0 + 1] (both trunks)
    float* base = ms + ((MR * ldx + 3) & ~3) + trunk * per_trunk;
    float* Wt0 = base;                       // [K0][64]
    float* Wt2 = Wt0 + K0 * MH;              // [64][64]
    float* b0 = Wt2 + MH * MH;               // [64]
    float* b2 = b0 + MH;                     // [64]
    float* h1 = b2 + MH;                     // [64][65]
    float* h2 = h1 + MR * MLD;               // [64][65]
    const int row0 = blockIdx.x * MR;
    const int nrows = min(MR, d.B - row0);
    for (int e = threadIdx.x; e < MR * K0; e += 2 * MNT) {
        const int r = e / K0, k = e - r * K0;
        float v = 0.f;
        if (r < nrows) {
            const long src = d.idx ? (long)d.idx[row0 + r] : (long)(row0 + r);
            v = d.obs[src * K0 + k];
        }
        xs[r * ldx + k] = v;
    }
    // trunk 0 = actor, 1 = critic
    const long w0 = trunk ? d.off_c0w : d.off_a0w, bb0 = trunk ? d.off_c0b : d.off_a0b;
    const long w2 = trunk ? d.off_c2w : d.off_a2w, bb2 = trunk ? d.off_c2b : d.off_a2b;
    mlp_load_wt(Wt0, d.P + w0, K0, tid);
    mlp_load_wt(Wt2, d.P + w2, MH, tid);
    if (tid < MH) { b0[tid] = d.P[bb0 + tid]; b2[tid] = d.P[bb2 + tid]; }
    __syncthreads();
    mlp_layer<true>(xs, ldx, K0, Wt0, b0, h1, tid);
    __syncthreads();
    mlp_layer<true>(h1, MLD, MH, Wt2, b2, h2, tid);
    __syncthreads();
    // saved activations: acts[(2 * trunk + layer)][B][64]
    float* a1 = d.acts + (long)(2 * trunk) * d.B * MH;
    float* a2 = a1 + (long)d.B * MH;
    for (int e = tid; e < nrows * MH; e += MNT) {
        const int r = e / MH, c = e - r * MH;
        a1[(long)(row0 + r) * MH + c] = h1[r * MLD + c];
        a2[(long)(row0 + r) * MH + c] = h2[r * MLD + c];
    }
    // heads: row 0 of the head matrix = critic_linear (reads the critic trunk), rows 1.. = distribution head
    const int o0 = trunk ? 0 : 1, o1 = trunk ? 1 : d.zs;
    for (int e = tid; e < nrows * (o1 - o0); e += MNT) {
        const int r = e / (o1 - o0), o = o0 + e % (o1 - o0);
        const float* w = d.P + d.off_hw + (long)o * MH;
        float s = d.P[d.off_hb + o];
        for (int c = 0; c < MH; ++c) s = fmaf(h2[r * MLD + c], __ldg(w + c), s);
        d.z[(long)(row0 + r) * d.zs + o] = s;
    }
}

// dW[c][k] = sum_r g[r][c] * in[r][k], db[c] = sum_r g[r][c]   (c < 64, k < K); thread = (quarter of the k range, c):
// a warp is 32 consecutive c of one k quarter -> g is a conflict-free row read, `in` a warp-wide broadcast
// (16-byte loads when the pitch and the quarter are 4-aligned)
__device__ __forceinline__ void mlp_wgrad(const float* __restrict__ g, const float* __restrict__ in, int ld_in, int K,
                                          float* __restrict__ dW, float* __restrict__ db, int tid) {
    const int c = tid & 63, q = tid >> 6;
    const int kq = (K + 3) / 4, k0 = q * kq, k1 = min(K, k0 + kq);
    const bool vec = (ld_in % 4 == 0) && (kq % 16 == 0);
    for (int kb = k0; kb < k1; kb += 16) {
        float acc[16];
#pragma unroll
        for (int j = 0; j < 16; ++j) acc[j] = 0.f;
        const int n = min(16, k1 - kb);
        if (vec) {
            for (int r = 0; r < MR; ++r) {
                const float gv = g[r * MLD + c];
                const float4* iv = reinterpret_cast<const float4*>(in + r * ld_in + kb);
#pragma unroll
                for (int qq = 0; qq < 4; ++qq) {
                    const float4 x = iv[qq];
                    acc[4 * qq + 0] = fmaf(gv, x.x, acc[4 * qq + 0]); acc[4 * qq + 1] = fmaf(gv, x.y, acc[4 * qq + 1]);
                    acc[4 * qq + 2] = fmaf(gv, x.z, acc[4 * qq + 2]); acc[4 * qq + 3] = fmaf(gv, x.w, acc[4 * qq + 3]);
                }
            }
        } else {
            for (int r = 0; r < MR; ++r) {
                const float gv = g[r * MLD + c];
#pragma unroll
                for (int j = 0; j < 16; ++j)
                    if (j < n) acc[j] = fmaf(gv, in[r * ld_in + kb + j], acc[j]);
            }
        }
        for (int j = 0; j < n; ++j) dW[(long)c * K + kb + j] = acc[j];
    }
    if (q == 0) {
        float s = 0.f;
        for (int r = 0; r < MR; ++r) s += g[r * MLD + c];
        db[c] = s;
    }
}

__global__ void __launch_bounds__(2 * MNT)
mlp_bwd_kernel(const MlpDesc d) {
    extern __shared__ __align__(16) float ms[];
    const int trunk = threadIdx.x >> 8, tid = threadIdx.x & 255;
    const int K0 = d.d_in, ldx = K0 + 1;
    const int per_trunk = MH * MH + 2 * MR * MLH + 2 * MR * MLD;
    float* xs = ms;                          // [64][K0 + 1]
    float* dzs = ms + ((MR * ldx + 3) & ~3); // [64][8]
    float* base = dzs + MR * 8 + trunk * per_trunk;
    float* W2 = base;                        // [64][64] as stored (out, in)
    float* h1 = W2 + MH * MH;                // [64][68]
    float* h2 = h1 + MR * MLH;               // [64][68]
    float* g2 = h2 + MR * MLH;               // [64][65]
    float* g1 = g2 + MR * MLD;               // [64][65]
    const int row0 = blockIdx.x * MR;
    const int nrows = min(MR, d.B - row0);
    float* G = d.partial ? d.partial + (long)blockIdx.x * d.span : d.G;
    for (int e = threadIdx.x; e < MR * K0; e += 2 * MNT) {
        const int r = e / K0, k = e - r * K0;
        float v = 0.f;
        if (r < nrows) {
            const long src = d.idx ? (long)d.idx[row0 + r] : (long)(row0 + r);
            v = d.obs[src * K0 + k];
        }
        xs[r * ldx + k] = v;
    }
    for (int e = threadIdx.x; e < MR * 8; e += 2 * MNT) {
        const int r = e >> 3, o = e & 7;
        dzs[e] = (r < nrows && o < d.zs) ? d.dz[(long)(row0 + r) * d.zs + o] : 0.f;
    }
    const long w0 = trunk ? d.off_c0w : d.off_a0w, bb0 = trunk ? d.off_c0b : d.off_a0b;
    const long w2 = trunk ? d.off_c2w : d.off_a2w, bb2 = trunk ? d.off_c2b : d.off_a2b;
    const float* a1 = d.acts + (long)(2 * trunk) * d.B * MH;
    const float* a2 = a1 + (long)d.B * MH;
    for (int e = tid; e < MR * MH; e += MNT) {
        const int r = e / MH, c = e - r * MH;
        const bool ok = r < nrows;
        h1[r * MLH + c] = ok ? a1[(long)(row0 + r) * MH + c] : 0.f;
        h2[r * MLH + c] = ok ? a2[(long)(row0 + r) * MH + c] : 0.f;
    }
    for (int e = tid; e < MH * MH; e += MNT) W2[e] = d.P[w2 + e];
    __syncthreads();
    const int o0 = trunk ? 0 : 1, o1 = trunk ? 1 : d.zs;
    // head weight / bias gradients of this trunk's outputs
    for (int e = tid; e < (o1 - o0) * MH; e += MNT) {
        const int o = o0 + e / MH, c = e % MH;
        float s = 0.f;
        for (int r = 0; r < MR; ++r) s = fmaf(dzs[r * 8 + o], h2[r * MLH + c], s);
        G[d.off_hw + (long)o * MH + c] = s;
    }
    if (tid >= o0 && tid < o1) {
        float s = 0.f;
        for (int r = 0; r < MR; ++r) s += dzs[r * 8 + tid];
        G[d.off_hb + tid] = s;
    }
    // g2 = (dz Wh) * (1 - h2^2)
    for (int e = tid; e < MR * MH; e += MNT) {
        const int r = e / MH, c = e - r * MH;
        float s = 0.f;
        for (int o = o0; o < o1; ++o) s = fmaf(dzs[r * 8 + o], __ldg(d.P + d.off_hw + (long)o * MH + c), s);
        const float h = h2[r * MLH + c];
        g2[r * MLD + c] = s * (1.f - h * h);
    }
    __syncthreads();
    mlp_wgrad(g2, h1, MLH, MH, G + w2, G + bb2, tid);
    // g1 = (g2 W2) * (1 - h1^2): thread = (16-column block of the layer-1 width, row)
    {
        const int r = tid & 63, kb = (tid >> 6) * 16;
        float acc[16];
#pragma unroll
        for (int j = 0; j < 16; ++j) acc[j] = 0.f;
        for (int c = 0; c < MH; ++c) {
            const float gv = g2[r * MLD + c];
            const float4* w = reinterpret_cast<const float4*>(W2 + c * MH + kb);
#pragma unroll
            for (int q = 0; q < 4; ++q) {
                const float4 wv = w[q];
                acc[4 * q + 0] = fmaf(gv, wv.x, acc[4 * q + 0]); acc[4 * q + 1] = fmaf(gv, wv.y, acc[4 * q + 1]);
                acc[4 * q + 2] = fmaf(gv, wv.z, acc[4 * q + 2]); acc[4 * q + 3] = fmaf(gv, wv.w, acc[4 * q + 3]);
            }
        }
#pragma unroll
        for (int j = 0; j < 16; ++j) {
            const float h = h1[r * MLH + kb + j];
            g1[r * MLD + kb + j] = acc[j] * (1.f - h * h);
        }
    }
    __syncthreads();
    mlp_wgrad(g1, xs, ldx, K0, G + w0, G + bb0, tid);
}

size_t fwd_smem(int d_in) {
    return (size_t)(((MR * (d_in + 1) + 3) & ~3) + 2 * (d_in * MH + MH * MH + 2 * MH + 2 * MR * MLD)) * sizeof(float);
}
size_t bwd_smem(int d_in) {
    return (size_t)(((MR * (d_in + 1) + 3) & ~3) + MR * 8 + 2 * (MH * MH + 2 * MR * MLH + 2 * MR * MLD)) * sizeof(float);
}

}  // namespace

bool mlp_fused_supported(int d_in, int H, int zs) { return H == MH && d_in >= 1 && d_in <= 128 && zs >= 2 && zs <= 8; }

int mlp_fused(int backward, const MlpDesc& d, cudaStream_t st) {
    A2CB_REQUIRE(mlp_fused_supported(d.d_in, d.H, d.zs), "mlp_fused: unsupported shape (d_in=%d H=%d zs=%d)", d.d_in, d.H, d.zs);
    A2CB_REQUIRE(d.B > 0 && d.P && d.obs && d.acts, "mlp_fused: null pointer or empty batch");
    const int ncta = (d.B + MR - 1) / MR;
    if (!backward) {
        A2CB_REQUIRE(d.z, "mlp_fused: z");
        const size_t smem = fwd_smem(d.d_in);
        static size_t cap = 0;
        if (smem > cap) {
            if (cudaFuncSetAttribute(mlp_fwd_kernel, cudaFuncAttributeMaxDynamicSharedMemorySize, (int)smem) != cudaSuccess) {
                set_error("mlp_fused: shared memory"); cudaGetLastError(); return A2CB_ERR_CUDA;
            }
            cap = smem;
        }
        mlp_fwd_kernel<<<ncta, 2 * MNT, smem, st>>>(d);
        return check_launch("mlp_fwd");
    }
    A2CB_REQUIRE(d.dz && d.G && (ncta == 1 || d.partial), "mlp_fused: backward pointers (partial slabs needed for B > %d)", MR);
    MlpDesc dd = d;
    if (ncta == 1) dd.partial = nullptr;
    const size_t smem = bwd_smem(d.d_in);
    static size_t capb = 0;
    if (smem > capb) {
        if (cudaFuncSetAttribute(mlp_bwd_kernel, cudaFuncAttributeMaxDynamicSharedMemorySize, (int)smem) != cudaSuccess) {
            set_error("mlp_fused: shared memory"); cudaGetLastError(); return A2CB_ERR_CUDA;
        }
        capb = smem;
    }
    mlp_bwd_kernel<<<ncta, 2 * MNT, smem, st>>>(dd);
    return check_launch("mlp_bwd");
}

}  // namespace a2cb
