// Affine-indexed GEMM engine (fp32 SIMT):
//
//     C[rowC(i) + colC(j)] = epi( alpha * sum_r  A[rowA(i) + redA(r)] * B[redB(r) + colB(j)] )
//
// Every index map is a three-digit affine map (Affine3), so im2col is never
// materialised: a convolution's patch matrix, its transposed-convolution (dgrad)
// gather over a zero-haloed gradient buffer, the weight-gradient reduction over
// (image, y, x), a Linear layer in the reference's [out, in] weight layout and the
// KFAC factor products are all instances of this one kernel with different maps
// (built and CPU-tested on the host side, a2c_ppo_acktr/_plans.py).  Rows of A may
// additionally be looked up through a minibatch index list (gather_row / gather_red),
// which fuses the shuffle-gather of observations -- and the uint8 -> fp32/255
// normalisation of Atari frames (reference model.py:190) -- into the first conv's load.
//
// This is the exact-fp32 engine: every layer shape runs on it and it is the numerical
// reference for the tcgen05 path.  256 threads, BMxBNx16 tiles, register micro-tiles,
// double-buffered shared memory with one barrier per k-tile; offsets are evaluated once
// per CTA (rows, columns) or once per k-tile by 2x16 threads (reduction), never per FMA.
#include "common.cuh"
#include "decls.cuh"
#include "gemm.cuh"
#include <stdlib.h>

namespace a2cb {

constexpr int kNT = 256;

__device__ __forceinline__ long affine_eval(const Affine3& a, long i, const int64_t* gather) {
    const unsigned ii = (unsigned)i, d1 = (unsigned)a.d1, d2 = (unsigned)a.d2;
    unsigned q0 = ii / d1;
    const unsigned rem = ii - q0 * d1;
    const unsigned q1 = rem / d2;
    const unsigned q2 = rem - q1 * d2;
    long g0 = gather ? (long)gather[q0] : (long)q0;
    return g0 * a.s0 + (long)q1 * a.s1 + (long)q2 * a.s2;
}

// ADT: 0 fp32 as is, 1 uint8 / 255, 2 fp32 / 255 (frames kept as fp32 0..255, reference layout).
// ADT 2 uses true IEEE division, so the operand equals torch's `inputs / 255.0` bit for bit
// (model.py:190).  ADT 1 keeps the byte as an exact integer-valued float and folds the 1/255 into
// the epilogue scale: sum(v w) / 255 instead of sum(fl(v / 255) w) -- equal to <= 1 ulp per
// product, no division per element, and exact in TF32 for the tensor-core engine.
template <int ADT>
__device__ __forceinline__ float load_a1(const void* A, long off) {
    if (ADT == 1) return (float)__ldg(reinterpret_cast<const uint8_t*>(A) + off);
    const float v = __ldg(reinterpret_cast<const float*>(A) + off);
    return ADT == 2 ? v / 255.0f : v;
}
template <int ADT>
__device__ __forceinline__ float4 load_a4(const void* A, long off) {
    if (ADT == 1) {
        const uchar4 u = __ldg(reinterpret_cast<const uchar4*>(reinterpret_cast<const uint8_t*>(A) + off));
        return make_float4((float)u.x, (float)u.y, (float)u.z, (float)u.w);
    }
    float4 v = __ldg(reinterpret_cast<const float4*>(reinterpret_cast<const float*>(A) + off));
    if (ADT == 2) { v.x /= 255.0f; v.y /= 255.0f; v.z /= 255.0f; v.w /= 255.0f; }
    return v;
}

template <int BM, int BN, int TM, int TN, int A_U8, int kBK>
__global__ void __launch_bounds__(kNT)
gemm_kernel(const GemmDesc g) {
    static_assert((BM / TM) * (BN / TN) == kNT, "thread tiling");
    static_assert(TM % 4 == 0 && TN % 4 == 0, "micro tile");
    constexpr int A_SLOTS = BM * kBK / 4;                    // float4 slots per A tile
    constexpr int B_SLOTS = kBK * BN / 4;
    constexpr int A_NV = (A_SLOTS + kNT - 1) / kNT;
    constexpr int B_NV = (B_SLOTS + kNT - 1) / kNT;

    __shared__ __align__(16) float As[2][kBK][BM + 4];
    __shared__ __align__(16) float Bs[2][kBK][BN + 4];
    __shared__ long rA[BM], rC[BM], rM[BM];
    __shared__ long cB[BN], cC[BN], cM[BN];
    __shared__ long kA[2][kBK], kB[2][kBK];
    __shared__ unsigned char rflag[BM];                      // 0 valid, 1 ones row, 2 out of range
    __shared__ unsigned char cvalid[BN];

    const int tid = threadIdx.x;
    const int batch = blockIdx.z / g.splits, split = blockIdx.z % g.splits;
    const long i0 = (long)blockIdx.y * BM, j0 = (long)blockIdx.x * BN;
    // symmetric product (KFAC factors): a tile that lies strictly below the diagonal is produced, transposed, by the
    // tile above it -- half the flops of the full product (reference kfac.py:42,60 computes the full a^T a / g^T g)
    if (g.sym && i0 > j0 + BN - 1) return;
    const long ktiles = (g.K + kBK - 1) / kBK;
    const long tps = (ktiles + g.splits - 1) / g.splits;
    const long kbeg = (long)split * tps * kBK;
    const long kend = min((long)g.K, kbeg + tps * kBK);

    const uint8_t* Ab = reinterpret_cast<const uint8_t*>(g.A) + g.batch_offA[batch] * (A_U8 == 1 ? 1 : 4);
    const float* Bb = g.B + g.batch_offB[batch];

    for (int t = tid; t < BM; t += kNT) {
        const long i = i0 + t;
        if (i < g.M) {
            rA[t] = affine_eval(g.rowA, i, g.gather_row);
            rC[t] = affine_eval(g.rowC, i, nullptr);
            // (symmetric products carry no mask: the mask tables hold the maps of the mirror image instead --
            //  colC of this tile's rows here, rowC of its columns below)
            rM[t] = g.mask_mode ? affine_eval(g.rowM, i, nullptr) : (g.sym ? affine_eval(g.colC, i, nullptr) : 0);
            rflag[t] = 0;
        } else {
            rA[t] = rC[t] = rM[t] = 0;
            rflag[t] = (g.ones_row && i == g.M) ? 1 : 2;
        }
    }
    for (int t = tid; t < BN; t += kNT) {
        const long j = j0 + t;
        if (j < g.N) {
            cB[t] = affine_eval(g.colB, j, nullptr);
            cC[t] = affine_eval(g.colC, j, nullptr);
            cM[t] = g.mask_mode ? affine_eval(g.colM, j, nullptr) : (g.sym ? affine_eval(g.rowC, j, nullptr) : 0);
            cvalid[t] = 1;
        } else {
            cB[t] = cC[t] = cM[t] = 0;
            cvalid[t] = 0;
        }
    }
    auto stage_red = [&](int buf, long k0) {
        if (tid < 2 * kBK) {
            const int t = tid & (kBK - 1);
            const long r = k0 + t;
            if (tid < kBK) kA[buf][t] = (r < kend) ? affine_eval(g.redA, r, g.gather_red) : 0;
            else kB[buf][t] = (r < kend) ? affine_eval(g.redB, r, nullptr) : 0;
        }
    };

    float4 a_reg[A_NV], b_reg[B_NV];
    auto load_tile = [&](int buf, long k0) {
        // ---- A ----
#pragma unroll
        for (int q = 0; q < A_NV; ++q) {
            const int s = tid + q * kNT;
            float4 v = make_float4(0.f, 0.f, 0.f, 0.f);
            if (A_SLOTS % kNT == 0 || s < A_SLOTS) {
                if (g.vecA == 1) {
                    const int il = s % BM, r = (s / BM) * 4;
                    if (k0 + r < kend) {
                        if (rflag[il] == 0) v = load_a4<A_U8>(Ab, rA[il] + kA[buf][r]);
                        else if (rflag[il] == 1) v = make_float4(1.f, 1.f, 1.f, 1.f);
                    }
                } else if (g.vecA == 2) {
                    const int il = (s % (BM / 4)) * 4, r = s / (BM / 4);
                    if (k0 + r < kend) {
                        if (rflag[il] == 0) v = load_a4<A_U8>(Ab, rA[il] + kA[buf][r]);
                        else if (rflag[il] == 1) v.x = 1.f;
                    }
                } else {
                    const int il = s % BM, r0 = s / BM;
                    float t4[4];
#pragma unroll
                    for (int c = 0; c < 4; ++c) {
                        const int r = r0 + c * (kBK / 4);
                        t4[c] = 0.f;
                        if (k0 + r < kend) {
                            if (rflag[il] == 0) t4[c] = load_a1<A_U8>(Ab, rA[il] + kA[buf][r]);
                            else if (rflag[il] == 1) t4[c] = 1.f;
                        }
                    }
                    v = make_float4(t4[0], t4[1], t4[2], t4[3]);
                }
            }
            a_reg[q] = v;
        }
        // ---- B ----
#pragma unroll
        for (int q = 0; q < B_NV; ++q) {
            const int s = tid + q * kNT;
            float4 v = make_float4(0.f, 0.f, 0.f, 0.f);
            if (B_SLOTS % kNT == 0 || s < B_SLOTS) {
                if (g.vecB == 1) {
                    const int jl = (s % (BN / 4)) * 4, r = s / (BN / 4);
                    if (k0 + r < kend && cvalid[jl])
                        v = __ldg(reinterpret_cast<const float4*>(Bb + kB[buf][r] + cB[jl]));
                } else if (g.vecB == 2) {
                    const int jl = s % BN, r = (s / BN) * 4;
                    if (k0 + r < kend && cvalid[jl])
                        v = __ldg(reinterpret_cast<const float4*>(Bb + kB[buf][r] + cB[jl]));
                } else {
                    const int jl = s % BN, r0 = s / BN;
                    float t4[4];
#pragma unroll
                    for (int c = 0; c < 4; ++c) {
                        const int r = r0 + c * (kBK / 4);
                        t4[c] = (k0 + r < kend && cvalid[jl]) ? __ldg(Bb + kB[buf][r] + cB[jl]) : 0.f;
                    }
                    v = make_float4(t4[0], t4[1], t4[2], t4[3]);
                }
            }
            b_reg[q] = v;
        }
    };
    auto store_tile = [&](int nb) {
#pragma unroll
        for (int q = 0; q < A_NV; ++q) {
            const int s = tid + q * kNT;
            if (A_SLOTS % kNT == 0 || s < A_SLOTS) {
                const float4 v = a_reg[q];
                if (g.vecA == 1) {
                    const int il = s % BM, r = (s / BM) * 4;
                    As[nb][r + 0][il] = v.x; As[nb][r + 1][il] = v.y;
                    As[nb][r + 2][il] = v.z; As[nb][r + 3][il] = v.w;
                } else if (g.vecA == 2) {
                    const int il = (s % (BM / 4)) * 4, r = s / (BM / 4);
                    *reinterpret_cast<float4*>(&As[nb][r][il]) = v;
                } else {
                    const int il = s % BM, r0 = s / BM;
                    As[nb][r0][il] = v.x; As[nb][r0 + kBK / 4][il] = v.y;
                    As[nb][r0 + 2 * (kBK / 4)][il] = v.z; As[nb][r0 + 3 * (kBK / 4)][il] = v.w;
                }
            }
        }
#pragma unroll
        for (int q = 0; q < B_NV; ++q) {
            const int s = tid + q * kNT;
            if (B_SLOTS % kNT == 0 || s < B_SLOTS) {
                const float4 v = b_reg[q];
                if (g.vecB == 1) {
                    const int jl = (s % (BN / 4)) * 4, r = s / (BN / 4);
                    *reinterpret_cast<float4*>(&Bs[nb][r][jl]) = v;
                } else if (g.vecB == 2) {
                    const int jl = s % BN, r = (s / BN) * 4;
                    Bs[nb][r + 0][jl] = v.x; Bs[nb][r + 1][jl] = v.y;
                    Bs[nb][r + 2][jl] = v.z; Bs[nb][r + 3][jl] = v.w;
                } else {
                    const int jl = s % BN, r0 = s / BN;
                    Bs[nb][r0][jl] = v.x; Bs[nb][r0 + kBK / 4][jl] = v.y;
                    Bs[nb][r0 + 2 * (kBK / 4)][jl] = v.z; Bs[nb][r0 + 3 * (kBK / 4)][jl] = v.w;
                }
            }
        }
    };

    float acc[TM][TN];
#pragma unroll
    for (int m = 0; m < TM; ++m)
#pragma unroll
        for (int n = 0; n < TN; ++n) acc[m][n] = 0.f;

    const int tx = tid % (BN / TN), ty = tid / (BN / TN);

    stage_red(0, kbeg);
    __syncthreads();
    if (kbeg < kend) {
        load_tile(0, kbeg);
        store_tile(0);
        stage_red(1, kbeg + kBK);
    }
    __syncthreads();

    for (long k0 = kbeg; k0 < kend; k0 += kBK) {
        const int cur = (int)(((k0 - kbeg) / kBK) & 1);
        const bool has_next = (k0 + kBK) < kend;
        if (has_next) load_tile(cur ^ 1, k0 + kBK);
#pragma unroll
        for (int kk = 0; kk < kBK; ++kk) {
            float a[TM], b[TN];
#pragma unroll
            for (int m = 0; m < TM; m += 4) {
                const float4 t = *reinterpret_cast<const float4*>(&As[cur][kk][ty * TM + m]);
                a[m] = t.x; a[m + 1] = t.y; a[m + 2] = t.z; a[m + 3] = t.w;
            }
#pragma unroll
            for (int n = 0; n < TN; n += 4) {
                const float4 t = *reinterpret_cast<const float4*>(&Bs[cur][kk][tx * TN + n]);
                b[n] = t.x; b[n + 1] = t.y; b[n + 2] = t.z; b[n + 3] = t.w;
            }
#pragma unroll
            for (int m = 0; m < TM; ++m)
#pragma unroll
                for (int n = 0; n < TN; ++n) acc[m][n] = fmaf(a[m], b[n], acc[m][n]);
        }
        if (has_next) {
            store_tile(cur ^ 1);
            stage_red(cur, k0 + 2 * kBK);
        }
        __syncthreads();
    }

    // ---- epilogue ----
    float* Cb = g.C + g.batch_offC[batch] + (long)split * g.split_stride;
    const float* Mb = g.mask ? g.mask + g.batch_offM[batch] : nullptr;
    const bool final_pass = (g.splits == 1);
#pragma unroll
    for (int m = 0; m < TM; ++m) {
        const int il = ty * TM + m;
        const unsigned char fl = rflag[il];
        if (fl == 2) continue;
#pragma unroll
        for (int n = 0; n < TN; ++n) {
            const int jl = tx * TN + n;
            if (!cvalid[jl]) continue;
            if (fl == 1) {       // logical all-ones row: never a byte operand, no 1/255
                g.C_ones[(long)split * g.split_stride + (j0 + jl) * g.ones_stride] = acc[m][n] * g.alpha;
                continue;
            }
            float v = acc[m][n] * (A_U8 == 1 ? g.alpha / 255.0f : g.alpha);
            float* dst = Cb + rC[il] + cC[jl];
            if (final_pass) {
                if (g.bias_mode == 1) v += __ldg(g.bias + j0 + jl);
                else if (g.bias_mode == 2) v += __ldg(g.bias + i0 + il);
                if (g.act == 1) v = fmaxf(v, 0.f);
                else if (g.act == 2) v = tanhf(v);
                if (g.mask_mode == 1) v = (__ldg(Mb + rM[il] + cM[jl]) > 0.f) ? v : 0.f;
                else if (g.mask_mode == 2) { const float t = __ldg(Mb + rM[il] + cM[jl]); v *= (1.f - t * t); }
                if (g.accumulate) v += *dst;
            }
            *dst = v;
            if (g.sym) {
                // mirror image of an element above the diagonal, when the tile that would hold it was skipped
                const long gi = i0 + il, gj = j0 + jl;
                if (gi < gj && (gj / BM) * BM > (gi / BN) * BN + BN - 1) Cb[cM[jl] + rM[il]] = v;
            }
        }
    }
}

template <int BM, int BN, int TM, int TN, int kBK>
static void launch_tile(const GemmDesc& g, cudaStream_t st) {
    const long Mtot = g.M + (g.ones_row ? 1 : 0);
    dim3 grid((unsigned)((g.N + BN - 1) / BN), (unsigned)((Mtot + BM - 1) / BM), (unsigned)(g.batch * g.splits));
    if (g.a_dtype == 1) gemm_kernel<BM, BN, TM, TN, 1, kBK><<<grid, kNT, 0, st>>>(g);
    else if (g.a_dtype == 2) gemm_kernel<BM, BN, TM, TN, 2, kBK><<<grid, kNT, 0, st>>>(g);
    else gemm_kernel<BM, BN, TM, TN, 0, kBK><<<grid, kNT, 0, st>>>(g);
}

// 0: fp32 SIMT engine only; 1: tensor-core (tcgen05, 3xTF32) engine wherever it applies;
// 2 (default): per-descriptor choice between the two from measured crossover points.
static int g_gemm_mode = -1;
int gemm_mode() {
    if (g_gemm_mode < 0) {
        const char* e = getenv("A2CB_GEMM");
        g_gemm_mode = 2;
        if (e && (e[0] == 's' || e[0] == 'S')) g_gemm_mode = 0;
        else if (e && (e[0] == 't' || e[0] == 'T')) g_gemm_mode = 1;
    }
    return g_gemm_mode;
}
void set_gemm_mode(int m) { g_gemm_mode = (m < 0 || m > 2) ? 2 : m; }

// Measured on B200 (tools/layer_bench.py, profiles/): with skinny outputs (N <= 64 channels) and
// reduction-contiguous operands the tensor-core engine wins from a few thousand rows upwards; the
// transposed-loader shapes (weight gradients) and the wide fc layers are still faster on the SIMT
// engine, whose split-K partials also keep their long reductions in round-to-nearest fp32.
static bool prefer_tc(const GemmDesc& g) {
    if (g.B_image) return g.vecA != 2 && g.M >= 2048 && g.K >= 64;      // B arrives by TMA: loaders only feed A
    return g.vecA != 2 && g.vecB != 1 && g.N <= 64 && g.N >= 16 && g.M >= 4096 && g.K >= 64;
}

int gemm(const GemmDesc& g, cudaStream_t st) {
    A2CB_REQUIRE(g.A && g.B && g.C, "gemm: null operand");
    A2CB_REQUIRE(g.M >= 0 && g.N >= 0 && g.K >= 0, "gemm: negative extent");
    A2CB_REQUIRE(g.a_dtype >= 0 && g.a_dtype <= 2, "gemm: a_dtype");
    A2CB_REQUIRE(g.batch >= 1 && g.batch <= 4 && g.splits >= 1, "gemm: batch/splits out of range");
    A2CB_REQUIRE(g.M < (1L << 31) && g.N < (1L << 31) && g.K < (1L << 31), "gemm: extent >= 2^31");
    A2CB_REQUIRE(!g.ones_row || g.C_ones, "gemm: ones_row without C_ones");
    A2CB_REQUIRE(!g.mask_mode || g.mask, "gemm: mask_mode without mask");
    A2CB_REQUIRE(!g.bias_mode || g.bias, "gemm: bias_mode without bias");
    A2CB_REQUIRE(g.splits == 1 || (!g.bias_mode && !g.act && !g.mask_mode && !g.accumulate),
                 "gemm: split-K partials cannot carry an epilogue");
    A2CB_REQUIRE(g.vecA != 1 || g.K % 4 == 0, "gemm: vecA=1 needs K %% 4 == 0");
    A2CB_REQUIRE(g.vecA != 2 || g.M % 4 == 0, "gemm: vecA=2 needs M %% 4 == 0");
    A2CB_REQUIRE(g.vecB != 1 || g.N % 4 == 0, "gemm: vecB=1 needs N %% 4 == 0");
    A2CB_REQUIRE(g.vecB != 2 || g.K % 4 == 0, "gemm: vecB=2 needs K %% 4 == 0");
    A2CB_REQUIRE(!g.sym || (g.M == g.N && !g.ones_row && !g.bias_mode && !g.act && !g.mask_mode && !g.accumulate && g.tile == 1001),
                 "gemm: sym needs a square product without epilogue on the SIMT engine (tile 1001)");
    if (g.M + g.ones_row == 0 || g.N == 0) return A2CB_OK;
    if (g.tile != 1001 && gemm_tc_supported(g)) {
        const int mode = gemm_mode();
        if (g.tile == 1000 || mode == 1 || (mode == 2 && prefer_tc(g))) return gemm_tc(g, st);
    }
    int bn = g.tile >= 1000 ? 0 : g.tile;
    if (bn == 0) {
        bn = g.N <= 32 ? 32 : (g.N <= 64 ? 64 : 128);
        // small problems: prefer 64x64 tiles over 128x128 when the big tile would leave most SMs idle
        const long Mtot = g.M + g.ones_row;
        const long big = ((Mtot + 127) / 128) * ((g.N + 127) / 128) * g.batch * g.splits;
        if (bn == 128 && big < kNumSM) bn = 65;
    }
    // deeper k-tiles (32) for the narrow tiles: half the barriers / offset stagings per reduction
    // element on latency-bound shapes; the 128x128 tile keeps 16 to stay inside 48 KiB of static smem
    if (bn == 32) launch_tile<128, 32, 4, 4, 32>(g, st);
    else if (bn == 64) launch_tile<128, 64, 8, 4, 16>(g, st);
    else if (bn == 65) launch_tile<64, 64, 4, 4, 32>(g, st);
    else if (bn == 128) launch_tile<128, 128, 8, 8, 16>(g, st);
    else { set_error("gemm: unsupported tile %d", bn); return A2CB_ERR_UNSUPPORTED; }
    return check_launch("gemm");
}

// dst[e] = (accumulate ? beta * dst[e] : 0) + act( sum_s src[s * stride + e] + bias[e % ncols] )  for e in [0, count): folds split-K
// partials (fixed order, deterministic) and applies the epilogue the split GEMM could not.
__global__ void __launch_bounds__(256)
reduce_splits_kernel(float* __restrict__ dst, const float* __restrict__ src, long count, long stride,
                     int splits, int accumulate, const float* __restrict__ bias, int ncols, int act, float beta) {
    const long e = (long)blockIdx.x * blockDim.x + threadIdx.x;
    if (e >= count) return;
    float s = 0.f;
    for (int k = 0; k < splits; ++k) s += __ldg(src + (long)k * stride + e);
    if (bias) s += __ldg(bias + (e % ncols));
    if (act == 1) s = fmaxf(s, 0.f);
    else if (act == 2) s = tanhf(s);
    dst[e] = accumulate ? beta * dst[e] + s : s;
}

// Same fold for MANY splits (weight-gradient slabs of 148 CTAs, bias column sums of ~1000 blocks): 32 consecutive
// elements per block, 8 split lanes each summing every 8th slab (coalesced 128-byte rows, 8 x fewer dependent
// loads per thread), then a fixed-order shared-memory fold -- still deterministic.
__global__ void __launch_bounds__(256)
reduce_splits_wide_kernel(float* __restrict__ dst, const float* __restrict__ src, long count, long stride,
                          int splits, int accumulate, const float* __restrict__ bias, int ncols, int act, float beta) {
    __shared__ float red[8][32];
    const int el = threadIdx.x & 31, sl = threadIdx.x >> 5;
    const long e = (long)blockIdx.x * 32 + el;
    float s0 = 0.f, s1 = 0.f;
    if (e < count) {
        int k = sl;
        for (; k + 8 < splits; k += 16) {
            s0 += __ldg(src + (long)k * stride + e);
            s1 += __ldg(src + (long)(k + 8) * stride + e);
        }
        if (k < splits) s0 += __ldg(src + (long)k * stride + e);
    }
    red[sl][el] = s0 + s1;
    __syncthreads();
    if (sl == 0 && e < count) {
        float s = red[0][el];
#pragma unroll
        for (int j = 1; j < 8; ++j) s += red[j][el];
        if (bias) s += __ldg(bias + (e % ncols));
        if (act == 1) s = fmaxf(s, 0.f);
        else if (act == 2) s = tanhf(s);
        dst[e] = accumulate ? beta * dst[e] + s : s;
    }
}

int reduce_splits(float* dst, const float* src, long count, long stride, int splits, int accumulate,
                  const float* bias, int ncols, int act, float beta, cudaStream_t st) {
    if (count == 0) return A2CB_OK;
    A2CB_REQUIRE(!bias || ncols > 0, "reduce_splits: bias needs ncols");
    if (splits >= 32) {
        reduce_splits_wide_kernel<<<(unsigned)((count + 31) / 32), 256, 0, st>>>(dst, src, count, stride, splits, accumulate,
                                                                                  bias, ncols, act, beta);
        return check_launch("reduce_splits");
    }
    reduce_splits_kernel<<<(unsigned)((count + 255) / 256), 256, 0, st>>>(dst, src, count, stride, splits, accumulate,
                                                                          bias, ncols, act, beta);
    return check_launch("reduce_splits");
}

}  // namespace a2cb
