// Affine-indexed GEMM engine, tensor-core path (tcgen05 / TMEM, sm_100a):
//
//     C[rowC(i) + colC(j)] = epi( alpha * sum_r  A[rowA(i) + redA(r)] * B[redB(r) + colB(j)] )
//
// Same descriptor, same index maps and same epilogue as the fp32 SIMT engine (gemm.cu); only the
// inner product moves to the 5th-generation tensor cores.  fp32 parity (1e-4 on losses and
// per-step parameters, BASELINE.json) rules out single-pass TF32, so every fp32
// operand is split on the fly into hi = tf32(x) and lo = tf32(x - hi) and three MMAs
//     D += Ah*Bh + Ah*Bl + Al*Bh
// accumulate into ONE fp32 TMEM accumulator (the dropped Al*Bl term is ~2^-22 relative).
//
// CTA = 128 x BN output tile, 256 threads:
//   * all 8 warps are loaders: gathered global loads (implicit im2col / minibatch index list /
//     uint8 -> fp32/255) -> hi/lo split in registers -> st.shared.v4 into the canonical K-major
//     no-swizzle UMMA layout (8-row x 16-byte core matrices, LBO = 128 B, SBO = 1 KiB).
//     Operands that are contiguous along the row rather than the reduction (weight gradients)
//     are transposed 4x4 across lanes with shuffles so that every store is one 16-byte row;
//   * one thread issues the tcgen05.mma chain of a stage and commits it to the stage's mbarrier;
//     the next stage's global loads and conversion overlap the asynchronous MMAs;
//   * the tensor core's fp32 accumulator TRUNCATES (measured: error grows ~7e-9 * K, biased toward
//     zero -- tools/tc_accuracy.py), which weight gradients with their long, cancelling reductions
//     cannot afford.  The reduction is therefore cut into chains of `tc_chunk` k-blocks that
//     alternate between two TMEM accumulators; a finished chain is drained with tcgen05.ld (32x32b)
//     and added to fp32 registers with round-to-nearest while the next chain is running;
//   * epilogue: bias / activation / mask / accumulate on the register sums, scattered through
//     rowC/colC.
#include "common.cuh"
#include "decls.cuh"
#include "gemm.cuh"

namespace a2cb {

constexpr int TBM = 128;
constexpr int TBK = 32;
constexpr int TNT = 256;
constexpr int TSTAGES = 2;
constexpr int TKC = TBK / 4;                 // 16-byte k-chunks per row

__device__ __forceinline__ long tc_affine(const Affine3& a, long i, const int64_t* gather) {
    const unsigned ii = (unsigned)i, d1 = (unsigned)a.d1, d2 = (unsigned)a.d2;
    unsigned q0 = ii / d1;
    const unsigned rem = ii - q0 * d1;
    const unsigned q1 = rem / d2;
    const unsigned q2 = rem - q1 * d2;
    const long g0 = gather ? (long)gather[q0] : (long)q0;
    return g0 * a.s0 + (long)q1 * a.s1 + (long)q2 * a.s2;
}

// ADT 1 (uint8 frames): the byte stays an exact integer-valued float -- exact in TF32, so the A
// operand needs no lo part and one of the three MMAs disappears; 1/255 is folded into the epilogue.
template <int ADT>
__device__ __forceinline__ float tc_load1(const void* A, long off) {
    if (ADT == 1) return (float)__ldg(reinterpret_cast<const uint8_t*>(A) + off);
    const float v = __ldg(reinterpret_cast<const float*>(A) + off);
    return ADT == 2 ? v / 255.0f : v;
}
template <int ADT>
__device__ __forceinline__ float4 tc_load4(const void* A, long off) {
    if (ADT == 1) {
        const uchar4 u = __ldg(reinterpret_cast<const uchar4*>(reinterpret_cast<const uint8_t*>(A) + off));
        return make_float4((float)u.x, (float)u.y, (float)u.z, (float)u.w);
    }
    float4 v = __ldg(reinterpret_cast<const float4*>(reinterpret_cast<const float*>(A) + off));
    if (ADT == 2) { v.x /= 255.0f; v.y /= 255.0f; v.z /= 255.0f; v.w /= 255.0f; }
    return v;
}

__device__ __forceinline__ uint32_t tc_smem(const void* p) { return (uint32_t)__cvta_generic_to_shared(p); }

__device__ __forceinline__ float tf32_rna(float x) {
    uint32_t r;
    asm("cvt.rna.tf32.f32 %0, %1;" : "=r"(r) : "f"(x));
    return __uint_as_float(r);
}
__device__ __forceinline__ void split4(const float4 v, float4& hi, float4& lo) {
    hi.x = tf32_rna(v.x); hi.y = tf32_rna(v.y); hi.z = tf32_rna(v.z); hi.w = tf32_rna(v.w);
    lo.x = tf32_rna(v.x - hi.x); lo.y = tf32_rna(v.y - hi.y);
    lo.z = tf32_rna(v.z - hi.z); lo.w = tf32_rna(v.w - hi.w);
}

// byte offset of the 16-byte chunk (row, kc) inside an operand tile in the canonical layout
__device__ __forceinline__ uint32_t tile_off(int row, int kc) {
    return (uint32_t)((((row >> 3) * TKC + kc) << 7) + ((row & 7) << 4));
}

// 4x4 transpose across the 4 lanes of a quad: lane c ends up with component c of lanes 0..3.
__device__ __forceinline__ float4 quad_transpose(const float4 v, int lane) {
    const int l4 = lane & 3, base = lane & ~3;
    float o0 = 0.f, o1 = 0.f, o2 = 0.f, o3 = 0.f;
#pragma unroll
    for (int t = 0; t < 4; ++t) {
        const int sel = (l4 - t) & 3;                       // component this lane contributes in round t
        const float send = sel == 0 ? v.x : (sel == 1 ? v.y : (sel == 2 ? v.z : v.w));
        const int j = (l4 + t) & 3;
        const float got = __shfl_sync(0xffffffffu, send, base + j);
        o0 = j == 0 ? got : o0; o1 = j == 1 ? got : o1; o2 = j == 2 ? got : o2; o3 = j == 3 ? got : o3;
    }
    return make_float4(o0, o1, o2, o3);
}

__device__ __forceinline__ uint64_t umma_desc(uint32_t smem_addr) {
    // K-major, SWIZZLE_NONE: start >> 4 | LBO(128 B) >> 4 << 16 | SBO(TKC * 128 B) >> 4 << 32 | version 1 << 46
    return (uint64_t)((smem_addr & 0x3FFFF) >> 4) | ((uint64_t)(128 >> 4) << 16) |
           ((uint64_t)((TKC * 128) >> 4) << 32) | (1ull << 46);
}
__device__ __forceinline__ void umma_tf32(uint32_t d_tmem, uint64_t a, uint64_t b, uint32_t idesc, uint32_t acc) {
    asm volatile(
        "{\n\t"
        ".reg .pred p;\n\t"
        "setp.ne.b32 p, %4, 0;\n\t"
        "tcgen05.mma.cta_group::1.kind::tf32 [%0], %1, %2, %3, p;\n\t"
        "}\n" ::"r"(d_tmem), "l"(a), "l"(b), "r"(idesc), "r"(acc)
        : "memory");
}
__device__ __forceinline__ void tc_mbar_init(uint64_t* bar, uint32_t count) {
    asm volatile("mbarrier.init.shared::cta.b64 [%0], %1;" ::"r"(tc_smem(bar)), "r"(count));
}
__device__ __forceinline__ void tc_mbar_wait(uint64_t* bar, uint32_t parity) {
    uint32_t ok = 0;
    while (!ok) {
        asm volatile(
            "{\n"
            ".reg .pred p;\n"
            "mbarrier.try_wait.parity.shared::cta.b64 p, [%1], %2;\n"
            "selp.u32 %0, 1, 0, p;\n"
            "}\n"
            : "=r"(ok)
            : "r"(tc_smem(bar)), "r"(parity)
            : "memory");
    }
}
__device__ __forceinline__ void tc_commit(uint64_t* bar) {
    asm volatile("tcgen05.commit.cta_group::1.mbarrier::arrive::one.shared::cluster.b64 [%0];" ::"r"(tc_smem(bar))
                 : "memory");
}

template <int BN, int ADT>
__global__ void __launch_bounds__(TNT)
gemm_tc_kernel(const GemmDesc g) {
    constexpr int A_TILE = TBM * TBK * 4;                    // bytes of one (hi or lo) A tile
    constexpr int B_TILE = BN * TBK * 4;
    constexpr int STAGE = 2 * A_TILE + 2 * B_TILE;
    constexpr int A_NV = TBM * TKC / TNT;                    // 16-byte slots per thread
    constexpr int B_SLOTS = BN * TKC;
    constexpr int B_NV = (B_SLOTS + TNT - 1) / TNT;

    extern __shared__ __align__(128) uint8_t tc_dyn[];
    __shared__ long rA[TBM], rC[TBM], rM[TBM];
    __shared__ long cB[BN], cC[BN], cM[BN];
    __shared__ long kA[2][TBK], kB[2][TBK];
    __shared__ unsigned char rflag[TBM];
    __shared__ unsigned char cvalid[BN];
    __shared__ __align__(8) uint64_t mbar[TSTAGES];
    __shared__ __align__(8) uint64_t bbar[TSTAGES];      // TMA arrival of a pre-tiled B block
    __shared__ uint32_t tmem_base_s;

    const int tid = threadIdx.x, lane = tid & 31, warp = tid >> 5;
    const int batch = blockIdx.z / g.splits, split = blockIdx.z % g.splits;
    const long i0 = (long)blockIdx.y * TBM, j0 = (long)blockIdx.x * BN;
    const long ktiles = (g.K + TBK - 1) / TBK;
    const long tps = (ktiles + g.splits - 1) / g.splits;
    const long kbeg = (long)split * tps * TBK;
    const long kend = min((long)g.K, kbeg + tps * TBK);
    const int nkb = kend > kbeg ? (int)((kend - kbeg + TBK - 1) / TBK) : 0;

    const uint8_t* Ab = reinterpret_cast<const uint8_t*>(g.A) + g.batch_offA[batch] * (ADT == 1 ? 1 : 4);
    const float* Bb = g.B + g.batch_offB[batch];

    // ---- one-time setup ----
    if (warp == 0) {
        asm volatile("tcgen05.alloc.cta_group::1.sync.aligned.shared::cta.b32 [%0], %1;" ::"r"(tc_smem(&tmem_base_s)),
                     "r"((uint32_t)(2 * BN)));
        asm volatile("tcgen05.relinquish_alloc_permit.cta_group::1.sync.aligned;" ::);
    }
    if (tid == 32) {
        for (int s = 0; s < TSTAGES; ++s) { tc_mbar_init(&mbar[s], 1); tc_mbar_init(&bbar[s], 1); }
        asm volatile("fence.mbarrier_init.release.cluster;" ::: "memory");
    }
    for (int t = tid; t < TBM; t += TNT) {
        const long i = i0 + t;
        if (i < g.M) {
            rA[t] = tc_affine(g.rowA, i, g.gather_row);
            rC[t] = tc_affine(g.rowC, i, nullptr);
            rM[t] = g.mask_mode ? tc_affine(g.rowM, i, nullptr) : 0;
            rflag[t] = 0;
        } else {
            rA[t] = rC[t] = rM[t] = 0;
            rflag[t] = (g.ones_row && i == g.M) ? 1 : 2;
        }
    }
    for (int t = tid; t < BN; t += TNT) {
        const long j = j0 + t;
        if (j < g.N) {
            cB[t] = tc_affine(g.colB, j, nullptr);
            cC[t] = tc_affine(g.colC, j, nullptr);
            cM[t] = g.mask_mode ? tc_affine(g.colM, j, nullptr) : 0;
            cvalid[t] = 1;
        } else {
            cB[t] = cC[t] = cM[t] = 0;
            cvalid[t] = 0;
        }
    }
    auto stage_red = [&](int buf, long k0) {
        if (tid < 2 * TBK) {
            const int t = tid & (TBK - 1);
            const long r = k0 + t;
            if (tid < TBK) kA[buf][t] = (r < kend) ? tc_affine(g.redA, r, g.gather_red) : 0;
            else kB[buf][t] = (r < kend) ? tc_affine(g.redB, r, nullptr) : 0;
        }
    };

    // pre-tiled B image: [batch][column tile][k-block][hi | lo][BN x 32 floats in tile_off order]
    const bool has_image = g.B_image != nullptr;
    const long n_ctiles = (g.N + BN - 1) / BN;
    const float* Bimg = has_image
        ? g.B_image + ((long)batch * n_ctiles + blockIdx.x) * ktiles * (2 * BN * TBK) : nullptr;
    float4 a_reg[A_NV], b_reg[B_NV];
    static_assert(A_NV == 4, "row-fast A mapping assumes 4 slots per thread");
    const int rowb = 16 * warp + (lane & 7), kcb = lane >> 3;   // row-fast A slots of this thread
    long ra_reg[2];
    unsigned char fl_reg[2];
    // Every slot ends up as (one row) x (4 consecutive r).  "row-fast" slots load that shape directly;
    // "transposed" slots load (4 consecutive rows) x (one r) and are transposed across a lane quad.
    auto load_tile = [&](int buf, long k0) {
#pragma unroll
        for (int q = 0; q < A_NV; ++q) {
            const int s = tid + q * TNT;
            float4 v = make_float4(0.f, 0.f, 0.f, 0.f);
            if (g.vecA == 2) {
                const int r = s & (TBK - 1), row = (s >> 5) * 4;
                if (k0 + r < kend) {
                    if (rflag[row] == 0) v = tc_load4<ADT>(Ab, rA[row] + kA[buf][r]);
                    else if (rflag[row] == 1) v.x = 1.f;
                }
                v = quad_transpose(v, lane);
            } else {
                // row-fast slots: a warp covers 16 rows; per instruction 8 rows x 64 contiguous bytes
                // (4 lines instead of 32), and every quarter-warp writes 8 distinct rows of a core matrix
                const int r = (kcb + 4 * (q >> 1)) * 4;
                const long ro = ra_reg[q & 1];
                const unsigned char f = fl_reg[q & 1];
                if (f == 0) {
                    if (g.vecA == 1) {
                        if (k0 + r < kend) v = tc_load4<ADT>(Ab, ro + kA[buf][r]);
                    } else {
                        if (k0 + r + 0 < kend) v.x = tc_load1<ADT>(Ab, ro + kA[buf][r + 0]);
                        if (k0 + r + 1 < kend) v.y = tc_load1<ADT>(Ab, ro + kA[buf][r + 1]);
                        if (k0 + r + 2 < kend) v.z = tc_load1<ADT>(Ab, ro + kA[buf][r + 2]);
                        if (k0 + r + 3 < kend) v.w = tc_load1<ADT>(Ab, ro + kA[buf][r + 3]);
                    }
                } else if (f == 1) {
                    v.x = (k0 + r + 0 < kend) ? 1.f : 0.f; v.y = (k0 + r + 1 < kend) ? 1.f : 0.f;
                    v.z = (k0 + r + 2 < kend) ? 1.f : 0.f; v.w = (k0 + r + 3 < kend) ? 1.f : 0.f;
                }
            }
            a_reg[q] = v;
        }
        if (has_image) return;
#pragma unroll
        for (int q = 0; q < B_NV; ++q) {
            const int s = tid + q * TNT;
            float4 v = make_float4(0.f, 0.f, 0.f, 0.f);
            if (g.vecB == 1) {
                // (B_SLOTS % TNT != 0 only for BN = 32, where every thread still takes part in the shuffles)
                const int r = s & (TBK - 1), col = (s >> 5) * 4;
                if (col < BN && k0 + r < kend && cvalid[col])
                    v = __ldg(reinterpret_cast<const float4*>(Bb + kB[buf][r] + cB[col]));
                v = quad_transpose(v, lane);
            } else if (B_SLOTS % TNT == 0 || s < B_SLOTS) {
                const int col = s % BN, r = (s / BN) * 4;
                if (cvalid[col]) {
                    if (g.vecB == 2) {
                        if (k0 + r < kend) v = __ldg(reinterpret_cast<const float4*>(Bb + kB[buf][r] + cB[col]));
                    } else {
                        if (k0 + r + 0 < kend) v.x = __ldg(Bb + kB[buf][r + 0] + cB[col]);
                        if (k0 + r + 1 < kend) v.y = __ldg(Bb + kB[buf][r + 1] + cB[col]);
                        if (k0 + r + 2 < kend) v.z = __ldg(Bb + kB[buf][r + 2] + cB[col]);
                        if (k0 + r + 3 < kend) v.w = __ldg(Bb + kB[buf][r + 3] + cB[col]);
                    }
                }
            }
            b_reg[q] = v;
        }
    };
    auto store_tile = [&](int stage) {
        uint8_t* Ah = tc_dyn + stage * STAGE;
        uint8_t* Al = Ah + A_TILE;
        uint8_t* Bh = Al + A_TILE;
        uint8_t* Bl = Bh + B_TILE;
#pragma unroll
        for (int q = 0; q < A_NV; ++q) {
            const int s = tid + q * TNT;
            int row, kc;
            if (g.vecA == 2) { row = (s >> 5) * 4 + (lane & 3); kc = (s & (TBK - 1)) >> 2; }
            else { row = rowb + 8 * (q & 1); kc = kcb + 4 * (q >> 1); }
            const uint32_t o = tile_off(row, kc);
            if (ADT == 1) {
                *reinterpret_cast<float4*>(Ah + o) = a_reg[q];     // bytes (and the ones row) are exact in TF32
            } else {
                float4 hi, lo;
                split4(a_reg[q], hi, lo);
                *reinterpret_cast<float4*>(Ah + o) = hi;
                *reinterpret_cast<float4*>(Al + o) = lo;
            }
        }
        if (has_image) return;
#pragma unroll
        for (int q = 0; q < B_NV; ++q) {
            const int s = tid + q * TNT;
            int col, kc;
            if (g.vecB == 1) { col = (s >> 5) * 4 + (lane & 3); kc = (s & (TBK - 1)) >> 2; }
            else { col = s % BN; kc = s / BN; }
            if (col < BN && (g.vecB == 1 || B_SLOTS % TNT == 0 || s < B_SLOTS)) {
                float4 hi, lo;
                split4(b_reg[q], hi, lo);
                const uint32_t o = tile_off(col, kc);
                *reinterpret_cast<float4*>(Bh + o) = hi;
                *reinterpret_cast<float4*>(Bl + o) = lo;
            }
        }
    };

    // instruction descriptor: D fp32, A/B tf32, both K-major, N = BN, M = 128
    const uint32_t idesc = (1u << 4) | (2u << 7) | (2u << 10) | ((uint32_t)(BN >> 3) << 17) | ((uint32_t)(TBM >> 4) << 24);

    stage_red(0, kbeg);
    asm volatile("tcgen05.fence::before_thread_sync;" ::: "memory");
    __syncthreads();
    asm volatile("tcgen05.fence::after_thread_sync;" ::: "memory");
    const uint32_t tmem_d = tmem_base_s;
    ra_reg[0] = rA[rowb]; ra_reg[1] = rA[rowb + 8];
    fl_reg[0] = rflag[rowb]; fl_reg[1] = rflag[rowb + 8];
    if (nkb > 0) {
        load_tile(0, kbeg);
        stage_red(1, kbeg + TBK);
    }
    __syncthreads();

    // fp32 register sums of the drained chains: this thread owns row (32 * (warp & 3) + lane) and the
    // column half (warp >> 2) of the tile
    constexpr int HALF = BN / 2;
    float acc[HALF];
#pragma unroll
    for (int n = 0; n < HALF; ++n) acc[n] = 0.f;
    const int CH = g.tc_chunk > 0 ? g.tc_chunk : 8;
    int drained = 0;
    auto drain = [&](int buf) {
#pragma unroll
        for (int c0 = 0; c0 < HALF; c0 += 16) {
            uint32_t v[16];
            const uint32_t taddr = tmem_d + ((uint32_t)(32 * (warp & 3)) << 16) +
                                   (uint32_t)(buf * BN + (warp >> 2) * HALF + c0);
            asm volatile(
                "tcgen05.ld.sync.aligned.32x32b.x16.b32"
                "{%0, %1, %2, %3, %4, %5, %6, %7, %8, %9, %10, %11, %12, %13, %14, %15}, [%16];\n"
                : "=r"(v[0]), "=r"(v[1]), "=r"(v[2]), "=r"(v[3]), "=r"(v[4]), "=r"(v[5]), "=r"(v[6]), "=r"(v[7]),
                  "=r"(v[8]), "=r"(v[9]), "=r"(v[10]), "=r"(v[11]), "=r"(v[12]), "=r"(v[13]), "=r"(v[14]), "=r"(v[15])
                : "r"(taddr));
            asm volatile("tcgen05.wait::ld.sync.aligned;" ::: "memory");
#pragma unroll
            for (int n = 0; n < 16; ++n) acc[c0 + n] += __uint_as_float(v[n]);
        }
    };

    for (int kb = 0; kb < nkb; ++kb) {
        const int s = kb % TSTAGES;
        if (kb >= TSTAGES) {
            tc_mbar_wait(&mbar[s], (uint32_t)(((kb / TSTAGES) - 1) & 1));
            // k-blocks <= kb - 2 are complete: drain every chain that ended there
            const int n_complete = (kb - 1) / CH;
            if (drained < n_complete) {
                asm volatile("tcgen05.fence::after_thread_sync;" ::: "memory");
                while (drained < n_complete) { drain(drained & 1); ++drained; }
                asm volatile("tcgen05.fence::before_thread_sync;" ::: "memory");
            }
        }
        if (has_image && tid == 0) {
            // stage s is free (its last MMAs were waited for above): fetch this k-block's B (hi | lo) with
            // ONE TMA bulk copy; it lands while the warps convert and store the A tile
            const uint32_t bytes = 2 * B_TILE;
            asm volatile("mbarrier.arrive.expect_tx.shared::cta.b64 _, [%0], %1;" ::"r"(tc_smem(&bbar[s])), "r"(bytes)
                         : "memory");
            asm volatile(
                "cp.async.bulk.shared::cluster.global.mbarrier::complete_tx::bytes [%0], [%1], %2, [%3];" ::"r"(
                    tc_smem(tc_dyn + s * STAGE + 2 * A_TILE)),
                "l"(Bimg + (kbeg / TBK + kb) * (2 * BN * TBK)), "r"(bytes), "r"(tc_smem(&bbar[s]))
                : "memory");
        }
        store_tile(s);
        if (kb + 1 < nkb) load_tile((kb + 1) & 1, kbeg + (long)(kb + 1) * TBK);
        stage_red(kb & 1, kbeg + (long)(kb + 2) * TBK);
        asm volatile("fence.proxy.async.shared::cta;" ::: "memory");
        __syncthreads();
        if (tid == 0) {
            if (has_image) tc_mbar_wait(&bbar[s], (uint32_t)((kb / TSTAGES) & 1));
            asm volatile("tcgen05.fence::after_thread_sync;" ::: "memory");
            const uint32_t ah = tc_smem(tc_dyn + s * STAGE), al = ah + A_TILE, bh = al + A_TILE, bl = bh + B_TILE;
            const uint32_t d = tmem_d + (uint32_t)(((kb / CH) & 1) * BN);
            const bool chain_start = (kb % CH) == 0;
#pragma unroll
            for (int ks = 0; ks < TBK / 8; ++ks) {
                const uint32_t o = ks * 256;                 // two 128-byte core matrices per K = 8 step
                const uint32_t acc0 = (!chain_start || ks > 0) ? 1u : 0u;
                if (ADT != 1) umma_tf32(d, umma_desc(al + o), umma_desc(bh + o), idesc, acc0);
                umma_tf32(d, umma_desc(ah + o), umma_desc(bl + o), idesc, ADT != 1 ? 1u : acc0);
                umma_tf32(d, umma_desc(ah + o), umma_desc(bh + o), idesc, 1u);
            }
            tc_commit(&mbar[s]);
        }
    }
    if (nkb > 0) {
        const int last = nkb - 1;
        tc_mbar_wait(&mbar[last % TSTAGES], (uint32_t)((last / TSTAGES) & 1));
        asm volatile("tcgen05.fence::after_thread_sync;" ::: "memory");
        const int n_chains = last / CH + 1;
        while (drained < n_chains) { drain(drained & 1); ++drained; }
    }

    // ---- epilogue on the register sums ----
    float* Cb = g.C + g.batch_offC[batch] + (long)split * g.split_stride;
    const float* Mb = g.mask ? g.mask + g.batch_offM[batch] : nullptr;
    const bool final_pass = (g.splits == 1);
    const int il = 32 * (warp & 3) + lane;
    const unsigned char fl = rflag[il];
    if (fl != 2) {
#pragma unroll
        for (int n = 0; n < HALF; ++n) {
            const int jl = (warp >> 2) * HALF + n;
            if (!cvalid[jl]) continue;
            if (fl == 1) {
                g.C_ones[(long)split * g.split_stride + (j0 + jl) * g.ones_stride] = acc[n] * g.alpha;
                continue;
            }
            float x = acc[n] * (ADT == 1 ? g.alpha / 255.0f : g.alpha);
            float* dst = Cb + rC[il] + cC[jl];
            if (final_pass) {
                if (g.bias_mode == 1) x += __ldg(g.bias + j0 + jl);
                else if (g.bias_mode == 2) x += __ldg(g.bias + i0 + il);
                if (g.act == 1) x = fmaxf(x, 0.f);
                else if (g.act == 2) x = tanhf(x);
                if (g.mask_mode == 1) x = (__ldg(Mb + rM[il] + cM[jl]) > 0.f) ? x : 0.f;
                else if (g.mask_mode == 2) { const float t = __ldg(Mb + rM[il] + cM[jl]); x *= (1.f - t * t); }
                if (g.accumulate) x += *dst;
            }
            *dst = x;
        }
    }

    asm volatile("tcgen05.fence::before_thread_sync;" ::: "memory");
    __syncthreads();
    if (warp == 0) {
        asm volatile("tcgen05.dealloc.cta_group::1.sync.aligned.b32 %0, %1;" ::"r"(tmem_d),
                     "r"((uint32_t)(2 * BN)));
    }
}

template <int BN>
static int launch_tc(const GemmDesc& g, cudaStream_t st) {
    constexpr int smem = TSTAGES * (2 * TBM * TBK * 4 + 2 * BN * TBK * 4);
    const long Mtot = g.M + (g.ones_row ? 1 : 0);
    dim3 grid((unsigned)((g.N + BN - 1) / BN), (unsigned)((Mtot + TBM - 1) / TBM), (unsigned)(g.batch * g.splits));
    static bool attr_set[3] = {false, false, false};
    if (!attr_set[g.a_dtype]) {
        cudaError_t e;
        if (g.a_dtype == 1) e = cudaFuncSetAttribute(gemm_tc_kernel<BN, 1>, cudaFuncAttributeMaxDynamicSharedMemorySize, smem);
        else if (g.a_dtype == 2) e = cudaFuncSetAttribute(gemm_tc_kernel<BN, 2>, cudaFuncAttributeMaxDynamicSharedMemorySize, smem);
        else e = cudaFuncSetAttribute(gemm_tc_kernel<BN, 0>, cudaFuncAttributeMaxDynamicSharedMemorySize, smem);
        if (e != cudaSuccess) { set_error("gemm_tc: cudaFuncSetAttribute: %s", cudaGetErrorString(e)); return A2CB_ERR_CUDA; }
        attr_set[g.a_dtype] = true;
    }
    if (g.a_dtype == 1) gemm_tc_kernel<BN, 1><<<grid, TNT, smem, st>>>(g);
    else if (g.a_dtype == 2) gemm_tc_kernel<BN, 2><<<grid, TNT, smem, st>>>(g);
    else gemm_tc_kernel<BN, 0><<<grid, TNT, smem, st>>>(g);
    return check_launch("gemm_tc");
}

// ---------------------------------------------------------------------------
// B image: weights are the same for every CTA of a GEMM and change once per optimiser step, so their
// gather (PyTorch [out, in, kh, kw] order -> engine K order), hi/lo split and UMMA tiling are done ONCE
// here instead of in every CTA and every k-block.
// ---------------------------------------------------------------------------
template <int BN>
__global__ void __launch_bounds__(256)
gemm_prep_b_kernel(const GemmDesc g, float* __restrict__ img) {
    const long ktiles = (g.K + TBK - 1) / TBK;
    const long n_ctiles = (g.N + BN - 1) / BN;
    const long chunks = (long)g.batch * n_ctiles * ktiles * BN * TKC;       // 16-byte (row, kc) chunks
    const long e = (long)blockIdx.x * blockDim.x + threadIdx.x;
    if (e >= chunks) return;
    const int kc = (int)(e % TKC);
    const int col = (int)((e / TKC) % BN);
    const long blk = e / ((long)TKC * BN);                                   // (batch, ctile, kb)
    const long kb = blk % ktiles;
    const long ct = (blk / ktiles) % n_ctiles;
    const int z = (int)(blk / (ktiles * n_ctiles));
    const long j = ct * BN + col;
    const float* Bb = g.B + g.batch_offB[z];
    float v[4] = {0.f, 0.f, 0.f, 0.f};
    if (j < g.N) {
        const long cb = tc_affine(g.colB, j, nullptr);
#pragma unroll
        for (int c = 0; c < 4; ++c) {
            const long r = kb * TBK + kc * 4 + c;
            if (r < g.K) v[c] = __ldg(Bb + tc_affine(g.redB, r, nullptr) + cb);
        }
    }
    float4 hi, lo;
    split4(make_float4(v[0], v[1], v[2], v[3]), hi, lo);
    float* dst = img + blk * (2 * BN * TBK);
    const uint32_t o = tile_off(col, kc) / 4;
    *reinterpret_cast<float4*>(dst + o) = hi;
    *reinterpret_cast<float4*>(dst + BN * TBK + o) = lo;
}

static int tc_bn(const GemmDesc& g) { return g.N <= 32 ? 32 : (g.N <= 64 ? 64 : 128); }

long gemm_b_image_floats(const GemmDesc& g) {
    const int bn = tc_bn(g);
    return (long)g.batch * ((g.N + bn - 1) / bn) * ((g.K + TBK - 1) / TBK) * 2 * bn * TBK;
}

int gemm_prep_b(const GemmDesc& g, cudaStream_t st) {
    A2CB_REQUIRE(g.B && g.B_image, "gemm_prep_b: B and B_image required");
    const int bn = tc_bn(g);
    const long chunks = gemm_b_image_floats(g) / 8;
    if (chunks == 0) return A2CB_OK;
    const unsigned grid = (unsigned)((chunks + 255) / 256);
    float* img = const_cast<float*>(g.B_image);
    if (bn == 32) gemm_prep_b_kernel<32><<<grid, 256, 0, st>>>(g, img);
    else if (bn == 64) gemm_prep_b_kernel<64><<<grid, 256, 0, st>>>(g, img);
    else gemm_prep_b_kernel<128><<<grid, 256, 0, st>>>(g, img);
    return check_launch("gemm_prep_b");
}

// Shapes the tensor-core path accepts; everything else stays on the fp32 SIMT engine.
bool gemm_tc_supported(const GemmDesc& g) {
    if (g.vecA == 2 && g.M % 4 != 0) return false;
    if (g.vecB == 1 && g.N % 4 != 0) return false;
    return true;
}

int gemm_tc(const GemmDesc& g, cudaStream_t st) {
    A2CB_REQUIRE(gemm_tc_supported(g), "gemm_tc: unsupported operand vectorisation");
    if (g.M + g.ones_row == 0 || g.N == 0) return A2CB_OK;
    if (g.N <= 32) return launch_tc<32>(g, st);
    if (g.N <= 64) return launch_tc<64>(g, st);
    return launch_tc<128>(g, st);
}

}  // namespace a2cb
