// GRU recurrence (reference a2c_ppo_acktr/model.py:111-166) on tcgen05: the whole [T, N] sequence in ONE cooperative
// launch per direction, the per-step product on the 5th-generation tensor cores.
//
// The mma.sync kernels of gru_persistent.cu spend ~1,100 instructions per warp and time step on operand handling:
// fp32 -> TF32 hi/lo splits of the RESIDENT weights (which never change during the launch), scalar loads of the
// exchanged state, a shared-memory fold of eight partial tiles.  Here
//   * the CTA's slice of W_hh (the 96 gate rows of its 32 hidden units) is converted ONCE, at kernel start, into two
//     fp16 planes w = (w_hi + w_lo) / s_W (scaled by a power of two so that w_lo stays a normal fp16: 22 significant
//     bits, the same 4 bytes per element as fp32) laid out as the canonical 128-byte-swizzled UMMA operand.  The SAME
//     image serves both directions: K-major A operand [gate rows x units] forward, MN-major A operand
//     [units x gate rows] backward (the 8-row x 128-byte swizzle atom is identical for both views);
//   * the exchanged operand of a step (forward: the masked state of the group's 8 environments; backward: the CTA's own
//     gate gradients) is written by its producers as fp16 hi / lo planes ALREADY in that shared-memory image layout, so a
//     consumer CTA fetches the whole state of a step with ONE bulk copy (cp.async.bulk global -> shared, completion on
//     an mbarrier) and one elected thread issues the 96 (forward) / 72 (backward) tcgen05.mma.kind::f16 of the step:
//     x w = x_hi w_hi + x_lo w_hi + x_hi w_lo with fp32 accumulation in TMEM, split over four accumulators so that the
//     tensor core's truncating accumulation sees chains of at most 24 instructions;
//   * forward: all-gather of the state through L2 (each CTA contributes 32 units, reads 512); backward: every CTA
//     multiplies ITS gate gradients (K = 96) by its weight slice for ALL units and the 16 partial products are summed
//     by their owners (reduce-scatter through L2, fixed order: deterministic) -- no gate gradient ever needs to be
//     gathered before the product;
//   * one release / acquire counter per environment group and step, as before; environments are independent.
// Ownership, saved quantities, buffers and the gate arithmetic are those of gru_persistent.cu, so either path feeds
// the same weight-gradient products and the same tests.
#include <cuda_fp16.h>
#include <stdlib.h>

#include "common.cuh"
#include "decls.cuh"
#include "gru.cuh"

namespace a2cb {

namespace {

constexpr int GT_NT = 256;                    // 8 warps: thread (warp e, lane u) owns (environment e, unit u)
constexpr int GT_RW = 8;                      // environments per group
constexpr int GT_UB = 32;                     // hidden units per CTA
constexpr int GT_ROWS = 3 * GT_UB;            // gate rows per CTA: row g * 32 + u
constexpr int GT_WBLK = GT_ROWS * 128;        // bytes of one 64-unit block of one weight plane
constexpr int GT_MAX_H = 512;
constexpr int GT_MAX_GROUPS = 64;             // environment groups per launch (N <= 512)
constexpr int GT_IMG_BYTES = 2 * (GT_MAX_H / 64) * 1024;              // state image of one group and step (both planes)
constexpr int GT_PART_FLOATS = (GT_MAX_H / 32) * (GT_MAX_H / 32) * GT_RW * GT_UB;   // partial carries of one group and step
constexpr int GT_CTRL = 4;                    // control warp: polls, issues the bulk copy and the MMAs
constexpr float GT_SX = 16.f;                 // power-of-two scale of the forward state planes

__device__ unsigned gt_flags[2][GT_MAX_GROUPS];
__device__ long long gt_dbg[2][16][16];            // A2CB_GRU_TC_DEBUG=1: clock64 stamps of CTA (0, 0), steps 16..31, per phase
#define GT_STAMP(dir, t, i) do { if (dbg && blockIdx.x == 0 && blockIdx.y == 0 && (t) >= 16 && (t) < 32) gt_dbg[dir][(t) - 16][i] = clock64(); } while (0)
__device__ uint4 gt_img[(size_t)GT_MAX_GROUPS * 2 * GT_IMG_BYTES / 16];
__device__ float gt_part[(size_t)GT_MAX_GROUPS * 2 * GT_PART_FLOATS];

__device__ __forceinline__ uint32_t gt_smem(const void* p) { return (uint32_t)__cvta_generic_to_shared(p); }
__device__ __forceinline__ float gt_sigmoid(float x) { return 1.f / (1.f + expf(-x)); }
__device__ __forceinline__ float gt_mask(const GruDesc& d, int t, int n) {
    const long row = (long)t * d.N + n;
    const long src = d.idx ? (long)d.idx[row] : row;
    return d.masks[src];
}

__device__ __forceinline__ void gt_mbar_init(uint64_t* bar, uint32_t count) {
    asm volatile("mbarrier.init.shared::cta.b64 [%0], %1;" ::"r"(gt_smem(bar)), "r"(count));
}
__device__ __forceinline__ void gt_mbar_wait(uint64_t* bar, uint32_t parity) {
    uint32_t ok = 0;
    while (!ok) {
        asm volatile(
            "{\n"
            ".reg .pred p;\n"
            "mbarrier.try_wait.parity.shared::cta.b64 p, [%1], %2;\n"
            "selp.u32 %0, 1, 0, p;\n"
            "}\n"
            : "=r"(ok)
            : "r"(gt_smem(bar)), "r"(parity)
            : "memory");
    }
}
__device__ __forceinline__ void gt_expect(uint64_t* bar, uint32_t bytes) {
    asm volatile("mbarrier.arrive.expect_tx.shared::cta.b64 _, [%0], %1;" ::"r"(gt_smem(bar)), "r"(bytes) : "memory");
}
__device__ __forceinline__ void gt_bulk_g2s(uint32_t dst, const void* src, uint32_t bytes, uint64_t* bar) {
    asm volatile("cp.async.bulk.shared::cluster.global.mbarrier::complete_tx::bytes [%0], [%1], %2, [%3];" ::"r"(dst),
                 "l"(src), "r"(bytes), "r"(gt_smem(bar))
                 : "memory");
}
__device__ __forceinline__ void gt_commit(uint64_t* bar) {
    asm volatile("tcgen05.commit.cta_group::1.mbarrier::arrive::one.shared::cluster.b64 [%0];" ::"r"(gt_smem(bar)) : "memory");
}
__device__ __forceinline__ void gt_mma(uint32_t d_tmem, uint64_t a, uint64_t b, uint32_t idesc, uint32_t acc) {
    asm volatile(
        "{\n\t"
        ".reg .pred p;\n\t"
        "setp.ne.b32 p, %4, 0;\n\t"
        "tcgen05.mma.cta_group::1.kind::f16 [%0], %1, %2, %3, p;\n\t"
        "}\n" ::"r"(d_tmem), "l"(a), "l"(b), "r"(idesc), "r"(acc)
        : "memory");
}
// A operand from tensor memory (lane = row, one 32-bit column = two consecutive fp16 reduction elements), B from shared memory
__device__ __forceinline__ void gt_mma_ts(uint32_t d_tmem, uint32_t a_tmem, uint64_t b, uint32_t idesc, uint32_t acc) {
    asm volatile(
        "{\n\t"
        ".reg .pred p;\n\t"
        "setp.ne.b32 p, %4, 0;\n\t"
        "tcgen05.mma.cta_group::1.kind::f16 [%0], [%1], %2, %3, p;\n\t"
        "}\n" ::"r"(d_tmem), "r"(a_tmem), "l"(b), "r"(idesc), "r"(acc)
        : "memory");
}
__device__ __forceinline__ void gt_st8(uint32_t taddr, const uint32_t* r) {
    asm volatile("tcgen05.st.sync.aligned.32x32b.x8.b32 [%0], {%1, %2, %3, %4, %5, %6, %7, %8};\n" ::"r"(taddr), "r"(r[0]), "r"(r[1]),
                 "r"(r[2]), "r"(r[3]), "r"(r[4]), "r"(r[5]), "r"(r[6]), "r"(r[7])
                 : "memory");
}
// shared-memory matrix descriptor: start >> 4 | LBO >> 4 << 16 | SBO >> 4 << 32 | version 1 << 46 | SWIZZLE_128B << 61
__device__ __forceinline__ uint64_t gt_desc(uint32_t addr, uint32_t lbo_bytes, uint32_t sbo_bytes) {
    return (uint64_t)((addr & 0x3FFFF) >> 4) | ((uint64_t)((lbo_bytes >> 4) & 0x3FFF) << 16) |
           ((uint64_t)((sbo_bytes >> 4) & 0x3FFF) << 32) | (1ull << 46) | (2ull << 61);
}
__device__ __forceinline__ bool gt_elect() {
    uint32_t pred;
    asm volatile(
        "{\n\t"
        ".reg .pred P1;\n\t"
        "elect.sync _|P1, 0xFFFFFFFF;\n\t"
        "selp.u32 %0, 1, 0, P1;\n\t"
        "}\n"
        : "=r"(pred));
    return pred != 0;
}
__device__ __forceinline__ void gt_ld8(uint32_t taddr, float* v) {
    uint32_t r[8];
    asm volatile("tcgen05.ld.sync.aligned.32x32b.x8.b32 {%0, %1, %2, %3, %4, %5, %6, %7}, [%8];\n"
                 : "=r"(r[0]), "=r"(r[1]), "=r"(r[2]), "=r"(r[3]), "=r"(r[4]), "=r"(r[5]), "=r"(r[6]), "=r"(r[7])
                 : "r"(taddr));
    asm volatile("tcgen05.wait::ld.sync.aligned;" ::: "memory");
#pragma unroll
    for (int i = 0; i < 8; ++i) v[i] = __uint_as_float(r[i]);
}
__device__ __forceinline__ unsigned gt_ld_acquire(const unsigned* p) {
    unsigned v;
    asm volatile("ld.acquire.gpu.global.u32 %0, [%1];" : "=r"(v) : "l"(p) : "memory");
    return v;
}
__device__ __forceinline__ void gt_release_add(unsigned* p) {
    asm volatile("red.release.gpu.global.add.u32 [%0], 1;" ::"l"(p) : "memory");
}
// x * s = hi + lo, both fp16 (hi the nearest fp16, lo the nearest fp16 of the exact remainder): packed {lo:hi}
__device__ __forceinline__ uint32_t gt_split(float xs) {
    const __half hi = __float2half_rn(xs);
    const __half lo = __float2half_rn(xs - __half2float(hi));
    return (uint32_t)__half_as_ushort(hi) | ((uint32_t)__half_as_ushort(lo) << 16);
}
// byte offset of element (row r, reduction index k) inside a K-major 128-byte-swizzled plane whose 64-element
// blocks are `blk` bytes apart (rows of 128 bytes, 16-byte chunk index xor (r & 7))
__device__ __forceinline__ uint32_t gt_swz(int r, int k, int blk) {
    return (uint32_t)((k >> 6) * blk + r * 128 + ((((k & 63) >> 3) ^ (r & 7)) << 4) + ((k & 7) << 1));
}
// power-of-two scale that maps `amax` into [2^e, 2^(e+1))
__device__ __forceinline__ float gt_pow2_scale(float amax, int e) {
    if (!(amax > 0.f) || !isfinite(amax)) return 1.f;
    int ex;
    frexpf(amax, &ex);                                    // amax = m * 2^ex, m in [0.5, 1)
    return ldexpf(1.f, e + 1 - ex);
}

// Converts the CTA's weight slice (rows g * H + ub + u, g < 3, u < 32, all H columns) into the two fp16 planes of the
// UMMA image at `img` (plane stride = (H / 64) * GT_WBLK); returns the scale (w * scale = hi + lo).
// `planes`: 3 = both planes (plane 1 behind plane 0), 2 = the lo plane only (at img), 0 = none (scale only).
__device__ float gt_build_weights(uint8_t* img, const float* __restrict__ Whh, int H, int ub, float* red, int planes = 3) {
    const int tid = threadIdx.x;
    float amax = 0.f;
    for (int e = tid; e < GT_ROWS * H; e += GT_NT) {
        const int row = e / H, k = e - row * H;
        amax = fmaxf(amax, fabsf(Whh[(long)((row >> 5) * H + ub + (row & 31)) * H + k]));
    }
    amax = warp_max(amax);
    __syncthreads();
    if ((tid & 31) == 0) red[tid >> 5] = amax;
    __syncthreads();
    amax = red[0];
#pragma unroll
    for (int w = 1; w < GT_NT / 32; ++w) amax = fmaxf(amax, red[w]);
    const float sw = gt_pow2_scale(amax, 13);             // |w| * sw < 2^14: w_lo is a normal fp16 down to |w| sw 2^-11 >= 2^-14
    const uint32_t plane = (uint32_t)(H / 64) * GT_WBLK;
    if (planes)
        for (int e = tid; e < GT_ROWS * H; e += GT_NT) {
            const int row = e / H, k = e - row * H;
            const uint32_t p = gt_split(Whh[(long)((row >> 5) * H + ub + (row & 31)) * H + k] * sw);
            const uint32_t off = gt_swz(row, k, GT_WBLK);
            if (planes == 3) {
                *reinterpret_cast<unsigned short*>(img + off) = (unsigned short)(p & 0xffffu);
                *reinterpret_cast<unsigned short*>(img + plane + off) = (unsigned short)(p >> 16);
            } else {
                *reinterpret_cast<unsigned short*>(img + off) = (unsigned short)(p >> 16);
            }
        }
    __syncthreads();
    return sw;
}

}  // namespace

// ---------------------------------------------------------------------------------------------------------------
// forward.  grid (H / 32 unit groups, resident environment groups); CTA (bx, by) owns units 32 bx .. 32 bx + 31 and
// the environment groups by, by + gridDim.y, ...  Shared memory: [state image][weight image][tail].
// ---------------------------------------------------------------------------------------------------------------
// TA: the hi plane of the weight slice lives in TENSOR MEMORY (128 lanes x H / 2 columns behind the accumulators) and is
// the A operand of two of the three products of every reduction step; only the lo plane stays in shared memory.  With
// both planes in shared memory every MMA streams 4 KiB of A through the shared-memory port (~59 cycles per instruction
// measured, 5,700 of the 10,200 cycles of a step).
template <bool TA>
__global__ void __launch_bounds__(GT_NT, 1)
gru_tc_fwd_kernel(const GruDesc d, const float* __restrict__ Whh, const float* __restrict__ bhh, const int dbg) {
    extern __shared__ uint8_t gt_dyn[];
    __shared__ uint64_t bar_state[4], bar_acc;
    __shared__ uint32_t tmem_slot;
    __shared__ float red[GT_NT / 32];
    const int H = d.H, nkb = H >> 6, ug = H >> 5;
    const uint32_t base = (gt_smem(gt_dyn) + 1023u) & ~1023u;
    uint8_t* sm = gt_dyn + (base - gt_smem(gt_dyn));
    const uint32_t st_bytes = 2u * nkb * 1024u;           // state image: [2 planes][nkb][8 rows x 128 B]
    const uint32_t w_plane = (uint32_t)nkb * GT_WBLK;
    const uint32_t s_state = base, s_w = base + st_bytes;
    constexpr int NPL = TA ? 1 : 2;                                        // weight planes in shared memory
    constexpr uint32_t TCOLS = TA ? 512u : 64u;                            // tensor memory: 64 accumulator columns (+ H / 2)
    float* xchg = reinterpret_cast<float*>(sm + st_bytes + NPL * w_plane);    // [3 gates][8 environments][32 units]
    const int tid = threadIdx.x, lane = tid & 31, warp = tid >> 5;
    const int ub = blockIdx.x * GT_UB, ku = ub + lane;
    const int ngrp = (d.N + GT_RW - 1) / GT_RW;
    const bool single = ngrp <= (int)gridDim.y;

    if (tid == 0) {
        for (int c = 0; c < 4; ++c) gt_mbar_init(&bar_state[c], 1);
        gt_mbar_init(&bar_acc, 1);
        asm volatile("fence.mbarrier_init.release.cluster;" ::: "memory");
    }
    if (warp == GT_CTRL) {
        asm volatile("tcgen05.alloc.cta_group::1.sync.aligned.shared::cta.b32 [%0], %1;" ::"r"(gt_smem(&tmem_slot)), "r"(TCOLS));
        asm volatile("tcgen05.relinquish_alloc_permit.cta_group::1.sync.aligned;" ::);
    }
    const float sw = gt_build_weights(sm + st_bytes, Whh, H, ub, red, TA ? 2 : 3);
    const float inv_scale = 1.f / (sw * GT_SX);
    asm volatile("fence.proxy.async.shared::cta;" ::: "memory");          // weight image: generic stores -> tensor-core reads
    asm volatile("tcgen05.fence::before_thread_sync;" ::: "memory");
    __syncthreads();
    asm volatile("tcgen05.fence::after_thread_sync;" ::: "memory");
    const uint32_t tmem = tmem_slot;
    if (TA) {
        // hi plane -> tensor memory: thread (warp w < 4, lane) owns row 32 w + lane (rows >= 96: zero); column
        // 64 + c holds reduction elements 2 c, 2 c + 1 of that row
        if (warp < 4) {
            const int row = 32 * warp + lane;
            const float* wr = Whh + (long)((row >> 5) * H + ub + (row & 31)) * H;       // rows >= 96 never dereferenced
            for (int c0 = 0; c0 < H / 2; c0 += 8) {
                uint32_t r[8];
#pragma unroll
                for (int j = 0; j < 8; ++j) {
                    uint32_t lo16 = 0u, hi16 = 0u;
                    if (row < GT_ROWS) {
                        lo16 = gt_split(wr[2 * (c0 + j)] * sw) & 0xffffu;
                        hi16 = gt_split(wr[2 * (c0 + j) + 1] * sw) & 0xffffu;
                    }
                    r[j] = lo16 | (hi16 << 16);
                }
                gt_st8(tmem + ((uint32_t)(32 * warp) << 16) + 64u + (uint32_t)c0, r);
            }
            asm volatile("tcgen05.wait::st.sync.aligned;" ::: "memory");
        }
        asm volatile("tcgen05.fence::before_thread_sync;" ::: "memory");
        __syncthreads();
        asm volatile("tcgen05.fence::after_thread_sync;" ::: "memory");
    }
    const float bias = warp < 3 ? bhh[warp * H + ku] : 0.f;               // drain warp g adds b_hh of gate g
    // the state of a step arrives in NCH chunks of k-blocks (one mbarrier each): the MMAs of accumulator c only wait
    // for chunk c, so most of the fetch hides behind the MMAs of the earlier chunks
    const int NCH = (nkb % 4 == 0) ? 4 : 1;
    const uint32_t ch_bytes = (uint32_t)(nkb / NCH) * 1024u;               // per plane and chunk

    // instruction descriptor: D fp32, A / B fp16, both K-major, N = 16, M = 128
    const uint32_t idesc = (1u << 4) | ((uint32_t)(16 >> 3) << 17) | ((uint32_t)(128 >> 4) << 24);
    const int k16 = H >> 4, per_acc = k16 >> 2;                            // 16-element reduction steps, steps per accumulator

    // writes this thread's masked state v of (environment `warp`, unit `lane`) into the group's image `img`
    auto put_state = [&](uint8_t* img, float v) {
        const uint32_t p = gt_split(v * GT_SX);
        const uint32_t q = __shfl_xor_sync(0xffffffffu, p, 1);
        // even lane: hi-plane word {hi(u), hi(u + 1)}; odd lane: lo-plane word {lo(u - 1), lo(u)}
        const uint32_t w0 = (lane & 1) ? ((q >> 16) | (p & 0xffff0000u)) : ((p & 0xffffu) | (q << 16));
        const uint32_t w1 = __shfl_down_sync(0xffffffffu, w0, 2), w2 = __shfl_down_sync(0xffffffffu, w0, 4),
                       w3 = __shfl_down_sync(0xffffffffu, w0, 6);
        if ((lane & 7) < 2) {                                             // lane 8 j: hi chunk of units 8 j .. 8 j + 7; 8 j + 1: lo chunk
            const int k = ub + (lane & ~7);
            const uint32_t off = (uint32_t)(lane & 1) * (uint32_t)nkb * 1024u + gt_swz(warp, k, 1024);
            *reinterpret_cast<uint4*>(img + off) = make_uint4(w0, w1, w2, w3);
        }
    };
    auto image_of = [&](int grp, int buf) {
        return reinterpret_cast<uint8_t*>(gt_img) + ((size_t)grp * 2 + buf) * GT_IMG_BYTES;
    };

    // ---- image of step 0: hm_0 = h0 * mask_0 ----
    float hm_cur = 0.f;                                                    // this thread's hm_t (single pass: kept in a register)
    for (int grp = blockIdx.y; grp < ngrp; grp += gridDim.y) {
        const int my = grp * GT_RW + warp;
        float v = 0.f;
        if (my < d.N) {
            v = d.h0[(long)my * H + ku] * gt_mask(d, 0, my);
            d.hm[(long)my * H + ku] = v;
        }
        put_state(image_of(grp, 0), v);
        hm_cur = v;
        __syncthreads();
        if (tid == 0) gt_release_add(&gt_flags[0][grp]);
    }

    struct GateIn { float r, z, n, mnext; };
    auto load_gi = [&](int t, int my) {
        GateIn q = {0.f, 0.f, 0.f, 0.f};
        if (my < d.N) {
            const float* gi = d.gi + ((long)t * d.N + my) * 3 * H;
            q.r = __ldg(gi + ku); q.z = __ldg(gi + H + ku); q.n = __ldg(gi + 2 * H + ku);
            q.mnext = (t + 1 < d.T) ? gt_mask(d, t + 1, my) : 0.f;
        }
        return q;
    };
    GateIn pre = load_gi(0, blockIdx.y * GT_RW + warp);
    uint32_t it = 0;                                                       // (step, pass) counter: barrier parities

    for (int t = 0; t < d.T; ++t) {
        for (int grp = blockIdx.y; grp < ngrp; grp += gridDim.y, ++it) {
            const int my = grp * GT_RW + warp;
            const bool live = my < d.N;
            const long row = (long)t * d.N + my;
            const GateIn q = single ? pre : load_gi(t, my);
            const float hprev = single ? hm_cur : (live ? d.hm[row * H + ku] : 0.f);
            if (warp == GT_CTRL) {
                // ===== control warp: wait for the group's state of step t, fetch it, issue the step's MMAs =====
                if (lane == 0) {
                    GT_STAMP(0, t, 0);
                    const unsigned target = (unsigned)ug * (unsigned)(t + 1);
                    while (gt_ld_acquire(&gt_flags[0][grp]) < target) {}
                    GT_STAMP(0, t, 1);
                }
                __syncwarp();
                if (gt_elect()) {
                    asm volatile("fence.proxy.async;" ::: "memory");     // other CTAs' generic stores -> these bulk copies
                    const uint8_t* src = image_of(grp, t & 1);
                    for (int c = 0; c < NCH; ++c) {
                        gt_expect(&bar_state[c], 2 * ch_bytes);
                        gt_bulk_g2s(s_state + c * ch_bytes, src + c * ch_bytes, ch_bytes, &bar_state[c]);
                        gt_bulk_g2s(s_state + nkb * 1024u + c * ch_bytes, src + nkb * 1024u + c * ch_bytes, ch_bytes, &bar_state[c]);
                    }
                }
                __syncwarp();
                for (int c = 0; c < NCH; ++c) {
                    gt_mbar_wait(&bar_state[c], it & 1u);
                    if (lane == 0 && c == 0) GT_STAMP(0, t, 2);
                    asm volatile("tcgen05.fence::after_thread_sync;" ::: "memory");
                    if (gt_elect()) {
                        const int s0 = c * (k16 / NCH), s1 = s0 + k16 / NCH;
                        for (int s = s0; s < s1; ++s) {
                            const uint32_t o = (uint32_t)(s >> 2) * GT_WBLK + (uint32_t)(s & 3) * 32u;
                            const uint32_t ob = (uint32_t)(s >> 2) * 1024u + (uint32_t)(s & 3) * 32u;
                            // B: 16 "environment" rows = the group's 8 (first atom) + 8 junk rows: the second atom is made
                            // to alias the start of the weight image (finite values) through SBO = size of the state image
                            const uint64_t bh = gt_desc(s_state + ob, 16, st_bytes), bl = gt_desc(s_state + nkb * 1024u + ob, 16, st_bytes);
                            const uint32_t dcol = tmem + (uint32_t)(s / per_acc) * 16u;
                            const uint32_t first = (s % per_acc) == 0 ? 0u : 1u;
                            if (TA) {
                                const uint64_t al = gt_desc(s_w + o, 16, 1024);
                                const uint32_t ah = tmem + 64u + 8u * (uint32_t)s;
                                gt_mma(dcol, al, bh, idesc, first);
                                gt_mma_ts(dcol, ah, bl, idesc, 1u);
                                gt_mma_ts(dcol, ah, bh, idesc, 1u);
                            } else {
                                const uint64_t ah = gt_desc(s_w + o, 16, 1024), al = gt_desc(s_w + w_plane + o, 16, 1024);
                                gt_mma(dcol, al, bh, idesc, first);
                                gt_mma(dcol, ah, bl, idesc, 1u);
                                gt_mma(dcol, ah, bh, idesc, 1u);
                            }
                        }
                        if (c == NCH - 1) gt_commit(&bar_acc);
                    }
                    __syncwarp();
                }
                if (lane == 0) GT_STAMP(0, t, 3);
            } else if (warp < 3) {
                // ===== drain: gate `warp`, row = lane: 4 accumulators x 8 environments =====
                gt_mbar_wait(&bar_acc, it & 1u);
                if (tid == 0) GT_STAMP(0, t, 4);
                asm volatile("tcgen05.fence::after_thread_sync;" ::: "memory");
                float s8[8] = {0.f, 0.f, 0.f, 0.f, 0.f, 0.f, 0.f, 0.f};
#pragma unroll
                for (int a = 0; a < 4; ++a) {
                    float v[8];
                    gt_ld8(tmem + ((uint32_t)(32 * warp) << 16) + (uint32_t)a * 16u, v);
#pragma unroll
                    for (int e = 0; e < 8; ++e) s8[e] += v[e];
                }
#pragma unroll
                for (int e = 0; e < 8; ++e) xchg[(warp * 8 + e) * 32 + lane] = s8[e] * inv_scale + bias;
                asm volatile("tcgen05.fence::before_thread_sync;" ::: "memory");
                if (tid == 0) GT_STAMP(0, t, 5);
            }
            __syncthreads();                                               // gate pre-activations of W_hh hm_t + b_hh are in xchg
            if (tid == 0) GT_STAMP(0, t, 6);
            const float ghr = xchg[(0 * 8 + warp) * 32 + lane], ghz = xchg[(1 * 8 + warp) * 32 + lane],
                        ghn = xchg[(2 * 8 + warp) * 32 + lane];
            const float rr = gt_sigmoid(q.r + ghr);
            const float zz = gt_sigmoid(q.z + ghz);
            const float nn = tanhf(q.n + rr * ghn);
            const float h = (1.f - zz) * nn + zz * hprev;
            const float hm_next = live ? h * q.mnext : 0.f;
            if (t + 1 < d.T) put_state(image_of(grp, (t + 1) & 1), hm_next);    // what the other CTAs wait for
            hm_cur = hm_next;
            if (tid == 0) GT_STAMP(0, t, 7);
            __syncthreads();
            if (tid == 0) GT_STAMP(0, t, 8);
            if (tid == 0 && t + 1 < d.T) gt_release_add(&gt_flags[0][grp]);
            if (tid == 0) GT_STAMP(0, t, 9);
            // everything below is consumed by later kernels only
            if (single && t + 1 < d.T) pre = load_gi(t + 1, my);
            if (live) {
                d.out[row * H + ku] = h;
                if (d.r) {
                    d.r[row * H + ku] = rr; d.z[row * H + ku] = zz; d.n[row * H + ku] = nn;
                    d.gh[row * 3 * H + 2 * H + ku] = ghn;                  // the only part of gh the backward pass reads
                }
                if (t + 1 < d.T) d.hm[(row + d.N) * H + ku] = hm_next;
            }
        }
    }
    asm volatile("tcgen05.fence::before_thread_sync;" ::: "memory");
    __syncthreads();
    if (warp == GT_CTRL)
        asm volatile("tcgen05.dealloc.cta_group::1.sync.aligned.b32 %0, %1;" ::"r"(tmem), "r"(TCOLS));
}

// ---------------------------------------------------------------------------------------------------------------
// backward through time.  Same grid and ownership.  Per step: (A) element-wise gate gradients of the CTA's units
// (needs the carry = dh z + the 16 partial products of the previous step that belong to its units), (B) the CTA's 96
// gate-gradient rows x its weight slice -> partial carries of ALL units (M = units, K = 96), written to the owners.
// Shared memory: [weight image][tail][gate-gradient operand: 2 planes x 2 blocks x 16 rows x 128 B].
// ---------------------------------------------------------------------------------------------------------------
// TA: both planes of the TRANSPOSED weight slice (M = units, K = the CTA's 96 gate rows) live in tensor memory: per
// 128-unit block and plane 48 columns (column c = gate rows 2 c, 2 c + 1), 384 columns behind the 64 accumulator columns
// for H = 512 -- no weight image in shared memory at all, every MMA reads only the 512-byte gate-gradient operand from it.
template <bool TA>
__global__ void __launch_bounds__(GT_NT, 1)
gru_tc_bwd_kernel(const GruDesc d, const float* __restrict__ Whh, float* __restrict__ carry_ws, const int swap_strides,
                  const int dbg) {
    extern __shared__ uint8_t gt_dyn[];
    __shared__ uint64_t bar_acc;
    __shared__ uint32_t tmem_slot;
    __shared__ float red[GT_NT / 32];
    const int H = d.H, nkb = H >> 6, ug = H >> 5, n_mb = H >> 7;
    const uint32_t base = (gt_smem(gt_dyn) + 1023u) & ~1023u;
    uint8_t* sm = gt_dyn + (base - gt_smem(gt_dyn));
    const uint32_t w_plane = (uint32_t)nkb * GT_WBLK;
    const uint32_t g_off = TA ? 0u : 2 * w_plane + 4096u;                  // gate-gradient operand (behind the image + 4 KiB tail)
    const uint32_t s_w = base, s_g = base + g_off;
    uint8_t* gop = sm + g_off;
    constexpr uint32_t G_PLANE = 2 * 2048;                                 // [2 blocks of 64 rows-of-K][16 rows x 128 B]
    constexpr uint32_t TCOLS = TA ? 512u : 64u;
    const int tid = threadIdx.x, lane = tid & 31, warp = tid >> 5;
    const int ub = blockIdx.x * GT_UB, ku = ub + lane;
    const int ngrp = (d.N + GT_RW - 1) / GT_RW;

    if (tid == 0) {
        gt_mbar_init(&bar_acc, 1);
        asm volatile("fence.mbarrier_init.release.cluster;" ::: "memory");
    }
    if (warp == GT_CTRL) {
        asm volatile("tcgen05.alloc.cta_group::1.sync.aligned.shared::cta.b32 [%0], %1;" ::"r"(gt_smem(&tmem_slot)), "r"(TCOLS));
        asm volatile("tcgen05.relinquish_alloc_permit.cta_group::1.sync.aligned;" ::);
    }
    const float sw = gt_build_weights(sm, Whh, H, ub, red, TA ? 0 : 3);
    for (int i = tid; i < (int)(2 * G_PLANE / 16); i += GT_NT) reinterpret_cast<uint4*>(gop)[i] = make_uint4(0u, 0u, 0u, 0u);
    asm volatile("fence.proxy.async.shared::cta;" ::: "memory");
    asm volatile("tcgen05.fence::before_thread_sync;" ::: "memory");
    __syncthreads();
    asm volatile("tcgen05.fence::after_thread_sync;" ::: "memory");
    const uint32_t tmem = tmem_slot;
    if (TA) {
        // thread (warp w < 4, lane) owns unit 128 mb + 32 w + lane of every block mb; gate row j <-> weight row
        // (j / 32) H + ub + j % 32: for a fixed j the lanes read consecutive columns (coalesced)
        if (warp < 4) {
            for (int mb = 0; mb < n_mb; ++mb) {
                const int u = mb * 128 + 32 * warp + lane;
                for (int j0 = 0; j0 < GT_ROWS; j0 += 16) {
                    uint32_t rh[8], rl[8];
#pragma unroll
                    for (int q = 0; q < 8; ++q) {
                        const int ja = j0 + 2 * q, jb = ja + 1;
                        const uint32_t pa = gt_split(Whh[(long)((ja >> 5) * H + ub + (ja & 31)) * H + u] * sw);
                        const uint32_t pb = gt_split(Whh[(long)((jb >> 5) * H + ub + (jb & 31)) * H + u] * sw);
                        rh[q] = (pa & 0xffffu) | (pb << 16);
                        rl[q] = (pa >> 16) | (pb & 0xffff0000u);
                    }
                    const uint32_t col = 64u + (uint32_t)(mb * 2) * 48u + (uint32_t)(j0 / 2);
                    gt_st8(tmem + ((uint32_t)(32 * warp) << 16) + col, rh);
                    gt_st8(tmem + ((uint32_t)(32 * warp) << 16) + col + 48u, rl);
                }
            }
            asm volatile("tcgen05.wait::st.sync.aligned;" ::: "memory");
        }
        asm volatile("tcgen05.fence::before_thread_sync;" ::: "memory");
        __syncthreads();
        asm volatile("tcgen05.fence::after_thread_sync;" ::: "memory");
    }
    // D fp32, B fp16 K-major, N = 16, M = 128; A fp16: K-major from tensor memory (TA) | MN-major (units contiguous) from
    // the shared-memory image
    const uint32_t idesc = (1u << 4) | (TA ? 0u : (1u << 15)) | ((uint32_t)(16 >> 3) << 17) | ((uint32_t)(128 >> 4) << 24);
    // MN-major SWIZZLE_128B: 64 units (128 B) x 8 gate rows per atom; LBO = next 64 units, SBO = next 8 gate rows
    const uint32_t a_lbo = swap_strides ? 1024u : (uint32_t)GT_WBLK, a_sbo = swap_strides ? (uint32_t)GT_WBLK : 1024u;

    struct PhaseA { float dout, r, z, n, hm, ghn, mnext; };
    auto load_a = [&](int t, int my) {
        PhaseA a = {0.f, 0.f, 0.f, 0.f, 0.f, 0.f, 0.f};
        if (my < d.N) {
            const long row = (long)t * d.N + my;
            a.dout = d.dout[row * H + ku];
            a.r = d.r[row * H + ku]; a.z = d.z[row * H + ku]; a.n = d.n[row * H + ku];
            a.hm = d.hm[row * H + ku];
            a.ghn = d.gh[row * 3 * H + 2 * H + ku];
            a.mnext = (t + 1 < d.T) ? gt_mask(d, t + 1, my) : 0.f;
        }
        return a;
    };
    auto part_of = [&](int grp, int buf) { return gt_part + ((size_t)grp * 2 + buf) * GT_PART_FLOATS; };
    const int my0 = blockIdx.y * GT_RW + warp;
    PhaseA pre = load_a(d.T - 1, my0);
    float carry0 = 0.f;                                                    // dh_{t+1} z_{t+1} of the first pass's (environment, unit)
    uint32_t it = 0;

    for (int t = d.T - 1; t >= 0; --t) {
        for (int grp = blockIdx.y; grp < ngrp; grp += gridDim.y) {
            const int my = grp * GT_RW + warp;
            const bool live = my < d.N;
            const bool first = grp == (int)blockIdx.y;
            const long row = (long)t * d.N + my;
            const PhaseA a = first ? pre : load_a(t, my);
            // ---- carry of step t + 1: own dh z + the unit groups' partial products, summed in a fixed order ----
            float dh = a.dout;
            if (t + 1 < d.T) {
                if (tid == 0) {
                    GT_STAMP(1, t, 0);
                    const unsigned target = (unsigned)ug * (unsigned)(d.T - 1 - t);
                    while (gt_ld_acquire(&gt_flags[1][grp]) < target) {}
                    GT_STAMP(1, t, 1);
                }
                __syncthreads();
                const float* p = part_of(grp, (t + 1) & 1) + (((size_t)blockIdx.x * ug) * GT_RW + warp) * GT_UB + lane;
                float sum = 0.f;
                for (int src = 0; src < ug; src += 4) {
                    const float v0 = __ldcg(p + (size_t)(src + 0) * GT_RW * GT_UB), v1 = __ldcg(p + (size_t)(src + 1) * GT_RW * GT_UB),
                                v2 = __ldcg(p + (size_t)(src + 2) * GT_RW * GT_UB), v3 = __ldcg(p + (size_t)(src + 3) * GT_RW * GT_UB);
                    sum += (v0 + v1) + (v2 + v3);
                }
                const float c = (first ? carry0 : (live ? carry_ws[(long)my * H + ku] : 0.f)) + sum;
                dh += c * a.mnext;
                if (tid == 0) GT_STAMP(1, t, 2);
            }
            // ---- phase A ----
            const float dn_pre = dh * (1.f - a.z) * (1.f - a.n * a.n);
            const float dz_pre = dh * (a.hm - a.n) * a.z * (1.f - a.z);
            const float dr_pre = dn_pre * a.ghn * a.r * (1.f - a.r);
            const float g0 = live ? dr_pre : 0.f, g1 = live ? dz_pre : 0.f, g2 = live ? dn_pre * a.r : 0.f;
            const float dhz = dh * a.z;
            if (first) carry0 = dhz;
            else if (live) carry_ws[(long)my * H + ku] = dhz;
            if (t > 0) {
                // ---- operand of the step: the CTA's gate gradients [8 environments x 96] as fp16 hi / lo, scaled by a
                //      power of two chosen from this step's largest magnitude ----
                float m = warp_max(fmaxf(fabsf(g0), fmaxf(fabsf(g1), fabsf(g2))));
                if (lane == 0) red[warp] = m;
                __syncthreads();
                m = red[0];
#pragma unroll
                for (int w = 1; w < GT_NT / 32; ++w) m = fmaxf(m, red[w]);
                const float sg = gt_pow2_scale(m, 12);
                const float gv[3] = {g0, g1, g2};
#pragma unroll
                for (int g = 0; g < 3; ++g) {
                    const uint32_t p = gt_split(gv[g] * sg);
                    const uint32_t off = gt_swz(warp, g * 32 + lane, 2048);
                    *reinterpret_cast<unsigned short*>(gop + off) = (unsigned short)(p & 0xffffu);
                    *reinterpret_cast<unsigned short*>(gop + G_PLANE + off) = (unsigned short)(p >> 16);
                }
                asm volatile("fence.proxy.async.shared::cta;" ::: "memory");
                asm volatile("tcgen05.fence::before_thread_sync;" ::: "memory");
                if (tid == 0) GT_STAMP(1, t, 3);
                __syncthreads();
                if (tid == 0) GT_STAMP(1, t, 4);
                if (warp == GT_CTRL) {
                    asm volatile("tcgen05.fence::after_thread_sync;" ::: "memory");
                    if (gt_elect()) {
                        for (int mb = 0; mb < n_mb; ++mb) {
                            const uint32_t dcol = tmem + (uint32_t)mb * 16u;
#pragma unroll
                            for (int j = 0; j < GT_ROWS / 16; ++j) {
                                const uint32_t ob = (uint32_t)(j >> 2) * 2048u + (uint32_t)(j & 3) * 32u;
                                const uint64_t bh = gt_desc(s_g + ob, 16, 1024), bl = gt_desc(s_g + G_PLANE + ob, 16, 1024);
                                if (TA) {
                                    const uint32_t ah = tmem + 64u + (uint32_t)(mb * 2) * 48u + 8u * (uint32_t)j, al = ah + 48u;
                                    gt_mma_ts(dcol, al, bh, idesc, j == 0 ? 0u : 1u);
                                    gt_mma_ts(dcol, ah, bl, idesc, 1u);
                                    gt_mma_ts(dcol, ah, bh, idesc, 1u);
                                } else {
                                    const uint32_t oa = (uint32_t)(2 * mb) * GT_WBLK + (uint32_t)j * 2048u;
                                    const uint64_t ah = gt_desc(s_w + oa, a_lbo, a_sbo), al = gt_desc(s_w + w_plane + oa, a_lbo, a_sbo);
                                    gt_mma(dcol, al, bh, idesc, j == 0 ? 0u : 1u);
                                    gt_mma(dcol, ah, bl, idesc, 1u);
                                    gt_mma(dcol, ah, bh, idesc, 1u);
                                }
                            }
                        }
                        gt_commit(&bar_acc);
                    }
                    __syncwarp();
                } else if (warp < 4) {
                    // ---- drain: row = unit mb * 128 + 32 warp + lane -> the partial carries of unit group 4 mb + warp ----
                    gt_mbar_wait(&bar_acc, it & 1u);
                    if (tid == 0) GT_STAMP(1, t, 5);
                    asm volatile("tcgen05.fence::after_thread_sync;" ::: "memory");
                    const float unscale = 1.f / (sw * sg);
                    float* pbase = part_of(grp, t & 1);
                    for (int mb = 0; mb < n_mb; ++mb) {
                        float v[8];
                        gt_ld8(tmem + ((uint32_t)(32 * warp) << 16) + (uint32_t)mb * 16u, v);
                        float* p = pbase + ((size_t)((4 * mb + warp) * ug + blockIdx.x) * GT_RW) * GT_UB + lane;
#pragma unroll
                        for (int e = 0; e < 8; ++e) __stcg(p + e * GT_UB, v[e] * unscale);
                    }
                    asm volatile("tcgen05.fence::before_thread_sync;" ::: "memory");
                    if (tid == 0) GT_STAMP(1, t, 6);
                }
                ++it;
                __syncthreads();
                if (tid == 0) GT_STAMP(1, t, 7);
                if (tid == 0) gt_release_add(&gt_flags[1][grp]);
                if (tid == 0) GT_STAMP(1, t, 8);
            }
            // ---- consumed by the weight-gradient products after this kernel ----
            if (live) {
                float* dgi = d.dgi + row * 3 * H;
                float* dgh = d.dgh + row * 3 * H;
                dgi[ku] = dr_pre; dgi[H + ku] = dz_pre; dgi[2 * H + ku] = dn_pre;
                dgh[ku] = dr_pre; dgh[H + ku] = dz_pre; dgh[2 * H + ku] = dn_pre * a.r;
            }
            if (first && t > 0) pre = load_a(t - 1, my);
        }
    }
    asm volatile("tcgen05.fence::before_thread_sync;" ::: "memory");
    __syncthreads();
    if (warp == GT_CTRL)
        asm volatile("tcgen05.dealloc.cta_group::1.sync.aligned.b32 %0, %1;" ::"r"(tmem), "r"(TCOLS));
}

// ---------------------------------------------------------------------------------------------------------------
// host side
// ---------------------------------------------------------------------------------------------------------------
// weights resident in tensor memory (default) or in shared memory only (A2CB_GRU_TC_TMEM=0)
static bool gt_tmem_weights() {
    static int mode = -1;
    if (mode < 0) { const char* e = getenv("A2CB_GRU_TC_TMEM"); mode = (e && e[0] == '0') ? 0 : 1; }
    return mode == 1;
}
static size_t gt_smem_fwd(int H) {
    return (size_t)2 * (H / 64) * 1024 + (size_t)(gt_tmem_weights() ? 1 : 2) * (H / 64) * GT_WBLK + 4096 + 1024;
}
static size_t gt_smem_bwd(int H) { return (gt_tmem_weights() ? 0 : (size_t)2 * (H / 64) * GT_WBLK + 4096) + 8192 + 1024; }

static int gt_sms() {
    static int sms = 0;
    if (!sms) {
        int dev = 0;
        cudaGetDevice(&dev);
        cudaDeviceGetAttribute(&sms, cudaDevAttrMultiProcessorCount, dev);
    }
    return sms;
}

// 2 (default): tcgen05 kernels where the shape allows; 1: mma.sync 3xTF32 kernels; 0: fp32 SIMT kernels
// (a2cb_set_gru_mode; initial value from A2CB_GRU_MODE, or the older A2CB_GRU_TC=0 / A2CB_GRU_MMA=0 switches)
static int g_gru_mode = -1;
int gru_mode() {
    if (g_gru_mode < 0) {
        g_gru_mode = 2;
        const char* e = getenv("A2CB_GRU_MODE");
        if (e && e[0] >= '0' && e[0] <= '2') g_gru_mode = e[0] - '0';
        else {
            const char* t = getenv("A2CB_GRU_TC");
            const char* m = getenv("A2CB_GRU_MMA");
            if (m && m[0] == '0') g_gru_mode = 0;
            else if (t && t[0] == '0') g_gru_mode = 1;
        }
    }
    return g_gru_mode;
}
void set_gru_mode(int mode) { g_gru_mode = mode < 0 ? -1 : (mode > 2 ? 2 : mode); }

// copies the clock64 stamps of the last debug launches (2 directions x 16 steps x 16 phases) to the host
int gru_tc_debug_stamps(long long* out) {
    return cudaMemcpyFromSymbol(out, gt_dbg, sizeof(long long) * 2 * 16 * 16) == cudaSuccess ? A2CB_OK : A2CB_ERR_CUDA;
}

bool gru_tc_supported(int N, int H) {
    if (gru_mode() != 2 || N <= 0 || H % 128 != 0 || H > GT_MAX_H) return false;
    if ((N + GT_RW - 1) / GT_RW > GT_MAX_GROUPS) return false;
    if (gt_smem_fwd(H) > 226 * 1024 || gt_smem_bwd(H) > 226 * 1024) return false;
    return H / GT_UB <= gt_sms();
}

int gru_tc(int backward, const GruDesc& d, const float* Whh, const float* bhh, cudaStream_t st) {
    A2CB_REQUIRE(gru_tc_supported(d.N, d.H), "gru_tc: unsupported shape (N=%d H=%d)", d.N, d.H);
    A2CB_REQUIRE(Whh && d.hm && d.masks, "gru_tc: null pointer");
    A2CB_REQUIRE(backward || (bhh && d.h0 && d.gi && d.out), "gru_tc: forward pointers");
    A2CB_REQUIRE(!backward || (d.r && d.z && d.n && d.gh && d.dout && d.dgi && d.dgh), "gru_tc: backward pointers");
    // dynamic shared memory: exactly what this H needs (the device limit of 227 KiB includes the kernels' few static bytes)
    static int attr_h[2] = {0, 0};
    if (attr_h[backward ? 1 : 0] < d.H) {
        const bool ta = gt_tmem_weights();
        const void* kfn = backward ? (ta ? (const void*)gru_tc_bwd_kernel<true> : (const void*)gru_tc_bwd_kernel<false>)
                                   : (ta ? (const void*)gru_tc_fwd_kernel<true> : (const void*)gru_tc_fwd_kernel<false>);
        const cudaError_t ea = cudaFuncSetAttribute(kfn, cudaFuncAttributeMaxDynamicSharedMemorySize,
                                                    (int)(backward ? gt_smem_bwd(d.H) : gt_smem_fwd(d.H)));
        if (ea != cudaSuccess) {
            set_error("gru_tc: shared memory attribute (%zu bytes): %s", backward ? gt_smem_bwd(d.H) : gt_smem_fwd(d.H),
                      cudaGetErrorString(ea));
            cudaGetLastError();
            return A2CB_ERR_CUDA;
        }
        attr_h[backward ? 1 : 0] = d.H;
    }
    static int swap = -1;                                                  // probe knob: exchange LBO / SBO of the MN-major operand
    if (swap < 0) { const char* e = getenv("A2CB_GRU_TC_SWAP"); swap = (e && e[0] == '1') ? 1 : 0; }
    const int ug = d.H / GT_UB;
    const int ngrp = (d.N + GT_RW - 1) / GT_RW;
    int gy = gt_sms() / ug;
    gy = gy < 1 ? 1 : (gy > ngrp ? ngrp : gy);
    A2CB_REQUIRE(!backward || ngrp <= gy || d.dcarry, "gru_tc: dcarry workspace [N, H] needed for N > %d", gy * GT_RW);
    {
        void* fptr = nullptr;
        cudaError_t e0 = cudaGetSymbolAddress(&fptr, gt_flags);
        if (e0 == cudaSuccess)
            e0 = cudaMemsetAsync(reinterpret_cast<unsigned*>(fptr) + (backward ? GT_MAX_GROUPS : 0), 0,
                                 GT_MAX_GROUPS * sizeof(unsigned), st);
        if (e0 != cudaSuccess) { set_error("gru_tc: flag reset: %s", cudaGetErrorString(e0)); return A2CB_ERR_CUDA; }
    }
    static int dbg = -1;
    if (dbg < 0) { const char* e = getenv("A2CB_GRU_TC_DEBUG"); dbg = (e && e[0] == '1') ? 1 : 0; }
    GruDesc dd = d;
    float* ws = d.dcarry;
    int sw = swap;
    void* args_f[] = {(void*)&dd, (void*)&Whh, (void*)&bhh, (void*)&dbg};
    void* args_b[] = {(void*)&dd, (void*)&Whh, (void*)&ws, (void*)&sw, (void*)&dbg};
    dim3 grid((unsigned)ug, (unsigned)gy);
    const bool ta = gt_tmem_weights();
    const void* kf = ta ? (const void*)gru_tc_fwd_kernel<true> : (const void*)gru_tc_fwd_kernel<false>;
    const void* kb = ta ? (const void*)gru_tc_bwd_kernel<true> : (const void*)gru_tc_bwd_kernel<false>;
    cudaError_t e = backward ? cudaLaunchCooperativeKernel(kb, grid, dim3(GT_NT), args_b, gt_smem_bwd(d.H), st)
                             : cudaLaunchCooperativeKernel(kf, grid, dim3(GT_NT), args_f, gt_smem_fwd(d.H), st);
    if (e != cudaSuccess) {
        set_error("gru_tc: cooperative launch failed: %s", cudaGetErrorString(e));
        cudaGetLastError();
        return A2CB_ERR_CUDA;
    }
    return check_launch("gru_tc");
}

}  // namespace a2cb
