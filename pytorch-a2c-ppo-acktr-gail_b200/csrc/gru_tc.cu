// GRU recurrence (reference a2c_ppo_acktr/model.py:111-166) on tcgen05: the whole [T, N] sequence in ONE cooperative
// launch per direction, the per-step product on the 5th-generation tensor cores.
//
// The mma.sync kernels of gru_persistent.cu spend ~1,100 instructions per warp and time step on operand handling:
// fp32 -> TF32 hi/lo splits of the RESIDENT weights (which never change during the launch), scalar loads of the
// exchanged state, a shared-memory fold of eight partial tiles.  Here
//   * the CTA's slice of W_hh (the 96 gate rows of its 32 hidden units) is converted ONCE, at kernel start, into two
//     fp16 planes w = (w_hi + w_lo) / s_W (scaled by a power of two so that w_lo stays a normal fp16: 22 significant
//     bits, the same 4 bytes per element as fp32).  Forward: the hi plane lives in TENSOR MEMORY (A operand of a
//     tcgen05.mma with the operand in TMEM: 21 cycles per M = 128, N = 16, K = 16 instruction, tools/mma_probe.cu), the lo
//     plane in shared memory as the canonical 128-byte-swizzled K-major UMMA operand (46 cycles per instruction: the
//     4 KiB A read, whatever N is).  Backward: both planes of the TRANSPOSED slice live in tensor memory;
//   * the exchanged operand of a step (forward: the masked state of the group's 8 environments; backward: the CTA's own
//     gate gradients) is split into fp16 hi / lo and laid out as ONE 16-row B operand: rows 0..7 the hi parts of the 8
//     environments, rows 8..15 their lo parts.  One instruction W_p x [x_hi | x_lo] therefore yields both products of a
//     weight plane p, and a reduction step costs TWO instructions (W_hi, W_lo) instead of three; the epilogue adds
//     accumulator columns e and e + 8: (W_hi + W_lo)(x_hi + x_lo), fp32 accumulation in TMEM, split over up to four
//     accumulators so that the tensor core's truncating accumulation sees chains of at most 24 instructions;
//   * forward: all-gather of the state through L2: producers write fp16 planes ALREADY in the shared-memory image layout,
//     a consumer CTA fetches the state of a step with bulk copies (cp.async.bulk global -> shared, one mbarrier per
//     chunk of k-blocks, so that the instructions of chunk c overlap the arrival of chunk c + 1).  A ninth warp does
//     nothing but poll, fetch and issue; the eight gate warps never wait for it except through the accumulator barrier;
//   * backward: every CTA multiplies ITS gate gradients (K = 96) by its weight slice for ALL units and the 16 partial
//     products are summed by their owners (reduce-scatter through L2, fixed order: deterministic) -- no gate gradient
//     ever needs to be gathered before the product;
//   * one release / acquire counter per environment group and step; environments are independent.
// Ownership, saved quantities, buffers and the gate arithmetic are those of gru_persistent.cu, so either path feeds
// the same weight-gradient products and the same tests.
#include <cuda_fp16.h>
#include <stdlib.h>

#include "common.cuh"
#include "decls.cuh"
#include "gru.cuh"

namespace a2cb {

namespace {

constexpr int GT_GW = 8;                      // gate warps: thread (warp e, lane u) owns (environment e, unit u)
constexpr int GT_NT = 32 * (GT_GW + 1);       // + the control warp
constexpr int GT_RW = 8;                      // environments per group
constexpr int GT_UB = 32;                     // hidden units per CTA
constexpr int GT_ROWS = 3 * GT_UB;            // gate rows per CTA: row g * 32 + u
constexpr int GT_WBLK = GT_ROWS * 128;        // bytes of one 64-unit block of one weight plane
constexpr int GT_MAX_H = 512;
constexpr int GT_MAX_GROUPS = 64;             // environment groups per launch (N <= 512)
constexpr int GT_IMG_BYTES = 2 * (GT_MAX_H / 64) * 1024;              // state image of one group and step (both planes)
constexpr int GT_PART_FLOATS = (GT_MAX_H / 32) * (GT_MAX_H / 32) * GT_RW * GT_UB;   // partial carries of one group and step
constexpr int GT_CTRL = GT_GW;                // control warp: polls, issues the bulk copies and the MMAs
constexpr uint32_t GT_ACC_COLS = 64;          // tensor-memory columns of the accumulators (4 x 16)
constexpr float GT_SX = 16.f;                 // power-of-two scale of the forward state planes

__device__ long long gt_dbg[2][16][16];            // A2CB_GRU_TC_DEBUG=1: clock64 stamps of CTA (0, 0), steps 16..31, per phase
#define GT_STAMP(dir, t, i) do { if (dbg && blockIdx.x == 0 && blockIdx.y == 0 && (t) >= 16 && (t) < 32) gt_dbg[dir][(t) - 16][i] = clock64(); } while (0)
__device__ uint4 gt_img[(size_t)GT_MAX_GROUPS * 2 * GT_IMG_BYTES / 16];
// backward exchange: [GT_MAX_GROUPS arrival counters][group][2 buffers][partial carries] -- one memset resets both
__device__ float gt_part_all[GT_MAX_GROUPS + (size_t)GT_MAX_GROUPS * 2 * GT_PART_FLOATS];

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
__device__ __forceinline__ void gt_mbar_arrive(uint64_t* bar) {
    asm volatile("mbarrier.arrive.shared::cta.b64 _, [%0];" ::"r"(gt_smem(bar)) : "memory");
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
// 16 accumulator columns of this thread's row; the wait "touches" the registers so that their uses stay behind it
__device__ __forceinline__ void gt_ld16_issue(uint32_t taddr, uint32_t* v) {
    asm volatile(
        "tcgen05.ld.sync.aligned.32x32b.x16.b32"
        "{%0, %1, %2, %3, %4, %5, %6, %7, %8, %9, %10, %11, %12, %13, %14, %15}, [%16];\n"
        : "=r"(v[0]), "=r"(v[1]), "=r"(v[2]), "=r"(v[3]), "=r"(v[4]), "=r"(v[5]), "=r"(v[6]), "=r"(v[7]), "=r"(v[8]),
          "=r"(v[9]), "=r"(v[10]), "=r"(v[11]), "=r"(v[12]), "=r"(v[13]), "=r"(v[14]), "=r"(v[15])
        : "r"(taddr));
}
__device__ __forceinline__ void gt_ld_wait16(uint32_t* v) {
    asm volatile("tcgen05.wait::ld.sync.aligned;"
                 : "+r"(v[0]), "+r"(v[1]), "+r"(v[2]), "+r"(v[3]), "+r"(v[4]), "+r"(v[5]), "+r"(v[6]), "+r"(v[7]), "+r"(v[8]),
                   "+r"(v[9]), "+r"(v[10]), "+r"(v[11]), "+r"(v[12]), "+r"(v[13]), "+r"(v[14]), "+r"(v[15])
                 :
                 : "memory");
}
__device__ __forceinline__ void gt_gate_barrier() { asm volatile("bar.sync 1, %0;" ::"n"(32 * GT_GW) : "memory"); }
// strong (L1-bypassing) 16-byte / 4-byte loads of words other CTAs are writing right now
__device__ __forceinline__ uint4 gt_ld_relaxed16(const void* p) {
    uint4 v;
    asm volatile("ld.relaxed.gpu.global.v4.u32 {%0, %1, %2, %3}, [%4];" : "=r"(v.x), "=r"(v.y), "=r"(v.z), "=r"(v.w) : "l"(p) : "memory");
    return v;
}
__device__ __forceinline__ uint32_t gt_ld_relaxed4(const void* p) {
    uint32_t v;
    asm volatile("ld.relaxed.gpu.global.u32 %0, [%1];" : "=r"(v) : "l"(p) : "memory");
    return v;
}
// 2-bit validity tag of exchange step t (the buffer t & 1 held step t -+ 2 before: the other tag; 0 = never written)
__device__ __forceinline__ uint32_t gt_tag(int t) { return 1u + (uint32_t)((t >> 1) & 1); }
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
// `k_first`: the image starts at reduction index k_first (a multiple of 64): only columns k >= k_first are written.
__device__ float gt_build_weights(uint8_t* img, const float* __restrict__ Whh, int H, int ub, float* red, int planes = 3,
                                  int k_first = 0) {
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
            if (k < k_first) continue;
            const uint32_t p = gt_split(Whh[(long)((row >> 5) * H + ub + (row & 31)) * H + k] * sw);
            const uint32_t off = gt_swz(row, k - k_first, GT_WBLK);
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

// Episode masks of the CTA's environments for every step and pass -> shared memory, once.  The per-step lookup is a
// DEPENDENT pair of loads (minibatch row list -> flat mask array) that sat at the end of every step, in front of the
// next step's wait: 11 % (forward) / 14 % (backward) of all warp-stall samples (profiles/r02n_gru_full.txt).
__device__ __forceinline__ void gt_fill_masks(float* msk, const GruDesc& d, int use_msk, int ngrp) {
    if (use_msk) {
        const int npass = (ngrp - (int)blockIdx.y + (int)gridDim.y - 1) / (int)gridDim.y;
        for (int i = threadIdx.x; i < npass * d.T * GT_RW; i += GT_NT) {
            const int e = i % GT_RW, t = (i / GT_RW) % d.T, p = i / (GT_RW * d.T);
            const int my = ((int)blockIdx.y + p * (int)gridDim.y) * GT_RW + e;
            msk[i] = my < d.N ? gt_mask(d, t, my) : 0.f;
        }
    }
    __syncthreads();
}

// tensor-memory columns to allocate: a power of two >= 32
__host__ __device__ constexpr uint32_t gt_tmem_cols(uint32_t need) {
    return need <= 32 ? 32u : need <= 64 ? 64u : need <= 128 ? 128u : need <= 256 ? 256u : 512u;
}

// ---------------------------------------------------------------------------------------------------------------
// forward.  grid (H / 32 unit groups, resident environment groups); CTA (bx, by) owns units 32 bx .. 32 bx + 31 and
// the environment groups by, by + gridDim.y, ...  Shared memory: [state image: hi plane, lo plane][lo weight plane][tail].
// Tensor memory: [accumulators: 64 columns][hi weight plane: 128 lanes (96 used) x H / 2 columns].
// NKB = H / 64 k-blocks (template: every descriptor offset of the issue loop is a compile-time constant).
// ---------------------------------------------------------------------------------------------------------------
template <int NKB>
__global__ void __launch_bounds__(GT_NT, 1)
gru_tc_fwd_kernel(const GruDesc d, const float* __restrict__ Whh, const float* __restrict__ bhh, const int use_msk, const int dbg) {
    constexpr int H = NKB * 64;
    constexpr int NCH = (NKB % 4 == 0) ? 4 : 2;                             // chunks of the state fetch = accumulators
    constexpr int SPC = NKB * 4 / NCH;                                      // 16-element reduction steps per chunk
    constexpr uint32_t ST_PLANE = (uint32_t)NKB * 1024u;                    // one plane of the state image: [NKB][8 rows x 128 B]
    constexpr uint32_t ST_BYTES = 2u * ST_PLANE;
    constexpr uint32_t W_PLANE = (uint32_t)(NKB - (4 * NKB < 56 - 4 * NKB ? NKB : (56 - 4 * NKB) / 4)) * GT_WBLK;   // lo k-blocks kept in shared memory
    // the lo weight plane goes to tensor memory too as far as the 512 columns reach (H = 512: 24 of its 32 reduction
    // steps); only the remaining k-blocks stay in shared memory and cost an instruction with a shared-memory A operand
    constexpr int LO_TS = (4 * NKB < 56 - 4 * NKB) ? 4 * NKB : 56 - 4 * NKB;   // reduction steps of the lo plane in tensor memory
    constexpr int LO_KB0 = LO_TS / 4;                                       // first k-block of the shared-memory lo image
    constexpr uint32_t HI_COL0 = GT_ACC_COLS, LO_COL0 = GT_ACC_COLS + (uint32_t)H / 2;
    constexpr uint32_t TCOLS = gt_tmem_cols(LO_COL0 + 8u * LO_TS);
    static_assert(LO_TS % 4 == 0 && LO_COL0 + 8u * LO_TS <= 512u, "gru_tc_fwd: tensor-memory budget");
    static_assert(NCH * 16 <= (int)GT_ACC_COLS && 2 * SPC <= 24, "gru_tc_fwd: accumulator chains");
    extern __shared__ uint8_t gt_dyn[];
    __shared__ uint64_t bar_state[4], bar_acc;
    __shared__ uint32_t tmem_slot;
    __shared__ float red[GT_NT / 32];
    const uint32_t base = (gt_smem(gt_dyn) + 1023u) & ~1023u;
    uint8_t* sm = gt_dyn + (base - gt_smem(gt_dyn));
    const uint32_t s_state = base, s_w = base + ST_BYTES;
    float* xchg = reinterpret_cast<float*>(sm + ST_BYTES + W_PLANE);        // [3 gates][8 environments][32 units]
    float* msk = xchg + 3 * GT_RW * GT_UB;                                  // [passes][T][8 environments] episode masks
    const int tid = threadIdx.x, lane = tid & 31, warp = tid >> 5;
    const bool gate = warp < GT_GW;
    const int ub = blockIdx.x * GT_UB, ku = ub + lane;
    const int ngrp = (d.N + GT_RW - 1) / GT_RW;
    const bool single = ngrp <= (int)gridDim.y;
    gt_fill_masks(msk, d, use_msk, ngrp);
    // mask of (pass p, step t, this warp's environment): from the table, or the dependent idx -> masks lookup
    auto mask_of = [&](int p, int t, int my) { return use_msk ? msk[(p * d.T + t) * GT_RW + warp] : gt_mask(d, t, my); };

    if (tid == 0) {
        for (int c = 0; c < 4; ++c) gt_mbar_init(&bar_state[c], GT_GW / NCH);
        gt_mbar_init(&bar_acc, 1);
        asm volatile("fence.mbarrier_init.release.cluster;" ::: "memory");
    }
    if (warp == GT_CTRL) {
        asm volatile("tcgen05.alloc.cta_group::1.sync.aligned.shared::cta.b32 [%0], %1;" ::"r"(gt_smem(&tmem_slot)), "r"(TCOLS));
        asm volatile("tcgen05.relinquish_alloc_permit.cta_group::1.sync.aligned;" ::);
    }
    const float sw = gt_build_weights(sm + ST_BYTES, Whh, H, ub, red, LO_KB0 < NKB ? 2 : 0, LO_KB0 * 64);   // rest of the lo plane
    const float inv_scale = 1.f / (sw * GT_SX);
    asm volatile("fence.proxy.async.shared::cta;" ::: "memory");          // weight image: generic stores -> tensor-core reads
    asm volatile("tcgen05.fence::before_thread_sync;" ::: "memory");
    __syncthreads();
    asm volatile("tcgen05.fence::after_thread_sync;" ::: "memory");
    const uint32_t tmem = tmem_slot;
    // weight planes -> tensor memory: thread (warp w < 4, lane) owns row 32 w + lane (rows >= 96: zero); column
    // HI_COL0 + c (LO_COL0 + c) holds the hi (lo) parts of reduction elements 2 c, 2 c + 1 of that row
    if (warp < 4) {
        const int row = 32 * warp + lane;
        const float* wr = Whh + (long)((row >> 5) * H + ub + (row & 31)) * H;       // rows >= 96 never dereferenced
        for (int c0 = 0; c0 < H / 2; c0 += 8) {
            uint32_t rh[8], rl[8];
#pragma unroll
            for (int j = 0; j < 8; ++j) {
                uint32_t pa = 0u, pb = 0u;
                if (row < GT_ROWS) {
                    pa = gt_split(wr[2 * (c0 + j)] * sw);
                    pb = gt_split(wr[2 * (c0 + j) + 1] * sw);
                }
                rh[j] = (pa & 0xffffu) | (pb << 16);
                rl[j] = (pa >> 16) | (pb & 0xffff0000u);
            }
            gt_st8(tmem + ((uint32_t)(32 * warp) << 16) + HI_COL0 + (uint32_t)c0, rh);
            if (c0 < 8 * LO_TS) gt_st8(tmem + ((uint32_t)(32 * warp) << 16) + LO_COL0 + (uint32_t)c0, rl);
        }
        asm volatile("tcgen05.wait::st.sync.aligned;" ::: "memory");
    }
    asm volatile("tcgen05.fence::before_thread_sync;" ::: "memory");
    __syncthreads();
    asm volatile("tcgen05.fence::after_thread_sync;" ::: "memory");

    // Writes this thread's masked state v of (environment `warp`, unit `lane`) into the group's exchange image for
    // step t.  Global layout: [8 environments][H / 4 pieces] x 16 bytes, piece (e, j) = {hi parts of units 4 j .. 4 j + 3,
    // their lo parts}; the two low mantissa bits of the piece's first two lo parts carry the step's validity tag, so a
    // consumer that finds the tag knows the whole 16-byte store has landed -- no flag, no fence, no barrier.
    auto put_state = [&](uint8_t* img, float v, int t) {
        const uint32_t p = gt_split(v * GT_SX);                            // {lo : hi}
        const uint32_t p1 = __shfl_down_sync(0xffffffffu, p, 1), p2 = __shfl_down_sync(0xffffffffu, p, 2),
                       p3 = __shfl_down_sync(0xffffffffu, p, 3);
        if ((lane & 3) == 0) {
            const uint32_t tag = gt_tag(t);
            uint4 w;
            w.x = (p & 0xffffu) | (p1 << 16);
            w.y = (p2 & 0xffffu) | (p3 << 16);
            w.z = (((p >> 16) | (p1 & 0xffff0000u)) & 0xfffefffeu) | (tag & 1u) | ((tag >> 1) << 16);
            w.w = (p2 >> 16) | (p3 & 0xffff0000u);
            *reinterpret_cast<uint4*>(img + ((size_t)warp * (H / 4) + (ub >> 2) + (lane >> 2)) * 16) = w;
        }
    };
    auto image_of = [&](int grp, int buf) {
        return reinterpret_cast<uint8_t*>(gt_img) + ((size_t)grp * 2 + buf) * GT_IMG_BYTES;
    };

    // ---- image of step 0: hm_0 = h0 * mask_0 ----
    float hm_cur = 0.f;                                                    // this thread's hm_t (single pass: kept in a register)
    for (int grp = blockIdx.y, ps = 0; grp < ngrp; grp += gridDim.y, ++ps) {
        const int my = grp * GT_RW + warp;
        if (gate) {
            float v = 0.f;
            if (my < d.N) {
                v = d.h0[(long)my * H + ku] * mask_of(ps, 0, my);
                d.hm[(long)my * H + ku] = v;
            }
            put_state(image_of(grp, 0), v, 0);
            hm_cur = v;
        }
    }

    if (warp == GT_CTRL) {
        // ===== control warp: per (step, group) wait for the chunks of the group's state that the gate warps fetch into
        // shared memory, issue the step's instructions.  It never meets the gate warps at a CTA barrier: they fetch the
        // state of the next (step, group) only after their second gate barrier, i.e. after the accumulators of this one
        // have been drained (and therefore after its instructions have read the state image). =====
        // instruction descriptor: D fp32, A / B fp16, both K-major, N = 16, M = 128
        const uint32_t idesc = (1u << 4) | ((uint32_t)(16 >> 3) << 17) | ((uint32_t)(128 >> 4) << 24);
        // B: 16 rows = the 8 environments' hi parts (first 8-row atom) and their lo parts (second atom, one plane further)
        const uint64_t b0 = gt_desc(s_state, 16, ST_PLANE);
        const uint64_t a0 = gt_desc(s_w, 16, 1024);                        // lo weight plane, k-blocks >= LO_KB0: 96 (128) rows, K-major
        uint32_t it = 0;                                                   // (step, pass) counter: barrier parities
        for (int t = 0; t < d.T; ++t) {
            for (int grp = blockIdx.y; grp < ngrp; grp += gridDim.y, ++it) {
                if (lane == 0) GT_STAMP(0, t, 0);
#pragma unroll
                for (int c = 0; c < NCH; ++c) {
                    gt_mbar_wait(&bar_state[c], it & 1u);                            // chunk c of the state image is in shared memory
                    if (lane == 0) { if (c == 0) GT_STAMP(0, t, 2); else GT_STAMP(0, t, 10 + c); }
                    asm volatile("tcgen05.fence::after_thread_sync;" ::: "memory");
                    if (gt_elect()) {
                        const uint32_t dcol = tmem + (uint32_t)c * 16u;
#pragma unroll
                        for (int j = 0; j < SPC; ++j) {
                            const int s = c * SPC + j;                                     // compile-time after unrolling
                            const uint32_t ob = (uint32_t)(s >> 2) * 1024u + (uint32_t)(s & 3) * 32u;
                            const uint64_t b = b0 + (uint64_t)(ob >> 4);
                            if (s < LO_TS) {                                               // W_lo [x_hi | x_lo]
                                gt_mma_ts(dcol, tmem + LO_COL0 + 8u * (uint32_t)s, b, idesc, j == 0 ? 0u : 1u);
                            } else {
                                const uint32_t o = (uint32_t)((s >> 2) - LO_KB0) * GT_WBLK + (uint32_t)(s & 3) * 32u;
                                gt_mma(dcol, a0 + (uint64_t)(o >> 4), b, idesc, j == 0 ? 0u : 1u);
                            }
                            gt_mma_ts(dcol, tmem + HI_COL0 + 8u * (uint32_t)s, b, idesc, 1u);  // W_hi [x_hi | x_lo]
                        }
                        if (c == NCH - 1) gt_commit(&bar_acc);
                    }
                    __syncwarp();
                    if (lane == 0 && c == 0) GT_STAMP(0, t, 10);
                }
                if (lane == 0) GT_STAMP(0, t, 3);
            }
        }
    } else {
        const float bias = warp < 3 ? bhh[warp * H + ku] : 0.f;           // drain warp g adds b_hh of gate g
        struct GateIn { float r, z, n, mnext; };
        auto load_gi = [&](int ps, int t, int my) {
            GateIn q = {0.f, 0.f, 0.f, 0.f};
            if (my < d.N) {
                const float* gi = d.gi + ((long)t * d.N + my) * 3 * H;
                q.r = __ldg(gi + ku); q.z = __ldg(gi + H + ku); q.n = __ldg(gi + 2 * H + ku);
                q.mnext = (t + 1 < d.T) ? mask_of(ps, t + 1, my) : 0.f;
            }
            return q;
        };
        GateIn pre = load_gi(0, 0, blockIdx.y * GT_RW + warp);
        uint32_t it = 0;
        for (int t = 0; t < d.T; ++t) {
            for (int grp = blockIdx.y, ps = 0; grp < ngrp; grp += gridDim.y, ++it, ++ps) {
                const int my = grp * GT_RW + warp;
                const bool live = my < d.N;
                const long row = (long)t * d.N + my;
                const GateIn q = single ? pre : load_gi(ps, t, my);
                if (single && t + 1 < d.T) pre = load_gi(0, t + 1, my);          // a whole step ahead of its use
                const float hprev = single ? hm_cur : (live ? d.hm[row * H + ku] : 0.f);
                // ---- fetch: every gate warp loads its share of the group's exchange image (strong 16-byte loads), waits
                //      until every piece carries this step's tag, scatters hi / lo parts into the shared-memory operand and
                //      arrives on its chunk's barrier: warps c * GT_GW / NCH .. share chunk c (a k-range) of both planes.
                //      (History of this exchange: a counter flag + bulk copies needed a cross-proxy fence -- fence.proxy.async
                //      measured 1,400 cycles -- and a red.release.gpu of ~1,000 cycles per step; an unfenced arrival counter
                //      in front of these loads, as the backward kernel uses, made this direction 7 % slower.) ----
                {
                    constexpr int WPC = GT_GW / NCH;                                 // warps per chunk
                    constexpr int JPC = H / 4 / NCH;                                 // pieces per environment and chunk
                    constexpr int PER = GT_RW * JPC / (32 * WPC);                    // pieces per lane
                    const int c = warp / WPC, wq = warp % WPC;
                    const uint8_t* src = image_of(grp, t & 1);
                    const uint32_t tag = gt_tag(t);
                    const uint4* ptr[PER];
                    uint32_t dst[PER];
                    uint4 w[PER];
#pragma unroll
                    for (int i = 0; i < PER; ++i) {
                        const int o = (wq * PER + i) * 32 + lane;                    // piece of the chunk: (environment, j)
                        const int e = o / JPC, j = c * JPC + (o - e * JPC);
                        ptr[i] = reinterpret_cast<const uint4*>(src + ((size_t)e * (H / 4) + j) * 16);
                        dst[i] = gt_swz(e, 4 * j, 1024);
                        w[i] = gt_ld_relaxed16(ptr[i]);
                    }
#pragma unroll
                    for (int i = 0; i < PER; ++i) {
                        while (((w[i].z & 1u) | ((w[i].z >> 15) & 2u)) != tag) w[i] = gt_ld_relaxed16(ptr[i]);
                        *reinterpret_cast<uint2*>(sm + dst[i]) = make_uint2(w[i].x, w[i].y);
                        *reinterpret_cast<uint2*>(sm + ST_PLANE + dst[i]) = make_uint2(w[i].z, w[i].w);
                    }
                    asm volatile("fence.proxy.async.shared::cta;" ::: "memory");     // these generic writes -> tensor-core reads
                    __syncwarp();
                    if (lane == 0) gt_mbar_arrive(&bar_state[c]);
                    if (tid == 0) GT_STAMP(0, t, 15);
                }
                if (warp < 3) {
                    // ===== drain: gate `warp`, row = lane: NCH accumulators x (8 environments hi | 8 lo) =====
                    gt_mbar_wait(&bar_acc, it & 1u);
                    if (tid == 0) GT_STAMP(0, t, 4);
                    asm volatile("tcgen05.fence::after_thread_sync;" ::: "memory");
                    uint32_t v[NCH][16];
#pragma unroll
                    for (int a = 0; a < NCH; ++a) gt_ld16_issue(tmem + ((uint32_t)(32 * warp) << 16) + (uint32_t)a * 16u, v[a]);
#pragma unroll
                    for (int a = 0; a < NCH; ++a) gt_ld_wait16(v[a]);
#pragma unroll
                    for (int e = 0; e < 8; ++e) {
                        float s8 = 0.f;
#pragma unroll
                        for (int a = 0; a < NCH; ++a) s8 += __uint_as_float(v[a][e]) + __uint_as_float(v[a][e + 8]);
                        xchg[(warp * 8 + e) * 32 + lane] = s8 * inv_scale + bias;
                    }
                    asm volatile("tcgen05.fence::before_thread_sync;" ::: "memory");
                    if (tid == 0) GT_STAMP(0, t, 5);
                }
                gt_gate_barrier();                                         // gate pre-activations of W_hh hm_t + b_hh are in xchg
                if (tid == 0) GT_STAMP(0, t, 6);
                const float ghr = xchg[(0 * 8 + warp) * 32 + lane], ghz = xchg[(1 * 8 + warp) * 32 + lane],
                            ghn = xchg[(2 * 8 + warp) * 32 + lane];
                const float rr = gt_sigmoid(q.r + ghr);
                const float zz = gt_sigmoid(q.z + ghz);
                const float nn = tanhf(q.n + rr * ghn);
                const float h = (1.f - zz) * nn + zz * hprev;
                const float hm_next = live ? h * q.mnext : 0.f;
                if (t + 1 < d.T) put_state(image_of(grp, (t + 1) & 1), hm_next, t + 1);   // what the other CTAs wait for
                hm_cur = hm_next;
                if (tid == 0) GT_STAMP(0, t, 7);
                // several groups per CTA: the next pass's drain must not overwrite xchg while this pass still reads it
                // (one group per CTA: the next drain needs every CTA's new state, this one's slowest warp included)
                if (!single) gt_gate_barrier();
                if (tid == 0) GT_STAMP(0, t, 8);
                if (tid == 0) GT_STAMP(0, t, 9);
                // everything below is consumed by later kernels only
                if (live) {
                    d.out[row * H + ku] = h;
                    if (d.r) {
                        d.r[row * H + ku] = rr; d.z[row * H + ku] = zz; d.n[row * H + ku] = nn;
                        d.gh[row * 3 * H + 2 * H + ku] = ghn;              // the only part of gh the backward pass reads
                    }
                    if (t + 1 < d.T) d.hm[(row + d.N) * H + ku] = hm_next;
                }
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
// Both planes of the TRANSPOSED weight slice live in tensor memory: per 128-unit block and plane 48 columns (column c =
// gate rows 2 c, 2 c + 1) behind the 64 accumulator columns -- no weight image in shared memory at all; every
// instruction reads only the 512-byte gate-gradient operand [8 environments hi | 8 lo] x 16 gate rows from it.
// ---------------------------------------------------------------------------------------------------------------
template <int NKB>
__global__ void __launch_bounds__(GT_NT, 1)
gru_tc_bwd_kernel(const GruDesc d, const float* __restrict__ Whh, float* __restrict__ carry_ws, const int use_msk, const int dbg) {
    constexpr int H = NKB * 64, UG = H / 32, NMB = H / 128;
    constexpr uint32_t TCOLS = gt_tmem_cols(GT_ACC_COLS + (uint32_t)NMB * 96u);
    constexpr uint32_t G_BYTES = 2 * 2048;                                 // [2 blocks of 64 gate rows][16 rows x 128 B]
    extern __shared__ uint8_t gt_dyn[];
    __shared__ uint64_t bar_acc;
    __shared__ uint32_t tmem_slot;
    __shared__ float red[GT_NT / 32];
    const uint32_t base = (gt_smem(gt_dyn) + 1023u) & ~1023u;
    uint8_t* gop = gt_dyn + (base - gt_smem(gt_dyn));
    const uint32_t s_g = base;
    float* msk = reinterpret_cast<float*>(gop + G_BYTES);                  // [passes][T][8 environments] episode masks
    const int tid = threadIdx.x, lane = tid & 31, warp = tid >> 5;
    const bool gate = warp < GT_GW;
    const int ub = blockIdx.x * GT_UB, ku = ub + lane;
    const int ngrp = (d.N + GT_RW - 1) / GT_RW;
    gt_fill_masks(msk, d, use_msk, ngrp);
    auto mask_of = [&](int p, int t, int my) { return use_msk ? msk[(p * d.T + t) * GT_RW + warp] : gt_mask(d, t, my); };

    if (tid == 0) {
        gt_mbar_init(&bar_acc, 1);
        asm volatile("fence.mbarrier_init.release.cluster;" ::: "memory");
    }
    if (warp == GT_CTRL) {
        asm volatile("tcgen05.alloc.cta_group::1.sync.aligned.shared::cta.b32 [%0], %1;" ::"r"(gt_smem(&tmem_slot)), "r"(TCOLS));
        asm volatile("tcgen05.relinquish_alloc_permit.cta_group::1.sync.aligned;" ::);
    }
    const float sw = gt_build_weights(gop, Whh, H, ub, red, 0);           // scale only
    for (int i = tid; i < (int)(G_BYTES / 16); i += GT_NT) reinterpret_cast<uint4*>(gop)[i] = make_uint4(0u, 0u, 0u, 0u);
    asm volatile("fence.proxy.async.shared::cta;" ::: "memory");
    asm volatile("tcgen05.fence::before_thread_sync;" ::: "memory");
    __syncthreads();
    asm volatile("tcgen05.fence::after_thread_sync;" ::: "memory");
    const uint32_t tmem = tmem_slot;
    // thread (warp w < 4, lane) owns unit 128 mb + 32 w + lane of every block mb; gate row j <-> weight row
    // (j / 32) H + ub + j % 32: for a fixed j the lanes read consecutive columns (coalesced)
    if (warp < 4) {
        for (int mb = 0; mb < NMB; ++mb) {
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
                const uint32_t col = GT_ACC_COLS + (uint32_t)mb * 96u + (uint32_t)(j0 / 2);
                gt_st8(tmem + ((uint32_t)(32 * warp) << 16) + col, rh);
                gt_st8(tmem + ((uint32_t)(32 * warp) << 16) + col + 48u, rl);
            }
        }
        asm volatile("tcgen05.wait::st.sync.aligned;" ::: "memory");
    }
    asm volatile("tcgen05.fence::before_thread_sync;" ::: "memory");
    __syncthreads();
    asm volatile("tcgen05.fence::after_thread_sync;" ::: "memory");
    // D fp32, A fp16 K-major from tensor memory, B fp16 K-major, N = 16, M = 128
    const uint32_t idesc = (1u << 4) | ((uint32_t)(16 >> 3) << 17) | ((uint32_t)(128 >> 4) << 24);
    const uint64_t b0 = gt_desc(s_g, 16, 1024);                            // rows 8..15 (the lo parts): the next 8-row atom

    struct PhaseA { float dout, r, z, n, hm, ghn, mnext; };
    auto load_a = [&](int ps, int t, int my) {
        PhaseA a = {0.f, 0.f, 0.f, 0.f, 0.f, 0.f, 0.f};
        if (gate && my < d.N) {
            const long row = (long)t * d.N + my;
            a.dout = d.dout[row * H + ku];
            a.r = d.r[row * H + ku]; a.z = d.z[row * H + ku]; a.n = d.n[row * H + ku];
            a.hm = d.hm[row * H + ku];
            a.ghn = d.gh[row * 3 * H + 2 * H + ku];
            a.mnext = (t + 1 < d.T) ? mask_of(ps, t + 1, my) : 0.f;
        }
        return a;
    };
    auto part_of = [&](int grp, int buf) { return gt_part_all + GT_MAX_GROUPS + ((size_t)grp * 2 + buf) * GT_PART_FLOATS; };
    unsigned* arrived = reinterpret_cast<unsigned*>(gt_part_all);        // per group: partial slabs stored so far (a HINT)
    const int my0 = blockIdx.y * GT_RW + warp;
    PhaseA pre = load_a(0, d.T - 1, my0);
    float carry0 = 0.f;                                                    // dh_{t+1} z_{t+1} of the first pass's (environment, unit)
    uint32_t it = 0;

    for (int t = d.T - 1; t >= 0; --t) {
        for (int grp = blockIdx.y, ps = 0; grp < ngrp; grp += gridDim.y, ++ps) {
            const int my = grp * GT_RW + warp;
            const bool live = gate && my < d.N;
            const bool first = grp == (int)blockIdx.y;
            const long row = (long)t * d.N + my;
            const PhaseA a = first ? pre : load_a(ps, t, my);
            // next step's operands of the first pass: issued NOW, a whole step ahead of their use (they come from HBM:
            // issued at the end of the step, their latency sat in front of the next step -- 15 % of all stall samples)
            if (first && t > 0) pre = load_a(0, t - 1, my);
            // ---- carry of step t + 1: own dh z + the unit groups' partial products, summed in a fixed order ----
            float dh = a.dout;
            if (t + 1 < d.T) {
                // The arrival counter is bumped WITHOUT a release fence (a red.release.gpu cost ~1,100 cycles per step): it
                // only tells when polling the data is worth while -- one thread polls one word instead of 256 threads
                // polling 16 each, which slowed the producers' stores down.  Validity comes from the tags below.
                if (tid == 0) {
                    GT_STAMP(1, t, 0);
                    const unsigned target = (unsigned)UG * (unsigned)(d.T - 1 - t);
                    while (gt_ld_relaxed4(arrived + grp) < target) {}
                    GT_STAMP(1, t, 1);
                }
                __syncthreads();
                if (gate) {
                    // the 16 partial products of step t + 1 for this (environment, unit), one per unit group: strong loads,
                    // repeated until every word carries that step's tag in its two low mantissa bits
                    const float* p = part_of(grp, (t + 1) & 1) + (((size_t)blockIdx.x * UG) * GT_RW + warp) * GT_UB + lane;
                    const uint32_t tag = gt_tag(t + 1);
                    uint32_t w[UG];
#pragma unroll
                    for (int src = 0; src < UG; ++src) w[src] = gt_ld_relaxed4(p + (size_t)src * GT_RW * GT_UB);
#pragma unroll
                    for (int src = 0; src < UG; ++src)
                        while ((w[src] & 3u) != tag) w[src] = gt_ld_relaxed4(p + (size_t)src * GT_RW * GT_UB);
                    float sum = 0.f;
#pragma unroll
                    for (int src = 0; src < UG; src += 4)
                        sum += (__uint_as_float(w[src] & ~3u) + __uint_as_float(w[src + 1] & ~3u)) +
                               (__uint_as_float(w[src + 2] & ~3u) + __uint_as_float(w[src + 3] & ~3u));
                    const float c = (first ? carry0 : (live ? carry_ws[(long)my * H + ku] : 0.f)) + sum;
                    dh += c * a.mnext;
                }
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
                // ---- operand of the step: the CTA's gate gradients [8 environments x 96] as fp16 hi (rows 0..7) / lo
                //      (rows 8..15), scaled by a power of two chosen from this step's largest magnitude ----
                float m = warp_max(fmaxf(fabsf(g0), fmaxf(fabsf(g1), fabsf(g2))));
                if (lane == 0) red[warp] = m;
                __syncthreads();
                m = red[0];
#pragma unroll
                for (int w = 1; w < GT_GW; ++w) m = fmaxf(m, red[w]);
                const float sg = gt_pow2_scale(m, 12);
                if (gate) {
                    const float gv[3] = {g0, g1, g2};
#pragma unroll
                    for (int g = 0; g < 3; ++g) {
                        const uint32_t p = gt_split(gv[g] * sg);
                        *reinterpret_cast<unsigned short*>(gop + gt_swz(warp, g * 32 + lane, 2048)) = (unsigned short)(p & 0xffffu);
                        *reinterpret_cast<unsigned short*>(gop + gt_swz(warp + 8, g * 32 + lane, 2048)) = (unsigned short)(p >> 16);
                    }
                }
                asm volatile("fence.proxy.async.shared::cta;" ::: "memory");
                asm volatile("tcgen05.fence::before_thread_sync;" ::: "memory");
                if (tid == 0) GT_STAMP(1, t, 3);
                __syncthreads();
                if (tid == 0) GT_STAMP(1, t, 4);
                if (warp == GT_CTRL) {
                    asm volatile("tcgen05.fence::after_thread_sync;" ::: "memory");
                    if (gt_elect()) {
#pragma unroll
                        for (int mb = 0; mb < NMB; ++mb) {
                            const uint32_t dcol = tmem + (uint32_t)mb * 16u;
#pragma unroll
                            for (int j = 0; j < GT_ROWS / 16; ++j) {
                                const uint64_t b = b0 + (uint64_t)(((uint32_t)(j >> 2) * 2048u + (uint32_t)(j & 3) * 32u) >> 4);
                                const uint32_t ah = tmem + GT_ACC_COLS + (uint32_t)mb * 96u + 8u * (uint32_t)j, al = ah + 48u;
                                gt_mma_ts(dcol, al, b, idesc, j == 0 ? 0u : 1u);            // W_lo^T [g_hi | g_lo]
                                gt_mma_ts(dcol, ah, b, idesc, 1u);                          // W_hi^T [g_hi | g_lo]
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
                    const uint32_t tag = gt_tag(t);
                    float* pbase = part_of(grp, t & 1);
                    uint32_t v[NMB][16];
#pragma unroll
                    for (int mb = 0; mb < NMB; ++mb) gt_ld16_issue(tmem + ((uint32_t)(32 * warp) << 16) + (uint32_t)mb * 16u, v[mb]);
#pragma unroll
                    for (int mb = 0; mb < NMB; ++mb) gt_ld_wait16(v[mb]);
#pragma unroll
                    for (int mb = 0; mb < NMB; ++mb) {
                        float* p = pbase + ((size_t)((4 * mb + warp) * UG + blockIdx.x) * GT_RW) * GT_UB + lane;
#pragma unroll
                        for (int e = 0; e < 8; ++e) {
                            const float val = (__uint_as_float(v[mb][e]) + __uint_as_float(v[mb][e + 8])) * unscale;
                            __stcg(reinterpret_cast<uint32_t*>(p) + e * GT_UB, (__float_as_uint(val) & ~3u) | tag);
                        }
                    }
                    asm volatile("tcgen05.fence::before_thread_sync;" ::: "memory");
                    if (tid == 0) GT_STAMP(1, t, 6);
                }
                ++it;
                __syncthreads();
                if (tid == 0) GT_STAMP(1, t, 7);
                if (tid == 0) asm volatile("red.relaxed.gpu.global.add.u32 [%0], 1;" ::"l"(arrived + grp) : "memory");
                if (tid == 0) GT_STAMP(1, t, 8);
            }
            // ---- consumed by the weight-gradient products after this kernel ----
            if (live) {
                float* dgi = d.dgi + row * 3 * H;
                float* dgh = d.dgh + row * 3 * H;
                dgi[ku] = dr_pre; dgi[H + ku] = dz_pre; dgi[2 * H + ku] = dn_pre;
                dgh[ku] = dr_pre; dgh[H + ku] = dz_pre; dgh[2 * H + ku] = dn_pre * a.r;
            }
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
// state image + the lo k-blocks that do not fit tensor memory + exchange tail; at least 120 KiB: one CTA per SM (each
// allocates up to 512 tensor-memory columns)
constexpr size_t GT_MSK_MAX = 64 * 1024;      // largest mask table ([passes][T][8] floats) kept in shared memory
static size_t gt_smem_fwd(int H, size_t msk_bytes) {
    const int nkb = H / 64, lo_ts = 4 * nkb < 56 - 4 * nkb ? 4 * nkb : 56 - 4 * nkb;
    const size_t need = (size_t)2 * nkb * 1024 + (size_t)(nkb - lo_ts / 4) * GT_WBLK + 4096 + 1024 + msk_bytes;
    return need > 120 * 1024 ? need : 120 * 1024;
}
// the backward kernel needs 5 KiB; it asks for more than half an SM's shared memory so that two CTAs (512 tensor-memory
// columns each) can never be placed on one SM
static size_t gt_smem_bwd(int, size_t msk_bytes) { return 120 * 1024 > 8192 + msk_bytes ? 120 * 1024 : 8192 + msk_bytes; }

template <int NKB>
static const void* gt_kernel(bool backward) {
    return backward ? (const void*)gru_tc_bwd_kernel<NKB> : (const void*)gru_tc_fwd_kernel<NKB>;
}
static const void* gt_kernel_for(int H, bool backward) {
    switch (H / 64) {
        case 2: return gt_kernel<2>(backward);
        case 4: return gt_kernel<4>(backward);
        case 6: return gt_kernel<6>(backward);
        default: return gt_kernel<8>(backward);
    }
}

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
    if (gt_smem_fwd(H, GT_MSK_MAX) > 226 * 1024 || gt_smem_bwd(H, GT_MSK_MAX) > 226 * 1024) return false;
    return H / GT_UB <= gt_sms();
}

int gru_tc(int backward, const GruDesc& d, const float* Whh, const float* bhh, cudaStream_t st) {
    A2CB_REQUIRE(gru_tc_supported(d.N, d.H), "gru_tc: unsupported shape (N=%d H=%d)", d.N, d.H);
    A2CB_REQUIRE(Whh && d.hm && d.masks, "gru_tc: null pointer");
    A2CB_REQUIRE(backward || (bhh && d.h0 && d.gi && d.out), "gru_tc: forward pointers");
    A2CB_REQUIRE(!backward || (d.r && d.z && d.n && d.gh && d.dout && d.dgi && d.dgh), "gru_tc: backward pointers");
    static bool attr_done[2][GT_MAX_H / 64 + 1] = {};
    if (!attr_done[backward ? 1 : 0][d.H / 64]) {
        const size_t most = backward ? gt_smem_bwd(d.H, GT_MSK_MAX) : gt_smem_fwd(d.H, GT_MSK_MAX);
        const cudaError_t ea = cudaFuncSetAttribute(gt_kernel_for(d.H, backward != 0), cudaFuncAttributeMaxDynamicSharedMemorySize, (int)most);
        if (ea != cudaSuccess) {
            set_error("gru_tc: shared memory attribute (%zu bytes): %s", most, cudaGetErrorString(ea));
            cudaGetLastError();
            return A2CB_ERR_CUDA;
        }
        attr_done[backward ? 1 : 0][d.H / 64] = true;
    }
    const int ug = d.H / GT_UB;
    const int ngrp = (d.N + GT_RW - 1) / GT_RW;
    int gy = gt_sms() / ug;
    gy = gy < 1 ? 1 : (gy > ngrp ? ngrp : gy);
    A2CB_REQUIRE(!backward || ngrp <= gy || d.dcarry, "gru_tc: dcarry workspace [N, H] needed for N > %d", gy * GT_RW);
    {
        // exchange buffers of the launch's groups: tag 0 = "never written" (the kernels validate every word they read)
        void* ptr = nullptr;
        cudaError_t e0 = backward ? cudaGetSymbolAddress(&ptr, gt_part_all) : cudaGetSymbolAddress(&ptr, gt_img);
        const size_t bytes = backward ? (GT_MAX_GROUPS + (size_t)ngrp * 2 * GT_PART_FLOATS) * sizeof(float) : (size_t)ngrp * 2 * GT_IMG_BYTES;
        if (e0 == cudaSuccess) e0 = cudaMemsetAsync(ptr, 0, bytes, st);
        if (e0 != cudaSuccess) { set_error("gru_tc: exchange buffer reset: %s", cudaGetErrorString(e0)); return A2CB_ERR_CUDA; }
    }
    static int dbg = -1;
    if (dbg < 0) { const char* e = getenv("A2CB_GRU_TC_DEBUG"); dbg = (e && e[0] == '1') ? 1 : 0; }
    GruDesc dd = d;
    float* ws = d.dcarry;
    // episode-mask table in shared memory when it fits: [passes per CTA][T][8 environments]
    size_t msk_bytes = (size_t)((ngrp + gy - 1) / gy) * (size_t)d.T * GT_RW * sizeof(float);
    int use_msk = msk_bytes <= GT_MSK_MAX ? 1 : 0;
    if (!use_msk) msk_bytes = 0;
    void* args_f[] = {(void*)&dd, (void*)&Whh, (void*)&bhh, (void*)&use_msk, (void*)&dbg};
    void* args_b[] = {(void*)&dd, (void*)&Whh, (void*)&ws, (void*)&use_msk, (void*)&dbg};
    dim3 grid((unsigned)ug, (unsigned)gy);
    cudaError_t e = cudaLaunchCooperativeKernel(gt_kernel_for(d.H, backward != 0), grid, dim3(GT_NT), backward ? args_b : args_f,
                                                backward ? gt_smem_bwd(d.H, msk_bytes) : gt_smem_fwd(d.H, msk_bytes), st);
    if (e != cudaSuccess) {
        set_error("gru_tc: cooperative launch failed: %s", cudaGetErrorString(e));
        cudaGetLastError();
        return A2CB_ERR_CUDA;
    }
    return check_launch("gru_tc");
}

}  // namespace a2cb
