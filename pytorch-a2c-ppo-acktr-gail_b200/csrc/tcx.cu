// TMA-fed tensor-core engine (tcgen05 / TMEM, sm_100a) for the dense layers of the policy update:
// CNNBase convolutions as per-tap GEMMs straight from NHWC activations, Linear layers, and all of their
// data / weight gradients (reference a2c_ppo_acktr/model.py:169-195 forward; autograd backward).
//
// Why a second engine next to gemm_tc.cu: that kernel gathers operands with SIMT loads (implicit
// im2col through affine index maps), splits every fp32 into TF32 hi/lo in registers and stores them to
// shared memory -- ~950 instructions per warp and k-block, which caps it at ~20-30 TFLOP/s.  Here
//   * activations live in HBM as two planes: x (plain fp32) and lo = x - trunc_tf32(x), written ONCE by
//     the producing epilogue.  kind::tf32 ignores the 13 low mantissa bits of its 32-bit operands, so the
//     x plane itself is the "hi" operand and  x*w = hi_x*hi_w + lo_x*hi_w + hi_x*lo_w  (+ ~2^-20) costs
//     three MMAs with no conversion pass at all;
//   * a convolution is a sum over filter taps of [pixels x C_in] x [C_in x C_out] products; the pixel
//     box of one tap (for a whole tile of images) is ONE multi-dimensional TMA box of the NHWC tensor
//     (stride-2 layers through a parity view, transposed convolutions through TMA's out-of-bounds zero
//     fill), landing in shared memory as 128-byte swizzled rows -- exactly the K-major UMMA operand;
//   * weight gradients reduce over pixels, i.e. both operands are "MN-major" ([pixels][channels] with
//     the reduction along rows); the same TMA boxes feed them through MN-major UMMA descriptors
//     (32-byte-atom 128B swizzle, the one layout the tensor core accepts for transposed 32-bit operands);
//   * warp-specialised CTA: warp 0 = TMA producer, warp 1 = MMA issuer (one thread), warps 2-9 =
//     accumulator drain + epilogue.  The tensor core's fp32 accumulation truncates (error grows with the
//     reduction length, tools/tc_accuracy.py), so the reduction is cut into chains that alternate between
//     two TMEM accumulators; the epilogue warps add finished chains into fp32 registers while the next
//     chain runs.
#include <cuda.h>
#include <stdlib.h>

#include "common.cuh"
#include "decls.cuh"
#include "tcx.cuh"

namespace a2cb {

namespace {

constexpr int TX_NT = 320;                    // warp 0 TMA, warp 1 MMA, warps 2..9 epilogue
constexpr int TX_EPI0 = 2;                    // first epilogue warp
constexpr int TX_ROWB = 128;                  // bytes per operand row (32 floats, one swizzle span)
constexpr int TX_MAX_STAGES = 6;
constexpr int TX_EPI_STAGING = 8 * 2048 + 4096;   // forward kernel: epilogue transposition slabs + row offsets (1 KiB multiple)

TcxKnobs g_knobs = {1 /* SWIZZLE_128B_BASE32B */, 512, 1024, (int)CU_TENSOR_MAP_SWIZZLE_128B_ATOM_32B};

__device__ __forceinline__ uint32_t tx_smem(const void* p) { return (uint32_t)__cvta_generic_to_shared(p); }

__device__ __forceinline__ void tx_mbar_init(uint64_t* bar, uint32_t count) {
    asm volatile("mbarrier.init.shared::cta.b64 [%0], %1;" ::"r"(tx_smem(bar)), "r"(count));
}
__device__ __forceinline__ void tx_mbar_wait(uint64_t* bar, uint32_t parity) {
    uint32_t ok = 0;
    while (!ok) {
        asm volatile(
            "{\n"
            ".reg .pred p;\n"
            "mbarrier.try_wait.parity.shared::cta.b64 p, [%1], %2;\n"
            "selp.u32 %0, 1, 0, p;\n"
            "}\n"
            : "=r"(ok)
            : "r"(tx_smem(bar)), "r"(parity)
            : "memory");
    }
}
__device__ __forceinline__ void tx_mbar_arrive(uint64_t* bar) {
    asm volatile("mbarrier.arrive.shared::cta.b64 _, [%0];" ::"r"(tx_smem(bar)) : "memory");
}
__device__ __forceinline__ void tx_expect(uint64_t* bar, uint32_t bytes) {
    asm volatile("mbarrier.arrive.expect_tx.shared::cta.b64 _, [%0], %1;" ::"r"(tx_smem(bar)), "r"(bytes) : "memory");
}
__device__ __forceinline__ void tx_tma5(uint32_t dst, const CUtensorMap* tm, uint64_t* bar, int c0, int c1, int c2,
                                        int c3, int c4) {
    asm volatile(
        "cp.async.bulk.tensor.5d.shared::cluster.global.tile.mbarrier::complete_tx::bytes [%0], [%1, {%3, %4, %5, %6, %7}], [%2];" ::
            "r"(dst), "l"(reinterpret_cast<uint64_t>(tm)), "r"(tx_smem(bar)), "r"(c0), "r"(c1), "r"(c2), "r"(c3), "r"(c4)
        : "memory");
}
__device__ __forceinline__ void tx_commit(uint64_t* bar) {
    asm volatile("tcgen05.commit.cta_group::1.mbarrier::arrive::one.shared::cluster.b64 [%0];" ::"r"(tx_smem(bar))
                 : "memory");
}
__device__ __forceinline__ void tx_mma(uint32_t d_tmem, uint64_t a, uint64_t b, uint32_t idesc, uint32_t acc) {
    asm volatile(
        "{\n\t"
        ".reg .pred p;\n\t"
        "setp.ne.b32 p, %4, 0;\n\t"
        "tcgen05.mma.cta_group::1.kind::tf32 [%0], %1, %2, %3, p;\n\t"
        "}\n" ::"r"(d_tmem), "l"(a), "l"(b), "r"(idesc), "r"(acc)
        : "memory");
}
__device__ __forceinline__ void tx_mma_f16(uint32_t d_tmem, uint64_t a, uint64_t b, uint32_t idesc, uint32_t acc) {
    asm volatile(
        "{\n\t"
        ".reg .pred p;\n\t"
        "setp.ne.b32 p, %4, 0;\n\t"
        "tcgen05.mma.cta_group::1.kind::f16 [%0], %1, %2, %3, p;\n\t"
        "}\n" ::"r"(d_tmem), "l"(a), "l"(b), "r"(idesc), "r"(acc)
        : "memory");
}
// shared-memory matrix descriptor (cute::UMMA::SmemDescriptor): start >> 4 | LBO >> 4 << 16 | SBO >> 4 << 32 |
// version 1 << 46 | layout type << 61
__device__ __forceinline__ uint64_t tx_desc(uint32_t addr, uint32_t lbo_bytes, uint32_t sbo_bytes, uint32_t layout) {
    return (uint64_t)((addr & 0x3FFFF) >> 4) | ((uint64_t)((lbo_bytes >> 4) & 0x3FFF) << 16) |
           ((uint64_t)((sbo_bytes >> 4) & 0x3FFF) << 32) | (1ull << 46) | ((uint64_t)layout << 61);
}
// One leader lane of a fully converged warp.  The single-thread instructions (TMA, tcgen05.mma / commit) sit under
// this predicate while the WHOLE warp runs the surrounding loop: with the loop itself inside `if (lane == 0)` the
// compiler cannot prove a single active thread and wraps every UTCHMMA / UTMALDG in an elect-and-loop sequence
// (~6 extra instructions and a branch per MMA -- the issuing thread, not the tensor pipe, became the bottleneck).
__device__ __forceinline__ bool tx_elect() {
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
// K-major SWIZZLE_128B descriptor with SBO = 1024 B: only the start address varies
__device__ __forceinline__ uint64_t tx_desc_k(uint32_t addr) {
    constexpr uint64_t kConst = ((uint64_t)1 << 16) | ((uint64_t)(1024 >> 4) << 32) | (1ull << 46) | (2ull << 61);
    return kConst | (uint64_t)((addr & 0x3FFFF) >> 4);
}
__device__ __forceinline__ void tx_ld16(uint32_t taddr, uint32_t* v) {
    asm volatile(
        "tcgen05.ld.sync.aligned.32x32b.x16.b32"
        "{%0, %1, %2, %3, %4, %5, %6, %7, %8, %9, %10, %11, %12, %13, %14, %15}, [%16];\n"
        : "=r"(v[0]), "=r"(v[1]), "=r"(v[2]), "=r"(v[3]), "=r"(v[4]), "=r"(v[5]), "=r"(v[6]), "=r"(v[7]), "=r"(v[8]),
          "=r"(v[9]), "=r"(v[10]), "=r"(v[11]), "=r"(v[12]), "=r"(v[13]), "=r"(v[14]), "=r"(v[15])
        : "r"(taddr));
    asm volatile("tcgen05.wait::ld.sync.aligned;" ::: "memory");
}
// the same load without the wait, and a wait that "touches" the destination registers so that the compiler cannot
// move their uses above it: several loads in flight per wait
__device__ __forceinline__ void tx_ld16_issue(uint32_t taddr, uint32_t* v) {
    asm volatile(
        "tcgen05.ld.sync.aligned.32x32b.x16.b32"
        "{%0, %1, %2, %3, %4, %5, %6, %7, %8, %9, %10, %11, %12, %13, %14, %15}, [%16];\n"
        : "=r"(v[0]), "=r"(v[1]), "=r"(v[2]), "=r"(v[3]), "=r"(v[4]), "=r"(v[5]), "=r"(v[6]), "=r"(v[7]), "=r"(v[8]),
          "=r"(v[9]), "=r"(v[10]), "=r"(v[11]), "=r"(v[12]), "=r"(v[13]), "=r"(v[14]), "=r"(v[15])
        : "r"(taddr));
}
__device__ __forceinline__ void tx_ld_wait16(uint32_t* v) {
    asm volatile("tcgen05.wait::ld.sync.aligned;"
                 : "+r"(v[0]), "+r"(v[1]), "+r"(v[2]), "+r"(v[3]), "+r"(v[4]), "+r"(v[5]), "+r"(v[6]), "+r"(v[7]), "+r"(v[8]),
                   "+r"(v[9]), "+r"(v[10]), "+r"(v[11]), "+r"(v[12]), "+r"(v[13]), "+r"(v[14]), "+r"(v[15])
                 :
                 : "memory");
}
__device__ __forceinline__ float tx_lo(float v) { return v - __uint_as_float(__float_as_uint(v) & 0xffffe000u); }

struct TxBars {
    uint64_t full[TX_MAX_STAGES], empty[TX_MAX_STAGES], acc_full[2], acc_empty[2];
    uint64_t fullB[TX_MAX_STAGES], emptyB[TX_MAX_STAGES];   // weight-gradient kernel: ring of the operand shared by all
                                                            // groups; TF32 patch mode: ring of the weight k-blocks
    uint32_t tmem_base;
};

__device__ __forceinline__ void tx_tile_base(const TcxTiling& t, int tile, int* base) {
    // selects instead of base[t.tr_dim] += ...: a dynamically indexed array lives in local memory
    const int tq = tile / t.t_div, tr = tile - tq * t.t_div;
    const int vr = tr * t.tr_step, vq = tq * t.tq_step;
    base[0] = 0;
#pragma unroll
    for (int d = 1; d < 5; ++d) base[d] = (t.tr_dim == d ? vr : 0) + (t.tq_dim == d ? vq : 0);
}

// ---------------------------------------------------------------------------------------------------
// forward-like: out tile [R <= 128 * MB pixels] x [BN channels], reduction over n_kb k-blocks of 32
// ---------------------------------------------------------------------------------------------------
// PATCH (bf16 mode, stride-1 taps): the taps of a tile are row-shifted windows of ONE input patch (box = the tile's
// pixels plus the tap halo, one image row = box[1] pixels), so the patch is loaded ONCE per tile and every tap's A
// operand is a UMMA descriptor whose start address is advanced by (dy * box[1] + dx) 128-byte rows inside it (the
// swizzle is a function of the absolute shared-memory address, for TMA's writes and the tensor core's reads alike).  Accumulator row m
// is patch pixel m: the columns / rows past the tile's valid extent are junk and never stored.  The weight image
// (all k-blocks, three planes) is loaded once per CTA and stays resident.  Bytes staged per tile: one patch instead
// of n_kb boxes + n_kb weight tiles.
template <int BN, int MB, bool BF16, bool PATCH = false>
__global__ void __launch_bounds__(TX_NT, 1)
tcx_fwd_kernel(const __grid_constant__ CUtensorMap tmA, const __grid_constant__ CUtensorMap tmAlo,
               const __grid_constant__ CUtensorMap tmB, const __grid_constant__ CUtensorMap tmBlo,
               const __grid_constant__ CUtensorMap tmB3, const __grid_constant__ TcxFwd p, const int stages,
               const int patch_base_offset, const int stagesB, const int merge_on) {
    constexpr int B_BYTES = BN * TX_ROWB;
    constexpr int HALF = BN / 2;
    // MERGE (<= 64 output channels): the weight planes of a stage are contiguous in shared memory, so ONE instruction
    // A x [B_hi | B_lo] (N = 2 BN; bf16: [b1 | b2 | b3], N = 3 BN) replaces two (three): an instruction with its A operand
    // in shared memory costs >= 46 cycles whatever N is (the 4 KiB A read, tools/mma_probe.cu), and A is read once per
    // plane instead of once per product.  The products land in NPL column blocks of the accumulator; the drain adds them.
    constexpr bool MERGE_T = BN <= 64;
    // A2CB_TCX_MERGE (A / B knob, default 7): bit 0 = bf16 instances, bit 1 = 32-channel TF32, bit 2 = 64-channel TF32
    const bool MERGE = MERGE_T && ((merge_on >> (BF16 ? 0 : (BN == 32 ? 1 : 2))) & 1) != 0;
    constexpr int NPL = MERGE_T ? (BF16 ? 3 : 2) : 1;              // column blocks per accumulator
    constexpr int W = NPL * BN;                                  // accumulator width in tensor-memory columns
    constexpr uint32_t TCOLS = (2 * MB * W <= 32) ? 32 : (2 * MB * W <= 64) ? 64 : (2 * MB * W <= 128) ? 128 : (2 * MB * W <= 256) ? 256 : 512;
    static_assert(2 * MB * W <= 512, "tcx_fwd: accumulators exceed tensor memory");
    extern __shared__ uint8_t tx_dyn[];
    __shared__ TxBars bars;
    // one plane of the A tile: the R rows TMA writes, rounded up to a swizzle atom (8 rows).  The last M block's MMAs
    // read up to 128 * MB rows, i.e. run on into the following operand planes of the stage: finite garbage that only
    // reaches accumulator rows >= R, which are never stored.
    const int A_BYTES = ((p.A.box[1] * p.A.box[2] * p.A.box[3] * p.A.box[4] + 7) & ~7) * TX_ROWB;
    const uint32_t dyn0 = (tx_smem(tx_dyn) + 1023u) & ~1023u;
    uint8_t* stage_base = tx_dyn + (dyn0 - tx_smem(tx_dyn));      // epilogue staging: 8 x 4 KiB slabs + row offsets
    const uint32_t dyn = dyn0 + TX_EPI_STAGING;                   // operand pipeline stages
    const int warp = threadIdx.x >> 5, lane = threadIdx.x & 31;
    const bool has_lo = !BF16 && p.A.lo != nullptr;
    constexpr int NB = BF16 ? 3 : 2;                              // planes of the weight tile
    // PATCH: a stage is one patch (all planes).  bf16: the weight image is resident ahead of the patch ring;
    // TF32: the weight k-blocks stream through their own ring behind it.
    const int stage_bytes = PATCH ? A_BYTES * (has_lo ? 2 : 1) : A_BYTES * (has_lo ? 2 : 1) + NB * B_BYTES;
    const int RBOX = p.A.box[1] * p.A.box[2] * p.A.box[3] * p.A.box[4];      // rows TMA writes
    const int R = PATCH ? p.patch : RBOX;                                      // accumulator rows the epilogue stores
    const uint32_t dynB = (PATCH && !BF16) ? dyn + (uint32_t)(stages * stage_bytes) : dyn;
    const uint32_t dynA = dyn + ((PATCH && BF16) ? (uint32_t)(p.n_kb * NB * B_BYTES) : 0u);
    const int chain = p.chain > 0 ? p.chain : p.n_kb;
    const int n_chains = (p.n_kb + chain - 1) / chain;
    // persistent: work item w = tile_m * n_n + tile_n, strided over the grid.  The three roles walk the same
    // sequence with running k-block / chain counters, so the epilogue of one tile overlaps the loads and MMAs
    // of the next (the two TMEM accumulators double as the hand-over buffers).
    const int n_n = (p.N + BN - 1) / BN;
    const int n_work = p.tiles.n_tiles * n_n;

    if (threadIdx.x == 0) {
        for (int s = 0; s < stages; ++s) { tx_mbar_init(&bars.full[s], 1); tx_mbar_init(&bars.empty[s], 1); }
        for (int b = 0; b < 2; ++b) { tx_mbar_init(&bars.acc_full[b], 1); tx_mbar_init(&bars.acc_empty[b], 256); }
        for (int s = 0; s < TX_MAX_STAGES; ++s) { tx_mbar_init(&bars.fullB[s], 1); tx_mbar_init(&bars.emptyB[s], 1); }
        asm volatile("fence.mbarrier_init.release.cluster;" ::: "memory");
    }
    if (warp == 1) {
        asm volatile("tcgen05.alloc.cta_group::1.sync.aligned.shared::cta.b32 [%0], %1;" ::"r"(tx_smem(&bars.tmem_base)),
                     "r"(TCOLS));
        asm volatile("tcgen05.relinquish_alloc_permit.cta_group::1.sync.aligned;" ::);
    }
    asm volatile("tcgen05.fence::before_thread_sync;" ::: "memory");
    __syncthreads();
    asm volatile("tcgen05.fence::after_thread_sync;" ::: "memory");
    const uint32_t tmem = bars.tmem_base;

    if (warp == 0) {
        // ===== TMA producer: the whole warp walks the loop, one elected lane issues =====
        const uint32_t bytes = (uint32_t)(RBOX * TX_ROWB * (has_lo ? 2 : 1) + NB * B_BYTES);
        int it = 0, s = 0;
        uint32_t ph = 0;                                   // parity of the NEXT completion of empty[s] to wait for
        if constexpr (PATCH && !BF16) {
            // patches (one per run of k-blocks with the same channel offset) and weight k-blocks, issued in the
            // order the MMA warp consumes them -- two rings, one program order, so neither can starve the other
            int itB = 0, sB = 0;
            uint32_t phB = 0;
            for (int w = blockIdx.x; w < n_work; w += gridDim.x) {
                int base[5];
                tx_tile_base(p.tiles, w, base);
                if (p.gather_dim > 0) base[p.gather_dim] = (int)__ldg(p.gather_idx + base[p.gather_dim]);
                for (int kb = 0; kb < p.n_kb; ++kb, ++itB) {
                    if (kb == 0 || p.kb_a[kb][0] != p.kb_a[kb - 1][0]) {
                        if (it >= stages) tx_mbar_wait(&bars.empty[s], ph ^ 1u);
                        if (tx_elect()) {
                            const uint32_t sa = dynA + (uint32_t)(s * stage_bytes);
                            tx_expect(&bars.full[s], (uint32_t)(RBOX * TX_ROWB * (has_lo ? 2 : 1)));
                            tx_tma5(sa, &tmA, &bars.full[s], p.kb_a[kb][0], base[1], base[2], base[3], base[4]);
                            if (has_lo) tx_tma5(sa + A_BYTES, &tmAlo, &bars.full[s], p.kb_a[kb][0], base[1], base[2], base[3], base[4]);
                        }
                        __syncwarp();
                        ++it;
                        if (++s == stages) { s = 0; ph ^= 1u; }
                    }
                    if (itB >= stagesB) tx_mbar_wait(&bars.emptyB[sB], phB ^ 1u);
                    if (tx_elect()) {
                        const uint32_t sb = dynB + (uint32_t)(sB * NB * B_BYTES);
                        tx_expect(&bars.fullB[sB], (uint32_t)(NB * B_BYTES));
                        tx_tma5(sb, &tmB, &bars.fullB[sB], p.kb_b[kb], 0, 0, 0, 0);
                        tx_tma5(sb + B_BYTES, &tmBlo, &bars.fullB[sB], p.kb_b[kb], 0, 0, 0, 0);
                    }
                    __syncwarp();
                    if (++sB == stagesB) { sB = 0; phB ^= 1u; }
                }
            }
        } else if constexpr (PATCH) {
            if (tx_elect()) {
                tx_expect(&bars.fullB[0], (uint32_t)(p.n_kb * NB * B_BYTES));
                for (int kb = 0; kb < p.n_kb; ++kb) {
                    const uint32_t sb = dynB + (uint32_t)(kb * NB * B_BYTES);
                    tx_tma5(sb, &tmB, &bars.fullB[0], p.kb_b[kb], 0, 0, 0, 0);
                    tx_tma5(sb + B_BYTES, &tmBlo, &bars.fullB[0], p.kb_b[kb], 0, 0, 0, 0);
                    tx_tma5(sb + 2 * B_BYTES, &tmB3, &bars.fullB[0], p.kb_b[kb], 0, 0, 0, 0);
                }
            }
            __syncwarp();
            for (int w = blockIdx.x; w < n_work; w += gridDim.x, ++it) {
                int base[5];
                tx_tile_base(p.tiles, w, base);
                if (p.gather_dim > 0) base[p.gather_dim] = (int)__ldg(p.gather_idx + base[p.gather_dim]);
                if (it >= stages) tx_mbar_wait(&bars.empty[s], ph ^ 1u);
                if (tx_elect()) {
                    tx_expect(&bars.full[s], (uint32_t)(RBOX * TX_ROWB));
                    tx_tma5(dynA + (uint32_t)(s * stage_bytes), &tmA, &bars.full[s], 0, base[1], base[2], base[3], base[4]);
                }
                __syncwarp();
                if (++s == stages) { s = 0; ph ^= 1u; }
            }
        } else
        for (int w = blockIdx.x; w < n_work; w += gridDim.x) {
            int base[5];
            tx_tile_base(p.tiles, w / n_n, base);
            if (p.gather_dim > 0) base[p.gather_dim] = (int)__ldg(p.gather_idx + base[p.gather_dim]);
            const int n0 = (w % n_n) * BN;
            for (int kb = 0; kb < p.n_kb; ++kb, ++it) {
                if (it >= stages) tx_mbar_wait(&bars.empty[s], ph ^ 1u);
                if (tx_elect()) {
                    const uint32_t sa = dyn + (uint32_t)(s * stage_bytes);
                    const uint32_t sb = sa + (uint32_t)(A_BYTES * (has_lo ? 2 : 1));
                    tx_expect(&bars.full[s], bytes);
                    const int c0 = p.kb_a[kb][0], c1 = p.kb_a[kb][1] + base[1], c2 = p.kb_a[kb][2] + base[2],
                              c3 = p.kb_a[kb][3] + base[3], c4 = base[4];
                    tx_tma5(sa, &tmA, &bars.full[s], c0, c1, c2, c3, c4);
                    if (has_lo) tx_tma5(sa + A_BYTES, &tmAlo, &bars.full[s], c0, c1, c2, c3, c4);
                    tx_tma5(sb, &tmB, &bars.full[s], p.kb_b[kb], n0, 0, 0, 0);
                    tx_tma5(sb + B_BYTES, &tmBlo, &bars.full[s], p.kb_b[kb], n0, 0, 0, 0);
                    if (BF16) tx_tma5(sb + 2 * B_BYTES, &tmB3, &bars.full[s], p.kb_b[kb], n0, 0, 0, 0);
                }
                __syncwarp();
                if (++s == stages) { s = 0; ph ^= 1u; }
            }
        }
    } else if (warp == 1) {
        // ===== MMA issuer: the whole warp waits on the barriers, one elected lane issues the MMAs =====
        // D fp32, A/B tf32 (format 2) or bf16 (format 1), both K-major, N = BN, M = 128
        constexpr uint32_t fmt = BF16 ? 1u : 2u;
        const uint32_t idesc = (1u << 4) | (fmt << 7) | (fmt << 10) | ((uint32_t)(BN >> 3) << 17) | ((uint32_t)(128 >> 4) << 24);
        const uint32_t idescW = (1u << 4) | (fmt << 7) | (fmt << 10) | ((uint32_t)(W >> 3) << 17) | ((uint32_t)(128 >> 4) << 24);   // N = W
        int s = 0, cc = 0;
        uint32_t ph = 0;                                   // parity of full[s] to wait for
        if constexpr (PATCH && !BF16) {
            int sB = 0;
            uint32_t phB = 0;
            const int pitch1 = p.A.box[1], pitch2 = p.A.box[1] * p.A.box[2];
            for (int w = blockIdx.x; w < n_work; w += gridDim.x) {
                int kc = 0, c = 0;
                for (int kb = 0; kb < p.n_kb; ++kb) {
                    const bool grp_first = kb == 0 || p.kb_a[kb][0] != p.kb_a[kb - 1][0];
                    const bool grp_last = kb == p.n_kb - 1 || p.kb_a[kb + 1][0] != p.kb_a[kb][0];
                    const int g = cc + c, buf = g & 1;
                    const bool first = kc == 0;
                    if (first && g >= 2) tx_mbar_wait(&bars.acc_empty[buf], (uint32_t)(((g >> 1) - 1) & 1));
                    if (grp_first) tx_mbar_wait(&bars.full[s], ph);
                    tx_mbar_wait(&bars.fullB[sB], phB);
                    asm volatile("tcgen05.fence::after_thread_sync;" ::: "memory");
                    const bool last = (kb == p.n_kb - 1) || (kc + 1 == chain);
                    if (tx_elect()) {
                        // tap shift of this k-block inside the patch, in whole 128-byte rows
                        const int shift = p.kb_a[kb][1] + p.kb_a[kb][2] * pitch1 + p.kb_a[kb][3] * pitch2;
                        const uint32_t sa = dynA + (uint32_t)(s * stage_bytes) + (uint32_t)(shift * TX_ROWB);
                        const uint32_t sb = dynB + (uint32_t)(sB * NB * B_BYTES);
#pragma unroll
                        for (int mb = 0; mb < MB; ++mb) {
                            const uint32_t d = tmem + (uint32_t)(buf * MB * W + mb * W);
                            const uint32_t ah = sa + (uint32_t)(mb * 128 * TX_ROWB);
#pragma unroll
                            for (int ks = 0; ks < 4; ++ks) {
                                const uint32_t o = ks * 32;
                                const uint64_t dah = tx_desc_k(ah + o), dbh = tx_desc_k(sb + o), dbl = tx_desc_k(sb + B_BYTES + o);
                                const uint32_t acc0 = (first && ks == 0) ? 0u : 1u;
                                if (MERGE) {
                                    tx_mma(d, dah, dbh, idescW, acc0);                       // A_hi x [B_hi | B_lo]
                                    if (has_lo) tx_mma(d, tx_desc_k(ah + A_BYTES + o), dbh, idesc, 1u);
                                    continue;
                                }
                                if (has_lo) { tx_mma(d, tx_desc_k(ah + A_BYTES + o), dbh, idesc, acc0); tx_mma(d, dah, dbl, idesc, 1u); }
                                else tx_mma(d, dah, dbl, idesc, acc0);
                                tx_mma(d, dah, dbh, idesc, 1u);
                            }
                        }
                        tx_commit(&bars.emptyB[sB]);
                        if (grp_last) tx_commit(&bars.empty[s]);
                        if (last) tx_commit(&bars.acc_full[buf]);
                    }
                    __syncwarp();
                    if (++sB == stagesB) { sB = 0; phB ^= 1u; }
                    if (grp_last) { if (++s == stages) { s = 0; ph ^= 1u; } }
                    if (last) { kc = 0; ++c; } else ++kc;
                }
                cc += n_chains;
            }
        } else if constexpr (PATCH) {
            tx_mbar_wait(&bars.fullB[0], 0u);
            for (int w = blockIdx.x; w < n_work; w += gridDim.x) {
                tx_mbar_wait(&bars.full[s], ph);
                asm volatile("tcgen05.fence::after_thread_sync;" ::: "memory");
                int kc = 0, c = 0;
                for (int kb = 0; kb < p.n_kb; ++kb) {
                    const int g = cc + c, buf = g & 1;
                    const bool first = kc == 0;
                    if (first && g >= 2) {
                        tx_mbar_wait(&bars.acc_empty[buf], (uint32_t)(((g >> 1) - 1) & 1));
                        asm volatile("tcgen05.fence::after_thread_sync;" ::: "memory");
                    }
                    const bool last = (kb == p.n_kb - 1) || (kc + 1 == chain);
                    if (tx_elect()) {
                        // tap (dx, dy) of this k-block: rows shifted by dy * (pixels per patch row) + dx
                        const uint32_t sa = dynA + (uint32_t)(s * stage_bytes) +
                                            (uint32_t)((p.kb_a[kb][1] + p.kb_a[kb][2] * p.A.box[1]) * TX_ROWB);
                        const uint32_t sb = dynB + (uint32_t)(kb * NB * B_BYTES);
#pragma unroll
                        for (int mb = 0; mb < MB; ++mb) {
                            const uint32_t d = tmem + (uint32_t)(buf * MB * W + mb * W);
                            const uint32_t ah = sa + (uint32_t)(mb * 128 * TX_ROWB);
                            const uint64_t bo = patch_base_offset ? ((uint64_t)((ah >> 7) & 7u) << 49) : 0ull;
#pragma unroll
                            for (int ks = 0; ks < 4; ++ks) {
                                const uint32_t o = ks * 32;
                                const uint64_t dah = tx_desc_k(ah + o) | bo;
                                if (MERGE) {                                                 // A x [b1 | b2 | b3]
                                    tx_mma_f16(d, dah, tx_desc_k(sb + o), idescW, (first && ks == 0) ? 0u : 1u);
                                    continue;
                                }
                                tx_mma_f16(d, dah, tx_desc_k(sb + 2 * B_BYTES + o), idesc, (first && ks == 0) ? 0u : 1u);
                                tx_mma_f16(d, dah, tx_desc_k(sb + B_BYTES + o), idesc, 1u);
                                tx_mma_f16(d, dah, tx_desc_k(sb + o), idesc, 1u);
                            }
                        }
                        if (kb == p.n_kb - 1) tx_commit(&bars.empty[s]);
                        if (last) tx_commit(&bars.acc_full[buf]);
                    }
                    __syncwarp();
                    if (last) { kc = 0; ++c; } else ++kc;
                }
                if (++s == stages) { s = 0; ph ^= 1u; }
                cc += n_chains;
            }
        } else
        for (int w = blockIdx.x; w < n_work; w += gridDim.x) {
            int kc = 0, c = 0;                             // position inside the chain, chain index inside the tile
            for (int kb = 0; kb < p.n_kb; ++kb) {
                const int g = cc + c, buf = g & 1;
                const bool first = kc == 0;
                if (first && g >= 2) tx_mbar_wait(&bars.acc_empty[buf], (uint32_t)(((g >> 1) - 1) & 1));
                tx_mbar_wait(&bars.full[s], ph);
                asm volatile("tcgen05.fence::after_thread_sync;" ::: "memory");
                const bool last = (kb == p.n_kb - 1) || (kc + 1 == chain);
                if (tx_elect()) {
                    const uint32_t sa = dyn + (uint32_t)(s * stage_bytes);
                    const uint32_t sb = sa + (uint32_t)(A_BYTES * (has_lo ? 2 : 1));
#pragma unroll
                    for (int mb = 0; mb < MB; ++mb) {
                        const uint32_t d = tmem + (uint32_t)(buf * MB * W + mb * W);
                        const uint32_t ah = sa + (uint32_t)(mb * 128 * TX_ROWB);
#pragma unroll
                        for (int ks = 0; ks < 4; ++ks) {
                            const uint32_t o = ks * 32;                  // 8 tf32 or 16 bf16 reduction elements = 32 bytes
                            const uint64_t dah = tx_desc_k(ah + o), dbh = tx_desc_k(sb + o), dbl = tx_desc_k(sb + B_BYTES + o);
                            if (MERGE) {
                                const uint32_t acc0 = (first && ks == 0) ? 0u : 1u;
                                if (BF16) { tx_mma_f16(d, dah, dbh, idescW, acc0); continue; }   // A x [b1 | b2 | b3]
                                tx_mma(d, dah, dbh, idescW, acc0);                                // A_hi x [B_hi | B_lo]
                                if (has_lo) tx_mma(d, tx_desc_k(ah + A_BYTES + o), dbh, idesc, 1u);
                                continue;
                            }
                            if (BF16) {
                                // exact bf16 operand x (b1 + b2 + b3): smallest term first
                                tx_mma_f16(d, dah, tx_desc_k(sb + 2 * B_BYTES + o), idesc, (first && ks == 0) ? 0u : 1u);
                                tx_mma_f16(d, dah, dbl, idesc, 1u);
                                tx_mma_f16(d, dah, dbh, idesc, 1u);
                                continue;
                            }
                            if (ks == 0) {
                                const uint32_t acc0 = first ? 0u : 1u;
                                if (has_lo) { tx_mma(d, tx_desc_k(ah + A_BYTES + o), dbh, idesc, acc0); tx_mma(d, dah, dbl, idesc, 1u); }
                                else tx_mma(d, dah, dbl, idesc, acc0);
                            } else {
                                if (has_lo) tx_mma(d, tx_desc_k(ah + A_BYTES + o), dbh, idesc, 1u);
                                tx_mma(d, dah, dbl, idesc, 1u);
                            }
                            tx_mma(d, dah, dbh, idesc, 1u);
                        }
                    }
                    tx_commit(&bars.empty[s]);
                    if (last) tx_commit(&bars.acc_full[buf]);
                }
                __syncwarp();
                if (++s == stages) { s = 0; ph ^= 1u; }
                if (last) { kc = 0; ++c; } else ++kc;
            }
            cc += n_chains;
        }
    } else {
        // ---- drain chains into registers, then the epilogue ----
        const int q = warp & 3, hf = (warp - TX_EPI0) >> 2;          // TMEM lane quadrant, column half
        const int b1 = p.A.box[1], b2 = p.A.box[2], b3 = p.A.box[3];
        int cc = 0;
        constexpr int NCH = (HALF + 31) / 32;                        // 32-column passes of the store phase
        float4 bsum[NCH];                                            // column sums of what this lane stores (bias gradient)
#pragma unroll
        for (int i = 0; i < NCH; ++i) bsum[i] = make_float4(0.f, 0.f, 0.f, 0.f);
        // Everything that does not depend on the tile is computed ONCE: the box coordinates of this thread's
        // accumulator rows (three runtime divisions each) and the uniform feature flags.  The epilogue is the
        // critical path of the small-N layers (~1400 instructions per warp and 256 x 32 tile before this, ncu source
        // view), so the per-tile and per-segment bookkeeping is kept to additions.
        int ri[MB][4];
        bool rin[MB];
#pragma unroll
        for (int mb = 0; mb < MB; ++mb) {
            int t = mb * 128 + 32 * q + lane;
            rin[mb] = t < R;
            ri[mb][0] = t % b1; t /= b1;
            ri[mb][1] = t % b2; t /= b2;
            ri[mb][2] = t % b3; t /= b3;
            ri[mb][3] = t;
        }
        const bool has_mask = p.mask != nullptr, has_bias = p.bias != nullptr, has_cs = p.colsum_part != nullptr;
        const bool has_cb = p.o_cb > 0, has_lo_out = p.out_lo != nullptr;
        constexpr int CW = HALF < 32 ? HALF : 32;                    // columns per store pass
        constexpr int LPR = CW / 4;                                  // lanes per row in the coalesced phase
        constexpr int RPI = 32 / LPR;                                // rows per instruction
        constexpr int NPASS = HALF / CW;
        const int c4 = lane % LPR, lrow = lane / LPR;
        for (int w = blockIdx.x; w < n_work; w += gridDim.x) {
            int base[5];
            tx_tile_base(p.tiles, w / n_n, base);
            const int n0 = (w % n_n) * BN;
            float acc[MB][HALF];
#pragma unroll
            for (int mb = 0; mb < MB; ++mb)
#pragma unroll
                for (int n = 0; n < HALF; ++n) acc[mb][n] = 0.f;
            // this lane's four channels of every store pass: channel, column-block shift, bias
            int chs[NPASS];
            long cbo[NPASS];
            float4 bia[NPASS];
#pragma unroll
            for (int ps = 0; ps < NPASS; ++ps) {
                int ch = n0 + hf * HALF + ps * CW + c4 * 4;
                long co = 0;
                const bool ok = ch < p.N;
                if (ok && has_cb) {
                    const int cb = ch / p.o_cb;
                    co = p.o_cb_off[cb];
                    ch -= cb * p.o_cb;
                }
                chs[ps] = ok ? ch : -1;
                cbo[ps] = co;
                bia[ps] = (ok && has_bias) ? __ldg(reinterpret_cast<const float4*>(p.bias + ch)) : make_float4(0.f, 0.f, 0.f, 0.f);
            }
            // output offsets of this thread's rows (both M blocks), shared with the warp through shared memory; the
            // relu' mask lines of the tile are pulled into L2 now, while the MMAs of the tile are still running
            long* rowoff = reinterpret_cast<long*>(stage_base + 8 * 2048) + (warp - TX_EPI0) * 64;
            __syncwarp();
#pragma unroll
            for (int mb = 0; mb < MB; ++mb) {
                long off0 = -1;
                if (rin[mb]) {
                    const int g1 = base[1] + ri[mb][0], g2 = base[2] + ri[mb][1], g3 = base[3] + ri[mb][2], g4 = base[4] + ri[mb][3];
                    if (g1 < p.o_ext[1] && g2 < p.o_ext[2] && g3 < p.o_ext[3] && g4 < p.o_ext[4])
                        off0 = p.o_base + g1 * p.o_stride[1] + g2 * p.o_stride[2] + g3 * p.o_stride[3] + g4 * p.o_stride[4];
                }
                rowoff[mb * 32 + lane] = off0;
                if (has_mask && off0 >= 0) {
#pragma unroll
                    for (int c0 = 0; c0 < HALF; c0 += 32) {
                        int ch = n0 + hf * HALF + c0;
                        if (ch >= p.N) continue;
                        long off = off0;
                        if (has_cb) {
                            const int cb = ch / p.o_cb;
                            off += p.o_cb_off[cb];
                            ch -= cb * p.o_cb;
                        }
                        asm volatile("prefetch.global.L2 [%0];" ::"l"(p.mask + off + ch));
                    }
                }
            }
            __syncwarp();
            for (int c = 0; c < n_chains; ++c, ++cc) {
                const int buf = cc & 1;
                tx_mbar_wait(&bars.acc_full[buf], (uint32_t)((cc >> 1) & 1));
                asm volatile("tcgen05.fence::after_thread_sync;" ::: "memory");
                // two 16-column loads in flight per wait (the pair is the two M blocks when a thread owns 16 columns)
                // (load l = ((mb * HALF / 16 + c0 / 16) * NPL + column block): the NPL blocks of a product are summed here)
                constexpr int NLD = MB * (HALF / 16) * NPL;
                static_assert(NLD % 2 == 0 || NLD == 1, "tcx_fwd: loads are drained in pairs");
#pragma unroll
                for (int l0 = 0; l0 < NLD; l0 += 2) {
                    uint32_t v[2][16];
#pragma unroll
                    for (int j = 0; j < 2; ++j) {
                        const int l = l0 + j;
                        if (l < NLD) {
                            const int pl = l % NPL, r = l / NPL;
                            const int mb = r / (HALF / 16), c0 = (r % (HALF / 16)) * 16;
                            if (pl > 0 && !MERGE) continue;
                            tx_ld16_issue(tmem + ((uint32_t)(32 * q) << 16) + (uint32_t)(buf * MB * W + mb * W + pl * BN + hf * HALF + c0), v[j]);
                        }
                    }
                    tx_ld_wait16(v[0]);
                    if (l0 + 1 < NLD) tx_ld_wait16(v[1]);              // already complete: only pins the registers
#pragma unroll
                    for (int j = 0; j < 2; ++j) {
                        const int l = l0 + j;
                        if (l < NLD && (l % NPL == 0 || MERGE)) {
                            const int r = l / NPL;
                            const int mb = r / (HALF / 16), c0 = (r % (HALF / 16)) * 16;
#pragma unroll
                            for (int n = 0; n < 16; ++n) acc[mb][c0 + n] += __uint_as_float(v[j][n]);
                        }
                    }
                }
                asm volatile("tcgen05.fence::before_thread_sync;" ::: "memory");
                tx_mbar_arrive(&bars.acc_empty[buf]);
            }
            // Stores: a thread owns a ROW of the accumulator, so writing straight from registers would touch 32
            // different 128-byte lines per instruction.  Each warp instead transposes 16 rows x CW columns at a time
            // through its private 2 KiB staging slab (16-byte chunks xor-swizzled by row: conflict-free both ways) and
            // then writes / reads global memory with LPR lanes per row: whole 128-byte (CW = 32) or 64-byte segments.
            float4* slab = reinterpret_cast<float4*>(stage_base + (warp - TX_EPI0) * 2048);
#pragma unroll
            for (int mb = 0; mb < MB; ++mb) {
#pragma unroll
                for (int ps = 0; ps < NPASS; ++ps) {
                  const int cc0 = ps * CW;
#pragma unroll
                  for (int hr = 0; hr < 2; ++hr) {                       // rows 16 hr .. 16 hr + 15 of this warp's 32
                    constexpr int NIT = 16 / RPI;
                    // addresses and the relu' mask of this lane's NIT segments first: the loads are independent of the
                    // accumulators, so their latency hides behind the slab round trip instead of serialising the stores
                    long offs[NIT];
                    float4 mk[NIT];
#pragma unroll
                    for (int it = 0; it < NIT; ++it) {
                        long off = rowoff[mb * 32 + 16 * hr + it * RPI + lrow];
                        off = (chs[ps] >= 0 && off >= 0) ? off + cbo[ps] + chs[ps] : -1;
                        offs[it] = off;
                        if (has_mask && off >= 0) mk[it] = __ldg(reinterpret_cast<const float4*>(p.mask + off));
                    }
                    __syncwarp();
                    if ((lane >> 4) == hr) {
                        const int lr = lane & 15;
#pragma unroll
                        for (int n = 0; n < CW; n += 4) {
                            float4 v = make_float4(acc[mb][cc0 + n] * p.alpha, acc[mb][cc0 + n + 1] * p.alpha,
                                                   acc[mb][cc0 + n + 2] * p.alpha, acc[mb][cc0 + n + 3] * p.alpha);
                            slab[lr * LPR + ((n >> 2) ^ (lr & (LPR - 1)))] = v;
                        }
                    }
                    __syncwarp();
#pragma unroll
                    for (int it = 0; it < NIT; ++it) {
                        const int lr = it * RPI + lrow;
                        const long off = offs[it];
                        if (off < 0) continue;
                        float4 v = slab[lr * LPR + (c4 ^ (lr & (LPR - 1)))];
                        const float4 b = bia[ps];
                        v.x += b.x; v.y += b.y; v.z += b.z; v.w += b.w;
                        if (p.act == 1) { v.x = fmaxf(v.x, 0.f); v.y = fmaxf(v.y, 0.f); v.z = fmaxf(v.z, 0.f); v.w = fmaxf(v.w, 0.f); }
                        if (has_mask) {
                            const float4 m = mk[it];
                            v.x = m.x > 0.f ? v.x : 0.f; v.y = m.y > 0.f ? v.y : 0.f;
                            v.z = m.z > 0.f ? v.z : 0.f; v.w = m.w > 0.f ? v.w : 0.f;
                        }
                        *reinterpret_cast<float4*>(p.out + off) = v;
                        if (has_lo_out)
                            *reinterpret_cast<float4*>(p.out_lo + off) = make_float4(tx_lo(v.x), tx_lo(v.y), tx_lo(v.z), tx_lo(v.w));
                        if (has_cs) { bsum[cc0 / 32].x += v.x; bsum[cc0 / 32].y += v.y; bsum[cc0 / 32].z += v.z; bsum[cc0 / 32].w += v.w; }
                    }
                  }
                }
            }
        }
        if (p.colsum_part) {
            // lanes with the same lane % LPR hold sums of the same 4 columns: fold them, then park the warp's HALF sums in
            // its own staging slab for the cross-warp fold after the final barrier
            constexpr int CWc = HALF < 32 ? HALF : 32;
            constexpr int LPRc = CWc / 4;
            float* wsum = reinterpret_cast<float*>(stage_base + (warp - TX_EPI0) * 2048);
            __syncwarp();
#pragma unroll
            for (int i = 0; i < NCH; ++i) {
                float4 t = bsum[i];
#pragma unroll
                for (int o = LPRc; o < 32; o <<= 1) {
                    t.x += __shfl_xor_sync(0xffffffffu, t.x, o); t.y += __shfl_xor_sync(0xffffffffu, t.y, o);
                    t.z += __shfl_xor_sync(0xffffffffu, t.z, o); t.w += __shfl_xor_sync(0xffffffffu, t.w, o);
                }
                if (lane < LPRc) *reinterpret_cast<float4*>(wsum + i * 32 + lane * 4) = t;
            }
        }
    }
    asm volatile("tcgen05.fence::before_thread_sync;" ::: "memory");
    __syncthreads();
    if (warp == 1) {
        asm volatile("tcgen05.dealloc.cta_group::1.sync.aligned.b32 %0, %1;" ::"r"(tmem), "r"(TCOLS));
    }
    if (p.colsum_part && threadIdx.x < BN && (int)threadIdx.x < p.N) {
        // column n lives in column half n / HALF: sum its four quadrant warps (fixed order) -> one partial row per CTA
        const int n = threadIdx.x, h = n / HALF, c = n - h * HALF;
        float s = 0.f;
#pragma unroll
        for (int qi = 0; qi < 4; ++qi) s += reinterpret_cast<const float*>(stage_base + (h * 4 + qi) * 2048)[c];
        p.colsum_part[(long)blockIdx.x * p.N + n] = s;
    }
}

// ---------------------------------------------------------------------------------------------------
// weight-gradient-like: G accumulators D_g[128 x BN] += sum over pixel rows of A_g^T B, both operands
// MN-major (rows = reduction).  One pipeline stage = (row tile, group).
// ---------------------------------------------------------------------------------------------------
template <int BN, int G>
__global__ void __launch_bounds__(TX_NT, 1)
tcx_wgrad_kernel(const __grid_constant__ CUtensorMap tmA, const __grid_constant__ CUtensorMap tmAlo,
                 const __grid_constant__ CUtensorMap tmB, const __grid_constant__ CUtensorMap tmBlo,
                 const __grid_constant__ TcxWgrad p, const int stages, const TcxKnobs knobs) {
    constexpr int NBB = BN / 32;
    constexpr int HALF = BN / 2;
    extern __shared__ uint8_t tx_dyn[];
    __shared__ TxBars bars;
    const uint32_t dyn = (tx_smem(tx_dyn) + 1023u) & ~1023u;
    uint8_t* dyn_ptr = tx_dyn + (dyn - tx_smem(tx_dyn));
    const int warp = threadIdx.x >> 5, lane = threadIdx.x & 31;
    const bool has_lo = p.A.lo != nullptr;
    const int R = p.A.box[1] * p.A.box[2] * p.A.box[3] * p.A.box[4];
    const int R8 = (R + 7) & ~7;
    const int box_bytes = R8 * TX_ROWB;
    const int a_planes = has_lo ? 2 : 1;
    // shared memory: [2 slots of the shared operand Bm: 2 NBB boxes each][`stages` stages of A: 4 a_planes boxes each].
    // Bm (the output gradient) is the same for every accumulator group of a row tile, so it is loaded once per row
    // tile into its own 2-deep ring instead of once per (row tile, group).
    const int slotB_bytes = box_bytes * 2 * NBB;
    const int stage_bytes = box_bytes * 4 * a_planes;
    const uint32_t dynA = dyn + 2u * (uint32_t)slotB_bytes;
    const int per = (p.rows.n_tiles + p.splits - 1) / p.splits;
    const int rt0 = blockIdx.x * per;
    const int rt1 = min(p.rows.n_tiles, rt0 + per);
    const int n_rt = max(0, rt1 - rt0);
    const int chain = p.chain > 0 ? p.chain : max(1, n_rt);
    const int n_chains = (n_rt + chain - 1) / chain;
    const int ota = blockIdx.y % p.ot_a, otb = blockIdx.y / p.ot_a;

    if (threadIdx.x == 0) {
        for (int s = 0; s < stages; ++s) { tx_mbar_init(&bars.full[s], 1); tx_mbar_init(&bars.empty[s], 1); }
        for (int b = 0; b < 2; ++b) {
            tx_mbar_init(&bars.acc_full[b], 1); tx_mbar_init(&bars.acc_empty[b], 256);
            tx_mbar_init(&bars.fullB[b], 1); tx_mbar_init(&bars.emptyB[b], 1);
        }
        asm volatile("fence.mbarrier_init.release.cluster;" ::: "memory");
    }
    if (warp == 1) {
        constexpr uint32_t cols = (2 * G * BN <= 32) ? 32 : (2 * G * BN <= 64) ? 64 : (2 * G * BN <= 128) ? 128 : (2 * G * BN <= 256) ? 256 : 512;
        asm volatile("tcgen05.alloc.cta_group::1.sync.aligned.shared::cta.b32 [%0], %1;" ::"r"(tx_smem(&bars.tmem_base)), "r"(cols));
        asm volatile("tcgen05.relinquish_alloc_permit.cta_group::1.sync.aligned;" ::);
    }
    // rows R..R8-1 of every box are never written by TMA but take part in the K = 8 MMA steps: zero them once
    if (R8 > R) {
        const int nbox = 2 * 2 * NBB + stages * 4 * a_planes;
        const int per_box = (R8 - R) * (TX_ROWB / 4);
        for (int e = threadIdx.x; e < nbox * per_box; e += TX_NT) {
            const int b = e / per_box, w = e - b * per_box;
            reinterpret_cast<float*>(dyn_ptr + (size_t)b * box_bytes + (size_t)R * TX_ROWB)[w] = 0.f;
        }
        asm volatile("fence.proxy.async.shared::cta;" ::: "memory");
    }
    asm volatile("tcgen05.fence::before_thread_sync;" ::: "memory");
    __syncthreads();
    asm volatile("tcgen05.fence::after_thread_sync;" ::: "memory");
    const uint32_t tmem = bars.tmem_base;

    if (warp == 0) {
        // ===== TMA producer (whole warp walks the loop, one elected lane issues) =====
        const uint32_t bytesA = (uint32_t)(R * TX_ROWB * 4 * a_planes), bytesB = (uint32_t)(R * TX_ROWB * 2 * NBB);
        int qn = 0, s = 0;
        uint32_t ph = 0;
        for (int i = 0; i < n_rt; ++i) {
            int base[5], ba[5];
            tx_tile_base(p.rows, rt0 + i, base);
#pragma unroll
            for (int d = 0; d < 5; ++d) ba[d] = base[d];
            if (p.gather_dim > 0) ba[p.gather_dim] = (int)__ldg(p.gather_idx + base[p.gather_dim]);
            const int sl = i & 1;
            if (i >= 2) tx_mbar_wait(&bars.emptyB[sl], (uint32_t)(((i >> 1) - 1) & 1));
            if (tx_elect()) {
                const uint32_t sb = dyn + (uint32_t)(sl * slotB_bytes);
                tx_expect(&bars.fullB[sl], bytesB);
#pragma unroll
                for (int j = 0; j < NBB; ++j) {
                    const int c0 = p.gb[j][0] + otb * p.ot_b_step, c1 = p.gb[j][1] + base[1], c2 = p.gb[j][2] + base[2],
                              c3 = p.gb[j][3] + base[3], c4 = base[4];
                    tx_tma5(sb + (uint32_t)(j * box_bytes), &tmB, &bars.fullB[sl], c0, c1, c2, c3, c4);
                    tx_tma5(sb + (uint32_t)((NBB + j) * box_bytes), &tmBlo, &bars.fullB[sl], c0, c1, c2, c3, c4);
                }
            }
            __syncwarp();
            for (int g = 0; g < G; ++g, ++qn) {
                if (qn >= stages) tx_mbar_wait(&bars.empty[s], ph ^ 1u);
                if (tx_elect()) {
                    const uint32_t sa = dynA + (uint32_t)(s * stage_bytes);
                    tx_expect(&bars.full[s], bytesA);
#pragma unroll
                    for (int j = 0; j < 4; ++j) {
                        const int c0 = p.ga[g][j][0] + ota * p.ot_a_step, c1 = p.ga[g][j][1] + ba[1],
                                  c2 = p.ga[g][j][2] + ba[2], c3 = p.ga[g][j][3] + ba[3], c4 = ba[4];
                        tx_tma5(sa + (uint32_t)(j * box_bytes), &tmA, &bars.full[s], c0, c1, c2, c3, c4);
                        if (has_lo) tx_tma5(sa + (uint32_t)((4 + j) * box_bytes), &tmAlo, &bars.full[s], c0, c1, c2, c3, c4);
                    }
                }
                __syncwarp();
                if (++s == stages) { s = 0; ph ^= 1u; }
            }
        }
    } else if (warp == 1) {
        // ===== MMA issuer (whole warp waits, one elected lane issues) =====
        // D fp32, A/B tf32, both MN-major, N = BN, M = 128
        const uint32_t idesc = (1u << 4) | (2u << 7) | (2u << 10) | (1u << 15) | (1u << 16) |
                               ((uint32_t)(BN >> 3) << 17) | ((uint32_t)(128 >> 4) << 24);
        // MN-major descriptors differ only in the start address: LBO = distance between 32-channel boxes
        const uint64_t dconst = tx_desc(0, (uint32_t)box_bytes, (uint32_t)knobs.mn_sbo_bytes, (uint32_t)knobs.mn_layout_type);
        const uint32_t kstep16 = (uint32_t)knobs.mn_kstep_bytes >> 4;
        const int ksteps = R8 / 8;
        int s = 0, kc = 0, c = 0;
        uint32_t ph = 0;
        for (int i = 0; i < n_rt; ++i) {
            const int buf = c & 1, sl = i & 1;
            const bool first = kc == 0;
            const bool last = (i == n_rt - 1) || (kc + 1 == chain);
            if (first && c >= 2) tx_mbar_wait(&bars.acc_empty[buf], (uint32_t)(((c >> 1) - 1) & 1));
            tx_mbar_wait(&bars.fullB[sl], (uint32_t)((i >> 1) & 1));
            for (int g = 0; g < G; ++g) {
                tx_mbar_wait(&bars.full[s], ph);
                asm volatile("tcgen05.fence::after_thread_sync;" ::: "memory");
                if (tx_elect()) {
                    const uint32_t sa = dynA + (uint32_t)(s * stage_bytes);
                    uint64_t dah = dconst | (uint64_t)((sa & 0x3FFFF) >> 4);
                    uint64_t dal = dconst | (uint64_t)(((sa + (uint32_t)(4 * box_bytes)) & 0x3FFFF) >> 4);
                    const uint32_t sb = dyn + (uint32_t)(sl * slotB_bytes);
                    uint64_t dbh = dconst | (uint64_t)((sb & 0x3FFFF) >> 4);
                    uint64_t dbl = dconst | (uint64_t)(((sb + (uint32_t)(NBB * box_bytes)) & 0x3FFFF) >> 4);
                    const uint32_t d = tmem + (uint32_t)(buf * G * BN + g * BN);
                    if (has_lo) {
                        tx_mma(d, dal, dbh, idesc, first ? 0u : 1u);
                        tx_mma(d, dah, dbl, idesc, 1u);
                    } else {
                        tx_mma(d, dah, dbl, idesc, first ? 0u : 1u);
                    }
                    tx_mma(d, dah, dbh, idesc, 1u);
                    for (int ks = 1; ks < ksteps; ++ks) {
                        dah += kstep16; dal += kstep16; dbh += kstep16; dbl += kstep16;
                        if (has_lo) tx_mma(d, dal, dbh, idesc, 1u);
                        tx_mma(d, dah, dbl, idesc, 1u);
                        tx_mma(d, dah, dbh, idesc, 1u);
                    }
                    tx_commit(&bars.empty[s]);
                    if (g == G - 1) {
                        tx_commit(&bars.emptyB[sl]);
                        if (last) tx_commit(&bars.acc_full[buf]);
                    }
                }
                __syncwarp();
                if (++s == stages) { s = 0; ph ^= 1u; }
            }
            if (last) { kc = 0; ++c; } else ++kc;
        }
    } else {
        const int q = warp & 3, hf = (warp - TX_EPI0) >> 2;
        float acc[G][HALF];
#pragma unroll
        for (int g = 0; g < G; ++g)
#pragma unroll
            for (int n = 0; n < HALF; ++n) acc[g][n] = 0.f;
        for (int c = 0; c < n_chains; ++c) {
            const int buf = c & 1;
            tx_mbar_wait(&bars.acc_full[buf], (uint32_t)((c >> 1) & 1));
            asm volatile("tcgen05.fence::after_thread_sync;" ::: "memory");
#pragma unroll
            for (int g = 0; g < G; ++g)
#pragma unroll
                for (int c0 = 0; c0 < HALF; c0 += 16) {
                    uint32_t v[16];
                    tx_ld16(tmem + ((uint32_t)(32 * q) << 16) + (uint32_t)(buf * G * BN + g * BN + hf * HALF + c0), v);
#pragma unroll
                    for (int n = 0; n < 16; ++n) acc[g][c0 + n] += __uint_as_float(v[n]);
                }
            asm volatile("tcgen05.fence::before_thread_sync;" ::: "memory");
            tx_mbar_arrive(&bars.acc_empty[buf]);
        }
        float* dst = p.partial + (long)blockIdx.x * p.split_stride;
#pragma unroll
        for (int g = 0; g < G; ++g) {
            const int m = g * 128 + 32 * q + lane;
            const int off = __ldg(p.row_off + (long)ota * (G * 128) + m);
            if (off < 0) continue;
#pragma unroll
            for (int n = 0; n < HALF; ++n) {
                const int col = otb * BN + hf * HALF + n;
                if (col < p.n_cols) dst[off + (long)col * p.col_stride] = acc[g][n] * p.alpha;
            }
        }
    }
    asm volatile("tcgen05.fence::before_thread_sync;" ::: "memory");
    __syncthreads();
    if (warp == 1) {
        constexpr uint32_t cols = (2 * G * BN <= 32) ? 32 : (2 * G * BN <= 64) ? 64 : (2 * G * BN <= 128) ? 128 : (2 * G * BN <= 256) ? 256 : 512;
        asm volatile("tcgen05.dealloc.cta_group::1.sync.aligned.b32 %0, %1;" ::"r"(tmem), "r"(cols));
    }
}

// ---------------------------------------------------------------------------------------------------
// host side
// ---------------------------------------------------------------------------------------------------
typedef CUresult (*EncodeTiledFn)(CUtensorMap*, CUtensorMapDataType, cuuint32_t, void*, const cuuint64_t*, const cuuint64_t*,
                                  const cuuint32_t*, const cuuint32_t*, CUtensorMapInterleave, CUtensorMapSwizzle,
                                  CUtensorMapL2promotion, CUtensorMapFloatOOBfill);

EncodeTiledFn encode_fn() {
    static EncodeTiledFn fn = nullptr;
    if (!fn) {
        void* p = nullptr;
        cudaDriverEntryPointQueryResult qres;
        if (cudaGetDriverEntryPoint("cuTensorMapEncodeTiled", &p, cudaEnableDefault, &qres) == cudaSuccess &&
            qres == cudaDriverEntryPointSuccess)
            fn = reinterpret_cast<EncodeTiledFn>(p);
    }
    return fn;
}

int make_map(CUtensorMap* out, const float* base, const int64_t* dim, const int64_t* stride, const int32_t* box, int swizzle,
             const char* what, int esize = 4) {
    EncodeTiledFn fn = encode_fn();
    if (!fn) { set_error("tcx: cuTensorMapEncodeTiled is not available"); return A2CB_ERR_CUDA; }
    cuuint64_t gd[5], gs[4];
    cuuint32_t bx[5], es[5] = {1, 1, 1, 1, 1};
    for (int d = 0; d < 5; ++d) { gd[d] = (cuuint64_t)dim[d]; bx[d] = (cuuint32_t)box[d]; }
    for (int d = 1; d < 5; ++d) gs[d - 1] = (cuuint64_t)stride[d] * esize;
    A2CB_REQUIRE(stride[0] == 1 && box[0] * esize == 128, "tcx(%s): innermost dimension must be contiguous with a 128-byte box", what);
    A2CB_REQUIRE(((uintptr_t)base & 15) == 0, "tcx(%s): base pointer must be 16-byte aligned", what);
    for (int d = 1; d < 5; ++d)
        A2CB_REQUIRE((stride[d] * esize) % 16 == 0 && box[d] >= 1 && box[d] <= 256 && dim[d] >= 1, "tcx(%s): bad dimension %d", what, d);
    CUresult r = fn(out, esize == 2 ? CU_TENSOR_MAP_DATA_TYPE_BFLOAT16 : CU_TENSOR_MAP_DATA_TYPE_FLOAT32, 5, const_cast<float*>(base), gd, gs, bx, es,
                    CU_TENSOR_MAP_INTERLEAVE_NONE, (CUtensorMapSwizzle)swizzle, CU_TENSOR_MAP_L2_PROMOTION_L2_256B,
                    CU_TENSOR_MAP_FLOAT_OOB_FILL_NONE);
    if (r != CUDA_SUCCESS) { set_error("tcx(%s): cuTensorMapEncodeTiled failed (%d)", what, (int)r); return A2CB_ERR_CUDA; }
    return A2CB_OK;
}

int view_maps(CUtensorMap* hi, CUtensorMap* lo, const TcxView& v, int swizzle, const char* what) {
    int rc = make_map(hi, v.x, v.dim, v.stride, v.box, swizzle, what);
    if (rc) return rc;
    if (v.lo) return make_map(lo, v.lo, v.dim, v.stride, v.box, swizzle, what);
    *lo = *hi;
    return A2CB_OK;
}

constexpr int kSmemBudget = 200 * 1024;
constexpr int kSmemMax = 232448 - 1024;          // 227 KiB opt-in limit minus the kernels' static shared memory

template <int BN, int MB, bool BF16 = false, bool PATCH = false>
int launch_fwd(const TcxFwd& p, const CUtensorMap* m, cudaStream_t st) {
    const int a_planes = (!BF16 && p.A.lo) ? 2 : 1;
    const int R = p.A.box[1] * p.A.box[2] * p.A.box[3] * p.A.box[4];
    const int R8 = (R + 7) & ~7;
    // patch mode: a stage is one patch (all planes).  bf16: the weight image (all k-blocks, three planes) is resident
    // next to the ring; TF32: the weight k-blocks stream through a ring of their own behind the patch ring.
    const int slotB = (BF16 ? 3 : 2) * BN * TX_ROWB;
    const int stagesB = (PATCH && !BF16) ? 4 : 0;
    const int resident = PATCH ? (BF16 ? p.n_kb * slotB : stagesB * slotB) : 0;
    const int stage = PATCH ? R8 * TX_ROWB * a_planes : R8 * TX_ROWB * a_planes + slotB;
    // the last stage's M-block reads may run past its own end: keep that many rows of slack inside the allocation
    int max_shift = 0;
    if (PATCH)
        for (int kb = 0; kb < p.n_kb; ++kb) {
            const int sh = p.kb_a[kb][1] + p.kb_a[kb][2] * p.A.box[1] + p.kb_a[kb][3] * p.A.box[1] * p.A.box[2];
            A2CB_REQUIRE(sh >= 0, "tcx_fwd: patch mode takes non-negative tap shifts");
            max_shift = sh > max_shift ? sh : max_shift;
        }
    int slack = (MB * 128 + max_shift - R8) * TX_ROWB;
    slack = slack > 0 ? slack : 0;
    if (PATCH && !BF16) slack = slack > stagesB * slotB ? slack - stagesB * slotB : 0;   // the weight ring follows the patches
    int stages = (kSmemMax - 1024 - TX_EPI_STAGING - slack - resident) / stage;
    stages = stages > TX_MAX_STAGES ? TX_MAX_STAGES : stages;
    A2CB_REQUIRE(stages >= 2, "tcx_fwd: tile does not fit shared memory");
    const int smem = stages * stage + 1024 + TX_EPI_STAGING + slack + resident;
    static bool attr = false;
    if (!attr) {
        cudaError_t e = cudaFuncSetAttribute(tcx_fwd_kernel<BN, MB, BF16, PATCH>, cudaFuncAttributeMaxDynamicSharedMemorySize, kSmemMax);
        if (e != cudaSuccess) { set_error("tcx_fwd: cudaFuncSetAttribute: %s", cudaGetErrorString(e)); return A2CB_ERR_CUDA; }
        attr = true;
    }
    // Measured on B200: the 128-byte swizzle of a UMMA operand is a function of the ABSOLUTE shared-memory address
    // (bits 7..9 xor-ed into bits 4..6), exactly like TMA's, so a descriptor whose start is advanced by whole rows reads
    // a row-shifted window of the patch as is; filling the descriptor's base-offset field with the start's row phase
    // (A2CB_TCX_PATCH_BO=1) applies the phase twice and gives wrong products.
    static int base_offset = -1;
    if (base_offset < 0) { const char* e = getenv("A2CB_TCX_PATCH_BO"); base_offset = (e && e[0] == '1') ? 1 : 0; }
    const long n_work = (long)p.tiles.n_tiles * ((p.N + BN - 1) / BN);
    dim3 grid((unsigned)(n_work < kNumSM ? n_work : kNumSM));
    static int merge_on = -1;
    if (merge_on < 0) { const char* e = getenv("A2CB_TCX_MERGE"); merge_on = (e && e[0] >= '0' && e[0] <= '7') ? e[0] - '0' : 7; }
    tcx_fwd_kernel<BN, MB, BF16, PATCH><<<grid, TX_NT, smem, st>>>(m[0], m[1], m[2], m[3], m[4], p, stages, base_offset, stagesB, merge_on);
    return check_launch("tcx_fwd");
}

template <int BN, int G>
int launch_wgrad(const TcxWgrad& p, const CUtensorMap* m, cudaStream_t st) {
    const int R = p.A.box[1] * p.A.box[2] * p.A.box[3] * p.A.box[4];
    const int R8 = (R + 7) & ~7;
    const int stage = R8 * TX_ROWB * 4 * (p.A.lo ? 2 : 1);          // A boxes of one (row tile, group)
    const int slots = 2 * R8 * TX_ROWB * 2 * (BN / 32);              // two slots of the shared operand
    int stages = (kSmemBudget - slots) / stage;
    stages = stages > TX_MAX_STAGES ? TX_MAX_STAGES : stages;
    A2CB_REQUIRE(stages >= 2, "tcx_wgrad: row tile of %d rows does not fit shared memory", R);
    const int smem = stages * stage + slots + 1024;
    static bool attr = false;
    if (!attr) {
        cudaError_t e = cudaFuncSetAttribute(tcx_wgrad_kernel<BN, G>, cudaFuncAttributeMaxDynamicSharedMemorySize, kSmemBudget + 1024);
        if (e != cudaSuccess) { set_error("tcx_wgrad: cudaFuncSetAttribute: %s", cudaGetErrorString(e)); return A2CB_ERR_CUDA; }
        attr = true;
    }
    dim3 grid((unsigned)p.splits, (unsigned)(p.ot_a * p.ot_b));
    tcx_wgrad_kernel<BN, G><<<grid, TX_NT, smem, st>>>(m[0], m[1], m[2], m[3], p, stages, g_knobs);
    return check_launch("tcx_wgrad");
}

}  // namespace

void tcx_set_knobs(const TcxKnobs& k) { g_knobs = k; }
TcxKnobs tcx_get_knobs() { return g_knobs; }

int tcx_fwd(const TcxFwd& p, cudaStream_t st) {
    A2CB_REQUIRE(p.A.x && p.B && p.B_lo && p.out, "tcx_fwd: null pointer");
    A2CB_REQUIRE(p.n_kb >= 1 && p.n_kb <= TCX_MAX_KB, "tcx_fwd: n_kb = %d out of range", p.n_kb);
    A2CB_REQUIRE(p.N > 0 && p.N % 4 == 0, "tcx_fwd: N = %d must be a positive multiple of 4", p.N);
    A2CB_REQUIRE(p.gather_dim == 0 || (p.gather_dim >= 1 && p.gather_dim <= 4 && p.gather_idx), "tcx_fwd: bad gather");
    A2CB_REQUIRE(!p.colsum_part || p.N <= 128, "tcx_fwd: column sums need a single column tile (N <= 128)");
    A2CB_REQUIRE(p.o_cb == 0 || (p.o_cb % 4 == 0 && !p.bias && (p.N + p.o_cb - 1) / p.o_cb <= 4), "tcx_fwd: bad column blocks");
    A2CB_REQUIRE(p.tiles.n_tiles > 0 && p.tiles.t_div > 0 && p.tiles.tr_dim >= 1 && p.tiles.tr_dim <= 4 &&
                 p.tiles.tq_dim >= 1 && p.tiles.tq_dim <= 4, "tcx_fwd: bad tiling");
    const long R = (long)p.A.box[1] * p.A.box[2] * p.A.box[3] * p.A.box[4];
    A2CB_REQUIRE(R >= 1 && (R <= 256 || (p.patch > 0 && R <= 512)), "tcx_fwd: %ld rows per tile (max 256)", R);
    const int BN = p.N <= 32 ? 32 : (p.N <= 64 ? 64 : 128);
    const int MB = R <= 128 ? 1 : 2;
    CUtensorMap m[5];
    if (p.bf16) {
        // exact bf16 operand, three-plane bf16 weight image, 64-channel k-blocks
        const bool patch = p.patch > 0;
        A2CB_REQUIRE(p.B3 && !p.A.lo && BN == 32 && (patch || MB == 2), "tcx_fwd: bf16 mode is built for 32 output channels and 129..256 rows per tile");
        A2CB_REQUIRE(!patch || (p.patch > 128 && p.patch <= 256 && p.N <= 32 && p.n_kb <= 8), "tcx_fwd: patch mode takes 129..256 accumulator rows, one column tile, <= 8 taps");
        int rc = make_map(&m[0], p.A.x, p.A.dim, p.A.stride, p.A.box, (int)CU_TENSOR_MAP_SWIZZLE_128B, "fwd A (bf16)", 2);
        if (rc) return rc;
        m[1] = m[0];
        const int64_t bd[5] = {p.b_cols, p.b_rows, 1, 1, 1};
        const int64_t bs[5] = {1, p.b_cols, p.b_cols * p.b_rows, p.b_cols * p.b_rows, p.b_cols * p.b_rows};
        const int32_t bb[5] = {64, BN, 1, 1, 1};
        A2CB_REQUIRE(p.b_cols % 8 == 0, "tcx_fwd: weight image pitch must be a multiple of 8 (bf16)");
        if ((rc = make_map(&m[2], p.B, bd, bs, bb, (int)CU_TENSOR_MAP_SWIZZLE_128B, "fwd B1", 2))) return rc;
        if ((rc = make_map(&m[3], p.B_lo, bd, bs, bb, (int)CU_TENSOR_MAP_SWIZZLE_128B, "fwd B2", 2))) return rc;
        if ((rc = make_map(&m[4], p.B3, bd, bs, bb, (int)CU_TENSOR_MAP_SWIZZLE_128B, "fwd B3", 2))) return rc;
        return patch ? launch_fwd<32, 2, true, true>(p, m, st) : launch_fwd<32, 2, true>(p, m, st);
    }
    int rc = view_maps(&m[0], &m[1], p.A, (int)CU_TENSOR_MAP_SWIZZLE_128B, "fwd A");
    if (rc) return rc;
    const int64_t bd[5] = {p.b_cols, p.b_rows, 1, 1, 1};
    const int64_t bs[5] = {1, p.b_cols, p.b_cols * p.b_rows, p.b_cols * p.b_rows, p.b_cols * p.b_rows};
    const int32_t bb[5] = {32, BN, 1, 1, 1};
    A2CB_REQUIRE(p.b_cols % 4 == 0, "tcx_fwd: weight image pitch must be a multiple of 4");
    if ((rc = make_map(&m[2], p.B, bd, bs, bb, (int)CU_TENSOR_MAP_SWIZZLE_128B, "fwd B"))) return rc;
    if ((rc = make_map(&m[3], p.B_lo, bd, bs, bb, (int)CU_TENSOR_MAP_SWIZZLE_128B, "fwd B_lo"))) return rc;
    m[4] = m[3];
    if (p.patch > 0) {
        A2CB_REQUIRE(MB == 2 && p.patch > 128 && p.patch <= 256 && p.N <= BN && BN <= 64,
                     "tcx_fwd: patch mode takes 129..256 accumulator rows and one column tile of <= 64 channels");
        return BN == 32 ? launch_fwd<32, 2, false, true>(p, m, st) : launch_fwd<64, 2, false, true>(p, m, st);
    }
    if (BN == 32) return MB == 1 ? launch_fwd<32, 1>(p, m, st) : launch_fwd<32, 2>(p, m, st);
    if (BN == 64) return MB == 1 ? launch_fwd<64, 1>(p, m, st) : launch_fwd<64, 2>(p, m, st);
    return MB == 1 ? launch_fwd<128, 1>(p, m, st) : launch_fwd<128, 2>(p, m, st);
}

int tcx_wgrad(const TcxWgrad& p, cudaStream_t st) {
    A2CB_REQUIRE(p.A.x && p.Bm.x && p.Bm.lo && p.partial && p.row_off, "tcx_wgrad: null pointer");
    A2CB_REQUIRE(p.G >= 1 && p.G <= TCX_MAX_G && p.nbb >= 1 && p.nbb <= 8, "tcx_wgrad: bad G / nbb");
    A2CB_REQUIRE(p.splits >= 1 && p.ot_a >= 1 && p.ot_b >= 1 && p.rows.n_tiles >= 1 && p.rows.t_div > 0, "tcx_wgrad: bad tiling");
    for (int d = 1; d < 5; ++d) A2CB_REQUIRE(p.A.box[d] == p.Bm.box[d] || true, "tcx_wgrad: box mismatch");
    const long RA = (long)p.A.box[1] * p.A.box[2] * p.A.box[3] * p.A.box[4];
    const long RB = (long)p.Bm.box[1] * p.Bm.box[2] * p.Bm.box[3] * p.Bm.box[4];
    A2CB_REQUIRE(RA == RB && RA >= 1, "tcx_wgrad: operand boxes enumerate %ld vs %ld rows", RA, RB);
    A2CB_REQUIRE(p.gather_dim == 0 || (p.gather_dim >= 1 && p.gather_dim <= 4 && p.gather_idx), "tcx_wgrad: bad gather");
    CUtensorMap m[4];
    int rc = view_maps(&m[0], &m[1], p.A, g_knobs.mn_tma_swizzle, "wgrad A");
    if (rc) return rc;
    if ((rc = view_maps(&m[2], &m[3], p.Bm, g_knobs.mn_tma_swizzle, "wgrad B"))) return rc;
    const int BN = 32 * p.nbb;
    if (BN == 32 && p.G == 1) return launch_wgrad<32, 1>(p, m, st);
    if (BN == 32 && p.G == 2) return launch_wgrad<32, 2>(p, m, st);
    if (BN == 32 && p.G == 5) return launch_wgrad<32, 5>(p, m, st);
    if (BN == 64 && p.G == 1) return launch_wgrad<64, 1>(p, m, st);
    if (BN == 64 && p.G == 4) return launch_wgrad<64, 4>(p, m, st);
    if (BN == 128 && p.G == 1) return launch_wgrad<128, 1>(p, m, st);
    if (BN == 256 && p.G == 1) return launch_wgrad<256, 1>(p, m, st);
    set_error("tcx_wgrad: no kernel instance for BN = %d, G = %d", BN, p.G);
    return A2CB_ERR_UNSUPPORTED;
}

}  // namespace a2cb
