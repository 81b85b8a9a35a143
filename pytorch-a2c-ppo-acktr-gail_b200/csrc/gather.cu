// Minibatch shuffle-gather of rollout rows (bit-exact copies).
//
// Replaces the 8 `aten::index` gathers per minibatch of
// RolloutStorage.feed_forward_generator (a2c_ppo_acktr/storage.py:126-140) and the
// per-env stack/flatten of recurrent_generator (storage.py:154-199).
//
// Row maps (flat row index of the [T*N, ...] view of a [T(+1), N, ...] tensor):
//   mode 0 (feed-forward): out row j <- src row idx[j]              (idx = perm chunk)
//   mode 1 (recurrent):    out row j <- src row (j / E) * N + idx[j % E]   (idx = env chunk,
//                          rows ordered time-major t*E + e, storage.py:176-199)
//
// Wide rows (observations, hidden states: >= 1 KiB, 16 B multiple) move as TMA 1-D
// bulk copies global -> shared -> global driven by ONE thread per CTA through an
// mbarrier ring, so the SM issues no per-byte instructions and the copy runs at HBM
// speed.  Narrow fields (actions, value_preds, returns, masks, old log-probs,
// advantages: 4-8 B per row) are batched into a single SIMT launch.
#include "common.cuh"
#include "decls.cuh"

namespace a2cb {

struct RowMap {
    const int64_t* idx;
    int mode;
    long E;
    long N;
    __device__ __forceinline__ long src_row(long j) const {
        if (mode == 0) return (long)idx[j];
        const long t = j / E;
        return t * N + (long)idx[j - t * E];
    }
};

// ---------------------------------------------------------------------------
// TMA bulk-copy gather
// ---------------------------------------------------------------------------
constexpr int kStages = 4;
constexpr int kLookahead = 2;       // loads in flight; kStages - kLookahead stores in flight
constexpr int kChunk = 16384;       // bytes per stage

__device__ __forceinline__ uint32_t smem_u32(const void* p) {
    return (uint32_t)__cvta_generic_to_shared(p);
}
__device__ __forceinline__ void mbar_init(uint64_t* bar, uint32_t count) {
    asm volatile("mbarrier.init.shared::cta.b64 [%0], %1;" ::"r"(smem_u32(bar)), "r"(count));
}
__device__ __forceinline__ void mbar_expect_tx(uint64_t* bar, uint32_t bytes) {
    asm volatile("mbarrier.arrive.expect_tx.shared::cta.b64 _, [%0], %1;" ::"r"(smem_u32(bar)), "r"(bytes)
                 : "memory");
}
__device__ __forceinline__ void mbar_wait(uint64_t* bar, uint32_t parity) {
    uint32_t ok = 0;
    while (!ok) {
        asm volatile(
            "{\n"
            ".reg .pred p;\n"
            "mbarrier.try_wait.parity.shared::cta.b64 p, [%1], %2;\n"
            "selp.u32 %0, 1, 0, p;\n"
            "}\n"
            : "=r"(ok)
            : "r"(smem_u32(bar)), "r"(parity)
            : "memory");
    }
}
__device__ __forceinline__ void bulk_g2s(void* smem_dst, const void* gsrc, uint32_t bytes, uint64_t* bar) {
    asm volatile(
        "cp.async.bulk.shared::cluster.global.mbarrier::complete_tx::bytes [%0], [%1], %2, [%3];" ::"r"(
            smem_u32(smem_dst)),
        "l"(gsrc), "r"(bytes), "r"(smem_u32(bar))
        : "memory");
}
__device__ __forceinline__ void bulk_s2g(void* gdst, const void* smem_src, uint32_t bytes) {
    asm volatile("cp.async.bulk.global.shared::cta.bulk_group [%0], [%1], %2;" ::"l"(gdst),
                 "r"(smem_u32(smem_src)), "r"(bytes)
                 : "memory");
}

__global__ void __launch_bounds__(32)
gather_rows_tma_kernel(const uint8_t* __restrict__ src, uint8_t* __restrict__ dst, long row_bytes,
                       long n_rows, RowMap map) {
    extern __shared__ __align__(128) uint8_t smem[];
    __shared__ __align__(8) uint64_t bars[kStages];
    if (threadIdx.x != 0) return;

    for (int s = 0; s < kStages; ++s) mbar_init(&bars[s], 1);
    asm volatile("fence.mbarrier_init.release.cluster;" ::: "memory");
    asm volatile("fence.proxy.async.shared::cta;" ::: "memory");

    const long chunks_per_row = (row_bytes + kChunk - 1) / kChunk;
    const long total = n_rows * chunks_per_row;
    // items owned by this CTA: w_k = blockIdx.x + k * gridDim.x
    const long n_items = (total > (long)blockIdx.x) ? (total - blockIdx.x + gridDim.x - 1) / gridDim.x : 0;

    for (long i = 0; i < n_items + kLookahead; ++i) {
        if (i < n_items) {
            const int stage = (int)(i % kStages);
            // the store that last read this stage (item i - kStages) must be done reading
            asm volatile("cp.async.bulk.wait_group.read %0;" ::"n"(kStages - kLookahead - 1) : "memory");
            const long w = blockIdx.x + i * gridDim.x;
            const long j = w / chunks_per_row, c = w - j * chunks_per_row;
            const long off = c * kChunk;
            const uint32_t bytes = (uint32_t)min((long)kChunk, row_bytes - off);
            mbar_expect_tx(&bars[stage], bytes);
            bulk_g2s(smem + stage * kChunk, src + map.src_row(j) * row_bytes + off, bytes, &bars[stage]);
        }
        const long k = i - kLookahead;
        if (k >= 0) {
            const int stage = (int)(k % kStages);
            mbar_wait(&bars[stage], (uint32_t)((k / kStages) & 1));
            const long w = blockIdx.x + k * gridDim.x;
            const long j = w / chunks_per_row, c = w - j * chunks_per_row;
            const long off = c * kChunk;
            const uint32_t bytes = (uint32_t)min((long)kChunk, row_bytes - off);
            asm volatile("fence.proxy.async.shared::cta;" ::: "memory");
            bulk_s2g(dst + j * row_bytes + off, smem + stage * kChunk, bytes);
            asm volatile("cp.async.bulk.commit_group;" ::: "memory");
        }
    }
    asm volatile("cp.async.bulk.wait_group 0;" ::: "memory");
}

// ---------------------------------------------------------------------------
// narrow fields: one launch for all of them
// ---------------------------------------------------------------------------
constexpr int kMaxFields = 12;
struct FieldPack {
    const uint8_t* src[kMaxFields];
    uint8_t* dst[kMaxFields];
    int row_bytes[kMaxFields];
    int n;
};

template <typename W>
__device__ __forceinline__ void copy_words(const uint8_t* s, uint8_t* d, int bytes, int lane, int nl) {
    const int nw = bytes / (int)sizeof(W);
    for (int w = lane; w < nw; w += nl)
        reinterpret_cast<W*>(d)[w] = reinterpret_cast<const W*>(s)[w];
}

// `G` threads cooperate on one output row (G = 1 for scalar fields, 32 for wider rows).
template <int G>
__global__ void __launch_bounds__(256)
gather_fields_kernel(FieldPack fp, long n_rows, RowMap map) {
    const long gid = ((long)blockIdx.x * blockDim.x + threadIdx.x) / G;
    const int lane = threadIdx.x % G;
    if (gid >= n_rows) return;
    const long sr = map.src_row(gid);
    for (int f = 0; f < fp.n; ++f) {
        const int rb = fp.row_bytes[f];
        const uint8_t* s = fp.src[f] + sr * rb;
        uint8_t* d = fp.dst[f] + gid * rb;
        const uintptr_t al = (uintptr_t)s | (uintptr_t)d | (uintptr_t)rb;
        if ((al & 15) == 0) copy_words<uint4>(s, d, rb, lane, G);
        else if ((al & 7) == 0) copy_words<uint2>(s, d, rb, lane, G);
        else if ((al & 3) == 0) copy_words<uint32_t>(s, d, rb, lane, G);
        else copy_words<uint8_t>(s, d, rb, lane, G);
    }
}

int gather_rows(const GatherField* fields, int n_fields, const int64_t* idx, long n_rows, int mode,
                long E, long N, cudaStream_t st) {
    if (n_rows == 0 || n_fields == 0) return A2CB_OK;
    A2CB_REQUIRE(mode == 0 || (mode == 1 && E > 0 && N > 0), "gather_rows: bad row map (mode=%d E=%ld N=%ld)", mode, E, N);
    RowMap map{idx, mode, E, N};
    FieldPack narrow, wide;
    narrow.n = 0;
    wide.n = 0;
    static bool smem_set = false;
    for (int f = 0; f < n_fields; ++f) {
        const long rb = fields[f].row_bytes;
        A2CB_REQUIRE(rb > 0 && rb < (1L << 31), "gather_rows: field %d row_bytes=%ld", f, rb);
        const bool tma_ok = rb >= 1024 && rb % 16 == 0 &&
                            ((((uintptr_t)fields[f].src | (uintptr_t)fields[f].dst) & 15) == 0);
        if (tma_ok) {
            if (!smem_set) {
                cudaFuncSetAttribute(gather_rows_tma_kernel, cudaFuncAttributeMaxDynamicSharedMemorySize,
                                     kStages * kChunk);
                smem_set = true;
            }
            const long chunks = n_rows * ((rb + kChunk - 1) / kChunk);
            // 3 CTAs (64 KiB of stages each) fit one SM; persistent over the work list
            const long grid = chunks < 3L * kNumSM ? chunks : 3L * kNumSM;
            gather_rows_tma_kernel<<<(unsigned)grid, 32, kStages * kChunk, st>>>(
                (const uint8_t*)fields[f].src, (uint8_t*)fields[f].dst, rb, n_rows, map);
            int rc = check_launch("gather_rows_tma");
            if (rc) return rc;
        } else {
            FieldPack& p = rb <= 32 ? narrow : wide;
            A2CB_REQUIRE(p.n < kMaxFields, "gather_rows: too many fields");
            p.src[p.n] = (const uint8_t*)fields[f].src;
            p.dst[p.n] = (uint8_t*)fields[f].dst;
            p.row_bytes[p.n] = (int)rb;
            p.n++;
        }
    }
    if (narrow.n) {
        const long grid = (n_rows + 255) / 256;
        gather_fields_kernel<1><<<(unsigned)grid, 256, 0, st>>>(narrow, n_rows, map);
        int rc = check_launch("gather_fields<1>");
        if (rc) return rc;
    }
    if (wide.n) {
        const long grid = (n_rows * 32 + 255) / 256;
        gather_fields_kernel<32><<<(unsigned)grid, 256, 0, st>>>(wide, n_rows, map);
        int rc = check_launch("gather_fields<32>");
        if (rc) return rc;
    }
    return A2CB_OK;
}

}  // namespace a2cb
