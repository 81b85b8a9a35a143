// Micro-probe (B200): what does ONE small tcgen05.mma.kind::f16 (M = 128, K = 16, N = 16..128) cost when
//   (a) all instructions accumulate into the same TMEM accumulator (a dependent chain),
//   (b) consecutive instructions rotate over 2..16 accumulators,
//   (c) the A operand comes from shared memory / from tensor memory?
// The GRU recurrence (csrc/gru_tc.cu) issues 96 / 72 such instructions per time step; its phase stamps show 59 / 42
// cycles per instruction against a data-path floor of 8 (M * N / 256).  Also times, single CTA, the exchange primitives
// of that kernel: a bulk copy L2 -> shared memory against plain 16-byte loads, and a store + red.release.gpu.
//
//   nvcc -gencode arch=compute_100a,code=sm_100a -O3 -o tools/mma_probe.bin tools/mma_probe.cu && tools/mma_probe.bin
#include <cuda_runtime.h>
#include <cuda_fp16.h>
#include <stdint.h>
#include <stdio.h>

__device__ __forceinline__ uint32_t smem_u32(const void* p) { return (uint32_t)__cvta_generic_to_shared(p); }
__device__ __forceinline__ uint64_t mk_desc(uint32_t addr, uint32_t lbo, uint32_t sbo) {
    return (uint64_t)((addr & 0x3FFFF) >> 4) | ((uint64_t)((lbo >> 4) & 0x3FFF) << 16) | ((uint64_t)((sbo >> 4) & 0x3FFF) << 32) |
           (1ull << 46) | (2ull << 61);
}
__device__ __forceinline__ void mma_ss(uint32_t d, uint64_t a, uint64_t b, uint32_t idesc, uint32_t acc) {
    asm volatile("{\n\t.reg .pred p;\n\tsetp.ne.b32 p, %4, 0;\n\ttcgen05.mma.cta_group::1.kind::f16 [%0], %1, %2, %3, p;\n\t}\n" ::"r"(d),
                 "l"(a), "l"(b), "r"(idesc), "r"(acc)
                 : "memory");
}
__device__ __forceinline__ void mma_ts(uint32_t d, uint32_t a, uint64_t b, uint32_t idesc, uint32_t acc) {
    asm volatile("{\n\t.reg .pred p;\n\tsetp.ne.b32 p, %4, 0;\n\ttcgen05.mma.cta_group::1.kind::f16 [%0], [%1], %2, %3, p;\n\t}\n" ::"r"(d),
                 "r"(a), "l"(b), "r"(idesc), "r"(acc)
                 : "memory");
}
__device__ __forceinline__ void mbar_wait(uint64_t* bar, uint32_t parity) {
    uint32_t ok = 0;
    while (!ok)
        asm volatile("{\n.reg .pred p;\nmbarrier.try_wait.parity.shared::cta.b64 p, [%1], %2;\nselp.u32 %0, 1, 0, p;\n}\n"
                     : "=r"(ok)
                     : "r"(smem_u32(bar)), "r"(parity)
                     : "memory");
}

__device__ __forceinline__ bool elect_one() {
    uint32_t pred;
    asm volatile("{\n\t.reg .pred P1;\n\telect.sync _|P1, 0xFFFFFFFF;\n\tselp.u32 %0, 1, 0, P1;\n\t}\n" : "=r"(pred));
    return pred != 0;
}

struct Case { int n, nacc, ts, count; };

__global__ void __launch_bounds__(128, 1) mma_probe(const Case* cases, int ncase, long long* out) {
    extern __shared__ uint8_t dyn[];
    __shared__ uint64_t bar;
    __shared__ uint32_t slot;
    const uint32_t base = (smem_u32(dyn) + 1023u) & ~1023u;
    uint8_t* sm = dyn + (base - smem_u32(dyn));
    const int tid = threadIdx.x, warp = tid >> 5;
    // A: 128 rows x 64 halves (16 KiB), B: up to 128 rows x 64 halves (16 KiB): small finite numbers
    for (int i = tid; i < 16384; i += 128) reinterpret_cast<__half*>(sm)[i] = __float2half(0.001f * (float)(i % 13));
    if (tid == 0) {
        asm volatile("mbarrier.init.shared::cta.b64 [%0], 1;" ::"r"(smem_u32(&bar)));
        asm volatile("fence.mbarrier_init.release.cluster;" ::: "memory");
    }
    if (warp == 0) {
        asm volatile("tcgen05.alloc.cta_group::1.sync.aligned.shared::cta.b32 [%0], 512;" ::"r"(smem_u32(&slot)));
        asm volatile("tcgen05.relinquish_alloc_permit.cta_group::1.sync.aligned;" ::);
    }
    asm volatile("fence.proxy.async.shared::cta;" ::: "memory");
    asm volatile("tcgen05.fence::before_thread_sync;" ::: "memory");
    __syncthreads();
    asm volatile("tcgen05.fence::after_thread_sync;" ::: "memory");
    const uint32_t tmem = slot;
    // A image in tensor memory: columns 448..479 (4 K-steps x 8 columns), any finite content
    {
        uint32_t r[8];
        for (int j = 0; j < 8; ++j) r[j] = 0x1c001c00u;
        for (int c = 0; c < 32; c += 8)
            asm volatile("tcgen05.st.sync.aligned.32x32b.x8.b32 [%0], {%1, %2, %3, %4, %5, %6, %7, %8};\n" ::"r"(
                             tmem + ((uint32_t)(32 * warp) << 16) + 448u + (uint32_t)c),
                         "r"(r[0]), "r"(r[1]), "r"(r[2]), "r"(r[3]), "r"(r[4]), "r"(r[5]), "r"(r[6]), "r"(r[7])
                         : "memory");
        asm volatile("tcgen05.wait::st.sync.aligned;" ::: "memory");
    }
    asm volatile("tcgen05.fence::before_thread_sync;" ::: "memory");
    __syncthreads();
    asm volatile("tcgen05.fence::after_thread_sync;" ::: "memory");
    uint32_t parity = 0;
    if (warp == 0) {
        // the whole (converged) warp walks the loops, ONE elected lane issues: with the loop under `if (tid == 0)` the
        // compiler wraps every UTCHMMA in an elect-and-branch sequence and the issue itself costs ~110 cycles per MMA
        for (int rep = 0; rep < 3; ++rep)                 // rep 0 warms the instruction cache
            for (int c = 0; c < ncase; ++c) {
                const Case cs = cases[c];
                const uint32_t idesc = (1u << 4) | ((uint32_t)(cs.n >> 3) << 17) | ((uint32_t)(128 >> 4) << 24);
                const uint64_t a0 = mk_desc(base, 16, 1024), b0 = mk_desc(base + 16384u, 16, 1024);
                long long t0 = 0, t1 = 0;
                __syncwarp();
                if (elect_one()) {
                    int acc = 0;
                    t0 = clock64();
                    if (cs.ts) {
                        for (int i = 0; i < cs.count; i += 4) {
#pragma unroll
                            for (int j = 0; j < 4; ++j) {
                                mma_ts(tmem + (uint32_t)(acc * cs.n), tmem + 448u + 8u * j, b0 + 2 * j, idesc, (i + j) >= cs.nacc ? 1u : 0u);
                                if (++acc == cs.nacc) acc = 0;
                            }
                        }
                    } else {
                        for (int i = 0; i < cs.count; i += 4) {
#pragma unroll
                            for (int j = 0; j < 4; ++j) {
                                mma_ss(tmem + (uint32_t)(acc * cs.n), a0 + 2 * j, b0 + 2 * j, idesc, (i + j) >= cs.nacc ? 1u : 0u);
                                if (++acc == cs.nacc) acc = 0;
                            }
                        }
                    }
                    t1 = clock64();
                    asm volatile("tcgen05.commit.cta_group::1.mbarrier::arrive::one.shared::cluster.b64 [%0];" ::"r"(smem_u32(&bar)) : "memory");
                }
                __syncwarp();
                mbar_wait(&bar, parity);
                parity ^= 1u;
                const long long t2 = clock64();
                asm volatile("tcgen05.fence::after_thread_sync;" ::: "memory");
                if (rep == 2 && t0 != 0) { out[2 * c] = t1 - t0; out[2 * c + 1] = t2 - t0; }
            }
    }
    asm volatile("tcgen05.fence::before_thread_sync;" ::: "memory");
    __syncthreads();
    if (warp == 0) asm volatile("tcgen05.dealloc.cta_group::1.sync.aligned.b32 %0, 512;" ::"r"(tmem));
}

// ---- exchange primitives, single CTA of 256 threads: src is L2-resident (written by a previous kernel) ----
__global__ void __launch_bounds__(256, 1) xchg_probe(const uint4* src, uint4* dst, unsigned* flag, long long* out) {
    __shared__ __align__(1024) uint8_t buf[16384];
    __shared__ uint64_t bar;
    const int tid = threadIdx.x;
    if (tid == 0) {
        asm volatile("mbarrier.init.shared::cta.b64 [%0], 1;" ::"r"(smem_u32(&bar)));
        asm volatile("fence.mbarrier_init.release.cluster;" ::: "memory");
    }
    __syncthreads();
    uint32_t parity = 0;
    for (int rep = 0; rep < 4; ++rep) {
        const uint4* s = src + rep * 4096;               // fresh 16 KiB per repetition (L1 never holds it)
        // (1) bulk copies: 16 KiB as one / as four copies, time to first completed 4 KiB and to all
        long long t0 = 0, t1 = 0, t2 = 0;
        if (tid == 0) {
            t0 = clock64();
            asm volatile("mbarrier.arrive.expect_tx.shared::cta.b64 _, [%0], %1;" ::"r"(smem_u32(&bar)), "r"(16384u) : "memory");
            for (int c = 0; c < 4; ++c)
                asm volatile("cp.async.bulk.shared::cluster.global.mbarrier::complete_tx::bytes [%0], [%1], %2, [%3];" ::"r"(
                                 smem_u32(buf) + c * 4096u),
                             "l"(reinterpret_cast<const uint8_t*>(s) + c * 4096), "r"(4096u), "r"(smem_u32(&bar))
                             : "memory");
            t1 = clock64();
            mbar_wait(&bar, parity);
            t2 = clock64();
        }
        parity ^= 1u;
        __syncthreads();
        // (2) plain loads: 256 threads x 4 x 16 bytes (ld.global.cg), stored to shared memory, CTA barrier
        const uint4* s2 = src + 4 * 4096 + rep * 4096;
        const long long u0 = clock64();
        uint4 v[4];
#pragma unroll
        for (int j = 0; j < 4; ++j) v[j] = __ldcg(s2 + j * 256 + tid);
#pragma unroll
        for (int j = 0; j < 4; ++j) reinterpret_cast<uint4*>(buf)[j * 256 + tid] = v[j];
        asm volatile("fence.proxy.async.shared::cta;" ::: "memory");
        __syncthreads();
        const long long u1 = clock64();
        // (3) release: 64 threads store 16 bytes each, CTA barrier, one red.release.gpu
        if (tid < 64) dst[rep * 64 + tid] = v[0];
        __syncthreads();
        const long long w0 = clock64();
        if (tid == 0) asm volatile("red.release.gpu.global.add.u32 [%0], 1;" ::"l"(flag) : "memory");
        const long long w1 = clock64();
        // (4) the same with an explicit fence + relaxed red
        if (tid < 64) dst[1024 + rep * 64 + tid] = v[1];
        __syncthreads();
        const long long x0 = clock64();
        if (tid == 0) {
            asm volatile("fence.acq_rel.gpu;" ::: "memory");
            asm volatile("red.relaxed.gpu.global.add.u32 [%0], 1;" ::"l"(flag + 32) : "memory");
        }
        const long long x1 = clock64();
        // (5) acquire poll of an already-set flag (one L2 round trip)
        long long y0 = 0, y1 = 0;
        if (tid == 0) {
            y0 = clock64();
            unsigned f;
            asm volatile("ld.acquire.gpu.global.u32 %0, [%1];" : "=r"(f) : "l"(flag + 64) : "memory");
            y1 = clock64() + (f == 12345u ? 1 : 0);
        }
        // (6) relaxed / volatile / 16-byte relaxed loads of L2-resident words, and the issue cost of a relaxed red
        long long z[5] = {0, 0, 0, 0, 0};
        if (tid == 0) {
            unsigned f;
            uint4 q;
            z[0] = clock64();
            asm volatile("ld.relaxed.gpu.global.u32 %0, [%1];" : "=r"(f) : "l"(flag + 96) : "memory");
            z[1] = clock64() + (f == 12345u ? 1 : 0);
            asm volatile("ld.volatile.global.u32 %0, [%1];" : "=r"(f) : "l"(flag + 128) : "memory");
            z[2] = clock64() + (f == 12345u ? 1 : 0);
            asm volatile("ld.relaxed.gpu.global.v4.u32 {%0, %1, %2, %3}, [%4];" : "=r"(q.x), "=r"(q.y), "=r"(q.z), "=r"(q.w) : "l"(src + 8 * 4096 + rep * 64) : "memory");
            z[3] = clock64() + (q.x == 12345u ? 1 : 0);
            asm volatile("red.relaxed.gpu.global.add.u32 [%0], 1;" ::"l"(flag + 160) : "memory");
            z[4] = clock64();
        }
        if (tid == 0 && rep == 3) { out[6] = z[1] - z[0]; out[7] = z[2] - z[1]; out[8] = z[3] - z[2]; out[9] = z[4] - z[3]; }
        if (tid == 0 && rep == 3) {
            out[0] = t1 - t0; out[1] = t2 - t0; out[2] = u1 - u0; out[3] = w1 - w0; out[4] = x1 - x0; out[5] = y1 - y0;
        }
        __syncthreads();
    }
}

// ---- 128 CTAs fetch 16 KiB each at the same moment (8 bulk copies of 2 KiB, as the GRU forward kernel does):
//      mode 0 = every CTA its own 16 KiB, mode 1 = the 16 CTAs of a group read the SAME 16 KiB (all-gather image) ----
__global__ void __launch_bounds__(32, 1) fetch_probe(const uint8_t* src, unsigned* counter, long long* out, int mode, int round) {
    __shared__ __align__(1024) uint8_t buf[16384];
    __shared__ uint64_t bar;
    if (threadIdx.x == 0) {
        asm volatile("mbarrier.init.shared::cta.b64 [%0], 1;" ::"r"(smem_u32(&bar)));
        asm volatile("fence.mbarrier_init.release.cluster;" ::: "memory");
        const uint8_t* s = src + (size_t)(mode ? (blockIdx.x / 16) : blockIdx.x) * 16384 + (size_t)round * (4u << 20);
        atomicAdd(counter, 1u);
        while (*(volatile unsigned*)counter < gridDim.x) {}
        const long long t0 = clock64();
        asm volatile("mbarrier.arrive.expect_tx.shared::cta.b64 _, [%0], %1;" ::"r"(smem_u32(&bar)), "r"(16384u) : "memory");
        for (int c = 0; c < 8; ++c)
            asm volatile("cp.async.bulk.shared::cluster.global.mbarrier::complete_tx::bytes [%0], [%1], %2, [%3];" ::"r"(
                             smem_u32(buf) + c * 2048u),
                         "l"(s + c * 2048), "r"(2048u), "r"(smem_u32(&bar))
                         : "memory");
        mbar_wait(&bar, 0);
        out[blockIdx.x] = clock64() - t0;
    }
}

__global__ void fill(uint4* p, int n) {
    for (int i = blockIdx.x * blockDim.x + threadIdx.x; i < n; i += gridDim.x * blockDim.x) p[i] = make_uint4(i, i, i, i);
}

int main() {
    Case h[64];
    int nc = 0;
    const int ns[] = {16, 32, 64, 128};
    for (int ts = 0; ts < 2; ++ts)
        for (int in = 0; in < 4; ++in)
            for (int nacc = 1; nacc <= 16; nacc *= 2) {
                if (nacc * ns[in] > 448) continue;
                h[nc++] = Case{ns[in], nacc, ts, 96};
            }
    Case* d;
    long long* out;
    cudaMalloc(&d, sizeof(h));
    cudaMalloc(&out, 4096);
    cudaMemcpy(d, h, sizeof(Case) * nc, cudaMemcpyHostToDevice);
    cudaFuncSetAttribute(mma_probe, cudaFuncAttributeMaxDynamicSharedMemorySize, 40960);
    mma_probe<<<1, 128, 40960>>>(d, nc, out);
    cudaError_t e = cudaDeviceSynchronize();
    if (e != cudaSuccess) { printf("mma_probe failed: %s\n", cudaGetErrorString(e)); return 1; }
    long long ho[512];
    cudaMemcpy(ho, out, sizeof(long long) * 2 * nc, cudaMemcpyDeviceToHost);
    for (int c = 0; c < nc; ++c)
        printf("{\"probe\": \"mma\", \"N\": %d, \"accumulators\": %d, \"a_from_tmem\": %d, \"count\": %d, \"issue_cycles\": %lld, "
               "\"done_cycles\": %lld, \"cycles_per_mma\": %.1f}\n",
               h[c].n, h[c].nacc, h[c].ts, h[c].count, ho[2 * c], ho[2 * c + 1], (double)ho[2 * c + 1] / h[c].count);
    uint4 *src, *dst;
    unsigned* flag;
    cudaMalloc(&src, 16 * 65536);
    cudaMalloc(&dst, 16 * 4096);
    cudaMalloc(&flag, 1024);
    cudaMemset(flag, 0, 1024);
    fill<<<64, 256>>>(src, 65536);
    xchg_probe<<<1, 256>>>(src, dst, flag, out);
    e = cudaDeviceSynchronize();
    if (e != cudaSuccess) { printf("xchg_probe failed: %s\n", cudaGetErrorString(e)); return 1; }
    cudaMemcpy(ho, out, sizeof(long long) * 10, cudaMemcpyDeviceToHost);
    printf("{\"probe\": \"exchange\", \"bulk_issue\": %lld, \"bulk_16k_done\": %lld, \"ldcg_16k_to_smem_bar\": %lld, \"red_release\": %lld, "
           "\"fence_plus_red\": %lld, \"ld_acquire\": %lld, \"ld_relaxed\": %lld, \"ld_volatile\": %lld, \"ld_relaxed_v4\": %lld, "
           "\"red_relaxed_issue\": %lld}\n",
           ho[0], ho[1], ho[2], ho[3], ho[4], ho[5], ho[6], ho[7], ho[8], ho[9]);
    // fetch probe: warm L2 with the source (fill wrote it), 2 modes x 2 rounds
    unsigned* counters;
    cudaMalloc(&counters, 64);
    cudaMemset(counters, 0, 64);
    uint8_t* big;
    cudaMalloc(&big, 16u << 20);
    fill<<<256, 256>>>(reinterpret_cast<uint4*>(big), (16 << 20) / 16);
    cudaDeviceSynchronize();
    for (int k = 0; k < 4; ++k) {
        const int mode = k & 1;
        fetch_probe<<<128, 32>>>(big, counters + k, out, mode, k);
        e = cudaDeviceSynchronize();
        if (e != cudaSuccess) { printf("fetch_probe failed: %s\n", cudaGetErrorString(e)); return 1; }
        cudaMemcpy(ho, out, sizeof(long long) * 128, cudaMemcpyDeviceToHost);
        long long mx = 0, mn = 1ll << 60, sum = 0;
        for (int i = 0; i < 128; ++i) { mx = ho[i] > mx ? ho[i] : mx; mn = ho[i] < mn ? ho[i] : mn; sum += ho[i]; }
        printf("{\"probe\": \"fetch16k_x128\", \"same_image_per_16_ctas\": %d, \"round\": %d, \"min\": %lld, \"mean\": %lld, \"max\": %lld}\n",
               mode, k, mn, sum / 128, mx);
    }
    return 0;
}
