// Peer-memory all-reduce of the gradient segments of the environment-sharded update.
//
// Why not ncclAllReduce: the update's kernels are persistent, one CTA per SM, with (nearly) all of an SM's shared memory.
// NCCL's kernel cannot become resident next to them, so an all-reduce "overlapped" on a side stream really runs in the gap
// between two of our kernels and holds the SMs it landed on for its whole duration -- transfer, hand-shakes and spinning
// included -- while the next persistent kernel starts with CTAs missing (measured: +0.26 ms per minibatch at N = 2 .. 8,
// the same at every N: it is residency, not bandwidth).  Here the exchange is three short data movers that never wait on
// the device:
//
//   scatter   buf slice j  -> inbox[rank] of rank j              (peer stores over NVLink, fire and forget)
//   reduce    sum_p inbox[p] (fixed order p = 0 .. W-1)  -> result slice `rank` of EVERY rank (peer stores)
//   gather    result -> buf                                       (local)
//
// and the two cross-GPU hand-shakes between them are STREAM memory operations (cuStreamWaitValue32 on a counter in local
// memory that the last CTA of the peers' previous phase bumps with red.release.sys): a waiting rank occupies no SM at all
// (A2CB_P2P_WAIT=spin: one polling warp instead).  Either way nothing may be launched for the FIRST time behind a pending
// hand-shake -- lazy module loading synchronises the context (tools/stream_probe.cu) -- so the kernels are loaded when the
// context is created.
// Every element is summed once, by one rank, in rank order: results are bit-identical on all ranks and independent of the
// grid.  The buffers ([inbox | result | counters], one cudaMalloc per rank) are mapped into the peers through CUDA IPC
// handles that the host side passes around (a2cb_p2p_ipc_handle / a2cb_p2p_open_ipc), or handed over as plain pointers by
// a caller that has its own mapping (a2cb_p2p_set_peers: same-process tests, symmetric-memory allocators).
#include <cuda.h>
#include <stdlib.h>
#include <string.h>

#include "common.cuh"
#include "decls.cuh"

namespace a2cb {

namespace {

constexpr int kMaxWorld = 16;
constexpr int kThreads = 256;
constexpr int kMaxCtas = 32;                // per phase: data movers, sized for NVLink latency x bandwidth, not for the SMs
constexpr int kCounterBytes = 256;          // [0] phase-A arrivals, [16] phase-B arrivals, [32] / [33] last-CTA tickets

struct Peers {
    char* base[1 + kMaxWorld];              // [0]: this rank's buffer, [1 + r]: the buffer of rank r as mapped here
};

struct P2P {
    int world, rank, device;
    long cap;                               // floats per region
    char* local;                            // cudaMalloc: [inbox cap][result cap][counters]
    Peers peers;
    bool opened[kMaxWorld];
    bool connected;
    unsigned seq;                           // all-reduces issued so far (identical on every rank: same program order)
    int wait_mode;                          // 0: stream memory operation (default), 1: one polling warp
};

typedef CUresult (*WaitValueFn)(CUstream, CUdeviceptr, cuuint32_t, unsigned int);
WaitValueFn wait_value_fn() {
    static WaitValueFn fn = nullptr;
    static bool tried = false;
    if (!tried) {
        tried = true;
        void* p = nullptr;
        cudaDriverEntryPointQueryResult q;
        if (cudaGetDriverEntryPoint("cuStreamWaitValue32", &p, cudaEnableDefault, &q) == cudaSuccess &&
            q == cudaDriverEntryPointSuccess)
            fn = reinterpret_cast<WaitValueFn>(p);
    }
    return fn;
}

__device__ __forceinline__ unsigned* counters(char* base, long cap) {
    return reinterpret_cast<unsigned*>(base + (size_t)cap * 8);
}

// Last CTA of the grid tells every rank (itself included) that this rank's phase is complete.  No shared memory: the
// kernels must fit next to CTAs that own all of it.
__device__ __forceinline__ void phase_done(const Peers& P, int world, long cap, int which) {
    __threadfence_system();                                    // this thread's peer stores are performed
    unsigned* mine = counters(P.base[0], cap);
    int last = 0;
    __syncthreads();
    if (threadIdx.x == 0) last = atomicAdd(mine + 32 + which, 1u) == gridDim.x - 1;
    last = __syncthreads_or(last);
    if (!last) return;
    __threadfence_system();
    if (threadIdx.x == 0) mine[32 + which] = 0;                // the next launch that takes tickets starts after this one
    if ((int)threadIdx.x < world) {
        unsigned* c = counters(P.base[1 + threadIdx.x], cap) + 16 * which;
        asm volatile("red.release.sys.global.add.u32 [%0], 1;" ::"l"(c) : "memory");
    }
}

__device__ __forceinline__ float4 ld_cg4(const float* p) { return __ldcg(reinterpret_cast<const float4*>(p)); }

__global__ void __launch_bounds__(kThreads) p2p_scatter_kernel(Peers P, const float* __restrict__ buf, long count, long S,
                                                               int world, int rank, long cap) {
    const long n4 = count >> 2;
    const long stride = (long)gridDim.x * kThreads;
    for (long i = (long)blockIdx.x * kThreads + threadIdx.x; i < n4; i += stride) {
        const long e = i << 2;
        const long j = e / S;                                  // S is a multiple of 4: a float4 never straddles slices
        float* dst = reinterpret_cast<float*>(P.base[1 + j]) + (long)rank * S + (e - j * S);
        *reinterpret_cast<float4*>(dst) = *reinterpret_cast<const float4*>(buf + e);
    }
    if (blockIdx.x == 0 && threadIdx.x < (count & 3)) {
        const long e = (n4 << 2) + threadIdx.x;
        const long j = e / S;
        reinterpret_cast<float*>(P.base[1 + j])[(long)rank * S + (e - j * S)] = buf[e];
    }
    phase_done(P, world, cap, 0);
}

__global__ void __launch_bounds__(kThreads) p2p_reduce_kernel(Peers P, long count, long S, int world, int rank, long cap) {
    const long lo = (long)rank * S;
    const long n = count - lo < S ? count - lo : S;            // may be <= 0 for the last ranks of a tiny segment
    const float* in = reinterpret_cast<const float*>(P.base[0]);
    const long n4 = n > 0 ? n >> 2 : 0;
    const long stride = (long)gridDim.x * kThreads;
    for (long i = (long)blockIdx.x * kThreads + threadIdx.x; i < n4; i += stride) {
        const long e = i << 2;
        float4 a = ld_cg4(in + e);
        for (int p = 1; p < world; ++p) {
            const float4 b = ld_cg4(in + (long)p * S + e);
            a.x += b.x; a.y += b.y; a.z += b.z; a.w += b.w;
        }
        for (int q = 0; q < world; ++q)
            *reinterpret_cast<float4*>(reinterpret_cast<float*>(P.base[1 + q]) + cap + lo + e) = a;
    }
    if (blockIdx.x == 0 && n > 0 && threadIdx.x < (n & 3)) {
        const long e = (n4 << 2) + threadIdx.x;
        float a = __ldcg(in + e);
        for (int p = 1; p < world; ++p) a += __ldcg(in + (long)p * S + e);
        for (int q = 0; q < world; ++q) reinterpret_cast<float*>(P.base[1 + q])[cap + lo + e] = a;
    }
    phase_done(P, world, cap, 1);
}

__global__ void __launch_bounds__(kThreads) p2p_gather_kernel(Peers P, float* __restrict__ buf, long count, long cap) {
    const float* res = reinterpret_cast<const float*>(P.base[0]) + cap;
    const long n4 = count >> 2;
    const long stride = (long)gridDim.x * kThreads;
    for (long i = (long)blockIdx.x * kThreads + threadIdx.x; i < n4; i += stride)
        *reinterpret_cast<float4*>(buf + (i << 2)) = ld_cg4(res + (i << 2));
    if (blockIdx.x == 0 && threadIdx.x < (count & 3)) {
        const long e = (n4 << 2) + threadIdx.x;
        buf[e] = __ldcg(res + e);
    }
}

// the hand-shake without stream memory operations: ONE resident warp polls the local counter
__global__ void p2p_wait_kernel(const unsigned* counter, unsigned target) {
    if (threadIdx.x == 0) {
        unsigned v;
        do {
            asm volatile("ld.acquire.sys.global.u32 %0, [%1];" : "=r"(v) : "l"(counter) : "memory");
            if ((int)(v - target) < 0) __nanosleep(200);
        } while ((int)(v - target) < 0);
    }
}

int wait_counter(P2P* c, int which, unsigned target, cudaStream_t st) {
    unsigned* addr = reinterpret_cast<unsigned*>(c->local + (size_t)c->cap * 8) + 16 * which;
    if (c->wait_mode == 0) {
        const CUresult r = wait_value_fn()((CUstream)st, (CUdeviceptr)addr, target, CU_STREAM_WAIT_VALUE_GEQ);
        if (r != CUDA_SUCCESS) { set_error("p2p: cuStreamWaitValue32 failed (%d)", (int)r); return A2CB_ERR_CUDA; }
        return A2CB_OK;
    }
    p2p_wait_kernel<<<1, 32, 0, st>>>(addr, target);
    return check_launch("p2p_wait");
}

int grid_for(long floats) {
    const long per_cta = (long)kThreads * 4 * 4;               // four float4 per thread
    long g = (floats + per_cta - 1) / per_cta;
    return (int)(g < 1 ? 1 : g > kMaxCtas ? kMaxCtas : g);
}

}  // namespace

int p2p_create(void** out, int world, int rank, long capacity_floats) {
    A2CB_REQUIRE(out && world >= 1 && world <= kMaxWorld && rank >= 0 && rank < world && capacity_floats > 0,
                 "p2p_create: bad arguments (world <= %d)", kMaxWorld);
    P2P* c = (P2P*)calloc(1, sizeof(P2P));
    A2CB_REQUIRE(c, "p2p_create: out of memory");
    c->world = world;
    c->rank = rank;
    c->cap = (capacity_floats + 4 * world + 255) / 256 * 256;  // every slice start stays 16-byte aligned
    cudaError_t e = cudaGetDevice(&c->device);
    const size_t bytes = (size_t)c->cap * 8 + kCounterBytes;
    if (e == cudaSuccess) e = cudaMalloc(&c->local, bytes);
    if (e == cudaSuccess) e = cudaMemset(c->local, 0, bytes);
    if (e == cudaSuccess) e = cudaDeviceSynchronize();
    if (e != cudaSuccess) { set_error("p2p_create: %s", cudaGetErrorString(e)); free(c); return A2CB_ERR_CUDA; }
    const char* w = getenv("A2CB_P2P_WAIT");
    c->wait_mode = (w && !strcmp(w, "spin")) || !wait_value_fn() ? 1 : 0;
    c->peers.base[0] = c->local;
    c->peers.base[1 + rank] = c->local;
    // CUDA loads kernels lazily, and a first launch may need a context-wide synchronisation: behind a pending hand-shake
    // that would stall the launching thread until the peers arrive (and dead-lock emulated ranks sharing a process).  Load
    // everything now: one empty launch of each kernel (world = 0: nobody is signalled).
    p2p_scatter_kernel<<<1, kThreads>>>(c->peers, nullptr, 0, 4, 0, rank, c->cap);
    p2p_reduce_kernel<<<1, kThreads>>>(c->peers, 0, 4, 0, rank, c->cap);
    p2p_gather_kernel<<<1, kThreads>>>(c->peers, nullptr, 0, c->cap);
    p2p_wait_kernel<<<1, 32>>>(reinterpret_cast<unsigned*>(c->local + (size_t)c->cap * 8), 0);
    e = cudaDeviceSynchronize();
    if (e != cudaSuccess) { set_error("p2p_create: warm-up launches: %s", cudaGetErrorString(e)); cudaFree(c->local); free(c); return A2CB_ERR_CUDA; }
    *out = c;
    return A2CB_OK;
}

int p2p_ipc_handle(void* ctx, void* out64) {
    P2P* c = (P2P*)ctx;
    A2CB_REQUIRE(c && out64, "p2p_ipc_handle: null pointer");
    static_assert(sizeof(cudaIpcMemHandle_t) == 64, "cudaIpcMemHandle_t is 64 bytes");
    cudaIpcMemHandle_t h;
    const cudaError_t e = cudaIpcGetMemHandle(&h, c->local);
    if (e != cudaSuccess) { set_error("p2p_ipc_handle: %s", cudaGetErrorString(e)); return A2CB_ERR_CUDA; }
    memcpy(out64, &h, 64);
    return A2CB_OK;
}

int p2p_open_ipc(void* ctx, const void* handles) {
    P2P* c = (P2P*)ctx;
    A2CB_REQUIRE(c && handles, "p2p_open_ipc: null pointer");
    for (int r = 0; r < c->world; ++r) {
        if (r == c->rank) continue;
        cudaIpcMemHandle_t h;
        memcpy(&h, (const char*)handles + 64 * r, 64);
        void* p = nullptr;
        const cudaError_t e = cudaIpcOpenMemHandle(&p, h, cudaIpcMemLazyEnablePeerAccess);
        if (e != cudaSuccess) {
            cudaGetLastError();
            set_error("p2p_open_ipc: rank %d cannot map the buffer of rank %d: %s", c->rank, r, cudaGetErrorString(e));
            return A2CB_ERR_CUDA;
        }
        c->peers.base[1 + r] = (char*)p;
        c->opened[r] = true;
    }
    c->connected = true;
    return A2CB_OK;
}

int p2p_set_peers(void* ctx, void* const* bases) {
    P2P* c = (P2P*)ctx;
    A2CB_REQUIRE(c && bases, "p2p_set_peers: null pointer");
    for (int r = 0; r < c->world; ++r) {
        A2CB_REQUIRE(r == c->rank || bases[r], "p2p_set_peers: no buffer for rank %d", r);
        if (r != c->rank) c->peers.base[1 + r] = (char*)bases[r];
    }
    c->connected = true;
    return A2CB_OK;
}

int p2p_local_base(void* ctx, void** base) {
    P2P* c = (P2P*)ctx;
    A2CB_REQUIRE(c && base, "p2p_local_base: null pointer");
    *base = c->local;
    return A2CB_OK;
}

long p2p_capacity(void* ctx) { return ctx ? ((P2P*)ctx)->cap - 4 * ((P2P*)ctx)->world : 0; }

int p2p_allreduce_sum_f32(void* ctx, float* buf, long count, cudaStream_t st) {
    P2P* c = (P2P*)ctx;
    A2CB_REQUIRE(c && buf && count > 0, "p2p_allreduce_sum_f32: bad arguments");
    A2CB_REQUIRE(c->connected, "p2p_allreduce_sum_f32: peers are not mapped yet (a2cb_p2p_open_ipc / a2cb_p2p_set_peers)");
    A2CB_REQUIRE(((uintptr_t)buf & 15) == 0, "p2p_allreduce_sum_f32: the segment must start 16-byte aligned");
    const int W = c->world;
    long S = (count + W - 1) / W;
    S = (S + 3) & ~3L;
    A2CB_REQUIRE(S * W <= c->cap, "p2p_allreduce_sum_f32: %ld floats exceed the capacity of the exchange buffer", count);
    const unsigned target = (++c->seq) * (unsigned)W;
    p2p_scatter_kernel<<<grid_for(count), kThreads, 0, st>>>(c->peers, buf, count, S, W, c->rank, c->cap);
    int rc = check_launch("p2p_scatter");
    if (rc == A2CB_OK) rc = wait_counter(c, 0, target, st);
    if (rc != A2CB_OK) return rc;
    p2p_reduce_kernel<<<grid_for(S), kThreads, 0, st>>>(c->peers, count, S, W, c->rank, c->cap);
    rc = check_launch("p2p_reduce");
    if (rc == A2CB_OK) rc = wait_counter(c, 1, target, st);
    if (rc != A2CB_OK) return rc;
    p2p_gather_kernel<<<grid_for(count), kThreads, 0, st>>>(c->peers, buf, count, c->cap);
    return check_launch("p2p_gather");
}

int p2p_destroy(void* ctx) {
    P2P* c = (P2P*)ctx;
    if (!c) return A2CB_OK;
    for (int r = 0; r < c->world; ++r)
        if (c->opened[r]) cudaIpcCloseMemHandle(c->peers.base[1 + r]);
    cudaFree(c->local);
    free(c);
    return A2CB_OK;
}

}  // namespace a2cb
