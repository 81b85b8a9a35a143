// Persistent GRU recurrence: the whole [T, N] sequence in ONE cooperative launch per direction.
//
// The per-step formulation (gru.cu + one tiny GEMM per step) is launch- and latency-bound: at the
// C5 minibatch shape (T = 128, N = 64 environments, H = 512) the 2 x 128 sequential GEMMs of
// 0.1 GFLOP each cost ~18 us apiece although they are ~1.5 us of fp32 math.  Here CTA (bx, by) owns 32
// hidden units and 8 environments: its slice of W_hh (96 gate rows forward, 32 columns backward, 192 KiB)
// stays resident in shared memory for all T steps, and per step it streams only ITS 8 rows of the previous
// state / gate gradients from L2 (16 / 48 KiB; a first version in which every CTA read all 64 rows was L2-
// bandwidth-bound at 17 / 50 MB per step).  16 x 8 = 128 CTAs, one per SM.  A warp takes 4 units, its lanes
// split the reduction dimension, partial sums are transpose-reduced with shuffles so that lane 4 r + j
// finishes (environment r, unit j), and the CTAs meet at ONE grid barrier per step.  Episode boundaries
// are the mask multiply, exactly as in gru.cu (reference model.py:111-166).
//
// Same buffers and the same saved quantities as the per-step path, so either path can feed the
// weight-gradient GEMMs and the tests compare both against the reference goldens.
#include <cooperative_groups.h>
#include <stdlib.h>

#include "common.cuh"
#include "decls.cuh"
#include "gru.cuh"

namespace cg = cooperative_groups;

namespace a2cb {

constexpr int PG_UW = 4;         // hidden units per warp (12 gate rows)
constexpr int PG_NW = 8;         // warps per CTA
constexpr int PG_UB = PG_UW * PG_NW;   // hidden units per CTA (32): H / 32 unit groups
constexpr int PG_RW = 8;         // environments per CTA and pass (all warps share them through L1)
constexpr int PG_NT = 32 * PG_NW;

__device__ __forceinline__ float pg_sigmoid(float x) { return 1.f / (1.f + expf(-x)); }

__device__ __forceinline__ float pg_mask(const GruDesc& d, int t, int n) {
    const long row = (long)t * d.N + n;
    const long src = d.idx ? (long)d.idx[row] : row;
    return d.masks[src];
}

__device__ __forceinline__ float pg_dot4(const float4 a, const float4 b, float acc) {
    acc = fmaf(a.x, b.x, acc); acc = fmaf(a.y, b.y, acc);
    acc = fmaf(a.z, b.z, acc); acc = fmaf(a.w, b.w, acc);
    return acc;
}

// Barrier among the CTAs that share an environment group (same blockIdx.y): only those exchange data --
// environments are independent -- so 16 CTAs meet instead of the whole grid, on their own monotonically
// increasing counter.  release / acquire at gpu scope order the state writes of every CTA before the reads
// of the others (the CTA-wide bar.sync on both sides extends that to all of its threads).
__device__ unsigned pg_counters[2][256];

__device__ __forceinline__ void pg_group_sync(unsigned* ctr, unsigned target) {
    __syncthreads();
    if (threadIdx.x == 0) {
        asm volatile("red.release.gpu.global.add.u32 [%0], 1;" ::"l"(ctr) : "memory");
        unsigned v;
        do {
            asm volatile("ld.acquire.gpu.global.u32 %0, [%1];" : "=r"(v) : "l"(ctr) : "memory");
        } while (v < target);
    }
    __syncthreads();
}

// The same barrier in two halves: arrive right after the stores the other CTAs wait for, then do work that nobody
// waits for (non-critical stores, prefetches of the next step's operands), then wait.
__device__ __forceinline__ void pg_group_arrive(unsigned* ctr) {
    __syncthreads();
    if (threadIdx.x == 0) asm volatile("red.release.gpu.global.add.u32 [%0], 1;" ::"l"(ctr) : "memory");
}
__device__ __forceinline__ void pg_group_wait(unsigned* ctr, unsigned target) {
    if (threadIdx.x == 0) {
        unsigned v;
        do {
            asm volatile("ld.acquire.gpu.global.u32 %0, [%1];" : "=r"(v) : "l"(ctr) : "memory");
        } while (v < target);
    }
    __syncthreads();
}

// Butterfly "transpose-reduce": every lane enters with NV partial sums v[0..NV) (NV = 32 * PER) of the
// same NV outputs and leaves with the complete sums of outputs lane * PER + i in v[i], i < PER.  Each of
// the five halving rounds exchanges the half a lane does not keep with its xor-partner.
template <int NV>
__device__ __forceinline__ void pg_reduce(float* v, int lane) {
#pragma unroll
    for (int s = 0; s < 5; ++s) {
        const int half = NV >> (s + 1);                      // values kept after this round
        const int off = 16 >> s;
        const bool upper = (lane & off) != 0;
#pragma unroll
        for (int i = 0; i < half; ++i) {
            const float a = v[i], b = v[i + half];
            const float send = upper ? a : b;
            const float keep = upper ? b : a;
            v[i] = keep + __shfl_xor_sync(0xffffffffu, send, off);
        }
    }
}

// ---------------------------------------------------------------------------
// forward: CTA (bx, by) owns hidden units 32 bx .. 32 bx + 31 (their 96 gate rows of W_hh, 192 KiB, stay in
// shared memory for all T steps) and the environment groups by, by + gridDim.y, ... of 8 environments.
// Warp w takes units 4 w .. 4 w + 3; its lanes split the reduction (lane l: k = 4 l .. 4 l + 3 of every
// 128).  All 8 warps read the same 8 rows of hm_t, so the CTA pulls 16 KiB per step from L2 and the other
// 7 reads hit L1 (every address is read only after its one write of this launch and was never cached
// before, so the non-coherent L1 cannot serve stale data).  The 96 partial sums of a warp are
// transpose-reduced so that lane 4 r + j ends up with the three gate pre-activations of (environment r,
// unit j) and applies the gate math.  One grid barrier per step.
// ---------------------------------------------------------------------------
__global__ void __launch_bounds__(PG_NT)
gru_persistent_fwd_kernel(const GruDesc d, const float* __restrict__ Whh, const float* __restrict__ bhh) {
    extern __shared__ __align__(16) float pg_smem[];
    const int H = d.H;
    unsigned* ctr = &pg_counters[0][blockIdx.y];
    unsigned epoch = 0;

    const int tid = threadIdx.x, lane = tid & 31, warp = tid >> 5;
    const int ub = blockIdx.x * PG_UB;                              // first unit of this CTA
    // Ws[(w * 4 + j) * 3 + g][H] = W_hh[g * H + ub + 4 w + j][:]
    for (int e = tid; e < 3 * PG_UB * H; e += PG_NT) {
        const int c = e / H, k = e - c * H;
        const int u = c / 3, g = c - u * 3;
        pg_smem[e] = Whh[(long)(g * H + ub + u) * H + k];
    }
    const float* Ws = pg_smem + (long)warp * 12 * H;
    const int rl = lane >> 2, jl = lane & 3;                        // this lane's (environment, unit) after the reduce
    const int ku = ub + PG_UW * warp + jl;
    const float b_r = bhh[ku], b_z = bhh[H + ku], b_n = bhh[2 * H + ku];
    const int ngrp = (d.N + PG_RW - 1) / PG_RW;

    // hm_0 = h0 * mask_0 for this CTA's units and environments
    for (int grp = blockIdx.y; grp < ngrp; grp += gridDim.y) {
        const int my = grp * PG_RW + rl;
        if (my < d.N) d.hm[(long)my * H + ku] = d.h0[(long)my * H + ku] * pg_mask(d, 0, my);
    }
    pg_group_sync(ctr, ++epoch * gridDim.x);

    for (int t = 0; t < d.T; ++t) {
        float* hm_t = d.hm + (long)t * d.N * H;
        for (int grp = blockIdx.y; grp < ngrp; grp += gridDim.y) {
            const int r0 = grp * PG_RW;
            const int my = r0 + rl;                                 // the environment this lane finishes
            const bool live = my < d.N;
            const long row = (long)t * d.N + my;
            float gi_r = 0.f, gi_z = 0.f, gi_n = 0.f, hprev = 0.f, mnext = 0.f;
            if (live) {                                             // issued early, consumed after the matmul
                const float* gi = d.gi + row * 3 * H;
                gi_r = __ldg(gi + ku); gi_z = __ldg(gi + H + ku); gi_n = __ldg(gi + 2 * H + ku);
                hprev = hm_t[(long)my * H + ku];
                mnext = (t + 1 < d.T) ? pg_mask(d, t + 1, my) : 0.f;
            }
            float acc[PG_RW * 12];
#pragma unroll
            for (int i = 0; i < PG_RW * 12; ++i) acc[i] = 0.f;
            // two register sets of 8 float4: the loads of chunk i + 1 are in flight while chunk i is multiplied
            float4 ha[PG_RW], hb[PG_RW];
            auto load_h = [&](float4* h, int k0) {
#pragma unroll
                for (int r = 0; r < PG_RW; ++r) {
                    const int n = min(r0 + r, d.N - 1);
                    h[r] = *reinterpret_cast<const float4*>(hm_t + (long)n * H + k0 + 4 * lane);
                }
            };
            auto mul_h = [&](const float4* h, int k0) {
#pragma unroll
                for (int c = 0; c < 12; ++c) {
                    const float4 w = *reinterpret_cast<const float4*>(Ws + c * H + k0 + 4 * lane);
#pragma unroll
                    for (int r = 0; r < PG_RW; ++r) acc[r * 12 + c] = pg_dot4(h[r], w, acc[r * 12 + c]);
                }
            };
            load_h(ha, 0);
            for (int k0 = 0; k0 < H; k0 += 256) {
                if (k0 + 128 < H) load_h(hb, k0 + 128);
                mul_h(ha, k0);
                if (k0 + 256 < H) load_h(ha, k0 + 256);
                if (k0 + 128 < H) mul_h(hb, k0 + 128);
            }
            pg_reduce<PG_RW * 12>(acc, lane);                      // acc[0..2] = (r, z, n) sums of (rl, jl)
            if (live) {
                const float ghr = acc[0] + b_r, ghz = acc[1] + b_z, ghn = acc[2] + b_n;
                const float rr = pg_sigmoid(gi_r + ghr);
                const float zz = pg_sigmoid(gi_z + ghz);
                const float nn = tanhf(gi_n + rr * ghn);
                const float h = (1.f - zz) * nn + zz * hprev;
                d.out[row * H + ku] = h;
                if (d.r) {
                    d.r[row * H + ku] = rr; d.z[row * H + ku] = zz; d.n[row * H + ku] = nn;
                    d.gh[row * 3 * H + 2 * H + ku] = ghn;            // the only part of gh the backward pass reads
                }
                if (t + 1 < d.T) d.hm[(row + d.N) * H + ku] = h * mnext;
            }
        }
        if (t + 1 < d.T) pg_group_sync(ctr, ++epoch * gridDim.x);
    }
}

// ---------------------------------------------------------------------------
// forward, cluster variant: the H / 32 CTAs of ONE environment group are a thread-block cluster.  The masked state
// hm_t of the group's 8 environments (8 x H floats) lives in every CTA's shared memory, double-buffered; after the
// gate math each CTA pushes its 8 x 32 slice of hm_{t+1} into all peers through distributed shared memory
// (st.shared::cluster, 16-byte stores) and the cluster meets at the hardware cluster barrier.  No global-memory
// round trip and no software barrier remain on the per-step critical path; clusters (environment groups) are
// completely independent, so there is no co-residency requirement beyond the cluster itself.
// Correct (same tests as the default kernel) but NOT the default: see gru_cluster_forward() for the measurement.
// ---------------------------------------------------------------------------
__device__ __forceinline__ uint32_t pg_smem_u32(const void* p) { return (uint32_t)__cvta_generic_to_shared(p); }
__device__ __forceinline__ void pg_st_cluster_v4(uint32_t local_addr, uint32_t peer, const float4 v) {
    uint32_t remote;
    asm volatile("mapa.shared::cluster.u32 %0, %1, %2;" : "=r"(remote) : "r"(local_addr), "r"(peer));
    asm volatile("st.shared::cluster.v4.f32 [%0], {%1, %2, %3, %4};" ::"r"(remote), "f"(v.x), "f"(v.y), "f"(v.z), "f"(v.w)
                 : "memory");
}
__device__ __forceinline__ void pg_cluster_sync() {
    asm volatile("barrier.cluster.arrive.release.aligned;" ::: "memory");
    asm volatile("barrier.cluster.wait.acquire.aligned;" ::: "memory");
}

__global__ void __launch_bounds__(PG_NT)
gru_cluster_fwd_kernel(const GruDesc d, const float* __restrict__ Whh, const float* __restrict__ bhh) {
    extern __shared__ __align__(16) float pg_smem[];
    const int H = d.H;
    float* hbuf = pg_smem + 3 * PG_UB * H;                          // [2][PG_RW][H] masked state of the group
    float* tile = hbuf + 2 * PG_RW * H;                             // [PG_RW][PG_UB] this CTA's slice, staged for the push
    const int tid = threadIdx.x, lane = tid & 31, warp = tid >> 5;
    const int ncta = gridDim.x;                                     // cluster size = unit groups
    const int ub = blockIdx.x * PG_UB;
    for (int e = tid; e < 3 * PG_UB * H; e += PG_NT) {
        const int c = e / H, k = e - c * H;
        const int u = c / 3, g = c - u * 3;
        pg_smem[e] = Whh[(long)(g * H + ub + u) * H + k];
    }
    const float* Ws = pg_smem + (long)warp * 12 * H;
    const int rl = lane >> 2, jl = lane & 3;
    const int ku = ub + PG_UW * warp + jl;
    const float b_r = bhh[ku], b_z = bhh[H + ku], b_n = bhh[2 * H + ku];
    const int r0 = blockIdx.y * PG_RW;
    const int my = r0 + rl;
    const bool live = my < d.N;

    // pushes tile[PG_RW][PG_UB] into hbuf[buf][:, ub .. ub + 31] of every CTA of the cluster (incl. this one)
    auto push = [&](int buf) {
        __syncthreads();                                            // tile complete
        // 64 float4 per peer; warp w serves peers w, w + 8, ...: lanes 0..31 take two float4 each
        for (int peer = warp; peer < ncta; peer += PG_NW) {
#pragma unroll
            for (int q = 0; q < 2; ++q) {
                const int f4 = lane + 32 * q;                       // 0..63: row = f4 / 8, 4-unit chunk = f4 % 8
                const int r = f4 >> 3, c4 = f4 & 7;
                const float4 v = *reinterpret_cast<const float4*>(tile + r * PG_UB + c4 * 4);
                pg_st_cluster_v4(pg_smem_u32(hbuf + ((long)buf * PG_RW + r) * H + ub + c4 * 4), (uint32_t)peer, v);
            }
        }
    };

    // hm_0 = h0 * mask_0
    {
        float v = 0.f;
        if (live) {
            v = d.h0[(long)my * H + ku] * pg_mask(d, 0, my);
            d.hm[(long)my * H + ku] = v;
        }
        tile[rl * PG_UB + PG_UW * warp + jl] = v;
        push(0);
        pg_cluster_sync();
    }

    for (int t = 0; t < d.T; ++t) {
        const float* hcur = hbuf + (long)(t & 1) * PG_RW * H;
        const long row = (long)t * d.N + my;
        float gi_r = 0.f, gi_z = 0.f, gi_n = 0.f, mnext = 0.f;
        if (live) {
            const float* gi = d.gi + row * 3 * H;
            gi_r = __ldg(gi + ku); gi_z = __ldg(gi + H + ku); gi_n = __ldg(gi + 2 * H + ku);
            mnext = (t + 1 < d.T) ? pg_mask(d, t + 1, my) : 0.f;
        }
        const float hprev = hcur[rl * H + ku];
        float acc[PG_RW * 12];
#pragma unroll
        for (int i = 0; i < PG_RW * 12; ++i) acc[i] = 0.f;
#pragma unroll 1
        for (int k0 = 0; k0 < H; k0 += 128) {
            float4 h[PG_RW];
#pragma unroll
            for (int r = 0; r < PG_RW; ++r) h[r] = *reinterpret_cast<const float4*>(hcur + r * H + k0 + 4 * lane);
#pragma unroll
            for (int c = 0; c < 12; ++c) {
                const float4 w = *reinterpret_cast<const float4*>(Ws + c * H + k0 + 4 * lane);
#pragma unroll
                for (int r = 0; r < PG_RW; ++r) acc[r * 12 + c] = pg_dot4(h[r], w, acc[r * 12 + c]);
            }
        }
        pg_reduce<PG_RW * 12>(acc, lane);
        float hm_next = 0.f;
        if (live) {
            const float ghr = acc[0] + b_r, ghz = acc[1] + b_z, ghn = acc[2] + b_n;
            const float rr = pg_sigmoid(gi_r + ghr);
            const float zz = pg_sigmoid(gi_z + ghz);
            const float nn = tanhf(gi_n + rr * ghn);
            const float hn = (1.f - zz) * nn + zz * hprev;
            d.out[row * H + ku] = hn;
            if (d.r) {
                d.r[row * H + ku] = rr; d.z[row * H + ku] = zz; d.n[row * H + ku] = nn;
                d.gh[row * 3 * H + 2 * H + ku] = ghn;
            }
            hm_next = hn * mnext;
            if (t + 1 < d.T) d.hm[(row + d.N) * H + ku] = hm_next;      // kept for the backward pass
        }
        if (t + 1 < d.T) {
            tile[rl * PG_UB + PG_UW * warp + jl] = hm_next;
            push((t + 1) & 1);
            pg_cluster_sync();
        }
    }
    pg_cluster_sync();                                               // no CTA exits while peers may still push into it
}

// ---------------------------------------------------------------------------
// backward through time: same ownership; the 32 columns of W_hh (all 3H gate rows, 192 KiB) of the CTA's
// units stay in shared memory.  Phase A (element-wise gate gradients of its units and environments, needs
// the carry it computed itself), grid barrier, phase B (carry = dh z + dgh_t W_hh restricted to its units,
// lanes split the 3H-long reduction, 32 partial sums transpose-reduced to one per lane).
// ---------------------------------------------------------------------------
__global__ void __launch_bounds__(PG_NT)
gru_persistent_bwd_kernel(const GruDesc d, const float* __restrict__ Whh, float* __restrict__ carry_ws) {
    extern __shared__ __align__(16) float pg_smem[];
    const int H = d.H, H3 = 3 * d.H;
    unsigned* ctr = &pg_counters[1][blockIdx.y];
    unsigned epoch = 0;

    const int tid = threadIdx.x, lane = tid & 31, warp = tid >> 5;
    const int ub = blockIdx.x * PG_UB;
    const int nit = H3 / 128;
    // Wc[w][(c / 128) * 4 + c % 4][lane = (c % 128) / 4][j] = W_hh[c][ub + 4 w + j]: one conflict-free float4 per lane
    for (int e = tid; e < H3 * PG_UB; e += PG_NT) {
        const int c = e / PG_UB, u = e - c * PG_UB;
        const int w = u >> 2, j = u & 3;
        const int it = c >> 7, l = (c & 127) >> 2, q = c & 3;
        pg_smem[((((long)w * nit + it) * 4 + q) * 32 + l) * 4 + j] = Whh[(long)c * H + ub + u];
    }
    __syncthreads();
    const float4* Wc = reinterpret_cast<const float4*>(pg_smem) + (long)warp * nit * 4 * 32;
    const int rl = lane >> 2, jl = lane & 3;
    const int ku = ub + PG_UW * warp + jl;
    const int ngrp = (d.N + PG_RW - 1) / PG_RW;
    // carry of (environment, unit): a register for the CTA's first environment group, the [N, H] workspace
    // for further groups
    float carry0 = 0.f;

    struct PhaseA { float dout, r, z, n, hm, ghn, mnext; };
    auto load_a = [&](int t, int my) {
        PhaseA a;
        const long row = (long)t * d.N + my;
        a.dout = d.dout[row * H + ku];
        a.r = d.r[row * H + ku]; a.z = d.z[row * H + ku]; a.n = d.n[row * H + ku];
        a.hm = d.hm[row * H + ku];
        a.ghn = d.gh[row * 3 * H + 2 * H + ku];
        a.mnext = (t + 1 < d.T) ? pg_mask(d, t + 1, my) : 0.f;
        return a;
    };
    const int my0 = blockIdx.y * PG_RW + rl;                 // this lane's environment in the first group
    PhaseA pre = {0.f, 0.f, 0.f, 0.f, 0.f, 0.f, 0.f};
    if (my0 < d.N) pre = load_a(d.T - 1, my0);

    for (int t = d.T - 1; t >= 0; --t) {
        // ---- phase A ----
        for (int grp = blockIdx.y; grp < ngrp; grp += gridDim.y) {
            const int my = grp * PG_RW + rl;
            if (my >= d.N) continue;
            const bool first = grp == (int)blockIdx.y;
            const long row = (long)t * d.N + my;
            const PhaseA a = first ? pre : load_a(t, my);      // the first group was prefetched before the last barrier
            float dh = a.dout;
            if (t + 1 < d.T) {
                const float c = first ? carry0 : carry_ws[(long)my * H + ku];
                dh += c * a.mnext;
            }
            const float dn_pre = dh * (1.f - a.z) * (1.f - a.n * a.n);
            const float dz_pre = dh * (a.hm - a.n) * a.z * (1.f - a.z);
            const float dr_pre = dn_pre * a.ghn * a.r * (1.f - a.r);
            float* dgi = d.dgi + row * 3 * H;
            float* dgh = d.dgh + row * 3 * H;
            dgi[ku] = dr_pre; dgi[H + ku] = dz_pre; dgi[2 * H + ku] = dn_pre;
            dgh[ku] = dr_pre; dgh[H + ku] = dz_pre; dgh[2 * H + ku] = dn_pre * a.r;
            const float dhz = dh * a.z;
            if (first) carry0 = dhz;
            else carry_ws[(long)my * H + ku] = dhz;
        }
        if (t == 0) break;                 // the gradient w.r.t. the initial state is not needed
        pg_group_sync(ctr, ++epoch * gridDim.x);
        if (my0 < d.N) pre = load_a(t - 1, my0);               // in flight during phase B
        // ---- phase B ----
        const float* dgh_t = d.dgh + (long)t * d.N * H3;
        for (int grp = blockIdx.y; grp < ngrp; grp += gridDim.y) {
            const int r0 = grp * PG_RW;
            float acc[PG_RW * PG_UW];
#pragma unroll
            for (int i = 0; i < PG_RW * PG_UW; ++i) acc[i] = 0.f;
            float4 ga[PG_RW], gb[PG_RW];
            auto load_g = [&](float4* g, int it) {
#pragma unroll
                for (int r = 0; r < PG_RW; ++r) {
                    const int n = min(r0 + r, d.N - 1);
                    g[r] = *reinterpret_cast<const float4*>(dgh_t + (long)n * H3 + it * 128 + 4 * lane);
                }
            };
            auto mul_g = [&](const float4* g, int it) {
                const float4* wp = Wc + (it * 4) * 32 + lane;
                const float4 w0 = wp[0], w1 = wp[32], w2 = wp[64], w3 = wp[96];
#pragma unroll
                for (int r = 0; r < PG_RW; ++r) {
                    float* a = acc + r * PG_UW;
                    a[0] = fmaf(g[r].x, w0.x, a[0]); a[1] = fmaf(g[r].x, w0.y, a[1]);
                    a[2] = fmaf(g[r].x, w0.z, a[2]); a[3] = fmaf(g[r].x, w0.w, a[3]);
                    a[0] = fmaf(g[r].y, w1.x, a[0]); a[1] = fmaf(g[r].y, w1.y, a[1]);
                    a[2] = fmaf(g[r].y, w1.z, a[2]); a[3] = fmaf(g[r].y, w1.w, a[3]);
                    a[0] = fmaf(g[r].z, w2.x, a[0]); a[1] = fmaf(g[r].z, w2.y, a[1]);
                    a[2] = fmaf(g[r].z, w2.z, a[2]); a[3] = fmaf(g[r].z, w2.w, a[3]);
                    a[0] = fmaf(g[r].w, w3.x, a[0]); a[1] = fmaf(g[r].w, w3.y, a[1]);
                    a[2] = fmaf(g[r].w, w3.z, a[2]); a[3] = fmaf(g[r].w, w3.w, a[3]);
                }
            };
            load_g(ga, 0);
            for (int it = 0; it < nit; it += 2) {
                if (it + 1 < nit) load_g(gb, it + 1);
                mul_g(ga, it);
                if (it + 2 < nit) load_g(ga, it + 2);
                if (it + 1 < nit) mul_g(gb, it + 1);
            }
            pg_reduce<PG_RW * PG_UW>(acc, lane);                   // acc[0] = sum of (rl, jl)
            const int my = r0 + rl;
            if (my < d.N) {
                if (grp == (int)blockIdx.y) carry0 += acc[0];
                else carry_ws[(long)my * H + ku] += acc[0];
            }
        }
    }
}

// ---------------------------------------------------------------------------
// Tensor-core variant of both recurrence kernels.  The per-step product is tiny ([8 environments] x [96 gate rows]
// x [H] per CTA) and its operands must stay where they are (W slice in shared memory, state in L2 / L1), so the
// tcgen05 path (M >= 64 rows per CTA, both operands through shared-memory descriptors) does not apply; warp-level
// mma.sync.m16n8k8 TF32 does: M = gate rows (forward) / hidden units (backward), N = the 8 environments, K split
// over the 8 warps, 3xTF32 (hi / lo split by truncation, in registers) for fp32-level accuracy, partial tiles
// folded through shared memory.  It replaces 1536 FFMA + a 96-value shuffle transpose per warp and step by 144 MMAs.
// Ownership, buffers, barriers and the gate math are those of the SIMT kernels above; the element-wise owner of
// (environment r, unit u) is thread 32 r + u, so every global access is a coalesced 128-byte row segment.
// ---------------------------------------------------------------------------
__device__ __forceinline__ void pm_split(float x, uint32_t& hi, uint32_t& lo) {
    hi = __float_as_uint(x) & 0xffffe000u;
    lo = __float_as_uint(x - __uint_as_float(hi));
}
__device__ __forceinline__ void pm_mma(float* c, const uint32_t* a, uint32_t b0, uint32_t b1) {
    asm volatile("mma.sync.aligned.m16n8k8.row.col.f32.tf32.tf32.f32 {%0, %1, %2, %3}, {%4, %5, %6, %7}, {%8, %9}, {%0, %1, %2, %3};"
                 : "+f"(c[0]), "+f"(c[1]), "+f"(c[2]), "+f"(c[3])
                 : "r"(a[0]), "r"(a[1]), "r"(a[2]), "r"(a[3]), "r"(b0), "r"(b1));
}
// one K = 8 step of MT m-tiles: A = W rows [mt * 16 + g (+ 8)][k0 + t (+ 4)] (pitch P), B = x[g][k0 + t (+ 4)]
template <int MT>
__device__ __forceinline__ void pm_kstep(float (*acc)[4], const float* __restrict__ Wsm, int P, int k0, float x0, float x1,
                                         int g, int t) {
    uint32_t bh0, bl0, bh1, bl1;
    pm_split(x0, bh0, bl0);
    pm_split(x1, bh1, bl1);
    // all A fragments of the k-step first, then the MMAs product-major: consecutive MMAs hit DIFFERENT accumulators,
    // so the three products of one tile do not wait for each other's latency
    uint32_t ah[MT][4], al[MT][4];
#pragma unroll
    for (int mt = 0; mt < MT; ++mt) {
        const float* w = Wsm + (long)(mt * 16 + g) * P + k0 + t;
        pm_split(w[0], ah[mt][0], al[mt][0]);
        pm_split(w[8 * P], ah[mt][1], al[mt][1]);
        pm_split(w[4], ah[mt][2], al[mt][2]);
        pm_split(w[8 * P + 4], ah[mt][3], al[mt][3]);
    }
#pragma unroll
    for (int mt = 0; mt < MT; ++mt) pm_mma(acc[mt], al[mt], bh0, bh1);
#pragma unroll
    for (int mt = 0; mt < MT; ++mt) pm_mma(acc[mt], ah[mt], bl0, bl1);
#pragma unroll
    for (int mt = 0; mt < MT; ++mt) pm_mma(acc[mt], ah[mt], bh0, bh1);
}

__global__ void __launch_bounds__(PG_NT)
gru_mma_fwd_kernel(const GruDesc d, const float* __restrict__ Whh, const float* __restrict__ bhh) {
    extern __shared__ __align__(16) float pg_smem[];
    const int H = d.H, P = H + 4;
    float* Ws = pg_smem;                                            // [96 gate rows: (u, g) -> u * 3 + g][P]
    float4* red = reinterpret_cast<float4*>(pg_smem + 3 * PG_UB * P);   // [8 warps][6 tiles][32 lanes]
    unsigned* ctr = &pg_counters[0][blockIdx.y];
    unsigned epoch = 0;
    const int tid = threadIdx.x, lane = tid & 31, warp = tid >> 5;
    const int g = lane >> 2, tq = lane & 3;
    const int ub = blockIdx.x * PG_UB;
    for (int e = tid; e < 3 * PG_UB * H; e += PG_NT) {
        const int c = e / H, k = e - c * H;
        const int u = c / 3, gg = c - u * 3;
        Ws[(long)c * P + k] = Whh[(long)(gg * H + ub + u) * H + k];
    }
    const int ro = warp, uo = lane;                                 // element-wise owner: (environment ro, unit uo)
    const int ku = ub + uo;
    const float b_r = bhh[ku], b_z = bhh[H + ku], b_n = bhh[2 * H + ku];
    const int ngrp = (d.N + PG_RW - 1) / PG_RW;
    const int ksteps = H / 8 / PG_NW;                               // K = 8 steps per warp

    for (int grp = blockIdx.y; grp < ngrp; grp += gridDim.y) {
        const int my = grp * PG_RW + ro;
        if (my < d.N) d.hm[(long)my * H + ku] = d.h0[(long)my * H + ku] * pg_mask(d, 0, my);
    }
    pg_group_sync(ctr, ++epoch * gridDim.x);

    // single environment group per CTA (the common case): the next step's gate inputs are prefetched while the CTA
    // waits at the barrier, and only the state store precedes the barrier arrival
    const bool single = ngrp <= (int)gridDim.y;
    struct GateIn { float r, z, n, mnext; };
    auto load_gi = [&](int t, int my) {
        GateIn q = {0.f, 0.f, 0.f, 0.f};
        if (my < d.N) {
            const float* gi = d.gi + ((long)t * d.N + my) * 3 * H;
            q.r = __ldg(gi + ku); q.z = __ldg(gi + H + ku); q.n = __ldg(gi + 2 * H + ku);
            q.mnext = (t + 1 < d.T) ? pg_mask(d, t + 1, my) : 0.f;
        }
        return q;
    };
    GateIn pre = load_gi(0, blockIdx.y * PG_RW + ro);

    for (int t = 0; t < d.T; ++t) {
        const float* hm_t = d.hm + (long)t * d.N * H;
        bool arrived = false;
        for (int grp = blockIdx.y; grp < ngrp; grp += gridDim.y) {
            const int r0 = grp * PG_RW;
            const int my = r0 + ro;
            const bool live = my < d.N;
            const long row = (long)t * d.N + my;
            const GateIn q = single ? pre : load_gi(t, my);
            const float hprev = live ? hm_t[(long)my * H + ku] : 0.f;
            float acc[6][4];
#pragma unroll
            for (int i = 0; i < 6; ++i) { acc[i][0] = acc[i][1] = acc[i][2] = acc[i][3] = 0.f; }
            const float* hrow = hm_t + (long)min(r0 + g, d.N - 1) * H + warp * ksteps * 8 + tq;
            // operands of all this warp's k-steps first (independent loads), then the MMAs
            float x0[8], x1[8];
#pragma unroll
            for (int ks = 0; ks < 8; ++ks)
                if (ks < ksteps) { x0[ks] = hrow[ks * 8]; x1[ks] = hrow[ks * 8 + 4]; }
#pragma unroll
            for (int ks = 0; ks < 8; ++ks)
                if (ks < ksteps) pm_kstep<6>(acc, Ws, P, (warp * ksteps + ks) * 8, x0[ks], x1[ks], g, tq);
            __syncthreads();                                         // previous readers of `red` are done
#pragma unroll
            for (int mt = 0; mt < 6; ++mt)
                red[(warp * 6 + mt) * 32 + lane] = make_float4(acc[mt][0], acc[mt][1], acc[mt][2], acc[mt][3]);
            __syncthreads();
            float gsum[3];
#pragma unroll
            for (int gg = 0; gg < 3; ++gg) {
                const int col = uo * 3 + gg, mt = col >> 4, i = col & 15;
                const int ln = (i & 7) * 4 + (ro >> 1), rg = (i >> 3) * 2 + (ro & 1);
                float v[PG_NW];
#pragma unroll
                for (int w = 0; w < PG_NW; ++w) v[w] = reinterpret_cast<const float*>(red + (w * 6 + mt) * 32 + ln)[rg];
                gsum[gg] = ((v[0] + v[1]) + (v[2] + v[3])) + ((v[4] + v[5]) + (v[6] + v[7]));
            }
            const float ghr = gsum[0] + b_r, ghz = gsum[1] + b_z, ghn = gsum[2] + b_n;
            const float rr = pg_sigmoid(q.r + ghr);
            const float zz = pg_sigmoid(q.z + ghz);
            const float nn = tanhf(q.n + rr * ghn);
            const float h = (1.f - zz) * nn + zz * hprev;
            if (live && t + 1 < d.T) d.hm[(row + d.N) * H + ku] = h * q.mnext;      // what the other CTAs wait for
            if (single && t + 1 < d.T) {
                pg_group_arrive(ctr);
                arrived = true;
                pre = load_gi(t + 1, my);                                           // in flight during the barrier wait
            }
            if (live) {
                d.out[row * H + ku] = h;
                if (d.r) {
                    d.r[row * H + ku] = rr; d.z[row * H + ku] = zz; d.n[row * H + ku] = nn;
                    d.gh[row * 3 * H + 2 * H + ku] = ghn;
                }
            }
        }
        if (t + 1 < d.T) {
            if (!arrived) pg_group_arrive(ctr);
            pg_group_wait(ctr, ++epoch * gridDim.x);
        }
    }
}

__global__ void __launch_bounds__(PG_NT)
gru_mma_bwd_kernel(const GruDesc d, const float* __restrict__ Whh, float* __restrict__ carry_ws) {
    extern __shared__ __align__(16) float pg_smem[];
    const int H = d.H, H3 = 3 * d.H, P = H3 + 4;
    float* Wt = pg_smem;                                            // [32 units][P]: Wt[u][c] = W_hh[c][ub + u]
    float4* red = reinterpret_cast<float4*>(pg_smem + PG_UB * P);   // [8 warps][2 tiles][32 lanes]
    unsigned* ctr = &pg_counters[1][blockIdx.y];
    unsigned epoch = 0;
    const int tid = threadIdx.x, lane = tid & 31, warp = tid >> 5;
    const int g = lane >> 2, tq = lane & 3;
    const int ub = blockIdx.x * PG_UB;
    for (int e = tid; e < H3 * PG_UB; e += PG_NT) {
        const int c = e / PG_UB, u = e - c * PG_UB;
        Wt[(long)u * P + c] = Whh[(long)c * H + ub + u];
    }
    __syncthreads();
    const int ro = warp, uo = lane;
    const int ku = ub + uo;
    const int ngrp = (d.N + PG_RW - 1) / PG_RW;
    const int ksteps = H3 / 8 / PG_NW;                              // 24 for H = 512
    float carry0 = 0.f;

    struct PhaseA { float dout, r, z, n, hm, ghn, mnext; };
    auto load_a = [&](int t, int my) {
        PhaseA a;
        const long row = (long)t * d.N + my;
        a.dout = d.dout[row * H + ku];
        a.r = d.r[row * H + ku]; a.z = d.z[row * H + ku]; a.n = d.n[row * H + ku];
        a.hm = d.hm[row * H + ku];
        a.ghn = d.gh[row * 3 * H + 2 * H + ku];
        a.mnext = (t + 1 < d.T) ? pg_mask(d, t + 1, my) : 0.f;
        return a;
    };
    const int my0 = blockIdx.y * PG_RW + ro;
    PhaseA pre = {0.f, 0.f, 0.f, 0.f, 0.f, 0.f, 0.f};
    if (my0 < d.N) pre = load_a(d.T - 1, my0);
    float dgi_keep[3] = {0.f, 0.f, 0.f};
    long keep_row = -1;

    for (int t = d.T - 1; t >= 0; --t) {
        // ---- phase A ----
        for (int grp = blockIdx.y; grp < ngrp; grp += gridDim.y) {
            const int my = grp * PG_RW + ro;
            if (my >= d.N) continue;
            const bool first = grp == (int)blockIdx.y;
            const long row = (long)t * d.N + my;
            const PhaseA a = first ? pre : load_a(t, my);
            float dh = a.dout;
            if (t + 1 < d.T) {
                const float c = first ? carry0 : carry_ws[(long)my * H + ku];
                dh += c * a.mnext;
            }
            const float dn_pre = dh * (1.f - a.z) * (1.f - a.n * a.n);
            const float dz_pre = dh * (a.hm - a.n) * a.z * (1.f - a.z);
            const float dr_pre = dn_pre * a.ghn * a.r * (1.f - a.r);
            float* dgi = d.dgi + row * 3 * H;
            float* dgh = d.dgh + row * 3 * H;
            dgh[ku] = dr_pre; dgh[H + ku] = dz_pre; dgh[2 * H + ku] = dn_pre * a.r;     // what the other CTAs wait for
            if (first) { dgi_keep[0] = dr_pre; dgi_keep[1] = dz_pre; dgi_keep[2] = dn_pre; keep_row = row; }
            else { dgi[ku] = dr_pre; dgi[H + ku] = dz_pre; dgi[2 * H + ku] = dn_pre; }
            const float dhz = dh * a.z;
            if (first) carry0 = dhz;
            else carry_ws[(long)my * H + ku] = dhz;
        }
        if (t > 0) pg_group_arrive(ctr);
        if (keep_row >= 0) {                                       // not needed by anyone inside this kernel
            float* dgi = d.dgi + keep_row * 3 * H;
            dgi[ku] = dgi_keep[0]; dgi[H + ku] = dgi_keep[1]; dgi[2 * H + ku] = dgi_keep[2];
            keep_row = -1;
        }
        if (t == 0) break;
        if (my0 < d.N) pre = load_a(t - 1, my0);                    // in flight during the barrier wait and phase B
        pg_group_wait(ctr, ++epoch * gridDim.x);
        // ---- phase B: carry[r][u] += sum_c dgh_t[r][c] W_hh[c][u] ----
        const float* dgh_t = d.dgh + (long)t * d.N * H3;
        for (int grp = blockIdx.y; grp < ngrp; grp += gridDim.y) {
            const int r0 = grp * PG_RW;
            float acc[2][4];
#pragma unroll
            for (int i = 0; i < 2; ++i) { acc[i][0] = acc[i][1] = acc[i][2] = acc[i][3] = 0.f; }
            const float* grow = dgh_t + (long)min(r0 + g, d.N - 1) * H3 + warp * ksteps * 8 + tq;
            for (int ks0 = 0; ks0 < ksteps; ks0 += 8) {
                float x0[8], x1[8];
#pragma unroll
                for (int ks = 0; ks < 8; ++ks)
                    if (ks0 + ks < ksteps) { x0[ks] = grow[(ks0 + ks) * 8]; x1[ks] = grow[(ks0 + ks) * 8 + 4]; }
#pragma unroll
                for (int ks = 0; ks < 8; ++ks)
                    if (ks0 + ks < ksteps) pm_kstep<2>(acc, Wt, P, (warp * ksteps + ks0 + ks) * 8, x0[ks], x1[ks], g, tq);
            }
            __syncthreads();
#pragma unroll
            for (int mt = 0; mt < 2; ++mt)
                red[(warp * 2 + mt) * 32 + lane] = make_float4(acc[mt][0], acc[mt][1], acc[mt][2], acc[mt][3]);
            __syncthreads();
            const int mt = uo >> 4, i = uo & 15;
            const int ln = (i & 7) * 4 + (ro >> 1), rg = (i >> 3) * 2 + (ro & 1);
            float v[PG_NW];
#pragma unroll
            for (int w = 0; w < PG_NW; ++w) v[w] = reinterpret_cast<const float*>(red + (w * 2 + mt) * 32 + ln)[rg];
            const float sum = ((v[0] + v[1]) + (v[2] + v[3])) + ((v[4] + v[5]) + (v[6] + v[7]));
            const int my = r0 + ro;
            if (my < d.N) {
                if (grp == (int)blockIdx.y) carry0 += sum;
                else carry_ws[(long)my * H + ku] += sum;
            }
        }
    }
}

static size_t pm_smem_fwd(int H) { return (size_t)(3 * PG_UB * (H + 4)) * sizeof(float) + (size_t)PG_NW * 6 * 32 * sizeof(float4); }
static size_t pm_smem_bwd(int H) { return (size_t)(PG_UB * (3 * H + 4)) * sizeof(float) + (size_t)PG_NW * 2 * 32 * sizeof(float4); }
static bool pm_enabled(int H) {
    const int mode = gru_mode() >= 1 ? 1 : 0;                       // a2cb_set_gru_mode(0) selects the SIMT kernels
    return mode == 1 && H % 64 == 0 && H / 8 / PG_NW <= 8 && pm_smem_fwd(H) <= 227 * 1024 && pm_smem_bwd(H) <= 227 * 1024;
}

static size_t pg_smem_bytes(int H) { return (size_t)(3 * PG_UB * H) * sizeof(float); }

static int pg_sms() {
    static int sms = 0;
    if (!sms) {
        int dev = 0;
        cudaGetDevice(&dev);
        cudaDeviceGetAttribute(&sms, cudaDevAttrMultiProcessorCount, dev);
    }
    return sms;
}

// Cooperative launch: every CTA must be co-resident (one per SM at this shared-memory footprint).
bool gru_persistent_supported(int N, int H) {
    if (gru_tc_supported(N, H)) return true;
    if (N <= 0 || H % 128 != 0 || H % PG_UB != 0) return false;
    const size_t smem = pg_smem_bytes(H);
    if (smem > 220 * 1024) return false;
    if (cudaFuncSetAttribute(gru_persistent_fwd_kernel, cudaFuncAttributeMaxDynamicSharedMemorySize, 220 * 1024) != cudaSuccess ||
        cudaFuncSetAttribute(gru_persistent_bwd_kernel, cudaFuncAttributeMaxDynamicSharedMemorySize, 220 * 1024) != cudaSuccess) {
        cudaGetLastError();
        return false;
    }
    int pf = 0, pb = 0;
    cudaOccupancyMaxActiveBlocksPerMultiprocessor(&pf, gru_persistent_fwd_kernel, PG_NT, smem);
    cudaOccupancyMaxActiveBlocksPerMultiprocessor(&pb, gru_persistent_bwd_kernel, PG_NT, smem);
    return (long)(H / PG_UB) <= (long)pg_sms() * (pf < pb ? pf : pb);
}

// cudaOccupancyMaxActiveClusters for the cluster forward kernel's footprint (1 CTA / SM): how many clusters of
// `cluster_size` CTAs the device can hold at once
int gru_cluster_probe(int cluster_size, int smem_bytes) {
    if (cudaFuncSetAttribute(gru_cluster_fwd_kernel, cudaFuncAttributeMaxDynamicSharedMemorySize, smem_bytes) != cudaSuccess ||
        cudaFuncSetAttribute(gru_cluster_fwd_kernel, cudaFuncAttributeNonPortableClusterSizeAllowed, 1) != cudaSuccess) {
        cudaGetLastError();
        return -1;
    }
    cudaLaunchConfig_t cfg = {};
    cfg.gridDim = dim3((unsigned)cluster_size, 64);
    cfg.blockDim = dim3(PG_NT);
    cfg.dynamicSmemBytes = (size_t)smem_bytes;
    cudaLaunchAttribute at[1];
    at[0].id = cudaLaunchAttributeClusterDimension;
    at[0].val.clusterDim.x = (unsigned)cluster_size; at[0].val.clusterDim.y = 1; at[0].val.clusterDim.z = 1;
    cfg.attrs = at;
    cfg.numAttrs = 1;
    int n = 0;
    if (cudaOccupancyMaxActiveClusters(&n, gru_cluster_fwd_kernel, &cfg) != cudaSuccess) { cudaGetLastError(); return -2; }
    return n;
}

int gru_persistent_variant(int N, int H) {
    if (gru_tc_supported(N, H)) return 2;
    if (!gru_persistent_supported(N, H)) return -1;
    return pm_enabled(H) ? 1 : 0;
}

// Cluster launch of the forward kernel when the H / 32 unit groups fit one cluster (H = 512 -> 16 CTAs, the
// non-portable maximum).  Returns 1 if launched, 0 if this shape / device cannot take it, < 0 on error.
static int gru_cluster_forward(const GruDesc& d, const float* Whh, const float* bhh, cudaStream_t st) {
    // Opt-in (A2CB_GRU_CLUSTER=1).  Measured on B200 at H = 512, 8 environment groups: 1.43 ms vs 0.82 ms for the
    // L2 / counter-barrier kernel -- the 8 clusters of 16 CTAs do not all become co-resident (GPCs of 16-20 SMs, one
    // die each), so the groups run in two waves; per wave the step time is 5.6 us, no better than the L2 exchange.
    static int mode = -1;
    if (mode < 0) { const char* e = getenv("A2CB_GRU_CLUSTER"); mode = (e && e[0] == '1') ? 1 : 0; }
    const int ug = d.H / PG_UB;
    if (!mode || ug < 2 || ug > 16 || (ug & (ug - 1)) != 0) return 0;
    const size_t smem = (size_t)(3 * PG_UB * d.H + 2 * PG_RW * d.H + PG_RW * PG_UB) * sizeof(float);
    if (smem > 227 * 1024) return 0;
    static bool attr = false;
    if (!attr) {
        if (cudaFuncSetAttribute(gru_cluster_fwd_kernel, cudaFuncAttributeMaxDynamicSharedMemorySize, 227 * 1024) != cudaSuccess ||
            cudaFuncSetAttribute(gru_cluster_fwd_kernel, cudaFuncAttributeNonPortableClusterSizeAllowed, 1) != cudaSuccess) {
            cudaGetLastError();
            return 0;
        }
        attr = true;
    }
    const int ngrp = (d.N + PG_RW - 1) / PG_RW;
    cudaLaunchConfig_t cfg = {};
    cfg.gridDim = dim3((unsigned)ug, (unsigned)ngrp);
    cfg.blockDim = dim3(PG_NT);
    cfg.dynamicSmemBytes = smem;
    cfg.stream = st;
    cudaLaunchAttribute at[1];
    at[0].id = cudaLaunchAttributeClusterDimension;
    at[0].val.clusterDim.x = (unsigned)ug; at[0].val.clusterDim.y = 1; at[0].val.clusterDim.z = 1;
    cfg.attrs = at;
    cfg.numAttrs = 1;
    int max_clusters = 0;
    if (cudaOccupancyMaxActiveClusters(&max_clusters, gru_cluster_fwd_kernel, &cfg) != cudaSuccess || max_clusters < 1) {
        cudaGetLastError();
        return 0;
    }
    cudaError_t e = cudaLaunchKernelEx(&cfg, gru_cluster_fwd_kernel, d, Whh, bhh);
    if (e != cudaSuccess) {
        set_error("gru_persistent: cluster launch failed: %s", cudaGetErrorString(e));
        cudaGetLastError();
        return A2CB_ERR_CUDA;
    }
    return check_launch("gru_cluster_fwd") == A2CB_OK ? 1 : A2CB_ERR_CUDA;
}

int gru_persistent(int backward, const GruDesc& d, const float* Whh, const float* bhh, cudaStream_t st) {
    if (gru_tc_supported(d.N, d.H)) return gru_tc(backward, d, Whh, bhh, st);      // tcgen05 kernels (gru_tc.cu)
    A2CB_REQUIRE(gru_persistent_supported(d.N, d.H), "gru_persistent: unsupported shape (N=%d H=%d)", d.N, d.H);
    A2CB_REQUIRE(Whh && d.hm && d.masks, "gru_persistent: null pointer");
    A2CB_REQUIRE(backward || (bhh && d.h0 && d.gi && d.out), "gru_persistent: forward pointers");
    A2CB_REQUIRE(!backward || (d.r && d.z && d.n && d.gh && d.dout && d.dgi && d.dgh), "gru_persistent: backward pointers");
    if (!backward) {
        const int rc = gru_cluster_forward(d, Whh, bhh, st);
        if (rc != 0) return rc > 0 ? A2CB_OK : rc;
    }
    const int ug = d.H / PG_UB;                                    // unit groups
    const int ngrp = (d.N + PG_RW - 1) / PG_RW;                    // environment groups
    int gy = pg_sms() / ug;
    gy = gy < 1 ? 1 : (gy > ngrp ? ngrp : gy);
    A2CB_REQUIRE(!backward || ngrp <= gy || d.dcarry, "gru_persistent: dcarry workspace [N, H] needed for N > %d", gy * PG_RW);
    A2CB_REQUIRE(gy <= 256, "gru_persistent: too many environment groups");
    const bool mma = pm_enabled(d.H);
    const size_t smem = mma ? (backward ? pm_smem_bwd(d.H) : pm_smem_fwd(d.H)) : pg_smem_bytes(d.H);
    if (mma) {
        static bool attr = false;
        if (!attr) {
            if (cudaFuncSetAttribute(gru_mma_fwd_kernel, cudaFuncAttributeMaxDynamicSharedMemorySize, 227 * 1024) != cudaSuccess ||
                cudaFuncSetAttribute(gru_mma_bwd_kernel, cudaFuncAttributeMaxDynamicSharedMemorySize, 227 * 1024) != cudaSuccess) {
                set_error("gru_persistent: shared memory attribute");
                cudaGetLastError();
                return A2CB_ERR_CUDA;
            }
            attr = true;
        }
    }
    dim3 grid((unsigned)ug, (unsigned)gy);
    {
        void* cptr = nullptr;
        cudaError_t e0 = cudaGetSymbolAddress(&cptr, pg_counters);
        if (e0 == cudaSuccess)
            e0 = cudaMemsetAsync(reinterpret_cast<unsigned*>(cptr) + (backward ? 256 : 0), 0, 256 * sizeof(unsigned), st);
        if (e0 != cudaSuccess) { set_error("gru_persistent: counter reset: %s", cudaGetErrorString(e0)); return A2CB_ERR_CUDA; }
    }
    GruDesc dd = d;
    float* ws = d.dcarry;
    void* args_f[] = {(void*)&dd, (void*)&Whh, (void*)&bhh};
    void* args_b[] = {(void*)&dd, (void*)&Whh, (void*)&ws};
    const void* kf = mma ? (const void*)gru_mma_fwd_kernel : (const void*)gru_persistent_fwd_kernel;
    const void* kb = mma ? (const void*)gru_mma_bwd_kernel : (const void*)gru_persistent_bwd_kernel;
    cudaError_t e = backward ? cudaLaunchCooperativeKernel(kb, grid, dim3(PG_NT), args_b, smem, st)
                             : cudaLaunchCooperativeKernel(kf, grid, dim3(PG_NT), args_f, smem, st);
    if (e != cudaSuccess) {
        set_error("gru_persistent: cooperative launch failed: %s", cudaGetErrorString(e));
        cudaGetLastError();
        return A2CB_ERR_CUDA;
    }
    return check_launch("gru_persistent");
}

}  // namespace a2cb
