// Persistent GRU recurrence: the whole [T, N] sequence in ONE cooperative launch per direction.
//
// The per-step formulation (gru.cu + one tiny GEMM per step) is launch- and latency-bound: at the
// C5 minibatch shape (T = 128, N = 64 environments, H = 512) the 2 x 128 sequential GEMMs of
// 0.1 GFLOP each cost ~18 us apiece although they are ~1 us of math.  Here every CTA owns KB = 8
// hidden units for one block of 64 environments and keeps ITS slice of W_hh resident in shared
// memory for all T steps (forward: the 3 x 8 gate rows, 48 KiB; backward: the 8 columns, 48 KiB);
// per step it streams the previous state (forward) / the gate gradients (backward) of its
// environment block from L2 through a padded shared-memory tile, does the matmul slice, applies the
// gate math in registers, and meets the other CTAs at one grid-wide barrier.  Episode boundaries are
// handled by the mask multiply, exactly as in gru.cu (reference model.py:111-166).
//
// Same buffers and the same saved quantities as the per-step path, so either path can feed the
// weight-gradient GEMMs and the tests compare both against the reference goldens.
#include <cooperative_groups.h>

#include "common.cuh"
#include "decls.cuh"
#include "gru.cuh"

namespace cg = cooperative_groups;

namespace a2cb {

constexpr int PG_KB = 8;        // hidden units per CTA
constexpr int PG_RB = 64;       // environments per CTA
constexpr int PG_NT = 256;      // threads: (row = tid % 64) x (unit pair = tid / 64)
constexpr int PG_KC = 64;       // reduction chunk staged in shared memory

__device__ __forceinline__ float pg_sigmoid(float x) { return 1.f / (1.f + expf(-x)); }

__device__ __forceinline__ float pg_mask(const GruDesc& d, int t, int n) {
    const long row = (long)t * d.N + n;
    const long src = d.idx ? (long)d.idx[row] : row;
    return d.masks[src];
}

// Loads src[(row0 + r) * ld + k0 + k] for r < 64, k < PG_KC into tile[k][r] (transposed, padded).
__device__ __forceinline__ void pg_load_tile(float (*tile)[PG_RB + 1], const float* __restrict__ src, long ld,
                                             int row0, int nrows, int k0, int tid) {
    // 64 rows x 16 float4 per chunk; consecutive lanes take consecutive float4 of one row (coalesced)
#pragma unroll
    for (int q = 0; q < (PG_RB * PG_KC / 4) / PG_NT; ++q) {
        const int s = tid + q * PG_NT;
        const int r = s / (PG_KC / 4), k4 = (s % (PG_KC / 4)) * 4;
        float4 v = make_float4(0.f, 0.f, 0.f, 0.f);
        if (r < nrows) v = *reinterpret_cast<const float4*>(src + (long)(row0 + r) * ld + k0 + k4);
        tile[k4 + 0][r] = v.x; tile[k4 + 1][r] = v.y; tile[k4 + 2][r] = v.z; tile[k4 + 3][r] = v.w;
    }
}

// ---------------------------------------------------------------------------
// forward
// ---------------------------------------------------------------------------
__global__ void __launch_bounds__(PG_NT)
gru_persistent_fwd_kernel(const GruDesc d, const float* __restrict__ Whh, const float* __restrict__ bhh) {
    extern __shared__ __align__(16) float pg_smem[];
    const int H = d.H;
    float* Ws = pg_smem;                                            // [3 * KB][H]: rows (unit j, gate g) -> j * 3 + g
    float (*tile)[PG_RB + 1] = reinterpret_cast<float (*)[PG_RB + 1]>(pg_smem + 3 * PG_KB * H);
    cg::grid_group grid = cg::this_grid();

    const int tid = threadIdx.x;
    const int kb = blockIdx.x * PG_KB;                              // first hidden unit of this CTA
    const int row0 = blockIdx.y * PG_RB;
    const int nrows = min(PG_RB, d.N - row0);
    const int r = tid % PG_RB, up = tid / PG_RB;                    // this thread: env row r, units 2*up, 2*up + 1

    for (int e = tid; e < 3 * PG_KB * H; e += PG_NT) {
        const int c = e / H, k = e - c * H;
        const int j = c / 3, g = c - j * 3;
        Ws[e] = Whh[(long)(g * H + kb + j) * H + k];
    }
    float bias[6];
#pragma unroll
    for (int c = 0; c < 6; ++c) {
        const int j = 2 * up + c / 3, g = c % 3;
        bias[c] = bhh[g * H + kb + j];
    }
    // hm_0 = h0 * mask_0 for this CTA's units (every CTA writes its own columns of hm[0])
    if (r < nrows) {
#pragma unroll
        for (int u = 0; u < 2; ++u) {
            const int k = kb + 2 * up + u;
            d.hm[(long)(row0 + r) * H + k] = d.h0[(long)(row0 + r) * H + k] * pg_mask(d, 0, row0 + r);
        }
    }
    __threadfence();
    grid.sync();

    for (int t = 0; t < d.T; ++t) {
        const float* hm_t = d.hm + (long)t * d.N * H;
        float acc[6] = {0.f, 0.f, 0.f, 0.f, 0.f, 0.f};
        for (int k0 = 0; k0 < H; k0 += PG_KC) {
            __syncthreads();
            pg_load_tile(tile, hm_t, H, row0, nrows, k0, tid);
            __syncthreads();
            const float* w = Ws + (long)(6 * up) * H + k0;
#pragma unroll 4
            for (int k = 0; k < PG_KC; k += 4) {
                const float h0v = tile[k][r], h1v = tile[k + 1][r], h2v = tile[k + 2][r], h3v = tile[k + 3][r];
#pragma unroll
                for (int c = 0; c < 6; ++c) {
                    const float4 wv = *reinterpret_cast<const float4*>(w + (long)c * H + k);
                    acc[c] = fmaf(h0v, wv.x, acc[c]); acc[c] = fmaf(h1v, wv.y, acc[c]);
                    acc[c] = fmaf(h2v, wv.z, acc[c]); acc[c] = fmaf(h3v, wv.w, acc[c]);
                }
            }
        }
        if (r < nrows) {
            const long row = (long)t * d.N + row0 + r;
            const float* gi = d.gi + row * 3 * H;
            const float mnext = (t + 1 < d.T) ? pg_mask(d, t + 1, row0 + r) : 0.f;
#pragma unroll
            for (int u = 0; u < 2; ++u) {
                const int k = kb + 2 * up + u;
                const float ghr = acc[3 * u + 0] + bias[3 * u + 0];
                const float ghz = acc[3 * u + 1] + bias[3 * u + 1];
                const float ghn = acc[3 * u + 2] + bias[3 * u + 2];
                const float hm = hm_t[(long)(row0 + r) * H + k];
                const float rr = pg_sigmoid(gi[k] + ghr);
                const float zz = pg_sigmoid(gi[H + k] + ghz);
                const float nn = tanhf(gi[2 * H + k] + rr * ghn);
                const float h = (1.f - zz) * nn + zz * hm;
                d.out[row * H + k] = h;
                if (d.r) {
                    d.r[row * H + k] = rr; d.z[row * H + k] = zz; d.n[row * H + k] = nn;
                    d.gh[row * 3 * H + 2 * H + k] = ghn;             // the only part of gh the backward pass reads
                }
                if (t + 1 < d.T) d.hm[(row + d.N) * H + k] = h * mnext;
            }
        }
        if (t + 1 < d.T) {
            __threadfence();
            grid.sync();
        }
    }
}

// ---------------------------------------------------------------------------
// backward through time
// ---------------------------------------------------------------------------
__global__ void __launch_bounds__(PG_NT)
gru_persistent_bwd_kernel(const GruDesc d, const float* __restrict__ Whh) {
    extern __shared__ __align__(16) float pg_smem[];
    const int H = d.H;
    float* Wc = pg_smem;                                            // [3H][KB]: W_hh[c][kb + j]
    float (*tile)[PG_RB + 1] = reinterpret_cast<float (*)[PG_RB + 1]>(pg_smem + 3 * H * PG_KB);
    cg::grid_group grid = cg::this_grid();

    const int tid = threadIdx.x;
    const int kb = blockIdx.x * PG_KB;
    const int row0 = blockIdx.y * PG_RB;
    const int nrows = min(PG_RB, d.N - row0);
    const int r = tid % PG_RB, up = tid / PG_RB;

    for (int e = tid; e < 3 * H * PG_KB; e += PG_NT) {
        const int c = e / PG_KB, j = e - c * PG_KB;
        Wc[e] = Whh[(long)c * H + kb + j];
    }
    float carry[2] = {0.f, 0.f};          // d loss / d hm_{t+1} for this thread's (row, 2 units)

    for (int t = d.T - 1; t >= 0; --t) {
        const long row = (long)t * d.N + row0 + r;
        float dhz[2] = {0.f, 0.f};
        // ---- phase A: gate gradients of this CTA's units ----
        if (r < nrows) {
            const float mnext = (t + 1 < d.T) ? pg_mask(d, t + 1, row0 + r) : 0.f;
#pragma unroll
            for (int u = 0; u < 2; ++u) {
                const int k = kb + 2 * up + u;
                float dh = d.dout[row * H + k];
                if (t + 1 < d.T) dh += carry[u] * mnext;
                const float rr = d.r[row * H + k], zz = d.z[row * H + k], nn = d.n[row * H + k];
                const float hm = d.hm[row * H + k];
                const float ghn = d.gh[row * 3 * H + 2 * H + k];
                const float dn_pre = dh * (1.f - zz) * (1.f - nn * nn);
                const float dz_pre = dh * (hm - nn) * zz * (1.f - zz);
                const float dr_pre = dn_pre * ghn * rr * (1.f - rr);
                float* dgi = d.dgi + row * 3 * H;
                float* dgh = d.dgh + row * 3 * H;
                dgi[k] = dr_pre; dgi[H + k] = dz_pre; dgi[2 * H + k] = dn_pre;
                dgh[k] = dr_pre; dgh[H + k] = dz_pre; dgh[2 * H + k] = dn_pre * rr;
                dhz[u] = dh * zz;
            }
        }
        if (t == 0) break;                 // the gradient w.r.t. the initial state is not needed
        __threadfence();
        grid.sync();
        // ---- phase B: carry[u] = dh * z + (dgh_t W_hh)[row, unit] over all 3H gate columns ----
        const float* dgh_t = d.dgh + (long)t * d.N * 3 * H;
        float acc0 = 0.f, acc1 = 0.f;
        for (int c0 = 0; c0 < 3 * H; c0 += PG_KC) {
            __syncthreads();
            pg_load_tile(tile, dgh_t, 3 * H, row0, nrows, c0, tid);
            __syncthreads();
            const float* w = Wc + (long)c0 * PG_KB + 2 * up;
#pragma unroll 8
            for (int c = 0; c < PG_KC; ++c) {
                const float g = tile[c][r];
                const float2 wv = *reinterpret_cast<const float2*>(w + c * PG_KB);
                acc0 = fmaf(g, wv.x, acc0);
                acc1 = fmaf(g, wv.y, acc1);
            }
        }
        carry[0] = dhz[0] + acc0;
        carry[1] = dhz[1] + acc1;
    }
}

static size_t pg_smem_bytes(int H) { return (size_t)(3 * PG_KB * H + PG_KC * (PG_RB + 1)) * sizeof(float); }

// Cooperative launch: every CTA must be co-resident.  Asks the runtime how many fit.
bool gru_persistent_supported(int N, int H) {
    if (N <= 0 || H % PG_KB != 0 || H % PG_KC != 0) return false;
    const size_t smem = pg_smem_bytes(H);
    if (smem > 220 * 1024) return false;
    static int sms = 0;
    if (!sms) {
        int dev = 0;
        cudaGetDevice(&dev);
        cudaDeviceGetAttribute(&sms, cudaDevAttrMultiProcessorCount, dev);
    }
    cudaFuncSetAttribute(gru_persistent_fwd_kernel, cudaFuncAttributeMaxDynamicSharedMemorySize, 220 * 1024);
    cudaFuncSetAttribute(gru_persistent_bwd_kernel, cudaFuncAttributeMaxDynamicSharedMemorySize, 220 * 1024);
    int pf = 0, pb = 0;
    cudaOccupancyMaxActiveBlocksPerMultiprocessor(&pf, gru_persistent_fwd_kernel, PG_NT, smem);
    cudaOccupancyMaxActiveBlocksPerMultiprocessor(&pb, gru_persistent_bwd_kernel, PG_NT, smem);
    const long ctas = (long)(H / PG_KB) * ((N + PG_RB - 1) / PG_RB);
    return ctas <= (long)sms * (pf < pb ? pf : pb);
}

int gru_persistent(int backward, const GruDesc& d, const float* Whh, const float* bhh, cudaStream_t st) {
    A2CB_REQUIRE(gru_persistent_supported(d.N, d.H), "gru_persistent: unsupported shape (N=%d H=%d)", d.N, d.H);
    A2CB_REQUIRE(Whh && d.hm && d.masks, "gru_persistent: null pointer");
    A2CB_REQUIRE(backward || (bhh && d.h0 && d.gi && d.out), "gru_persistent: forward pointers");
    A2CB_REQUIRE(!backward || (d.r && d.z && d.n && d.gh && d.dout && d.dgi && d.dgh), "gru_persistent: backward pointers");
    const size_t smem = pg_smem_bytes(d.H);
    static bool attr[2] = {false, false};
    if (!attr[backward]) {
        cudaError_t e = backward
            ? cudaFuncSetAttribute(gru_persistent_bwd_kernel, cudaFuncAttributeMaxDynamicSharedMemorySize, 220 * 1024)
            : cudaFuncSetAttribute(gru_persistent_fwd_kernel, cudaFuncAttributeMaxDynamicSharedMemorySize, 220 * 1024);
        if (e != cudaSuccess) { set_error("gru_persistent: %s", cudaGetErrorString(e)); return A2CB_ERR_CUDA; }
        attr[backward] = true;
    }
    dim3 grid((unsigned)(d.H / PG_KB), (unsigned)((d.N + PG_RB - 1) / PG_RB));
    GruDesc dd = d;
    void* args_f[] = {(void*)&dd, (void*)&Whh, (void*)&bhh};
    void* args_b[] = {(void*)&dd, (void*)&Whh};
    cudaError_t e = backward
        ? cudaLaunchCooperativeKernel((const void*)gru_persistent_bwd_kernel, grid, dim3(PG_NT), args_b, smem, st)
        : cudaLaunchCooperativeKernel((const void*)gru_persistent_fwd_kernel, grid, dim3(PG_NT), args_f, smem, st);
    if (e != cudaSuccess) {
        set_error("gru_persistent: cooperative launch failed: %s", cudaGetErrorString(e));
        cudaGetLastError();
        return A2CB_ERR_CUDA;
    }
    return check_launch("gru_persistent");
}

}  // namespace a2cb
