// Element-wise helpers around the TMA-fed tensor-core engine (tcx.cu): operand planes, weight images,
// the frame layout the first convolution reads, and bias gradients.  All HBM-bound.
#include <cuda_bf16.h>

#include "common.cuh"
#include "decls.cuh"
#include "gemm.cuh"
#include "tcx.cuh"

namespace a2cb {

namespace {

__device__ __forceinline__ float ax_lo(float v) { return v - __uint_as_float(__float_as_uint(v) & 0xffffe000u); }

// lo[i] = x[i] - trunc_tf32(x[i])
__global__ void __launch_bounds__(256)
tcx_lo_kernel(float4* __restrict__ lo, const float4* __restrict__ x, long n4) {
    for (long i = (long)blockIdx.x * blockDim.x + threadIdx.x; i < n4; i += (long)gridDim.x * blockDim.x) {
        const float4 v = x[i];
        lo[i] = make_float4(ax_lo(v.x), ax_lo(v.y), ax_lo(v.z), ax_lo(v.w));
    }
}

// img[n][k] = src[ntab[n] + ktab[k]] (0 where either table entry is negative), img_lo = its lo plane
__global__ void __launch_bounds__(256)
tcx_prep_kernel(float* __restrict__ img, float* __restrict__ img_lo, const float* __restrict__ src, long N, long K,
                const int32_t* __restrict__ ntab, const int32_t* __restrict__ ktab) {
    const long e = (long)blockIdx.x * blockDim.x + threadIdx.x;
    if (e >= N * K) return;
    const long n = e / K, k = e - n * K;
    const int a = __ldg(ntab + n), b = __ldg(ktab + k);
    const float v = (a >= 0 && b >= 0) ? __ldg(src + (long)a + b) : 0.f;
    img[e] = v;
    img_lo[e] = ax_lo(v);
}

// Frames [*, C, H, W] (uint8 or fp32, rows picked through idx) -> space-to-depth fp32 [B, H/S, W/S, C*S*S]
// with channel (c, dy, dx): the stride-S first convolution becomes a stride-1 convolution over S x S blocks.
template <typename T>
__global__ void __launch_bounds__(256)
tcx_s2d_kernel(float* __restrict__ out, float* __restrict__ out_lo, const T* __restrict__ frames,
               const int64_t* __restrict__ idx, int B, int C, int H, int W, int S, int scatter) {
    const int Hb = H / S, Wb = W / S, CC = C * S * S;
    const long total = (long)B * Hb * Wb * CC;
    for (long e = (long)blockIdx.x * blockDim.x + threadIdx.x; e < total; e += (long)gridDim.x * blockDim.x) {
        const int ch = (int)(e % CC);
        long t = e / CC;
        const int bx = (int)(t % Wb); t /= Wb;
        const int by = (int)(t % Hb); t /= Hb;
        const long b = t;
        const int dx = ch % S, dy = (ch / S) % S, c = ch / (S * S);
        const long row = idx ? (long)idx[b] : b;
        const float v = (float)frames[((row * C + c) * H + (by * S + dy)) * W + bx * S + dx];
        const long o = scatter ? e + (row - b) * ((long)Hb * Wb * CC) : e;     // scatter: image slot = source row
        out[o] = v;
        if (out_lo) out_lo[o] = ax_lo(v);
    }
}

// Several weight images in one launch: jobs[j] describes one image, `first` = index of its first element in the
// concatenated element range.
struct PrepJob { float* img; float* img_lo; const float* src; const int32_t* ntab; const int32_t* ktab; long N, K, first;
                 float* img3; long bf16; };

__global__ void __launch_bounds__(256)
tcx_prep_multi_kernel(const PrepJob* __restrict__ jobs, int n_jobs, long total) {
    const long e0 = (long)blockIdx.x * blockDim.x + threadIdx.x;
    if (e0 >= total) return;
    int j = 0;
    while (j + 1 < n_jobs && e0 >= jobs[j + 1].first) ++j;
    const PrepJob q = jobs[j];
    const long e = e0 - q.first;
    const long n = e / q.K, k = e - n * q.K;
    const int a = __ldg(q.ntab + n), b = __ldg(q.ktab + k);
    const float v = (a >= 0 && b >= 0) ? __ldg(q.src + (long)a + b) : 0.f;
    if (q.bf16) {
        // w = b1 + b2 + b3 with three bf16 terms (8 + 8 + 8 significant bits)
        const __nv_bfloat16 b1 = __float2bfloat16_rn(v);
        const float r1 = v - __bfloat162float(b1);
        const __nv_bfloat16 b2 = __float2bfloat16_rn(r1);
        const __nv_bfloat16 b3 = __float2bfloat16_rn(r1 - __bfloat162float(b2));
        reinterpret_cast<__nv_bfloat16*>(q.img)[e] = b1;
        reinterpret_cast<__nv_bfloat16*>(q.img_lo)[e] = b2;
        reinterpret_cast<__nv_bfloat16*>(q.img3)[e] = b3;
        return;
    }
    q.img[e] = v;
    q.img_lo[e] = ax_lo(v);
}

// Atari geometry (C = 4, 84 x 84, S = 4): one block per (image, block row).  The 16 source rows (c, dy) of
// 84 pixels are read as coalesced 4-byte / 16-byte words into shared memory and written back as 21 x 64
// contiguous floats -- both sides of the transposition are fully coalesced.
//
// Frame ring (valid != nullptr; uint8 only): `frames` is [slots, ring_n, 84, 84] with every 84 x 84 frame stored ONCE;
// the 4-stack of rollout row (t, n) = row / ring_n, row % ring_n is frames of slots t .. t + 3 (channel c <-> slot
// t + c, newest last), of which only the newest valid[row] are real: the older ones are zero, exactly as
// VecPyTorchFrameStack clears the stack when an episode ends (reference envs.py:243-250).
template <typename T>
__global__ void __launch_bounds__(352)
tcx_s2d_atari_kernel(float4* __restrict__ out, float4* __restrict__ out_lo, uint2* __restrict__ out_b16,
                     const T* __restrict__ frames, const int64_t* __restrict__ idx, int scatter,
                     const uint8_t* __restrict__ valid, int ring_n, int by_per) {
    // One CTA = `by_per` of the 21 block rows of one image (a block row: 16 source rows of 84 pixels -> 21 output pixels
    // of 64 channels = 5.4 KB of fp32 + 2.7 KB of bf16).  One block row per CTA (172,032 CTAs for a 64-environment chunk)
    // left the kernel bound by CTA turnover at 3.3 TB/s of writes; the tile is double-buffered so that a block row costs
    // one barrier.
    __shared__ float4 tile[2][16][21];
    const int groups = 21 / by_per;
    const int b = blockIdx.x / groups, by0 = (blockIdx.x % groups) * by_per;
    const long row = idx ? (long)idx[b] : b;
    const long slot = scatter ? row : b;                                 // scatter: image slot = source row
    const int i = threadIdx.x;
    const int r = i / 21, bx = i % 21;                                   // load role: r = c * 4 + dy
    const int sbx = i >> 4, sr = i & 15;                                 // store role
    const bool ring = sizeof(T) == 1 && valid != nullptr;
    long t = 0, n = 0;
    int first_real = 0;
    if (ring) {
        t = row / ring_n;
        n = row - t * ring_n;
        first_real = 4 - (int)valid[row];
    }
    for (int k = 0; k < by_per; ++k) {
        const int by = by0 + k;
        if (i < 336) {
            float4 v;
            if (ring) {
                const int c = r >> 2;
                const long src = (((t + c) * ring_n + n) * 84 + (by * 4 + (r & 3))) * 84 + bx * 4;
                uchar4 u = make_uchar4(0, 0, 0, 0);
                if (c >= first_real) u = *reinterpret_cast<const uchar4*>(reinterpret_cast<const uint8_t*>(frames) + src);
                v = make_float4((float)u.x, (float)u.y, (float)u.z, (float)u.w);
            } else {
                const long src = ((row * 4 + (r >> 2)) * 84 + (by * 4 + (r & 3))) * 84 + bx * 4;
                if (sizeof(T) == 1) {
                    const uchar4 u = *reinterpret_cast<const uchar4*>(reinterpret_cast<const uint8_t*>(frames) + src);
                    v = make_float4((float)u.x, (float)u.y, (float)u.z, (float)u.w);
                } else {
                    v = *reinterpret_cast<const float4*>(reinterpret_cast<const float*>(frames) + src);
                }
            }
            tile[k & 1][r][bx] = v;
        }
        __syncthreads();
        if (i < 336) {
            const float4 v = tile[k & 1][sr][sbx];
            const long o = ((slot * 21 + by) * 21 + sbx) * 16 + sr;
            out[o] = v;
            if (out_lo) out_lo[o] = make_float4(ax_lo(v.x), ax_lo(v.y), ax_lo(v.z), ax_lo(v.w));
            if (out_b16) {
                const __nv_bfloat162 p0 = __floats2bfloat162_rn(v.x, v.y), p1 = __floats2bfloat162_rn(v.z, v.w);
                uint2 u;
                u.x = *reinterpret_cast<const uint32_t*>(&p0);
                u.y = *reinterpret_cast<const uint32_t*>(&p1);
                out_b16[o] = u;
            }
        }
    }
}

// part[blk][c] = sum over this block's rows of x[row * pitch + c]   (C a multiple of 4, row-contiguous).
// Thread = (channel quad, row lane); float4 loads, four independent accumulator sets, shared-memory fold
// across the row lanes in a fixed order (deterministic).
template <bool LO>
__global__ void __launch_bounds__(256)
tcx_colsum_kernel(float* __restrict__ part, const float* __restrict__ x, float* __restrict__ lo, long rows, int C, long pitch,
                  long rows_per_blk) {
    __shared__ float4 red[256];
    const long r0 = (long)blockIdx.x * rows_per_blk, r1 = min(rows, r0 + rows_per_blk);
    const int nq = C / 4;
    const int qw = nq < 256 ? nq : 256;            // channel quads covered per pass
    const int rl_n = 256 / qw;                     // row lanes
    const int ql = threadIdx.x % qw, rl = threadIdx.x / qw;
    for (int q0 = 0; q0 < nq; q0 += qw) {
        float4 s0 = make_float4(0.f, 0.f, 0.f, 0.f), s1 = s0, s2 = s0, s3 = s0;
        if (rl < rl_n && q0 + ql < nq) {
            const float* base = x + (long)(q0 + ql) * 4;
            long r = r0 + rl;
            float* lob = LO ? lo + (long)(q0 + ql) * 4 : nullptr;
            auto put_lo = [&](long rr, const float4 v) {
                if (LO) *reinterpret_cast<float4*>(lob + rr * pitch) = make_float4(ax_lo(v.x), ax_lo(v.y), ax_lo(v.z), ax_lo(v.w));
            };
            for (; r + 3L * rl_n < r1; r += 4L * rl_n) {
                const float4 a = *reinterpret_cast<const float4*>(base + r * pitch);
                const float4 b = *reinterpret_cast<const float4*>(base + (r + rl_n) * pitch);
                const float4 c = *reinterpret_cast<const float4*>(base + (r + 2L * rl_n) * pitch);
                const float4 d = *reinterpret_cast<const float4*>(base + (r + 3L * rl_n) * pitch);
                put_lo(r, a); put_lo(r + rl_n, b); put_lo(r + 2L * rl_n, c); put_lo(r + 3L * rl_n, d);
                s0.x += a.x; s0.y += a.y; s0.z += a.z; s0.w += a.w;
                s1.x += b.x; s1.y += b.y; s1.z += b.z; s1.w += b.w;
                s2.x += c.x; s2.y += c.y; s2.z += c.z; s2.w += c.w;
                s3.x += d.x; s3.y += d.y; s3.z += d.z; s3.w += d.w;
            }
            for (; r < r1; r += rl_n) {
                const float4 a = *reinterpret_cast<const float4*>(base + r * pitch);
                put_lo(r, a);
                s0.x += a.x; s0.y += a.y; s0.z += a.z; s0.w += a.w;
            }
        }
        red[threadIdx.x] = make_float4((s0.x + s1.x) + (s2.x + s3.x), (s0.y + s1.y) + (s2.y + s3.y),
                                       (s0.z + s1.z) + (s2.z + s3.z), (s0.w + s1.w) + (s2.w + s3.w));
        __syncthreads();
        if (rl == 0 && q0 + ql < nq) {
            float4 t = red[ql];
            for (int j = 1; j < rl_n; ++j) {
                const float4 u = red[j * qw + ql];
                t.x += u.x; t.y += u.y; t.z += u.z; t.w += u.w;
            }
            *reinterpret_cast<float4*>(part + (long)blockIdx.x * C + (long)(q0 + ql) * 4) = t;
        }
        __syncthreads();
    }
}

// Weight + bias gradient of a narrow Linear layer (the policy / value heads, N <= 8 outputs):
//   part[blk][n * K + k] = sum over the block's rows of gy[row * gy_stride + n] * x[row * K + k]
//   part[blk][N * K + n] = sum of gy[row * gy_stride + n]
// x is streamed once with coalesced loads (HBM-bound); the N gradients of a row are warp-uniform broadcasts.
template <int N>
__global__ void __launch_bounds__(256)
tcx_narrow_wgrad_kernel(float* __restrict__ part, const float* __restrict__ x, const float* __restrict__ gy, long rows,
                        int K, int gy_stride, long rows_per_blk) {
    const long r0 = (long)blockIdx.x * rows_per_blk, r1 = min(rows, r0 + rows_per_blk);
    float* dst = part + (long)blockIdx.x * (N * K + N);
    for (int k = threadIdx.x; k < K; k += 256) {
        float acc[N];
#pragma unroll
        for (int n = 0; n < N; ++n) acc[n] = 0.f;
        float bsum[N];
#pragma unroll
        for (int n = 0; n < N; ++n) bsum[n] = 0.f;
#pragma unroll 4
        for (long r = r0; r < r1; ++r) {
            const float xv = x[r * K + k];
#pragma unroll
            for (int n = 0; n < N; ++n) {
                const float g = __ldg(gy + r * gy_stride + n);
                acc[n] = fmaf(g, xv, acc[n]);
                bsum[n] += g;
            }
        }
#pragma unroll
        for (int n = 0; n < N; ++n) dst[n * K + k] = acc[n];
        if (k < 1) {
#pragma unroll
            for (int n = 0; n < N; ++n) dst[N * K + n] = bsum[n];
        }
    }
}

// Narrow Linear layer forward (heads): z[row][o] = bias[o] + sum_k x[row][k] W[o][k], o < N <= 8.  One warp per row,
// lanes split k with 16-byte loads (x is streamed once: HBM-bound), N warp reductions.
template <int N>
__global__ void __launch_bounds__(256)
tcx_narrow_fwd_kernel(float* __restrict__ z, const float* __restrict__ x, const float* __restrict__ W,
                      const float* __restrict__ bias, long rows, int K, int z_stride) {
    const int lane = threadIdx.x & 31;
    const long row = (long)blockIdx.x * 8 + (threadIdx.x >> 5);
    if (row >= rows) return;
    float acc[N];
#pragma unroll
    for (int o = 0; o < N; ++o) acc[o] = 0.f;
    const float4* xr = reinterpret_cast<const float4*>(x + row * K);
    for (int k4 = lane; k4 < K / 4; k4 += 32) {
        const float4 xv = xr[k4];
#pragma unroll
        for (int o = 0; o < N; ++o) {
            const float4 wv = __ldg(reinterpret_cast<const float4*>(W + (long)o * K) + k4);
            acc[o] = fmaf(xv.x, wv.x, acc[o]); acc[o] = fmaf(xv.y, wv.y, acc[o]);
            acc[o] = fmaf(xv.z, wv.z, acc[o]); acc[o] = fmaf(xv.w, wv.w, acc[o]);
        }
    }
#pragma unroll
    for (int o = 0; o < N; ++o) acc[o] = warp_sum(acc[o]);
    if (lane == 0) {
#pragma unroll
        for (int o = 0; o < N; ++o) z[row * z_stride + o] = acc[o] + __ldg(bias + o);
    }
}

// Narrow Linear layer data gradient: gx[row][k] = sum_o dz[row][o] W[o][k], optionally masked by relu'(mask) and
// with the lo plane of the result.  Thread = (row, 4 consecutive k); W (N x K) is read through L1.
template <int N>
__global__ void __launch_bounds__(256)
tcx_narrow_dgrad_kernel(float4* __restrict__ gx, float4* __restrict__ gx_lo, const float* __restrict__ dz,
                        const float* __restrict__ W, const float4* __restrict__ mask, long rows, int K, int dz_stride) {
    const int k4n = K / 4;
    const long e = (long)blockIdx.x * blockDim.x + threadIdx.x;
    if (e >= rows * k4n) return;
    const long row = e / k4n;
    const int k4 = (int)(e - row * k4n);
    float4 s = make_float4(0.f, 0.f, 0.f, 0.f);
#pragma unroll
    for (int o = 0; o < N; ++o) {
        const float g = __ldg(dz + row * dz_stride + o);
        const float4 wv = __ldg(reinterpret_cast<const float4*>(W + (long)o * K) + k4);
        s.x = fmaf(g, wv.x, s.x); s.y = fmaf(g, wv.y, s.y); s.z = fmaf(g, wv.z, s.z); s.w = fmaf(g, wv.w, s.w);
    }
    if (mask) {
        const float4 m = mask[e];
        s.x = m.x > 0.f ? s.x : 0.f; s.y = m.y > 0.f ? s.y : 0.f; s.z = m.z > 0.f ? s.z : 0.f; s.w = m.w > 0.f ? s.w : 0.f;
    }
    gx[e] = s;
    if (gx_lo) gx_lo[e] = make_float4(ax_lo(s.x), ax_lo(s.y), ax_lo(s.z), ax_lo(s.w));
}

}  // namespace

int tcx_lo_plane(float* lo, const float* x, long n, cudaStream_t st) {
    A2CB_REQUIRE(lo && x && n >= 0 && n % 4 == 0, "tcx_lo_plane: bad arguments");
    if (n == 0) return A2CB_OK;
    const long n4 = n / 4;
    const unsigned grid = (unsigned)min((n4 + 255) / 256, (long)kNumSM * 16);
    tcx_lo_kernel<<<grid, 256, 0, st>>>(reinterpret_cast<float4*>(lo), reinterpret_cast<const float4*>(x), n4);
    return check_launch("tcx_lo_plane");
}

int tcx_prep_weights(float* img, float* img_lo, const float* src, long N, long K, const int32_t* ntab,
                     const int32_t* ktab, cudaStream_t st) {
    A2CB_REQUIRE(img && img_lo && src && ntab && ktab && N > 0 && K > 0, "tcx_prep_weights: bad arguments");
    const unsigned grid = (unsigned)((N * K + 255) / 256);
    tcx_prep_kernel<<<grid, 256, 0, st>>>(img, img_lo, src, N, K, ntab, ktab);
    return check_launch("tcx_prep_weights");
}

int tcx_s2d(float* out, float* out_lo, void* out_b16, const void* frames, int src_is_u8, const int64_t* idx, int B, int C,
            int H, int W, int S, cudaStream_t st, int scatter, const uint8_t* valid, int ring_n) {
    A2CB_REQUIRE(out && frames && B > 0 && S > 0 && H % S == 0 && W % S == 0, "tcx_s2d: bad arguments");
    A2CB_REQUIRE(!scatter || idx, "tcx_s2d: scatter mode needs a row list");
    A2CB_REQUIRE(!valid || (src_is_u8 && C == 4 && H == 84 && W == 84 && S == 4 && ring_n > 0),
                 "tcx_s2d: the frame ring holds uint8 84 x 84 frames stacked 4 deep");
    if (C == 4 && H == 84 && W == 84 && S == 4) {
        static int by_per = 0;                                   // block rows per CTA: 1, 3, 7 or 21 (A2CB_S2D_BY)
        if (by_per == 0) {
            const char* e = getenv("A2CB_S2D_BY");
            const int v = e ? atoi(e) : 3;
            by_per = (v == 1 || v == 3 || v == 7 || v == 21) ? v : 3;
        }
        const unsigned grid = (unsigned)((long)B * (21 / by_per));
        if (src_is_u8)
            tcx_s2d_atari_kernel<uint8_t><<<grid, 352, 0, st>>>(reinterpret_cast<float4*>(out), reinterpret_cast<float4*>(out_lo),
                                                                 reinterpret_cast<uint2*>(out_b16),
                                                                 reinterpret_cast<const uint8_t*>(frames), idx, scatter,
                                                                 valid, ring_n, by_per);
        else
            tcx_s2d_atari_kernel<float><<<grid, 352, 0, st>>>(reinterpret_cast<float4*>(out), reinterpret_cast<float4*>(out_lo),
                                                               reinterpret_cast<uint2*>(out_b16),
                                                               reinterpret_cast<const float*>(frames), idx, scatter,
                                                               nullptr, 0, by_per);
        return check_launch("tcx_s2d");
    }
    A2CB_REQUIRE(!out_b16, "tcx_s2d: the bf16 copy is only produced for the Atari geometry");
    const long total = (long)B * C * H * W;
    const unsigned grid = (unsigned)min((total + 255) / 256, (long)kNumSM * 32);
    if (src_is_u8)
        tcx_s2d_kernel<uint8_t><<<grid, 256, 0, st>>>(out, out_lo, reinterpret_cast<const uint8_t*>(frames), idx, B, C, H, W, S, scatter);
    else
        tcx_s2d_kernel<float><<<grid, 256, 0, st>>>(out, out_lo, reinterpret_cast<const float*>(frames), idx, B, C, H, W, S, scatter);
    return check_launch("tcx_s2d");
}

int tcx_colsum_blocks(long rows) { return (int)min((rows + 15) / 16, (long)kNumSM * 8); }

int tcx_colsum(float* part, const float* x, float* lo, long rows, int C, long pitch, cudaStream_t st) {
    A2CB_REQUIRE(part && x && rows > 0 && C > 0 && C % 4 == 0 && pitch % 4 == 0, "tcx_colsum: bad arguments");
    const int blocks = tcx_colsum_blocks(rows);
    const long per = (rows + blocks - 1) / blocks;
    if (lo) tcx_colsum_kernel<true><<<blocks, 256, 0, st>>>(part, x, lo, rows, C, pitch, per);
    else tcx_colsum_kernel<false><<<blocks, 256, 0, st>>>(part, x, nullptr, rows, C, pitch, per);
    return check_launch("tcx_colsum");
}

int tcx_narrow_fwd(float* z, const float* x, const float* W, const float* bias, long rows, int K, int N, int z_stride,
                   cudaStream_t st) {
    A2CB_REQUIRE(z && x && W && bias && rows > 0 && K > 0 && K % 4 == 0 && N >= 1 && N <= 8, "tcx_narrow_fwd: bad arguments");
    const unsigned grid = (unsigned)((rows + 7) / 8);
    switch (N) {
#define A2CB_NF(n) case n: tcx_narrow_fwd_kernel<n><<<grid, 256, 0, st>>>(z, x, W, bias, rows, K, z_stride); break;
        A2CB_NF(1) A2CB_NF(2) A2CB_NF(3) A2CB_NF(4) A2CB_NF(5) A2CB_NF(6) A2CB_NF(7) A2CB_NF(8)
#undef A2CB_NF
    }
    return check_launch("tcx_narrow_fwd");
}

int tcx_narrow_dgrad(float* gx, float* gx_lo, const float* dz, const float* W, const float* mask, long rows, int K, int N,
                     int dz_stride, cudaStream_t st) {
    A2CB_REQUIRE(gx && dz && W && rows > 0 && K > 0 && K % 4 == 0 && N >= 1 && N <= 8, "tcx_narrow_dgrad: bad arguments");
    const unsigned grid = (unsigned)((rows * (K / 4) + 255) / 256);
    switch (N) {
#define A2CB_ND(n) case n: tcx_narrow_dgrad_kernel<n><<<grid, 256, 0, st>>>(reinterpret_cast<float4*>(gx), reinterpret_cast<float4*>(gx_lo), dz, W, reinterpret_cast<const float4*>(mask), rows, K, dz_stride); break;
        A2CB_ND(1) A2CB_ND(2) A2CB_ND(3) A2CB_ND(4) A2CB_ND(5) A2CB_ND(6) A2CB_ND(7) A2CB_ND(8)
#undef A2CB_ND
    }
    return check_launch("tcx_narrow_dgrad");
}

int tcx_prep_multi(const void* jobs, int n_jobs, long total, cudaStream_t st) {
    A2CB_REQUIRE(jobs && n_jobs > 0 && total > 0, "tcx_prep_multi: bad arguments");
    tcx_prep_multi_kernel<<<(unsigned)((total + 255) / 256), 256, 0, st>>>(reinterpret_cast<const PrepJob*>(jobs), n_jobs, total);
    return check_launch("tcx_prep_multi");
}

int tcx_narrow_wgrad_blocks(long rows) { return (int)min((rows + 63) / 64, (long)kNumSM * 4); }

int tcx_narrow_wgrad(float* part, const float* x, const float* gy, long rows, int K, int N, int gy_stride, cudaStream_t st) {
    A2CB_REQUIRE(part && x && gy && rows > 0 && K > 0 && N >= 1 && N <= 8, "tcx_narrow_wgrad: bad arguments (N = %d)", N);
    const int blocks = tcx_narrow_wgrad_blocks(rows);
    const long per = (rows + blocks - 1) / blocks;
    switch (N) {
#define A2CB_NW(n) case n: tcx_narrow_wgrad_kernel<n><<<blocks, 256, 0, st>>>(part, x, gy, rows, K, gy_stride, per); break;
        A2CB_NW(1) A2CB_NW(2) A2CB_NW(3) A2CB_NW(4) A2CB_NW(5) A2CB_NW(6) A2CB_NW(7) A2CB_NW(8)
#undef A2CB_NW
    }
    return check_launch("tcx_narrow_wgrad");
}

}  // namespace a2cb
