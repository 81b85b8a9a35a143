// Element-wise helpers around the TMA-fed tensor-core engine (tcx.cu): operand planes, weight images,
// the frame layout the first convolution reads, and bias gradients.  All HBM-bound.
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
               const int64_t* __restrict__ idx, int B, int C, int H, int W, int S) {
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
        out[e] = v;
        if (out_lo) out_lo[e] = ax_lo(v);
    }
}

// part[blk][c] = sum over this block's rows of x[row * pitch + c]   (C <= 2048 channels, row-contiguous)
__global__ void __launch_bounds__(256)
tcx_colsum_kernel(float* __restrict__ part, const float* __restrict__ x, long rows, int C, long pitch, long rows_per_blk) {
    __shared__ float red[256];
    const long r0 = (long)blockIdx.x * rows_per_blk, r1 = min(rows, r0 + rows_per_blk);
    // thread -> (channel lane cl, row lane rl): consecutive threads read consecutive channels of a row
    const int cw = C < 256 ? C : 256;              // channels covered per pass
    const int rl_n = 256 / cw;                     // rows in flight
    const int cl = threadIdx.x % cw, rl = threadIdx.x / cw;
    for (int c0 = 0; c0 < C; c0 += cw) {
        float s = 0.f;
        if (rl < rl_n && c0 + cl < C)
            for (long r = r0 + rl; r < r1; r += rl_n) s += x[r * pitch + c0 + cl];
        red[threadIdx.x] = s;
        __syncthreads();
        if (rl == 0 && c0 + cl < C) {
            float tsum = 0.f;
            for (int j = 0; j < rl_n; ++j) tsum += red[j * cw + cl];
            part[(long)blockIdx.x * C + c0 + cl] = tsum;
        }
        __syncthreads();
    }
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

int tcx_s2d(float* out, float* out_lo, const void* frames, int src_is_u8, const int64_t* idx, int B, int C, int H, int W,
            int S, cudaStream_t st) {
    A2CB_REQUIRE(out && frames && B > 0 && S > 0 && H % S == 0 && W % S == 0, "tcx_s2d: bad arguments");
    const long total = (long)B * C * H * W;
    const unsigned grid = (unsigned)min((total + 255) / 256, (long)kNumSM * 32);
    if (src_is_u8)
        tcx_s2d_kernel<uint8_t><<<grid, 256, 0, st>>>(out, out_lo, reinterpret_cast<const uint8_t*>(frames), idx, B, C, H, W, S);
    else
        tcx_s2d_kernel<float><<<grid, 256, 0, st>>>(out, out_lo, reinterpret_cast<const float*>(frames), idx, B, C, H, W, S);
    return check_launch("tcx_s2d");
}

int tcx_colsum_blocks(long rows) { return (int)min((rows + 255) / 256, (long)kNumSM * 2); }

int tcx_colsum(float* part, const float* x, long rows, int C, long pitch, cudaStream_t st) {
    A2CB_REQUIRE(part && x && rows > 0 && C > 0, "tcx_colsum: bad arguments");
    const int blocks = tcx_colsum_blocks(rows);
    const long per = (rows + blocks - 1) / blocks;
    tcx_colsum_kernel<<<blocks, 256, 0, st>>>(part, x, rows, C, pitch, per);
    return check_launch("tcx_colsum");
}

}  // namespace a2cb
