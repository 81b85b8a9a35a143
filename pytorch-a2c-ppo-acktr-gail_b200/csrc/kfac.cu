// Small kernels of the ACKTR / KFAC update (reference a2c_ppo_acktr/algo/kfac.py).  The heavy parts --
// the factor products a^T a, g^T g (kfac.py:29-64) and the Kronecker preconditioning chain
// Q_g (Q_g^T dW Q_a / (d_g d_a^T + damping)) Q_a^T (kfac.py:216-227) -- are descriptors of the GEMM
// engine; the eigendecompositions stay a library call (torch.linalg.eigh -> cuSOLVER).  Here:
//   * spatial_sum:   s[b, c] = sum over pixels of a conv layer's output gradient (AddBias g-statistics,
//                    kfac.py:59-61), reading the zero-haloed NHWC gradient buffers in place;
//   * kfac_scale:    v[i, j] /= d_g[i] * d_a[j] + damping                              (kfac.py:223-225)
//   * kfac_nu:       vg = sum(v * grad) * lr^2 over the flat buffers (deterministic two-level sum),
//                    nu = min(1, sqrt(kl_clip / vg))                                    (kfac.py:229-234)
//   * u8_to_f32:     frames / 255 as fp32 (the B operand of the conv1 a-factor product)
#include "common.cuh"
#include "decls.cuh"

namespace a2cb {

__global__ void __launch_bounds__(256)
spatial_sum_kernel(float* __restrict__ out, const float* __restrict__ in, int B, int P, int C, int ow,
                   long img_stride, long pitch) {
    const long e = (long)blockIdx.x * blockDim.x + threadIdx.x;
    if (e >= (long)B * C) return;
    const int b = (int)(e / C), c = (int)(e - (long)b * C);
    const float* src = in + (long)b * img_stride + c;
    float s = 0.f;
    for (int p = 0; p < P; ++p) s += __ldg(src + (long)(p / ow) * pitch + (long)(p % ow) * C);
    out[e] = s;
}

int spatial_sum(float* out, const float* in, int B, int P, int C, int ow, long img_stride, long pitch,
                cudaStream_t st) {
    A2CB_REQUIRE(out && in && B >= 0 && P > 0 && C > 0 && ow > 0, "spatial_sum: bad argument");
    if (B == 0) return A2CB_OK;
    spatial_sum_kernel<<<(unsigned)(((long)B * C + 255) / 256), 256, 0, st>>>(out, in, B, P, C, ow, img_stride, pitch);
    return check_launch("spatial_sum");
}

__global__ void __launch_bounds__(256)
kfac_scale_kernel(float* __restrict__ v, const float* __restrict__ dg, const float* __restrict__ da, int rows,
                  int cols, float damping) {
    const long e = (long)blockIdx.x * blockDim.x + threadIdx.x;
    if (e >= (long)rows * cols) return;
    const int i = (int)(e / cols), j = (int)(e - (long)i * cols);
    const float a = da ? __ldg(da + j) : 1.f;
    v[e] = v[e] / (__ldg(dg + i) * a + damping);
}

int kfac_scale(float* v, const float* dg, const float* da, int rows, int cols, float damping, cudaStream_t st) {
    A2CB_REQUIRE(v && dg && rows > 0 && cols > 0, "kfac_scale: bad argument");
    kfac_scale_kernel<<<(unsigned)(((long)rows * cols + 255) / 256), 256, 0, st>>>(v, dg, da, rows, cols, damping);
    return check_launch("kfac_scale");
}

__global__ void __launch_bounds__(256)
kfac_nu_kernel(const float* __restrict__ v, const float* __restrict__ g, long n, double* __restrict__ partials,
               float* __restrict__ out, unsigned* counter, float lr, float kl_clip) {
    __shared__ double red[32];
    __shared__ bool is_last;
    double s = 0.0;
    for (long e = (long)blockIdx.x * blockDim.x + threadIdx.x; e < n; e += (long)gridDim.x * blockDim.x)
        s += (double)(v[e] * g[e] * lr * lr);
    s = block_sum(s, red);
    if (threadIdx.x == 0) {
        partials[blockIdx.x] = s;
        __threadfence();
        is_last = (atomicAdd(counter, 1u) == gridDim.x - 1);
    }
    __syncthreads();
    if (!is_last) return;
    __threadfence();
    if (threadIdx.x == 0) {
        double S = 0.0;
        for (unsigned k = 0; k < gridDim.x; ++k) S += partials[k];
        const double nu = fmin(1.0, sqrt((double)kl_clip / S));
        out[0] = (float)nu;
        out[1] = (float)S;
        *counter = 0;
    }
}

int kfac_nu(const float* v, const float* g, long n, double* scratch, float* out, float lr, float kl_clip,
            cudaStream_t st) {
    A2CB_REQUIRE(v && g && scratch && out && n > 0, "kfac_nu: bad argument");
    long want = (n + 255) / 256;
    const int grid = (int)(want > 148 ? 148 : want);
    unsigned* counter = reinterpret_cast<unsigned*>(scratch + 148);
    kfac_nu_kernel<<<grid, 256, 0, st>>>(v, g, n, scratch, out, counter, lr, kl_clip);
    return check_launch("kfac_nu");
}

__global__ void __launch_bounds__(256)
u8_to_f32_kernel(float* __restrict__ dst, const uint8_t* __restrict__ src, long n, float div) {
    const long e = (long)blockIdx.x * blockDim.x + threadIdx.x;
    if (e < n) dst[e] = (float)src[e] / div;
}
__global__ void __launch_bounds__(256)
f32_div_kernel(float* __restrict__ dst, const float* __restrict__ src, long n, float div) {
    const long e = (long)blockIdx.x * blockDim.x + threadIdx.x;
    if (e < n) dst[e] = src[e] / div;
}

int frames_to_f32(float* dst, const void* src, long n, int src_is_u8, float div, cudaStream_t st) {
    A2CB_REQUIRE(dst && src && n >= 0, "frames_to_f32: bad argument");
    if (n == 0) return A2CB_OK;
    const unsigned grid = (unsigned)((n + 255) / 256);
    if (src_is_u8) u8_to_f32_kernel<<<grid, 256, 0, st>>>(dst, (const uint8_t*)src, n, div);
    else f32_div_kernel<<<grid, 256, 0, st>>>(dst, (const float*)src, n, div);
    return check_launch("frames_to_f32");
}

}  // namespace a2cb
