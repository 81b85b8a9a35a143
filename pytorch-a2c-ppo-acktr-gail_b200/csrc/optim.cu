// Gradient-norm clipping and optimiser steps over the FLAT parameter buffer.
//
// Replaces nn.utils.clip_grad_norm_ + optim.Adam.step (reference algo/ppo.py:79-84),
// + optim.RMSprop.step (algo/a2c_acktr.py:70-78) and the inner optim.SGD of KFAC
// (algo/kfac.py:139-142,241).  Arithmetic follows torch 2.11's single-tensor rules
// (torch/optim/adam.py, rmsprop.py, sgd.py; torch/nn/utils/clip_grad.py:165-182):
//   coef = min(1, max_norm / (||g||_2 + 1e-6));  g *= coef   (always multiplied)
//   Adam:    m += (g - m)(1 - b1);  v = b2 v + (1 - b2) g g;
//            p -= (lr / (1 - b1^t)) * m / (sqrt(v) / sqrt(1 - b2^t) + eps)
//   RMSprop: s = a s + (1 - a) g g;  p -= lr * g / (sqrt(s) + eps)
//   SGD-m:   buf = g (first step) | mom * buf + g;  p -= lr * buf
// Two launches per step instead of torch's ~10 foreach kernels: a deterministic
// two-level norm reduction, then one fused element-wise update that also writes the
// clipped gradient back (p.grad after update() matches the reference).
#include "common.cuh"
#include "decls.cuh"

namespace a2cb {

__global__ void __launch_bounds__(256)
grad_sqnorm_kernel(const float* __restrict__ g, long n, double* __restrict__ partials,
                   float* __restrict__ norm_out, unsigned* counter) {
    __shared__ double red[32];
    __shared__ bool is_last;
    double s = 0.0;
    const long n4 = n / 4;
    const float4* g4 = reinterpret_cast<const float4*>(g);
    for (long e = (long)blockIdx.x * blockDim.x + threadIdx.x; e < n4; e += (long)gridDim.x * blockDim.x) {
        const float4 v = __ldg(g4 + e);
        s += (double)v.x * v.x + (double)v.y * v.y + (double)v.z * v.z + (double)v.w * v.w;
    }
    if (blockIdx.x == 0)
        for (long e = n4 * 4 + threadIdx.x; e < n; e += blockDim.x) s += (double)g[e] * g[e];
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
        norm_out[0] = (float)sqrt(S);
        norm_out[1] = (float)S;
        *counter = 0;
    }
}

constexpr int kNormBlocks = 148;

int grad_norm(const float* g, long n, double* scratch, float* norm_out, cudaStream_t st) {
    A2CB_REQUIRE(g && scratch && norm_out && n >= 0, "grad_norm: bad argument");
    A2CB_REQUIRE(((uintptr_t)g & 15) == 0, "grad_norm: gradient buffer must be 16-byte aligned");
    long want = (n / 4 + 255) / 256;
    int grid = (int)(want < 1 ? 1 : (want > kNormBlocks ? kNormBlocks : want));
    unsigned* counter = reinterpret_cast<unsigned*>(scratch + kNormBlocks);
    grad_sqnorm_kernel<<<grid, 256, 0, st>>>(g, n, scratch, norm_out, counter);
    return check_launch("grad_norm");
}

struct OptimArgs {
    float lr_eff;        // Adam: lr / bias_correction1; RMSprop / SGD: lr
    float eps;
    float b1, b2;        // Adam betas | RMSprop: b2 = alpha | SGD: b1 = momentum
    float omb1, omb2;    // (1 - b1), (1 - b2) rounded from double on the host, as torch passes them
    float bc2_sqrt;      // Adam: sqrt(1 - b2^t)
    float max_norm;      // <= 0: no clipping
    int first_step;      // SGD momentum: buf = g
    const float* dyn;    // optional device scalars {lr_eff, bc2_sqrt, first_step} overriding the fields
                         // above, so that a CUDA graph of many optimiser steps needs no re-capture
};

template <int KIND>   // 0 Adam, 1 RMSprop, 2 SGD-momentum
__global__ void __launch_bounds__(256)
optim_step_kernel(float* __restrict__ p, float* __restrict__ g, float* __restrict__ s1,
                  float* __restrict__ s2, long n, const float* __restrict__ norm, OptimArgs a) {
    if (a.dyn) { a.lr_eff = a.dyn[0]; a.bc2_sqrt = a.dyn[1]; a.first_step = a.dyn[2] != 0.f; }
    float coef = 1.f;
    if (a.max_norm > 0.f) coef = fminf(1.f, a.max_norm / (norm[0] + 1e-6f));
    else if (a.max_norm < 0.f) coef = norm[0];              // KFAC: norm[0] holds nu (kfac.py:234-239)
    const long e = (long)blockIdx.x * blockDim.x + threadIdx.x;
    if (e >= n) return;
    float gi = g[e];
    if (a.max_norm != 0.f) { gi = __fmul_rn(gi, coef); g[e] = gi; }
    float pi = p[e];
    if (KIND == 0) {
        float m = s1[e], v = s2[e];
        m = __fadd_rn(m, __fmul_rn(__fsub_rn(gi, m), a.omb1));              // lerp_
        v = __fadd_rn(__fmul_rn(v, a.b2), __fmul_rn(__fmul_rn(gi, gi), a.omb2));
        const float denom = __fadd_rn(__fdiv_rn(__fsqrt_rn(v), a.bc2_sqrt), a.eps);
        pi = __fadd_rn(pi, __fmul_rn(-a.lr_eff, __fdiv_rn(m, denom)));
        s1[e] = m; s2[e] = v;
    } else if (KIND == 1) {
        float sq = s1[e];
        sq = __fadd_rn(__fmul_rn(sq, a.b2), __fmul_rn(__fmul_rn(gi, gi), a.omb2));
        const float denom = __fadd_rn(__fsqrt_rn(sq), a.eps);
        pi = __fadd_rn(pi, __fmul_rn(-a.lr_eff, __fdiv_rn(gi, denom)));
        s1[e] = sq;
    } else {
        float buf = a.first_step ? gi : __fadd_rn(__fmul_rn(s1[e], a.b1), gi);
        pi = __fadd_rn(pi, __fmul_rn(-a.lr_eff, buf));
        s1[e] = buf;
    }
    p[e] = pi;
}

int optim_step(int kind, float* p, float* g, float* s1, float* s2, long n, const float* norm,
               float lr_eff, float eps, float b1, float b2, float omb1, float omb2, float bc2_sqrt,
               float max_norm, int first_step, const float* dyn, cudaStream_t st) {
    A2CB_REQUIRE(p && g && n >= 0, "optim_step: null pointer");
    A2CB_REQUIRE(kind >= 0 && kind <= 2, "optim_step: kind");
    A2CB_REQUIRE(s1 && (kind != 0 || s2), "optim_step: missing optimiser state");
    A2CB_REQUIRE(max_norm == 0.f || norm, "optim_step: clipping needs the norm buffer");
    if (n == 0) return A2CB_OK;
    OptimArgs a{lr_eff, eps, b1, b2, omb1, omb2, bc2_sqrt, max_norm, first_step, dyn};
    const unsigned grid = (unsigned)((n + 255) / 256);
    if (kind == 0) optim_step_kernel<0><<<grid, 256, 0, st>>>(p, g, s1, s2, n, norm, a);
    else if (kind == 1) optim_step_kernel<1><<<grid, 256, 0, st>>>(p, g, s1, s2, n, norm, a);
    else optim_step_kernel<2><<<grid, 256, 0, st>>>(p, g, s1, s2, n, norm, a);
    return check_launch("optim_step");
}

}  // namespace a2cb
