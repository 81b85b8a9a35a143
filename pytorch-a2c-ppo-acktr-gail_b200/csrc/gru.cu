// GRU recurrence of NNBase._forward_gru (reference a2c_ppo_acktr/model.py:111-166, torch.nn.GRU gate
// order r, z, n):
//     hm_t = h_{t-1} * mask_t
//     r = sigmoid(gi_r + gh_r);  z = sigmoid(gi_z + gh_z);  n = tanh(gi_n + r * gh_n)
//     h_t = (1 - z) * n + z * hm_t
// The reference re-launches cuDNN per done-free segment after a host sync (model.py:129-157); stepping
// every t with h * mask[t] is the same recurrence, so the masking lives in the
// gate kernels and no host round trip is needed.  The two matmuls per step run on the GEMM engine:
//   gi = x W_ih^T + b_ih   for ALL T*N rows at once (one GEMM),
//   gh = hm_t W_hh^T + b_hh per step (the only sequential GEMM),
// and these element-wise kernels fuse everything else, forward and backward-through-time.
#include "common.cuh"
#include "decls.cuh"
#include "gru.cuh"

namespace a2cb {

__device__ __forceinline__ float sigmoidf_(float x) { return 1.f / (1.f + expf(-x)); }

__device__ __forceinline__ float row_mask(const GruDesc& d, int t, int n) {
    const long row = (long)t * d.N + n;
    const long src = d.idx ? (long)d.idx[row] : row;
    return d.masks[src];
}

// dst[n, :] = src[n, :] * mask[0, n]      (hm_0 = hxs * masks[0])
__global__ void __launch_bounds__(256)
gru_init_kernel(const GruDesc d) {
    const long e = (long)blockIdx.x * blockDim.x + threadIdx.x;
    if (e >= (long)d.N * d.H) return;
    const int n = (int)(e / d.H);
    d.hm[e] = d.h0[e] * row_mask(d, 0, n);
}

__global__ void __launch_bounds__(256)
gru_fwd_kernel(const GruDesc d) {
    const long e = (long)blockIdx.x * blockDim.x + threadIdx.x;
    const int H = d.H;
    if (e >= (long)d.N * H) return;
    const int n = (int)(e / H), k = (int)(e - (long)n * H);
    const int t = d.t;
    const long row = (long)t * d.N + n;
    const float* gi = d.gi + row * 3 * H;
    float* gh = d.gh + row * 3 * H;
    float ghr, ghz, ghn;
    if (d.gh_part) {
        // fold the split-K partials of this step's hm_t W_hh^T (fixed order) and add b_hh
        ghr = d.gh_bias[k]; ghz = d.gh_bias[H + k]; ghn = d.gh_bias[2 * H + k];
        const long stride = (long)d.N * 3 * H;
        const float* p = d.gh_part + (long)n * 3 * H;
        float sr = 0.f, sz = 0.f, sn = 0.f;
        for (int s = 0; s < d.gh_splits; ++s, p += stride) { sr += p[k]; sz += p[H + k]; sn += p[2 * H + k]; }
        ghr += sr; ghz += sz; ghn += sn;
        gh[k] = ghr; gh[H + k] = ghz; gh[2 * H + k] = ghn;       // kept for the backward pass
    } else {
        ghr = gh[k]; ghz = gh[H + k]; ghn = gh[2 * H + k];
    }
    const float hm = d.hm[row * H + k];
    const float r = sigmoidf_(gi[k] + ghr);
    const float z = sigmoidf_(gi[H + k] + ghz);
    const float nn = tanhf(gi[2 * H + k] + r * ghn);
    const float h = (1.f - z) * nn + z * hm;
    d.out[row * H + k] = h;
    if (d.r) { d.r[row * H + k] = r; d.z[row * H + k] = z; d.n[row * H + k] = nn; }
    if (t + 1 < d.T) d.hm[(row + d.N) * H + k] = h * row_mask(d, t + 1, n);
}

// Backward of step t.  dcarry holds d(loss)/d(hm_{t+1}) on entry (ignored at t = T-1) and
// dh * z on exit; the engine then accumulates dgh_t W_hh onto it.
__global__ void __launch_bounds__(256)
gru_bwd_kernel(const GruDesc d) {
    const long e = (long)blockIdx.x * blockDim.x + threadIdx.x;
    const int H = d.H;
    if (e >= (long)d.N * H) return;
    const int n = (int)(e / H), k = (int)(e - (long)n * H);
    const int t = d.t;
    const long row = (long)t * d.N + n;
    float dh = d.dout[row * H + k];
    if (t + 1 < d.T) {
        float c = d.dcarry[e];
        if (d.dc_part) {
            const long stride = (long)d.N * H;
            float s = 0.f;
            for (int q = 0; q < d.dc_splits; ++q) s += d.dc_part[q * stride + e];
            c += s;
        }
        dh += c * row_mask(d, t + 1, n);
    }
    const float r = d.r[row * H + k], z = d.z[row * H + k], nn = d.n[row * H + k];
    const float hm = d.hm[row * H + k];
    const float ghn = d.gh[row * 3 * H + 2 * H + k];
    const float dn_pre = dh * (1.f - z) * (1.f - nn * nn);
    const float dz_pre = dh * (hm - nn) * z * (1.f - z);
    const float dr_pre = dn_pre * ghn * r * (1.f - r);
    float* dgi = d.dgi + row * 3 * H;
    float* dgh = d.dgh + row * 3 * H;
    dgi[k] = dr_pre; dgi[H + k] = dz_pre; dgi[2 * H + k] = dn_pre;
    dgh[k] = dr_pre; dgh[H + k] = dz_pre; dgh[2 * H + k] = dn_pre * r;
    d.dcarry[e] = dh * z;
}

int gru_op(int which, const GruDesc& d, cudaStream_t st) {
    A2CB_REQUIRE(d.N > 0 && d.H > 0 && d.T > 0 && d.t >= 0 && d.t < d.T, "gru: bad shape (N=%d H=%d T=%d t=%d)", d.N, d.H, d.T, d.t);
    A2CB_REQUIRE(d.masks && d.hm, "gru: null pointer");
    const unsigned grid = (unsigned)(((long)d.N * d.H + 255) / 256);
    if (which == 0) {
        A2CB_REQUIRE(d.h0, "gru_init: h0");
        gru_init_kernel<<<grid, 256, 0, st>>>(d);
        return check_launch("gru_init");
    }
    if (which == 1) {
        A2CB_REQUIRE(d.gi && d.gh && d.out, "gru_fwd: null pointer");
        gru_fwd_kernel<<<grid, 256, 0, st>>>(d);
        return check_launch("gru_fwd");
    }
    A2CB_REQUIRE(d.gh && d.r && d.z && d.n && d.dout && d.dgi && d.dgh && d.dcarry, "gru_bwd: null pointer");
    gru_bwd_kernel<<<grid, 256, 0, st>>>(d);
    return check_launch("gru_bwd");
}

}  // namespace a2cb
