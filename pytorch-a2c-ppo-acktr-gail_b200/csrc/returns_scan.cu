// Reverse scan over [num_steps x num_processes] for GAE / discounted returns.
//
// Replaces the per-time-step Python loop of the reference
// (a2c_ppo_acktr/storage.py:66-105): four variants selected by
// (use_gae, use_proper_time_limits).  Layout is the reference's: every tensor is
// [T(+1), N, 1] fp32 row-major, element (t, n) at t*N + n.
//
// Two schedules, one arithmetic:
//  * env-parallel ("sequential" in time): one thread owns V in {1,4} adjacent
//    environments and walks t = T-1 .. 0.  Every fp32 operation is issued in the
//    reference's order with explicit round-to-nearest intrinsics (no FMA
//    contraction), so the result is bit-identical to the torch CPU loop.  Loads
//    are coalesced (float4 when N % 4 == 0) and prefetched U steps ahead; this is
//    the HBM-bound schedule used when N is large.
//  * time-parallel: one warp owns one environment; each lane composes the affine
//    maps acc_t = a_t * acc_{t+1} + c_t of a contiguous chunk of time, a warp
//    shuffle scan composes the 32 chunk maps, then every lane replays its chunk
//    with the reference op order from its incoming carry.  Used when N is too
//    small to fill the machine (C1: T=2048, N=1).  Differs from the sequential
//    result only by the re-association of the carry (<= ~4e-6 relative).
#include "common.cuh"
#include "decls.cuh"

namespace a2cb {

template <bool GAE, bool PTL>
struct StepMath {
    // One reverse-time step with the reference's exact operation order.
    //   acc    : gae_{t+1} (GAE) or returns_{t+1} (non-GAE)
    //   vnext  : value_preds[t+1] (GAE only)
    // returns the value written to returns[t]; updates acc.
    __device__ __forceinline__ static float apply(float& acc, float vnext, float r, float v,
                                                  float m, float b, float g, float gl) {
        if (GAE) {
            // delta = r + gamma * V[t+1] * m - V[t]         storage.py:76-78 / 95-97
            float x = __fmul_rn(__fmul_rn(g, vnext), m);
            float delta = __fsub_rn(__fadd_rn(r, x), v);
            // gae = delta + gamma * lambda * m * gae        storage.py:79-80 / 98-99
            float y = __fmul_rn(__fmul_rn(gl, m), acc);
            acc = __fadd_rn(delta, y);
            if (PTL) acc = __fmul_rn(acc, b);              // storage.py:82
            return __fadd_rn(acc, v);                       // storage.py:83 / 100
        } else {
            float x = __fmul_rn(__fmul_rn(acc, g), m);
            float y = __fadd_rn(x, r);                      // storage.py:104-105
            if (PTL) {
                // (...) * bad + (1 - bad) * V[t]            storage.py:87-89
                y = __fadd_rn(__fmul_rn(y, b), __fmul_rn(__fsub_rn(1.0f, b), v));
            }
            acc = y;
            return y;
        }
    }
    // Affine form acc_t = a * acc_{t+1} + c of the same step (time-parallel carry only).
    __device__ __forceinline__ static void affine(float vnext, float r, float v, float m, float b,
                                                  float g, float gl, float& a, float& c) {
        if (GAE) {
            float delta = (r + (g * vnext) * m) - v;
            a = gl * m;
            c = delta;
            if (PTL) { a *= b; c *= b; }
        } else {
            a = g * m;
            c = r;
            if (PTL) { a *= b; c = r * b + (1.0f - b) * v; }
        }
    }
};

template <int V> struct FVec { float v[V]; };

template <int V>
__device__ __forceinline__ FVec<V> ldv(const float* p) {
    FVec<V> o;
    if (V == 4) {
        float4 t = __ldg(reinterpret_cast<const float4*>(p));
        o.v[0] = t.x; o.v[1 % V] = t.y; o.v[2 % V] = t.z; o.v[3 % V] = t.w;
    } else {
#pragma unroll
        for (int i = 0; i < V; ++i) o.v[i] = __ldg(p + i);
    }
    return o;
}
template <int V>
__device__ __forceinline__ void stv(float* p, const FVec<V>& o) {
    if (V == 4) {
        float4 t = make_float4(o.v[0], o.v[1 % V], o.v[2 % V], o.v[3 % V]);
        __stcs(reinterpret_cast<float4*>(p), t);
    } else {
#pragma unroll
        for (int i = 0; i < V; ++i) p[i] = o.v[i];
    }
}

// ---------------------------------------------------------------------------
// env-parallel schedule
// ---------------------------------------------------------------------------
template <bool GAE, bool PTL, int V, int U>
__global__ void __launch_bounds__(128)
returns_scan_env_kernel(const float* __restrict__ rewards, float* value_preds, float* returns,
                        const float* __restrict__ masks, const float* __restrict__ bad_masks,
                        const float* __restrict__ next_value, int T, long N, float g, float gl) {
    const long n0 = ((long)blockIdx.x * blockDim.x + threadIdx.x) * V;
    if (n0 >= N) return;
    constexpr bool NEED_V = GAE || PTL;

    FVec<V> acc, vnext;
    {
        FVec<V> nv = ldv<V>(next_value + n0);
        if (GAE) {
            stv<V>(value_preds + (long)T * N + n0, nv);   // storage.py:74 / 93
            vnext = nv;
#pragma unroll
            for (int i = 0; i < V; ++i) acc.v[i] = 0.0f;
        } else {
            stv<V>(returns + (long)T * N + n0, nv);       // storage.py:85 / 102
            acc = nv;
            vnext = nv;
        }
    }

    for (int t0 = T - 1; t0 >= 0; t0 -= U) {
        FVec<V> r[U], v[U], m[U], b[U];
#pragma unroll
        for (int u = 0; u < U; ++u) {
            const int t = t0 - u;
            if (t >= 0) {
                const long o = (long)t * N + n0;
                r[u] = ldv<V>(rewards + o);
                if (NEED_V) v[u] = ldv<V>(value_preds + o);
                m[u] = ldv<V>(masks + o + N);
                if (PTL) b[u] = ldv<V>(bad_masks + o + N);
            }
        }
#pragma unroll
        for (int u = 0; u < U; ++u) {
            const int t = t0 - u;
            if (t >= 0) {
                FVec<V> out;
#pragma unroll
                for (int i = 0; i < V; ++i) {
                    const float vv = NEED_V ? v[u].v[i] : 0.0f;
                    const float bb = PTL ? b[u].v[i] : 1.0f;
                    out.v[i] = StepMath<GAE, PTL>::apply(acc.v[i], vnext.v[i], r[u].v[i], vv,
                                                         m[u].v[i], bb, g, gl);
                    vnext.v[i] = vv;
                }
                stv<V>(returns + (long)t * N + n0, out);
            }
        }
    }
}

// ---------------------------------------------------------------------------
// time-parallel schedule: one warp per environment
// ---------------------------------------------------------------------------
template <bool GAE, bool PTL>
__global__ void __launch_bounds__(128)
returns_scan_time_kernel(const float* __restrict__ rewards, float* value_preds, float* returns,
                         const float* __restrict__ masks, const float* __restrict__ bad_masks,
                         const float* __restrict__ next_value, int T, long N, float g, float gl) {
    const long n = ((long)blockIdx.x * blockDim.x + threadIdx.x) >> 5;
    const int lane = threadIdx.x & 31;
    if (n >= N) return;
    constexpr bool NEED_V = GAE || PTL;

    const float nv = __ldg(next_value + n);
    if (lane == 0) {
        if (GAE) value_preds[(long)T * N + n] = nv;
        else returns[(long)T * N + n] = nv;
    }
    // reversed time s = T-1-t; lane owns s in [s_lo, s_hi)
    const int L = (T + 31) / 32;
    const int s_lo = min(lane * L, T), s_hi = min(s_lo + L, T);

    // pass 1: compose the chunk's affine map  acc_out = A * acc_in + C
    float A = 1.0f, C = 0.0f;
    for (int s = s_lo; s < s_hi; ++s) {
        const int t = T - 1 - s;
        const long o = (long)t * N + n;
        const float r = __ldg(rewards + o);
        const float v = NEED_V ? value_preds[o] : 0.0f;
        // value_preds[T] is being written by lane 0 in this kernel: use nv directly
        const float vn = GAE ? (t + 1 == T ? nv : value_preds[o + N]) : 0.0f;
        const float m = __ldg(masks + o + N);
        const float b = PTL ? __ldg(bad_masks + o + N) : 1.0f;
        float a, c;
        StepMath<GAE, PTL>::affine(vn, r, v, m, b, g, gl, a, c);
        A = a * A;
        C = a * C + c;
    }
    // inclusive scan of the maps across lanes (lane l := F_l o ... o F_0)
#pragma unroll
    for (int d = 1; d < 32; d <<= 1) {
        const float pA = __shfl_up_sync(0xffffffffu, A, d);
        const float pC = __shfl_up_sync(0xffffffffu, C, d);
        if (lane >= d) { C = A * pC + C; A = A * pA; }
    }
    const float acc0 = GAE ? 0.0f : nv;
    float carry = A * acc0 + C;                     // value after this lane's chunk
    carry = __shfl_up_sync(0xffffffffu, carry, 1);  // incoming = value after previous lane
    if (lane == 0) carry = acc0;

    // pass 2: replay the chunk with the reference op order
    float acc = carry;
    for (int s = s_lo; s < s_hi; ++s) {
        const int t = T - 1 - s;
        const long o = (long)t * N + n;
        const float r = __ldg(rewards + o);
        const float v = NEED_V ? value_preds[o] : 0.0f;
        const float vn = GAE ? (t + 1 == T ? nv : value_preds[o + N]) : 0.0f;
        const float m = __ldg(masks + o + N);
        const float b = PTL ? __ldg(bad_masks + o + N) : 1.0f;
        returns[o] = StepMath<GAE, PTL>::apply(acc, vn, r, v, m, b, g, gl);
    }
}

template <bool GAE, bool PTL>
static int launch_scan(const float* rewards, float* value_preds, float* returns, const float* masks,
                       const float* bad_masks, const float* next_value, int T, long N, float g,
                       float gl, int schedule, cudaStream_t st) {
    if (schedule == 2) {
        const long threads = N * 32;
        const int block = 128;
        const long grid = (threads + block - 1) / block;
        returns_scan_time_kernel<GAE, PTL><<<(unsigned)grid, block, 0, st>>>(
            rewards, value_preds, returns, masks, bad_masks, next_value, T, N, g, gl);
    } else {
        const bool vec4 = (N % 4 == 0) && N >= 4 * 128 &&
                          ((((uintptr_t)rewards | (uintptr_t)value_preds | (uintptr_t)returns |
                             (uintptr_t)masks | (uintptr_t)bad_masks | (uintptr_t)next_value) & 15) == 0);
        const int block = 128;
        if (vec4) {
            const long grid = (N / 4 + block - 1) / block;
            returns_scan_env_kernel<GAE, PTL, 4, 4><<<(unsigned)grid, block, 0, st>>>(
                rewards, value_preds, returns, masks, bad_masks, next_value, T, N, g, gl);
        } else {
            const long grid = (N + block - 1) / block;
            returns_scan_env_kernel<GAE, PTL, 1, 8><<<(unsigned)grid, block, 0, st>>>(
                rewards, value_preds, returns, masks, bad_masks, next_value, T, N, g, gl);
        }
    }
    return check_launch("returns_scan");
}

int returns_scan(const float* rewards, float* value_preds, float* returns, const float* masks,
                 const float* bad_masks, const float* next_value, int T, long N, double gamma,
                 double gae_lambda, int use_gae, int use_ptl, int schedule, cudaStream_t st) {
    // gamma * gae_lambda is formed in Python double precision before it meets the
    // fp32 tensor (storage.py:79), so round once, here.
    const float g = (float)gamma;
    const float gl = (float)(gamma * gae_lambda);
    if (schedule == 0) {
        // auto: the time-parallel schedule wins while one-thread-per-env cannot fill
        // 148 SMs; from ~16k envs upwards the coalesced env-parallel walk is HBM-bound.
        schedule = (N < 16384 && T >= 16) ? 2 : 1;
    }
    if (use_gae) {
        if (use_ptl) return launch_scan<true, true>(rewards, value_preds, returns, masks, bad_masks, next_value, T, N, g, gl, schedule, st);
        return launch_scan<true, false>(rewards, value_preds, returns, masks, bad_masks, next_value, T, N, g, gl, schedule, st);
    }
    if (use_ptl) return launch_scan<false, true>(rewards, value_preds, returns, masks, bad_masks, next_value, T, N, g, gl, schedule, st);
    return launch_scan<false, false>(rewards, value_preds, returns, masks, bad_masks, next_value, T, N, g, gl, schedule, st);
}

}  // namespace a2cb
