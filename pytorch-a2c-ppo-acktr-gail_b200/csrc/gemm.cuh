// Descriptor of the affine-indexed GEMM engine (mirrors a2cb_gemm_t in include/a2cb.h).
#pragma once
#include <stdint.h>

namespace a2cb {

// offset(i) = q0 * s0 + q1 * s1 + q2 * s2   with  q0 = i / d1, q1 = (i % d1) / d2, q2 = i % d2
struct Affine3 {
    int64_t d1, d2;
    int64_t s0, s1, s2;
};

struct GemmDesc {
    const void* A;
    const float* B;
    float* C;
    float* C_ones;
    const float* bias;
    const float* mask;
    const int64_t* gather_row;   // replaces q0 of the ROW index in rowA
    const int64_t* gather_red;   // replaces q0 of the REDUCTION index in redA
    int64_t M, N, K;
    Affine3 rowA, rowC, rowM;
    Affine3 redA, redB;
    Affine3 colB, colC, colM;
    int32_t a_dtype;      // 0 fp32, 1 uint8 / 255.0f, 2 fp32 / 255.0f
    int32_t ones_row;     // logical extra row M of A == 1.0; outputs to C_ones[j * ones_stride]
    int64_t ones_stride;
    int32_t bias_mode;    // 0 none, 1 bias[j], 2 bias[i]
    int32_t act;          // 0 none, 1 relu, 2 tanh
    int32_t mask_mode;    // 0 none, 1 *= (mask > 0), 2 *= (1 - mask^2)
    int32_t accumulate;   // C += result
    float alpha;
    int32_t batch;        // 1..4 problems sharing the maps
    int64_t batch_offA[4], batch_offB[4], batch_offC[4], batch_offM[4];
    int32_t splits;       // reduction split across CTAs; split s writes at + s * split_stride
    int32_t vecA;         // 0 scalar, 1 four consecutive r contiguous & aligned, 2 four consecutive rows
    int64_t split_stride;
    int32_t vecB;         // 0 scalar, 1 four consecutive columns contiguous & aligned
    int32_t tile;         // 0 auto, else preferred BN (32 / 64 / 128)
    int32_t tc_chunk;     // tensor-core engine: k-blocks (of 32) per TMEM accumulation chain, 0 = default 8
    int32_t sym;          // symmetric product (SYRK): lower tiles skipped, upper tiles also store their mirror image
    const float* B_image; // optional: B pre-split (hi | lo) and pre-tiled in the UMMA smem layout by
                          // gemm_prep_b(); the tensor-core engine then fetches it with one TMA bulk copy
                          // per k-block instead of gathering + splitting it in every CTA
};

}  // namespace a2cb
