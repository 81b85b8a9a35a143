// Descriptors of the TMA-fed tensor-core engine (tcx.cu); mirrored by a2cb_tcx_*_t in include/a2cb.h.
#pragma once
#include <stdint.h>

namespace a2cb {

constexpr int TCX_MAX_KB = 64;      // k-blocks (32 reduction elements each) of a forward-like op
constexpr int TCX_MAX_G = 8;        // accumulator groups of a weight-gradient op

// A strided 5-D view of an fp32 tensor, innermost dimension first.  dim[0] is the channel axis
// (stride 1); box[0] is always 32 floats = one 128-byte swizzle row.
struct TcxView {
    const float* x;                 // plain fp32 values; doubles as the TF32 "hi" operand (the tensor core
                                    // ignores the 13 low mantissa bits)
    const float* lo;                // x - trunc_tf32(x), or NULL when x is exact in TF32 (uint8 frames)
    int64_t dim[5];
    int64_t stride[5];              // in elements
    int32_t box[5];
    int32_t pad_;
};

// tile index t -> (tq, tr) = (t / t_div, t % t_div); base[tr_dim] += tr * tr_step; base[tq_dim] += tq * tq_step
struct TcxTiling {
    int32_t n_tiles, t_div, tr_dim, tr_step, tq_dim, tq_step;
};

// out[o_base + sum_d (base_d + i_d) * o_stride[d] + n] = epi(alpha * sum_k A[...] * B[n][k])
struct TcxFwd {
    TcxView A;
    const float* B;                 // weight image [b_rows = N][b_cols = Ktot], K contiguous
    const float* B_lo;
    int64_t b_rows, b_cols;
    int32_t n_kb;
    int32_t chain;                  // k-blocks per TMEM accumulation chain (<= 0: one chain)
    int16_t kb_a[TCX_MAX_KB][4];    // per k-block coordinate offsets of dims 0..3 of A
    int32_t kb_b[TCX_MAX_KB];       // per k-block column offset into B
    TcxTiling tiles;
    float* out;
    float* out_lo;                  // optional lo plane of the output
    const float* mask;              // optional: out = mask > 0 ? out : 0 (same indexing as out)
    const float* bias;              // optional [N]
    int64_t o_stride[5];            // element strides of out for box dims 1..4 (index 0 unused)
    int64_t o_base;
    int32_t o_ext[5];               // valid extents of (base_d + i_d) for box dims 1..4
    int32_t N;                      // output channels (multiple of 4)
    float alpha;
    int32_t act;                    // 0 none, 1 relu
    // optional column blocks: output column n belongs to block n / o_cb and is stored at channel n % o_cb of an
    // output shifted by o_cb_off[block] (the four stride-parity classes of a transposed stride-2 convolution
    // share their input boxes, so they run as ONE product with 4 x C_in columns).  o_cb = 0: off.
    int32_t o_cb;
    // optional image gather: the A coordinate of dimension gather_dim becomes gather_idx[coordinate] (rows of a
    // minibatch inside a tensor that holds the whole rollout); output coordinates are not affected.  0 = off.
    int32_t gather_dim;
    int64_t o_cb_off[4];
    const int64_t* gather_idx;
    // optional: column sums of the stored values (= bias gradient of the layer that produced this gradient), one
    // partial row [N] per CTA at colsum_part[blockIdx.x * N + n]; single column tile only (N <= BN)
    float* colsum_part;
    // bf16 mode (operands exactly representable in bf16, e.g. uint8 frames): A.x and the weight image are bf16; the
    // image has THREE planes B, B_lo, B3 (w = b1 + b2 + b3 to 24 bits), k-blocks are 64 channels (128-byte rows) and
    // every product costs three kind::f16 MMAs at twice the TF32 rate with half the operand bytes.
    const float* B3;
    int32_t bf16;
    // patch mode (bf16 mode only; 0 = off): A.box is the input PATCH of a tile (tile pixels + tap halo), kb_a[kb][1..2]
    // are the tap shifts (dx, dy) inside it, and `patch` = accumulator rows to store (patch pixels enumerated box[1] per
    // row; pixels past o_ext are junk).  One TMA box per tile instead of one per tap.
    int32_t patch;
};

// partial[split][row_off[m] + n * col_stride] = alpha * sum_rows A[row, m] * Bm[row, n]
// Accumulator row m = g * 128 + 32 * j + c: channel c of box j of group g.
struct TcxWgrad {
    TcxView A;                      // operand whose boxes differ per group (input activations under tap shifts)
    TcxView Bm;                     // operand shared by all groups (output gradient)
    int32_t G;                      // accumulator groups per CTA
    int32_t nbb;                    // B boxes (32 channels each); BN = 32 * nbb
    int16_t ga[TCX_MAX_G][4][4];    // coordinate offsets (dims 0..3) of box j of group g
    int16_t gb[8][4];               // coordinate offsets of B box j
    TcxTiling rows;                 // row tiles (the reduction); n_tiles of them
    int32_t splits;                 // row-tile ranges -> grid.x
    int32_t chain;                  // row tiles per accumulation chain
    int32_t ot_a, ot_b;             // output tiles along A channels / B channels -> grid.y = ot_a * ot_b
    int32_t ot_a_step, ot_b_step;   // added to the dim-0 coordinate of A / B boxes per output tile
    float* partial;
    int64_t split_stride;
    const int32_t* row_off;         // [ot_a][G * 128] element offsets, -1 = unused row
    int64_t col_stride;
    int32_t n_cols;                 // valid B channels in total
    float alpha;
    const int64_t* gather_idx;      // optional image gather of the A boxes (see TcxFwd), dimension gather_dim
    int32_t gather_dim;
    int32_t pad_;
};

// hardware-semantics knobs of the MN-major path (host-overridable for probing)
struct TcxKnobs {
    int32_t mn_layout_type;         // UMMA smem descriptor layout type for MN-major fp32 operands
    int32_t mn_sbo_bytes;           // stride between K atoms
    int32_t mn_kstep_bytes;         // start-address advance per K = 8 step
    int32_t mn_tma_swizzle;         // CUtensorMapSwizzle used for MN-major boxes
};

}  // namespace a2cb
