/*
 * a2cb.h -- C ABI of liba2cb.so, the sm_100a implementation of the rollout-batch
 * policy-update hot path of ikostrikov/pytorch-a2c-ppo-acktr-gail.
 *
 * The reference has no FFI of its own: its seam is the Python class API of the
 * `a2c_ppo_acktr` package.  Each entry point below names the
 * reference code (file:line under the reference tree) whose arithmetic it replaces;
 * the Python classes in pytorch-a2c-ppo-acktr-gail_b200/a2c_ppo_acktr/ bind these
 * through ctypes (see INTEGRATION.md for the binding a maintainer would add).
 *
 * Conventions
 *  - plain C types only: device pointers, sizes, a CUstream; no torch types;
 *  - every function returns 0 on success, a negative A2CB_ERR_* otherwise;
 *    a2cb_last_error() returns a thread-local description of the last failure;
 *  - caller owns all memory and workspaces; no hidden allocation and no host
 *    synchronisation unless a function's comment says so;
 *  - all floating tensors are fp32, row-major, in the reference's layouts.
 */
#ifndef A2CB_H_
#define A2CB_H_

#include <stdint.h>

#ifdef __cplusplus
extern "C" {
#endif

#define A2CB_ABI_VERSION 1

#define A2CB_OK 0
#define A2CB_ERR_INVALID (-1)     /* bad argument */
#define A2CB_ERR_CUDA (-2)        /* CUDA runtime / launch failure */
#define A2CB_ERR_UNSUPPORTED (-3) /* shape or mode not implemented */
#define A2CB_ERR_NCCL (-4)

typedef void* a2cb_stream_t; /* cudaStream_t */

int a2cb_abi_version(void);
const char* a2cb_last_error(void);
/* Number of kernels this library has launched in the calling process (bench.py's
 * "gpu_launches" evidence). */
int64_t a2cb_launch_count(void);
/* Bookkeeping for launches that happen through a replayed CUDA graph (captured launches are counted
 * when recorded; the host wrapper subtracts them once and adds them back per replay). */
void a2cb_add_launches(int64_t n);

/* ---------------------------------------------------------------------------
 * (1) RolloutStorage.compute_returns              a2c_ppo_acktr/storage.py:66-105
 *
 * rewards [T,N,1]; value_preds, returns, masks, bad_masks [T+1,N,1]; next_value [N,1].
 * Writes returns[0..T) and, as the reference does, value_preds[T] = next_value
 * (use_gae) or returns[T] = next_value (otherwise).
 * gamma / gae_lambda are doubles because the reference multiplies them in Python
 * double precision before rounding to fp32 (storage.py:79).
 * schedule: 0 auto, 1 env-parallel (bit-exact to the reference's operation order),
 *           2 time-parallel warp scan (<= 1e-5 relative).
 * ------------------------------------------------------------------------- */
int a2cb_returns_scan(const float* rewards, float* value_preds, float* returns, const float* masks,
                      const float* bad_masks, const float* next_value, int32_t T, int64_t N,
                      double gamma, double gae_lambda, int32_t use_gae,
                      int32_t use_proper_time_limits, int32_t schedule, a2cb_stream_t stream);

/* ---------------------------------------------------------------------------
 * (2) minibatch shuffle-gather                    a2c_ppo_acktr/storage.py:126-140
 *                                                 (recurrent: storage.py:154-199)
 * Copies rows bit-exactly.  Field f is a [rows, row_bytes] byte matrix.
 *   mode 0: out row j <- src row idx[j]                     (feed_forward_generator)
 *   mode 1: out row j <- src row (j / E) * N + idx[j % E]   (recurrent_generator,
 *           rows ordered t*E + e; idx holds E environment numbers)
 * idx is a device pointer to int64.
 * ------------------------------------------------------------------------- */
typedef struct {
    const void* src;
    void* dst;
    int64_t row_bytes;
} a2cb_gather_field_t;

int a2cb_gather_rows(const a2cb_gather_field_t* fields, int32_t n_fields, const int64_t* idx,
                     int64_t n_out_rows, int32_t mode, int64_t E, int64_t N, a2cb_stream_t stream);

/* ---------------------------------------------------------------------------
 * (3) affine-indexed GEMM engine: conv / linear / GRU forward, dgrad and wgrad
 *     reference call sites: a2c_ppo_acktr/model.py:176-180,185,190 (CNNBase),
 *     model.py:209-218 (MLPBase), model.py:113,153-155 (GRU), distributions.py:69-72,
 *     83-87 (heads) and their autograd backward; also kfac.py:29-64,216-227.
 *
 *   C[rowC(i) + colC(j)] = epi( alpha * sum_r A[rowA(i) + redA(r)] * B[redB(r) + colB(j)] )
 *
 * Each map is a 3-digit affine map: off(i) = q0*s0 + q1*s1 + q2*s2 with q0 = i / d1,
 * q1 = (i % d1) / d2, q2 = i % d2 (element units).  gather_row / gather_red optionally
 * replace q0 of rowA / redA by a lookup in a device int64 list (minibatch rows of the
 * rollout buffer), which fuses the observation gather (storage.py:127) and -- with
 * a_dtype = 1 -- the uint8 -> fp32 / 255 normalisation (model.py:190) into the load.
 * epi: + bias, activation, ReLU/tanh-derivative masking, optional accumulate.
 * With splits > 1 raw partial sums are written at C + s * split_stride for
 * a2cb_reduce_splits to fold.  ones_row appends a logical all-ones row M to A whose
 * outputs (column sums of B = bias gradients) go to C_ones[j * ones_stride].
 * ------------------------------------------------------------------------- */
typedef struct {
    int64_t d1, d2;
    int64_t s0, s1, s2;
} a2cb_affine3_t;

typedef struct {
    const void* A;
    const float* B;
    float* C;
    float* C_ones;
    const float* bias;
    const float* mask;
    const int64_t* gather_row;
    const int64_t* gather_red;
    int64_t M, N, K;
    a2cb_affine3_t rowA, rowC, rowM;
    a2cb_affine3_t redA, redB;
    a2cb_affine3_t colB, colC, colM;
    int32_t a_dtype;    /* 0 fp32, 1 uint8 / 255.0f, 2 fp32 / 255.0f */
    int32_t ones_row;
    int64_t ones_stride;
    int32_t bias_mode;  /* 0 none, 1 bias[j], 2 bias[i] */
    int32_t act;        /* 0 none, 1 relu, 2 tanh */
    int32_t mask_mode;  /* 0 none, 1 *= (mask > 0), 2 *= (1 - mask^2) */
    int32_t accumulate;
    float alpha;
    int32_t batch;      /* 1..4 problems sharing the maps, base offsets below */
    int64_t batch_offA[4], batch_offB[4], batch_offC[4], batch_offM[4];
    int32_t splits;
    int32_t vecA;       /* 0 scalar, 1 four consecutive r contiguous+aligned, 2 four consecutive rows */
    int64_t split_stride;
    int32_t vecB;       /* 0 scalar, 1 four consecutive columns contiguous+aligned, 2 four consecutive r */
    int32_t tile;       /* 0 auto, else SIMT tile: 32 / 64 / 128 (128 x tile) or 65 (64 x 64); 1000 / 1001 force engine */
    int32_t tc_chunk;   /* tensor-core engine: k-blocks (32 reduction indices each) accumulated in TMEM before
                           the partial sum is drained into fp32 registers (the tensor core's accumulator
                           truncates; short chains keep long, cancelling reductions exact enough).  0 = 8 */
    int32_t sym;        /* 1: the product is symmetric (M == N, B is A transposed through the maps: KFAC's a^T a and
                           g^T g, reference kfac.py:29-64): tiles strictly below the diagonal are skipped and the tiles
                           above it also store their mirror image C[rowC(j) + colC(i)] (SIMT engine only) */
    const float* B_image; /* optional image of B produced by a2cb_gemm_prep_b (weights: once per optimiser step);
                             the tensor-core engine then loads B with one TMA bulk copy per k-block */
} a2cb_gemm_t;

int a2cb_gemm(const a2cb_gemm_t* desc, a2cb_stream_t stream);
/* Writes desc->B_image: for every (batch, column tile, k-block) the B tile split into hi = tf32(b),
 * lo = tf32(b - hi) and laid out exactly as the tensor-core engine wants it in shared memory
 * (8-row x 16-byte core matrices, K-major).  a2cb_gemm_b_image_floats gives the buffer size. */
int a2cb_gemm_prep_b(const a2cb_gemm_t* desc, a2cb_stream_t stream);
int64_t a2cb_gemm_b_image_floats(const a2cb_gemm_t* desc);
/* Engine behind a2cb_gemm: 1 = tcgen05 tensor cores (3xTF32 split operands, chunked fp32 TMEM
 * accumulation) wherever supported; 0 = exact-fp32 SIMT engine; 2 (default) = per-descriptor choice
 * from measured crossover points.  Also A2CB_GEMM=simt|tc|auto, or per descriptor tile = 1000 (force
 * tensor cores) / 1001 (force SIMT). */
int a2cb_set_gemm_mode(int32_t mode);
int a2cb_get_gemm_mode(void);
/* dst[e] = (accumulate ? beta * dst[e] : 0) + act( sum_{s < splits} src[s * stride + e] + bias[e % ncols] ),
 * e in [0, count)  (bias may be NULL; act: 0 none, 1 relu, 2 tanh).  With beta = stat_decay this is also
 * the KFAC running average m <- decay * m + (1 - decay) * new (kfac.py:67-71). */
int a2cb_reduce_splits(float* dst, const float* src, int64_t count, int64_t stride, int32_t splits,
                       int32_t accumulate, const float* bias, int32_t ncols, int32_t act, float beta,
                       a2cb_stream_t stream);

/* ---------------------------------------------------------------------------
 * (3b) TMA-fed tensor-core engine: CNNBase / Linear forward, data and weight gradients
 *      reference call sites: a2c_ppo_acktr/model.py:176-180,185,190 (CNNBase), model.py:113
 *      (GRU input product) and their autograd backward.
 *
 * Operands are fp32 tensors stored as TWO planes, x and lo = x - trunc_tf32(x); the tensor core ignores
 * the 13 low mantissa bits of a kind::tf32 operand, so x itself is the "hi" operand and three MMAs
 * (x*w_hi + lo*w_hi + x*w_lo) reproduce the fp32 product to ~2^-20.  Activations are NHWC; a
 * convolution is a sum over filter taps of [pixels x C_in] x [C_in x C_out] products whose pixel box is one
 * multi-dimensional TMA box of a strided 5-D view (dimension 0 = channels, 32 per box).
 *
 *   a2cb_tcx_fwd:   out[o_base + sum_d (base_d + i_d) o_stride[d] + n] =
 *                       epi( alpha * sum_kb sum_{c<32} A[box(kb)][row i][c] * B[n][kb_b[kb] + c] )
 *                   epi: + bias[n], relu, relu' masking; optionally also writes the lo plane of out.
 *   a2cb_tcx_wgrad: partial[split][row_off[m] + n col_stride] = alpha * sum_rows A[row][m] * Bm[row][n]
 *                   (m = g 128 + 32 j + c: channel c of box j of accumulator group g), reduction over
 *                   row tiles split across CTAs; fold the splits with a2cb_reduce_splits.
 * ------------------------------------------------------------------------- */
#define A2CB_TCX_MAX_KB 64
#define A2CB_TCX_MAX_G 8
typedef struct {
    const float* x;
    const float* lo;      /* NULL: x is exact in TF32 (uint8 frames) */
    int64_t dim[5];       /* extents, innermost (channels) first */
    int64_t stride[5];    /* elements; stride[0] = 1 */
    int32_t box[5];       /* box[0] = 32 */
    int32_t pad_;
} a2cb_tcx_view_t;
/* tile t -> (tq, tr) = (t / t_div, t % t_div); base[tr_dim] += tr * tr_step; base[tq_dim] += tq * tq_step */
typedef struct { int32_t n_tiles, t_div, tr_dim, tr_step, tq_dim, tq_step; } a2cb_tcx_tiling_t;
typedef struct {
    a2cb_tcx_view_t A;
    const float* B;       /* weight image [b_rows = N][b_cols], reduction index contiguous (a2cb_tcx_prep_weights) */
    const float* B_lo;
    int64_t b_rows, b_cols;
    int32_t n_kb;
    int32_t chain;        /* k-blocks per TMEM accumulation chain (the tensor core's accumulator truncates) */
    int16_t kb_a[A2CB_TCX_MAX_KB][4];
    int32_t kb_b[A2CB_TCX_MAX_KB];
    a2cb_tcx_tiling_t tiles;
    float* out;
    float* out_lo;
    const float* mask;
    const float* bias;
    int64_t o_stride[5];
    int64_t o_base;
    int32_t o_ext[5];
    int32_t N;
    float alpha;
    int32_t act;
    int32_t o_cb;         /* column blocks: column n -> channel n % o_cb of an output shifted by o_cb_off[n / o_cb]; 0 = off */
    int32_t gather_dim;   /* image gather: A coordinate of this dimension -> gather_idx[coordinate]; 0 = off */
    int64_t o_cb_off[4];
    const int64_t* gather_idx; /* device list (minibatch rows inside a whole-rollout tensor, storage.py:127,181) */
    float* colsum_part;   /* optional [CTAs][N] partial column sums of the stored values (bias gradients); N <= 128 */
    const float* B3;      /* bf16 mode: third plane of the weight image */
    int32_t bf16;         /* 1: A.x and B / B_lo / B3 hold bf16 (exact operands, e.g. uint8 frames); 64-channel k-blocks */
    int32_t patch;        /* bf16 mode, stride-1 taps: > 0 = A.box is the tile's input PATCH (pixels + tap halo), loaded once
                             per tile; kb_a[kb][1..2] = tap shift inside it; value = accumulator rows to store */
} a2cb_tcx_fwd_t;
typedef struct {
    a2cb_tcx_view_t A;
    a2cb_tcx_view_t Bm;
    int32_t G, nbb;
    int16_t ga[A2CB_TCX_MAX_G][4][4];
    int16_t gb[8][4];
    a2cb_tcx_tiling_t rows;
    int32_t splits, chain;
    int32_t ot_a, ot_b, ot_a_step, ot_b_step;
    float* partial;
    int64_t split_stride;
    const int32_t* row_off;
    int64_t col_stride;
    int32_t n_cols;
    float alpha;
    const int64_t* gather_idx; /* image gather of the A boxes, see a2cb_tcx_fwd_t */
    int32_t gather_dim;
    int32_t pad_;
} a2cb_tcx_wgrad_t;
int a2cb_tcx_fwd(const a2cb_tcx_fwd_t* desc, a2cb_stream_t stream);
int a2cb_tcx_wgrad(const a2cb_tcx_wgrad_t* desc, a2cb_stream_t stream);
/* lo[i] = x[i] - trunc_tf32(x[i]), n a multiple of 4 */
int a2cb_tcx_lo_plane(float* lo, const float* x, int64_t n, a2cb_stream_t stream);
/* img[n][k] = src[ntab[n] + ktab[k]] (0 where a table entry is negative), img_lo = its lo plane: weights in
 * the reference's PyTorch layouts -> the reduction-contiguous images a2cb_tcx_fwd reads (once per optimiser
 * step).  ntab [N], ktab [K]: device int32 element offsets. */
int a2cb_tcx_prep_weights(float* img, float* img_lo, const float* src, int64_t N, int64_t K, const int32_t* ntab,
                          const int32_t* ktab, a2cb_stream_t stream);
/* Several weight images in one launch.  jobs: device array of a2cb_tcx_prep_job_t (first = index of the image's first
 * element in the concatenated element range of all jobs, ascending), total = sum of N K. */
typedef struct { float* img; float* img_lo; const float* src; const int32_t* ntab; const int32_t* ktab; int64_t N, K, first;
                 float* img3; int64_t bf16; /* bf16 = 1: img, img_lo, img3 are bf16 planes b1, b2, b3 with w = b1 + b2 + b3 */
} a2cb_tcx_prep_job_t;
int a2cb_tcx_prep_multi(const a2cb_tcx_prep_job_t* jobs, int32_t n_jobs, int64_t total, a2cb_stream_t stream);
/* frames [*, C, H, W] (uint8 or fp32; rows through idx, NULL = identity) -> fp32 space-to-depth
 * [B, H/S, W/S, C*S*S], channel (c, dy, dx): fuses the minibatch gather of storage.py:127 with the layout
 * change that turns the stride-S first convolution into a stride-1 one.  out_lo optional. */
int a2cb_tcx_s2d(float* out, float* out_lo, const void* frames, int32_t src_is_u8, const int64_t* idx, int32_t B,
                 int32_t C, int32_t H, int32_t W, int32_t S, a2cb_stream_t stream);
/* same with an additional bf16 copy of the output (Atari geometry only) */
int a2cb_tcx_s2d_b16(float* out, float* out_lo, void* out_b16, const void* frames, int32_t src_is_u8, const int64_t* idx,
                     int32_t B, int32_t C, int32_t H, int32_t W, int32_t S, a2cb_stream_t stream);
/* part[blk][c] = partial column sums of x[rows][C] (row pitch in elements); a2cb_tcx_colsum_blocks(rows)
 * partials, folded by a2cb_reduce_splits: bias gradients.  lo (optional): also writes the lo plane of x in the
 * same pass (the GRU's gate gradients need both). */
int a2cb_tcx_colsum(float* part, const float* x, float* lo, int64_t rows, int32_t C, int64_t pitch, a2cb_stream_t stream);
int a2cb_tcx_colsum_blocks(int64_t rows);
/* Weight + bias gradient of a narrow Linear layer (policy / value heads, distributions.py:69-72, model.py:185;
 * N <= 8 outputs): part[blk][n K + k] = sum_rows gy[row gy_stride + n] x[row K + k], part[blk][N K + n] = sum_rows gy;
 * a2cb_tcx_narrow_wgrad_blocks(rows) partial slabs of N K + N floats, folded by a2cb_reduce_splits. */
int a2cb_tcx_narrow_wgrad(float* part, const float* x, const float* gy, int64_t rows, int32_t K, int32_t N,
                          int32_t gy_stride, a2cb_stream_t stream);
int a2cb_tcx_narrow_wgrad_blocks(int64_t rows);
/* The same narrow layer forward, z[row z_stride + o] = bias[o] + sum_k x[row K + k] W[o K + k], and its data gradient
 * gx[row K + k] = sum_o dz[row dz_stride + o] W[o K + k] (optionally masked by relu'(mask) and with its lo plane). */
int a2cb_tcx_narrow_fwd(float* z, const float* x, const float* W, const float* bias, int64_t rows, int32_t K, int32_t N,
                        int32_t z_stride, a2cb_stream_t stream);
int a2cb_tcx_narrow_dgrad(float* gx, float* gx_lo, const float* dz, const float* W, const float* mask, int64_t rows,
                          int32_t K, int32_t N, int32_t dz_stride, a2cb_stream_t stream);
/* sizeof of a2cb_tcx_fwd_t (0), a2cb_tcx_wgrad_t (1), a2cb_tcx_view_t (2) as compiled: lets a binding check its mirror */
int a2cb_tcx_sizeof(int32_t which);
/* hardware-semantics knobs of the MN-major (weight-gradient) path, for probing: UMMA layout type, stride
 * between K atoms, start-address advance per K = 8 step, CUtensorMapSwizzle of the boxes */
int a2cb_tcx_set_knobs(int32_t mn_layout_type, int32_t mn_sbo_bytes, int32_t mn_kstep_bytes, int32_t mn_tma_swizzle);

/* ---------------------------------------------------------------------------
 * (3c) fused MLPBase                                   a2c_ppo_acktr/model.py:198-229
 * Both 2-layer tanh trunks (hidden width 64), critic_linear and the distribution head in one launch per direction;
 * a CTA takes 64 rows through the whole network with all weights and activations in shared memory (exact fp32).
 * forward: z[b] = (value, head outputs), saves the post-tanh activations; backward: dz -> every weight / bias
 * gradient of the trunks and heads (per-CTA slabs folded by a2cb_reduce_splits when B > 64).
 * ------------------------------------------------------------------------- */
typedef struct {
    const float* P;
    float* G;
    const float* obs;
    const int64_t* idx;
    float* acts;          /* [4][B][64]: actor l1, actor l2, critic l1, critic l2 */
    float* z;             /* [B][zs] */
    const float* dz;      /* [B][zs] */
    float* partial;       /* [ceil(B / 64)][span] */
    int64_t off_a0w, off_a0b, off_a2w, off_a2b, off_c0w, off_c0b, off_c2w, off_c2b, off_hw, off_hb;
    int64_t span;
    int32_t B, d_in, H, zs;
} a2cb_mlp_t;
int a2cb_mlp_fused(int32_t backward, const a2cb_mlp_t* desc, a2cb_stream_t stream);
int a2cb_mlp_fused_supported(int32_t d_in, int32_t H, int32_t zs);

/* ---------------------------------------------------------------------------
 * (4) fused loss epilogue                a2c_ppo_acktr/algo/ppo.py:35-37,61-77,86-88
 *                                         algo/a2c_acktr.py:48-51,55-64
 *                                         distributions.py:18-44, model.py:76-77
 * One thread per row: Categorical / DiagGaussian log-prob + entropy from the head outputs
 * z[b] = (value, logits[A] | mean[A]); PPO clipped surrogate + (clipped) value loss, or the
 * A2C / Fisher-sample losses; analytic gradient dz of
 *     total = c_v * value_loss + action_loss - c_e * entropy
 * and the three loss means (warp-shuffle + fixed-order block reduction, deterministic).
 * Rollout fields are read through the minibatch index list idx (NULL = identity), which
 * fuses the narrow-field gathers of storage.py:131-140.  inv_B = 1 / global minibatch size.
 * ------------------------------------------------------------------------- */
#define A2CB_MAX_ACTIONS 32
typedef struct {
    const float* z;
    const float* logstd;
    const int64_t* idx;
    const void* actions;      /* int64 [rows, act_stride] (Categorical) | fp32 [rows, act_stride] */
    const float* old_logp;
    const float* old_value;
    const float* returns;
    const float* adv;
    const float* noise;
    float* dz;
    float* dlogstd;
    float* logp_out;
    float* ent_out;
    float* loss_out;          /* [3] value_loss, action_loss, dist_entropy of this minibatch */
    float* loss_accum;        /* [3] running sums over minibatches */
    float* scratch;           /* a2cb_loss_scratch_floats(B) floats, zeroed once by the caller */
    int32_t B, A;
    int32_t z_stride, dz_stride;
    int32_t act_stride, dlogstd_stride;
    int32_t dist;             /* 0 Categorical, 1 DiagGaussian */
    int32_t mode;             /* 0 PPO, 1 A2C, 2 Fisher sample (ACKTR), 3 forward only,
                                 4 VJP: old_value / old_logp hold dL/dvalue, dL/dlogp per row, noise[0] = dL/d(mean entropy) */
    int32_t use_clipped_value_loss;
    float clip, c_v, c_e;
    float inv_B;
    int32_t n_valid;          /* rows [n_valid, B) are padding of a sharded minibatch; < 0: all valid */
    int32_t valid_mod;        /* > 0: row b is valid iff (b % valid_mod) < n_valid (recurrent, padded per env) */
    int32_t pad2_;
    float* dlogstd_rows;      /* optional [B, A]: per-row gradient w.r.t. the broadcast log-std (DiagGaussian) --
                                 the output gradient of the reference's logstd AddBias module, for its KFAC g-factor */
} a2cb_loss_t;

int a2cb_loss_epilogue(const a2cb_loss_t* desc, a2cb_stream_t stream);
int64_t a2cb_loss_scratch_floats(int32_t B);

/* advantage statistics (ppo.py:35-37): moments[0..2] = (sum, sum of squares, count) of
 * returns - value_preds over n rows, fp64 (all-reduce them across ranks when sharded);
 * scratch: 2*128 doubles + 1 counter word, zeroed once.  a2cb_adv_normalize then writes
 * adv = (x - mean) / (unbiased_std + 1e-5). */
int a2cb_adv_moments(const float* returns, const float* value_preds, int64_t n, double* scratch,
                     double* moments, a2cb_stream_t stream);
int a2cb_adv_normalize(const float* returns, const float* value_preds, int64_t n, const double* moments,
                       float* adv, a2cb_stream_t stream);

/* ---------------------------------------------------------------------------
 * clip_grad_norm_ + optimiser step over the flat parameter buffer
 *     a2c_ppo_acktr/algo/ppo.py:79-84, algo/a2c_acktr.py:70-78, algo/kfac.py:139-142,241
 * a2cb_grad_norm: norm_out[0] = ||g||_2, norm_out[1] = ||g||_2^2  (scratch: 148 doubles + counter).
 * a2cb_optim_step: kind 0 Adam, 1 RMSprop (b2 = alpha), 2 SGD with momentum (b1 = momentum).
 *   lr_eff = lr / (1 - b1^t) for Adam, lr otherwise; omb1/omb2 = 1-b1 / 1-b2 rounded from
 *   double; bc2_sqrt = sqrt(1 - b2^t).  max_norm > 0 rescales g by min(1, max_norm/(norm+1e-6))
 *   in place first (reading norm[0] on the device: no host sync); max_norm < 0 rescales g by
 *   norm[0] itself (the KFAC trust-region factor nu).
 * ------------------------------------------------------------------------- */
int a2cb_grad_norm(const float* grads, int64_t n, double* scratch, float* norm_out, a2cb_stream_t stream);
int a2cb_optim_step(int32_t kind, float* params, float* grads, float* state1, float* state2, int64_t n,
                    const float* norm, float lr_eff, float eps, float b1, float b2, float omb1, float omb2,
                    float bc2_sqrt, float max_norm, int32_t first_step, a2cb_stream_t stream);

/* ---------------------------------------------------------------------------
 * Op lists: one minibatch step of PPO.update / A2C_ACKTR.update (forward GEMMs, loss
 * epilogue, backward GEMMs, split-K folds, clip + optimiser) is a short list of the
 * descriptors above; a2cb_run_ops launches them back to back on `stream` from C, so the
 * per-minibatch Python cost is one call (ppo.py:51-88 is ~200 Python-dispatched kernels).
 * Ops flagged A2CB_OP_RUNTIME_IDX get their minibatch index list (gather_row / gather_red /
 * loss idx) replaced by the `idx` argument, so one list serves every minibatch of an epoch.
 * ------------------------------------------------------------------------- */
#define A2CB_OP_GEMM 1
#define A2CB_OP_REDUCE 2
#define A2CB_OP_LOSS 3
#define A2CB_OP_GRAD_NORM 4
#define A2CB_OP_OPTIM 5
#define A2CB_OP_FILL 6
#define A2CB_OP_GRU_INIT 7
#define A2CB_OP_GRU_FWD 8
#define A2CB_OP_GRU_BWD 9
#define A2CB_OP_SPATIAL_SUM 10
#define A2CB_OP_GEMM_PREP_B 11
#define A2CB_OP_GRU_SEQ_FWD 12
#define A2CB_OP_GRU_SEQ_BWD 13
#define A2CB_OP_TCX_FWD 14
#define A2CB_OP_TCX_WGRAD 15
#define A2CB_OP_TCX_LO 16
#define A2CB_OP_TCX_PREP 17
#define A2CB_OP_TCX_S2D 18
#define A2CB_OP_TCX_COLSUM 19
#define A2CB_OP_TCX_NARROW_WGRAD 20
#define A2CB_OP_TCX_PREP_MULTI 21
#define A2CB_OP_MLP_FWD 22
#define A2CB_OP_MLP_BWD 23
#define A2CB_OP_TCX_NARROW_FWD 24
#define A2CB_OP_TCX_NARROW_DGRAD 25
#define A2CB_OP_SAMPLE 26
#define A2CB_OP_COMM_FORK 27   /* a2cb_comm_op_t: all-reduce buf[0..count) on the side stream, ordered after the ops so far */
#define A2CB_OP_COMM_JOIN 28   /* a2cb_comm_op_t: the launching stream waits for every all-reduce forked since the last join */
#define A2CB_OP_RUNTIME_IDX 1

typedef struct { float* dst; const float* src; int64_t count; int64_t stride; int32_t splits; int32_t accumulate;
                 const float* bias; int32_t ncols; int32_t act; float beta; int32_t pad_; } a2cb_reduce_t;
typedef struct { const float* grads; int64_t n; double* scratch; float* norm_out; } a2cb_grad_norm_t;
typedef struct {
    float* params; float* grads; float* state1; float* state2; int64_t n; const float* norm;
    int32_t kind; int32_t first_step;
    float lr_eff, eps, b1, b2, omb1, omb2, bc2_sqrt, max_norm;
    const float* dyn;   /* optional DEVICE scalars {lr_eff, bc2_sqrt, first_step}: they override the fields above,
                           so a CUDA graph holding a whole epoch of optimiser steps is replayed without re-capture */
} a2cb_optim_t;
typedef struct { void* dst; int64_t bytes; int32_t value; int32_t pad_; } a2cb_fill_t;
typedef struct { float* out; const float* in; int64_t img_stride; int64_t pitch; int32_t B, P, C, ow; } a2cb_spatial_sum_t;
typedef struct { float* lo; const float* x; int64_t n; } a2cb_tcx_lo_t;
typedef struct { float* img; float* img_lo; const float* src; int64_t N, K; const int32_t* ntab; const int32_t* ktab; } a2cb_tcx_prep_t;
typedef struct { float* out; float* out_lo; const void* frames; const int64_t* idx; int32_t src_is_u8, B, C, H, W, S;
                 void* out_b16; /* optional bf16 copy of out (uint8 frames are exact in bf16) */
                 int32_t scatter; /* 1: image b is written to slot idx[b] of out (in-place re-layout of a row subset) */
                 int32_t ring_n;  /* frame ring (valid != NULL): environments per ring slot */
                 const uint8_t* valid; /* frame ring: frames = uint8 [slots, ring_n, 84, 84], each frame stored ONCE; the stack
                                          of row (t, n) is slots t .. t + 3, of which the newest valid[row] (1..4) are real and
                                          the older ones read as zero (a2c_ppo_acktr/envs.py:218-259 VecPyTorchFrameStack) */
               } a2cb_tcx_s2d_t;
typedef struct { float* part; const float* x; int64_t rows; int64_t pitch; int32_t C; int32_t pad_; float* lo; } a2cb_tcx_colsum_t;
typedef struct { float* part; const float* x; const float* gy; int64_t rows; int32_t K, N, gy_stride, pad_; } a2cb_tcx_narrow_wgrad_t;
typedef struct { const a2cb_tcx_prep_job_t* jobs; int64_t total; int32_t n_jobs; int32_t pad_; } a2cb_tcx_prep_multi_t;
typedef struct { float* z; const float* x; const float* W; const float* bias; int64_t rows; int32_t K, N, z_stride, pad_; } a2cb_tcx_narrow_fwd_t;
typedef struct { float* gx; float* gx_lo; const float* dz; const float* W; const float* mask; int64_t rows; int32_t K, N, dz_stride, pad_; } a2cb_tcx_narrow_dgrad_t;
typedef struct { int32_t kind; int32_t flags; const void* desc; } a2cb_op_t;

int a2cb_run_ops(const a2cb_op_t* ops, int32_t n_ops, const int64_t* idx, a2cb_stream_t stream);

/* ---------------------------------------------------------------------------
 * GRU recurrence of NNBase._forward_gru        a2c_ppo_acktr/model.py:111-166 (+ nn.GRU, :89-95)
 * Gate kernels for one time-step of a [T, N] sequence (rows time-major t*N + n); the two matmuls
 * (gi for all rows, gh per step) are a2cb_gemm descriptors in the same op list.  Episode boundaries
 * are handled by multiplying the carried state with masks[t] inside the kernels (equivalent to the
 * reference's per-segment cuDNN calls, without its host sync at model.py:133-137).
 *   which = A2CB_OP_GRU_INIT: hm[0] = h0 * mask[0]
 *           A2CB_OP_GRU_FWD : step t forward (saves r, z, n when non-NULL), writes hm[t+1]
 *           A2CB_OP_GRU_BWD : step t backward-through-time: dgi[t], dgh[t], dcarry = dh * z
 * ------------------------------------------------------------------------- */
typedef struct {
    const float* gi;
    float* gh;
    float* hm;
    const float* h0;
    const float* masks;
    const int64_t* idx;
    float* out;
    float* r; float* z; float* n;
    const float* dout;
    float* dgi; float* dgh;
    float* dcarry;
    const float* gh_part;   /* optional split-K partials [gh_splits][N,3H] of this step's hm W_hh^T: folded (+ gh_bias) */
    const float* gh_bias;   /*   by the forward gate kernel itself, which then also writes gh                        */
    const float* dc_part;   /* optional split-K partials [dc_splits][N,H] of dgh_{t+1} W_hh, folded by the backward kernel */
    int32_t T, N, H, t;
    int32_t gh_splits, dc_splits;
} a2cb_gru_t;

int a2cb_gru_op(int32_t which, const a2cb_gru_t* desc, a2cb_stream_t stream);

/* Whole-sequence GRU in ONE cooperative (persistent) launch per direction: every CTA keeps its slice of
 * W_hh in shared memory for all T steps and the CTAs meet at a grid barrier once per step.  Same buffers
 * as the per-step ops (gi must already hold x W_ih^T + b_ih; backward leaves dgi / dgh for the weight-
 * gradient GEMMs).  a2cb_gru_seq_supported says whether all CTAs of an [N, H] problem can be co-resident. */
typedef struct { a2cb_gru_t d; const float* W_hh; const float* b_hh; } a2cb_gru_seq_t;
int a2cb_gru_seq(int32_t backward, const a2cb_gru_seq_t* desc, a2cb_stream_t stream);
int a2cb_gru_seq_supported(int32_t N, int32_t H);
/* which kernel pair a2cb_gru_seq runs for this shape: 0 fp32 SIMT, 1 mma.sync 3xTF32, 2 tcgen05 (fp16 hi/lo operand
 * pairs, state exchanged as shared-memory images), -1 unsupported */
int a2cb_gru_seq_variant(int32_t N, int32_t H);
/* selects the kernel family of a2cb_gru_seq: 2 (default) tcgen05 where the shape allows, 1 mma.sync, 0 fp32 SIMT,
 * < 0 back to the default / environment (A2CB_GRU_MODE) */
int a2cb_set_gru_mode(int32_t mode);
/* diagnostics: clock64 stamps [2 directions][16 steps][16 phases] of CTA (0, 0) of the last launches run with
 * A2CB_GRU_TC_DEBUG=1; number of thread-block clusters of `cluster_size` CTAs with `smem_bytes` of dynamic shared
 * memory each that can be co-resident on this device (cudaOccupancyMaxActiveClusters) */
int a2cb_gru_tc_debug_stamps(int64_t* out);
int a2cb_gru_cluster_probe(int32_t cluster_size, int32_t smem_bytes);

/* ---------------------------------------------------------------------------
 * (5) ACKTR / KFAC helpers                              a2c_ppo_acktr/algo/kfac.py
 * The factor products (compute_cov_a / compute_cov_g, kfac.py:29-64) and the preconditioning chain
 * (kfac.py:216-227) are a2cb_gemm descriptors (implicit im2col, running average folded into
 * a2cb_reduce_splits with beta = stat_decay).  These entry points cover the rest:
 *   a2cb_spatial_sum: s[b,c] = sum_p g[b,p,c] over a (haloed) NHWC gradient image  (kfac.py:59-61)
 *   a2cb_kfac_scale:  v[i,j] /= d_g[i] * d_a[j] + damping (d_a NULL -> 1)          (kfac.py:223-225)
 *   a2cb_kfac_nu:     out[0] = nu = min(1, sqrt(kl_clip / sum(v * grad * lr^2)))   (kfac.py:229-234)
 *                     (scratch: 148 doubles + counter, zeroed once; a2cb_optim_step with max_norm < 0
 *                     then scales by norm[0] = nu before the SGD-momentum step, kfac.py:236-241)
 *   a2cb_frames_to_f32: dst = src / div as fp32 (uint8 or fp32 source)
 * ------------------------------------------------------------------------- */
int a2cb_spatial_sum(float* out, const float* in, int32_t B, int32_t P, int32_t C, int32_t ow, int64_t img_stride,
                     int64_t pitch, a2cb_stream_t stream);
int a2cb_kfac_scale(float* v, const float* dg, const float* da, int32_t rows, int32_t cols, float damping,
                    a2cb_stream_t stream);
int a2cb_kfac_nu(const float* v, const float* grads, int64_t n, double* scratch, float* out, float lr, float kl_clip,
                 a2cb_stream_t stream);
int a2cb_frames_to_f32(float* dst, const void* src, int64_t n, int32_t src_is_u8, float div, a2cb_stream_t stream);

/* ---------------------------------------------------------------------------
 * NCCL gradient exchange of the environment-sharded update          SURVEY.md section 8e (no reference counterpart:
 * the reference is single-device; correctness = the single-device result up to the summation order of the all-reduce)
 * All entry points take an ALREADY INITIALISED communicator (ncclComm_t as void*); a2cb_nccl_unique_id /
 * a2cb_nccl_comm_init_rank wrap ncclGetUniqueId / ncclCommInitRank for callers that have none (the 128-byte id is
 * produced on rank 0 and distributed by the host side).  libnccl.so.2 is resolved with dlopen at first use
 * (a2cb_nccl_load(path) pins a specific copy).  a2cb_comm_ctx_* own a side stream and events: inside an op list
 * A2CB_OP_COMM_FORK hands a finished gradient segment to ncclAllReduce(sum, fp32, in place) on the side stream while
 * the remaining weight-gradient kernels run, A2CB_OP_COMM_JOIN orders the optimiser ops after all of them.
 * ------------------------------------------------------------------------- */
typedef void* a2cb_nccl_comm_t;
typedef void* a2cb_comm_ctx_t;
typedef struct { a2cb_comm_ctx_t ctx; float* buf; int64_t count; } a2cb_comm_op_t;
int a2cb_nccl_load(const char* path);
int a2cb_nccl_unique_id(void* out128);
int a2cb_nccl_comm_init_rank(a2cb_nccl_comm_t* comm, int32_t world, int32_t rank, const void* id128);
int a2cb_nccl_comm_destroy(a2cb_nccl_comm_t comm);
int a2cb_nccl_allreduce_sum_f32(a2cb_nccl_comm_t comm, float* buf, int64_t count, a2cb_stream_t stream);
int a2cb_comm_ctx_create(a2cb_comm_ctx_t* ctx, a2cb_nccl_comm_t comm);
int a2cb_comm_ctx_destroy(a2cb_comm_ctx_t ctx);
int a2cb_comm_fork_allreduce(a2cb_comm_ctx_t ctx, float* buf, int64_t count, a2cb_stream_t stream);
int a2cb_comm_join(a2cb_comm_ctx_t ctx, a2cb_stream_t stream);

/* Column copies between rollout-shaped tensors (csrc/colcopy.cu): dst[dst_col(j)][t] = src[src_col(j)][t] for j < n_cols,
 * t < rows, every field with its own byte strides per column (cs) and per row (rs) and w bytes per element; dst_cols /
 * src_cols are device arrays of n_cols indices or NULL (column j).  Used by the balanced sharded recurrent update to pack
 * the rollout columns of migrating environments into records (storage.py:10-30 layout, time-major -> environment-major),
 * to unpack received records into guest columns and to copy uploaded columns into the shadow rollout: one launch for all
 * fields.  Elements of >= 256 bytes with 16-byte aligned pointers / strides move as 16-byte words. */
typedef struct {
    void* dst; const void* src;
    int64_t dst_cs, dst_rs, src_cs, src_rs;
    int64_t w;
    int32_t rows; int32_t reserved_;
} a2cb_colcopy_t;
int a2cb_copy_columns(const a2cb_colcopy_t* fields, int32_t n_fields, const int64_t* dst_cols, const int64_t* src_cols,
                      int32_t n_cols, a2cb_stream_t stream);

/* Peer-memory all-reduce (csrc/p2p.cu): the same sum over NVLink peer mappings instead of NCCL.  The update's kernels
 * are persistent and own (nearly) all of every SM's shared memory, so an NCCL kernel on the side stream cannot become
 * resident next to them and serialises with them; this exchange is three short copy / sum kernels without shared memory
 * (scatter slices into the owners' inboxes, sum in rank order and broadcast, copy back) whose two cross-GPU hand-shakes
 * are stream memory operations (no SM held while a rank waits).  One rank sums each element: results are bit-identical
 * on every rank.  a2cb_p2p_create allocates [inbox | result | counters] on the current device for segments of up to
 * capacity_floats; the 64-byte CUDA IPC handles are passed around by the host side (world * 64 bytes, in rank order) and
 * mapped by a2cb_p2p_open_ipc -- or the caller hands over its own mappings (a2cb_p2p_set_peers: one base pointer per
 * rank, from a2cb_p2p_local_base; same-process tests, symmetric-memory allocators).  Every rank must issue the same
 * sequence of a2cb_p2p_allreduce_sum_f32 calls (buf 16-byte aligned, in place).  a2cb_comm_ctx_set_p2p makes
 * A2CB_OP_COMM_FORK use it for every segment that fits (larger ones still go through NCCL). */
typedef void* a2cb_p2p_t;
int a2cb_p2p_create(a2cb_p2p_t* p2p, int32_t world, int32_t rank, int64_t capacity_floats);
int a2cb_p2p_ipc_handle(a2cb_p2p_t p2p, void* out64);
int a2cb_p2p_open_ipc(a2cb_p2p_t p2p, const void* handles);
int a2cb_p2p_set_peers(a2cb_p2p_t p2p, void* const* bases);
int a2cb_p2p_local_base(a2cb_p2p_t p2p, void** base);
int64_t a2cb_p2p_capacity(a2cb_p2p_t p2p);
int a2cb_p2p_allreduce_sum_f32(a2cb_p2p_t p2p, float* buf, int64_t count, a2cb_stream_t stream);
int a2cb_p2p_destroy(a2cb_p2p_t p2p);
int a2cb_comm_ctx_set_p2p(a2cb_comm_ctx_t ctx, a2cb_p2p_t p2p);

/* ---------------------------------------------------------------------------
 * GAIL discriminator                      a2c_ppo_acktr/algo/gail.py:11-110, main.py:141-155
 * trunk = Linear(D, H) - Tanh - Linear(H, H) - Tanh - Linear(H, 1), D = Ds + Da <= 128, H <= 128, flat parameters
 * [W1 (H x D) | b1 | W2 (H x H) | b2 | w3 (H) | b3] (the reference's state_dict order, row-major).
 *  a2cb_gail_disc(0, ..): ONE minibatch step of Discriminator.update (gail.py:57-96): policy rows gathered from the
 *    rollout buffer through idx, expert rows, mixup rows (gail.py:31-55 compute_grad_pen, alpha given by the caller);
 *    BCE-with-logits(expert, 1) + BCE-with-logits(policy, 0) + lambda mean((||dD/dmixup||_2 - 1)^2) and ALL parameter
 *    gradients incl. the double-backward terms; writes a2cb_gail_ctas(B) slabs of a2cb_gail_slab_floats(D, H) floats
 *    (gradients in parameter order, then {BCE loss, penalty}) to be summed (a2cb_reduce_splits).
 *  a2cb_gail_disc(1, ..): raw rewards log s - log(1 - s), s = sigmoid(D(state, action)), of B rows (gail.py:98-103).
 *  a2cb_gail_normalize: the sequential rest of predict_reward for T steps of N environments in one launch:
 *    returns = returns * masks * gamma + r (returns_init: 0 until `returns` exists), RunningMeanStd merge of the batch
 *    moments of `returns` into rms = double {mean, var, count}, out = r / sqrt(var + 1e-8)  (gail.py:104-110).
 * ------------------------------------------------------------------------- */
typedef struct {
    const float* P; float* slabs; const float* pol_state; const float* pol_action; const int64_t* idx;
    const float* exp_state; const float* exp_action; const float* alpha; float* raw_out;
    int32_t B, Ds, Da, H; float lambda; int32_t pad_;
} a2cb_gail_t;
int a2cb_gail_disc(int32_t reward, const a2cb_gail_t* desc, a2cb_stream_t stream);
int a2cb_gail_slab_floats(int32_t D, int32_t H);
int a2cb_gail_ctas(int32_t B);
int a2cb_gail_normalize(float* out, const float* raw, const float* masks, int32_t T, int32_t N, float gamma, float* returns,
                        int32_t* returns_init, double* rms, int32_t update_rms, a2cb_stream_t stream);

/* ---------------------------------------------------------------------------
 * Actor-side head of Policy.act          a2c_ppo_acktr/model.py:54-66, distributions.py:18-44,59-95
 * From the head outputs z[b] = (value, logits[A] | mean[A]): draws the action (Categorical: inverse CDF of the
 * softmax; DiagGaussian: mean + exp(logstd) * N(0,1)) or takes the mode (deterministic: argmax | mean), and writes
 * value, action and log-prob to their destinations (e.g. the [step] slots of the rollout buffer, storage.py:46-58).
 * rng = DEVICE {seed, offset, ticket, row_base} (4 x uint64, ticket zeroed once; row_base = global index of row 0): Philox4x32-10 keyed by seed, counter
 * (row, draw, offset); every sampling launch advances offset by one on the device.
 * ------------------------------------------------------------------------- */
typedef struct {
    const float* z; const float* logstd; float* value_out; void* action_out; float* logp_out; uint64_t* rng;
    int32_t B, A, z_stride, dist, deterministic, pad_;
} a2cb_sample_t;
int a2cb_sample_actions(const a2cb_sample_t* desc, a2cb_stream_t stream);

/* ---------------------------------------------------------------------------
 * Rollout ingest (the host -> device half of RolloutStorage, reference storage.py:35-58 / envs.py:167-190).
 * Uploads the columns of the listed environments of a time-major [rows, N, col_bytes] tensor from (pinned) host
 * memory into the device tensor of the same layout: one strided DMA copy per run of adjacent environments,
 * asynchronous on `stream`.  envs is a HOST list (it is consumed before the call returns).  The recurrent PPO
 * update pulls the frames of each minibatch's environments this way while earlier minibatches compute.
 * ------------------------------------------------------------------------- */
int a2cb_upload_env_columns(void* dst, const void* src_host, int64_t col_bytes, int64_t rows, int64_t pitch_bytes,
                            const int64_t* envs, int32_t n_envs, a2cb_stream_t stream);

#ifdef __cplusplus
}
#endif
#endif /* A2CB_H_ */
