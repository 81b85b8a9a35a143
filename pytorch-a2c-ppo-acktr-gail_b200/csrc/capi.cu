// extern "C" surface of liba2cb.so -- see include/a2cb.h for the contract.
#include "../../include/a2cb.h"
#include "common.cuh"
#include "decls.cuh"
#include "gemm.cuh"
#include "loss.cuh"
#include "gru.cuh"
#include "tcx.cuh"
#include "mlp_fused.cuh"

#include <algorithm>
#include <atomic>
#include <vector>
#include <stdarg.h>

namespace a2cb {

static thread_local char g_err[512] = "";
static std::atomic<int64_t> g_launches{0};

void set_error(const char* fmt, ...) {
    va_list ap;
    va_start(ap, fmt);
    vsnprintf(g_err, sizeof(g_err), fmt, ap);
    va_end(ap);
}

int check_launch(const char* what) {
    g_launches.fetch_add(1, std::memory_order_relaxed);
    cudaError_t e = cudaPeekAtLastError();
    if (e != cudaSuccess) {
        cudaGetLastError();
        set_error("%s: %s", what, cudaGetErrorString(e));
        return A2CB_ERR_CUDA;
    }
    return A2CB_OK;
}

}  // namespace a2cb

using namespace a2cb;

extern "C" {

int a2cb_abi_version(void) { return A2CB_ABI_VERSION; }
const char* a2cb_last_error(void) { return g_err; }
int64_t a2cb_launch_count(void) { return g_launches.load(); }
void a2cb_add_launches(int64_t n) { g_launches.fetch_add(n); }

int a2cb_returns_scan(const float* rewards, float* value_preds, float* returns, const float* masks,
                      const float* bad_masks, const float* next_value, int32_t T, int64_t N,
                      double gamma, double gae_lambda, int32_t use_gae,
                      int32_t use_proper_time_limits, int32_t schedule, a2cb_stream_t stream) {
    A2CB_REQUIRE((T == 0 || rewards) && value_preds && returns && (T == 0 || masks) && next_value, "returns_scan: null pointer");
    A2CB_REQUIRE(T == 0 || !use_proper_time_limits || bad_masks, "returns_scan: bad_masks required");
    A2CB_REQUIRE(T >= 0 && N >= 0 && schedule >= 0 && schedule <= 2, "returns_scan: bad T/N/schedule");
    if (N == 0) return A2CB_OK;
    return returns_scan(rewards, value_preds, returns, masks, bad_masks, next_value, T, (long)N, gamma,
                        gae_lambda, use_gae, use_proper_time_limits, schedule, (cudaStream_t)stream);
}

int a2cb_gather_rows(const a2cb_gather_field_t* fields, int32_t n_fields, const int64_t* idx,
                     int64_t n_out_rows, int32_t mode, int64_t E, int64_t N, a2cb_stream_t stream) {
    A2CB_REQUIRE(n_fields >= 0 && n_out_rows >= 0, "gather_rows: negative size");
    A2CB_REQUIRE(n_fields == 0 || n_out_rows == 0 || (fields && idx), "gather_rows: null pointer");
    static_assert(sizeof(a2cb_gather_field_t) == sizeof(GatherField), "layout");
    return gather_rows(reinterpret_cast<const GatherField*>(fields), n_fields, idx, (long)n_out_rows, mode,
                       (long)E, (long)N, (cudaStream_t)stream);
}

int a2cb_gemm(const a2cb_gemm_t* desc, a2cb_stream_t stream) {
    A2CB_REQUIRE(desc, "gemm: null descriptor");
    static_assert(sizeof(a2cb_gemm_t) == sizeof(GemmDesc), "a2cb_gemm_t and GemmDesc must share a layout");
    return gemm(*reinterpret_cast<const GemmDesc*>(desc), (cudaStream_t)stream);
}

int a2cb_gemm_prep_b(const a2cb_gemm_t* desc, a2cb_stream_t stream) {
    A2CB_REQUIRE(desc, "gemm_prep_b: null descriptor");
    return gemm_prep_b(*reinterpret_cast<const GemmDesc*>(desc), (cudaStream_t)stream);
}
int64_t a2cb_gemm_b_image_floats(const a2cb_gemm_t* desc) {
    return desc ? gemm_b_image_floats(*reinterpret_cast<const GemmDesc*>(desc)) : 0;
}

int a2cb_set_gemm_mode(int32_t mode) {
    A2CB_REQUIRE(mode >= 0 && mode <= 2, "set_gemm_mode: 0 (fp32 SIMT), 1 (tcgen05 3xTF32) or 2 (auto)");
    set_gemm_mode(mode);
    return A2CB_OK;
}
int a2cb_get_gemm_mode(void) { return gemm_mode(); }

int a2cb_reduce_splits(float* dst, const float* src, int64_t count, int64_t stride, int32_t splits,
                       int32_t accumulate, const float* bias, int32_t ncols, int32_t act, float beta,
                       a2cb_stream_t stream) {
    A2CB_REQUIRE(dst && src && count >= 0 && splits >= 1, "reduce_splits: bad argument");
    return reduce_splits(dst, src, (long)count, (long)stride, splits, accumulate, bias, ncols, act, beta,
                         (cudaStream_t)stream);
}

int a2cb_loss_epilogue(const a2cb_loss_t* desc, a2cb_stream_t stream) {
    A2CB_REQUIRE(desc, "loss: null descriptor");
    static_assert(sizeof(a2cb_loss_t) == sizeof(LossDesc), "a2cb_loss_t and LossDesc must share a layout");
    static_assert(A2CB_MAX_ACTIONS == kMaxActions, "max actions");
    return loss_epilogue(*reinterpret_cast<const LossDesc*>(desc), (cudaStream_t)stream);
}
int64_t a2cb_loss_scratch_floats(int32_t B) { return loss_scratch_floats(B); }

int a2cb_adv_moments(const float* returns, const float* value_preds, int64_t n, double* scratch,
                     double* moments, a2cb_stream_t stream) {
    return adv_moments(returns, value_preds, (long)n, scratch, moments, (cudaStream_t)stream);
}
int a2cb_adv_normalize(const float* returns, const float* value_preds, int64_t n, const double* moments,
                       float* adv, a2cb_stream_t stream) {
    return adv_normalize(returns, value_preds, (long)n, moments, adv, (cudaStream_t)stream);
}

int a2cb_grad_norm(const float* grads, int64_t n, double* scratch, float* norm_out, a2cb_stream_t stream) {
    return grad_norm(grads, (long)n, scratch, norm_out, (cudaStream_t)stream);
}
int a2cb_optim_step(int32_t kind, float* params, float* grads, float* state1, float* state2, int64_t n,
                    const float* norm, float lr_eff, float eps, float b1, float b2, float omb1, float omb2,
                    float bc2_sqrt, float max_norm, int32_t first_step, a2cb_stream_t stream) {
    return optim_step(kind, params, grads, state1, state2, (long)n, norm, lr_eff, eps, b1, b2, omb1, omb2,
                      bc2_sqrt, max_norm, first_step, nullptr, (cudaStream_t)stream);
}

int a2cb_spatial_sum(float* out, const float* in, int32_t B, int32_t P, int32_t C, int32_t ow, int64_t img_stride,
                     int64_t pitch, a2cb_stream_t stream) {
    return spatial_sum(out, in, B, P, C, ow, (long)img_stride, (long)pitch, (cudaStream_t)stream);
}
int a2cb_kfac_scale(float* v, const float* dg, const float* da, int32_t rows, int32_t cols, float damping,
                    a2cb_stream_t stream) {
    return kfac_scale(v, dg, da, rows, cols, damping, (cudaStream_t)stream);
}
int a2cb_kfac_nu(const float* v, const float* grads, int64_t n, double* scratch, float* out, float lr, float kl_clip,
                 a2cb_stream_t stream) {
    return kfac_nu(v, grads, (long)n, scratch, out, lr, kl_clip, (cudaStream_t)stream);
}
int a2cb_frames_to_f32(float* dst, const void* src, int64_t n, int32_t src_is_u8, float div, a2cb_stream_t stream) {
    return frames_to_f32(dst, src, (long)n, src_is_u8, div, (cudaStream_t)stream);
}

int a2cb_sample_actions(const a2cb_sample_t* desc, a2cb_stream_t stream) {
    A2CB_REQUIRE(desc, "sample_actions: null descriptor");
    return sample_actions(desc, (cudaStream_t)stream);
}

int a2cb_upload_env_columns(void* dst, const void* src_host, int64_t col_bytes, int64_t rows, int64_t pitch_bytes,
                            const int64_t* envs, int32_t n_envs, a2cb_stream_t stream) {
    A2CB_REQUIRE(dst && src_host && envs && col_bytes > 0 && rows > 0 && n_envs > 0 && pitch_bytes >= col_bytes &&
                 pitch_bytes % col_bytes == 0, "upload_env_columns: bad arguments");
    const int64_t N = pitch_bytes / col_bytes;
    std::vector<int64_t> e(envs, envs + n_envs);
    std::sort(e.begin(), e.end());
    for (size_t i = 0; i < e.size();) {
        A2CB_REQUIRE(e[i] >= 0 && e[i] < N, "upload_env_columns: environment %lld out of range", (long long)e[i]);
        size_t j = i + 1;
        while (j < e.size() && e[j] == e[j - 1] + 1) ++j;               // run of adjacent environments: one copy
        const int64_t off = e[i] * col_bytes;
        const cudaError_t err = cudaMemcpy2DAsync(static_cast<char*>(dst) + off, (size_t)pitch_bytes,
                                                  static_cast<const char*>(src_host) + off, (size_t)pitch_bytes,
                                                  (size_t)((int64_t)(j - i) * col_bytes), (size_t)rows,
                                                  cudaMemcpyHostToDevice, (cudaStream_t)stream);
        if (err != cudaSuccess) { set_error("upload_env_columns: %s", cudaGetErrorString(err)); return A2CB_ERR_CUDA; }
        i = j;
    }
    return 0;
}

int a2cb_gru_op(int32_t which, const a2cb_gru_t* desc, a2cb_stream_t stream) {
    A2CB_REQUIRE(desc && which >= A2CB_OP_GRU_INIT && which <= A2CB_OP_GRU_BWD, "gru_op: bad argument");
    static_assert(sizeof(a2cb_gru_t) == sizeof(GruDesc), "a2cb_gru_t and GruDesc must share a layout");
    return gru_op(which - A2CB_OP_GRU_INIT, *reinterpret_cast<const GruDesc*>(desc), (cudaStream_t)stream);
}

int a2cb_gru_seq(int32_t backward, const a2cb_gru_seq_t* desc, a2cb_stream_t stream) {
    A2CB_REQUIRE(desc, "gru_seq: null descriptor");
    return gru_persistent(backward != 0, *reinterpret_cast<const GruDesc*>(&desc->d), desc->W_hh, desc->b_hh,
                          (cudaStream_t)stream);
}
int a2cb_gru_seq_supported(int32_t N, int32_t H) { return gru_persistent_supported(N, H) ? 1 : 0; }
int a2cb_gru_seq_variant(int32_t N, int32_t H) { return gru_persistent_variant(N, H); }
int a2cb_set_gru_mode(int32_t mode) { set_gru_mode(mode); return A2CB_OK; }
int a2cb_gru_tc_debug_stamps(int64_t* out) { return gru_tc_debug_stamps(reinterpret_cast<long long*>(out)); }
int a2cb_gru_cluster_probe(int32_t cluster_size, int32_t smem_bytes) { return gru_cluster_probe(cluster_size, smem_bytes); }

int a2cb_nccl_load(const char* path) { return nccl_load_library(path); }
int a2cb_nccl_unique_id(void* out128) { return nccl_unique_id(out128); }
int a2cb_nccl_comm_init_rank(a2cb_nccl_comm_t* comm, int32_t world, int32_t rank, const void* id128) {
    return nccl_comm_init_rank(comm, world, rank, id128);
}
int a2cb_nccl_comm_destroy(a2cb_nccl_comm_t comm) { return nccl_comm_destroy(comm); }
int a2cb_nccl_allreduce_sum_f32(a2cb_nccl_comm_t comm, float* buf, int64_t count, a2cb_stream_t stream) {
    return nccl_allreduce_sum_f32(comm, buf, (long)count, (cudaStream_t)stream);
}
int a2cb_comm_ctx_create(a2cb_comm_ctx_t* ctx, a2cb_nccl_comm_t comm) { return comm_ctx_create(ctx, comm); }
int a2cb_comm_ctx_destroy(a2cb_comm_ctx_t ctx) { return comm_ctx_destroy(ctx); }
int a2cb_comm_fork_allreduce(a2cb_comm_ctx_t ctx, float* buf, int64_t count, a2cb_stream_t stream) {
    return comm_fork_allreduce(ctx, buf, (long)count, (cudaStream_t)stream);
}
int a2cb_comm_join(a2cb_comm_ctx_t ctx, a2cb_stream_t stream) { return comm_join(ctx, (cudaStream_t)stream); }
int a2cb_comm_ctx_set_p2p(a2cb_comm_ctx_t ctx, a2cb_p2p_t p2p) { return comm_ctx_set_p2p(ctx, p2p); }
int a2cb_p2p_create(a2cb_p2p_t* p2p, int32_t world, int32_t rank, int64_t capacity_floats) {
    return p2p_create(p2p, world, rank, (long)capacity_floats);
}
int a2cb_p2p_ipc_handle(a2cb_p2p_t p2p, void* out64) { return p2p_ipc_handle(p2p, out64); }
int a2cb_p2p_open_ipc(a2cb_p2p_t p2p, const void* handles) { return p2p_open_ipc(p2p, handles); }
int a2cb_p2p_set_peers(a2cb_p2p_t p2p, void* const* bases) { return p2p_set_peers(p2p, bases); }
int a2cb_p2p_local_base(a2cb_p2p_t p2p, void** base) { return p2p_local_base(p2p, base); }
int64_t a2cb_p2p_capacity(a2cb_p2p_t p2p) { return (int64_t)p2p_capacity(p2p); }
int a2cb_p2p_allreduce_sum_f32(a2cb_p2p_t p2p, float* buf, int64_t count, a2cb_stream_t stream) {
    return p2p_allreduce_sum_f32(p2p, buf, (long)count, (cudaStream_t)stream);
}
int a2cb_p2p_destroy(a2cb_p2p_t p2p) { return p2p_destroy(p2p); }
int a2cb_copy_columns(const a2cb_colcopy_t* fields, int32_t n_fields, const int64_t* dst_cols, const int64_t* src_cols,
                      int32_t n_cols, a2cb_stream_t stream) {
    static_assert(sizeof(a2cb_colcopy_t) == sizeof(ColCopyField), "a2cb_colcopy_t mirrors ColCopyField");
    return copy_columns(reinterpret_cast<const ColCopyField*>(fields), n_fields, reinterpret_cast<const long*>(dst_cols),
                        reinterpret_cast<const long*>(src_cols), n_cols, (cudaStream_t)stream);
}

int a2cb_gail_disc(int32_t reward, const a2cb_gail_t* desc, a2cb_stream_t stream) {
    A2CB_REQUIRE(desc, "gail_disc: null descriptor");
    return gail_disc(reward, desc, (cudaStream_t)stream);
}
int a2cb_gail_slab_floats(int32_t D, int32_t H) { return gail_slab_floats(D, H); }
int a2cb_gail_ctas(int32_t B) { return gail_ctas(B); }
int a2cb_gail_normalize(float* out, const float* raw, const float* masks, int32_t T, int32_t N, float gamma, float* returns,
                        int32_t* returns_init, double* rms, int32_t update_rms, a2cb_stream_t stream) {
    return gail_normalize(out, raw, masks, T, N, gamma, returns, returns_init, rms, update_rms, (cudaStream_t)stream);
}

int a2cb_mlp_fused(int32_t backward, const a2cb_mlp_t* desc, a2cb_stream_t stream) {
    A2CB_REQUIRE(desc, "mlp_fused: null descriptor");
    static_assert(sizeof(a2cb_mlp_t) == sizeof(MlpDesc), "a2cb_mlp_t out of sync with MlpDesc");
    return mlp_fused(backward != 0, *reinterpret_cast<const MlpDesc*>(desc), (cudaStream_t)stream);
}
int a2cb_mlp_fused_supported(int32_t d_in, int32_t H, int32_t zs) { return mlp_fused_supported(d_in, H, zs) ? 1 : 0; }

int a2cb_tcx_fwd(const a2cb_tcx_fwd_t* desc, a2cb_stream_t stream) {
    A2CB_REQUIRE(desc, "tcx_fwd: null descriptor");
    static_assert(sizeof(a2cb_tcx_fwd_t) == sizeof(TcxFwd), "a2cb_tcx_fwd_t out of sync with TcxFwd");
    return tcx_fwd(*reinterpret_cast<const TcxFwd*>(desc), (cudaStream_t)stream);
}
int a2cb_tcx_wgrad(const a2cb_tcx_wgrad_t* desc, a2cb_stream_t stream) {
    A2CB_REQUIRE(desc, "tcx_wgrad: null descriptor");
    static_assert(sizeof(a2cb_tcx_wgrad_t) == sizeof(TcxWgrad), "a2cb_tcx_wgrad_t out of sync with TcxWgrad");
    return tcx_wgrad(*reinterpret_cast<const TcxWgrad*>(desc), (cudaStream_t)stream);
}
int a2cb_tcx_lo_plane(float* lo, const float* x, int64_t n, a2cb_stream_t stream) {
    return tcx_lo_plane(lo, x, (long)n, (cudaStream_t)stream);
}
int a2cb_tcx_prep_weights(float* img, float* img_lo, const float* src, int64_t N, int64_t K, const int32_t* ntab,
                          const int32_t* ktab, a2cb_stream_t stream) {
    return tcx_prep_weights(img, img_lo, src, (long)N, (long)K, ntab, ktab, (cudaStream_t)stream);
}
int a2cb_tcx_s2d(float* out, float* out_lo, const void* frames, int32_t src_is_u8, const int64_t* idx, int32_t B,
                 int32_t C, int32_t H, int32_t W, int32_t S, a2cb_stream_t stream) {
    return tcx_s2d(out, out_lo, nullptr, frames, src_is_u8, idx, B, C, H, W, S, (cudaStream_t)stream);
}
int a2cb_tcx_s2d_b16(float* out, float* out_lo, void* out_b16, const void* frames, int32_t src_is_u8, const int64_t* idx,
                     int32_t B, int32_t C, int32_t H, int32_t W, int32_t S, a2cb_stream_t stream) {
    return tcx_s2d(out, out_lo, out_b16, frames, src_is_u8, idx, B, C, H, W, S, (cudaStream_t)stream);
}
int a2cb_tcx_colsum(float* part, const float* x, float* lo, int64_t rows, int32_t C, int64_t pitch, a2cb_stream_t stream) {
    return tcx_colsum(part, x, lo, (long)rows, C, (long)pitch, (cudaStream_t)stream);
}
int a2cb_tcx_prep_multi(const a2cb_tcx_prep_job_t* jobs, int32_t n_jobs, int64_t total, a2cb_stream_t stream) {
    return tcx_prep_multi(jobs, n_jobs, (long)total, (cudaStream_t)stream);
}
int a2cb_tcx_narrow_wgrad(float* part, const float* x, const float* gy, int64_t rows, int32_t K, int32_t N,
                          int32_t gy_stride, a2cb_stream_t stream) {
    return tcx_narrow_wgrad(part, x, gy, (long)rows, K, N, gy_stride, (cudaStream_t)stream);
}
int a2cb_tcx_narrow_fwd(float* z, const float* x, const float* W, const float* bias, int64_t rows, int32_t K, int32_t N,
                        int32_t z_stride, a2cb_stream_t stream) {
    return tcx_narrow_fwd(z, x, W, bias, (long)rows, K, N, z_stride, (cudaStream_t)stream);
}
int a2cb_tcx_narrow_dgrad(float* gx, float* gx_lo, const float* dz, const float* W, const float* mask, int64_t rows,
                          int32_t K, int32_t N, int32_t dz_stride, a2cb_stream_t stream) {
    return tcx_narrow_dgrad(gx, gx_lo, dz, W, mask, (long)rows, K, N, dz_stride, (cudaStream_t)stream);
}
int a2cb_tcx_narrow_wgrad_blocks(int64_t rows) { return tcx_narrow_wgrad_blocks((long)rows); }
int a2cb_tcx_sizeof(int32_t which) {
    return which == 0 ? (int)sizeof(TcxFwd) : which == 1 ? (int)sizeof(TcxWgrad) : (int)sizeof(TcxView);
}
int a2cb_tcx_colsum_blocks(int64_t rows) { return tcx_colsum_blocks((long)rows); }
int a2cb_tcx_set_knobs(int32_t mn_layout_type, int32_t mn_sbo_bytes, int32_t mn_kstep_bytes, int32_t mn_tma_swizzle) {
    TcxKnobs k = {mn_layout_type, mn_sbo_bytes, mn_kstep_bytes, mn_tma_swizzle};
    tcx_set_knobs(k);
    return A2CB_OK;
}

int a2cb_run_ops(const a2cb_op_t* ops, int32_t n_ops, const int64_t* idx, a2cb_stream_t stream) {
    A2CB_REQUIRE(n_ops >= 0 && (n_ops == 0 || ops), "run_ops: bad op list");
    cudaStream_t st = (cudaStream_t)stream;
    for (int i = 0; i < n_ops; ++i) {
        const a2cb_op_t& op = ops[i];
        const bool rt = (op.flags & A2CB_OP_RUNTIME_IDX) != 0;
        A2CB_REQUIRE(op.desc, "run_ops: op %d has no descriptor", i);
        A2CB_REQUIRE(!rt || idx, "run_ops: op %d wants a runtime index list but none was given", i);
        int rc = A2CB_OK;
        switch (op.kind) {
            case A2CB_OP_GEMM: {
                GemmDesc g = *reinterpret_cast<const GemmDesc*>(op.desc);
                if (rt) {
                    if (g.gather_row) g.gather_row = idx;
                    if (g.gather_red) g.gather_red = idx;
                }
                rc = gemm(g, st);
                break;
            }
            case A2CB_OP_REDUCE: {
                const a2cb_reduce_t& r = *reinterpret_cast<const a2cb_reduce_t*>(op.desc);
                rc = reduce_splits(r.dst, r.src, (long)r.count, (long)r.stride, r.splits, r.accumulate, r.bias, r.ncols, r.act, r.beta, st);
                break;
            }
            case A2CB_OP_LOSS: {
                LossDesc d = *reinterpret_cast<const LossDesc*>(op.desc);
                if (rt) d.idx = idx;
                rc = loss_epilogue(d, st);
                break;
            }
            case A2CB_OP_GRAD_NORM: {
                const a2cb_grad_norm_t& n = *reinterpret_cast<const a2cb_grad_norm_t*>(op.desc);
                rc = grad_norm(n.grads, (long)n.n, n.scratch, n.norm_out, st);
                break;
            }
            case A2CB_OP_OPTIM: {
                const a2cb_optim_t& o = *reinterpret_cast<const a2cb_optim_t*>(op.desc);
                rc = optim_step(o.kind, o.params, o.grads, o.state1, o.state2, (long)o.n, o.norm, o.lr_eff, o.eps,
                                o.b1, o.b2, o.omb1, o.omb2, o.bc2_sqrt, o.max_norm, o.first_step, o.dyn, st);
                break;
            }
            case A2CB_OP_COMM_FORK: {
                const a2cb_comm_op_t& q = *reinterpret_cast<const a2cb_comm_op_t*>(op.desc);
                rc = comm_fork_allreduce(q.ctx, q.buf, (long)q.count, st);
                break;
            }
            case A2CB_OP_COMM_JOIN: {
                const a2cb_comm_op_t& q = *reinterpret_cast<const a2cb_comm_op_t*>(op.desc);
                rc = comm_join(q.ctx, st);
                break;
            }
            case A2CB_OP_FILL: {
                const a2cb_fill_t& f = *reinterpret_cast<const a2cb_fill_t*>(op.desc);
                cudaError_t e = cudaMemsetAsync(f.dst, f.value, (size_t)f.bytes, st);
                if (e != cudaSuccess) { set_error("run_ops: memset: %s", cudaGetErrorString(e)); rc = A2CB_ERR_CUDA; }
                break;
            }
            case A2CB_OP_GEMM_PREP_B: {
                rc = gemm_prep_b(*reinterpret_cast<const GemmDesc*>(op.desc), st);
                break;
            }
            case A2CB_OP_SPATIAL_SUM: {
                const a2cb_spatial_sum_t& q = *reinterpret_cast<const a2cb_spatial_sum_t*>(op.desc);
                rc = spatial_sum(q.out, q.in, q.B, q.P, q.C, q.ow, (long)q.img_stride, (long)q.pitch, st);
                break;
            }
            case A2CB_OP_GRU_SEQ_FWD:
            case A2CB_OP_GRU_SEQ_BWD: {
                const a2cb_gru_seq_t& q = *reinterpret_cast<const a2cb_gru_seq_t*>(op.desc);
                GruDesc d = *reinterpret_cast<const GruDesc*>(&q.d);
                if (rt) d.idx = idx;
                rc = gru_persistent(op.kind == A2CB_OP_GRU_SEQ_BWD, d, q.W_hh, q.b_hh, st);
                break;
            }
            case A2CB_OP_TCX_FWD: {
                if (!rt) { rc = tcx_fwd(*reinterpret_cast<const TcxFwd*>(op.desc), st); break; }
                TcxFwd d = *reinterpret_cast<const TcxFwd*>(op.desc);
                d.gather_idx = idx;
                rc = tcx_fwd(d, st);
                break;
            }
            case A2CB_OP_TCX_WGRAD: {
                if (!rt) { rc = tcx_wgrad(*reinterpret_cast<const TcxWgrad*>(op.desc), st); break; }
                TcxWgrad d = *reinterpret_cast<const TcxWgrad*>(op.desc);
                d.gather_idx = idx;
                rc = tcx_wgrad(d, st);
                break;
            }
            case A2CB_OP_TCX_LO: {
                const a2cb_tcx_lo_t& q = *reinterpret_cast<const a2cb_tcx_lo_t*>(op.desc);
                rc = tcx_lo_plane(q.lo, q.x, (long)q.n, st);
                break;
            }
            case A2CB_OP_TCX_PREP: {
                const a2cb_tcx_prep_t& q = *reinterpret_cast<const a2cb_tcx_prep_t*>(op.desc);
                rc = tcx_prep_weights(q.img, q.img_lo, q.src, (long)q.N, (long)q.K, q.ntab, q.ktab, st);
                break;
            }
            case A2CB_OP_TCX_S2D: {
                const a2cb_tcx_s2d_t& q = *reinterpret_cast<const a2cb_tcx_s2d_t*>(op.desc);
                rc = tcx_s2d(q.out, q.out_lo, q.out_b16, q.frames, q.src_is_u8, rt ? idx : q.idx, q.B, q.C, q.H, q.W, q.S, st, q.scatter,
                             q.valid, q.ring_n);
                break;
            }
            case A2CB_OP_TCX_COLSUM: {
                const a2cb_tcx_colsum_t& q = *reinterpret_cast<const a2cb_tcx_colsum_t*>(op.desc);
                rc = tcx_colsum(q.part, q.x, q.lo, (long)q.rows, q.C, (long)q.pitch, st);
                break;
            }
            case A2CB_OP_SAMPLE:
                rc = sample_actions(op.desc, st);
                break;
            case A2CB_OP_MLP_FWD:
            case A2CB_OP_MLP_BWD: {
                MlpDesc q = *reinterpret_cast<const MlpDesc*>(op.desc);
                if (rt) q.idx = idx;
                rc = mlp_fused(op.kind == A2CB_OP_MLP_BWD, q, st);
                break;
            }
            case A2CB_OP_TCX_NARROW_FWD: {
                const a2cb_tcx_narrow_fwd_t& q = *reinterpret_cast<const a2cb_tcx_narrow_fwd_t*>(op.desc);
                rc = tcx_narrow_fwd(q.z, q.x, q.W, q.bias, (long)q.rows, q.K, q.N, q.z_stride, st);
                break;
            }
            case A2CB_OP_TCX_NARROW_DGRAD: {
                const a2cb_tcx_narrow_dgrad_t& q = *reinterpret_cast<const a2cb_tcx_narrow_dgrad_t*>(op.desc);
                rc = tcx_narrow_dgrad(q.gx, q.gx_lo, q.dz, q.W, q.mask, (long)q.rows, q.K, q.N, q.dz_stride, st);
                break;
            }
            case A2CB_OP_TCX_PREP_MULTI: {
                const a2cb_tcx_prep_multi_t& q = *reinterpret_cast<const a2cb_tcx_prep_multi_t*>(op.desc);
                rc = tcx_prep_multi(q.jobs, q.n_jobs, (long)q.total, st);
                break;
            }
            case A2CB_OP_TCX_NARROW_WGRAD: {
                const a2cb_tcx_narrow_wgrad_t& q = *reinterpret_cast<const a2cb_tcx_narrow_wgrad_t*>(op.desc);
                rc = tcx_narrow_wgrad(q.part, q.x, q.gy, (long)q.rows, q.K, q.N, q.gy_stride, st);
                break;
            }
            case A2CB_OP_GRU_INIT:
            case A2CB_OP_GRU_FWD:
            case A2CB_OP_GRU_BWD: {
                GruDesc d = *reinterpret_cast<const GruDesc*>(op.desc);
                if (rt) d.idx = idx;
                rc = gru_op(op.kind - A2CB_OP_GRU_INIT, d, st);
                break;
            }
            default:
                set_error("run_ops: unknown op kind %d at %d", op.kind, i);
                return A2CB_ERR_INVALID;
        }
        if (rc != A2CB_OK) return rc;
    }
    return A2CB_OK;
}

}  // extern "C"
