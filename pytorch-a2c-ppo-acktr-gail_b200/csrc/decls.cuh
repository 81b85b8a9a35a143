// Internal C++ entry points implemented by the .cu files and wrapped by capi.cu.
#pragma once
#include <cuda_runtime.h>
#include <stdint.h>

namespace a2cb {

// returns_scan.cu
int returns_scan(const float* rewards, float* value_preds, float* returns, const float* masks,
                 const float* bad_masks, const float* next_value, int T, long N, double gamma,
                 double gae_lambda, int use_gae, int use_ptl, int schedule, cudaStream_t st);

// gather.cu
struct GatherField { const void* src; void* dst; long row_bytes; };
int gather_rows(const GatherField* fields, int n_fields, const int64_t* idx, long n_rows, int mode,
                long E, long N, cudaStream_t st);

// gemm.cu
struct GemmDesc;
int gemm(const GemmDesc& g, cudaStream_t st);
int gemm_mode();
void set_gemm_mode(int m);
// gemm_tc.cu
bool gemm_tc_supported(const GemmDesc& g);
int gemm_tc(const GemmDesc& g, cudaStream_t st);
int gemm_prep_b(const GemmDesc& g, cudaStream_t st);
long gemm_b_image_floats(const GemmDesc& g);
int reduce_splits(float* dst, const float* src, long count, long stride, int splits, int accumulate,
                  const float* bias, int ncols, int act, float beta, cudaStream_t st);

// loss.cu
struct LossDesc;
int loss_epilogue(const LossDesc& d, cudaStream_t st);
long loss_scratch_floats(int B);
int adv_moments(const float* returns, const float* values, long n, double* scratch, double* moments,
                cudaStream_t st);
int adv_normalize(const float* returns, const float* values, long n, const double* moments, float* adv,
                  cudaStream_t st);

// optim.cu
int grad_norm(const float* g, long n, double* scratch, float* norm_out, cudaStream_t st);
int optim_step(int kind, float* p, float* g, float* s1, float* s2, long n, const float* norm,
               float lr_eff, float eps, float b1, float b2, float omb1, float omb2, float bc2_sqrt,
               float max_norm, int first_step, const float* dyn, cudaStream_t st);

// kfac.cu
int spatial_sum(float* out, const float* in, int B, int P, int C, int ow, long img_stride, long pitch,
                cudaStream_t st);
int kfac_scale(float* v, const float* dg, const float* da, int rows, int cols, float damping, cudaStream_t st);
int kfac_nu(const float* v, const float* g, long n, double* scratch, float* out, float lr, float kl_clip,
            cudaStream_t st);
int frames_to_f32(float* dst, const void* src, long n, int src_is_u8, float div, cudaStream_t st);

// gru.cu
struct GruDesc;
int gru_op(int which, const GruDesc& d, cudaStream_t st);   // which: 0 init, 1 forward step, 2 backward step

// gru_persistent.cu
bool gru_persistent_supported(int N, int H);
int gru_persistent_variant(int N, int H);     // 0 SIMT, 1 mma.sync 3xTF32, 2 tcgen05 (gru_tc.cu), -1 unsupported
// nccl_glue.cu
int nccl_load_library(const char* path);
int nccl_unique_id(void* out128);
int nccl_comm_init_rank(void** comm, int world, int rank, const void* id128);
int nccl_comm_destroy(void* comm);
int nccl_allreduce_sum_f32(void* comm, float* buf, long count, cudaStream_t st);
int comm_ctx_create(void** ctx, void* comm);
int comm_ctx_destroy(void* ctx);
int comm_fork_allreduce(void* ctx, float* buf, long count, cudaStream_t st);
int comm_join(void* ctx, cudaStream_t st);
int comm_ctx_set_p2p(void* ctx, void* p2p);
// colcopy.cu
struct ColCopyField {
    char* dst;
    const char* src;
    long dst_cs, dst_rs, src_cs, src_rs;   // byte strides per column / per row
    long w;                                // bytes per (column, row) element
    int rows;
    int wide;                              // set by copy_columns
};
int copy_columns(const ColCopyField* fields, int n_fields, const long* dst_cols, const long* src_cols, int n_cols, cudaStream_t st);
// p2p.cu
int p2p_create(void** out, int world, int rank, long capacity_floats);
int p2p_ipc_handle(void* ctx, void* out64);
int p2p_open_ipc(void* ctx, const void* handles);
int p2p_set_peers(void* ctx, void* const* bases);
int p2p_local_base(void* ctx, void** base);
long p2p_capacity(void* ctx);
int p2p_allreduce_sum_f32(void* ctx, float* buf, long count, cudaStream_t st);
int p2p_destroy(void* ctx);
// gail.cu
int gail_slab_floats(int D, int H);
int gail_ctas(int B);
int gail_disc(int reward, const void* desc, cudaStream_t st);
int gail_normalize(float* out, const float* raw, const float* masks, int T, int N, float gamma, float* returns,
                   int* returns_init, double* rms, int update_rms, cudaStream_t st);
// gru_tc.cu
bool gru_tc_supported(int N, int H);
int gru_tc_debug_stamps(long long* out);
int gru_cluster_probe(int cluster_size, int smem_bytes);
int gru_mode();
void set_gru_mode(int mode);
int gru_tc(int backward, const GruDesc& d, const float* Whh, const float* bhh, cudaStream_t st);
int gru_persistent(int backward, const GruDesc& d, const float* Whh, const float* bhh, cudaStream_t st);

// tcx.cu / tcx_aux.cu
struct TcxFwd;
struct TcxWgrad;
struct TcxKnobs;
int tcx_fwd(const TcxFwd& p, cudaStream_t st);
int tcx_wgrad(const TcxWgrad& p, cudaStream_t st);
void tcx_set_knobs(const TcxKnobs& k);
int tcx_lo_plane(float* lo, const float* x, long n, cudaStream_t st);
int tcx_prep_weights(float* img, float* img_lo, const float* src, long N, long K, const int32_t* ntab,
                     const int32_t* ktab, cudaStream_t st);
int tcx_s2d(float* out, float* out_lo, void* out_b16, const void* frames, int src_is_u8, const int64_t* idx, int B, int C,
            int H, int W, int S, cudaStream_t st, int scatter = 0, const uint8_t* valid = nullptr, int ring_n = 0);
int tcx_colsum(float* part, const float* x, float* lo, long rows, int C, long pitch, cudaStream_t st);
int tcx_colsum_blocks(long rows);
int tcx_prep_multi(const void* jobs, int n_jobs, long total, cudaStream_t st);
int tcx_narrow_wgrad(float* part, const float* x, const float* gy, long rows, int K, int N, int gy_stride, cudaStream_t st);
int tcx_narrow_wgrad_blocks(long rows);
int tcx_narrow_fwd(float* z, const float* x, const float* W, const float* bias, long rows, int K, int N, int z_stride,
                   cudaStream_t st);
int tcx_narrow_dgrad(float* gx, float* gx_lo, const float* dz, const float* W, const float* mask, long rows, int K, int N,
                     int dz_stride, cudaStream_t st);

// actor.cu
int sample_actions(const void* desc, cudaStream_t st);   // desc: a2cb_sample_t

// mlp_fused.cu
struct MlpDesc;
bool mlp_fused_supported(int d_in, int H, int zs);
int mlp_fused(int backward, const MlpDesc& d, cudaStream_t st);

}  // namespace a2cb
