// Internal launcher declarations (definitions in kernels_simt.cu / lstm_tc.cu / gemm_tc.cu).
#pragma once
#include "lagvae_common.cuh"

namespace lagvae {

const char* last_error();

int embed_gather(const int64_t* x, int64_t x_ld, int t_off, int B, int ns, int Tn, const float* table,
                 int ni, DropSpec drop, float* out, cudaStream_t st);
int embed_scatter_add(const int64_t* x, int64_t x_ld, int t_off, int B, int ns, int Tn,
                      const float* dX, int ni, DropSpec drop, float* dTable, int64_t skip_row,
                      cudaStream_t st);
int lstm_point_fwd(float* gates_t, const float* c_prev, float* c_t, float* h_t, float* hdrop_t,
                   DropSpec drop, int t, int Tn, int Bd, int nh, cudaStream_t st);
int lstm_point_bwd(const float* gates_t, const float* c_t, const float* c_prev, const float* dh_ext_t,
                   DropSpec drop, const float* dh_rec, float* dc, float* dgates_t, int t, int Tn,
                   int Bd, int nh, cudaStream_t st);
int head_reparam_kl(const float* h_last, const float* w_lin, const float* eps, int B, int nh, int nz,
                    int ns, float* mu, float* logvar, float* z, float* kl, cudaStream_t st);
int reparam_kl_bwd(const float* dz, const float* eps, const float* mu, const float* logvar,
                   const float* g_kl, int B, int nz, int ns, float* dml, cudaStream_t st);
int vec_add(const float* a, const float* b, float* o, int n, cudaStream_t st);
int tanh_copy(const float* a, float* o, int n, cudaStream_t st);
int dc0_total(const float* dc_init, const float* dh_init, const float* h0, float* o, int n, cudaStream_t st);
int time_sum(const float* src, int Tn, int Bd, int ncol, float* out, cudaStream_t st);
int add_row_periodic(float* x, const float* bias, int64_t rows, int ncol, int period, cudaStream_t st);
int logits_batch_major(const float* src, int64_t ld, int V, int Tn, int Bd, float* out, cudaStream_t st);
int col_sum(const float* src, int rows, int ncol, float* out1, float* out2, cudaStream_t st);
int ce_fwd(const float* logits, int64_t ld, int V, const int64_t* x, int64_t x_ld, int Tn, int Bd,
           int ns, float* lse_out, float* loss_row, cudaStream_t st);
int ce_bwd(float* logits, int64_t ld, int V, const int64_t* x, int64_t x_ld, int Tn, int Bd, int ns,
           const float* lse, const float* g_rec, cudaStream_t st);
int ce_bwd_split(const float* logits, int64_t ld, int V, const int64_t* x, int64_t x_ld, int Tn, int Bd, int ns,
                 const float* lse, const float* g_rec, uint16_t* hi, uint16_t* lo, int64_t ld_out, cudaStream_t st);
int ce_fused(const float* logits, int64_t ld, int V, const int64_t* x, int64_t x_ld, int Tn, int Bd, int ns, float g_row,
             float* lse_out, float* loss_row, uint16_t* hi, uint16_t* lo, int64_t ld_out, bool* done, cudaStream_t st);
int embed_gather_split(const int64_t* x, int64_t x_ld, int t_off, int B, int ns, int Tn, const float* table, int ni,
                       DropSpec drop, float* out, uint16_t* hi, uint16_t* lo, int64_t ld_out, cudaStream_t st);
int wait_flag(unsigned* flag, long long max_cycles, cudaStream_t st);
int finalize_loss(const float* loss_row, const float* kl, int B, int ns, int Tn, float klw, float* loss,
                  float* rec, float* kl_out, float* scalars, cudaStream_t st);
int combine_upstream(const float* gl, const float* gr, const float* gk, float klw, int B, float* g_rec,
                     float* g_kl, cudaStream_t st);
int fill(float* p, float v, int64_t n, cudaStream_t st);
int clip_sgd_step(float* const* h_params, float* const* h_grads, const int64_t* h_counts, int n_seg,
                  int n_update, float max_norm, float lr, int scale_all, float* out_norm, void* scratch,
                  cudaStream_t st);
int mi_estimate(const float* mu, const float* logvar, const float* eps, int B, int nz, float* out,
                cudaStream_t st);
int dropout_mask(uint64_t seed, uint32_t sid, int64_t n, float p, uint8_t* out, cudaStream_t st);

}  // namespace lagvae
