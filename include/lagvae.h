/* lagvae.h — C-ABI of liblagvae.so: B200 (sm_100a) kernels for the aggressive-inner-loop hot path
 * of jxhe/vae-lagging-encoder.
 *
 * The reference has NO native/FFI interface (it is pure Python on torch); the
 * boundary a maintainer binds is therefore the set of torch library calls the reference makes on
 * the path.  Each entry point below names the reference call site (file:line, relative to the
 * reference root) whose arithmetic it replaces.  INTEGRATION.md shows the ctypes stub.
 *
 * Conventions
 *  - plain pointers and sizes only; every pointer is a DEVICE pointer unless its name starts with
 *    `h_`; `stream` is a cudaStream_t passed as void*.
 *  - every function returns 0 on success, non-zero on failure (LAGVAE_E_*); the message of the last
 *    failure on the calling thread is returned by lagvae_last_error().  No exceptions, no
 *    allocation inside (workspaces are caller-owned; sizes come from *_workspace_bytes queries).
 *  - there is NO CPU fallback: every compute entry point launches CUDA kernels and fails with
 *    LAGVAE_E_CUDA when no sm_100 device/context is usable.
 *  - activations are TIME-MAJOR: row r = t*Bd + bd, where Bd = B*ns decoder rows, bd = b*ns + s
 *    (sample-major inside a sentence, as dec_lstm.py:87-94).
 *  - gate order along 4*nh is PyTorch's i,f,g,o.
 */
#ifndef LAGVAE_H_
#define LAGVAE_H_

#include <stddef.h>
#include <stdint.h>

#ifdef __cplusplus
extern "C" {
#endif

#define LAGVAE_OK 0
#define LAGVAE_E_ARG 1     /* bad argument / unsupported shape */
#define LAGVAE_E_CUDA 2    /* CUDA runtime / launch failure, or no usable sm_100 device */
#define LAGVAE_E_WORKSPACE 3

#define LAGVAE_ABI_VERSION 2

int lagvae_abi_version(void);
const char* lagvae_last_error(void);
/* 0 if the current device is sm_100 (B200) and kernels can launch; LAGVAE_E_CUDA otherwise. */
int lagvae_device_check(void);
/* number of kernels launched by this library on this process since load (bench `gpu_launches`) */
int64_t lagvae_launch_count(void);

/* ---------------------------------------------------------------------------------------------
 * Text-path dimensions.  Be = B encoder rows, Bd = B*ns decoder rows, T = token columns of x
 * (incl. <s>, </s>), Td = T-1 decoder steps.
 * ------------------------------------------------------------------------------------------- */
typedef struct lagvae_text_dims {
  int32_t B, T, ns;
  int32_t V, ni, nh, nz;
} lagvae_text_dims;

/* The 13 parameter tensors in reference state_dict / vae.parameters() order:
 * encoder: embed.weight[V,ni] lstm.weight_ih_l0[4nh,ni] lstm.weight_hh_l0[4nh,nh] lstm.bias_ih_l0[4nh]
 *          lstm.bias_hh_l0[4nh] linear.weight[2nz,nh]
 * decoder: embed.weight[V,ni] trans_linear.weight[nh,nz] lstm.weight_ih_l0[4nh,ni+nz]
 *          lstm.weight_hh_l0[4nh,nh] lstm.bias_ih_l0[4nh] lstm.bias_hh_l0[4nh] pred_linear.weight[V,nh]
 * All fp32, contiguous row-major.  The same struct carries gradient pointers. */
#define LAGVAE_TEXT_NPARAM 13
typedef struct lagvae_text_params {
  float* p[LAGVAE_TEXT_NPARAM];
} lagvae_text_params;

/* Dropout control (dec_lstm.py:30-31,81,106).  mode 0: identity (eval()).  mode 1: caller-provided
 * keep masks (uint8, 1 = keep): mask_in [B,Td,ni] (applied before the ns expansion, dec_lstm.py:81
 * precedes :87), mask_out [Bd,Td,nh].  mode 2: in-kernel counter-based Philox4x32-10 keyed by
 * (seed, stream id), element index = logical index in the shapes above.  seed_dev (mode 2, may be NULL): a device
 * word ADDED to `seed` (mod 2^64) when the kernels run — a step captured in a CUDA graph freezes the host scalar `seed`,
 * and draws a fresh mask at every replay by bumping that word inside the graph (the forward and the backward of one step
 * must see the same value). */
typedef struct lagvae_dropout {
  int32_t mode;
  float p_in, p_out;
  const uint8_t* mask_in;
  const uint8_t* mask_out;
  uint64_t seed;
  const uint64_t* seed_dev;
} lagvae_dropout;

/* Opaque per-shape plan: workspace carving + kernel selection. */
typedef struct lagvae_text_plan lagvae_text_plan;

/* flags for lagvae_text_plan_create */
#define LAGVAE_PLAN_DEFAULT 0u
#define LAGVAE_PLAN_FORCE_SIMT 1u   /* use the fp32 SIMT kernels for every contraction */
#define LAGVAE_PLAN_INFERENCE 2u    /* forward only: no stash for backward */

size_t lagvae_text_workspace_bytes(const lagvae_text_dims* d, uint32_t flags);
/* `workspace` (device, >= workspace_bytes, 256-B aligned) stays owned by the caller and must
 * outlive the plan.  h_plan_out receives the handle. */
int lagvae_text_plan_create(const lagvae_text_dims* d, uint32_t flags, void* workspace,
                            size_t workspace_bytes, lagvae_text_plan** h_plan_out);
void lagvae_text_plan_destroy(lagvae_text_plan* plan);

/* VAE.loss forward  — modules/vae.py:79-98 = encoder.py:40-57 (enc_lstm.py:47-64, reparam 59-79,
 * KL :55) + dec_lstm.py:113-148 (decode 66-111, CrossEntropyLoss :47,143) + vae.py:95,98.
 *  x      int64 [B,T] row-major (driver-owned, never written)
 *  eps    fp32 [B,ns,nz]  the N(0,1) draw of encoder.py:77
 *  out_loss/out_rec/out_kl  fp32 [B];  out_mu/out_logvar fp32 [B,nz]; out_z fp32 [B,ns,nz]
 *  (any out_* except loss/rec/kl may be NULL). */
int lagvae_text_loss_forward(lagvae_text_plan* plan, const lagvae_text_params* params,
                             const int64_t* x, const float* eps, float kl_weight,
                             const lagvae_dropout* drop, float* out_loss, float* out_rec,
                             float* out_kl, float* out_mu, float* out_logvar, float* out_z,
                             void* stream);

/* Backward of the above — the autograd walk of text.py:382-384.  Must follow a
 * lagvae_text_loss_forward on the same plan (uses its stash).  g_loss/g_rec/g_kl: upstream
 * gradients of the three outputs, fp32 [B], any may be NULL (= zeros).  `grads` receives the 13
 * gradients (overwritten, not accumulated); decoder.embed.weight row V-1 gets zeros
 * (padding_idx=-1, dec_lstm.py:28). */
#define LAGVAE_BWD_DEFAULT 0u
/* The decoder WEIGHT gradients (pred_linear, decoder lstm weight_ih/hh, decoder embedding) feed only the clip norm in the aggressive
 * loop (text.py:385; the decoder is not stepped, text.py:387): compute them with ONE bf16 pass instead of three
 * (norm error ~1e-5).  Everything that reaches the encoder update stays fp32-grade. */
#define LAGVAE_BWD_DECODER_WGRAD_NORM_ONLY 1u
int lagvae_text_loss_backward(lagvae_text_plan* plan, const lagvae_text_params* params,
                              const int64_t* x, const float* g_loss, const float* g_rec,
                              const float* g_kl, const lagvae_text_params* grads, uint32_t flags,
                              void* stream);

/* Encoder forward only: enc_lstm.py:47-64 (VAE.encode_stats vae.py:33-40). */
int lagvae_text_encode_stats(lagvae_text_plan* plan, const lagvae_text_params* params,
                             const int64_t* x, float* out_mu, float* out_logvar, void* stream);

/* Decoder only: LSTMDecoder.reconstruct_error — dec_lstm.py:113-148 (log_probability :151-161 is its
 * negation).  z fp32 [B*ns, nz]; out_rec_rows fp32 [B*ns] (row b*ns+s).  Forward only. */
int lagvae_text_reconstruct_error(lagvae_text_plan* plan, const lagvae_text_params* params,
                                  const int64_t* x, const float* z, const lagvae_dropout* drop,
                                  float* out_rec_rows, void* stream);

/* Decoder only: LSTMDecoder.decode — dec_lstm.py:66-111.  `input` int64 [B, T-1] row-major (the plan's T minus one: the
 * decoder consumes T-1 tokens), z fp32 [B*ns, nz]; out_logits fp32 [B*ns, T-1, V] row-major, the tensor the reference returns
 * (vocabulary projection of every position, dec_lstm.py:109).  Forward only. */
int lagvae_text_decode_logits(lagvae_text_plan* plan, const lagvae_text_params* params, const int64_t* input,
                              const float* z, const lagvae_dropout* drop, float* out_logits, void* stream);

/* clip_grad_norm_(all params, max_norm) + SGD(momentum 0) on the first `n_update` tensors —
 * text.py:385 + text.py:387 (optim.SGD text.py:325).  segs: `n_seg` (ptr,count) pairs describing
 * the gradient tensors; params/grads pair up by index.  out_norm (device fp32[1]) receives the
 * pre-clip total L2 norm.  scale_all_grads != 0 also multiplies every gradient by the clip
 * coefficient in place (what clip_grad_norm_ does); 0 skips that write for tensors that are not
 * updated (their clipped value is dead in the aggressive loop). scratch: device, >= 16 KiB. */
int lagvae_clip_sgd_step(float* const* h_params, float* const* h_grads, const int64_t* h_counts,
                         int n_seg, int n_update, float max_norm, float lr, int scale_all_grads,
                         float* out_norm, void* scratch, void* stream);

/* clip_grad_norm_(all params, max_norm) + torch.optim.Adam step (amsgrad False, no weight decay) on the first
 * `n_update` tensors — image.py:312 + image.py:314 (optim.Adam(lr=0.001) image.py:267; toy.py:289-290).  The image model
 * has 248 parameter tensors, so the (param, grad, exp_avg, exp_avg_sq, count) tuples live in a device-resident table built
 * once: `lagvae_adam_table_create` uploads it into caller-owned device memory (>= lagvae_adam_table_bytes, 256-B aligned;
 * synchronous, call it outside graph capture); `lagvae_clip_adam_step` is then three launches and capturable: Adam's step
 * count (bias correction) lives on the device, starts at `initial_step` and is incremented by every call.  exp_avg /
 * exp_avg_sq are caller-owned fp32 tensors (the optimizer state, zero before the first step).  scale_all_grads as in
 * lagvae_clip_sgd_step.  out_norm: device fp32[1] or NULL. */
typedef struct lagvae_adam_table lagvae_adam_table;
size_t lagvae_adam_table_bytes(const int64_t* h_counts, int n_seg);
int lagvae_adam_table_create(float* const* h_params, float* const* h_grads, float* const* h_exp_avg,
                             float* const* h_exp_avg_sq, const int64_t* h_counts, int n_seg, int n_update,
                             int initial_step, void* device_mem, size_t device_bytes, void* stream,
                             lagvae_adam_table** out);
void lagvae_adam_table_destroy(lagvae_adam_table* table);
int lagvae_clip_adam_step(lagvae_adam_table* table, float max_norm, float lr, float beta1, float beta2, float eps,
                          int scale_all_grads, float* out_norm, void* stream);

/* The data-path collective of the batch-sharded inner step (the single-process reference has no
 * counterpart: its loop body is text.py:371-391): ONE sum all-reduce of the flat fp32 gradient bucket per inner step over
 * NCCL / NVLink, the communicator owned by the library.  NCCL is bound at run time: `lagvae_comm_load` dlopen()s the
 * libnccl.so.2 the host side points it to (torch's bundled copy), `lagvae_comm_unique_id` (rank 0) fills a 128-byte id that
 * the caller distributes to the other ranks out of band, `lagvae_comm_init` is collective over the `world` ranks (one
 * process per GPU, current device).  `lagvae_allreduce_bucket` enqueues the all-reduce of bucket[0, count) on `stream`. */
typedef struct lagvae_comm lagvae_comm;
int lagvae_comm_load(const char* libnccl_path);
int lagvae_comm_unique_id(void* out_128_bytes);
int lagvae_comm_init(const void* uid_128_bytes, int rank, int world, lagvae_comm** out);
int lagvae_allreduce_bucket(lagvae_comm* comm, float* bucket, int64_t count, void* stream);
void lagvae_comm_destroy(lagvae_comm* comm);

/* MI estimate from posterior stats — encoder.py:111-145 (+utils.py:3-16).  mu/logvar [B,nz], eps
 * [B,nz] (the draw of encoder.py:128).  out_mi: device fp32[1]. */
int lagvae_mi_estimate(const float* mu, const float* logvar, const float* eps, int B, int nz,
                       float* out_mi, void* stream);

/* Fused aggressive inner step — one iteration of text.py:371-391 without the host round trips:
 * zero-grad, loss fwd, Σloss, backward of mean(loss), clip(max_norm) over all 13 grads, SGD on the
 * 6 encoder tensors.  `grad_ws` is a caller-owned flat fp32 buffer of lagvae_text_param_count()
 * floats.  out_scalars (device fp32[4]) = {Σloss, Σrec, ΣKL, pre-clip grad norm}. */
int64_t lagvae_text_param_count(const lagvae_text_dims* d);
int lagvae_text_inner_step(lagvae_text_plan* plan, const lagvae_text_params* params,
                           const int64_t* x, const float* eps, float kl_weight,
                           const lagvae_dropout* drop, float max_norm, float lr, float* grad_ws,
                           float* out_loss, float* out_scalars, void* stream);

/* The decoder-update step that closes every outer iteration — text.py:407-424: zero-grad, loss fwd, backward of mean(loss),
 * clip(max_norm) over all 13 grads, then `dec_optimizer.step()` and, once the aggressive phase is over (`update_encoder`
 * != 0, text.py:421-422), `enc_optimizer.step()` too — SGD(lr, momentum 0) both (text.py:325-326).  Same buffers and
 * out_scalars as lagvae_text_inner_step; all gradients are computed at full (3-pass) precision here. */
int lagvae_text_outer_step(lagvae_text_plan* plan, const lagvae_text_params* params, const int64_t* x,
                           const float* eps, float kl_weight, const lagvae_dropout* drop, float max_norm,
                           float lr, int update_encoder, float* grad_ws, float* out_loss, float* out_scalars,
                           void* stream);

/* The decoder weights do not change inside the aggressive loop (text.py:387 steps the encoder only), so their bf16 hi/lo
 * tensor-core operand copies (W_pred 82 MB, the x-columns of the decoder W_ih) need not be re-split at every inner step.
 * The caller declares an EPOCH for the decoder weights: while consecutive calls on this plan see the same non-zero epoch the
 * cached splits are reused; any change of the value (or 0 = "unknown, never cache", the default) re-splits.  The host side
 * derives the epoch from the tensors' data pointers and torch version counters, which every in-place update bumps
 * (optimizer steps, load_state_dict); lagvae_text_outer_step invalidates the cache itself. */
int lagvae_text_decoder_weights_epoch(lagvae_text_plan* plan, uint64_t epoch);

/* Data-parallel overlap hook (no counterpart in the single-process reference).  The decoder
 * gradients (the last 7 tensors of vae.parameters(), 70% of the bucket) are final before the encoder LSTM
 * backward starts.  With the hook enabled lagvae_text_loss_backward records an internal event at that
 * point and lagvae_text_wait_decoder_grads makes `stream` wait for it, so that the caller can all-reduce
 * the decoder part of the flat gradient on a side stream while the encoder backward is still running. */
int lagvae_text_decoder_grads_event(lagvae_text_plan* plan, int enable);
int lagvae_text_wait_decoder_grads(lagvae_text_plan* plan, void* stream);

/* ---------------------------------------------------------------------------------------------
 * Building blocks (exported for the unit/parity tests; also what the plan calls internally)
 * ------------------------------------------------------------------------------------------- */

/* C[M,N] (ldc) = alpha * sum_k A(m,k) B(n,k) + beta*C, A(m,k)=A[m*a_rs+k*a_cs], B(n,k)=B[n*b_rs+k*b_cs]
 * + optional bias_n[N] + optional bias_rows[(m % bias_period), N].  fp32 SIMT reference-grade GEMM
 * (replaces the cuBLAS sgemm calls behind nn.Linear / the LSTM projections). */
int lagvae_gemm_f32(const float* A, int64_t a_rs, int64_t a_cs, const float* B, int64_t b_rs,
                    int64_t b_cs, float* C, int64_t ldc, int M, int N, int K, float alpha, float beta,
                    const float* bias_n, const float* bias_rows, int bias_period, void* stream);

/* Tensor-core GEMM on tcgen05 (UMMA, fp32 accumulate in TMEM, TMA operand staging).  Operands are
 * bf16 split pairs (hi, lo) with x = hi + lo + O(2^-17 |x|); passes = 3 computes
 * hi*hi + hi*lo + lo*hi (fp32-grade), passes = 1 computes hi*hi only.
 *  A: a_mn_major == 0 -> [M, lda] row-major (K contiguous), else [K, lda] row-major (M contiguous)
 *  B: b_mn_major == 0 -> [N, ldb] row-major (K contiguous), else [K, ldb] row-major (N contiguous)
 *  lda/ldb in elements, multiples of 8; pointers 16-B aligned.  C fp32 [M, ldc]; epilogue as above.
 *  out_row_map (int32[M] or NULL) scatters output row m to C row out_row_map[m]. */
int lagvae_gemm_tc(const uint16_t* A_hi, const uint16_t* A_lo, int64_t lda, int a_mn_major,
                   const uint16_t* B_hi, const uint16_t* B_lo, int64_t ldb, int b_mn_major, float* C,
                   int64_t ldc, int M, int N, int K, int passes, float alpha, float beta,
                   const float* bias_n, const float* bias_rows, int bias_period,
                   const int32_t* out_row_map, void* stream);

/* LSTM recurrence of nn.LSTM (enc_lstm.py:60 / dec_lstm.py:104) given the input projection.
 * tier 0: launch-per-step fp32 SIMT; tier 1: persistent tcgen05 kernel (one cooperative launch for all
 * Tn steps; needs nh % 8 == 0, nh/8 <= #SMs, Bd <= 512 and a workspace of lagvae_lstm_workspace_bytes).
 *  gates [Tn*Bd, 4nh]: pre-activations (x W_ihᵀ + b) on entry, activated i,f,g,o on exit (the stash);
 *  h0/c0 [Bd,nh] or NULL (zeros); c_all/h_all [Tn*Bd,nh]; hdrop_all (or NULL) = h * dropout_out keep/scale
 *  (uses drop->p_out / mask_out / seed).  Backward: dh_ext [Tn*Bd,nh] gradient wrt the emitted h (scaled by
 *  the same dropout) or NULL; dh_last [Bd,nh] gradient on the final h or NULL; on exit dgates [Tn*Bd,4nh],
 *  dc = d c_{-1}, dh_rec = d h_{-1} (when want_init). */
size_t lagvae_lstm_workspace_bytes(int nh, int Bd);
int lagvae_lstm_forward(int tier, int nh, int Tn, int Bd, const float* w_hh, const float* h0,
                        const float* c0, float* gates, float* c_all, float* h_all, float* hdrop_all,
                        const lagvae_dropout* drop, void* workspace, size_t workspace_bytes, void* stream);
int lagvae_lstm_backward(int tier, int nh, int Tn, int Bd, const float* w_hh, const float* c0,
                         const float* gates, const float* c_all, const float* dh_ext, const float* dh_last,
                         const lagvae_dropout* drop, float* dc, float* dh_rec, float* dgates, int want_init,
                         void* workspace, size_t workspace_bytes, void* stream);

/* Profiling aid: when set (device uint64 buffer, >= 8*(Tn+1) words), CTA 0 of the persistent LSTM kernels records
 * clock64() at 5 points of every time step ([step][8]: start, first operand stage landed, MMAs issued,
 * accumulators complete, epilogue stores issued).  NULL disables (default). */
void lagvae_debug_trace_buffer(void* dev_u64, size_t words);

/* Which recurrence kernel the most recent LSTM forward (direction 0) / backward (direction 1) launch of this process
 * used: "v2/cs<N>" (persistent cluster kernel, cluster size N), "v1" (persistent, no cluster), "steps" (launch per
 * time step, fp32 SIMT) or "none".  A variant that was selected and then fails to launch is an error, never a silent
 * fallback; tests and bench.py assert the variant they mean to measure (nn.LSTM at enc_lstm.py:60, dec_lstm.py:104). */
const char* lagvae_lstm_variant(int direction);

/* fp32 [rows, cols] (ld) -> bf16 hi/lo [rows, ld_out] (zero padded columns cols..ld_out). */
int lagvae_split_bf16(const float* src, int64_t ld, int rows, int cols, uint16_t* hi, uint16_t* lo,
                      int64_t ld_out, void* stream);

/* ---------------------------------------------------------------------------------------------
 * Image path (ResNetEncoderV2 enc_resnet_v2.py:93-126, PixelCNNDecoderV2 dec_pixelcnn_v2.py:123-195).
 * Activations are NHWC fp32: a [B,H,W,C] tensor is the matrix [B*H*W rows, C]; a convolution is
 * im2col (identity for 1x1 stride 1) + lagvae_gemm_auto against the weight reshaped to [Cout, kh*kw*Cin]
 * (tap-major, channel-minor); its backward is two more GEMMs + col2im.
 * ------------------------------------------------------------------------------------------- */
/* nn.Conv2d patch gather / its adjoint (replaces cuDNN implicit-GEMM convolution).  Output spatial size
 * Ho = (H + 2 pad - kh)/stride + 1.  col [B*Ho*Wo, kh*kw*C]; dx [B,H,W,C] is overwritten. */
int lagvae_im2col(const float* x, int B, int H, int W, int C, int kh, int kw, int stride, int pad, float* col,
                  void* stream);
int lagvae_col2im(const float* dcol, int B, int H, int W, int C, int kh, int kw, int stride, int pad, float* dx,
                  void* stream);
/* nn.BatchNorm2d in train(): batch statistics over the R rows (biased variance), running stats updated in place
 * with the unbiased variance (momentum 0.1 default); scratch: device, >= 16*C bytes.  Backward returns dx, dgamma,
 * dbeta. */
int lagvae_bn_train_fwd(const float* x, int64_t R, int C, const float* gamma, const float* beta, float eps,
                        float momentum, float* y, float* save_mean, float* save_invstd, float* running_mean,
                        float* running_var, void* scratch, void* stream);
int lagvae_bn_train_bwd(const float* x, const float* dy, int64_t R, int C, const float* gamma, const float* save_mean,
                        const float* save_invstd, float* dx, float* dgamma, float* dbeta, void* scratch, void* stream);
/* eval()-mode BatchNorm: y = (x - mean) * invstd * gamma + beta with caller-provided per-channel statistics. */
int lagvae_bn_apply(const float* x, int64_t R, int C, const float* mean, const float* invstd, const float* gamma,
                    const float* beta, float* y, void* stream);
/* nn.ELU(alpha=1) of (a + b) (b may be NULL: plain ELU; with b: the residual add of ResNetBlock / PixelCNNBlock fused
 * in); backward from the OUTPUT y. */
int lagvae_elu_fwd(const float* a, const float* b_or_null, float* y, int64_t n, void* stream);
int lagvae_elu_bwd(const float* y, const float* dy, float* dx, int64_t n, void* stream);
int lagvae_add(const float* a, const float* b, float* out, int64_t n, void* stream);
/* nn.Sigmoid + the Bernoulli NLL of dec_pixelcnn_v2.py:172-195 (eps = 1e-12 inside both logs): logits [B*ns, P],
 * x [B, P] -> nll [B*ns]; backward: dlogits = g[row] * d nll / d logit. */
int lagvae_bernoulli_nll_fwd(const float* logits, const float* x, int B, int ns, int P, float* nll, void* stream);
int lagvae_bernoulli_nll_bwd(const float* logits, const float* x, const float* g, int B, int ns, int P, float* dlogits,
                             void* stream);
/* Reparameterise + analytic KL on given (mu, logvar) [B,nz] (encoder.py:55,72-79): z [B,ns,nz], kl [B]; backward
 * writes dml [B, 2nz] = (dmu | dlogvar) from dz (may be NULL) and the upstream gradient of KL (may be NULL). */
int lagvae_reparam_kl_fwd(const float* mu, const float* logvar, const float* eps, int B, int nz, int ns, float* z,
                          float* kl, void* stream);
int lagvae_reparam_kl_bwd(const float* dz, const float* eps, const float* mu, const float* logvar, const float* g_kl,
                          int B, int nz, int ns, float* dml, void* stream);
/* GEMM dispatcher used by the image layers and nn.Linear: C[M,N] = alpha * Σ_k A(m,k) B(n,k) + beta C + bias_n with
 * strided fp32 operands; stages split-bf16 copies into `scratch` and runs lagvae_gemm_tc (3 passes) when the problem
 * is tensor-core sized, else lagvae_gemm_f32. */
size_t lagvae_gemm_auto_scratch_bytes(int M, int N, int K);
int lagvae_gemm_auto(const float* A, int64_t a_rs, int64_t a_cs, const float* B, int64_t b_rs, int64_t b_cs, float* C,
                     int64_t ldc, int M, int N, int K, float alpha, float beta, const float* bias_n, void* scratch,
                     size_t scratch_bytes, void* stream);

/* ---------------------------------------------------------------------------------------------
 * Im2col-free convolutions on tcgen05 (csrc/conv_tc.cu) — the PixelCNNBlock layers (dec_pixelcnn_v2.py:32-62: 1x1 64->32,
 * MaskedConv2d k x k 32->32 :12-30, 1x1 32->64) and the head 1x1 64->64 (:145): Cin, Cout in {32, 64}, k odd <= 7,
 * stride 1, pad k/2, NHWC, tiles of whole image rows (H*W geometry must pass lagvae_convtc_supported).
 * Activations are staged once per layer in the "cat" format: bf16 [B,H,W,2C] = [hi(C) | lo(C)] per pixel
 * (lagvae_split_cat, or the fused BatchNorm/ELU kernels below); shifted operand tiles are 4-D TMA boxes with zero fill
 * outside the image.  mask_mode: 0 plain, 1 mask 'A', 2 mask 'B' (all input channels masked).  Only live taps are
 * multiplied in forward/dgrad; wgrad returns ALL k*k taps in the torch layout [Cout,Cin,kh,kw] (autograd of the
 * reference produces gradients for masked taps as well).  (Cin, Cout) always name the FORWARD convolution.
 * ------------------------------------------------------------------------------------------- */
int lagvae_convtc_supported(int B, int H, int W, int Cin, int Cout, int kh, int kw);
/* fp32 [rows, C] -> bf16 [rows, 2C] = [hi | lo], C in {32, 64} */
int lagvae_split_cat(const float* x, int64_t rows, int C, uint16_t* cat, void* stream);
size_t lagvae_convtc_wbuf_bytes(int Cin, int Cout, int kh, int kw);
/* w: fp32 [Cout,Cin,kh,kw] (torch layout) -> bf16 forward + dgrad weight tiles in wbuf (128-B aligned). */
int lagvae_convtc_prepare_weights(const float* w, int Cout, int Cin, int kh, int kw, int mask_mode, void* wbuf, void* stream);
/* y fp32 [B,H,W,Cout] = conv(x) (+ addend fp32 [B,H,W,Cout] when given); stats_or_null: device double[2 Cout] receiving
 * sum(y) | sum(y^2) per channel (zeroed inside) — the batch statistics of the BatchNorm that follows, accumulated in
 * the epilogue. */
int lagvae_convtc_forward(const uint16_t* xcat, const void* wbuf, int B, int H, int W, int Cin, int Cout, int kh, int kw,
                          int mask_mode, const float* addend_or_null, float* y, double* stats_or_null, void* stream);
/* dx fp32 [B,H,W,Cin] = conv_transpose(dy) (+ addend) from dycat [B,H,W,2 Cout] */
int lagvae_convtc_dgrad(const uint16_t* dycat, const void* wbuf, int B, int H, int W, int Cin, int Cout, int kh, int kw,
                        int mask_mode, const float* addend_or_null, float* dx, void* stream);
size_t lagvae_convtc_wgrad_scratch_bytes(int Cin, int Cout, int kh, int kw);
int lagvae_convtc_wgrad(const uint16_t* dycat, const uint16_t* xcat, int B, int H, int W, int Cin, int Cout, int kh, int kw,
                        float* dw, void* scratch, void* stream);

/* Fused BatchNorm2d(train) + residual + ELU (csrc/image_fused.cu; dec_pixelcnn_v2.py:40-49,61), C in {32, 64}, R rows.
 * forward: `stats` = double[2C] sums left by lagvae_convtc_forward; out = [ELU]( (y-mean)*invstd*gamma + beta [+ residual] )
 * written as fp32 and/or in the cat format; save_mean/save_invstd [C] for the backward; running statistics updated as
 * nn.BatchNorm2d does (unbiased variance).  backward: dout = gradient of `out`; the activation output is read back from
 * fp32 or cat (needed when elu != 0); emits dy (fp32 and/or cat), the residual-branch gradient dres = dout*ELU'(out),
 * dgamma, dbeta.  scratch: device, >= 16*C bytes. */
int lagvae_bnact_fwd(const float* y, const double* stats, int64_t R, int C, const float* gamma, const float* beta, float eps,
                     float momentum, const float* residual_or_null, int elu, float* out_f32_or_null, uint16_t* out_cat_or_null,
                     float* save_mean, float* save_invstd, float* running_mean, float* running_var, void* stream);
/* eval(): fill `stats` so that lagvae_bnact_fwd normalises with the given running statistics (pass NULL running pointers
 * there so nothing is updated). */
int lagvae_bn_eval_stats(const float* running_mean, const float* running_var, int64_t R, int C, double* stats, void* stream);
int lagvae_bnact_bwd(const float* dout, const float* out_f32_or_null, const uint16_t* out_cat_or_null, const float* y, int64_t R,
                     int C, const float* gamma, const float* save_mean, const float* save_invstd, int elu, float* dy_f32_or_null,
                     uint16_t* dy_cat_or_null, float* dres_or_null, float* dgamma, float* dbeta, void* scratch, void* stream);

/* PixelCNNBlock (dec_pixelcnn_v2.py:32-62) as one call per direction (csrc/image_plan.cu): conv1x1 C->Cm, BN, ELU, masked
 * k x k Cm->Cm (mask 'B'), BN, ELU, conv1x1 Cm->C, BN, out = ELU(. + x), training-mode BatchNorm.  `stash` (caller-owned,
 * 256-B aligned, lagvae_pixelblock_stash_bytes) keeps what the backward needs; `scratch` (lagvae_pixelblock_scratch_bytes)
 * is transient.  x/out fp32 [B,H,W,C]; xcat/outcat: their bf16 [hi|lo] operand copies (xcat NULL: made inside; outcat NULL:
 * not produced).  w2 must already be masked (MaskedConv2d does it in place, :29); wgrad returns all taps. */
typedef struct lagvae_pixelblock_dims {
  int32_t B, H, W, C, Cm, k;
  float eps, momentum;
  int32_t eval;   /* != 0: eval() — normalise with the running statistics, update nothing (forward only) */
} lagvae_pixelblock_dims;
typedef struct lagvae_pixelblock_params {
  const float *w1, *g1, *b1, *w2, *g2, *b2, *w3, *g3, *b3;   /* conv weights (torch layout), BN weight / bias */
  float *rm1, *rv1, *rm2, *rv2, *rm3, *rv3;                  /* BN running_mean / running_var (updated by forward) */
} lagvae_pixelblock_params;
typedef struct lagvae_pixelblock_grads {
  float *dw1, *dg1, *db1, *dw2, *dg2, *db2, *dw3, *dg3, *db3;
} lagvae_pixelblock_grads;
size_t lagvae_pixelblock_stash_bytes(const lagvae_pixelblock_dims* d);
size_t lagvae_pixelblock_scratch_bytes(const lagvae_pixelblock_dims* d);
int lagvae_pixelblock_forward(const lagvae_pixelblock_dims* d, const lagvae_pixelblock_params* p, const float* x,
                              const uint16_t* xcat_or_null, void* stash, float* out, uint16_t* outcat_or_null, void* stream);
int lagvae_pixelblock_backward(const lagvae_pixelblock_dims* d, const lagvae_pixelblock_params* p, const float* dout, const float* out,
                               const void* stash, const uint16_t* xcat_or_null, float* dx, const lagvae_pixelblock_grads* g,
                               void* scratch, void* stream);

/* materialise the Philox dropout keep-mask the kernels would use (tests feed it to the oracle) */
int lagvae_dropout_mask(uint64_t seed, uint32_t stream_id, int64_t n, float p, uint8_t* out_keep,
                        void* stream);

#ifdef __cplusplus
}
#endif
#endif /* LAGVAE_H_ */
