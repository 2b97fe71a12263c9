// Text-path plan: workspace carving and the launch sequence of VAE.loss forward / backward /
// fused inner step.  Host-side C++ only; every arithmetic stage is a kernel in
// kernels_simt.cu / gemm_tc.cu / lstm_tc.cu.
#include <algorithm>
#include <cstdlib>
#include <map>
#include <cstring>
#include <mutex>
#include <new>
#include <vector>

#include "kernels.cuh"
#include "lstm_tc.cuh"

using namespace lagvae;

// parameter slots (reference state_dict order)
enum { E_EMB = 0, E_WIH, E_WHH, E_BIH, E_BHH, E_LIN, D_EMB, D_TRANS, D_WIH, D_WHH, D_BIH, D_BHH, D_PRED };

struct Mat {  // stored row-major fp32 matrix
  const float* p;
  int64_t rows, cols, ld;
};
struct Staged {
  Mat m;
  TcOperand tc;  // valid when tc.hi != nullptr
};

struct lagvae_text_plan {
  lagvae_text_dims d;
  uint32_t flags;
  bool use_tc;
  int Bd, Te, Td;
  int64_t re, rd;
  int64_t ldl;  // leading dimension of the logits buffer (V padded to 8: 16-B aligned rows)
  char* base;
  size_t bytes;
  // forward stash
  float *xe, *gates_e, *c_e, *h_e, *bsum_e, *mu, *logvar, *z, *kl, *eps;
  float *xd, *zb, *bsum_d, *c0, *h0, *gates_d, *c_d, *h_d, *hdrop_d, *logits, *lse, *loss_row, *scalars;
  // backward scratch
  float *g_rec, *g_kl, *dh_d, *dgates_d, *dx_d, *dc, *dh_rec, *dzb, *dc0t, *dz, *dml, *dh_last;
  float *dgates_e, *dx_e, *dc_e, *dh_rec_e;
  void* clip_scratch;
  // tensor-core operand arena (bf16 hi/lo staging), bump allocated per pass
  char* arena;
  size_t arena_bytes, arena_off;
  // decoder weights do not change inside the aggressive loop (text.py:387 steps the encoder only): their bf16 hi/lo splits are
  // cached across calls while the caller-declared epoch (lagvae_text_decoder_weights_epoch) stays the same; 0 = never cache
  // The cache sits at the START of the workspace (same address for every plan of an engine: plans of different (B,T,ns)
  // share one workspace) and its state is shared through a registry keyed by that address (WCacheState below).
  char* wcache;
  size_t wcache_bytes;
  uint64_t dec_epoch;
  struct WCacheState* wc;
  LstmTcState* lstm_tc;  // persistent tcgen05 LSTM state (nullptr when unsupported / SIMT)
  cudaEvent_t dec_ev;    // optional: recorded when the decoder gradients are final (data-parallel overlap hook)
  bool dec_ev_recorded;
  // side stream for the norm-only dW_pred GEMM: it runs on the 20 SMs the 128-CTA persistent decoder recurrence leaves
  // idle (scripts/microbench/overlap_probe.py: both launch orders overlap fully) and is joined before the decoder
  // gradients are declared final
  cudaStream_t side;
  cudaEvent_t side_fork, side_join;
  size_t arena_floor;    // staging below this offset is still read by the side stream
  unsigned* side_gate;   // device word: set by the decoder backward recurrence once its grid is resident (wait_flag on `side`)
  // state carried from forward to backward
  Staged st_xe, st_xd, st_h;   // forward's bf16 hi/lo operand copies of xe, xd and hdrop/h_d: the backward GEMMs reuse them
  size_t fwd_arena_end;        // ... so the backward's bump allocation starts behind them
  lagvae_dropout drop;
  float kl_weight;
  bool have_forward;
  // fused steps only: the upstream gradient of every reconstruction row is known before the forward runs (> 0), so the
  // forward's cross-entropy pass also emits dlogits as the split-bf16 operand of the backward GEMMs (st_dl)
  float ce_upstream;
  Staged st_dl;
  bool dl_ready;
  int dec_wgrad_passes;  // 3 = fp32-grade; 1 = single bf16 pass (fused inner step: these gradients only feed the clip norm)
};

struct WCacheState {
  uint64_t epoch = 0;            // epoch the cached splits belong to (0 = nothing cached)
  bool filled[2] = {false, false};
};

namespace {

std::mutex g_wc_mutex;
std::map<void*, WCacheState*> g_wc_registry;   // workspace base -> shared cache state (entries live for the process)
WCacheState* wcache_state_for(void* base) {
  std::lock_guard<std::mutex> lk(g_wc_mutex);
  auto it = g_wc_registry.find(base);
  if (it != g_wc_registry.end()) return it->second;
  WCacheState* s = new WCacheState();
  g_wc_registry[base] = s;
  return s;
}

struct Carver {
  char* base;
  size_t off;
  template <class T>
  T* take(int64_t n) {
    off = (size_t)round_up((int64_t)off, 256);
    T* p = base ? (T*)(base + off) : nullptr;
    off += (size_t)n * sizeof(T);
    return p;
  }
};

size_t arena_need(const lagvae_text_dims& d) {
  // generous bound: every staged operand of the backward pass (the larger pass), hi+lo bf16
  const int64_t Bd = (int64_t)d.B * d.ns, re = (int64_t)d.T * d.B, rd = (int64_t)(d.T - 1) * Bd;
  auto pad8 = [](int64_t c) { return round_up(c, 8); };
  int64_t el = 0;
  el += rd * pad8(d.V);                        // dlogits
  el += (int64_t)d.V * pad8(d.nh);             // W_pred
  el += 2 * rd * pad8(d.nh);                   // hdrop, h_d
  el += rd * pad8(4 * d.nh) + rd * pad8(d.ni); // dgates_d, xd
  el += re * pad8(4 * d.nh) + re * pad8(d.ni) + re * pad8(d.nh);  // dgates_e, xe, h_e
  el += 2 * ((int64_t)4 * d.nh * pad8(d.ni + d.nz) + (int64_t)4 * d.nh * pad8(d.nh));  // LSTM weights
  el += Bd * pad8(d.nh) * 4;
  el += re * pad8(d.ni) + rd * pad8(d.ni) + rd * pad8(d.nh);   // forward's xe | xd | h copies kept alive under the backward pass
  return (size_t)el * 2 * sizeof(uint16_t) + (64 << 10);
}

void carve(lagvae_text_plan* P, char* base) {
  const lagvae_text_dims& d = P->d;
  const int64_t B = d.B, Bd = P->Bd, re = P->re, rd = P->rd, nh = d.nh, ni = d.ni, nz = d.nz, V = d.V;
  Carver c{base, 0};
  // first: the cross-call decoder-weight cache — its size depends on (V, ni, nh) only, so every plan of an engine finds it at
  // the same address
  P->wcache_bytes = P->use_tc ? (size_t)(((int64_t)4 * nh * round_up(ni, 8) + V * round_up(nh, 8)) * 2 * sizeof(uint16_t) + 2048) : 0;
  P->wcache = c.take<char>((int64_t)P->wcache_bytes);
  P->xe = c.take<float>(re * ni);
  P->gates_e = c.take<float>(re * 4 * nh);
  P->c_e = c.take<float>(re * nh);
  P->h_e = c.take<float>(re * nh);
  P->bsum_e = c.take<float>(4 * nh);
  P->mu = c.take<float>(B * nz);
  P->logvar = c.take<float>(B * nz);
  P->z = c.take<float>(Bd * nz);
  P->kl = c.take<float>(B);
  P->eps = c.take<float>(Bd * nz);
  P->xd = c.take<float>(rd * ni);
  P->zb = c.take<float>(Bd * 4 * nh);
  P->bsum_d = c.take<float>(4 * nh);
  P->c0 = c.take<float>(Bd * nh);
  P->h0 = c.take<float>(Bd * nh);
  P->gates_d = c.take<float>(rd * 4 * nh);
  P->c_d = c.take<float>(rd * nh);
  P->h_d = c.take<float>(rd * nh);
  P->hdrop_d = c.take<float>(rd * nh);
  P->ldl = round_up(V, 8);
  P->logits = c.take<float>(rd * P->ldl);
  P->lse = c.take<float>(rd);
  P->loss_row = c.take<float>(rd);
  P->scalars = c.take<float>(8);
  P->g_rec = c.take<float>(B);
  P->g_kl = c.take<float>(B);
  P->dh_d = c.take<float>(rd * nh);
  P->dgates_d = c.take<float>(rd * 4 * nh);
  P->dx_d = c.take<float>(rd * ni);
  P->dc = c.take<float>(Bd * nh);
  P->dh_rec = c.take<float>(Bd * nh);
  P->dzb = c.take<float>(Bd * 4 * nh);
  P->dc0t = c.take<float>(Bd * nh);
  P->dz = c.take<float>(Bd * nz);
  P->dml = c.take<float>(B * 2 * nz);
  P->dh_last = c.take<float>(B * nh);
  P->dgates_e = c.take<float>(re * 4 * nh);
  P->dx_e = c.take<float>(re * ni);
  P->dc_e = c.take<float>(B * nh);
  P->dh_rec_e = c.take<float>(B * nh);
  P->clip_scratch = c.take<char>(16384);
  P->side_gate = c.take<unsigned>(64);
  P->arena_bytes = P->use_tc ? arena_need(d) : 0;
  P->arena = c.take<char>((int64_t)P->arena_bytes);
  c.off = (size_t)round_up((int64_t)c.off, 256);
  P->bytes = c.off;
}

// start of a forward-type entry point: nothing staged, nothing kept for a backward pass
void reset_pass(lagvae_text_plan* P) {
  P->arena_off = 0;
  P->fwd_arena_end = 0;
  P->st_xe = P->st_xd = P->st_h = P->st_dl = Staged{Mat{nullptr, 0, 0, 0}, TcOperand{nullptr, nullptr, 0, 0}};
  P->have_forward = false;
  P->dl_ready = false;
}

bool dims_ok(const lagvae_text_dims* d) {
  return d && d->B > 0 && d->T >= 2 && d->ns > 0 && d->V > 1 && d->ni > 0 && d->nh > 0 && d->nz > 0;
}

bool want_tc(const lagvae_text_dims& d, uint32_t flags) {
  if (flags & LAGVAE_PLAN_FORCE_SIMT) return false;
  // the tcgen05 path pays off (and is validated) for tensor-core sized problems only
  return d.nh >= 64 && d.ni >= 32 && d.V >= 128;
}

Staged stage(lagvae_text_plan* P, Mat m, cudaStream_t st, int* status) {
  Staged s;
  s.m = m;
  s.tc = TcOperand{nullptr, nullptr, 0, 0};
  if (!P->use_tc) return s;
  const int64_t ldo = round_up(m.cols, 8);
  const size_t need = (size_t)m.rows * ldo * sizeof(uint16_t);
  size_t off = (size_t)round_up((int64_t)P->arena_off, 256);
  if (off + 2 * need + 512 > P->arena_bytes) {
    set_error("text plan: tensor-core staging arena exhausted");
    *status = LAGVAE_E_WORKSPACE;
    return s;
  }
  uint16_t* hi = (uint16_t*)(P->arena + off);
  off = (size_t)round_up((int64_t)(off + need), 256);
  uint16_t* lo = (uint16_t*)(P->arena + off);
  P->arena_off = off + need;
  const int r = split_bf16_launch(m.p, m.ld, (int)m.rows, (int)m.cols, hi, lo, ldo, st);
  if (r != LAGVAE_OK) *status = r;
  s.tc = TcOperand{hi, lo, ldo, 0};
  return s;
}

Staged stage_alloc(lagvae_text_plan* P, int64_t rows, int64_t cols, int* status);

// embedding gather (+ input dropout) producing the fp32 rows and, on the tensor-core tier, the bf16 hi/lo operand in the SAME
// pass (k_embed_gather_split); falls back to gather + stage when the table rows are not float4-aligned
Staged gather_staged(lagvae_text_plan* P, const int64_t* x, int64_t x_ld, int B, int ns, int Tn, const float* table, int ni,
                     DropSpec drop, float* out, int64_t rows, cudaStream_t st, int* status) {
  static const bool off = [] { const char* e = getenv("LAGVAE_GATHER_SPLIT"); return e && e[0] == '0'; }();
  if (P->use_tc && !off && (ni & 3) == 0 && ((uintptr_t)table & 15) == 0 && ((uintptr_t)out & 15) == 0) {
    Staged s = stage_alloc(P, rows, ni, status);
    if (*status != LAGVAE_OK) return s;
    s.m = Mat{out, rows, ni, ni};
    const int r = embed_gather_split(x, x_ld, 0, B, ns, Tn, table, ni, drop, out, const_cast<uint16_t*>(s.tc.hi),
                                     const_cast<uint16_t*>(s.tc.lo), s.tc.ld, st);
    if (r != LAGVAE_OK) *status = r;
    return s;
  }
  const int r = embed_gather(x, x_ld, 0, B, ns, Tn, table, ni, drop, out, st);
  if (r != LAGVAE_OK) {
    *status = r;
    return Staged{Mat{out, rows, ni, ni}, TcOperand{nullptr, nullptr, 0, 0}};
  }
  return stage(P, Mat{out, rows, ni, ni}, st, status);
}

// decoder weight operands through the cross-call cache: which = 0 -> x-columns of W_ih ([4nh, ni], ld ni+nz), 1 -> W_pred [V, nh]
Staged stage_dec_weight(lagvae_text_plan* P, int which, Mat m, cudaStream_t st, int* status) {
  if (!P->use_tc || P->dec_epoch == 0 || P->wc == nullptr) return stage(P, m, st, status);
  WCacheState* wc = P->wc;
  if (wc->epoch != P->dec_epoch) {         // the weights changed (or nothing cached yet): both entries are stale
    wc->epoch = P->dec_epoch;
    wc->filled[0] = wc->filled[1] = false;
  }
  const lagvae_text_dims& d = P->d;
  const int64_t ld0 = round_up(d.ni, 8), ld1 = round_up(d.nh, 8);
  uint16_t* base0 = (uint16_t*)P->wcache;
  uint16_t* base1 = (uint16_t*)round_up((int64_t)(uintptr_t)(base0 + (int64_t)2 * 4 * d.nh * ld0), 256);
  uint16_t* hi = which == 0 ? base0 : base1;
  const int64_t ldo = which == 0 ? ld0 : ld1;
  uint16_t* lo = hi + m.rows * ldo;
  if (!wc->filled[which]) {                // filled on the first request within the epoch (same stream order as its readers)
    const int r = split_bf16_launch(m.p, m.ld, (int)m.rows, (int)m.cols, hi, lo, ldo, st);
    if (r != LAGVAE_OK) *status = r;
    wc->filled[which] = true;
  }
  Staged s;
  s.m = m;
  s.tc = TcOperand{hi, lo, ldo, 0};
  return s;
}

// arena allocation only (the producer kernel writes the hi/lo parts itself); fp32 view is absent
Staged stage_alloc(lagvae_text_plan* P, int64_t rows, int64_t cols, int* status) {
  Staged s;
  s.m = Mat{nullptr, rows, cols, cols};
  s.tc = TcOperand{nullptr, nullptr, 0, 0};
  const int64_t ldo = round_up(cols, 8);
  const size_t need = (size_t)rows * ldo * sizeof(uint16_t);
  size_t off = (size_t)round_up((int64_t)P->arena_off, 256);
  if (off + 2 * need + 512 > P->arena_bytes) {
    set_error("text plan: tensor-core staging arena exhausted");
    *status = LAGVAE_E_WORKSPACE;
    return s;
  }
  uint16_t* hi = (uint16_t*)(P->arena + off);
  off = (size_t)round_up((int64_t)(off + need), 256);
  uint16_t* lo = (uint16_t*)(P->arena + off);
  P->arena_off = off + need;
  s.tc = TcOperand{hi, lo, ldo, 0};
  return s;
}

// C[M,N] = alpha * opA · opBᵀ (+beta C)(+bias).  A used as [M,K]: stored [M,K] (a_t = false) or
// stored [K,M] (a_t = true).  B used as [N,K]: stored [N,K] (b_t = false) or stored [K,N].
int mm(lagvae_text_plan* P, const Staged& A, bool a_t, const Staged& B, bool b_t, float* C, int64_t ldc,
       int M, int N, int K, float alpha, float beta, const float* bias_n, const float* bias_rows,
       int bias_period, int passes, cudaStream_t st) {
  const bool tc_ok = P->use_tc && A.tc.hi && B.tc.hi && M >= 32 && N >= 16 && K >= 16;
  if (tc_ok) {
    TcOperand a = A.tc, b = B.tc;
    a.mn_major = a_t ? 1 : 0;
    b.mn_major = b_t ? 1 : 0;
    return gemm_tc(a, b, C, ldc, M, N, K, passes, alpha, beta, bias_n, bias_rows, bias_period, nullptr, st);
  }
  return gemm_f32(A.m.p, a_t ? 1 : A.m.ld, a_t ? A.m.ld : 1, B.m.p, b_t ? 1 : B.m.ld, b_t ? B.m.ld : 1, C,
                  ldc, M, N, K, alpha, beta, bias_n, bias_rows, bias_period, st);
}
// sub-view of a staged matrix: rows [r0, r0+nr), cols [c0, c0+nc) (c0 multiple of 8 for tc)
Staged sub(const Staged& s, int64_t r0, int64_t nr, int64_t c0, int64_t nc) {
  Staged o = s;
  o.m.p = s.m.p + r0 * s.m.ld + c0;
  o.m.rows = nr;
  o.m.cols = nc;
  if (s.tc.hi) {
    if (c0 % 8 != 0) {
      o.tc.hi = o.tc.lo = nullptr;
    } else {
      o.tc.hi = s.tc.hi + r0 * s.tc.ld + c0;
      o.tc.lo = s.tc.lo + r0 * s.tc.ld + c0;
    }
  }
  return o;
}

DropSpec spec_in(const lagvae_dropout& d) {
  DropSpec s{};
  const bool on = d.mode != 0 && d.p_in > 0.f;
  s.mode = on ? d.mode : 0;
  s.p = d.p_in;
  s.scale = on ? 1.f / (1.f - d.p_in) : 1.f;
  s.mask = d.mask_in;
  s.seed = d.seed;
  s.seed_dev = s.mode == 2 ? d.seed_dev : nullptr;
  s.sid = 1;
  return s;
}
DropSpec spec_out(const lagvae_dropout& d) {
  DropSpec s{};
  const bool on = d.mode != 0 && d.p_out > 0.f;
  s.mode = on ? d.mode : 0;
  s.p = d.p_out;
  s.scale = on ? 1.f / (1.f - d.p_out) : 1.f;
  s.mask = d.mask_out;
  s.seed = d.seed;
  s.seed_dev = s.mode == 2 ? d.seed_dev : nullptr;
  s.sid = 2;
  return s;
}
DropSpec spec_none() {
  DropSpec s{};
  s.mode = 0;
  s.scale = 1.f;
  return s;
}

// ---- LSTM recurrence, launch-per-step tier (fp32 SIMT GEMM + point-wise cell) -------------------
// gates [Tn*Bd, 4nh] holds the input projection on entry, activated gates on exit.
int lstm_forward_steps(const float* w_hh, const float* h0, const float* c0, float* gates, float* c_all,
                       float* h_all, float* hdrop_all, DropSpec drop, int Tn, int Bd, int nh,
                       cudaStream_t st) {
  const int64_t gs = (int64_t)Bd * 4 * nh, hs = (int64_t)Bd * nh;
  lstm_note_variant(0, "steps");
  for (int t = 0; t < Tn; ++t) {
    const float* hp = t ? h_all + (t - 1) * hs : h0;
    const float* cp = t ? c_all + (t - 1) * hs : c0;
    if (hp)
      LV_TRY(gemm_f32(hp, nh, 1, w_hh, nh, 1, gates + t * gs, 4 * nh, Bd, 4 * nh, nh, 1.f, 1.f, nullptr,
                      nullptr, 0, st));
    LV_TRY(lstm_point_fwd(gates + t * gs, cp, c_all + t * hs, h_all + t * hs,
                          hdrop_all ? hdrop_all + t * hs : nullptr, drop, t, Tn, Bd, nh, st));
  }
  return LAGVAE_OK;
}
// dh_ext: [Tn*Bd, nh] gradient wrt the emitted h (scaled by `drop` keep factors) or nullptr;
// dh_last: [Bd, nh] extra gradient on the final h only (encoder) or nullptr.
// On exit dc holds d c_{-1} and dh_rec holds d h_{-1} (when want_init).
int lstm_backward_steps(const float* w_hh, const float* c0, const float* gates, const float* c_all,
                        const float* dh_ext, DropSpec drop, const float* dh_last, float* dc,
                        float* dh_rec, float* dgates, int Tn, int Bd, int nh, bool want_init,
                        cudaStream_t st) {
  const int64_t gs = (int64_t)Bd * 4 * nh, hs = (int64_t)Bd * nh;
  LV_TRY(fill(dc, 0.f, hs, st));
  lstm_note_variant(1, "steps");
  for (int t = Tn - 1; t >= 0; --t) {
    // dh_last (gradient on the final h only) adds to dh_ext at t = Tn-1; it rides in the dh_rec slot there
    const float* ext = dh_ext ? dh_ext + t * hs : (t == Tn - 1 ? dh_last : nullptr);
    const DropSpec ds = dh_ext ? drop : spec_none();
    const float* cp = t ? c_all + (t - 1) * hs : c0;
    const float* rec = t == Tn - 1 ? (dh_ext ? dh_last : nullptr) : dh_rec;
    LV_TRY(lstm_point_bwd(gates + t * gs, c_all + t * hs, cp, ext, ds, rec, dc,
                          dgates + t * gs, t, Tn, Bd, nh, st));
    if (t > 0 || want_init)  // dh_{t-1} = dG_t · W_hh
      LV_TRY(gemm_f32(dgates + t * gs, 4 * nh, 1, w_hh, 1, nh, dh_rec, nh, Bd, nh, 4 * nh, 1.f, 0.f, nullptr,
                      nullptr, 0, st));
  }
  return LAGVAE_OK;
}

// fork (optional): recorded on `st` right before the recurrence is launched, i.e. after the input-projection GEMM — work
// that a side stream runs behind it overlaps the recurrence and nothing else
int encoder_forward(lagvae_text_plan* P, const lagvae_text_params* w, const int64_t* x, cudaStream_t st,
                    cudaEvent_t fork = nullptr) {
  const lagvae_text_dims& d = P->d;
  const int nh = d.nh, ni = d.ni;
  int status = LAGVAE_OK;
  Staged sx = gather_staged(P, x, d.T, d.B, 1, P->Te, w->p[E_EMB], ni, spec_none(), P->xe, P->re, st, &status);   // enc_lstm.py:58
  LV_TRY(status);
  LV_TRY(vec_add(w->p[E_BIH], w->p[E_BHH], P->bsum_e, 4 * nh, st));
  P->st_xe = sx;                                  // kept for the backward pass (dW_ih = dGᵀ·X)
  const size_t keep = P->arena_off;
  Staged sw = stage(P, Mat{w->p[E_WIH], 4 * nh, ni, ni}, st, &status);
  LV_TRY(status);
  // input projection for all steps (first half of nn.LSTM, enc_lstm.py:60)
  LV_TRY(mm(P, sx, false, sw, false, P->gates_e, 4 * nh, (int)P->re, 4 * nh, ni, 1.f, 0.f, P->bsum_e, nullptr,
            0, 3, st));
  P->arena_off = keep;                            // the W_ih copy is dead once the GEMM is enqueued (stream order)
  P->fwd_arena_end = keep;
  if (fork) LV_CUDA(cudaEventRecord(fork, st));
  if (P->lstm_tc)
    LV_TRY(lstm_tc_forward(P->lstm_tc, w->p[E_WHH], nullptr, nullptr, P->gates_e, P->c_e, P->h_e, nullptr, spec_none(),
                           P->Te, d.B, st));
  else
    LV_TRY(lstm_forward_steps(w->p[E_WHH], nullptr, nullptr, P->gates_e, P->c_e, P->h_e, nullptr, spec_none(),
                              P->Te, d.B, nh, st));
  return LAGVAE_OK;
}


// decoder forward up to per-token CE (dec_lstm.py:66-148).  z: device [Bd, nz].
// x_ld: row stride of the token tensor (d.T for the [B,T] training batches, T-1 for LSTMDecoder.decode's `input`);
// with_ce = false stops after the vocabulary projection (targets are not read).
// The z-independent front of the decoder forward: embedding + input dropout (dec_lstm.py:80-81, 87-91) and the x-columns of
// the input projection (first half of nn.LSTM, :104).  Returns whether the tensor-core GEMM ran (then the row-periodic z
// bias is still to be added).  `cap` > 0 limits the GEMM grid (side stream under a persistent recurrence).
int decoder_xproj(lagvae_text_plan* P, const lagvae_text_params* w, const int64_t* x, const lagvae_dropout& dr,
                  cudaStream_t st, int64_t x_ld, int cap, bool* ran_tc) {
  const lagvae_text_dims& d = P->d;
  const int nh = d.nh, ni = d.ni, nz = d.nz, B = d.B, ns = d.ns, Td = P->Td;
  int status = LAGVAE_OK;
  *ran_tc = false;
  Staged sxd = gather_staged(P, x, x_ld, B, ns, Td, w->p[D_EMB], ni, spec_in(dr), P->xd, P->rd, st, &status);   // :80-81 (+:87-91)
  LV_TRY(status);
  P->st_xd = sxd;
  Staged swd = stage_dec_weight(P, 0, Mat{w->p[D_WIH], 4 * nh, ni, ni + nz}, st, &status);  // x-columns of W_ih
  LV_TRY(status);
  if (P->use_tc && sxd.tc.hi && swd.tc.hi) {
    gemm_tc_set_grid_cap(cap);
    const int r = mm(P, sxd, false, swd, false, P->gates_d, 4 * nh, (int)P->rd, 4 * nh, ni, 1.f, 0.f, nullptr, nullptr, 0, 3, st);
    gemm_tc_set_grid_cap(0);
    LV_TRY(r);
    *ran_tc = true;
  }
  return LAGVAE_OK;
}

// xproj_state: 0 = run the front here; 1 = decoder_xproj already ran (tensor-core GEMM done, bias pending); 2 = it ran but
// fell to the non-tensor-core tier (GEMM + bias still to do here)
int decoder_forward(lagvae_text_plan* P, const lagvae_text_params* w, const int64_t* x, const float* z,
                    const lagvae_dropout& dr, cudaStream_t st, int64_t x_ld = -1, bool with_ce = true, int xproj_state = 0) {
  const lagvae_text_dims& d = P->d;
  const int nh = d.nh, ni = d.ni, nz = d.nz, V = d.V, B = d.B, ns = d.ns, Bd = P->Bd, Td = P->Td;
  int status = LAGVAE_OK;
  // ---- decoder: dec_lstm.py:66-111
  const DropSpec dout = spec_out(dr);
  if (x_ld < 0) x_ld = d.T;
  if (xproj_state == 0) {
    bool ran = false;
    LV_TRY(decoder_xproj(P, w, x, dr, st, x_ld, 0, &ran));
    xproj_state = ran ? 1 : 2;
  }
  LV_TRY(vec_add(w->p[D_BIH], w->p[D_BHH], P->bsum_d, 4 * nh, st));
  // z enters every step through the last nz input columns (:84,97): time-invariant row bias
  LV_TRY(gemm_f32(z, nz, 1, w->p[D_WIH] + ni, ni + nz, 1, P->zb, 4 * nh, Bd, 4 * nh, nz, 1.f, 0.f,
                  P->bsum_d, nullptr, 0, st));
  LV_TRY(gemm_f32(z, nz, 1, w->p[D_TRANS], nz, 1, P->c0, nh, Bd, nh, nz, 1.f, 0.f, nullptr, nullptr, 0,
                  st));                                                                      // :100
  LV_TRY(tanh_copy(P->c0, P->h0, Bd * nh, st));                                              // :101
  const float* rec_bias = nullptr;
  if (xproj_state == 1) {
    // tensor-core tier: plain GEMM (decoder_xproj); the row-periodic z bias is added by the persistent recurrence kernel as it
    // loads the pre-activations (or in one streaming pass, k_add_row_periodic, on the tiers that cannot)
    if (P->lstm_tc) rec_bias = P->zb;
    else LV_TRY(add_row_periodic(P->gates_d, P->zb, P->rd, 4 * nh, Bd, st));
  } else {
    Staged swd = stage_dec_weight(P, 0, Mat{w->p[D_WIH], 4 * nh, ni, ni + nz}, st, &status);
    LV_TRY(status);
    LV_TRY(mm(P, P->st_xd, false, swd, false, P->gates_d, 4 * nh, (int)P->rd, 4 * nh, ni, 1.f, 0.f, nullptr, P->zb, Bd,
              3, st));
  }
  float* hdrop = dout.mode ? P->hdrop_d : nullptr;
  if (P->lstm_tc)
    LV_TRY(lstm_tc_forward(P->lstm_tc, w->p[D_WHH], P->h0, P->c0, P->gates_d, P->c_d, P->h_d, hdrop, dout, Td, Bd, st, rec_bias));
  else
    LV_TRY(lstm_forward_steps(w->p[D_WHH], P->h0, P->c0, P->gates_d, P->c_d, P->h_d, hdrop, dout, Td, Bd, nh,
                              st));                                                          // :104,106
  // vocabulary projection (:109) + cross entropy (:143-148)
  Staged sh = stage(P, Mat{hdrop ? hdrop : P->h_d, P->rd, nh, nh}, st, &status);
  P->st_h = sh;
  P->fwd_arena_end = P->arena_off;     // xe | xd | h stay alive for the backward pass
  Staged swp = stage_dec_weight(P, 1, Mat{w->p[D_PRED], V, nh, nh}, st, &status);
  LV_TRY(status);
  LV_TRY(mm(P, sh, false, swp, false, P->logits, P->ldl, (int)P->rd, V, nh, 1.f, 0.f, nullptr, nullptr, 0, 3, st));
  if (with_ce) {
    static const bool fused_env = [] { const char* e = getenv("LAGVAE_FUSED_CE"); return !(e && e[0] == '0'); }();
    bool done = false;
    if (fused_env && P->ce_upstream > 0.f && P->use_tc && P->rd >= 32 && nh >= 16 && V >= 32) {
      // same bump allocation the backward makes first (arena_off == fwd_arena_end here), so it finds the operand in place
      Staged sdl = stage_alloc(P, P->rd, V, &status);
      LV_TRY(status);
      LV_TRY(ce_fused(P->logits, P->ldl, V, x, d.T, Td, Bd, ns, P->ce_upstream / (float)ns, P->lse, P->loss_row,
                      const_cast<uint16_t*>(sdl.tc.hi), const_cast<uint16_t*>(sdl.tc.lo), sdl.tc.ld, &done, st));
      if (done) {
        P->st_dl = sdl;
        P->dl_ready = true;
      }
    }
    if (!done) LV_TRY(ce_fwd(P->logits, P->ldl, V, x, d.T, Td, Bd, ns, P->lse, P->loss_row, st));
  }
  return LAGVAE_OK;
}

}  // namespace

// =============================================================================================
// C ABI
// =============================================================================================
extern "C" {

int lagvae_abi_version(void) { return LAGVAE_ABI_VERSION; }
const char* lagvae_last_error(void) { return lagvae::last_error(); }
int64_t lagvae_launch_count(void) { return lagvae::g_launches.load(); }

int lagvae_device_check(void) {
  int dev = 0;
  LV_CUDA(cudaGetDevice(&dev));
  cudaDeviceProp prop;
  LV_CUDA(cudaGetDeviceProperties(&prop, dev));
  if (prop.major != 10) {
    set_error("lagvae: device %d is sm_%d%d; this library contains sm_100a code only", dev, prop.major,
              prop.minor);
    return LAGVAE_E_CUDA;
  }
  return LAGVAE_OK;
}

int64_t lagvae_text_param_count(const lagvae_text_dims* d) {
  if (!dims_ok(d)) return -1;
  const int64_t V = d->V, ni = d->ni, nh = d->nh, nz = d->nz;
  return V * ni + 4 * nh * ni + 4 * nh * nh + 8 * nh + 2 * nz * nh      // encoder
         + V * ni + nh * nz + 4 * nh * (ni + nz) + 4 * nh * nh + 8 * nh + V * nh;  // decoder
}

size_t lagvae_text_workspace_bytes(const lagvae_text_dims* d, uint32_t flags) {
  if (!dims_ok(d)) return 0;
  lagvae_text_plan P{};
  P.d = *d;
  P.flags = flags;
  P.use_tc = want_tc(*d, flags);
  P.Bd = d->B * d->ns;
  P.Te = d->T;
  P.Td = d->T - 1;
  P.re = (int64_t)P.Te * d->B;
  P.rd = (int64_t)P.Td * P.Bd;
  carve(&P, nullptr);
  return P.bytes + lstm_tc_workspace_bytes(*d, P.use_tc);
}

int lagvae_text_plan_create(const lagvae_text_dims* d, uint32_t flags, void* workspace,
                            size_t workspace_bytes, lagvae_text_plan** out) {
  LV_CHECK_ARG(dims_ok(d), "text plan: bad dims");
  LV_CHECK_ARG(out != nullptr && workspace != nullptr, "text plan: null workspace/out");
  LV_CHECK_ARG(((uintptr_t)workspace & 255) == 0, "text plan: workspace must be 256-B aligned");
  const size_t need = lagvae_text_workspace_bytes(d, flags);
  if (workspace_bytes < need) {
    set_error("text plan: workspace too small (%zu < %zu)", workspace_bytes, need);
    return LAGVAE_E_WORKSPACE;
  }
  LV_TRY(lagvae_device_check());
  lagvae_text_plan* P = new (std::nothrow) lagvae_text_plan{};
  LV_CHECK_ARG(P != nullptr, "text plan: host allocation failed");
  P->d = *d;
  P->flags = flags;
  P->use_tc = want_tc(*d, flags);
  P->Bd = d->B * d->ns;
  P->Te = d->T;
  P->Td = d->T - 1;
  P->re = (int64_t)P->Te * d->B;
  P->rd = (int64_t)P->Td * P->Bd;
  P->base = (char*)workspace;
  carve(P, P->base);
  P->lstm_tc = nullptr;
  P->side = nullptr;
  P->side_fork = P->side_join = nullptr;
  P->arena_floor = 0;
  // LAGVAE_NO_LSTM_TC=1 keeps the tensor-core GEMMs but runs the recurrence on the launch-per-step tier
  const char* no_rec = getenv("LAGVAE_NO_LSTM_TC");
  const bool rec_tc = P->use_tc && !(no_rec && no_rec[0] == '1');
  const int r = lstm_tc_create(*d, rec_tc, P->base + P->bytes, workspace_bytes - P->bytes, &P->lstm_tc);
  if (r != LAGVAE_OK) {
    delete P;
    return r;
  }
  P->have_forward = false;
  P->ce_upstream = 0.f;
  P->dl_ready = false;
  P->dec_wgrad_passes = 3;
  P->dec_epoch = 0;
  P->wc = wcache_state_for(workspace);
  *out = P;
  return LAGVAE_OK;
}

void lagvae_text_plan_destroy(lagvae_text_plan* P) {
  if (!P) return;
  lstm_tc_destroy(P->lstm_tc);
  if (P->dec_ev) cudaEventDestroy(P->dec_ev);
  if (P->side_fork) cudaEventDestroy(P->side_fork);
  if (P->side_join) cudaEventDestroy(P->side_join);
  if (P->side) cudaStreamDestroy(P->side);
  delete P;
}

int lagvae_text_decoder_weights_epoch(lagvae_text_plan* P, uint64_t epoch) {
  LV_CHECK_ARG(P, "decoder_weights_epoch: null plan");
  P->dec_epoch = epoch;
  return LAGVAE_OK;
}

int lagvae_text_decoder_grads_event(lagvae_text_plan* P, int enable) {
  LV_CHECK_ARG(P, "decoder_grads_event: null plan");
  if (enable && !P->dec_ev) {
    LV_CUDA(cudaEventCreateWithFlags(&P->dec_ev, cudaEventDisableTiming));
    P->dec_ev_recorded = false;
  } else if (!enable && P->dec_ev) {
    LV_CUDA(cudaEventDestroy(P->dec_ev));
    P->dec_ev = nullptr;
    P->dec_ev_recorded = false;
  }
  return LAGVAE_OK;
}

int lagvae_text_wait_decoder_grads(lagvae_text_plan* P, void* stream) {
  LV_CHECK_ARG(P, "wait_decoder_grads: null plan");
  LV_CHECK_ARG(P->dec_ev && P->dec_ev_recorded, "wait_decoder_grads: hook not enabled or no backward recorded yet");
  LV_CUDA(cudaStreamWaitEvent((cudaStream_t)stream, P->dec_ev, 0));
  return LAGVAE_OK;
}

int lagvae_text_encode_stats(lagvae_text_plan* P, const lagvae_text_params* w, const int64_t* x,
                             float* out_mu, float* out_logvar, void* stream) {
  LV_CHECK_ARG(P && w && x && out_mu && out_logvar, "encode_stats: null argument");
  cudaStream_t st = (cudaStream_t)stream;
  reset_pass(P);
  LV_TRY(encoder_forward(P, w, x, st));
  const float* h_last = P->h_e + (int64_t)(P->Te - 1) * P->d.B * P->d.nh;
  LV_TRY(head_reparam_kl(h_last, w->p[E_LIN], nullptr, P->d.B, P->d.nh, P->d.nz, 1, out_mu, out_logvar,
                         nullptr, nullptr, st));
  return LAGVAE_OK;
}

int lagvae_text_loss_forward(lagvae_text_plan* P, const lagvae_text_params* w, const int64_t* x,
                             const float* eps, float kl_weight, const lagvae_dropout* drop,
                             float* out_loss, float* out_rec, float* out_kl, float* out_mu,
                             float* out_logvar, float* out_z, void* stream) {
  LV_CHECK_ARG(P && w && x && eps && out_loss && out_rec && out_kl, "loss_forward: null argument");
  cudaStream_t st = (cudaStream_t)stream;
  const lagvae_text_dims& d = P->d;
  const int nh = d.nh, ni = d.ni, nz = d.nz, V = d.V, B = d.B, ns = d.ns, Bd = P->Bd, Td = P->Td;
  lagvae_dropout dr{};
  if (drop) dr = *drop;
  LV_CHECK_ARG(dr.mode >= 0 && dr.mode <= 2, "loss_forward: bad dropout mode");
  LV_CHECK_ARG(dr.mode != 1 || ((dr.p_in <= 0.f || dr.mask_in) && (dr.p_out <= 0.f || dr.mask_out)),
               "loss_forward: dropout mode 1 needs masks");
  LV_CHECK_ARG(dr.p_in < 1.f && dr.p_out < 1.f, "loss_forward: dropout p must be < 1");
  P->drop = dr;
  P->kl_weight = kl_weight;
  reset_pass(P);
  int status = LAGVAE_OK;

  // ---- encoder: enc_lstm.py:47-64 ; reparameterise + KL: encoder.py:40-79
  // The decoder's embedding + x-projection do not depend on z: they run on a side stream UNDER the encoder recurrence, on
  // the 20 SMs its 128 persistent CTAs leave idle (GEMM grid capped to them; scripts/microbench/overlap_probe.py: the
  // forward recurrence and a 20-CTA GEMM overlap fully in both launch orders).  LAGVAE_OVERLAP_XPROJ=0 turns it off.
  static const bool ov_env = [] { const char* e = getenv("LAGVAE_OVERLAP_XPROJ"); return !(e && e[0] == '0'); }();
  const bool overlap = ov_env && P->use_tc && P->lstm_tc && nh >= 256 && B <= 128 && P->rd >= 128;
  int xstate = 0;
  if (overlap) {
    if (!P->side) {
      LV_CUDA(cudaStreamCreateWithFlags(&P->side, cudaStreamNonBlocking));
      LV_CUDA(cudaEventCreateWithFlags(&P->side_fork, cudaEventDisableTiming));
      LV_CUDA(cudaEventCreateWithFlags(&P->side_join, cudaEventDisableTiming));
    }
    LV_TRY(encoder_forward(P, w, x, st, P->side_fork));
    LV_CUDA(cudaStreamWaitEvent(P->side, P->side_fork, 0));
    bool ran = false;
    const int r = decoder_xproj(P, w, x, dr, P->side, d.T, 20, &ran);   // 148 SMs - 128 recurrence CTAs
    LV_CUDA(cudaEventRecord(P->side_join, P->side));                     // joined even on error: no dangling fork
    LV_CUDA(cudaStreamWaitEvent(st, P->side_join, 0));
    LV_TRY(r);
    xstate = ran ? 1 : 2;
  } else {
    LV_TRY(encoder_forward(P, w, x, st));
  }
  LV_CUDA(cudaMemcpyAsync(P->eps, eps, sizeof(float) * Bd * nz, cudaMemcpyDeviceToDevice, st));
  const float* h_last = P->h_e + (int64_t)(P->Te - 1) * B * nh;
  LV_TRY(head_reparam_kl(h_last, w->p[E_LIN], P->eps, B, nh, nz, ns, P->mu, P->logvar, P->z, P->kl, st));

  LV_TRY(decoder_forward(P, w, x, P->z, dr, st, -1, true, xstate));
  LV_TRY(finalize_loss(P->loss_row, P->kl, B, ns, Td, kl_weight, out_loss, out_rec, out_kl, P->scalars, st));
  if (out_mu) LV_CUDA(cudaMemcpyAsync(out_mu, P->mu, sizeof(float) * B * nz, cudaMemcpyDeviceToDevice, st));
  if (out_logvar)
    LV_CUDA(cudaMemcpyAsync(out_logvar, P->logvar, sizeof(float) * B * nz, cudaMemcpyDeviceToDevice, st));
  if (out_z) LV_CUDA(cudaMemcpyAsync(out_z, P->z, sizeof(float) * Bd * nz, cudaMemcpyDeviceToDevice, st));
  P->have_forward = true;
  return LAGVAE_OK;
}

int lagvae_text_reconstruct_error(lagvae_text_plan* P, const lagvae_text_params* w, const int64_t* x,
                                  const float* z, const lagvae_dropout* drop, float* out_rec_rows,
                                  void* stream) {
  LV_CHECK_ARG(P && w && x && z && out_rec_rows, "reconstruct_error: null argument");
  cudaStream_t st = (cudaStream_t)stream;
  lagvae_dropout dr{};
  if (drop) dr = *drop;
  LV_CHECK_ARG(dr.mode >= 0 && dr.mode <= 2 && dr.p_in < 1.f && dr.p_out < 1.f, "reconstruct_error: bad dropout");
  reset_pass(P);
  LV_TRY(decoder_forward(P, w, x, z, dr, st));
  LV_TRY(time_sum(P->loss_row, P->Td, P->Bd, 1, out_rec_rows, st));   // dec_lstm.py:148
  return LAGVAE_OK;
}

int lagvae_text_decode_logits(lagvae_text_plan* P, const lagvae_text_params* w, const int64_t* input,
                              const float* z, const lagvae_dropout* drop, float* out_logits, void* stream) {
  LV_CHECK_ARG(P && w && input && z && out_logits, "decode_logits: null argument");
  cudaStream_t st = (cudaStream_t)stream;
  lagvae_dropout dr{};
  if (drop) dr = *drop;
  LV_CHECK_ARG(dr.mode >= 0 && dr.mode <= 2 && dr.p_in < 1.f && dr.p_out < 1.f, "decode_logits: bad dropout");
  reset_pass(P);
  LV_TRY(decoder_forward(P, w, input, z, dr, st, P->Td, false));
  LV_TRY(logits_batch_major(P->logits, P->ldl, P->d.V, P->Td, P->Bd, out_logits, st));
  return LAGVAE_OK;
}

int lagvae_text_loss_backward(lagvae_text_plan* P, const lagvae_text_params* w, const int64_t* x,
                              const float* g_loss, const float* g_rec, const float* g_kl,
                              const lagvae_text_params* gr, uint32_t flags, void* stream) {
  LV_CHECK_ARG(P && w && x && gr, "loss_backward: null argument");
  LV_CHECK_ARG(P->have_forward, "loss_backward: no forward stash on this plan");
  P->dec_wgrad_passes = (flags & LAGVAE_BWD_DECODER_WGRAD_NORM_ONLY) ? 1 : 3;
  cudaStream_t st = (cudaStream_t)stream;
  const lagvae_text_dims& d = P->d;
  const int nh = d.nh, ni = d.ni, nz = d.nz, V = d.V, B = d.B, ns = d.ns, Bd = P->Bd, Td = P->Td, Te = P->Te;
  const int64_t re = P->re, rd = P->rd;
  const DropSpec din = spec_in(P->drop), dout = spec_out(P->drop);
  int status = LAGVAE_OK;
  P->arena_off = P->fwd_arena_end;   // forward's copies of xe | xd | h stay alive below this offset
  P->arena_floor = P->fwd_arena_end;
  P->have_forward = false;  // logits are consumed in place
  // Norm-only dW_pred (fused inner step: one bf16 pass, 0.39 ms at full width) on a side stream UNDER the two backward
  // recurrences, on the 20 SMs their 128 persistent CTAs leave idle.  The backward recurrence runs clusters of 4: with
  // 18-19 SMs per GPC exactly 4 clusters fit per GPC, so all 32 clusters fit only if the GEMM's CTAs sit on the LEFT-OVER
  // SMs — round 1 started the GEMM first and the step got slower (10.22 vs 9.49 ms).  Now the side stream waits on a device
  // flag that the recurrence sets once its whole grid is resident (k_wait_flag), and it is joined at the END of the backward
  // pass (the encoder recurrence re-uses the same cluster placement).  Not used when the data-parallel hook wants the
  // decoder gradients early.  LAGVAE_SIDE_WGRAD=0 turns it off; LAGVAE_SIDE_WGRAD_FRAC=f sends only the first f·V rows.
  static const bool side_env = [] { const char* e = getenv("LAGVAE_SIDE_WGRAD"); return !(e && e[0] == '0'); }();
  // The full-precision (3-pass) dW_pred of the module path is 2.4x the work: only its first rows fit under the recurrences
  // (LAGVAE_SIDE_WGRAD_FRAC3, default 0.35), the rest runs at full width as before.
  auto frac_env = [](const char* name, float dflt) { const char* e = getenv(name); const float f = e ? (float)atof(e) : dflt;
                                                     return f < 0.f ? 0.f : (f > 1.f ? 1.f : f); };
  static const float side_frac1 = frac_env("LAGVAE_SIDE_WGRAD_FRAC", 1.f);
  static const float side_frac3 = frac_env("LAGVAE_SIDE_WGRAD_FRAC3", 0.35f);
  // calibrated at B*ns = 32, V = 20001 (the recurrences' duration does not grow with either, the GEMM's work does)
  const float side_frac = (P->dec_wgrad_passes == 1 ? side_frac1 : side_frac3) * std::min(1.f, 32.f / (float)Bd) *
                          std::min(1.f, 20480.f / (float)V);
  const bool side_wgrad = side_env && side_frac > 0.f && P->use_tc && P->lstm_tc && nh >= 256 && Bd <= 64 && !P->dec_ev &&
                          V >= 1024;
  bool side_pending = false;
  int side_rows = 0;
  Staged side_sdl{}, side_sh{};

  LV_TRY(combine_upstream(g_loss, g_rec, g_kl, P->kl_weight, B, P->g_rec, P->g_kl, st));

  // ---- CE + vocabulary projection backward (autograd of dec_lstm.py:109,143-148)
  const float* hd = dout.mode ? P->hdrop_d : P->h_d;
  {
    // tensor-core path: dlogits is produced directly as the split-bf16 operand (no fp32 round trip)
    const bool tc_dl = P->use_tc && rd >= 32 && nh >= 16 && V >= 32;
    Staged sdl;
    if (tc_dl) {
      sdl = stage_alloc(P, rd, V, &status);
      LV_TRY(status);
      // fused steps: the forward's cross-entropy pass already wrote it (same allocation, upstream known in advance)
      const bool have_dl = P->dl_ready && P->st_dl.tc.hi == sdl.tc.hi && P->st_dl.tc.lo == sdl.tc.lo;
      P->dl_ready = false;
      if (!have_dl)
        LV_TRY(ce_bwd_split(P->logits, P->ldl, V, x, d.T, Td, Bd, ns, P->lse, P->g_rec, const_cast<uint16_t*>(sdl.tc.hi),
                            const_cast<uint16_t*>(sdl.tc.lo), sdl.tc.ld, st));
    } else {
      LV_TRY(ce_bwd(P->logits, P->ldl, V, x, d.T, Td, Bd, ns, P->lse, P->g_rec, st));
      sdl = stage(P, Mat{P->logits, rd, V, P->ldl}, st, &status);
    }
    Staged swp = stage_dec_weight(P, 1, Mat{w->p[D_PRED], V, nh, nh}, st, &status);
    Staged sh = P->st_h.tc.hi ? P->st_h : stage(P, Mat{hd, rd, nh, nh}, st, &status);   // forward's copy of hdrop / h_d
    LV_TRY(status);
    // dH_drop [rd, nh] = dlogits · W_pred
    LV_TRY(mm(P, sdl, false, swp, true, P->dh_d, nh, (int)rd, nh, V, 1.f, 0.f, nullptr, nullptr, 0, 3, st));
    // dW_pred [V, nh] = dlogitsᵀ · H_drop
    if (side_wgrad && tc_dl && sdl.tc.hi && sh.tc.hi) {
      if (!P->side) {
        LV_CUDA(cudaStreamCreateWithFlags(&P->side, cudaStreamNonBlocking));
        LV_CUDA(cudaEventCreateWithFlags(&P->side_fork, cudaEventDisableTiming));
        LV_CUDA(cudaEventCreateWithFlags(&P->side_join, cudaEventDisableTiming));
      }
      const int Vs = side_frac >= 1.f ? V : (int)((int64_t)(side_frac * (float)V) / 128 * 128);   // rows [0, Vs) on the side stream
      if (Vs < V)    // the rest at full width, here
        LV_TRY(mm(P, sub(sdl, 0, rd, Vs, V - Vs), true, sh, true, gr->p[D_PRED] + (int64_t)Vs * nh, nh, V - Vs, nh, (int)rd,
                  1.f, 0.f, nullptr, nullptr, 0, P->dec_wgrad_passes, st));
      LV_CUDA(cudaMemsetAsync(P->side_gate, 0, sizeof(unsigned), st));
      LV_CUDA(cudaEventRecord(P->side_fork, st));          // operands staged, gate cleared; the recurrence is launched next
      side_rows = Vs;
      side_sdl = sdl;
      side_sh = sh;
      side_pending = Vs > 0;
      P->arena_floor = P->arena_off;                       // dlogits staging stays live until the join
    } else {
      LV_TRY(mm(P, sdl, true, sh, true, gr->p[D_PRED], nh, V, nh, (int)rd, 1.f, 0.f, nullptr, nullptr, 0,
                P->dec_wgrad_passes, st));
    }
  }
  P->arena_off = P->arena_floor;  // dlogits staging no longer needed (unless the side stream still reads it)

  // ---- decoder LSTM backward (cuDNN RNN backward in the reference)
  // the cluster recurrence kernel also emits Σ_t dG (-> dzb) and dG as the bf16 hi/lo operand of the weight-gradient GEMMs
  static const bool no_extras = [] { const char* e = getenv("LAGVAE_LSTM_NO_EXTRAS"); return e && e[0] == '1'; }();
  Staged sdg_k = P->lstm_tc && P->use_tc && !no_extras ? stage_alloc(P, rd, 4 * nh, &status) : Staged{};
  LV_TRY(status);
  bool extras = false;
  if (P->lstm_tc)
    LV_TRY(lstm_tc_backward(P->lstm_tc, w->p[D_WHH], P->c0, P->gates_d, P->c_d, P->dh_d, dout, nullptr, P->dc, P->dh_rec,
                            P->dgates_d, Td, Bd, true, st, P->dzb, const_cast<uint16_t*>(sdg_k.tc.hi),
                            const_cast<uint16_t*>(sdg_k.tc.lo), &extras, side_pending ? P->side_gate : nullptr));
  else
    LV_TRY(lstm_backward_steps(w->p[D_WHH], P->c0, P->gates_d, P->c_d, P->dh_d, dout, nullptr, P->dc,
                               P->dh_rec, P->dgates_d, Td, Bd, nh, true, st));
  if (side_pending) {
    // enqueued AFTER the recurrence launch: fork -> wait until the recurrence's grid is resident -> GEMM on the idle SMs
    LV_CUDA(cudaStreamWaitEvent(P->side, P->side_fork, 0));
    // the gate only when the cluster kernel (the one that sets the flag) was the one launched; k_wait_flag gives up after ~1 ms
    const bool gated = P->lstm_tc && strncmp(lstm_last_variant(1), "v2", 2) == 0;
    int rs = gated ? wait_flag(P->side_gate, 2000000LL, P->side) : LAGVAE_OK;
    if (rs == LAGVAE_OK) {
      gemm_tc_set_grid_cap(20);                                     // 148 SMs - 128 recurrence CTAs
      rs = mm(P, sub(side_sdl, 0, rd, 0, side_rows), true, side_sh, true, gr->p[D_PRED], nh, side_rows, nh, (int)rd, 1.f, 0.f,
              nullptr, nullptr, 0, P->dec_wgrad_passes, P->side);
      gemm_tc_set_grid_cap(0);
    }
    LV_CUDA(cudaEventRecord(P->side_join, P->side));
    if (rs != LAGVAE_OK) {
      cudaStreamWaitEvent(st, P->side_join, 0);
      return rs;
    }
  }
  if (!extras) LV_TRY(time_sum(P->dgates_d, Td, Bd, 4 * nh, P->dzb, st));
  LV_TRY(col_sum(P->dzb, Bd, 4 * nh, gr->p[D_BIH], gr->p[D_BHH], st));
  {
    Staged sdg = sdg_k;
    sdg.m = Mat{P->dgates_d, rd, 4 * nh, 4 * nh};
    if (!extras) sdg = stage(P, Mat{P->dgates_d, rd, 4 * nh, 4 * nh}, st, &status);
    Staged sxd = P->st_xd.tc.hi ? P->st_xd : stage(P, Mat{P->xd, rd, ni, ni}, st, &status);
    // h_d itself (not its dropped-out version) pairs with dG in dW_hh: forward's copy serves only when dropout_out is off.
    // When it is staged here, h0 is staged in FRONT of it: [h0 ; h_0 .. h_{T-1}] makes dW_hh ONE GEMM over all T steps
    // (rows t pair with row t of that operand) instead of a tensor GEMM over t >= 1 plus an fp32 GEMM for the h0 term.
    Staged shd, shd_prev{};
    if (P->st_h.tc.hi && !dout.mode) {
      shd = P->st_h;
    } else if (P->use_tc) {
      Staged ext = stage_alloc(P, rd + Bd, nh, &status);
      LV_TRY(status);
      LV_TRY(split_bf16_launch(P->h0, nh, Bd, nh, const_cast<uint16_t*>(ext.tc.hi), const_cast<uint16_t*>(ext.tc.lo), ext.tc.ld, st));
      LV_TRY(split_bf16_launch(P->h_d, nh, (int)rd, nh, const_cast<uint16_t*>(ext.tc.hi) + (int64_t)Bd * ext.tc.ld,
                               const_cast<uint16_t*>(ext.tc.lo) + (int64_t)Bd * ext.tc.ld, ext.tc.ld, st));
      shd_prev = ext;
      shd_prev.m = Mat{nullptr, rd, nh, nh};
      shd = sub(ext, Bd, rd, 0, nh);
      shd.m = Mat{P->h_d, rd, nh, nh};
    } else {
      shd = stage(P, Mat{P->h_d, rd, nh, nh}, st, &status);
    }
    Staged swx = stage_dec_weight(P, 0, Mat{w->p[D_WIH], 4 * nh, ni, ni + nz}, st, &status);
    LV_TRY(status);
    // dW_ih[:, :ni] = dGᵀ · X   ;  dW_ih[:, ni:] = dzbᵀ · z
    LV_TRY(mm(P, sdg, true, sxd, true, gr->p[D_WIH], ni + nz, 4 * nh, ni, (int)rd, 1.f, 0.f, nullptr, nullptr, 0,
              P->dec_wgrad_passes, st));
    LV_TRY(gemm_f32(P->dzb, 1, 4 * nh, P->z, 1, nz, gr->p[D_WIH] + ni, ni + nz, 4 * nh, nz, Bd, 1.f, 0.f, nullptr,
                    nullptr, 0, st));
    // dW_hh = Σ_t dG_tᵀ h_{t-1}: rows t>=1 pair with h_d[t-1]; t=0 pairs with h0
    if (shd_prev.tc.hi && sdg.tc.hi && rd >= 32 && nh >= 16) {
      LV_TRY(mm(P, sdg, true, sub(shd_prev, 0, rd, 0, nh), true, gr->p[D_WHH], nh, 4 * nh, nh, (int)rd, 1.f, 0.f, nullptr,
                nullptr, 0, P->dec_wgrad_passes, st));
    } else {
      if (Td > 1)
        LV_TRY(mm(P, sub(sdg, Bd, rd - Bd, 0, 4 * nh), true, sub(shd, 0, rd - Bd, 0, nh), true, gr->p[D_WHH], nh,
                  4 * nh, nh, (int)(rd - Bd), 1.f, 0.f, nullptr, nullptr, 0, P->dec_wgrad_passes, st));
      LV_TRY(gemm_f32(P->dgates_d, 1, 4 * nh, P->h0, 1, nh, gr->p[D_WHH], nh, 4 * nh, nh, Bd, 1.f,
                      Td > 1 ? 1.f : 0.f, nullptr, nullptr, 0, st));
    }
    // dX = dG · W_ih[:, :ni]  -> dense decoder embedding gradient (row V-1 = padding_idx, no grad).  It feeds ONLY the
    // decoder embedding gradient — a decoder weight gradient like the three above: one bf16 pass when they are norm-only
    LV_TRY(mm(P, sdg, false, swx, true, P->dx_d, ni, (int)rd, ni, 4 * nh, 1.f, 0.f, nullptr, nullptr, 0,
              P->dec_wgrad_passes, st));
  }
  LV_TRY(fill(gr->p[D_EMB], 0.f, (int64_t)V * ni, st));
  LV_TRY(embed_scatter_add(x, d.T, 0, B, ns, Td, P->dx_d, ni, din, gr->p[D_EMB], V - 1, st));
  // initial state: c0 = z W_transᵀ, h0 = tanh(c0)  (dec_lstm.py:100-101)
  LV_TRY(dc0_total(P->dc, P->dh_rec, P->h0, P->dc0t, Bd * nh, st));
  LV_TRY(gemm_f32(P->dc0t, 1, nh, P->z, 1, nz, gr->p[D_TRANS], nz, nh, nz, Bd, 1.f, 0.f, nullptr, nullptr, 0, st));
  // dz = dzb · W_ih[:, ni:] + dc0_tot · W_trans
  LV_TRY(gemm_f32(P->dzb, 4 * nh, 1, w->p[D_WIH] + ni, 1, ni + nz, P->dz, nz, Bd, nz, 4 * nh, 1.f, 0.f, nullptr,
                  nullptr, 0, st));
  LV_TRY(gemm_f32(P->dc0t, nh, 1, w->p[D_TRANS], 1, nz, P->dz, nz, Bd, nz, nh, 1.f, 1.f, nullptr, nullptr, 0, st));

  // all 7 decoder gradients are final here: data-parallel callers may start reducing them (lagvae.h)
  if (P->dec_ev) {
    LV_CUDA(cudaEventRecord(P->dec_ev, st));
    P->dec_ev_recorded = true;
  }

  // ---- reparameterisation + KL + head backward
  LV_TRY(reparam_kl_bwd(P->dz, P->eps, P->mu, P->logvar, P->g_kl, B, nz, ns, P->dml, st));
  const float* h_last = P->h_e + (int64_t)(Te - 1) * B * nh;
  LV_TRY(gemm_f32(P->dml, 2 * nz, 1, w->p[E_LIN], 1, nh, P->dh_last, nh, B, nh, 2 * nz, 1.f, 0.f, nullptr,
                  nullptr, 0, st));
  LV_TRY(gemm_f32(P->dml, 1, 2 * nz, h_last, 1, nh, gr->p[E_LIN], nh, 2 * nz, nh, B, 1.f, 0.f, nullptr, nullptr,
                  0, st));

  // ---- encoder LSTM backward
  P->arena_off = P->arena_floor;       // == fwd_arena_end unless the side stream still reads the dlogits staging
  Staged sdg_ke = P->lstm_tc && P->use_tc && !no_extras ? stage_alloc(P, re, 4 * nh, &status) : Staged{};
  LV_TRY(status);
  bool extras_e = false;
  if (P->lstm_tc)
    LV_TRY(lstm_tc_backward(P->lstm_tc, w->p[E_WHH], nullptr, P->gates_e, P->c_e, nullptr, spec_none(), P->dh_last,
                            P->dc_e, P->dh_rec_e, P->dgates_e, Te, B, false, st, P->dzb, const_cast<uint16_t*>(sdg_ke.tc.hi),
                            const_cast<uint16_t*>(sdg_ke.tc.lo), &extras_e));
  else
    LV_TRY(lstm_backward_steps(w->p[E_WHH], nullptr, P->gates_e, P->c_e, nullptr, spec_none(), P->dh_last,
                               P->dc_e, P->dh_rec_e, P->dgates_e, Te, B, nh, false, st));
  if (!extras_e) LV_TRY(time_sum(P->dgates_e, Te, B, 4 * nh, P->dzb, st));   // Σ_t first (131 K threads), then Σ_b over B rows
  LV_TRY(col_sum(P->dzb, B, 4 * nh, gr->p[E_BIH], gr->p[E_BHH], st));
  {
    Staged sdg = sdg_ke;
    sdg.m = Mat{P->dgates_e, re, 4 * nh, 4 * nh};
    if (!extras_e) sdg = stage(P, Mat{P->dgates_e, re, 4 * nh, 4 * nh}, st, &status);
    Staged sxe = P->st_xe.tc.hi ? P->st_xe : stage(P, Mat{P->xe, re, ni, ni}, st, &status);
    Staged she = stage(P, Mat{P->h_e, re, nh, nh}, st, &status);
    Staged swx = stage(P, Mat{w->p[E_WIH], 4 * nh, ni, ni}, st, &status);
    LV_TRY(status);
    LV_TRY(mm(P, sdg, true, sxe, true, gr->p[E_WIH], ni, 4 * nh, ni, (int)re, 1.f, 0.f, nullptr, nullptr, 0, 3, st));
    if (Te > 1)
      LV_TRY(mm(P, sub(sdg, B, re - B, 0, 4 * nh), true, sub(she, 0, re - B, 0, nh), true, gr->p[E_WHH], nh,
                4 * nh, nh, (int)(re - B), 1.f, 0.f, nullptr, nullptr, 0, 3, st));
    else
      LV_TRY(fill(gr->p[E_WHH], 0.f, (int64_t)4 * nh * nh, st));
    LV_TRY(mm(P, sdg, false, swx, true, P->dx_e, ni, (int)re, ni, 4 * nh, 1.f, 0.f, nullptr, nullptr, 0, 3, st));
  }
  LV_TRY(fill(gr->p[E_EMB], 0.f, (int64_t)V * ni, st));
  LV_TRY(embed_scatter_add(x, d.T, 0, B, 1, Te, P->dx_e, ni, spec_none(), gr->p[E_EMB], -1, st));
  if (side_pending) {   // dW_pred ran under the two recurrences: join before the gradients are handed back
    LV_CUDA(cudaStreamWaitEvent(st, P->side_join, 0));
    P->arena_floor = P->fwd_arena_end;
  }
  return LAGVAE_OK;
}

int lagvae_clip_sgd_step(float* const* h_params, float* const* h_grads, const int64_t* h_counts, int n_seg,
                         int n_update, float max_norm, float lr, int scale_all_grads, float* out_norm,
                         void* scratch, void* stream) {
  LV_CHECK_ARG(h_grads && h_counts && scratch && (n_update == 0 || h_params), "clip_sgd_step: null argument");
  return clip_sgd_step(h_params, h_grads, h_counts, n_seg, n_update, max_norm, lr, scale_all_grads, out_norm,
                       scratch, (cudaStream_t)stream);
}

int lagvae_mi_estimate(const float* mu, const float* logvar, const float* eps, int B, int nz, float* out_mi,
                       void* stream) {
  LV_CHECK_ARG(mu && logvar && eps && out_mi && B > 0 && nz > 0, "mi_estimate: bad argument");
  return mi_estimate(mu, logvar, eps, B, nz, out_mi, (cudaStream_t)stream);
}

static int text_step(lagvae_text_plan* P, const lagvae_text_params* w, const int64_t* x, const float* eps, float kl_weight,
                     const lagvae_dropout* drop, float max_norm, float lr, bool upd_enc, bool upd_dec, float* grad_ws,
                     float* out_loss, float* out_scalars, void* stream) {
  cudaStream_t st = (cudaStream_t)stream;
  const lagvae_text_dims& d = P->d;
  const int64_t V = d.V, ni = d.ni, nh = d.nh, nz = d.nz;
  const int64_t counts[LAGVAE_TEXT_NPARAM] = {V * ni, 4 * nh * ni, 4 * nh * nh, 4 * nh, 4 * nh, 2 * nz * nh,
                                              V * ni, nh * nz, 4 * nh * (ni + nz), 4 * nh * nh, 4 * nh, 4 * nh,
                                              V * nh};
  lagvae_text_params g;
  int64_t off = 0;
  for (int i = 0; i < LAGVAE_TEXT_NPARAM; ++i) {
    g.p[i] = grad_ws + off;
    off += counts[i];
  }
  // text.py:379 / :411 loss; :381 Σloss; :382 / :413 mean(dim=-1) -> upstream 1/B
  P->ce_upstream = 1.f / (float)d.B;
  const int rf = lagvae_text_loss_forward(P, w, x, eps, kl_weight, drop, out_loss, P->dml /*rec tmp*/, P->dh_last /*kl tmp*/,
                                          nullptr, nullptr, nullptr, stream);
  P->ce_upstream = 0.f;
  LV_TRY(rf);
  LV_CUDA(cudaMemcpyAsync(out_scalars, P->scalars, 3 * sizeof(float), cudaMemcpyDeviceToDevice, st));
  LV_TRY(fill(P->dc0t, 1.f / (float)d.B, d.B, st));
  // Decoder WEIGHT gradients that are not applied (aggressive inner loop: text.py:387 steps the encoder only) enter the
  // update through the clip norm alone (text.py:385), for which one bf16 pass is ample (norm error ~1e-5); when the
  // decoder is stepped (text.py:424) they are computed at full 3-pass precision.
  LV_TRY(lagvae_text_loss_backward(P, w, x, P->dc0t, nullptr, nullptr, &g,
                                   upd_dec ? LAGVAE_BWD_DEFAULT : LAGVAE_BWD_DECODER_WGRAD_NORM_ONLY, stream));
  // text.py:385 / :414 clip over all 13 grads; then SGD on the selected halves (updated tensors first in the segment list)
  float* pp[LAGVAE_TEXT_NPARAM];
  float* gg[LAGVAE_TEXT_NPARAM];
  int64_t cc[LAGVAE_TEXT_NPARAM];
  int n = 0, n_upd = 0;
  auto push = [&](int lo, int hi) {
    for (int i = lo; i < hi; ++i, ++n) {
      pp[n] = const_cast<float*>(w->p[i]);
      gg[n] = g.p[i];
      cc[n] = counts[i];
    }
  };
  if (upd_enc) { push(0, 6); n_upd = n; }
  if (upd_dec) { push(6, LAGVAE_TEXT_NPARAM); n_upd = n; }
  if (!upd_enc) push(0, 6);
  if (!upd_dec) push(6, LAGVAE_TEXT_NPARAM);
  LV_TRY(clip_sgd_step(pp, gg, cc, LAGVAE_TEXT_NPARAM, n_upd, max_norm, lr, 0, out_scalars + 3, P->clip_scratch, st));
  return LAGVAE_OK;
}

int lagvae_text_inner_step(lagvae_text_plan* P, const lagvae_text_params* w, const int64_t* x,
                           const float* eps, float kl_weight, const lagvae_dropout* drop, float max_norm,
                           float lr, float* grad_ws, float* out_loss, float* out_scalars, void* stream) {
  LV_CHECK_ARG(P && w && x && eps && grad_ws && out_loss && out_scalars, "inner_step: null argument");
  return text_step(P, w, x, eps, kl_weight, drop, max_norm, lr, true, false, grad_ws, out_loss, out_scalars, stream);
}

int lagvae_text_outer_step(lagvae_text_plan* P, const lagvae_text_params* w, const int64_t* x,
                           const float* eps, float kl_weight, const lagvae_dropout* drop, float max_norm,
                           float lr, int update_encoder, float* grad_ws, float* out_loss, float* out_scalars,
                           void* stream) {
  LV_CHECK_ARG(P && w && x && eps && grad_ws && out_loss && out_scalars, "outer_step: null argument");
  const int r = text_step(P, w, x, eps, kl_weight, drop, max_norm, lr, update_encoder != 0, true, grad_ws, out_loss,
                          out_scalars, stream);
  if (P->wc) P->wc->epoch = 0;   // the decoder weights were just stepped in place: whatever epoch the caller declares next, re-split
  return r;
}

size_t lagvae_lstm_workspace_bytes(int nh, int Bd) {
  lagvae_text_dims d{Bd, 2, 1, 2, 1, nh, 1};
  return lstm_tc_workspace_bytes(d, true) + 256;
}

static int lstm_state_for(int tier, int nh, int Bd, void* ws, size_t ws_bytes, LstmTcState** st_out) {
  *st_out = nullptr;
  if (tier == 0) return LAGVAE_OK;
  LV_CHECK_ARG(ws && ((uintptr_t)ws & 255) == 0, "lstm: tier 1 needs a 256-B aligned workspace");
  lagvae_text_dims d{Bd, 2, 1, 2, 1, nh, 1};
  LV_TRY(lstm_tc_create(d, true, ws, ws_bytes, st_out));
  if (!*st_out) {
    set_error("lstm: shape nh=%d Bd=%d is not covered by the persistent tcgen05 kernel", nh, Bd);
    return LAGVAE_E_ARG;
  }
  return LAGVAE_OK;
}

int lagvae_lstm_forward(int tier, int nh, int Tn, int Bd, const float* w_hh, const float* h0, const float* c0,
                        float* gates, float* c_all, float* h_all, float* hdrop_all, const lagvae_dropout* drop,
                        void* ws, size_t ws_bytes, void* stream) {
  LV_CHECK_ARG(w_hh && gates && c_all && h_all && nh > 0 && Tn > 0 && Bd > 0, "lstm_forward: bad argument");
  cudaStream_t st = (cudaStream_t)stream;
  lagvae_dropout dr{};
  if (drop) dr = *drop;
  const DropSpec ds = hdrop_all ? spec_out(dr) : spec_none();
  LstmTcState* s = nullptr;
  LV_TRY(lstm_state_for(tier, nh, Bd, ws, ws_bytes, &s));
  int r;
  if (s) {
    r = lstm_tc_forward(s, w_hh, h0, c0, gates, c_all, h_all, hdrop_all, ds, Tn, Bd, st);
    lstm_tc_destroy(s);
  } else {
    r = lstm_forward_steps(w_hh, h0, c0, gates, c_all, h_all, hdrop_all, ds, Tn, Bd, nh, st);
  }
  return r;
}

int lagvae_lstm_backward(int tier, int nh, int Tn, int Bd, const float* w_hh, const float* c0,
                         const float* gates, const float* c_all, const float* dh_ext, const float* dh_last,
                         const lagvae_dropout* drop, float* dc, float* dh_rec, float* dgates, int want_init,
                         void* ws, size_t ws_bytes, void* stream) {
  LV_CHECK_ARG(w_hh && gates && c_all && dc && dh_rec && dgates && nh > 0 && Tn > 0 && Bd > 0,
               "lstm_backward: bad argument");
  cudaStream_t st = (cudaStream_t)stream;
  lagvae_dropout dr{};
  if (drop) dr = *drop;
  const DropSpec ds = dh_ext ? spec_out(dr) : spec_none();
  LstmTcState* s = nullptr;
  LV_TRY(lstm_state_for(tier, nh, Bd, ws, ws_bytes, &s));
  int r;
  if (s) {
    r = lstm_tc_backward(s, w_hh, c0, gates, c_all, dh_ext, ds, dh_last, dc, dh_rec, dgates, Tn, Bd, want_init != 0, st);
    lstm_tc_destroy(s);
  } else {
    r = lstm_backward_steps(w_hh, c0, gates, c_all, dh_ext, ds, dh_last, dc, dh_rec, dgates, Tn, Bd, nh,
                            want_init != 0, st);
  }
  return r;
}

void lagvae_debug_trace_buffer(void* dev_u64, size_t words) { lstm_tc_set_debug(dev_u64, words); }

const char* lagvae_lstm_variant(int direction) { return lstm_last_variant(direction); }

int lagvae_gemm_f32(const float* A, int64_t a_rs, int64_t a_cs, const float* B, int64_t b_rs, int64_t b_cs,
                    float* C, int64_t ldc, int M, int N, int K, float alpha, float beta, const float* bias_n,
                    const float* bias_rows, int bias_period, void* stream) {
  LV_CHECK_ARG(A && B && C, "gemm_f32: null argument");
  return gemm_f32(A, a_rs, a_cs, B, b_rs, b_cs, C, ldc, M, N, K, alpha, beta, bias_n, bias_rows, bias_period,
                  (cudaStream_t)stream);
}

int lagvae_gemm_tc(const uint16_t* A_hi, const uint16_t* A_lo, int64_t lda, int a_mn_major,
                   const uint16_t* B_hi, const uint16_t* B_lo, int64_t ldb, int b_mn_major, float* C,
                   int64_t ldc, int M, int N, int K, int passes, float alpha, float beta, const float* bias_n,
                   const float* bias_rows, int bias_period, const int32_t* out_row_map, void* stream) {
  LV_CHECK_ARG(A_hi && B_hi && C, "gemm_tc: null argument");
  TcOperand a{A_hi, A_lo, lda, a_mn_major}, b{B_hi, B_lo, ldb, b_mn_major};
  return gemm_tc(a, b, C, ldc, M, N, K, passes, alpha, beta, bias_n, bias_rows, bias_period, out_row_map,
                 (cudaStream_t)stream);
}

int lagvae_split_bf16(const float* src, int64_t ld, int rows, int cols, uint16_t* hi, uint16_t* lo,
                      int64_t ld_out, void* stream) {
  LV_CHECK_ARG(src && hi && lo && ld_out >= cols, "split_bf16: bad argument");
  return split_bf16_launch(src, ld, rows, cols, hi, lo, ld_out, (cudaStream_t)stream);
}

int lagvae_dropout_mask(uint64_t seed, uint32_t stream_id, int64_t n, float p, uint8_t* out_keep, void* stream) {
  LV_CHECK_ARG(out_keep && n >= 0, "dropout_mask: bad argument");
  return dropout_mask(seed, stream_id, n, p, out_keep, (cudaStream_t)stream);
}

}  // extern "C"
