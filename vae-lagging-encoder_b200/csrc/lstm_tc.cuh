// Persistent tcgen05 LSTM recurrence (forward and backward) — interface used by text_plan.cu.
#pragma once
#include "lagvae_common.cuh"

namespace lagvae {

struct LstmTcState;  // opaque; nullptr = not available for this shape -> launch-per-step tier

size_t lstm_tc_workspace_bytes(const lagvae_text_dims& d, bool use_tc);
// returns LAGVAE_OK with *out == nullptr when the shape is not covered by the persistent kernels
int lstm_tc_create(const lagvae_text_dims& d, bool use_tc, void* ws, size_t ws_bytes, LstmTcState** out);
void lstm_tc_destroy(LstmTcState* s);
void lstm_note_variant(int dir, const char* what);      // dir 0 = forward, 1 = backward
const char* lstm_last_variant(int dir);
void lstm_tc_set_debug(void* dev_u64, size_t words);   // optional clock64 trace buffer of CTA 0 ([step][8])
// Same contracts as lstm_forward_steps / lstm_backward_steps in text_plan.cu.
int lstm_tc_forward(LstmTcState* s, const float* w_hh, const float* h0, const float* c0, float* gates,
                    float* c_all, float* h_all, float* hdrop_all, DropSpec drop, int Tn, int Bd,
                    cudaStream_t st,
                    // optional [Bd, 4nh]: added to the pre-activations of every time step (decoder z bias)
                    const float* row_bias = nullptr);
int lstm_tc_backward(LstmTcState* s, const float* w_hh, const float* c0, const float* gates,
                     const float* c_all, const float* dh_ext, DropSpec drop, const float* dh_last, float* dc,
                     float* dh_rec, float* dgates, int Tn, int Bd, bool want_init, cudaStream_t st,
                     // optional in-kernel extras (cluster kernel, Bd <= 64): dgsum [Bd,4nh] = sum over time of dG; dg_hi / dg_lo
                     // [Tn*Bd,4nh] = dG as the bf16 hi/lo GEMM operand; *extras_done tells whether they were produced
                     float* dgsum = nullptr, uint16_t* dg_hi = nullptr, uint16_t* dg_lo = nullptr, bool* extras_done = nullptr,
                     // optional device word set to 1 by the cluster kernel once its whole grid is running
                     unsigned* started = nullptr);

}  // namespace lagvae
