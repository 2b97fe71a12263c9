// Persistent tcgen05 LSTM recurrence — placeholder translation unit (kernel lands in a later commit).
#include "lstm_tc.cuh"

namespace lagvae {

size_t lstm_tc_workspace_bytes(const lagvae_text_dims&, bool) { return 0; }
int lstm_tc_create(const lagvae_text_dims&, bool, void*, size_t, LstmTcState** out) {
  *out = nullptr;
  return LAGVAE_OK;
}
void lstm_tc_destroy(LstmTcState*) {}
int lstm_tc_pack_weights(LstmTcState*, int, const float*, int, int64_t, const float*, cudaStream_t) {
  set_error("lstm_tc: not available");
  return LAGVAE_E_ARG;
}
int lstm_tc_forward(LstmTcState*, int, const float*, const float*, float*, float*, float*, float*, DropSpec, int,
                    int, cudaStream_t) {
  set_error("lstm_tc: not available");
  return LAGVAE_E_ARG;
}
int lstm_tc_backward(LstmTcState*, int, const float*, const float*, const float*, const float*, DropSpec,
                     const float*, float*, float*, float*, int, int, bool, cudaStream_t) {
  set_error("lstm_tc: not available");
  return LAGVAE_E_ARG;
}

}  // namespace lagvae
