// PixelCNNBlock as ONE C-ABI call per direction (dec_pixelcnn_v2.py:32-62):
//   out = ELU( BN3(conv1x1_{Cm->C}( ELU(BN2(maskedconv_kxk_{Cm->Cm}( ELU(BN1(conv1x1_{C->Cm}(x))) ))) )) + x )
// forward  = 3 x [weight tiles, im2col-free tcgen05 convolution with BatchNorm statistics in the epilogue, fused
//                 BN/(residual)/ELU pass emitting the next operand in the bf16 cat format]           (10 launches)
// backward = 3 x [fused BN/ELU backward (reduce + apply), tcgen05 wgrad (+ deterministic reduce), tcgen05 dgrad]; the
//            residual-branch gradient rides in the epilogue of the last dgrad                         (15 launches)
// The host loop around it (one call per block instead of ~45 ctypes calls + ~35 allocations) is what keeps the
// driver-faithful (eager, unmodified image.py) inner step device-bound.
#include "kernels.cuh"

#include <algorithm>

namespace lagvae {
namespace {

struct Carver {   // bump allocator over a caller-owned buffer; with base == nullptr it only measures
  char* base;
  size_t off = 0;
  template <typename T>
  T* take(size_t n) {
    T* q = base ? (T*)(base + off) : nullptr;
    off += (size_t)round_up((int64_t)(n * sizeof(T)), 256);
    return q;
  }
};

struct Stash {
  float *y1, *y2, *y3;
  uint16_t *a1cat, *a2cat, *xcat;
  uint8_t *wb1, *wb2, *wb3;
  double* stats;   // [3][2C]
  float *sm1, *si1, *sm2, *si2, *sm3, *si3;
  size_t bytes;
};
Stash carve_stash(const lagvae_pixelblock_dims* d, void* base) {
  const size_t R = (size_t)d->B * d->H * d->W, C = d->C, Cm = d->Cm;
  Carver c{(char*)base};
  Stash s{};
  s.y1 = c.take<float>(R * Cm);
  s.y2 = c.take<float>(R * Cm);
  s.y3 = c.take<float>(R * C);
  s.a1cat = c.take<uint16_t>(R * 2 * Cm);
  s.a2cat = c.take<uint16_t>(R * 2 * Cm);
  s.xcat = c.take<uint16_t>(R * 2 * C);
  s.wb1 = c.take<uint8_t>(lagvae_convtc_wbuf_bytes(d->C, d->Cm, 1, 1));
  s.wb2 = c.take<uint8_t>(lagvae_convtc_wbuf_bytes(d->Cm, d->Cm, d->k, d->k));
  s.wb3 = c.take<uint8_t>(lagvae_convtc_wbuf_bytes(d->Cm, d->C, 1, 1));
  s.stats = c.take<double>(3 * 2 * C);
  s.sm1 = c.take<float>(Cm); s.si1 = c.take<float>(Cm);
  s.sm2 = c.take<float>(Cm); s.si2 = c.take<float>(Cm);
  s.sm3 = c.take<float>(C); s.si3 = c.take<float>(C);
  s.bytes = c.off;
  return s;
}

struct Scratch {
  uint16_t *dy3cat, *dy2cat, *dy1cat;
  float *dpre3, *da2, *da1;
  void *wg, *bn;
  size_t bytes;
};
size_t wgrad_scratch_max(const lagvae_pixelblock_dims* d) {
  const size_t a = lagvae_convtc_wgrad_scratch_bytes(d->C, d->Cm, 1, 1), b = lagvae_convtc_wgrad_scratch_bytes(d->Cm, d->Cm, d->k, d->k),
               c = lagvae_convtc_wgrad_scratch_bytes(d->Cm, d->C, 1, 1);
  return std::max(a, std::max(b, c));
}
Scratch carve_scratch(const lagvae_pixelblock_dims* d, void* base) {
  const size_t R = (size_t)d->B * d->H * d->W, C = d->C, Cm = d->Cm;
  Carver c{(char*)base};
  Scratch s{};
  s.dy3cat = c.take<uint16_t>(R * 2 * C);
  s.dy2cat = c.take<uint16_t>(R * 2 * Cm);
  s.dy1cat = c.take<uint16_t>(R * 2 * Cm);
  s.dpre3 = c.take<float>(R * C);
  s.da2 = c.take<float>(R * Cm);
  s.da1 = c.take<float>(R * Cm);
  s.wg = c.take<uint8_t>(wgrad_scratch_max(d));
  s.bn = c.take<uint8_t>(16 * C + 256);
  s.bytes = c.off;
  return s;
}
bool dims_ok(const lagvae_pixelblock_dims* d) {
  return d && (d->C == 32 || d->C == 64) && (d->Cm == 32 || d->Cm == 64) &&
         lagvae_convtc_supported(d->B, d->H, d->W, d->C, d->Cm, 1, 1) && lagvae_convtc_supported(d->B, d->H, d->W, d->Cm, d->Cm, d->k, d->k);
}

}  // namespace
}  // namespace lagvae

using namespace lagvae;

extern "C" {

size_t lagvae_pixelblock_stash_bytes(const lagvae_pixelblock_dims* d) {
  if (!dims_ok(d)) return 0;
  return carve_stash(d, nullptr).bytes + 256;
}
size_t lagvae_pixelblock_scratch_bytes(const lagvae_pixelblock_dims* d) {
  if (!dims_ok(d)) return 0;
  return carve_scratch(d, nullptr).bytes + 256;
}

int lagvae_pixelblock_forward(const lagvae_pixelblock_dims* d, const lagvae_pixelblock_params* p, const float* x,
                              const uint16_t* xcat_or_null, void* stash, float* out, uint16_t* outcat_or_null, void* stream) {
  LV_CHECK_ARG(dims_ok(d) && p && x && stash && out && ((uintptr_t)stash & 255) == 0, "pixelblock_forward: bad argument");
  const Stash s = carve_stash(d, stash);
  const int B = d->B, H = d->H, W = d->W, C = d->C, Cm = d->Cm, k = d->k;
  const int64_t R = (int64_t)B * H * W;
  const uint16_t* xcat = xcat_or_null;
  if (!xcat) {
    LV_TRY(lagvae_split_cat(x, R, C, s.xcat, stream));
    xcat = s.xcat;
  }
  LV_TRY(lagvae_convtc_prepare_weights(p->w1, Cm, C, 1, 1, 0, s.wb1, stream));
  LV_TRY(lagvae_convtc_prepare_weights(p->w2, Cm, Cm, k, k, 2, s.wb2, stream));
  LV_TRY(lagvae_convtc_prepare_weights(p->w3, C, Cm, 1, 1, 0, s.wb3, stream));
  double *st1 = s.stats, *st2 = s.stats + 2 * C, *st3 = s.stats + 4 * C;
  const bool ev = d->eval != 0;
  if (ev) {   // eval(): running statistics, nothing is updated (dec_pixelcnn_v2 under vae.eval())
    LV_TRY(lagvae_bn_eval_stats(p->rm1, p->rv1, R, Cm, st1, stream));
    LV_TRY(lagvae_bn_eval_stats(p->rm2, p->rv2, R, Cm, st2, stream));
    LV_TRY(lagvae_bn_eval_stats(p->rm3, p->rv3, R, C, st3, stream));
  }
  LV_TRY(lagvae_convtc_forward(xcat, s.wb1, B, H, W, C, Cm, 1, 1, 0, nullptr, s.y1, ev ? nullptr : st1, stream));
  LV_TRY(lagvae_bnact_fwd(s.y1, st1, R, Cm, p->g1, p->b1, d->eps, d->momentum, nullptr, 1, nullptr, s.a1cat, s.sm1, s.si1,
                          ev ? nullptr : p->rm1, ev ? nullptr : p->rv1, stream));
  LV_TRY(lagvae_convtc_forward(s.a1cat, s.wb2, B, H, W, Cm, Cm, k, k, 2, nullptr, s.y2, ev ? nullptr : st2, stream));
  LV_TRY(lagvae_bnact_fwd(s.y2, st2, R, Cm, p->g2, p->b2, d->eps, d->momentum, nullptr, 1, nullptr, s.a2cat, s.sm2, s.si2,
                          ev ? nullptr : p->rm2, ev ? nullptr : p->rv2, stream));
  LV_TRY(lagvae_convtc_forward(s.a2cat, s.wb3, B, H, W, Cm, C, 1, 1, 0, nullptr, s.y3, ev ? nullptr : st3, stream));
  LV_TRY(lagvae_bnact_fwd(s.y3, st3, R, C, p->g3, p->b3, d->eps, d->momentum, x, 1, out, outcat_or_null, s.sm3, s.si3,
                          ev ? nullptr : p->rm3, ev ? nullptr : p->rv3, stream));
  return LAGVAE_OK;
}

int lagvae_pixelblock_backward(const lagvae_pixelblock_dims* d, const lagvae_pixelblock_params* p, const float* dout, const float* out,
                               const void* stash, const uint16_t* xcat_or_null, float* dx, const lagvae_pixelblock_grads* g,
                               void* scratch, void* stream) {
  LV_CHECK_ARG(dims_ok(d) && p && dout && out && stash && dx && g && scratch && ((uintptr_t)stash & 255) == 0 &&
               ((uintptr_t)scratch & 255) == 0, "pixelblock_backward: bad argument");
  LV_CHECK_ARG(d->eval == 0, "pixelblock_backward: eval()-mode BatchNorm backward is not part of the path");
  const Stash s = carve_stash(d, const_cast<void*>(stash));
  const Scratch w = carve_scratch(d, scratch);
  const int B = d->B, H = d->H, W = d->W, C = d->C, Cm = d->Cm, k = d->k;
  const int64_t R = (int64_t)B * H * W;
  const uint16_t* xcat = xcat_or_null ? xcat_or_null : s.xcat;
  // BN3 + residual + ELU
  LV_TRY(lagvae_bnact_bwd(dout, out, nullptr, s.y3, R, C, p->g3, s.sm3, s.si3, 1, nullptr, w.dy3cat, w.dpre3, g->dg3, g->db3, w.bn, stream));
  LV_TRY(lagvae_convtc_wgrad(w.dy3cat, s.a2cat, B, H, W, Cm, C, 1, 1, g->dw3, w.wg, stream));
  LV_TRY(lagvae_convtc_dgrad(w.dy3cat, s.wb3, B, H, W, Cm, C, 1, 1, 0, nullptr, w.da2, stream));
  // BN2 + ELU, masked convolution
  LV_TRY(lagvae_bnact_bwd(w.da2, nullptr, s.a2cat, s.y2, R, Cm, p->g2, s.sm2, s.si2, 1, nullptr, w.dy2cat, nullptr, g->dg2, g->db2, w.bn, stream));
  LV_TRY(lagvae_convtc_wgrad(w.dy2cat, s.a1cat, B, H, W, Cm, Cm, k, k, g->dw2, w.wg, stream));
  LV_TRY(lagvae_convtc_dgrad(w.dy2cat, s.wb2, B, H, W, Cm, Cm, k, k, 2, nullptr, w.da1, stream));
  // BN1 + ELU, first 1x1; the residual-branch gradient dpre3 is added in the dgrad epilogue
  LV_TRY(lagvae_bnact_bwd(w.da1, nullptr, s.a1cat, s.y1, R, Cm, p->g1, s.sm1, s.si1, 1, nullptr, w.dy1cat, nullptr, g->dg1, g->db1, w.bn, stream));
  LV_TRY(lagvae_convtc_wgrad(w.dy1cat, xcat, B, H, W, C, Cm, 1, 1, g->dw1, w.wg, stream));
  LV_TRY(lagvae_convtc_dgrad(w.dy1cat, s.wb1, B, H, W, C, Cm, 1, 1, 0, w.dpre3, dx, stream));
  return LAGVAE_OK;
}

}  // extern "C"
