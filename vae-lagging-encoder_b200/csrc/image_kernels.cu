// Image-path kernels (ResNetEncoderV2, PixelCNNDecoderV2) — correctness-first tier.
// Activations are NHWC ("rows x channels": row = (b*H + y)*W + x), so every 1x1 convolution IS a GEMM on the
// activation matrix; k x k (masked / strided) convolutions gather their taps with im2col and run the same GEMM
// kernels (tcgen05 split-bf16 or fp32 SIMT, lagvae_gemm_auto); training-mode BatchNorm, ELU, residual add and the
// fused sigmoid + Bernoulli NLL are element-wise / column-reduction kernels over [rows, C].
#include "kernels.cuh"

namespace lagvae {

// ---------------------------------------------------------------------------------------------
// im2col / col2im  (nn.Conv2d geometry: enc_resnet_v2.py:14-17,36-40,101; dec_pixelcnn_v2.py:39-47,71-73)
// col[(b, oy, ox), (ty*kw + tx)*C + c] = x[b, oy*stride - pad + ty, ox*stride - pad + tx, c]  (0 outside)
// ---------------------------------------------------------------------------------------------
__global__ void k_im2col(const float* __restrict__ x, int B, int H, int W, int C, int kh, int kw, int stride, int pad,
                         int Ho, int Wo, float* __restrict__ col) {
  const int64_t n = (int64_t)B * Ho * Wo * kh * kw * C;
  for (int64_t i = blockIdx.x * (int64_t)blockDim.x + threadIdx.x; i < n; i += (int64_t)gridDim.x * blockDim.x) {
    const int c = (int)(i % C);
    int64_t r = i / C;
    const int tap = (int)(r % (kh * kw));
    r /= (kh * kw);
    const int ox = (int)(r % Wo);
    r /= Wo;
    const int oy = (int)(r % Ho);
    const int b = (int)(r / Ho);
    const int iy = oy * stride - pad + tap / kw, ix = ox * stride - pad + tap % kw;
    col[i] = (iy >= 0 && iy < H && ix >= 0 && ix < W) ? x[(((int64_t)b * H + iy) * W + ix) * C + c] : 0.f;
  }
}
// gather form of col2im (no atomics): dx[b,iy,ix,c] = sum over taps whose window covers (iy,ix)
__global__ void k_col2im(const float* __restrict__ dcol, int B, int H, int W, int C, int kh, int kw, int stride, int pad,
                         int Ho, int Wo, float* __restrict__ dx) {
  const int64_t n = (int64_t)B * H * W * C;
  const int K = kh * kw * C;
  for (int64_t i = blockIdx.x * (int64_t)blockDim.x + threadIdx.x; i < n; i += (int64_t)gridDim.x * blockDim.x) {
    const int c = (int)(i % C);
    int64_t r = i / C;
    const int ix = (int)(r % W);
    r /= W;
    const int iy = (int)(r % H);
    const int b = (int)(r / H);
    float acc = 0.f;
    for (int ty = 0; ty < kh; ++ty) {
      const int ny = iy + pad - ty;
      if (ny < 0 || ny % stride) continue;
      const int oy = ny / stride;
      if (oy >= Ho) continue;
      for (int tx = 0; tx < kw; ++tx) {
        const int nx = ix + pad - tx;
        if (nx < 0 || nx % stride) continue;
        const int ox = nx / stride;
        if (ox >= Wo) continue;
        acc += dcol[(((int64_t)b * Ho + oy) * Wo + ox) * K + (ty * kw + tx) * C + c];
      }
    }
    dx[i] = acc;
  }
}

// ---------------------------------------------------------------------------------------------
// column sums over [R, C]: out[j] = Σ_r f(r, j) for up to 2 functions, fp32 per thread, fp64 atomics per block
// ---------------------------------------------------------------------------------------------
template <int MODE>  // 0: (x, x^2)   1: (dy, dy * xhat)
__global__ void __launch_bounds__(256)
k_col_reduce2(const float* __restrict__ a, const float* __restrict__ b, const float* __restrict__ mean,
              const float* __restrict__ invstd, int64_t R, int C, int rows_per_block, double* __restrict__ acc) {
  __shared__ float s0[8][33], s1[8][33];
  const int cx = threadIdx.x & 31, ry = threadIdx.x >> 5;
  const int c = blockIdx.x * 32 + cx;
  const int64_t r0 = (int64_t)blockIdx.y * rows_per_block;
  const int64_t r1 = r0 + rows_per_block < R ? r0 + rows_per_block : R;
  float p0 = 0.f, p1 = 0.f;
  if (c < C) {
    const float m = MODE == 1 ? mean[c] : 0.f, is = MODE == 1 ? invstd[c] : 0.f;
    for (int64_t r = r0 + ry; r < r1; r += 8) {
      const float v = a[r * C + c];
      if (MODE == 0) {
        p0 += v;
        p1 = fmaf(v, v, p1);
      } else {
        const float xh = (b[r * C + c] - m) * is;
        p0 += v;
        p1 = fmaf(v, xh, p1);
      }
    }
  }
  s0[ry][cx] = p0;
  s1[ry][cx] = p1;
  __syncthreads();
  if (ry == 0 && c < C) {
    double t0 = 0.0, t1 = 0.0;
#pragma unroll
    for (int k = 0; k < 8; ++k) { t0 += (double)s0[k][cx]; t1 += (double)s1[k][cx]; }
    atomicAdd(acc + c, t0);
    atomicAdd(acc + C + c, t1);
  }
}

// BatchNorm2d, train(): batch statistics (biased variance), running stats updated with the unbiased variance
// (PyTorch defaults eps=1e-5, momentum=0.1 — SURVEY A.8)
__global__ void k_bn_finalize(const double* __restrict__ acc, int64_t R, int C, float eps, float momentum,
                              float* __restrict__ save_mean, float* __restrict__ save_invstd,
                              float* __restrict__ running_mean, float* __restrict__ running_var) {
  const int c = blockIdx.x * blockDim.x + threadIdx.x;
  if (c >= C) return;
  const double mean = acc[c] / (double)R;
  double var = acc[C + c] / (double)R - mean * mean;
  if (var < 0.0) var = 0.0;
  save_mean[c] = (float)mean;
  save_invstd[c] = (float)(1.0 / sqrt(var + (double)eps));
  if (running_mean) running_mean[c] = (1.f - momentum) * running_mean[c] + momentum * (float)mean;
  if (running_var) {
    const double unb = R > 1 ? var * (double)R / (double)(R - 1) : var;
    running_var[c] = (1.f - momentum) * running_var[c] + momentum * (float)unb;
  }
}
// y = (x - mean) * invstd * gamma + beta, optionally followed by ELU(alpha=1)
__global__ void k_bn_apply(const float* __restrict__ x, int64_t n, int C, const float* __restrict__ mean,
                           const float* __restrict__ invstd, const float* __restrict__ gamma,
                           const float* __restrict__ beta, float* __restrict__ y) {
  for (int64_t i = blockIdx.x * (int64_t)blockDim.x + threadIdx.x; i < n; i += (int64_t)gridDim.x * blockDim.x) {
    const int c = (int)(i % C);
    y[i] = (x[i] - mean[c]) * invstd[c] * gamma[c] + beta[c];
  }
}
// dx = gamma*invstd*(dy - Σdy/R - xhat*Σ(dy*xhat)/R); dgamma = Σ dy*xhat; dbeta = Σ dy
__global__ void k_bn_bwd_apply(const float* __restrict__ x, const float* __restrict__ dy, int64_t n, int C, int64_t R,
                               const float* __restrict__ mean, const float* __restrict__ invstd,
                               const float* __restrict__ gamma, const double* __restrict__ acc, float* __restrict__ dx) {
  for (int64_t i = blockIdx.x * (int64_t)blockDim.x + threadIdx.x; i < n; i += (int64_t)gridDim.x * blockDim.x) {
    const int c = (int)(i % C);
    const float xh = (x[i] - mean[c]) * invstd[c];
    const float sdy = (float)(acc[c] / (double)R), sdyx = (float)(acc[C + c] / (double)R);
    dx[i] = gamma[c] * invstd[c] * (dy[i] - sdy - xh * sdyx);
  }
}
__global__ void k_bn_param_grads(const double* __restrict__ acc, int C, float* __restrict__ dgamma, float* __restrict__ dbeta) {
  const int c = blockIdx.x * blockDim.x + threadIdx.x;
  if (c >= C) return;
  dbeta[c] = (float)acc[c];
  dgamma[c] = (float)acc[C + c];
}

static int grid_for(int64_t n) { return (int)std::min<int64_t>(cdiv(n, 256), 148 * 16); }

int bn_train_fwd(const float* x, int64_t R, int C, const float* gamma, const float* beta, float eps, float momentum,
                 float* y, float* save_mean, float* save_invstd, float* running_mean, float* running_var, double* scratch,
                 cudaStream_t st) {
  LV_CUDA(cudaMemsetAsync(scratch, 0, sizeof(double) * 2 * C, st));
  const int rpb = (int)std::max<int64_t>(8, cdiv(R, 64));
  dim3 grid((unsigned)cdiv(C, 32), (unsigned)cdiv(R, rpb));
  k_col_reduce2<0><<<grid, 256, 0, st>>>(x, nullptr, nullptr, nullptr, R, C, rpb, scratch);
  LV_LAUNCH_CHECK();
  k_bn_finalize<<<(int)cdiv(C, 128), 128, 0, st>>>(scratch, R, C, eps, momentum, save_mean, save_invstd, running_mean, running_var);
  LV_LAUNCH_CHECK();
  k_bn_apply<<<grid_for(R * C), 256, 0, st>>>(x, R * C, C, save_mean, save_invstd, gamma, beta, y);
  LV_LAUNCH_CHECK();
  return LAGVAE_OK;
}
int bn_train_bwd(const float* x, const float* dy, int64_t R, int C, const float* gamma, const float* save_mean,
                 const float* save_invstd, float* dx, float* dgamma, float* dbeta, double* scratch, cudaStream_t st) {
  LV_CUDA(cudaMemsetAsync(scratch, 0, sizeof(double) * 2 * C, st));
  const int rpb = (int)std::max<int64_t>(8, cdiv(R, 64));
  dim3 grid((unsigned)cdiv(C, 32), (unsigned)cdiv(R, rpb));
  k_col_reduce2<1><<<grid, 256, 0, st>>>(dy, x, save_mean, save_invstd, R, C, rpb, scratch);
  LV_LAUNCH_CHECK();
  k_bn_param_grads<<<(int)cdiv(C, 128), 128, 0, st>>>(scratch, C, dgamma, dbeta);
  LV_LAUNCH_CHECK();
  k_bn_bwd_apply<<<grid_for(R * C), 256, 0, st>>>(x, dy, R * C, C, R, save_mean, save_invstd, gamma, scratch, dx);
  LV_LAUNCH_CHECK();
  return LAGVAE_OK;
}

// ---------------------------------------------------------------------------------------------
// ELU(alpha=1) (+ optional residual add before it) — enc_resnet_v2.py:32,69; dec_pixelcnn_v2.py:41,45,49,61
// ---------------------------------------------------------------------------------------------
__global__ void k_elu_fwd(const float* __restrict__ a, const float* __restrict__ b, float* __restrict__ y, int64_t n) {
  for (int64_t i = blockIdx.x * (int64_t)blockDim.x + threadIdx.x; i < n; i += (int64_t)gridDim.x * blockDim.x) {
    const float v = a[i] + (b ? b[i] : 0.f);
    y[i] = v > 0.f ? v : expm1f(v);
  }
}
__global__ void k_elu_bwd(const float* __restrict__ y, const float* __restrict__ dy, float* __restrict__ dx, int64_t n) {
  for (int64_t i = blockIdx.x * (int64_t)blockDim.x + threadIdx.x; i < n; i += (int64_t)gridDim.x * blockDim.x) {
    const float o = y[i];
    dx[i] = dy[i] * (o > 0.f ? 1.f : o + 1.f);
  }
}
__global__ void k_add(const float* __restrict__ a, const float* __restrict__ b, float* __restrict__ o, int64_t n) {
  for (int64_t i = blockIdx.x * (int64_t)blockDim.x + threadIdx.x; i < n; i += (int64_t)gridDim.x * blockDim.x) o[i] = a[i] + b[i];
}

// ---------------------------------------------------------------------------------------------
// sigmoid + Bernoulli NLL with the reference's +1e-12 inside both logs (dec_pixelcnn_v2.py:145-152,173,193-195)
// logits [B*ns, P] (row b*ns+s), x [B, P];  nll[b*ns+s] = -Σ_p x log(p+eps) + (1-x) log(1-p+eps)
// ---------------------------------------------------------------------------------------------
__global__ void __launch_bounds__(256)
k_bernoulli_nll_fwd(const float* __restrict__ logits, const float* __restrict__ x, int ns, int P, float* __restrict__ nll) {
  __shared__ float red[32];
  const int row = blockIdx.x, b = row / ns;
  float a = 0.f;
  for (int p = threadIdx.x; p < P; p += blockDim.x) {
    const float pr = 1.f / (1.f + expf(-logits[(int64_t)row * P + p]));
    const float t = x[(int64_t)b * P + p];
    a += logf(pr + 1e-12f) * t + logf(1.f - pr + 1e-12f) * (1.f - t);
  }
  a = block_sum(a, red);
  if (threadIdx.x == 0) nll[row] = -a;
}
__global__ void k_bernoulli_nll_bwd(const float* __restrict__ logits, const float* __restrict__ x, const float* __restrict__ g,
                                    int ns, int P, int64_t n, float* __restrict__ dlogits) {
  for (int64_t i = blockIdx.x * (int64_t)blockDim.x + threadIdx.x; i < n; i += (int64_t)gridDim.x * blockDim.x) {
    const int row = (int)(i / P), p = (int)(i % P), b = row / ns;
    const float pr = 1.f / (1.f + expf(-logits[i]));
    const float t = x[(int64_t)b * P + p];
    const float dpr = -(t / (pr + 1e-12f) - (1.f - t) / (1.f - pr + 1e-12f));
    dlogits[i] = g[row] * dpr * pr * (1.f - pr);
  }
}

}  // namespace lagvae

using namespace lagvae;

extern "C" {

int lagvae_im2col(const float* x, int B, int H, int W, int C, int kh, int kw, int stride, int pad, float* col, void* stream) {
  LV_CHECK_ARG(x && col && B > 0 && H > 0 && W > 0 && C > 0 && kh > 0 && kw > 0 && stride > 0 && pad >= 0, "im2col: bad argument");
  const int Ho = (H + 2 * pad - kh) / stride + 1, Wo = (W + 2 * pad - kw) / stride + 1;
  LV_CHECK_ARG(Ho > 0 && Wo > 0, "im2col: empty output");
  const int64_t n = (int64_t)B * Ho * Wo * kh * kw * C;
  k_im2col<<<grid_for(n), 256, 0, (cudaStream_t)stream>>>(x, B, H, W, C, kh, kw, stride, pad, Ho, Wo, col);
  LV_LAUNCH_CHECK();
  return LAGVAE_OK;
}
int lagvae_col2im(const float* dcol, int B, int H, int W, int C, int kh, int kw, int stride, int pad, float* dx, void* stream) {
  LV_CHECK_ARG(dcol && dx && B > 0 && H > 0 && W > 0 && C > 0 && kh > 0 && kw > 0 && stride > 0 && pad >= 0, "col2im: bad argument");
  const int Ho = (H + 2 * pad - kh) / stride + 1, Wo = (W + 2 * pad - kw) / stride + 1;
  k_col2im<<<grid_for((int64_t)B * H * W * C), 256, 0, (cudaStream_t)stream>>>(dcol, B, H, W, C, kh, kw, stride, pad, Ho, Wo, dx);
  LV_LAUNCH_CHECK();
  return LAGVAE_OK;
}
int lagvae_bn_train_fwd(const float* x, int64_t R, int C, const float* gamma, const float* beta, float eps, float momentum,
                        float* y, float* save_mean, float* save_invstd, float* running_mean, float* running_var,
                        void* scratch, void* stream) {
  LV_CHECK_ARG(x && gamma && beta && y && save_mean && save_invstd && scratch && R > 0 && C > 0, "bn_train_fwd: bad argument");
  return bn_train_fwd(x, R, C, gamma, beta, eps, momentum, y, save_mean, save_invstd, running_mean, running_var,
                      (double*)scratch, (cudaStream_t)stream);
}
int lagvae_bn_train_bwd(const float* x, const float* dy, int64_t R, int C, const float* gamma, const float* save_mean,
                        const float* save_invstd, float* dx, float* dgamma, float* dbeta, void* scratch, void* stream) {
  LV_CHECK_ARG(x && dy && gamma && save_mean && save_invstd && dx && dgamma && dbeta && scratch && R > 0 && C > 0,
               "bn_train_bwd: bad argument");
  return bn_train_bwd(x, dy, R, C, gamma, save_mean, save_invstd, dx, dgamma, dbeta, (double*)scratch, (cudaStream_t)stream);
}
int lagvae_bn_apply(const float* x, int64_t R, int C, const float* mean, const float* invstd, const float* gamma,
                    const float* beta, float* y, void* stream) {
  LV_CHECK_ARG(x && mean && invstd && gamma && beta && y && R > 0 && C > 0, "bn_apply: bad argument");
  k_bn_apply<<<grid_for(R * C), 256, 0, (cudaStream_t)stream>>>(x, R * C, C, mean, invstd, gamma, beta, y);
  LV_LAUNCH_CHECK();
  return LAGVAE_OK;
}
int lagvae_elu_fwd(const float* a, const float* b_or_null, float* y, int64_t n, void* stream) {
  LV_CHECK_ARG(a && y && n >= 0, "elu_fwd: bad argument");
  if (n == 0) return LAGVAE_OK;
  k_elu_fwd<<<grid_for(n), 256, 0, (cudaStream_t)stream>>>(a, b_or_null, y, n);
  LV_LAUNCH_CHECK();
  return LAGVAE_OK;
}
int lagvae_elu_bwd(const float* y, const float* dy, float* dx, int64_t n, void* stream) {
  LV_CHECK_ARG(y && dy && dx && n >= 0, "elu_bwd: bad argument");
  if (n == 0) return LAGVAE_OK;
  k_elu_bwd<<<grid_for(n), 256, 0, (cudaStream_t)stream>>>(y, dy, dx, n);
  LV_LAUNCH_CHECK();
  return LAGVAE_OK;
}
int lagvae_add(const float* a, const float* b, float* out, int64_t n, void* stream) {
  LV_CHECK_ARG(a && b && out && n >= 0, "add: bad argument");
  if (n == 0) return LAGVAE_OK;
  k_add<<<grid_for(n), 256, 0, (cudaStream_t)stream>>>(a, b, out, n);
  LV_LAUNCH_CHECK();
  return LAGVAE_OK;
}
int lagvae_bernoulli_nll_fwd(const float* logits, const float* x, int B, int ns, int P, float* nll, void* stream) {
  LV_CHECK_ARG(logits && x && nll && B > 0 && ns > 0 && P > 0, "bernoulli_nll_fwd: bad argument");
  k_bernoulli_nll_fwd<<<B * ns, 256, 0, (cudaStream_t)stream>>>(logits, x, ns, P, nll);
  LV_LAUNCH_CHECK();
  return LAGVAE_OK;
}
int lagvae_bernoulli_nll_bwd(const float* logits, const float* x, const float* g, int B, int ns, int P, float* dlogits,
                             void* stream) {
  LV_CHECK_ARG(logits && x && g && dlogits && B > 0 && ns > 0 && P > 0, "bernoulli_nll_bwd: bad argument");
  const int64_t n = (int64_t)B * ns * P;
  k_bernoulli_nll_bwd<<<grid_for(n), 256, 0, (cudaStream_t)stream>>>(logits, x, g, ns, P, n, dlogits);
  LV_LAUNCH_CHECK();
  return LAGVAE_OK;
}

// reparameterise + KL on given posterior stats (encoder.py:55,72-79) — image encoders produce (mu, logvar) themselves
__global__ void k_reparam_kl_fwd(const float* __restrict__ mu, const float* __restrict__ logvar, const float* __restrict__ eps,
                                 int B, int nz, int ns, float* __restrict__ z, float* __restrict__ kl) {
  __shared__ float red[32];
  const int b = blockIdx.x;
  float part = 0.f;
  for (int j = threadIdx.x; j < nz; j += blockDim.x) {
    const float m = mu[(int64_t)b * nz + j], lv = logvar[(int64_t)b * nz + j], sd = expf(0.5f * lv);
    for (int s = 0; s < ns; ++s) {
      const int64_t o = ((int64_t)b * ns + s) * nz + j;
      z[o] = m + eps[o] * sd;
    }
    part += 0.5f * (m * m + expf(lv) - lv - 1.f);
  }
  part = block_sum(part, red);
  if (threadIdx.x == 0) kl[b] = part;
}
int lagvae_reparam_kl_fwd(const float* mu, const float* logvar, const float* eps, int B, int nz, int ns, float* z, float* kl,
                          void* stream) {
  LV_CHECK_ARG(mu && logvar && eps && z && kl && B > 0 && nz > 0 && ns > 0, "reparam_kl_fwd: bad argument");
  k_reparam_kl_fwd<<<B, 128, 0, (cudaStream_t)stream>>>(mu, logvar, eps, B, nz, ns, z, kl);
  LV_LAUNCH_CHECK();
  return LAGVAE_OK;
}
int lagvae_reparam_kl_bwd(const float* dz, const float* eps, const float* mu, const float* logvar, const float* g_kl, int B,
                          int nz, int ns, float* dml, void* stream) {
  LV_CHECK_ARG(eps && mu && logvar && dml && B > 0 && nz > 0 && ns > 0, "reparam_kl_bwd: bad argument");
  return reparam_kl_bwd(dz, eps, mu, logvar, g_kl, B, nz, ns, dml, (cudaStream_t)stream);
}

// C = alpha * opA · opBᵀ (+ beta C)(+ bias_n): tcgen05 split-bf16 (3 passes) when the problem is tensor-core sized and
// the operands can be staged into `scratch` (>= lagvae_gemm_auto_scratch_bytes), fp32 SIMT otherwise.
size_t lagvae_gemm_auto_scratch_bytes(int M, int N, int K) {
  const size_t a = (size_t)round_up(M, 8) * (size_t)round_up(K, 8), b = (size_t)round_up(N, 8) * (size_t)round_up(K, 8);
  return (a + b) * 2 /*hi, lo*/ * sizeof(uint16_t) + 8 * 256;
}
int lagvae_gemm_auto(const float* A, int64_t a_rs, int64_t a_cs, const float* Bm, int64_t b_rs, int64_t b_cs, float* Cm,
                     int64_t ldc, int M, int N, int K, float alpha, float beta, const float* bias_n, void* scratch,
                     size_t scratch_bytes, void* stream) {
  LV_CHECK_ARG(A && Bm && Cm && M > 0 && N > 0 && K > 0, "gemm_auto: bad argument");
  cudaStream_t st = (cudaStream_t)stream;
  const bool contiguous = (a_rs == 1 || a_cs == 1) && (b_rs == 1 || b_cs == 1);
  const bool big = M >= 64 && N >= 16 && K >= 32 && (int64_t)M * N * K >= (int64_t)1 << 22;
  if (contiguous && big && scratch && scratch_bytes >= lagvae_gemm_auto_scratch_bytes(M, N, K)) {
    // stage A and B as stored (row-major [rows, ld]) into hi/lo bf16 with ld padded to 8
    const bool a_mn = a_cs != 1, b_mn = b_cs != 1;   // stored [K, M] / [K, N]
    const int64_t a_rows = a_mn ? K : M, a_cols = a_mn ? M : K, a_ld = a_mn ? a_cs : a_rs;
    const int64_t b_rows = b_mn ? K : N, b_cols = b_mn ? N : K, b_ld = b_mn ? b_cs : b_rs;
    const int64_t a_ldo = round_up(a_cols, 8), b_ldo = round_up(b_cols, 8);
    char* p = (char*)(((uintptr_t)scratch + 255) & ~(uintptr_t)255);
    auto take = [&](int64_t elems) { uint16_t* q = (uint16_t*)p; p += round_up(elems * 2, 256); return q; };
    uint16_t *ah = take(a_rows * a_ldo), *al = take(a_rows * a_ldo), *bh = take(b_rows * b_ldo), *bl = take(b_rows * b_ldo);
    if ((size_t)(p - (char*)scratch) <= scratch_bytes) {
      LV_TRY(split_bf16_launch(A, a_ld, (int)a_rows, (int)a_cols, ah, al, a_ldo, st));
      LV_TRY(split_bf16_launch(Bm, b_ld, (int)b_rows, (int)b_cols, bh, bl, b_ldo, st));
      TcOperand ta{ah, al, a_ldo, a_mn ? 1 : 0}, tb{bh, bl, b_ldo, b_mn ? 1 : 0};
      return gemm_tc(ta, tb, Cm, ldc, M, N, K, 3, alpha, beta, bias_n, nullptr, 0, nullptr, st);
    }
  }
  return gemm_f32(A, a_rs, a_cs, Bm, b_rs, b_cs, Cm, ldc, M, N, K, alpha, beta, bias_n, nullptr, 0, st);
}

}  // extern "C"
