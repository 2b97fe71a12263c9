// Fused BatchNorm(train) + residual + ELU kernels of the PixelCNN blocks (dec_pixelcnn_v2.py:32-62:
// conv -> BN -> ELU -> masked conv -> BN -> ELU -> conv -> BN, out = ELU(main(x) + x)).
//
// The convolution kernels (conv_tc.cu) leave the per-channel sums (Σy, Σy²) of their output in a double[2C] buffer, so
// the BatchNorm statistics cost no extra pass; these kernels then do, in ONE pass over the activation each:
//   forward : mean / invstd from the sums (every block recomputes the 32-64 values; block 0 also updates the running
//             statistics exactly as nn.BatchNorm2d does: momentum 0.1, unbiased variance), normalise, (+ residual),
//             ELU, and emit the next layer's operand directly in the bf16 [hi | lo] "cat" format (and/or fp32).
//   backward: pass 1 reduces Σdpre and Σdpre·xhat per channel (dpre = dout · ELU'(out) recomputed from the stored
//             output); pass 2 applies dy = γ·invstd·(dpre − Σdpre/R − xhat·Σ(dpre·xhat)/R) and emits dy in the cat
//             format for the dgrad / wgrad kernels (+ fp32, + the residual-branch gradient dpre).
// All HBM accesses are 16-byte (fp32 x4) / 8-byte (bf16 x4) vectors.
#include "kernels.cuh"

namespace lagvae {

namespace {

__device__ __forceinline__ float elu1(float v) { return v > 0.f ? v : expm1f(v); }

__device__ __forceinline__ void store_cat4(__nv_bfloat16* cat, int64_t r, int C, int c0, const float v[4]) {
  __align__(8) __nv_bfloat16 hi[4], lo[4];
#pragma unroll
  for (int e = 0; e < 4; ++e) split_bf16(v[e], hi[e], lo[e]);
  *(uint2*)(cat + r * 2 * C + c0) = *(const uint2*)hi;
  *(uint2*)(cat + r * 2 * C + C + c0) = *(const uint2*)lo;
}
__device__ __forceinline__ void load_cat4(const __nv_bfloat16* cat, int64_t r, int C, int c0, float v[4]) {
  __align__(8) __nv_bfloat16 hi[4], lo[4];
  *(uint2*)hi = *(const uint2*)(cat + r * 2 * C + c0);
  *(uint2*)lo = *(const uint2*)(cat + r * 2 * C + C + c0);
#pragma unroll
  for (int e = 0; e < 4; ++e) v[e] = __bfloat162float(hi[e]) + __bfloat162float(lo[e]);
}

template <int C>
__global__ void __launch_bounds__(256)
k_bnact_fwd(const float* __restrict__ y, const double* __restrict__ stats, int64_t R, const float* __restrict__ gamma,
            const float* __restrict__ beta, float eps, float momentum, const float* __restrict__ res, int elu,
            float* __restrict__ out, __nv_bfloat16* __restrict__ cat, float* __restrict__ save_mean,
            float* __restrict__ save_invstd, float* __restrict__ running_mean, float* __restrict__ running_var) {
  __shared__ float s_mean[C], s_scale[C], s_beta[C];
  if (threadIdx.x < C) {
    const int c = threadIdx.x;
    const double mean = stats[c] / (double)R;
    double var = stats[C + c] / (double)R - mean * mean;
    if (var < 0.0) var = 0.0;
    const float invstd = (float)(1.0 / sqrt(var + (double)eps));
    s_mean[c] = (float)mean;
    s_scale[c] = invstd * gamma[c];
    s_beta[c] = beta[c];
    if (blockIdx.x == 0) {
      save_mean[c] = (float)mean;
      save_invstd[c] = invstd;
      if (running_mean) running_mean[c] = (1.f - momentum) * running_mean[c] + momentum * (float)mean;
      if (running_var) {
        const double unb = R > 1 ? var * (double)R / (double)(R - 1) : var;
        running_var[c] = (1.f - momentum) * running_var[c] + momentum * (float)unb;
      }
    }
  }
  __syncthreads();
  constexpr int TPR = C / 4;
  const int64_t n = R * TPR;
  for (int64_t i = blockIdx.x * (int64_t)blockDim.x + threadIdx.x; i < n; i += (int64_t)gridDim.x * blockDim.x) {
    const int64_t r = i / TPR;
    const int c0 = (int)(i % TPR) * 4;
    const float4 yv = *(const float4*)(y + r * C + c0);
    float v[4] = {yv.x, yv.y, yv.z, yv.w};
#pragma unroll
    for (int e = 0; e < 4; ++e) v[e] = (v[e] - s_mean[c0 + e]) * s_scale[c0 + e] + s_beta[c0 + e];
    if (res) {
      const float4 rv = *(const float4*)(res + r * C + c0);
      v[0] += rv.x; v[1] += rv.y; v[2] += rv.z; v[3] += rv.w;
    }
    if (elu) {
#pragma unroll
      for (int e = 0; e < 4; ++e) v[e] = elu1(v[e]);
    }
    if (out) *(float4*)(out + r * C + c0) = make_float4(v[0], v[1], v[2], v[3]);
    if (cat) store_cat4(cat, r, C, c0, v);
  }
}

// dpre of 4 channels of row r: dout * ELU'(out) (out from fp32 or from the cat operand copy)
template <int C>
__device__ __forceinline__ void load_dpre(const float* dout, const float* out, const __nv_bfloat16* ocat, int elu, int64_t r,
                                          int c0, float d[4]) {
  const float4 dv = *(const float4*)(dout + r * C + c0);
  d[0] = dv.x; d[1] = dv.y; d[2] = dv.z; d[3] = dv.w;
  if (elu) {
    float o[4];
    if (out) {
      const float4 ov = *(const float4*)(out + r * C + c0);
      o[0] = ov.x; o[1] = ov.y; o[2] = ov.z; o[3] = ov.w;
    } else {
      load_cat4(ocat, r, C, c0, o);
    }
#pragma unroll
    for (int e = 0; e < 4; ++e) d[e] *= (o[e] > 0.f ? 1.f : o[e] + 1.f);
  }
}

template <int C>
__global__ void __launch_bounds__(256)
k_bnact_bwd_reduce(const float* __restrict__ dout, const float* __restrict__ out, const __nv_bfloat16* __restrict__ ocat,
                   const float* __restrict__ y, int64_t R, const float* __restrict__ save_mean,
                   const float* __restrict__ save_invstd, int elu, double* __restrict__ acc) {
  constexpr int TPR = C / 4, RL = 256 / TPR;
  __shared__ float sm[256][9];
  const int cg = threadIdx.x % TPR, rl = threadIdx.x / TPR, c0 = cg * 4;
  float mean[4], istd[4], sd[4] = {0.f, 0.f, 0.f, 0.f}, sx[4] = {0.f, 0.f, 0.f, 0.f};
#pragma unroll
  for (int e = 0; e < 4; ++e) { mean[e] = save_mean[c0 + e]; istd[e] = save_invstd[c0 + e]; }
  for (int64_t r = (int64_t)blockIdx.x * RL + rl; r < R; r += (int64_t)gridDim.x * RL) {
    float d[4];
    load_dpre<C>(dout, out, ocat, elu, r, c0, d);
    const float4 yv = *(const float4*)(y + r * C + c0);
    const float yy[4] = {yv.x, yv.y, yv.z, yv.w};
#pragma unroll
    for (int e = 0; e < 4; ++e) {
      sd[e] += d[e];
      sx[e] = fmaf(d[e], (yy[e] - mean[e]) * istd[e], sx[e]);
    }
  }
#pragma unroll
  for (int e = 0; e < 4; ++e) { sm[threadIdx.x][e] = sd[e]; sm[threadIdx.x][4 + e] = sx[e]; }
  __syncthreads();
  if (threadIdx.x < 2 * C) {
    const int which = threadIdx.x / C, c = threadIdx.x % C;
    double t = 0.0;
    for (int k = 0; k < RL; ++k) t += (double)sm[k * TPR + c / 4][which * 4 + (c & 3)];
    atomicAdd(acc + which * C + c, t);
  }
}

template <int C>
__global__ void __launch_bounds__(256)
k_bnact_bwd_apply(const float* __restrict__ dout, const float* __restrict__ out, const __nv_bfloat16* __restrict__ ocat,
                  const float* __restrict__ y, int64_t R, const float* __restrict__ gamma, const float* __restrict__ save_mean,
                  const float* __restrict__ save_invstd, int elu, const double* __restrict__ acc, float* __restrict__ dy,
                  __nv_bfloat16* __restrict__ dycat, float* __restrict__ dres, float* __restrict__ dgamma,
                  float* __restrict__ dbeta) {
  __shared__ float s_mean[C], s_istd[C], s_g[C], s_sd[C], s_sx[C];
  if (threadIdx.x < C) {
    const int c = threadIdx.x;
    s_mean[c] = save_mean[c];
    s_istd[c] = save_invstd[c];
    s_g[c] = gamma[c] * save_invstd[c];
    s_sd[c] = (float)(acc[c] / (double)R);
    s_sx[c] = (float)(acc[C + c] / (double)R);
    if (blockIdx.x == 0) {
      dbeta[c] = (float)acc[c];
      dgamma[c] = (float)acc[C + c];
    }
  }
  __syncthreads();
  constexpr int TPR = C / 4;
  const int64_t n = R * TPR;
  for (int64_t i = blockIdx.x * (int64_t)blockDim.x + threadIdx.x; i < n; i += (int64_t)gridDim.x * blockDim.x) {
    const int64_t r = i / TPR;
    const int c0 = (int)(i % TPR) * 4;
    float d[4];
    load_dpre<C>(dout, out, ocat, elu, r, c0, d);
    const float4 yv = *(const float4*)(y + r * C + c0);
    const float yy[4] = {yv.x, yv.y, yv.z, yv.w};
    float g[4];
#pragma unroll
    for (int e = 0; e < 4; ++e) {
      const float xh = (yy[e] - s_mean[c0 + e]) * s_istd[c0 + e];
      g[e] = s_g[c0 + e] * (d[e] - s_sd[c0 + e] - xh * s_sx[c0 + e]);
    }
    if (dres) *(float4*)(dres + r * C + c0) = make_float4(d[0], d[1], d[2], d[3]);
    if (dy) *(float4*)(dy + r * C + c0) = make_float4(g[0], g[1], g[2], g[3]);
    if (dycat) store_cat4(dycat, r, C, c0, g);
  }
}

// eval(): sums that make k_bnact_fwd reproduce the running statistics (mean = Σ/R, var = Σ²/R - mean²)
__global__ void k_bn_eval_stats(const float* __restrict__ rm, const float* __restrict__ rv, int64_t R, int C, double* __restrict__ stats) {
  const int c = blockIdx.x * blockDim.x + threadIdx.x;
  if (c >= C) return;
  const double m = (double)rm[c], v = (double)rv[c];
  stats[c] = m * (double)R;
  stats[C + c] = (v + m * m) * (double)R;
}

int grid_rows(int64_t n) { return (int)std::min<int64_t>(cdiv(n, 256), 148 * 8); }

}  // namespace

}  // namespace lagvae

using namespace lagvae;

extern "C" {

int lagvae_bnact_fwd(const float* y, const double* stats, int64_t R, int C, const float* gamma, const float* beta, float eps,
                     float momentum, const float* residual_or_null, int elu, float* out_f32_or_null, uint16_t* out_cat_or_null,
                     float* save_mean, float* save_invstd, float* running_mean, float* running_var, void* stream) {
  LV_CHECK_ARG(y && stats && gamma && beta && save_mean && save_invstd && R > 0 && (C == 32 || C == 64) &&
               (out_f32_or_null || out_cat_or_null), "bnact_fwd: bad argument");
  cudaStream_t st = (cudaStream_t)stream;
  const int grid = grid_rows(R * (C / 4));
  if (C == 32)
    k_bnact_fwd<32><<<grid, 256, 0, st>>>(y, stats, R, gamma, beta, eps, momentum, residual_or_null, elu, out_f32_or_null,
                                          (__nv_bfloat16*)out_cat_or_null, save_mean, save_invstd, running_mean, running_var);
  else
    k_bnact_fwd<64><<<grid, 256, 0, st>>>(y, stats, R, gamma, beta, eps, momentum, residual_or_null, elu, out_f32_or_null,
                                          (__nv_bfloat16*)out_cat_or_null, save_mean, save_invstd, running_mean, running_var);
  LV_LAUNCH_CHECK();
  return LAGVAE_OK;
}

int lagvae_bn_eval_stats(const float* running_mean, const float* running_var, int64_t R, int C, double* stats, void* stream) {
  LV_CHECK_ARG(running_mean && running_var && stats && R > 0 && C > 0, "bn_eval_stats: bad argument");
  k_bn_eval_stats<<<(int)cdiv(C, 128), 128, 0, (cudaStream_t)stream>>>(running_mean, running_var, R, C, stats);
  LV_LAUNCH_CHECK();
  return LAGVAE_OK;
}

int lagvae_bnact_bwd(const float* dout, const float* out_f32_or_null, const uint16_t* out_cat_or_null, const float* y, int64_t R,
                     int C, const float* gamma, const float* save_mean, const float* save_invstd, int elu, float* dy_f32_or_null,
                     uint16_t* dy_cat_or_null, float* dres_or_null, float* dgamma, float* dbeta, void* scratch, void* stream) {
  LV_CHECK_ARG(dout && y && gamma && save_mean && save_invstd && dgamma && dbeta && scratch && R > 0 && (C == 32 || C == 64) &&
               (!elu || out_f32_or_null || out_cat_or_null), "bnact_bwd: bad argument");
  cudaStream_t st = (cudaStream_t)stream;
  double* acc = (double*)scratch;
  LV_CUDA(cudaMemsetAsync(acc, 0, 2 * C * sizeof(double), st));
  const __nv_bfloat16* ocat = (const __nv_bfloat16*)out_cat_or_null;
  const int rgrid = (int)std::min<int64_t>(cdiv(R, 256 / (C / 4)), 148 * 4);
  const int grid = grid_rows(R * (C / 4));
  if (C == 32) {
    k_bnact_bwd_reduce<32><<<rgrid, 256, 0, st>>>(dout, out_f32_or_null, ocat, y, R, save_mean, save_invstd, elu, acc);
    LV_LAUNCH_CHECK();
    k_bnact_bwd_apply<32><<<grid, 256, 0, st>>>(dout, out_f32_or_null, ocat, y, R, gamma, save_mean, save_invstd, elu, acc,
                                                dy_f32_or_null, (__nv_bfloat16*)dy_cat_or_null, dres_or_null, dgamma, dbeta);
  } else {
    k_bnact_bwd_reduce<64><<<rgrid, 256, 0, st>>>(dout, out_f32_or_null, ocat, y, R, save_mean, save_invstd, elu, acc);
    LV_LAUNCH_CHECK();
    k_bnact_bwd_apply<64><<<grid, 256, 0, st>>>(dout, out_f32_or_null, ocat, y, R, gamma, save_mean, save_invstd, elu, acc,
                                                dy_f32_or_null, (__nv_bfloat16*)dy_cat_or_null, dres_or_null, dgamma, dbeta);
  }
  LV_LAUNCH_CHECK();
  return LAGVAE_OK;
}

}  // extern "C"
