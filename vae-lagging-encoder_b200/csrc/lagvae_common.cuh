// Shared host/device helpers for liblagvae (sm_100a only).
#pragma once
#include <cuda_runtime.h>
#include <cuda_bf16.h>
#include <stdint.h>
#include <stdio.h>
#include <atomic>
#include <string>

#include "../../include/lagvae.h"

namespace lagvae {

// ---- error plumbing -------------------------------------------------------------------------
void set_error(const char* fmt, ...);
extern std::atomic<int64_t> g_launches;

#define LV_CHECK_ARG(cond, ...)                 \
  do {                                          \
    if (!(cond)) {                              \
      lagvae::set_error(__VA_ARGS__);           \
      return LAGVAE_E_ARG;                      \
    }                                           \
  } while (0)

#define LV_CUDA(expr)                                                                   \
  do {                                                                                  \
    cudaError_t _e = (expr);                                                            \
    if (_e != cudaSuccess) {                                                            \
      lagvae::set_error("%s:%d %s -> %s", __FILE__, __LINE__, #expr, cudaGetErrorString(_e)); \
      return LAGVAE_E_CUDA;                                                             \
    }                                                                                   \
  } while (0)

#define LV_LAUNCH_CHECK()                         \
  do {                                            \
    lagvae::g_launches.fetch_add(1);              \
    LV_CUDA(cudaGetLastError());                  \
  } while (0)

#define LV_TRY(expr)            \
  do {                          \
    int _s = (expr);            \
    if (_s != LAGVAE_OK) return _s; \
  } while (0)

static inline int64_t cdiv(int64_t a, int64_t b) { return (a + b - 1) / b; }
static inline int64_t round_up(int64_t a, int64_t b) { return cdiv(a, b) * b; }

// ---- device helpers ---------------------------------------------------------------------------
__device__ __forceinline__ float sigmoidf_(float x) { return 1.0f / (1.0f + expf(-x)); }

__device__ __forceinline__ float warp_sum(float v) {
#pragma unroll
  for (int o = 16; o > 0; o >>= 1) v += __shfl_xor_sync(0xffffffffu, v, o);
  return v;
}
__device__ __forceinline__ float warp_max(float v) {
#pragma unroll
  for (int o = 16; o > 0; o >>= 1) v = fmaxf(v, __shfl_xor_sync(0xffffffffu, v, o));
  return v;
}

// block-wide sum; `red` must hold >= 32 floats; result valid in every thread
__device__ __forceinline__ float block_sum(float v, float* red) {
  const int lane = threadIdx.x & 31, w = threadIdx.x >> 5, nw = (blockDim.x + 31) >> 5;
  v = warp_sum(v);
  __syncthreads();
  if (lane == 0) red[w] = v;
  __syncthreads();
  float r = (lane < nw) ? red[lane] : 0.f;
  r = warp_sum(r);
  return r;
}
__device__ __forceinline__ float block_max(float v, float* red) {
  const int lane = threadIdx.x & 31, w = threadIdx.x >> 5, nw = (blockDim.x + 31) >> 5;
  v = warp_max(v);
  __syncthreads();
  if (lane == 0) red[w] = v;
  __syncthreads();
  float r = (lane < nw) ? red[lane] : -INFINITY;
  r = warp_max(r);
  return r;
}

// bf16 hi/lo split: x ~= hi + lo with |x - hi - lo| <= 2^-17 |x|
__device__ __forceinline__ void split_bf16(float x, __nv_bfloat16& hi, __nv_bfloat16& lo) {
  hi = __float2bfloat16_rn(x);
  lo = __float2bfloat16_rn(x - __bfloat162float(hi));
}

// ---- Philox4x32-10 (counter-based dropout; lagvae_dropout mode 2) -----------------------------
__host__ __device__ __forceinline__ uint32_t mulhi32(uint32_t a, uint32_t b) {
#ifdef __CUDA_ARCH__
  return __umulhi(a, b);
#else
  return (uint32_t)(((uint64_t)a * b) >> 32);
#endif
}
__host__ __device__ __forceinline__ void philox4x32_10(uint32_t c[4], uint32_t k0, uint32_t k1) {
#pragma unroll
  for (int r = 0; r < 10; ++r) {
    const uint32_t hi0 = mulhi32(0xD2511F53u, c[0]), lo0 = 0xD2511F53u * c[0];
    const uint32_t hi1 = mulhi32(0xCD9E8D57u, c[2]), lo1 = 0xCD9E8D57u * c[2];
    const uint32_t n0 = hi1 ^ c[1] ^ k0, n1 = lo1, n2 = hi0 ^ c[3] ^ k1, n3 = lo0;
    c[0] = n0; c[1] = n1; c[2] = n2; c[3] = n3;
    k0 += 0x9E3779B9u; k1 += 0xBB67AE85u;
  }
}
// keep decision for logical element `idx` of dropout stream `sid`
__host__ __device__ __forceinline__ bool philox_keep(uint64_t seed, uint32_t sid, uint64_t idx, float p) {
  uint32_t c[4] = {(uint32_t)(idx >> 2), (uint32_t)(idx >> 34), sid, 0x1a9fae5u};
  philox4x32_10(c, (uint32_t)seed, (uint32_t)(seed >> 32));
  const uint32_t r = c[idx & 3];
  const float u = (float)(r >> 8) * (1.0f / 16777216.0f);   // [0,1)
  return u >= p;
}

// the four keep decisions of the aligned element group idx0 .. idx0+3 (idx0 % 4 == 0): ONE Philox block, same values as
// philox_keep(idx0 + j)
__host__ __device__ __forceinline__ void philox_keep4(uint64_t seed, uint32_t sid, uint64_t idx0, float p, bool keep[4]) {
  uint32_t c[4] = {(uint32_t)(idx0 >> 2), (uint32_t)(idx0 >> 34), sid, 0x1a9fae5u};
  philox4x32_10(c, (uint32_t)seed, (uint32_t)(seed >> 32));
#pragma unroll
  for (int j = 0; j < 4; ++j) keep[j] = (float)(c[j] >> 8) * (1.0f / 16777216.0f) >= p;
}

struct DropSpec {  // resolved view of lagvae_dropout for one of the two masks
  int mode;        // 0 none, 1 mask, 2 philox
  float p, scale;  // scale = 1/(1-p)
  const uint8_t* mask;
  uint64_t seed;
  uint32_t sid;
  const uint64_t* seed_dev;   // optional device word added to `seed` at run time (graph replays), lagvae.h
};
__device__ __forceinline__ float drop_factor(const DropSpec& d, uint64_t idx) {
  if (d.mode == 0) return 1.0f;
  if (d.mode == 1) return d.mask[idx] ? d.scale : 0.0f;
  const uint64_t key = d.seed_dev ? d.seed + __ldg((const unsigned long long*)d.seed_dev) : d.seed;
  return philox_keep(key, d.sid, idx, d.p) ? d.scale : 0.0f;
}
// same for an aligned group of four elements (idx0 % 4 == 0)
__device__ __forceinline__ void drop_factor4(const DropSpec& d, uint64_t idx0, float f[4]) {
  if (d.mode == 0) {
    f[0] = f[1] = f[2] = f[3] = 1.0f;
  } else if (d.mode == 1) {
#pragma unroll
    for (int j = 0; j < 4; ++j) f[j] = d.mask[idx0 + j] ? d.scale : 0.0f;
  } else {
    const uint64_t key = d.seed_dev ? d.seed + __ldg((const unsigned long long*)d.seed_dev) : d.seed;
    bool k[4];
    philox_keep4(key, d.sid, idx0, d.p, k);
#pragma unroll
    for (int j = 0; j < 4; ++j) f[j] = k[j] ? d.scale : 0.0f;
  }
}

// ---- internal launchers shared between translation units --------------------------------------
int gemm_f32(const float* A, int64_t a_rs, int64_t a_cs, const float* B, int64_t b_rs, int64_t b_cs,
             float* C, int64_t ldc, int M, int N, int K, float alpha, float beta, const float* bias_n,
             const float* bias_rows, int bias_period, cudaStream_t st);

struct TcOperand {
  const uint16_t* hi;
  const uint16_t* lo;
  int64_t ld;
  int mn_major;
};
int gemm_tc(const TcOperand& A, const TcOperand& B, float* C, int64_t ldc, int M, int N, int K,
            int passes, float alpha, float beta, const float* bias_n, const float* bias_rows,
            int bias_period, const int32_t* out_row_map, cudaStream_t st);
void gemm_tc_set_grid_cap(int max_ctas);   // 0 = default (one CTA per SM); applies to this thread's next launches
int split_bf16_launch(const float* src, int64_t ld, int rows, int cols, uint16_t* hi, uint16_t* lo,
                      int64_t ld_out, cudaStream_t st);

}  // namespace lagvae
