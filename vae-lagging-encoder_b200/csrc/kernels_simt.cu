// fp32 SIMT kernels of liblagvae: the element-wise / reduction / gather-scatter stages of the
// aggressive inner step, and a strided fp32 GEMM used for (a) shapes the tcgen05 path does not
// cover (toy nh=50, nz-sized contractions) and (b) as the on-device cross-check of gemm_tc.
// Reference call sites are cited per kernel (paths relative to the reference root).
#include "lagvae_common.cuh"
#include "kernels.cuh"

#include <stdarg.h>

namespace lagvae {

// ---------------------------------------------------------------------------------------------
// error plumbing
// ---------------------------------------------------------------------------------------------
static thread_local char g_err[1024] = "";
std::atomic<int64_t> g_launches{0};
void set_error(const char* fmt, ...) {
  va_list ap;
  va_start(ap, fmt);
  vsnprintf(g_err, sizeof(g_err), fmt, ap);
  va_end(ap);
}
const char* last_error() { return g_err; }

// ---------------------------------------------------------------------------------------------
// strided fp32 GEMM: C = alpha * A·Bᵀ + beta*C (+bias).  64x64x16 tiles, 4x4 register micro-tile.
// ---------------------------------------------------------------------------------------------
constexpr int GB_M = 64, GB_N = 64, GB_K = 16;

// SPLITK: blockIdx.z owns K range [z*kchunk, (z+1)*kchunk) and atomically adds alpha*partial into C
// (C pre-initialised to beta*C + bias by k_gemm_prep) — for skinny outputs with a long contraction.
template <bool SPLITK>
__global__ void __launch_bounds__(256)
k_gemm_f32(const float* __restrict__ A, int64_t a_rs, int64_t a_cs, const float* __restrict__ B,
           int64_t b_rs, int64_t b_cs, float* __restrict__ C, int64_t ldc, int M, int N, int K,
           float alpha, float beta, const float* __restrict__ bias_n,
           const float* __restrict__ bias_rows, int bias_period, int kchunk) {
  __shared__ float As[GB_K][GB_M + 4];
  __shared__ float Bs[GB_K][GB_N + 4];
  const int tid = threadIdx.x;
  const int m0 = blockIdx.y * GB_M, n0 = blockIdx.x * GB_N;
  const int tx = tid & 15, ty = tid >> 4;  // 16 x 16 threads, each 4x4 outputs
  float acc[4][4];
#pragma unroll
  for (int i = 0; i < 4; ++i)
#pragma unroll
    for (int j = 0; j < 4; ++j) acc[i][j] = 0.f;

  // loader mapping: pick the thread->element order that walks the contiguous dimension
  const bool a_kfast = (a_cs == 1), b_kfast = (b_cs == 1);
  const int kbeg = SPLITK ? blockIdx.z * kchunk : 0;
  const int kend = SPLITK ? min(K, kbeg + kchunk) : K;
  for (int k0 = kbeg; k0 < kend; k0 += GB_K) {
#pragma unroll
    for (int r = 0; r < 4; ++r) {
      const int e = tid + r * 256;  // 0..1023 = 64 x 16
      int mm, kk;
      if (a_kfast) { kk = e & 15; mm = e >> 4; } else { mm = e & 63; kk = e >> 6; }
      const int gm = m0 + mm, gk = k0 + kk;
      As[kk][mm] = (gm < M && gk < kend) ? A[(int64_t)gm * a_rs + (int64_t)gk * a_cs] : 0.f;
      int nn, kb;
      if (b_kfast) { kb = e & 15; nn = e >> 4; } else { nn = e & 63; kb = e >> 6; }
      const int gn = n0 + nn, gkb = k0 + kb;
      Bs[kb][nn] = (gn < N && gkb < kend) ? B[(int64_t)gn * b_rs + (int64_t)gkb * b_cs] : 0.f;
    }
    __syncthreads();
#pragma unroll
    for (int kk = 0; kk < GB_K; ++kk) {
      float a[4], b[4];
#pragma unroll
      for (int i = 0; i < 4; ++i) a[i] = As[kk][ty * 4 + i];
#pragma unroll
      for (int j = 0; j < 4; ++j) b[j] = Bs[kk][tx * 4 + j];
#pragma unroll
      for (int i = 0; i < 4; ++i)
#pragma unroll
        for (int j = 0; j < 4; ++j) acc[i][j] = fmaf(a[i], b[j], acc[i][j]);
    }
    __syncthreads();
  }
#pragma unroll
  for (int i = 0; i < 4; ++i) {
    const int gm = m0 + ty * 4 + i;
    if (gm >= M) continue;
#pragma unroll
    for (int j = 0; j < 4; ++j) {
      const int gn = n0 + tx * 4 + j;
      if (gn >= N) continue;
      float v = alpha * acc[i][j];
      float* c = C + (int64_t)gm * ldc + gn;
      if (SPLITK) {
        atomicAdd(c, v);
      } else {
        if (bias_n) v += bias_n[gn];
        if (bias_rows) v += bias_rows[(int64_t)(gm % bias_period) * N + gn];
        if (beta != 0.f) v += beta * (*c);
        *c = v;
      }
    }
  }
}

__global__ void k_gemm_prep(float* __restrict__ C, int64_t ldc, int M, int N, float beta,
                            const float* __restrict__ bias_n, const float* __restrict__ bias_rows, int bias_period) {
  const int i = blockIdx.x * blockDim.x + threadIdx.x;
  if (i >= M * N) return;
  const int m = i / N, n = i - m * N;
  float* c = C + (int64_t)m * ldc + n;
  float v = beta != 0.f ? beta * (*c) : 0.f;
  if (bias_n) v += bias_n[n];
  if (bias_rows) v += bias_rows[(int64_t)(m % bias_period) * N + n];
  *c = v;
}

int gemm_f32(const float* A, int64_t a_rs, int64_t a_cs, const float* B, int64_t b_rs, int64_t b_cs,
             float* C, int64_t ldc, int M, int N, int K, float alpha, float beta, const float* bias_n,
             const float* bias_rows, int bias_period, cudaStream_t st) {
  if (M <= 0 || N <= 0) return LAGVAE_OK;
  LV_CHECK_ARG(K >= 0 && (bias_rows == nullptr || bias_period > 0), "gemm_f32: bad K/bias_period");
  dim3 grid((unsigned)cdiv(N, GB_N), (unsigned)cdiv(M, GB_M));
  LV_CHECK_ARG(grid.y <= 65535, "gemm_f32: M too large (%d)", M);
  if ((int64_t)grid.x * grid.y <= 16 && K >= 512) {   // skinny output, long contraction: split K over the grid
    const int kchunk = 128;
    grid.z = (unsigned)cdiv(K, kchunk);
    k_gemm_prep<<<(int)cdiv((int64_t)M * N, 256), 256, 0, st>>>(C, ldc, M, N, beta, bias_n, bias_rows, bias_period);
    LV_LAUNCH_CHECK();
    k_gemm_f32<true><<<grid, 256, 0, st>>>(A, a_rs, a_cs, B, b_rs, b_cs, C, ldc, M, N, K, alpha, beta, bias_n,
                                            bias_rows, bias_period, kchunk);
    LV_LAUNCH_CHECK();
    return LAGVAE_OK;
  }
  k_gemm_f32<false><<<grid, 256, 0, st>>>(A, a_rs, a_cs, B, b_rs, b_cs, C, ldc, M, N, K, alpha, beta, bias_n,
                                           bias_rows, bias_period, 0);
  LV_LAUNCH_CHECK();
  return LAGVAE_OK;
}

// ---------------------------------------------------------------------------------------------
// fp32 -> bf16 hi/lo split (operand staging for gemm_tc)
// ---------------------------------------------------------------------------------------------
__device__ __forceinline__ uint2 pack_bf16x4(const __nv_bfloat16 (&v)[4]) {
  uint2 r;
  r.x = (uint32_t)__bfloat16_as_ushort(v[0]) | ((uint32_t)__bfloat16_as_ushort(v[1]) << 16);
  r.y = (uint32_t)__bfloat16_as_ushort(v[2]) | ((uint32_t)__bfloat16_as_ushort(v[3]) << 16);
  return r;
}
// 4 elements per thread: float4 load (when the source row is 16-B aligned), two 8-B stores
__global__ void k_split_bf16(const float* __restrict__ src, int64_t ld, int rows, int cols,
                             __nv_bfloat16* __restrict__ hi, __nv_bfloat16* __restrict__ lo,
                             int64_t ld_out) {
  const int64_t q_per_row = ld_out >> 2;   // ld_out % 8 == 0
  const int64_t n = (int64_t)rows * q_per_row;
  const bool vec = ((ld & 3) == 0) && ((((uintptr_t)src) & 15) == 0);
  for (int64_t i = blockIdx.x * (int64_t)blockDim.x + threadIdx.x; i < n;
       i += (int64_t)gridDim.x * blockDim.x) {
    const int64_t r = i / q_per_row;
    const int c = (int)(i - r * q_per_row) * 4;
    float v[4];
    if (vec && c + 3 < cols) {
      const float4 x = *(const float4*)(src + r * ld + c);
      v[0] = x.x; v[1] = x.y; v[2] = x.z; v[3] = x.w;
    } else {
#pragma unroll
      for (int j = 0; j < 4; ++j) v[j] = (c + j < cols) ? src[r * ld + c + j] : 0.f;
    }
    __nv_bfloat16 h[4], l[4];
#pragma unroll
    for (int j = 0; j < 4; ++j) split_bf16(v[j], h[j], l[j]);
    *(uint2*)(hi + r * ld_out + c) = pack_bf16x4(h);
    *(uint2*)(lo + r * ld_out + c) = pack_bf16x4(l);
  }
}
int split_bf16_launch(const float* src, int64_t ld, int rows, int cols, uint16_t* hi, uint16_t* lo,
                      int64_t ld_out, cudaStream_t st) {
  if (rows <= 0) return LAGVAE_OK;
  const int64_t n = (int64_t)rows * (ld_out / 4);
  const int blocks = (int)std::min<int64_t>(cdiv(n, 256), 148 * 16);
  k_split_bf16<<<blocks, 256, 0, st>>>(src, ld, rows, cols, (__nv_bfloat16*)hi, (__nv_bfloat16*)lo, ld_out);
  LV_LAUNCH_CHECK();
  return LAGVAE_OK;
}

// ---------------------------------------------------------------------------------------------
// embedding gather (+inverted dropout)  — enc_lstm.py:58; dec_lstm.py:80-81 (+expansion :87-91)
// out[(t*Bd + b*ns + s), j] = table[x[b, t_off+t], j] * keep((b*Tn + t)*ni + j)
// ---------------------------------------------------------------------------------------------
__global__ void k_embed_gather(const int64_t* __restrict__ x, int64_t x_ld, int t_off, int B, int ns,
                               int Tn, const float* __restrict__ table, int ni, DropSpec drop,
                               float* __restrict__ out) {
  const int Bd = B * ns;
  const int64_t n = (int64_t)Tn * Bd * ni;
  for (int64_t i = blockIdx.x * (int64_t)blockDim.x + threadIdx.x; i < n;
       i += (int64_t)gridDim.x * blockDim.x) {
    const int j = (int)(i % ni);
    const int64_t row = i / ni;
    const int bd = (int)(row % Bd), t = (int)(row / Bd);
    const int b = bd / ns;
    const int64_t tok = x[(int64_t)b * x_ld + t_off + t];
    const float f = drop_factor(drop, ((uint64_t)b * Tn + t) * ni + j);
    out[i] = table[tok * ni + j] * f;
  }
}
int embed_gather(const int64_t* x, int64_t x_ld, int t_off, int B, int ns, int Tn, const float* table,
                 int ni, DropSpec drop, float* out, cudaStream_t st) {
  const int64_t n = (int64_t)Tn * B * ns * ni;
  if (n <= 0) return LAGVAE_OK;
  const int blocks = (int)std::min<int64_t>(cdiv(n, 256), 148 * 16);
  k_embed_gather<<<blocks, 256, 0, st>>>(x, x_ld, t_off, B, ns, Tn, table, ni, drop, out);
  LV_LAUNCH_CHECK();
  return LAGVAE_OK;
}

// Same gather, four columns per thread, emitting the fp32 rows AND the bf16 hi/lo tensor-core operand [rows, ld_out] of the
// input projection in one pass (gather -> split -> GEMM were three passes over the tensor).  ni % 4 == 0,
// ld_out % 4 == 0, ld_out >= ni (pad columns zero).  One Philox block serves the four elements.
__global__ void __launch_bounds__(256)
k_embed_gather_split(const int64_t* __restrict__ x, int64_t x_ld, int t_off, int B, int ns, int Tn,
                     const float* __restrict__ table, int ni, DropSpec drop, float* __restrict__ out,
                     __nv_bfloat16* __restrict__ hi, __nv_bfloat16* __restrict__ lo, int ld_out) {
  const int Bd = B * ns;
  const int q_row = ld_out >> 2;                          // float4 groups per output row (incl. padding)
  const int64_t n = (int64_t)Tn * Bd * q_row;
  for (int64_t i = blockIdx.x * (int64_t)blockDim.x + threadIdx.x; i < n; i += (int64_t)gridDim.x * blockDim.x) {
    const int64_t row = i / q_row;
    const int j = (int)(i - row * q_row) * 4;
    float v[4] = {0.f, 0.f, 0.f, 0.f};
    if (j < ni) {
      const int bd = (int)(row % Bd), t = (int)(row / Bd);
      const int b = bd / ns;
      const int64_t tok = x[(int64_t)b * x_ld + t_off + t];
      const float4 e = *(const float4*)(table + tok * ni + j);
      float f[4];
      drop_factor4(drop, ((uint64_t)b * Tn + t) * ni + j, f);
      v[0] = e.x * f[0]; v[1] = e.y * f[1]; v[2] = e.z * f[2]; v[3] = e.w * f[3];
      *(float4*)(out + row * ni + j) = make_float4(v[0], v[1], v[2], v[3]);
    }
    __nv_bfloat16 a[4], c[4];
#pragma unroll
    for (int k = 0; k < 4; ++k) split_bf16(v[k], a[k], c[k]);
    *(uint2*)(hi + row * ld_out + j) = pack_bf16x4(a);
    *(uint2*)(lo + row * ld_out + j) = pack_bf16x4(c);
  }
}
int embed_gather_split(const int64_t* x, int64_t x_ld, int t_off, int B, int ns, int Tn, const float* table, int ni,
                       DropSpec drop, float* out, uint16_t* hi, uint16_t* lo, int64_t ld_out, cudaStream_t st) {
  const int64_t n = (int64_t)Tn * B * ns * (ld_out / 4);
  if (n <= 0) return LAGVAE_OK;
  const int blocks = (int)std::min<int64_t>(cdiv(n, 256), 148 * 16);
  k_embed_gather_split<<<blocks, 256, 0, st>>>(x, x_ld, t_off, B, ns, Tn, table, ni, drop, out, (__nv_bfloat16*)hi,
                                               (__nv_bfloat16*)lo, (int)ld_out);
  LV_LAUNCH_CHECK();
  return LAGVAE_OK;
}

// dense embedding gradient: dTable[x[b,t], :] += dX[row, :] * keep  (autograd of nn.Embedding,
// text.py:384; decoder row V-1 skipped: padding_idx=-1 -> V-1, dec_lstm.py:28)
__global__ void k_embed_scatter_add(const int64_t* __restrict__ x, int64_t x_ld, int t_off, int B,
                                    int ns, int Tn, const float* __restrict__ dX, int ni,
                                    DropSpec drop, float* __restrict__ dTable, int64_t skip_row) {
  const int Bd = B * ns;
  const int64_t n = (int64_t)Tn * Bd * ni;
  for (int64_t i = blockIdx.x * (int64_t)blockDim.x + threadIdx.x; i < n;
       i += (int64_t)gridDim.x * blockDim.x) {
    const int j = (int)(i % ni);
    const int64_t row = i / ni;
    const int bd = (int)(row % Bd), t = (int)(row / Bd);
    const int b = bd / ns;
    const int64_t tok = x[(int64_t)b * x_ld + t_off + t];
    if (tok == skip_row) continue;
    const float f = drop_factor(drop, ((uint64_t)b * Tn + t) * ni + j);
    if (f != 0.f) atomicAdd(dTable + tok * ni + j, dX[i] * f);
  }
}
int embed_scatter_add(const int64_t* x, int64_t x_ld, int t_off, int B, int ns, int Tn,
                      const float* dX, int ni, DropSpec drop, float* dTable, int64_t skip_row,
                      cudaStream_t st) {
  const int64_t n = (int64_t)Tn * B * ns * ni;
  if (n <= 0) return LAGVAE_OK;
  const int blocks = (int)std::min<int64_t>(cdiv(n, 256), 148 * 16);
  k_embed_scatter_add<<<blocks, 256, 0, st>>>(x, x_ld, t_off, B, ns, Tn, dX, ni, drop, dTable, skip_row);
  LV_LAUNCH_CHECK();
  return LAGVAE_OK;
}

// ---------------------------------------------------------------------------------------------
// LSTM cell, one time step (nn.LSTM semantics; enc_lstm.py:60, dec_lstm.py:104,106)
// gates_t [Bd,4nh]: pre-activations in, activated (i,f,g,o) out (stash for backward)
// ---------------------------------------------------------------------------------------------
__global__ void k_lstm_point_fwd(float* __restrict__ gates_t, const float* __restrict__ c_prev,
                                 float* __restrict__ c_t, float* __restrict__ h_t,
                                 float* __restrict__ hdrop_t, DropSpec drop, int t, int Tn, int Bd,
                                 int nh) {
  const int n = Bd * nh;
  for (int i = blockIdx.x * blockDim.x + threadIdx.x; i < n; i += gridDim.x * blockDim.x) {
    const int bd = i / nh, u = i - bd * nh;
    float* g = gates_t + (int64_t)bd * 4 * nh;
    const float ig = sigmoidf_(g[u]), fg = sigmoidf_(g[nh + u]), gg = tanhf(g[2 * nh + u]),
                og = sigmoidf_(g[3 * nh + u]);
    const float cp = c_prev ? c_prev[i] : 0.f;
    const float c = fg * cp + ig * gg;
    const float h = og * tanhf(c);
    g[u] = ig; g[nh + u] = fg; g[2 * nh + u] = gg; g[3 * nh + u] = og;
    c_t[i] = c;
    h_t[i] = h;
    if (hdrop_t) hdrop_t[i] = h * drop_factor(drop, ((uint64_t)bd * Tn + t) * nh + u);
  }
}
int lstm_point_fwd(float* gates_t, const float* c_prev, float* c_t, float* h_t, float* hdrop_t,
                   DropSpec drop, int t, int Tn, int Bd, int nh, cudaStream_t st) {
  const int n = Bd * nh;
  k_lstm_point_fwd<<<(int)std::min<int64_t>(cdiv(n, 256), 148 * 8), 256, 0, st>>>(
      gates_t, c_prev, c_t, h_t, hdrop_t, drop, t, Tn, Bd, nh);
  LV_LAUNCH_CHECK();
  return LAGVAE_OK;
}

// backward of one step: dh = dh_ext*keep + dh_rec; dc += dh*o*(1-tanh²c); gate grads; dc <- dc*f
__global__ void k_lstm_point_bwd(const float* __restrict__ gates_t, const float* __restrict__ c_t,
                                 const float* __restrict__ c_prev, const float* __restrict__ dh_ext_t,
                                 DropSpec drop, const float* __restrict__ dh_rec,
                                 float* __restrict__ dc, float* __restrict__ dgates_t, int t, int Tn,
                                 int Bd, int nh) {
  const int n = Bd * nh;
  for (int i = blockIdx.x * blockDim.x + threadIdx.x; i < n; i += gridDim.x * blockDim.x) {
    const int bd = i / nh, u = i - bd * nh;
    const float* g = gates_t + (int64_t)bd * 4 * nh;
    const float ig = g[u], fg = g[nh + u], gg = g[2 * nh + u], og = g[3 * nh + u];
    float dh = dh_rec ? dh_rec[i] : 0.f;
    if (dh_ext_t) dh += dh_ext_t[i] * drop_factor(drop, ((uint64_t)bd * Tn + t) * nh + u);
    const float tc = tanhf(c_t[i]);
    const float dct = dc[i] + dh * og * (1.f - tc * tc);
    const float cp = c_prev ? c_prev[i] : 0.f;
    float* dg = dgates_t + (int64_t)bd * 4 * nh;
    dg[u] = dct * gg * ig * (1.f - ig);
    dg[nh + u] = dct * cp * fg * (1.f - fg);
    dg[2 * nh + u] = dct * ig * (1.f - gg * gg);
    dg[3 * nh + u] = dh * tc * og * (1.f - og);
    dc[i] = dct * fg;
  }
}
int lstm_point_bwd(const float* gates_t, const float* c_t, const float* c_prev, const float* dh_ext_t,
                   DropSpec drop, const float* dh_rec, float* dc, float* dgates_t, int t, int Tn,
                   int Bd, int nh, cudaStream_t st) {
  const int n = Bd * nh;
  k_lstm_point_bwd<<<(int)std::min<int64_t>(cdiv(n, 256), 148 * 8), 256, 0, st>>>(
      gates_t, c_t, c_prev, dh_ext_t, drop, dh_rec, dc, dgates_t, t, Tn, Bd, nh);
  LV_LAUNCH_CHECK();
  return LAGVAE_OK;
}

// ---------------------------------------------------------------------------------------------
// encoder head + reparameterise + KL  — enc_lstm.py:62 (Linear nh->2nz, no bias, chunk),
// encoder.py:72-79 (z = mu + eps*exp(.5 logvar)), encoder.py:55 (KL).  One block per batch row.
// ---------------------------------------------------------------------------------------------
__global__ void __launch_bounds__(256)
k_head_reparam_kl(const float* __restrict__ h_last, const float* __restrict__ w_lin,
                  const float* __restrict__ eps, int nh, int nz, int ns, float* __restrict__ mu,
                  float* __restrict__ logvar, float* __restrict__ z, float* __restrict__ kl) {
  extern __shared__ float sm[];  // [nh] h row, [2nz] ml, [32] red
  float* hs = sm;
  float* ml = sm + nh;
  float* red = ml + 2 * nz;
  const int b = blockIdx.x, tid = threadIdx.x, lane = tid & 31, w = tid >> 5, nw = blockDim.x >> 5;
  for (int k = tid; k < nh; k += blockDim.x) hs[k] = h_last[(int64_t)b * nh + k];
  __syncthreads();
  for (int j = w; j < 2 * nz; j += nw) {
    const float* wr = w_lin + (int64_t)j * nh;
    float a = 0.f;
    for (int k = lane; k < nh; k += 32) a = fmaf(hs[k], wr[k], a);
    a = warp_sum(a);
    if (lane == 0) ml[j] = a;
  }
  __syncthreads();
  float part = 0.f;
  for (int j = tid; j < nz; j += blockDim.x) {
    const float m = ml[j], lv = ml[nz + j];
    if (mu) mu[(int64_t)b * nz + j] = m;
    if (logvar) logvar[(int64_t)b * nz + j] = lv;
    if (z) {
      const float sd = expf(0.5f * lv);
      for (int s = 0; s < ns; ++s) {
        const int64_t o = ((int64_t)b * ns + s) * nz + j;
        z[o] = m + eps[o] * sd;
      }
    }
    part += 0.5f * (m * m + expf(lv) - lv - 1.f);
  }
  part = block_sum(part, red);
  if (tid == 0 && kl) kl[b] = part;
}
int head_reparam_kl(const float* h_last, const float* w_lin, const float* eps, int B, int nh, int nz,
                    int ns, float* mu, float* logvar, float* z, float* kl, cudaStream_t st) {
  const size_t smem = (size_t)(nh + 2 * nz + 32) * sizeof(float);
  LV_CHECK_ARG(smem <= 48 * 1024, "head_reparam_kl: nh too large for smem (%d)", nh);
  k_head_reparam_kl<<<B, 256, smem, st>>>(h_last, w_lin, eps, nh, nz, ns, mu, logvar, z, kl);
  LV_LAUNCH_CHECK();
  return LAGVAE_OK;
}

// backward of reparam + KL: dml[b, 0:nz] = dmu, dml[b, nz:2nz] = dlogvar
//   dmu = Σ_s dz + g_kl*mu ;  dlogvar = Σ_s dz*eps*.5*std + g_kl*.5*(e^lv - 1)
__global__ void k_reparam_kl_bwd(const float* __restrict__ dz, const float* __restrict__ eps,
                                 const float* __restrict__ mu, const float* __restrict__ logvar,
                                 const float* __restrict__ g_kl, int B, int nz, int ns,
                                 float* __restrict__ dml) {
  const int n = B * nz;
  for (int i = blockIdx.x * blockDim.x + threadIdx.x; i < n; i += gridDim.x * blockDim.x) {
    const int b = i / nz, j = i - b * nz;
    const float lv = logvar[i], sd = expf(0.5f * lv), gk = g_kl ? g_kl[b] : 0.f;
    float sdz = 0.f, sdze = 0.f;
    for (int s = 0; s < ns; ++s) {
      const int64_t o = ((int64_t)b * ns + s) * nz + j;
      const float d = dz ? dz[o] : 0.f;
      sdz += d;
      sdze += d * eps[o];
    }
    dml[(int64_t)b * 2 * nz + j] = sdz + gk * mu[i];
    dml[(int64_t)b * 2 * nz + nz + j] = sdze * 0.5f * sd + gk * 0.5f * (expf(lv) - 1.f);
  }
}
int reparam_kl_bwd(const float* dz, const float* eps, const float* mu, const float* logvar,
                   const float* g_kl, int B, int nz, int ns, float* dml, cudaStream_t st) {
  k_reparam_kl_bwd<<<(int)cdiv(B * nz, 128), 128, 0, st>>>(dz, eps, mu, logvar, g_kl, B, nz, ns, dml);
  LV_LAUNCH_CHECK();
  return LAGVAE_OK;
}

// ---------------------------------------------------------------------------------------------
// small element-wise helpers
// ---------------------------------------------------------------------------------------------
// x[r, :] += bias[r % period, :]  (ncol % 4 == 0, 16-byte aligned): the time-invariant z contribution of the decoder input
// projection (dec_lstm.py:84,97) added in ONE streaming pass over the [T'*Bd, 4nh] pre-activations.  Doing it in the GEMM
// epilogue (per-row bias loads + a runtime modulo per store) made that GEMM 2.1x slower (235 vs 110 us, profiles/r2_ncu_full.md).
__global__ void __launch_bounds__(256) k_add_row_periodic(float* __restrict__ x, const float* __restrict__ bias, int64_t rows, int ncol,
                                                          int period) {
  const int64_t n4 = rows * (ncol >> 2);
  const int c4 = ncol >> 2;
  for (int64_t i = blockIdx.x * (int64_t)blockDim.x + threadIdx.x; i < n4; i += (int64_t)gridDim.x * blockDim.x) {
    const int64_t r = i / c4;
    const int c = (int)(i - r * c4);
    float4 v = ((float4*)x)[i];
    const float4 b = ((const float4*)bias)[(r % period) * c4 + c];
    v.x += b.x; v.y += b.y; v.z += b.z; v.w += b.w;
    ((float4*)x)[i] = v;
  }
}
int add_row_periodic(float* x, const float* bias, int64_t rows, int ncol, int period, cudaStream_t st) {
  LV_CHECK_ARG((ncol & 3) == 0 && ((uintptr_t)x & 15) == 0 && ((uintptr_t)bias & 15) == 0, "add_row_periodic: needs 16-B aligned rows");
  k_add_row_periodic<<<148 * 8, 256, 0, st>>>(x, bias, rows, ncol, period);
  LV_LAUNCH_CHECK();
  return LAGVAE_OK;
}
__global__ void k_vec_add(const float* a, const float* b, float* o, int n) {
  for (int i = blockIdx.x * blockDim.x + threadIdx.x; i < n; i += gridDim.x * blockDim.x) o[i] = a[i] + b[i];
}
int vec_add(const float* a, const float* b, float* o, int n, cudaStream_t st) {
  k_vec_add<<<(int)cdiv(n, 256), 256, 0, st>>>(a, b, o, n);
  LV_LAUNCH_CHECK();
  return LAGVAE_OK;
}
__global__ void k_tanh(const float* a, float* o, int n) {
  for (int i = blockIdx.x * blockDim.x + threadIdx.x; i < n; i += gridDim.x * blockDim.x) o[i] = tanhf(a[i]);
}
int tanh_copy(const float* a, float* o, int n, cudaStream_t st) {  // dec_lstm.py:101
  k_tanh<<<(int)cdiv(n, 256), 256, 0, st>>>(a, o, n);
  LV_LAUNCH_CHECK();
  return LAGVAE_OK;
}
// dc0_tot = dc_init + dh_init * (1 - h0^2)   (backward of h0 = tanh(c0), c0 shared; dec_lstm.py:100-101)
__global__ void k_dc0_total(const float* dc_init, const float* dh_init, const float* h0, float* o, int n) {
  for (int i = blockIdx.x * blockDim.x + threadIdx.x; i < n; i += gridDim.x * blockDim.x) {
    const float h = h0[i];
    o[i] = dc_init[i] + dh_init[i] * (1.f - h * h);
  }
}
int dc0_total(const float* dc_init, const float* dh_init, const float* h0, float* o, int n, cudaStream_t st) {
  k_dc0_total<<<(int)cdiv(n, 256), 256, 0, st>>>(dc_init, dh_init, h0, o, n);
  LV_LAUNCH_CHECK();
  return LAGVAE_OK;
}
// out[bd, j] = Σ_t src[(t*Bd + bd), j]    (d zb: the z-columns / bias see every time step)
__global__ void k_time_sum(const float* __restrict__ src, int Tn, int Bd, int ncol, float* __restrict__ out) {
  const int64_t n = (int64_t)Bd * ncol;
  for (int64_t i = blockIdx.x * (int64_t)blockDim.x + threadIdx.x; i < n; i += (int64_t)gridDim.x * blockDim.x) {
    float a0 = 0.f, a1 = 0.f, a2 = 0.f, a3 = 0.f;   // four independent chains: the loop is load-latency bound
    int t = 0;
    for (; t + 4 <= Tn; t += 4) {
      a0 += src[(int64_t)t * n + i];
      a1 += src[(int64_t)(t + 1) * n + i];
      a2 += src[(int64_t)(t + 2) * n + i];
      a3 += src[(int64_t)(t + 3) * n + i];
    }
    for (; t < Tn; ++t) a0 += src[(int64_t)t * n + i];
    out[i] = (a0 + a1) + (a2 + a3);
  }
}
// time-major logits [Tn*Bd, ld] -> batch-major [Bd, Tn, V] (the layout LSTMDecoder.decode returns, dec_lstm.py:109-111)
__global__ void __launch_bounds__(256) k_logits_batch_major(const float* __restrict__ src, int64_t ld, int V, int Tn, int Bd,
                                                            float* __restrict__ out) {
  const int64_t row = blockIdx.x;                 // output row = bd * Tn + t
  const int bd = (int)(row / Tn), t = (int)(row - (int64_t)bd * Tn);
  const float* s = src + ((int64_t)t * Bd + bd) * ld;
  float* o = out + row * V;
  for (int v = threadIdx.x; v < V; v += blockDim.x) o[v] = s[v];
}
int logits_batch_major(const float* src, int64_t ld, int V, int Tn, int Bd, float* out, cudaStream_t st) {
  k_logits_batch_major<<<(unsigned)((int64_t)Tn * Bd), 256, 0, st>>>(src, ld, V, Tn, Bd, out);
  LV_LAUNCH_CHECK();
  return LAGVAE_OK;
}
int time_sum(const float* src, int Tn, int Bd, int ncol, float* out, cudaStream_t st) {
  const int64_t n = (int64_t)Bd * ncol;
  k_time_sum<<<(int)std::min<int64_t>(cdiv(n, 128), 148 * 16), 128, 0, st>>>(src, Tn, Bd, ncol, out);
  LV_LAUNCH_CHECK();
  return LAGVAE_OK;
}
// out1[j] = out2[j] = Σ_r src[r, j]   (LSTM bias grads: b_ih and b_hh get identical gradients)
__global__ void k_col_sum(const float* __restrict__ src, int rows, int ncol, float* __restrict__ out1,
                          float* __restrict__ out2) {
  const int j = blockIdx.x * blockDim.x + threadIdx.x;
  if (j >= ncol) return;
  float a = 0.f;
  for (int r = 0; r < rows; ++r) a += src[(int64_t)r * ncol + j];
  out1[j] = a;
  if (out2) out2[j] = a;
}
int col_sum(const float* src, int rows, int ncol, float* out1, float* out2, cudaStream_t st) {
  k_col_sum<<<(int)cdiv(ncol, 128), 128, 0, st>>>(src, rows, ncol, out1, out2);
  LV_LAUNCH_CHECK();
  return LAGVAE_OK;
}

// ---------------------------------------------------------------------------------------------
// cross entropy over materialised logits — dec_lstm.py:47,143-148 (CrossEntropyLoss(weight=1,
// reduce=False) then sum over time).  One block per (t,bd) row.
// ---------------------------------------------------------------------------------------------
// online (single-pass) log-sum-exp: each thread keeps a running (max, sum), merged warp- then block-wide
__device__ __forceinline__ void lse_merge(float& m, float& s, float m2, float s2) {
  const float nm = fmaxf(m, m2);
  const float a = (m == -INFINITY) ? 0.f : s * expf(m - nm);
  const float b = (m2 == -INFINITY) ? 0.f : s2 * expf(m2 - nm);
  m = nm;
  s = a + b;
}
__global__ void __launch_bounds__(256)
k_ce_fwd(const float* __restrict__ logits, int64_t ld, int V, const int64_t* __restrict__ x,
         int64_t x_ld, int Bd, int ns, float* __restrict__ lse_out, float* __restrict__ loss_row) {
  __shared__ float red_m[8], red_s[8];
  const int row = blockIdx.x;
  const int bd = row % Bd, t = row / Bd;
  const float* l = logits + (int64_t)row * ld;
  float m = -INFINITY, s = 0.f;
  const bool vec = ((ld & 3) == 0) && ((((uintptr_t)logits) & 15) == 0);
  if (vec) {   // float4 loads (rows are 16-B aligned), exponentials against the running max of the 4-group
    const int V4 = V >> 2;
    const float4* l4 = (const float4*)l;
    for (int q = threadIdx.x; q < V4; q += blockDim.x) {
      const float4 x = l4[q];
      const float gm = fmaxf(fmaxf(x.x, x.y), fmaxf(x.z, x.w));
      if (gm > m) { s *= expf(m - gm); m = gm; }
      s += expf(x.x - m) + expf(x.y - m) + expf(x.z - m) + expf(x.w - m);
    }
    for (int v = (V4 << 2) + threadIdx.x; v < V; v += blockDim.x) {
      const float xv = l[v];
      if (xv > m) { s = s * expf(m - xv) + 1.f; m = xv; } else { s += expf(xv - m); }
    }
  } else {
    for (int v = threadIdx.x; v < V; v += blockDim.x) {
      const float xv = l[v];
      if (xv > m) { s = s * expf(m - xv) + 1.f; m = xv; } else { s += expf(xv - m); }
    }
  }
#pragma unroll
  for (int o = 16; o > 0; o >>= 1) {
    const float m2 = __shfl_xor_sync(0xffffffffu, m, o), s2 = __shfl_xor_sync(0xffffffffu, s, o);
    lse_merge(m, s, m2, s2);
  }
  const int lane = threadIdx.x & 31, w = threadIdx.x >> 5;
  if (lane == 0) { red_m[w] = m; red_s[w] = s; }
  __syncthreads();
  if (threadIdx.x == 0) {
    for (int i = 1; i < (int)(blockDim.x >> 5); ++i) lse_merge(m, s, red_m[i], red_s[i]);
    const float lse = m + logf(s);
    const int64_t tgt = x[(int64_t)(bd / ns) * x_ld + 1 + t];  // tgt = x[:,1:]  dec_lstm.py:127
    lse_out[row] = lse;
    loss_row[row] = lse - l[tgt];
  }
}
int ce_fwd(const float* logits, int64_t ld, int V, const int64_t* x, int64_t x_ld, int Tn, int Bd,
           int ns, float* lse_out, float* loss_row, cudaStream_t st) {
  k_ce_fwd<<<Tn * Bd, 256, 0, st>>>(logits, ld, V, x, x_ld, Bd, ns, lse_out, loss_row);
  LV_LAUNCH_CHECK();
  return LAGVAE_OK;
}
// in place: logits <- (softmax - onehot) * g_row,  g_row = g_rec[b] / ns
__global__ void __launch_bounds__(256)
k_ce_bwd(float* __restrict__ logits, int64_t ld, int V, const int64_t* __restrict__ x, int64_t x_ld,
         int Bd, int ns, const float* __restrict__ lse, const float* __restrict__ g_rec) {
  const int row = blockIdx.x;
  const int bd = row % Bd, t = row / Bd, b = bd / ns;
  const float g = g_rec[b] / (float)ns;
  const float L = lse[row];
  const int64_t tgt = x[(int64_t)b * x_ld + 1 + t];
  float* l = logits + (int64_t)row * ld;
  for (int v = threadIdx.x; v < V; v += blockDim.x) {
    float p = expf(l[v] - L);
    if (v == tgt) p -= 1.f;
    l[v] = p * g;
  }
}
int ce_bwd(float* logits, int64_t ld, int V, const int64_t* x, int64_t x_ld, int Tn, int Bd, int ns,
           const float* lse, const float* g_rec, cudaStream_t st) {
  k_ce_bwd<<<Tn * Bd, 256, 0, st>>>(logits, ld, V, x, x_ld, Bd, ns, lse, g_rec);
  LV_LAUNCH_CHECK();
  return LAGVAE_OK;
}

// same, but emits dlogits directly as the split-bf16 (hi, lo) GEMM operand [rows, ld_out] (pad columns = 0):
// saves the fp32 write + re-read of the 509 MB tensor
__global__ void __launch_bounds__(256)
k_ce_bwd_split(const float* __restrict__ logits, int64_t ld, int V, const int64_t* __restrict__ x, int64_t x_ld,
               int Bd, int ns, const float* __restrict__ lse, const float* __restrict__ g_rec,
               __nv_bfloat16* __restrict__ hi, __nv_bfloat16* __restrict__ lo, int64_t ld_out) {
  const int row = blockIdx.x;
  const int bd = row % Bd, t = row / Bd, b = bd / ns;
  const float g = g_rec[b] / (float)ns;
  const float L = lse[row];
  const int64_t tgt = x[(int64_t)b * x_ld + 1 + t];
  const float* l = logits + (int64_t)row * ld;
  __nv_bfloat16* h = hi + (int64_t)row * ld_out;
  __nv_bfloat16* w = lo + (int64_t)row * ld_out;
  const bool vec = ((ld & 3) == 0) && ((((uintptr_t)logits) & 15) == 0);
  for (int q = threadIdx.x; q < (int)(ld_out >> 2); q += blockDim.x) {
    const int v0 = q * 4;
    float xv[4];
    if (vec && v0 + 3 < V) {
      const float4 x4 = *(const float4*)(l + v0);
      xv[0] = x4.x; xv[1] = x4.y; xv[2] = x4.z; xv[3] = x4.w;
    } else {
#pragma unroll
      for (int j = 0; j < 4; ++j) xv[j] = (v0 + j < V) ? l[v0 + j] : -INFINITY;
    }
    __nv_bfloat16 a[4], c[4];
#pragma unroll
    for (int j = 0; j < 4; ++j) {
      float d = 0.f;
      if (v0 + j < V) {
        float p = expf(xv[j] - L);
        if (v0 + j == tgt) p -= 1.f;
        d = p * g;
      }
      split_bf16(d, a[j], c[j]);
    }
    *(uint2*)(h + v0) = pack_bf16x4(a);
    *(uint2*)(w + v0) = pack_bf16x4(c);
  }
}
int ce_bwd_split(const float* logits, int64_t ld, int V, const int64_t* x, int64_t x_ld, int Tn, int Bd, int ns,
                 const float* lse, const float* g_rec, uint16_t* hi, uint16_t* lo, int64_t ld_out, cudaStream_t st) {
  k_ce_bwd_split<<<Tn * Bd, 256, 0, st>>>(logits, ld, V, x, x_ld, Bd, ns, lse, g_rec, (__nv_bfloat16*)hi,
                                           (__nv_bfloat16*)lo, ld_out);
  LV_LAUNCH_CHECK();
  return LAGVAE_OK;
}

// Forward AND backward of the cross entropy in one pass over the logits, for callers that know the upstream gradient of
// every reconstruction row before the forward runs (the fused steps: text.py:382 / :413 take mean(dim=-1), so it is 1/B):
// the row (V fp32 = 80 KB at Yahoo's V) is held in registers between the log-sum-exp and the (softmax - onehot)·g
// emission, so the 509 MB logits tensor is read from HBM once instead of twice.  One 512-thread block per (t, bd) row,
// NV4 float4 per thread (V <= 2048·NV4).
template <int NV4>
__global__ void __launch_bounds__(512)
k_ce_fused(const float* __restrict__ logits, int64_t ld, int V, const int64_t* __restrict__ x, int64_t x_ld, int Bd, int ns,
           float g, float* __restrict__ lse_out, float* __restrict__ loss_row, __nv_bfloat16* __restrict__ hi,
           __nv_bfloat16* __restrict__ lo, int64_t ld_out) {
  __shared__ float red[16];
  __shared__ float bcast[2];
  const int row = blockIdx.x;
  const int bd = row % Bd, t = row / Bd, b = bd / ns;
  const float* l = logits + (int64_t)row * ld;
  const int lane = threadIdx.x & 31, w = threadIdx.x >> 5;
  float4 r[NV4];
#pragma unroll
  for (int i = 0; i < NV4; ++i) {
    const int v0 = (threadIdx.x + i * 512) * 4;
    if (v0 + 3 < V) {
      r[i] = *(const float4*)(l + v0);
    } else {
      r[i].x = (v0 + 0 < V) ? l[v0 + 0] : -INFINITY;
      r[i].y = (v0 + 1 < V) ? l[v0 + 1] : -INFINITY;
      r[i].z = (v0 + 2 < V) ? l[v0 + 2] : -INFINITY;
      r[i].w = -INFINITY;
    }
  }
  float m = -INFINITY;
#pragma unroll
  for (int i = 0; i < NV4; ++i) m = fmaxf(m, fmaxf(fmaxf(r[i].x, r[i].y), fmaxf(r[i].z, r[i].w)));
#pragma unroll
  for (int o = 16; o > 0; o >>= 1) m = fmaxf(m, __shfl_xor_sync(0xffffffffu, m, o));
  if (lane == 0) red[w] = m;
  __syncthreads();
  if (w == 0) {
    float v = (lane < 16) ? red[lane] : -INFINITY;
#pragma unroll
    for (int o = 8; o > 0; o >>= 1) v = fmaxf(v, __shfl_xor_sync(0xffffffffu, v, o));
    if (lane == 0) bcast[0] = v;
  }
  __syncthreads();
  m = bcast[0];
  float s = 0.f;
#pragma unroll
  for (int i = 0; i < NV4; ++i) s += (expf(r[i].x - m) + expf(r[i].y - m)) + (expf(r[i].z - m) + expf(r[i].w - m));
#pragma unroll
  for (int o = 16; o > 0; o >>= 1) s += __shfl_xor_sync(0xffffffffu, s, o);
  if (lane == 0) red[w] = s;
  __syncthreads();
  const int64_t tgt = x[(int64_t)b * x_ld + 1 + t];   // tgt = x[:,1:]  dec_lstm.py:127
  if (w == 0) {
    float v = (lane < 16) ? red[lane] : 0.f;
#pragma unroll
    for (int o = 8; o > 0; o >>= 1) v += __shfl_xor_sync(0xffffffffu, v, o);
    if (lane == 0) {
      const float L = m + logf(v);
      bcast[1] = L;
      lse_out[row] = L;
      loss_row[row] = L - l[tgt];
    }
  }
  __syncthreads();
  const float L = bcast[1];
  __nv_bfloat16* h = hi + (int64_t)row * ld_out;
  __nv_bfloat16* c = lo + (int64_t)row * ld_out;
#pragma unroll
  for (int i = 0; i < NV4; ++i) {
    const int v0 = (threadIdx.x + i * 512) * 4;
    if (v0 < ld_out) {
      const float xv[4] = {r[i].x, r[i].y, r[i].z, r[i].w};
      __nv_bfloat16 a[4], e[4];
#pragma unroll
      for (int j = 0; j < 4; ++j) {
        float d = 0.f;
        if (v0 + j < V) {
          float p = expf(xv[j] - L);
          if (v0 + j == tgt) p -= 1.f;
          d = p * g;
        }
        split_bf16(d, a[j], e[j]);
      }
      *(uint2*)(h + v0) = pack_bf16x4(a);
      *(uint2*)(c + v0) = pack_bf16x4(e);
    }
  }
}
// returns LAGVAE_OK and sets *done when the shape is supported (otherwise the caller keeps the two-kernel path)
int ce_fused(const float* logits, int64_t ld, int V, const int64_t* x, int64_t x_ld, int Tn, int Bd, int ns, float g_row,
             float* lse_out, float* loss_row, uint16_t* hi, uint16_t* lo, int64_t ld_out, bool* done, cudaStream_t st) {
  *done = false;
  const bool vec = ((ld & 3) == 0) && ((((uintptr_t)logits) & 15) == 0) && ((ld_out & 3) == 0);
  if (!vec || ld_out > 2048 * 10 || V < 1) return LAGVAE_OK;
  auto* h = (__nv_bfloat16*)hi;
  auto* c = (__nv_bfloat16*)lo;
  if (ld_out <= 2048 * 3)
    k_ce_fused<3><<<Tn * Bd, 512, 0, st>>>(logits, ld, V, x, x_ld, Bd, ns, g_row, lse_out, loss_row, h, c, ld_out);
  else if (ld_out <= 2048 * 6)
    k_ce_fused<6><<<Tn * Bd, 512, 0, st>>>(logits, ld, V, x, x_ld, Bd, ns, g_row, lse_out, loss_row, h, c, ld_out);
  else
    k_ce_fused<10><<<Tn * Bd, 512, 0, st>>>(logits, ld, V, x, x_ld, Bd, ns, g_row, lse_out, loss_row, h, c, ld_out);
  LV_LAUNCH_CHECK();
  *done = true;
  return LAGVAE_OK;
}

// Side-stream gate: returns once *flag != 0 (set by a persistent kernel of another stream when its whole grid is resident),
// consumes the flag, and gives up after max_cycles so that a launch that never sets it cannot hang the stream.
__global__ void k_wait_flag(unsigned* flag, long long max_cycles) {
  const long long t0 = clock64();
  while (*(volatile unsigned*)flag == 0u && clock64() - t0 < max_cycles) __nanosleep(500);
  *(volatile unsigned*)flag = 0u;
}
int wait_flag(unsigned* flag, long long max_cycles, cudaStream_t st) {
  k_wait_flag<<<1, 1, 0, st>>>(flag, max_cycles);
  LV_LAUNCH_CHECK();
  return LAGVAE_OK;
}

// rec[b] = mean_s Σ_t loss_row ; loss[b] = rec + klw*KL   (dec_lstm.py:148, vae.py:95,98);
// scalars[0..2] = Σloss, Σrec, ΣKL (text.py:381 reads Σloss).  Single block.
__global__ void __launch_bounds__(1024)
k_finalize_loss(const float* __restrict__ loss_row, const float* __restrict__ kl, int B, int ns, int Tn,
                float klw, float* __restrict__ loss, float* __restrict__ rec, float* __restrict__ kl_out,
                float* __restrict__ scalars) {
  __shared__ float red[32];
  const int Bd = B * ns;
  const int lane = threadIdx.x & 31, warp = threadIdx.x >> 5, nwarps = blockDim.x >> 5;
  float sl = 0.f, sr = 0.f, sk = 0.f;
  // one warp per sentence: the lanes stride over (sample, time) — fixed summation order, Tn/32 dependent loads instead of Tn
  for (int b = warp; b < B; b += nwarps) {
    float r = 0.f;
    for (int i = lane; i < ns * Tn; i += 32) {
      const int s_ = i / Tn, t = i - s_ * Tn;
      r += loss_row[(int64_t)t * Bd + b * ns + s_];
    }
    r = warp_sum(r) / (float)ns;
    if (lane == 0) {
      const float k = kl[b];
      const float l = r + klw * k;
      loss[b] = l;
      rec[b] = r;
      if (kl_out && kl_out != kl) kl_out[b] = k;
      sl += l; sr += r; sk += k;
    }
  }
  sl = block_sum(sl, red);
  sr = block_sum(sr, red);
  sk = block_sum(sk, red);
  if (threadIdx.x == 0 && scalars) { scalars[0] = sl; scalars[1] = sr; scalars[2] = sk; }
}
int finalize_loss(const float* loss_row, const float* kl, int B, int ns, int Tn, float klw, float* loss,
                  float* rec, float* kl_out, float* scalars, cudaStream_t st) {
  k_finalize_loss<<<1, 1024, 0, st>>>(loss_row, kl, B, ns, Tn, klw, loss, rec, kl_out, scalars);
  LV_LAUNCH_CHECK();
  return LAGVAE_OK;
}

// g_rec[b] = g_loss[b] + g_rec_in[b];  g_kl[b] = klw*g_loss[b] + g_kl_in[b]   (chain rule of vae.py:98)
__global__ void k_combine_upstream(const float* gl, const float* gr, const float* gk, float klw, int B,
                                   float* g_rec, float* g_kl) {
  const int b = blockIdx.x * blockDim.x + threadIdx.x;
  if (b >= B) return;
  const float l = gl ? gl[b] : 0.f;
  g_rec[b] = l + (gr ? gr[b] : 0.f);
  g_kl[b] = klw * l + (gk ? gk[b] : 0.f);
}
int combine_upstream(const float* gl, const float* gr, const float* gk, float klw, int B, float* g_rec,
                     float* g_kl, cudaStream_t st) {
  k_combine_upstream<<<(int)cdiv(B, 128), 128, 0, st>>>(gl, gr, gk, klw, B, g_rec, g_kl);
  LV_LAUNCH_CHECK();
  return LAGVAE_OK;
}
__global__ void k_fill(float* p, float v, int64_t n) {
  for (int64_t i = blockIdx.x * (int64_t)blockDim.x + threadIdx.x; i < n; i += (int64_t)gridDim.x * blockDim.x) p[i] = v;
}
int fill(float* p, float v, int64_t n, cudaStream_t st) {
  if (n <= 0) return LAGVAE_OK;
  k_fill<<<(int)std::min<int64_t>(cdiv(n, 256), 148 * 8), 256, 0, st>>>(p, v, n);
  LV_LAUNCH_CHECK();
  return LAGVAE_OK;
}

// ---------------------------------------------------------------------------------------------
// clip_grad_norm_ + SGD  — text.py:385,387 (torch.nn.utils.clip_grad_norm_, optim.SGD momentum 0)
// ---------------------------------------------------------------------------------------------
struct SegTable {
  float* p[16];
  float* g[16];
  int64_t n[16];
  int nseg, nupd;
};
// partial Σg² per block into scratch[blockIdx.x] (double), deterministic two-stage reduce.  Per-thread
// accumulation is fp32 in 4 independent lanes over <= ~1 K elements (B200 fp64 rate is far too low to use it
// in the streaming loop); the cross-thread / cross-block reduction is fp64.
__global__ void __launch_bounds__(256) k_sumsq(SegTable tb, double* __restrict__ partial) {
  __shared__ double red[256];
  double a = 0.0;
  for (int s = 0; s < tb.nseg; ++s) {
    const float* g = tb.g[s];
    const int64_t n = tb.n[s];
    float p0 = 0.f, p1 = 0.f, p2 = 0.f, p3 = 0.f;
    if ((((uintptr_t)g) & 15) == 0) {
      const int64_t n4 = n >> 2;
      const float4* g4 = (const float4*)g;
      for (int64_t i = blockIdx.x * (int64_t)blockDim.x + threadIdx.x; i < n4; i += (int64_t)gridDim.x * blockDim.x) {
        const float4 v = g4[i];
        p0 = fmaf(v.x, v.x, p0); p1 = fmaf(v.y, v.y, p1); p2 = fmaf(v.z, v.z, p2); p3 = fmaf(v.w, v.w, p3);
      }
      for (int64_t i = (n4 << 2) + blockIdx.x * (int64_t)blockDim.x + threadIdx.x; i < n; i += (int64_t)gridDim.x * blockDim.x)
        p0 = fmaf(g[i], g[i], p0);
    } else {
      for (int64_t i = blockIdx.x * (int64_t)blockDim.x + threadIdx.x; i < n; i += (int64_t)gridDim.x * blockDim.x)
        p0 = fmaf(g[i], g[i], p0);
    }
    a += (double)p0 + (double)p1 + (double)p2 + (double)p3;
  }
  red[threadIdx.x] = a;
  __syncthreads();
  for (int o = 128; o > 0; o >>= 1) {
    if (threadIdx.x < o) red[threadIdx.x] += red[threadIdx.x + o];
    __syncthreads();
  }
  if (threadIdx.x == 0) partial[blockIdx.x] = red[0];
}
__global__ void __launch_bounds__(256)
k_norm_finish(const double* __restrict__ partial, int nblk, float max_norm, float* out_norm, float* coef_out) {
  __shared__ double red[256];
  double a = 0.0;
  for (int i = threadIdx.x; i < nblk; i += 256) a += partial[i];
  red[threadIdx.x] = a;
  __syncthreads();
  for (int o = 128; o > 0; o >>= 1) {
    if (threadIdx.x < o) red[threadIdx.x] += red[threadIdx.x + o];
    __syncthreads();
  }
  if (threadIdx.x == 0) {
    const float nrm = (float)sqrt(red[0]);
    const float c = max_norm / (nrm + 1e-6f);   // clip_coef; clamped to 1.0
    if (out_norm) *out_norm = nrm;
    *coef_out = c < 1.f ? c : 1.f;
  }
}
__global__ void __launch_bounds__(256)
k_clip_sgd(SegTable tb, const float* __restrict__ coef_p, float lr, int scale_all) {
  const float coef = *coef_p;
  const int last = scale_all ? tb.nseg : tb.nupd;
  for (int s = 0; s < last; ++s) {
    float* g = tb.g[s];
    float* p = tb.p[s];
    const int64_t n = tb.n[s];
    const bool upd = s < tb.nupd;
    int64_t done = 0;
    if (((((uintptr_t)g) | ((uintptr_t)p)) & 15) == 0) {      // 16-byte vector body (pure HBM traffic: g r/w, p r/w)
      const int64_t n4 = n >> 2;
      float4* g4 = (float4*)g;
      float4* p4 = (float4*)p;
      for (int64_t i = blockIdx.x * (int64_t)blockDim.x + threadIdx.x; i < n4; i += (int64_t)gridDim.x * blockDim.x) {
        float4 gv = g4[i];
        gv.x *= coef; gv.y *= coef; gv.z *= coef; gv.w *= coef;   // clip_grad_norm_ multiplies even when coef == 1
        g4[i] = gv;
        if (upd) {
          float4 pv = p4[i];
          pv.x -= lr * gv.x; pv.y -= lr * gv.y; pv.z -= lr * gv.z; pv.w -= lr * gv.w;
          p4[i] = pv;
        }
      }
      done = n4 << 2;
    }
    for (int64_t i = done + blockIdx.x * (int64_t)blockDim.x + threadIdx.x; i < n; i += (int64_t)gridDim.x * blockDim.x) {
      const float gc = g[i] * coef;
      g[i] = gc;
      if (upd) p[i] = p[i] - lr * gc;
    }
  }
}
int clip_sgd_step(float* const* h_params, float* const* h_grads, const int64_t* h_counts, int n_seg,
                  int n_update, float max_norm, float lr, int scale_all, float* out_norm, void* scratch,
                  cudaStream_t st) {
  LV_CHECK_ARG(n_seg > 0 && n_seg <= 16 && n_update <= n_seg, "clip_sgd_step: n_seg must be 1..16");
  SegTable tb;
  tb.nseg = n_seg;
  tb.nupd = n_update;
  for (int i = 0; i < n_seg; ++i) {
    tb.p[i] = h_params ? h_params[i] : nullptr;
    tb.g[i] = h_grads[i];
    tb.n[i] = h_counts[i];
  }
  const int nblk = 148 * 8;  // 1184 partials * 8 B = 9472 B of scratch, coef at +12288
  double* partial = (double*)scratch;
  float* coef = (float*)((char*)scratch + 12288);
  k_sumsq<<<nblk, 256, 0, st>>>(tb, partial);
  LV_LAUNCH_CHECK();
  k_norm_finish<<<1, 256, 0, st>>>(partial, nblk, max_norm, out_norm, coef);
  LV_LAUNCH_CHECK();
  k_clip_sgd<<<148 * 4, 256, 0, st>>>(tb, coef, lr, scale_all);
  LV_LAUNCH_CHECK();
  return LAGVAE_OK;
}

// ---------------------------------------------------------------------------------------------
// MI estimate — encoder.py:111-145 (+ utils.py:3-16 log_sum_exp).  Single block; B x B pairs.
// ---------------------------------------------------------------------------------------------
__global__ void __launch_bounds__(256)
k_mi(const float* __restrict__ mu, const float* __restrict__ logvar, const float* __restrict__ eps,
     int B, int nz, float* __restrict__ out) {
  extern __shared__ float sm[];  // [B] log_qz , [32] red
  float* lq = sm;
  float* red = sm + B;
  const float LOG2PI = 1.8378770664093453f;
  // neg_entropy (encoder.py:125)
  float ne = 0.f;
  for (int b = threadIdx.x; b < B; b += blockDim.x) {
    float s = 0.f;
    for (int j = 0; j < nz; ++j) s += 1.f + logvar[(int64_t)b * nz + j];
    ne += -0.5f * nz * LOG2PI - 0.5f * s;
  }
  ne = block_sum(ne, red) / (float)B;
  // row a: logsumexp_b log_density[a,b] (encoder.py:134-143); one warp per row a
  const int lane = threadIdx.x & 31, w = threadIdx.x >> 5, nw = blockDim.x >> 5;
  for (int a = w; a < B; a += nw) {
    float m = -INFINITY, s = 0.f;  // online logsumexp over b (per lane), merged below
    for (int b = lane; b < B; b += 32) {
      float q = 0.f, slv = 0.f;
      for (int j = 0; j < nz; ++j) {
        const float lva = logvar[(int64_t)a * nz + j];
        const float za = mu[(int64_t)a * nz + j] + eps[(int64_t)a * nz + j] * expf(0.5f * lva);
        const float lvb = logvar[(int64_t)b * nz + j];
        const float d = za - mu[(int64_t)b * nz + j];
        q += d * d / expf(lvb);
        slv += lvb;
      }
      const float ld = -0.5f * q - 0.5f * (nz * LOG2PI + slv);
      const float nm = fmaxf(m, ld);
      s = s * expf(m - nm) + expf(ld - nm);
      m = nm;
    }
    const float gm = warp_max(m);
    float ss = (m == -INFINITY) ? 0.f : s * expf(m - gm);
    ss = warp_sum(ss);
    if (lane == 0) lq[a] = gm + logf(ss) - logf((float)B);
  }
  __syncthreads();
  float acc = 0.f;
  for (int a = threadIdx.x; a < B; a += blockDim.x) acc += lq[a];
  acc = block_sum(acc, red) / (float)B;
  if (threadIdx.x == 0) *out = ne - acc;
}
int mi_estimate(const float* mu, const float* logvar, const float* eps, int B, int nz, float* out,
                cudaStream_t st) {
  const size_t smem = (size_t)(B + 32) * sizeof(float);
  LV_CHECK_ARG(smem <= 48 * 1024, "mi_estimate: B too large (%d)", B);
  k_mi<<<1, 256, smem, st>>>(mu, logvar, eps, B, nz, out);
  LV_LAUNCH_CHECK();
  return LAGVAE_OK;
}

__global__ void k_dropout_mask(uint64_t seed, uint32_t sid, int64_t n, float p, uint8_t* out) {
  for (int64_t i = blockIdx.x * (int64_t)blockDim.x + threadIdx.x; i < n; i += (int64_t)gridDim.x * blockDim.x)
    out[i] = philox_keep(seed, sid, (uint64_t)i, p) ? 1 : 0;
}
int dropout_mask(uint64_t seed, uint32_t sid, int64_t n, float p, uint8_t* out, cudaStream_t st) {
  if (n <= 0) return LAGVAE_OK;
  k_dropout_mask<<<(int)std::min<int64_t>(cdiv(n, 256), 148 * 8), 256, 0, st>>>(seed, sid, n, p, out);
  LV_LAUNCH_CHECK();
  return LAGVAE_OK;
}

}  // namespace lagvae
