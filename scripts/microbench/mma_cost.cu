// Microbenchmark: cycles per tcgen05.mma (kind::f16, bf16, SWIZZLE_128B K-major B) as a function of (M, N) and of the
// A operand source (SS: shared memory descriptor, TS: tensor memory), single CTA, back-to-back issue from one elected
// lane, accumulating into NACC rotating TMEM slots.  Also a known-answer check of the TS operand layout.
//   nvcc -gencode arch=compute_100a,code=sm_100a -O3 -I ../../vae-lagging-encoder_b200/csrc mma_cost.cu -o mma_cost
#include <cstdio>
#include <cstdlib>
#include <cuda_bf16.h>
#include <cuda_runtime.h>
#include "sm100_ptx.cuh"
using namespace lagvae;

constexpr int A_STAGE = 128 * 128;       // bytes per A stage (<= 128 rows x 64 k)
constexpr int NSTAGE = 4;

template <int M, int N, bool TS>
__global__ void __launch_bounds__(128, 1) k(int iters, int nacc, int same_ab, unsigned long long* out) {
  extern __shared__ uint8_t smem_raw[];
  constexpr int B_STAGE = N * 128;
  const uint32_t base = (ptx::smem_u32(smem_raw) + 1023u) & ~1023u;
  const uint32_t a_base = base, b_base = base + NSTAGE * A_STAGE, bar = b_base + NSTAGE * B_STAGE;
  uint32_t* slot = (uint32_t*)(smem_raw + (bar + 64 - ptx::smem_u32(smem_raw)));
  for (int i = threadIdx.x; i < (NSTAGE * (A_STAGE + B_STAGE)) / 4; i += blockDim.x)
    ((uint32_t*)(smem_raw + (base - ptx::smem_u32(smem_raw))))[i] = 0x3f803f80u;
  const int warp = threadIdx.x >> 5;
  if (threadIdx.x == 0) { ptx::mbar_init(bar, 1); ptx::fence_mbar_init(); }
  if (warp == 0) ptx::tmem_alloc<512>(bar + 64);
  ptx::fence_proxy_async_smem();
  ptx::tc_fence_before();
  __syncthreads();
  ptx::tc_fence_after();
  const uint32_t tmem = *slot;
  if (TS) {   // A operand region: columns [256, 384) = 16 K-slices of 8 columns
    uint32_t r[8];
    for (int j = 0; j < 8; ++j) r[j] = 0x3f803f80u;
    for (int c = 0; c < 128; c += 8) ptx::tmem_st8(tmem + ((uint32_t)(warp * 32) << 16) + 256u + c, r);
    ptx::tmem_st_wait();
    ptx::tc_fence_before();
    __syncthreads();
    ptx::tc_fence_after();
  }
  if (warp == 1) {
    constexpr uint32_t idesc = ptx::make_idesc_bf16_f32(M, N, 0, 0);
    unsigned long long t0 = 0, t1 = 0;
    for (int rep = 0; rep < 2; ++rep) {
      if (ptx::elect_one()) {
        t0 = clock64();
        for (int i = 0; i < iters; ++i) {
          const int kk = same_ab ? 0 : (i & 3);
          const int st = same_ab ? 0 : ((i >> 2) & 3);
          const uint64_t bd = ptx::make_smem_desc_sw128(b_base + st * B_STAGE + kk * 32, 16, 1024);
          const uint32_t d = tmem + (uint32_t)((i % nacc) * N);
          if (TS) {
            ptx::umma_f16_ts(d, tmem + 256u + (uint32_t)((st * 4 + kk) * 8), bd, idesc, 1u);
          } else {
            const uint64_t ad = ptx::make_smem_desc_sw128(a_base + st * A_STAGE + kk * 32, 16, 1024);
            ptx::umma_f16(d, ad, bd, idesc, 1u);
          }
        }
        ptx::umma_commit(bar);
      }
      __syncwarp();
      ptx::mbar_wait(bar, (uint32_t)rep);
      t1 = clock64();
    }
    if (threadIdx.x == 32) out[0] = t1 - t0;
  }
  ptx::tc_fence_before();
  __syncthreads();
  if (warp == 0) ptx::tmem_dealloc<512>(tmem);
}

template <int M, int N, bool TS>
void run(int nacc, int same_ab) {
  if (nacc * N > 256) return;
  unsigned long long* d;
  cudaMalloc(&d, 8);
  const int smem = NSTAGE * (A_STAGE + N * 128) + 2048;
  cudaFuncSetAttribute(k<M, N, TS>, cudaFuncAttributeMaxDynamicSharedMemorySize, smem);
  const int iters = 2048;
  k<M, N, TS><<<1, 128, smem>>>(iters, nacc, same_ab, d);
  unsigned long long h = 0;
  cudaError_t e = cudaMemcpy(&h, d, 8, cudaMemcpyDeviceToHost);
  printf("%s M=%3d N=%3d nacc=%d %s : %7.1f cycles/MMA  (%6.0f MAC/clk)%s\n", TS ? "TS" : "SS", M, N, nacc,
         same_ab ? "same-tile " : "4x4 tiles ", (double)h / iters, (double)M * N * 16 * iters / (double)h,
         e == cudaSuccess ? "" : cudaGetErrorString(e));
  fflush(stdout);
  if (e != cudaSuccess) exit(1);
  cudaFree(d);
}

// ---- known-answer check of the TS form: D[128 x 64] = A[128 x 64] * B[64 x 64]^T, A written to TMEM with tcgen05.st
__device__ __host__ inline float a_val(int m, int kk) { return (float)(((m * 7 + kk * 3) % 13) - 6); }
__device__ __host__ inline float b_val(int n, int kk) { return (float)(((n * 5 + kk) % 11) - 5); }

__global__ void __launch_bounds__(128, 1) k_check(float* out) {
  extern __shared__ uint8_t smem_raw[];
  const uint32_t base = (ptx::smem_u32(smem_raw) + 1023u) & ~1023u;
  uint8_t* gen = smem_raw + (base - ptx::smem_u32(smem_raw));
  const uint32_t b_base = base, bar = base + 8192;
  uint32_t* slot = (uint32_t*)(gen + 8192 + 64);
  const int warp = threadIdx.x >> 5, lane = threadIdx.x & 31;
  // B[n][k] K-major SWIZZLE_128B: row n = 128 bytes, 16-byte chunk c stored at chunk (c ^ (n & 7))
  for (int i = threadIdx.x; i < 64 * 64; i += 128) {
    const int n = i >> 6, kk = i & 63;
    const int off = n * 128 + (((kk >> 3) ^ (n & 7)) << 4) + (kk & 7) * 2;
    *(__nv_bfloat16*)(gen + off) = __float2bfloat16(b_val(n, kk));
  }
  if (threadIdx.x == 0) { ptx::mbar_init(bar, 1); ptx::fence_mbar_init(); }
  if (warp == 0) ptx::tmem_alloc<512>(bar + 64);
  ptx::fence_proxy_async_smem();
  ptx::tc_fence_before();
  __syncthreads();
  ptx::tc_fence_after();
  const uint32_t tmem = *slot;
  const int m = threadIdx.x;
  for (int c = 0; c < 32; c += 8) {     // 32 columns = 64 k
    uint32_t r[8];
    for (int j = 0; j < 8; ++j) {
      const int kk = 2 * (c + j);
      __nv_bfloat162 v = __floats2bfloat162_rn(a_val(m, kk), a_val(m, kk + 1));   // .x (low half) = even k
      r[j] = *(uint32_t*)&v;
    }
    ptx::tmem_st8(tmem + ((uint32_t)(warp * 32) << 16) + 256u + c, r);
  }
  ptx::tmem_st_wait();
  ptx::tc_fence_before();
  __syncthreads();
  ptx::tc_fence_after();
  if (warp == 1) {
    constexpr uint32_t idesc = ptx::make_idesc_bf16_f32(128, 64, 0, 0);
    if (ptx::elect_one()) {
      for (int kq = 0; kq < 4; ++kq)
        ptx::umma_f16_ts(tmem, tmem + 256u + kq * 8, ptx::make_smem_desc_sw128(b_base + kq * 32, 16, 1024), idesc, kq ? 1u : 0u);
      ptx::umma_commit(bar);
    }
    __syncwarp();
  }
  ptx::mbar_wait(bar, 0);
  ptx::tc_fence_after();
  for (int c = 0; c < 64; c += 32) {
    uint32_t r[32];
    ptx::tmem_ld32(tmem + ((uint32_t)(warp * 32) << 16) + c, r);
    ptx::tmem_ld_wait();
    for (int j = 0; j < 32; ++j) out[m * 64 + c + j] = __uint_as_float(r[j]);
  }
  (void)lane;
  ptx::tc_fence_before();
  __syncthreads();
  if (warp == 0) ptx::tmem_dealloc<512>(tmem);
}

static void check_ts() {
  float* d;
  cudaMalloc(&d, 128 * 64 * 4);
  cudaFuncSetAttribute(k_check, cudaFuncAttributeMaxDynamicSharedMemorySize, 16384);
  k_check<<<1, 128, 16384>>>(d);
  static float h[128 * 64];
  cudaError_t e = cudaMemcpy(h, d, sizeof(h), cudaMemcpyDeviceToHost);
  if (e != cudaSuccess) { printf("TS check: %s\n", cudaGetErrorString(e)); exit(1); }
  int bad = 0;
  for (int m = 0; m < 128; ++m)
    for (int n = 0; n < 64; ++n) {
      float ref = 0.f;
      for (int kk = 0; kk < 64; ++kk) ref += a_val(m, kk) * b_val(n, kk);
      if (ref != h[m * 64 + n] && bad++ < 5) printf("  mismatch D[%d][%d] = %g, expected %g\n", m, n, h[m * 64 + n], ref);
    }
  printf("TS known-answer check (M=128 N=64 K=64, A from TMEM lane=row, column=k/2): %s (%d mismatches)\n",
         bad ? "FAILED" : "exact", bad);
  fflush(stdout);
  cudaFree(d);
}

int main() {
  check_ts();
  for (int nacc = 1; nacc <= 2; ++nacc) {
    run<64, 8, false>(nacc, 0); run<64, 32, false>(nacc, 0); run<64, 64, false>(nacc, 0); run<64, 128, false>(nacc, 0);
    run<64, 256, false>(nacc, 0);
    run<128, 16, false>(nacc, 0); run<128, 32, false>(nacc, 0); run<128, 64, false>(nacc, 0); run<128, 128, false>(nacc, 0);
    run<128, 256, false>(nacc, 0);
    run<128, 16, true>(nacc, 0); run<128, 32, true>(nacc, 0); run<128, 64, true>(nacc, 0); run<128, 128, true>(nacc, 0);
    run<128, 256, true>(nacc, 0);
    run<64, 32, true>(nacc, 0); run<64, 64, true>(nacc, 0); run<64, 128, true>(nacc, 0);
  }
  run<64, 32, false>(1, 1); run<128, 64, false>(1, 1); run<128, 64, true>(1, 1);
  return 0;
}
