// Microbenchmark: cycles per tcgen05.mma (kind::f16, bf16, SS operands, SWIZZLE_128B K-major) as a function of
// (M, N), single CTA, back-to-back issue from one elected lane, accumulating into NACC rotating TMEM slots.
//   nvcc -gencode arch=compute_100a,code=sm_100a -O3 -I ../../vae-lagging-encoder_b200/csrc mma_cost.cu -o mma_cost
#include <cstdio>
#include <cuda_runtime.h>
#include "sm100_ptx.cuh"
using namespace lagvae;

template <int M, int N>
__global__ void __launch_bounds__(128, 1) k(int iters, int nacc, int same_ab, unsigned long long* out) {
  extern __shared__ uint8_t smem_raw[];
  const uint32_t base = (ptx::smem_u32(smem_raw) + 1023u) & ~1023u;
  const uint32_t a_base = base, b_base = base + 65536, bar = base + 65536 + 65536;
  uint32_t* slot = (uint32_t*)(smem_raw + (bar + 64 - ptx::smem_u32(smem_raw)));
  for (int i = threadIdx.x; i < 131072 / 4; i += blockDim.x) ((uint32_t*)(smem_raw + (base - ptx::smem_u32(smem_raw))))[i] = 0x3f803f80u;
  const int warp = threadIdx.x >> 5;
  if (threadIdx.x == 0) { ptx::mbar_init(bar, 1); ptx::fence_mbar_init(); }
  if (warp == 0) ptx::tmem_alloc<512>(bar + 64);
  ptx::fence_proxy_async_smem();
  ptx::tc_fence_before();
  __syncthreads();
  ptx::tc_fence_after();
  const uint32_t tmem = *slot;
  if (warp == 1) {
    constexpr uint32_t idesc = ptx::make_idesc_bf16_f32(M, N, 0, 0);
    unsigned long long t0 = 0, t1 = 0;
    for (int rep = 0; rep < 2; ++rep) {
      if (ptx::elect_one()) {
        t0 = clock64();
        for (int i = 0; i < iters; ++i) {
          const int kk = same_ab ? 0 : (i & 3);
          const int st = same_ab ? 0 : ((i >> 2) & 3);
          const uint64_t ad = ptx::make_smem_desc_sw128(a_base + st * 16384 + kk * 32, 16, 1024);
          const uint64_t bd = ptx::make_smem_desc_sw128(b_base + st * 16384 + kk * 32, 16, 1024);
          ptx::umma_f16(tmem + (uint32_t)((i % nacc) * N), ad, bd, idesc, 1u);
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

template <int M, int N>
void run(int nacc, int same_ab) {
  unsigned long long* d;
  cudaMalloc(&d, 8);
  const int smem = 131072 + 2048;
  cudaFuncSetAttribute(k<M, N>, cudaFuncAttributeMaxDynamicSharedMemorySize, smem);
  const int iters = 2048;
  k<M, N><<<1, 128, smem>>>(iters, nacc, same_ab, d);
  unsigned long long h = 0;
  cudaError_t e = cudaMemcpy(&h, d, 8, cudaMemcpyDeviceToHost);
  printf("M=%3d N=%3d nacc=%d %s : %7.1f cycles/MMA  (%6.0f MAC/clk)%s\n", M, N, nacc, same_ab ? "same-tile " : "4x4 tiles ",
         (double)h / iters, (double)M * N * 16 * iters / (double)h, e == cudaSuccess ? "" : cudaGetErrorString(e));
  cudaFree(d);
}

int main() {
  for (int nacc = 1; nacc <= 2; ++nacc) {
    run<64, 8>(nacc, 0); run<64, 16>(nacc, 0); run<64, 32>(nacc, 0); run<64, 64>(nacc, 0); run<64, 128>(nacc, 0); run<64, 256>(nacc, 0);
    run<128, 16>(nacc, 0); run<128, 32>(nacc, 0); run<128, 64>(nacc, 0); run<128, 128>(nacc, 0); run<128, 256>(nacc, 0);
  }
  run<64, 32>(1, 1); run<64, 256>(1, 1); run<128, 32>(1, 1); run<128, 256>(1, 1);
  return 0;
}
