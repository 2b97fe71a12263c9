// Im2col-free masked convolution on tcgen05 (sm_100a) for the PixelCNN stack of the image path
// (MaskedConv2d dec_pixelcnn_v2.py:12-30 inside PixelCNNBlock :32-62, 32 -> 32 channels,
// k = 7 / 5 / 3, stride 1, pad k/2, 28 x 28 images) — forward, data gradient and weight gradient.
//
// Operand format ("cat"): an NHWC activation [B, H, W, 32] fp32 is staged ONCE per layer as bf16 [B, H, W, 64] with
// channels [hi(32) | lo(32)], x = hi + lo + O(2^-17 |x|).  One pixel is then one 128-byte SWIZZLE_128B row, and a 4-D
// TMA box (64 ch, W, RPT rows, 1 image) whose W / H coordinates are offset by the tap (dy, dx) IS the shifted operand
// tile: out-of-image coordinates are zero-filled by the TMA unit, so padding, shifting and gathering cost no
// instructions and no HBM traffic beyond the activation itself (no patch matrix is ever materialised).
//
//   forward / dgrad (k_conv_tc): implicit GEMM, M = RPT*W output pixels (112 of a 128-row UMMA tile), N = 64, K = 64 per
//       live tap:  D[:, 0:32]  = [x_hi | x_lo] . [w_hi | w_hi]^T = x_hi w_hi + x_lo w_hi
//                  D[:, 32:64] = [x_hi | x_lo] . [w_lo | 0   ]^T = x_hi w_lo
//       (the three split-bf16 passes in ONE 128x64x16 instruction stream); only LIVE taps are multiplied (mask B:
//       25 / 13 / 5 of 49 / 25 / 9).  Warp-specialised: TMA producer, single-lane MMA issuer, 4 epilogue warps that add the
//       two halves, store fp32 NHWC rows and (optionally) accumulate the BatchNorm batch statistics of the output.
//   wgrad (k_conv_wgrad_tc): D_tap[64 x 64] += dycat^T . xcat(shifted by tap) over all pixels (K = pixels, both operands
//       MN-major straight from the NHWC rows); the four 32 x 32 quadrants are hi.hi, hi.lo, lo.hi, lo.lo and their sum is
//       dW[co, ci, tap].  ALL k*k taps are computed (the reference's autograd produces gradients for masked taps too,
//       SURVEY §7 quirk 6d).  Up to 8 taps share one TMEM allocation (two taps per N = 128 instruction); the pixel
//       dimension is split over CTAs and reduced deterministically by k_conv_wgrad_reduce.
#include "kernels.cuh"
#include "sm100_ptx.cuh"

#include <algorithm>

namespace lagvae {

namespace {

constexpr int CC = 32;          // channels of the masked convolutions
constexpr int MAX_TAPS = 49;
constexpr int NTHREADS = 192;   // warp 0 TMA, warp 1 MMA, warps 2-5 epilogue

__device__ __forceinline__ void tma_load_4d(uint32_t dst, const CUtensorMap* m, uint32_t bar, int32_t c0, int32_t c1,
                                            int32_t c2, int32_t c3) {
  asm volatile(
      "cp.async.bulk.tensor.4d.shared::cluster.global.mbarrier::complete_tx::bytes [%0], [%1, {%3, %4, %5, %6}], [%2];"
      ::"r"(dst), "l"((uint64_t)m), "r"(bar), "r"(c0), "r"(c1), "r"(c2), "r"(c3)
      : "memory");
}
// mbarrier wait that cannot hang the GPU: a pipeline bug traps after ~2 s instead of spinning forever
__device__ __forceinline__ void mbar_wait_b(uint32_t bar, uint32_t parity) {
  if (ptx::mbar_try_wait(bar, parity)) return;
  const long long t0 = clock64();
  while (!ptx::mbar_try_wait(bar, parity)) {
    if (clock64() - t0 > 4000000000LL) __trap();
  }
}
__device__ __forceinline__ void named_bar_sync(int id, int nthreads) {
  asm volatile("bar.sync %0, %1;" ::"r"(id), "r"(nthreads) : "memory");
}

struct TapList {
  int n;
  signed char dy[MAX_TAPS], dx[MAX_TAPS];
};

// ---------------------------------------------------------------------------------------------------------------
// forward / dgrad
// ---------------------------------------------------------------------------------------------------------------
constexpr int A_SLOT = 16384;                 // 128 rows x 128 B (RPT*W <= 128 rows are loaded)
constexpr int B_SLOT = 8192;                  // 64 rows x 128 B: [w_hi | w_hi] rows 0-31, [w_lo | 0] rows 32-63
constexpr int F_STAGE = A_SLOT + B_SLOT;
constexpr int F_STAGES = 6;
constexpr int CSTRIDE = 36;
constexpr int F_CST_BYTES = 4 * 32 * CSTRIDE * 4;
constexpr int F_SMEM = F_STAGES * F_STAGE + 1024 + 256 + F_CST_BYTES;
constexpr int F_TMEM = 128;                   // 2 accumulators x 64 columns

struct ConvArgs {
  float* Y;          // [B*H*W, 32]
  double* stats;     // [64]: sum(y) | sum(y^2) per output channel, or null
  int B, H, W, RPT;
  TapList taps;      // input pixel = output pixel + (dy, dx)
};

__global__ void __launch_bounds__(NTHREADS, 1)
k_conv_tc(const __grid_constant__ CUtensorMap tm_x, const __grid_constant__ CUtensorMap tm_w, const ConvArgs g) {
  extern __shared__ uint8_t smem_raw[];
  const uint32_t smem_base = (ptx::smem_u32(smem_raw) + 1023u) & ~1023u;
  const uint32_t bar_base = smem_base + F_STAGES * F_STAGE;
  auto full_bar = [&](int s) { return bar_base + 8u * s; };
  auto empty_bar = [&](int s) { return bar_base + 8u * (F_STAGES + s); };
  auto tfull_bar = [&](int b) { return bar_base + 8u * (2 * F_STAGES + b); };
  auto tempty_bar = [&](int b) { return bar_base + 8u * (2 * F_STAGES + 2 + b); };
  const uint32_t tmem_slot = bar_base + 8u * (2 * F_STAGES + 4);
  uint32_t* tmem_slot_ptr = (uint32_t*)(smem_raw + (tmem_slot - ptx::smem_u32(smem_raw)));

  const int warp = threadIdx.x >> 5, lane = threadIdx.x & 31;
  const int tpi = g.H / g.RPT;                 // tiles per image
  const int num_tiles = g.B * tpi;
  const int nrows = g.RPT * g.W;               // valid rows of a tile
  const int ntaps = g.taps.n;

  if (warp == 0 && lane == 0) {
    ptx::prefetch_tmap(&tm_x);
    ptx::prefetch_tmap(&tm_w);
    for (int s = 0; s < F_STAGES; ++s) {
      ptx::mbar_init(full_bar(s), 1);
      ptx::mbar_init(empty_bar(s), 1);
    }
    for (int b = 0; b < 2; ++b) {
      ptx::mbar_init(tfull_bar(b), 1);
      ptx::mbar_init(tempty_bar(b), 4);
    }
    ptx::fence_mbar_init();
  }
  if (warp == 1) ptx::tmem_alloc<F_TMEM>(tmem_slot);
  ptx::tc_fence_before();
  __syncthreads();
  ptx::tc_fence_after();
  const uint32_t tmem_base = *tmem_slot_ptr;

  if (warp == 0) {
    // ================================ TMA producer ================================
    int stage = 0;
    uint32_t phase = 0;
    const uint32_t tx_bytes = (uint32_t)(nrows * 128 + B_SLOT);
    for (int tile = blockIdx.x; tile < num_tiles; tile += gridDim.x) {
      const int b = tile / tpi, oy0 = (tile % tpi) * g.RPT;
      for (int t = 0; t < ntaps; ++t) {
        mbar_wait_b(empty_bar(stage), phase ^ 1u);
        const uint32_t sa = smem_base + stage * F_STAGE, sb = sa + A_SLOT;
        const int dy = g.taps.dy[t], dx = g.taps.dx[t];
        if (ptx::elect_one()) {
          ptx::mbar_expect_tx(full_bar(stage), tx_bytes);
          tma_load_4d(sa, &tm_x, full_bar(stage), 0, dx, oy0 + dy, b);   // shifted window, zero-filled outside the image
          ptx::tma_load_2d(sb, &tm_w, full_bar(stage), 0, t * 64);
        }
        __syncwarp();
        if (++stage == F_STAGES) { stage = 0; phase ^= 1u; }
      }
    }
  } else if (warp == 1) {
    // ================================ MMA issuer ================================
    constexpr uint32_t idesc = ptx::make_idesc_bf16_f32(128, 64, 0, 0);
    int stage = 0;
    uint32_t phase = 0;
    int acc = 0;
    uint32_t acc_phase = 0;
    for (int tile = blockIdx.x; tile < num_tiles; tile += gridDim.x) {
      mbar_wait_b(tempty_bar(acc), acc_phase ^ 1u);
      ptx::tc_fence_after();
      const uint32_t d_tmem = tmem_base + (uint32_t)(acc * 64);
      for (int t = 0; t < ntaps; ++t) {
        mbar_wait_b(full_bar(stage), phase);
        ptx::tc_fence_after();
        const uint32_t sa = smem_base + stage * F_STAGE, sb = sa + A_SLOT;
        if (ptx::elect_one()) {
#pragma unroll
          for (int k = 0; k < 4; ++k) {
            const uint64_t da = ptx::make_smem_desc_sw128(sa + k * 32u, 16u, 1024u);
            const uint64_t db = ptx::make_smem_desc_sw128(sb + k * 32u, 16u, 1024u);
            ptx::umma_f16(d_tmem, da, db, idesc, (t | k) ? 1u : 0u);
          }
          ptx::umma_commit(empty_bar(stage));
          if (t == ntaps - 1) ptx::umma_commit(tfull_bar(acc));
        }
        __syncwarp();
        if (++stage == F_STAGES) { stage = 0; phase ^= 1u; }
      }
      if (++acc == 2) { acc = 0; acc_phase ^= 1u; }
    }
  } else {
    // ================================ epilogue ================================
    const int quad = warp & 3;
    float* cst = (float*)(smem_raw + (bar_base + 256 - ptx::smem_u32(smem_raw))) + (warp - 2) * 32 * CSTRIDE;
    float s1 = 0.f, s2 = 0.f;   // BatchNorm statistics of channel `lane` over this warp's rows
    int acc = 0;
    uint32_t acc_phase = 0;
    for (int tile = blockIdx.x; tile < num_tiles; tile += gridDim.x) {
      mbar_wait_b(tfull_bar(acc), acc_phase);
      ptx::tc_fence_after();
      uint32_t r1[32], r2[32];
      const uint32_t t0 = tmem_base + ((uint32_t)(quad * 32) << 16) + (uint32_t)(acc * 64);
      ptx::tmem_ld32(t0, r1);
      ptx::tmem_ld32(t0 + 32u, r2);
      ptx::tmem_ld_wait();
      ptx::tc_fence_before();
      __syncwarp();
      if (lane == 0) ptx::mbar_arrive(tempty_bar(acc));   // accumulator is in registers: release it early
      float* mine = cst + lane * CSTRIDE;
#pragma unroll
      for (int j = 0; j < 32; j += 4)
        *(float4*)(mine + j) = make_float4(__uint_as_float(r1[j]) + __uint_as_float(r2[j]),
                                           __uint_as_float(r1[j + 1]) + __uint_as_float(r2[j + 1]),
                                           __uint_as_float(r1[j + 2]) + __uint_as_float(r2[j + 2]),
                                           __uint_as_float(r1[j + 3]) + __uint_as_float(r2[j + 3]));
      __syncwarp();
      const int64_t row0 = (int64_t)tile * nrows;
      const int cc = (lane & 7) * 4;
#pragma unroll
      for (int i = 0; i < 8; ++i) {
        const int rr = i * 4 + (lane >> 3), row = quad * 32 + rr;
        if (row < nrows) *(float4*)(g.Y + (row0 + row) * CC + cc) = *(const float4*)(cst + rr * CSTRIDE + cc);
      }
      if (g.stats) {
        const int lim = min(32, nrows - quad * 32);
        for (int rr = 0; rr < lim; ++rr) {
          const float v = cst[rr * CSTRIDE + lane];
          s1 += v;
          s2 = fmaf(v, v, s2);
        }
      }
      __syncwarp();
      if (++acc == 2) { acc = 0; acc_phase ^= 1u; }
    }
    if (g.stats) {
      atomicAdd(g.stats + lane, (double)s1);
      atomicAdd(g.stats + CC + lane, (double)s2);
    }
  }
  ptx::tc_fence_before();
  __syncthreads();
  if (warp == 1) {
    ptx::tc_fence_after();
    ptx::tmem_dealloc<F_TMEM>(tmem_base);
  }
}

// ---------------------------------------------------------------------------------------------------------------
// weight gradient
// ---------------------------------------------------------------------------------------------------------------
constexpr int W_ASLOT = 16384;                // dycat tile: RPT*W pixel rows x 128 B
constexpr int W_BSTAGE = 2 * 16384;           // two shifted xcat tiles (two taps -> one N = 128 instruction)
constexpr int W_NSB = 4;
constexpr int W_PART_BYTES = 64 * 33 * 4;
constexpr int W_SMEM = 2 * W_ASLOT + W_NSB * W_BSTAGE + 1024 + 256 + W_PART_BYTES;
constexpr int W_TMEM = 512;                   // 4 tap pairs x 128 columns
constexpr int W_TPG_MAX = 8;                  // taps per CTA group

struct WgradArgs {
  float* partial;    // [S][ntaps][32][32]
  int B, H, W, RPT;
  int tpg, S;        // taps per group, split of the pixel-tile dimension
  TapList taps;      // x pixel = dy pixel + (dy, dx), ALL k*k taps
};

__global__ void __launch_bounds__(NTHREADS, 1)
k_conv_wgrad_tc(const __grid_constant__ CUtensorMap tm_dy, const __grid_constant__ CUtensorMap tm_x, const WgradArgs g) {
  extern __shared__ uint8_t smem_raw[];
  const uint32_t smem_base = (ptx::smem_u32(smem_raw) + 1023u) & ~1023u;
  const uint32_t sB0 = smem_base + 2 * W_ASLOT;
  const uint32_t bar_base = sB0 + W_NSB * W_BSTAGE;
  auto a_full = [&](int s) { return bar_base + 8u * s; };
  auto a_empty = [&](int s) { return bar_base + 8u * (2 + s); };
  auto b_full = [&](int s) { return bar_base + 8u * (4 + s); };
  auto b_empty = [&](int s) { return bar_base + 8u * (4 + W_NSB + s); };
  const uint32_t acc_full = bar_base + 8u * (4 + 2 * W_NSB);
  const uint32_t tmem_slot = bar_base + 8u * (5 + 2 * W_NSB);
  uint32_t* tmem_slot_ptr = (uint32_t*)(smem_raw + (tmem_slot - ptx::smem_u32(smem_raw)));
  float* part = (float*)(smem_raw + (bar_base + 256 - ptx::smem_u32(smem_raw)));   // [64][33]

  const int warp = threadIdx.x >> 5, lane = threadIdx.x & 31;
  const int grp = blockIdx.x / g.S, s = blockIdx.x % g.S;
  const int t0 = grp * g.tpg;
  const int nt = min(g.tpg, g.taps.n - t0);
  const int npairs = (nt + 1) >> 1;
  const int tpi = g.H / g.RPT;
  const int nblocks = g.B * tpi;
  const int nrows = g.RPT * g.W;
  const int ksteps = nrows >> 4;
  const bool has_work = s < nblocks;

  if (warp == 0 && lane == 0) {
    ptx::prefetch_tmap(&tm_dy);
    ptx::prefetch_tmap(&tm_x);
    for (int i = 0; i < 2; ++i) {
      ptx::mbar_init(a_full(i), 1);
      ptx::mbar_init(a_empty(i), 1);
    }
    for (int i = 0; i < W_NSB; ++i) {
      ptx::mbar_init(b_full(i), 1);
      ptx::mbar_init(b_empty(i), 1);
    }
    ptx::mbar_init(acc_full, 1);
    ptx::fence_mbar_init();
  }
  if (warp == 1) ptx::tmem_alloc<W_TMEM>(tmem_slot);
  ptx::tc_fence_before();
  __syncthreads();
  ptx::tc_fence_after();
  const uint32_t tmem_base = *tmem_slot_ptr;

  if (warp == 0) {
    // ================================ TMA producer ================================
    int aslot = 0, stage = 0;
    uint32_t aphase = 0, phase = 0;
    const uint32_t a_bytes = (uint32_t)(nrows * 128);
    for (int blk = s; blk < nblocks; blk += g.S) {
      const int b = blk / tpi, oy0 = (blk % tpi) * g.RPT;
      mbar_wait_b(a_empty(aslot), aphase ^ 1u);
      if (ptx::elect_one()) {
        ptx::mbar_expect_tx(a_full(aslot), a_bytes);
        tma_load_4d(smem_base + aslot * W_ASLOT, &tm_dy, a_full(aslot), 0, 0, oy0, b);
      }
      __syncwarp();
      for (int p = 0; p < npairs; ++p) {
        const int ta = t0 + 2 * p, tb = min(ta + 1, t0 + nt - 1);   // an odd group repeats its last tap (result ignored)
        mbar_wait_b(b_empty(stage), phase ^ 1u);
        const uint32_t sb = sB0 + stage * W_BSTAGE;
        if (ptx::elect_one()) {
          ptx::mbar_expect_tx(b_full(stage), 2 * a_bytes);
          tma_load_4d(sb, &tm_x, b_full(stage), 0, g.taps.dx[ta], oy0 + g.taps.dy[ta], b);
          tma_load_4d(sb + 16384u, &tm_x, b_full(stage), 0, g.taps.dx[tb], oy0 + g.taps.dy[tb], b);
        }
        __syncwarp();
        if (++stage == W_NSB) { stage = 0; phase ^= 1u; }
      }
      aslot ^= 1;
      if (aslot == 0) aphase ^= 1u;
    }
  } else if (warp == 1) {
    // ================================ MMA issuer ================================
    // A = dycat tile, MN-major (M = 64 channels contiguous, K = pixel rows); B = two xcat tiles, MN-major, N = 128
    constexpr uint32_t idesc = ptx::make_idesc_bf16_f32(64, 128, 1, 1);
    int aslot = 0, stage = 0;
    uint32_t aphase = 0, phase = 0;
    int it = 0;
    for (int blk = s; blk < nblocks; blk += g.S, ++it) {
      const bool last_blk = blk + g.S >= nblocks;
      mbar_wait_b(a_full(aslot), aphase);
      ptx::tc_fence_after();
      const uint32_t sa = smem_base + aslot * W_ASLOT;
      for (int p = 0; p < npairs; ++p) {
        mbar_wait_b(b_full(stage), phase);
        ptx::tc_fence_after();
        const uint32_t sb = sB0 + stage * W_BSTAGE;
        if (ptx::elect_one()) {
          for (int k = 0; k < ksteps; ++k) {
            const uint64_t da = ptx::make_smem_desc_sw128(sa + k * 2048u, 8192u, 1024u);
            const uint64_t db = ptx::make_smem_desc_sw128(sb + k * 2048u, 16384u, 1024u);
            ptx::umma_f16(tmem_base + (uint32_t)(p * 128), da, db, idesc, (it | k) ? 1u : 0u);
          }
          ptx::umma_commit(b_empty(stage));
          if (p == npairs - 1) {
            ptx::umma_commit(a_empty(aslot));
            if (last_blk) ptx::umma_commit(acc_full);
          }
        }
        __syncwarp();
        if (++stage == W_NSB) { stage = 0; phase ^= 1u; }
      }
      aslot ^= 1;
      if (aslot == 0) aphase ^= 1u;
    }
  } else {
    // ================================ epilogue ================================
    // TMEM layout for M = 64: row i -> lane 32*(i/16) + i%16, i.e. this warp's lanes 0-15 hold rows 16*quad ... +15.
    const int quad = warp & 3;
    const int et = (warp - 2) * 32 + lane;    // 0..127
    if (has_work) {
      mbar_wait_b(acc_full, 0u);
      ptx::tc_fence_after();
    }
    for (int j = 0; j < nt; ++j) {
      const int tap = t0 + j;
      if (has_work) {
        uint32_t r1[32], r2[32];
        const uint32_t tb = tmem_base + ((uint32_t)(quad * 32) << 16) + (uint32_t)((j >> 1) * 128 + (j & 1) * 64);
        ptx::tmem_ld32(tb, r1);
        ptx::tmem_ld32(tb + 32u, r2);
        ptx::tmem_ld_wait();
        if (lane < 16) {
          float* dst = part + (quad * 16 + lane) * 33;
#pragma unroll
          for (int c = 0; c < 32; ++c) dst[c] = __uint_as_float(r1[c]) + __uint_as_float(r2[c]);
        }
      }
      named_bar_sync(1, 128);
      float* out = g.partial + ((int64_t)s * g.taps.n + tap) * (CC * CC);
#pragma unroll
      for (int q = 0; q < 8; ++q) {
        const int idx = et + 128 * q, co = idx >> 5, ci = idx & 31;
        out[idx] = has_work ? part[co * 33 + ci] + part[(32 + co) * 33 + ci] : 0.f;
      }
      named_bar_sync(1, 128);
    }
  }
  ptx::tc_fence_before();
  __syncthreads();
  if (warp == 1) {
    ptx::tc_fence_after();
    ptx::tmem_dealloc<W_TMEM>(tmem_base);
  }
}

// dW[co, ci, ty, tx] (torch layout [32, 32, kh, kw]) = sum over the S pixel partitions
__global__ void k_conv_wgrad_reduce(const float* __restrict__ partial, int S, int ntaps, float* __restrict__ dw) {
  const int i = blockIdx.x * blockDim.x + threadIdx.x;      // over (tap, co, ci)
  if (i >= ntaps * CC * CC) return;
  float a = 0.f;
  for (int s = 0; s < S; ++s) a += partial[(int64_t)s * ntaps * CC * CC + i];
  const int tap = i / (CC * CC), co = (i / CC) % CC, ci = i % CC;
  dw[((int64_t)co * CC + ci) * ntaps + tap] = a;
}

// ---------------------------------------------------------------------------------------------------------------
// operand staging
// ---------------------------------------------------------------------------------------------------------------
// fp32 [rows, 32] -> bf16 [rows, 64] = [hi | lo]; 4 threads per row, 8 channels each
__global__ void k_split_cat32(const float* __restrict__ x, int64_t rows, __nv_bfloat16* __restrict__ cat) {
  const int64_t n = rows * 4;
  for (int64_t i = blockIdx.x * (int64_t)blockDim.x + threadIdx.x; i < n; i += (int64_t)gridDim.x * blockDim.x) {
    const int64_t r = i >> 2;
    const int c0 = (int)(i & 3) * 8;
    const float4 a = *(const float4*)(x + r * CC + c0), b = *(const float4*)(x + r * CC + c0 + 4);
    const float v[8] = {a.x, a.y, a.z, a.w, b.x, b.y, b.z, b.w};
    __align__(16) __nv_bfloat16 hi[8], lo[8];
#pragma unroll
    for (int e = 0; e < 8; ++e) split_bf16(v[e], hi[e], lo[e]);
    *(uint4*)(cat + r * 64 + c0) = *(const uint4*)hi;
    *(uint4*)(cat + r * 64 + 32 + c0) = *(const uint4*)lo;
  }
}

// weight tiles.  w: torch layout [co = 32, ci = 32, kh, kw].  For live tap t = (ty, tx):
//   forward tile t [64 rows][64]: row n < 32: [hi(w[n, :, t]) | hi(w[n, :, t])], row 32 + n: [lo(w[n, :, t]) | 0]   (K = ci)
//   dgrad   tile t [64 rows][64]: row n < 32: [hi(w[:, n, t]) | hi(w[:, n, t])], row 32 + n: [lo(w[:, n, t]) | 0]   (K = co)
__global__ void k_conv_wprep(const float* __restrict__ w, int kh, int kw, TapList taps, int pad,
                             __nv_bfloat16* __restrict__ wf, __nv_bfloat16* __restrict__ wd) {
  const int i = blockIdx.x * blockDim.x + threadIdx.x;      // over (tap, row 0..63, col 0..63)
  if (i >= taps.n * 4096) return;
  const int t = i >> 12, row = (i >> 6) & 63, col = i & 63;
  const int ty = taps.dy[t] + pad, tx = taps.dx[t] + pad;
  const int n = row & 31, k = col & 31;
  const float vf = w[(((int64_t)n * CC + k) * kh + ty) * kw + tx];     // w[co = n, ci = k]
  const float vd = w[(((int64_t)k * CC + n) * kh + ty) * kw + tx];     // w[co = k, ci = n]
  __nv_bfloat16 hi, lo;
  split_bf16(vf, hi, lo);
  wf[i] = row < 32 ? hi : (col < 32 ? lo : __float2bfloat16_rn(0.f));
  split_bf16(vd, hi, lo);
  wd[i] = row < 32 ? hi : (col < 32 ? lo : __float2bfloat16_rn(0.f));
}

// ---------------------------------------------------------------------------------------------------------------
// host side
// ---------------------------------------------------------------------------------------------------------------
int make_tmap_nhwc64(CUtensorMap* out, const void* base, int B, int H, int W, int box_rows) {
  PFN_encodeTiled enc = get_encode_tiled();
  if (!enc) {
    set_error("cuTensorMapEncodeTiled entry point unavailable");
    return LAGVAE_E_CUDA;
  }
  LV_CHECK_ARG(((uintptr_t)base & 127) == 0, "conv tensor map: base must be 128-B aligned");
  cuuint64_t gdim[4] = {64, (cuuint64_t)W, (cuuint64_t)H, (cuuint64_t)B};
  cuuint64_t gstr[3] = {128, (cuuint64_t)W * 128, (cuuint64_t)H * W * 128};
  cuuint32_t box[4] = {64, (cuuint32_t)W, (cuuint32_t)box_rows, 1};
  cuuint32_t estr[4] = {1, 1, 1, 1};
  CUresult r = enc(out, CU_TENSOR_MAP_DATA_TYPE_BFLOAT16, 4, const_cast<void*>(base), gdim, gstr, box, estr,
                   CU_TENSOR_MAP_INTERLEAVE_NONE, CU_TENSOR_MAP_SWIZZLE_128B, CU_TENSOR_MAP_L2_PROMOTION_L2_256B,
                   CU_TENSOR_MAP_FLOAT_OOB_FILL_NONE);
  if (r != CUDA_SUCCESS) {
    set_error("cuTensorMapEncodeTiled (4-D NHWC) failed (%d) B=%d H=%d W=%d rows=%d", (int)r, B, H, W, box_rows);
    return LAGVAE_E_CUDA;
  }
  return LAGVAE_OK;
}

int sm_count() {
  static int n = 0;
  if (!n) {
    int dev = 0;
    cudaGetDevice(&dev);
    cudaDeviceGetAttribute(&n, cudaDevAttrMultiProcessorCount, dev);
    if (n <= 0) n = 148;
  }
  return n;
}

// rows of whole image lines per tile: the largest RPT with RPT*W <= 128, H % RPT == 0 and (RPT*W) % 16 == 0
int rows_per_tile(int H, int W) {
  for (int r = std::min(H, 128 / std::max(W, 1)); r >= 1; --r)
    if (H % r == 0 && (r * W) % 16 == 0) return r;
  return 0;
}

// mask_mode 0: plain convolution (all taps); 1: mask 'A' (rows above the centre + left of the centre);
// 2: mask 'B' (as 'A' plus the centre) — dec_pixelcnn_v2.py:16-20 with every input channel masked
TapList live_taps(int kh, int kw, int mask_mode, bool negate) {
  TapList t{};
  const int cy = kh / 2, cx = kw / 2;
  for (int ty = 0; ty < kh; ++ty)
    for (int tx = 0; tx < kw; ++tx) {
      bool live = true;
      if (mask_mode != 0) live = ty < cy || (ty == cy && (tx < cx || (mask_mode == 2 && tx == cx)));
      if (!live) continue;
      t.dy[t.n] = (signed char)(negate ? cy - ty : ty - cy);
      t.dx[t.n] = (signed char)(negate ? cx - tx : tx - cx);
      ++t.n;
    }
  return t;
}

bool conv32_ok(int B, int H, int W, int kh, int kw) {
  return B > 0 && H > 0 && W > 0 && W <= 128 && kh == kw && (kh & 1) && kh * kw <= MAX_TAPS && rows_per_tile(H, W) > 0;
}

template <typename K>
int set_smem(K kern, int bytes, bool* done) {
  if (!*done) {
    LV_CUDA(cudaFuncSetAttribute(kern, cudaFuncAttributeMaxDynamicSharedMemorySize, bytes));
    *done = true;
  }
  return LAGVAE_OK;
}

int conv32_run(const uint16_t* acat, const __nv_bfloat16* wtiles, int B, int H, int W, const TapList& taps, float* y,
               double* stats, cudaStream_t st) {
  const int rpt = rows_per_tile(H, W);
  CUtensorMap tm_x, tm_w;
  LV_TRY(make_tmap_nhwc64(&tm_x, acat, B, H, W, rpt));
  LV_TRY(make_tmap_bf16_2d(&tm_w, wtiles, (uint64_t)taps.n * 64, 64, 64, 64, 64));
  ConvArgs g{};
  g.Y = y; g.stats = stats; g.B = B; g.H = H; g.W = W; g.RPT = rpt; g.taps = taps;
  static bool configured = false;
  LV_TRY(set_smem(k_conv_tc, F_SMEM, &configured));
  if (stats) LV_CUDA(cudaMemsetAsync(stats, 0, 2 * CC * sizeof(double), st));
  const int tiles = B * (H / rpt);
  k_conv_tc<<<std::min(tiles, sm_count()), NTHREADS, F_SMEM, st>>>(tm_x, tm_w, g);
  LV_LAUNCH_CHECK();
  return LAGVAE_OK;
}

}  // namespace

}  // namespace lagvae

using namespace lagvae;

extern "C" {

int lagvae_conv32_supported(int B, int H, int W, int kh, int kw) { return conv32_ok(B, H, W, kh, kw) ? 1 : 0; }

int lagvae_split_cat32(const float* x, int64_t rows, uint16_t* cat, void* stream) {
  LV_CHECK_ARG(x && cat && rows > 0, "split_cat32: bad argument");
  const int64_t n = rows * 4;
  k_split_cat32<<<(int)std::min<int64_t>(cdiv(n, 256), 148 * 16), 256, 0, (cudaStream_t)stream>>>(x, rows, (__nv_bfloat16*)cat);
  LV_LAUNCH_CHECK();
  return LAGVAE_OK;
}

size_t lagvae_conv32_wbuf_bytes(int kh, int kw) { return (size_t)2 * kh * kw * 4096 * sizeof(uint16_t) + 256; }

int lagvae_conv32_prepare_weights(const float* w, int kh, int kw, int mask_mode, void* wbuf, void* stream) {
  LV_CHECK_ARG(w && wbuf && kh == kw && (kh & 1) && kh * kw <= MAX_TAPS && mask_mode >= 0 && mask_mode <= 2 &&
               ((uintptr_t)wbuf & 127) == 0, "conv32_prepare_weights: bad argument");
  const TapList taps = live_taps(kh, kw, mask_mode, false);
  __nv_bfloat16* wf = (__nv_bfloat16*)wbuf;
  __nv_bfloat16* wd = wf + (size_t)kh * kw * 4096;
  const int n = taps.n * 4096;
  k_conv_wprep<<<(int)cdiv(n, 256), 256, 0, (cudaStream_t)stream>>>(w, kh, kw, taps, kh / 2, wf, wd);
  LV_LAUNCH_CHECK();
  return LAGVAE_OK;
}

int lagvae_conv32_forward(const uint16_t* xcat, const void* wbuf, int B, int H, int W, int kh, int kw, int mask_mode,
                          float* y, double* stats_or_null, void* stream) {
  LV_CHECK_ARG(xcat && wbuf && y && conv32_ok(B, H, W, kh, kw) && mask_mode >= 0 && mask_mode <= 2, "conv32_forward: bad argument");
  return conv32_run(xcat, (const __nv_bfloat16*)wbuf, B, H, W, live_taps(kh, kw, mask_mode, false), y, stats_or_null,
                    (cudaStream_t)stream);
}

int lagvae_conv32_dgrad(const uint16_t* dycat, const void* wbuf, int B, int H, int W, int kh, int kw, int mask_mode,
                        float* dx, void* stream) {
  LV_CHECK_ARG(dycat && wbuf && dx && conv32_ok(B, H, W, kh, kw) && mask_mode >= 0 && mask_mode <= 2, "conv32_dgrad: bad argument");
  // dx[p] = sum_t dy[p - off_t] . w_t : the same implicit GEMM with negated offsets and the transposed weight tiles
  return conv32_run(dycat, (const __nv_bfloat16*)wbuf + (size_t)kh * kw * 4096, B, H, W, live_taps(kh, kw, mask_mode, true), dx,
                    nullptr, (cudaStream_t)stream);
}

static void wgrad_geometry(int ntaps, int* tpg, int* ngroups, int* S) {
  *ngroups = (int)cdiv(ntaps, W_TPG_MAX);
  *tpg = (int)cdiv(ntaps, *ngroups);
  *S = std::max(1, sm_count() / *ngroups);
}

size_t lagvae_conv32_wgrad_scratch_bytes(int kh, int kw) {
  int tpg, ng, S;
  wgrad_geometry(kh * kw, &tpg, &ng, &S);
  return (size_t)S * kh * kw * CC * CC * sizeof(float) + 256;
}

int lagvae_conv32_wgrad(const uint16_t* dycat, const uint16_t* xcat, int B, int H, int W, int kh, int kw, float* dw,
                        void* scratch, void* stream) {
  LV_CHECK_ARG(dycat && xcat && dw && scratch && conv32_ok(B, H, W, kh, kw), "conv32_wgrad: bad argument");
  cudaStream_t st = (cudaStream_t)stream;
  const int rpt = rows_per_tile(H, W);
  CUtensorMap tm_dy, tm_x;
  LV_TRY(make_tmap_nhwc64(&tm_dy, dycat, B, H, W, rpt));
  LV_TRY(make_tmap_nhwc64(&tm_x, xcat, B, H, W, rpt));
  WgradArgs g{};
  int ng;
  wgrad_geometry(kh * kw, &g.tpg, &ng, &g.S);
  g.S = std::min(g.S, B * (H / rpt));
  g.partial = (float*)scratch; g.B = B; g.H = H; g.W = W; g.RPT = rpt;
  g.taps = live_taps(kh, kw, 0, false);
  static bool configured = false;
  LV_TRY(set_smem(k_conv_wgrad_tc, W_SMEM, &configured));
  k_conv_wgrad_tc<<<ng * g.S, NTHREADS, W_SMEM, st>>>(tm_dy, tm_x, g);
  LV_LAUNCH_CHECK();
  const int n = kh * kw * CC * CC;
  k_conv_wgrad_reduce<<<(int)cdiv(n, 256), 256, 0, st>>>(g.partial, g.S, kh * kw, dw);
  LV_LAUNCH_CHECK();
  return LAGVAE_OK;
}

}  // extern "C"
