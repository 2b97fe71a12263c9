// Im2col-free convolutions on tcgen05 (sm_100a) for the PixelCNN stack of the image path (SURVEY §8 row a17:
// PixelCNNBlock dec_pixelcnn_v2.py:32-62 = 1x1 64->32, MaskedConv2d k x k 32->32 (:12-30), 1x1 32->64; head 1x1 64->64
// :145-147) — forward, data gradient and weight gradient; stride 1, pad k/2, channel counts 32 / 64.
//
// Operand format ("cat"): an NHWC activation [B, H, W, C] fp32 is staged ONCE per layer as bf16 [B, H, W, 2C] with
// channels [hi(C) | lo(C)], x = hi + lo + O(2^-17 |x|).  A 64-channel chunk of one pixel is one 128-byte SWIZZLE_128B
// row, and a 4-D TMA box (64 ch, W, RPT rows, 1 image) whose W / H coordinates are offset by the tap (dy, dx) IS the
// shifted operand tile: out-of-image coordinates are zero-filled by the TMA unit, so padding, shifting and gathering
// cost no instructions and no HBM traffic beyond the activation itself (no patch matrix is ever materialised).
//
//   forward / dgrad (k_conv_tc<COUT>): implicit GEMM, M = RPT*W output pixels (112 of a 128-row UMMA tile), N = 2 COUT,
//       K = 64 per item (item = live tap x 64-wide operand chunk).  C_in = 32 (one chunk [x_hi | x_lo] per tap):
//           D[:, 0:COUT]      += [x_hi | x_lo] . [w_hi | w_hi]^T = x_hi w_hi + x_lo w_hi
//           D[:, COUT:2 COUT] += [x_hi | x_lo] . [w_lo | 0   ]^T = x_hi w_lo
//       C_in = 64 (chunks x_hi, x_lo):  x_hi . [w_hi ; w_lo]^T  then  x_lo . [w_hi ; 0]^T  — the three split-bf16 passes in
//       ONE instruction stream; only LIVE taps are multiplied (mask B: 25 / 13 / 5 of 49 / 25 / 9).  Warp-specialised: TMA
//       producer, single-lane MMA issuer, 4 epilogue warps that add the two halves (+ an optional fp32 addend: the
//       residual-branch gradient), store fp32 NHWC rows and (optionally) accumulate the BatchNorm batch statistics of
//       the output.
//   wgrad (k_conv_wgrad_tc): D_tap[2 C_out x 2 C_in] += dycat^T . xcat(shifted by tap) over all pixels (K = pixels, both
//       operands MN-major straight from the NHWC rows); the four quadrants are hi.hi, hi.lo, lo.hi, lo.lo and their sum
//       is dW[co, ci, tap].  ALL k*k taps are computed (the reference's autograd produces gradients for masked taps too,
//       SURVEY §7 quirk 6d).  One N = 128 instruction covers two taps (C_in = 32) or the hi and lo chunk of one tap
//       (C_in = 64); up to 4 such units share the TMEM allocation; the pixel dimension is split over CTAs and reduced
//       deterministically (with the quadrant sum) by k_conv_wgrad_reduce.
#include "kernels.cuh"
#include "sm100_ptx.cuh"

#include <algorithm>

namespace lagvae {

namespace {

constexpr int MAX_TAPS = 49;
constexpr int MAX_ITEMS = 2 * MAX_TAPS;   // (tap, 64-wide operand chunk)
constexpr int NTHREADS = 192;             // warp 0 TMA, warp 1 MMA, warps 2-5 epilogue

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

struct TapList {   // input pixel = output pixel + (dy, dx)
  int n;
  signed char dy[MAX_TAPS], dx[MAX_TAPS];
};
struct ItemList {  // one K = 64 step of the implicit GEMM: tap shift + channel offset of the operand chunk
  int n;
  signed char dy[MAX_ITEMS], dx[MAX_ITEMS];
  unsigned char c0[MAX_ITEMS];
};

// ---------------------------------------------------------------------------------------------------------------
// forward / dgrad
// ---------------------------------------------------------------------------------------------------------------
constexpr int A_SLOT = 16384;                 // 128 rows x 128 B (RPT*W <= 128 rows are loaded)
constexpr int CSTRIDE = 36;
constexpr int F_CST_BYTES = 4 * 32 * CSTRIDE * 4;
template <int COUT>
struct FCfg {
  static constexpr int N = 2 * COUT;                      // accumulator columns: [hi-weight part | lo-weight part]
  static constexpr int B_SLOT = N * 128;                  // weight tile: N rows x 128 B
  static constexpr int STAGE = A_SLOT + B_SLOT;
  static constexpr int STAGES = COUT == 32 ? 6 : 5;
  static constexpr int TMEM = 2 * N;                      // double-buffered accumulator
  static constexpr int SMEM = STAGES * STAGE + 1024 + 256 + F_CST_BYTES;
};

struct ConvArgs {
  float* Y;              // [B*H*W, COUT]
  const float* addend;   // optional fp32 [B*H*W, COUT] added to the result
  double* stats;         // [2 COUT]: sum(y) | sum(y^2) per output channel (of the stored value), or null
  int B, H, W, RPT;
  ItemList items;
};

template <int COUT>
__global__ void __launch_bounds__(NTHREADS, 1)
k_conv_tc(const __grid_constant__ CUtensorMap tm_x, const __grid_constant__ CUtensorMap tm_w, const ConvArgs g) {
  using F = FCfg<COUT>;
  constexpr int STAGES = F::STAGES;
  extern __shared__ uint8_t smem_raw[];
  const uint32_t smem_base = (ptx::smem_u32(smem_raw) + 1023u) & ~1023u;
  const uint32_t bar_base = smem_base + STAGES * F::STAGE;
  auto full_bar = [&](int s) { return bar_base + 8u * s; };
  auto empty_bar = [&](int s) { return bar_base + 8u * (STAGES + s); };
  auto tfull_bar = [&](int b) { return bar_base + 8u * (2 * STAGES + b); };
  auto tempty_bar = [&](int b) { return bar_base + 8u * (2 * STAGES + 2 + b); };
  const uint32_t tmem_slot = bar_base + 8u * (2 * STAGES + 4);
  uint32_t* tmem_slot_ptr = (uint32_t*)(smem_raw + (tmem_slot - ptx::smem_u32(smem_raw)));

  const int warp = threadIdx.x >> 5, lane = threadIdx.x & 31;
  const int tpi = g.H / g.RPT;                 // tiles per image
  const int num_tiles = g.B * tpi;
  const int nrows = g.RPT * g.W;               // valid rows of a tile
  const int nitems = g.items.n;

  if (warp == 0 && lane == 0) {
    ptx::prefetch_tmap(&tm_x);
    ptx::prefetch_tmap(&tm_w);
    for (int s = 0; s < STAGES; ++s) {
      ptx::mbar_init(full_bar(s), 1);
      ptx::mbar_init(empty_bar(s), 1);
    }
    for (int b = 0; b < 2; ++b) {
      ptx::mbar_init(tfull_bar(b), 1);
      ptx::mbar_init(tempty_bar(b), 4);
    }
    ptx::fence_mbar_init();
  }
  if (warp == 1) ptx::tmem_alloc<F::TMEM>(tmem_slot);
  ptx::tc_fence_before();
  __syncthreads();
  ptx::tc_fence_after();
  const uint32_t tmem_base = *tmem_slot_ptr;

  if (warp == 0) {
    // ================================ TMA producer ================================
    int stage = 0;
    uint32_t phase = 0;
    const uint32_t tx_bytes = (uint32_t)(nrows * 128 + F::B_SLOT);
    for (int tile = blockIdx.x; tile < num_tiles; tile += gridDim.x) {
      const int b = tile / tpi, oy0 = (tile % tpi) * g.RPT;
      for (int t = 0; t < nitems; ++t) {
        mbar_wait_b(empty_bar(stage), phase ^ 1u);
        const uint32_t sa = smem_base + stage * F::STAGE, sb = sa + A_SLOT;
        const int dy = g.items.dy[t], dx = g.items.dx[t], c0 = g.items.c0[t];
        if (ptx::elect_one()) {
          ptx::mbar_expect_tx(full_bar(stage), tx_bytes);
          tma_load_4d(sa, &tm_x, full_bar(stage), c0, dx, oy0 + dy, b);   // shifted window, zero-filled outside the image
          ptx::tma_load_2d(sb, &tm_w, full_bar(stage), 0, t * F::N);
        }
        __syncwarp();
        if (++stage == STAGES) { stage = 0; phase ^= 1u; }
      }
    }
  } else if (warp == 1) {
    // ================================ MMA issuer ================================
    constexpr uint32_t idesc = ptx::make_idesc_bf16_f32(128, F::N, 0, 0);
    int stage = 0;
    uint32_t phase = 0;
    int acc = 0;
    uint32_t acc_phase = 0;
    for (int tile = blockIdx.x; tile < num_tiles; tile += gridDim.x) {
      mbar_wait_b(tempty_bar(acc), acc_phase ^ 1u);
      ptx::tc_fence_after();
      const uint32_t d_tmem = tmem_base + (uint32_t)(acc * F::N);
      for (int t = 0; t < nitems; ++t) {
        mbar_wait_b(full_bar(stage), phase);
        ptx::tc_fence_after();
        const uint32_t sa = smem_base + stage * F::STAGE, sb = sa + A_SLOT;
        if (ptx::elect_one()) {
#pragma unroll
          for (int k = 0; k < 4; ++k) {
            const uint64_t da = ptx::make_smem_desc_sw128(sa + k * 32u, 16u, 1024u);
            const uint64_t db = ptx::make_smem_desc_sw128(sb + k * 32u, 16u, 1024u);
            ptx::umma_f16(d_tmem, da, db, idesc, (t | k) ? 1u : 0u);
          }
          ptx::umma_commit(empty_bar(stage));
          if (t == nitems - 1) ptx::umma_commit(tfull_bar(acc));
        }
        __syncwarp();
        if (++stage == STAGES) { stage = 0; phase ^= 1u; }
      }
      if (++acc == 2) { acc = 0; acc_phase ^= 1u; }
    }
  } else {
    // ================================ epilogue ================================
    const int quad = warp & 3;
    float* cst = (float*)(smem_raw + (bar_base + 256 - ptx::smem_u32(smem_raw))) + (warp - 2) * 32 * CSTRIDE;
    constexpr int NG = COUT / 32;
    float s1[NG], s2[NG];   // BatchNorm statistics of channel 32*grp + lane over this warp's rows
#pragma unroll
    for (int q = 0; q < NG; ++q) s1[q] = s2[q] = 0.f;
    int acc = 0;
    uint32_t acc_phase = 0;
    for (int tile = blockIdx.x; tile < num_tiles; tile += gridDim.x) {
      mbar_wait_b(tfull_bar(acc), acc_phase);
      ptx::tc_fence_after();
      const int64_t row0 = (int64_t)tile * nrows;
      const int cc = (lane & 7) * 4;
#pragma unroll
      for (int grp = 0; grp < NG; ++grp) {
        uint32_t r1[32], r2[32];
        const uint32_t t0 = tmem_base + ((uint32_t)(quad * 32) << 16) + (uint32_t)(acc * F::N + grp * 32);
        ptx::tmem_ld32(t0, r1);
        ptx::tmem_ld32(t0 + (uint32_t)COUT, r2);
        ptx::tmem_ld_wait();
        if (grp == NG - 1) {   // accumulator is in registers: release it
          ptx::tc_fence_before();
          __syncwarp();
          if (lane == 0) ptx::mbar_arrive(tempty_bar(acc));
        }
        float* mine = cst + lane * CSTRIDE;
#pragma unroll
        for (int j = 0; j < 32; j += 4)
          *(float4*)(mine + j) = make_float4(__uint_as_float(r1[j]) + __uint_as_float(r2[j]),
                                             __uint_as_float(r1[j + 1]) + __uint_as_float(r2[j + 1]),
                                             __uint_as_float(r1[j + 2]) + __uint_as_float(r2[j + 2]),
                                             __uint_as_float(r1[j + 3]) + __uint_as_float(r2[j + 3]));
        __syncwarp();
        float4 ad[8];
        if (g.addend) {   // all eight loads in flight before the first store (Y and addend may alias as far as nvcc knows)
#pragma unroll
          for (int i = 0; i < 8; ++i) {
            const int rr = i * 4 + (lane >> 3), row = quad * 32 + rr;
            ad[i] = row < nrows ? __ldg((const float4*)(g.addend + (row0 + row) * COUT + grp * 32 + cc)) : make_float4(0.f, 0.f, 0.f, 0.f);
          }
        }
#pragma unroll
        for (int i = 0; i < 8; ++i) {
          const int rr = i * 4 + (lane >> 3), row = quad * 32 + rr;
          if (row < nrows) {
            float4 v = *(const float4*)(cst + rr * CSTRIDE + cc);
            const int64_t o = (row0 + row) * COUT + grp * 32 + cc;
            if (g.addend) {
              v.x += ad[i].x; v.y += ad[i].y; v.z += ad[i].z; v.w += ad[i].w;
              if (g.stats) *(float4*)(cst + rr * CSTRIDE + cc) = v;
            }
            *(float4*)(g.Y + o) = v;
          }
        }
        if (g.stats) {
          __syncwarp();
          const int lim = min(32, nrows - quad * 32);
          for (int rr = 0; rr < lim; ++rr) {
            const float v = cst[rr * CSTRIDE + lane];
            s1[grp] += v;
            s2[grp] = fmaf(v, v, s2[grp]);
          }
        }
        __syncwarp();
      }
      if (++acc == 2) { acc = 0; acc_phase ^= 1u; }
    }
    if (g.stats) {
#pragma unroll
      for (int grp = 0; grp < NG; ++grp) {
        atomicAdd(g.stats + grp * 32 + lane, (double)s1[grp]);
        atomicAdd(g.stats + COUT + grp * 32 + lane, (double)s2[grp]);
      }
    }
  }
  ptx::tc_fence_before();
  __syncthreads();
  if (warp == 1) {
    ptx::tc_fence_after();
    ptx::tmem_dealloc<F::TMEM>(tmem_base);
  }
}

// ---------------------------------------------------------------------------------------------------------------
// weight gradient
// ---------------------------------------------------------------------------------------------------------------
constexpr int W_ASLOT = 2 * 16384;            // dycat tile: 1 or 2 chunks of [RPT*W pixel rows x 128 B]
constexpr int W_BSTAGE = 2 * 16384;           // two xcat chunks (two taps, or hi + lo of one tap) -> one N = 128 instruction
constexpr int W_NSB = 4;
constexpr int W_SMEM = 2 * W_ASLOT + W_NSB * W_BSTAGE + 1024 + 256;
constexpr int W_TMEM = 512;                   // 4 units x 128 columns
constexpr int W_UNITS_MAX = 4;

struct WgradArgs {
  float* partial;    // [S][n_groups * upg][M][128] raw accumulators
  int B, H, W, RPT;
  int a_chunks;      // 64-wide chunks of dycat (C_out / 32): M = 64 * a_chunks
  int x_chunks;      // 64-wide chunks of xcat (C_in / 32)
  int tpg, upg, S;   // taps per group, units per group, split of the pixel-tile dimension
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

  const int warp = threadIdx.x >> 5, lane = threadIdx.x & 31;
  const int grp = blockIdx.x / g.S, s = blockIdx.x % g.S;
  const int t0 = grp * g.tpg;
  const int nt = min(g.tpg, g.taps.n - t0);
  const int nunits = g.x_chunks == 1 ? (nt + 1) >> 1 : nt;
  const int tpi = g.H / g.RPT;
  const int nblocks = g.B * tpi;
  const int nrows = g.RPT * g.W;
  const int ksteps = nrows >> 4;
  const bool has_work = s < nblocks;
  const int M = 64 * g.a_chunks;

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
    const uint32_t c_bytes = (uint32_t)(nrows * 128);
    for (int blk = s; blk < nblocks; blk += g.S) {
      const int b = blk / tpi, oy0 = (blk % tpi) * g.RPT;
      mbar_wait_b(a_empty(aslot), aphase ^ 1u);
      if (ptx::elect_one()) {
        ptx::mbar_expect_tx(a_full(aslot), c_bytes * (uint32_t)g.a_chunks);
        for (int c = 0; c < g.a_chunks; ++c)
          tma_load_4d(smem_base + aslot * W_ASLOT + c * 16384u, &tm_dy, a_full(aslot), c * 64, 0, oy0, b);
      }
      __syncwarp();
      for (int u = 0; u < nunits; ++u) {
        int ta, tb, ca, cb;   // (tap, channel offset) of the two chunks of this unit
        if (g.x_chunks == 1) {
          ta = t0 + 2 * u; tb = min(ta + 1, t0 + nt - 1);   // an odd group repeats its last tap (result ignored)
          ca = cb = 0;
        } else {
          ta = tb = t0 + u; ca = 0; cb = 64;
        }
        mbar_wait_b(b_empty(stage), phase ^ 1u);
        const uint32_t sb = sB0 + stage * W_BSTAGE;
        if (ptx::elect_one()) {
          ptx::mbar_expect_tx(b_full(stage), 2 * c_bytes);
          tma_load_4d(sb, &tm_x, b_full(stage), ca, g.taps.dx[ta], oy0 + g.taps.dy[ta], b);
          tma_load_4d(sb + 16384u, &tm_x, b_full(stage), cb, g.taps.dx[tb], oy0 + g.taps.dy[tb], b);
        }
        __syncwarp();
        if (++stage == W_NSB) { stage = 0; phase ^= 1u; }
      }
      aslot ^= 1;
      if (aslot == 0) aphase ^= 1u;
    }
  } else if (warp == 1) {
    // ================================ MMA issuer ================================
    // A = dycat tile, MN-major (64-channel chunks 16 KiB apart, K = pixel rows); B = two xcat chunks, MN-major, N = 128
    const uint32_t idesc = ptx::make_idesc_bf16_f32(M, 128, 1, 1);
    int aslot = 0, stage = 0;
    uint32_t aphase = 0, phase = 0;
    int it = 0;
    for (int blk = s; blk < nblocks; blk += g.S, ++it) {
      const bool last_blk = blk + g.S >= nblocks;
      mbar_wait_b(a_full(aslot), aphase);
      ptx::tc_fence_after();
      const uint32_t sa = smem_base + aslot * W_ASLOT;
      for (int u = 0; u < nunits; ++u) {
        mbar_wait_b(b_full(stage), phase);
        ptx::tc_fence_after();
        const uint32_t sb = sB0 + stage * W_BSTAGE;
        if (ptx::elect_one()) {
          for (int k = 0; k < ksteps; ++k) {
            const uint64_t da = ptx::make_smem_desc_sw128(sa + k * 2048u, 16384u, 1024u);
            const uint64_t db = ptx::make_smem_desc_sw128(sb + k * 2048u, 16384u, 1024u);
            ptx::umma_f16(tmem_base + (uint32_t)(u * 128), da, db, idesc, (it | k) ? 1u : 0u);
          }
          ptx::umma_commit(b_empty(stage));
          if (u == nunits - 1) {
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
    // ================================ epilogue: raw accumulators -> partial[s][unit][row][128] ================================
    // TMEM layouts: M = 128: row i -> lane i;  M = 64: row i -> lane 32*(i/16) + i%16 (lanes 16-31 of a quadrant unused)
    const int quad = warp & 3;
    int row;
    bool valid;
    if (M == 128) { row = quad * 32 + lane; valid = true; }
    else { row = quad * 16 + (lane & 15); valid = lane < 16; }
    if (has_work) {
      mbar_wait_b(acc_full, 0u);
      ptx::tc_fence_after();
    }
    const int total_units = (int)(gridDim.x / g.S) * g.upg;
    for (int u = 0; u < nunits; ++u) {
      float* out = g.partial + (((int64_t)s * total_units + grp * g.upg + u) * M + row) * 128;
#pragma unroll 1
      for (int c = 0; c < 4; ++c) {
        uint32_t r[32];
        if (has_work) {
          ptx::tmem_ld32(tmem_base + ((uint32_t)(quad * 32) << 16) + (uint32_t)(u * 128 + c * 32), r);
          ptx::tmem_ld_wait();
        } else {
#pragma unroll
          for (int j = 0; j < 32; ++j) r[j] = 0u;
        }
        if (valid) {
#pragma unroll
          for (int j = 0; j < 32; j += 4)
            *(float4*)(out + c * 32 + j) = make_float4(__uint_as_float(r[j]), __uint_as_float(r[j + 1]),
                                                       __uint_as_float(r[j + 2]), __uint_as_float(r[j + 3]));
        }
      }
    }
  }
  ptx::tc_fence_before();
  __syncthreads();
  if (warp == 1) {
    ptx::tc_fence_after();
    ptx::tmem_dealloc<W_TMEM>(tmem_base);
  }
}

// dW[co, ci, ty, tx] (torch layout [Cout, Cin, kh, kw]) = sum over the S pixel partitions of the four quadrants
// hi.hi + hi.lo + lo.hi + lo.lo of the raw accumulator of that tap.  32 outputs x 8 partition lanes per block: the S
// partitions are summed in a fixed order (lane-strided, then a fixed tree) -> deterministic.
__global__ void __launch_bounds__(256)
k_conv_wgrad_reduce(const float* __restrict__ partial, int S, int total_units, int M, int Cout, int Cin, int ntaps, int tpg,
                    int upg, float* __restrict__ dw) {
  __shared__ float red[8][33];
  const int o = threadIdx.x & 31, sl = threadIdx.x >> 5;
  const int i = blockIdx.x * 32 + o;                        // over (tap, co, ci)
  const bool ok = i < ntaps * Cout * Cin;
  float a = 0.f;
  int tap = 0, co = 0, ci = 0;
  if (ok) {
    tap = i / (Cout * Cin); co = (i / Cin) % Cout; ci = i % Cin;
    const int grp = tap / tpg, j = tap % tpg;
    int u, cb;
    if (Cin == 32) { u = grp * upg + (j >> 1); cb = (j & 1) * 64; }
    else { u = grp * upg + j; cb = 0; }
    for (int s = sl; s < S; s += 8) {
      const float* P = partial + ((int64_t)s * total_units + u) * M * 128;
      a += (P[co * 128 + cb + ci] + P[co * 128 + cb + Cin + ci]) + (P[(Cout + co) * 128 + cb + ci] + P[(Cout + co) * 128 + cb + Cin + ci]);
    }
  }
  red[sl][o] = a;
  __syncthreads();
  if (sl == 0 && ok) {
    float t = 0.f;
#pragma unroll
    for (int k = 0; k < 8; ++k) t += red[k][o];
    dw[((int64_t)co * Cin + ci) * ntaps + tap] = t;
  }
}

// ---------------------------------------------------------------------------------------------------------------
// operand staging
// ---------------------------------------------------------------------------------------------------------------
// fp32 [rows, C] -> bf16 [rows, 2C] = [hi | lo]; C/8 threads per row, 8 channels each
template <int C>
__global__ void k_split_cat(const float* __restrict__ x, int64_t rows, __nv_bfloat16* __restrict__ cat) {
  constexpr int TPR = C / 8;
  const int64_t n = rows * TPR;
  for (int64_t i = blockIdx.x * (int64_t)blockDim.x + threadIdx.x; i < n; i += (int64_t)gridDim.x * blockDim.x) {
    const int64_t r = i / TPR;
    const int c0 = (int)(i % TPR) * 8;
    const float4 a = *(const float4*)(x + r * C + c0), b = *(const float4*)(x + r * C + c0 + 4);
    const float v[8] = {a.x, a.y, a.z, a.w, b.x, b.y, b.z, b.w};
    __align__(16) __nv_bfloat16 hi[8], lo[8];
#pragma unroll
    for (int e = 0; e < 8; ++e) split_bf16(v[e], hi[e], lo[e]);
    *(uint4*)(cat + r * 2 * C + c0) = *(const uint4*)hi;
    *(uint4*)(cat + r * 2 * C + C + c0) = *(const uint4*)lo;
  }
}

// weight tiles for k_conv_tc.  w: torch layout [Cout_w, Cin_w, kh, kw].  Wm[n, k] = w[n, k, tap] (forward: N = Cout_w,
// K = Cin_w) or w[k, n, tap] (dgrad: N = Cin_w, K = Cout_w).  Per live tap, per 64-wide operand chunk, a tile of 2N rows:
//   K = 32 (one chunk [x_hi | x_lo]):  row n: [hi(Wm[n, :]) | hi(Wm[n, :])]      row N + n: [lo(Wm[n, :]) | 0]
//   K = 64, chunk 0 (x_hi):            row n: hi(Wm[n, :])                       row N + n: lo(Wm[n, :])
//           chunk 1 (x_lo):            row n: hi(Wm[n, :])                       row N + n: 0
__global__ void k_conv_wprep(const float* __restrict__ w, int Cout_w, int Cin_w, int kh, int kw, TapList taps, int pad,
                             int n_fwd, __nv_bfloat16* __restrict__ out_fwd, __nv_bfloat16* __restrict__ out_dgrad) {
  int i = blockIdx.x * blockDim.x + threadIdx.x;            // forward tiles first, then the dgrad tiles (one launch)
  const int dgrad = i >= n_fwd;
  if (dgrad) i -= n_fwd;
  const int N = dgrad ? Cin_w : Cout_w, K = dgrad ? Cout_w : Cin_w;
  const int chunks = K / 32;
  const int per_item = 2 * N * 64;                          // over (tap, chunk, row 0..2N-1, col 0..63)
  if (i >= taps.n * chunks * per_item) return;
  const int item = i / per_item, t = item / chunks, ch = item % chunks;
  const int row = (i % per_item) >> 6, col = i & 63;
  const int ty = taps.dy[t] + pad, tx = taps.dx[t] + pad;
  const int n = row % N, k = K == 32 ? (col & 31) : col;
  const int co = dgrad ? k : n, ci = dgrad ? n : k;
  const float v = w[(((int64_t)co * Cin_w + ci) * kh + ty) * kw + tx];
  __nv_bfloat16 hi, lo;
  split_bf16(v, hi, lo);
  const __nv_bfloat16 zero = __float2bfloat16_rn(0.f);
  __nv_bfloat16 o;
  if (K == 32) o = row < N ? hi : (col < 32 ? lo : zero);
  else o = row < N ? hi : (ch == 0 ? lo : zero);
  (dgrad ? out_dgrad : out_fwd)[i] = o;
}

// ---------------------------------------------------------------------------------------------------------------
// host side
// ---------------------------------------------------------------------------------------------------------------
// bf16 NHWC "cat" tensor [B, H, W, C2] viewed as a 4-D map; box = 64 channels x W x box_rows x 1 image
int make_tmap_nhwc_cat(CUtensorMap* out, const void* base, int B, int H, int W, int C2, int box_rows) {
  PFN_encodeTiled enc = get_encode_tiled();
  if (!enc) {
    set_error("cuTensorMapEncodeTiled entry point unavailable");
    return LAGVAE_E_CUDA;
  }
  LV_CHECK_ARG(((uintptr_t)base & 127) == 0, "conv tensor map: base must be 128-B aligned");
  const cuuint64_t pix = (cuuint64_t)C2 * 2;
  cuuint64_t gdim[4] = {(cuuint64_t)C2, (cuuint64_t)W, (cuuint64_t)H, (cuuint64_t)B};
  cuuint64_t gstr[3] = {pix, (cuuint64_t)W * pix, (cuuint64_t)H * W * pix};
  cuuint32_t box[4] = {64, (cuuint32_t)W, (cuuint32_t)box_rows, 1};
  cuuint32_t estr[4] = {1, 1, 1, 1};
  CUresult r = enc(out, CU_TENSOR_MAP_DATA_TYPE_BFLOAT16, 4, const_cast<void*>(base), gdim, gstr, box, estr,
                   CU_TENSOR_MAP_INTERLEAVE_NONE, CU_TENSOR_MAP_SWIZZLE_128B, CU_TENSOR_MAP_L2_PROMOTION_L2_256B,
                   CU_TENSOR_MAP_FLOAT_OOB_FILL_NONE);
  if (r != CUDA_SUCCESS) {
    set_error("cuTensorMapEncodeTiled (4-D NHWC) failed (%d) B=%d H=%d W=%d C2=%d rows=%d", (int)r, B, H, W, C2, box_rows);
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

bool chan_ok(int c) { return c == 32 || c == 64; }
bool convtc_ok(int B, int H, int W, int Cin, int Cout, int kh, int kw) {
  return B > 0 && H > 0 && W > 0 && W <= 128 && chan_ok(Cin) && chan_ok(Cout) && kh == kw && (kh & 1) && kh * kw <= MAX_TAPS &&
         rows_per_tile(H, W) > 0;
}

template <typename K>
int set_smem(K kern, int bytes, bool* done) {
  if (!*done) {
    LV_CUDA(cudaFuncSetAttribute(kern, cudaFuncAttributeMaxDynamicSharedMemorySize, bytes));
    *done = true;
  }
  return LAGVAE_OK;
}

// tiles of one direction inside wbuf: forward first, dgrad after kh*kw*(Cin/32) forward items
size_t fwd_tile_elems(int Cin, int Cout, int ntaps_all) { return (size_t)ntaps_all * (Cin / 32) * 2 * Cout * 64; }
size_t dgrad_tile_elems(int Cin, int Cout, int ntaps_all) { return (size_t)ntaps_all * (Cout / 32) * 2 * Cin * 64; }

// Y[B*H*W, Nout] = implicit GEMM of the Kin-channel cat tensor with the prepared tiles
int convtc_run(const uint16_t* acat, const __nv_bfloat16* wtiles, int B, int H, int W, int Kin, int Nout, const TapList& taps,
               const float* addend, float* y, double* stats, cudaStream_t st) {
  const int rpt = rows_per_tile(H, W);
  const int chunks = Kin / 32;
  CUtensorMap tm_x, tm_w;
  LV_TRY(make_tmap_nhwc_cat(&tm_x, acat, B, H, W, 2 * Kin, rpt));
  LV_TRY(make_tmap_bf16_2d(&tm_w, wtiles, (uint64_t)taps.n * chunks * 2 * Nout, 64, 64, 64, 2 * Nout));
  ConvArgs g{};
  g.Y = y; g.addend = addend; g.stats = stats; g.B = B; g.H = H; g.W = W; g.RPT = rpt;
  for (int t = 0; t < taps.n; ++t)
    for (int c = 0; c < chunks; ++c) {
      g.items.dy[g.items.n] = taps.dy[t];
      g.items.dx[g.items.n] = taps.dx[t];
      g.items.c0[g.items.n] = (unsigned char)(c * 64);
      ++g.items.n;
    }
  if (stats) LV_CUDA(cudaMemsetAsync(stats, 0, 2 * Nout * sizeof(double), st));
  const int tiles = B * (H / rpt);
  const int grid = std::min(tiles, sm_count());
  if (Nout == 32) {
    static bool configured = false;
    LV_TRY(set_smem(k_conv_tc<32>, FCfg<32>::SMEM, &configured));
    k_conv_tc<32><<<grid, NTHREADS, FCfg<32>::SMEM, st>>>(tm_x, tm_w, g);
  } else {
    static bool configured = false;
    LV_TRY(set_smem(k_conv_tc<64>, FCfg<64>::SMEM, &configured));
    k_conv_tc<64><<<grid, NTHREADS, FCfg<64>::SMEM, st>>>(tm_x, tm_w, g);
  }
  LV_LAUNCH_CHECK();
  return LAGVAE_OK;
}

void wgrad_geometry(int ntaps, int Cin, int* tpg, int* upg, int* ngroups, int* S) {
  const int taps_max = Cin == 32 ? 2 * W_UNITS_MAX : W_UNITS_MAX;
  *ngroups = (int)cdiv(ntaps, taps_max);
  *tpg = (int)cdiv(ntaps, *ngroups);
  *upg = Cin == 32 ? (*tpg + 1) / 2 : *tpg;
  *S = std::max(1, sm_count() / *ngroups);
}

}  // namespace

}  // namespace lagvae

using namespace lagvae;

extern "C" {

int lagvae_convtc_supported(int B, int H, int W, int Cin, int Cout, int kh, int kw) {
  return convtc_ok(B, H, W, Cin, Cout, kh, kw) ? 1 : 0;
}

int lagvae_split_cat(const float* x, int64_t rows, int C, uint16_t* cat, void* stream) {
  LV_CHECK_ARG(x && cat && rows > 0 && chan_ok(C) && ((uintptr_t)x & 15) == 0 && ((uintptr_t)cat & 15) == 0, "split_cat: bad argument");
  const int64_t n = rows * (C / 8);
  const int grid = (int)std::min<int64_t>(cdiv(n, 256), 148 * 16);
  if (C == 32) k_split_cat<32><<<grid, 256, 0, (cudaStream_t)stream>>>(x, rows, (__nv_bfloat16*)cat);
  else k_split_cat<64><<<grid, 256, 0, (cudaStream_t)stream>>>(x, rows, (__nv_bfloat16*)cat);
  LV_LAUNCH_CHECK();
  return LAGVAE_OK;
}

size_t lagvae_convtc_wbuf_bytes(int Cin, int Cout, int kh, int kw) {
  return (fwd_tile_elems(Cin, Cout, kh * kw) + dgrad_tile_elems(Cin, Cout, kh * kw)) * sizeof(uint16_t) + 256;
}

int lagvae_convtc_prepare_weights(const float* w, int Cout, int Cin, int kh, int kw, int mask_mode, void* wbuf, void* stream) {
  LV_CHECK_ARG(w && wbuf && chan_ok(Cin) && chan_ok(Cout) && kh == kw && (kh & 1) && kh * kw <= MAX_TAPS && mask_mode >= 0 &&
               mask_mode <= 2 && ((uintptr_t)wbuf & 127) == 0, "convtc_prepare_weights: bad argument");
  const TapList taps = live_taps(kh, kw, mask_mode, false);
  __nv_bfloat16* wf = (__nv_bfloat16*)wbuf;
  __nv_bfloat16* wd = wf + fwd_tile_elems(Cin, Cout, kh * kw);
  const int nf = taps.n * (Cin / 32) * 2 * Cout * 64, nd = taps.n * (Cout / 32) * 2 * Cin * 64;
  k_conv_wprep<<<(int)cdiv(nf + nd, 256), 256, 0, (cudaStream_t)stream>>>(w, Cout, Cin, kh, kw, taps, kh / 2, nf, wf, wd);
  LV_LAUNCH_CHECK();
  return LAGVAE_OK;
}

int lagvae_convtc_forward(const uint16_t* xcat, const void* wbuf, int B, int H, int W, int Cin, int Cout, int kh, int kw,
                          int mask_mode, const float* addend_or_null, float* y, double* stats_or_null, void* stream) {
  LV_CHECK_ARG(xcat && wbuf && y && convtc_ok(B, H, W, Cin, Cout, kh, kw) && mask_mode >= 0 && mask_mode <= 2, "convtc_forward: bad argument");
  return convtc_run(xcat, (const __nv_bfloat16*)wbuf, B, H, W, Cin, Cout, live_taps(kh, kw, mask_mode, false), addend_or_null, y,
                    stats_or_null, (cudaStream_t)stream);
}

int lagvae_convtc_dgrad(const uint16_t* dycat, const void* wbuf, int B, int H, int W, int Cin, int Cout, int kh, int kw,
                        int mask_mode, const float* addend_or_null, float* dx, void* stream) {
  LV_CHECK_ARG(dycat && wbuf && dx && convtc_ok(B, H, W, Cin, Cout, kh, kw) && mask_mode >= 0 && mask_mode <= 2, "convtc_dgrad: bad argument");
  // dx[p] = sum_t dy[p - off_t] . w_t : the same implicit GEMM with negated offsets and the transposed weight tiles
  return convtc_run(dycat, (const __nv_bfloat16*)wbuf + fwd_tile_elems(Cin, Cout, kh * kw), B, H, W, Cout, Cin,
                    live_taps(kh, kw, mask_mode, true), addend_or_null, dx, nullptr, (cudaStream_t)stream);
}

size_t lagvae_convtc_wgrad_scratch_bytes(int Cin, int Cout, int kh, int kw) {
  int tpg, upg, ng, S;
  wgrad_geometry(kh * kw, Cin, &tpg, &upg, &ng, &S);
  return (size_t)S * ng * upg * (2 * Cout) * 128 * sizeof(float) + 256;
}

int lagvae_convtc_wgrad(const uint16_t* dycat, const uint16_t* xcat, int B, int H, int W, int Cin, int Cout, int kh, int kw,
                        float* dw, void* scratch, void* stream) {
  LV_CHECK_ARG(dycat && xcat && dw && scratch && convtc_ok(B, H, W, Cin, Cout, kh, kw) && ((uintptr_t)scratch & 15) == 0,
               "convtc_wgrad: bad argument");
  cudaStream_t st = (cudaStream_t)stream;
  const int rpt = rows_per_tile(H, W);
  CUtensorMap tm_dy, tm_x;
  LV_TRY(make_tmap_nhwc_cat(&tm_dy, dycat, B, H, W, 2 * Cout, rpt));
  LV_TRY(make_tmap_nhwc_cat(&tm_x, xcat, B, H, W, 2 * Cin, rpt));
  WgradArgs g{};
  int ng;
  wgrad_geometry(kh * kw, Cin, &g.tpg, &g.upg, &ng, &g.S);
  g.S = std::min(g.S, B * (H / rpt));
  g.partial = (float*)scratch; g.B = B; g.H = H; g.W = W; g.RPT = rpt;
  g.a_chunks = Cout / 32; g.x_chunks = Cin / 32;
  g.taps = live_taps(kh, kw, 0, false);
  static bool configured = false;
  LV_TRY(set_smem(k_conv_wgrad_tc, W_SMEM, &configured));
  k_conv_wgrad_tc<<<ng * g.S, NTHREADS, W_SMEM, st>>>(tm_dy, tm_x, g);
  LV_LAUNCH_CHECK();
  const int n = kh * kw * Cout * Cin;
  k_conv_wgrad_reduce<<<(int)cdiv(n, 32), 256, 0, st>>>(g.partial, g.S, ng * g.upg, 2 * Cout, Cout, Cin, kh * kw, g.tpg, g.upg, dw);
  LV_LAUNCH_CHECK();
  return LAGVAE_OK;
}

}  // extern "C"
