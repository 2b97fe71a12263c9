// Persistent LSTM recurrence on tcgen05 (sm_100a): ONE launch runs all T time steps of the forward
// (or backward) recurrence of nn.LSTM (enc_lstm.py:60, dec_lstm.py:104 and their cuDNN backward).
//
// Decomposition (nh = 1024 -> 128 CTAs, one per SM, co-resident via cooperative launch):
//   forward : CTA c owns hidden units [8c, 8c+8)  = 32 gate columns (i,f,g,o x 8).  Its W_hh slice
//             [32 x nh] lives in shared memory for the whole kernel as split-bf16 (hi, lo) UMMA
//             B-operand tiles (128 KiB at nh = 1024).  Per step the previous hidden state h_{t-1}
//             [Bd x nh] (bf16 hi/lo, written by all CTAs) is streamed through a cp.async.cg ring
//             (L2 only) as the A operand; tcgen05.mma M=64 x N=32 x K=16 accumulates in TMEM
//             (hi*hi + hi*lo + lo*hi, fp32).  Four epilogue warps read TMEM (tcgen05.ld), add the
//             input projection, apply the gate non-linearities and the cell update in fp32, and
//             publish h_t (fp32 stash + bf16 hi/lo for the next step).
//   backward: CTA c owns the same 8 units; resident operand = W_hhᵀ slice [8 x 4nh] (128 KiB);
//             streamed operand = dG_{t+1} [Bd x 4nh] (bf16 hi/lo); UMMA M=64 x N=8.
//   One grid-wide barrier (atomic counter, release/acquire) separates time steps.
// Rows of the 64-row UMMA tile beyond the batch are zero (never written) — TMEM layout for M=64:
// row i -> lane 32*(i/16) + i%16 (cute/atom/mma_traits_sm100.hpp, tmem_frg M_MMA == 64).
#include "lstm_tc.cuh"
#include "kernels.cuh"
#include "sm100_ptx.cuh"

#include <cstdio>
#include <cstdlib>
#include <new>

namespace lagvae {

namespace {

constexpr int NTHREADS = 192;             // warps 0-3 epilogue, warp 4 MMA issuer, warp 5 TMA producer
constexpr int MAX_NS = 16;                // ring stages (barrier slots)
constexpr int MAX_MT = 4;                 // m-tiles of 64 batch rows (Bd <= 256; forward uses 128 TMEM columns per m-tile)
constexpr int SMEM_LIMIT = 232448;        // 227 KiB

struct TMaps {
  CUtensorMap m[4];   // [slot][part] views of the streamed operand: [Bd rows, KP cols] bf16, box 64 x rows_alloc, SWIZZLE_128B
};

struct RecArgs {
  int nh, Bd, Tn, KP, KB, NS, m_tiles;
  int part_bytes;                // bytes of one operand part of a ring stage = rows_alloc x 128 (rows_alloc = Bd rounded to 8, <= 64)
  unsigned long long* dbg;       // optional clock64 trace of CTA 0: [step][8]
  unsigned* bar;                 // grid barrier counter (host-zeroed)
  __nv_bfloat16* abuf;           // streamed operand (h or dG), bf16 hi/lo.  v1: [2 slots][2 parts][Bd][KP] (tensor-map view);
                                 // v2: "tile image" [2 slots][KB k-blocks][m_tiles][2 parts][rows_alloc][128 B], see img_off()
  const float* w_hh;             // [4nh, nh]
  // forward
  const float* h0;
  const float* c0;
  float* gates;                  // [Tn*Bd, 4nh] pre-activations in, activated out
  float* c_all;
  float* h_all;
  float* hdrop_all;
  DropSpec drop;
  // backward
  const float* dh_ext;           // [Tn*Bd, nh] or null
  const float* dh_last;          // [Bd, nh] or null
  float* dc;                     // [Bd, nh]
  float* dh_rec_out;             // [Bd, nh] (d h_{-1}) when want_init
  float* dgates;                 // [Tn*Bd, 4nh]
  int want_init;
  float* dgsum;                  // backward, optional: [Bd, 4nh] = sum over time of dG (the bias gradients' and dz's input), accumulated in
                                 // registers across the steps and written once — replaces a 105 MB re-read of dgates (time_sum)
  __nv_bfloat16* dg_hi;          // backward, optional: dG as the bf16 hi / lo tensor-core operand [Tn*Bd, 4nh] of the weight-gradient
  __nv_bfloat16* dg_lo;          // GEMMs, written next to the fp32 copy — replaces a k_split_bf16 pass over dgates
  int bulk_stages;               // v2: ring stages per cp.async.bulk copy (LAGVAE_LSTM_BULK_STAGES, default 4)
  int grouped;                   // v2: NS and the stages per step are multiples of bulk_stages -> the group partition is the same in
                                 // every step and only the group LEADERS' barriers are used (one commit / wait per group)
  int prod_fence;                // 1: producer-side fence.proxy.async before the arrival (round-1 behaviour; LAGVAE_LSTM_PROD_FENCE=1)
  const float* row_bias;         // forward, optional: [Bd, 4nh] added to the pre-activations of EVERY time step as they are loaded
                                 // (the decoder's z contribution, dec_lstm.py:84,97) — saves a read-modify-write pass over gates
  unsigned* started;             // optional: set to 1 once every CTA of the persistent grid is running (a side stream gates work on it
                                 // so that it takes the SMs the clusters leave over and never the ones they need)
};

__device__ __forceinline__ unsigned ld_acquire_u32(const unsigned* p) {
  unsigned v;
  asm volatile("ld.acquire.gpu.global.u32 %0, [%1];" : "=r"(v) : "l"(p) : "memory");
  return v;
}
// grid barrier on a monotonically increasing counter: bar.sync orders the CTA's stores before thread 0's
// release-add (cumulative), the acquire-poll orders them before every later load of the other CTAs
__device__ __forceinline__ void grid_barrier(unsigned* bar, unsigned target) {
  __syncthreads();
  if (threadIdx.x == 0) {
    asm volatile("red.release.gpu.global.add.u32 [%0], 1;" ::"l"(bar) : "memory");
    while (ld_acquire_u32(bar) < target) {
    }
  }
  __syncthreads();
}

__device__ __forceinline__ void tmem_alloc_rt(uint32_t smem_dst, uint32_t ncols) {
  asm volatile("tcgen05.alloc.cta_group::1.sync.aligned.shared::cta.b32 [%0], %1;" ::"r"(smem_dst), "r"(ncols)
               : "memory");
  asm volatile("tcgen05.relinquish_alloc_permit.cta_group::1.sync.aligned;" ::: "memory");
}
__device__ __forceinline__ void tmem_dealloc_rt(uint32_t taddr, uint32_t ncols) {
  asm volatile("tcgen05.dealloc.cta_group::1.sync.aligned.b32 %0, %1;" ::"r"(taddr), "r"(ncols) : "memory");
}
__device__ __forceinline__ void tmem_ld8(uint32_t taddr, uint32_t (&r)[8]) {
  asm volatile("tcgen05.ld.sync.aligned.32x32b.x8.b32 {%0, %1, %2, %3, %4, %5, %6, %7}, [%8];"
               : "=r"(r[0]), "=r"(r[1]), "=r"(r[2]), "=r"(r[3]), "=r"(r[4]), "=r"(r[5]), "=r"(r[6]), "=r"(r[7])
               : "r"(taddr)
               : "memory");
}

// swizzled (SWIZZLE_128B, K-major) byte offset of 16-byte chunk c of row r inside a tile of 128-B rows
__device__ __forceinline__ uint32_t sw128(int r, int c) { return (uint32_t)(r * 128 + ((c ^ (r & 7)) << 4)); }

// v2 operand layout in GLOBAL memory = the shared-memory image of the ring stages: stage (kb, mt) is one contiguous block
// [hi part rows ; lo part rows] x 128 B with SWIZZLE_128B already applied by the producing CTAs, so the consumer moves a
// whole stage with ONE cp.async.bulk (no tensor map; a tensor-map box of 32 separate 128-B rows cost ~2.5 K cycles of
// latency and ~10 B/clk per SM: profiles/r2a_exchange_probe_*.txt).  Byte offset inside a slot of the 16-B chunk that
// holds K elements [k0, k0 + 8) of batch row b:
__device__ __forceinline__ uint32_t img_off(int k0, int b, int part, int m_tiles, int rows_alloc) {
  const int kb = k0 >> 6, c = (k0 & 63) >> 3, mt = b >> 6, r = b & 63;
  return (uint32_t)((((kb * m_tiles + mt) * 2 + part) * rows_alloc + r) * 128 + ((c ^ (r & 7)) << 4));
}
__device__ __forceinline__ void bulk_g2s(uint32_t dst, const void* src, uint32_t bytes, uint32_t bar) {
  asm volatile("cp.async.bulk.shared::cluster.global.mbarrier::complete_tx::bytes [%0], [%1], %2, [%3];" ::"r"(dst), "l"(src),
               "r"(bytes), "r"(bar)
               : "memory");
}
__device__ __forceinline__ void store_bf16x4_img(uint8_t* slot, uint32_t off_hi, uint32_t off_lo, const float (&v)[4]) {
  __nv_bfloat16 hi[4], lo[4];
#pragma unroll
  for (int j = 0; j < 4; ++j) split_bf16(v[j], hi[j], lo[j]);
  uint2 h, l;
  h.x = (uint32_t)__bfloat16_as_ushort(hi[0]) | ((uint32_t)__bfloat16_as_ushort(hi[1]) << 16);
  h.y = (uint32_t)__bfloat16_as_ushort(hi[2]) | ((uint32_t)__bfloat16_as_ushort(hi[3]) << 16);
  l.x = (uint32_t)__bfloat16_as_ushort(lo[0]) | ((uint32_t)__bfloat16_as_ushort(lo[1]) << 16);
  l.y = (uint32_t)__bfloat16_as_ushort(lo[2]) | ((uint32_t)__bfloat16_as_ushort(lo[3]) << 16);
  *(uint2*)(slot + off_hi) = h;
  *(uint2*)(slot + off_lo) = l;
}

struct Smem {
  uint32_t w_base, a_base, bar_base;
  __device__ uint32_t full(int s) const { return bar_base + 8u * s; }
  __device__ uint32_t empty(int s, int NS) const { return bar_base + 8u * (NS + s); }
  __device__ uint32_t acc(int mt, int NS) const { return bar_base + 8u * (2 * NS + mt); }
  __device__ uint32_t tmem_slot(int NS) const { return bar_base + 8u * (2 * NS + MAX_MT); }
};

// ---- producer: stream rows [mt*64, ...) of the current slot through the ring -------------------
struct PipeState {
  int stage;
  uint32_t phase;
};

// whole producer warp (uniform); one elected lane issues two TMA loads (hi, lo part) per k-block.  TMA writes shared
// memory through the async proxy, so the UMMA consumer needs no generic->async proxy fence per stage (that fence
// waited for every in-flight copy and serialised the ring at one L2 round trip per k-block: profiles/README.md).
__device__ __forceinline__ void producer_pass(const RecArgs& a, const Smem& sm, PipeState& ps, const TMaps& tm,
                                              int slot, int mt) {
  for (int kb = 0; kb < a.KB; ++kb) {
    ptx::mbar_wait(sm.empty(ps.stage, a.NS), ps.phase ^ 1u);
    const uint32_t sbase = sm.a_base + ps.stage * 2 * a.part_bytes;
    if (ptx::elect_one()) {
      ptx::mbar_expect_tx(sm.full(ps.stage), 2u * (uint32_t)a.part_bytes);
      ptx::tma_load_2d(sbase, &tm.m[slot * 2 + 0], sm.full(ps.stage), kb * 64, mt * 64);
      ptx::tma_load_2d(sbase + a.part_bytes, &tm.m[slot * 2 + 1], sm.full(ps.stage), kb * 64, mt * 64);
    }
    __syncwarp();
    if (++ps.stage == a.NS) { ps.stage = 0; ps.phase ^= 1u; }
  }
}

// ---- MMA issuer: one K sweep for m-tile mt ------------------------------------------------------
// Weight tile of one k-block = [2*NB rows x 64 k]: rows [0,NB) = bf16 hi of the NB resident columns, rows
// [NB,2NB) = their lo parts.  Per K sub-step TWO MMAs instead of three:
//     D[:, 0:2NB] += A_hi · [W_hi ; W_lo]ᵀ      (N = 2NB: hi·hi and hi·lo in one instruction)
//     D[:, 0:NB]  += A_lo · W_hiᵀ               (N = NB)
// and the epilogue adds D[:, 0:NB] + D[:, NB:2NB].  (tcgen05.mma with M=64 costs ~55 cycles whatever N <= 64
// is — measured with the clock64 trace — so the instruction count is what matters.)
constexpr int NACC = 2;  // accumulator slots per m-tile (alternating K sub-steps)

// called by the WHOLE MMA warp (uniform control flow); one elected lane issues the tensor-core work
template <int NB>
__device__ __forceinline__ void mma_pass(const RecArgs& a, const Smem& sm, PipeState& ps, uint32_t d_tmem,
                                         uint32_t acc_bar) {
  constexpr uint32_t idesc_all = ptx::make_idesc_bf16_f32(64, 2 * NB, 0, 0);
  constexpr uint32_t idesc_hi = ptx::make_idesc_bf16_f32(64, NB, 0, 0);
  constexpr int WT2 = 2 * NB * 128;  // bytes of one merged weight tile
  for (int kb = 0; kb < a.KB; ++kb) {
    ptx::mbar_wait(sm.full(ps.stage), ps.phase);
    ptx::tc_fence_after();
    const uint32_t sa = sm.a_base + ps.stage * 2 * a.part_bytes;
    const uint32_t sw = sm.w_base + (uint32_t)kb * WT2;
    if (ptx::elect_one()) {
#pragma unroll
      for (int k = 0; k < 4; ++k) {
        const uint64_t a_hi = ptx::make_smem_desc_sw128(sa + k * 32, 16, 1024);
        const uint64_t a_lo = ptx::make_smem_desc_sw128(sa + a.part_bytes + k * 32, 16, 1024);
        const uint64_t b = ptx::make_smem_desc_sw128(sw + k * 32, 16, 1024);
        const uint32_t d = d_tmem + (uint32_t)((k & (NACC - 1)) * 2 * NB);
        ptx::umma_f16(d, a_hi, b, idesc_all, (kb > 0 || k >= NACC) ? 1u : 0u);
        ptx::umma_f16(d, a_lo, b, idesc_hi, 1u);
      }
      ptx::umma_commit(sm.empty(ps.stage, a.NS));
      if (kb == a.KB - 1) ptx::umma_commit(acc_bar);
    }
    __syncwarp();
    if (++ps.stage == a.NS) { ps.stage = 0; ps.phase ^= 1u; }
  }
}

// fast, accurate-to-~3e-7 gate non-linearities (ex2.approx based; the precise tanhf/expf paths cost ~4x more
// instructions and the cell epilogue is on the per-time-step critical path)
__device__ __forceinline__ float fsigmoid(float x) { return __fdividef(1.f, 1.f + __expf(-x)); }
__device__ __forceinline__ float ftanh(float x) {
  const float t = __expf(-2.f * fabsf(x));
  return copysignf(__fdividef(1.f - t, 1.f + t), x);
}
__device__ __forceinline__ void store_bf16x4(__nv_bfloat16* dst_hi, __nv_bfloat16* dst_lo, const float (&v)[4]) {
  __nv_bfloat16 hi[4], lo[4];
#pragma unroll
  for (int j = 0; j < 4; ++j) split_bf16(v[j], hi[j], lo[j]);
  uint2 h, l;
  h.x = (uint32_t)__bfloat16_as_ushort(hi[0]) | ((uint32_t)__bfloat16_as_ushort(hi[1]) << 16);
  h.y = (uint32_t)__bfloat16_as_ushort(hi[2]) | ((uint32_t)__bfloat16_as_ushort(hi[3]) << 16);
  l.x = (uint32_t)__bfloat16_as_ushort(lo[0]) | ((uint32_t)__bfloat16_as_ushort(lo[1]) << 16);
  l.y = (uint32_t)__bfloat16_as_ushort(lo[2]) | ((uint32_t)__bfloat16_as_ushort(lo[3]) << 16);
  *(uint2*)dst_hi = h;
  *(uint2*)dst_lo = l;
}

__device__ __forceinline__ void store_bf16x8(__nv_bfloat16* dst_hi, __nv_bfloat16* dst_lo, const float (&v)[8]) {
  __align__(16) __nv_bfloat16 hi[8], lo[8];
#pragma unroll
  for (int j = 0; j < 8; ++j) split_bf16(v[j], hi[j], lo[j]);
  *(uint4*)dst_hi = *(const uint4*)hi;
  *(uint4*)dst_lo = *(const uint4*)lo;
}

__device__ __forceinline__ void common_prologue(const RecArgs& a, const Smem& sm, int tmem_cols, uint32_t* slot_ptr,
                                                uint32_t full_count = 1) {
  const int warp = threadIdx.x >> 5;
  // ring rows beyond the batch are never written: the UMMA reads 64 rows, but D row i depends on A row i only and
  // rows >= Bd of D are never read back, so whatever aliases there (next part / next stage / W tiles) is harmless
  if (threadIdx.x == 0) {
    for (int s = 0; s < a.NS; ++s) {
      ptx::mbar_init(sm.full(s), full_count);
      ptx::mbar_init(sm.empty(s, a.NS), 1);
    }
    for (int m = 0; m < MAX_MT; ++m) ptx::mbar_init(sm.acc(m, a.NS), 1);
    ptx::fence_mbar_init();
  }
  if (warp == 4) tmem_alloc_rt(sm.tmem_slot(a.NS), (uint32_t)tmem_cols);
  ptx::fence_proxy_async_smem();
  ptx::tc_fence_before();
  __syncthreads();
  ptx::tc_fence_after();
  (void)slot_ptr;
}

// =================================================================================================
// forward
// =================================================================================================
__global__ void __launch_bounds__(NTHREADS, 1) k_lstm_fwd_tc(const RecArgs a, const __grid_constant__ TMaps tm) {
  extern __shared__ uint8_t smem_raw[];
  constexpr int NB = 32, WT2 = 2 * NB * 128, TCOLS = NACC * 2 * NB;  // 128 TMEM columns per m-tile
  Smem sm;
  sm.a_base = (ptx::smem_u32(smem_raw) + 1023u) & ~1023u;
  sm.w_base = sm.a_base + (uint32_t)a.NS * 2 * a.part_bytes;
  sm.bar_base = sm.w_base + (uint32_t)a.KB * WT2;
  uint8_t* gen_base = smem_raw + (sm.w_base - ptx::smem_u32(smem_raw));
  const int warp = threadIdx.x >> 5, lane = threadIdx.x & 31;
  const int nh = a.nh, Bd = a.Bd, u0 = blockIdx.x * 8;
  const int need_cols = a.m_tiles * TCOLS;
  const int tmem_cols = need_cols <= 128 ? 128 : (need_cols <= 256 ? 256 : 512);

  // resident W_hh slice: merged tile (kb) = [64 rows x 64 k] bf16: row n = gate*8 + uu (hi), row 32 + n (lo)
  for (int id = threadIdx.x; id < 32 * (a.KP / 8); id += NTHREADS) {
    const int n = id / (a.KP / 8), ck = id % (a.KP / 8);
    const int k0 = ck * 8;
    const float* src = a.w_hh + (int64_t)((n >> 3) * nh + u0 + (n & 7)) * nh + k0;
    float v[8];
#pragma unroll
    for (int j = 0; j < 8; ++j) v[j] = (k0 + j < nh) ? src[j] : 0.f;
    const int kb = k0 >> 6, c = (k0 & 63) >> 3;
    uint8_t* tile = gen_base + (size_t)kb * WT2;
    store_bf16x8((__nv_bfloat16*)(tile + sw128(n, c)), (__nv_bfloat16*)(tile + sw128(NB + n, c)), v);
  }
  common_prologue(a, sm, tmem_cols, nullptr);
  const uint32_t tmem_base = *(uint32_t*)(gen_base + (sm.tmem_slot(a.NS) - sm.w_base));

  // publish h_{-1} (own 8 units, all rows) into slot 1
  const int64_t slot_elems = (int64_t)2 * Bd * a.KP;
  for (int b = threadIdx.x; b < Bd; b += NTHREADS) {
    float v[8];
#pragma unroll
    for (int j = 0; j < 8; ++j) v[j] = a.h0 ? a.h0[(int64_t)b * nh + u0 + j] : 0.f;
    __nv_bfloat16* d = a.abuf + slot_elems + (int64_t)b * a.KP + u0;
    store_bf16x8(d, d + (int64_t)Bd * a.KP, v);
  }
  ptx::fence_proxy_async_all();   // generic stores above are read by other CTAs' TMA (async proxy) after the barrier
  unsigned epoch = 0;
  grid_barrier(a.bar, (++epoch) * gridDim.x);

  PipeState ps{0, 0};
  const bool trace = a.dbg != nullptr && blockIdx.x == 0;
  for (int t = 0; t < a.Tn; ++t) {
    if (trace && threadIdx.x == 0) a.dbg[t * 8 + 0] = clock64();                 // step start (barrier exit)
    __nv_bfloat16* wr = a.abuf + (int64_t)(t & 1) * slot_elems;
    if (warp == 5) {
      ptx::fence_proxy_async_all();   // h_{t-1} was written with generic stores by other SMs (acquired at the barrier)
      for (int mt = 0; mt < a.m_tiles; ++mt) producer_pass(a, sm, ps, tm, (t + 1) & 1, mt);
    } else if (warp == 4) {
      if (trace && lane == 0) {   // first stage of this step landed?
        ptx::mbar_wait(sm.full(ps.stage), ps.phase);
        a.dbg[t * 8 + 1] = clock64();
      }
      for (int mt = 0; mt < a.m_tiles; ++mt)
        mma_pass<NB>(a, sm, ps, tmem_base + (uint32_t)(mt * TCOLS), sm.acc(mt, a.NS));
      if (trace && lane == 0) a.dbg[t * 8 + 2] = clock64();                      // all MMAs issued
    } else {
      // epilogue: TMEM lane l (< 16) of warp w holds batch row 16w + l.  Lane l handles units 0-3 of that row,
      // lane l + 16 handles units 4-7 (values handed over by shuffle): all 32 lanes of the warp do cell maths.
      float* gates_t = a.gates + (int64_t)t * Bd * 4 * nh;
      const int half = lane >> 4, ub = u0 + half * 4;
      for (int mt = 0; mt < a.m_tiles; ++mt) {
        const int b = mt * 64 + warp * 16 + (lane & 15);
        const bool valid = b < Bd;
        float pre[16], cp[4];
        if (valid) {  // prefetch the input projection and c_{t-1} while the MMAs run
          const float* g = gates_t + (int64_t)b * 4 * nh + ub;
#pragma unroll
          for (int q = 0; q < 4; ++q) {
            const float4 x0 = *(const float4*)(g + q * nh);
            pre[q * 4 + 0] = x0.x; pre[q * 4 + 1] = x0.y; pre[q * 4 + 2] = x0.z; pre[q * 4 + 3] = x0.w;
          }
          const float* cpp = t ? a.c_all + ((int64_t)(t - 1) * Bd + b) * nh + ub : (a.c0 ? a.c0 + (int64_t)b * nh + ub : nullptr);
#pragma unroll
          for (int j = 0; j < 4; ++j) cp[j] = cpp ? cpp[j] : 0.f;
        }
        ptx::mbar_wait(sm.acc(mt, a.NS), (uint32_t)(t & 1));
        ptx::tc_fence_after();
        if (trace && threadIdx.x == 0 && mt == 0) a.dbg[t * 8 + 3] = clock64();  // accumulators complete
        float acc[32];
#pragma unroll
        for (int j = 0; j < 32; ++j) acc[j] = 0.f;
#pragma unroll
        for (int sl = 0; sl < NACC * 2; ++sl) {   // NACC slots x {A·W_hi (+A_lo·W_hi), A_hi·W_lo}
          uint32_t r[32];
          ptx::tmem_ld32(tmem_base + ((uint32_t)(warp * 32) << 16) + (uint32_t)(mt * TCOLS + sl * 32), r);
          ptx::tmem_ld_wait();
#pragma unroll
          for (int j = 0; j < 32; ++j) acc[j] += __uint_as_float(r[j]);
        }
        // hand units 4-7 to the upper half-warp
        float my[16];
#pragma unroll
        for (int q = 0; q < 4; ++q)
#pragma unroll
          for (int j = 0; j < 4; ++j) {
            const float up = __shfl_sync(0xffffffffu, acc[q * 8 + 4 + j], lane & 15);
            my[q * 4 + j] = half ? up : acc[q * 8 + j];
          }
        if (valid) {
          float hv[4], cv[4], act[16];
#pragma unroll
          for (int j = 0; j < 4; ++j) {
            const float ig = fsigmoid(pre[j] + my[j]);
            const float fg = fsigmoid(pre[4 + j] + my[4 + j]);
            const float gg = ftanh(pre[8 + j] + my[8 + j]);
            const float og = fsigmoid(pre[12 + j] + my[12 + j]);
            const float c = fg * cp[j] + ig * gg;
            cv[j] = c;
            hv[j] = og * ftanh(c);
            act[j] = ig; act[4 + j] = fg; act[8 + j] = gg; act[12 + j] = og;
          }
          __nv_bfloat16* d = wr + (int64_t)b * a.KP + ub;
          store_bf16x4(d, d + (int64_t)Bd * a.KP, hv);          // next step's operand first
          float* g = gates_t + (int64_t)b * 4 * nh + ub;
#pragma unroll
          for (int q = 0; q < 4; ++q)
            *(float4*)(g + q * nh) = make_float4(act[q * 4], act[q * 4 + 1], act[q * 4 + 2], act[q * 4 + 3]);
          const int64_t o = ((int64_t)t * Bd + b) * nh + ub;
          *(float4*)(a.c_all + o) = make_float4(cv[0], cv[1], cv[2], cv[3]);
          *(float4*)(a.h_all + o) = make_float4(hv[0], hv[1], hv[2], hv[3]);
          if (a.hdrop_all) {
            float hd[4];
#pragma unroll
            for (int j = 0; j < 4; ++j) hd[j] = hv[j] * drop_factor(a.drop, ((uint64_t)b * a.Tn + t) * nh + ub + j);
            *(float4*)(a.hdrop_all + o) = make_float4(hd[0], hd[1], hd[2], hd[3]);
          }
        }
      }
      ptx::fence_proxy_async_all();   // this thread's h_t stores -> visible to the async proxy (next step's TMA)
      ptx::tc_fence_before();
      if (trace && threadIdx.x == 0) a.dbg[t * 8 + 4] = clock64();               // epilogue stores issued
    }
    grid_barrier(a.bar, (++epoch) * gridDim.x);
    ptx::tc_fence_after();
  }
  if (warp == 4) tmem_dealloc_rt(tmem_base, (uint32_t)tmem_cols);
}

// =================================================================================================
// backward
// =================================================================================================
__global__ void __launch_bounds__(NTHREADS, 1) k_lstm_bwd_tc(const RecArgs a, const __grid_constant__ TMaps tm) {
  extern __shared__ uint8_t smem_raw[];
  constexpr int NB = 8, WT2 = 2 * NB * 128, TCOLS = NACC * 2 * NB;   // 32 TMEM columns per m-tile
  Smem sm;
  sm.a_base = (ptx::smem_u32(smem_raw) + 1023u) & ~1023u;
  sm.w_base = sm.a_base + (uint32_t)a.NS * 2 * a.part_bytes;
  sm.bar_base = sm.w_base + (uint32_t)a.KB * WT2;
  uint8_t* gen_base = smem_raw + (sm.w_base - ptx::smem_u32(smem_raw));
  const int warp = threadIdx.x >> 5, lane = threadIdx.x & 31;
  const int nh = a.nh, Bd = a.Bd, Tn = a.Tn, u0 = blockIdx.x * 8;
  const int need_cols = a.m_tiles * TCOLS;
  const int tmem_cols = need_cols <= 32 ? 32 : (need_cols <= 64 ? 64 : (need_cols <= 128 ? 128 : 256));

  // resident W_hhᵀ slice: rows n = unit uu, K index = gate column k in [0, 4nh): W_hh[k, u0+uu]
  for (int id = threadIdx.x; id < 8 * (a.KP / 8); id += NTHREADS) {
    const int n = id % 8, ck = id / 8;   // consecutive threads -> consecutive units (32-B segments of a W_hh row)
    const int k0 = ck * 8;
    float v[8];
#pragma unroll
    for (int j = 0; j < 8; ++j) v[j] = (k0 + j < 4 * nh) ? a.w_hh[(int64_t)(k0 + j) * nh + u0 + n] : 0.f;
    const int kb = k0 >> 6, c = (k0 & 63) >> 3;
    uint8_t* tile = gen_base + (size_t)kb * WT2;   // merged tile: rows [0,8) hi, rows [8,16) lo
    store_bf16x8((__nv_bfloat16*)(tile + sw128(n, c)), (__nv_bfloat16*)(tile + sw128(NB + n, c)), v);
  }
  common_prologue(a, sm, tmem_cols, nullptr);
  const uint32_t tmem_base = *(uint32_t*)(gen_base + (sm.tmem_slot(a.NS) - sm.w_base));
  for (int i = threadIdx.x; i < Bd * 8; i += NTHREADS) a.dc[(int64_t)(i >> 3) * nh + u0 + (i & 7)] = 0.f;
  __syncthreads();

  const int64_t slot_elems = (int64_t)2 * Bd * a.KP;
  unsigned epoch = 0;
  PipeState ps{0, 0};
  const int nsteps = Tn + (a.want_init ? 1 : 0);
  const bool trace = a.dbg != nullptr && blockIdx.x == 0;
  for (int s = 0; s < nsteps; ++s) {
    if (trace && threadIdx.x == 0) a.dbg[s * 8 + 0] = clock64();
    const int t = Tn - 1 - s;            // t = -1 on the extra step that only produces d h_{-1}
    const bool has_rec = s > 0;
    __nv_bfloat16* wr = a.abuf + (int64_t)(s & 1) * slot_elems;
    if (warp == 5) {
      if (has_rec) {
        ptx::fence_proxy_async_all();
        for (int mt = 0; mt < a.m_tiles; ++mt) producer_pass(a, sm, ps, tm, (s + 1) & 1, mt);
      }
    } else if (warp == 4) {
      if (has_rec) {
        if (trace && lane == 0) {
          ptx::mbar_wait(sm.full(ps.stage), ps.phase);
          a.dbg[s * 8 + 1] = clock64();
        }
        for (int mt = 0; mt < a.m_tiles; ++mt)
          mma_pass<NB>(a, sm, ps, tmem_base + (uint32_t)(mt * TCOLS), sm.acc(mt, a.NS));
        if (trace && lane == 0) a.dbg[s * 8 + 2] = clock64();
      }
    } else {
      for (int mt = 0; mt < a.m_tiles; ++mt) {
        const int b = mt * 64 + warp * 16 + lane;
        const bool valid = lane < 16 && b < Bd;
        float gt[32], cc[8], cp[8], dcv[8], dhx[8];
        if (valid && t >= 0) {
          const float* g = a.gates + ((int64_t)t * Bd + b) * 4 * nh + u0;
#pragma unroll
          for (int q = 0; q < 4; ++q) {
            const float4 x0 = *(const float4*)(g + q * nh), x1 = *(const float4*)(g + q * nh + 4);
            gt[q * 8 + 0] = x0.x; gt[q * 8 + 1] = x0.y; gt[q * 8 + 2] = x0.z; gt[q * 8 + 3] = x0.w;
            gt[q * 8 + 4] = x1.x; gt[q * 8 + 5] = x1.y; gt[q * 8 + 6] = x1.z; gt[q * 8 + 7] = x1.w;
          }
          const int64_t o = ((int64_t)t * Bd + b) * nh + u0;
          const float* cpp = t ? a.c_all + o - (int64_t)Bd * nh : (a.c0 ? a.c0 + (int64_t)b * nh + u0 : nullptr);
#pragma unroll
          for (int j = 0; j < 8; ++j) {
            cc[j] = a.c_all[o + j];
            cp[j] = cpp ? cpp[j] : 0.f;
            dcv[j] = a.dc[(int64_t)b * nh + u0 + j];
            float e = 0.f;
            if (a.dh_ext) e = a.dh_ext[o + j] * drop_factor(a.drop, ((uint64_t)b * Tn + t) * nh + u0 + j);
            if (t == Tn - 1 && a.dh_last) e += a.dh_last[(int64_t)b * nh + u0 + j];
            dhx[j] = e;
          }
        }
        float rec[8];
#pragma unroll
        for (int j = 0; j < 8; ++j) rec[j] = 0.f;
        if (has_rec) {
          ptx::mbar_wait(sm.acc(mt, a.NS), (uint32_t)((s - 1) & 1));
          ptx::tc_fence_after();
          if (trace && threadIdx.x == 0 && mt == 0) a.dbg[s * 8 + 3] = clock64();
#pragma unroll
          for (int sl = 0; sl < NACC * 2; ++sl) {   // NACC slots x {cols 0-7, cols 8-15}
            uint32_t r[8];
            tmem_ld8(tmem_base + ((uint32_t)(warp * 32) << 16) + (uint32_t)(mt * TCOLS + sl * 8), r);
            ptx::tmem_ld_wait();
#pragma unroll
            for (int j = 0; j < 8; ++j) rec[j] += __uint_as_float(r[j]);
          }
        }
        if (valid) {
          if (t < 0) {
#pragma unroll
            for (int j = 0; j < 8; ++j) a.dh_rec_out[(int64_t)b * nh + u0 + j] = rec[j];
          } else {
            float dg[32], dcn[8];
#pragma unroll
            for (int j = 0; j < 8; ++j) {
              const float ig = gt[j], fg = gt[8 + j], gg = gt[16 + j], og = gt[24 + j];
              const float dh = rec[j] + dhx[j];
              const float tc = ftanh(cc[j]);
              const float dct = dcv[j] + dh * og * (1.f - tc * tc);
              dg[j] = dct * gg * ig * (1.f - ig);
              dg[8 + j] = dct * cp[j] * fg * (1.f - fg);
              dg[16 + j] = dct * ig * (1.f - gg * gg);
              dg[24 + j] = dh * tc * og * (1.f - og);
              dcn[j] = dct * fg;
            }
            float* go = a.dgates + ((int64_t)t * Bd + b) * 4 * nh + u0;
            __nv_bfloat16* wb = wr + (int64_t)b * a.KP + u0;
#pragma unroll
            for (int q = 0; q < 4; ++q) {
              *(float4*)(go + q * nh) = make_float4(dg[q * 8], dg[q * 8 + 1], dg[q * 8 + 2], dg[q * 8 + 3]);
              *(float4*)(go + q * nh + 4) = make_float4(dg[q * 8 + 4], dg[q * 8 + 5], dg[q * 8 + 6], dg[q * 8 + 7]);
              float seg[8];
#pragma unroll
              for (int j = 0; j < 8; ++j) seg[j] = dg[q * 8 + j];
              store_bf16x8(wb + q * nh, wb + q * nh + (int64_t)Bd * a.KP, seg);
            }
#pragma unroll
            for (int j = 0; j < 8; ++j) a.dc[(int64_t)b * nh + u0 + j] = dcn[j];
          }
        }
      }
      ptx::fence_proxy_async_all();   // dG_t stores -> async proxy (next step's TMA)
      ptx::tc_fence_before();
      if (trace && threadIdx.x == 0) a.dbg[s * 8 + 4] = clock64();
    }
    grid_barrier(a.bar, (++epoch) * gridDim.x);
    ptx::tc_fence_after();
  }
  if (warp == 4) tmem_dealloc_rt(tmem_base, (uint32_t)tmem_cols);
}


// =================================================================================================
// v2: cluster K-split recurrence.  tcgen05.mma with M=64 costs ~58 cycles for any N <= 64 and N/2 cycles above
// (scripts/microbench/mma_cost.cu), so the per-step MMA time is (#instructions) x 58: the v1 kernels (N = 32 / 8,
// full K per CTA: 128-192 / 512-768 MMAs per step) are instruction-count bound.  v2 keeps the SAME unit ownership
// and global data layout but lets a thread-block CLUSTER split K:
//   forward : cluster of 4 CTAs = 32 units (128 gate columns).  CTA rank r holds W_hh[128 cols, K-slice r] (hi+lo
//             merged: N = 256) and multiplies its quarter of h_{t-1}: 16 K sub-steps x 2 MMAs per time step.
//   backward: cluster of 8 CTAs = 64 units.  CTA rank r holds W_hhᵀ[64 units, K-slice r of the 4nh gate columns]
//             (N = 128) and multiplies its eighth of dG_{t+1}: 32 K sub-steps x 2 MMAs per time step (v1: 512).
//   The partial products [Bd x NC] go TMEM -> shared memory; after a cluster barrier every CTA sums the NC/CS
//   columns it owns over the CS peers through distributed shared memory (ld.shared::cluster) and runs the cell.
// =================================================================================================
__device__ __forceinline__ uint32_t cluster_ctarank() {
  uint32_t r;
  asm volatile("mov.u32 %0, %%cluster_ctarank;" : "=r"(r));
  return r;
}
__device__ __forceinline__ void cluster_sync_all() {
  asm volatile("barrier.cluster.arrive.release.aligned;" ::: "memory");
  asm volatile("barrier.cluster.wait.acquire.aligned;" ::: "memory");
}
__device__ __forceinline__ float4 ld_dsmem_f4(uint32_t local_addr, uint32_t rank) {
  uint32_t ra;
  asm volatile("mapa.shared::cluster.u32 %0, %1, %2;" : "=r"(ra) : "r"(local_addr), "r"(rank));
  float4 v;
  asm volatile("ld.shared::cluster.v4.f32 {%0, %1, %2, %3}, [%4];" : "=f"(v.x), "=f"(v.y), "=f"(v.z), "=f"(v.w) : "r"(ra) : "memory");
  return v;
}

__device__ __forceinline__ uint32_t mapa_u32(uint32_t local_addr, uint32_t rank) {
  uint32_t ra;
  asm volatile("mapa.shared::cluster.u32 %0, %1, %2;" : "=r"(ra) : "r"(local_addr), "r"(rank));
  return ra;
}
__device__ __forceinline__ void st_cluster_f4(uint32_t cluster_addr, float x, float y, float z, float w) {
  asm volatile("st.shared::cluster.v4.f32 [%0], {%1, %2, %3, %4};" ::"r"(cluster_addr), "f"(x), "f"(y), "f"(z), "f"(w) : "memory");
}
// Receive-slot protocol: a slot word holds RX_EMPTY (a NaN payload no arithmetic produces) until the contributor's
// 16-byte DSMEM store lands; the owner polls the data itself and re-arms the slot after reading it.  No fence, barrier
// or mbarrier round trip sits between the contributor's tcgen05.ld and the owner's cell update.
constexpr uint32_t RX_EMPTY = 0xFFFFA5C3u;
__device__ __forceinline__ float4 ld_shared_volatile_f4(uint32_t addr) {
  float4 v;
  asm volatile("ld.volatile.shared.v4.f32 {%0, %1, %2, %3}, [%4];" : "=f"(v.x), "=f"(v.y), "=f"(v.z), "=f"(v.w) : "r"(addr) : "memory");
  return v;
}
__device__ __forceinline__ void st_shared_u4(uint32_t addr, uint32_t x) {
  asm volatile("st.shared.v4.b32 [%0], {%1, %1, %1, %1};" ::"r"(addr), "r"(x) : "memory");
}
__device__ __forceinline__ bool rx_ready(const float4& v) {
  return __float_as_uint(v.x) != RX_EMPTY && __float_as_uint(v.y) != RX_EMPTY && __float_as_uint(v.z) != RX_EMPTY &&
         __float_as_uint(v.w) != RX_EMPTY;
}

// v2 operand load: the 4 epilogue warps (idle until the accumulators complete) copy the CTA's K slice of the published
// operand from global memory (L2) into the ring with plain 16-byte loads and stores — the global tile image IS the
// shared-memory image — G stages (16 chunks per thread) in flight at a time, then a CTA-local generic->async proxy fence
// and one mbarrier arrival per stage for the MMA warp.  Measured against cp.async.bulk / tensor TMA for this access
// pattern (128 CTAs pulling 64 KB each out of the same 128 KB every ~6 us; profiles/r2b_exchange_probe_*.txt): the bulk
// copies cost ~2.3 K cycles to the first byte plus ~0.7 K per 8 KB copy (P6: 10.2 K cycles per exchange) against ~2.0 K
// cycles for the whole 64 KB with LDG.128 (P15: 6.5 K), and a global-scope fence.proxy.async (724 cycles) disappears
// because no async-proxy operation reads global memory any more.
template <bool FWD, int CS_>
struct V2Cfg {
  static constexpr int CS = CS_;                  // cluster size = K split (forward 4; backward 8, or 4 when 16 clusters of 8 do not fit)
  static constexpr int NOWN = FWD ? 32 : 8;       // result columns owned per CTA (8 units x 4 gates | 8 units)
  static constexpr int NC = CS * NOWN;            // columns of the cluster (128 | 64)
  static constexpr int NALL = 2 * NC;             // merged hi+lo weight rows = MMA N (256 | 128)
  static constexpr int WT = NALL * 128;           // bytes of one weight tile (one 64-wide k-block)
  static constexpr int RROW = NOWN + 4;           // floats per row of a receive slot (owned columns + bank padding)
};

template <bool FWD, int CS_>
__global__ void __launch_bounds__(NTHREADS, 1) k_lstm_v2(const RecArgs a) {
  using C = V2Cfg<FWD, CS_>;
  extern __shared__ uint8_t smem_raw[];
  const int warp = threadIdx.x >> 5, lane = threadIdx.x & 31;
  const int nh = a.nh, Bd = a.Bd, Tn = a.Tn;
  const uint32_t rank = cluster_ctarank();
  const int cl = blockIdx.x / C::CS;                 // cluster index
  const int u0 = blockIdx.x * 8;                     // owned units (same ownership as v1)
  const int KBS = a.KB / C::CS;                      // k-blocks of this CTA's K slice
  const int rows_alloc = a.part_bytes / 128;
  Smem sm;
  sm.a_base = (ptx::smem_u32(smem_raw) + 1023u) & ~1023u;
  sm.w_base = sm.a_base + (uint32_t)a.NS * 2 * a.part_bytes;
  // "receive slots": every CTA of the cluster PUSHES the partial sums of the columns another CTA owns straight into that
  // CTA's shared memory (st.shared::cluster) and arrives on its mbarrier; the owner then reduces locally.
  const int rows_tot = a.m_tiles * rows_alloc;
  const uint32_t r_base = sm.w_base + (uint32_t)KBS * C::WT;                       // [nslots][rows_tot][RROW] fp32
  // "stacked M" (Bd <= 32): the ring stage is [hi rows ; lo rows] contiguously, so ONE M=64 MMA whose A descriptor
  // starts at the hi part covers both operand parts: D rows [0,ra) = A_hi·[W_hi;W_lo], rows [ra,2ra) = A_lo·[W_hi;W_lo];
  // the epilogue adds D[b,0:NC] + D[b,NC:2NC] + D[ra+b,0:NC].  Halves the tcgen05.mma count per time step.
  const bool stack = a.m_tiles == 1 && 2 * rows_alloc <= 64;
  const int nslots = C::CS * (stack ? 2 : 1);                                      // contributors per owned value
  sm.bar_base = r_base + (uint32_t)(nslots * rows_tot * C::RROW * 4);
  uint8_t* gen_w = smem_raw + (sm.w_base - ptx::smem_u32(smem_raw));
  const int tmem_need = a.m_tiles * C::NALL;
  const int tmem_cols = tmem_need <= 128 ? 128 : (tmem_need <= 256 ? 256 : 512);

  // ---- resident weight slice -> merged [hi rows ; lo rows] UMMA B tiles -------------------------------------
  {
    const int kslice0 = (int)rank * KBS * 64;
    const int nchunk = KBS * 8;                                // 16-B chunks (8 k) per row
    for (int id = threadIdx.x; id < C::NC * nchunk; id += NTHREADS) {
      int j, ck;
      if (FWD) { j = id / nchunk; ck = id % nchunk; } else { j = id % C::NC; ck = id / C::NC; }
      const int k0 = kslice0 + ck * 8;
      float v[8];
      if (FWD) {   // column j = q*32 + gate*8 + uu  <->  W_hh row gate*nh + (8 CS) cl + 8q + uu ; K index = hidden unit
        const int q = j >> 5, gate = (j >> 3) & 3, uu = j & 7;
        const float* src = a.w_hh + (int64_t)(gate * nh + 8 * C::CS * cl + 8 * q + uu) * nh + k0;
#pragma unroll
        for (int e = 0; e < 8; ++e) v[e] = (k0 + e < nh) ? src[e] : 0.f;
      } else {     // row j = unit NC*cl + j ; K index = gate column k: W_hh[k, unit]
#pragma unroll
        for (int e = 0; e < 8; ++e) v[e] = (k0 + e < 4 * nh) ? a.w_hh[(int64_t)(k0 + e) * nh + C::NC * cl + j] : 0.f;
      }
      uint8_t* tile = gen_w + (size_t)(ck >> 3) * C::WT;
      const int c = ck & 7;
      store_bf16x8((__nv_bfloat16*)(tile + sw128(j, c)), (__nv_bfloat16*)(tile + sw128(C::NC + j, c)), v);
    }
  }
  for (int i = threadIdx.x; i < nslots * rows_tot * C::RROW / 4; i += NTHREADS) st_shared_u4(r_base + 16u * i, RX_EMPTY);
  common_prologue(a, sm, tmem_cols, nullptr);
  const uint32_t tmem_base = *(uint32_t*)(gen_w + (sm.tmem_slot(a.NS) - sm.w_base));
  const uint32_t stage_bytes = 2u * (uint32_t)a.part_bytes;
  const size_t slot_bytes_g = (size_t)a.KB * a.m_tiles * stage_bytes;            // one operand slot in global memory
  uint8_t* const abuf8 = (uint8_t*)a.abuf;
  if (FWD) {   // publish h_{-1} (own 8 units, all rows) into slot 1
    for (int b = threadIdx.x; b < Bd; b += NTHREADS) {
      float v[8];
#pragma unroll
      for (int j = 0; j < 8; ++j) v[j] = a.h0 ? a.h0[(int64_t)b * nh + u0 + j] : 0.f;
      uint8_t* slot = abuf8 + slot_bytes_g;
      store_bf16x8((__nv_bfloat16*)(slot + img_off(u0, b, 0, a.m_tiles, rows_alloc)),
                   (__nv_bfloat16*)(slot + img_off(u0, b, 1, a.m_tiles, rows_alloc)), v);
    }
  } else {
    for (int i = threadIdx.x; i < Bd * 8; i += NTHREADS) a.dc[(int64_t)(i >> 3) * nh + u0 + (i & 7)] = 0.f;
  }
  ptx::fence_proxy_async_all();
  cluster_sync_all();                                // peers' smem (receive slots, barriers) exist before any DSMEM access
  grid_barrier(a.bar, gridDim.x);                    // epoch 1: initial operand published; step s ends epoch s + 2
  if (a.started != nullptr && blockIdx.x == 0 && threadIdx.x == 0) *(volatile unsigned*)a.started = 1u;

  PipeState ps{0, 0};      // ring position (producer warp and MMA warp each advance their own copy)
  const bool trace = a.dbg != nullptr && blockIdx.x == 0;
  const int nsteps = FWD ? Tn : Tn + (a.want_init ? 1 : 0);
  int acc_par = 0;
  float dgs[16];                 // backward: running sum over time of this thread's 16 dG values (single-item threads only)
#pragma unroll
  for (int j = 0; j < 16; ++j) dgs[j] = 0.f;
  // Per-step synchronisation (no CTA-wide barrier inside the loop): the epilogue warps arrive on the grid counter as
  // soon as the next operand is stored; only the TMA producer warp polls it.  The MMA warp is gated by the ring's
  // mbarriers, the epilogue warps by the accumulator mbarrier, and every re-use (TMEM accumulator, receive slots,
  // operand double buffer) is ordered behind the arrival of ALL CTAs for the previous step.
  for (int s = 0; s < nsteps; ++s) {
    const int t = FWD ? s : Tn - 1 - s;              // backward: t = -1 on the extra step that only produces d h_{-1}
    const bool has_rec = FWD ? true : s > 0;
    const int rd_slot = (s + 1) & 1;
    uint8_t* const wr = abuf8 + (size_t)(s & 1) * slot_bytes_g;
    constexpr uint32_t idesc_all = ptx::make_idesc_bf16_f32(64, C::NALL, 0, 0);
    constexpr uint32_t idesc_hi = ptx::make_idesc_bf16_f32(64, C::NC, 0, 0);

    // cell inputs of this thread's first (batch row, unit quad) item: issued now so that their L2 latency hides
    // under the MMAs / cluster barrier (epilogue part 2 consumes them)
    const int items = Bd * 2;
    float pin[16], pc[4], pc2[4], pdc[4], pe[4];
    auto load_inputs = [&](int it) {
      const int b = it >> 1, ub = u0 + (it & 1) * 4;
      if (FWD) {
        const float* g = a.gates + ((int64_t)t * Bd + b) * 4 * nh + ub;
#pragma unroll
        for (int q = 0; q < 4; ++q) {
          const float4 x0 = *(const float4*)(g + q * nh);
          pin[q * 4 + 0] = x0.x; pin[q * 4 + 1] = x0.y; pin[q * 4 + 2] = x0.z; pin[q * 4 + 3] = x0.w;
        }
        if (a.row_bias != nullptr) {
          const float* rb = a.row_bias + (int64_t)b * 4 * nh + ub;
#pragma unroll
          for (int q = 0; q < 4; ++q) {
            const float4 z0 = __ldg((const float4*)(rb + q * nh));
            pin[q * 4 + 0] += z0.x; pin[q * 4 + 1] += z0.y; pin[q * 4 + 2] += z0.z; pin[q * 4 + 3] += z0.w;
          }
        }
        const float* cpp = t ? a.c_all + ((int64_t)(t - 1) * Bd + b) * nh + ub : (a.c0 ? a.c0 + (int64_t)b * nh + ub : nullptr);
#pragma unroll
        for (int j = 0; j < 4; ++j) pc[j] = cpp ? cpp[j] : 0.f;
      } else if (t >= 0) {
        const float* g = a.gates + ((int64_t)t * Bd + b) * 4 * nh + ub;
#pragma unroll
        for (int q = 0; q < 4; ++q) {
          const float4 x0 = *(const float4*)(g + q * nh);
          pin[q * 4 + 0] = x0.x; pin[q * 4 + 1] = x0.y; pin[q * 4 + 2] = x0.z; pin[q * 4 + 3] = x0.w;
        }
        const int64_t o = ((int64_t)t * Bd + b) * nh + ub;
        const float* cpp = t ? a.c_all + o - (int64_t)Bd * nh : (a.c0 ? a.c0 + (int64_t)b * nh + ub : nullptr);
#pragma unroll
        for (int j = 0; j < 4; ++j) {
          pc[j] = a.c_all[o + j];
          pc2[j] = cpp ? cpp[j] : 0.f;
          pdc[j] = a.dc[(int64_t)b * nh + ub + j];
          float e = 0.f;
          if (a.dh_ext) e = a.dh_ext[o + j] * drop_factor(a.drop, ((uint64_t)b * Tn + t) * nh + ub + j);
          if (t == Tn - 1 && a.dh_last) e += a.dh_last[(int64_t)b * nh + ub + j];
          pe[j] = e;
        }
      }
    };
    if (warp == 5) {
      // ------------------------------- producer: this CTA's K slice of the published operand, global (L2) -> ring
      const int nst = KBS * a.m_tiles;
      // While the other CTAs are still publishing: make sure the ring slots this step starts with are free (their MMAs of the
      // previous step committed long ago).  Every mbarrier.try_wait costs ~90 cycles even when the phase is complete; done
      // here they are off the critical path, done after the counter they delayed the first copy by 0.4-0.7 K cycles
      // (probe P9 vs P23a: 2.7 K vs 1.4 K cycles from "seen" to 64 KB landed).
      const int npre = has_rec ? min(nst, a.NS) : 0;
      {
        PipeState q = ps;
        for (int j = 0; j < npre; ++j) {
          if (!a.grouped || (j % a.bulk_stages) == 0) ptx::mbar_wait(sm.empty(q.stage, a.NS), q.phase ^ 1u);
          if (++q.stage == a.NS) { q.stage = 0; q.phase ^= 1u; }
        }
      }
      if (s > 0) {                       // every CTA has published its part of this step's operand
        if (lane == 0) {
          const unsigned target = (unsigned)(s + 1) * gridDim.x;
          // (measured and rejected, profiles/README.md r2c: pipelined polls — 4 relaxed loads in flight, one every 160 cycles —
          // made the step 17 % SLOWER, the extra requests queue in front of the arrivals on the counter's L2 line; sleeping
          // 1-4 us before the first poll changed nothing)
          while (ld_acquire_u32(a.bar) < target) {
          }
        }
        __syncwarp();
      }
      if (trace && lane == 0) a.dbg[s * 8 + 0] = clock64();   // step start = grid counter complete
      if (has_rec) {
        // The operand was written with generic stores by other SMs (ordered before this point by their release and the
        // acquire above) and is read through the async proxy: ONE consumer-side proxy fence, restricted to the global
        // state space (fence.proxy.async over all state spaces: 724 cycles; .global: 20 — profiles/r2b_exchange_probe_d0.txt).
        asm volatile("fence.proxy.async.global;" ::: "memory");
        // The K slice of this CTA is ONE contiguous region of the tile image (k-block major: stage i = kb * m_tiles + mt).  It
        // moves as a few LARGE cp.async.bulk copies of `a.bulk_stages` ring stages each: on this access pattern a bulk copy
        // costs ~1.4 K cycles to land 64 KB but separate 8 KB copies retire only one every ~0.7 K cycles (probe P6 vs
        // P23a).  The first stage of a group carries the group's transaction bytes; the MMA warp consumes stages in order,
        // so it has passed the leader's barrier before it touches a later stage of the group.
        const uint8_t* gsrc = abuf8 + (size_t)rd_slot * slot_bytes_g + (size_t)rank * nst * stage_bytes;
        for (int i = 0; i < nst;) {
          const int glen = min(min(a.bulk_stages, nst - i), a.NS - ps.stage);     // contiguous in the ring, too
          for (int j = 0; j < glen; ++j)
            if (i + j >= npre && (!a.grouped || j == 0))
              ptx::mbar_wait(sm.empty(ps.stage + j, a.NS), ps.phase ^ 1u);   // slot re-used within the step
          if (ptx::elect_one()) {
            ptx::mbar_expect_tx(sm.full(ps.stage), (uint32_t)glen * stage_bytes);
            bulk_g2s(sm.a_base + (uint32_t)ps.stage * stage_bytes, gsrc + (size_t)i * stage_bytes, (uint32_t)glen * stage_bytes,
                     sm.full(ps.stage));
            if (!a.grouped)
              for (int j = 1; j < glen; ++j) ptx::mbar_arrive(sm.full(ps.stage + j));
          }
          __syncwarp();
          i += glen;
          ps.stage += glen;
          if (ps.stage >= a.NS) { ps.stage = 0; ps.phase ^= 1u; }
        }
      }
    } else if (warp == 4) {
      // ------------------------------- MMA issuer: 2 instructions per K sub-step (see mma_pass)
      if (has_rec) {
        // Stages in k-block major order (stage i = kb * m_tiles + mt, the order of the tile image; each m-tile has its own
        // accumulator), consumed in the producer's GROUPS: one barrier wait + one tcgen05 fence per group, then all its MMAs
        // back to back.  tcgen05.mma issue is synchronous with the tensor pipe (no deep queue): every cycle the issuing
        // thread spends between two MMAs is a cycle the pipe idles, and the per-stage wait / fence / elect / commit sequence
        // cost ~250 cycles per 4 MMAs (measured: 127 cycles per M=64 N=128 MMA in the loop against 64 back to back).
        const int nst = KBS * a.m_tiles;
        for (int i = 0; i < nst;) {
          const int glen = min(min(a.bulk_stages, nst - i), a.NS - ps.stage);
          ptx::mbar_wait(sm.full(ps.stage), ps.phase);            // the group leader's barrier carries the group's bytes
          ptx::tc_fence_after();
          if (trace && i == 0 && lane == 0) a.dbg[s * 8 + 1] = clock64();
          if (ptx::elect_one()) {
            // the issue loop is the tensor pipe's feed: keep it to a few integer instructions per MMA.  A descriptor's start
            // address field counts 16-byte units, so the four K sub-steps of a stage (32 bytes apart) are desc + 2 k.
            int kb = i / a.m_tiles, mt = i - kb * a.m_tiles;
            for (int j = 0; j < glen; ++j) {
              const uint32_t d = tmem_base + (uint32_t)(mt * C::NALL);
              const uint32_t sa = sm.a_base + (uint32_t)(ps.stage + j) * 2 * a.part_bytes;
              const uint64_t a_hi0 = ptx::make_smem_desc_sw128(sa, 16, 1024);
              const uint64_t b0 = ptx::make_smem_desc_sw128(sm.w_base + (uint32_t)kb * C::WT, 16, 1024);
              if (stack) {
#pragma unroll
                for (int k = 0; k < 4; ++k) ptx::umma_f16(d, a_hi0 + (uint64_t)(2 * k), b0 + (uint64_t)(2 * k), idesc_all, (kb | k) ? 1u : 0u);
              } else {
                const uint64_t a_lo0 = ptx::make_smem_desc_sw128(sa + a.part_bytes, 16, 1024);
#pragma unroll
                for (int k = 0; k < 4; ++k) {
                  ptx::umma_f16(d, a_hi0 + (uint64_t)(2 * k), b0 + (uint64_t)(2 * k), idesc_all, (kb | k) ? 1u : 0u);
                  ptx::umma_f16(d, a_lo0 + (uint64_t)(2 * k), b0 + (uint64_t)(2 * k), idesc_hi, 1u);
                }
              }
              if (++mt == a.m_tiles) { mt = 0; ++kb; }
            }
            kb = i / a.m_tiles;
            mt = i - kb * a.m_tiles;
            // tcgen05.commit goes through the tensor pipe's queue like an MMA (~70 cycles each, measured): one per GROUP when
            // the partition is step-invariant (only the leaders' barriers are then ever waited on), else one per stage
            for (int j = 0; j < glen; ++j) {
              if (!a.grouped || j == 0) ptx::umma_commit(sm.empty(ps.stage + j, a.NS));
              if (kb == KBS - 1) ptx::umma_commit(sm.acc(mt, a.NS));
              if (++mt == a.m_tiles) { mt = 0; ++kb; }
            }
          }
          __syncwarp();
          i += glen;
          ps.stage += glen;
          if (ps.stage >= a.NS) { ps.stage = 0; ps.phase ^= 1u; }
        }
        if (trace && lane == 0) a.dbg[s * 8 + 2] = clock64();
      }
    } else {
      // =============================== epilogue warps 0-3 ===============================
      if ((int)threadIdx.x < items) load_inputs((int)threadIdx.x);   // L2 latency hides under the MMAs
      if (has_rec) {
        // ---- part 1: partial products TMEM -> registers -> PUSH to the owner's receive slot (DSMEM store)
        const int nrow_mma = stack ? 2 * rows_alloc : rows_alloc;          // MMA rows of an m-tile that carry data
        for (int mt = 0; mt < a.m_tiles; ++mt) {
          ptx::mbar_wait(sm.acc(mt, a.NS), (uint32_t)acc_par);
          ptx::tc_fence_after();
          if (trace && threadIdx.x == 0 && mt == 0) a.dbg[s * 8 + 3] = clock64();
          if (warp * 16 < nrow_mma) {                                      // warp-uniform
            const int rloc = warp * 16 + (lane & 15);                      // MMA row (TMEM lane 32*warp + lane%16)
            const bool lo_row = stack && rloc >= rows_alloc;               // stacked M: rows [ra, 2ra) = A_lo · [W_hi ; W_lo]
            const bool valid = lane < 16 && rloc < nrow_mma;
            const int brow = mt * rows_alloc + (lo_row ? rloc - rows_alloc : rloc);
            const int slot = stack ? 2 * (int)rank + (lo_row ? 1 : 0) : (int)rank;
            const uint32_t dst_local = r_base + (uint32_t)(((slot * rows_tot + brow) * C::RROW) * 4);
            const uint32_t tbase = tmem_base + ((uint32_t)(warp * 32) << 16) + (uint32_t)(mt * C::NALL);
#pragma unroll
            for (int qq = 0; qq < C::CS; ++qq) {
              const uint32_t q = (rank + 1u + (uint32_t)qq) % (uint32_t)C::CS;   // peers first, own columns last
              const uint32_t dst = mapa_u32(dst_local, q);
              if (FWD) {
                uint32_t r1[32], r2[32];
                ptx::tmem_ld32(tbase + q * 32u, r1);
                ptx::tmem_ld32(tbase + (uint32_t)C::NC + q * 32u, r2);
                ptx::tmem_ld_wait();
                if (valid) {
#pragma unroll
                  for (int j = 0; j < 32; j += 4) {
                    float v[4];
#pragma unroll
                    for (int e = 0; e < 4; ++e)
                      v[e] = __uint_as_float(r1[j + e]) + (lo_row ? 0.f : __uint_as_float(r2[j + e]));   // lo·lo dropped
                    st_cluster_f4(dst + (uint32_t)j * 4u, v[0], v[1], v[2], v[3]);
                  }
                }
              } else {
                uint32_t r1[8], r2[8];
                tmem_ld8(tbase + q * 8u, r1);
                tmem_ld8(tbase + (uint32_t)C::NC + q * 8u, r2);
                ptx::tmem_ld_wait();
                if (valid) {
#pragma unroll
                  for (int j = 0; j < 8; j += 4) {
                    float v[4];
#pragma unroll
                    for (int e = 0; e < 4; ++e)
                      v[e] = __uint_as_float(r1[j + e]) + (lo_row ? 0.f : __uint_as_float(r2[j + e]));
                    st_cluster_f4(dst + (uint32_t)j * 4u, v[0], v[1], v[2], v[3]);
                  }
                }
              }
            }
          }
        }
        ptx::tc_fence_before();
        acc_par ^= 1;
        if (trace && threadIdx.x == 0) a.dbg[s * 8 + 5] = clock64();       // own partials pushed
      }

      // ---- part 2: local reduce of the receive slots + LSTM cell, 4 units per thread.  Only the bf16 operand of the
      // next time step is stored before the grid-barrier arrival; everything else (fp32 stashes for the later
      // kernels) is stored after it, off the inter-SM critical path (single-item threads only).
      const bool defer = items <= 128;
      const uint32_t slot_bytes = (uint32_t)(rows_tot * C::RROW * 4);
      if (FWD) {
        float hv[4], cv[4], act[16];
        auto store_rest = [&](int it) {
          const int b = it >> 1, ub = u0 + (it & 1) * 4;
          float* g = a.gates + ((int64_t)t * Bd + b) * 4 * nh + ub;
#pragma unroll
          for (int q = 0; q < 4; ++q)
            *(float4*)(g + q * nh) = make_float4(act[q * 4], act[q * 4 + 1], act[q * 4 + 2], act[q * 4 + 3]);
          const int64_t o = ((int64_t)t * Bd + b) * nh + ub;
          *(float4*)(a.c_all + o) = make_float4(cv[0], cv[1], cv[2], cv[3]);
          *(float4*)(a.h_all + o) = make_float4(hv[0], hv[1], hv[2], hv[3]);
          if (a.hdrop_all) {
            float hd[4];
#pragma unroll
            for (int j = 0; j < 4; ++j) hd[j] = hv[j] * drop_factor(a.drop, ((uint64_t)b * Tn + t) * nh + ub + j);
            *(float4*)(a.hdrop_all + o) = make_float4(hd[0], hd[1], hd[2], hd[3]);
          }
        };
        for (int it = threadIdx.x; it < items; it += 128) {     // item = (batch row, unit quad)
          if (it != (int)threadIdx.x) load_inputs(it);
          const int b = it >> 1, uq = it & 1, ub = u0 + uq * 4;
          const uint32_t rrow = r_base + (uint32_t)((((b >> 6) * rows_alloc + (b & 63)) * C::RROW + uq * 4) * 4);
          float my[16];
          {
            float4 pv[2 * C::CS][4];
            bool ok;
            unsigned spins = 0;
            do {                                  // all loads in flight, then test; retry until every contributor landed
              if (++spins > (1u << 26)) asm volatile("trap;");   // a lost contribution must not hang the device
              ok = true;
#pragma unroll
              for (int sl = 0; sl < 2 * C::CS; ++sl)
                if (sl < nslots) {
#pragma unroll
                  for (int g = 0; g < 4; ++g) pv[sl][g] = ld_shared_volatile_f4(rrow + (uint32_t)sl * slot_bytes + g * 32u);
                }
#pragma unroll
              for (int sl = 0; sl < 2 * C::CS; ++sl)
                if (sl < nslots) {
#pragma unroll
                  for (int g = 0; g < 4; ++g) ok = ok && rx_ready(pv[sl][g]);
                }
            } while (!ok);
#pragma unroll
            for (int g = 0; g < 4; ++g) {
              float4 acc4 = pv[0][g];
#pragma unroll
              for (int sl = 1; sl < 2 * C::CS; ++sl)
                if (sl < nslots) { acc4.x += pv[sl][g].x; acc4.y += pv[sl][g].y; acc4.z += pv[sl][g].z; acc4.w += pv[sl][g].w; }
              my[g * 4 + 0] = acc4.x; my[g * 4 + 1] = acc4.y; my[g * 4 + 2] = acc4.z; my[g * 4 + 3] = acc4.w;
            }
#pragma unroll
            for (int sl = 0; sl < 2 * C::CS; ++sl)
              if (sl < nslots) {
#pragma unroll
                for (int g = 0; g < 4; ++g) st_shared_u4(rrow + (uint32_t)sl * slot_bytes + g * 32u, RX_EMPTY);   // re-arm
              }
          }
          if (trace && threadIdx.x == 0) a.dbg[s * 8 + 6] = clock64();     // every contributor's partial has landed
#pragma unroll
          for (int j = 0; j < 4; ++j) {
            const float ig = fsigmoid(pin[j] + my[j]);
            const float fg = fsigmoid(pin[4 + j] + my[4 + j]);
            const float gg = ftanh(pin[8 + j] + my[8 + j]);
            const float og = fsigmoid(pin[12 + j] + my[12 + j]);
            const float c = fg * pc[j] + ig * gg;
            cv[j] = c;
            hv[j] = og * ftanh(c);
            act[j] = ig; act[4 + j] = fg; act[8 + j] = gg; act[12 + j] = og;
          }
          {                                                       // critical: operand of step s+1 on every SM
            const uint32_t oh = img_off(ub & ~7, b, 0, a.m_tiles, rows_alloc) + (uint32_t)((ub & 4) << 1);
            store_bf16x4_img(wr, oh, oh + (uint32_t)a.part_bytes, hv);
          }
          if (!defer) store_rest(it);
        }
        if (trace && threadIdx.x == 0) a.dbg[s * 8 + 7] = clock64();       // cell done, operand stores issued
        // The operand of the next step is read through the async proxy (cp.async.bulk) on other SMs.  The generic->async proxy
        // fence sits on the CONSUMER side only (TMA warp: acquire of the counter, then fence.proxy.async, then the copies): the
        // release/acquire pair orders the generic stores before the acquire, the consumer's proxy fence orders its later
        // async-proxy reads after it.  A second fence here waited for the store acknowledgements a second time (~1.4 K
        // cycles per time step, profiles/r2a_exchange_probe_d0.txt P0 vs P1; data validated in the probe and by the tests).
        if (a.prod_fence) ptx::fence_proxy_async_all();
        if (trace && threadIdx.x == 0) a.dbg[s * 8 + 4] = clock64();
        asm volatile("bar.sync 1, 128;" ::: "memory");                     // the 4 epilogue warps
        if (threadIdx.x == 0) asm volatile("red.release.gpu.global.add.u32 [%0], 1;" ::"l"(a.bar) : "memory");
        // the release's membar drains the SM's write queue: keep the non-critical stores behind it
        asm volatile("bar.sync 1, 128;" ::: "memory");
        if (defer && (int)threadIdx.x < items) store_rest((int)threadIdx.x);
      } else {
        float dg[16], dcn[4], rec[4];
        auto store_rest = [&](int it) {
          const int b = it >> 1, ub = u0 + (it & 1) * 4;
          if (t < 0) {
            *(float4*)(a.dh_rec_out + (int64_t)b * nh + ub) = make_float4(rec[0], rec[1], rec[2], rec[3]);
          } else {
            const int64_t go_off = ((int64_t)t * Bd + b) * 4 * nh + ub;
            float* go = a.dgates + go_off;
#pragma unroll
            for (int q = 0; q < 4; ++q)
              *(float4*)(go + q * nh) = make_float4(dg[q * 4], dg[q * 4 + 1], dg[q * 4 + 2], dg[q * 4 + 3]);
            if (a.dg_hi) {
#pragma unroll
              for (int q = 0; q < 4; ++q) {
                const float seg[4] = {dg[q * 4], dg[q * 4 + 1], dg[q * 4 + 2], dg[q * 4 + 3]};
                store_bf16x4(a.dg_hi + go_off + q * nh, a.dg_lo + go_off + q * nh, seg);
              }
            }
            if (a.dgsum) {
#pragma unroll
              for (int j = 0; j < 16; ++j) dgs[j] += dg[j];
            }
            *(float4*)(a.dc + (int64_t)b * nh + ub) = make_float4(dcn[0], dcn[1], dcn[2], dcn[3]);
          }
        };
        for (int it = threadIdx.x; it < items; it += 128) {
          if (it != (int)threadIdx.x) load_inputs(it);
          const int b = it >> 1, uq = it & 1, ub = u0 + uq * 4;
          rec[0] = rec[1] = rec[2] = rec[3] = 0.f;
          if (has_rec) {
            const uint32_t rrow = r_base + (uint32_t)((((b >> 6) * rows_alloc + (b & 63)) * C::RROW + uq * 4) * 4);
            float4 pv[2 * C::CS];
            bool ok;
            unsigned spins = 0;
            do {
              if (++spins > (1u << 26)) asm volatile("trap;");
              ok = true;
#pragma unroll
              for (int sl = 0; sl < 2 * C::CS; ++sl)
                if (sl < nslots) pv[sl] = ld_shared_volatile_f4(rrow + (uint32_t)sl * slot_bytes);
#pragma unroll
              for (int sl = 0; sl < 2 * C::CS; ++sl)
                if (sl < nslots) ok = ok && rx_ready(pv[sl]);
            } while (!ok);
#pragma unroll
            for (int sl = 0; sl < 2 * C::CS; ++sl)
              if (sl < nslots) {
                rec[0] += pv[sl].x; rec[1] += pv[sl].y; rec[2] += pv[sl].z; rec[3] += pv[sl].w;
                st_shared_u4(rrow + (uint32_t)sl * slot_bytes, RX_EMPTY);   // re-arm
              }
            if (trace && threadIdx.x == 0) a.dbg[s * 8 + 6] = clock64();
          }
          if (t >= 0) {
#pragma unroll
            for (int j = 0; j < 4; ++j) {
              const float cc = pc[j], cpv = pc2[j], dcv = pdc[j], e = pe[j];
              const float ig = pin[j], fg = pin[4 + j], gg = pin[8 + j], og = pin[12 + j];
              const float dh = rec[j] + e;
              const float tc = ftanh(cc);
              const float dct = dcv + dh * og * (1.f - tc * tc);
              dg[j] = dct * gg * ig * (1.f - ig);
              dg[4 + j] = dct * cpv * fg * (1.f - fg);
              dg[8 + j] = dct * ig * (1.f - gg * gg);
              dg[12 + j] = dh * tc * og * (1.f - og);
              dcn[j] = dct * fg;
            }
#pragma unroll
            for (int q = 0; q < 4; ++q) {       // critical: operand of step s+1 on every SM (K index = gate q * nh + unit)
              float seg[4] = {dg[q * 4], dg[q * 4 + 1], dg[q * 4 + 2], dg[q * 4 + 3]};
              const int k = q * nh + ub;
              const uint32_t oh = img_off(k & ~7, b, 0, a.m_tiles, rows_alloc) + (uint32_t)((k & 4) << 1);
              store_bf16x4_img(wr, oh, oh + (uint32_t)a.part_bytes, seg);
            }
          }
          if (!defer) store_rest(it);
        }
        if (trace && threadIdx.x == 0) a.dbg[s * 8 + 7] = clock64();
        if (a.prod_fence) ptx::fence_proxy_async_all();          // see the forward branch
        if (trace && threadIdx.x == 0) a.dbg[s * 8 + 4] = clock64();
        asm volatile("bar.sync 1, 128;" ::: "memory");
        if (threadIdx.x == 0) asm volatile("red.release.gpu.global.add.u32 [%0], 1;" ::"l"(a.bar) : "memory");
        // the release's membar drains the SM's write queue: keep the non-critical stores behind it
        asm volatile("bar.sync 1, 128;" ::: "memory");
        if (defer && (int)threadIdx.x < items) store_rest((int)threadIdx.x);
      }
    }
  }
  if (!FWD && a.dgsum && warp < 4 && (int)threadIdx.x < Bd * 2) {     // host side guarantees Bd * 2 <= 128 when dgsum is requested
    const int b = (int)threadIdx.x >> 1, ub = u0 + ((int)threadIdx.x & 1) * 4;
#pragma unroll
    for (int q = 0; q < 4; ++q)
      *(float4*)(a.dgsum + (int64_t)b * 4 * nh + q * nh + ub) = make_float4(dgs[q * 4], dgs[q * 4 + 1], dgs[q * 4 + 2], dgs[q * 4 + 3]);
  }
  cluster_sync_all();                              // no CTA may exit while a peer can still read its shared memory
  if (warp == 4) tmem_dealloc_rt(tmem_base, (uint32_t)tmem_cols);
}

}  // namespace

// =================================================================================================
// host side
// =================================================================================================
struct LstmTcState {
  int nh, G, KPf, KPb, max_bd;
  int64_t wf, wb;  // resident weight bytes (forward / backward)
  __nv_bfloat16* abuf;
  unsigned* bar;
  bool configured;
};

constexpr int64_t MISC_BYTES = 1024 /*align slack*/ + 8 * (2 * MAX_NS + MAX_MT + 2) + 64;

// which recurrence kernel the last forward / backward launch used ("v2/cs2", "v1", "steps", ...): exported through
// lagvae_lstm_variant so that tests and bench.py can assert that the intended kernel ran (no silent fallback)
static char g_variant[2][32] = {"none", "none"};
void lstm_note_variant(int dir, const char* what) { snprintf(g_variant[dir & 1], sizeof(g_variant[0]), "%s", what); }
const char* lstm_last_variant(int dir) { return g_variant[dir & 1]; }

static unsigned long long* g_dbg = nullptr;   // optional device trace buffer (lagvae_debug_trace_buffer)
static size_t g_dbg_words = 0;
void lstm_tc_set_debug(void* p, size_t words) {
  g_dbg = (unsigned long long*)p;
  g_dbg_words = words;
}

// ring geometry for a launch: stage = 2 parts (hi, lo) of rows_alloc x 128 B
static bool ring_geometry(int64_t wbytes, int Bd, int KB, int* part_bytes, int* ns, size_t* smem) {
  const int rows_alloc = Bd >= 64 ? 64 : (int)round_up(Bd, 8);
  const int64_t stage = 2 * (int64_t)rows_alloc * 128;
  int64_t n = (SMEM_LIMIT - wbytes - MISC_BYTES) / stage;
  n = std::min<int64_t>(n, MAX_NS);
  n = std::min<int64_t>(n, std::max(KB, 2));
  if (n < 2) return false;
  *part_bytes = rows_alloc * 128;
  *ns = (int)n;
  *smem = (size_t)(wbytes + n * stage + MISC_BYTES);
  return true;
}

static bool shape_supported(const lagvae_text_dims& d) {
  const int nh = d.nh;
  if (nh % 8 != 0 || nh < 64) return false;
  int sms = 0, dev = 0;
  if (cudaGetDevice(&dev) != cudaSuccess) return false;
  cudaDeviceGetAttribute(&sms, cudaDevAttrMultiProcessorCount, dev);
  if (nh / 8 > sms) return false;
  if ((int64_t)d.B * d.ns > 64 * MAX_MT) return false;
  const int KPf = (int)round_up(nh, 64), KPb = (int)round_up(4 * nh, 64);
  const int64_t wf = (int64_t)(KPf / 64) * 2 * 32 * 128, wb = (int64_t)(KPb / 64) * 2 * 8 * 128;
  int pb, ns;
  size_t sm;
  return ring_geometry(wf, 64, KPf / 64, &pb, &ns, &sm) && ring_geometry(wb, 64, KPb / 64, &pb, &ns, &sm);
}

size_t lstm_tc_workspace_bytes(const lagvae_text_dims& d, bool use_tc) {
  if (!use_tc) return 0;
  const int64_t Bd = (int64_t)d.B * d.ns, KPb = round_up(4 * d.nh, 64);
  int64_t rows = round_up(Bd, 64);                                        // v2 tile image: whole row tiles per k-block
  if (Bd < 64) {
    rows = 8;
    while (rows < Bd) rows *= 2;
  }
  return (size_t)(2 * 2 * rows * KPb * 2 + 1024);   // sized for the backward operand (>= forward's)
}

int lstm_tc_create(const lagvae_text_dims& d, bool use_tc, void* ws, size_t ws_bytes, LstmTcState** out) {
  *out = nullptr;
  if (!use_tc) return LAGVAE_OK;
  if (!shape_supported(d)) return LAGVAE_OK;
  if (ws_bytes < lstm_tc_workspace_bytes(d, true)) {
    set_error("lstm_tc: workspace too small");
    return LAGVAE_E_WORKSPACE;
  }
  LstmTcState* s = new (std::nothrow) LstmTcState{};
  LV_CHECK_ARG(s != nullptr, "lstm_tc: host allocation failed");
  s->nh = d.nh;
  s->G = d.nh / 8;
  s->KPf = (int)round_up(d.nh, 64);
  s->KPb = (int)round_up(4 * d.nh, 64);
  s->wf = (int64_t)(s->KPf / 64) * 2 * 32 * 128;
  s->wb = (int64_t)(s->KPb / 64) * 2 * 8 * 128;
  s->max_bd = d.B * d.ns;
  s->bar = (unsigned*)ws;
  s->abuf = (__nv_bfloat16*)((char*)ws + 1024);
  s->configured = false;
  *out = s;
  return LAGVAE_OK;
}

void lstm_tc_destroy(LstmTcState* s) { delete s; }

static int configure(LstmTcState* s) {
  if (s->configured) return LAGVAE_OK;
  LV_CUDA(cudaFuncSetAttribute(k_lstm_fwd_tc, cudaFuncAttributeMaxDynamicSharedMemorySize, SMEM_LIMIT));
  LV_CUDA(cudaFuncSetAttribute(k_lstm_bwd_tc, cudaFuncAttributeMaxDynamicSharedMemorySize, SMEM_LIMIT));
  s->configured = true;
  return LAGVAE_OK;
}

static int prod_fence_env() {
  static const int v = [] { const char* e = getenv("LAGVAE_LSTM_PROD_FENCE"); return (e && e[0] == '1') ? 1 : 0; }();
  return v;
}

static int bulk_stages_env() {
  static const int v = [] { const char* e = getenv("LAGVAE_LSTM_BULK_STAGES"); const int n = e ? atoi(e) : 4; return n < 1 ? 1 : (n > 16 ? 16 : n); }();
  return v;
}

// ---- v2 (cluster K-split) geometry / launch ---------------------------------------------------------------
template <bool FWD, int CS_>
static bool v2_geometry(const LstmTcState* s, int Bd, RecArgs* a, size_t* smem) {
  using C = V2Cfg<FWD, CS_>;
  static const bool off = [] { const char* e = getenv("LAGVAE_LSTM_V1"); return e && e[0] == '1'; }();
  if (off) return false;
  const int nh = s->nh;
  const int K = FWD ? nh : 4 * nh;
  if (nh % (8 * C::CS) != 0 || K % (64 * C::CS) != 0 || nh < 256) return false;
  const int m_tiles = (int)cdiv(Bd, 64);
  if (m_tiles * C::NALL > 512) return false;
  const int KBS = K / 64 / C::CS;
  int rows_alloc = 64;                                   // rows of one operand part of a stage: a power of two (8..64), so that
  if (Bd < 64) {                                         // a stage is 1, 2, 4 or 8 quarters of 2 KB for the operand copy
    rows_alloc = 8;
    while (rows_alloc < Bd) rows_alloc *= 2;
  }
  const int64_t stage = 2 * (int64_t)rows_alloc * 128;
  const int64_t wbytes = (int64_t)KBS * C::WT;
  const bool stack = m_tiles == 1 && 2 * rows_alloc <= 64;
  const int64_t pbytes = (int64_t)C::CS * (stack ? 2 : 1) * m_tiles * rows_alloc * C::RROW * 4;   // receive slots
  int64_t n = (SMEM_LIMIT - wbytes - pbytes - MISC_BYTES) / stage;
  n = std::min<int64_t>(n, MAX_NS);
  n = std::min<int64_t>(n, std::max(KBS * m_tiles, 2));
  if (n < 2) return false;
  // the loaders keep a group of stages in registers (16 chunks of 16 B per thread = floor(16 / (stage / 2 KB)) stages) and
  // store it before arriving on any of its barriers: the ring must hold a whole group

  const int bulk = bulk_stages_env();
  const int nst_step = KBS * m_tiles;
  if (n >= bulk && nst_step % bulk == 0) n -= n % bulk;      // step-invariant group partition (see RecArgs::grouped)
  a->grouped = (n % bulk == 0 && nst_step % bulk == 0) ? 1 : 0;
  a->part_bytes = rows_alloc * 128;
  a->NS = (int)n;
  a->KP = K;
  a->KB = K / 64;
  a->m_tiles = m_tiles;
  *smem = (size_t)(wbytes + pbytes + n * stage + MISC_BYTES);
  return true;
}

template <bool FWD, int CS_>
static int v2_launch(const LstmTcState* s, const RecArgs& a, const TMaps& tm, size_t smem, cudaStream_t st, bool* launched) {
  using C = V2Cfg<FWD, CS_>;
  // per-device decision, made ONCE from an occupancy query (no trial launches): 1 = all clusters of this size are
  // co-resident (the per-step grid barrier needs that), -1 = they are not (the caller picks another cluster size).
  // After a positive decision a failing launch is an ERROR: nothing falls back silently (lagvae_lstm_variant reports
  // what ran).  The kernel spins on a grid-wide counter, so the launch is always cooperative.
  // The decision depends on the shared-memory footprint (a small nh leaves room for two CTAs per SM, nh = 1024 does not):
  // it is re-made whenever the footprint differs from the one it was made for.
  static int state[16] = {0};
  static size_t state_smem[16] = {0};
  int dev = 0;
  LV_CUDA(cudaGetDevice(&dev));
  int& stt = state[dev & 15];
  if (state_smem[dev & 15] != smem) {
    state_smem[dev & 15] = smem;
    stt = 0;
  }
  *launched = false;
  if (stt < 0) return LAGVAE_OK;
  auto kern = k_lstm_v2<FWD, CS_>;
  cudaLaunchConfig_t cfg{};
  cfg.gridDim = dim3(s->G);
  cfg.blockDim = dim3(NTHREADS);
  cfg.dynamicSmemBytes = smem;
  cfg.stream = st;
  cudaLaunchAttribute at[2];
  at[0].id = cudaLaunchAttributeClusterDimension;
  at[0].val.clusterDim.x = C::CS;
  at[0].val.clusterDim.y = 1;
  at[0].val.clusterDim.z = 1;
  at[1].id = cudaLaunchAttributeCooperative;
  at[1].val.cooperative = 1;
  cfg.attrs = at;
  if (stt == 0) {
    if (cudaFuncSetAttribute(kern, cudaFuncAttributeMaxDynamicSharedMemorySize, SMEM_LIMIT) != cudaSuccess) {
      cudaGetLastError();
      stt = -1;
      return LAGVAE_OK;
    }
    int ncl = 0;
    cfg.numAttrs = 1;
    if (cudaOccupancyMaxActiveClusters(&ncl, kern, &cfg) != cudaSuccess || ncl * C::CS < s->G) {
      cudaGetLastError();
      stt = -1;
      return LAGVAE_OK;
    }
    stt = 1;
  }
  cfg.numAttrs = 2;
  (void)tm;   // v2 reads its operand with cp.async.bulk from the tile-image layout: no tensor maps
  cudaError_t e = cudaLaunchKernelEx(&cfg, kern, a);
  if (e != cudaSuccess) {
    cudaGetLastError();
    set_error("lstm v2 (%s, cluster of %d) cooperative launch failed: %s", FWD ? "forward" : "backward", C::CS,
              cudaGetErrorString(e));
    return LAGVAE_E_CUDA;
  }
  *launched = true;
  char nm[32];
  snprintf(nm, sizeof(nm), "v2/cs%d", C::CS);
  lstm_note_variant(FWD ? 0 : 1, nm);
  g_launches.fetch_add(1);
  return LAGVAE_OK;
}

static int make_maps(const LstmTcState* s, int Bd, int KP, int part_bytes, TMaps* tm) {
  for (int slot = 0; slot < 2; ++slot)
    for (int part = 0; part < 2; ++part) {
      const __nv_bfloat16* base = s->abuf + ((int64_t)slot * 2 + part) * Bd * KP;
      LV_TRY(make_tmap_bf16_2d(&tm->m[slot * 2 + part], base, (uint64_t)Bd, (uint64_t)KP, (uint64_t)KP, 64,
                               (uint32_t)(part_bytes / 128)));
    }
  return LAGVAE_OK;
}

int lstm_tc_forward(LstmTcState* s, const float* w_hh, const float* h0, const float* c0, float* gates,
                    float* c_all, float* h_all, float* hdrop_all, DropSpec drop, int Tn, int Bd,
                    cudaStream_t st, const float* row_bias) {
  LV_CHECK_ARG(s && Bd <= s->max_bd && Tn > 0, "lstm_tc_forward: bad arguments");
  LV_TRY(configure(s));
  RecArgs a{};
  a.nh = s->nh; a.Bd = Bd; a.Tn = Tn;
  size_t smem = 0;
  // cluster of 2 when the stacked-M mode applies (Bd <= 32: same MMA time as a cluster of 4, half the partial-sum
  // traffic), cluster of 4 otherwise
  static const int force_cs = [] { const char* e = getenv("LAGVAE_LSTM_FWD_CS"); return e ? atoi(e) : 0; }();
  const int want_cs = force_cs ? force_cs : (Bd <= 32 ? 2 : 4);
  bool v2 = false;
  int v2cs = 0;
  if (want_cs == 2) { v2 = v2_geometry<true, 2>(s, Bd, &a, &smem); v2cs = v2 ? 2 : 0; }
  if (!v2) { v2 = v2_geometry<true, 4>(s, Bd, &a, &smem); v2cs = v2 ? 4 : 0; }
  // clusters of 4 need CS x rows x 144 B of receive slots next to the 128 KB of resident weights: beyond 64 rows that no
  // longer fits, clusters of 2 (half the receive slots) still do up to 128 rows — and beat the non-cluster v1 kernel by far
  if (!v2 && want_cs != 2) { v2 = v2_geometry<true, 2>(s, Bd, &a, &smem); v2cs = v2 ? 2 : 0; }
  if (!v2) {
    a.KP = s->KPf; a.KB = s->KPf / 64;
    a.m_tiles = (int)cdiv(Bd, 64);
    LV_CHECK_ARG(ring_geometry(s->wf, Bd, a.KB, &a.part_bytes, &a.NS, &smem), "lstm_tc_forward: no ring geometry");
  }
  a.prod_fence = prod_fence_env();
  a.bulk_stages = bulk_stages_env();
  a.dbg = (g_dbg && g_dbg_words >= (size_t)Tn * 8) ? g_dbg : nullptr;
  a.bar = s->bar; a.abuf = s->abuf; a.w_hh = w_hh; a.h0 = h0; a.c0 = c0; a.gates = gates; a.c_all = c_all;
  a.h_all = h_all; a.hdrop_all = hdrop_all; a.drop = drop;
  a.row_bias = v2 ? row_bias : nullptr;          // the cluster kernel adds it as it loads the pre-activations
  LV_CUDA(cudaMemsetAsync(s->bar, 0, 256, st));
  // the K padding columns of the streamed buffer must be zero
  if (s->KPf != s->nh) LV_CUDA(cudaMemsetAsync(s->abuf, 0, (size_t)4 * Bd * s->KPf * 2, st));
  TMaps tm;
  LV_TRY(make_maps(s, Bd, a.KP, a.part_bytes, &tm));
  if (v2) {
    bool launched = false;
    if (v2cs == 2) {
      LV_TRY((v2_launch<true, 2>(s, a, tm, smem, st, &launched)));
      if (!launched && v2_geometry<true, 4>(s, Bd, &a, &smem)) {
        LV_TRY(make_maps(s, Bd, a.KP, a.part_bytes, &tm));
        LV_TRY((v2_launch<true, 4>(s, a, tm, smem, st, &launched)));
      }
    } else {
      LV_TRY((v2_launch<true, 4>(s, a, tm, smem, st, &launched)));
    }
    if (launched) return LAGVAE_OK;
    a.KP = s->KPf; a.KB = s->KPf / 64;   // cluster launch unavailable: v1 geometry
    a.m_tiles = (int)cdiv(Bd, 64);
    LV_CHECK_ARG(ring_geometry(s->wf, Bd, a.KB, &a.part_bytes, &a.NS, &smem), "lstm_tc_forward: no ring geometry");
    LV_TRY(make_maps(s, Bd, a.KP, a.part_bytes, &tm));
  }
  a.row_bias = nullptr;                          // v1 does not know it: one streaming pass first
  if (row_bias) LV_TRY(add_row_periodic(gates, row_bias, (int64_t)Tn * Bd, 4 * s->nh, Bd, st));
  void* args[] = {(void*)&a, (void*)&tm};
  LV_CUDA(cudaLaunchCooperativeKernel((const void*)k_lstm_fwd_tc, dim3(s->G), dim3(NTHREADS), args, smem, st));
  lstm_note_variant(0, "v1");
  g_launches.fetch_add(1);
  return LAGVAE_OK;
}

int lstm_tc_backward(LstmTcState* s, const float* w_hh, const float* c0, const float* gates,
                     const float* c_all, const float* dh_ext, DropSpec drop, const float* dh_last, float* dc,
                     float* dh_rec, float* dgates, int Tn, int Bd, bool want_init, cudaStream_t st, float* dgsum,
                     uint16_t* dg_hi, uint16_t* dg_lo, bool* extras_done, unsigned* started) {
  if (extras_done) *extras_done = false;
  LV_CHECK_ARG(s && Bd <= s->max_bd && Tn > 0, "lstm_tc_backward: bad arguments");
  LV_TRY(configure(s));
  RecArgs a{};
  a.nh = s->nh; a.Bd = Bd; a.Tn = Tn;
  size_t smem = 0;
  static int cs_ok = 8;   // (process-wide; one device model per process) 16 clusters of 8 CTAs do not fit on every B200 (GPC sizes): fall back to clusters of 4
  bool v2 = cs_ok == 8 && v2_geometry<false, 8>(s, Bd, &a, &smem);
  int v2cs = v2 ? 8 : 0;
  if (!v2 && cs_ok >= 4) {
    v2 = v2_geometry<false, 4>(s, Bd, &a, &smem);
    v2cs = v2 ? 4 : 0;
  }
  if (!v2) {
    a.KP = s->KPb; a.KB = s->KPb / 64;
    a.m_tiles = (int)cdiv(Bd, 64);
    LV_CHECK_ARG(ring_geometry(s->wb, Bd, a.KB, &a.part_bytes, &a.NS, &smem), "lstm_tc_backward: no ring geometry");
  }
  a.prod_fence = prod_fence_env();
  a.bulk_stages = bulk_stages_env();
  a.dbg = (g_dbg && g_dbg_words >= (size_t)(Tn + 1) * 8) ? g_dbg : nullptr;
  a.bar = s->bar; a.abuf = s->abuf; a.w_hh = w_hh; a.c0 = c0; a.gates = const_cast<float*>(gates);
  a.c_all = const_cast<float*>(c_all); a.dh_ext = dh_ext; a.drop = drop; a.dh_last = dh_last; a.dc = dc;
  a.dh_rec_out = dh_rec; a.dgates = dgates; a.want_init = want_init ? 1 : 0;
  // the in-kernel extras (time sum of dG, dG as bf16 hi/lo operand) exist in the cluster kernel for single-item threads only
  const bool extras = v2 && extras_done != nullptr && Bd * 2 <= 128 && dgsum != nullptr && dg_hi != nullptr && dg_lo != nullptr;
  a.started = v2 ? started : nullptr;
  a.dgsum = extras ? dgsum : nullptr;
  a.dg_hi = extras ? (__nv_bfloat16*)dg_hi : nullptr;
  a.dg_lo = extras ? (__nv_bfloat16*)dg_lo : nullptr;
  LV_CUDA(cudaMemsetAsync(s->bar, 0, 256, st));
  if (s->KPb != 4 * s->nh) LV_CUDA(cudaMemsetAsync(s->abuf, 0, (size_t)4 * Bd * s->KPb * 2, st));
  TMaps tm;
  LV_TRY(make_maps(s, Bd, a.KP, a.part_bytes, &tm));
  if (v2) {
    bool launched = false;
    if (v2cs == 8) {
      LV_TRY((v2_launch<false, 8>(s, a, tm, smem, st, &launched)));
      if (launched) { if (extras) *extras_done = true; return LAGVAE_OK; }
      cs_ok = 4;
      if (v2_geometry<false, 4>(s, Bd, &a, &smem)) {
        LV_TRY(make_maps(s, Bd, a.KP, a.part_bytes, &tm));
        LV_TRY((v2_launch<false, 4>(s, a, tm, smem, st, &launched)));
        if (launched) { if (extras) *extras_done = true; return LAGVAE_OK; }
      }
    } else {
      LV_TRY((v2_launch<false, 4>(s, a, tm, smem, st, &launched)));
      if (launched) { if (extras) *extras_done = true; return LAGVAE_OK; }
    }
    // no cluster size fits THIS footprint: v1 for this call only (another shape may fit again; v2_launch caches its own decision)
    a.dgsum = nullptr; a.dg_hi = a.dg_lo = nullptr;
    a.KP = s->KPb; a.KB = s->KPb / 64;
    a.m_tiles = (int)cdiv(Bd, 64);
    LV_CHECK_ARG(ring_geometry(s->wb, Bd, a.KB, &a.part_bytes, &a.NS, &smem), "lstm_tc_backward: no ring geometry");
    LV_TRY(make_maps(s, Bd, a.KP, a.part_bytes, &tm));
  }
  void* args[] = {(void*)&a, (void*)&tm};
  LV_CUDA(cudaLaunchCooperativeKernel((const void*)k_lstm_bwd_tc, dim3(s->G), dim3(NTHREADS), args, smem, st));
  lstm_note_variant(1, "v1");
  g_launches.fetch_add(1);
  return LAGVAE_OK;
}

}  // namespace lagvae
