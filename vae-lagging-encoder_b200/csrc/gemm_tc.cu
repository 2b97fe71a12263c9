// Split-bf16 tensor-core GEMM on tcgen05 (sm_100a).
//
//   C[M,N] = alpha * Σ_k A(m,k) B(n,k) (+ beta C) (+ bias),   A = A_hi + A_lo,  B = B_hi + B_lo (bf16)
//   passes = 3:  A_hi·B_hi + A_hi·B_lo + A_lo·B_hi   (relative error ~2^-16 per product: fp32-grade,
//                needed for the 1e-4 parity bar, SURVEY §7 hard part 2)
//   passes = 1:  A_hi·B_hi
//
// Persistent warp-specialised kernel, one CTA per SM:
//   warp 0      TMA producer   (cp.async.bulk.tensor 2-D, SWIZZLE_128B, 3-stage mbarrier ring)
//   warp 1      MMA issuer     (one lane issues tcgen05.mma 128x128x16, fp32 accumulators in TMEM,
//                               double-buffered: 2 x 128 columns)
//   warps 2-5   epilogue       (tcgen05.ld 32x32b.x32 -> registers -> bias/alpha/beta -> global)
// Both operands may be K-major (row-major [rows, K]) or MN-major (row-major [K, rows]); the latter is
// what every weight-gradient GEMM (dW = dYᵀ·X) needs, so no transposes are ever materialised.
// Replaces the cuBLAS sgemm calls under nn.Linear / nn.LSTM projections of the reference
// (dec_lstm.py:109, enc_lstm.py:60, dec_lstm.py:104 and their autograd backward, text.py:384).
#include "lagvae_common.cuh"
#include "sm100_ptx.cuh"

#include <cstdlib>
#include <mutex>

namespace lagvae {

// ---------------------------------------------------------------------------------------------
// host: tensor maps
// ---------------------------------------------------------------------------------------------
PFN_encodeTiled get_encode_tiled() {
  static PFN_encodeTiled fn = nullptr;
  static std::once_flag once;
  std::call_once(once, [] {
    void* p = nullptr;
    cudaDriverEntryPointQueryResult q;
    if (cudaGetDriverEntryPoint("cuTensorMapEncodeTiled", &p, cudaEnableDefault, &q) == cudaSuccess &&
        q == cudaDriverEntryPointSuccess)
      fn = (PFN_encodeTiled)p;
  });
  return fn;
}

int make_tmap_bf16_2d(CUtensorMap* out, const void* base, uint64_t rows, uint64_t cols, uint64_t ld,
                      uint32_t box_cols, uint32_t box_rows) {
  PFN_encodeTiled enc = get_encode_tiled();
  if (!enc) {
    set_error("cuTensorMapEncodeTiled entry point unavailable");
    return LAGVAE_E_CUDA;
  }
  LV_CHECK_ARG(((uintptr_t)base & 15) == 0 && (ld % 8) == 0, "tensor map: base must be 16-B aligned, ld %% 8 == 0");
  cuuint64_t gdim[2] = {cols, rows};
  cuuint64_t gstr[1] = {ld * 2};
  cuuint32_t box[2] = {box_cols, box_rows};
  cuuint32_t estr[2] = {1, 1};
  CUresult r = enc(out, CU_TENSOR_MAP_DATA_TYPE_BFLOAT16, 2, const_cast<void*>(base), gdim, gstr, box, estr,
                   CU_TENSOR_MAP_INTERLEAVE_NONE, CU_TENSOR_MAP_SWIZZLE_128B, CU_TENSOR_MAP_L2_PROMOTION_L2_256B,
                   CU_TENSOR_MAP_FLOAT_OOB_FILL_NONE);
  if (r != CUDA_SUCCESS) {
    set_error("cuTensorMapEncodeTiled failed (%d) rows=%llu cols=%llu ld=%llu", (int)r,
              (unsigned long long)rows, (unsigned long long)cols, (unsigned long long)ld);
    return LAGVAE_E_CUDA;
  }
  return LAGVAE_OK;
}

// ---------------------------------------------------------------------------------------------
// device kernel
// ---------------------------------------------------------------------------------------------
namespace {

constexpr int BM = 128, BK = 64, UK = 16;
constexpr int TILE_BYTES = BM * BK * 2;            // 16 KiB: one 128-row operand part
// Two tile shapes: 128x128 (3-stage ring of 64 KiB) and 128x256 (2-stage ring of 96 KiB).  The split-bf16 operands
// (hi + lo) make these GEMMs L2->SM bandwidth bound (profiles/README.md): the wide tile moves 25% fewer operand bytes
// per MMA and is used whenever the problem still yields >= 2 waves of tiles.
template <int BN>
struct TileCfg {
  static constexpr int STAGES = BN == 128 ? 3 : 2;
  static constexpr int B_PART = BN * BK * 2;                         // one B operand part
  static constexpr int STAGE_BYTES = 2 * TILE_BYTES + 2 * B_PART;    // A_hi A_lo B_hi B_lo
  static constexpr int TMEM_COLS = 2 * BN;                           // double-buffered accumulator
};
constexpr int CSTRIDE = 36;                        // floats per staged row (32 + 4 pad: conflict-free float4 both ways)
constexpr int CSTAGE_BYTES = 4 * 32 * CSTRIDE * 4; // epilogue staging: 4 warps x 32 rows
template <int BN>
constexpr int smem_bytes() { return TileCfg<BN>::STAGES * TileCfg<BN>::STAGE_BYTES + 1024 /*align slack*/ + 256 /*barriers*/ + CSTAGE_BYTES; }
constexpr int NTHREADS = 192;

struct GemmArgs {
  float* C;
  int64_t ldc;
  int M, N, K;
  int passes;
  float alpha, beta;
  const float* bias_n;
  const float* bias_rows;
  int bias_period;
  const int32_t* row_map;
};

template <bool A_MN, bool B_MN, int BN>
__global__ void __launch_bounds__(NTHREADS, 1)
k_gemm_tc(const __grid_constant__ CUtensorMap tm_a_hi, const __grid_constant__ CUtensorMap tm_a_lo,
          const __grid_constant__ CUtensorMap tm_b_hi, const __grid_constant__ CUtensorMap tm_b_lo,
          const GemmArgs g) {
  constexpr int STAGES = TileCfg<BN>::STAGES, STAGE_BYTES = TileCfg<BN>::STAGE_BYTES, B_PART = TileCfg<BN>::B_PART;
  constexpr int TMEM_COLS = TileCfg<BN>::TMEM_COLS;
  extern __shared__ uint8_t smem_raw[];
  const uint32_t smem_base = (ptx::smem_u32(smem_raw) + 1023u) & ~1023u;  // SWIZZLE_128B: 1024-B aligned
  const uint32_t bar_base = smem_base + STAGES * STAGE_BYTES;
  // barriers: full[3] empty[3] tfull[2] tempty[2] ; then tmem ptr
  auto full_bar = [&](int s) { return bar_base + 8u * s; };
  auto empty_bar = [&](int s) { return bar_base + 8u * (STAGES + s); };
  auto tfull_bar = [&](int b) { return bar_base + 8u * (2 * STAGES + b); };
  auto tempty_bar = [&](int b) { return bar_base + 8u * (2 * STAGES + 2 + b); };
  const uint32_t tmem_slot = bar_base + 8u * (2 * STAGES + 4);
  uint32_t* tmem_slot_ptr = (uint32_t*)(smem_raw + (tmem_slot - ptx::smem_u32(smem_raw)));

  const int warp = threadIdx.x >> 5, lane = threadIdx.x & 31;
  const int m_blks = (g.M + BM - 1) / BM, n_blks = (g.N + BN - 1) / BN;
  const int num_tiles = m_blks * n_blks;
  const int k_blks = (g.K + BK - 1) / BK;
  const bool three = g.passes == 3;

  if (warp == 0 && lane == 0) {
    ptx::prefetch_tmap(&tm_a_hi);
    ptx::prefetch_tmap(&tm_b_hi);
    if (three) {
      ptx::prefetch_tmap(&tm_a_lo);
      ptx::prefetch_tmap(&tm_b_lo);
    }
    for (int s = 0; s < STAGES; ++s) {
      ptx::mbar_init(full_bar(s), 1);
      ptx::mbar_init(empty_bar(s), 1);
    }
    for (int b = 0; b < 2; ++b) {
      ptx::mbar_init(tfull_bar(b), 1);
      ptx::mbar_init(tempty_bar(b), 4);  // one arrival per epilogue warp
    }
    ptx::fence_mbar_init();
  }
  if (warp == 1) ptx::tmem_alloc<TMEM_COLS>(tmem_slot);
  ptx::tc_fence_before();
  __syncthreads();
  ptx::tc_fence_after();
  const uint32_t tmem_base = *tmem_slot_ptr;

  if (warp == 0) {
    // ================================ TMA producer ================================
    // the whole warp runs the (uniform) loop; one elected lane issues the copies
    {
      int stage = 0;
      uint32_t phase = 0;
      const uint32_t tx_bytes = (three ? 2u : 1u) * (uint32_t)(TILE_BYTES + B_PART);
      for (int tile = blockIdx.x; tile < num_tiles; tile += gridDim.x) {
        const int mb = tile % m_blks, nb = tile / m_blks;
        const int m0 = mb * BM, n0 = nb * BN;
        for (int kb = 0; kb < k_blks; ++kb) {
          ptx::mbar_wait(empty_bar(stage), phase ^ 1u);
          const uint32_t sa_hi = smem_base + stage * STAGE_BYTES, sa_lo = sa_hi + TILE_BYTES;
          const uint32_t sb_hi = sa_hi + 2 * TILE_BYTES, sb_lo = sb_hi + B_PART;
          const int k0 = kb * BK;
          if (ptx::elect_one()) {
            ptx::mbar_expect_tx(full_bar(stage), tx_bytes);
            if (!A_MN) {
              ptx::tma_load_2d(sa_hi, &tm_a_hi, full_bar(stage), k0, m0);
              if (three) ptx::tma_load_2d(sa_lo, &tm_a_lo, full_bar(stage), k0, m0);
            } else {  // stored [K, M]: two 64(M) x 64(K) boxes
              ptx::tma_load_2d(sa_hi, &tm_a_hi, full_bar(stage), m0, k0);
              ptx::tma_load_2d(sa_hi + TILE_BYTES / 2, &tm_a_hi, full_bar(stage), m0 + 64, k0);
              if (three) {
                ptx::tma_load_2d(sa_lo, &tm_a_lo, full_bar(stage), m0, k0);
                ptx::tma_load_2d(sa_lo + TILE_BYTES / 2, &tm_a_lo, full_bar(stage), m0 + 64, k0);
              }
            }
            if (!B_MN) {
              ptx::tma_load_2d(sb_hi, &tm_b_hi, full_bar(stage), k0, n0);
              if (three) ptx::tma_load_2d(sb_lo, &tm_b_lo, full_bar(stage), k0, n0);
            } else {   // stored [K, N]: BN/64 boxes of 64(N) x 64(K)
#pragma unroll
              for (int j = 0; j < BN / 64; ++j) {
                ptx::tma_load_2d(sb_hi + j * 8192, &tm_b_hi, full_bar(stage), n0 + 64 * j, k0);
                if (three) ptx::tma_load_2d(sb_lo + j * 8192, &tm_b_lo, full_bar(stage), n0 + 64 * j, k0);
              }
            }
          }
          __syncwarp();
          if (++stage == STAGES) { stage = 0; phase ^= 1u; }
        }
      }
    }
  } else if (warp == 1) {
    // ================================ MMA issuer ================================
    // warp-uniform loop; one elected lane issues tcgen05.mma / tcgen05.commit
    {
      constexpr uint32_t idesc = ptx::make_idesc_bf16_f32(BM, BN, A_MN ? 1 : 0, B_MN ? 1 : 0);
      // K-major: 8-row groups 1024 B apart, +32 B per UMMA_K inside the 128-B swizzle row.
      // MN-major: 64-element chunks 8192 B apart (LBO), 8-row K groups 1024 B apart (SBO), +2048 B per UMMA_K.
      constexpr uint32_t A_LBO = A_MN ? 8192u : 16u, A_SBO = 1024u, A_KSTEP = A_MN ? 2048u : 32u;
      constexpr uint32_t B_LBO = B_MN ? 8192u : 16u, B_SBO = 1024u, B_KSTEP = B_MN ? 2048u : 32u;
      int stage = 0;
      uint32_t phase = 0;
      int acc = 0;
      uint32_t acc_phase = 0;
      for (int tile = blockIdx.x; tile < num_tiles; tile += gridDim.x) {
        ptx::mbar_wait(tempty_bar(acc), acc_phase ^ 1u);
        ptx::tc_fence_after();
        const uint32_t d_tmem = tmem_base + (uint32_t)(acc * BN);
        for (int kb = 0; kb < k_blks; ++kb) {
          ptx::mbar_wait(full_bar(stage), phase);
          ptx::tc_fence_after();
          const uint32_t sa_hi = smem_base + stage * STAGE_BYTES, sa_lo = sa_hi + TILE_BYTES;
          const uint32_t sb_hi = sa_hi + 2 * TILE_BYTES, sb_lo = sb_hi + B_PART;
          if (ptx::elect_one()) {
#pragma unroll
            for (int k = 0; k < BK / UK; ++k) {
              const uint64_t da_hi = ptx::make_smem_desc_sw128(sa_hi + k * A_KSTEP, A_LBO, A_SBO);
              const uint64_t db_hi = ptx::make_smem_desc_sw128(sb_hi + k * B_KSTEP, B_LBO, B_SBO);
              ptx::umma_f16(d_tmem, da_hi, db_hi, idesc, (kb | k) ? 1u : 0u);
              if (three) {
                const uint64_t da_lo = ptx::make_smem_desc_sw128(sa_lo + k * A_KSTEP, A_LBO, A_SBO);
                const uint64_t db_lo = ptx::make_smem_desc_sw128(sb_lo + k * B_KSTEP, B_LBO, B_SBO);
                ptx::umma_f16(d_tmem, da_hi, db_lo, idesc, 1u);
                ptx::umma_f16(d_tmem, da_lo, db_hi, idesc, 1u);
              }
            }
            ptx::umma_commit(empty_bar(stage));  // smem slot free once these MMAs retire
            if (kb == k_blks - 1) ptx::umma_commit(tfull_bar(acc));  // accumulator complete
          }
          __syncwarp();
          if (++stage == STAGES) { stage = 0; phase ^= 1u; }
        }
        if (++acc == 2) { acc = 0; acc_phase ^= 1u; }
      }
    }
  } else {
    // ================================ epilogue ================================
    // TMEM -> registers (thread = accumulator row) -> per-warp smem transpose -> row-contiguous global stores:
    // every store instruction of a warp covers 4 rows x 128 B (vector path) or 1 row x 128 B (unaligned C).
    const int quad = warp & 3;  // TMEM lane quadrant this warp may read
    float* cst = (float*)(smem_raw + (bar_base + 256 - ptx::smem_u32(smem_raw))) + (warp - 2) * 32 * CSTRIDE;
    const bool vec_ok = ((g.ldc & 3) == 0) && ((((uintptr_t)g.C) & 15) == 0) && ((((uintptr_t)g.bias_n) & 15) == 0) &&
                        (g.bias_rows == nullptr || (((g.N & 3) == 0) && ((((uintptr_t)g.bias_rows) & 15) == 0)));
    int acc = 0;
    uint32_t acc_phase = 0;
    for (int tile = blockIdx.x; tile < num_tiles; tile += gridDim.x) {
      const int mb = tile % m_blks, nb = tile / m_blks;
      const int row0 = mb * BM + quad * 32;
      const int n0 = nb * BN;
      ptx::mbar_wait(tfull_bar(acc), acc_phase);
      ptx::tc_fence_after();
#pragma unroll 1
      for (int c = 0; c < BN / 32; ++c) {
        uint32_t r[32];
        ptx::tmem_ld32(tmem_base + ((uint32_t)(quad * 32) << 16) + (uint32_t)(acc * BN + c * 32), r);
        ptx::tmem_ld_wait();
        float* mine = cst + lane * CSTRIDE;
#pragma unroll
        for (int j = 0; j < 32; j += 4)
          *(float4*)(mine + j) = make_float4(g.alpha * __uint_as_float(r[j]), g.alpha * __uint_as_float(r[j + 1]),
                                             g.alpha * __uint_as_float(r[j + 2]), g.alpha * __uint_as_float(r[j + 3]));
        __syncwarp();
        const int cb = n0 + c * 32;
        if (vec_ok) {
          const int cc = (lane & 7) * 4, col = cb + cc;
#pragma unroll
          for (int i = 0; i < 8; ++i) {
            const int rr = i * 4 + (lane >> 3), row = row0 + rr;
            if (row < g.M && col < g.N) {
              float4 v = *(const float4*)(cst + rr * CSTRIDE + cc);
              float* vp = &v.x;
              const int64_t orow = g.row_map ? (int64_t)g.row_map[row] : (int64_t)row;
              float* dst = g.C + orow * g.ldc + col;
              const float* brow = g.bias_rows ? g.bias_rows + (int64_t)(row % g.bias_period) * g.N : nullptr;
              if (col + 3 < g.N) {
                if (g.bias_n) { const float4 b = *(const float4*)(g.bias_n + col); v.x += b.x; v.y += b.y; v.z += b.z; v.w += b.w; }
                if (brow) { const float4 b = *(const float4*)(brow + col); v.x += b.x; v.y += b.y; v.z += b.z; v.w += b.w; }
                if (g.beta != 0.f) { const float4 o = *(const float4*)dst; v.x += g.beta * o.x; v.y += g.beta * o.y; v.z += g.beta * o.z; v.w += g.beta * o.w; }
                *(float4*)dst = v;
              } else {
                for (int q = 0; q < 4 && col + q < g.N; ++q) {
                  float x = vp[q];
                  if (g.bias_n) x += g.bias_n[col + q];
                  if (brow) x += brow[col + q];
                  if (g.beta != 0.f) x += g.beta * dst[q];
                  dst[q] = x;
                }
              }
            }
          }
        } else {
          const int col = cb + lane;
          if (col < g.N) {
            const float bn = g.bias_n ? g.bias_n[col] : 0.f;
            for (int rr = 0; rr < 32; ++rr) {
              const int row = row0 + rr;
              if (row >= g.M) break;
              float x = cst[rr * CSTRIDE + lane] + bn;
              if (g.bias_rows) x += g.bias_rows[(int64_t)(row % g.bias_period) * g.N + col];
              const int64_t orow = g.row_map ? (int64_t)g.row_map[row] : (int64_t)row;
              float* dst = g.C + orow * g.ldc + col;
              if (g.beta != 0.f) x += g.beta * (*dst);
              *dst = x;
            }
          }
        }
        __syncwarp();
      }
      ptx::tc_fence_before();
      __syncwarp();
      if (lane == 0) ptx::mbar_arrive(tempty_bar(acc));
      if (++acc == 2) { acc = 0; acc_phase ^= 1u; }
    }
  }
  ptx::tc_fence_before();
  __syncthreads();
  if (warp == 1) {
    ptx::tc_fence_after();
    ptx::tmem_dealloc<TMEM_COLS>(tmem_base);
  }
}

template <bool A_MN, bool B_MN, int BN>
int launch(const CUtensorMap* tm, const GemmArgs& g, int grid, cudaStream_t st) {
  auto kern = k_gemm_tc<A_MN, B_MN, BN>;
  static bool configured = false;
  if (!configured) {
    LV_CUDA(cudaFuncSetAttribute(kern, cudaFuncAttributeMaxDynamicSharedMemorySize, smem_bytes<BN>()));
    configured = true;
  }
  kern<<<grid, NTHREADS, smem_bytes<BN>(), st>>>(tm[0], tm[1], tm[2], tm[3], g);
  LV_LAUNCH_CHECK();
  return LAGVAE_OK;
}

int num_sms() {
  static int n = 0;
  if (!n) {
    int dev = 0;
    cudaGetDevice(&dev);
    cudaDeviceGetAttribute(&n, cudaDevAttrMultiProcessorCount, dev);
    if (n <= 0) n = 148;
  }
  return n;
}

}  // namespace

// Optional cap on the persistent grid of the NEXT gemm_tc launches of this thread (0 = one CTA per SM): lets a GEMM run on
// the SMs a co-resident persistent kernel leaves idle (text_plan.cu: weight-gradient GEMM under the LSTM recurrence).
static thread_local int tl_grid_cap = 0;
void gemm_tc_set_grid_cap(int max_ctas) { tl_grid_cap = max_ctas > 0 ? max_ctas : 0; }

int gemm_tc(const TcOperand& A, const TcOperand& B, float* C, int64_t ldc, int M, int N, int K, int passes,
            float alpha, float beta, const float* bias_n, const float* bias_rows, int bias_period,
            const int32_t* out_row_map, cudaStream_t st) {
  if (M <= 0 || N <= 0) return LAGVAE_OK;
  LV_CHECK_ARG(K > 0, "gemm_tc: K must be > 0");
  LV_CHECK_ARG(passes == 1 || passes == 3, "gemm_tc: passes must be 1 or 3");
  LV_CHECK_ARG(A.hi && B.hi && (passes == 1 || (A.lo && B.lo)), "gemm_tc: missing operand part");
  LV_CHECK_ARG(bias_rows == nullptr || bias_period > 0, "gemm_tc: bias_period");
  // tile shape: the wide 128x256 tile when it still gives every SM at least two tiles
  static const int force_bn = [] { const char* e = getenv("LAGVAE_GEMM_BN"); return e ? atoi(e) : 0; }();
  const int64_t tiles256 = cdiv(M, BM) * cdiv(N, 256);
  const int BN = force_bn ? force_bn : ((N >= 256 && tiles256 >= 2 * num_sms()) ? 256 : 128);
  CUtensorMap tm[4];
  // K-major operand: tensor [rows=MN, cols=K], box 64(K) x rows(128 | BN).
  // MN-major operand: tensor [rows=K, cols=MN], box 64(MN) x 64(K rows).
  const void* parts[4] = {A.hi, passes == 3 ? A.lo : A.hi, B.hi, passes == 3 ? B.lo : B.hi};
  for (int i = 0; i < 4; ++i) {
    const bool isA = i < 2;
    const TcOperand& o = isA ? A : B;
    const uint64_t mn = isA ? M : N;
    if (!o.mn_major)
      LV_TRY(make_tmap_bf16_2d(&tm[i], parts[i], mn, K, o.ld, BK, isA ? BM : BN));
    else
      LV_TRY(make_tmap_bf16_2d(&tm[i], parts[i], K, mn, o.ld, 64, BK));
  }
  GemmArgs g{C, ldc, M, N, K, passes, alpha, beta, bias_n, bias_rows, bias_period, out_row_map};
  const int tiles = (int)(cdiv(M, BM) * cdiv(N, BN));
  const int grid = std::min(tiles, tl_grid_cap ? std::min(tl_grid_cap, num_sms()) : num_sms());
  const int key = (A.mn_major ? 2 : 0) | (B.mn_major ? 1 : 0);
  if (BN == 256) {
    switch (key) {
      case 0: return launch<false, false, 256>(tm, g, grid, st);
      case 1: return launch<false, true, 256>(tm, g, grid, st);
      case 2: return launch<true, false, 256>(tm, g, grid, st);
      default: return launch<true, true, 256>(tm, g, grid, st);
    }
  }
  switch (key) {
    case 0: return launch<false, false, 128>(tm, g, grid, st);
    case 1: return launch<false, true, 128>(tm, g, grid, st);
    case 2: return launch<true, false, 128>(tm, g, grid, st);
    default: return launch<true, true, 128>(tm, g, grid, st);
  }
}

}  // namespace lagvae
