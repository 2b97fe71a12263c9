// Exchange-protocol probe for the persistent LSTM recurrence (sm_100a).  No tensor-core work: it isolates the per-time-step
// "publish 1 KB per CTA -> make it visible -> every CTA pulls its 64 KB K-slice into shared memory" loop that bounds
// k_lstm_v2 (profiles/README.md: 5.8 K of 12.2 K cycles per time step), with the forward geometry at Bd = 32:
// 128 CTAs, K-split over pairs (rank = cta & 1 reads hidden units [512 rank, 512 rank + 512) = 8 k-blocks of 64).
// Every protocol is validated: producers write values derived from (step, row, unit) and consumers check what landed.
//
//   P0  generic stores -> fence.proxy.async -> bar -> red.release(counter) | poll ld.acquire(counter) -> fence.proxy.async -> TMA x16
//   P1  P0 without the producer-side fence.proxy.async
//   P2  per-CTA flags (st.release) | warp polls the 64 flags of its K-slice, TMA per k-block as soon as its 8 producers are in
//   P3  P2 but all 64 flags before the first TMA
//   P4  P2 with the slice staged in shared memory and written by ONE bulk TMA store per part (async proxy end to end)
//   P5  P2 without the producer-side fence.proxy.async
//   --- "tile image" global layout: the operand lives in global memory as the shared-memory image [k-block][part][32 rows][128 B,
//       SWIZZLE_128B applied by the producers], so a k-block stage is ONE contiguous 8 KB cp.async.bulk (no tensor map, no 32
//       separate 128-B rows per box) ---
//   P6  counter, consumer-side proxy fence only, 8 x 8 KB bulk copies
//   P7  P6 with one counter per K-slice (64 arrivals each)
//   P8  per-CTA flags polled with ld.acquire.v2 (no separate fence.acq_rel), bulk copy per ready k-block
//   P9  P7 with ONE 64 KB bulk copy
//   P10 P7, every storing warp arrives on its own (no CTA-wide bar.sync before the release)
//   P11 P6 with the k-block order staggered per CTA (start at (cta / 2) % 8): spreads the L2 slices the 64 readers of a slice hit
//   P12 P6 with 2 x 32 KB copies
//   P13 P6 WITHOUT any proxy fence (validation tells whether async-proxy reads see acquired generic writes)
//   P14 P6 with fence.proxy.async.global on the consumer side
//   P15 P6 but the 4 epilogue warps copy the slice with generic LDG.128 -> STS.128 (image layout == smem layout) and a
//       CTA-local fence.proxy.async.shared::cta; no global proxy fence at all
//   P16 P11 + P12: staggered start, 4 x 16 KB copies
//   P20/P21/P22 = P9 / P15 / P6 with the operand written to 4 replicas (different addresses -> different L2 slices); a consumer
//       reads replica (cta / 2) % 4: every line is read by 16 SMs instead of 64 (tests the L2 hot-spot hypothesis)
//   P23a/b/c clusters of 1 / 2 / 4 CTAs that consume the SAME K slice: every member loads 1/CS of it with a multicast bulk copy
//   P18 P15 while another warp spins on ld.acquire.gpu of the counter (next step's target) during the copy
//   P19 P18 with the spinning warp using ld.relaxed.gpu + nanosleep(64)
//   P17 validity-in-data: no counter, no release, no global fence.  Producers write whole 16-B chunks whose first bf16 carries
//       a 1-bit tag in its LSB (tag = parity of the slot's reuse count, 4 slots); the 4 epilogue warps poll-load their K-slice
//       with LDG.128 (L2), re-issue only the chunks whose tag is stale, store to shared memory (image layout == smem layout),
//       fence.proxy.async.shared::cta, arrive.  Chain = store one-way + load round trip.
//
// build: nvcc -gencode arch=compute_100a,code=sm_100a -O3 -std=c++17 -o exchange_probe exchange_probe.cu
// run  : ./exchange_probe [steps=400] [delay_cycles=0]
#include <cstdio>
#include <cstdlib>
#include <cstring>
#include <vector>

#include <cuda_bf16.h>

#include "../../vae-lagging-encoder_b200/csrc/sm100_ptx.cuh"

using namespace lagvae;

#define CK(x)                                                                                     \
  do {                                                                                            \
    cudaError_t e_ = (x);                                                                         \
    if (e_ != cudaSuccess) {                                                                      \
      printf("CUDA error %s at %s:%d\n", cudaGetErrorString(e_), __FILE__, __LINE__);             \
      exit(1);                                                                                    \
    }                                                                                             \
  } while (0)

constexpr int G = 128, BD = 32, NH = 1024, KB = 8, NS = 8, NSLOT = 4, NREP = 4;
constexpr int PART_BYTES = BD * 128;                 // one operand part of a ring stage: 32 rows x 128 B
constexpr int NTHREADS = 192;

struct Maps {
  CUtensorMap ld[NSLOT * 2];   // [slot][part] load view: [BD rows, NH cols] bf16, box 64 x 32, SWIZZLE_128B
  CUtensorMap st[NSLOT * 2];   // [slot][part] store view: box 8 x 32, no swizzle
};

struct Args {
  __nv_bfloat16* abuf;         // [NSLOT][2][BD][NH]
  unsigned* counter;
  unsigned* flags;             // [G]
  unsigned long long* trace;   // [steps][8] for CTA 0
  unsigned* errors;
  int steps, delay;
};

__device__ __forceinline__ unsigned ld_acquire_u32(const unsigned* p) {
  unsigned v;
  asm volatile("ld.acquire.gpu.global.u32 %0, [%1];" : "=r"(v) : "l"(p) : "memory");
  return v;
}
__device__ __forceinline__ uint2 ld_relaxed_v2(const unsigned* p) {
  uint2 v;
  asm volatile("ld.relaxed.gpu.global.v2.u32 {%0, %1}, [%2];" : "=r"(v.x), "=r"(v.y) : "l"(p) : "memory");
  return v;
}
__device__ __forceinline__ uint2 ld_acquire_v2(const unsigned* p) {
  uint2 v;
  asm volatile("ld.acquire.gpu.global.v2.u32 {%0, %1}, [%2];" : "=r"(v.x), "=r"(v.y) : "l"(p) : "memory");
  return v;
}
__device__ __forceinline__ void bulk_g2s(uint32_t dst, const void* src, uint32_t bytes, uint32_t bar) {
  asm volatile("cp.async.bulk.shared::cluster.global.mbarrier::complete_tx::bytes [%0], [%1], %2, [%3];" ::"r"(dst), "l"(src),
               "r"(bytes), "r"(bar)
               : "memory");
}
__device__ __forceinline__ unsigned short bf16_bits(int step, int part, int b, int u) {
  // exactly representable pattern, distinct per (step, part, row, unit)
  return (unsigned short)((step * 131 + part * 17 + b * 7 + u * 3) & 0x7fff);
}

template <int PROTO>
__global__ void __launch_bounds__(NTHREADS, 1) k_probe(const Args a, const __grid_constant__ Maps tm) {
  extern __shared__ uint8_t smem_raw[];
  const uint32_t base = (ptx::smem_u32(smem_raw) + 1023u) & ~1023u;
  uint8_t* gen = smem_raw + (base - ptx::smem_u32(smem_raw));
  const uint32_t ring = base;                                  // NS stages x 2 parts x 4 KB
  const uint32_t stage_sm = ring + NS * 2 * PART_BYTES;         // staging for the TMA store: 2 parts x 32 rows x 16 B
  const uint32_t bars = stage_sm + 1024;
  auto full = [&](int s) { return bars + 8u * s; };
  auto empty = [&](int s) { return bars + 8u * (NS + s); };
  const uint32_t done_bar = bars + 8u * (2 * NS);
  const int warp = threadIdx.x >> 5, lane = threadIdx.x & 31;
  const int rank = blockIdx.x & 1, u0 = blockIdx.x * 8;
  if (threadIdx.x == 0) {
    for (int s = 0; s < NS; ++s) { ptx::mbar_init(full(s), 1); ptx::mbar_init(empty(s), 1); }
    ptx::mbar_init(done_bar, 1);
    ptx::fence_mbar_init();
  }
  __syncthreads();
  const size_t slot_elems = (size_t)2 * BD * NH;
  int stage = 0;
  uint32_t phase = 0;
  const bool tr = a.trace != nullptr && blockIdx.x == 0;
  unsigned errs = 0;

  for (int s = 0; s <= a.steps; ++s) {
    // ---------------- consume: operand of step s was published at the end of step s-1 (s = 0: nothing to read)
    if (s > 0) {
      const int rd = (s - 1) & (NSLOT - 1);
      if (warp == 5) {
        if (PROTO <= 1) {
          if (lane == 0) {
            const unsigned target = (unsigned)s * G;
            unsigned spins = 0;
            while (ld_acquire_u32(a.counter) < target) if (++spins > (1u << 24)) asm volatile("trap;");
          }
          __syncwarp();
          if (tr && lane == 0) a.trace[s * 8 + 1] = clock64();
          ptx::fence_proxy_async_all();
          for (int kb = 0; kb < KB; ++kb) {
            ptx::mbar_wait(empty(stage), phase ^ 1u);
            if (ptx::elect_one()) {
              const int kx = (rank * KB + kb) * 64;
              ptx::mbar_expect_tx(full(stage), 2u * PART_BYTES);
              ptx::tma_load_2d(ring + stage * 2 * PART_BYTES, &tm.ld[rd * 2 + 0], full(stage), kx, 0);
              ptx::tma_load_2d(ring + stage * 2 * PART_BYTES + PART_BYTES, &tm.ld[rd * 2 + 1], full(stage), kx, 0);
            }
            __syncwarp();
            if (++stage == NS) { stage = 0; phase ^= 1u; }
          }
        } else {
          // lane l watches producers 64 rank + 2l, 2l+1 ; k-block kb <-> lanes [4kb, 4kb+4)
          const unsigned* fp = a.flags + 64 * rank + 2 * lane;
          unsigned issued = 0;     // bit kb
          unsigned spins = 0;
          bool first = true;
          while (issued != 0xffu) {
            const uint2 f = ld_relaxed_v2(fp);
            const bool ok = f.x >= (unsigned)s && f.y >= (unsigned)s;
            const unsigned m = __ballot_sync(0xffffffffu, ok);
            unsigned ready = 0;
#pragma unroll
            for (int kb = 0; kb < KB; ++kb) if (((m >> (4 * kb)) & 0xfu) == 0xfu) ready |= 1u << kb;
            if (PROTO == 3 && ready != 0xffu) ready = 0;
            unsigned todo = ready & ~issued;
            if (todo) {
              asm volatile("fence.acq_rel.gpu;" ::: "memory");
              ptx::fence_proxy_async_all();
              if (tr && lane == 0 && first) { a.trace[s * 8 + 1] = clock64(); first = false; }
              // in k-block order (deterministic accumulation order in the real kernel): stop at the first gap
              for (int kb = 0; kb < KB; ++kb) {
                if (issued & (1u << kb)) continue;
                if (!(todo & (1u << kb))) break;
                ptx::mbar_wait(empty(stage), phase ^ 1u);
                if (ptx::elect_one()) {
                  const int kx = (rank * KB + kb) * 64;
                  ptx::mbar_expect_tx(full(stage), 2u * PART_BYTES);
                  ptx::tma_load_2d(ring + stage * 2 * PART_BYTES, &tm.ld[rd * 2 + 0], full(stage), kx, 0);
                  ptx::tma_load_2d(ring + stage * 2 * PART_BYTES + PART_BYTES, &tm.ld[rd * 2 + 1], full(stage), kx, 0);
                }
                __syncwarp();
                issued |= 1u << kb;
                if (++stage == NS) { stage = 0; phase ^= 1u; }
              }
            }
            if (++spins > (1u << 22)) asm volatile("trap;");
          }
        }
      } else if (warp == 4) {
        // "MMA" warp: consume the stages in order, validate the first 16 bytes of 32 rows of each part
        int st2 = stage;
        uint32_t ph2 = phase;
        for (int kb = 0; kb < KB; ++kb) {
          ptx::mbar_wait(full(st2), ph2);
          if (tr && lane == 0 && kb == 0) a.trace[s * 8 + 2] = clock64();
          // row = lane, chunk 0 of the swizzled tile holds units [64 kb' , +8) of that row: chunk index c ^ (row & 7)
          for (int part = 0; part < 2; ++part) {
            const uint8_t* tile = gen + (st2 * 2 * PART_BYTES + part * PART_BYTES);
            const int c = 3;                                           // logical 16-B chunk 3 = units 24..31 of the k-block
            const uint4 v = *(const uint4*)(tile + lane * 128 + ((c ^ (lane & 7)) << 4));
            const int ug = (rank * KB + kb) * 64 + c * 8;
            const unsigned short w0 = (unsigned short)(v.x & 0xffff);
            if (w0 != bf16_bits(s - 1, part, lane, ug)) ++errs;
          }
          __syncwarp();
          if (lane == 0) ptx::mbar_arrive(empty(st2));
          if (++st2 == NS) { st2 = 0; ph2 ^= 1u; }
        }
        if (tr && lane == 0) a.trace[s * 8 + 3] = clock64();
        if (lane == 0) ptx::mbar_arrive(done_bar);
        stage = st2; phase = ph2;
      }
      if (warp < 4) ptx::mbar_wait(done_bar, (uint32_t)((s - 1) & 1));
      if (warp == 5) { /* stage/phase already advanced */ }
      if (warp < 4) {  // keep the ring bookkeeping of the epilogue warps in step (unused by them)
      }
    }
    if (s == a.steps) break;
    // ---------------- produce the operand of step s+1 into slot s & 3
    if (warp < 4) {
      if (a.delay > 0) {
        const long long t0 = clock64();
        while (clock64() - t0 < a.delay) {
        }
      }
      if (tr && threadIdx.x == 0) a.trace[s * 8 + 4] = clock64();
      const int wr = s & (NSLOT - 1);
      __nv_bfloat16* wbase = a.abuf + (size_t)wr * slot_elems;
      if (threadIdx.x < 64) {
        const int b = threadIdx.x >> 1, uq = threadIdx.x & 1, ub = u0 + uq * 4;
        uint2 hi, lo;
        hi.x = bf16_bits(s, 0, b, ub) | ((uint32_t)bf16_bits(s, 0, b, ub + 1) << 16);
        hi.y = bf16_bits(s, 0, b, ub + 2) | ((uint32_t)bf16_bits(s, 0, b, ub + 3) << 16);
        lo.x = bf16_bits(s, 1, b, ub) | ((uint32_t)bf16_bits(s, 1, b, ub + 1) << 16);
        lo.y = bf16_bits(s, 1, b, ub + 2) | ((uint32_t)bf16_bits(s, 1, b, ub + 3) << 16);
        if (PROTO == 4) {
          uint8_t* sg = gen + (stage_sm - base);
          *(uint2*)(sg + b * 16 + uq * 8) = hi;
          *(uint2*)(sg + 512 + b * 16 + uq * 8) = lo;
        } else {
          *(uint2*)(wbase + (size_t)b * NH + ub) = hi;
          *(uint2*)(wbase + (size_t)BD * NH + (size_t)b * NH + ub) = lo;
        }
      }
      if (PROTO == 4) {
        ptx::fence_proxy_async_smem();
        asm volatile("bar.sync 1, 128;" ::: "memory");
        if (threadIdx.x == 0) {
          asm volatile("cp.async.bulk.tensor.2d.global.shared::cta.bulk_group [%0, {%2, %3}], [%1];" ::"l"((uint64_t)&tm.st[wr * 2 + 0]),
                       "r"(stage_sm), "r"(u0), "r"(0)
                       : "memory");
          asm volatile("cp.async.bulk.tensor.2d.global.shared::cta.bulk_group [%0, {%2, %3}], [%1];" ::"l"((uint64_t)&tm.st[wr * 2 + 1]),
                       "r"(stage_sm + 512), "r"(u0), "r"(0)
                       : "memory");
          asm volatile("cp.async.bulk.commit_group;" ::: "memory");
          asm volatile("cp.async.bulk.wait_group 0;" ::: "memory");
          if (tr) a.trace[s * 8 + 5] = clock64();
          ptx::fence_proxy_async_all();
          asm volatile("st.release.gpu.global.u32 [%0], %1;" ::"l"(a.flags + blockIdx.x), "r"((unsigned)(s + 1)) : "memory");
          if (tr) a.trace[s * 8 + 6] = clock64();
        }
      } else {
        if (PROTO == 0 || PROTO == 2 || PROTO == 3) ptx::fence_proxy_async_all();
        if (tr && threadIdx.x == 0) a.trace[s * 8 + 5] = clock64();
        asm volatile("bar.sync 1, 128;" ::: "memory");
        if (threadIdx.x == 0) {
          if (PROTO <= 1)
            asm volatile("red.release.gpu.global.add.u32 [%0], 1;" ::"l"(a.counter) : "memory");
          else
            asm volatile("st.release.gpu.global.u32 [%0], %1;" ::"l"(a.flags + blockIdx.x), "r"((unsigned)(s + 1)) : "memory");
          if (tr) a.trace[s * 8 + 6] = clock64();
        }
      }
    }
  }
  if (errs) atomicAdd(a.errors, errs);
}


// ---------------------------------------------------------------------------------------------------------------------
// tile-image layout protocols (P6..P10)
// global operand: [NSLOT][16 k-blocks][2 parts][32 rows][128 B]; producer CTA j owns the 16-B chunk j % 8 of every row of k-block j / 8
template <int PROTO>
__global__ void __launch_bounds__(NTHREADS, 1) k_probe2(const Args a) {
  extern __shared__ uint8_t smem_raw[];
  const uint32_t base = (ptx::smem_u32(smem_raw) + 1023u) & ~1023u;
  uint8_t* gen = smem_raw + (base - ptx::smem_u32(smem_raw));
  const uint32_t ring = base;
  const uint32_t bars = ring + NS * 2 * PART_BYTES;
  auto full = [&](int s) { return bars + 8u * s; };
  auto empty = [&](int s) { return bars + 8u * (NS + s); };
  const uint32_t done_bar = bars + 8u * (2 * NS);
  const int warp = threadIdx.x >> 5, lane = threadIdx.x & 31;
  const int rank = blockIdx.x & 1, u0 = blockIdx.x * 8;
  if (threadIdx.x == 0) {
    for (int s = 0; s < NS; ++s) { ptx::mbar_init(full(s), 1); ptx::mbar_init(empty(s), 1); }
    ptx::mbar_init(done_bar, 1);
    ptx::fence_mbar_init();
  }
  __syncthreads();
  constexpr size_t SLOT_BYTES = (size_t)16 * 2 * PART_BYTES;      // 128 KB
  constexpr uint32_t STAGE_BYTES = 2 * PART_BYTES;                // 8 KB
  uint8_t* gbase = (uint8_t*)a.abuf;
  int stage = 0;
  uint32_t phase = 0;
  const bool tr = a.trace != nullptr && blockIdx.x == 0;
  unsigned errs = 0;
  const unsigned arrivals_per_cta = (PROTO == 10) ? 2u : 1u;
  unsigned* my_counter = (PROTO == 6) ? a.counter : a.counter + 32 * (blockIdx.x / 64);        // slice the CTA PRODUCES into
  const unsigned* rd_counter = (PROTO == 6) ? a.counter : a.counter + 32 * rank;               // slice it CONSUMES
  const unsigned per_step = (PROTO == 6 ? (unsigned)G : 64u) * arrivals_per_cta;

  for (int s = 0; s <= a.steps; ++s) {
    if (s > 0) {
      constexpr bool REPL = (PROTO == 20 || PROTO == 21 || PROTO == 22);
      const uint8_t* src = gbase + (size_t)((s - 1) & (NSLOT - 1)) * SLOT_BYTES + (size_t)rank * KB * STAGE_BYTES +
                           (REPL ? (size_t)((blockIdx.x >> 1) & (NREP - 1)) * NSLOT * SLOT_BYTES : 0);
      if (warp == 5) {
        if (PROTO == 15 || PROTO == 21) {
          // warp 5 idle: the epilogue warps do the copy below
        } else if (PROTO == 18 || PROTO == 19) {
          // warp 5 keeps polling the counter for the NEXT step while the epilogue warps copy (what k_lstm_v2's poller does)
          if (lane == 0 && s < a.steps) {
            const unsigned target = (unsigned)(s + 1) * per_step;
            unsigned spins = 0;
            if (PROTO == 18) {
              while (ld_acquire_u32(rd_counter) < target) if (++spins > (1u << 24)) asm volatile("trap;");
            } else {
              unsigned v;
              do {
                asm volatile("ld.relaxed.gpu.global.u32 %0, [%1];" : "=r"(v) : "l"(rd_counter) : "memory");
                if (v < target) __nanosleep(64);
                if (++spins > (1u << 24)) asm volatile("trap;");
              } while (v < target);
            }
          }
          __syncwarp();
        } else if (PROTO != 8) {
          if (lane == 0) {
            const unsigned target = (unsigned)s * per_step;
            unsigned spins = 0;
            while (ld_acquire_u32(rd_counter) < target) if (++spins > (1u << 24)) asm volatile("trap;");
          }
          __syncwarp();
          if (tr && lane == 0) a.trace[s * 8 + 1] = clock64();
          if (PROTO == 14) asm volatile("fence.proxy.async.global;" ::: "memory");
          else if (PROTO != 13) ptx::fence_proxy_async_all();
          if (tr && lane == 0) a.trace[s * 8 + 7] = clock64();
          if (PROTO == 9 || PROTO == 20) {
            // one 64 KB copy: all NS = 8 stages are consecutive in the ring; stage 0's barrier carries the whole transaction
            for (int kb = 0; kb < KB; ++kb) ptx::mbar_wait(empty(kb), phase ^ 1u);
            if (ptx::elect_one()) {
              ptx::mbar_expect_tx(full(0), KB * STAGE_BYTES);
              bulk_g2s(ring, src, KB * STAGE_BYTES, full(0));
              for (int kb = 1; kb < KB; ++kb) ptx::mbar_arrive(full(kb));     // trivially complete: data guarded by full(0)
            }
            __syncwarp();
            phase ^= 1u;
          } else if (PROTO == 12 || PROTO == 16) {
            // G copies of (8 / G) stages each; the first stage of a group carries the group's transaction
            constexpr int GRP = (PROTO == 12) ? 4 : 2;         // stages per copy
            const int start = (PROTO == 16) ? ((blockIdx.x >> 1) & 7) & ~(GRP - 1) : 0;
            for (int kb = 0; kb < KB; ++kb) ptx::mbar_wait(empty(kb), phase ^ 1u);
            if (ptx::elect_one()) {
              for (int g = 0; g < KB / GRP; ++g) {
                const int k0 = (start + g * GRP) & 7;
                ptx::mbar_expect_tx(full(k0), GRP * STAGE_BYTES);
                bulk_g2s(ring + k0 * STAGE_BYTES, src + (size_t)k0 * STAGE_BYTES, GRP * STAGE_BYTES, full(k0));
                for (int j = 1; j < GRP; ++j) ptx::mbar_arrive(full(k0 + j));
              }
            }
            __syncwarp();
            phase ^= 1u;
          } else {
            const int start = (PROTO == 11) ? ((blockIdx.x >> 1) & 7) : 0;
            for (int i = 0; i < KB; ++i) {
              const int kb = (start + i) & 7;
              // stage index == k-block index (NS == KB): the consumer walks them in the same rotated order
              ptx::mbar_wait(empty(kb), phase ^ 1u);
              if (ptx::elect_one()) {
                ptx::mbar_expect_tx(full(kb), STAGE_BYTES);
                bulk_g2s(ring + kb * STAGE_BYTES, src + (size_t)kb * STAGE_BYTES, STAGE_BYTES, full(kb));
              }
              __syncwarp();
            }
            phase ^= 1u;
          }
        } else {
          const unsigned* fp = a.flags + 64 * rank + 2 * lane;
          unsigned issued = 0, spins = 0;
          bool first = true;
          while (issued != 0xffu) {
            const uint2 f = ld_acquire_v2(fp);
            const bool ok = f.x >= (unsigned)s && f.y >= (unsigned)s;
            const unsigned m = __ballot_sync(0xffffffffu, ok);
            unsigned ready = 0;
#pragma unroll
            for (int kb = 0; kb < KB; ++kb) if (((m >> (4 * kb)) & 0xfu) == 0xfu) ready |= 1u << kb;
            const unsigned todo = ready & ~issued;
            if (todo) {
              ptx::fence_proxy_async_all();
              if (tr && lane == 0 && first) { a.trace[s * 8 + 1] = clock64(); first = false; }
              for (int kb = 0; kb < KB; ++kb) {
                if (issued & (1u << kb)) continue;
                if (!(todo & (1u << kb))) break;
                ptx::mbar_wait(empty(stage), phase ^ 1u);
                if (ptx::elect_one()) {
                  ptx::mbar_expect_tx(full(stage), STAGE_BYTES);
                  bulk_g2s(ring + stage * STAGE_BYTES, src + (size_t)kb * STAGE_BYTES, STAGE_BYTES, full(stage));
                }
                __syncwarp();
                issued |= 1u << kb;
                if (++stage == NS) { stage = 0; phase ^= 1u; }
              }
            }
            if (++spins > (1u << 22)) asm volatile("trap;");
          }
        }
      } else if (warp == 4) {
        if (PROTO == 8) {
          int st2 = stage;
          uint32_t ph2 = phase;
          for (int kb = 0; kb < KB; ++kb) {
            ptx::mbar_wait(full(st2), ph2);
            if (tr && lane == 0 && kb == 0) a.trace[s * 8 + 2] = clock64();
            for (int part = 0; part < 2; ++part) {
              const uint8_t* tile = gen + (st2 * STAGE_BYTES + part * PART_BYTES);
              const int c = 3;
              const uint4 v = *(const uint4*)(tile + lane * 128 + ((c ^ (lane & 7)) << 4));
              const int ug = (rank * KB + kb) * 64 + c * 8;
              if ((unsigned short)(v.x & 0xffff) != bf16_bits(s - 1, part, lane, ug)) ++errs;
            }
            __syncwarp();
            if (lane == 0) ptx::mbar_arrive(empty(st2));
            if (++st2 == NS) { st2 = 0; ph2 ^= 1u; }
          }
          stage = st2; phase = ph2;
        } else {
          // stage index == k-block index; walk in the producer's order, group leaders carry the data barrier
          const int GRP = (PROTO == 9 || PROTO == 20) ? 8 : (PROTO == 12 ? 4 : (PROTO == 16 ? 2 : 1));
          const int start = (PROTO == 11) ? ((blockIdx.x >> 1) & 7) : (PROTO == 16 ? (((blockIdx.x >> 1) & 7) & ~(GRP - 1)) : 0);
          for (int i = 0; i < KB; ++i) {
            const int kb = (start + i) & 7;
            ptx::mbar_wait(full(kb & ~(GRP - 1)), phase);
            ptx::mbar_wait(full(kb), phase);
            if (tr && lane == 0 && i == 0) a.trace[s * 8 + 2] = clock64();
            for (int part = 0; part < 2; ++part) {
              const uint8_t* tile = gen + (kb * STAGE_BYTES + part * PART_BYTES);
              const int c = 3;
              const uint4 v = *(const uint4*)(tile + lane * 128 + ((c ^ (lane & 7)) << 4));
              const int ug = (rank * KB + kb) * 64 + c * 8;
              if ((unsigned short)(v.x & 0xffff) != bf16_bits(s - 1, part, lane, ug)) ++errs;
            }
            __syncwarp();
            if (lane == 0) ptx::mbar_arrive(empty(kb));
          }
          phase ^= 1u;
        }
        if (tr && lane == 0) a.trace[s * 8 + 3] = clock64();
        if (lane == 0) ptx::mbar_arrive(done_bar);
      } else if (PROTO == 15 || PROTO == 18 || PROTO == 19 || PROTO == 21) {
        // epilogue warps: generic-proxy copy of the 64 KB slice (image layout == shared-memory layout)
        if (threadIdx.x == 0) {
          const unsigned target = (unsigned)s * per_step;
          unsigned spins = 0;
          while (ld_acquire_u32(rd_counter) < target) if (++spins > (1u << 24)) asm volatile("trap;");
          if (tr) a.trace[s * 8 + 1] = clock64();
        }
        asm volatile("bar.sync 1, 128;" ::: "memory");
        const uint4* g4 = (const uint4*)src;
        uint4* s4 = (uint4*)gen;
#pragma unroll
        for (int h = 0; h < 2; ++h) {
          uint4 v[16];
#pragma unroll
          for (int j = 0; j < 16; ++j) v[j] = __ldcg(g4 + (h * 16 + j) * 128 + threadIdx.x);
#pragma unroll
          for (int j = 0; j < 16; ++j) s4[(h * 16 + j) * 128 + threadIdx.x] = v[j];
        }
        ptx::fence_proxy_async_smem();
        asm volatile("bar.sync 1, 128;" ::: "memory");
        if (threadIdx.x == 0) {
          if (tr) a.trace[s * 8 + 2] = clock64();
          for (int kb = 0; kb < KB; ++kb) ptx::mbar_arrive(full(kb));
        }
      }
      if (warp < 4) ptx::mbar_wait(done_bar, (uint32_t)((s - 1) & 1));
    }
    if (s == a.steps) break;
    if (warp < 4) {
      if (a.delay > 0) {
        const long long t0 = clock64();
        while (clock64() - t0 < a.delay) {
        }
      }
      if (tr && threadIdx.x == 0) a.trace[s * 8 + 4] = clock64();
      uint8_t* wslot = gbase + (size_t)(s & (NSLOT - 1)) * SLOT_BYTES + (size_t)(blockIdx.x / 8) * STAGE_BYTES;
      const int c = blockIdx.x & 7;
      if (threadIdx.x < 64) {
        const int b = threadIdx.x >> 1, uq = threadIdx.x & 1, ub = u0 + uq * 4;
        uint2 hi, lo;
        hi.x = bf16_bits(s, 0, b, ub) | ((uint32_t)bf16_bits(s, 0, b, ub + 1) << 16);
        hi.y = bf16_bits(s, 0, b, ub + 2) | ((uint32_t)bf16_bits(s, 0, b, ub + 3) << 16);
        lo.x = bf16_bits(s, 1, b, ub) | ((uint32_t)bf16_bits(s, 1, b, ub + 1) << 16);
        lo.y = bf16_bits(s, 1, b, ub + 2) | ((uint32_t)bf16_bits(s, 1, b, ub + 3) << 16);
        const uint32_t off = (uint32_t)(b * 128 + ((c ^ (b & 7)) << 4) + uq * 8);
        constexpr bool REPLW = (PROTO == 20 || PROTO == 21 || PROTO == 22);
#pragma unroll
        for (int r = 0; r < (REPLW ? NREP : 1); ++r) {
          *(uint2*)(wslot + (size_t)r * NSLOT * SLOT_BYTES + off) = hi;
          *(uint2*)(wslot + (size_t)r * NSLOT * SLOT_BYTES + PART_BYTES + off) = lo;
        }
      }
      if (tr && threadIdx.x == 0) a.trace[s * 8 + 5] = clock64();
      if (PROTO == 10) {
        if (warp < 2) {
          __syncwarp();
          if (lane == 0) asm volatile("red.release.gpu.global.add.u32 [%0], 1;" ::"l"(my_counter) : "memory");
        }
        if (tr && threadIdx.x == 0) a.trace[s * 8 + 6] = clock64();
      } else {
        asm volatile("bar.sync 1, 128;" ::: "memory");
        if (threadIdx.x == 0) {
          if (PROTO == 8)
            asm volatile("st.release.gpu.global.u32 [%0], %1;" ::"l"(a.flags + blockIdx.x), "r"((unsigned)(s + 1)) : "memory");
          else
            asm volatile("red.release.gpu.global.add.u32 [%0], 1;" ::"l"(my_counter) : "memory");
          if (tr) a.trace[s * 8 + 6] = clock64();
        }
      }
    }
  }
  if (errs) atomicAdd(a.errors, errs);
}

// ---------------------------------------------------------------------------------------------------------------------
// P17: validity-in-data ("tagged chunks")
__device__ __forceinline__ uint4 ld_cg_u4(const void* p) {
  uint4 v;
  asm volatile("ld.relaxed.gpu.global.v4.u32 {%0, %1, %2, %3}, [%4];" : "=r"(v.x), "=r"(v.y), "=r"(v.z), "=r"(v.w) : "l"(p) : "memory");
  return v;
}
__global__ void __launch_bounds__(NTHREADS, 1) k_probe3(const Args a) {
  extern __shared__ uint8_t smem_raw[];
  const uint32_t base = (ptx::smem_u32(smem_raw) + 1023u) & ~1023u;
  uint8_t* gen = smem_raw + (base - ptx::smem_u32(smem_raw));
  const uint32_t bars = base + NS * 2 * PART_BYTES;
  auto full = [&](int s) { return bars + 8u * s; };
  const uint32_t done_bar = bars + 8u * (2 * NS);
  const int warp = threadIdx.x >> 5, lane = threadIdx.x & 31;
  const int rank = blockIdx.x & 1, u0 = blockIdx.x * 8;
  if (threadIdx.x == 0) {
    for (int s = 0; s < NS; ++s) ptx::mbar_init(full(s), 128);
    ptx::mbar_init(done_bar, 1);
    ptx::fence_mbar_init();
  }
  __syncthreads();
  constexpr size_t SLOT_BYTES = (size_t)16 * 2 * PART_BYTES;
  constexpr uint32_t STAGE_BYTES = 2 * PART_BYTES;
  uint8_t* gbase = (uint8_t*)a.abuf;
  const bool tr = a.trace != nullptr && blockIdx.x == 0;
  unsigned errs = 0;
  uint32_t phase = 0;
  for (int s = 0; s <= a.steps; ++s) {
    if (s > 0) {
      const int pidx = s - 1;                                             // publish index of the data read now
      const uint32_t want = (uint32_t)(((pidx / NSLOT) + 1) & 1);
      const uint8_t* src = gbase + (size_t)(pidx & (NSLOT - 1)) * SLOT_BYTES + (size_t)rank * KB * STAGE_BYTES;
      if (warp < 4) {
        const uint4* g4 = (const uint4*)src;
        uint4* s4 = (uint4*)gen;
        unsigned spins = 0;
#pragma unroll
        for (int h = 0; h < 2; ++h) {
          uint4 v[16];
#pragma unroll
          for (int j = 0; j < 16; ++j) v[j] = ld_cg_u4(g4 + (h * 16 + j) * 128 + threadIdx.x);
          bool all_ok;
          do {
            all_ok = true;
#pragma unroll
            for (int j = 0; j < 16; ++j)
              if ((v[j].x & 1u) != want) {
                v[j] = ld_cg_u4(g4 + (h * 16 + j) * 128 + threadIdx.x);
                all_ok = false;
              }
            if (++spins > (1u << 22)) asm volatile("trap;");
          } while (!all_ok);
          if (tr && threadIdx.x == 0 && h == 0) a.trace[s * 8 + 1] = clock64();
#pragma unroll
          for (int j = 0; j < 16; ++j) s4[(h * 16 + j) * 128 + threadIdx.x] = v[j];
          ptx::fence_proxy_async_smem();
#pragma unroll
          for (int j = 0; j < 4; ++j) ptx::mbar_arrive(full(h * 4 + j));   // 4 chunks of this thread per stage: stage = (h*16 + j) / 4
        }
      } else if (warp == 4) {
        for (int kb = 0; kb < KB; ++kb) {
          ptx::mbar_wait(full(kb), phase);
          if (tr && lane == 0 && kb == 0) a.trace[s * 8 + 2] = clock64();
          for (int part = 0; part < 2; ++part) {
            const uint8_t* tile = gen + (kb * STAGE_BYTES + part * PART_BYTES);
            const int c = 3;
            const uint4 v = *(const uint4*)(tile + lane * 128 + ((c ^ (lane & 7)) << 4));
            const int ug = (rank * KB + kb) * 64 + c * 8;
            if ((unsigned short)((v.x & 0xfffe)) != (bf16_bits(s - 1, part, lane, ug) & 0xfffe)) ++errs;
          }
          __syncwarp();
        }
        phase ^= 1u;
        if (tr && lane == 0) a.trace[s * 8 + 3] = clock64();
        if (lane == 0) ptx::mbar_arrive(done_bar);
      }
      if (warp < 4) ptx::mbar_wait(done_bar, (uint32_t)((s - 1) & 1));
    }
    if (s == a.steps) break;
    if (warp < 4) {
      if (a.delay > 0) {
        const long long t0 = clock64();
        while (clock64() - t0 < a.delay) {
        }
      }
      if (tr && threadIdx.x == 0) a.trace[s * 8 + 4] = clock64();
      const int pidx = s;                                                 // publish index
      const uint32_t tag = (uint32_t)(((pidx / NSLOT) + 1) & 1);
      uint8_t* wslot = gbase + (size_t)(pidx & (NSLOT - 1)) * SLOT_BYTES + (size_t)(blockIdx.x / 8) * STAGE_BYTES;
      const int c = blockIdx.x & 7;
      if (threadIdx.x < 64) {        // thread = (row b, part): one whole 16-B chunk each
        const int b = threadIdx.x >> 1, part = threadIdx.x & 1;
        uint4 v;
        v.x = ((bf16_bits(s, part, b, u0) & 0xfffe) | tag) | ((uint32_t)bf16_bits(s, part, b, u0 + 1) << 16);
        v.y = bf16_bits(s, part, b, u0 + 2) | ((uint32_t)bf16_bits(s, part, b, u0 + 3) << 16);
        v.z = bf16_bits(s, part, b, u0 + 4) | ((uint32_t)bf16_bits(s, part, b, u0 + 5) << 16);
        v.w = bf16_bits(s, part, b, u0 + 6) | ((uint32_t)bf16_bits(s, part, b, u0 + 7) << 16);
        *(uint4*)(wslot + part * PART_BYTES + b * 128 + ((c ^ (b & 7)) << 4)) = v;
      }
      if (tr && threadIdx.x == 0) { a.trace[s * 8 + 5] = clock64(); a.trace[s * 8 + 6] = clock64(); }
    }
  }
  if (errs) atomicAdd(a.errors, errs);
}

// ---------------------------------------------------------------------------------------------------------------------
// P23/P24: cluster multicast.  Clusters of CS CTAs whose members need the SAME K slice (rank = cluster index & 1): every member
// loads 1/CS of the 64 KB slice with cp.async.bulk ... .multicast::cluster into ALL members' rings, so the slice crosses the
// L2 -> SM fabric once per cluster instead of once per CTA.  Counter protocol as P6 (consumer-side .global proxy fence).
__device__ __forceinline__ uint32_t cluster_rank() {
  uint32_t r;
  asm volatile("mov.u32 %0, %%cluster_ctarank;" : "=r"(r));
  return r;
}
__device__ __forceinline__ void bulk_g2s_mc(uint32_t dst, const void* src, uint32_t bytes, uint32_t bar, uint16_t mask) {
  asm volatile("cp.async.bulk.shared::cluster.global.mbarrier::complete_tx::bytes.multicast::cluster [%0], [%1], %2, [%3], %4;" ::"r"(dst),
               "l"(src), "r"(bytes), "r"(bar), "h"(mask)
               : "memory");
}
template <int CS>
__global__ void __launch_bounds__(NTHREADS, 1) k_probe_mc(const Args a) {
  extern __shared__ uint8_t smem_raw[];
  const uint32_t base = (ptx::smem_u32(smem_raw) + 1023u) & ~1023u;
  uint8_t* gen = smem_raw + (base - ptx::smem_u32(smem_raw));
  const uint32_t ring = base;
  const uint32_t bars = ring + NS * 2 * PART_BYTES;
  const uint32_t full0 = bars;                      // ONE barrier for the whole slice: CS multicast copies complete_tx on it
  const uint32_t done_bar = bars + 8u * (2 * NS);
  const int warp = threadIdx.x >> 5, lane = threadIdx.x & 31;
  const uint32_t crank = cluster_rank();
  const int cl = blockIdx.x / CS;
  const int rank = cl & 1;                          // K slice this whole cluster consumes
  const int u0 = blockIdx.x * 8;
  if (threadIdx.x == 0) {
    ptx::mbar_init(full0, 1);
    ptx::mbar_init(done_bar, 1);
    ptx::fence_mbar_init();
  }
  __syncthreads();
  asm volatile("barrier.cluster.arrive.release.aligned;" ::: "memory");
  asm volatile("barrier.cluster.wait.acquire.aligned;" ::: "memory");
  constexpr size_t SLOT_BYTES = (size_t)16 * 2 * PART_BYTES;
  constexpr uint32_t STAGE_BYTES = 2 * PART_BYTES, SLICE = KB * STAGE_BYTES, PIECE = SLICE / CS;
  uint8_t* gbase = (uint8_t*)a.abuf;
  const bool tr = a.trace != nullptr && blockIdx.x == 0;
  unsigned errs = 0;
  uint32_t phase = 0;
  for (int s = 0; s <= a.steps; ++s) {
    if (s > 0) {
      const uint8_t* src = gbase + (size_t)((s - 1) & (NSLOT - 1)) * SLOT_BYTES + (size_t)rank * SLICE;
      if (warp == 5) {
        if (lane == 0) {
          const unsigned target = (unsigned)s * G;
          unsigned spins = 0;
          while (ld_acquire_u32(a.counter) < target) if (++spins > (1u << 24)) asm volatile("trap;");
        }
        __syncwarp();
        if (tr && lane == 0) a.trace[s * 8 + 1] = clock64();
        asm volatile("fence.proxy.async.global;" ::: "memory");
        if (ptx::elect_one()) {
          ptx::mbar_expect_tx(full0, SLICE);                                   // all CS pieces land here (own + peers' multicasts)
          bulk_g2s_mc(ring + crank * PIECE, src + (size_t)crank * PIECE, PIECE, full0, (uint16_t)((1u << CS) - 1));
        }
        __syncwarp();
      } else if (warp == 4) {
        ptx::mbar_wait(full0, phase);
        if (tr && lane == 0) a.trace[s * 8 + 2] = clock64();
        for (int kb = 0; kb < KB; ++kb)
          for (int part = 0; part < 2; ++part) {
            const uint8_t* tile = gen + (kb * STAGE_BYTES + part * PART_BYTES);
            const int c = 3;
            const uint4 v = *(const uint4*)(tile + lane * 128 + ((c ^ (lane & 7)) << 4));
            const int ug = (rank * KB + kb) * 64 + c * 8;
            if ((unsigned short)(v.x & 0xffff) != bf16_bits(s - 1, part, lane, ug)) ++errs;
          }
        __syncwarp();
        phase ^= 1u;
        if (tr && lane == 0) a.trace[s * 8 + 3] = clock64();
        if (lane == 0) ptx::mbar_arrive(done_bar);
      }
      if (warp < 4) ptx::mbar_wait(done_bar, (uint32_t)((s - 1) & 1));
      // a peer may only overwrite my ring (next step's multicast) after I have consumed this step's data
      asm volatile("barrier.cluster.arrive.release.aligned;" ::: "memory");
      asm volatile("barrier.cluster.wait.acquire.aligned;" ::: "memory");
    }
    if (s == a.steps) break;
    if (warp < 4) {
      if (a.delay > 0) {
        const long long t0 = clock64();
        while (clock64() - t0 < a.delay) {
        }
      }
      if (tr && threadIdx.x == 0) a.trace[s * 8 + 4] = clock64();
      uint8_t* wslot = gbase + (size_t)(s & (NSLOT - 1)) * SLOT_BYTES + (size_t)(blockIdx.x / 8) * STAGE_BYTES;
      const int c = blockIdx.x & 7;
      if (threadIdx.x < 64) {
        const int b = threadIdx.x >> 1, uq = threadIdx.x & 1, ub = u0 + uq * 4;
        uint2 hi, lo;
        hi.x = bf16_bits(s, 0, b, ub) | ((uint32_t)bf16_bits(s, 0, b, ub + 1) << 16);
        hi.y = bf16_bits(s, 0, b, ub + 2) | ((uint32_t)bf16_bits(s, 0, b, ub + 3) << 16);
        lo.x = bf16_bits(s, 1, b, ub) | ((uint32_t)bf16_bits(s, 1, b, ub + 1) << 16);
        lo.y = bf16_bits(s, 1, b, ub + 2) | ((uint32_t)bf16_bits(s, 1, b, ub + 3) << 16);
        const uint32_t off = (uint32_t)(b * 128 + ((c ^ (b & 7)) << 4) + uq * 8);
        *(uint2*)(wslot + off) = hi;
        *(uint2*)(wslot + PART_BYTES + off) = lo;
      }
      if (tr && threadIdx.x == 0) a.trace[s * 8 + 5] = clock64();
      asm volatile("bar.sync 1, 128;" ::: "memory");
      if (threadIdx.x == 0) {
        asm volatile("red.release.gpu.global.add.u32 [%0], 1;" ::"l"(a.counter) : "memory");
        if (tr) a.trace[s * 8 + 6] = clock64();
      }
    }
  }
  if (errs) atomicAdd(a.errors, errs);
}

typedef CUresult (*PFN_encodeTiled)(CUtensorMap*, CUtensorMapDataType, cuuint32_t, void*, const cuuint64_t*, const cuuint64_t*,
                                    const cuuint32_t*, const cuuint32_t*, CUtensorMapInterleave, CUtensorMapSwizzle,
                                    CUtensorMapL2promotion, CUtensorMapFloatOOBfill);

static PFN_encodeTiled get_enc() {
  void* fn = nullptr;
  cudaDriverEntryPointQueryResult q;
  CK(cudaGetDriverEntryPoint("cuTensorMapEncodeTiled", &fn, cudaEnableDefault, &q));
  return (PFN_encodeTiled)fn;
}

static void make_map(CUtensorMap* m, void* basep, uint32_t box_cols, uint32_t box_rows, bool swz) {
  static PFN_encodeTiled enc = get_enc();
  cuuint64_t dims[2] = {NH, BD};
  cuuint64_t strides[1] = {(cuuint64_t)NH * 2};
  cuuint32_t box[2] = {box_cols, box_rows};
  cuuint32_t es[2] = {1, 1};
  CUresult r = enc(m, CU_TENSOR_MAP_DATA_TYPE_BFLOAT16, 2, basep, dims, strides, box, es, CU_TENSOR_MAP_INTERLEAVE_NONE,
                   swz ? CU_TENSOR_MAP_SWIZZLE_128B : CU_TENSOR_MAP_SWIZZLE_NONE, CU_TENSOR_MAP_L2_PROMOTION_L2_128B,
                   CU_TENSOR_MAP_FLOAT_OOB_FILL_NONE);
  if (r != CUDA_SUCCESS) { printf("encode failed %d\n", (int)r); exit(1); }
}

static size_t g_smem_pad = 0;   // extra dynamic shared memory (shrinks the L1: the real kernel runs at the 227 KB carve-out)
template <int PROTO, bool V2 = false>
static void run(const char* name, Args a, const Maps& tm, int steps) {
  const size_t smem = NS * 2 * PART_BYTES + 1024 + 8 * (2 * NS + 2) + 1024 + 64 + g_smem_pad;
  const void* kfn = PROTO == 17 ? (const void*)k_probe3 : (V2 ? (const void*)k_probe2<PROTO> : (const void*)k_probe<PROTO>);
  if (PROTO == 17) CK(cudaMemset(a.abuf, 0, (size_t)NSLOT * 2 * BD * NH * 2));
  CK(cudaFuncSetAttribute(kfn, cudaFuncAttributeMaxDynamicSharedMemorySize, (int)smem));
  cudaEvent_t e0, e1;
  CK(cudaEventCreate(&e0));
  CK(cudaEventCreate(&e1));
  float best = 1e30f;
  for (int rep = 0; rep < 3; ++rep) {
    CK(cudaMemset(a.counter, 0, 256));
    CK(cudaMemset(a.flags, 0, G * 4));
    CK(cudaMemset(a.errors, 0, 4));
    CK(cudaMemset(a.trace, 0, (size_t)(steps + 1) * 8 * 8));
    void* args[] = {(void*)&a, (void*)&tm};
    CK(cudaEventRecord(e0));
    CK(cudaLaunchCooperativeKernel(kfn, dim3(G), dim3(NTHREADS), args, smem, 0));
    CK(cudaEventRecord(e1));
    CK(cudaDeviceSynchronize());
    float ms = 0;
    CK(cudaEventElapsedTime(&ms, e0, e1));
    if (ms < best) best = ms;
  }
  unsigned errs = 0;
  CK(cudaMemcpy(&errs, a.errors, 4, cudaMemcpyDeviceToHost));
  std::vector<unsigned long long> tr((size_t)(steps + 1) * 8);
  CK(cudaMemcpy(tr.data(), a.trace, tr.size() * 8, cudaMemcpyDeviceToHost));
  // CTA 0 phases relative to "stores start" (col 4) of the same step: fence done (5), signalled (6); next step: flags seen (1),
  // first stage landed (2), all landed (3)
  double d_f = 0, d_s = 0, d_seen = 0, d_first = 0, d_all = 0, d_loop = 0, d_cf = 0;
  int n = 0;
  for (int s = 5; s + 1 < steps; ++s) {
    const unsigned long long t4 = tr[s * 8 + 4];
    d_f += (double)(tr[s * 8 + 5] - t4);
    d_s += (double)(tr[s * 8 + 6] - t4);
    d_seen += (double)(tr[(s + 1) * 8 + 1] - t4);
    if (tr[(s + 1) * 8 + 7]) d_cf += (double)(tr[(s + 1) * 8 + 7] - tr[(s + 1) * 8 + 1]);
    d_first += (double)(tr[(s + 1) * 8 + 2] - t4);
    d_all += (double)(tr[(s + 1) * 8 + 3] - t4);
    d_loop += (double)(tr[(s + 1) * 8 + 4] - t4);
    ++n;
  }
  int clk = 0;
  CK(cudaDeviceGetAttribute(&clk, cudaDevAttrClockRate, 0));
  printf("%-44s %8.3f us/step  errors %u | CTA0 cycles from store start: fence %5.0f signal %5.0f seen %5.0f first stage %5.0f all landed %5.0f loop %5.0f | consumer fence %4.0f\n",
         name, best * 1e3 / steps, errs, d_f / n, d_s / n, d_seen / n, d_first / n, d_all / n, d_loop / n, d_cf / n);
}

template <int CS>
static void run_mc(const char* name, Args a, int steps) {
  const size_t smem = NS * 2 * PART_BYTES + 1024 + 8 * (2 * NS + 2) + 1024 + 64 + g_smem_pad;
  auto kfn = k_probe_mc<CS>;
  CK(cudaFuncSetAttribute(kfn, cudaFuncAttributeMaxDynamicSharedMemorySize, (int)smem));
  cudaEvent_t e0, e1;
  CK(cudaEventCreate(&e0));
  CK(cudaEventCreate(&e1));
  float best = 1e30f;
  for (int rep = 0; rep < 3; ++rep) {
    CK(cudaMemset(a.counter, 0, 256));
    CK(cudaMemset(a.errors, 0, 4));
    CK(cudaMemset(a.trace, 0, (size_t)(steps + 1) * 8 * 8));
    cudaLaunchConfig_t cfg{};
    cfg.gridDim = dim3(G);
    cfg.blockDim = dim3(NTHREADS);
    cfg.dynamicSmemBytes = smem;
    cudaLaunchAttribute at[2];
    at[0].id = cudaLaunchAttributeClusterDimension;
    at[0].val.clusterDim.x = CS;
    at[0].val.clusterDim.y = 1;
    at[0].val.clusterDim.z = 1;
    at[1].id = cudaLaunchAttributeCooperative;
    at[1].val.cooperative = 1;
    cfg.attrs = at;
    cfg.numAttrs = 2;
    CK(cudaEventRecord(e0));
    CK(cudaLaunchKernelEx(&cfg, kfn, a));
    CK(cudaEventRecord(e1));
    CK(cudaDeviceSynchronize());
    float ms = 0;
    CK(cudaEventElapsedTime(&ms, e0, e1));
    if (ms < best) best = ms;
  }
  unsigned errs = 0;
  CK(cudaMemcpy(&errs, a.errors, 4, cudaMemcpyDeviceToHost));
  std::vector<unsigned long long> tr((size_t)(steps + 1) * 8);
  CK(cudaMemcpy(tr.data(), a.trace, tr.size() * 8, cudaMemcpyDeviceToHost));
  double d_s = 0, d_seen = 0, d_first = 0, d_loop = 0;
  int n = 0;
  for (int s = 5; s + 1 < steps; ++s) {
    const unsigned long long t4 = tr[s * 8 + 4];
    d_s += (double)(tr[s * 8 + 6] - t4);
    d_seen += (double)(tr[(s + 1) * 8 + 1] - t4);
    d_first += (double)(tr[(s + 1) * 8 + 2] - t4);
    d_loop += (double)(tr[(s + 1) * 8 + 4] - t4);
    ++n;
  }
  printf("%-44s %8.3f us/step  errors %u | CTA0 cycles from store start: signal %5.0f seen %5.0f slice landed %5.0f loop %5.0f\n", name,
         best * 1e3 / steps, errs, d_s / n, d_seen / n, d_first / n, d_loop / n);
}

int main(int argc, char** argv) {
  const int steps = argc > 1 ? atoi(argv[1]) : 400;
  const int delay = argc > 2 ? atoi(argv[2]) : 0;
  g_smem_pad = argc > 3 ? (size_t)atoi(argv[3]) * 1024 : 0;
  Args a{};
  a.steps = steps;
  a.delay = delay;
  CK(cudaMalloc(&a.abuf, (size_t)NREP * NSLOT * 2 * BD * NH * 2));
  CK(cudaMemset(a.abuf, 0, (size_t)NREP * NSLOT * 2 * BD * NH * 2));
  CK(cudaMalloc(&a.counter, 256));
  CK(cudaMalloc(&a.flags, G * 4));
  CK(cudaMalloc(&a.errors, 4));
  CK(cudaMalloc(&a.trace, (size_t)(steps + 1) * 8 * 8));
  Maps tm;
  for (int slot = 0; slot < NSLOT; ++slot)
    for (int part = 0; part < 2; ++part) {
      void* b = a.abuf + ((size_t)slot * 2 + part) * BD * NH;
      make_map(&tm.ld[slot * 2 + part], b, 64, BD, true);
      make_map(&tm.st[slot * 2 + part], b, 8, BD, false);
    }
  cudaDeviceProp prop;
  CK(cudaGetDeviceProperties(&prop, 0));
  printf("device %s, %d SMs, steps %d, simulated compute delay %d cycles, smem pad %zu KB\n", prop.name, prop.multiProcessorCount, steps, delay,
         g_smem_pad >> 10);
  run<0>("P0 counter + both proxy fences (current)", a, tm, steps);
  run<1>("P1 counter, consumer-side proxy fence only", a, tm, steps);
  run<2>("P2 per-CTA flags, TMA per ready k-block", a, tm, steps);
  run<3>("P3 per-CTA flags, all 64 before first TMA", a, tm, steps);
  run<4>("P4 flags + bulk TMA store from smem", a, tm, steps);
  run<5>("P5 flags, consumer-side proxy fence only", a, tm, steps);
  run<6, true>("P6 tile image: counter + 8 x 8KB bulk", a, tm, steps);
  run<7, true>("P7 tile image: per-slice counter + 8 bulk", a, tm, steps);
  run<8, true>("P8 tile image: flags(ld.acquire) + bulk/kb", a, tm, steps);
  run<9, true>("P9 tile image: per-slice counter + 1 x 64KB", a, tm, steps);
  run<10, true>("P10 tile image: per-slice ctr, warp arrivals", a, tm, steps);
  run<11, true>("P11 P6 + staggered k-block start", a, tm, steps);
  run<12, true>("P12 P6 with 2 x 32KB copies", a, tm, steps);
  run<13, true>("P13 P6 with NO proxy fence", a, tm, steps);
  run<14, true>("P14 P6 with fence.proxy.async.global", a, tm, steps);
  run<15, true>("P15 generic LDG->STS copy by 4 warps", a, tm, steps);
  run<16, true>("P16 staggered start, 4 x 16KB copies", a, tm, steps);
  run<17, true>("P17 validity-in-data, LDG poll by 4 warps", a, tm, steps);
  run<20, true>("P20 P9 (1 x 64KB bulk) + 4 replicas", a, tm, steps);
  run<21, true>("P21 P15 (LDG copy) + 4 replicas", a, tm, steps);
  run<22, true>("P22 P6 (8 x 8KB bulk) + 4 replicas", a, tm, steps);
  run_mc<1>("P23a cluster 1 (unicast 64KB, .global fence)", a, steps);
  run_mc<2>("P23b cluster 2, 2 x 32KB multicast", a, steps);
  run_mc<4>("P23c cluster 4, 4 x 16KB multicast", a, steps);
  run<18, true>("P18 P15 + concurrent ld.acquire spinner", a, tm, steps);
  run<19, true>("P19 P15 + relaxed/nanosleep spinner", a, tm, steps);
  return 0;
}
