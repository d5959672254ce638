// Dense projection Y[M,N] = X[M,K] * W[N,K]^T for the decoder's QKV / O / gate_up / down / lm_head
// linears (reference operator: F.linear in $VLLM/model_executor/layers/utils.py:92-98, which
// dispatches to cuBLASLt).  Hand-written for sm_100a:
//
//   * operands are staged by TMA (cp.async.bulk.tensor, 128-byte swizzle) into a ring of shared
//     memory stages guarded by full/empty mbarriers;
//   * one elected thread issues tcgen05.mma (kind::f16, bf16 x bf16 -> fp32) with the accumulator
//     in tensor memory, double-buffered so the epilogue of tile i overlaps the main loop of i+1;
//   * four epilogue warps read the accumulator back with tcgen05.ld (one weight row per thread),
//     transpose 32 x 32 blocks through shared memory and hand them to TMA STORES
//     (cp.async.bulk.tensor, clipped by the hardware at the M / N edges) -- bf16, fp32 logits or
//     fp32 split-K partial planes;
//   * the kernel is persistent: grid = #SMs, tiles are walked with a grouped raster so that a
//     wave shares activation tiles through L2.
//
// Orientation ("swap-AB"): the WEIGHT tile is the MMA's M operand (128 weight rows = 128 TMEM
// lanes) and the ACTIVATION tile is the N operand (BLOCK_X tokens = BLOCK_X fp32 TMEM columns).
// A decode step with M = 1..256 tokens is therefore ONE tile column: every weight byte is read
// from HBM exactly once and no tensor-core work is spent on padding rows; a prefill step uses
// 128 x 256 tiles.  Weight streaming is the roofline for decode (HBM), the tensor pipe for prefill.
#include "common.cuh"
#include "runtime.h"
#include "../../include/b200.h"

#include <stdlib.h>

namespace b200 {

constexpr int kBlockW = 128;  // weight rows per tile  (UMMA M)
constexpr int kBlockK = 64;   // 64 bf16 = 128 bytes = one swizzle row
constexpr int kUmmaK = 16;
constexpr int kGemmThreads = 192;  // warp0 TMA, warp1 MMA+TMEM alloc, warps 2-5 epilogue
// epilogue staging: every epilogue warp owns two 4 KB buffers (a 32 x 32 block of fp32, or of bf16
// in the first half) that alternate between "being filled" and "being read by a TMA store"
constexpr int kStgBufBytes = 4096;
constexpr int kStgWarpBytes = 2 * kStgBufBytes;
constexpr int kStgBytes = 4 * kStgWarpBytes;
constexpr int kRingBudget = 192 * 1024;  // operand ring; + staging + barriers + alignment <= 227 KB
constexpr int kBarBytes = 256;

template <int BLOCK_X>
struct GemmCfg {
  static constexpr int kABytes = kBlockW * kBlockK * 2;
  static constexpr int kBBytes = BLOCK_X * kBlockK * 2;
  static constexpr int kStageBytes = kABytes + kBBytes;
  // as many stages as fit in 192 KB: deep rings keep >100 KB of weight loads in flight per SM,
  // which is what an HBM-bound decode tile needs (latency x per-SM share of 6.5 TB/s).
  static constexpr int kStages = kRingBudget / kStageBytes > 12 ? 12 : kRingBudget / kStageBytes;
  static constexpr int kTmemCols = 2 * BLOCK_X < 32 ? 32 : 2 * BLOCK_X;
  static constexpr int kSmemBytes = kStages * kStageBytes + kStgBytes + 1024 /*align*/ + kBarBytes;
};

struct TileCoord {
  int split, m_tile, n_tile;
};

// Grouped raster: tiles are ordered (m-group, n, m-in-group, split) so that concurrently running
// CTAs share a handful of activation tiles and each weight tile is pulled from HBM once per group.
__device__ __forceinline__ TileCoord decode_unit(int u, int n_tiles, int m_tiles, int splits,
                                                 int group_m) {
  TileCoord c;
  c.split = u % splits;
  int t = u / splits;
  const int per_group = group_m * n_tiles;
  const int g = t / per_group;
  const int first_m = g * group_m;
  const int gm = min(group_m, m_tiles - first_m);
  const int r = t - g * per_group;
  c.n_tile = r / gm;
  c.m_tile = first_m + r % gm;
  return c;
}

// TMEM accumulator (this warp's 32 lanes = 32 weight rows x BLOCK_X token columns) -> global memory.
// A thread owns ONE output column n and BLOCK_X rows m of Y[m, n]: written directly that would be
// one 2-byte store per element at stride ldy.  Instead each 32-column chunk is transposed through
// this warp's staging buffer ([m][n] rows of 64 / 128 / 32 bytes, conflict-free: the 32 lanes of a
// store instruction fill one row) and leaves as ONE TMA store per chunk; the hardware clips rows
// >= M and columns >= N, so tail tiles need no predicates.  Two buffers alternate so that chunk i+1
// is filled while the store of chunk i still reads its buffer.
// EPI 0: bf16 / fp32 store; EPI 1: SiLU*mul of interleaved gate/up rows (see the kernel comment).
// FP8 1: W8A8 dequantisation -- every accumulator is multiplied by the token's dynamic activation
//        scale xs[m] and the weight row's scale ws (already loaded by the caller) before any
//        rounding, as cutlass_scaled_mm does in the reference engine.
// n0: first output column (weight row) of this warp; `which` (in/out): staging buffer to fill next.
// DirectOut: fallback for outputs whose row pitch is not a multiple of 16 bytes (a TMA tensor map
// cannot describe them): plain per-thread stores, one element each -- correct, slow, and only ever
// taken by odd-shaped test matrices (every tensor of the engine is 16-byte pitched).
struct DirectOut {
  void* ptr;      // nullptr: TMA-store path
  int64_t ldy;
};

template <int BLOCK_X, int EPI, int FP8 = 0>
__device__ __forceinline__ void epilogue_store_tile(uint32_t t_addr, const CUtensorMap* tmap_y, uint8_t* stg,
                                                    int& which, int n0, int m_base, int split, int M, int N,
                                                    int out_f32, int lane, DirectOut dir,
                                                    const float* __restrict__ xs = nullptr, float ws = 1.f) {
  constexpr int CH = BLOCK_X >= 32 ? 32 : 16;
#pragma unroll 1
  for (int c0 = 0; c0 < BLOCK_X; c0 += CH) {
    if (m_base + c0 >= M) break;  // warp-uniform: rest of the tile is padding
    uint32_t r[CH];
    if constexpr (CH == 32) tmem_ld_32x32(t_addr + c0, r);
    else tmem_ld_32x16(t_addr + c0, r);
    float xs_lane = 0.f;
    if constexpr (FP8) {
      const int m = m_base + c0 + lane;   // lane j holds the scale of column j of this chunk
      xs_lane = (lane < CH && m < M) ? xs[m] : 0.f;
    }
    // the store that read this buffer two chunks ago must have drained it before it is refilled
    if (lane == 0) bulk_wait_read<1>();
    __syncwarp();
    tmem_ld_wait();
    if constexpr (FP8) {
#pragma unroll
      for (int j = 0; j < CH; ++j)
        r[j] = __float_as_uint(__uint_as_float(r[j]) * (__shfl_sync(0xffffffffu, xs_lane, j) * ws));
    }
    if (dir.ptr != nullptr) {   // warp-uniform
      const int n = n0 + lane;
      if constexpr (EPI == 1) {
        bf16* dst = reinterpret_cast<bf16*>(dir.ptr) + (size_t)(m_base + c0) * dir.ldy + (n >> 1);
        const bool odd = lane & 1;
#pragma unroll
        for (int j = 0; j < CH; j += 2) {
          const uint32_t send = odd ? r[j] : r[j + 1];
          const uint32_t recv = __shfl_xor_sync(0xffffffffu, send, 1);
          const float g = round_bf16(__uint_as_float(odd ? recv : r[j]));
          const float up = round_bf16(__uint_as_float(odd ? r[j + 1] : recv));
          const float sg = round_bf16(g / (1.f + __expf(-g)));
          const int m = m_base + c0 + j + (odd ? 1 : 0);
          if (n < N && m < M) dst[(size_t)(j + (odd ? 1 : 0)) * dir.ldy] = __float2bfloat16_rn(sg * up);
        }
      } else if (n < N) {
        const size_t off = ((size_t)split * M + (size_t)(m_base + c0)) * dir.ldy + n;
#pragma unroll
        for (int j = 0; j < CH; ++j) {
          if (m_base + c0 + j >= M) break;
          if (out_f32) reinterpret_cast<float*>(dir.ptr)[off + (size_t)j * dir.ldy] = __uint_as_float(r[j]);
          else reinterpret_cast<bf16*>(dir.ptr)[off + (size_t)j * dir.ldy] = __float2bfloat16_rn(__uint_as_float(r[j]));
        }
      }
      continue;
    }
    uint8_t* buf = stg + which * kStgBufBytes;
    const uint32_t sbuf = smem_u32(buf);
    if constexpr (EPI == 1) {
      // lane pair (2p, 2p+1) = (gate, up) rows of output channel n0/2 + p; staging rows are 16 channels
      const bool odd = lane & 1;
      const uint32_t col = sbuf + (uint32_t)(lane >> 1) * 2u;
#pragma unroll
      for (int j = 0; j < CH; j += 2) {
        // even lane (gate row) computes token j, odd lane (up row) computes token j+1
        const uint32_t send = odd ? r[j] : r[j + 1];
        const uint32_t recv = __shfl_xor_sync(0xffffffffu, send, 1);
        const float g = round_bf16(__uint_as_float(odd ? recv : r[j]));
        const float up = round_bf16(__uint_as_float(odd ? r[j + 1] : recv));
        const float sg = round_bf16(g / (1.f + __expf(-g)));
        const bf16 v = __float2bfloat16_rn(sg * up);
        st_shared_u16(col + (uint32_t)(j + (odd ? 1 : 0)) * 32u, *reinterpret_cast<const uint16_t*>(&v));
      }
    } else if (out_f32) {
      const uint32_t col = sbuf + (uint32_t)lane * 4u;
#pragma unroll
      for (int j = 0; j < CH; ++j) st_shared_u32(col + (uint32_t)j * 128u, r[j]);
    } else {
      const uint32_t col = sbuf + (uint32_t)lane * 2u;
#pragma unroll
      for (int j = 0; j < CH; ++j) {
        const bf16 v = __float2bfloat16_rn(__uint_as_float(r[j]));
        st_shared_u16(col + (uint32_t)j * 64u, *reinterpret_cast<const uint16_t*>(&v));
      }
    }
    fence_proxy_async();   // generic-proxy writes above -> visible to the TMA (async proxy) read below
    __syncwarp();
    if (lane == 0 && n0 < N) {
      tma_store_3d(tmap_y, buf, EPI == 1 ? (n0 >> 1) : n0, m_base + c0, split);
      bulk_commit();
    }
    which ^= 1;
  }
}

// EPI 0: store the accumulator (bf16, or fp32 / split-K partials).
// EPI 1: SiLU*mul epilogue for the gate_up projection ($VLLM activation.py:133-143 fused into the
//        producing GEMM): W rows are interleaved (row 2i = gate_i, row 2i+1 = up_i), so the two
//        values of an output channel sit in adjacent TMEM lanes = adjacent threads of one warp;
//        a lane pair swaps halves with one shuffle per two tokens and writes
//        act[m, i] = bf16(bf16(silu(bf16(g))) * bf16(u)) -- the same roundings as the unfused path.
// MC 2: the kernel runs as clusters of two CTAs that own ADJACENT weight tiles of the same
//       (m-tile, k-split).  Each CTA loads its own weight tile and HALF of the shared activation
//       tile, multicast into both CTAs' shared memory, so the activation bytes an SM has to ingest
//       from L2 are halved (the measured limiter of decode-shaped tiles).  Stage release is a
//       multicast tcgen05.commit: a slot is reused only after BOTH CTAs' MMAs have read it.
template <int BLOCK_X, int EPI, int MC>
__global__ void __launch_bounds__(kGemmThreads, 1)
gemm_bf16_tcgen05_kernel(const __grid_constant__ CUtensorMap tmap_w,
                         const __grid_constant__ CUtensorMap tmap_x,
                         const __grid_constant__ CUtensorMap tmap_y, int M, int N, int K, int out_f32,
                         int splits, int n_tiles, int m_tiles, int group_m, DirectOut dir) {
  using Cfg = GemmCfg<BLOCK_X>;
  constexpr int STAGES = Cfg::kStages;
  extern __shared__ uint8_t smem_raw[];
  uint8_t* smem = reinterpret_cast<uint8_t*>(
      (reinterpret_cast<uintptr_t>(smem_raw) + 1023) & ~static_cast<uintptr_t>(1023));
  uint8_t* stg_base = smem + STAGES * Cfg::kStageBytes;   // epilogue staging (1024-byte aligned)
  uint64_t* full_bar = reinterpret_cast<uint64_t*>(stg_base + kStgBytes);
  uint64_t* empty_bar = full_bar + STAGES;
  uint64_t* tmem_full = empty_bar + STAGES;
  uint64_t* tmem_empty = tmem_full + 2;
  uint32_t* tmem_slot = reinterpret_cast<uint32_t*>(tmem_empty + 2);

  const int warp = threadIdx.x >> 5;
  const int lane = threadIdx.x & 31;
  const uint32_t crank = MC > 1 ? cluster_ctarank() : 0u;
  pdl_launch_dependents();  // the next kernel may start its prologue while this one runs

  if (threadIdx.x == 0) {
    tma_prefetch_desc(&tmap_w);
    tma_prefetch_desc(&tmap_x);
    tma_prefetch_desc(&tmap_y);
    for (int i = 0; i < STAGES; ++i) {
      mbar_init(&full_bar[i], 1);
      mbar_init(&empty_bar[i], MC);  // one multicast commit per CTA of the cluster
    }
    for (int i = 0; i < 2; ++i) {
      mbar_init(&tmem_full[i], 1);
      mbar_init(&tmem_empty[i], 4);  // one arrive per epilogue warp
    }
    fence_mbar_init();
  }
  if (warp == 1) {
    tmem_alloc(tmem_slot, Cfg::kTmemCols);
    tmem_relinquish();
  }
  tc_fence_before();
  __syncthreads();
  if constexpr (MC > 1) cluster_sync_all();  // peers' barriers are initialised before any remote arrive
  tc_fence_after();
  const uint32_t tmem_base = *tmem_slot;
  // everything above (barrier init, TMEM allocation, descriptor prefetch) overlapped with the tail
  // of the previous kernel; from here on we touch its outputs
  pdl_wait();

  const int kb_per_split = (K / kBlockK) / splits;
  // with MC == 2 a "unit" is a PAIR of adjacent weight tiles; this CTA takes tile 2*pair + crank
  const int n_units = (n_tiles + MC - 1) / MC;
  const int total_units = n_units * m_tiles * splits;
  const int first_unit = MC > 1 ? (int)(blockIdx.x / MC) : (int)blockIdx.x;
  const int unit_stride = MC > 1 ? (int)(gridDim.x / MC) : (int)gridDim.x;

  if (warp == 0) {
    // ------------------------------------------------------------------ TMA producer
    if (elect_one()) {
      const uint64_t hint_w = (m_tiles > 1) ? kEvictNormal : kEvictFirst;
      int stage = 0;
      uint32_t phase = 0;
      for (int u = first_unit; u < total_units; u += unit_stride) {
        const TileCoord c = decode_unit(u, n_units, m_tiles, splits, group_m);
        const int n_tile = c.n_tile * MC + (int)crank;
        const int k0 = c.split * kb_per_split * kBlockK;
        for (int kb = 0; kb < kb_per_split; ++kb) {
          mbar_wait_guarded(&empty_bar[stage], phase ^ 1);
          uint8_t* a = smem + stage * Cfg::kStageBytes;
          uint8_t* b = a + Cfg::kABytes;
          mbar_expect_tx(&full_bar[stage], Cfg::kStageBytes);
          tma_load_2d(a, &tmap_w, &full_bar[stage], k0 + kb * kBlockK, n_tile * kBlockW, hint_w);
          if constexpr (MC > 1) {
            constexpr int HALF = BLOCK_X / MC;  // rows of the activation tile this CTA fetches
            tma_load_2d_mc(b + crank * (HALF * kBlockK * 2), &tmap_x, &full_bar[stage],
                           k0 + kb * kBlockK, c.m_tile * BLOCK_X + (int)crank * HALF,
                           (uint16_t)((1u << MC) - 1), kEvictLast);
          } else {
            tma_load_2d(b, &tmap_x, &full_bar[stage], k0 + kb * kBlockK, c.m_tile * BLOCK_X,
                        kEvictLast);
          }
          if (++stage == STAGES) { stage = 0; phase ^= 1; }
        }
      }
    }
  } else if (warp == 1) {
    // ------------------------------------------------------------------ MMA issuer
    if (elect_one()) {
      constexpr uint32_t idesc = umma_idesc_bf16(kBlockW, BLOCK_X);
      int stage = 0;
      uint32_t phase = 0;
      int acc = 0;
      uint32_t acc_phase = 0;
      for (int u = first_unit; u < total_units; u += unit_stride) {
        mbar_wait_guarded(&tmem_empty[acc], acc_phase ^ 1);
        tc_fence_after();
        const uint32_t d_tmem = tmem_base + (uint32_t)(acc * BLOCK_X);
        for (int kb = 0; kb < kb_per_split; ++kb) {
          mbar_wait_guarded(&full_bar[stage], phase);
          tc_fence_after();
          const uint32_t a_addr = smem_u32(smem + stage * Cfg::kStageBytes);
          const uint32_t b_addr = a_addr + Cfg::kABytes;
#pragma unroll
          for (int k = 0; k < kBlockK / kUmmaK; ++k) {
            umma_bf16(d_tmem, umma_desc_k_sw128(a_addr + k * kUmmaK * 2),
                      umma_desc_k_sw128(b_addr + k * kUmmaK * 2), idesc, (kb | k) != 0);
          }
          // frees the smem stage (in every CTA that multicasts into it) once these MMAs retire
          if constexpr (MC > 1) umma_commit_mc(&empty_bar[stage], (uint16_t)((1u << MC) - 1));
          else umma_commit(&empty_bar[stage]);
          if (++stage == STAGES) { stage = 0; phase ^= 1; }
        }
        umma_commit(&tmem_full[acc]);  // accumulator complete -> epilogue
        if (++acc == 2) { acc = 0; acc_phase ^= 1; }
      }
    }
  } else {
    // ------------------------------------------------------------------ epilogue (warps 2..5)
    const int quad = warp & 3;  // a warp may only touch TMEM lanes [32*(warp%4), +32)
    uint8_t* stg = stg_base + quad * kStgWarpBytes;
    int which = 0;
    int acc = 0;
    uint32_t acc_phase = 0;
    for (int u = first_unit; u < total_units; u += unit_stride) {
      const TileCoord c = decode_unit(u, n_units, m_tiles, splits, group_m);
      mbar_wait_guarded(&tmem_full[acc], acc_phase);
      tc_fence_after();
      const int n0 = (c.n_tile * MC + (int)crank) * kBlockW + quad * 32;
      const int m_base = c.m_tile * BLOCK_X;
      const uint32_t t_addr = tmem_base + ((uint32_t)(quad * 32) << 16) + (uint32_t)(acc * BLOCK_X);
      epilogue_store_tile<BLOCK_X, EPI>(t_addr, &tmap_y, stg, which, n0, m_base, c.split, M, N, out_f32, lane, dir);
      tc_fence_before();
      __syncwarp();
      if (lane == 0) mbar_arrive(&tmem_empty[acc]);
      if (++acc == 2) { acc = 0; acc_phase ^= 1; }
    }
    if (lane == 0) bulk_wait<0>();   // every TMA store of this warp has been performed
  }

  tc_fence_before();
  __syncthreads();
  if constexpr (MC > 1) cluster_sync_all();  // no CTA leaves while its peer may still write into it
  if (warp == 1) {
    tc_fence_after();
    tmem_dealloc(tmem_base, Cfg::kTmemCols);
  }
}

// =================================================================================================
// 2-SM variant (tcgen05.mma.cta_group::2): a CTA PAIR computes one 256 (weight rows) x BLOCK_X
// (tokens) tile.  Each CTA stages its own 128 weight rows and HALF of the activation tile; the
// pair's tensor cores read both halves, so per SM the shared-memory operand traffic of a k-block
// drops from 16 KB + BLOCK_X*128 B to 16 KB + BLOCK_X*64 B and the ring gets deeper.  The leader
// (cluster rank 0) issues every MMA; data arrival from both CTAs is tracked on the leader's full
// barriers (TMA .cta_group::2 + remote expect_tx), stage release and accumulator hand-off are
// multicast commits to both CTAs, and both CTAs' epilogue warps release the accumulator on the
// leader's barrier.
template <int BLOCK_X>
struct Gemm2Cfg {
  static constexpr int kHalf = BLOCK_X / 2;
  static constexpr int kABytes = kBlockW * kBlockK * 2;
  static constexpr int kBBytes = kHalf * kBlockK * 2;
  static constexpr int kStageBytes = kABytes + kBBytes;
  static constexpr int kStages = kRingBudget / kStageBytes > 10 ? 10 : kRingBudget / kStageBytes;
  static constexpr int kTmemCols = 2 * BLOCK_X < 32 ? 32 : 2 * BLOCK_X;
  static constexpr int kSmemBytes = kStages * kStageBytes + kStgBytes + 1024 + kBarBytes;
};

// FP8 1: e4m3 x e4m3 -> fp32 (kind::f8f6f4).  A k-block is still one 128-byte swizzle row, i.e.
// 128 elements instead of 64, and one MMA consumes 32 bytes of K either way, so the pipeline is
// byte-for-byte the bf16 one; only the element counts handed to TMA, the instruction kind and the
// dequantising epilogue differ.
// GROUPED 1: mixture-of-experts grouped GEMM.  The token rows live in a permuted row space in which
// every m-tile (BLOCK_X rows) belongs to ONE expert (moe.cu: moe_align_kernel); W stacks the
// experts' [N, K] matrices along its rows, so tile (m, n) multiplies rows of X with weight rows
// tile_expert[m] * N + n*128...  The number of m-tiles in use is read from device memory after the
// dependency wait (no host round trip, graph capturable); `xscale`, when given, is the per-row
// routing weight applied to the fp32 accumulator (MUL_ROUTED_WEIGHT of the reference's fused_moe).
struct GroupInfo {
  const int32_t* tile_expert;   // [max m_tiles]
  const int32_t* num_m_tiles;   // device scalar
};

template <int BLOCK_X, int EPI, int FP8, int GROUPED = 0>
__device__ __forceinline__ void gemm_2sm_body(const CUtensorMap& tmap_w, const CUtensorMap& tmap_x,
                                              const CUtensorMap& tmap_y, int M, int N, int K,
                                              int out_f32, int splits, int n_tiles, int m_tiles,
                                              int group_m, const float* __restrict__ xscale,
                                              const float* __restrict__ wscale, int wscale_stride,
                                              DirectOut dir, GroupInfo grp = GroupInfo{nullptr, nullptr}) {
  using Cfg = Gemm2Cfg<BLOCK_X>;
  constexpr int kElemsK = FP8 ? 2 * kBlockK : kBlockK;   // elements per 128-byte k-block
  constexpr int STAGES = Cfg::kStages;
  // bits 16.. of group_m: L2 eviction-priority overrides for the two operand streams (A/B runs,
  // B200_GEMM_HINTS; 0 = the built-in choice)
  const int hint_code = group_m >> 16;
  group_m &= 0xffff;
  extern __shared__ uint8_t smem_raw[];
  uint8_t* smem = reinterpret_cast<uint8_t*>(
      (reinterpret_cast<uintptr_t>(smem_raw) + 1023) & ~static_cast<uintptr_t>(1023));
  uint8_t* stg_base = smem + STAGES * Cfg::kStageBytes;   // epilogue staging (1024-byte aligned)
  uint64_t* full_bar = reinterpret_cast<uint64_t*>(stg_base + kStgBytes);
  uint64_t* empty_bar = full_bar + STAGES;
  uint64_t* tmem_full = empty_bar + STAGES;
  uint64_t* tmem_empty = tmem_full + 2;
  uint32_t* tmem_slot = reinterpret_cast<uint32_t*>(tmem_empty + 2);

  const int warp = threadIdx.x >> 5;
  const int lane = threadIdx.x & 31;
  const uint32_t crank = cluster_ctarank();
  const bool leader = crank == 0;
  pdl_launch_dependents();

  if (threadIdx.x == 0) {
    tma_prefetch_desc(&tmap_w);
    tma_prefetch_desc(&tmap_x);
    tma_prefetch_desc(&tmap_y);
    for (int i = 0; i < STAGES; ++i) {
      mbar_init(&full_bar[i], 2);   // (leader's copy is the live one) one expect_tx arrive per CTA
      mbar_init(&empty_bar[i], 1);  // one multicast commit from the leader
    }
    for (int i = 0; i < 2; ++i) {
      mbar_init(&tmem_full[i], 1);   // multicast commit from the leader
      mbar_init(&tmem_empty[i], 8);  // (leader's copy) 4 epilogue warps x 2 CTAs
    }
    fence_mbar_init();
  }
  if (warp == 1) {
    tmem_alloc_2sm(tmem_slot, Cfg::kTmemCols);
    tmem_relinquish_2sm();
  }
  tc_fence_before();
  __syncthreads();
  cluster_sync_all();
  tc_fence_after();
  const uint32_t tmem_base = *tmem_slot;
  pdl_wait();

  if constexpr (GROUPED) {
    const int used = *grp.num_m_tiles;   // written by the align kernel this grid depends on
    m_tiles = used < m_tiles ? used : m_tiles;
  }
  const int kb_per_split = (K / kElemsK) / splits;
  const int n_units = (n_tiles + 1) / 2;  // a unit = a pair of adjacent weight tiles (256 rows)
  const int total_units = n_units * m_tiles * splits;
  const int first_unit = (int)(blockIdx.x / 2);
  const int unit_stride = (int)(gridDim.x / 2);

  if (warp == 0) {
    // ------------------------------------------------------------------ TMA producer (both CTAs)
    if (elect_one()) {
      uint64_t hint_w = (m_tiles > 1) ? kEvictNormal : kEvictFirst;
      uint64_t hint_x = kEvictLast;
      {
        const uint64_t tab[4] = {0, kEvictFirst, kEvictNormal, kEvictLast};
        if (hint_code & 3) hint_w = tab[hint_code & 3];
        if ((hint_code >> 2) & 3) hint_x = tab[(hint_code >> 2) & 3];
      }
      int stage = 0;
      uint32_t phase = 0;
      for (int u = first_unit; u < total_units; u += unit_stride) {
        const TileCoord c = decode_unit(u, n_units, m_tiles, splits, group_m);
        const int n_tile = c.n_tile * 2 + (int)crank;
        int w_row = n_tile * kBlockW;
        if constexpr (GROUPED) w_row += grp.tile_expert[c.m_tile] * N;
        const int k0 = c.split * kb_per_split * kElemsK;
        for (int kb = 0; kb < kb_per_split; ++kb) {
          mbar_wait_guarded(&empty_bar[stage], phase ^ 1);
          uint8_t* a = smem + stage * Cfg::kStageBytes;
          uint8_t* b = a + Cfg::kABytes;
          // this CTA's bytes are announced on the LEADER's barrier, where the TMAs complete
          mbar_expect_tx_cluster(mapa_u32(&full_bar[stage], 0), Cfg::kStageBytes);
          tma_load_2d_2sm(a, &tmap_w, &full_bar[stage], k0 + kb * kElemsK, w_row, hint_w);
          tma_load_2d_2sm(b, &tmap_x, &full_bar[stage], k0 + kb * kElemsK,
                          c.m_tile * BLOCK_X + (int)crank * Cfg::kHalf, hint_x);
          if (++stage == STAGES) { stage = 0; phase ^= 1; }
        }
      }
    }
  } else if (warp == 1) {
    // ------------------------------------------------------------------ MMA issuer (leader only)
    if (leader && elect_one()) {
      constexpr uint32_t idesc = FP8 ? umma_idesc_e4m3(2 * kBlockW, BLOCK_X) : umma_idesc_bf16(2 * kBlockW, BLOCK_X);
      int stage = 0;
      uint32_t phase = 0;
      int acc = 0;
      uint32_t acc_phase = 0;
      for (int u = first_unit; u < total_units; u += unit_stride) {
        mbar_wait_guarded(&tmem_empty[acc], acc_phase ^ 1);
        tc_fence_after();
        const uint32_t d_tmem = tmem_base + (uint32_t)(acc * BLOCK_X);
        for (int kb = 0; kb < kb_per_split; ++kb) {
          mbar_wait_guarded(&full_bar[stage], phase);
          tc_fence_after();
          const uint32_t a_addr = smem_u32(smem + stage * Cfg::kStageBytes);
          const uint32_t b_addr = a_addr + Cfg::kABytes;
#pragma unroll
          for (int k = 0; k < kBlockK / kUmmaK; ++k) {   // 4 MMAs x 32 bytes of K per k-block
            if constexpr (FP8)
              umma_e4m3_2sm(d_tmem, umma_desc_k_sw128(a_addr + k * kUmmaK * 2),
                            umma_desc_k_sw128(b_addr + k * kUmmaK * 2), idesc, (kb | k) != 0);
            else
              umma_bf16_2sm(d_tmem, umma_desc_k_sw128(a_addr + k * kUmmaK * 2),
                            umma_desc_k_sw128(b_addr + k * kUmmaK * 2), idesc, (kb | k) != 0);
          }
          umma_commit_2sm_mc(&empty_bar[stage], 3);  // both CTAs' slots are free once these retire
          if (++stage == STAGES) { stage = 0; phase ^= 1; }
        }
        umma_commit_2sm_mc(&tmem_full[acc], 3);      // accumulator complete in both CTAs' TMEM
        if (++acc == 2) { acc = 0; acc_phase ^= 1; }
      }
    }
  } else {
    // ------------------------------------------------------------------ epilogue (both CTAs)
    const int quad = warp & 3;
    uint8_t* stg = stg_base + quad * kStgWarpBytes;
    int which = 0;
    int acc = 0;
    uint32_t acc_phase = 0;
    for (int u = first_unit; u < total_units; u += unit_stride) {
      const TileCoord c = decode_unit(u, n_units, m_tiles, splits, group_m);
      mbar_wait_guarded(&tmem_full[acc], acc_phase);
      tc_fence_after();
      const int n0 = (c.n_tile * 2 + (int)crank) * kBlockW + quad * 32;
      const int n = n0 + lane;
      const int m_base = c.m_tile * BLOCK_X;
      const uint32_t t_addr = tmem_base + ((uint32_t)(quad * 32) << 16) + (uint32_t)(acc * BLOCK_X);
      if constexpr (FP8) {
        const float ws = n < N ? wscale[(size_t)n * wscale_stride] : 0.f;
        epilogue_store_tile<BLOCK_X, EPI, 1>(t_addr, &tmap_y, stg, which, n0, m_base, c.split, M, N, out_f32, lane,
                                             dir, xscale, ws);
      } else if (GROUPED && xscale != nullptr) {
        epilogue_store_tile<BLOCK_X, EPI, 1>(t_addr, &tmap_y, stg, which, n0, m_base, c.split, M, N, out_f32, lane,
                                             dir, xscale, 1.f);
      } else {
        epilogue_store_tile<BLOCK_X, EPI>(t_addr, &tmap_y, stg, which, n0, m_base, c.split, M, N, out_f32, lane, dir);
      }
      tc_fence_before();
      __syncwarp();
      if (lane == 0) mbar_arrive_cluster(mapa_u32(&tmem_empty[acc], 0));
      if (++acc == 2) { acc = 0; acc_phase ^= 1; }
    }
    if (lane == 0) bulk_wait<0>();   // every TMA store of this warp has been performed
  }

  tc_fence_before();
  __syncthreads();
  cluster_sync_all();
  if (warp == 1) {
    tc_fence_after();
    tmem_dealloc_2sm(tmem_base, Cfg::kTmemCols);
  }
}

template <int BLOCK_X, int EPI>
__global__ void __launch_bounds__(kGemmThreads, 1)
gemm_bf16_tcgen05_2sm_kernel(const __grid_constant__ CUtensorMap tmap_w,
                             const __grid_constant__ CUtensorMap tmap_x,
                             const __grid_constant__ CUtensorMap tmap_y, int M, int N, int K, int out_f32,
                             int splits, int n_tiles, int m_tiles, int group_m, DirectOut dir) {
  gemm_2sm_body<BLOCK_X, EPI, 0>(tmap_w, tmap_x, tmap_y, M, N, K, out_f32, splits, n_tiles, m_tiles,
                                 group_m, nullptr, nullptr, 0, dir);
}

// W8A8 projection of the fp8 checkpoints GPUStack deploys (BASELINE config "Llama-3-70B fp8"):
// Y[m,n] = (sum_k Xq[m,k] * Wq[n,k]) * xscale[m] * wscale[n * wscale_stride]
// (reference operator: cutlass_scaled_mm behind Fp8LinearMethod.apply,
// $VLLM/model_executor/layers/quantization/fp8.py:437-470).
template <int BLOCK_X, int EPI>
__global__ void __launch_bounds__(kGemmThreads, 1)
gemm_fp8_tcgen05_2sm_kernel(const __grid_constant__ CUtensorMap tmap_w,
                            const __grid_constant__ CUtensorMap tmap_x,
                            const __grid_constant__ CUtensorMap tmap_y, int M, int N, int K, int out_f32,
                            int splits, int n_tiles, int m_tiles, int group_m,
                            const float* __restrict__ xscale, const float* __restrict__ wscale,
                            int wscale_stride, DirectOut dir) {
  gemm_2sm_body<BLOCK_X, EPI, 1>(tmap_w, tmap_x, tmap_y, M, N, K, out_f32, splits, n_tiles, m_tiles,
                                 group_m, xscale, wscale, wscale_stride, dir);
}

// Grouped (mixture-of-experts) projection: see GroupInfo above.
template <int BLOCK_X, int EPI>
__global__ void __launch_bounds__(kGemmThreads, 1)
gemm_bf16_grouped_2sm_kernel(const __grid_constant__ CUtensorMap tmap_w,
                             const __grid_constant__ CUtensorMap tmap_x,
                             const __grid_constant__ CUtensorMap tmap_y, int M, int N, int K, int n_tiles,
                             int m_tiles, int group_m, const float* __restrict__ row_scale,
                             const int32_t* __restrict__ tile_expert,
                             const int32_t* __restrict__ num_m_tiles, DirectOut dir) {
  gemm_2sm_body<BLOCK_X, EPI, 0, 1>(tmap_w, tmap_x, tmap_y, M, N, K, 0, 1, n_tiles, m_tiles, group_m,
                                    row_scale, nullptr, 0, dir, GroupInfo{tile_expert, num_m_tiles});
}

// Tensor map of the output for the epilogue's TMA stores: [splits][M][N] (bf16, or fp32 when out_f32)
// or, for the fused SiLU*mul epilogue, the [M][N/2] bf16 activation; box = one warp's chunk.
template <int BLOCK_X, int EPI>
static int make_out_tmap(CUtensorMap* ty, void* Y, int M, int N, int64_t ldy, int out_f32, int splits,
                         DirectOut* dir) {
  constexpr int CH = BLOCK_X >= 32 ? 32 : 16;
  const int es = (EPI == 1 || !out_f32) ? 2 : 4;
  dir->ptr = nullptr;
  dir->ldy = ldy;
  if ((reinterpret_cast<uintptr_t>(Y) & 15) || ((ldy * es) & 15)) {
    // not describable by a tensor map: the kernel stores element by element; the map passed along is
    // a placeholder over the (always aligned) first 16 bytes and is never used
    dir->ptr = Y;
    return make_tmap_out_3d(ty, reinterpret_cast<void*>(reinterpret_cast<uintptr_t>(Y) & ~(uintptr_t)15), es,
                            /*cols*/ 8, /*rows*/ 1, /*planes*/ 1, /*ld*/ 8, /*plane*/ 8, /*box*/ 8, 1);
  }
  if constexpr (EPI == 1)
    return make_tmap_out_3d(ty, Y, 2, (uint64_t)(N / 2), (uint64_t)M, 1, (uint64_t)ldy, 0, 16, CH);
  return make_tmap_out_3d(ty, Y, out_f32 ? 4 : 2, (uint64_t)N, (uint64_t)M, (uint64_t)splits, (uint64_t)ldy,
                          (uint64_t)M * (uint64_t)ldy, 32, CH);
}

static int g_gemm_multicast = -1;  // -1: read B200_GEMM_MULTICAST once (default on)

static bool gemm_multicast_enabled() {
  if (g_gemm_multicast < 0) {
    const char* e = getenv("B200_GEMM_MULTICAST");
    g_gemm_multicast = (e && e[0] == '0') ? 0 : 1;
  }
  return g_gemm_multicast != 0;
}

template <int BLOCK_X, int EPI, int MC>
static int launch_gemm_mc(const CUtensorMap& tw, const CUtensorMap& tx, void* Y, int M, int N, int K,
                          int64_t ldy, int out_f32, int splits, cudaStream_t stream) {
  using Cfg = GemmCfg<BLOCK_X>;
  auto kernel = gemm_bf16_tcgen05_kernel<BLOCK_X, EPI, MC>;
  static bool attr_set = false;
  if (!attr_set) {
    cudaError_t e = cudaFuncSetAttribute(kernel, cudaFuncAttributeMaxDynamicSharedMemorySize, Cfg::kSmemBytes);
    if (e != cudaSuccess) {
      b200_set_error("gemm: cudaFuncSetAttribute(smem=%d): %s", Cfg::kSmemBytes, cudaGetErrorString(e));
      return B200_ERR_CUDA;
    }
    attr_set = true;
  }
  CUtensorMap ty;
  DirectOut dir;
  const int rc = make_out_tmap<BLOCK_X, EPI>(&ty, Y, M, N, ldy, out_f32, splits, &dir);
  if (rc) return rc;
  const int n_tiles = (N + kBlockW - 1) / kBlockW;
  const int m_tiles = (M + BLOCK_X - 1) / BLOCK_X;
  const int n_units = (n_tiles + MC - 1) / MC;
  const int total = n_units * m_tiles * splits;
  const int sms = device_info().sm_count > 0 ? device_info().sm_count : 148;
  const int max_clusters = sms / MC;
  const int clusters = total < max_clusters ? total : max_clusters;
  const int group_m = m_tiles < 8 ? m_tiles : 8;
  cudaError_t e = launch_pdl(kernel, dim3(clusters * MC), dim3(kGemmThreads), (size_t)Cfg::kSmemBytes, stream, MC,
                             tw, tx, ty, M, N, K, out_f32, splits, n_tiles, m_tiles, group_m, dir);
  if (e != cudaSuccess) {
    b200_set_error("b200_gemm_bf16: launch failed: %s", cudaGetErrorString(e));
    (void)cudaGetLastError();
    return B200_ERR_CUDA;
  }
  return b200_check_launch("b200_gemm_bf16");
}

// m-tiles per raster group: a group's activation tiles stay L2-resident while the weight matrix
// streams past once per group (B200_GEMM_GROUP_M overrides for experiments).  16 tiles measured best
// on the K = 4096 prefill shapes (16 x 256 tokens x 8 KB = 32 MB of activations); for long K the
// group is shrunk so that its activation rows stay below ~64 MB -- a 16-tile group of the down
// projection (K = 14336) is 117 MB, does not fit the 126 MB L2 next to the weights, and round 1
// measured 3.1x the algorithmic DRAM traffic for it.
static int gemm_group_m(int m_tiles, int K) {
  static int v = -1;
  if (v < 0) {
    const char* e = getenv("B200_GEMM_GROUP_M");
    v = (e && atoi(e) > 0) ? atoi(e) : 0;
  }
  int g = v;
  if (g <= 0) {
    g = 16;
    const long long tile_bytes = 256LL * (long long)K * 2;   // one m-tile of activations (bf16 element count)
    while (g > 4 && (long long)g * tile_bytes > (64LL << 20)) g >>= 1;
  }
  static int hints = -1;
  if (hints < 0) {
    // "w,x" with 1 = evict-first, 2 = normal, 3 = evict-last, 0 = built-in (weights normal, activations last)
    const char* e = getenv("B200_GEMM_HINTS");
    hints = 0;
    if (e && e[0] >= '0' && e[0] <= '3' && e[1] == ',' && e[2] >= '0' && e[2] <= '3')
      hints = (e[0] - '0') | ((e[2] - '0') << 2);
  }
  return (m_tiles < g ? m_tiles : g) | (hints << 16);
}

static bool gemm_2sm_enabled() {
  static int v = -1;
  if (v < 0) {
    const char* e = getenv("B200_GEMM_2SM");   // default on; B200_GEMM_2SM=0 falls back to 1-CTA tiles
    v = (e && e[0] == '0') ? 0 : 1;
  }
  return v != 0;
}

template <int BLOCK_X, int EPI>
static int launch_gemm_2sm(const void* X, const void* W, void* Y, int M, int N, int K, int64_t ldx,
                           int64_t ldw, int64_t ldy, int out_f32, int splits, cudaStream_t stream) {
  using Cfg = Gemm2Cfg<BLOCK_X>;
  auto kernel = gemm_bf16_tcgen05_2sm_kernel<BLOCK_X, EPI>;
  static bool attr_set = false;
  if (!attr_set) {
    cudaError_t e = cudaFuncSetAttribute(kernel, cudaFuncAttributeMaxDynamicSharedMemorySize, Cfg::kSmemBytes);
    if (e != cudaSuccess) {
      b200_set_error("gemm(2sm): cudaFuncSetAttribute(smem=%d): %s", Cfg::kSmemBytes, cudaGetErrorString(e));
      return B200_ERR_CUDA;
    }
    attr_set = true;
  }
  CUtensorMap tw, tx;
  int rc = make_tmap_bf16_2d(&tw, W, (uint64_t)N, (uint64_t)K, (uint64_t)ldw, kBlockW, kBlockK);
  if (rc) return rc;
  rc = make_tmap_bf16_2d(&tx, X, (uint64_t)M, (uint64_t)K, (uint64_t)ldx, BLOCK_X / 2, kBlockK);
  if (rc) return rc;
  CUtensorMap ty;
  DirectOut dir;
  rc = make_out_tmap<BLOCK_X, EPI>(&ty, Y, M, N, ldy, out_f32, splits, &dir);
  if (rc) return rc;
  const int n_tiles = (N + kBlockW - 1) / kBlockW;
  const int m_tiles = (M + BLOCK_X - 1) / BLOCK_X;
  const int total = ((n_tiles + 1) / 2) * m_tiles * splits;
  const int sms = device_info().sm_count > 0 ? device_info().sm_count : 148;
  const int clusters = total < sms / 2 ? total : sms / 2;
  const int group_m = gemm_group_m(m_tiles, K);
  cudaError_t e = launch_pdl(kernel, dim3(clusters * 2), dim3(kGemmThreads), (size_t)Cfg::kSmemBytes, stream, 2,
                             tw, tx, ty, M, N, K, out_f32, splits, n_tiles, m_tiles, group_m, dir);
  if (e != cudaSuccess) {
    b200_set_error("b200_gemm_bf16(2sm): launch failed: %s", cudaGetErrorString(e));
    (void)cudaGetLastError();
    return B200_ERR_CUDA;
  }
  return b200_check_launch("b200_gemm_bf16(2sm)");
}

template <int BLOCK_X, int EPI>
static int launch_gemm_fp8(const void* X, const void* W, void* Y, int M, int N, int K, int64_t ldx, int64_t ldw,
                           int64_t ldy, int out_f32, int splits, const float* xscale, const float* wscale,
                           int wscale_stride, cudaStream_t stream) {
  using Cfg = Gemm2Cfg<BLOCK_X>;
  auto kernel = gemm_fp8_tcgen05_2sm_kernel<BLOCK_X, EPI>;
  static bool attr_set = false;
  if (!attr_set) {
    cudaError_t e = cudaFuncSetAttribute(kernel, cudaFuncAttributeMaxDynamicSharedMemorySize, Cfg::kSmemBytes);
    if (e != cudaSuccess) {
      b200_set_error("gemm(fp8): cudaFuncSetAttribute(smem=%d): %s", Cfg::kSmemBytes, cudaGetErrorString(e));
      return B200_ERR_CUDA;
    }
    attr_set = true;
  }
  CUtensorMap tw, tx;
  int rc = make_tmap_u8_2d(&tw, W, (uint64_t)N, (uint64_t)K, (uint64_t)ldw, kBlockW, 2 * kBlockK);
  if (rc) return rc;
  rc = make_tmap_u8_2d(&tx, X, (uint64_t)M, (uint64_t)K, (uint64_t)ldx, BLOCK_X / 2, 2 * kBlockK);
  if (rc) return rc;
  CUtensorMap ty;
  DirectOut dir;
  rc = make_out_tmap<BLOCK_X, EPI>(&ty, Y, M, N, ldy, out_f32, splits, &dir);
  if (rc) return rc;
  const int n_tiles = (N + kBlockW - 1) / kBlockW;
  const int m_tiles = (M + BLOCK_X - 1) / BLOCK_X;
  const int total = ((n_tiles + 1) / 2) * m_tiles * splits;
  const int sms = device_info().sm_count > 0 ? device_info().sm_count : 148;
  const int clusters = total < sms / 2 ? total : sms / 2;
  const int group_m = gemm_group_m(m_tiles, K / 2);   // a k-block is 128 one-byte elements
  cudaError_t e = launch_pdl(kernel, dim3(clusters * 2), dim3(kGemmThreads), (size_t)Cfg::kSmemBytes, stream, 2,
                             tw, tx, ty, M, N, K, out_f32, splits, n_tiles, m_tiles, group_m, xscale, wscale,
                             wscale_stride, dir);
  if (e != cudaSuccess) {
    b200_set_error("b200_gemm_fp8: launch failed: %s", cudaGetErrorString(e));
    (void)cudaGetLastError();
    return B200_ERR_CUDA;
  }
  return b200_check_launch("b200_gemm_fp8");
}

template <int BLOCK_X, int EPI = 0>
static int launch_gemm(const void* X, const void* W, void* Y, int M, int N, int K, int64_t ldx,
                       int64_t ldw, int64_t ldy, int out_f32, int splits, cudaStream_t stream) {
  const int n_tiles = (N + kBlockW - 1) / kBlockW;
  if constexpr (BLOCK_X >= 32) {
    if (gemm_2sm_enabled() && n_tiles >= 2)
      return launch_gemm_2sm<BLOCK_X, EPI>(X, W, Y, M, N, K, ldx, ldw, ldy, out_f32, splits, stream);
  }
  // pairs of weight tiles share one multicast activation tile; not worth it for a single tile
  const bool mc = gemm_multicast_enabled() && n_tiles >= 2;
  CUtensorMap tw, tx;
  int rc = make_tmap_bf16_2d(&tw, W, (uint64_t)N, (uint64_t)K, (uint64_t)ldw, kBlockW, kBlockK);
  if (rc) return rc;
  rc = make_tmap_bf16_2d(&tx, X, (uint64_t)M, (uint64_t)K, (uint64_t)ldx, mc ? BLOCK_X / 2 : BLOCK_X, kBlockK);
  if (rc) return rc;
  if (mc) return launch_gemm_mc<BLOCK_X, EPI, 2>(tw, tx, Y, M, N, K, ldy, out_f32, splits, stream);
  return launch_gemm_mc<BLOCK_X, EPI, 1>(tw, tx, Y, M, N, K, ldy, out_f32, splits, stream);
}

}  // namespace b200

using namespace b200;

// ---- decode-shape planning (M <= 256), from the sweep in tools/gemm_sweep.py ------------------
// A tile's cost is dominated by the bytes its SM has to ingest (weight tile + activation tile, the
// latter from L2), so the plan spreads one GEMM over ~one wave of CTAs (units = n_tiles * m_tiles
// * splits close to the SM count) and prefers splitting K (no re-read) over splitting M (weights
// re-read from L2) when K is long.
static int pow2_cover(int M) {
  int b = 16;
  while (b < M && b < 256) b <<= 1;
  return b;
}

// Activation-tile width for a given split count: start from the smallest width covering M and
// cut it into 2 or 4 m-tiles while that still fits one wave of CTAs (few weight tiles = small N,
// e.g. tensor-parallel shards: the weights are re-read from L2, not from HBM).
static int plan_block_x(int M, int N, int splits) {
  if (M > 256) return 256;
  const int sms = device_info().sm_count > 0 ? device_info().sm_count : 148;
  const int base = pow2_cover(M);
  const int units = ((N + kBlockW - 1) / kBlockW) * splits;
  const int budget = sms + sms / 12;
  if (base >= 128 && units * 4 <= budget) return base / 4;
  if (base >= 64 && units * 2 <= budget) return base / 2;
  return base;
}

extern "C" int b200_gemm_pick_block_x(int M, int N, int K) {
  (void)K;
  return plan_block_x(M, N, 1);
}

// K-splits for the row-parallel projections (O, down), whose fp32 partials are consumed for free
// by b200_fused_add_rms_norm_partials.
extern "C" int b200_gemm_suggest_splits(int M, int N, int K) {
  if (M <= 0 || N <= 0 || K <= 0 || K % kBlockK) return 1;
  if (M > 256) return 1;  // prefill: tensor-bound, plenty of tiles
  const int sms = device_info().sm_count > 0 ? device_info().sm_count : 148;
  const int n_tiles = (N + kBlockW - 1) / kBlockW;
  const int kb = K / kBlockK;
  int f = sms / n_tiles;             // how many units per weight tile one wave can hold
  if (f < 2) return 1;
  if (K < 8192 && M > 32) f /= 2;    // short K: let two m-tiles take half of the factor
  int splits = 1;
  while (splits * 2 <= f && splits < 8 && kb % (splits * 2) == 0 && kb / (splits * 2) >= 8)
    splits *= 2;
  return splits;
}

extern "C" int b200_gemm_bf16_ex(const void* X, const void* W, void* Y, int M, int N, int K,
                                 int64_t ldx, int64_t ldw, int64_t ldy, int out_f32, int splits,
                                 int block_x, void* stream) {
  B200_REQUIRE(M >= 0 && N > 0 && K > 0, "gemm: bad shape M=%d N=%d K=%d", M, N, K);
  B200_REQUIRE(K % kBlockK == 0, "gemm: K=%d must be a multiple of %d", K, kBlockK);
  B200_REQUIRE(splits >= 1 && (K / kBlockK) % splits == 0,
               "gemm: splits=%d must divide K/%d=%d", splits, kBlockK, K / kBlockK);
  B200_REQUIRE(splits == 1 || out_f32, "gemm: split-K output must be fp32 partials");
  B200_REQUIRE(ldx >= K && ldw >= K && ldy >= N, "gemm: leading dimensions too small");
  if (M == 0) return B200_OK;
  const DeviceInfo& d = device_info();
  if (!d.valid || d.cc_major != 10) {
    b200_set_error("gemm: tcgen05 kernel needs an sm_100 device (found cc %d.%d)", d.cc_major,
                   d.cc_minor);
    return B200_ERR_UNSUPPORTED;
  }
  cudaStream_t s = (cudaStream_t)stream;
  if (block_x == 0) block_x = plan_block_x(M, N, splits);
  switch (block_x) {
    case 16: return launch_gemm<16>(X, W, Y, M, N, K, ldx, ldw, ldy, out_f32, splits, s);
    case 32: return launch_gemm<32>(X, W, Y, M, N, K, ldx, ldw, ldy, out_f32, splits, s);
    case 64: return launch_gemm<64>(X, W, Y, M, N, K, ldx, ldw, ldy, out_f32, splits, s);
    case 128: return launch_gemm<128>(X, W, Y, M, N, K, ldx, ldw, ldy, out_f32, splits, s);
    case 256: return launch_gemm<256>(X, W, Y, M, N, K, ldx, ldw, ldy, out_f32, splits, s);
    default:
      b200_set_error("gemm: block_x=%d unsupported (16,32,64,128,256)", block_x);
      return B200_ERR_INVALID;
  }
}

extern "C" int b200_gemm_bf16_silu_mul(const void* X, const void* W, void* Y, int M, int N, int K,
                                       int64_t ldx, int64_t ldw, int64_t ldy, void* stream) {
  B200_REQUIRE(M >= 0 && N > 0 && K > 0 && N % 2 == 0, "gemm_silu_mul: bad shape M=%d N=%d K=%d", M, N, K);
  B200_REQUIRE(K % kBlockK == 0, "gemm_silu_mul: K=%d must be a multiple of %d", K, kBlockK);
  B200_REQUIRE(ldx >= K && ldw >= K && ldy >= N / 2, "gemm_silu_mul: leading dimensions too small");
  if (M == 0) return B200_OK;
  const DeviceInfo& d = device_info();
  if (!d.valid || d.cc_major != 10) {
    b200_set_error("gemm: tcgen05 kernel needs an sm_100 device (found cc %d.%d)", d.cc_major, d.cc_minor);
    return B200_ERR_UNSUPPORTED;
  }
  cudaStream_t s = (cudaStream_t)stream;
  switch (plan_block_x(M, N, 1)) {
    case 16: return launch_gemm<16, 1>(X, W, Y, M, N, K, ldx, ldw, ldy, 0, 1, s);
    case 32: return launch_gemm<32, 1>(X, W, Y, M, N, K, ldx, ldw, ldy, 0, 1, s);
    case 64: return launch_gemm<64, 1>(X, W, Y, M, N, K, ldx, ldw, ldy, 0, 1, s);
    case 128: return launch_gemm<128, 1>(X, W, Y, M, N, K, ldx, ldw, ldy, 0, 1, s);
    default: return launch_gemm<256, 1>(X, W, Y, M, N, K, ldx, ldw, ldy, 0, 1, s);
  }
}

extern "C" int b200_gemm_bf16(const void* X, const void* W, void* Y, int M, int N, int K,
                              int64_t ldx, int64_t ldw, int64_t ldy, int out_f32, int splits,
                              void* stream) {
  return b200_gemm_bf16_ex(X, W, Y, M, N, K, ldx, ldw, ldy, out_f32, splits, 0, stream);
}

// W8A8 fp8 (e4m3) projection.  X [M,K] and W [N,K] are e4m3 bytes, xscale [M] the dynamic per-token
// activation scales, wscale the weight scale(s): stride 0 = per-tensor, 1 = per output channel.
// silu_mul != 0 applies the fused SiLU*mul epilogue over interleaved gate/up rows (Y is [M, N/2]).
extern "C" int b200_gemm_fp8(const void* X, const void* W, void* Y, int M, int N, int K, int64_t ldx,
                             int64_t ldw, int64_t ldy, const float* xscale, const float* wscale,
                             int wscale_stride, int out_f32, int splits, int silu_mul, void* stream) {
  constexpr int kElems = 2 * kBlockK;
  B200_REQUIRE(M >= 0 && N > 0 && K > 0, "gemm_fp8: bad shape M=%d N=%d K=%d", M, N, K);
  B200_REQUIRE(K % kElems == 0, "gemm_fp8: K=%d must be a multiple of %d", K, kElems);
  B200_REQUIRE(N > kBlockW, "gemm_fp8: N=%d must exceed %d (CTA-pair tiles)", N, kBlockW);
  B200_REQUIRE(splits >= 1 && (K / kElems) % splits == 0, "gemm_fp8: splits=%d must divide K/%d=%d", splits,
               kElems, K / kElems);
  B200_REQUIRE(splits == 1 || out_f32, "gemm_fp8: split-K output must be fp32 partials");
  B200_REQUIRE(!silu_mul || (splits == 1 && !out_f32 && N % 2 == 0), "gemm_fp8: silu_mul needs bf16 output, no split-K");
  B200_REQUIRE(ldx >= K && ldw >= K && ldy >= (silu_mul ? N / 2 : N), "gemm_fp8: leading dimensions too small");
  B200_REQUIRE(xscale && wscale && (wscale_stride == 0 || wscale_stride == 1), "gemm_fp8: bad scale arguments");
  if (M == 0) return B200_OK;
  const DeviceInfo& d = device_info();
  if (!d.valid || d.cc_major != 10) {
    b200_set_error("gemm_fp8: tcgen05 kernel needs an sm_100 device (found cc %d.%d)", d.cc_major, d.cc_minor);
    return B200_ERR_UNSUPPORTED;
  }
  cudaStream_t s = (cudaStream_t)stream;
  int bx = plan_block_x(M, N, splits);
  if (bx < 32) bx = 32;
#define B200_FP8_CASE(BX)                                                                                   \
  case BX:                                                                                                  \
    return silu_mul ? launch_gemm_fp8<BX, 1>(X, W, Y, M, N, K, ldx, ldw, ldy, 0, 1, xscale, wscale,         \
                                             wscale_stride, s)                                              \
                    : launch_gemm_fp8<BX, 0>(X, W, Y, M, N, K, ldx, ldw, ldy, out_f32, splits, xscale,      \
                                             wscale, wscale_stride, s);
  switch (bx) {
    B200_FP8_CASE(32)
    B200_FP8_CASE(64)
    B200_FP8_CASE(128)
    default:
      B200_FP8_CASE(256)
  }
#undef B200_FP8_CASE
}

namespace b200 {
template <int BLOCK_X, int EPI>
static int launch_gemm_grouped(const void* X, const void* W, void* Y, int M, int N, int K, int E, int64_t ldx,
                               int64_t ldw, int64_t ldy, const float* row_scale, const int32_t* tile_expert,
                               const int32_t* num_m_tiles, cudaStream_t stream) {
  using Cfg = Gemm2Cfg<BLOCK_X>;
  auto kernel = gemm_bf16_grouped_2sm_kernel<BLOCK_X, EPI>;
  static bool attr_set = false;
  if (!attr_set) {
    cudaError_t e = cudaFuncSetAttribute(kernel, cudaFuncAttributeMaxDynamicSharedMemorySize, Cfg::kSmemBytes);
    if (e != cudaSuccess) {
      b200_set_error("gemm(grouped): cudaFuncSetAttribute(smem=%d): %s", Cfg::kSmemBytes, cudaGetErrorString(e));
      return B200_ERR_CUDA;
    }
    attr_set = true;
  }
  CUtensorMap tw, tx;
  int rc = make_tmap_bf16_2d(&tw, W, (uint64_t)E * (uint64_t)N, (uint64_t)K, (uint64_t)ldw, kBlockW, kBlockK);
  if (rc) return rc;
  rc = make_tmap_bf16_2d(&tx, X, (uint64_t)M, (uint64_t)K, (uint64_t)ldx, BLOCK_X / 2, kBlockK);
  if (rc) return rc;
  CUtensorMap ty;
  DirectOut dir;
  rc = make_out_tmap<BLOCK_X, EPI>(&ty, Y, M, N, ldy, 0, 1, &dir);
  if (rc) return rc;
  const int n_tiles = (N + kBlockW - 1) / kBlockW;
  const int m_tiles = (M + BLOCK_X - 1) / BLOCK_X;      // capacity; the kernel clamps to the tiles in use
  const int total = ((n_tiles + 1) / 2) * m_tiles;
  const int sms = device_info().sm_count > 0 ? device_info().sm_count : 148;
  const int clusters = total < sms / 2 ? total : sms / 2;
  const int group_m = gemm_group_m(m_tiles, K);
  cudaError_t e = launch_pdl_cls(kPdlMoe, kernel, dim3(clusters * 2), dim3(kGemmThreads), (size_t)Cfg::kSmemBytes,
                                 stream, 2, tw, tx, ty, M, N, K, n_tiles, m_tiles, group_m, row_scale,
                                 tile_expert, num_m_tiles, dir);
  if (e != cudaSuccess) {
    b200_set_error("b200_moe_grouped_gemm: launch failed: %s", cudaGetErrorString(e));
    (void)cudaGetLastError();
    return B200_ERR_CUDA;
  }
  return b200_check_launch("b200_moe_grouped_gemm");
}
}  // namespace b200

// Grouped expert projection over the permuted row space of b200_moe_align (block_x must be the
// one the rows were aligned to).  X [capacity_rows, K] bf16, W [E*N, K] bf16 (experts stacked),
// Y [capacity_rows, N] bf16 (or [capacity_rows, N/2] with silu_mul over interleaved gate/up rows);
// row_scale (optional, fp32 [capacity_rows]) multiplies the fp32 accumulator of each row.
extern "C" int b200_moe_grouped_gemm(const void* X, const void* W, void* Y, int capacity_rows, int N, int K,
                                     int num_experts, int64_t ldx, int64_t ldw, int64_t ldy,
                                     const float* row_scale, const int32_t* tile_expert,
                                     const int32_t* num_tiles, int block_x, int silu_mul, void* stream) {
  B200_REQUIRE(capacity_rows > 0 && N > 0 && K > 0 && num_experts >= 1, "moe_grouped_gemm: bad shape");
  B200_REQUIRE(K % kBlockK == 0, "moe_grouped_gemm: K=%d must be a multiple of %d", K, kBlockK);
  B200_REQUIRE(N % (2 * kBlockW) == 0, "moe_grouped_gemm: N=%d per expert must be a multiple of %d", N, 2 * kBlockW);
  B200_REQUIRE(block_x > 0 && capacity_rows % block_x == 0, "moe_grouped_gemm: capacity_rows must be a multiple of block_x");
  B200_REQUIRE(!silu_mul || row_scale == nullptr, "moe_grouped_gemm: silu_mul and row_scale are exclusive");
  B200_REQUIRE(ldx >= K && ldw >= K && ldy >= (silu_mul ? N / 2 : N), "moe_grouped_gemm: leading dimensions too small");
  B200_REQUIRE(tile_expert && num_tiles, "moe_grouped_gemm: null tile tables");
  const DeviceInfo& d = device_info();
  if (!d.valid || d.cc_major != 10) {
    b200_set_error("moe_grouped_gemm: tcgen05 kernel needs an sm_100 device (found cc %d.%d)", d.cc_major, d.cc_minor);
    return B200_ERR_UNSUPPORTED;
  }
  cudaStream_t s = (cudaStream_t)stream;
#define B200_GRP_CASE(BX)                                                                                     \
  case BX:                                                                                                    \
    return silu_mul ? launch_gemm_grouped<BX, 1>(X, W, Y, capacity_rows, N, K, num_experts, ldx, ldw, ldy,    \
                                                 nullptr, tile_expert, num_tiles, s)                          \
                    : launch_gemm_grouped<BX, 0>(X, W, Y, capacity_rows, N, K, num_experts, ldx, ldw, ldy,    \
                                                 row_scale, tile_expert, num_tiles, s);
  switch (block_x) {
    B200_GRP_CASE(32)
    B200_GRP_CASE(64)
    B200_GRP_CASE(128)
    B200_GRP_CASE(256)
    default:
      b200_set_error("moe_grouped_gemm: block_x=%d unsupported (32,64,128,256)", block_x);
      return B200_ERR_INVALID;
  }
#undef B200_GRP_CASE
}
