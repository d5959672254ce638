// Causal paged-KV prefill attention on the 5th-generation tensor cores (reference operator:
// FlashInferImpl.forward -> trtllm_batch_context_with_kv_cache,
// $VLLM/v1/attention/backends/flashinfer.py:1095-1230; on cc 10.x the reference engine runs
// FlashInfer's TRT-LLM-gen FMHA cubins for this step).
//
// One work item = one 128-row query tile of one (sequence, kv-head): the G = Hq/Hkv query heads that
// share the kv-head times 128/G consecutive query tokens, so that S = Q K^T is a single UMMA M = 128
// tile and every K/V byte staged in shared memory serves the whole GQA group.
//
//   warp 4    TMA producer: the Q tile once per work item, then K_j / V_j blocks of 64 tokens (4 KV pages of 16
//             tokens, each page two 2 KB boxes of the 128-byte-swizzled layout) through a 4-slot
//             mbarrier ring; pages past the end of the sequence are read from the all-zero null block;
//   warp 5    MMA issuer (one elected thread): S_j = Q K_j^T into one of two 64-column TMEM buffers,
//             O += P_j V_j into a 128-column TMEM accumulator.  S_{j+1} is issued BEFORE waiting for
//             P_j, so the tensor pipe works on block j+1 while the softmax warps are busy with j;
//   warps 0-3 softmax: thread = one query row (TMEM lane).  tcgen05.ld S, scale / causal mask,
//             online softmax in base 2 with LAZY rescaling (the running reference maximum is only
//             moved when the row maximum grows by more than 2^8, so the O accumulator in TMEM is
//             rescaled rarely; P values stay <= 256, exact in fp32 / fine in bf16), P rounded to bf16
//             and written back over the S buffer's tensor-memory columns, from where the PV MMA reads
//             it as its A operand (P_TMEM; the shared-memory route is kept for A/B runs); at the end
//             O / l -> bf16 -> global.  The hand-off to the MMA warp is one mbarrier PER S BUFFER.
//
// V is consumed as an MN-major B operand straight from its [token][d] page layout (no transpose).
// Two CTAs share an SM (112 KB of shared memory, 256 TMEM columns each): one CTA's softmax / epilogue
// overlaps the other's MMAs.  Work per launch: 4 * T * ctx/2 * Hq * D FLOP (tensor-bound).
//
// PERSISTENT tile loop: the grid is 2 CTAs per SM and CTA c walks the (heaviest-first) list of
// (tile, kv-head) work items in a boustrophedon static schedule (see work_of below).  Barriers, tensor memory and the K/V ring
// live for the whole kernel; all mbarrier phases are running counters.  While a tile finishes (last
// softmax block, PV, epilogue) the producer is already streaming the next tile's K/V into the ring
// and -- once the last QK^T of the tile has retired (q_empty) -- its Q; every role loads the NEXT
// item's metadata while it works on the current one.  The per-tile cost of a fresh CTA (barrier
// init, TMEM allocation, three dependent metadata loads, cold Q/K latency, teardown) is paid once
// per CTA instead of once per tile.  B200_ATTN_PERSIST=0 launches one CTA per work item instead.
#include "common.cuh"
#include "runtime.h"
#include "../../include/b200.h"

#include <math.h>
#include <stdlib.h>

namespace b200 {

constexpr int kFaThreads = 192;
constexpr int kFaD = 128;
constexpr int kUmmaKBf16 = 16;                         // K elements per tcgen05.mma (kind::f16)
constexpr int kFaKv = 64;                              // KV tokens per ring slot
constexpr int kFaPage = 16;                            // tokens per KV page (vLLM block_size)
constexpr int kFaSlotBytes = kFaKv * kFaD * 2;         // 16 KB = 2 d-halves x [64 tokens][128 B]
constexpr int kFaHalfBytes = kFaKv * 128;              // 8 KB
constexpr int kFaSlots = 4;                            // K_j, V_j, K_j+1, V_j+1
constexpr int kFaQBytes = 128 * kFaD * 2;              // 32 KB = 2 d-halves x [128 rows][128 B]
constexpr int kFaQHalfBytes = 128 * 128;               // 16 KB
constexpr int kFaPBytes = 128 * kFaKv * 2;             // 16 KB = [128 rows][128 B]
constexpr int kFaOffKV = kFaQBytes;
constexpr int kFaOffP = kFaOffKV + kFaSlots * kFaSlotBytes;
constexpr int kFaOffBar = kFaOffP + kFaPBytes;         // 114,688
constexpr int kFaSmemBytes = 115712;                   // (228 KB - 2 x 1 KB reserved) / 2 CTAs per SM
constexpr int kFaMaxSlack = kFaSmemBytes - kFaOffBar - 256;
constexpr int kFaTmemCols = 256;                       // S0 [0,64)  S1 [64,128)  O [128,256)
constexpr float kFaRescaleThreshold = 8.f;             // log2 units

// P_TMEM: P_j is written back into the first 32 columns of the TMEM buffer S_j came from (bf16 pairs
// per 32-bit column, tcgen05.st) and the PV MMA reads its A operand from tensor memory: no shared-
// memory round trip, no proxy fence, and the softmax of block j+1 no longer waits for PV_j to retire
// (it only ever waits for p_empty when the O accumulator has to be rescaled).  P_TMEM = false keeps
// P in swizzled shared memory (B200_ATTN_P_TMEM=0, A/B runs).
template <int G, bool P_TMEM>
__global__ void __launch_bounds__(kFaThreads, 2)
paged_attn_prefill_tc_kernel(const __grid_constant__ CUtensorMap tmap_q,
                             const __grid_constant__ CUtensorMap tmap_k,
                             const __grid_constant__ CUtensorMap tmap_v,
                             const int32_t* __restrict__ block_table, int max_blocks,
                             const int32_t* __restrict__ cu_q, const int32_t* __restrict__ ctx_lens,
                             const int32_t* __restrict__ tile_seq, const int32_t* __restrict__ tile_q0,
                             int num_tiles, bf16* __restrict__ out, int Hq, int Hkv, float scale_log2) {
  constexpr int TPT = 128 / G;   // query tokens per tile
  extern __shared__ uint8_t fa_raw[];
  uint8_t* smem = reinterpret_cast<uint8_t*>((reinterpret_cast<uintptr_t>(fa_raw) + 1023) &
                                             ~static_cast<uintptr_t>(1023));
  if (smem - fa_raw > kFaMaxSlack) __trap();   // the swizzled tiles need 1024-byte alignment
  uint64_t* bars = reinterpret_cast<uint64_t*>(smem + kFaOffBar);
  uint64_t* q_full = bars;
  uint64_t* kv_full = bars + 1;     // [4]
  uint64_t* kv_empty = bars + 5;    // [4]
  uint64_t* s_full = bars + 9;      // [2]
  uint64_t* p_full = bars + 11;
  uint64_t* p_empty = bars + 12;
  uint64_t* o_full = bars + 13;
  uint64_t* q_empty = bars + 14;    // the tile's last Q K^T has retired: the Q buffer may be reloaded
  uint64_t* o_empty = bars + 15;    // the epilogue has read O out of tensor memory (4 warp arrivals)
  uint64_t* p_full_b = bars + 16;   // p_full of the odd S buffer (p_full itself: the even one)
  uint32_t* tmem_slot = reinterpret_cast<uint32_t*>(bars + 17);

  const int warp = threadIdx.x >> 5;
  const int lane = threadIdx.x & 31;
  pdl_launch_dependents();

  if (threadIdx.x == 0) {
    tma_prefetch_desc(&tmap_q);
    tma_prefetch_desc(&tmap_k);
    tma_prefetch_desc(&tmap_v);
    mbar_init(q_full, 1);
    for (int i = 0; i < kFaSlots; ++i) {
      mbar_init(&kv_full[i], 1);
      mbar_init(&kv_empty[i], 1);
    }
    mbar_init(&s_full[0], 1);
    mbar_init(&s_full[1], 1);
    mbar_init(p_full, 4);     // one arrive per softmax warp; ONE BARRIER PER S BUFFER (see the softmax loop)
    mbar_init(p_full_b, 4);
    mbar_init(p_empty, 1);
    mbar_init(o_full, 1);
    mbar_init(q_empty, 1);
    mbar_init(o_empty, 4);
    fence_mbar_init();
  }
  if (warp == 5) {
    tmem_alloc(tmem_slot, kFaTmemCols);
    tmem_relinquish();
  }
  tc_fence_before();
  __syncthreads();
  tc_fence_after();
  const uint32_t tmem_base = *tmem_slot;
  pdl_wait();   // Q, the KV cache and the index buffers were written by preceding kernels / copies

  // one work item = (tile, kv-head); tiles are listed heaviest first, the kv-heads of a tile are adjacent
  struct Work {
    int seq, kvh, q0, q_begin, q_len, ctx, total_len, n_blocks, n_pages_seq;
  };
  const int total_work = num_tiles * Hkv;
  auto load_work = [&](int w) {
    Work m;
    if (w >= total_work) w = total_work - 1;               // prefetch past the end: harmless re-read
    const int tile = w / Hkv;
    m.kvh = w - tile * Hkv;
    m.seq = __ldg(tile_seq + tile);
    m.q0 = __ldg(tile_q0 + tile);
    m.q_begin = __ldg(cu_q + m.seq);
    m.q_len = __ldg(cu_q + m.seq + 1) - m.q_begin;
    m.ctx = __ldg(ctx_lens + m.seq);
    m.total_len = m.ctx + m.q_len;
    const int kv_end = min(m.total_len, m.ctx + m.q0 + TPT);   // causal horizon of the tile (exclusive)
    m.n_blocks = (kv_end + kFaKv - 1) / kFaKv;                 // >= 1
    m.n_pages_seq = (m.total_len + kFaPage - 1) / kFaPage;
    return m;
  };
  // Static schedule over the heaviest-first list: round r hands work r * P + c to CTA c on even rounds and
  // r * P + (P - 1 - c) on odd ones (boustrophedon), so a CTA that got the heaviest item of one round
  // gets the lightest of the next -- per-CTA sums stay level although causal tiles differ by up to
  // ctx / 64 blocks.  (Plain c, c + P, c + 2P ... left a 29 % spread on one 8192-token prompt.)
  const int n_cta = gridDim.x;
  const int n_rounds = (total_work + n_cta - 1) / n_cta;
  auto work_of = [&](int rd) { return rd * n_cta + ((rd & 1) ? (n_cta - 1 - (int)blockIdx.x) : (int)blockIdx.x); };

  if (warp == 4) {
    // ------------------------------------------------------------------ TMA producer
    // The whole warp walks the block table (lane l holds the page id of page base + l, one coalesced
    // load per 32 pages = 8 KV blocks; a dependent global load per TMA would serialise the ring at
    // L2 latency); lane 0 waits for the slot and issues the copies.
    constexpr int PPB = kFaKv / kFaPage;   // pages per KV block
    int slot = 0;
    uint32_t phase = 0, it = 0;
    Work nxt = load_work(work_of(0));
    for (int rd = 0; rd < n_rounds; ++rd) {
      const Work m = nxt;
      nxt = load_work(work_of(rd + 1));     // in flight while this tile streams
      if (work_of(rd) >= total_work) break; // only in the last, partial round
      if (lane == 0) {
        if (it > 0) mbar_wait_guarded(q_empty, (it - 1) & 1);   // previous tile's Q K^T all retired
        mbar_expect_tx(q_full, kFaQBytes);
#pragma unroll
        for (int g = 0; g < G; ++g) {
#pragma unroll
          for (int h = 0; h < 2; ++h)
            tma_load_2d(smem + h * kFaQHalfBytes + g * TPT * 128, &tmap_q, q_full, (m.kvh * G + g) * kFaD + h * 64,
                        m.q_begin + m.q0, kEvictFirst);
        }
      }
      const int32_t* bt = block_table + (size_t)m.seq * max_blocks;
      int my_blk = 0;
      for (int j = 0; j < m.n_blocks; ++j) {
        if ((j & 7) == 0) {   // refill: pages [j*4, j*4 + 32)
          const int page = j * PPB + lane;
          my_blk = page < m.n_pages_seq ? __ldg(bt + page) : 0;   // block 0: the all-zero null block
        }
        int rows[PPB];
#pragma unroll
        for (int p = 0; p < PPB; ++p)
          rows[p] = (__shfl_sync(0xffffffffu, my_blk, ((j & 7) * PPB + p)) * Hkv + m.kvh) * kFaPage;
        if (lane == 0) {
#pragma unroll
          for (int kv = 0; kv < 2; ++kv) {   // K_j then V_j
            mbar_wait_guarded(&kv_empty[slot], phase ^ 1);
            mbar_expect_tx(&kv_full[slot], kFaSlotBytes);
            uint8_t* dst = smem + kFaOffKV + slot * kFaSlotBytes;
#pragma unroll
            for (int p = 0; p < PPB; ++p) {
#pragma unroll
              for (int h = 0; h < 2; ++h)
                tma_load_2d(dst + h * kFaHalfBytes + p * (kFaPage * 128), kv ? &tmap_v : &tmap_k, &kv_full[slot],
                            h * 64, rows[p], kEvictNormal);
            }
            if (++slot == kFaSlots) { slot = 0; phase ^= 1; }
          }
        }
        __syncwarp();
      }
      ++it;
    }
  } else if (warp == 5) {
    // ------------------------------------------------------------------ MMA issuer
    if (elect_one()) {
      constexpr uint32_t idesc_s = umma_idesc_bf16(128, kFaKv);        // S[128 x 64]  = Q K^T
      constexpr uint32_t idesc_o = umma_idesc_bf16_bmn(128, kFaD);     // O[128 x 128] += P V (V MN-major)
      const uint32_t q_addr = smem_u32(smem);
      const uint32_t kv_addr = smem_u32(smem + kFaOffKV);
      const uint32_t p_addr = smem_u32(smem + kFaOffP);
      // gb: running KV-block counter of this CTA (across tiles).  Block gb uses S buffer gb & 1, ring
      // slots 2gb (K) and 2gb + 1 (V), and the gb-th phase of p_full / p_empty.
      auto issue_s = [&](uint32_t gb) {
        const uint32_t slot = (2 * gb) & 3;
        mbar_wait_guarded(&kv_full[slot], ((2 * gb) >> 2) & 1);
        tc_fence_after();
        const uint32_t k_addr = kv_addr + slot * kFaSlotBytes;
        const uint32_t d = tmem_base + (gb & 1) * kFaKv;
#pragma unroll
        for (int k = 0; k < kFaD / kUmmaKBf16; ++k) {   // 8 MMAs x 16 head-dim elements
          const uint32_t half = k >> 2, off = (k & 3) * 32;
          umma_bf16(d, umma_desc_k_sw128(q_addr + half * kFaQHalfBytes + off),
                    umma_desc_k_sw128(k_addr + half * kFaHalfBytes + off), idesc_s, k != 0);
        }
        umma_commit(&kv_empty[slot]);     // K slot reusable once these MMAs retire
        umma_commit(&s_full[gb & 1]);     // S ready for the softmax warps
      };
      uint32_t gb = 0, it = 0;
      Work nxt = load_work(work_of(0));
      for (int rd = 0; rd < n_rounds; ++rd) {
        const int n_blocks = nxt.n_blocks;
        nxt = load_work(work_of(rd + 1));
        if (work_of(rd) >= total_work) break;
        mbar_wait_guarded(q_full, it & 1);
        tc_fence_after();
        issue_s(gb);
        if (n_blocks == 1) umma_commit(q_empty);
        for (int j = 0; j < n_blocks; ++j, ++gb) {
          // S buffer (gb+1)&1 held S_{gb-1}, which the softmax warps finished reading before p_full(gb-1)
          if (j + 1 < n_blocks) {
            issue_s(gb + 1);
            if (j + 2 == n_blocks) umma_commit(q_empty);   // the tile's last Q K^T: Q may be reloaded
          }
          mbar_wait_guarded((gb & 1) ? p_full_b : p_full, (gb >> 1) & 1);
          tc_fence_after();
          if (j == 0 && it > 0) {           // PV_0 overwrites O: the previous tile's epilogue must have read it
            mbar_wait_guarded(o_empty, (it - 1) & 1);
            tc_fence_after();
          }
          const uint32_t slot = (2 * gb + 1) & 3;
          mbar_wait_guarded(&kv_full[slot], ((2 * gb + 1) >> 2) & 1);
          tc_fence_after();
          const uint32_t v_addr = kv_addr + slot * kFaSlotBytes;
#pragma unroll
          for (int k = 0; k < kFaKv / kUmmaKBf16; ++k) {   // 4 MMAs x 16 KV tokens
            const uint64_t v_desc = umma_desc_mn_sw128(v_addr + k * (kUmmaKBf16 * 128), kFaHalfBytes, 1024);
            if constexpr (P_TMEM)   // A = P in TMEM: 16 bf16 of K = 8 columns per MMA
              umma_bf16_ts(tmem_base + 2 * kFaKv, tmem_base + (gb & 1) * kFaKv + k * 8, v_desc, idesc_o,
                           (j | k) != 0);
            else
              umma_bf16(tmem_base + 2 * kFaKv, umma_desc_k_sw128(p_addr + k * 32), v_desc, idesc_o, (j | k) != 0);
          }
          umma_commit(&kv_empty[slot]);     // V slot reusable
          umma_commit(p_empty);             // P buffer and the O accumulator are stable again
        }
        umma_commit(o_full);
        ++it;
      }
    }
  } else {
    // ------------------------------------------------------------------ softmax / epilogue (warps 0..3)
    const int row = warp * 32 + lane;        // TMEM lane = tile row = (head g, token tok)
    const int g = row / TPT, tok = row % TPT;
    const uint32_t t_lane = tmem_base + ((uint32_t)(warp * 32) << 16);
    const uint32_t p_row = smem_u32(smem + kFaOffP) + (uint32_t)row * 128u;
    const uint32_t sw = (uint32_t)(row & 7);
    const bool wide = (reinterpret_cast<uintptr_t>(out) & 31u) == 0;   // 256-bit stores: one full sector per thread
    uint32_t gb = 0, it = 0;
    Work nxt = load_work(work_of(0));
    for (int rd = 0; rd < n_rounds; ++rd) {
      const Work m = nxt;
      nxt = load_work(work_of(rd + 1));
      if (work_of(rd) >= total_work) break;
      const int qi = m.q0 + tok;
      const bool valid = qi < m.q_len;
      const int qpos = m.ctx + qi;
      const int total_len = m.total_len;
      float m_used = -INFINITY, l = 0.f;

      for (int j = 0; j < m.n_blocks; ++j, ++gb) {
        const uint32_t s_col = (gb & 1) * kFaKv;
        mbar_wait_guarded(&s_full[gb & 1], (gb >> 1) & 1);
        tc_fence_after();
        uint32_t s0[32], s1[32];
        tmem_ld_32x32(t_lane + s_col, s0);
        tmem_ld_32x32(t_lane + s_col + 32, s1);
        tmem_ld_wait();
        const int kv0 = j * kFaKv;
        // block-uniform: does any row of the tile see a masked column in this block?
        const bool need_mask = (kv0 + kFaKv - 1 > m.ctx + m.q0) || (kv0 + kFaKv - 1 >= total_len);
        if (need_mask) {   // ONE warp-uniform branch per block (the condition is tile-uniform), not one per column
#pragma unroll
          for (int c = 0; c < 32; ++c) {
            const int ka = kv0 + c, kb = kv0 + 32 + c;
            if (ka > qpos || ka >= total_len) s0[c] = 0xff800000u;   // -inf
            if (kb > qpos || kb >= total_len) s1[c] = 0xff800000u;
          }
        }
        // row maximum on the raw scores (scale > 0 commutes with max); the scale itself is folded
        // into the exponent's FFMA below
        float mx = -INFINITY;
#pragma unroll
        for (int c = 0; c < 32; ++c) mx = fmaxf(mx, fmaxf(__uint_as_float(s0[c]), __uint_as_float(s1[c])));
        mx *= scale_log2;
        // lazy rescale: move the reference maximum only when the row maximum outgrew it by > 2^8
        // (always on the first block, where m_used = -inf and the factor is exp2(-inf) = 0)
        float alpha = 1.f;
        bool rescale = false;
        if (mx > m_used + kFaRescaleThreshold) {
          alpha = ex2_approx(m_used - mx);
          m_used = mx;
          l *= alpha;
          rescale = j > 0;
        }
        uint32_t pk[32];
        float sum = 0.f;
        const float neg_m = -m_used;
#pragma unroll
        for (int c = 0; c < 32; c += 2) {
          const float p0 = ex2_approx(fmaf(__uint_as_float(s0[c]), scale_log2, neg_m));
          const float p1 = ex2_approx(fmaf(__uint_as_float(s0[c + 1]), scale_log2, neg_m));
          const float p2 = ex2_approx(fmaf(__uint_as_float(s1[c]), scale_log2, neg_m));
          const float p3 = ex2_approx(fmaf(__uint_as_float(s1[c + 1]), scale_log2, neg_m));
          sum += (p0 + p1) + (p2 + p3);
          pk[c >> 1] = pack_bf16(p0, p1);
          pk[16 + (c >> 1)] = pack_bf16(p2, p3);
        }
        l += sum;
        const bool any_rescale = __any_sync(0xffffffffu, rescale);
        // the previous block's PV (possibly the previous tile's last) must have retired before the
        // shared-memory P buffer is overwritten / before O is rescaled (any_rescale implies j > 0)
        if ((j > 0 || it > 0) && (!P_TMEM || any_rescale)) {
          mbar_wait_guarded(p_empty, (gb - 1) & 1);
          tc_fence_after();
        }
        if (any_rescale) {
#pragma unroll 1
          for (int c4 = 0; c4 < kFaD / 32; ++c4) {
            uint32_t o[32];
            tmem_ld_32x32(t_lane + (uint32_t)(2 * kFaKv + c4 * 32), o);
            tmem_ld_wait();
#pragma unroll
            for (int i = 0; i < 32; ++i) o[i] = __float_as_uint(__uint_as_float(o[i]) * alpha);
            tmem_st_32x32(t_lane + (uint32_t)(2 * kFaKv + c4 * 32), o);
          }
          tmem_st_wait();
        }
        if constexpr (P_TMEM) {
          // P row (64 bf16 = 32 packed columns) over the first half of the S buffer this thread just read
          tmem_st_32x32(t_lane + s_col, pk);
          tmem_st_wait();
        } else {
          // P row (64 bf16 = 8 chunks of 16 B) into the 128B-swizzled K-major A-operand layout
#pragma unroll
          for (int ch = 0; ch < 8; ++ch)
            st_shared_v4(p_row + (((uint32_t)ch ^ sw) << 4),
                         make_uint4(pk[4 * ch], pk[4 * ch + 1], pk[4 * ch + 2], pk[4 * ch + 3]));
          fence_proxy_async();   // generic-proxy smem writes -> visible to the tensor core's async-proxy reads
        }
        // p_full is one barrier PER S BUFFER.  With a single count-4 barrier a fast warp could finish
        // block gb + 1 (its S is issued one block ahead and is already waiting) and arrive a second time
        // while a slow warp is still inside block gb: the phase would complete early, PV would read P rows
        // that are not written yet and the next Q K^T would overwrite S rows still being read (seen on
        // hardware as NaN rows once the persistent loop let the warps drift apart).  With one barrier per
        // buffer the second arrival lands on the other barrier, and block gb + 2 (same barrier again)
        // cannot start before the MMA warp has consumed phase gb.
        tc_fence_before();
        __syncwarp();
        if (lane == 0) mbar_arrive((gb & 1) ? p_full_b : p_full);
      }

      // ---- epilogue: O / l -> bf16 -> out[token, head, :]
      mbar_wait_guarded(o_full, it & 1);
      tc_fence_after();
      const float inv = l > 0.f ? 1.f / l : 0.f;
      bf16* orow = out + ((size_t)(m.q_begin + qi) * Hq + (size_t)(m.kvh * G + g)) * kFaD;
      auto store32 = [&](const uint32_t (&o)[32], bf16* dst) {            // 32 fp32 columns -> 64 bytes of the row
        uint4 wv[4];
#pragma unroll
        for (int v = 0; v < 4; ++v) {
          wv[v].x = pack_bf16(__uint_as_float(o[8 * v + 0]) * inv, __uint_as_float(o[8 * v + 1]) * inv);
          wv[v].y = pack_bf16(__uint_as_float(o[8 * v + 2]) * inv, __uint_as_float(o[8 * v + 3]) * inv);
          wv[v].z = pack_bf16(__uint_as_float(o[8 * v + 4]) * inv, __uint_as_float(o[8 * v + 5]) * inv);
          wv[v].w = pack_bf16(__uint_as_float(o[8 * v + 6]) * inv, __uint_as_float(o[8 * v + 7]) * inv);
        }
        if (wide) {
          st_v8(dst, wv[0], wv[1]);
          st_v8(dst + 16, wv[2], wv[3]);
        } else {
#pragma unroll
          for (int v = 0; v < 4; ++v) st_v4(dst + v * 8, wv[v]);
        }
      };
#pragma unroll 1
      for (int c4 = 0; c4 < kFaD / 32; c4 += 2) {   // two TMEM loads in flight per wait
        uint32_t o0[32], o1[32];
        tmem_ld_32x32(t_lane + (uint32_t)(2 * kFaKv + c4 * 32), o0);
        tmem_ld_32x32(t_lane + (uint32_t)(2 * kFaKv + c4 * 32 + 32), o1);
        tmem_ld_wait();
        if (c4 + 2 >= kFaD / 32) {   // O is in registers: the next tile's first PV may overwrite the accumulator
          tc_fence_before();
          __syncwarp();
          if (lane == 0) mbar_arrive(o_empty);
        }
        if (valid) {
          store32(o0, orow + c4 * 32);
          store32(o1, orow + c4 * 32 + 32);
        }
      }
      ++it;
    }
  }

  tc_fence_before();
  __syncthreads();
  if (warp == 5) {
    tc_fence_after();
    tmem_dealloc(tmem_base, kFaTmemCols);
  }
}

template <int G, bool P_TMEM>
static int launch_prefill_tc(const void* q, int64_t q_stride, int num_tokens, const void* k_cache,
                             const void* v_cache, int num_blocks, const int32_t* block_table, int max_blocks,
                             const int32_t* cu_q, const int32_t* ctx_lens, const int32_t* tile_seq,
                             const int32_t* tile_q0, int num_tiles, void* out, int Hq, int Hkv, float scale,
                             cudaStream_t stream) {
  auto kernel = paged_attn_prefill_tc_kernel<G, P_TMEM>;
  static bool attr_set = false;
  if (!attr_set) {
    cudaError_t e = cudaFuncSetAttribute(kernel, cudaFuncAttributeMaxDynamicSharedMemorySize, kFaSmemBytes);
    if (e == cudaSuccess)
      e = cudaFuncSetAttribute(kernel, cudaFuncAttributePreferredSharedMemoryCarveout, cudaSharedmemCarveoutMaxShared);
    if (e != cudaSuccess) {
      b200_set_error("attn_prefill(tc): cudaFuncSetAttribute(smem=%d): %s", kFaSmemBytes, cudaGetErrorString(e));
      return B200_ERR_CUDA;
    }
    attr_set = true;
  }
  CUtensorMap tq, tk, tv;
  int rc = make_tmap_bf16_2d(&tq, q, (uint64_t)num_tokens, (uint64_t)Hq * kFaD, (uint64_t)q_stride, 128 / G, 64);
  if (rc) return rc;
  const uint64_t kv_rows = (uint64_t)num_blocks * (uint64_t)Hkv * kFaPage;
  rc = make_tmap_bf16_2d(&tk, k_cache, kv_rows, kFaD, kFaD, kFaPage, 64);
  if (rc) return rc;
  rc = make_tmap_bf16_2d(&tv, v_cache, kv_rows, kFaD, kFaD, kFaPage, 64);
  if (rc) return rc;
  const float scale_log2 = scale * 1.4426950408889634f;
  static int persist = -1;
  if (persist < 0) {
    const char* pe = getenv("B200_ATTN_PERSIST");   // =0: one CTA per (tile, kv-head) instead of the persistent loop
    persist = (pe && pe[0] == '0') ? 0 : 1;
  }
  const long long total_work = (long long)num_tiles * Hkv;
  const int sms = device_info().sm_count > 0 ? device_info().sm_count : 148;
  const int grid = (int)(persist ? (total_work < 2LL * sms ? total_work : 2LL * sms) : total_work);
  cudaError_t e = launch_pdl(kernel, dim3(grid), dim3(kFaThreads), (size_t)kFaSmemBytes, stream, 1, tq, tk, tv,
                             block_table, max_blocks, cu_q, ctx_lens, tile_seq, tile_q0, num_tiles, (bf16*)out, Hq,
                             Hkv, scale_log2);
  if (e != cudaSuccess) {
    b200_set_error("b200_paged_attn_prefill(tc): launch failed: %s", cudaGetErrorString(e));
    (void)cudaGetLastError();
    return B200_ERR_CUDA;
  }
  return b200_check_launch("b200_paged_attn_prefill(tc)");
}

// Called by b200_paged_attn_prefill (attention.cu).  Returns 1 when this kernel does not cover the
// configuration (the caller then runs the mma.sync kernel), 0 on success, < 0 on error.
int paged_attn_prefill_tc(const void* q, int64_t q_stride, int num_tokens, const void* k_cache, const void* v_cache,
                          int num_blocks, const int32_t* block_table, int max_blocks, const int32_t* cu_q,
                          const int32_t* ctx_lens, const int32_t* tile_seq, const int32_t* tile_q0, int num_tiles,
                          void* out, int Hq, int Hkv, int D, float scale, cudaStream_t stream) {
  static int enabled = -1;
  if (enabled < 0) {
    const char* e = getenv("B200_ATTN_PREFILL_TC");   // =0: keep the mma.sync kernels (A/B runs)
    enabled = (e && e[0] == '0') ? 0 : 1;
  }
  const int G = Hq / Hkv;
  const DeviceInfo& d = device_info();
  if (!enabled || D != kFaD || (G != 4 && G != 8) || !d.valid || d.cc_major != 10) return 1;
  if (num_tokens <= 0 || num_blocks <= 0 || (q_stride % 8) != 0) return 1;
  static int p_tmem = -1;
  if (p_tmem < 0) {
    const char* e = getenv("B200_ATTN_P_TMEM");   // =0: P through shared memory (A/B runs)
    p_tmem = (e && e[0] == '0') ? 0 : 1;
  }
#define B200_FA_LAUNCH(GG, PT)                                                                                       \
  launch_prefill_tc<GG, PT>(q, q_stride, num_tokens, k_cache, v_cache, num_blocks, block_table, max_blocks, cu_q,    \
                            ctx_lens, tile_seq, tile_q0, num_tiles, out, Hq, Hkv, scale, stream)
  if (G == 4) return p_tmem ? B200_FA_LAUNCH(4, true) : B200_FA_LAUNCH(4, false);
  return p_tmem ? B200_FA_LAUNCH(8, true) : B200_FA_LAUNCH(8, false);
#undef B200_FA_LAUNCH
}

}  // namespace b200
