// Paged-KV attention for the decoder path (reference operator: FlashInferImpl.forward,
// $VLLM/v1/attention/backends/flashinfer.py:1095-1230 -> trtllm_batch_{decode,context}_with_kv_cache).
//
// KV cache layout (ours): k_cache, v_cache = [num_blocks][Hkv][16 tokens][D] bf16, so one
// (block, kv-head) page is a contiguous 16*D*2-byte run (4 KB at D=128) that a warp copies with
// coalesced 16-byte cp.async into XOR-swizzled shared memory (conflict-free ldmatrix).
//
// Decode kernel: HBM-bound (B*ctx*Hkv*D*2*2 bytes per layer, ~2 FLOP/byte).  One CTA per
// (sequence, kv-head, split); the G = Hq/Hkv query heads of the group form the M rows of
// mma.sync m16n8k16, so every K/V byte is fetched from HBM once per group.  Each of the 4 warps
// owns an independent 3-stage cp.async ring over its pages (2 CTAs/SM -> ~128 KB in flight per
// SM), runs an online softmax in fp32; warps are merged in shared memory, KV splits by the last CTA
// of the (sequence, kv-head) group to finish (arrival counter in the workspace) -- no merge launch.
//
// Prefill kernel: causal, ragged, chunk-capable.  One CTA per (query tile, kv-head); warp w
// holds 16 query tokens of one head of the group, K/V are streamed in 64-token chunks through a
// double-buffered CTA-wide cp.async pipeline shared by all heads of the group.
#include "common.cuh"
#include "runtime.h"
#include "../../include/b200.h"

#include <math.h>
#include <stdlib.h>

namespace b200 {

constexpr int kPage = 16;  // tokens per KV block (vLLM's default block_size)

// byte offset of 16-byte chunk `c` of row `r` in a swizzled [rows][D] bf16 tile
template <int D>
__device__ __forceinline__ uint32_t swz(int r, int c) {
  return (uint32_t)(r * (D * 2) + ((c ^ (r & 7)) << 4));
}

// ------------------------------------------------------------------------------------ decode
constexpr int kDecWarps = 4;
constexpr int kDecStages = 3;

template <int D>
struct DecCfg {
  static constexpr int kPageBytes = kPage * D * 2;          // one K (or V) page
  static constexpr int kStageBytes = 2 * kPageBytes;        // K + V
  static constexpr int kQBytes = 16 * D * 2;
  static constexpr int kWarpBytes = kDecStages * kStageBytes;
  static constexpr int kSmemBytes = kQBytes + kDecWarps * kWarpBytes;
};

template <int D>
__device__ __forceinline__ void load_page(uint8_t* dst, const bf16* __restrict__ src, int lane) {
  constexpr int CPR = D / 8;               // 16-byte chunks per row
  constexpr int CHUNKS = kPage * CPR;
#pragma unroll
  for (int i = 0; i < CHUNKS / 32; ++i) {
    const int ci = lane + 32 * i;
    const int r = ci / CPR, c = ci % CPR;
    cp_async_16(dst + swz<D>(r, c), src + (size_t)ci * 8);
  }
}

template <int D>
__global__ void __launch_bounds__(kDecWarps * 32, 2)
paged_attn_decode_kernel(const bf16* __restrict__ q, int64_t q_stride,
                         const bf16* __restrict__ k_cache, const bf16* __restrict__ v_cache,
                         const int32_t* __restrict__ block_table, int max_blocks,
                         const int32_t* __restrict__ seq_lens, bf16* __restrict__ out,
                         float* __restrict__ ws_o, float* __restrict__ ws_ml, int num_splits,
                         int Hq, int Hkv, float scale_log2, uint32_t* __restrict__ split_ctr) {
  pdl_launch_dependents();
  pdl_wait();
  using Cfg = DecCfg<D>;
  constexpr int KS = D / 16;  // k-steps of QK^T
  constexpr int NT = D / 8;   // n-tiles of PV
  extern __shared__ __align__(1024) uint8_t smem[];
  const int kvh = blockIdx.x, split = blockIdx.y, b = blockIdx.z;
  const int G = Hq / Hkv;
  const int warp = threadIdx.x >> 5, lane = threadIdx.x & 31;
  const int seq_len = seq_lens[b];
  const int n_pages = (seq_len + kPage - 1) / kPage;
  const int pps = (n_pages + num_splits - 1) / num_splits;
  const int p_begin = split * pps;
  const int p_end = min(n_pages, p_begin + pps);

  // ---- stage Q (G rows, zero-padded to 16) in swizzled smem
  uint8_t* q_smem = smem;
  {
    constexpr int CPR = D / 8;
    for (int ci = threadIdx.x; ci < 16 * CPR; ci += blockDim.x) {
      const int r = ci / CPR, c = ci % CPR;
      uint4 v = make_uint4(0, 0, 0, 0);
      if (r < G) v = ld_v4(q + (size_t)b * q_stride + (size_t)(kvh * G + r) * D + c * 8);
      *reinterpret_cast<uint4*>(q_smem + swz<D>(r, c)) = v;
    }
  }
  __syncthreads();
  uint32_t qf[KS][4];
  {
    const int i = lane >> 3, rr = lane & 7;
    const int row = rr + (i & 1) * 8;
#pragma unroll
    for (int ks = 0; ks < KS; ++ks)
      ldsm_x4(smem_u32(q_smem + swz<D>(row, 2 * ks + (i >> 1))), qf[ks][0], qf[ks][1], qf[ks][2],
              qf[ks][3]);
  }

  uint8_t* my = smem + Cfg::kQBytes + warp * Cfg::kWarpBytes;
  const int32_t* bt = block_table + (size_t)b * max_blocks;
  const int first = p_begin + warp;
  const int n_my = first < p_end ? (p_end - first + kDecWarps - 1) / kDecWarps : 0;

  auto issue = [&](int i) {
    const int page = first + i * kDecWarps;
    const int blk = __ldg(bt + page);
    const size_t base = ((size_t)blk * Hkv + kvh) * (size_t)(kPage * D);
    uint8_t* st = my + (i % kDecStages) * Cfg::kStageBytes;
    load_page<D>(st, k_cache + base, lane);
    load_page<D>(st + Cfg::kPageBytes, v_cache + base, lane);
  };

#pragma unroll
  for (int i = 0; i < kDecStages - 1; ++i) {
    if (i < n_my) issue(i);
    cp_async_commit();
  }

  float o[NT][4];
#pragma unroll
  for (int t = 0; t < NT; ++t) o[t][0] = o[t][1] = o[t][2] = o[t][3] = 0.f;
  float m0 = -INFINITY, m1 = -INFINITY, l0 = 0.f, l1 = 0.f;

  for (int i = 0; i < n_my; ++i) {
    cp_async_wait<kDecStages - 2>();
    __syncwarp();
    if (i + kDecStages - 1 < n_my) issue(i + kDecStages - 1);
    cp_async_commit();

    const uint8_t* ks_ = my + (i % kDecStages) * Cfg::kStageBytes;
    const uint8_t* vs_ = ks_ + Cfg::kPageBytes;
    const uint32_t k_base = smem_u32(ks_), v_base = smem_u32(vs_);

    // S = Q K^T : 16 (heads, padded) x 16 (tokens)
    float s[2][4] = {{0.f, 0.f, 0.f, 0.f}, {0.f, 0.f, 0.f, 0.f}};
    {
      const int mi = lane >> 3, rr = lane & 7;
      const int tok = rr + (mi >> 1) * 8;
#pragma unroll
      for (int ks = 0; ks < KS; ++ks) {
        uint32_t b0, b1, b2, b3;
        ldsm_x4(k_base + swz<D>(tok, 2 * ks + (mi & 1)), b0, b1, b2, b3);
        mma_bf16_16816(s[0], qf[ks], b0, b1);
        mma_bf16_16816(s[1], qf[ks], b2, b3);
      }
    }
    // scale, mask the tail of the last page, online softmax (fp32, base-2)
    const int tok0 = (first + i * kDecWarps) * kPage + 2 * (lane & 3);
    float mx0 = -INFINITY, mx1 = -INFINITY;
#pragma unroll
    for (int t = 0; t < 2; ++t) {
#pragma unroll
      for (int e = 0; e < 4; ++e) {
        const int tok = tok0 + t * 8 + (e & 1);
        const float v = tok < seq_len ? s[t][e] * scale_log2 : -INFINITY;
        s[t][e] = v;
        if (e < 2) mx0 = fmaxf(mx0, v); else mx1 = fmaxf(mx1, v);
      }
    }
    mx0 = fmaxf(mx0, __shfl_xor_sync(0xffffffffu, mx0, 1));
    mx0 = fmaxf(mx0, __shfl_xor_sync(0xffffffffu, mx0, 2));
    mx1 = fmaxf(mx1, __shfl_xor_sync(0xffffffffu, mx1, 1));
    mx1 = fmaxf(mx1, __shfl_xor_sync(0xffffffffu, mx1, 2));
    const float mn0 = fmaxf(m0, mx0), mn1 = fmaxf(m1, mx1);
    const float a0 = exp2f(m0 - mn0), a1 = exp2f(m1 - mn1);
    m0 = mn0; m1 = mn1;
    float rs0 = 0.f, rs1 = 0.f;
#pragma unroll
    for (int t = 0; t < 2; ++t) {
      s[t][0] = exp2f(s[t][0] - mn0); s[t][1] = exp2f(s[t][1] - mn0);
      s[t][2] = exp2f(s[t][2] - mn1); s[t][3] = exp2f(s[t][3] - mn1);
      rs0 += s[t][0] + s[t][1];
      rs1 += s[t][2] + s[t][3];
    }
    l0 = l0 * a0 + rs0;
    l1 = l1 * a1 + rs1;
#pragma unroll
    for (int t = 0; t < NT; ++t) { o[t][0] *= a0; o[t][1] *= a0; o[t][2] *= a1; o[t][3] *= a1; }

    // O += P V
    const uint32_t pa[4] = {pack_bf16(s[0][0], s[0][1]), pack_bf16(s[0][2], s[0][3]),
                            pack_bf16(s[1][0], s[1][1]), pack_bf16(s[1][2], s[1][3])};
    {
      const int mi = lane >> 3, rr = lane & 7;
      const int tok = rr + (mi & 1) * 8;
#pragma unroll
      for (int c = 0; c < NT; c += 2) {
        uint32_t b0, b1, b2, b3;
        ldsm_x4_t(v_base + swz<D>(tok, c + (mi >> 1)), b0, b1, b2, b3);
        mma_bf16_16816(o[c], pa, b0, b1);
        mma_bf16_16816(o[c + 1], pa, b2, b3);
      }
    }
  }
  cp_async_wait<0>();
  l0 += __shfl_xor_sync(0xffffffffu, l0, 1);
  l0 += __shfl_xor_sync(0xffffffffu, l0, 2);
  l1 += __shfl_xor_sync(0xffffffffu, l1, 1);
  l1 += __shfl_xor_sync(0xffffffffu, l1, 2);

  // ---- each warp parks (m, l, O) for its real rows in its own (now idle) ring, then merge
  __syncwarp();
  float* po = reinterpret_cast<float*>(my);              // [G][D]
  float* pml = po + 16 * D;                               // [16][2]
  {
    const int r0 = lane >> 2, r1 = r0 + 8, cb = 2 * (lane & 3);
#pragma unroll
    for (int t = 0; t < NT; ++t) {
      if (r0 < G) { po[r0 * D + t * 8 + cb] = o[t][0]; po[r0 * D + t * 8 + cb + 1] = o[t][1]; }
      if (r1 < G) { po[r1 * D + t * 8 + cb] = o[t][2]; po[r1 * D + t * 8 + cb + 1] = o[t][3]; }
    }
    if ((lane & 3) == 0) {
      if (r0 < G) { pml[r0 * 2] = m0; pml[r0 * 2 + 1] = l0; }
      if (r1 < G) { pml[r1 * 2] = m1; pml[r1 * 2 + 1] = l1; }
    }
  }
  __syncthreads();
  const uint8_t* rings = smem + Cfg::kQBytes;
  for (int idx = threadIdx.x; idx < G * D; idx += blockDim.x) {
    const int g = idx / D, d = idx % D;
    float mm = -INFINITY;
#pragma unroll
    for (int w = 0; w < kDecWarps; ++w) {
      const float* wml = reinterpret_cast<const float*>(rings + w * Cfg::kWarpBytes) + 16 * D;
      mm = fmaxf(mm, wml[g * 2]);
    }
    float num = 0.f, den = 0.f;
    if (mm > -INFINITY) {
#pragma unroll
      for (int w = 0; w < kDecWarps; ++w) {
        const float* wo = reinterpret_cast<const float*>(rings + w * Cfg::kWarpBytes);
        const float* wml = wo + 16 * D;
        const float sc = exp2f(wml[g * 2] - mm);
        num += sc * wo[g * D + d];
        den += sc * wml[g * 2 + 1];
      }
    }
    const int h = kvh * G + g;
    if (num_splits == 1) {
      out[((size_t)b * Hq + h) * D + d] = __float2bfloat16_rn(den > 0.f ? num / den : 0.f);
    } else {
      const size_t slot = ((size_t)b * Hq + h) * num_splits + split;
      ws_o[slot * D + d] = num;
      if (d == 0) { ws_ml[slot * 2] = mm; ws_ml[slot * 2 + 1] = den; }
    }
  }
  if (num_splits == 1) return;
  // ---- cross-split merge by the LAST CTA of this (sequence, kv-head) to arrive: no extra launch.
  // Every CTA publishes its partials (release), bumps the group's counter, and the one that sees
  // num_splits - 1 earlier arrivals reads all partials back (L2) and writes the final rows; it
  // also resets the counter, so the workspace is reusable by the next launch / graph replay.
  __shared__ int s_last;
  __threadfence();
  __syncthreads();
  if (threadIdx.x == 0) {
    const uint32_t prev = atomicAdd(split_ctr + (size_t)b * Hkv + kvh, 1u);
    s_last = (prev == (uint32_t)(num_splits - 1));
    if (s_last) split_ctr[(size_t)b * Hkv + kvh] = 0u;
  }
  __syncthreads();
  if (!s_last) return;
  __threadfence();
  for (int idx = threadIdx.x; idx < G * D; idx += blockDim.x) {
    const int g = idx / D, d = idx % D;
    const size_t row = (size_t)b * Hq + kvh * G + g;
    const float* ml = ws_ml + row * num_splits * 2;
    float mm = -INFINITY;
    for (int sp = 0; sp < num_splits; ++sp) mm = fmaxf(mm, __ldcg(ml + sp * 2));
    float num = 0.f, den = 0.f;
    for (int sp = 0; sp < num_splits; ++sp) {
      const float m = __ldcg(ml + sp * 2);
      if (m == -INFINITY) continue;
      const float sc = exp2f(m - mm);
      den += sc * __ldcg(ml + sp * 2 + 1);
      num += sc * __ldcg(ws_o + (row * num_splits + sp) * D + d);
    }
    out[row * D + d] = __float2bfloat16_rn(den > 0.f ? num / den : 0.f);
  }
}

// ----------------------------------------------------------------------------------- prefill
constexpr int kPfChunk = 64;  // KV tokens per pipeline stage

template <int D>
struct PfCfg {
  static constexpr int kTileBytes = kPfChunk * D * 2;   // K (or V) chunk
  static constexpr int kStageBytes = 2 * kTileBytes;
  static constexpr int kSmemBytes = 2 * kStageBytes;
};

template <int D>
__global__ void __launch_bounds__(256, 1)
paged_attn_prefill_kernel(const bf16* __restrict__ q, int64_t q_stride,
                          const bf16* __restrict__ k_cache, const bf16* __restrict__ v_cache,
                          const int32_t* __restrict__ block_table, int max_blocks,
                          const int32_t* __restrict__ cu_q, const int32_t* __restrict__ ctx_lens,
                          const int32_t* __restrict__ tile_seq, const int32_t* __restrict__ tile_q0,
                          bf16* __restrict__ out, int Hq, int Hkv, int QT, float scale_log2) {
  pdl_launch_dependents();
  pdl_wait();
  using Cfg = PfCfg<D>;
  constexpr int KS = D / 16;
  constexpr int NT = D / 8;
  constexpr int CPR = D / 8;
  extern __shared__ __align__(1024) uint8_t smem[];
  const int tile = blockIdx.x, kvh = blockIdx.y;
  const int G = Hq / Hkv;
  const int warp = threadIdx.x >> 5, lane = threadIdx.x & 31;
  const int seq = tile_seq[tile], q0 = tile_q0[tile];
  const int q_begin = cu_q[seq];
  const int q_len = cu_q[seq + 1] - q_begin;
  const int ctx = ctx_lens[seq];
  const int total_len = ctx + q_len;
  const int g = warp % G, qsub = warp / G;
  const int head = kvh * G + g;
  const int tile_tokens = 16 * QT;
  // causal horizon of the whole tile / of this warp (exclusive)
  const int tile_kv_end = min(total_len, ctx + q0 + tile_tokens);
  const int warp_kv_end = min(total_len, ctx + q0 + qsub * 16 + 16);
  const int n_chunks = (tile_kv_end + kPfChunk - 1) / kPfChunk;
  const int n_pages_seq = (total_len + kPage - 1) / kPage;
  const int32_t* bt = block_table + (size_t)seq * max_blocks;

  // ---- Q fragments straight from global memory (one-off, 4-byte loads)
  const int r0 = lane >> 2, r1 = r0 + 8, cb = 2 * (lane & 3);
  const int qi0 = q0 + qsub * 16 + r0, qi1 = q0 + qsub * 16 + r1;  // offsets inside the chunk
  uint32_t qf[KS][4];
  {
    const bf16* qp0 = q + (size_t)(q_begin + qi0) * q_stride + (size_t)head * D;
    const bf16* qp1 = q + (size_t)(q_begin + qi1) * q_stride + (size_t)head * D;
    const bool v0 = qi0 < q_len, v1 = qi1 < q_len;
#pragma unroll
    for (int ks = 0; ks < KS; ++ks) {
      qf[ks][0] = v0 ? *reinterpret_cast<const uint32_t*>(qp0 + ks * 16 + cb) : 0u;
      qf[ks][1] = v1 ? *reinterpret_cast<const uint32_t*>(qp1 + ks * 16 + cb) : 0u;
      qf[ks][2] = v0 ? *reinterpret_cast<const uint32_t*>(qp0 + ks * 16 + 8 + cb) : 0u;
      qf[ks][3] = v1 ? *reinterpret_cast<const uint32_t*>(qp1 + ks * 16 + 8 + cb) : 0u;
    }
  }

  auto issue = [&](int c) {
    uint8_t* st = smem + (c & 1) * Cfg::kStageBytes;
    for (int ci = threadIdx.x; ci < kPfChunk * CPR; ci += blockDim.x) {
      const int r = ci / CPR, cc = ci % CPR;
      const int tok = c * kPfChunk + r;
      const int page = tok / kPage;
      const bool valid = page < n_pages_seq;
      const int blk = valid ? __ldg(bt + page) : 0;
      const size_t off = (((size_t)blk * Hkv + kvh) * kPage + (tok % kPage)) * D + cc * 8;
      cp_async_16_zfill(st + swz<D>(r, cc), k_cache + off, valid);
      cp_async_16_zfill(st + Cfg::kTileBytes + swz<D>(r, cc), v_cache + off, valid);
    }
  };

  float o[NT][4];
#pragma unroll
  for (int t = 0; t < NT; ++t) o[t][0] = o[t][1] = o[t][2] = o[t][3] = 0.f;
  float m0 = -INFINITY, m1 = -INFINITY, l0 = 0.f, l1 = 0.f;
  const int qpos0 = ctx + qi0, qpos1 = ctx + qi1;

  if (n_chunks > 0) issue(0);
  cp_async_commit();
  for (int c = 0; c < n_chunks; ++c) {
    if (c + 1 < n_chunks) issue(c + 1);
    cp_async_commit();
    cp_async_wait<1>();
    __syncthreads();
    if (c * kPfChunk < warp_kv_end) {
      const uint32_t k_base = smem_u32(smem + (c & 1) * Cfg::kStageBytes);
      const uint32_t v_base = k_base + Cfg::kTileBytes;
      float s[8][4];
#pragma unroll
      for (int t = 0; t < 8; ++t) s[t][0] = s[t][1] = s[t][2] = s[t][3] = 0.f;
      {
        const int mi = lane >> 3, rr = lane & 7;
#pragma unroll
        for (int tp = 0; tp < 4; ++tp) {  // 16 tokens per pair of n-tiles
          const int tok = tp * 16 + rr + (mi >> 1) * 8;
#pragma unroll
          for (int ks = 0; ks < KS; ++ks) {
            uint32_t b0, b1, b2, b3;
            ldsm_x4(k_base + swz<D>(tok, 2 * ks + (mi & 1)), b0, b1, b2, b3);
            mma_bf16_16816(s[2 * tp], qf[ks], b0, b1);
            mma_bf16_16816(s[2 * tp + 1], qf[ks], b2, b3);
          }
        }
      }
      float mx0 = -INFINITY, mx1 = -INFINITY;
#pragma unroll
      for (int t = 0; t < 8; ++t) {
#pragma unroll
        for (int e = 0; e < 4; ++e) {
          const int kv = c * kPfChunk + t * 8 + cb + (e & 1);
          const int qp = e < 2 ? qpos0 : qpos1;
          const float v = kv <= qp ? s[t][e] * scale_log2 : -INFINITY;
          s[t][e] = v;
          if (e < 2) mx0 = fmaxf(mx0, v); else mx1 = fmaxf(mx1, v);
        }
      }
      mx0 = fmaxf(mx0, __shfl_xor_sync(0xffffffffu, mx0, 1));
      mx0 = fmaxf(mx0, __shfl_xor_sync(0xffffffffu, mx0, 2));
      mx1 = fmaxf(mx1, __shfl_xor_sync(0xffffffffu, mx1, 1));
      mx1 = fmaxf(mx1, __shfl_xor_sync(0xffffffffu, mx1, 2));
      const float mn0 = fmaxf(m0, mx0), mn1 = fmaxf(m1, mx1);
      // a fully masked row keeps m = -inf; use 0 as the subtraction base to avoid inf - inf
      const float base0 = mn0 == -INFINITY ? 0.f : mn0, base1 = mn1 == -INFINITY ? 0.f : mn1;
      const float a0 = exp2f(m0 - base0), a1 = exp2f(m1 - base1);
      m0 = mn0; m1 = mn1;
      float rs0 = 0.f, rs1 = 0.f;
#pragma unroll
      for (int t = 0; t < 8; ++t) {
        s[t][0] = exp2f(s[t][0] - base0); s[t][1] = exp2f(s[t][1] - base0);
        s[t][2] = exp2f(s[t][2] - base1); s[t][3] = exp2f(s[t][3] - base1);
        rs0 += s[t][0] + s[t][1];
        rs1 += s[t][2] + s[t][3];
      }
      l0 = l0 * a0 + rs0;
      l1 = l1 * a1 + rs1;
#pragma unroll
      for (int t = 0; t < NT; ++t) { o[t][0] *= a0; o[t][1] *= a0; o[t][2] *= a1; o[t][3] *= a1; }
      {
        const int mi = lane >> 3, rr = lane & 7;
#pragma unroll
        for (int j = 0; j < 4; ++j) {  // k-steps of 16 tokens
          const uint32_t pa[4] = {pack_bf16(s[2 * j][0], s[2 * j][1]),
                                  pack_bf16(s[2 * j][2], s[2 * j][3]),
                                  pack_bf16(s[2 * j + 1][0], s[2 * j + 1][1]),
                                  pack_bf16(s[2 * j + 1][2], s[2 * j + 1][3])};
          const int tok = j * 16 + rr + (mi & 1) * 8;
#pragma unroll
          for (int cc = 0; cc < NT; cc += 2) {
            uint32_t b0, b1, b2, b3;
            ldsm_x4_t(v_base + swz<D>(tok, cc + (mi >> 1)), b0, b1, b2, b3);
            mma_bf16_16816(o[cc], pa, b0, b1);
            mma_bf16_16816(o[cc + 1], pa, b2, b3);
          }
        }
      }
    }
    __syncthreads();  // everyone is done with stage (c & 1) before chunk c+2 overwrites it
  }
  cp_async_wait<0>();
  l0 += __shfl_xor_sync(0xffffffffu, l0, 1);
  l0 += __shfl_xor_sync(0xffffffffu, l0, 2);
  l1 += __shfl_xor_sync(0xffffffffu, l1, 1);
  l1 += __shfl_xor_sync(0xffffffffu, l1, 2);
  const float i0 = l0 > 0.f ? 1.f / l0 : 0.f, i1 = l1 > 0.f ? 1.f / l1 : 0.f;
  const size_t orow = (size_t)Hq * D;
  if (qi0 < q_len) {
    bf16* op = out + (size_t)(q_begin + qi0) * orow + (size_t)head * D + cb;
#pragma unroll
    for (int t = 0; t < NT; ++t)
      *reinterpret_cast<uint32_t*>(op + t * 8) = pack_bf16(o[t][0] * i0, o[t][1] * i0);
  }
  if (qi1 < q_len) {
    bf16* op = out + (size_t)(q_begin + qi1) * orow + (size_t)head * D + cb;
#pragma unroll
    for (int t = 0; t < NT; ++t)
      *reinterpret_cast<uint32_t*>(op + t * 8) = pack_bf16(o[t][2] * i1, o[t][3] * i1);
  }
}

// ----------------------------------------------------------------------------- prefill, v2
// Same tiling as above (CTA = 16*QT query tokens x G heads, one warp per 16x1), restructured for
// occupancy: Q fragments are re-read from (swizzled) shared memory instead of living in 32
// registers, K/V stream in 32-token chunks through a 3-stage cp.async ring with ONE barrier per
// chunk, and the kernel fits 128 registers so two CTAs share an SM (16 warps instead of 8).
constexpr int kPf2Chunk = 32;
constexpr int kPf2Stages = 3;

template <int D>
struct Pf2Cfg {
  static constexpr int kQWarpBytes = 16 * D * 2;               // one warp's 16 x D query tile
  static constexpr int kTileBytes = kPf2Chunk * D * 2;         // K (or V) chunk
  static constexpr int kStageBytes = 2 * kTileBytes;
  static constexpr int smem_bytes(int warps) { return warps * kQWarpBytes + kPf2Stages * kStageBytes; }
};

template <int D>
__global__ void __launch_bounds__(256, 2)
paged_attn_prefill_v2_kernel(const bf16* __restrict__ q, int64_t q_stride,
                             const bf16* __restrict__ k_cache, const bf16* __restrict__ v_cache,
                             const int32_t* __restrict__ block_table, int max_blocks,
                             const int32_t* __restrict__ cu_q, const int32_t* __restrict__ ctx_lens,
                             const int32_t* __restrict__ tile_seq, const int32_t* __restrict__ tile_q0,
                             bf16* __restrict__ out, int Hq, int Hkv, int QT, float scale_log2) {
  pdl_launch_dependents();
  pdl_wait();
  using Cfg = Pf2Cfg<D>;
  constexpr int KS = D / 16;
  constexpr int NT = D / 8;
  constexpr int CPR = D / 8;
  extern __shared__ __align__(1024) uint8_t smem[];
  const int tile = blockIdx.x, kvh = blockIdx.y;
  const int G = Hq / Hkv;
  const int nwarps = blockDim.x >> 5;
  const int warp = threadIdx.x >> 5, lane = threadIdx.x & 31;
  const int seq = tile_seq[tile], q0 = tile_q0[tile];
  const int q_begin = cu_q[seq];
  const int q_len = cu_q[seq + 1] - q_begin;
  const int ctx = ctx_lens[seq];
  const int total_len = ctx + q_len;
  const int g = warp % G, qsub = warp / G;
  const int head = kvh * G + g;
  const int tile_tokens = 16 * QT;
  const int tile_kv_end = min(total_len, ctx + q0 + tile_tokens);
  const int warp_kv_end = min(total_len, ctx + q0 + qsub * 16 + 16);
  const int n_chunks = (tile_kv_end + kPf2Chunk - 1) / kPf2Chunk;
  const int n_pages_seq = (total_len + kPage - 1) / kPage;
  const int32_t* bt = block_table + (size_t)seq * max_blocks;

  uint8_t* q_smem = smem + warp * Cfg::kQWarpBytes;
  uint8_t* kv_smem = smem + nwarps * Cfg::kQWarpBytes;

  // ---- this warp's 16 x D query tile -> swizzled smem (cp.async group 0, with chunk 0)
  for (int ci = lane; ci < 16 * CPR; ci += 32) {
    const int r = ci / CPR, c = ci % CPR;
    const int qi = q0 + qsub * 16 + r;
    const bool valid = qi < q_len;
    const bf16* src = q + (size_t)(q_begin + (valid ? qi : 0)) * q_stride + (size_t)head * D + c * 8;
    cp_async_16_zfill(q_smem + swz<D>(r, c), src, valid);
  }

  auto issue = [&](int c) {
    uint8_t* st = kv_smem + (c % kPf2Stages) * Cfg::kStageBytes;
    for (int ci = threadIdx.x; ci < kPf2Chunk * CPR; ci += blockDim.x) {
      const int r = ci / CPR, cc = ci % CPR;
      const int tok = c * kPf2Chunk + r;
      const int page = tok / kPage;
      const bool valid = page < n_pages_seq;
      const int blk = valid ? __ldg(bt + page) : 0;
      const size_t off = (((size_t)blk * Hkv + kvh) * kPage + (tok % kPage)) * D + cc * 8;
      cp_async_16_zfill(st + swz<D>(r, cc), k_cache + off, valid);
      cp_async_16_zfill(st + Cfg::kTileBytes + swz<D>(r, cc), v_cache + off, valid);
    }
  };

  const int r0 = lane >> 2, r1 = r0 + 8, cb = 2 * (lane & 3);
  const int qi0 = q0 + qsub * 16 + r0, qi1 = q0 + qsub * 16 + r1;
  const int qpos0 = ctx + qi0, qpos1 = ctx + qi1;
  float o[NT][4];
#pragma unroll
  for (int t = 0; t < NT; ++t) o[t][0] = o[t][1] = o[t][2] = o[t][3] = 0.f;
  float m0 = -INFINITY, m1 = -INFINITY, l0 = 0.f, l1 = 0.f;

#pragma unroll
  for (int c = 0; c < kPf2Stages - 1; ++c) {
    if (c < n_chunks) issue(c);
    cp_async_commit();
  }
  const uint32_t q_base = smem_u32(q_smem);
  const int mi = lane >> 3, rr = lane & 7;
  const int q_row = rr + (mi & 1) * 8;          // A-fragment row supplied by this lane

  for (int c = 0; c < n_chunks; ++c) {
    cp_async_wait<kPf2Stages - 2>();
    __syncthreads();                           // chunk c landed; everyone is done with chunk c-1
    if (c + kPf2Stages - 1 < n_chunks) issue(c + kPf2Stages - 1);
    cp_async_commit();
    if (c * kPf2Chunk >= warp_kv_end) continue;  // fully masked for this warp (warp-uniform)

    const uint32_t k_base = smem_u32(kv_smem + (c % kPf2Stages) * Cfg::kStageBytes);
    const uint32_t v_base = k_base + Cfg::kTileBytes;
    float s[4][4];
#pragma unroll
    for (int t = 0; t < 4; ++t) s[t][0] = s[t][1] = s[t][2] = s[t][3] = 0.f;
#pragma unroll
    for (int ks = 0; ks < KS; ++ks) {
      uint32_t qf[4];
      ldsm_x4(q_base + swz<D>(q_row, 2 * ks + (mi >> 1)), qf[0], qf[1], qf[2], qf[3]);
#pragma unroll
      for (int tp = 0; tp < 2; ++tp) {
        const int tok = tp * 16 + rr + (mi >> 1) * 8;
        uint32_t b0, b1, b2, b3;
        ldsm_x4(k_base + swz<D>(tok, 2 * ks + (mi & 1)), b0, b1, b2, b3);
        mma_bf16_16816(s[2 * tp], qf, b0, b1);
        mma_bf16_16816(s[2 * tp + 1], qf, b2, b3);
      }
    }
    float mx0 = -INFINITY, mx1 = -INFINITY;
#pragma unroll
    for (int t = 0; t < 4; ++t) {
#pragma unroll
      for (int e = 0; e < 4; ++e) {
        const int kv = c * kPf2Chunk + t * 8 + cb + (e & 1);
        const int qp = e < 2 ? qpos0 : qpos1;
        const float v = kv <= qp ? s[t][e] * scale_log2 : -INFINITY;
        s[t][e] = v;
        if (e < 2) mx0 = fmaxf(mx0, v); else mx1 = fmaxf(mx1, v);
      }
    }
    mx0 = fmaxf(mx0, __shfl_xor_sync(0xffffffffu, mx0, 1));
    mx0 = fmaxf(mx0, __shfl_xor_sync(0xffffffffu, mx0, 2));
    mx1 = fmaxf(mx1, __shfl_xor_sync(0xffffffffu, mx1, 1));
    mx1 = fmaxf(mx1, __shfl_xor_sync(0xffffffffu, mx1, 2));
    const float mn0 = fmaxf(m0, mx0), mn1 = fmaxf(m1, mx1);
    const float base0 = mn0 == -INFINITY ? 0.f : mn0, base1 = mn1 == -INFINITY ? 0.f : mn1;
    const float a0 = exp2f(m0 - base0), a1 = exp2f(m1 - base1);
    m0 = mn0; m1 = mn1;
    float rs0 = 0.f, rs1 = 0.f;
#pragma unroll
    for (int t = 0; t < 4; ++t) {
      s[t][0] = exp2f(s[t][0] - base0); s[t][1] = exp2f(s[t][1] - base0);
      s[t][2] = exp2f(s[t][2] - base1); s[t][3] = exp2f(s[t][3] - base1);
      rs0 += s[t][0] + s[t][1];
      rs1 += s[t][2] + s[t][3];
    }
    l0 = l0 * a0 + rs0;
    l1 = l1 * a1 + rs1;
#pragma unroll
    for (int t = 0; t < NT; ++t) { o[t][0] *= a0; o[t][1] *= a0; o[t][2] *= a1; o[t][3] *= a1; }
#pragma unroll
    for (int j = 0; j < 2; ++j) {
      const uint32_t pa[4] = {pack_bf16(s[2 * j][0], s[2 * j][1]), pack_bf16(s[2 * j][2], s[2 * j][3]),
                              pack_bf16(s[2 * j + 1][0], s[2 * j + 1][1]),
                              pack_bf16(s[2 * j + 1][2], s[2 * j + 1][3])};
      const int tok = j * 16 + rr + (mi & 1) * 8;
#pragma unroll
      for (int cc = 0; cc < NT; cc += 2) {
        uint32_t b0, b1, b2, b3;
        ldsm_x4_t(v_base + swz<D>(tok, cc + (mi >> 1)), b0, b1, b2, b3);
        mma_bf16_16816(o[cc], pa, b0, b1);
        mma_bf16_16816(o[cc + 1], pa, b2, b3);
      }
    }
  }
  cp_async_wait<0>();
  l0 += __shfl_xor_sync(0xffffffffu, l0, 1);
  l0 += __shfl_xor_sync(0xffffffffu, l0, 2);
  l1 += __shfl_xor_sync(0xffffffffu, l1, 1);
  l1 += __shfl_xor_sync(0xffffffffu, l1, 2);
  const float i0 = l0 > 0.f ? 1.f / l0 : 0.f, i1 = l1 > 0.f ? 1.f / l1 : 0.f;
  const size_t orow = (size_t)Hq * D;
  if (qi0 < q_len) {
    bf16* op = out + (size_t)(q_begin + qi0) * orow + (size_t)head * D + cb;
#pragma unroll
    for (int t = 0; t < NT; ++t)
      *reinterpret_cast<uint32_t*>(op + t * 8) = pack_bf16(o[t][0] * i0, o[t][1] * i0);
  }
  if (qi1 < q_len) {
    bf16* op = out + (size_t)(q_begin + qi1) * orow + (size_t)head * D + cb;
#pragma unroll
    for (int t = 0; t < NT; ++t)
      *reinterpret_cast<uint32_t*>(op + t * 8) = pack_bf16(o[t][2] * i1, o[t][3] * i1);
  }
}

template <typename K>
static int set_smem_attr(K kernel, int bytes, const char* what) {
  cudaError_t e = cudaFuncSetAttribute(kernel, cudaFuncAttributeMaxDynamicSharedMemorySize, bytes);
  if (e != cudaSuccess) {
    b200_set_error("%s: cudaFuncSetAttribute(smem=%d): %s", what, bytes, cudaGetErrorString(e));
    return B200_ERR_CUDA;
  }
  return B200_OK;
}

}  // namespace b200

using namespace b200;

extern "C" int b200_attn_decode_splits(int B, int Hkv, int max_seq_len, int block_size) {
  if (B <= 0 || Hkv <= 0 || max_seq_len <= 0 || block_size <= 0) return 1;
  const int sms = device_info().sm_count > 0 ? device_info().sm_count : 148;
  const int pages = (max_seq_len + block_size - 1) / block_size;
  const int base = B * Hkv;                     // CTAs without splitting
  const int slots = 2 * sms;                    // resident CTAs (2 per SM)
  // one wave is enough once the unsplit grid nearly fills the machine: a split costs every CTA its
  // fixed part (Q staging, in-CTA merge, partials) again -- measured on the TP=8 shard shape
  // (B=256, one kv-head per rank) 3 splits ran at 0.3 of the HBM roofline
  if (base * 4 >= slots * 3) return 1;
  int want = (2 * slots + base - 1) / base;     // otherwise: about two waves
  int cap = pages / 8;                          // >= 8 pages (2 per warp) per split
  if (cap < 1) cap = 1;
  if (want > cap) want = cap;
  if (want > 32) want = 32;
  return want < 1 ? 1 : want;
}

// partial O [B*Hq*splits][D] + (m, l) [B*Hq*splits][2] fp32, then one arrival counter per (sequence,
// kv-head) group.  The counters sit at the START of the workspace (zeroed once by the caller,
// b200_attn_decode_workspace_init; self-resetting afterwards) so that one buffer sized for the worst
// batch serves every batch size.
static size_t attn_decode_ctr_bytes(int B, int Hq) { return ((size_t)B * Hq * sizeof(uint32_t) + 255) & ~(size_t)255; }

extern "C" size_t b200_attn_decode_workspace_bytes(int B, int Hq, int D, int num_splits) {
  if (num_splits <= 1) return 16;
  return attn_decode_ctr_bytes(B, Hq) + (size_t)B * Hq * num_splits * (size_t)(D + 2) * sizeof(float);
}

extern "C" int b200_paged_attn_decode(const void* q, int64_t q_stride, const void* k_cache,
                                      const void* v_cache, const int32_t* block_table,
                                      int max_blocks, const int32_t* seq_lens, void* out,
                                      void* workspace, int num_splits, int B, int Hq, int Hkv,
                                      int D, int block_size, float scale, void* stream) {
  B200_REQUIRE(B >= 0 && Hq > 0 && Hkv > 0 && Hq % Hkv == 0, "attn_decode: bad head counts");
  B200_REQUIRE(Hq / Hkv <= 16, "attn_decode: group size %d > 16 unsupported", Hq / Hkv);
  B200_REQUIRE(block_size == kPage, "attn_decode: block_size must be %d", kPage);
  B200_REQUIRE(D == 128 || D == 64, "attn_decode: head_dim %d unsupported (64, 128)", D);
  B200_REQUIRE(num_splits >= 1 && (num_splits == 1 || workspace), "attn_decode: bad splits");
  B200_REQUIRE(q_stride % 8 == 0, "attn_decode: q_stride must be a multiple of 8");
  if (B == 0) return B200_OK;
  cudaStream_t s = (cudaStream_t)stream;
  const float scale_log2 = scale * 1.4426950408889634f;
  // counters live in the first bytes of the workspace, sized for THIS call's B (<= the B the buffer was
  // sized and zeroed for: a prefix of a zeroed, self-resetting array is itself zeroed)
  uint32_t* split_ctr = reinterpret_cast<uint32_t*>(workspace);
  float* ws_o = workspace ? reinterpret_cast<float*>(reinterpret_cast<uint8_t*>(workspace) + attn_decode_ctr_bytes(B, Hq))
                          : nullptr;
  float* ws_ml = ws_o ? ws_o + (size_t)B * Hq * num_splits * D : nullptr;
  dim3 grid(Hkv, num_splits, B);
  int rc;
  if (D == 128) {
    static bool once = false;
    if (!once) {
      rc = set_smem_attr(paged_attn_decode_kernel<128>, DecCfg<128>::kSmemBytes, "attn_decode");
      if (rc) return rc;
      once = true;
    }
    (void)launch_pdl(paged_attn_decode_kernel<128>, dim3(grid), dim3(kDecWarps * 32), DecCfg<128>::kSmemBytes, s, 1, 
        (const bf16*)q, q_stride, (const bf16*)k_cache, (const bf16*)v_cache, block_table,
        max_blocks, seq_lens, (bf16*)out, ws_o, ws_ml, num_splits, Hq, Hkv, scale_log2, split_ctr);
    rc = b200_check_launch("b200_paged_attn_decode");
  } else {
    static bool once = false;
    if (!once) {
      rc = set_smem_attr(paged_attn_decode_kernel<64>, DecCfg<64>::kSmemBytes, "attn_decode");
      if (rc) return rc;
      once = true;
    }
    (void)launch_pdl(paged_attn_decode_kernel<64>, dim3(grid), dim3(kDecWarps * 32), DecCfg<64>::kSmemBytes, s, 1, 
        (const bf16*)q, q_stride, (const bf16*)k_cache, (const bf16*)v_cache, block_table,
        max_blocks, seq_lens, (bf16*)out, ws_o, ws_ml, num_splits, Hq, Hkv, scale_log2, split_ctr);
    rc = b200_check_launch("b200_paged_attn_decode");
  }
  return rc;
}

extern "C" int b200_attn_prefill_tile(int Hq, int Hkv) {
  if (Hq <= 0 || Hkv <= 0 || Hq % Hkv) return 0;
  const int G = Hq / Hkv;
  if (G > 8) return 0;
  return G <= 4 ? 32 : 16;
}

namespace b200 {
// attention_prefill_tc.cu: tcgen05 / TMEM kernel (D = 128, GQA group of 4 or 8); returns 1 when the
// configuration is not covered
int paged_attn_prefill_tc(const void* q, int64_t q_stride, int num_tokens, const void* k_cache, const void* v_cache,
                          int num_blocks, const int32_t* block_table, int max_blocks, const int32_t* cu_q,
                          const int32_t* ctx_lens, const int32_t* tile_seq, const int32_t* tile_q0, int num_tiles,
                          void* out, int Hq, int Hkv, int D, float scale, cudaStream_t stream);
}  // namespace b200

extern "C" int b200_paged_attn_prefill(const void* q, int64_t q_stride, int num_tokens, const void* k_cache,
                                       const void* v_cache, int num_blocks, const int32_t* block_table,
                                       int max_blocks, const int32_t* cu_q,
                                       const int32_t* ctx_lens, const int32_t* tile_seq,
                                       const int32_t* tile_q0, int num_tiles, void* out, int Hq,
                                       int Hkv, int D, int block_size, float scale, void* stream) {
  B200_REQUIRE(Hq > 0 && Hkv > 0 && Hq % Hkv == 0, "attn_prefill: bad head counts");
  const int tile = b200_attn_prefill_tile(Hq, Hkv);
  B200_REQUIRE(tile > 0, "attn_prefill: group size %d > 8 unsupported", Hq / Hkv);
  B200_REQUIRE(block_size == kPage, "attn_prefill: block_size must be %d", kPage);
  B200_REQUIRE(D == 128 || D == 64, "attn_prefill: head_dim %d unsupported (64, 128)", D);
  B200_REQUIRE(q_stride % 2 == 0, "attn_prefill: q_stride must be even");
  if (num_tiles <= 0) return B200_OK;
  cudaStream_t s = (cudaStream_t)stream;
  {
    // tensor-core path first (the kernels below are the mma.sync fallback for other head shapes)
    const int rc_tc = paged_attn_prefill_tc(q, q_stride, num_tokens, k_cache, v_cache, num_blocks, block_table,
                                            max_blocks, cu_q, ctx_lens, tile_seq, tile_q0, num_tiles, out, Hq, Hkv,
                                            D, scale, s);
    if (rc_tc <= 0) return rc_tc;
  }
  const float scale_log2 = scale * 1.4426950408889634f;
  const int G = Hq / Hkv, QT = tile / 16;
  dim3 grid(num_tiles, Hkv);
  int rc;
  static int use_v2 = -1;
  if (use_v2 < 0) {
    const char* e = getenv("B200_ATTN_PREFILL_V1");
    use_v2 = (e && e[0] == '1') ? 0 : 1;
  }
  if (D == 128 && use_v2) {
    const int warps = G * QT;
    const int smem_bytes = Pf2Cfg<128>::smem_bytes(warps);
    static bool once2 = false;
    if (!once2) {
      rc = set_smem_attr(paged_attn_prefill_v2_kernel<128>, Pf2Cfg<128>::smem_bytes(8), "attn_prefill_v2");
      if (rc) return rc;
      once2 = true;
    }
    (void)launch_pdl(paged_attn_prefill_v2_kernel<128>, dim3(grid), dim3(32 * warps), (size_t)smem_bytes, s, 1,
                     (const bf16*)q, q_stride, (const bf16*)k_cache, (const bf16*)v_cache, block_table,
                     max_blocks, cu_q, ctx_lens, tile_seq, tile_q0, (bf16*)out, Hq, Hkv, QT, scale_log2);
  } else if (D == 128) {
    static bool once = false;
    if (!once) {
      rc = set_smem_attr(paged_attn_prefill_kernel<128>, PfCfg<128>::kSmemBytes, "attn_prefill");
      if (rc) return rc;
      once = true;
    }
    (void)launch_pdl(paged_attn_prefill_kernel<128>, dim3(grid), dim3(32 * G * QT), PfCfg<128>::kSmemBytes, s, 1, 
        (const bf16*)q, q_stride, (const bf16*)k_cache, (const bf16*)v_cache, block_table,
        max_blocks, cu_q, ctx_lens, tile_seq, tile_q0, (bf16*)out, Hq, Hkv, QT, scale_log2);
  } else {
    static bool once = false;
    if (!once) {
      rc = set_smem_attr(paged_attn_prefill_kernel<64>, PfCfg<64>::kSmemBytes, "attn_prefill");
      if (rc) return rc;
      once = true;
    }
    (void)launch_pdl(paged_attn_prefill_kernel<64>, dim3(grid), dim3(32 * G * QT), PfCfg<64>::kSmemBytes, s, 1, 
        (const bf16*)q, q_stride, (const bf16*)k_cache, (const bf16*)v_cache, block_table,
        max_blocks, cu_q, ctx_lens, tile_seq, tile_q0, (bf16*)out, Hq, Hkv, QT, scale_log2);
  }
  return b200_check_launch("b200_paged_attn_prefill");
}
