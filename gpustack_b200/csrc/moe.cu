// Mixture-of-experts routing for the Mixtral family (BASELINE config "Mixtral-8x7B"; reference:
// MixtralMoE -> FusedMoE in $VLLM/model_executor/models/mixtral.py and
// $VLLM/model_executor/layers/fused_moe/: topk_softmax, moe_align_block_size, moe_sum).
//
// A step's tokens are routed to top-k experts, (token, k) pairs are grouped by expert into a
// permuted row space in which every expert owns a whole number of GEMM tiles (block_x rows), the
// two expert projections run as GROUPED tcgen05 GEMMs over that space (gemm_tcgen05.cu), and the k
// results of a token are summed back.  All bookkeeping stays on the device (no host sync), so the
// whole layer is CUDA-graph capturable: tile counts are read by the GEMM from device memory.
//
// These kernels are index / byte movers: HBM- or latency-bound, no tensor cores.
#include "common.cuh"
#include "runtime.h"
#include "../../include/b200.h"

namespace b200 {

constexpr int kMaxExperts = 64;
constexpr int kMaxTopK = 8;

// ---------------------------------------------------------------------------------------------
// topk_w[t, :] = renormalised top-k of softmax(float(logits[t, :])); ties -> lowest expert index.
__global__ void __launch_bounds__(128)
moe_route_kernel(const bf16* __restrict__ logits, int64_t ld, int32_t* __restrict__ topk_ids,
                 float* __restrict__ topk_w, int T, int E, int K, int renormalize) {
  pdl_launch_dependents();
  pdl_wait();
  const int t = blockIdx.x * blockDim.x + threadIdx.x;
  if (t >= T) return;
  float p[kMaxExperts];
  float mx = -INFINITY;
  for (int e = 0; e < E; ++e) {
    p[e] = __bfloat162float(logits[(size_t)t * ld + e]);
    mx = fmaxf(mx, p[e]);
  }
  float sum = 0.f;
  for (int e = 0; e < E; ++e) {
    p[e] = expf(p[e] - mx);
    sum += p[e];
  }
  const float inv = 1.f / sum;
  float wsum = 0.f;
  float w[kMaxTopK];
  int id[kMaxTopK];
  for (int j = 0; j < K; ++j) {
    int best = 0;
    float bv = -1.f;
    for (int e = 0; e < E; ++e)
      if (p[e] > bv) { bv = p[e]; best = e; }
    id[j] = best;
    w[j] = bv * inv;
    wsum += w[j];
    p[best] = -2.f;
  }
  const float rn = renormalize ? 1.f / wsum : 1.f;
  for (int j = 0; j < K; ++j) {
    topk_ids[(size_t)t * K + j] = id[j];
    topk_w[(size_t)t * K + j] = w[j] * rn;
  }
}

// ---------------------------------------------------------------------------------------------
// Groups the T*K (token, k) pairs by expert.  Row space: expert e owns rows
// [row0[e], row0[e] + pad(count[e], block_x)); tile i of the space belongs to tile_expert[i].
//   pair_row[pair]  = row of the pair                       (for the combine)
//   row_pair[row]   = pair of the row, -1 for padding rows  (for the gather), for rows < capacity
//   row_weight[row] = routing weight of the row, 0 for padding
//   num_tiles[0]    = tiles in use, num_tiles[1] = rows in use
// One CTA: the pair count of a step is at most a few 10^4 and E <= 64.
__global__ void __launch_bounds__(1024)
moe_align_kernel(const int32_t* __restrict__ topk_ids, const float* __restrict__ topk_w, int n_pairs, int E,
                 int block_x, int capacity_rows, int32_t* __restrict__ pair_row, int32_t* __restrict__ row_pair,
                 float* __restrict__ row_weight, int32_t* __restrict__ tile_expert, int32_t* __restrict__ num_tiles) {
  pdl_launch_dependents();
  pdl_wait();
  __shared__ int count[kMaxExperts];
  __shared__ int row0[kMaxExperts + 1];
  __shared__ int cursor[kMaxExperts];
  for (int e = threadIdx.x; e < E; e += blockDim.x) { count[e] = 0; cursor[e] = 0; }
  __syncthreads();
  for (int i = threadIdx.x; i < n_pairs; i += blockDim.x) atomicAdd(&count[topk_ids[i]], 1);
  __syncthreads();
  if (threadIdx.x == 0) {
    int r = 0, tile = 0;
    for (int e = 0; e < E; ++e) {
      row0[e] = r;
      const int tiles = (count[e] + block_x - 1) / block_x;
      for (int i = 0; i < tiles; ++i) tile_expert[tile++] = e;
      r += tiles * block_x;
    }
    row0[E] = r;
    num_tiles[0] = tile;
    num_tiles[1] = r;
  }
  __syncthreads();
  const int used = row0[E];
  for (int r = threadIdx.x; r < capacity_rows; r += blockDim.x) {
    row_pair[r] = -1;
    row_weight[r] = 0.f;
  }
  __syncthreads();
  // rows inside an expert are assigned in pair order WITHIN a warp-stride pass only up to atomics:
  // the order does not matter, every output row depends on its own input row alone.
  for (int i = threadIdx.x; i < n_pairs; i += blockDim.x) {
    const int e = topk_ids[i];
    const int r = row0[e] + atomicAdd(&cursor[e], 1);
    pair_row[i] = r;
    row_pair[r] = i;
    row_weight[r] = topk_w[i];
  }
  (void)used;
}

// ---------------------------------------------------------------------------------------------
// xp[row, :] = x[row_pair[row] / K, :]  (zero for padding rows), rows < num_tiles[1]
__global__ void __launch_bounds__(256)
moe_gather_kernel(const bf16* __restrict__ x, const int32_t* __restrict__ row_pair,
                  const int32_t* __restrict__ num_tiles, bf16* __restrict__ xp, int H, int K, int capacity_rows) {
  pdl_launch_dependents();
  pdl_wait();
  const int rows = min(num_tiles[1], capacity_rows);
  const int vec_per_row = H >> 3;
  const size_t total = (size_t)rows * vec_per_row;
  for (size_t idx = (size_t)blockIdx.x * blockDim.x + threadIdx.x; idx < total;
       idx += (size_t)gridDim.x * blockDim.x) {
    const size_t r = idx / vec_per_row;
    const int vi = (int)(idx % vec_per_row);
    const int pair = row_pair[r];
    uint4 v = make_uint4(0, 0, 0, 0);
    if (pair >= 0) v = ld_v4(x + (size_t)(pair / K) * H + (size_t)vi * 8);
    st_v4(xp + r * (size_t)H + (size_t)vi * 8, v);
  }
}

// ---------------------------------------------------------------------------------------------
// out[t, :] = bf16(sum_j float(yp[pair_row[t*K + j], :]))      (moe_sum)
__global__ void __launch_bounds__(256)
moe_combine_kernel(const bf16* __restrict__ yp, const int32_t* __restrict__ pair_row, bf16* __restrict__ out,
                   int T, int H, int K) {
  pdl_launch_dependents();
  pdl_wait();
  const int vec_per_row = H >> 3;
  const size_t total = (size_t)T * vec_per_row;
  for (size_t idx = (size_t)blockIdx.x * blockDim.x + threadIdx.x; idx < total;
       idx += (size_t)gridDim.x * blockDim.x) {
    const size_t t = idx / vec_per_row;
    const int vi = (int)(idx % vec_per_row);
    float acc[8] = {0.f, 0.f, 0.f, 0.f, 0.f, 0.f, 0.f, 0.f};
    for (int j = 0; j < K; ++j) {
      const int r = pair_row[t * K + j];
      const uint4 a = ld_v4(yp + (size_t)r * H + (size_t)vi * 8);
      acc[0] += bf16_lo(a.x); acc[1] += bf16_hi(a.x);
      acc[2] += bf16_lo(a.y); acc[3] += bf16_hi(a.y);
      acc[4] += bf16_lo(a.z); acc[5] += bf16_hi(a.z);
      acc[6] += bf16_lo(a.w); acc[7] += bf16_hi(a.w);
    }
    uint4 o;
    o.x = pack_bf16(acc[0], acc[1]);
    o.y = pack_bf16(acc[2], acc[3]);
    o.z = pack_bf16(acc[4], acc[5]);
    o.w = pack_bf16(acc[6], acc[7]);
    st_v4(out + t * (size_t)H + (size_t)vi * 8, o);
  }
}

static int moe_grid(size_t work_items, int threads) {
  const int sms = device_info().sm_count > 0 ? device_info().sm_count : 148;
  size_t blocks = (work_items + threads - 1) / threads;
  const size_t cap = (size_t)sms * 8;
  if (blocks > cap) blocks = cap;
  return blocks < 1 ? 1 : (int)blocks;
}

}  // namespace b200

using namespace b200;

extern "C" int b200_moe_route(const void* router_logits, int64_t ld, int32_t* topk_ids, float* topk_w, int T,
                              int E, int K, int renormalize, void* stream) {
  B200_REQUIRE(ld >= E, "moe_route: row stride %lld < E=%d", (long long)ld, E);
  B200_REQUIRE(T >= 0 && E >= 1 && E <= kMaxExperts && K >= 1 && K <= kMaxTopK && K <= E,
               "moe_route: unsupported E=%d K=%d (E <= %d, K <= %d)", E, K, kMaxExperts, kMaxTopK);
  if (T == 0) return B200_OK;
  (void)launch_pdl_cls(kPdlMoe, moe_route_kernel, dim3((T + 127) / 128), dim3(128), 0, (cudaStream_t)stream, 1,
                   (const bf16*)router_logits, ld, topk_ids, topk_w, T, E, K, renormalize);
  return b200_check_launch("b200_moe_route");
}

extern "C" int b200_moe_capacity_rows(int n_pairs, int E, int block_x) {
  // every expert may waste up to block_x - 1 padding rows
  const int worst = n_pairs + E * (block_x - 1);
  return (worst + block_x - 1) / block_x * block_x;
}

extern "C" int b200_moe_align(const int32_t* topk_ids, const float* topk_w, int n_pairs, int E, int block_x,
                              int capacity_rows, int32_t* pair_row, int32_t* row_pair, float* row_weight,
                              int32_t* tile_expert, int32_t* num_tiles, void* stream) {
  B200_REQUIRE(n_pairs >= 0 && E >= 1 && E <= kMaxExperts, "moe_align: unsupported E=%d", E);
  B200_REQUIRE(block_x >= 16 && block_x <= 256 && (block_x & (block_x - 1)) == 0, "moe_align: bad block_x=%d", block_x);
  B200_REQUIRE(capacity_rows >= b200_moe_capacity_rows(n_pairs, E, block_x),
               "moe_align: capacity_rows=%d too small for %d pairs", capacity_rows, n_pairs);
  (void)launch_pdl_cls(kPdlMoe, moe_align_kernel, dim3(1), dim3(1024), 0, (cudaStream_t)stream, 1, topk_ids, topk_w, n_pairs, E,
                   block_x, capacity_rows, pair_row, row_pair, row_weight, tile_expert, num_tiles);
  return b200_check_launch("b200_moe_align");
}

extern "C" int b200_moe_gather(const void* x, const int32_t* row_pair, const int32_t* num_tiles, void* xp, int H,
                               int K, int capacity_rows, void* stream) {
  B200_REQUIRE(H > 0 && H % 8 == 0 && K >= 1, "moe_gather: H=%d must be a multiple of 8", H);
  if (capacity_rows == 0) return B200_OK;
  const size_t items = (size_t)capacity_rows * (H / 8);
  (void)launch_pdl_cls(kPdlMoe, moe_gather_kernel, dim3(moe_grid(items, 256)), dim3(256), 0, (cudaStream_t)stream, 1,
                   (const bf16*)x, row_pair, num_tiles, (bf16*)xp, H, K, capacity_rows);
  return b200_check_launch("b200_moe_gather");
}

extern "C" int b200_moe_combine(const void* yp, const int32_t* pair_row, void* out, int T, int H, int K,
                                void* stream) {
  B200_REQUIRE(H > 0 && H % 8 == 0 && K >= 1, "moe_combine: H=%d must be a multiple of 8", H);
  if (T == 0) return B200_OK;
  const size_t items = (size_t)T * (H / 8);
  (void)launch_pdl_cls(kPdlMoe, moe_combine_kernel, dim3(moe_grid(items, 256)), dim3(256), 0, (cudaStream_t)stream, 1,
                   (const bf16*)yp, pair_row, (bf16*)out, T, H, K);
  return b200_check_launch("b200_moe_combine");
}
