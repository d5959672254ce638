// Element-wise and gather kernels of the decoder path: SiLU*mul, embedding gather, row gather,
// split-K reduction and greedy argmax.  All are HBM-bound: 16-byte vector accesses, grid sized in
// multiples of the SM count, no shared-memory staging (no reuse to exploit).
#include "common.cuh"
#include "runtime.h"
#include "../../include/b200.h"

#include <float.h>

namespace b200 {

// ---------------------------------------------------------------------------------------------
// out[t,i] = bf16(bf16(silu(g)) * u), g = x[t,i], u = x[t,I+i]   ($VLLM activation.py:133-143)
__global__ void __launch_bounds__(256)
silu_and_mul_kernel(const bf16* __restrict__ x, bf16* __restrict__ out, int T, int I) {
  pdl_launch_dependents();
  pdl_wait();
  const int vec_per_row = I >> 3;
  const size_t total = (size_t)T * vec_per_row;
  for (size_t idx = (size_t)blockIdx.x * blockDim.x + threadIdx.x; idx < total;
       idx += (size_t)gridDim.x * blockDim.x) {
    const size_t t = idx / vec_per_row;
    const int vi = (int)(idx % vec_per_row);
    const bf16* g = x + t * (size_t)(2 * I) + (size_t)vi * 8;
    const uint4 ga = ld_cg_v4(g);
    const uint4 ua = ld_cg_v4(g + I);
    const uint32_t gw[4] = {ga.x, ga.y, ga.z, ga.w};
    const uint32_t uw[4] = {ua.x, ua.y, ua.z, ua.w};
    uint32_t ow[4];
#pragma unroll
    for (int j = 0; j < 4; ++j) {
      const float g0 = bf16_lo(gw[j]), g1 = bf16_hi(gw[j]);
      const float s0 = round_bf16(g0 / (1.f + __expf(-g0)));
      const float s1 = round_bf16(g1 / (1.f + __expf(-g1)));
      ow[j] = pack_bf16(s0 * bf16_lo(uw[j]), s1 * bf16_hi(uw[j]));
    }
    st_v4(out + t * (size_t)I + (size_t)vi * 8, make_uint4(ow[0], ow[1], ow[2], ow[3]));
  }
}

// ---------------------------------------------------------------------------------------------
// out[t] = table[ids[t]-vocab_start] or 0 outside [vocab_start, vocab_end)
__global__ void __launch_bounds__(256)
embedding_kernel(const int32_t* __restrict__ ids, const bf16* __restrict__ table,
                 bf16* __restrict__ out, int T, int H, int vocab_start, int vocab_end) {
  pdl_launch_dependents();
  pdl_wait();
  const int vec_per_row = H >> 3;
  const size_t total = (size_t)T * vec_per_row;
  for (size_t idx = (size_t)blockIdx.x * blockDim.x + threadIdx.x; idx < total;
       idx += (size_t)gridDim.x * blockDim.x) {
    const size_t t = idx / vec_per_row;
    const int vi = (int)(idx % vec_per_row);
    const int id = ids[t];
    uint4 v = make_uint4(0, 0, 0, 0);
    if (id >= vocab_start && id < vocab_end)
      v = ld_v4(table + (size_t)(id - vocab_start) * H + (size_t)vi * 8);
    st_v4(out + t * (size_t)H + (size_t)vi * 8, v);
  }
}

__global__ void __launch_bounds__(256)
gather_rows_kernel(const bf16* __restrict__ x, const int32_t* __restrict__ rows,
                   bf16* __restrict__ out, int n, int H) {
  pdl_launch_dependents();
  pdl_wait();
  const int vec_per_row = H >> 3;
  const size_t total = (size_t)n * vec_per_row;
  for (size_t idx = (size_t)blockIdx.x * blockDim.x + threadIdx.x; idx < total;
       idx += (size_t)gridDim.x * blockDim.x) {
    const size_t i = idx / vec_per_row;
    const int vi = (int)(idx % vec_per_row);
    st_v4(out + i * (size_t)H + (size_t)vi * 8,
          ld_v4(x + (size_t)rows[i] * H + (size_t)vi * 8));
  }
}

// ---------------------------------------------------------------------------------------------
// Y[m,n] = bf16(sum_s P[s][m][n]);  P fp32 [S][M][N] dense
__global__ void __launch_bounds__(256)
splitk_reduce_kernel(const float* __restrict__ P, int S, bf16* __restrict__ Y, int M, int N,
                     int64_t ldy) {
  pdl_launch_dependents();
  pdl_wait();
  const int vec_per_row = N >> 2;
  const size_t total = (size_t)M * vec_per_row;
  const size_t plane = (size_t)M * N;
  for (size_t idx = (size_t)blockIdx.x * blockDim.x + threadIdx.x; idx < total;
       idx += (size_t)gridDim.x * blockDim.x) {
    const size_t m = idx / vec_per_row;
    const int vi = (int)(idx % vec_per_row);
    float4 acc = make_float4(0.f, 0.f, 0.f, 0.f);
    for (int s = 0; s < S; ++s) {
      const float4 p = ld_cg_f4(reinterpret_cast<const float4*>(P + s * plane + m * N) + vi);
      acc.x += p.x; acc.y += p.y; acc.z += p.z; acc.w += p.w;
    }
    uint2 o;
    o.x = pack_bf16(acc.x, acc.y);
    o.y = pack_bf16(acc.z, acc.w);
    *reinterpret_cast<uint2*>(Y + m * ldy + (size_t)vi * 4) = o;
  }
}

// ---------------------------------------------------------------------------------------------
// Sampler: one CTA per row, ONE pass over the fp32 logits.
//   temperature 0 (or temps == nullptr): greedy -- lowest index of the row maximum;
//   temperature T > 0: exact softmax(logits / T) sampling by the Gumbel-max trick,
//       token = argmax_v ( logit_v / T - log(-log(u_v)) ),  u_v = hash(row seed, GLOBAL vocab index),
//     so sampling costs what greedy costs, stays inside the captured decode graph, and composes with
//     vocabulary-parallel logits: every rank perturbs its shard with noise keyed by the global index and
//     the cross-rank (max, argmax) exchange that serves greedy picks the sampled token.
//     (Reference: $VLLM/v1/sample/sampler.py draws from softmax(logits / T); nucleus (top_p < 1) rows are
//     re-sampled by the host path.)  The row seed changes every step (request seed x position).
__device__ __forceinline__ uint32_t fmix32(uint32_t h) {
  h ^= h >> 16; h *= 0x85ebca6bu; h ^= h >> 13; h *= 0xc2b2ae35u; h ^= h >> 16;
  return h;
}
__device__ __forceinline__ float gumbel_key(float logit, float inv_t, uint32_t seed, uint32_t gidx) {
  const uint32_t h = fmix32(fmix32(gidx * 0x9e3779b1u + seed) ^ (seed * 0x85ebca6bu + 0x632be5abu));
  const float u = ((float)(h >> 8) + 0.5f) * (1.0f / 16777216.0f);      // (0, 1)
  return logit * inv_t - __logf(-__logf(u));
}

__global__ void __launch_bounds__(1024)
argmax_f32_kernel(const float* __restrict__ logits, int64_t ld, int32_t* __restrict__ out_idx,
                  float* __restrict__ out_val, int V, const float* __restrict__ temps,
                  const uint32_t* __restrict__ seeds, int vocab_offset) {
  pdl_launch_dependents();
  pdl_wait();
  __shared__ float s_val[32];
  __shared__ int s_idx[32];
  const float* row = logits + (size_t)blockIdx.x * ld;
  const float temp = temps ? temps[blockIdx.x] : 0.f;
  const bool sample = temp > 0.f;
  const float inv_t = sample ? 1.f / temp : 1.f;
  const uint32_t seed = (sample && seeds) ? seeds[blockIdx.x] : 0u;
  float best = -FLT_MAX;
  int best_i = 0x7fffffff;
  const int nvec = ((reinterpret_cast<uintptr_t>(row) & 15) == 0) ? (V >> 2) : 0;
  for (int vi = threadIdx.x; vi < nvec; vi += blockDim.x) {
    const float4 v = ld_cg_f4(reinterpret_cast<const float4*>(row) + vi);
    float f[4] = {v.x, v.y, v.z, v.w};
#pragma unroll
    for (int j = 0; j < 4; ++j) {
      const int i = vi * 4 + j;
      if (sample) f[j] = gumbel_key(f[j], inv_t, seed, (uint32_t)(vocab_offset + i));
      if (f[j] > best || (f[j] == best && i < best_i)) { best = f[j]; best_i = i; }
    }
  }
  for (int i = nvec * 4 + threadIdx.x; i < V; i += blockDim.x) {
    float f = row[i];
    if (sample) f = gumbel_key(f, inv_t, seed, (uint32_t)(vocab_offset + i));
    if (f > best || (f == best && i < best_i)) { best = f; best_i = i; }
  }
#pragma unroll
  for (int o = 16; o > 0; o >>= 1) {
    const float ov = __shfl_xor_sync(0xffffffffu, best, o);
    const int oi = __shfl_xor_sync(0xffffffffu, best_i, o);
    if (ov > best || (ov == best && oi < best_i)) { best = ov; best_i = oi; }
  }
  const int warp = threadIdx.x >> 5, lane = threadIdx.x & 31;
  if (lane == 0) { s_val[warp] = best; s_idx[warp] = best_i; }
  __syncthreads();
  if (warp == 0) {
    const int nwarps = blockDim.x >> 5;
    best = lane < nwarps ? s_val[lane] : -FLT_MAX;
    best_i = lane < nwarps ? s_idx[lane] : 0x7fffffff;
#pragma unroll
    for (int o = 16; o > 0; o >>= 1) {
      const float ov = __shfl_xor_sync(0xffffffffu, best, o);
      const int oi = __shfl_xor_sync(0xffffffffu, best_i, o);
      if (ov > best || (ov == best && oi < best_i)) { best = ov; best_i = oi; }
    }
    if (lane == 0) {
      out_idx[blockIdx.x] = best_i;
      if (out_val) out_val[blockIdx.x] = best;
    }
  }
}

static int ew_grid(size_t work_items, int threads) {
  const int sms = device_info().sm_count > 0 ? device_info().sm_count : 148;
  size_t blocks = (work_items + threads - 1) / threads;
  const size_t cap = (size_t)sms * 8;  // 8 resident CTAs of 256 threads per SM
  if (blocks > cap) blocks = cap;
  return blocks < 1 ? 1 : (int)blocks;
}

}  // namespace b200

using namespace b200;

extern "C" int b200_silu_and_mul(const void* x, void* out, int T, int I, void* stream) {
  B200_REQUIRE(T >= 0 && I > 0 && I % 8 == 0, "silu_and_mul: I=%d must be a multiple of 8", I);
  if (T == 0) return B200_OK;
  const size_t items = (size_t)T * (I / 8);
  (void)launch_pdl(silu_and_mul_kernel, dim3(ew_grid(items, 256)), dim3(256), 0, (cudaStream_t)stream, 1,
      (const bf16*)x, (bf16*)out, T, I);
  return b200_check_launch("b200_silu_and_mul");
}

extern "C" int b200_embedding(const int32_t* ids, const void* table, void* out, int T, int H,
                              int vocab_start, int vocab_end, void* stream) {
  B200_REQUIRE(T >= 0 && H > 0 && H % 8 == 0, "embedding: H=%d must be a multiple of 8", H);
  if (T == 0) return B200_OK;
  const size_t items = (size_t)T * (H / 8);
  (void)launch_pdl(embedding_kernel, dim3(ew_grid(items, 256)), dim3(256), 0, (cudaStream_t)stream, 1,
      ids, (const bf16*)table, (bf16*)out, T, H, vocab_start, vocab_end);
  return b200_check_launch("b200_embedding");
}

extern "C" int b200_gather_rows(const void* x, const int32_t* rows, void* out, int n, int H,
                                void* stream) {
  B200_REQUIRE(n >= 0 && H > 0 && H % 8 == 0, "gather_rows: H=%d must be a multiple of 8", H);
  if (n == 0) return B200_OK;
  const size_t items = (size_t)n * (H / 8);
  (void)launch_pdl(gather_rows_kernel, dim3(ew_grid(items, 256)), dim3(256), 0, (cudaStream_t)stream, 1,
      (const bf16*)x, rows, (bf16*)out, n, H);
  return b200_check_launch("b200_gather_rows");
}

extern "C" int b200_splitk_reduce(const float* partials, int splits, void* Y, int M, int N,
                                  int64_t ldy, void* stream) {
  B200_REQUIRE(splits >= 1 && M >= 0 && N > 0 && N % 4 == 0 && ldy % 4 == 0,
               "splitk_reduce: N=%d and ldy must be multiples of 4", N);
  if (M == 0) return B200_OK;
  const size_t items = (size_t)M * (N / 4);
  (void)launch_pdl(splitk_reduce_kernel, dim3(ew_grid(items, 256)), dim3(256), 0, (cudaStream_t)stream, 1,
      partials, splits, (bf16*)Y, M, N, ldy);
  return b200_check_launch("b200_splitk_reduce");
}

extern "C" int b200_argmax_f32(const float* logits, int64_t ld, int32_t* out_idx, float* out_val,
                               int B, int V, void* stream) {
  B200_REQUIRE(B >= 0 && V > 0, "argmax: bad shape");
  if (B == 0) return B200_OK;
  (void)launch_pdl_cls(kPdlSampler, argmax_f32_kernel, dim3(B), dim3(1024), 0, (cudaStream_t)stream, 1, logits, ld, out_idx,
                       out_val, V, (const float*)nullptr, (const uint32_t*)nullptr, 0);
  return b200_check_launch("b200_argmax_f32");
}

// Greedy (temps[b] == 0) or Gumbel-max temperature sampling (temps[b] > 0, seed seeds[b]) of row b in one
// pass; vocab_offset = global index of column 0 (vocabulary-parallel shards); out_val = the winning key.
extern "C" int b200_sample_f32(const float* logits, int64_t ld, const float* temps, const uint32_t* seeds,
                               int vocab_offset, int32_t* out_idx, float* out_val, int B, int V, void* stream) {
  B200_REQUIRE(B >= 0 && V > 0 && vocab_offset >= 0, "sample: bad shape");
  if (B == 0) return B200_OK;
  (void)launch_pdl_cls(kPdlSampler, argmax_f32_kernel, dim3(B), dim3(1024), 0, (cudaStream_t)stream, 1, logits, ld, out_idx,
                       out_val, V, temps, seeds, vocab_offset);
  return b200_check_launch("b200_sample_f32");
}
