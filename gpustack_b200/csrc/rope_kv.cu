// Neox-style rotary embedding on q/k (in place) fused with the paged KV-cache scatter.
// Reference operators: $VLLM/model_executor/layers/rotary_embedding/base.py:150-215 (rotation,
// cos|sin cache in the model dtype) and $VLLM/_custom_ops.py:2591-2610 (reshape_and_cache_flash).
// One CTA per token; each work item is a 16-byte vector (8 bf16).  Rotation is done in fp32 and
// rounded once (the reference rounds every bf16 op; the difference is < 2 bf16 ulp).
#include "common.cuh"
#include "runtime.h"
#include "../../include/b200.h"

namespace b200 {

__device__ __forceinline__ void unpack8(const uint4& a, float (&f)[8]) {
  f[0] = bf16_lo(a.x); f[1] = bf16_hi(a.x); f[2] = bf16_lo(a.y); f[3] = bf16_hi(a.y);
  f[4] = bf16_lo(a.z); f[5] = bf16_hi(a.z); f[6] = bf16_lo(a.w); f[7] = bf16_hi(a.w);
}
__device__ __forceinline__ uint4 pack8(const float (&f)[8]) {
  uint4 o;
  o.x = pack_bf16(f[0], f[1]); o.y = pack_bf16(f[2], f[3]);
  o.z = pack_bf16(f[4], f[5]); o.w = pack_bf16(f[6], f[7]);
  return o;
}

// PARTIALS = true: the QKV projection arrives as `splits` fp32 split-K partial sums
// [splits][T][(Hq+2Hkv)*D]; they are summed, rounded to bf16 (what the un-split GEMM would have
// stored) and then rotated -- the split-K reduction costs no extra pass.
template <bool PARTIALS>
__device__ __forceinline__ void load8(const bf16* qkv, const float* partials, int splits, size_t plane,
                                      size_t off, float (&f)[8], const bf16* bias = nullptr, int col = 0) {
  if constexpr (!PARTIALS) {
    unpack8(ld_cg_v4(qkv + off), f);
    return;
  }
#pragma unroll
  for (int j = 0; j < 8; ++j) f[j] = 0.f;
  for (int s = 0; s < splits; ++s) {
    const float4* p = reinterpret_cast<const float4*>(partials + (size_t)s * plane + off);
    const float4 a = ld_cg_f4(p), b = ld_cg_f4(p + 1);
    f[0] += a.x; f[1] += a.y; f[2] += a.z; f[3] += a.w;
    f[4] += b.x; f[5] += b.y; f[6] += b.z; f[7] += b.w;
  }
  if (bias != nullptr) {  // QKV bias (Qwen2): added in fp32, the sum is rounded once
    float bf[8];
    unpack8(ld_v4(bias + col), bf);
#pragma unroll
    for (int j = 0; j < 8; ++j) f[j] += bf[j];
  }
#pragma unroll
  for (int j = 0; j < 8; ++j) f[j] = round_bf16(f[j]);
}

template <bool PARTIALS>
__global__ void __launch_bounds__(256)
rope_kv_write_kernel(bf16* __restrict__ qkv, const float* __restrict__ partials, int splits, int T,
                     const bf16* __restrict__ bias, const int32_t* __restrict__ positions,
                     const bf16* __restrict__ cos_sin, bf16* __restrict__ k_cache,
                     bf16* __restrict__ v_cache, const int32_t* __restrict__ slot_mapping, int Hq,
                     int Hkv, int D, int block_size) {
  pdl_launch_dependents();
  pdl_wait();
  const int t = blockIdx.x;
  const int half = D >> 1;
  const int vec_per_half = half >> 3;           // 16-byte vectors in half a head
  const int rope_items = (Hq + Hkv) * vec_per_half;
  const int v_items = Hkv * (D >> 3);
  const int pos = positions[t];
  const int slot = slot_mapping[t];
  const size_t row = (size_t)t * (size_t)(Hq + 2 * Hkv) * D;
  const size_t plane = (size_t)T * (size_t)(Hq + 2 * Hkv) * D;
  const bf16* cs = cos_sin + (size_t)pos * D;
  const int blk = slot >= 0 ? slot / block_size : 0;
  const int off = slot >= 0 ? slot % block_size : 0;

  for (int it = threadIdx.x; it < rope_items + v_items; it += blockDim.x) {
    if (it < rope_items) {
      const int head = it / vec_per_half;       // 0..Hq-1 = q heads, then k heads
      const int vi = it % vec_per_half;
      bf16* p1 = qkv + row + (size_t)head * D + vi * 8;
      bf16* p2 = p1 + half;
      float x1[8], x2[8], c[8], s[8], o1[8], o2[8];
      load8<PARTIALS>(qkv, partials, splits, plane, (size_t)(p1 - qkv), x1, bias, head * D + vi * 8);
      load8<PARTIALS>(qkv, partials, splits, plane, (size_t)(p2 - qkv), x2, bias, head * D + half + vi * 8);
      unpack8(ld_v4(cs + vi * 8), c);
      unpack8(ld_v4(cs + half + vi * 8), s);
#pragma unroll
      for (int j = 0; j < 8; ++j) {
        o1[j] = x1[j] * c[j] - x2[j] * s[j];
        o2[j] = x2[j] * c[j] + x1[j] * s[j];
      }
      const uint4 r1 = pack8(o1), r2 = pack8(o2);
      st_v4(p1, r1);
      st_v4(p2, r2);
      if (head >= Hq && slot >= 0) {
        const int kh = head - Hq;
        bf16* kc = k_cache + (((size_t)blk * Hkv + kh) * block_size + off) * D + vi * 8;
        st_v4(kc, r1);
        st_v4(kc + half, r2);
      }
    } else if (slot >= 0 || PARTIALS) {
      const int j = it - rope_items;
      const int vh = j / (D >> 3);
      const int vi = j % (D >> 3);
      const size_t voff = row + (size_t)(Hq + Hkv + vh) * D + vi * 8;
      bf16* dst = v_cache + (((size_t)blk * Hkv + vh) * block_size + off) * D + vi * 8;
      if constexpr (!PARTIALS) {
        st_v4(dst, ld_v4(qkv + voff));
      } else {
        float vf[8];
        load8<true>(qkv, partials, splits, plane, voff, vf, bias, (Hq + Hkv + vh) * D + vi * 8);
        const uint4 vv = pack8(vf);
        if (slot >= 0) st_v4(dst, vv);
        st_v4(qkv + voff, vv);
      }
    }
  }
}

}  // namespace b200

using namespace b200;

static int rope_launch(void* qkv, const float* partials, int splits, const void* bias, const int32_t* positions,
                       const void* cos_sin, void* k_cache, void* v_cache, const int32_t* slot_mapping,
                       int T, int Hq, int Hkv, int D, int block_size, void* stream, const char* what) {
  B200_REQUIRE(T >= 0 && Hq > 0 && Hkv > 0 && block_size > 0, "%s: bad shape", what);
  B200_REQUIRE(D % 16 == 0 && D <= 256, "%s: head_dim %d must be a multiple of 16", what, D);
  if (T == 0) return B200_OK;
  if (partials != nullptr)
    (void)launch_pdl(rope_kv_write_kernel<true>, dim3(T), dim3(256), 0, (cudaStream_t)stream, 1, (bf16*)qkv,
                     partials, splits, T, (const bf16*)bias, positions, (const bf16*)cos_sin, (bf16*)k_cache, (bf16*)v_cache,
                     slot_mapping, Hq, Hkv, D, block_size);
  else
    (void)launch_pdl(rope_kv_write_kernel<false>, dim3(T), dim3(256), 0, (cudaStream_t)stream, 1, (bf16*)qkv,
                     partials, splits, T, (const bf16*)bias, positions, (const bf16*)cos_sin, (bf16*)k_cache, (bf16*)v_cache,
                     slot_mapping, Hq, Hkv, D, block_size);
  return b200_check_launch(what);
}

extern "C" int b200_rope_kv_write(void* qkv, const int32_t* positions, const void* cos_sin,
                                  void* k_cache, void* v_cache, const int32_t* slot_mapping, int T,
                                  int Hq, int Hkv, int D, int block_size, void* stream) {
  return rope_launch(qkv, nullptr, 0, nullptr, positions, cos_sin, k_cache, v_cache, slot_mapping, T, Hq, Hkv, D,
                     block_size, stream, "b200_rope_kv_write");
}

extern "C" int b200_rope_kv_write_partials(const float* partials, int splits, const void* qkv_bias,
                                           void* qkv_out, const int32_t* positions, const void* cos_sin, void* k_cache,
                                           void* v_cache, const int32_t* slot_mapping, int T, int Hq,
                                           int Hkv, int D, int block_size, void* stream) {
  B200_REQUIRE(partials != nullptr && splits >= 1, "rope_kv_write_partials: partials required");
  return rope_launch(qkv_out, partials, splits, qkv_bias, positions, cos_sin, k_cache, v_cache, slot_mapping, T, Hq,
                     Hkv, D, block_size, stream, "b200_rope_kv_write_partials");
}
