// Dynamic per-token fp8 (e4m3) activation quantisation for the W8A8 projections
// (reference: QuantFP8.forward_native, $VLLM/model_executor/layers/quantization/input_quant_fp8.py:188-235,
// the semantics behind ops.scaled_fp8_quant(use_per_token_if_dynamic=True)):
//
//     scale[t] = max(max_k |x[t,k]| / 448, 1 / (448 * 512))
//     q[t,k]   = e4m3_rne_sat(float(x[t,k]) * (1 / scale[t]))
//
// Memory-bound: one CTA per token row, 16-byte loads, 8-byte stores.  The row is read twice (absmax,
// then convert); the second read is an L1/L2 hit (a row is 8-57 KB), so HBM sees 2 + 1 bytes per
// element.  Division and reciprocal are the IEEE-rounded ones, so the result is bit-identical to
// the PyTorch restatement in oracle/llama_oracle.py.
#include "common.cuh"
#include "runtime.h"
#include "../../include/b200.h"

#include <cuda_fp8.h>

namespace b200 {

constexpr float kFp8Max = 448.f;
constexpr float kFp8MinScale = 1.f / (448.f * 512.f);

__device__ __forceinline__ uint32_t pack_e4m3x4(float a, float b, float c, float d) {
  const uint32_t lo = __nv_cvt_float2_to_fp8x2(make_float2(a, b), __NV_SATFINITE, __NV_E4M3);
  const uint32_t hi = __nv_cvt_float2_to_fp8x2(make_float2(c, d), __NV_SATFINITE, __NV_E4M3);
  return lo | (hi << 16);
}

__global__ void __launch_bounds__(256)
quant_fp8_per_token_kernel(const bf16* __restrict__ x, int64_t ldx, uint8_t* __restrict__ q, int64_t ldq,
                           float* __restrict__ scales, int T, int K) {
  pdl_launch_dependents();
  pdl_wait();
  __shared__ float red[8];
  __shared__ float s_inv;
  const int row = blockIdx.x;
  const bf16* xr = x + (size_t)row * ldx;
  uint8_t* qr = q + (size_t)row * ldq;
  const int nvec = K >> 3;

  float amax = 0.f;
  for (int vi = threadIdx.x; vi < nvec; vi += blockDim.x) {
    const uint4 a = ld_v4(xr + (size_t)vi * 8);
    const uint32_t w[4] = {a.x, a.y, a.z, a.w};
#pragma unroll
    for (int j = 0; j < 4; ++j) amax = fmaxf(amax, fmaxf(fabsf(bf16_lo(w[j])), fabsf(bf16_hi(w[j]))));
  }
  amax = warp_max(amax);
  const int warp = threadIdx.x >> 5, lane = threadIdx.x & 31;
  if (lane == 0) red[warp] = amax;
  __syncthreads();
  if (warp == 0) {
    float t = lane < (int)(blockDim.x >> 5) ? red[lane] : 0.f;
    t = warp_max(t);
    if (lane == 0) {
      const float scale = fmaxf(__fdiv_rn(t, kFp8Max), kFp8MinScale);
      scales[row] = scale;
      s_inv = __frcp_rn(scale);
    }
  }
  __syncthreads();
  const float inv = s_inv;
  for (int vi = threadIdx.x; vi < nvec; vi += blockDim.x) {
    const uint4 a = ld_v4(xr + (size_t)vi * 8);
    uint2 o;
    o.x = pack_e4m3x4(__fmul_rn(bf16_lo(a.x), inv), __fmul_rn(bf16_hi(a.x), inv),
                      __fmul_rn(bf16_lo(a.y), inv), __fmul_rn(bf16_hi(a.y), inv));
    o.y = pack_e4m3x4(__fmul_rn(bf16_lo(a.z), inv), __fmul_rn(bf16_hi(a.z), inv),
                      __fmul_rn(bf16_lo(a.w), inv), __fmul_rn(bf16_hi(a.w), inv));
    *reinterpret_cast<uint2*>(qr + (size_t)vi * 8) = o;
  }
}

}  // namespace b200

using namespace b200;

extern "C" int b200_quant_fp8_per_token(const void* x, int64_t ldx, void* q, int64_t ldq, float* scales,
                                        int T, int K, void* stream) {
  B200_REQUIRE(T >= 0 && K > 0 && K % 8 == 0, "quant_fp8: K=%d must be a multiple of 8", K);
  B200_REQUIRE(ldx >= K && ldq >= K && ldx % 8 == 0 && ldq % 8 == 0, "quant_fp8: bad leading dimensions");
  B200_REQUIRE(x && q && scales, "quant_fp8: null pointer");
  if (T == 0) return B200_OK;
  (void)launch_pdl(quant_fp8_per_token_kernel, dim3(T), dim3(256), 0, (cudaStream_t)stream, 1,
                   (const bf16*)x, ldx, (uint8_t*)q, ldq, scales, T, K);
  return b200_check_launch("b200_quant_fp8_per_token");
}
