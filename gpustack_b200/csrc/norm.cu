// RMSNorm family (memory-bound, one CTA per token row, 16-byte vector loads, values kept in
// registers between the reduction and the normalisation so each byte is read exactly once).
// Semantics follow $VLLM/ir/ops/layernorm.py:10-58: fp32 statistics over the UNROUNDED sum,
// normalised value rounded to bf16, then multiplied by the bf16 weight and rounded again.
#include "common.cuh"
#include "runtime.h"
#include "../../include/b200.h"

#include <stdlib.h>

namespace b200 {

constexpr int kMaxVecPerThread = 4;  // H <= 8 * 4 * 1024

__device__ __forceinline__ float block_sum(float v, float* red) {
  v = warp_sum(v);
  const int warp = threadIdx.x >> 5, lane = threadIdx.x & 31;
  const int nwarps = (blockDim.x + 31) >> 5;
  if (lane == 0) red[warp] = v;
  __syncthreads();
  float t = (lane < nwarps) ? red[lane] : 0.f;
  t = warp_sum(t);
  return t;  // every thread holds the total
}

// MODE 0: y = norm(x)                      (no residual)
// MODE 1: z = x + residual                 (bf16 x)
// MODE 2: z = sum_s partials[s] + residual (fp32 partials; residual may be null)
template <int MODE>
__global__ void __launch_bounds__(512)
rms_norm_kernel(const bf16* __restrict__ x, const float* __restrict__ partials, int splits,
                const bf16* __restrict__ residual, const bf16* __restrict__ w,
                bf16* y, bf16* residual_out, int T, int H, float eps) {
  pdl_launch_dependents();
  pdl_wait();
  __shared__ float red[32];
  const int row = blockIdx.x;
  const int nvec = H >> 3;
  const size_t row_off = (size_t)row * H;
  float v[kMaxVecPerThread][8];
  float ss = 0.f;

#pragma unroll
  for (int i = 0; i < kMaxVecPerThread; ++i) {
    const int vi = threadIdx.x + i * blockDim.x;
    if (vi < nvec) {
      const size_t off = row_off + (size_t)vi * 8;
      if (MODE == 2) {
#pragma unroll
        for (int j = 0; j < 8; ++j) v[i][j] = 0.f;
        for (int s = 0; s < splits; ++s) {
          const float4* p =
              reinterpret_cast<const float4*>(partials + (size_t)s * T * H + off);
          float4 a = ld_cg_f4(p), b = ld_cg_f4(p + 1);   // written by the split-K GEMM just before
          v[i][0] += a.x; v[i][1] += a.y; v[i][2] += a.z; v[i][3] += a.w;
          v[i][4] += b.x; v[i][5] += b.y; v[i][6] += b.z; v[i][7] += b.w;
        }
      } else {
        uint4 xa = ld_cg_v4(x + off);
        v[i][0] = bf16_lo(xa.x); v[i][1] = bf16_hi(xa.x);
        v[i][2] = bf16_lo(xa.y); v[i][3] = bf16_hi(xa.y);
        v[i][4] = bf16_lo(xa.z); v[i][5] = bf16_hi(xa.z);
        v[i][6] = bf16_lo(xa.w); v[i][7] = bf16_hi(xa.w);
      }
      if (MODE != 0 && residual != nullptr) {
        uint4 ra = ld_cg_v4(residual + off);
        v[i][0] += bf16_lo(ra.x); v[i][1] += bf16_hi(ra.x);
        v[i][2] += bf16_lo(ra.y); v[i][3] += bf16_hi(ra.y);
        v[i][4] += bf16_lo(ra.z); v[i][5] += bf16_hi(ra.z);
        v[i][6] += bf16_lo(ra.w); v[i][7] += bf16_hi(ra.w);
      }
      if (MODE != 0) {
        uint4 o;
        o.x = pack_bf16(v[i][0], v[i][1]);
        o.y = pack_bf16(v[i][2], v[i][3]);
        o.z = pack_bf16(v[i][4], v[i][5]);
        o.w = pack_bf16(v[i][6], v[i][7]);
        st_v4(residual_out + off, o);
      }
#pragma unroll
      for (int j = 0; j < 8; ++j) ss += v[i][j] * v[i][j];
    }
  }
  const float total = block_sum(ss, red);
  const float inv = rsqrtf(total / (float)H + eps);

#pragma unroll
  for (int i = 0; i < kMaxVecPerThread; ++i) {
    const int vi = threadIdx.x + i * blockDim.x;
    if (vi < nvec) {
      uint4 wa = ld_v4(w + (size_t)vi * 8);
      float wf[8] = {bf16_lo(wa.x), bf16_hi(wa.x), bf16_lo(wa.y), bf16_hi(wa.y),
                     bf16_lo(wa.z), bf16_hi(wa.z), bf16_lo(wa.w), bf16_hi(wa.w)};
      float o[8];
#pragma unroll
      for (int j = 0; j < 8; ++j) o[j] = round_bf16(v[i][j] * inv) * wf[j];
      uint4 ov;
      ov.x = pack_bf16(o[0], o[1]);
      ov.y = pack_bf16(o[2], o[3]);
      ov.z = pack_bf16(o[4], o[5]);
      ov.w = pack_bf16(o[6], o[7]);
      st_v4(y + row_off + (size_t)vi * 8, ov);
    }
  }
}

static int norm_threads(int H) {
  static int vec_per_thread = -1;   // B200_NORM_VPT: 16-byte vectors per thread (tuning knob)
  if (vec_per_thread < 0) {
    const char* e = getenv("B200_NORM_VPT");
    vec_per_thread = (e && atoi(e) >= 1 && atoi(e) <= kMaxVecPerThread) ? atoi(e) : 2;  // measured: 2 -> 5.9 TB/s (1: 4.4, 4: 5.5)
  }
  int nvec = H / 8;
  int t = ((nvec + vec_per_thread - 1) / vec_per_thread + 31) / 32 * 32;
  if (t > 512) t = 512;
  while (t * kMaxVecPerThread < nvec) t += 32;
  return t < 32 ? 32 : t;
}

static int check_norm_args(int T, int H) {
  B200_REQUIRE(T >= 0 && H > 0, "rms_norm: bad shape T=%d H=%d", T, H);
  B200_REQUIRE(H % 8 == 0, "rms_norm: H=%d must be a multiple of 8", H);
  B200_REQUIRE(H <= 8 * kMaxVecPerThread * 1024, "rms_norm: H=%d too large", H);
  return B200_OK;
}

}  // namespace b200

using namespace b200;

extern "C" int b200_rms_norm(const void* x, const void* w, void* y, int T, int H, float eps,
                             void* stream) {
  int rc = check_norm_args(T, H);
  if (rc) return rc;
  if (T == 0) return B200_OK;
  int threads = norm_threads(H);
  // with 512 threads and nvec up to 2048 the per-thread vector count stays <= 4
  (void)launch_pdl(rms_norm_kernel<0>, dim3(T), dim3(threads), 0, (cudaStream_t)stream, 1, 
      (const bf16*)x, nullptr, 0, nullptr, (const bf16*)w, (bf16*)y, nullptr, T, H, eps);
  return b200_check_launch("b200_rms_norm");
}

extern "C" int b200_fused_add_rms_norm(const void* x, const void* residual, const void* w, void* y,
                                       void* residual_out, int T, int H, float eps, void* stream) {
  int rc = check_norm_args(T, H);
  if (rc) return rc;
  B200_REQUIRE(residual && residual_out, "fused_add_rms_norm: residual pointers required");
  if (T == 0) return B200_OK;
  int threads = norm_threads(H);
  (void)launch_pdl(rms_norm_kernel<1>, dim3(T), dim3(threads), 0, (cudaStream_t)stream, 1, 
      (const bf16*)x, nullptr, 0, (const bf16*)residual, (const bf16*)w, (bf16*)y,
      (bf16*)residual_out, T, H, eps);
  return b200_check_launch("b200_fused_add_rms_norm");
}

extern "C" int b200_fused_add_rms_norm_partials(const float* partials, int splits,
                                                const void* residual, const void* w, void* y,
                                                void* residual_out, int T, int H, float eps,
                                                void* stream) {
  int rc = check_norm_args(T, H);
  if (rc) return rc;
  B200_REQUIRE(splits >= 1 && residual_out, "fused_add_rms_norm_partials: bad arguments");
  if (T == 0) return B200_OK;
  int threads = norm_threads(H);
  (void)launch_pdl(rms_norm_kernel<2>, dim3(T), dim3(threads), 0, (cudaStream_t)stream, 1, 
      nullptr, partials, splits, (const bf16*)residual, (const bf16*)w, (bf16*)y,
      (bf16*)residual_out, T, H, eps);
  return b200_check_launch("b200_fused_add_rms_norm_partials");
}
