// Tensor-parallel all-reduce for the row-parallel linears (O and down projections), fused with the
// residual add + RMSNorm that always follows them in the decoder layer.
// Reference operator: CudaCommunicator.all_reduce ($VLLM/distributed/device_communicators/
// cuda_communicator.py:241-290; the "custom all-reduce" one-shot path of custom_all_reduce.py:50-58)
// followed by fused_add_rms_norm ($VLLM/model_executor/layers/layernorm.py:104-150).
//
// One-shot over NVLink peer memory: every rank's row-parallel GEMM writes its partial [T,H] bf16
// straight into a peer-mapped staging buffer; after a flag barrier (release/acquire at system
// scope, one flag word per (CTA, peer)) each rank reads the same rows from all peers, sums them in
// fp32 in rank order (bit-identical on every rank), adds the residual and normalises -- one pass,
// no second exchange.  NVSwitch gives every peer full bandwidth, so for decode-sized messages
// (<= 2 MiB) latency, not link count, is what matters.  Staging buffers are double-buffered by the
// caller (O-proj / down-proj alternate), which makes the start barrier sufficient.
#include "common.cuh"
#include "runtime.h"
#include "../../include/b200.h"

namespace b200 {

constexpr int kMaxWorld = 8;
constexpr int kArMaxVec = 4;

struct PeerPtrs {
  const bf16* data[kMaxWorld];
  uint32_t* flags[kMaxWorld];
};

__device__ __forceinline__ void st_release_sys(uint32_t* p, uint32_t v) {
  asm volatile("st.release.sys.global.u32 [%0], %1;" ::"l"(p), "r"(v) : "memory");
}
__device__ __forceinline__ uint32_t ld_acquire_sys(const uint32_t* p) {
  uint32_t v;
  asm volatile("ld.acquire.sys.global.u32 %0, [%1];" : "=r"(v) : "l"(p) : "memory");
  return v;
}
__device__ __forceinline__ uint4 ld_peer_v4(const void* p) {
  uint4 r;
  asm volatile("ld.global.cg.v4.u32 {%0,%1,%2,%3}, [%4];"
               : "=r"(r.x), "=r"(r.y), "=r"(r.z), "=r"(r.w)
               : "l"(p)
               : "memory");
  return r;
}

__device__ __forceinline__ float ar_block_sum(float v, float* red) {
  v = warp_sum(v);
  const int warp = threadIdx.x >> 5, lane = threadIdx.x & 31;
  const int nwarps = (blockDim.x + 31) >> 5;
  __syncthreads();  // red[] may still be read by the previous row
  if (lane == 0) red[warp] = v;
  __syncthreads();
  float t = (lane < nwarps) ? red[lane] : 0.f;
  return warp_sum(t);
}

__device__ __forceinline__ void st_release_gpu(uint32_t* p, uint32_t v) {
  asm volatile("st.release.gpu.global.u32 [%0], %1;" ::"l"(p), "r"(v) : "memory");
}
__device__ __forceinline__ uint32_t ld_acquire_gpu(const uint32_t* p) {
  uint32_t v;
  asm volatile("ld.acquire.gpu.global.u32 %0, [%1];" : "=r"(v) : "l"(p) : "memory");
  return v;
}

// flags layout on every rank (uint32 words):
//   [0, world)            arrival words, written by the peers (word r = rank r's epoch)
//   [world]               local "go" word, written by CTA 0 once every peer has arrived
//   [world+1, +grid)      private per-CTA call counters (every launch uses the SAME fixed grid, so
//                         all of them equal the call epoch; kept on the device so that CUDA-graph
//                         replays need no host-side epoch)
// Only CTA 0 talks to the other GPUs (one system-scope release/acquire handshake per call); the
// other CTAs wait on the local go word at GPU scope.  The release/acquire chain
// peer GEMM -> peer signal -> CTA 0 acquire.sys -> go release.gpu -> CTA acquire.gpu
// makes the peers' partial sums visible to every CTA.
__global__ void __launch_bounds__(512)
allreduce_add_rms_norm_kernel(PeerPtrs peers, int world, int rank, int flag_blocks,
                              const bf16* __restrict__ residual, const bf16* __restrict__ w,
                              bf16* y, bf16* residual_out, int T, int H, float eps) {
  pdl_launch_dependents();
  pdl_wait();
  __shared__ float red[32];
  uint32_t* my_flags = peers.flags[rank];
  uint32_t* go = my_flags + world;
  uint32_t epoch = 0;
  if (threadIdx.x == 0) {
    uint32_t* ctr = my_flags + world + 1 + blockIdx.x;
    epoch = *ctr + 1;
    *ctr = epoch;
  }
  if (blockIdx.x == 0) {
    __shared__ uint32_t s_epoch;
    if (threadIdx.x == 0) s_epoch = epoch;
    __syncthreads();
    epoch = s_epoch;
    if (threadIdx.x < world) {
      if ((int)threadIdx.x != rank) st_release_sys(peers.flags[threadIdx.x] + rank, epoch);
      const uint32_t* mine = my_flags + threadIdx.x;
      uint64_t t0 = 0;
      uint32_t spins = 0;
      while ((int)threadIdx.x != rank && (int32_t)(ld_acquire_sys(mine) - epoch) < 0) {
        if ((++spins & 1023u) == 0) {
          uint64_t now;
          asm volatile("mov.u64 %0, %%globaltimer;" : "=l"(now));
          if (t0 == 0) t0 = now;
          else if (now - t0 > 10000000000ull) __trap();  // a peer never arrived: fail loudly
        }
      }
    }
    __syncthreads();
    if (threadIdx.x == 0) st_release_gpu(go, epoch);
  } else {
    if (threadIdx.x == 0) {
      uint64_t t0 = 0;
      uint32_t spins = 0;
      while ((int32_t)(ld_acquire_gpu(go) - epoch) < 0) {
        if ((++spins & 4095u) == 0) {
          uint64_t now;
          asm volatile("mov.u64 %0, %%globaltimer;" : "=l"(now));
          if (t0 == 0) t0 = now;
          else if (now - t0 > 10000000000ull) __trap();
        }
      }
    }
    __syncthreads();
  }

  const int nvec = H >> 3;
  for (int row = blockIdx.x; row < T; row += gridDim.x) {
    const size_t row_off = (size_t)row * H;
    float v[kArMaxVec][8];
    float ss = 0.f;
#pragma unroll
    for (int i = 0; i < kArMaxVec; ++i) {
      const int vi = threadIdx.x + i * blockDim.x;
      if (vi < nvec) {
        const size_t off = row_off + (size_t)vi * 8;
#pragma unroll
        for (int j = 0; j < 8; ++j) v[i][j] = 0.f;
        for (int p = 0; p < world; ++p) {
          const uint4 a = ld_peer_v4(peers.data[p] + off);
          v[i][0] += bf16_lo(a.x); v[i][1] += bf16_hi(a.x);
          v[i][2] += bf16_lo(a.y); v[i][3] += bf16_hi(a.y);
          v[i][4] += bf16_lo(a.z); v[i][5] += bf16_hi(a.z);
          v[i][6] += bf16_lo(a.w); v[i][7] += bf16_hi(a.w);
        }
        if (w == nullptr) {
          uint4 o;
          o.x = pack_bf16(v[i][0], v[i][1]); o.y = pack_bf16(v[i][2], v[i][3]);
          o.z = pack_bf16(v[i][4], v[i][5]); o.w = pack_bf16(v[i][6], v[i][7]);
          st_v4(y + off, o);
          continue;
        }
        if (residual != nullptr) {
          const uint4 ra = ld_v4(residual + off);
          v[i][0] += bf16_lo(ra.x); v[i][1] += bf16_hi(ra.x);
          v[i][2] += bf16_lo(ra.y); v[i][3] += bf16_hi(ra.y);
          v[i][4] += bf16_lo(ra.z); v[i][5] += bf16_hi(ra.z);
          v[i][6] += bf16_lo(ra.w); v[i][7] += bf16_hi(ra.w);
        }
        uint4 o;
        o.x = pack_bf16(v[i][0], v[i][1]); o.y = pack_bf16(v[i][2], v[i][3]);
        o.z = pack_bf16(v[i][4], v[i][5]); o.w = pack_bf16(v[i][6], v[i][7]);
        st_v4(residual_out + off, o);
#pragma unroll
        for (int j = 0; j < 8; ++j) ss += v[i][j] * v[i][j];
      }
    }
    if (w == nullptr) continue;
    const float total = ar_block_sum(ss, red);
    const float inv = rsqrtf(total / (float)H + eps);
#pragma unroll
    for (int i = 0; i < kArMaxVec; ++i) {
      const int vi = threadIdx.x + i * blockDim.x;
      if (vi < nvec) {
        const uint4 wa = ld_v4(w + (size_t)vi * 8);
        const float wf[8] = {bf16_lo(wa.x), bf16_hi(wa.x), bf16_lo(wa.y), bf16_hi(wa.y),
                             bf16_lo(wa.z), bf16_hi(wa.z), bf16_lo(wa.w), bf16_hi(wa.w)};
        float o[8];
#pragma unroll
        for (int j = 0; j < 8; ++j) o[j] = round_bf16(v[i][j] * inv) * wf[j];
        uint4 ov;
        ov.x = pack_bf16(o[0], o[1]); ov.y = pack_bf16(o[2], o[3]);
        ov.z = pack_bf16(o[4], o[5]); ov.w = pack_bf16(o[6], o[7]);
        st_v4(y + row_off + (size_t)vi * 8, ov);
      }
    }
  }
}

}  // namespace b200

using namespace b200;

// The kernel is ALWAYS launched with this many CTAs (rows are looped), so that the private
// per-CTA call counters stay equal to the call epoch.
#define B200_AR_BLOCKS 128

extern "C" int b200_allreduce_flag_words(int world) { return world + 1 + B200_AR_BLOCKS + 32; }

extern "C" int b200_allreduce_add_rms_norm(void* const* peers_host, void* const* flags_host,
                                           int world, int rank, uint32_t epoch_unused,
                                           const void* residual, const void* w, void* y,
                                           void* residual_out, int T, int H, float eps,
                                           void* stream) {
  (void)epoch_unused;  // epochs are kept on the device so that CUDA-graph replays stay valid
  B200_REQUIRE(world >= 2 && world <= kMaxWorld && rank >= 0 && rank < world,
               "allreduce: bad world/rank %d/%d", world, rank);
  B200_REQUIRE(T >= 0 && H > 0 && H % 8 == 0 && H <= 8 * kArMaxVec * 1024, "allreduce: bad H=%d", H);
  B200_REQUIRE(w == nullptr || residual_out != nullptr, "allreduce: residual_out required");
  if (T == 0) return B200_OK;
  PeerPtrs pp;
  for (int i = 0; i < kMaxWorld; ++i) {
    pp.data[i] = i < world ? (const bf16*)peers_host[i] : nullptr;
    pp.flags[i] = i < world ? (uint32_t*)flags_host[i] : nullptr;
  }
  int threads = (H / 8 + 31) / 32 * 32;
  if (threads > 512) threads = 512;
  const int grid = B200_AR_BLOCKS;
  (void)launch_pdl(allreduce_add_rms_norm_kernel, dim3(grid), dim3(threads), 0, (cudaStream_t)stream, 1, 
      pp, world, rank, B200_AR_BLOCKS, (const bf16*)residual, (const bf16*)w, (bf16*)y,
      (bf16*)residual_out, T, H, eps);
  return b200_check_launch("b200_allreduce_add_rms_norm");
}
