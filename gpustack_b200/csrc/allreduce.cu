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

// =================================================================================================
// Two-shot, sequence-parallel variant for large messages (prefill chunks; decode batches whose
// (world-1) x message no longer fits the one-shot budget): ONE kernel per row-parallel projection
//
//   barrier A      every rank's GEMM partial [T,H] is in its staging buffer
//   reduce-scatter rank r sums rows [r*T/world, (r+1)*T/world) of all peers' partials (peer loads)
//   add + RMSNorm  on those rows only: the residual stream stays sharded by token rows between
//                  calls (each rank ever reads / writes only its own rows of `residual`), so the
//                  norm work is 1/world of the replicated version
//   all-gather     the normalised rows are pushed into every peer's xn buffer (peer stores)
//   barrier B      all pushes have landed; xn is complete on every rank when the kernel ends
//
// Per rank and direction T*H*2*(world-1)/world bytes cross NVLink in each of the two phases (the
// algorithmic minimum of an all-reduce), against (world-1) x T*H*2 inbound for the one-shot kernel.
// Flag words (uint32, one buffer per rank): [0,8) barrier-A arrivals, [8,16) barrier-B arrivals,
// [16] local go word, [17] local done counter, [32, 32+grid) per-CTA call counters (as above).
struct TwoShotPtrs {
  const bf16* stage[kMaxWorld];
  bf16* xn[kMaxWorld];
  uint32_t* flags[kMaxWorld];
};

__device__ __forceinline__ void spin_until(const uint32_t* p, uint32_t target, bool sys) {
  uint64_t t0 = 0;
  uint32_t spins = 0;
  while ((int32_t)((sys ? ld_acquire_sys(p) : ld_acquire_gpu(p)) - target) < 0) {
    if ((++spins & 1023u) == 0) {
      uint64_t now;
      asm volatile("mov.u64 %0, %%globaltimer;" : "=l"(now));
      if (t0 == 0) t0 = now;
      else if (now - t0 > 10000000000ull) __trap();   // a peer / CTA never arrived: fail loudly
    }
  }
}

__global__ void __launch_bounds__(512)
allreduce_rs_norm_ag_kernel(TwoShotPtrs P, int world, int rank, const bf16* __restrict__ residual,
                            const bf16* __restrict__ w, bf16* residual_out, int T, int H, float eps) {
  pdl_launch_dependents();
  pdl_wait();
  __shared__ float red[32];
  __shared__ uint32_t s_epoch;
  uint32_t* my_flags = P.flags[rank];
  uint32_t* go = my_flags + 16;
  uint32_t* done = my_flags + 17;
  if (threadIdx.x == 0) {
    uint32_t* ctr = my_flags + 32 + blockIdx.x;
    s_epoch = *ctr + 1;
    *ctr = s_epoch;
  }
  __syncthreads();
  const uint32_t epoch = s_epoch;

  // ---- barrier A (CTA 0 talks to the peers, the other CTAs wait for its go word)
  if (blockIdx.x == 0) {
    if ((int)threadIdx.x < world && (int)threadIdx.x != rank) {
      st_release_sys(P.flags[threadIdx.x] + rank, epoch);
      spin_until(my_flags + threadIdx.x, epoch, true);
    }
    __syncthreads();
    if (threadIdx.x == 0) st_release_gpu(go, epoch);
  } else {
    if (threadIdx.x == 0) spin_until(go, epoch, false);
    __syncthreads();
  }

  // ---- reduce-scatter + add + norm on this rank's rows, all-gather by peer stores
  const int rpr = (T + world - 1) / world;
  const int r0 = rank * rpr;
  const int r1 = min(T, r0 + rpr);
  const int nvec = H >> 3;
  for (int row = r0 + blockIdx.x; row < r1; row += gridDim.x) {
    const size_t row_off = (size_t)row * H;
    float v[kArMaxVec][8];
    float ss = 0.f;
#pragma unroll
    for (int i = 0; i < kArMaxVec; ++i) {
      const int vi = threadIdx.x + i * blockDim.x;
      if (vi < nvec) {
        const size_t off = row_off + (size_t)vi * 8;
        uint4 a[kMaxWorld];
#pragma unroll
        for (int p = 0; p < kMaxWorld; ++p)
          if (p < world) a[p] = ld_peer_v4(P.stage[p] + off);   // all peer loads in flight together
#pragma unroll
        for (int j = 0; j < 8; ++j) v[i][j] = 0.f;
#pragma unroll
        for (int p = 0; p < kMaxWorld; ++p) {
          if (p < world) {
            v[i][0] += bf16_lo(a[p].x); v[i][1] += bf16_hi(a[p].x);
            v[i][2] += bf16_lo(a[p].y); v[i][3] += bf16_hi(a[p].y);
            v[i][4] += bf16_lo(a[p].z); v[i][5] += bf16_hi(a[p].z);
            v[i][6] += bf16_lo(a[p].w); v[i][7] += bf16_hi(a[p].w);
          }
        }
        if (residual != nullptr) {
          const uint4 ra = ld_cg_v4(residual + off);
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
        const size_t off = row_off + (size_t)vi * 8;
#pragma unroll
        for (int p = 0; p < kMaxWorld; ++p)
          if (p < world) st_v4(P.xn[p] + off, ov);   // push the normalised row to every rank
      }
    }
  }

  // ---- barrier B: this rank's pushes are visible system-wide, then every peer's are
  __threadfence_system();
  __syncthreads();
  if (blockIdx.x != 0) {
    if (threadIdx.x == 0) atomicAdd(done, 1u);
  } else {
    if (threadIdx.x == 0) spin_until(done, epoch * (gridDim.x - 1), false);
    __syncthreads();
    if ((int)threadIdx.x < world && (int)threadIdx.x != rank) {
      st_release_sys(P.flags[threadIdx.x] + 8 + rank, epoch);
      spin_until(my_flags + 8 + threadIdx.x, epoch, true);
    }
  }
}

// =================================================================================================
// Greedy sampling under vocabulary-parallel logits: every rank has the (max, argmax) of its vocab
// shard for each of the n sequences (argmax_f32_kernel wrote them into peer-mapped arrays); one CTA
// exchanges them through NVLink peer loads and keeps the global winner (ties -> lowest vocabulary
// index, i.e. lowest rank).  Replaces two NCCL all-gathers + four framework kernels per step
// (reference: the logits all-gather of $VLLM/model_executor/layers/logits_processor.py).
// flags (uint32 per rank): [0,8) barrier-A arrivals, [8,16) barrier-B arrivals, [16] call counter.
struct ArgmaxPtrs {
  const float* val[kMaxWorld];
  const int32_t* idx[kMaxWorld];
  uint32_t* flags[kMaxWorld];
};

__global__ void __launch_bounds__(1024)
tp_argmax_select_kernel(ArgmaxPtrs P, int world, int rank, int vocab_per_rank, int32_t* __restrict__ out_idx,
                        int n) {
  pdl_launch_dependents();
  pdl_wait();
  __shared__ uint32_t s_epoch;
  uint32_t* my_flags = P.flags[rank];
  if (threadIdx.x == 0) {
    s_epoch = my_flags[16] + 1;
    my_flags[16] = s_epoch;
  }
  __syncthreads();
  const uint32_t epoch = s_epoch;
  if ((int)threadIdx.x < world && (int)threadIdx.x != rank) {
    st_release_sys(P.flags[threadIdx.x] + rank, epoch);       // my (val, idx) arrays are complete
    spin_until(my_flags + threadIdx.x, epoch, true);
  }
  __syncthreads();
  for (int i = threadIdx.x; i < n; i += blockDim.x) {
    float best = -INFINITY;
    int32_t bi = 0;
    for (int p = 0; p < world; ++p) {
      float v;
      int32_t k;
      asm volatile("ld.global.cg.f32 %0, [%1];" : "=f"(v) : "l"(P.val[p] + i) : "memory");
      asm volatile("ld.global.cg.s32 %0, [%1];" : "=r"(k) : "l"(P.idx[p] + i) : "memory");
      if (p == 0 || v > best) { best = v; bi = k + p * vocab_per_rank; }
    }
    out_idx[i] = bi;
  }
  __syncthreads();
  // barrier B: nobody overwrites its arrays (next step's argmax) before every peer has read them
  if ((int)threadIdx.x < world && (int)threadIdx.x != rank) {
    st_release_sys(P.flags[threadIdx.x] + 8 + rank, epoch);
    spin_until(my_flags + 8 + threadIdx.x, epoch, true);
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

extern "C" int b200_allreduce_twoshot_flag_words(void) { return 32 + B200_AR_BLOCKS + 32; }

// Two-shot sequence-parallel all-reduce + residual add + RMSNorm (see the kernel comment).
// stage_host[r] / xn_host[r] / flags_host[r]: rank r's staging, gathered-output and flag buffers as
// mapped in THIS process (own pointers for r == rank, CUDA-IPC mappings otherwise).  residual /
// residual_out are local and only rows [rank*ceil(T/world), ...) of them are read / written.
extern "C" int b200_allreduce_rs_norm_ag(void* const* stage_host, void* const* xn_host, void* const* flags_host,
                                         int world, int rank, const void* residual, const void* w,
                                         void* residual_out, int T, int H, float eps, void* stream) {
  B200_REQUIRE(world >= 2 && world <= kMaxWorld && rank >= 0 && rank < world, "allreduce(2shot): bad world/rank %d/%d",
               world, rank);
  B200_REQUIRE(T >= 0 && H > 0 && H % 8 == 0 && H <= 8 * kArMaxVec * 512, "allreduce(2shot): bad H=%d", H);
  B200_REQUIRE(w != nullptr && residual_out != nullptr, "allreduce(2shot): norm weight and residual_out required");
  if (T == 0) return B200_OK;
  TwoShotPtrs pp;
  for (int i = 0; i < kMaxWorld; ++i) {
    pp.stage[i] = i < world ? (const bf16*)stage_host[i] : nullptr;
    pp.xn[i] = i < world ? (bf16*)xn_host[i] : nullptr;
    pp.flags[i] = i < world ? (uint32_t*)flags_host[i] : nullptr;
  }
  int threads = (H / 8 + 31) / 32 * 32;
  if (threads > 512) threads = 512;
  (void)launch_pdl(allreduce_rs_norm_ag_kernel, dim3(B200_AR_BLOCKS), dim3(threads), 0, (cudaStream_t)stream, 1, pp,
                   world, rank, (const bf16*)residual, (const bf16*)w, (bf16*)residual_out, T, H, eps);
  return b200_check_launch("b200_allreduce_rs_norm_ag");
}

// Global greedy token of n sequences from per-rank (max, argmax) arrays.  val_host[r] / idx_host[r] /
// flags_host[r]: rank r's fp32 maxima [n], int32 LOCAL argmax [n] and flag words (>= 32 zeroed uint32)
// as mapped in this process; vocab_per_rank = rows of one lm_head shard (ceil(V / world)).
extern "C" int b200_tp_argmax_select(void* const* val_host, void* const* idx_host, void* const* flags_host,
                                     int world, int rank, int vocab_per_rank, int32_t* out_idx, int n,
                                     void* stream) {
  B200_REQUIRE(world >= 2 && world <= kMaxWorld && rank >= 0 && rank < world, "tp_argmax: bad world/rank %d/%d",
               world, rank);
  B200_REQUIRE(n >= 0 && out_idx != nullptr, "tp_argmax: bad arguments");
  if (n == 0) return B200_OK;
  ArgmaxPtrs pp;
  for (int i = 0; i < kMaxWorld; ++i) {
    pp.val[i] = i < world ? (const float*)val_host[i] : nullptr;
    pp.idx[i] = i < world ? (const int32_t*)idx_host[i] : nullptr;
    pp.flags[i] = i < world ? (uint32_t*)flags_host[i] : nullptr;
  }
  int threads = (n + 31) / 32 * 32;
  if (threads > 1024) threads = 1024;
  if (threads < 32) threads = 32;
  (void)launch_pdl(tp_argmax_select_kernel, dim3(1), dim3(threads), 0, (cudaStream_t)stream, 1, pp, world, rank,
                   vocab_per_rank, out_idx, n);
  return b200_check_launch("b200_tp_argmax_select");
}
