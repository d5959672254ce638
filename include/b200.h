/*
 * b200.h — C ABI of libb200.so, the sm_100a kernel library behind the engine.
 *
 * GPUStack has no in-process engine ABI: its engine boundary is process + CLI + HTTP
 * (reference: gpustack/worker/backends/custom.py:141-159, gpustack/worker/serve_manager.py:1000-1048,
 * gpustack/routes/openai.py:92-181).  That boundary is implemented by gpustack_b200/{cli,server}.py.
 * This header is the INTERNAL boundary the north star asks for ("Python over PyTorch tensors
 * calling the kernels through a thin C-ABI extension"): one entry point per operator on the
 * decoder hot path, each citing the operator of the reference's vLLM backend it replaces
 * ($VLLM = site-packages/vllm 0.22.0, the engine gpustack/worker/backends/vllm.py:403-407 launches).
 *
 * Conventions
 *   - every pointer is a DEVICE pointer unless its name ends in _host; the caller owns all memory;
 *   - kernels never allocate and never synchronise; `stream` is a cudaStream_t passed as void*;
 *   - return 0 on success, negative on error (B200_ERR_*), message via b200_last_error();
 *   - bf16 tensors are row-major, innermost dimension contiguous, 16-byte aligned rows.
 */
#ifndef B200_H_
#define B200_H_

#include <stddef.h>
#include <stdint.h>

#ifdef __cplusplus
extern "C" {
#endif

#define B200_ABI_VERSION 1

#define B200_ERR_INVALID (-1)
#define B200_ERR_CUDA (-2)
#define B200_ERR_UNSUPPORTED (-3)

/* ---- runtime ------------------------------------------------------------------------------ */
int b200_abi_version(void);
const char* b200_last_error(void);
/* sm count, compute capability and L2 size of the current device. */
int b200_device_info(int* sm_count, int* cc_major, int* cc_minor, int* l2_bytes);

/* ---- normalisation: $VLLM/model_executor/layers/layernorm.py:104-150,
 *      math $VLLM/ir/ops/layernorm.py:10-58 ------------------------------------------------- */
/* y = bf16(bf16(x * rsqrt(mean(x^2) + eps)) * w); x,y [T,H] bf16, w [H] bf16. */
int b200_rms_norm(const void* x, const void* w, void* y, int T, int H, float eps, void* stream);
/* z = x + residual (fp32); residual_out = bf16(z); y = rmsnorm(z) as above.
 * y may alias x, residual_out may alias residual (the reference op is in-place). */
int b200_fused_add_rms_norm(const void* x, const void* residual, const void* w, void* y,
                            void* residual_out, int T, int H, float eps, void* stream);
/* Same, but x arrives as `splits` fp32 partial sums [splits][T][H] (split-K GEMM output);
 * the partials are summed in fp32 before the residual add.  residual may be NULL (first layer). */
int b200_fused_add_rms_norm_partials(const float* partials, int splits, const void* residual,
                                     const void* w, void* y, void* residual_out, int T, int H,
                                     float eps, void* stream);

/* ---- rotary embedding + KV-cache write:
 *      $VLLM/model_executor/layers/rotary_embedding/base.py:150-215 (neox style) and
 *      $VLLM/_custom_ops.py:2591-2610 (reshape_and_cache_flash) -------------------------------- */
/* qkv [T, (Hq+2*Hkv)*D] bf16: q and k heads are rotated in place; rotated k and v are scattered
 * into k_cache/v_cache laid out [num_blocks][Hkv][block_size][D] at slot_mapping[t]
 * (= block*block_size + offset; negative = skip).  cos_sin [max_pos][D] bf16 = cos | sin halves. */
int b200_rope_kv_write(void* qkv, const int32_t* positions, const void* cos_sin, void* k_cache,
                       void* v_cache, const int32_t* slot_mapping, int T, int Hq, int Hkv, int D,
                       int block_size, void* stream);

/* Same, with the QKV projection arriving as `splits` fp32 split-K partial sums
 * [splits][T][(Hq+2Hkv)*D]: they are summed, the optional QKV bias [(Hq+2Hkv)*D] bf16 (Qwen2) is added
 * in fp32, the result is rounded to bf16 (what the un-split GEMM would have stored), rotated,
 * written to qkv_out (bf16, same layout) and scattered into the cache. */
int b200_rope_kv_write_partials(const float* partials, int splits, const void* qkv_bias, void* qkv_out,
                                const int32_t* positions, const void* cos_sin, void* k_cache,
                                void* v_cache, const int32_t* slot_mapping, int T, int Hq, int Hkv,
                                int D, int block_size, void* stream);

/* ---- activation: $VLLM/model_executor/layers/activation.py:133-143 ----------------------- */
/* out[t, i] = bf16(bf16(silu(x[t, i])) * x[t, I + i]);  x [T, 2I] bf16, out [T, I] bf16. */
int b200_silu_and_mul(const void* x, void* out, int T, int I, void* stream);

/* ---- embedding + greedy sampler:
 *      $VLLM/model_executor/layers/vocab_parallel_embedding.py, $VLLM/v1/sample/sampler.py -- */
/* out[t] = table[ids[t] - vocab_start] if vocab_start <= ids[t] < vocab_end else 0. */
int b200_embedding(const int32_t* ids, const void* table, void* out, int T, int H, int vocab_start,
                   int vocab_end, void* stream);
/* out_idx[b] = lowest index of the row maximum of logits[b, :V] (fp32, row stride ld);
 * out_val[b] (optional) = that maximum. */
int b200_argmax_f32(const float* logits, int64_t ld, int32_t* out_idx, float* out_val, int B, int V,
                    void* stream);
/* Sampler, one pass over the logits: row b is greedy when temps[b] == 0 (or temps is NULL), otherwise an
 * exact softmax(logits / temps[b]) sample by the Gumbel-max trick with the counter-based noise
 * u = hash(seeds[b], vocab_offset + column) -- deterministic per (seed, position), and composable with
 * vocabulary-parallel shards (the cross-rank (max, argmax) exchange of greedy picks the sample).
 * out_val (optional): the winning key of each row. */
int b200_sample_f32(const float* logits, int64_t ld, const float* temps, const uint32_t* seeds,
                    int vocab_offset, int32_t* out_idx, float* out_val, int B, int V, void* stream);
/* gather rows: out[i] = x[rows[i]]; x [*,H] bf16 (used to pick each sequence's last token). */
int b200_gather_rows(const void* x, const int32_t* rows, void* out, int n, int H, void* stream);

/* ---- dense projection (tcgen05/TMEM, TMA-fed):
 *      $VLLM/model_executor/layers/utils.py:92-98 (F.linear of QKV/O/gate_up/down/lm_head) ---- */
/* Y[M,N] = X[M,K] * W[N,K]^T.  X,W bf16 (K contiguous, ldx/ldw in elements), fp32 accumulate.
 *   out_f32 == 0: Y bf16 [M, ldy];  out_f32 == 1: Y fp32 [M, ldy].
 *   splits > 1 : K is cut into `splits` equal ranges and Y must be fp32 [splits][M][ldy]
 *                partial sums (reduced by b200_fused_add_rms_norm_partials / b200_splitk_reduce).
 * K % 64 == 0 required.  M, N arbitrary (tails are zero-filled by TMA and masked on store). */
int b200_gemm_bf16(const void* X, const void* W, void* Y, int M, int N, int K, int64_t ldx,
                   int64_t ldw, int64_t ldy, int out_f32, int splits, void* stream);
/* Same with an explicit activation-tile width block_x in {16,32,64,128,256} (0 = pick from M, see
 * b200_gemm_pick_block_x); used by the tuning sweep in tools/gemm_sweep.py. */
int b200_gemm_bf16_ex(const void* X, const void* W, void* Y, int M, int N, int K, int64_t ldx,
                      int64_t ldw, int64_t ldy, int out_f32, int splits, int block_x, void* stream);
int b200_gemm_pick_block_x(int M, int N, int K);
/* gate_up projection with the SiLU*mul activation ($VLLM activation.py:133-143) fused into the
 * epilogue.  W [N=2I, K] holds gate and up rows INTERLEAVED (row 2i = gate_i, row 2i+1 = up_i);
 * Y [M, ldy >= I] bf16 = bf16(bf16(silu(bf16(g))) * bf16(u)): bit-identical to
 * b200_gemm_bf16 followed by b200_silu_and_mul. */
int b200_gemm_bf16_silu_mul(const void* X, const void* W, void* Y, int M, int N, int K, int64_t ldx,
                            int64_t ldw, int64_t ldy, void* stream);
/* W8A8 fp8 (e4m3) projection -- the linear layers of the fp8 checkpoints GPUStack deploys
 * (BASELINE config "Llama-3-70B fp8"; reference operator: cutlass_scaled_mm behind
 * Fp8LinearMethod.apply, $VLLM/model_executor/layers/quantization/fp8.py:437-470):
 *   Y[m,n] = (sum_k Xq[m,k] * Wq[n,k]) * xscale[m] * wscale[n * wscale_stride]
 * Xq [M,K], Wq [N,K] e4m3 bytes (K-major, K % 128 == 0, N > 128); xscale [M] fp32 dynamic
 * per-token scales (b200_quant_fp8_per_token); wscale fp32, stride 0 = per-tensor, 1 = per output
 * channel.  out_f32/splits as b200_gemm_bf16.  silu_mul != 0: W rows are interleaved gate/up and
 * Y [M, N/2] bf16 gets SiLU*mul of the dequantised, bf16-rounded pair (as b200_gemm_bf16_silu_mul). */
int b200_gemm_fp8(const void* X, const void* W, void* Y, int M, int N, int K, int64_t ldx, int64_t ldw,
                  int64_t ldy, const float* xscale, const float* wscale, int wscale_stride, int out_f32,
                  int splits, int silu_mul, void* stream);
/* Dynamic per-token e4m3 quantisation of bf16 activations (QuantFP8.forward_native,
 * $VLLM/model_executor/layers/quantization/input_quant_fp8.py:188-235):
 *   scales[t] = max(absmax(x[t,:]) / 448, 1/(448*512));  q[t,k] = e4m3(x[t,k] * (1/scales[t])), RNE, saturating. */
int b200_quant_fp8_per_token(const void* x, int64_t ldx, void* q, int64_t ldq, float* scales, int T, int K,
                             void* stream);
/* ---- mixture of experts (Mixtral family): MixtralMoE -> FusedMoE, $VLLM/model_executor/models/
 * mixtral.py, $VLLM/model_executor/layers/fused_moe/ (topk_softmax, moe_align_block_size,
 * fused_experts, moe_sum).  Everything stays on the device; tile counts are read by the grouped
 * GEMM from device memory, so a MoE layer is CUDA-graph capturable. ---- */
/* topk_w[t,:] = (renormalised) top-K of softmax(float(router_logits[t,:E])), router_logits bf16 with row
 * stride ld >= E; topk_ids int32 [T,K]; ties go to the lowest expert index.  E <= 64, K <= 8. */
int b200_moe_route(const void* router_logits, int64_t ld, int32_t* topk_ids, float* topk_w, int T, int E,
                   int K, int renormalize, void* stream);
/* Rows the permuted space can need for n_pairs (token,k) pairs: every expert is padded to whole
 * tiles of block_x rows. */
int b200_moe_capacity_rows(int n_pairs, int E, int block_x);
/* Group the pairs by expert.  pair_row [n_pairs] = row of each pair; row_pair [capacity_rows] = pair of
 * each row (-1 = padding); row_weight [capacity_rows] = routing weight (0 = padding); tile_expert
 * [capacity_rows / block_x] = owner of each tile; num_tiles[0] = tiles in use, [1] = rows in use. */
int b200_moe_align(const int32_t* topk_ids, const float* topk_w, int n_pairs, int E, int block_x,
                   int capacity_rows, int32_t* pair_row, int32_t* row_pair, float* row_weight,
                   int32_t* tile_expert, int32_t* num_tiles, void* stream);
/* xp[row,:] = x[row_pair[row] / K, :] (zeros for padding rows) for the rows in use. */
int b200_moe_gather(const void* x, const int32_t* row_pair, const int32_t* num_tiles, void* xp, int H, int K,
                    int capacity_rows, void* stream);
/* Grouped tcgen05 GEMM over the permuted rows: Y[row,:] = X[row,:] * W[tile_expert[row / block_x]]^T.
 * W [num_experts * N, K] bf16 (experts stacked, N % 256 == 0, K % 64 == 0).  row_scale (optional)
 * multiplies each row's fp32 accumulator (routing weight); silu_mul != 0: W rows are interleaved
 * gate/up and Y is [capacity_rows, N/2] (as b200_gemm_bf16_silu_mul). */
int b200_moe_grouped_gemm(const void* X, const void* W, void* Y, int capacity_rows, int N, int K,
                          int num_experts, int64_t ldx, int64_t ldw, int64_t ldy, const float* row_scale,
                          const int32_t* tile_expert, const int32_t* num_tiles, int block_x, int silu_mul,
                          void* stream);
/* out[t,:] = bf16(sum_j float(yp[pair_row[t*K+j], :]))  (moe_sum). */
int b200_moe_combine(const void* yp, const int32_t* pair_row, void* out, int T, int H, int K, void* stream);
/* Y[m,n] = bf16(sum_s partials[s][m][n]). */
int b200_splitk_reduce(const float* partials, int splits, void* Y, int M, int N, int64_t ldy,
                       void* stream);
/* Heuristic used by the engine: how many K-splits b200_gemm_bf16 should use for this shape. */
int b200_gemm_suggest_splits(int M, int N, int K);

/* ---- paged-KV attention: $VLLM/v1/attention/backends/flashinfer.py:1095-1230 -------------- */
/* Decode: one query token per sequence.  q [B, Hq, D] bf16 with row stride q_stride (elements),
 * out [B, Hq*D] bf16, block_table [B, max_blocks] int32, seq_lens [B] int32 (incl. the new
 * token).  workspace: b200_attn_decode_workspace_bytes() bytes, ZEROED ONCE by the caller before its
 * first use (it starts with the self-resetting arrival counters of the in-kernel split merge; a
 * buffer sized for a larger batch may be reused for smaller ones); num_splits from
 * b200_attn_decode_splits() or caller-chosen (>=1). */
int b200_attn_decode_splits(int B, int Hkv, int max_seq_len, int block_size);
size_t b200_attn_decode_workspace_bytes(int B, int Hq, int D, int num_splits);
int b200_paged_attn_decode(const void* q, int64_t q_stride, const void* k_cache,
                           const void* v_cache, const int32_t* block_table, int max_blocks,
                           const int32_t* seq_lens, void* out, void* workspace, int num_splits,
                           int B, int Hq, int Hkv, int D, int block_size, float scale,
                           void* stream);
/* Prefill / chunked prefill: ragged batch of query chunks.  Sequence s contributes tokens
 * [cu_q[s], cu_q[s+1]) of q; its chunk starts at context position ctx_lens[s] (tokens already in
 * the cache) and attends causally to cache positions [0, ctx_lens[s] + i].  K/V of the chunk must
 * already be in the cache (b200_rope_kv_write runs first).  tile_seq/tile_q0 [num_tiles] list the
 * (sequence, first query offset) of each 32-token query tile, built on the host.  num_tokens = rows
 * of q, num_blocks = pages of the cache tensors (bounds of the TMA tensor maps of the tcgen05 kernel
 * that serves head_dim 128 with GQA groups of 4 or 8; other shapes run the mma.sync kernel). */
/* query-tile size (tokens) the prefill kernel uses for this head grouping: 32 (group <= 4),
 * 16 (group <= 8), 0 = unsupported. */
int b200_attn_prefill_tile(int Hq, int Hkv);
int b200_paged_attn_prefill(const void* q, int64_t q_stride, int num_tokens, const void* k_cache,
                            const void* v_cache, int num_blocks, const int32_t* block_table, int max_blocks,
                            const int32_t* cu_q, const int32_t* ctx_lens, const int32_t* tile_seq,
                            const int32_t* tile_q0, int num_tiles, void* out, int Hq, int Hkv, int D,
                            int block_size, float scale, void* stream);

/* ---- tensor-parallel all-reduce over NVLink peer memory:
 *      $VLLM/distributed/device_communicators/cuda_communicator.py:241-290 (custom all-reduce) */
/* Zero-filled device allocation made directly with cudaMalloc (so that it has its own IPC handle;
 * framework allocators sub-allocate) and its release. */
int b200_dev_malloc(size_t bytes, void** dev_ptr_out);
int b200_dev_free(void* dev_ptr);
int b200_ipc_get_handle(void* dev_ptr, void* handle_out_host /* 64 bytes */);
int b200_ipc_open_handle(const void* handle_host /* 64 bytes */, void** dev_ptr_out);
int b200_ipc_close_handle(void* dev_ptr);
/* uint32 words each rank's flag buffer must hold (zero-initialised once, peer-mapped). */
int b200_allreduce_flag_words(int world);
/* One-shot all-reduce(sum) of bf16 [T,H] fused with residual add + RMSNorm of the reduced rows:
 * peers[r] = rank r's staging buffer (already filled by that rank's row-parallel GEMM);
 * flags[r] = rank r's flag words.  Every rank reduces all rows (redundantly, in rank order, so
 * results are bit-identical across ranks) -- no second exchange.  w == NULL: plain all-reduce
 * into y.  Arrival epochs live on the device (CUDA-graph replays stay valid); `epoch` is unused.
 * The caller alternates between two staging buffers on consecutive calls. */
int b200_allreduce_add_rms_norm(void* const* peers_host, void* const* flags_host, int world,
                                int rank, uint32_t epoch, const void* residual, const void* w,
                                void* y, void* residual_out, int T, int H, float eps, void* stream);

/* Two-shot, sequence-parallel form for large messages (prefill chunks, big decode batches): ONE
 * kernel = barrier, reduce-scatter by token rows over peer loads, residual add + RMSNorm on this
 * rank's rows only, all-gather of the normalised rows by peer stores, barrier.
 * stage_host[r] / xn_host[r] / flags_host[r]: rank r's GEMM-partial staging buffer [T,H] bf16, its
 * gathered-output buffer [T,H] bf16 (every rank's copy is complete when the kernel ends) and its
 * flag buffer (b200_allreduce_twoshot_flag_words() zero-initialised uint32 words), as mapped in the
 * calling process.  The residual stream is sharded by rows between calls: only rows
 * [rank*ceil(T/world), (rank+1)*ceil(T/world)) of residual / residual_out are read / written. */
int b200_allreduce_twoshot_flag_words(void);
/* Greedy sampling over vocabulary-parallel logits: rank r holds, for each of n sequences, the max
 * (val) and LOCAL argmax (idx) of its lm_head shard in peer-mapped arrays; one CTA per rank reads
 * every rank's pair over NVLink and writes the global argmax (idx + r * vocab_per_rank; ties -> the
 * lowest vocabulary index) to out_idx.  flags_host[r]: >= 32 zero-initialised uint32 words. */
int b200_tp_argmax_select(void* const* val_host, void* const* idx_host, void* const* flags_host,
                          int world, int rank, int vocab_per_rank, int32_t* out_idx, int n, void* stream);
int b200_allreduce_rs_norm_ag(void* const* stage_host, void* const* xn_host, void* const* flags_host,
                              int world, int rank, const void* residual, const void* w,
                              void* residual_out, int T, int H, float eps, void* stream);

#ifdef __cplusplus
}
#endif
#endif /* B200_H_ */
