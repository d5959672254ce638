"""Exact VRAM model of this engine, for GPUStack's placement.

GPUStack places a custom backend with `VRAM = weight_file_bytes * 1.2 + 2 GiB`
(gpustack/policies/utils.py:283-360; used by CustomBackendResourceFitSelector,
gpustack/policies/candidate_selectors/custom_backend_resource_fit_selector.py:105-121, claim split
evenly over the chosen GPUs :289-292) unless the model carries `env.GPUSTACK_MODEL_VRAM_CLAIM`
(utils.py:241-253).  That heuristic knows nothing about the KV cache the deployment asks for.  This
module computes what the engine really allocates -- weights (per dtype / quantisation), the
preallocated activation workspace, CUDA-graph pool, and either the KV cache needed for a requested
(max_model_len x max_num_seqs) or the whole `gpu_memory_utilization` share the engine would take --
and renders the env entry a maintainer puts into the model spec.  `b200-engine estimate <dir> [--kv-tokens N]`
prints it (default KV capacity: one sequence of max_model_len = the smallest claim the engine starts with).
"""
from __future__ import annotations

import math
from dataclasses import asdict, dataclass
from typing import Optional

from .config import EngineConfig, ModelConfig

GiB = 1024 ** 3
GRAPH_POOL_BYTES = 1 * GiB          # what ModelRunner.profile_num_blocks reserves for graphs / allocator slack
CUDA_CONTEXT_BYTES = 600 * 1024 ** 2


@dataclass
class VramEstimate:
    tensor_parallel_size: int
    weights: int                 # per GPU
    workspace: int               # per GPU: activations, logits, split-K partials, staging, MoE / fp8 buffers
    kv_tokens: int               # KV capacity the claim is sized for (default: ONE sequence of max_model_len,
                                 # the engine's start-up condition, runner.py profile_num_blocks)
    kv_cache_requested: int      # per GPU bytes for kv_tokens
    kv_bytes_per_token: int      # per GPU
    graphs_and_context: int
    per_gpu_min: int             # smallest claim with which the engine starts and holds the requested KV
    total_min: int               # = per_gpu_min * tp  -> GPUSTACK_MODEL_VRAM_CLAIM
    reference_heuristic: int     # what GPUStack would claim on its own (weights_on_disk * 1.2 + 2 GiB)

    def env(self) -> dict:
        return {"GPUSTACK_MODEL_VRAM_CLAIM": str(self.total_min)}

    def to_dict(self) -> dict:
        d = asdict(self)
        d["env"] = self.env()
        return d


def weight_bytes_per_gpu(m: ModelConfig, tp: int, quantization: Optional[str] = None) -> int:
    """Resident weight bytes of one rank (embedding and lm_head vocab-parallel, norms replicated)."""
    H, I, L = m.hidden_size, m.intermediate_size, m.num_hidden_layers
    q = quantization or m.quantization
    lin_b = 1 if q == "fp8" else 2
    attn = ((m.q_size + 2 * m.kv_size) * H + m.q_size * H) // tp
    if m.num_local_experts:
        mlp = m.num_local_experts * 3 * I * H // tp
        router = 128 * H * 2                                  # padded to one GEMM tile, replicated, bf16
    else:
        mlp, router = 3 * I * H // tp, 0
    bias = (m.q_size + 2 * m.kv_size) // tp * 2 if m.attention_bias else 0
    per_layer = (attn + mlp) * lin_b + router + bias + 2 * H * 2 + (16 if q == "fp8" else 0)
    vocab_local = math.ceil(m.vocab_size / tp)
    embed = vocab_local * H * 2
    lm_head = 0 if m.tie_word_embeddings else vocab_local * H * 2     # tied: the embedding shard itself
    return L * per_layer + embed + lm_head + H * 2


def workspace_bytes_per_gpu(m: ModelConfig, e: EngineConfig, tp: int, quantization: Optional[str] = None) -> int:
    """Mirror of ModelRunner._alloc_workspace (runner.py)."""
    T = max(e.max_num_batched_tokens, e.max_num_seqs)
    S = e.max_num_seqs
    H, I = m.hidden_size, m.intermediate_size // tp
    hq, hkv, D = m.num_attention_heads // tp, m.num_key_value_heads // tp, m.head_dim
    qkv_w = (hq + 2 * hkv) * D
    v_local = math.ceil(m.vocab_size / tp)
    S_smp = min(T, S * (1 + e.num_speculative_tokens))       # rows one step may sample (runner.max_sample)
    act = T * (3 * H + qkv_w + hq * D + I) * 2 + S_smp * H * 2
    logits = S_smp * v_local * 4
    partials = max(8 * min(256, T) * max(H, qkv_w), T * qkv_w if m.attention_bias else 0) * 4
    max_len = e.max_model_len or m.max_position_embeddings
    rope = max_len * D * 2
    staging = (3 * T + 6 * S + 3 * S_smp + S * math.ceil(max_len / e.block_size) + 2 * T) * 4
    extra = 0
    if (quantization or m.quantization) == "fp8":
        extra += T * max(H, hq * D, I) + T * 4
    if m.num_local_experts:
        pairs = T * m.num_experts_per_tok
        cap = pairs + m.num_local_experts * 255 + 255
        extra += cap * (2 * H + I) * 2 + cap * 12 + pairs * 12 + T * 128 * 2
    return act + logits + partials + rope + staging + extra


def estimate(m: ModelConfig, e: EngineConfig, weights_on_disk: Optional[int] = None,
             kv_tokens: Optional[int] = None) -> VramEstimate:
    tp = max(1, e.tensor_parallel_size)
    w = weight_bytes_per_gpu(m, tp, e.quantization)
    ws = workspace_bytes_per_gpu(m, e, tp, e.quantization)
    kv_tok = m.kv_bytes_per_token(tp)
    max_len = min(e.max_model_len or m.max_position_embeddings, m.max_position_embeddings)
    kv_tokens = max_len if kv_tokens is None else max(int(kv_tokens), max_len)
    blocks = math.ceil(kv_tokens / e.block_size) + 1 + (e.max_num_seqs if kv_tokens > max_len else 0)
    kv = blocks * e.block_size * kv_tok
    fixed = GRAPH_POOL_BYTES + CUDA_CONTEXT_BYTES
    # the engine sizes its KV cache as (free - (1 - util) * total - 1 GiB): the claim that yields
    # `kv` bytes of cache on a GPU of any size is therefore everything it allocates, divided by util
    per_gpu = int(math.ceil((w + ws + kv + fixed) / max(0.05, e.gpu_memory_utilization)))
    disk = weights_on_disk if weights_on_disk is not None else w * tp
    return VramEstimate(tensor_parallel_size=tp, weights=w, workspace=ws, kv_tokens=kv_tokens, kv_cache_requested=kv,
                        kv_bytes_per_token=kv_tok, graphs_and_context=fixed, per_gpu_min=per_gpu,
                        total_min=per_gpu * tp, reference_heuristic=int(disk * 1.2) + 2 * GiB)
