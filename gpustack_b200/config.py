"""Model and engine configuration.

ModelConfig is read from the HF-format `config.json` inside the local model directory that
GPUStack mounts and passes as {{model_path}} (gpustack/worker/backends/custom.py:91-134).
EngineConfig mirrors the subset of vLLM's CLI that GPUStack's launcher composes
(gpustack/worker/backends/vllm.py:395-507): --max-model-len, --tensor-parallel-size,
--gpu-memory-utilization, --max-num-seqs, --max-num-batched-tokens, --served-model-name.
"""
from __future__ import annotations

import json
from dataclasses import dataclass, field
from pathlib import Path
from typing import Optional

SUPPORTED_ARCHS = ("LlamaForCausalLM", "MistralForCausalLM", "Qwen2ForCausalLM", "MixtralForCausalLM")


@dataclass
class ModelConfig:
    vocab_size: int
    hidden_size: int
    intermediate_size: int
    num_hidden_layers: int
    num_attention_heads: int
    num_key_value_heads: int
    head_dim: int
    rms_norm_eps: float = 1e-5
    rope_theta: float = 500000.0
    max_position_embeddings: int = 8192
    tie_word_embeddings: bool = False
    attention_bias: bool = False
    architecture: str = "LlamaForCausalLM"
    eos_token_id: Optional[int] = None
    bos_token_id: Optional[int] = None
    eos_token_ids: tuple = ()                # every id that ends a turn (config.json list + generation_config.json)
    rope_scaling: Optional[dict] = None      # None = plain RoPE; {"rope_type": "llama3" | "linear", ...}
    quantization: Optional[str] = None       # "fp8" when config.json carries an fp8 quantization_config
    num_local_experts: int = 0               # Mixtral: experts per layer (0 = dense MLP)
    sliding_window: Optional[int] = None     # caps max_model_len (full attention == windowed below it)
    num_experts_per_tok: int = 0

    def __post_init__(self):
        ids = [int(t) for t in self.eos_token_ids]
        if self.eos_token_id is not None and int(self.eos_token_id) not in ids:
            ids.insert(0, int(self.eos_token_id))
        self.eos_token_ids = tuple(ids)
        if self.eos_token_id is None and ids:
            self.eos_token_id = ids[0]

    @property
    def q_size(self) -> int:
        return self.num_attention_heads * self.head_dim

    @property
    def kv_size(self) -> int:
        return self.num_key_value_heads * self.head_dim

    def kv_bytes_per_token(self, tp: int = 1) -> int:
        return 2 * self.num_hidden_layers * (self.num_key_value_heads // tp) * self.head_dim * 2

    def weight_bytes(self, tp: int = 1, quantization: Optional[str] = None) -> int:
        """bytes streamed per decode step on one rank (layers + lm_head + norms); the decoder linears
        are 1 byte/parameter under fp8, lm_head and norms stay bf16."""
        H, I = self.hidden_size, self.intermediate_size
        lin_b = 1 if (quantization or self.quantization) == "fp8" else 2
        mlp = 3 * I * H * max(1, self.num_local_experts)                      # all experts hot
        per_layer = (self.q_size + 2 * self.kv_size) * H + self.q_size * H + mlp
        router = self.num_local_experts * H * 2 * self.num_hidden_layers      # replicated, bf16
        norms = (2 * self.num_hidden_layers + 1) * H * 2                      # replicated
        return self.num_hidden_layers * per_layer // tp * lin_b + self.vocab_size * H // tp * 2 + norms + router

    @staticmethod
    def from_dict(d: dict) -> "ModelConfig":
        archs = d.get("architectures") or ["LlamaForCausalLM"]
        arch = archs[0]
        if arch not in SUPPORTED_ARCHS:
            raise ValueError(f"unsupported architecture {arch}; supported: {SUPPORTED_ARCHS}")
        heads = d["num_attention_heads"]
        eos = d.get("eos_token_id")
        eos_all = [int(t) for t in eos] if isinstance(eos, list) else ([] if eos is None else [int(eos)])
        eos = eos_all[0] if eos_all else None
        rope = d.get("rope_theta")
        rp = d.get("rope_parameters") if isinstance(d.get("rope_parameters"), dict) else None
        if rope is None and rp is not None:
            rope = rp.get("rope_theta")
        # legacy `rope_scaling` dict, or the transformers>=5 `rope_parameters` dict (same keys)
        sc = d.get("rope_scaling") if isinstance(d.get("rope_scaling"), dict) else rp
        rope_type = None if sc is None else sc.get("rope_type", sc.get("type"))
        scaling = None
        if rope_type not in (None, "default"):
            if rope_type == "llama3":
                scaling = {"rope_type": "llama3", "factor": float(sc["factor"]),
                           "low_freq_factor": float(sc.get("low_freq_factor", 1.0)),
                           "high_freq_factor": float(sc.get("high_freq_factor", 4.0)),
                           "original_max_position_embeddings": int(sc["original_max_position_embeddings"])}
            elif rope_type == "linear":
                scaling = {"rope_type": "linear", "factor": float(sc["factor"])}
            else:
                raise ValueError(f"rope scaling type {rope_type!r} is not supported (supported: default, llama3, linear)")
        quant = None
        qc = d.get("quantization_config")
        if isinstance(qc, dict):
            # per-tensor fp8 checkpoints with dynamic activation scales (what `--quantization fp8`
            # produces and what $VLLM/.../quantization/fp8.py:80-140 reads); anything else is refused
            method = qc.get("quant_method")
            if method != "fp8" or qc.get("activation_scheme", "dynamic") != "dynamic" or qc.get("weight_block_size"):
                raise ValueError(f"unsupported quantization_config {qc}: only quant_method=fp8, "
                                 "activation_scheme=dynamic, per-tensor weights are supported")
            quant = "fp8"
        n_exp = int(d.get("num_local_experts") or 0) if arch == "MixtralForCausalLM" else 0
        if n_exp and quant:
            raise ValueError("fp8 quantisation of mixture-of-experts models is not supported yet")
        # sliding-window attention is not implemented: a model that uses it is served with the
        # context capped at the window, where full attention is the same computation
        window = d.get("sliding_window") if d.get("use_sliding_window", True) else None
        return ModelConfig(
            quantization=quant, num_local_experts=n_exp, sliding_window=int(window) if window else None,
            num_experts_per_tok=int(d.get("num_experts_per_tok") or 0) if n_exp else 0,
            vocab_size=d["vocab_size"], hidden_size=d["hidden_size"],
            intermediate_size=d["intermediate_size"], num_hidden_layers=d["num_hidden_layers"],
            num_attention_heads=heads, num_key_value_heads=d.get("num_key_value_heads", heads),
            head_dim=d.get("head_dim") or d["hidden_size"] // heads,
            rms_norm_eps=d.get("rms_norm_eps", 1e-5), rope_theta=float(rope or 10000.0),
            max_position_embeddings=d.get("max_position_embeddings", 8192),
            tie_word_embeddings=bool(d.get("tie_word_embeddings", False)),
            attention_bias=bool(d.get("attention_bias", arch == "Qwen2ForCausalLM")),
            architecture=arch, eos_token_id=eos, bos_token_id=d.get("bos_token_id"),
            eos_token_ids=tuple(eos_all), rope_scaling=scaling)

    @staticmethod
    def from_pretrained(model_path: str) -> "ModelConfig":
        with open(Path(model_path) / "config.json") as f:
            cfg = ModelConfig.from_dict(json.load(f))
        # chat checkpoints list their turn terminators (<|eot_id|>, ...) in generation_config.json
        gen = Path(model_path) / "generation_config.json"
        if gen.exists():
            try:
                e = json.loads(gen.read_text()).get("eos_token_id")
            except (OSError, ValueError):
                e = None
            extra = [int(t) for t in e] if isinstance(e, list) else ([] if e is None else [int(e)])
            cfg.eos_token_ids = tuple(dict.fromkeys(list(cfg.eos_token_ids) + extra))
            if cfg.eos_token_id is None and cfg.eos_token_ids:
                cfg.eos_token_id = cfg.eos_token_ids[0]
        return cfg

    def rope_inv_freq(self):
        """Inverse RoPE frequencies [head_dim/2] (fp32 torch tensor), with the checkpoint's frequency
        scaling applied ($VLLM/model_executor/layers/rotary_embedding/llama3_rope.py,
        linear_scaling_rope.py; same arithmetic as transformers' ROPE_INIT_FUNCTIONS)."""
        import math

        import torch

        D = self.head_dim
        inv = 1.0 / (self.rope_theta ** (torch.arange(0, D, 2, dtype=torch.float) / D))
        sc = self.rope_scaling
        if not sc:
            return inv
        if sc["rope_type"] == "linear":
            return inv / sc["factor"]
        factor, lo, hi = sc["factor"], sc["low_freq_factor"], sc["high_freq_factor"]
        old = sc["original_max_position_embeddings"]
        low_wl, high_wl = old / lo, old / hi
        wl = 2 * math.pi / inv
        scaled = torch.where(wl > low_wl, inv / factor, inv)
        smooth = (old / wl - lo) / (hi - lo)
        mid = (1 - smooth) * inv / factor + smooth * inv
        is_mid = (wl >= high_wl) & (wl <= low_wl)
        return torch.where(is_mid, mid, scaled)

    def to_hf_dict(self) -> dict:
        extra = {"quantization_config": {"quant_method": "fp8", "activation_scheme": "dynamic"}} \
            if self.quantization == "fp8" else {}
        if self.num_local_experts:
            extra.update({"num_local_experts": self.num_local_experts, "num_experts_per_tok": self.num_experts_per_tok,
                          "sliding_window": None})
        model_type = {"MixtralForCausalLM": "mixtral", "Qwen2ForCausalLM": "qwen2", "MistralForCausalLM": "mistral"} \
            .get(self.architecture, "llama")
        return {
            **extra,
            "architectures": [self.architecture], "model_type": model_type, "vocab_size": self.vocab_size,
            "hidden_size": self.hidden_size, "intermediate_size": self.intermediate_size,
            "num_hidden_layers": self.num_hidden_layers, "num_attention_heads": self.num_attention_heads,
            "num_key_value_heads": self.num_key_value_heads, "head_dim": self.head_dim,
            "rms_norm_eps": self.rms_norm_eps, "rope_theta": self.rope_theta,
            "max_position_embeddings": self.max_position_embeddings,
            "tie_word_embeddings": self.tie_word_embeddings, "attention_bias": self.attention_bias,
            "hidden_act": "silu", "torch_dtype": "bfloat16", "bos_token_id": self.bos_token_id or 1,
            "eos_token_id": (list(self.eos_token_ids) if len(self.eos_token_ids) > 1 else
                             (self.eos_token_id if self.eos_token_id is not None else 2)),
            **({"rope_scaling": dict(self.rope_scaling)} if self.rope_scaling else {}),
        }


# Named shapes used by tests and bench (random-init; there are no checkpoints offline).
def llama3_8b() -> ModelConfig:
    return ModelConfig(vocab_size=128256, hidden_size=4096, intermediate_size=14336, num_hidden_layers=32,
                       num_attention_heads=32, num_key_value_heads=8, head_dim=128, rms_norm_eps=1e-5,
                       rope_theta=500000.0, max_position_embeddings=8192, eos_token_id=128001,
                       bos_token_id=128000)


def qwen2_05b() -> ModelConfig:
    """BASELINE.json configs[0] model shape (Qwen2-0.5B: q/k/v bias, tied embeddings, GQA group of 7)."""
    return ModelConfig(vocab_size=151936, hidden_size=896, intermediate_size=4864, num_hidden_layers=24,
                       num_attention_heads=14, num_key_value_heads=2, head_dim=64, rms_norm_eps=1e-6,
                       rope_theta=1e6, max_position_embeddings=32768, tie_word_embeddings=True,
                       attention_bias=True, architecture="Qwen2ForCausalLM", eos_token_id=151643,
                       bos_token_id=151643)


def llama3_70b() -> ModelConfig:
    return ModelConfig(vocab_size=128256, hidden_size=8192, intermediate_size=28672, num_hidden_layers=80,
                       num_attention_heads=64, num_key_value_heads=8, head_dim=128, rms_norm_eps=1e-5,
                       rope_theta=500000.0, max_position_embeddings=8192, eos_token_id=128001,
                       bos_token_id=128000)


def tiny_mixtral() -> ModelConfig:
    return ModelConfig(vocab_size=512, hidden_size=512, intermediate_size=256, num_hidden_layers=2,
                       num_attention_heads=8, num_key_value_heads=2, head_dim=64, rope_theta=1e6,
                       max_position_embeddings=512, eos_token_id=2, bos_token_id=1,
                       architecture="MixtralForCausalLM", num_local_experts=8, num_experts_per_tok=2)


def mixtral_8x7b() -> ModelConfig:
    return ModelConfig(vocab_size=32000, hidden_size=4096, intermediate_size=14336, num_hidden_layers=32,
                       num_attention_heads=32, num_key_value_heads=8, head_dim=128, rms_norm_eps=1e-5,
                       rope_theta=1e6, max_position_embeddings=32768, eos_token_id=2, bos_token_id=1,
                       architecture="MixtralForCausalLM", num_local_experts=8, num_experts_per_tok=2)


def tiny_llama(head_dim: int = 64) -> ModelConfig:
    if head_dim == 128:
        return ModelConfig(vocab_size=640, hidden_size=512, intermediate_size=1024, num_hidden_layers=2,
                           num_attention_heads=4, num_key_value_heads=1, head_dim=128,
                           max_position_embeddings=1024, eos_token_id=2, bos_token_id=1)
    return ModelConfig(vocab_size=512, hidden_size=512, intermediate_size=1024, num_hidden_layers=2,
                       num_attention_heads=8, num_key_value_heads=2, head_dim=64,
                       max_position_embeddings=512, eos_token_id=2, bos_token_id=1)


@dataclass
class EngineConfig:
    model: str = ""
    served_model_name: Optional[str] = None
    max_model_len: Optional[int] = None
    tensor_parallel_size: int = 1
    gpu_memory_utilization: float = 0.90
    max_num_seqs: int = 256
    max_num_batched_tokens: int = 8192
    num_gpu_blocks: Optional[int] = None      # override the memory-derived KV block count
    enforce_eager: bool = False               # skip CUDA-graph capture of decode steps
    enable_prefix_caching: bool = True        # reuse full KV blocks across requests (reference engine default)
    quantization: Optional[str] = None        # None | "fp8" (W8A8: online per-tensor weights, per-token activations)
    mixed_steps: bool = True                  # prefill steps also carry the running sequences' decode tokens
    kv_host_cache_gib: float = 0.0            # extended KV cache: pinned host memory for evicted prefix blocks (0 = off)
    # --speculative-config as a dict ({"method": "ngram", "num_speculative_tokens": k, ...}), None = off
    speculative_config: Optional[dict] = None
    seed: int = 0
    block_size: int = 16
    cuda_graph_sizes: list = field(default_factory=list)

    def speculative(self):
        """Parsed speculative config, None when speculation is off or the method needs a draft model."""
        from .spec_decode import SpeculativeConfig

        c = SpeculativeConfig.parse(self.speculative_config)
        return c if (c is not None and c.supported) else None

    @property
    def num_speculative_tokens(self) -> int:
        c = self.speculative()
        return c.num_speculative_tokens if c is not None else 0

    def graph_buckets(self) -> list:
        if self.cuda_graph_sizes:
            return sorted(set(min(b, self.max_num_seqs) for b in self.cuda_graph_sizes))
        b = [1, 2, 4, 8]
        b += list(range(16, 129, 8))
        b += list(range(144, 513, 16))
        b += list(range(544, 1025, 32))
        b = [x for x in b if x <= self.max_num_seqs]
        if self.max_num_seqs not in b:
            b.append(self.max_num_seqs)
        return sorted(set(b))
