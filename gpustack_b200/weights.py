"""Weight loading: HF-format safetensors directory -> fused, tensor-parallel-sharded bf16 tensors.

Sharding follows the reference engine's Megatron layout (SURVEY.md section 8e;
$VLLM/model_executor/models/llama.py:81-347): QKV and gate_up are column-parallel (each rank owns
Hq/t query heads, Hkv/t kv heads and I/t MLP columns), O and down are row-parallel (in-dim / t),
lm_head is vocab-parallel, the embedding table and the norms are replicated.
"""
from __future__ import annotations

import json
from dataclasses import dataclass
from pathlib import Path
from typing import Callable, Dict, List, Optional

import torch

from .config import ModelConfig

BF16 = torch.bfloat16


@dataclass
class LayerWeights:
    input_norm: torch.Tensor      # [H]
    qkv: torch.Tensor             # [(Hq+2Hkv)/t * D, H]
    qkv_bias: Optional[torch.Tensor]
    o: torch.Tensor               # [H, Hq/t * D]
    post_norm: torch.Tensor       # [H]
    gate_up: torch.Tensor         # [2I/t, H]  rows interleaved: 2i = gate_i, 2i+1 = up_i (fused SiLU*mul epilogue)
    down: torch.Tensor            # [H, I/t]


@dataclass
class ModelWeights:
    embed: torch.Tensor           # [V, H] replicated
    layers: List[LayerWeights]
    final_norm: torch.Tensor      # [H]
    lm_head: torch.Tensor         # [V/t (padded to a multiple of 8), H]
    vocab_start: int
    vocab_end: int


def check_tp(cfg: ModelConfig, tp: int) -> None:
    # same divisibility rule the reference scheduler enforces
    # (gpustack/policies/candidate_selectors/base_candidate_selector.py:1007-1082)
    if cfg.num_attention_heads % tp or cfg.num_key_value_heads % tp:
        raise ValueError(f"tensor_parallel_size={tp} must divide attention heads "
                         f"({cfg.num_attention_heads}) and kv heads ({cfg.num_key_value_heads})")
    if cfg.intermediate_size % (tp * 64):
        raise ValueError(f"intermediate_size {cfg.intermediate_size} not divisible by 64*tp")


def shard_state_dict(cfg: ModelConfig, get: Callable[[str], torch.Tensor], rank: int, tp: int,
                     device: str) -> ModelWeights:
    """`get(name)` returns the full HF tensor (any device); the result lives on `device`."""
    check_tp(cfg, tp)
    D = cfg.head_dim
    hq, hkv, I = cfg.num_attention_heads // tp, cfg.num_key_value_heads // tp, cfg.intermediate_size // tp

    def dev(t):
        return t.to(device=device, dtype=BF16).contiguous()

    layers = []
    for l in range(cfg.num_hidden_layers):
        p = f"model.layers.{l}."
        q = get(p + "self_attn.q_proj.weight")[rank * hq * D:(rank + 1) * hq * D]
        k = get(p + "self_attn.k_proj.weight")[rank * hkv * D:(rank + 1) * hkv * D]
        v = get(p + "self_attn.v_proj.weight")[rank * hkv * D:(rank + 1) * hkv * D]
        bias = None
        if cfg.attention_bias:
            bias = dev(torch.cat([get(p + "self_attn.q_proj.bias")[rank * hq * D:(rank + 1) * hq * D],
                                  get(p + "self_attn.k_proj.bias")[rank * hkv * D:(rank + 1) * hkv * D],
                                  get(p + "self_attn.v_proj.bias")[rank * hkv * D:(rank + 1) * hkv * D]]))
        g = get(p + "mlp.gate_proj.weight")[rank * I:(rank + 1) * I]
        u = get(p + "mlp.up_proj.weight")[rank * I:(rank + 1) * I]
        layers.append(LayerWeights(
            input_norm=dev(get(p + "input_layernorm.weight")),
            qkv=dev(torch.cat([q, k, v], dim=0)), qkv_bias=bias,
            o=dev(get(p + "self_attn.o_proj.weight")[:, rank * hq * D:(rank + 1) * hq * D]),
            post_norm=dev(get(p + "post_attention_layernorm.weight")),
            gate_up=dev(torch.stack([g, u], dim=1).reshape(2 * I, -1)),
            down=dev(get(p + "mlp.down_proj.weight")[:, rank * I:(rank + 1) * I])))
    embed = dev(get("model.embed_tokens.weight"))
    V = cfg.vocab_size
    per = (V + tp - 1) // tp
    v0, v1 = rank * per, min(V, (rank + 1) * per)
    lm_full = embed if cfg.tie_word_embeddings else get("lm_head.weight")
    lm = dev(lm_full[v0:v1]) if (tp > 1 or not cfg.tie_word_embeddings) else embed
    return ModelWeights(embed=embed, layers=layers, final_norm=dev(get("model.norm.weight")),
                        lm_head=lm, vocab_start=v0, vocab_end=v1)


def load_safetensors_dir(cfg: ModelConfig, model_path: str, rank: int, tp: int, device: str) -> ModelWeights:
    from safetensors import safe_open

    root = Path(model_path)
    index = root / "model.safetensors.index.json"
    if index.exists():
        weight_map = json.loads(index.read_text())["weight_map"]
    else:
        weight_map = {}
        for f in sorted(root.glob("*.safetensors")):
            with safe_open(str(f), framework="pt") as sf:
                for k in sf.keys():
                    weight_map[k] = f.name
    if not weight_map:
        raise FileNotFoundError(f"no *.safetensors files under {model_path}")
    handles: Dict[str, object] = {}

    def get(name: str) -> torch.Tensor:
        fname = weight_map[name]
        if fname not in handles:
            handles[fname] = safe_open(str(root / fname), framework="pt")
        return handles[fname].get_tensor(name)

    return shard_state_dict(cfg, get, rank, tp, device)


def random_state_dict(cfg: ModelConfig, seed: int = 0, std: float = 0.02, device: str = "cuda") -> Dict[str, torch.Tensor]:
    """Seeded N(0, std) bf16 weights in HF naming (synthetic checkpoints for benchmarks: no real
    weights exist offline).  Norm weights are 1 + N(0, 0.1)."""
    g = torch.Generator(device=device)
    g.manual_seed(seed)

    def rnd(*shape, s=std):
        return (torch.randn(*shape, generator=g, device=device, dtype=torch.float32) * s).to(BF16)

    H, I = cfg.hidden_size, cfg.intermediate_size
    w = {"model.embed_tokens.weight": rnd(cfg.vocab_size, H)}
    for l in range(cfg.num_hidden_layers):
        p = f"model.layers.{l}."
        w[p + "input_layernorm.weight"] = (1.0 + rnd(H, s=0.1).float()).to(BF16)
        w[p + "self_attn.q_proj.weight"] = rnd(cfg.q_size, H)
        w[p + "self_attn.k_proj.weight"] = rnd(cfg.kv_size, H)
        w[p + "self_attn.v_proj.weight"] = rnd(cfg.kv_size, H)
        if cfg.attention_bias:
            for n, sz in (("q", cfg.q_size), ("k", cfg.kv_size), ("v", cfg.kv_size)):
                w[p + f"self_attn.{n}_proj.bias"] = rnd(sz)
        w[p + "self_attn.o_proj.weight"] = rnd(H, cfg.q_size)
        w[p + "post_attention_layernorm.weight"] = (1.0 + rnd(H, s=0.1).float()).to(BF16)
        w[p + "mlp.gate_proj.weight"] = rnd(I, H)
        w[p + "mlp.up_proj.weight"] = rnd(I, H)
        w[p + "mlp.down_proj.weight"] = rnd(H, I)
    w["model.norm.weight"] = (1.0 + rnd(H, s=0.1).float()).to(BF16)
    if not cfg.tie_word_embeddings:
        w["lm_head.weight"] = rnd(cfg.vocab_size, H)
    return w


def save_checkpoint(cfg: ModelConfig, state: Dict[str, torch.Tensor], out_dir: str, shard_bytes: int = 4 << 30) -> None:
    """Write an HF-format directory (config.json + sharded safetensors + index) so that the same
    synthetic weights can be served by this engine and by the reference's vLLM backend."""
    from safetensors.torch import save_file

    out = Path(out_dir)
    out.mkdir(parents=True, exist_ok=True)
    (out / "config.json").write_text(json.dumps(cfg.to_hf_dict(), indent=1))
    shards, cur, cur_bytes = [], {}, 0
    for k, v in state.items():
        n = v.numel() * v.element_size()
        if cur and cur_bytes + n > shard_bytes:
            shards.append(cur)
            cur, cur_bytes = {}, 0
        cur[k] = v.detach().to("cpu").contiguous()
        cur_bytes += n
    if cur:
        shards.append(cur)
    weight_map = {}
    for i, sh in enumerate(shards):
        name = f"model-{i + 1:05d}-of-{len(shards):05d}.safetensors"
        save_file(sh, str(out / name), metadata={"format": "pt"})
        for k in sh:
            weight_map[k] = name
    (out / "model.safetensors.index.json").write_text(
        json.dumps({"metadata": {}, "weight_map": weight_map}, indent=1))
