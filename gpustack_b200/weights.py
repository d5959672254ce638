"""Weight loading: HF-format safetensors directory -> fused, tensor-parallel-sharded bf16 (or
fp8 W8A8) tensors.

Sharding follows the reference engine's Megatron layout (SURVEY.md section 8e;
$VLLM/model_executor/models/llama.py:81-347): QKV and gate_up are column-parallel (each rank owns
Hq/t query heads, Hkv/t kv heads and I/t MLP columns), O and down are row-parallel (in-dim / t),
lm_head and the embedding table are vocab-parallel (rows [rank*ceil(V/t), ...)), the norms are replicated.

fp8 (`--quantization fp8`, or a checkpoint whose config.json carries quantization_config
{quant_method: fp8, activation_scheme: dynamic}): the four decoder linears of every layer are held
as e4m3 with ONE fp32 scale per fused weight, exactly as the reference engine prepares them --
bf16 checkpoints are quantised per tensor after fusing and sharding
($VLLM/model_executor/layers/quantization/fp8.py:537-549), per-tensor fp8 checkpoints are
re-quantised to the largest shard scale of a fused weight (fp8.py:406-428,
requantize_with_max_scale).  Embedding, norms and lm_head stay bf16.
"""
from __future__ import annotations

import json
from dataclasses import dataclass
from pathlib import Path
from typing import Callable, Dict, List, Optional

import torch

from .config import ModelConfig

BF16 = torch.bfloat16
FP8 = torch.float8_e4m3fn
FP8_MAX = 448.0
FP8_MIN_SCALE = 1.0 / (448.0 * 512.0)


@dataclass
class LayerWeights:
    input_norm: torch.Tensor      # [H]
    qkv: torch.Tensor             # [(Hq+2Hkv)/t * D, H]
    qkv_bias: Optional[torch.Tensor]
    o: torch.Tensor               # [H, Hq/t * D]
    post_norm: torch.Tensor       # [H]
    gate_up: Optional[torch.Tensor]   # [2I/t, H]  rows interleaved: 2i = gate_i, 2i+1 = up_i (fused SiLU*mul epilogue)
    down: Optional[torch.Tensor]      # [H, I/t]
    # fp8 mode: the four matrices above are float8_e4m3fn and these hold their fp32 scales ([1])
    qkv_scale: Optional[torch.Tensor] = None
    o_scale: Optional[torch.Tensor] = None
    gate_up_scale: Optional[torch.Tensor] = None
    down_scale: Optional[torch.Tensor] = None
    # Mixtral (gate_up / down are None): router replicated, experts stacked along the rows and
    # tensor-parallel over I exactly like the dense MLP
    router: Optional[torch.Tensor] = None      # [128, H]: E router rows, zero-padded to a GEMM tile
    w13: Optional[torch.Tensor] = None         # [E * 2I/t, H]  per expert: interleaved gate (w1) / up (w3) rows
    w2: Optional[torch.Tensor] = None          # [E * H, I/t]


@dataclass
class ModelWeights:
    embed: torch.Tensor           # [V/t, H]: this rank's vocabulary rows (the whole table when t == 1)
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


def quantize_fp8_per_tensor(w: torch.Tensor, scale: Optional[torch.Tensor] = None):
    """(e4m3 weight, fp32 scale[1]); scale = max(absmax/448, 1/(448*512)) unless given.  Load-time
    torch ops on the weight's device (same arithmetic as ops.scaled_fp8_quant(w, scale=None))."""
    wf = w.to(torch.float32)
    if scale is None:
        scale = (wf.abs().max() / FP8_MAX).clamp(min=FP8_MIN_SCALE)
    scale = scale.to(torch.float32).reshape(1)
    q = (wf * scale.reciprocal()).clamp(-FP8_MAX, FP8_MAX).to(FP8)
    return q.contiguous(), scale


def shard_state_dict(cfg: ModelConfig, get: Callable[[str], torch.Tensor], rank: int, tp: int,
                     device: str, quantization: Optional[str] = None,
                     has: Optional[Callable[[str], bool]] = None) -> ModelWeights:
    """`get(name)` returns the full HF tensor (any device); the result lives on `device`.
    quantization: None (bf16) or "fp8"; `has(name)` tells whether the checkpoint holds a tensor
    (needed to find `<proj>.weight_scale` / `.input_scale` of fp8 checkpoints)."""
    check_tp(cfg, tp)
    if quantization not in (None, "fp8"):
        raise ValueError(f"unsupported quantization {quantization!r} (supported: fp8)")
    D = cfg.head_dim
    hq, hkv, I = cfg.num_attention_heads // tp, cfg.num_key_value_heads // tp, cfg.intermediate_size // tp
    fp8 = quantization == "fp8"

    def dev(t):
        return t.to(device=device, dtype=BF16).contiguous()

    def shard_scale(name: str) -> Optional[torch.Tensor]:
        """per-tensor weight_scale of an fp8 checkpoint tensor, None for a bf16 one."""
        sname = name[: -len("weight")] + "weight_scale"
        if has is None or not has(sname):
            return None
        if has(name[: -len("weight")] + "input_scale"):
            raise ValueError("fp8 checkpoints with static activation scales are not supported "
                             "(activation_scheme must be dynamic)")
        sc = get(sname).to(torch.float32)
        if sc.numel() != 1:
            raise ValueError(f"{sname}: only per-tensor fp8 weight scales are supported")
        return sc.reshape(1).to(device)

    def fuse(names, slicers, interleave: bool = False):
        """Fused weight of this rank from the checkpoint tensors `names`, each cut by its slicer.
        bf16 mode -> (bf16 tensor, None); fp8 mode -> (e4m3 tensor, fp32 scale[1])."""
        parts = [sl(get(n)).to(device) for n, sl in zip(names, slicers)]
        scales = [shard_scale(n) for n in names] if fp8 else [None] * len(names)

        def merge(ts):
            if interleave:
                return torch.stack(ts, dim=1).reshape(2 * ts[0].shape[0], -1)
            return torch.cat(ts, dim=0) if len(ts) > 1 else ts[0]

        if not fp8:
            return dev(merge(parts)), None
        if all(sc is None for sc in scales):                      # online quantisation of a bf16 checkpoint
            return quantize_fp8_per_tensor(merge([p.to(BF16) for p in parts]))
        if any(sc is None for sc in scales) or any(p.dtype != FP8 for p in parts):
            raise ValueError(f"{names}: mixed fp8 / bf16 shards in one fused weight")
        if len(parts) == 1:
            return parts[0].contiguous(), scales[0]
        smax = torch.stack(scales).max().reshape(1)                # requantize_with_max_scale
        return quantize_fp8_per_tensor(merge([p.to(torch.float32) * sc for p, sc in zip(parts, scales)]), smax)

    layers = []
    for l in range(cfg.num_hidden_layers):
        p = f"model.layers.{l}."
        qsl = lambda t: t[rank * hq * D:(rank + 1) * hq * D]          # noqa: E731
        ksl = lambda t: t[rank * hkv * D:(rank + 1) * hkv * D]        # noqa: E731
        isl = lambda t: t[rank * I:(rank + 1) * I]                    # noqa: E731
        bias = None
        if cfg.attention_bias:
            bias = dev(torch.cat([qsl(get(p + "self_attn.q_proj.bias")), ksl(get(p + "self_attn.k_proj.bias")),
                                  ksl(get(p + "self_attn.v_proj.bias"))]))
        qkv_w, qkv_s = fuse([p + f"self_attn.{n}_proj.weight" for n in "qkv"], [qsl, ksl, ksl])
        o_w, o_s = fuse([p + "self_attn.o_proj.weight"], [lambda t: t[:, rank * hq * D:(rank + 1) * hq * D]])
        moe = {}
        gu_w = gu_s = down_w = down_s = None
        if cfg.num_local_experts:
            if fp8:
                raise ValueError("fp8 quantisation of mixture-of-experts models is not supported yet")
            E = cfg.num_local_experts
            m = p + "block_sparse_moe."
            if has is not None and not has(m + "gate.weight") and has(p + "mlp.experts.gate_up_proj"):
                # transformers >= 5 export: experts fused as [E, 2I, H] (gate rows, then up rows) / [E, H, I]
                gup, dn = get(p + "mlp.experts.gate_up_proj"), get(p + "mlp.experts.down_proj")
                Ifull = cfg.intermediate_size
                w1 = [gup[e, :Ifull] for e in range(E)]
                w3 = [gup[e, Ifull:] for e in range(E)]
                w2 = [dn[e] for e in range(E)]
                router = get(p + "mlp.gate.weight")
            else:
                w1 = [get(m + f"experts.{e}.w1.weight") for e in range(E)]
                w3 = [get(m + f"experts.{e}.w3.weight") for e in range(E)]
                w2 = [get(m + f"experts.{e}.w2.weight") for e in range(E)]
                router = get(m + "gate.weight")
            w13 = torch.cat([torch.stack([isl(w1[e]).to(device), isl(w3[e]).to(device)], dim=1).reshape(2 * I, -1)
                             for e in range(E)], dim=0)
            router_pad = torch.zeros(128, router.shape[1], dtype=BF16)     # one full GEMM weight tile
            router_pad[:E] = router.to(BF16)
            moe = dict(router=dev(router_pad), w13=dev(w13),
                       w2=dev(torch.cat([w2[e][:, rank * I:(rank + 1) * I].to(device) for e in range(E)], dim=0)))
        else:
            gu_w, gu_s = fuse([p + "mlp.gate_proj.weight", p + "mlp.up_proj.weight"], [isl, isl], interleave=True)
            down_w, down_s = fuse([p + "mlp.down_proj.weight"], [lambda t: t[:, rank * I:(rank + 1) * I]])
        layers.append(LayerWeights(
            input_norm=dev(get(p + "input_layernorm.weight")),
            qkv=qkv_w, qkv_bias=bias, o=o_w,
            post_norm=dev(get(p + "post_attention_layernorm.weight")),
            gate_up=gu_w, down=down_w,
            qkv_scale=qkv_s, o_scale=o_s, gate_up_scale=gu_s, down_scale=down_s, **moe))
    V = cfg.vocab_size
    per = (V + tp - 1) // tp
    v0, v1 = rank * per, min(V, (rank + 1) * per)
    # vocab-parallel embedding (VocabParallelEmbedding of the reference engine): each rank looks up
    # only its own rows, the partial result is summed by the first fused all-reduce + norm of the step
    embed = dev(get("model.embed_tokens.weight")[v0:v1])
    # tied embeddings: the lm_head shard IS the embedding shard (no second copy)
    lm = embed if cfg.tie_word_embeddings else dev(get("lm_head.weight")[v0:v1])
    return ModelWeights(embed=embed, layers=layers, final_norm=dev(get("model.norm.weight")),
                        lm_head=lm, vocab_start=v0, vocab_end=v1)


def load_safetensors_dir(cfg: ModelConfig, model_path: str, rank: int, tp: int, device: str,
                         quantization: Optional[str] = None) -> ModelWeights:
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

    return shard_state_dict(cfg, get, rank, tp, device, quantization=quantization,
                            has=lambda name: name in weight_map)


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
        if cfg.num_local_experts:
            w[p + "block_sparse_moe.gate.weight"] = rnd(cfg.num_local_experts, H, s=0.2)
            for e in range(cfg.num_local_experts):
                w[p + f"block_sparse_moe.experts.{e}.w1.weight"] = rnd(I, H)
                w[p + f"block_sparse_moe.experts.{e}.w3.weight"] = rnd(I, H)
                w[p + f"block_sparse_moe.experts.{e}.w2.weight"] = rnd(H, I)
            continue
        w[p + "mlp.gate_proj.weight"] = rnd(I, H)
        w[p + "mlp.up_proj.weight"] = rnd(I, H)
        w[p + "mlp.down_proj.weight"] = rnd(H, I)
    w["model.norm.weight"] = (1.0 + rnd(H, s=0.1).float()).to(BF16)
    if not cfg.tie_word_embeddings:
        w["lm_head.weight"] = rnd(cfg.vocab_size, H)
    return w


def lazy_random_source(cfg: ModelConfig, seed: int = 0, std: float = 0.02, device: str = "cuda"):
    """(get, has) for shard_state_dict that synthesise each checkpoint tensor on demand instead of
    materialising the whole state dict first -- a 47 B-parameter Mixtral would otherwise be resident
    twice while its shards are fused.  Every tensor is N(0, std) (router 10x wider, norms 1 + N(0, 0.1))
    from a generator seeded by a hash of (seed, name), so any rank / any call order sees the same values."""
    import zlib

    H, I = cfg.hidden_size, cfg.intermediate_size
    shapes: Dict[str, tuple] = {"model.embed_tokens.weight": (cfg.vocab_size, H), "model.norm.weight": (H,)}
    if not cfg.tie_word_embeddings:
        shapes["lm_head.weight"] = (cfg.vocab_size, H)
    for l in range(cfg.num_hidden_layers):
        p = f"model.layers.{l}."
        shapes[p + "input_layernorm.weight"] = (H,)
        shapes[p + "post_attention_layernorm.weight"] = (H,)
        for n, rows in (("q", cfg.q_size), ("k", cfg.kv_size), ("v", cfg.kv_size)):
            shapes[p + f"self_attn.{n}_proj.weight"] = (rows, H)
            if cfg.attention_bias:
                shapes[p + f"self_attn.{n}_proj.bias"] = (rows,)
        shapes[p + "self_attn.o_proj.weight"] = (H, cfg.q_size)
        if cfg.num_local_experts:
            shapes[p + "block_sparse_moe.gate.weight"] = (cfg.num_local_experts, H)
            for e in range(cfg.num_local_experts):
                shapes[p + f"block_sparse_moe.experts.{e}.w1.weight"] = (I, H)
                shapes[p + f"block_sparse_moe.experts.{e}.w3.weight"] = (I, H)
                shapes[p + f"block_sparse_moe.experts.{e}.w2.weight"] = (H, I)
        else:
            shapes[p + "mlp.gate_proj.weight"] = (I, H)
            shapes[p + "mlp.up_proj.weight"] = (I, H)
            shapes[p + "mlp.down_proj.weight"] = (H, I)

    def get(name: str) -> torch.Tensor:
        g = torch.Generator(device=device)
        g.manual_seed(zlib.crc32(f"{seed}/{name}".encode()) | (zlib.crc32(f"{name}#{seed}".encode()) << 32))
        shape = shapes[name]
        x = torch.randn(*shape, generator=g, device=device, dtype=torch.float32)
        if name.endswith("norm.weight"):
            return (1.0 + x * 0.1).to(BF16)
        return (x * (10 * std if name.endswith("block_sparse_moe.gate.weight") else std)).to(BF16)

    return get, (lambda name: name in shapes)


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
