"""CPU oracle for the decoder hot path -- TEST INFRASTRUCTURE, NOT PRODUCT CODE.

Only tests/, __graft_entry__.smoke() and bench.py's cpu_baseline / --impl reference leg may import
this module.  The engine (gpustack_b200/) never does: it has no CPU path.

GPUStack itself contains no inference arithmetic; the arithmetic of "the
reference's vLLM backend" lives in third-party vLLM (pinned vllm==0.10.1.1 in the reference's
pyproject.toml:97, 0.22.0 installed here).  Each function below restates one operator of that
engine in plain PyTorch-CPU fp32 with bf16 rounding at the points where the engine stores bf16,
citing the vLLM file it follows ($VLLM = site-packages/vllm).

Pinning: tests/golden/make_golden.py imports vLLM's own pure-PyTorch implementations
(vllm.ir.ops.layernorm, RotaryEmbedding.forward_static, SiluAndMul.forward_native) and HF
transformers' LlamaForCausalLM in the build container and commits their outputs as fixtures;
tests/test_oracle_golden.py checks this oracle against them.  The reference publishes no golden
vectors for inference, so those generated fixtures are the pin.
"""
from __future__ import annotations

import math
from dataclasses import dataclass
from typing import Dict, List, Optional, Sequence

import torch

BF16 = torch.bfloat16


def _bf(x: torch.Tensor) -> torch.Tensor:
    """round to bf16 and come back to fp32 (models a bf16 store)."""
    return x.to(BF16).to(torch.float32)


# ------------------------------------------------------------------------------- operators
def rms_norm(x: torch.Tensor, w: torch.Tensor, eps: float) -> torch.Tensor:
    """$VLLM/ir/ops/layernorm.py:10-21 -- fp32 stats, cast to weight dtype, multiply."""
    xf = x.to(torch.float32)
    var = xf.pow(2).mean(dim=-1, keepdim=True)
    xn = xf * torch.rsqrt(var + eps)
    return (xn.to(w.dtype) * w).to(x.dtype)


def fused_add_rms_norm(x, residual, w, eps):
    """$VLLM/ir/ops/layernorm.py:40-58 -- returns (normed, new_residual)."""
    z = x.to(torch.float32) + residual.to(torch.float32)
    new_res = z.to(x.dtype)
    var = z.pow(2).mean(dim=-1, keepdim=True)
    zn = z * torch.rsqrt(var + eps)
    return (zn.to(w.dtype) * w).to(x.dtype), new_res


def cos_sin_cache(head_dim: int, max_pos: int, base: float, dtype=BF16) -> torch.Tensor:
    """$VLLM/model_executor/layers/rotary_embedding/base.py:71-93 -- [max_pos, head_dim] = cos|sin."""
    inv_freq = 1.0 / (base ** (torch.arange(0, head_dim, 2, dtype=torch.float) / head_dim))
    t = torch.arange(max_pos, dtype=torch.float)
    freqs = torch.einsum("i,j -> ij", t, inv_freq)
    return torch.cat((freqs.cos(), freqs.sin()), dim=-1).to(dtype)


def rope_neox(x: torch.Tensor, positions: torch.Tensor, cache: torch.Tensor) -> torch.Tensor:
    """$VLLM/model_executor/layers/rotary_embedding/base.py:150-180 + common.py:144-183.
    x [T, heads, D]; cos/sin come from the cache in the model dtype; the rotation itself is done
    here in fp32 and rounded once (the engine's kernel does the same; vLLM's native path rounds
    every bf16 op, which differs by < 2 bf16 ulp -- covered by the stated tolerance)."""
    D = x.shape[-1]
    cs = cache.index_select(0, positions.long()).to(torch.float32)
    cos, sin = cs[:, : D // 2].unsqueeze(1), cs[:, D // 2 :].unsqueeze(1)
    xf = x.to(torch.float32)
    x1, x2 = xf[..., : D // 2], xf[..., D // 2 :]
    return torch.cat((x1 * cos - x2 * sin, x2 * cos + x1 * sin), dim=-1).to(x.dtype)


def silu_and_mul(x: torch.Tensor) -> torch.Tensor:
    """$VLLM/model_executor/layers/activation.py:133-143 (forward_native): silu in the model dtype,
    then the product."""
    d = x.shape[-1] // 2
    return torch.nn.functional.silu(x[..., :d]) * x[..., d:]


def linear(x: torch.Tensor, w: torch.Tensor, b: Optional[torch.Tensor] = None) -> torch.Tensor:
    """$VLLM/model_executor/layers/utils.py:92-98 -- F.linear: fp32 accumulation (+ bias in fp32),
    one rounding to bf16."""
    y = x.to(torch.float32) @ w.to(torch.float32).t()
    if b is not None:
        y = y + b.to(torch.float32)
    return y.to(x.dtype)


# ------------------------------------------------------------------------------- fp8 (W8A8)
FP8 = torch.float8_e4m3fn
FP8_MAX = 448.0
FP8_MIN_SCALE = 1.0 / (448.0 * 512.0)


def quant_fp8(x: torch.Tensor, per_token: bool = True):
    """$VLLM/model_executor/layers/quantization/input_quant_fp8.py:188-235 (QuantFP8.forward_native,
    dynamic): scale = max(absmax / 448, 1/(448*512)) per token row (activations) or over the whole
    tensor (online weight quantisation, fp8.py:537-549); q = e4m3(clamp(float(x) * (1/scale)))."""
    if per_token:
        amax = x.abs().max(dim=-1, keepdim=True)[0].to(torch.float32)
    else:
        amax = x.abs().max().to(torch.float32).reshape(1)
    scale = (amax / FP8_MAX).clamp(min=FP8_MIN_SCALE)
    q = (x.to(torch.float32) * scale.reciprocal()).clamp(-FP8_MAX, FP8_MAX).to(FP8)
    return q, scale


def fp8_linear(x: torch.Tensor, wq: torch.Tensor, w_scale: torch.Tensor, b: Optional[torch.Tensor] = None) -> torch.Tensor:
    """Fp8LinearMethod.apply -> cutlass_scaled_mm ($VLLM/.../quantization/fp8.py:437-470): dynamic
    per-token activation scales, e4m3 x e4m3 products accumulated in fp32, multiplied by both
    scales (+ bias) in fp32, one rounding to bf16.  `wq` may be e4m3 or its exact fp32 image."""
    xq, xs = quant_fp8(x, per_token=True)
    y = (xq.to(torch.float32) @ wq.to(torch.float32).t()) * xs * w_scale.to(torch.float32)
    if b is not None:
        y = y + b.to(torch.float32)
    return y.to(x.dtype)


def moe_topk(probs: torch.Tensor, k: int):
    """top-k of each row, ties to the lowest index (what the reference's topk_softmax kernel and
    this repo's moe_route_kernel do; torch.topk leaves tie order unspecified)."""
    p = probs.clone()
    vals, idx = [], []
    for _ in range(k):
        v, i = p.max(dim=-1)          # max returns the first maximal index on CPU
        vals.append(v); idx.append(i)
        p.scatter_(1, i.unsqueeze(1), -1.0)
    return torch.stack(vals, dim=1), torch.stack(idx, dim=1)


def attention_ref(q, k, v, scale: float, q_pos: torch.Tensor) -> torch.Tensor:
    """Causal GQA attention over a contiguous K/V history.
    q [Tq, Hq, D], k/v [Tk, Hkv, D]; query i may see keys 0..q_pos[i].  fp32 softmax.
    ($VLLM/model_executor/models/llama.py:161 scale = head_dim**-0.5; flashinfer.py:1095-1230.)"""
    Tq, Hq, D = q.shape
    Tk, Hkv, _ = k.shape
    G = Hq // Hkv
    qf = q.to(torch.float32).view(Tq, Hkv, G, D)
    kf, vf = k.to(torch.float32), v.to(torch.float32)
    s = torch.einsum("thgd,khd->thgk", qf, kf) * scale
    mask = torch.arange(Tk).view(1, 1, 1, Tk) > q_pos.view(Tq, 1, 1, 1)
    s = s.masked_fill(mask, float("-inf"))
    p = torch.softmax(s, dim=-1)
    o = torch.einsum("thgk,khd->thgd", p, vf)
    return o.reshape(Tq, Hq, D).to(q.dtype)


def gather_paged(cache: torch.Tensor, block_table_row: Sequence[int], n_tokens: int) -> torch.Tensor:
    """cache [num_blocks, Hkv, block, D] -> contiguous [n_tokens, Hkv, D] for one sequence."""
    bs = cache.shape[2]
    pages = [cache[int(b)] for b in block_table_row[: (n_tokens + bs - 1) // bs]]
    kv = torch.cat(pages, dim=1).transpose(0, 1)  # [tokens, Hkv, D]
    return kv[:n_tokens]


# ----------------------------------------------------------------------------------- model
@dataclass
class LlamaConfig:
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
    attention_bias: bool = False      # Qwen2: bias on q/k/v projections
    num_local_experts: int = 0        # Mixtral: experts per MoE layer (0 = dense MLP)
    num_experts_per_tok: int = 0

    @property
    def q_size(self):
        return self.num_attention_heads * self.head_dim

    @property
    def kv_size(self):
        return self.num_key_value_heads * self.head_dim


TINY = LlamaConfig(vocab_size=512, hidden_size=512, intermediate_size=1024, num_hidden_layers=2,
                   num_attention_heads=8, num_key_value_heads=2, head_dim=64, max_position_embeddings=512)
TINY128 = LlamaConfig(vocab_size=640, hidden_size=512, intermediate_size=1024, num_hidden_layers=2,
                      num_attention_heads=4, num_key_value_heads=1, head_dim=128,
                      max_position_embeddings=1024)
QWEN2_05B_SHAPE = LlamaConfig(vocab_size=151936, hidden_size=896, intermediate_size=4864,
                              num_hidden_layers=24, num_attention_heads=14, num_key_value_heads=2,
                              head_dim=64, rms_norm_eps=1e-6, rope_theta=1e6,
                              max_position_embeddings=32768, tie_word_embeddings=True, attention_bias=True)
TINY_QWEN2 = LlamaConfig(vocab_size=1000, hidden_size=896, intermediate_size=1024, num_hidden_layers=2,
                         num_attention_heads=14, num_key_value_heads=2, head_dim=64, rms_norm_eps=1e-6,
                         rope_theta=1e6, max_position_embeddings=512, tie_word_embeddings=True,
                         attention_bias=True)
TINY_MIXTRAL = LlamaConfig(vocab_size=512, hidden_size=512, intermediate_size=256, num_hidden_layers=2,
                           num_attention_heads=8, num_key_value_heads=2, head_dim=64, rope_theta=1e6,
                           max_position_embeddings=512, num_local_experts=8, num_experts_per_tok=2)
LLAMA3_8B = LlamaConfig(vocab_size=128256, hidden_size=4096, intermediate_size=14336,
                        num_hidden_layers=32, num_attention_heads=32, num_key_value_heads=8, head_dim=128)


def init_weights(cfg: LlamaConfig, seed: int = 0, std: float = 0.02, device="cpu") -> Dict[str, torch.Tensor]:
    """Seeded N(0, std) bf16 weights in HF naming, norm weights 1 + N(0, 0.1) so the norm multiply
    is exercised.  Generated layer by layer on `device` (use the GPU for the 8B shape)."""
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
            w[p + "self_attn.q_proj.bias"] = rnd(cfg.q_size, s=0.5)
            w[p + "self_attn.k_proj.bias"] = rnd(cfg.kv_size, s=0.5)
            w[p + "self_attn.v_proj.bias"] = rnd(cfg.kv_size, s=0.5)
        w[p + "self_attn.o_proj.weight"] = rnd(H, cfg.q_size)
        w[p + "post_attention_layernorm.weight"] = (1.0 + rnd(H, s=0.1).float()).to(BF16)
        if cfg.num_local_experts:
            # classic Mixtral checkpoint naming: w1 = gate, w3 = up, w2 = down; the router weight is
            # drawn wider so that the top-k choice is well separated for most tokens
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


class OracleLlama:
    """Restates LlamaModel.forward / LlamaDecoderLayer.forward
    ($VLLM/model_executor/models/llama.py:316-333, 395-434) with a plain contiguous KV history per
    sequence.  bf16 tensors between operators, fp32 inside them."""

    def __init__(self, cfg: LlamaConfig, weights: Dict[str, torch.Tensor], fp32_cache: bool = False,
                 quantization: Optional[str] = None):
        """fp32_cache=True keeps an fp32 copy of every matmul weight (same arithmetic, no per-call
        upcast) -- used when this oracle is timed as the CPU baseline.
        quantization="fp8": the decoder linears run W8A8 as the reference engine's online fp8 mode
        does (`--quantization fp8`): each MERGED weight (qkv, gate_up) and o / down gets one
        per-tensor scale (fp8.py:537-549), activations per-token dynamic scales; lm_head stays bf16."""
        self.cfg = cfg
        self.w = {k: v.to("cpu") for k, v in weights.items()}
        self.q8: Dict[str, tuple] = {}
        if quantization == "fp8":
            for l in range(cfg.num_hidden_layers):
                p = f"model.layers.{l}."
                for group in (("self_attn.q_proj", "self_attn.k_proj", "self_attn.v_proj"), ("self_attn.o_proj",),
                              ("mlp.gate_proj", "mlp.up_proj"), ("mlp.down_proj",)):
                    names = [p + g + ".weight" for g in group]
                    merged = torch.cat([self.w[n] for n in names], dim=0)
                    mq, ms = quant_fp8(merged, per_token=False)
                    r0 = 0
                    for n in names:
                        rows = self.w[n].shape[0]
                        self.q8[n] = (mq[r0:r0 + rows].to(torch.float32), ms)
                        r0 += rows
        elif quantization is not None:
            raise ValueError(f"unsupported quantization {quantization!r}")
        self.wf = {k: v.to(torch.float32) for k, v in self.w.items() if v.dim() == 2 and "embed" not in k} \
            if fp32_cache else {}
        self.cache = cos_sin_cache(cfg.head_dim, cfg.max_position_embeddings, cfg.rope_theta)
        self.lm_head = self.w.get("lm_head.weight", self.w["model.embed_tokens.weight"])
        self.kv: Dict[int, List] = {}

    def reset(self):
        self.kv = {}

    def forward(self, seq_id: int, tokens: Sequence[int], last_only: bool = False) -> torch.Tensor:
        """Append `tokens` to sequence `seq_id`; returns fp32 logits [len(tokens), vocab]
        (only the last position when last_only)."""
        cfg, w = self.cfg, self.w
        hist = self.kv.setdefault(seq_id, [None] * cfg.num_hidden_layers)
        start = 0 if hist[0] is None else hist[0][0].shape[0]
        T = len(tokens)
        pos = torch.arange(start, start + T)
        h = w["model.embed_tokens.weight"][torch.tensor(list(tokens))]
        res = None
        scale = cfg.head_dim ** -0.5
        for l in range(cfg.num_hidden_layers):
            p = f"model.layers.{l}."
            if res is None:
                res, x = h, rms_norm(h, w[p + "input_layernorm.weight"], cfg.rms_norm_eps)
            else:
                x, res = fused_add_rms_norm(h, res, w[p + "input_layernorm.weight"], cfg.rms_norm_eps)
            q = self._lin(x, p + "self_attn.q_proj.weight").view(T, cfg.num_attention_heads, cfg.head_dim)
            k = self._lin(x, p + "self_attn.k_proj.weight").view(T, cfg.num_key_value_heads, cfg.head_dim)
            v = self._lin(x, p + "self_attn.v_proj.weight").view(T, cfg.num_key_value_heads, cfg.head_dim)
            q, k = rope_neox(q, pos, self.cache), rope_neox(k, pos, self.cache)
            if hist[l] is None:
                hist[l] = (k, v)
            else:
                hist[l] = (torch.cat((hist[l][0], k)), torch.cat((hist[l][1], v)))
            a = attention_ref(q, hist[l][0], hist[l][1], scale, pos).reshape(T, cfg.q_size)
            h = self._lin(a, p + "self_attn.o_proj.weight")
            x, res = fused_add_rms_norm(h, res, w[p + "post_attention_layernorm.weight"], cfg.rms_norm_eps)
            if cfg.num_local_experts:
                h = self._moe(x, p + "block_sparse_moe.")
                continue
            gu = torch.cat((self._lin(x, p + "mlp.gate_proj.weight"), self._lin(x, p + "mlp.up_proj.weight")), dim=-1)
            h = self._lin(silu_and_mul(gu), p + "mlp.down_proj.weight")
        x, _ = fused_add_rms_norm(h, res, w["model.norm.weight"], cfg.rms_norm_eps)
        if last_only:
            x = x[-1:]
        lm = self.wf.get("lm_head.weight")
        return x.to(torch.float32) @ (lm if lm is not None else self.lm_head.to(torch.float32)).t()

    def _moe(self, x: torch.Tensor, p: str) -> torch.Tensor:
        """MixtralMoE.forward ($VLLM/model_executor/models/mixtral.py) -> FusedMoE: the router is a
        plain linear in the model dtype; fused_topk takes softmax in fp32, the top-k (ties -> lowest
        index) and renormalises; fused_experts computes silu(w1 x) * w3 x and w2 (.) per expert with
        the routing weight applied to the fp32 accumulator before the bf16 store
        (MUL_ROUTED_WEIGHT), and moe_sum adds the k rows of a token."""
        cfg = self.cfg
        logits = self._lin(x, p + "gate.weight")                      # bf16 [T, E]
        probs = torch.softmax(logits.to(torch.float32), dim=-1)
        topw, topi = moe_topk(probs, cfg.num_experts_per_tok)
        topw = topw / topw.sum(dim=-1, keepdim=True)
        out = torch.zeros(x.shape[0], cfg.hidden_size, dtype=torch.float32)
        for e in range(cfg.num_local_experts):
            tok, slot = torch.nonzero(topi == e, as_tuple=True)
            if tok.numel() == 0:
                continue
            xe = x[tok]
            gu = torch.cat((self._lin(xe, p + f"experts.{e}.w1.weight"), self._lin(xe, p + f"experts.{e}.w3.weight")), dim=-1)
            act = silu_and_mul(gu)
            w2 = self.wf.get(p + f"experts.{e}.w2.weight")
            w2 = w2 if w2 is not None else self.w[p + f"experts.{e}.w2.weight"].to(torch.float32)
            y = (act.to(torch.float32) @ w2.t()) * topw[tok, slot].unsqueeze(-1)
            out.index_add_(0, tok, _bf(y))                            # bf16 rows, summed in fp32
        return out.to(x.dtype)

    def _lin(self, x: torch.Tensor, name: str) -> torch.Tensor:
        wf = self.wf.get(name)
        b = self.w.get(name.replace(".weight", ".bias"))
        if name in self.q8:
            return fp8_linear(x, self.q8[name][0], self.q8[name][1], b)
        if wf is None:
            return linear(x, self.w[name], b)
        y = x.to(torch.float32) @ wf.t()
        if b is not None:
            y = y + b.to(torch.float32)
        return y.to(x.dtype)

    def generate(self, seq_id: int, prompt: Sequence[int], max_new_tokens: int) -> List[int]:
        """greedy decode ($VLLM/v1/sample/sampler.py: argmax over fp32 logits, lowest index wins)."""
        out: List[int] = []
        logits = self.forward(seq_id, prompt, last_only=True)[-1]
        for _ in range(max_new_tokens):
            nxt = int(torch.argmax(logits))
            out.append(nxt)
            if len(out) == max_new_tokens:
                break
            logits = self.forward(seq_id, [nxt], last_only=True)[-1]
        return out


def to_hf_state_dict(weights: Dict[str, torch.Tensor]) -> Dict[str, torch.Tensor]:
    return dict(weights)
