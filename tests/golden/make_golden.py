"""Generate the golden fixtures that pin oracle/llama_oracle.py.

Run in the BUILD container (no GPU needed):  python tests/golden/make_golden.py
Sources of truth imported here (never at test time on the GPU box):
  * vLLM 0.22.0's own pure-PyTorch operator definitions -- the engine GPUStack's vLLM backend
    launches (gpustack/worker/backends/vllm.py:403-407): vllm.ir.ops.layernorm.{rms_norm,
    fused_add_rms_norm}, RotaryEmbedding.forward_static (neox), SiluAndMul.forward_native;
  * HF transformers LlamaForCausalLM (eager attention) for whole-model logits and greedy tokens
    of the seeded tiny Llama configs.
Outputs: tests/golden/ops_vllm.pt, fp8_quant_vllm.pt, fp8_linear_vllm.pt, tiny_llama_hf.pt, tiny_qwen2_hf.pt,
tiny_mixtral_hf.pt (a few hundred KB).
"""
import sys
from pathlib import Path

import torch

ROOT = Path(__file__).resolve().parents[2]
sys.path.insert(0, str(ROOT))
from oracle import llama_oracle as O  # noqa: E402

OUT = Path(__file__).resolve().parent
BF = torch.bfloat16


def ops_from_vllm():
    from vllm.ir.ops.layernorm import fused_add_rms_norm, rms_norm
    from vllm.model_executor.layers.activation import SiluAndMul
    from vllm.model_executor.layers.rotary_embedding.base import RotaryEmbedding

    g = torch.Generator().manual_seed(1234)
    out = {}
    x = (torch.randn(6, 512, generator=g) * 1.5).to(BF)
    r = torch.randn(6, 512, generator=g).to(BF)
    w = (1 + 0.1 * torch.randn(512, generator=g)).to(BF)
    out["norm_x"], out["norm_res"], out["norm_w"] = x, r, w
    out["rms_norm"] = rms_norm(x, w, 1e-5)
    y, nr = fused_add_rms_norm(x.clone(), r.clone(), w, 1e-5)
    out["fused_y"], out["fused_res"] = y, nr

    for D, theta in ((128, 500000.0), (64, 1e6)):
        T, Hq, Hkv, max_pos = 9, 4, 2, 256
        q = torch.randn(T, Hq * D, generator=g).to(BF)
        k = torch.randn(T, Hkv * D, generator=g).to(BF)
        pos = torch.randint(0, max_pos, (T,), generator=g)
        inv_freq = 1.0 / (theta ** (torch.arange(0, D, 2, dtype=torch.float) / D))
        freqs = torch.einsum("i,j -> ij", torch.arange(max_pos, dtype=torch.float), inv_freq)
        cache = torch.cat((freqs.cos(), freqs.sin()), dim=-1).to(BF)  # base.py:82-93 then .to(dtype)
        rq, rk = RotaryEmbedding.forward_static(pos, q.clone(), k.clone(), D, D, cache, True)
        out[f"rope{D}"] = {"q": q, "k": k, "pos": pos, "theta": theta, "max_pos": max_pos,
                           "cache": cache, "rq": rq, "rk": rk}

    xs = (torch.randn(5, 2 * 384, generator=g) * 2).to(BF)
    out["silu_x"] = xs
    out["silu_y"] = SiluAndMul.forward_native(xs)
    torch.save(out, OUT / "ops_vllm.pt")
    print("wrote ops_vllm.pt")


def fp8_quant_from_vllm():
    """Dynamic fp8 quantisation as the reference engine's W8A8 linears apply it (QuantFP8.forward_native;
    the module needs a vLLM config context to be constructed, the method itself only reads four
    attributes, so it is called on a plain namespace)."""
    from types import SimpleNamespace

    from vllm.model_executor.layers.quantization.input_quant_fp8 import QuantFP8
    from vllm.model_executor.layers.quantization.utils.quant_utils import GroupShape

    g = torch.Generator().manual_seed(4321)
    x = (torch.randn(7, 384, generator=g) * 2.5).to(BF)
    x[2] = 0                                   # all-zero row -> minimum scale
    x[3, 5] = 3000.0                           # outlier: the rest of the row collapses towards 0
    x[4] = x[4] * 1e-4
    out = {"x": x}
    for key, shape in (("token", GroupShape.PER_TOKEN), ("tensor", GroupShape.PER_TENSOR)):
        ns = SimpleNamespace(static=False, group_shape=shape, is_group_quant=False, num_token_padding=None)
        q, scale = QuantFP8.forward_native(ns, x)
        out[f"q_{key}"] = q.view(torch.uint8).clone()
        out[f"scale_{key}"] = scale.clone()
    torch.save(out, OUT / "fp8_quant_vllm.pt")
    print("wrote fp8_quant_vllm.pt")


def fp8_linear_from_vllm():
    """The W8A8 linear as the reference engine's pure-PyTorch fallback computes it
    (ChannelWiseTorchFP8ScaledMMLinearKernel.apply_scaled_mm: torch._scaled_mm of the e4m3 codes
    with unit scales in fp32, then * x_scale * w_scale (+ bias), one cast to bf16) -- the definition
    cutlass_scaled_mm implements on the GPU.  The method reads no instance state."""
    from types import SimpleNamespace

    from vllm.model_executor.kernels.linear.scaled_mm.pytorch import ChannelWiseTorchFP8ScaledMMLinearKernel as K

    g = torch.Generator().manual_seed(777)
    x = (torch.randn(7, 256, generator=g) * 1.7).to(BF)
    w = (torch.randn(40, 256, generator=g) * 0.05).to(BF)
    b = (torch.randn(40, generator=g) * 0.3).to(BF)
    xq, xs = O.quant_fp8(x, per_token=True)           # bit-pinned to QuantFP8 by fp8_quant_vllm.pt
    wq, ws = O.quant_fp8(w, per_token=False)
    out = {"x": x, "w": w, "bias": b, "wq": wq.view(torch.uint8).clone(), "w_scale": ws.clone()}
    for key, bias in (("y", None), ("y_bias", b)):
        out[key] = K.apply_scaled_mm(SimpleNamespace(), A=xq, B=wq.t(), out_dtype=BF, As=xs, Bs=ws.reshape(1, 1),
                                     bias=bias, output_shape=[7, 40]).clone()
    torch.save(out, OUT / "fp8_linear_vllm.pt")
    print("wrote fp8_linear_vllm.pt")


def tiny_llama_from_hf():
    from transformers import LlamaConfig, LlamaForCausalLM

    out = {}
    for name, cfg in (("tiny", O.TINY), ("tiny128", O.TINY128)):
        weights = O.init_weights(cfg, seed=7)
        hf_cfg = LlamaConfig(
            vocab_size=cfg.vocab_size, hidden_size=cfg.hidden_size, intermediate_size=cfg.intermediate_size,
            num_hidden_layers=cfg.num_hidden_layers, num_attention_heads=cfg.num_attention_heads,
            num_key_value_heads=cfg.num_key_value_heads, head_dim=cfg.head_dim, rms_norm_eps=cfg.rms_norm_eps,
            rope_theta=cfg.rope_theta, max_position_embeddings=cfg.max_position_embeddings,
            tie_word_embeddings=False, attention_bias=False, mlp_bias=False, hidden_act="silu",
            attn_implementation="eager")
        g = torch.Generator().manual_seed(42)
        prompt = torch.randint(3, cfg.vocab_size, (1, 23), generator=g)
        res = {"prompt": prompt[0].tolist()}
        for dt, key in ((torch.float32, "fp32"), (BF, "bf16")):
            model = LlamaForCausalLM(hf_cfg).to(dt).eval()
            missing = model.load_state_dict({k: v.to(dt) for k, v in weights.items()}, strict=True)
            with torch.no_grad():
                logits = model(prompt).logits[0].float()
                gen = model.generate(prompt, max_new_tokens=12, do_sample=False, pad_token_id=0)
            res[f"logits_{key}"] = logits[-4:].clone()      # last 4 positions are enough to pin
            res[f"greedy_{key}"] = gen[0, prompt.shape[1]:].tolist()
        out[name] = res
    torch.save(out, OUT / "tiny_llama_hf.pt")
    print("wrote tiny_llama_hf.pt")


def tiny_qwen2_from_hf():
    """Qwen2 family (BASELINE.json configs[0]): q/k/v bias + tied embeddings, GQA group of 7."""
    from transformers import Qwen2Config, Qwen2ForCausalLM

    cfg = O.TINY_QWEN2
    weights = O.init_weights(cfg, seed=7)
    hf_cfg = Qwen2Config(
        vocab_size=cfg.vocab_size, hidden_size=cfg.hidden_size, intermediate_size=cfg.intermediate_size,
        num_hidden_layers=cfg.num_hidden_layers, num_attention_heads=cfg.num_attention_heads,
        num_key_value_heads=cfg.num_key_value_heads, rms_norm_eps=cfg.rms_norm_eps, rope_theta=cfg.rope_theta,
        max_position_embeddings=cfg.max_position_embeddings, tie_word_embeddings=True, use_sliding_window=False,
        attn_implementation="eager")
    g = torch.Generator().manual_seed(43)
    prompt = torch.randint(3, cfg.vocab_size, (1, 19), generator=g)
    res = {"prompt": prompt[0].tolist()}
    for dt, key in ((torch.float32, "fp32"), (BF, "bf16")):
        model = Qwen2ForCausalLM(hf_cfg).to(dt).eval()
        sd = {k: v.to(dt) for k, v in weights.items()}
        sd["lm_head.weight"] = sd["model.embed_tokens.weight"]
        model.load_state_dict(sd, strict=True)
        with torch.no_grad():
            res[f"logits_{key}"] = model(prompt).logits[0].float()[-4:].clone()
            gen = model.generate(prompt, max_new_tokens=10, do_sample=False, pad_token_id=0)
        res[f"greedy_{key}"] = gen[0, prompt.shape[1]:].tolist()
    torch.save(res, OUT / "tiny_qwen2_hf.pt")
    print("wrote tiny_qwen2_hf.pt")


def tiny_mixtral_from_hf():
    """Mixtral family (BASELINE.json configs[4]): 8 experts, top-2 routing with renormalisation.
    transformers 5.x keeps the experts fused in memory (mlp.experts.gate_up_proj [E, 2I, H] = gate
    rows then up rows, mlp.experts.down_proj [E, H, I], router mlp.gate.weight); checkpoints -- and
    the oracle -- use the classic block_sparse_moe.experts.{e}.w1/w3/w2 names."""
    from transformers import MixtralConfig, MixtralForCausalLM

    cfg = O.TINY_MIXTRAL
    weights = O.init_weights(cfg, seed=7)
    hf_cfg = MixtralConfig(
        vocab_size=cfg.vocab_size, hidden_size=cfg.hidden_size, intermediate_size=cfg.intermediate_size,
        num_hidden_layers=cfg.num_hidden_layers, num_attention_heads=cfg.num_attention_heads,
        num_key_value_heads=cfg.num_key_value_heads, head_dim=cfg.head_dim, rms_norm_eps=cfg.rms_norm_eps,
        rope_theta=cfg.rope_theta, max_position_embeddings=cfg.max_position_embeddings,
        num_local_experts=cfg.num_local_experts, num_experts_per_tok=cfg.num_experts_per_tok,
        sliding_window=None, tie_word_embeddings=False, router_jitter_noise=0.0, attn_implementation="eager")
    sd = {}
    E = cfg.num_local_experts
    for k, v in weights.items():
        if "block_sparse_moe" not in k:
            sd[k] = v
    for l in range(cfg.num_hidden_layers):
        p = f"model.layers.{l}."
        sd[p + "mlp.gate.weight"] = weights[p + "block_sparse_moe.gate.weight"]
        sd[p + "mlp.experts.gate_up_proj"] = torch.stack([
            torch.cat([weights[p + f"block_sparse_moe.experts.{e}.w1.weight"],
                       weights[p + f"block_sparse_moe.experts.{e}.w3.weight"]]) for e in range(E)])
        sd[p + "mlp.experts.down_proj"] = torch.stack([weights[p + f"block_sparse_moe.experts.{e}.w2.weight"]
                                                       for e in range(E)])
    g = torch.Generator().manual_seed(44)
    prompt = torch.randint(3, cfg.vocab_size, (1, 29), generator=g)
    res = {"prompt": prompt[0].tolist()}
    for dt, key in ((torch.float32, "fp32"), (BF, "bf16")):
        model = MixtralForCausalLM(hf_cfg).to(dt).eval()
        model.load_state_dict({k: v.to(dt) for k, v in sd.items()}, strict=True)
        with torch.no_grad():
            res[f"logits_{key}"] = model(prompt).logits[0].float()[-4:].clone()
            gen = model.generate(prompt, max_new_tokens=10, do_sample=False, pad_token_id=0)
        res[f"greedy_{key}"] = gen[0, prompt.shape[1]:].tolist()
    torch.save(res, OUT / "tiny_mixtral_hf.pt")
    print("wrote tiny_mixtral_hf.pt")


if __name__ == "__main__":
    ops_from_vllm()
    fp8_quant_from_vllm()
    fp8_linear_from_vllm()
    tiny_llama_from_hf()
    tiny_qwen2_from_hf()
    tiny_mixtral_from_hf()
