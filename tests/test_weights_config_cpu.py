"""Weight sharding, checkpoint round trip and config parsing (CPU): the tensor-parallel shards must
reassemble to exactly the un-sharded computation (QKV/gate_up column-parallel,
O/down row-parallel + sum, vocab-parallel lm_head), and an HF-format directory written by
save_checkpoint must load back bit-identically."""
import json

import pytest
import torch

from gpustack_b200.config import EngineConfig, ModelConfig, llama3_8b, tiny_llama
from gpustack_b200.weights import (check_tp, load_safetensors_dir, random_state_dict, save_checkpoint,
                                   shard_state_dict)


def _cfg(bias=False, tie=False):
    return ModelConfig(vocab_size=300, hidden_size=128, intermediate_size=256, num_hidden_layers=2,
                       num_attention_heads=8, num_key_value_heads=4, head_dim=16, attention_bias=bias,
                       tie_word_embeddings=tie, architecture="Qwen2ForCausalLM" if bias else "LlamaForCausalLM")


@pytest.mark.parametrize("tp", [1, 2, 4])
def test_shards_reassemble_to_the_full_projection(tp):
    cfg = _cfg(bias=True)
    sd = random_state_dict(cfg, seed=1, device="cpu")
    x = torch.randn(5, cfg.hidden_size)
    shards = [shard_state_dict(cfg, lambda k: sd[k], r, tp, "cpu") for r in range(tp)]
    D, hq, hkv, I = cfg.head_dim, cfg.num_attention_heads // tp, cfg.num_key_value_heads // tp, cfg.intermediate_size // tp
    p = "model.layers.1."
    full_q = x @ sd[p + "self_attn.q_proj.weight"].float().t()
    full_k = x @ sd[p + "self_attn.k_proj.weight"].float().t()
    full_v = x @ sd[p + "self_attn.v_proj.weight"].float().t()
    qs, ks, vs, gates, ups = [], [], [], [], []
    for sh in shards:
        lw = sh.layers[1]
        y = x @ lw.qkv.float().t()
        qs.append(y[:, : hq * D]); ks.append(y[:, hq * D:(hq + hkv) * D]); vs.append(y[:, (hq + hkv) * D:])
        gu = x @ lw.gate_up.float().t()           # interleaved rows: 2i = gate_i, 2i+1 = up_i
        gates.append(gu[:, 0::2]); ups.append(gu[:, 1::2])
        assert lw.qkv_bias.shape[0] == (hq + 2 * hkv) * D
    assert torch.equal(torch.cat(qs, 1), full_q) and torch.equal(torch.cat(ks, 1), full_k)
    assert torch.equal(torch.cat(vs, 1), full_v)
    assert torch.equal(torch.cat(gates, 1), x @ sd[p + "mlp.gate_proj.weight"].float().t())
    assert torch.equal(torch.cat(ups, 1), x @ sd[p + "mlp.up_proj.weight"].float().t())
    bias_q = torch.cat([sh.layers[1].qkv_bias[: hq * D] for sh in shards])
    assert torch.equal(bias_q, sd[p + "self_attn.q_proj.bias"])
    # row-parallel: partial products sum to the full product
    a = torch.randn(5, cfg.q_size)
    parts = sum(a[:, r * hq * D:(r + 1) * hq * D] @ shards[r].layers[1].o.float().t() for r in range(tp))
    torch.testing.assert_close(parts, a @ sd[p + "self_attn.o_proj.weight"].float().t(), rtol=1e-5, atol=1e-5)
    h = torch.randn(5, cfg.intermediate_size)
    parts = sum(h[:, r * I:(r + 1) * I] @ shards[r].layers[1].down.float().t() for r in range(tp))
    torch.testing.assert_close(parts, h @ sd[p + "mlp.down_proj.weight"].float().t(), rtol=1e-5, atol=1e-5)
    # vocab-parallel lm_head covers the vocabulary exactly once, in order
    assert [s.vocab_start for s in shards] == [r * ((300 + tp - 1) // tp) for r in range(tp)]
    assert shards[-1].vocab_end == 300
    assert torch.equal(torch.cat([s.lm_head for s in shards]), sd["lm_head.weight"])


def test_tied_embeddings_share_storage_on_one_gpu():
    cfg = _cfg(tie=True)
    sd = random_state_dict(cfg, seed=2, device="cpu")
    assert "lm_head.weight" not in sd
    w = shard_state_dict(cfg, lambda k: sd[k], 0, 1, "cpu")
    assert w.lm_head.data_ptr() == w.embed.data_ptr()
    w1 = shard_state_dict(cfg, lambda k: sd[k], 1, 2, "cpu")       # TP: a row-range view, still no copy
    assert w1.lm_head.untyped_storage().data_ptr() == w1.embed.untyped_storage().data_ptr()
    assert torch.equal(w1.lm_head, sd["model.embed_tokens.weight"][w1.vocab_start:w1.vocab_end])


def test_checkpoint_round_trip(tmp_path):
    cfg = _cfg()
    sd = random_state_dict(cfg, seed=3, device="cpu")
    save_checkpoint(cfg, sd, str(tmp_path / "ckpt"), shard_bytes=200_000)     # forces several shards + index
    assert (tmp_path / "ckpt" / "model.safetensors.index.json").exists()
    cfg2 = ModelConfig.from_pretrained(str(tmp_path / "ckpt"))
    assert cfg2.hidden_size == cfg.hidden_size and cfg2.head_dim == cfg.head_dim and cfg2.num_key_value_heads == 4
    a = load_safetensors_dir(cfg2, str(tmp_path / "ckpt"), 0, 1, "cpu")
    b = shard_state_dict(cfg, lambda k: sd[k], 0, 1, "cpu")
    for la, lb in zip(a.layers, b.layers):
        for f in ("qkv", "o", "gate_up", "down", "input_norm", "post_norm"):
            assert torch.equal(getattr(la, f), getattr(lb, f))
    assert torch.equal(a.embed, b.embed) and torch.equal(a.lm_head, b.lm_head)


def test_tp_divisibility_rule_matches_the_reference_scheduler():
    with pytest.raises(ValueError):
        check_tp(llama3_8b(), 16)           # 8 kv heads: TP <= 8
    with pytest.raises(ValueError):
        check_tp(llama3_8b(), 3)
    for t in (1, 2, 4, 8):
        check_tp(llama3_8b(), t)


def test_config_parsing_and_byte_accounting():
    c = llama3_8b()
    assert c.weight_bytes(1) == 15_009_849_344 and c.kv_bytes_per_token(1) == 131_072
    assert c.kv_bytes_per_token(8) == 16_384
    hf = {"architectures": ["LlamaForCausalLM"], "vocab_size": 100, "hidden_size": 64, "intermediate_size": 128,
          "num_hidden_layers": 1, "num_attention_heads": 4, "rope_theta": 1e4, "eos_token_id": [7, 9]}
    m = ModelConfig.from_dict(hf)
    assert m.head_dim == 16 and m.num_key_value_heads == 4 and m.eos_token_id == 7
    with pytest.raises(ValueError):
        ModelConfig.from_dict({**hf, "architectures": ["DeepseekV3ForCausalLM"]})
    with pytest.raises(ValueError):   # yarn / dynamic scaling change attention scaling too: refused, never ignored
        ModelConfig.from_dict({**hf, "rope_scaling": {"rope_type": "yarn", "factor": 8.0}})
    with pytest.raises(ValueError):   # the transformers>=5 spelling must not slip through either
        ModelConfig.from_dict({**hf, "rope_parameters": {"rope_type": "yarn", "factor": 8.0, "rope_theta": 1e4}})
    e = EngineConfig(max_num_seqs=100)
    b = e.graph_buckets()
    assert b[0] == 1 and b[-1] == 100 and b == sorted(set(b))
    assert tiny_llama(128).head_dim == 128


# ---------------------------------------------------------------------------------- fp8 (W8A8)
def _deq(w, s):
    return w.to(torch.float32) * s


def test_online_fp8_quantisation_matches_the_oracle():
    """`--quantization fp8` on a bf16 checkpoint: one per-tensor scale per FUSED projection, same
    bytes as oracle.quant_fp8 of the fused weight (vLLM fp8.py:537-549)."""
    from oracle import llama_oracle as O

    cfg = _cfg(bias=True)
    sd = random_state_dict(cfg, seed=4, device="cpu")
    w = shard_state_dict(cfg, lambda k: sd[k], 0, 1, "cpu", quantization="fp8", has=lambda k: k in sd)
    p = "model.layers.0."
    lw = w.layers[0]
    assert lw.qkv.dtype == torch.float8_e4m3fn and lw.qkv_scale.shape == (1,) and lw.qkv_bias.dtype == torch.bfloat16
    fused = torch.cat([sd[p + f"self_attn.{n}_proj.weight"] for n in "qkv"])
    rq, rs = O.quant_fp8(fused, per_token=False)
    assert torch.equal(lw.qkv.view(torch.uint8), rq.view(torch.uint8)) and torch.equal(lw.qkv_scale, rs)
    g, u = sd[p + "mlp.gate_proj.weight"], sd[p + "mlp.up_proj.weight"]
    rq, rs = O.quant_fp8(torch.cat([g, u]), per_token=False)
    I = cfg.intermediate_size
    assert torch.equal(lw.gate_up.view(torch.uint8)[0::2], rq.view(torch.uint8)[:I])       # interleaved rows
    assert torch.equal(lw.gate_up.view(torch.uint8)[1::2], rq.view(torch.uint8)[I:]) and torch.equal(lw.gate_up_scale, rs)
    for name, t, sc in (("self_attn.o_proj", lw.o, lw.o_scale), ("mlp.down_proj", lw.down, lw.down_scale)):
        rq, rs = O.quant_fp8(sd[p + name + ".weight"], per_token=False)
        assert torch.equal(t.view(torch.uint8), rq.view(torch.uint8)) and torch.equal(sc, rs)
    assert w.embed.dtype == torch.bfloat16 and w.lm_head.dtype == torch.bfloat16 and w.final_norm.dtype == torch.bfloat16
    # quantisation error of the dequantised weight stays within one e4m3 step (2^-4 relative)
    err = (_deq(lw.o, lw.o_scale) - sd[p + "self_attn.o_proj.weight"].float()).abs().max()
    assert float(err) <= float(lw.o_scale) * 448 * 2 ** -4


def test_fp8_checkpoint_round_trip_requantises_fused_shards_to_the_max_scale(tmp_path):
    """A per-tensor fp8 checkpoint (weight + weight_scale per projection, as `--quantization fp8`
    exports and llm-compressor's FP8_DYNAMIC-per-tensor writes): single projections are taken
    as they are, fused ones are re-quantised to the largest shard scale (vLLM fp8.py:406-428)."""
    from oracle import llama_oracle as O

    cfg = _cfg()
    sd = random_state_dict(cfg, seed=5, device="cpu")
    q8 = {}
    for k, v in sd.items():
        if k.endswith("_proj.weight"):
            mul = 3.0 if "k_proj" in k else 1.0                    # make the shard scales differ
            wq, ws = O.quant_fp8((v.float() * mul).to(torch.bfloat16), per_token=False)
            q8[k] = wq
            q8[k[: -len("weight")] + "weight_scale"] = ws.reshape(())
        else:
            q8[k] = v
    cfg8 = ModelConfig.from_dict({**cfg.to_hf_dict(), "quantization_config": {"quant_method": "fp8",
                                                                             "activation_scheme": "dynamic"}})
    assert cfg8.quantization == "fp8" and cfg8.to_hf_dict()["quantization_config"]["quant_method"] == "fp8"
    save_checkpoint(cfg8, q8, str(tmp_path / "ckpt8"))
    cfg_l = ModelConfig.from_pretrained(str(tmp_path / "ckpt8"))
    assert cfg_l.quantization == "fp8"
    w = load_safetensors_dir(cfg_l, str(tmp_path / "ckpt8"), 0, 1, "cpu", quantization=cfg_l.quantization)
    p = "model.layers.1."
    lw = w.layers[1]
    assert torch.equal(lw.o.view(torch.uint8), q8[p + "self_attn.o_proj.weight"].view(torch.uint8))
    assert torch.equal(lw.o_scale, q8[p + "self_attn.o_proj.weight_scale"].reshape(1))
    scs = [q8[p + f"self_attn.{n}_proj.weight_scale"] for n in "qkv"]
    assert torch.equal(lw.qkv_scale, torch.stack(scs).max().reshape(1)) and float(scs[1]) > float(scs[0])
    deq_ref = torch.cat([_deq(q8[p + f"self_attn.{n}_proj.weight"], q8[p + f"self_attn.{n}_proj.weight_scale"])
                         for n in "qkv"])
    # the shard that owned the max scale is reproduced exactly, the others within one e4m3 step of the new scale
    nq = cfg.q_size
    kk = slice(nq, nq + cfg.kv_size)
    assert torch.equal(_deq(lw.qkv, lw.qkv_scale)[kk], deq_ref[kk])
    assert float((_deq(lw.qkv, lw.qkv_scale) - deq_ref).abs().max()) <= float(lw.qkv_scale) * 448 * 2 ** -4
    # tensor-parallel shards quantise their own slice
    w2 = load_safetensors_dir(cfg_l, str(tmp_path / "ckpt8"), 1, 2, "cpu", quantization="fp8")
    assert w2.layers[0].down.shape == (cfg.hidden_size, cfg.intermediate_size // 2)
    assert w2.layers[0].down.dtype == torch.float8_e4m3fn


def test_unsupported_quantisation_configs_are_refused():
    base = _cfg().to_hf_dict()
    for qc in ({"quant_method": "awq"}, {"quant_method": "fp8", "activation_scheme": "static"},
               {"quant_method": "fp8", "weight_block_size": [128, 128]}):
        with pytest.raises(ValueError):
            ModelConfig.from_dict({**base, "quantization_config": qc})
    cfg = _cfg()
    sd = random_state_dict(cfg, seed=6, device="cpu")
    with pytest.raises(ValueError):
        shard_state_dict(cfg, lambda k: sd[k], 0, 1, "cpu", quantization="int4")
    # static activation scales in the checkpoint
    from oracle import llama_oracle as O
    q8 = dict(sd)
    for k in list(sd):
        if k.endswith("_proj.weight"):
            q8[k], s = O.quant_fp8(sd[k], per_token=False)
            q8[k[:-6] + "weight_scale"] = s
            q8[k[:-6] + "input_scale"] = torch.ones(1)
    with pytest.raises(ValueError, match="static activation"):
        shard_state_dict(cfg, lambda k: q8[k], 0, 1, "cpu", quantization="fp8", has=lambda k: k in q8)


# ------------------------------------------------------------------------------------ Mixtral
def test_mixtral_expert_weights_are_stacked_interleaved_and_tp_sharded():
    from gpustack_b200.config import tiny_mixtral

    cfg = tiny_mixtral()
    sd = random_state_dict(cfg, seed=8, device="cpu")
    E, I, H = cfg.num_local_experts, cfg.intermediate_size, cfg.hidden_size
    full = shard_state_dict(cfg, lambda k: sd[k], 0, 1, "cpu", has=lambda k: k in sd)
    lw = full.layers[1]
    p = "model.layers.1.block_sparse_moe."
    assert lw.gate_up is None and lw.down is None
    assert lw.router.shape == (128, H) and torch.equal(lw.router[:E], sd[p + "gate.weight"]) and not lw.router[E:].any()
    assert lw.w13.shape == (E * 2 * I, H) and lw.w2.shape == (E * H, I)
    for e in (0, 5):
        blk = lw.w13[e * 2 * I:(e + 1) * 2 * I]
        assert torch.equal(blk[0::2], sd[p + f"experts.{e}.w1.weight"]) and torch.equal(blk[1::2], sd[p + f"experts.{e}.w3.weight"])
        assert torch.equal(lw.w2[e * H:(e + 1) * H], sd[p + f"experts.{e}.w2.weight"])
    # tensor parallel over I: the two ranks' partial expert outputs sum to the full one
    tp = 2
    shards = [shard_state_dict(cfg, lambda k: sd[k], r, tp, "cpu", has=lambda k: k in sd) for r in range(tp)]
    x = torch.randn(3, H)
    e = 3
    ref_g = x @ sd[p + f"experts.{e}.w1.weight"].float().t()
    ref_u = x @ sd[p + f"experts.{e}.w3.weight"].float().t()
    ref = (torch.nn.functional.silu(ref_g) * ref_u) @ sd[p + f"experts.{e}.w2.weight"].float().t()
    acc = 0
    for sh in shards:
        l2 = sh.layers[1]
        Il = I // tp
        blk = l2.w13[e * 2 * Il:(e + 1) * 2 * Il].float()
        gu = x @ blk.t()
        acc = acc + (torch.nn.functional.silu(gu[:, 0::2]) * gu[:, 1::2]) @ l2.w2[e * H:(e + 1) * H].float().t()
    torch.testing.assert_close(acc, ref, rtol=1e-4, atol=1e-4)


def test_mixtral_config_and_transformers5_fused_checkpoint_layout(tmp_path):
    from gpustack_b200.config import mixtral_8x7b, tiny_mixtral

    cfg = tiny_mixtral()
    d = cfg.to_hf_dict()
    assert d["model_type"] == "mixtral" and d["num_local_experts"] == 8 and d["num_experts_per_tok"] == 2
    back = ModelConfig.from_dict(d)
    assert back.num_local_experts == 8 and back.num_experts_per_tok == 2 and back.architecture == "MixtralForCausalLM"
    big = mixtral_8x7b()
    # 46,702,792,704 parameters
    n = big.vocab_size * big.hidden_size * 2 + big.num_hidden_layers * (
        (big.q_size + 2 * big.kv_size) * big.hidden_size + big.q_size * big.hidden_size + 8 * 3 * big.intermediate_size * big.hidden_size
        + 8 * big.hidden_size + 2 * big.hidden_size) + big.hidden_size
    assert n == 46_702_792_704
    # sliding window: served with the context capped at the window
    assert ModelConfig.from_dict({**d, "sliding_window": 256}).sliding_window == 256
    assert ModelConfig.from_dict({**d, "sliding_window": 256, "use_sliding_window": False}).sliding_window is None
    with pytest.raises(ValueError):
        ModelConfig.from_dict({**d, "quantization_config": {"quant_method": "fp8"}})
    # classic checkpoint names and the transformers-5 fused export load to the same tensors
    sd = random_state_dict(cfg, seed=9, device="cpu")
    fused = {k: v for k, v in sd.items() if "block_sparse_moe" not in k}
    E = cfg.num_local_experts
    for l in range(cfg.num_hidden_layers):
        p = f"model.layers.{l}."
        fused[p + "mlp.gate.weight"] = sd[p + "block_sparse_moe.gate.weight"]
        fused[p + "mlp.experts.gate_up_proj"] = torch.stack([torch.cat([sd[p + f"block_sparse_moe.experts.{e}.w1.weight"],
                                                                        sd[p + f"block_sparse_moe.experts.{e}.w3.weight"]])
                                                             for e in range(E)])
        fused[p + "mlp.experts.down_proj"] = torch.stack([sd[p + f"block_sparse_moe.experts.{e}.w2.weight"] for e in range(E)])
    save_checkpoint(cfg, sd, str(tmp_path / "classic"))
    save_checkpoint(cfg, fused, str(tmp_path / "fused"))
    a = load_safetensors_dir(ModelConfig.from_pretrained(str(tmp_path / "classic")), str(tmp_path / "classic"), 0, 1, "cpu")
    b = load_safetensors_dir(ModelConfig.from_pretrained(str(tmp_path / "fused")), str(tmp_path / "fused"), 0, 1, "cpu")
    for la, lb in zip(a.layers, b.layers):
        assert torch.equal(la.router, lb.router) and torch.equal(la.w13, lb.w13) and torch.equal(la.w2, lb.w2)


# --------------------------------------------------------------------- VRAM model
def _resident_bytes(w):
    """bytes the shard occupies once every tensor is its own device buffer (on the CPU the row
    slices are views of the full checkpoint tensors); a tied lm_head is a view of the embedding."""
    tensors = [w.embed, w.final_norm]
    if w.lm_head.untyped_storage().data_ptr() != w.embed.untyped_storage().data_ptr():
        tensors.append(w.lm_head)
    for lw in w.layers:
        tensors += [t for t in vars(lw).values() if isinstance(t, torch.Tensor)]
    return sum(t.numel() * t.element_size() for t in tensors)


@pytest.mark.parametrize("kind", ["llama", "qwen2-bias-tied", "mixtral", "fp8"])
@pytest.mark.parametrize("tp", [1, 2])
def test_vram_model_weight_bytes_equal_the_loaded_shards(kind, tp):
    from gpustack_b200.config import tiny_mixtral
    from gpustack_b200.resource_fit import weight_bytes_per_gpu

    cfg = {"llama": _cfg(), "qwen2-bias-tied": _cfg(bias=True, tie=True), "mixtral": tiny_mixtral(), "fp8": _cfg()}[kind]
    quant = "fp8" if kind == "fp8" else None
    sd = random_state_dict(cfg, seed=11, device="cpu")
    for rank in range(tp):
        w = shard_state_dict(cfg, lambda k: sd[k], rank, tp, "cpu", quantization=quant, has=lambda k: k in sd)
        assert weight_bytes_per_gpu(cfg, tp, quant) == _resident_bytes(w)


def test_vram_estimate_and_cli(capsys):
    from gpustack_b200.cli import main
    from gpustack_b200.resource_fit import GiB, estimate

    m = llama3_8b()
    e1 = estimate(m, EngineConfig(max_model_len=8192))
    assert e1.weights == 16_060_522_496 and e1.kv_tokens == 8192 and e1.kv_bytes_per_token == 131_072
    assert e1.total_min == e1.per_gpu_min and 20 * GiB < e1.total_min < 22 * GiB
    assert e1.env() == {"GPUSTACK_MODEL_VRAM_CLAIM": str(e1.total_min)}
    assert e1.reference_heuristic == int(16_060_522_496 * 1.2) + 2 * GiB          # GPUStack's own formula
    e8 = estimate(m, EngineConfig(max_model_len=8192, tensor_parallel_size=8), kv_tokens=1_000_000)
    assert e8.kv_bytes_per_token == 16_384 and e8.total_min == 8 * e8.per_gpu_min
    assert e8.kv_cache_requested >= 1_000_000 * 16_384 and e8.weights < e1.weights / 4
    f8 = estimate(m, EngineConfig(max_model_len=8192, quantization="fp8"))
    assert f8.weights < 0.6 * e1.weights
    assert main(["estimate", "random:llama3-8b", "--max-model-len", "8192", "--kv-tokens", "500000"]) == 0
    out = json.loads(capsys.readouterr().out)
    assert out["kv_tokens"] == 500000 and int(out["env"]["GPUSTACK_MODEL_VRAM_CLAIM"]) == out["total_min"]


def test_lazy_random_source_is_deterministic_and_complete():
    from gpustack_b200.config import tiny_mixtral
    from gpustack_b200.weights import lazy_random_source

    for cfg in (tiny_mixtral(), _cfg(bias=True, tie=True)):
        get, has = lazy_random_source(cfg, seed=3, device="cpu")
        w = shard_state_dict(cfg, get, 0, 1, "cpu", has=has)
        get2, has2 = lazy_random_source(cfg, seed=3, device="cpu")
        w2 = shard_state_dict(cfg, get2, 0, 1, "cpu", has=has2)
        assert torch.equal(w.layers[1].qkv, w2.layers[1].qkv) and torch.equal(w.embed, w2.embed)
        assert not has("model.layers.0.self_attn.q_proj.weight_scale")
        get3, _ = lazy_random_source(cfg, seed=4, device="cpu")
        assert not torch.equal(get("model.norm.weight"), get3("model.norm.weight"))
        assert abs(float(get("model.norm.weight").float().mean()) - 1.0) < 0.1
        if cfg.num_local_experts:
            assert w.layers[0].w13.shape[0] == cfg.num_local_experts * 2 * cfg.intermediate_size
            assert float(w.layers[0].router[:8].float().std()) > 5 * float(w.layers[0].w2.float().std())


def test_rope_scaling_matches_transformers_and_eos_ids_are_all_kept(tmp_path):
    """Llama-3.1-style and linear RoPE scaling (legacy `rope_scaling` and the transformers>=5
    `rope_parameters` spelling) give the inverse frequencies transformers computes; every EOS id of
    config.json + generation_config.json ends a turn."""
    import json

    import torch
    from transformers import LlamaConfig
    from transformers.modeling_rope_utils import ROPE_INIT_FUNCTIONS

    hf = tiny_llama(128).to_hf_dict()
    l3 = {"rope_type": "llama3", "factor": 8.0, "low_freq_factor": 1.0, "high_freq_factor": 4.0,
          "original_max_position_embeddings": 8192}
    for spelling in ("rope_scaling", "rope_parameters"):
        d = {**hf, "max_position_embeddings": 131072, "rope_theta": 500000.0}
        d[spelling] = {**l3, "rope_theta": 500000.0} if spelling == "rope_parameters" else l3
        m = ModelConfig.from_dict(d)
        assert m.rope_scaling["rope_type"] == "llama3"
        ours = m.rope_inv_freq()
        ref_cfg = LlamaConfig(hidden_size=m.hidden_size, num_attention_heads=m.num_attention_heads, head_dim=m.head_dim,
                              max_position_embeddings=131072, rope_theta=500000.0,
                              rope_scaling={**l3, "rope_theta": 500000.0})
        ref, _ = ROPE_INIT_FUNCTIONS["llama3"](ref_cfg, "cpu")
        torch.testing.assert_close(ours, ref.float(), rtol=1e-6, atol=0)
        assert not torch.equal(ours, ModelConfig.from_dict(hf).rope_inv_freq() * 0 + ours[0])  # really scaled
    lin = ModelConfig.from_dict({**hf, "rope_scaling": {"type": "linear", "factor": 4.0}})
    base = ModelConfig.from_dict(hf)
    torch.testing.assert_close(lin.rope_inv_freq(), base.rope_inv_freq() / 4.0)
    assert ModelConfig.from_dict({**hf, "rope_scaling": {"rope_type": "default"}}).rope_scaling is None
    # round trip through to_hf_dict keeps the scaling
    assert ModelConfig.from_dict(lin.to_hf_dict()).rope_scaling == lin.rope_scaling

    # EOS ids: list in config.json, more in generation_config.json
    d = tmp_path / "m"
    d.mkdir()
    (d / "config.json").write_text(json.dumps({**hf, "eos_token_id": [7, 9]}))
    (d / "generation_config.json").write_text(json.dumps({"eos_token_id": [7, 11]}))
    m = ModelConfig.from_pretrained(str(d))
    assert m.eos_token_id == 7 and m.eos_token_ids == (7, 9, 11)
