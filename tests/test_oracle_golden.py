"""Pin the CPU oracle against fixtures produced by the reference engine's own pure-PyTorch
operator definitions (vLLM) and by HF transformers (tests/golden/make_golden.py).  CPU only."""
from pathlib import Path

import pytest
import torch

from oracle import llama_oracle as O

GOLD = Path(__file__).resolve().parent / "golden"
BF = torch.bfloat16


@pytest.fixture(scope="module")
def ops_gold():
    return torch.load(GOLD / "ops_vllm.pt", weights_only=False)


@pytest.fixture(scope="module")
def hf_gold():
    return torch.load(GOLD / "tiny_llama_hf.pt", weights_only=False)


def test_rms_norm_bit_exact(ops_gold):
    g = ops_gold
    assert torch.equal(O.rms_norm(g["norm_x"], g["norm_w"], 1e-5), g["rms_norm"])


def test_fused_add_rms_norm_bit_exact(ops_gold):
    g = ops_gold
    y, r = O.fused_add_rms_norm(g["norm_x"], g["norm_res"], g["norm_w"], 1e-5)
    assert torch.equal(r, g["fused_res"])
    assert torch.equal(y, g["fused_y"])


def test_silu_and_mul_bit_exact(ops_gold):
    assert torch.equal(O.silu_and_mul(ops_gold["silu_x"]), ops_gold["silu_y"])


@pytest.mark.parametrize("D", [128, 64])
def test_rope_matches_vllm_native(ops_gold, D):
    r = ops_gold[f"rope{D}"]
    cache = O.cos_sin_cache(D, r["max_pos"], r["theta"])
    assert torch.equal(cache, r["cache"])  # same cos|sin table, bit for bit
    T = r["q"].shape[0]
    q = O.rope_neox(r["q"].view(T, -1, D), r["pos"], cache).reshape(T, -1)
    k = O.rope_neox(r["k"].view(T, -1, D), r["pos"], cache).reshape(T, -1)
    # oracle rounds once (fp32 rotation); vLLM native rounds each bf16 op, so where x1*cos - x2*sin
    # cancels the two differ by up to one bf16 ulp of the OPERANDS (|x| < 4 -> ulp 2^-6 = 0.0156)
    torch.testing.assert_close(q.float(), r["rq"].float(), atol=1.6e-2, rtol=1.6e-2)
    torch.testing.assert_close(k.float(), r["rk"].float(), atol=1.6e-2, rtol=1.6e-2)


@pytest.mark.parametrize("name,cfg", [("tiny", O.TINY), ("tiny128", O.TINY128)])
def test_whole_model_against_hf(hf_gold, name, cfg):
    g = hf_gold[name]
    model = O.OracleLlama(cfg, O.init_weights(cfg, seed=7))
    logits = model.forward(0, g["prompt"])[-4:]
    ref32, ref16 = g["logits_fp32"], g["logits_bf16"]
    scale = float(ref32.abs().max())
    # stated tolerance: max-abs-diff <= 2e-2 * max|logit| for a bf16 pipeline
    assert float((logits - ref32).abs().max()) <= 2e-2 * scale
    assert float((logits - ref16).abs().max()) <= 3e-2 * scale
    # greedy tokens: exact wherever HF's own fp32 and bf16 runs agree (i.e. no near-tie)
    model.reset()
    mine = model.generate(0, g["prompt"], 12)
    agree = [a == b for a, b in zip(g["greedy_fp32"], g["greedy_bf16"])]
    n_stable = agree.index(False) if False in agree else len(agree)
    assert mine[:n_stable] == g["greedy_fp32"][:n_stable]


def test_paged_gather_and_attention_shapes():
    g = torch.Generator().manual_seed(0)
    cache = torch.randn(6, 2, 16, 64, generator=g).to(BF)
    kv = O.gather_paged(cache, [4, 1, 3], 40)
    assert kv.shape == (40, 2, 64)
    assert torch.equal(kv[17], cache[1, :, 1])
    q = torch.randn(3, 8, 64, generator=g).to(BF)
    o = O.attention_ref(q, kv, kv, 0.125, torch.tensor([37, 38, 39]))
    assert o.shape == (3, 8, 64)
    # causality: the first query must ignore keys 38, 39
    o2 = O.attention_ref(q[:1], kv[:38], kv[:38], 0.125, torch.tensor([37]))
    assert torch.equal(o[:1], o2)


def test_qwen2_family_against_hf():
    """q/k/v bias, tied embeddings, GQA group of 7 (the Qwen2-0.5B family of BASELINE.json configs[0])."""
    g = torch.load(GOLD / "tiny_qwen2_hf.pt", weights_only=False)
    model = O.OracleLlama(O.TINY_QWEN2, O.init_weights(O.TINY_QWEN2, seed=7))
    logits = model.forward(0, g["prompt"])[-4:]
    scale = float(g["logits_fp32"].abs().max())
    assert float((logits - g["logits_fp32"]).abs().max()) <= 2e-2 * scale
    model.reset()
    mine = model.generate(0, g["prompt"], 10)
    agree = [a == b for a, b in zip(g["greedy_fp32"], g["greedy_bf16"])]
    n_stable = agree.index(False) if False in agree else len(agree)
    assert mine[:n_stable] == g["greedy_fp32"][:n_stable]


def test_fp8_quant_against_vllm_native():
    """oracle.quant_fp8 == vLLM's QuantFP8.forward_native bit for bit (dynamic per-token and
    per-tensor), including the all-zero row (minimum scale) and the outlier row."""
    g = torch.load(GOLD / "fp8_quant_vllm.pt", weights_only=False)
    for key, per_token in (("token", True), ("tensor", False)):
        q, s = O.quant_fp8(g["x"], per_token=per_token)
        assert torch.equal(s.reshape(-1), g[f"scale_{key}"].reshape(-1))
        assert torch.equal(q.view(torch.uint8), g[f"q_{key}"])
    assert float(g["scale_token"][2]) == float(torch.tensor(1.0 / (448.0 * 512.0), dtype=torch.float32))


def test_fp8_linear_is_the_dequantised_product():
    """fp8_linear == fp32 matmul of the dequantised operands (the definition cutlass_scaled_mm
    implements), and fp8 model mode stays close to -- but is not -- the bf16 model."""
    g = torch.Generator().manual_seed(3)
    x = torch.randn(6, 256, generator=g).to(BF)
    w = (torch.randn(40, 256, generator=g) * 0.05).to(BF)
    wq, ws = O.quant_fp8(w, per_token=False)
    y = O.fp8_linear(x, wq, ws)
    xq, xs = O.quant_fp8(x)
    ref = ((xq.float() * xs) @ (wq.float() * ws).t())
    assert float((y.float() - ref).abs().max()) <= 1e-2 * float(ref.abs().max())
    assert float((y.float() - O.linear(x, w).float()).abs().max()) <= 8e-2 * float(ref.abs().max())
    wts = O.init_weights(O.TINY, seed=7)
    p = list(range(5, 40))
    la = O.OracleLlama(O.TINY, wts).forward(0, p)[-1]
    lb = O.OracleLlama(O.TINY, wts, quantization="fp8").forward(0, p)[-1]
    d = float((la - lb).abs().max()) / float(la.abs().max())
    assert 1e-3 < d < 0.25


def test_mixtral_family_against_hf():
    """8 experts, top-2 softmax routing with renormalisation (BASELINE.json configs[4] family)."""
    g = torch.load(GOLD / "tiny_mixtral_hf.pt", weights_only=False)
    model = O.OracleLlama(O.TINY_MIXTRAL, O.init_weights(O.TINY_MIXTRAL, seed=7))
    logits = model.forward(0, g["prompt"])[-4:]
    scale = float(g["logits_fp32"].abs().max())
    assert float((logits - g["logits_fp32"]).abs().max()) <= 2e-2 * scale
    model.reset()
    mine = model.generate(0, g["prompt"], 10)
    agree = [a == b for a, b in zip(g["greedy_fp32"], g["greedy_bf16"])]
    n_stable = agree.index(False) if False in agree else len(agree)
    assert mine[:n_stable] == g["greedy_fp32"][:n_stable]


def test_moe_topk_ties_go_to_the_lowest_index():
    p = torch.tensor([[0.1, 0.3, 0.3, 0.3], [0.25, 0.25, 0.25, 0.25]])
    v, i = O.moe_topk(p, 2)
    assert i.tolist() == [[1, 2], [0, 1]] and torch.equal(v, torch.tensor([[0.3, 0.3], [0.25, 0.25]]))


def test_fp8_linear_against_vllm_torch_fallback():
    """oracle.fp8_linear == vLLM's ChannelWiseTorchFP8ScaledMMLinearKernel.apply_scaled_mm bit for
    bit (with and without bias): the restated W8A8 linear is the reference's own definition."""
    g = torch.load(GOLD / "fp8_linear_vllm.pt", weights_only=False)
    wq = g["wq"].view(torch.float8_e4m3fn)
    rq, rs = O.quant_fp8(g["w"], per_token=False)
    assert torch.equal(rq.view(torch.uint8), g["wq"]) and torch.equal(rs, g["w_scale"])
    assert torch.equal(O.fp8_linear(g["x"], wq, g["w_scale"]), g["y"])
    assert torch.equal(O.fp8_linear(g["x"], wq, g["w_scale"], g["bias"]), g["y_bias"])
