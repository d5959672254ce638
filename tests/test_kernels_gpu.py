"""Parity of every libb200.so kernel against the CPU oracle (oracle/llama_oracle.py), through the
C ABI, on seeded inputs.  Tolerances: bf16 results atol 1e-3*scale / rtol 1.6e-2 (vLLM's own bf16
tolerance, $VLLM/ir/tolerances.py:12-38); integer results (argmax, gathers, cache scatter) exact.
"""
import json
import math

import pytest
import torch

from oracle import llama_oracle as O

pytestmark = pytest.mark.gpu

BF = torch.bfloat16


@pytest.fixture(scope="module")
def ops():
    from gpustack_b200 import ops as _ops

    _ops.device_info()
    return _ops


def _close(got, ref, atol, rtol, what, out_dir=None):
    got, ref = got.float().cpu(), ref.float().cpu()
    err = (got - ref).abs()
    tol = atol + rtol * ref.abs()
    bad = err > tol
    if bad.any():
        idx = bad.nonzero()[:8].tolist()
        info = {
            "what": what, "max_err": float(err.max()), "ref_absmax": float(ref.abs().max()),
            "bad_frac": float(bad.float().mean()), "first_bad": idx,
            "got": [float(got[tuple(i)]) for i in idx], "ref": [float(ref[tuple(i)]) for i in idx],
        }
        if got.dim() == 2:
            info["bad_rows"] = bad.any(dim=1).nonzero().flatten()[:16].tolist()
            info["bad_cols"] = bad.any(dim=0).nonzero().flatten()[:16].tolist()
            info["n_bad_rows"] = int(bad.any(dim=1).sum()); info["n_bad_cols"] = int(bad.any(dim=0).sum())
        if out_dir is not None:
            with open(out_dir / "parity_failures.jsonl", "a") as f:
                f.write(json.dumps(info) + "\n")
        pytest.fail(json.dumps(info))


# --------------------------------------------------------------------------------- rmsnorm
@pytest.mark.parametrize("T,H", [(1, 4096), (7, 4096), (300, 4096), (5, 896), (3, 512), (2, 8192)])
def test_rms_norm(ops, T, H, out_dir):
    g = torch.Generator().manual_seed(T * 1000 + H)
    x = (torch.randn(T, H, generator=g) * 2).to(BF)
    w = (1 + 0.1 * torch.randn(H, generator=g)).to(BF)
    ref = O.rms_norm(x, w, 1e-5)
    got = ops.rms_norm(x.cuda(), w.cuda(), 1e-5)
    _close(got, ref, 1e-2, 1.6e-2, f"rms_norm {T}x{H}", out_dir)


@pytest.mark.parametrize("T,H", [(1, 4096), (33, 4096), (257, 4096), (4, 896)])
def test_fused_add_rms_norm(ops, T, H, out_dir):
    g = torch.Generator().manual_seed(T + H)
    x = torch.randn(T, H, generator=g).to(BF)
    r = torch.randn(T, H, generator=g).to(BF)
    w = (1 + 0.1 * torch.randn(H, generator=g)).to(BF)
    ref_y, ref_r = O.fused_add_rms_norm(x, r, w, 1e-5)
    xc, rc = x.cuda(), r.cuda()
    y, rr = ops.fused_add_rms_norm(xc, rc, w.cuda(), 1e-5, out=xc, residual_out=rc)  # in place
    _close(rr, ref_r, 0, 0, "fused_add residual (exact)", out_dir)
    _close(y, ref_y, 1e-2, 1.6e-2, "fused_add_rms_norm", out_dir)


def test_fused_add_rms_norm_partials(ops, out_dir):
    g = torch.Generator().manual_seed(5)
    S, T, H = 4, 19, 4096
    parts = torch.randn(S, T, H, generator=g)
    r = torch.randn(T, H, generator=g).to(BF)
    w = (1 + 0.1 * torch.randn(H, generator=g)).to(BF)
    z = parts.sum(0) + r.float()
    ref_r = z.to(BF)
    ref_y = ((z * torch.rsqrt(z.pow(2).mean(-1, keepdim=True) + 1e-5)).to(BF) * w)
    y = torch.empty(T, H, dtype=BF, device="cuda"); ro = torch.empty_like(y)
    ops.fused_add_rms_norm_partials(parts.cuda(), r.cuda(), w.cuda(), 1e-5, y, ro)
    _close(ro, ref_r, 1e-2, 8e-3, "partials residual", out_dir)
    _close(y, ref_y, 2e-2, 1.6e-2, "partials norm", out_dir)


# ------------------------------------------------------------------------ rope + kv write
@pytest.mark.parametrize("Hq,Hkv,D", [(32, 8, 128), (14, 2, 64), (4, 1, 128)])
def test_rope_kv_write(ops, Hq, Hkv, D, out_dir):
    g = torch.Generator().manual_seed(Hq)
    T, nblk, bs = 37, 12, 16
    qkv = torch.randn(T, (Hq + 2 * Hkv) * D, generator=g).to(BF)
    pos = torch.randint(0, 500, (T,), generator=g, dtype=torch.int32)
    cache = O.cos_sin_cache(D, 512, 500000.0)
    slots = torch.randperm(nblk * bs, generator=g)[:T].to(torch.int32)
    slots[3] = -1
    kc = torch.zeros(nblk, Hkv, bs, D, dtype=BF, device="cuda"); vc = torch.zeros_like(kc)
    q = qkv[:, : Hq * D].view(T, Hq, D)
    k = qkv[:, Hq * D : (Hq + Hkv) * D].view(T, Hkv, D)
    v = qkv[:, (Hq + Hkv) * D :].view(T, Hkv, D)
    rq, rk = O.rope_neox(q, pos, cache), O.rope_neox(k, pos, cache)
    dq = qkv.cuda()
    ops.rope_kv_write(dq, pos.cuda(), cache.cuda(), kc, vc, slots.cuda(), Hq, Hkv, D)
    got = dq.cpu()
    _close(got[:, : Hq * D].view(T, Hq, D), rq, 1e-2, 8e-3, "rope q", out_dir)
    _close(got[:, Hq * D : (Hq + Hkv) * D].view(T, Hkv, D), rk, 1e-2, 8e-3, "rope k", out_dir)
    assert torch.equal(got[:, (Hq + Hkv) * D :], qkv[:, (Hq + Hkv) * D :])
    kcc, vcc = kc.cpu(), vc.cpu()
    gk = got[:, Hq * D : (Hq + Hkv) * D].view(T, Hkv, D)
    for t in range(T):
        s = int(slots[t])
        if s < 0:
            continue
        assert torch.equal(kcc[s // bs, :, s % bs], gk[t]), f"k cache slot {s}"
        assert torch.equal(vcc[s // bs, :, s % bs], v[t]), f"v cache slot {s}"
    used = set(int(s) for s in slots if s >= 0)
    for s in range(nblk * bs):
        if s not in used:
            assert float(kcc[s // bs, :, s % bs].abs().sum()) == 0.0


def test_rope_kv_write_from_splitk_partials(ops, out_dir):
    """split-K QKV partials reduced inside the RoPE kernel == reduce-to-bf16 then RoPE."""
    g = torch.Generator().manual_seed(77)
    T, Hq, Hkv, D, S, nblk = 21, 32, 8, 128, 4, 8
    W = (Hq + 2 * Hkv) * D
    parts = torch.randn(S, T, W, generator=g) * 0.5
    pos = torch.randint(0, 400, (T,), generator=g, dtype=torch.int32)
    cache = O.cos_sin_cache(D, 512, 500000.0)
    slots = torch.randperm(nblk * 16, generator=g)[:T].to(torch.int32); slots[5] = -1
    kc1 = torch.zeros(nblk, Hkv, 16, D, dtype=BF, device="cuda"); vc1 = torch.zeros_like(kc1)
    kc2 = torch.zeros_like(kc1); vc2 = torch.zeros_like(kc1)
    acc = parts[0].clone()
    for s_ in range(1, S):
        acc = acc + parts[s_]          # same sequential fp32 order as the kernel
    ref_qkv = acc.to(BF).cuda()
    ops.rope_kv_write(ref_qkv, pos.cuda(), cache.cuda(), kc1, vc1, slots.cuda(), Hq, Hkv, D)
    out = torch.empty(T, W, dtype=BF, device="cuda")
    ops.rope_kv_write_partials(parts.cuda(), out, pos.cuda(), cache.cuda(), kc2, vc2, slots.cuda(), Hq, Hkv, D)
    assert torch.equal(out, ref_qkv) and torch.equal(kc1, kc2) and torch.equal(vc1, vc2)
    # with a QKV bias (Qwen2): added in fp32 before the single bf16 rounding
    bias = (torch.randn(W, generator=g) * 0.5).to(BF)
    ref_b = (acc + bias.float()).to(BF).cuda()
    kc3 = torch.zeros_like(kc1); vc3 = torch.zeros_like(kc1); kc4 = torch.zeros_like(kc1); vc4 = torch.zeros_like(kc1)
    ops.rope_kv_write(ref_b, pos.cuda(), cache.cuda(), kc3, vc3, slots.cuda(), Hq, Hkv, D)
    ops.rope_kv_write_partials(parts.cuda(), out, pos.cuda(), cache.cuda(), kc4, vc4, slots.cuda(), Hq, Hkv, D,
                               bias=bias.cuda())
    assert torch.equal(out, ref_b) and torch.equal(kc3, kc4) and torch.equal(vc3, vc4)


@pytest.mark.parametrize("T,I", [(1, 14336), (65, 14336), (9, 4864), (300, 1024)])
def test_silu_and_mul(ops, T, I, out_dir):
    g = torch.Generator().manual_seed(I)
    x = (torch.randn(T, 2 * I, generator=g) * 2).to(BF)
    ref = O.silu_and_mul(x)
    _close(ops.silu_and_mul(x.cuda()), ref, 1e-2, 1.6e-2, "silu_and_mul", out_dir)


def test_embedding_gather_argmax(ops):
    g = torch.Generator().manual_seed(1)
    V, H, T = 1000, 512, 77
    table = torch.randn(V, H, generator=g).to(BF)
    ids = torch.randint(0, V, (T,), generator=g, dtype=torch.int32)
    got = ops.embedding(ids.cuda(), table.cuda()).cpu()
    assert torch.equal(got, table[ids.long()])
    # vocab-parallel shard: rows outside [300, 600) come back as zeros
    shard = ops.embedding(ids.cuda(), table[300:600].contiguous().cuda(), vocab_start=300, vocab_end=600).cpu()
    ref = table[ids.long()].clone(); ref[(ids < 300) | (ids >= 600)] = 0
    assert torch.equal(shard, ref)
    rows = torch.randint(0, T, (13,), generator=g, dtype=torch.int32)
    assert torch.equal(ops.gather_rows(got.cuda(), rows.cuda()).cpu(), got[rows.long()])
    for B, Vv in [(1, 128256), (5, 128256), (3, 1001), (2, 7)]:
        lg = torch.randn(B, Vv, generator=g)
        lg[0, Vv // 2] = 50.0; lg[0, Vv - 1] = 50.0  # tie: lowest index must win
        idx = ops.argmax(lg.cuda()).cpu()
        assert idx.tolist() == [int(torch.argmax(r)) if i else Vv // 2 for i, r in enumerate(lg)]


def test_one_pass_sampler_is_greedy_at_t0_exact_softmax_above_and_shards_compose(ops):
    """b200_sample_f32: temperature 0 rows == argmax (lowest index on ties); temperature T rows are
    Gumbel-max draws whose empirical distribution matches softmax(logits / T); the draw is a pure
    function of (seed, global vocabulary index), so two vocabulary shards combined by their winning
    keys (what tp_argmax_select_kernel does across ranks) reproduce the unsharded token exactly."""
    g = torch.Generator().manual_seed(5)
    V, N = 384, 40000
    logits = (torch.randn(V, generator=g) * 2.0)
    logits[7] = logits[300] = logits.max() + 0.5          # a tie for greedy: index 7 must win
    T = 0.8
    rows = logits.repeat(N, 1).contiguous().cuda()
    temps = torch.full((N,), T, device="cuda"); temps[:100] = 0.0
    seeds = torch.arange(N, dtype=torch.int64).mul(2654435761).remainder(2 ** 31).to(torch.int32).cuda()
    vals = torch.empty(N, device="cuda")
    idx = ops.sample(rows, temps, seeds, out_val=vals).cpu()
    assert idx[:100].tolist() == [7] * 100
    again = ops.sample(rows, temps, seeds).cpu()
    assert torch.equal(idx, again)                          # deterministic in (seed, row)
    p = torch.softmax(logits.double() / T, dim=0)
    freq = torch.bincount(idx[100:].long(), minlength=V).double() / (N - 100)
    tv = 0.5 * float((freq - p).abs().sum())
    assert tv < 0.05, tv                                    # sampling noise at N = 4e4, V = 384 is ~0.03
    assert abs(float(freq[7]) - float(p[7])) < 0.01 and abs(float(freq[300]) - float(p[300])) < 0.01
    # vocabulary-parallel composition: shard [0, 160) and [160, 384) with global offsets
    i0, v0 = torch.empty(N, dtype=torch.int32, device="cuda"), torch.empty(N, device="cuda")
    i1, v1 = torch.empty(N, dtype=torch.int32, device="cuda"), torch.empty(N, device="cuda")
    ops.sample(rows[:, :160].contiguous(), temps, seeds, out_idx=i0, out_val=v0, vocab_offset=0)
    ops.sample(rows[:, 160:].contiguous(), temps, seeds, out_idx=i1, out_val=v1, vocab_offset=160)
    combined = torch.where(v1 > v0, i1 + 160, i0).cpu()
    assert torch.equal(combined, idx)


# ------------------------------------------------------------------------------------ GEMM
GEMM_SHAPES = [
    # M, N, K, splits, f32
    (1, 256, 128, 1, 0), (16, 128, 64, 1, 0), (5, 384, 512, 1, 0),
    (1, 4096, 4096, 1, 0), (16, 6144, 4096, 1, 0), (33, 4096, 4096, 1, 0), (64, 4096, 14336, 1, 0),
    (100, 28672, 4096, 1, 0), (128, 4096, 4096, 1, 0), (256, 6144, 4096, 1, 0),
    (200, 1000, 896, 1, 0),        # N tail (not a multiple of 128), M tail
    (64, 16032, 4096, 1, 0),       # vocab shard of TP=8
    (512, 6144, 4096, 1, 0), (1000, 4096, 4096, 1, 0), (2048, 4096, 14336, 1, 0),
    (8, 128256, 4096, 1, 1),       # lm_head, fp32 logits
    (64, 4096, 14336, 4, 1), (7, 4096, 4096, 8, 1), (256, 4096, 4096, 2, 1),  # split-K partials
]


@pytest.mark.parametrize("M,N,K,splits,f32", GEMM_SHAPES)
def test_gemm(ops, M, N, K, splits, f32, out_dir):
    g = torch.Generator(device="cuda").manual_seed(M * 7 + N + K)
    x = torch.randn(M, K, generator=g, device="cuda").to(BF)
    w = (torch.randn(N, K, generator=g, device="cuda") * 0.02).to(BF)
    ref = x.double() @ w.double().t()
    scale = 0.02 * math.sqrt(K)
    if splits > 1:
        parts = ops.gemm(x, w, splits=splits)
        assert parts.shape == (splits, M, N) and parts.dtype == torch.float32
        _close(parts.double().sum(0), ref, 2e-3 * scale, 2e-3, f"gemm splitK {M}x{N}x{K}/{splits}", out_dir)
        y = ops.splitk_reduce(parts)
        _close(y, ref, 1e-2 * scale, 1.6e-2, "splitk_reduce", out_dir)
    elif f32:
        y = ops.gemm(x, w, out_dtype=torch.float32)
        _close(y, ref, 2e-3 * scale, 2e-3, f"gemm f32 {M}x{N}x{K}", out_dir)
    else:
        y = ops.gemm(x, w)
        _close(y, ref, 1e-2 * scale, 1.6e-2, f"gemm {M}x{N}x{K}", out_dir)


@pytest.mark.parametrize("M,I,K", [(1, 256, 128), (7, 1024, 512), (64, 14336, 4096), (200, 1792, 4096),
                                   (256, 14336, 4096), (1000, 14336, 4096), (33, 104, 64)])
def test_gemm_silu_mul_fused_equals_unfused(ops, M, I, K, out_dir):
    """gate_up GEMM with the SiLU*mul epilogue is bit-identical to GEMM followed by silu_and_mul,
    and within tolerance of the CPU oracle."""
    g = torch.Generator(device="cuda").manual_seed(M + I)
    x = torch.randn(M, K, generator=g, device="cuda").to(BF)
    wg = (torch.randn(I, K, generator=g, device="cuda") * 0.03).to(BF)
    wu = (torch.randn(I, K, generator=g, device="cuda") * 0.03).to(BF)
    fused = ops.gemm_silu_mul(x, ops.interleave_gate_up(wg, wu))
    unfused = ops.silu_and_mul(ops.gemm(x, torch.cat([wg, wu], dim=0).contiguous()))
    assert torch.equal(fused, unfused)
    if M * I <= 64 * 14336:
        ref = O.silu_and_mul(torch.cat([O.linear(x.cpu(), wg.cpu()), O.linear(x.cpu(), wu.cpu())], dim=-1))
        _close(fused, ref, 2e-2, 3.2e-2, f"gemm_silu_mul {M}x{I}x{K}", out_dir)


# ------------------------------------------------------------------------------- fp8 (W8A8)
@pytest.mark.parametrize("T,K", [(1, 4096), (7, 896), (300, 14336), (33, 512), (5, 28672)])
def test_quant_fp8_per_token_is_bit_exact(ops, T, K):
    """bytes and scales equal the oracle's (= vLLM QuantFP8.forward_native, pinned by
    tests/golden/fp8_quant_vllm.pt); includes an all-zero row, an outlier row and a tiny row."""
    g = torch.Generator().manual_seed(T + K)
    x = (torch.randn(T, K, generator=g) * 2.0).to(BF)
    if T > 4:
        x[1] = 0
        x[2, 3] = 3000.0
        x[3] = x[3] * 1e-4
    q, s = ops.quant_fp8_per_token(x.cuda())
    rq, rs = O.quant_fp8(x, per_token=True)
    assert torch.equal(s.cpu(), rs.reshape(-1))
    assert torch.equal(q.cpu().view(torch.uint8), rq.view(torch.uint8))
    # strided input (a column slice) and a wider output buffer
    big = (torch.randn(T, K + 64, generator=g)).to(BF).cuda()
    q2, s2 = ops.quant_fp8_per_token(big[:, 64:])
    rq2, rs2 = O.quant_fp8(big[:, 64:].cpu(), per_token=True)
    assert torch.equal(s2.cpu(), rs2.reshape(-1)) and torch.equal(q2.cpu().view(torch.uint8), rq2.view(torch.uint8))


FP8_GEMM_SHAPES = [
    # M, N, K, splits, f32, per_channel
    (1, 256, 128, 1, 0, 0), (5, 384, 512, 1, 0, 0), (16, 6144, 4096, 1, 0, 0), (33, 4096, 4096, 1, 0, 1),
    (64, 4096, 14336, 1, 0, 0), (100, 28672, 4096, 1, 0, 0), (128, 4096, 4096, 1, 0, 0),
    (256, 6144, 4096, 1, 0, 0), (200, 1000, 896, 1, 0, 1),           # M / N tails
    (512, 6144, 4096, 1, 0, 0), (1000, 4096, 4096, 1, 0, 0), (2048, 4096, 14336, 1, 0, 0),
    (8, 1280, 8192, 1, 1, 0),                                        # fp32 output (QKV + bias path)
    (64, 4096, 14336, 4, 1, 0), (7, 4096, 4096, 8, 1, 1),            # split-K partials
]


@pytest.mark.parametrize("M,N,K,splits,f32,per_channel", FP8_GEMM_SHAPES)
def test_gemm_fp8(ops, M, N, K, splits, f32, per_channel, out_dir):
    """e4m3 x e4m3 tcgen05 GEMM with the dequantising epilogue == exact fp64 product of the same
    quantised operands times the scales (the only slack is fp32 accumulation order + one bf16
    rounding; same tolerances as the bf16 GEMM test)."""
    g = torch.Generator(device="cuda").manual_seed(M * 5 + N + K)
    x = torch.randn(M, K, generator=g, device="cuda").to(BF)
    w = (torch.randn(N, K, generator=g, device="cuda") * 0.02).to(BF)
    xq, xs = ops.quant_fp8_per_token(x)
    if per_channel:
        ws = (w.float().abs().amax(dim=1) / 448.0).clamp(min=1e-8)
        wq = (w.float() / ws[:, None]).clamp(-448, 448).to(torch.float8_e4m3fn)
    else:
        ws = (w.float().abs().max() / 448.0).reshape(1)
        wq = (w.float() / ws).clamp(-448, 448).to(torch.float8_e4m3fn)
    ref = (xq.double() @ wq.double().t()) * xs.double()[:, None] * ws.double()[None, :]
    scale = 0.02 * math.sqrt(K)
    if splits > 1:
        parts = ops.gemm_fp8(xq, xs, wq, ws, splits=splits)
        assert parts.shape == (splits, M, N) and parts.dtype == torch.float32
        _close(parts.double().sum(0), ref, 2e-3 * scale, 2e-3, f"gemm_fp8 splitK {M}x{N}x{K}/{splits}", out_dir)
    elif f32:
        y = ops.gemm_fp8(xq, xs, wq, ws, out_dtype=torch.float32)
        _close(y, ref, 2e-3 * scale, 2e-3, f"gemm_fp8 f32 {M}x{N}x{K}", out_dir)
    else:
        y = ops.gemm_fp8(xq, xs, wq, ws)
        _close(y, ref, 1e-2 * scale, 1.6e-2, f"gemm_fp8 {M}x{N}x{K}", out_dir)
        # and it tracks the unquantised bf16 product to fp8 accuracy
        full = x.double() @ w.double().t()
        assert float((y.double() - full).abs().max()) <= 0.15 * float(full.abs().max())


@pytest.mark.parametrize("M,I,K", [(1, 256, 128), (7, 1024, 512), (64, 14336, 4096), (256, 14336, 4096),
                                   (1000, 3584, 8192)])
def test_gemm_fp8_silu_mul_fused_equals_unfused(ops, M, I, K, out_dir):
    g = torch.Generator(device="cuda").manual_seed(M + I + 1)
    x = torch.randn(M, K, generator=g, device="cuda").to(BF)
    wg = (torch.randn(I, K, generator=g, device="cuda") * 0.03).to(BF)
    wu = (torch.randn(I, K, generator=g, device="cuda") * 0.03).to(BF)
    xq, xs = ops.quant_fp8_per_token(x)
    ws = (torch.cat([wg, wu]).float().abs().max() / 448.0).reshape(1)
    q = lambda t: (t.float() / ws).clamp(-448, 448).to(torch.float8_e4m3fn)      # noqa: E731
    fused = ops.gemm_fp8(xq, xs, ops.interleave_gate_up(q(wg).view(torch.uint8), q(wu).view(torch.uint8))
                         .view(torch.float8_e4m3fn), ws, silu_mul=True)
    unfused = ops.silu_and_mul(ops.gemm_fp8(xq, xs, torch.cat([q(wg), q(wu)], dim=0).contiguous(), ws))
    assert fused.shape == (M, I) and torch.equal(fused, unfused)


def test_gemm_fp8_vs_cpu_oracle(ops, out_dir):
    g = torch.Generator().manual_seed(3)
    x = torch.randn(40, 512, generator=g).to(BF); w = (torch.randn(640, 512, generator=g) * 0.05).to(BF)
    wq, ws = O.quant_fp8(w, per_token=False)
    xq, xs = ops.quant_fp8_per_token(x.cuda())
    y = ops.gemm_fp8(xq, xs, wq.cuda(), ws.cuda())
    _close(y, O.fp8_linear(x, wq, ws), 1e-2, 1.6e-2, "gemm_fp8 vs oracle.fp8_linear", out_dir)


# ------------------------------------------------------------------------------------- MoE
@pytest.mark.parametrize("T,E,K", [(1, 8, 2), (7, 8, 2), (64, 8, 2), (300, 8, 2), (33, 4, 1), (129, 16, 4)])
def test_moe_route_and_align(ops, T, E, K):
    """routing == oracle (softmax in fp32, top-k with ties to the lowest index, renormalised), and the
    permuted row space is a valid grouping: every pair owns one row inside a tile of its expert."""
    from gpustack_b200.lib import load

    g = torch.Generator().manual_seed(T + E)
    logits = (torch.randn(T, 128, generator=g) * 2).to(BF)          # padded router output, E live columns
    logits[0, :E] = 0.5                                              # an all-tie row
    ids, w = ops.moe_route(logits.cuda(), K, num_experts=E)
    rw, ri = O.moe_topk(torch.softmax(logits[:, :E].float(), dim=-1), K)
    rw = rw / rw.sum(dim=-1, keepdim=True)
    assert torch.equal(ids.cpu().view(T, K), ri.to(torch.int32))
    assert float((w.cpu().view(T, K) - rw).abs().max()) <= 1e-6
    for bx in (32, 128):
        n_pairs = T * K
        cap = ops.moe_capacity_rows(n_pairs, E, bx)
        i32 = lambda n: torch.empty(n, dtype=torch.int32, device="cuda")          # noqa: E731
        pair_row, row_pair, tile_expert, num_tiles = i32(n_pairs), i32(cap), i32(cap // bx), torch.zeros(2, dtype=torch.int32, device="cuda")
        row_w = torch.empty(cap, dtype=torch.float32, device="cuda")
        rc = load().b200_moe_align(ids.data_ptr(), w.data_ptr(), n_pairs, E, bx, cap, pair_row.data_ptr(), row_pair.data_ptr(),
                                   row_w.data_ptr(), tile_expert.data_ptr(), num_tiles.data_ptr(),
                                   torch.cuda.current_stream().cuda_stream)
        assert rc == 0
        pr, rp, te, nt, rwt = pair_row.cpu(), row_pair.cpu(), tile_expert.cpu(), num_tiles.cpu().tolist(), row_w.cpu()
        counts = torch.bincount(ids.cpu().long(), minlength=E)
        assert nt[0] == int(((counts + bx - 1) // bx).sum()) and nt[1] == nt[0] * bx <= cap
        assert len(set(pr.tolist())) == n_pairs and int(pr.max()) < nt[1]
        for i in range(n_pairs):
            r = int(pr[i])
            assert int(rp[r]) == i and int(te[r // bx]) == int(ids[i]) and float(rwt[r]) == float(w[i])
        assert int((rp[: nt[1]] >= 0).sum()) == n_pairs and bool((rwt[rp < 0] == 0).all())
        assert te[: nt[0]].tolist() == sorted(te[: nt[0]].tolist())


@pytest.mark.parametrize("T,E,K,H,I", [(1, 8, 2, 512, 256), (7, 8, 2, 512, 256), (64, 8, 2, 512, 256),
                                         (300, 8, 2, 512, 256), (40, 4, 1, 1024, 512), (257, 8, 2, 4096, 1792),
                                         (2100, 8, 2, 1024, 512)])
def test_moe_experts_vs_oracle(ops, T, E, K, H, I, out_dir):
    """route -> align -> gather -> grouped gate_up+SiLU -> grouped down (x routing weight) -> combine
    against the oracle's per-expert loop (same roundings: bf16 gate/up, bf16 act, fp32 acc x weight
    -> bf16 row, fp32 sum of the k rows -> bf16)."""
    g = torch.Generator().manual_seed(T * 3 + E + H)
    x = torch.randn(T, H, generator=g).to(BF)
    router = (torch.randn(E, H, generator=g) * 0.2).to(BF)
    w1 = [(torch.randn(I, H, generator=g) * 0.03).to(BF) for _ in range(E)]
    w3 = [(torch.randn(I, H, generator=g) * 0.03).to(BF) for _ in range(E)]
    w2 = [(torch.randn(H, I, generator=g) * 0.03).to(BF) for _ in range(E)]
    ws = ops.MoEWorkspace(T, K, E, H, I, "cuda")
    xc = x.cuda()
    logits = ops.gemm(xc, ops.pad_router(router.cuda()), out=ws.logits[:T])
    _close(logits[:, :E], O.linear(x, router), 1e-2, 1.6e-2, "router logits (padded one-tile GEMM)", out_dir)
    assert not logits[:, E:].any()                                  # zero rows of the padded router weight
    w13 = torch.cat([ops.interleave_gate_up(w1[e], w3[e]) for e in range(E)]).cuda()
    w2s = torch.cat(w2).cuda()
    out = torch.empty(T, H, dtype=BF, device="cuda")
    ops.moe_experts(xc, logits, w13, w2s, K, ws, out)
    # oracle on the GPU's own router logits (a near-tie between experts is a legitimate flip otherwise)
    lg = logits[:, :E].cpu()
    topw, topi = O.moe_topk(torch.softmax(lg.float(), dim=-1), K)
    topw = topw / topw.sum(dim=-1, keepdim=True)
    ref = torch.zeros(T, H)
    for e in range(E):
        tok, slot = torch.nonzero(topi == e, as_tuple=True)
        if tok.numel() == 0:
            continue
        act = O.silu_and_mul(torch.cat((O.linear(x[tok], w1[e]), O.linear(x[tok], w3[e])), dim=-1))
        y = (act.float() @ w2[e].float().t()) * topw[tok, slot].unsqueeze(-1)
        ref.index_add_(0, tok, y.to(BF).float())
    _close(out, ref.to(BF), 2e-2 * float(ref.abs().max()), 3.2e-2, f"moe_experts T={T} E={E} K={K}", out_dir)
    nt = ws.num_tiles.cpu().tolist()
    assert nt[1] >= T * K


def test_gemm_vs_cpu_oracle(ops, out_dir):
    g = torch.Generator().manual_seed(3)
    x = torch.randn(40, 512, generator=g).to(BF); w = (torch.randn(640, 512, generator=g) * 0.05).to(BF)
    _close(ops.gemm(x.cuda(), w.cuda()), O.linear(x, w), 1e-2, 1.6e-2, "gemm vs oracle.linear", out_dir)


def test_gemm_strided_views(ops, out_dir):
    """column-sliced activations (q part of qkv) and row-sliced weights are passed as views."""
    g = torch.Generator(device="cuda").manual_seed(9)
    big = torch.randn(50, 1024, generator=g, device="cuda").to(BF)
    w = (torch.randn(300, 2048, generator=g, device="cuda") * 0.03).to(BF)
    x = big[:, 256:768]            # ld 1024, K 512
    wv = w[:, 512:1024]            # ld 2048
    ref = x.double() @ wv.double().t()
    _close(ops.gemm(x, wv), ref, 1e-2, 1.6e-2, "gemm strided", out_dir)


# ------------------------------------------------------------------------------- attention
def _make_paged(g, seq_lens, Hkv, D, extra_blocks=5):
    bs = 16
    n_pages = [(n + bs - 1) // bs for n in seq_lens]
    total = sum(n_pages) + extra_blocks
    perm = torch.randperm(total, generator=g)
    max_blocks = max(n_pages) + 2
    bt = torch.zeros(len(seq_lens), max_blocks, dtype=torch.int32)
    kc = torch.randn(total, Hkv, bs, D, generator=g).to(BF)
    vc = torch.randn(total, Hkv, bs, D, generator=g).to(BF)
    off = 0
    for i, n in enumerate(n_pages):
        bt[i, :n] = perm[off : off + n].to(torch.int32)
        off += n
    return kc, vc, bt


@pytest.mark.parametrize("Hq,Hkv,D,seq_lens,splits", [
    (32, 8, 128, [1, 17, 576, 33, 640, 100], 1),
    (32, 8, 128, [2000, 5, 1023], 4),
    (14, 2, 64, [77, 300, 16, 15], 1),
    (14, 2, 64, [900, 31], 3),
    (4, 1, 128, [8192], 16),
    (8, 8, 128, [50, 130], 1),
    (64, 8, 128, [300, 17, 2049], 1),                     # Llama-3-70B heads: group of 8
    (16, 2, 128, [4100, 640], 4),
])
def test_paged_attn_decode(ops, Hq, Hkv, D, seq_lens, splits, out_dir):
    g = torch.Generator().manual_seed(sum(seq_lens))
    B = len(seq_lens)
    kc, vc, bt = _make_paged(g, seq_lens, Hkv, D)
    stride = (Hq + 2 * Hkv) * D
    qkv = torch.randn(B, stride, generator=g).to(BF)
    scale = D ** -0.5
    ref = torch.empty(B, Hq * D, dtype=BF)
    for b, n in enumerate(seq_lens):
        k = O.gather_paged(kc, bt[b].tolist(), n); v = O.gather_paged(vc, bt[b].tolist(), n)
        q = qkv[b, : Hq * D].view(1, Hq, D)
        ref[b] = O.attention_ref(q, k, v, scale, torch.tensor([n - 1])).reshape(-1)
    dq = qkv.cuda()
    out = torch.empty(B, Hq * D, dtype=BF, device="cuda")
    ws = ops.attn_decode_workspace(B, Hq, D, splits, "cuda")
    ops.paged_attn_decode(dq, stride, kc.cuda(), vc.cuda(), bt.cuda(), torch.tensor(seq_lens, dtype=torch.int32).cuda(),
                          out, Hq, Hkv, D, scale, num_splits=splits, workspace=ws)
    _close(out, ref, 2e-2, 1.6e-2, f"attn_decode Hq{Hq} D{D} splits{splits}", out_dir)


@pytest.mark.parametrize("Hq,Hkv,D,ctx_q", [
    (32, 8, 128, [(0, 512), (0, 1), (0, 33), (0, 100)]),
    (32, 8, 128, [(512, 64), (100, 17), (31, 300)]),     # chunked prefill / prefix continuation
    (14, 2, 64, [(0, 130), (20, 50)]),
    (4, 1, 128, [(0, 1500)]),
    (16, 2, 128, [(0, 70), (5, 9)]),                      # group of 8 -> 16-token tiles
    # 528 (tile, kv-head) work items > 2 CTAs x 148 SMs: every CTA of the persistent kernel walks several
    # tiles with different (odd and even) block counts, so the running barrier phases carry across tiles
    (32, 8, 128, [(0, 700), (64, 600), (0, 300), (17, 450)]),
])
def test_paged_attn_prefill(ops, Hq, Hkv, D, ctx_q, out_dir):
    g = torch.Generator().manual_seed(len(ctx_q) * 31 + Hq)
    totals = [c + n for c, n in ctx_q]
    kc, vc, bt = _make_paged(g, totals, Hkv, D)
    q_lens = [n for _, n in ctx_q]
    T = sum(q_lens)
    stride = (Hq + 2 * Hkv) * D
    qkv = torch.randn(T, stride, generator=g).to(BF)
    scale = D ** -0.5
    cu = [0]
    for n in q_lens:
        cu.append(cu[-1] + n)
    ref = torch.empty(T, Hq * D, dtype=BF)
    for s, (c, n) in enumerate(ctx_q):
        k = O.gather_paged(kc, bt[s].tolist(), c + n); v = O.gather_paged(vc, bt[s].tolist(), c + n)
        q = qkv[cu[s] : cu[s + 1], : Hq * D].reshape(n, Hq, D)
        ref[cu[s] : cu[s + 1]] = O.attention_ref(q, k, v, scale, torch.arange(c, c + n)).reshape(n, -1)
    tile = ops.attn_prefill_tile(Hq, Hkv)
    ts, tq = ops.build_prefill_tiles(q_lens, tile)
    i32 = lambda a: torch.tensor(a, dtype=torch.int32, device="cuda")
    out = torch.zeros(T, Hq * D, dtype=BF, device="cuda")
    ops.paged_attn_prefill(qkv.cuda(), stride, kc.cuda(), vc.cuda(), bt.cuda(), i32(cu), i32([c for c, _ in ctx_q]),
                           i32(ts), i32(tq), out, Hq, Hkv, D, scale)
    _close(out, ref, 2e-2, 1.6e-2, f"attn_prefill Hq{Hq} D{D}", out_dir)


@pytest.mark.parametrize("batch,seqlen", [(16, 512), (4, 2048)])
def test_paged_attn_prefill_is_race_free_at_bench_shape(ops, batch, seqlen):
    """BASELINE-sized prefill step (8192 tokens, 2048 work items on 296 persistent CTAs, hundreds of them
    single-block tiles processed back to back): 25 launches with a cold L2 must be bit-identical (the kernel's
    reduction order is fixed), NaN-free, and right against plain fp32 attention.  Guards the softmax -> MMA
    hand-off: with ONE count-4 `p_full` barrier a fast softmax warp could arrive twice and PV consumed P rows
    that were not written yet -- a few NaN rows per launch at exactly this shape."""
    Hq, Hkv, D, page = 32, 8, 128, 16
    dev = torch.device("cuda")
    g = torch.Generator(device=dev).manual_seed(batch * 7919 + seqlen)
    pages_per = seqlen // page
    total = batch * pages_per + 1
    kc = torch.randn(total, Hkv, page, D, generator=g, device=dev).to(BF)
    vc = torch.randn(total, Hkv, page, D, generator=g, device=dev).to(BF)
    kc[0].zero_(); vc[0].zero_()                                   # block 0 = the engine's null block
    bt = (torch.randperm(total - 1, generator=g, device=dev) + 1).to(torch.int32).reshape(batch, pages_per).contiguous()
    stride = (Hq + 2 * Hkv) * D
    T = batch * seqlen
    qkv = torch.randn(T, stride, generator=g, device=dev).to(BF)
    cu = torch.arange(0, T + 1, seqlen, dtype=torch.int32, device=dev)
    ctx = torch.zeros(batch, dtype=torch.int32, device=dev)
    ts, tq = ops.build_prefill_tiles([seqlen] * batch, ops.attn_prefill_tile(Hq, Hkv))
    ts = torch.tensor(ts, dtype=torch.int32, device=dev); tq = torch.tensor(tq, dtype=torch.int32, device=dev)
    flush = torch.zeros(192 << 20, dtype=torch.uint8, device=dev)
    scale = D ** -0.5
    first = None
    for it in range(25):
        flush.add_(1)
        out = torch.full((T, Hq * D), float("nan"), dtype=BF, device=dev)
        ops.paged_attn_prefill(qkv, stride, kc, vc, bt, cu, ctx, ts, tq, out, Hq, Hkv, D, scale)
        torch.cuda.synchronize()
        bad = torch.isnan(out.float()).reshape(T, Hq, D).any(-1).nonzero()[:8].tolist()
        assert not bad, f"launch {it}: NaN (token, head) {bad}"
        if first is None:
            first = out
        else:
            assert torch.equal(out, first), f"launch {it} differs from launch 0"
    for sq, head in ((0, 0), (batch - 1, 13), (batch // 2, 31)):     # spot checks against fp32 attention
        kvh = head // (Hq // Hkv)
        pages = bt[sq].long()
        kk = kc[pages, kvh].reshape(-1, D).float(); vv = vc[pages, kvh].reshape(-1, D).float()
        qq = qkv[sq * seqlen:(sq + 1) * seqlen, head * D:(head + 1) * D].float()
        att = (qq @ kk.T) * scale
        att = att.masked_fill(torch.triu(torch.ones(seqlen, seqlen, dtype=torch.bool, device=dev), 1), float("-inf"))
        ref = torch.softmax(att, -1) @ vv
        got = first[sq * seqlen:(sq + 1) * seqlen, head * D:(head + 1) * D].float()
        assert (got - ref).abs().max().item() < 2e-2, (sq, head)
