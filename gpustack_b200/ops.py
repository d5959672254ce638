"""Operator wrappers: torch CUDA tensors in, libb200.so kernels on torch's current stream.

Each function mirrors one entry point of include/b200.h (and through it one operator of the
reference's vLLM backend).  torch is used for memory and streams only; there is deliberately no
eager/PyTorch fallback -- a missing library or a non-CUDA tensor raises.
"""
from __future__ import annotations

import ctypes as C
from typing import Optional, Sequence

import torch

from .lib import B200Error, check, load

BLOCK_SIZE = 16  # tokens per KV page (vLLM default block_size)


class KernelTimer:
    """Optional per-launch CUDA-event timing of the GEMM kernel (bench.py's live roofline).
    Events are recorded on the launching stream around each launch; nothing is recorded while a
    CUDA graph is being captured."""

    def __init__(self):
        self.records = []   # (tag, flops, bytes, start_event, end_event)

    def summary(self):
        torch.cuda.synchronize()
        agg = {}
        for tag, flops, nbytes, e0, e1 in self.records:
            a = agg.setdefault(tag, {"launches": 0, "ms": 0.0, "flops": 0.0, "bytes": 0.0})
            a["launches"] += 1; a["ms"] += e0.elapsed_time(e1); a["flops"] += flops; a["bytes"] += nbytes
        return agg


TIMER: Optional[KernelTimer] = None


def _stream() -> int:
    return torch.cuda.current_stream().cuda_stream


def _p(t: Optional[torch.Tensor]) -> Optional[int]:
    return None if t is None else t.data_ptr()


def _req(t: torch.Tensor, dtype, name: str) -> None:
    if not t.is_cuda:
        raise B200Error(f"{name}: expected a CUDA tensor (there is no CPU path)")
    if t.dtype != dtype:
        raise B200Error(f"{name}: expected {dtype}, got {t.dtype}")
    if not t.is_contiguous():
        raise B200Error(f"{name}: tensor must be contiguous")


def device_info() -> dict:
    lib = load()
    v = [C.c_int() for _ in range(4)]
    check(lib.b200_device_info(*[C.byref(x) for x in v]), "device_info")
    return {"sm_count": v[0].value, "cc": (v[1].value, v[2].value), "l2_bytes": v[3].value}


# ------------------------------------------------------------------------------ normalisation
def rms_norm(x: torch.Tensor, w: torch.Tensor, eps: float, out: Optional[torch.Tensor] = None):
    _req(x, torch.bfloat16, "x"); _req(w, torch.bfloat16, "w")
    T, H = x.shape
    y = torch.empty_like(x) if out is None else out
    check(load().b200_rms_norm(_p(x), _p(w), _p(y), T, H, eps, _stream()), "rms_norm")
    return y


def fused_add_rms_norm(x, residual, w, eps, out=None, residual_out=None):
    """returns (y, residual_out); pass out=x, residual_out=residual for the in-place form."""
    _req(x, torch.bfloat16, "x"); _req(residual, torch.bfloat16, "residual"); _req(w, torch.bfloat16, "w")
    T, H = x.shape
    y = torch.empty_like(x) if out is None else out
    r = torch.empty_like(x) if residual_out is None else residual_out
    check(load().b200_fused_add_rms_norm(_p(x), _p(residual), _p(w), _p(y), _p(r), T, H, eps, _stream()),
          "fused_add_rms_norm")
    return y, r


def fused_add_rms_norm_partials(partials, residual, w, eps, out, residual_out):
    _req(partials, torch.float32, "partials"); _req(w, torch.bfloat16, "w")
    S, T, H = partials.shape
    check(load().b200_fused_add_rms_norm_partials(_p(partials), S, _p(residual), _p(w), _p(out),
                                                  _p(residual_out), T, H, eps, _stream()),
          "fused_add_rms_norm_partials")
    return out, residual_out


# ------------------------------------------------------------------- rope + kv-cache scatter
def rope_kv_write(qkv, positions, cos_sin, k_cache, v_cache, slot_mapping, Hq, Hkv, D):
    _req(qkv, torch.bfloat16, "qkv"); _req(positions, torch.int32, "positions")
    _req(cos_sin, torch.bfloat16, "cos_sin"); _req(slot_mapping, torch.int32, "slot_mapping")
    _req(k_cache, torch.bfloat16, "k_cache"); _req(v_cache, torch.bfloat16, "v_cache")
    T = qkv.shape[0]
    if qkv.shape[1] != (Hq + 2 * Hkv) * D:
        raise B200Error("rope_kv_write: qkv width mismatch")
    check(load().b200_rope_kv_write(_p(qkv), _p(positions), _p(cos_sin), _p(k_cache), _p(v_cache),
                                    _p(slot_mapping), T, Hq, Hkv, D, k_cache.shape[2], _stream()),
          "rope_kv_write")


def rope_kv_write_partials(partials, qkv_out, positions, cos_sin, k_cache, v_cache, slot_mapping, Hq, Hkv, D,
                           bias=None):
    """partials [S,T,(Hq+2Hkv)*D] fp32 (split-K QKV GEMM) -> rotated bf16 qkv_out + cache scatter."""
    _req(partials, torch.float32, "partials"); _req(qkv_out, torch.bfloat16, "qkv_out")
    S, T, W = partials.shape
    if W != (Hq + 2 * Hkv) * D or qkv_out.shape[1] != W:
        raise B200Error("rope_kv_write_partials: qkv width mismatch")
    check(load().b200_rope_kv_write_partials(_p(partials), S, _p(bias), _p(qkv_out), _p(positions), _p(cos_sin), _p(k_cache),
                                             _p(v_cache), _p(slot_mapping), T, Hq, Hkv, D, k_cache.shape[2],
                                             _stream()), "rope_kv_write_partials")


def silu_and_mul(x, out=None):
    _req(x, torch.bfloat16, "x")
    T, I2 = x.shape
    y = torch.empty((T, I2 // 2), dtype=x.dtype, device=x.device) if out is None else out
    check(load().b200_silu_and_mul(_p(x), _p(y), T, I2 // 2, _stream()), "silu_and_mul")
    return y


def embedding(ids, table, out=None, vocab_start=0, vocab_end=None):
    _req(ids, torch.int32, "ids"); _req(table, torch.bfloat16, "table")
    T, H = ids.shape[0], table.shape[1]
    y = torch.empty((T, H), dtype=table.dtype, device=table.device) if out is None else out
    ve = vocab_start + table.shape[0] if vocab_end is None else vocab_end
    check(load().b200_embedding(_p(ids), _p(table), _p(y), T, H, vocab_start, ve, _stream()), "embedding")
    return y


def argmax(logits, out_idx=None, out_val=None):
    _req(logits, torch.float32, "logits")
    B, V = logits.shape
    idx = torch.empty((B,), dtype=torch.int32, device=logits.device) if out_idx is None else out_idx
    check(load().b200_argmax_f32(_p(logits), logits.stride(0), _p(idx), _p(out_val), B, V, _stream()),
          "argmax")
    return idx


def sample(logits, temps, seeds, out_idx=None, out_val=None, vocab_offset: int = 0):
    """Greedy (temps == 0) or Gumbel-max temperature sampling per row in ONE pass over the logits.
    temps: fp32 [B]; seeds: int32/uint32 [B] (bit patterns); returns int32 token ids (shard-local)."""
    _req(logits, torch.float32, "logits"); _req(temps, torch.float32, "temps"); _req(seeds, torch.int32, "seeds")
    B, V = logits.shape
    idx = torch.empty((B,), dtype=torch.int32, device=logits.device) if out_idx is None else out_idx
    check(load().b200_sample_f32(_p(logits), logits.stride(0), _p(temps), _p(seeds), vocab_offset, _p(idx), _p(out_val),
                                 B, V, _stream()), "sample")
    return idx


def gather_rows(x, rows, out=None):
    _req(x, torch.bfloat16, "x"); _req(rows, torch.int32, "rows")
    n, H = rows.shape[0], x.shape[1]
    y = torch.empty((n, H), dtype=x.dtype, device=x.device) if out is None else out
    check(load().b200_gather_rows(_p(x), _p(rows), _p(y), n, H, _stream()), "gather_rows")
    return y


# --------------------------------------------------------------------------------------- GEMM
def gemm_suggest_splits(M: int, N: int, K: int) -> int:
    return load().b200_gemm_suggest_splits(M, N, K)


def gemm(x, w, out=None, out_dtype=torch.bfloat16, splits: int = 1, block_x: int = 0):
    """y = x @ w.T (x [M,K], w [N,K], bf16).  splits>1 returns fp32 partials [splits, M, N]."""
    if not (x.is_cuda and w.is_cuda) or x.dtype != torch.bfloat16 or w.dtype != torch.bfloat16:
        raise B200Error("gemm: bf16 CUDA tensors required")
    if x.stride(-1) != 1 or w.stride(-1) != 1:
        raise B200Error("gemm: K must be the contiguous dimension")
    M, K = x.shape
    N = w.shape[0]
    out_f32 = 1 if (out_dtype == torch.float32 or splits > 1) else 0
    if out is None:
        shape = (splits, M, N) if splits > 1 else (M, N)
        out = torch.empty(shape, dtype=torch.float32 if out_f32 else torch.bfloat16, device=x.device)
    ldy = out.stride(-2)
    timed = TIMER is not None and not torch.cuda.is_current_stream_capturing()
    if timed:
        e0, e1 = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
        e0.record()
    check(load().b200_gemm_bf16_ex(_p(x), _p(w), _p(out), M, N, K, x.stride(0), w.stride(0), ldy, out_f32,
                                   splits, block_x, _stream()), "gemm")
    if timed:
        e1.record()
        tag = "gemm_prefill" if M > 256 else "gemm_decode"
        TIMER.records.append((tag, 2.0 * M * N * K, 2.0 * (N * K + M * K) + (4.0 if out_f32 else 2.0) * M * N * splits,
                              e0, e1))
    return out


def interleave_gate_up(gate: torch.Tensor, up: torch.Tensor) -> torch.Tensor:
    """[I,H] gate and up weights -> [2I,H] with row 2i = gate_i, row 2i+1 = up_i (the layout
    gemm_silu_mul expects)."""
    return torch.stack((gate, up), dim=1).reshape(2 * gate.shape[0], gate.shape[1]).contiguous()


def gemm_silu_mul(x, w_interleaved, out=None):
    """act = silu(x @ gate.T) * (x @ up.T) in one kernel; w_interleaved from interleave_gate_up."""
    if not (x.is_cuda and w_interleaved.is_cuda) or x.dtype != torch.bfloat16 or w_interleaved.dtype != torch.bfloat16:
        raise B200Error("gemm_silu_mul: bf16 CUDA tensors required")
    M, K = x.shape
    N = w_interleaved.shape[0]
    if out is None:
        out = torch.empty((M, N // 2), dtype=torch.bfloat16, device=x.device)
    timed = TIMER is not None and not torch.cuda.is_current_stream_capturing()
    if timed:
        e0, e1 = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
        e0.record()
    check(load().b200_gemm_bf16_silu_mul(_p(x), _p(w_interleaved), _p(out), M, N, K, x.stride(0),
                                         w_interleaved.stride(0), out.stride(0), _stream()), "gemm_silu_mul")
    if timed:
        e1.record()
        TIMER.records.append(("gemm_prefill" if M > 256 else "gemm_decode", 2.0 * M * N * K,
                              2.0 * (N * K + M * K) + 1.0 * M * N, e0, e1))
    return out


# --------------------------------------------------------------------------------------- fp8 (W8A8)
FP8 = torch.float8_e4m3fn


def quant_fp8_per_token(x, out=None, scales=None):
    """Dynamic per-token e4m3 quantisation: returns (q [T,K] float8_e4m3fn, scales [T] fp32)."""
    if not x.is_cuda or x.dtype != torch.bfloat16 or x.stride(-1) != 1:
        raise B200Error("quant_fp8_per_token: bf16 CUDA tensor with contiguous rows required")
    T, K = x.shape
    q = torch.empty((T, K), dtype=FP8, device=x.device) if out is None else out
    sc = torch.empty((T,), dtype=torch.float32, device=x.device) if scales is None else scales
    check(load().b200_quant_fp8_per_token(_p(x), x.stride(0), _p(q), q.stride(0), _p(sc), T, K, _stream()),
          "quant_fp8_per_token")
    return q, sc


def gemm_fp8(xq, xscale, wq, wscale, out=None, out_dtype=torch.bfloat16, splits: int = 1, silu_mul: bool = False):
    """y = (xq @ wq.T) * xscale[:,None] * wscale  (e4m3 operands, fp32 accumulate).  wscale: fp32
    tensor with 1 element (per-tensor) or N elements (per output channel).  silu_mul: wq holds
    interleaved gate/up rows and y = silu(gate) * up, [M, N/2] bf16."""
    if not (xq.is_cuda and wq.is_cuda) or xq.dtype != FP8 or wq.dtype != FP8:
        raise B200Error("gemm_fp8: float8_e4m3fn CUDA tensors required (there is no CPU path)")
    if xq.stride(-1) != 1 or wq.stride(-1) != 1:
        raise B200Error("gemm_fp8: K must be the contiguous dimension")
    _req(xscale, torch.float32, "xscale"); _req(wscale, torch.float32, "wscale")
    M, K = xq.shape
    N = wq.shape[0]
    if xscale.numel() != M or wscale.numel() not in (1, N):
        raise B200Error("gemm_fp8: xscale must have M elements, wscale 1 or N")
    out_f32 = 1 if (out_dtype == torch.float32 or splits > 1) else 0
    if out is None:
        shape = (splits, M, N) if splits > 1 else (M, N // 2 if silu_mul else N)
        out = torch.empty(shape, dtype=torch.float32 if out_f32 else torch.bfloat16, device=xq.device)
    check(load().b200_gemm_fp8(_p(xq), _p(wq), _p(out), M, N, K, xq.stride(0), wq.stride(0), out.stride(-2),
                               _p(xscale), _p(wscale), 0 if wscale.numel() == 1 else 1, out_f32, splits,
                               1 if silu_mul else 0, _stream()), "gemm_fp8")
    return out


# --------------------------------------------------------------------------------------- MoE
ROUTER_PAD = 128   # the router weight is zero-padded to one full weight tile of the GEMM


def pad_router(router: torch.Tensor) -> torch.Tensor:
    """[E, H] -> [128, H] with zero rows, so the router projection is an ordinary one-tile GEMM."""
    out = torch.zeros(ROUTER_PAD, router.shape[1], dtype=router.dtype, device=router.device)
    out[: router.shape[0]] = router
    return out


def moe_block_x(n_pairs: int, num_experts: int) -> int:
    """Row-tile width of the permuted MoE row space: about the mean rows per expert, so that the
    padding (< block_x rows per expert) stays a small fraction of the grouped GEMMs."""
    avg = n_pairs / max(1, num_experts)
    for bx in (256, 128, 64):
        if avg >= 0.75 * bx:
            return bx
    return 32


def moe_capacity_rows(n_pairs: int, num_experts: int, block_x: int) -> int:
    return load().b200_moe_capacity_rows(n_pairs, num_experts, block_x)


class MoEWorkspace:
    """Device buffers of one MoE dispatch (sized for `max_tokens` tokens, reused by every layer)."""

    def __init__(self, max_tokens: int, top_k: int, num_experts: int, H: int, I: int, device):
        self.top_k, self.E, self.H, self.I = top_k, num_experts, H, I
        n_pairs = max_tokens * top_k
        cap = max(moe_capacity_rows(n_pairs, num_experts, bx) for bx in (32, 64, 128, 256))
        i32 = lambda *s: torch.empty(*s, dtype=torch.int32, device=device)   # noqa: E731
        self.capacity = cap
        self.logits = torch.empty(max_tokens, ROUTER_PAD, dtype=torch.bfloat16, device=device)
        self.topk_ids, self.topk_w = i32(n_pairs), torch.empty(n_pairs, dtype=torch.float32, device=device)
        self.pair_row, self.row_pair = i32(n_pairs), i32(cap)
        self.row_weight = torch.empty(cap, dtype=torch.float32, device=device)
        self.tile_expert, self.num_tiles = i32(cap // 32 + 1), torch.zeros(2, dtype=torch.int32, device=device)
        self.xp = torch.empty(cap, H, dtype=torch.bfloat16, device=device)
        self.act = torch.empty(cap, I, dtype=torch.bfloat16, device=device)
        self.yp = torch.empty(cap, H, dtype=torch.bfloat16, device=device)


def moe_route(router_logits, top_k: int, renormalize: bool = True, out_ids=None, out_w=None,
              num_experts: Optional[int] = None):
    """router_logits [T, >=E] bf16 (columns beyond num_experts are padding of the router GEMM)."""
    if not router_logits.is_cuda or router_logits.dtype != torch.bfloat16 or router_logits.stride(1) != 1:
        raise B200Error("moe_route: bf16 CUDA logits with contiguous rows required")
    T = router_logits.shape[0]
    E = router_logits.shape[1] if num_experts is None else num_experts
    ids = torch.empty(T * top_k, dtype=torch.int32, device=router_logits.device) if out_ids is None else out_ids
    w = torch.empty(T * top_k, dtype=torch.float32, device=router_logits.device) if out_w is None else out_w
    check(load().b200_moe_route(_p(router_logits), router_logits.stride(0), _p(ids), _p(w), T, E, top_k,
                                1 if renormalize else 0, _stream()), "moe_route")
    return ids, w


def moe_experts(x, router_logits, w13, w2, top_k: int, ws: MoEWorkspace, out, renormalize: bool = True):
    """out[t] = sum_k weight_k * down_k(silu(gate_k x) * up_k x) over the top-k experts of token t.
    x [T,H] bf16; router_logits [T,>=E] bf16; w13 [E*2I, H] (per expert: interleaved gate/up rows);
    w2 [E*H, I].  Seven launches: route, align, gather, grouped gate_up+SiLU, grouped down (x routing
    weight), combine -- no host synchronisation."""
    T, H = x.shape
    E, I = ws.E, ws.I
    n_pairs = T * top_k
    bx = moe_block_x(n_pairs, E)
    cap = moe_capacity_rows(n_pairs, E, bx)
    if cap > ws.capacity:
        raise B200Error(f"moe_experts: {T} tokens exceed the workspace ({ws.capacity} rows)")
    L, st = load(), _stream()
    moe_route(router_logits, top_k, renormalize, out_ids=ws.topk_ids, out_w=ws.topk_w, num_experts=E)
    check(L.b200_moe_align(_p(ws.topk_ids), _p(ws.topk_w), n_pairs, E, bx, cap, _p(ws.pair_row), _p(ws.row_pair),
                           _p(ws.row_weight), _p(ws.tile_expert), _p(ws.num_tiles), st), "moe_align")
    check(L.b200_moe_gather(_p(x), _p(ws.row_pair), _p(ws.num_tiles), _p(ws.xp), H, top_k, cap, st), "moe_gather")
    check(L.b200_moe_grouped_gemm(_p(ws.xp), _p(w13), _p(ws.act), cap, 2 * I, H, E, H, w13.stride(0), I, None,
                                  _p(ws.tile_expert), _p(ws.num_tiles), bx, 1, st), "moe_grouped_gemm(gate_up)")
    check(L.b200_moe_grouped_gemm(_p(ws.act), _p(w2), _p(ws.yp), cap, H, I, E, I, w2.stride(0), H, _p(ws.row_weight),
                                  _p(ws.tile_expert), _p(ws.num_tiles), bx, 0, st), "moe_grouped_gemm(down)")
    check(L.b200_moe_combine(_p(ws.yp), _p(ws.pair_row), _p(out), T, H, top_k, st), "moe_combine")
    return out


MOE_LAUNCHES = 6


def splitk_reduce(partials, out=None):
    _req(partials, torch.float32, "partials")
    S, M, N = partials.shape
    y = torch.empty((M, N), dtype=torch.bfloat16, device=partials.device) if out is None else out
    check(load().b200_splitk_reduce(_p(partials), S, _p(y), M, N, y.stride(0), _stream()), "splitk_reduce")
    return y


# ---------------------------------------------------------------------------------- attention
def attn_decode_splits(B, Hkv, max_seq_len) -> int:
    return load().b200_attn_decode_splits(B, Hkv, max_seq_len, BLOCK_SIZE)


def attn_decode_workspace(B, Hq, D, num_splits, device) -> torch.Tensor:
    n = load().b200_attn_decode_workspace_bytes(B, Hq, D, num_splits)
    # zeroed ONCE: the first bytes hold the per-(sequence, kv-head) arrival counters of the in-kernel
    # split merge, which reset themselves after every launch
    return torch.zeros((n + 3) // 4, dtype=torch.float32, device=device)


def paged_attn_decode(q, q_stride, k_cache, v_cache, block_table, seq_lens, out, Hq, Hkv, D, scale,
                      num_splits=1, workspace=None):
    """q: [B, >=Hq*D] view into the qkv buffer (row stride q_stride elements)."""
    B = seq_lens.shape[0]
    _req(block_table, torch.int32, "block_table"); _req(seq_lens, torch.int32, "seq_lens")
    check(load().b200_paged_attn_decode(_p(q), q_stride, _p(k_cache), _p(v_cache), _p(block_table),
                                        block_table.shape[1], _p(seq_lens), _p(out), _p(workspace),
                                        num_splits, B, Hq, Hkv, D, k_cache.shape[2], scale, _stream()),
          "paged_attn_decode")
    return out


def attn_prefill_tile(Hq, Hkv) -> int:
    t = load().b200_attn_prefill_tile(Hq, Hkv)
    if t <= 0:
        raise B200Error(f"prefill attention: unsupported head grouping {Hq}/{Hkv}")
    return t


def build_prefill_tiles(q_lens: Sequence[int], tile: int, ctx_lens: Optional[Sequence[int]] = None):
    """Host-side tile list for b200_paged_attn_prefill: (tile_seq, tile_q0) int32 lists, ordered
    heaviest first (a causal tile's work grows with its last key position), so the long tiles do
    not end up alone in the last wave."""
    tiles = []
    for s, n in enumerate(q_lens):
        c = 0 if ctx_lens is None else int(ctx_lens[s])
        for q0 in range(0, n, tile):
            tiles.append((c + q0 + tile, s, q0))
    tiles.sort(key=lambda t: -t[0])
    return [t[1] for t in tiles], [t[2] for t in tiles]


def paged_attn_prefill(q, q_stride, k_cache, v_cache, block_table, cu_q, ctx_lens, tile_seq, tile_q0,
                       out, Hq, Hkv, D, scale):
    for t, n in ((block_table, "block_table"), (cu_q, "cu_q"), (ctx_lens, "ctx_lens"),
                 (tile_seq, "tile_seq"), (tile_q0, "tile_q0")):
        _req(t, torch.int32, n)
    check(load().b200_paged_attn_prefill(_p(q), q_stride, q.shape[0], _p(k_cache), _p(v_cache), k_cache.shape[0],
                                         _p(block_table), block_table.shape[1], _p(cu_q), _p(ctx_lens), _p(tile_seq),
                                         _p(tile_q0), tile_seq.shape[0], _p(out), Hq, Hkv, D,
                                         k_cache.shape[2], scale, _stream()), "paged_attn_prefill")
    return out
