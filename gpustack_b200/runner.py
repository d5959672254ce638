"""GPU model runner: executes prefill and decode steps of a Llama-class decoder with libb200.so.

Restates the reference engine's LlamaModel.forward / LlamaDecoderLayer.forward
($VLLM/model_executor/models/llama.py:316-333, 395-434) as a fixed sequence of C-ABI kernel
launches on torch's current stream:

    embed -> 32 x [ (add+)RMSNorm -> QKV GEMM -> RoPE+KV-write -> paged attention -> O GEMM ->
                    add+RMSNorm -> gate_up GEMM -> SiLU*mul -> down GEMM ] -> add+RMSNorm ->
    last-token gather -> lm_head GEMM (fp32 logits) -> argmax

Decode steps are captured once per batch bucket in CUDA graphs (the reference engine does the same
through torch.compile/cudagraph, $VLLM/v1/worker/gpu_model_runner.py); step inputs travel in ONE
pinned int32 staging buffer copied host->device with a single cudaMemcpyAsync.
"""
from __future__ import annotations

import math
import os
import time
from typing import Dict, List, Optional, Sequence

import numpy as np
import torch

from . import ops
from .config import EngineConfig, ModelConfig
from .lib import B200Error
from .weights import ModelWeights

BF16 = torch.bfloat16
_SPLITK_MAX_M = 256


def _align(n: int, a: int = 4) -> int:
    return (n + a - 1) // a * a


class ModelRunner:
    def __init__(self, mcfg: ModelConfig, ecfg: EngineConfig, weights: ModelWeights, device: str = "cuda",
                 tp_group=None, tp_ctx=None):
        if not torch.cuda.is_available():
            raise B200Error("ModelRunner needs a CUDA device: this engine has no CPU path")
        self.mcfg, self.ecfg, self.w = mcfg, ecfg, weights
        self.device = torch.device(device)
        self.tp = ecfg.tensor_parallel_size
        self.tp_group = tp_group
        self.Hq = mcfg.num_attention_heads // self.tp
        self.Hkv = mcfg.num_key_value_heads // self.tp
        self.D = mcfg.head_dim
        self.H = mcfg.hidden_size
        self.I = mcfg.intermediate_size // self.tp
        self.qkv_width = (self.Hq + 2 * self.Hkv) * self.D
        self.scale = self.D ** -0.5
        self.max_model_len = ecfg.max_model_len or mcfg.max_position_embeddings
        self.block_size = ecfg.block_size
        self.max_blocks = (self.max_model_len + self.block_size - 1) // self.block_size
        self.max_seqs = ecfg.max_num_seqs
        self.max_tokens = max(ecfg.max_num_batched_tokens, self.max_seqs)
        # rows one step may sample: one per sequence, (1 + k) per sequence when speculative decoding
        # verifies k drafts per step (spec_decode.py)
        self.max_sample = min(self.max_tokens, self.max_seqs * (1 + ecfg.num_speculative_tokens))
        self.prefill_tile = ops.attn_prefill_tile(self.Hq, self.Hkv)
        self.v_local = weights.lm_head.shape[0]
        self.fp8 = weights.layers[0].qkv.dtype == ops.FP8
        self._alloc_workspace()
        self.k_cache: List[torch.Tensor] = []
        self.v_cache: List[torch.Tensor] = []
        self.num_blocks = 0
        self.graphs: Dict[int, torch.cuda.CUDAGraph] = {}
        self.graph_pool = None
        self._gen = torch.Generator(device=self.device)
        self._gen.manual_seed(ecfg.seed)
        self.oneshot = None
        self.twoshot = None
        self._ar_toggle = 0
        self._tp_mode = "nccl"
        # one-shot moves (tp-1) x the message over NVLink: right for decode-sized messages only
        self._oneshot_max_bytes = int(os.environ.get("B200_TP_ONESHOT_MAX_BYTES", str(8 << 20)))
        ar = os.environ.get("B200_TP_ALLREDUCE", "fused")     # fused | oneshot (no two-shot) | nccl
        if self.tp > 1 and tp_ctx is not None and ar in ("fused", "oneshot"):
            from .tp_comm import OneShotAllReduce, TwoShotAllReduce

            self.oneshot = OneShotAllReduce(tp_ctx, min(512, self.max_tokens), self.H, self.device)
            if ar == "fused":
                self.twoshot = TwoShotAllReduce(tp_ctx, self.max_tokens, self.H, self.device)
                self.xn = self.twoshot.xn          # the peers push normalised rows straight into it
        self.peer_argmax = None
        if self.tp > 1 and tp_ctx is not None and os.environ.get("B200_TP_SAMPLER", "peer") == "peer":
            from .tp_comm import PeerArgmax

            per = (mcfg.vocab_size + self.tp - 1) // self.tp
            self.peer_argmax = PeerArgmax(tp_ctx, self.max_sample, per, self.device)
        self.stats = {"prefill_steps": 0, "decode_steps": 0, "graph_replays": 0, "kernel_launches": 0,
                      "h2d_bytes": 0, "d2h_bytes": 0, "decode_kv_tokens": 0, "prefill_tokens": 0}
        self.launches_per_decode = 0

    # ------------------------------------------------------------------------------ memory
    def _alloc_workspace(self):
        dev, T, H = self.device, self.max_tokens, self.H
        e = lambda *s, dt=BF16: torch.empty(*s, dtype=dt, device=dev)
        self.resid = e(T, H)
        self.hbuf = e(T, H)
        self.xn = e(T, H)
        self.qkv = e(T, self.qkv_width)
        self.attn_out = e(T, self.Hq * self.D)
        self.act = e(T, self.I)
        self.moe_ws = None
        if self.mcfg.num_local_experts:
            self.moe_ws = ops.MoEWorkspace(T, self.mcfg.num_experts_per_tok, self.mcfg.num_local_experts, H, self.I, dev)
        if self.fp8:
            kmax = max(H, self.Hq * self.D, self.I)
            self.xq = e(T * kmax, dt=ops.FP8)          # flat; viewed [T, K] per projection
            self.xs = e(T, dt=torch.float32)
        self.x_last = e(self.max_sample, H)
        self.logits = e(self.max_sample, self.v_local, dt=torch.float32)
        self.out_ids = torch.zeros(self.max_sample, dtype=torch.int32, device=dev)
        self.out_vals = torch.zeros(self.max_sample, dtype=torch.float32, device=dev)
        has_bias = self.w.layers[0].qkv_bias is not None
        n_part = max(8 * min(_SPLITK_MAX_M, T) * max(H, self.qkv_width), T * self.qkv_width if has_bias else 0)
        self.partials = e(n_part, dt=torch.float32)   # flat: [splits, T, W]
        self.cos_sin = self._cos_sin_cache().to(dev)
        # one int32 staging area for all step inputs (host pinned + device mirror)
        n_tiles_max = T // self.prefill_tile + self.max_seqs + 1
        self.stage_words = 3 * _align(T) + 6 * _align(self.max_seqs + 1) + 3 * _align(self.max_sample + 1) \
            + 2 * _align(n_tiles_max) + self.max_seqs * self.max_blocks + 64
        self.stage_host = torch.zeros(self.stage_words, dtype=torch.int32, pin_memory=True)
        self.stage_np = self.stage_host.numpy()
        self.stage_dev = torch.zeros(self.stage_words, dtype=torch.int32, device=dev)
        self.out_host = torch.zeros(self.max_sample, dtype=torch.int32, pin_memory=True)
        # fixed decode layout inside the staging area
        S = self.max_seqs
        self.d_ids, self.d_pos, self.d_slots, self.d_lens = 0, S, 2 * S, 3 * S
        self.d_temp, self.d_seed = 4 * S, 5 * S          # in-kernel sampler: fp32 temperatures (bit patterns), uint32 seeds
        self.d_bt = 6 * S
        self._attn_ws: Optional[torch.Tensor] = None

    def _cos_sin_cache(self) -> torch.Tensor:
        # $VLLM/model_executor/layers/rotary_embedding/base.py:71-93; kept in the model dtype
        inv_freq = self.mcfg.rope_inv_freq()    # plain, llama3- or linearly-scaled frequencies
        t = torch.arange(self.max_model_len, dtype=torch.float)
        freqs = torch.einsum("i,j -> ij", t, inv_freq)
        return torch.cat((freqs.cos(), freqs.sin()), dim=-1).to(BF16)

    def _decode_attn_plan(self, B: int):
        """(KV splits, workspace) of the split-KV decode attention for a batch of B sequences.  ONE
        workspace, sized for the worst batch, serves every batch size (eager, mixed and graph steps
        never overlap on the stream)."""
        splits = ops.attn_decode_splits(B, self.Hkv, self.max_model_len)
        if self._attn_ws is None:
            from .lib import load

            L = load()
            need = max(L.b200_attn_decode_workspace_bytes(b, self.Hq, self.D,
                                                          ops.attn_decode_splits(b, self.Hkv, self.max_model_len))
                       for b in range(1, self.max_seqs + 1))
            # zeroed once: it starts with the self-resetting arrival counters of the in-kernel split merge
            self._attn_ws = torch.zeros((need + 3) // 4, dtype=torch.float32, device=self.device)
        return splits, self._attn_ws

    def kv_block_bytes(self) -> int:
        return 2 * self.mcfg.num_hidden_layers * self.Hkv * self.block_size * self.D * 2

    def profile_num_blocks(self) -> int:
        """KV blocks that fit in gpu_memory_utilization of the device after weights + workspace."""
        if self.ecfg.num_gpu_blocks:
            return int(self.ecfg.num_gpu_blocks)
        torch.cuda.synchronize()
        free, total = torch.cuda.mem_get_info()
        budget = free - int(total * (1.0 - self.ecfg.gpu_memory_utilization)) - (1 << 30)  # graphs etc.
        n = budget // self.kv_block_bytes()
        need = self.max_blocks + 1
        if n < need:
            raise B200Error(f"not enough GPU memory for the KV cache: {n} blocks available, "
                            f"{need} needed for one sequence of max_model_len={self.max_model_len}")
        return int(n)

    def init_kv_cache(self, num_blocks: int) -> None:
        self.num_blocks = num_blocks
        shape = (num_blocks, self.Hkv, self.block_size, self.D)
        self.k_cache = [torch.zeros(shape, dtype=BF16, device=self.device) for _ in range(self.mcfg.num_hidden_layers)]
        self.v_cache = [torch.zeros(shape, dtype=BF16, device=self.device) for _ in range(self.mcfg.num_hidden_layers)]

    # ------------------------------------------------------------------------------ extended KV cache (host tier)
    def init_host_kv(self, slots: int) -> None:
        """Pinned host memory for `slots` KV blocks (all layers, K and V of this rank's kv-heads) plus a
        device staging area; see kv_cache.BlockPool (host tier) for who decides what moves."""
        self.host_slots = int(slots)
        if self.host_slots <= 0:
            return
        L = self.mcfg.num_hidden_layers
        shape = (L, 2, self.Hkv, self.block_size, self.D)
        self._swap_batch = 32
        self.host_kv = torch.empty((self.host_slots,) + shape, dtype=BF16, pin_memory=True)
        self.swap_stage = torch.empty((self._swap_batch,) + shape, dtype=BF16, device=self.device)

    def kv_swap(self, ops_: Sequence[Sequence[int]]) -> None:
        """Execute the pool's swap operations IN ORDER on the current stream (before the step's kernels):
        (0, block, slot) copies a device block to a host slot, (1, block, slot) the other way.  Runs of
        the same kind are batched: one gather / scatter per layer through the staging area, one pinned
        transfer per block."""
        i, n_ops = 0, len(ops_)
        L = self.mcfg.num_hidden_layers
        while i < n_ops:
            kind = int(ops_[i][0])
            j = i
            while j < n_ops and int(ops_[j][0]) == kind and j - i < self._swap_batch:
                j += 1
            blocks = torch.tensor([int(o[1]) for o in ops_[i:j]], dtype=torch.long, device=self.device)
            slots = [int(o[2]) for o in ops_[i:j]]
            n = j - i
            st = self.swap_stage[:n]
            if kind == 0:
                for l in range(L):
                    st[:, l, 0].copy_(self.k_cache[l].index_select(0, blocks))
                    st[:, l, 1].copy_(self.v_cache[l].index_select(0, blocks))
                for q, slot in enumerate(slots):
                    self.host_kv[slot].copy_(st[q], non_blocking=True)
                self.stats["kv_swap_out_blocks"] = self.stats.get("kv_swap_out_blocks", 0) + n
            else:
                for q, slot in enumerate(slots):
                    st[q].copy_(self.host_kv[slot], non_blocking=True)
                for l in range(L):
                    self.k_cache[l].index_copy_(0, blocks, st[:, l, 0])
                    self.v_cache[l].index_copy_(0, blocks, st[:, l, 1])
                self.stats["kv_swap_in_blocks"] = self.stats.get("kv_swap_in_blocks", 0) + n
            i = j

    # ------------------------------------------------------------------------------ layers
    def _quant(self, x: torch.Tensor, T: int):
        """bf16 activations -> (e4m3 [T,K], per-token scales [T]) in the shared fp8 staging buffer."""
        K = x.shape[1]
        self.stats["kernel_launches"] += 1
        return ops.quant_fp8_per_token(x, out=self.xq[:T * K].view(T, K), scales=self.xs[:T])

    def _row_parallel(self, x: torch.Tensor, w: torch.Tensor, T: int, norm_w: torch.Tensor,
                      w_scale: Optional[torch.Tensor] = None):
        """O / down projection followed by (all-reduce +) residual add + RMSNorm -> self.xn."""
        eps = self.mcfg.rms_norm_eps
        if w_scale is not None:
            # W8A8: quantise the activations per token, dequantise in the GEMM epilogue
            xq, xs = self._quant(x, T)

            def mm(out):
                return ops.gemm_fp8(xq, xs, w, w_scale, out=out)
            splitk = False
        else:
            def mm(out):
                return ops.gemm(x, w, out=out)
            splitk = True
        if self.tp > 1:
            mm(self._tp_partial_out(T))
            self._tp_reduce_norm(T, norm_w)
            self.stats["kernel_launches"] += 1
            return
        splits = ops.gemm_suggest_splits(T, w.shape[0], w.shape[1]) if (splitk and T <= _SPLITK_MAX_M) else 1
        if splits > 1:
            parts = self.partials[:splits * T * self.H].view(splits, T, self.H)
            ops.gemm(x, w, out=parts, splits=splits)
            ops.fused_add_rms_norm_partials(parts, self.resid[:T], norm_w, eps, self.xn[:T], self.resid[:T])
        else:
            mm(self.hbuf[:T])
            ops.fused_add_rms_norm(self.hbuf[:T], self.resid[:T], norm_w, eps, out=self.xn[:T],
                                   residual_out=self.resid[:T])
        self.stats["kernel_launches"] += 2

    # ---- tensor parallel: where a row-parallel projection writes its partial sum, and the fused
    #      (all-reduce + residual add + RMSNorm) that consumes it.  ONE mode per step: the two-shot
    #      kernel keeps the residual stream sharded by token rows between calls.
    def _tp_select_mode(self, T: int) -> None:
        ar = self.oneshot
        if ar is not None and T <= ar.max_tokens and T * self.H * 2 * (self.tp - 1) <= self._oneshot_max_bytes:
            self._tp_mode = "oneshot"
        elif self.twoshot is not None and T <= self.twoshot.max_tokens:
            self._tp_mode = "twoshot"
        else:
            self._tp_mode = "nccl"

    def _tp_partial_out(self, T: int) -> torch.Tensor:
        if self._tp_mode == "oneshot":
            return self.oneshot.stage(self._ar_toggle, T)
        if self._tp_mode == "twoshot":
            return self.twoshot.stage(T)
        return self.hbuf[:T]

    def _tp_reduce_norm(self, T: int, norm_w: torch.Tensor, first: bool = False) -> None:
        """xn[:T] = rmsnorm(sum over ranks of the partial + resid) * norm_w; resid updated.
        first=True: the partial is the vocab-parallel embedding lookup and there is no residual yet
        (resid := the summed embeddings)."""
        eps = self.mcfg.rms_norm_eps
        res_in = None if first else self.resid[:T]
        if self._tp_mode == "oneshot":
            # ONE kernel pulls every rank's partial over NVLink, sums, adds the residual, normalises
            which = self._ar_toggle
            self._ar_toggle ^= 1
            self.oneshot.allreduce_add_rms_norm(which, res_in, norm_w, self.xn[:T], self.resid[:T], T, eps)
            self.stats["kernel_launches"] += 1
        elif self._tp_mode == "twoshot":
            self.twoshot.rs_norm_ag(res_in, norm_w, self.resid[:T], T, eps)
            self.stats["kernel_launches"] += 1
        elif first:
            torch.distributed.all_reduce(self.hbuf[:T], group=self.tp_group)
            self.resid[:T].copy_(self.hbuf[:T])
            ops.rms_norm(self.resid[:T], norm_w, eps, out=self.xn[:T])
            self.stats["kernel_launches"] += 3
        else:
            torch.distributed.all_reduce(self.hbuf[:T], group=self.tp_group)
            ops.fused_add_rms_norm(self.hbuf[:T], self.resid[:T], norm_w, eps, out=self.xn[:T],
                                   residual_out=self.resid[:T])
            self.stats["kernel_launches"] += 2

    def _embed(self, ids: torch.Tensor, T: int) -> None:
        """Token embeddings -> resid[:T] and the first norm -> xn[:T].  TP: vocab-parallel lookup
        (rows outside this rank's shard contribute zeros), summed by the step's first fused
        all-reduce + RMSNorm (VocabParallelEmbedding + all-reduce of the reference engine)."""
        m, w = self.mcfg, self.w
        if self.tp > 1:
            self._tp_select_mode(T)
            ops.embedding(ids, w.embed, out=self._tp_partial_out(T), vocab_start=w.vocab_start, vocab_end=w.vocab_end)
            self._tp_reduce_norm(T, w.layers[0].input_norm, first=True)
        else:
            ops.embedding(ids, w.embed, out=self.resid[:T])
            ops.rms_norm(self.resid[:T], w.layers[0].input_norm, m.rms_norm_eps, out=self.xn[:T])
        self.stats["kernel_launches"] += 2

    def _layers(self, T: int, attn_fn) -> None:
        """Runs all decoder layers after self._embed() (resid[:T] = embeddings, xn[:T] = first norm);
        leaves the final-normed hidden states in self.xn[:T]."""
        m, w = self.mcfg, self.w
        pos, slots = self._pos, self._slots
        for l, lw in enumerate(w.layers):
            if self.fp8:
                self._layer_fp8(T, l, lw, attn_fn)
                continue
            qs = ops.gemm_suggest_splits(T, lw.qkv.shape[0], lw.qkv.shape[1]) if T <= _SPLITK_MAX_M else 1
            if qs > 1 or lw.qkv_bias is not None:
                # fp32 (split-K) partials are reduced -- and the Qwen2 QKV bias added, one rounding --
                # inside the RoPE + KV-write kernel
                parts = self.partials[:qs * T * self.qkv_width].view(qs, T, self.qkv_width)
                ops.gemm(self.xn[:T], lw.qkv, out=parts if qs > 1 else parts[0], out_dtype=torch.float32, splits=qs)
                ops.rope_kv_write_partials(parts, self.qkv[:T], pos, self.cos_sin, self.k_cache[l],
                                           self.v_cache[l], slots, self.Hq, self.Hkv, self.D, bias=lw.qkv_bias)
            else:
                ops.gemm(self.xn[:T], lw.qkv, out=self.qkv[:T])
                ops.rope_kv_write(self.qkv[:T], pos, self.cos_sin, self.k_cache[l], self.v_cache[l], slots,
                                  self.Hq, self.Hkv, self.D)
            attn_fn(l)
            self._row_parallel(self.attn_out[:T], lw.o, T, lw.post_norm)
            nxt = w.layers[l + 1].input_norm if l + 1 < len(w.layers) else w.final_norm
            if lw.router is not None:
                self._moe(T, lw, nxt)
                self.stats["kernel_launches"] += 2
                continue
            ops.gemm_silu_mul(self.xn[:T], lw.gate_up, out=self.act[:T])   # SiLU*mul in the epilogue
            self._row_parallel(self.act[:T], lw.down, T, nxt)
            self.stats["kernel_launches"] += 4

    def _moe(self, T: int, lw, norm_w: torch.Tensor) -> None:
        """Mixtral sparse MLP (MixtralMoE.forward, $VLLM/model_executor/models/mixtral.py): router
        linear -> top-k softmax routing -> grouped expert GEMMs over the permuted rows -> combine,
        then (all-reduce +) residual add + RMSNorm -> self.xn.  Experts are tensor-parallel over I
        like the dense MLP, so the combine output is this rank's partial sum."""
        eps = self.mcfg.rms_norm_eps
        ws = self.moe_ws
        ops.gemm(self.xn[:T], lw.router, out=ws.logits[:T])
        self.stats["kernel_launches"] += 1 + ops.MOE_LAUNCHES
        if self.tp > 1:
            ops.moe_experts(self.xn[:T], ws.logits[:T], lw.w13, lw.w2, self.mcfg.num_experts_per_tok, ws,
                            out=self._tp_partial_out(T))
            self._tp_reduce_norm(T, norm_w)
            return
        ops.moe_experts(self.xn[:T], ws.logits[:T], lw.w13, lw.w2, self.mcfg.num_experts_per_tok, ws, out=self.hbuf[:T])
        ops.fused_add_rms_norm(self.hbuf[:T], self.resid[:T], norm_w, eps, out=self.xn[:T], residual_out=self.resid[:T])
        self.stats["kernel_launches"] += 1

    def _layer_fp8(self, T: int, l: int, lw, attn_fn) -> None:
        """One decoder layer with W8A8 linears (Fp8LinearMethod.apply,
        $VLLM/model_executor/layers/quantization/fp8.py:437-470): every projection input is
        quantised per token, the e4m3 GEMM dequantises in its epilogue; norms, RoPE, attention and
        the residual stream stay bf16 exactly as in the bf16 path."""
        w = self.w
        pos, slots = self._pos, self._slots
        xq, xs = self._quant(self.xn[:T], T)
        if lw.qkv_bias is not None:
            # fp32 result + Qwen2 bias folded (one rounding) into the RoPE + KV-write kernel
            parts = self.partials[:T * self.qkv_width].view(1, T, self.qkv_width)
            ops.gemm_fp8(xq, xs, lw.qkv, lw.qkv_scale, out=parts[0], out_dtype=torch.float32)
            ops.rope_kv_write_partials(parts, self.qkv[:T], pos, self.cos_sin, self.k_cache[l], self.v_cache[l],
                                       slots, self.Hq, self.Hkv, self.D, bias=lw.qkv_bias)
        else:
            ops.gemm_fp8(xq, xs, lw.qkv, lw.qkv_scale, out=self.qkv[:T])
            ops.rope_kv_write(self.qkv[:T], pos, self.cos_sin, self.k_cache[l], self.v_cache[l], slots,
                              self.Hq, self.Hkv, self.D)
        attn_fn(l)
        self._row_parallel(self.attn_out[:T], lw.o, T, lw.post_norm, w_scale=lw.o_scale)
        xq, xs = self._quant(self.xn[:T], T)
        ops.gemm_fp8(xq, xs, lw.gate_up, lw.gate_up_scale, out=self.act[:T], silu_mul=True)
        nxt = w.layers[l + 1].input_norm if l + 1 < len(w.layers) else w.final_norm
        self._row_parallel(self.act[:T], lw.down, T, nxt, w_scale=lw.down_scale)
        self.stats["kernel_launches"] += 4

    def _sample(self, n: int, rows: Optional[torch.Tensor], temps: torch.Tensor, seeds: torch.Tensor) -> None:
        """Next token of n sequences into self.out_ids[:n]; rows = last-token row of each.  temps [n]
        fp32 / seeds [n] int32 drive the one-pass sampler: 0 = greedy, > 0 = Gumbel-max temperature
        sampling (exact softmax(logits / T) draws at the cost of an argmax, inside the decode graph,
        and -- the noise being keyed by the global vocabulary index -- valid across vocab shards)."""
        x = self.xn[:n] if rows is None else ops.gather_rows(self.xn, rows, out=self.x_last[:n])
        lg = self.logits[:n]
        ops.gemm(x, self.w.lm_head, out=lg, out_dtype=torch.float32)
        if self.tp == 1:
            ops.sample(lg, temps, seeds, out_idx=self.out_ids[:n])
            self.stats["kernel_launches"] += 2 + (rows is not None)
            return
        v0 = self.w.vocab_start
        if self.peer_argmax is not None:
            # local (max key, argmax) straight into peer-mapped arrays, then ONE kernel picks the global winner
            pa = self.peer_argmax
            ops.sample(lg, temps, seeds, out_idx=pa.idxs[:n], out_val=pa.vals[:n], vocab_offset=v0)
            pa.select(self.out_ids, n)
            self.stats["kernel_launches"] += 3 + (rows is not None)
            return
        ops.sample(lg, temps, seeds, out_idx=self.out_ids[:n], out_val=self.out_vals[:n], vocab_offset=v0)
        vals = torch.empty(self.tp, n, dtype=torch.float32, device=self.device)
        idxs = torch.empty(self.tp, n, dtype=torch.int32, device=self.device)
        torch.distributed.all_gather_into_tensor(vals, self.out_vals[:n].contiguous(), group=self.tp_group)
        torch.distributed.all_gather_into_tensor(idxs, (self.out_ids[:n] + self.w.vocab_start).contiguous(),
                                                 group=self.tp_group)
        best = vals.argmax(dim=0, keepdim=True)   # ties -> lowest rank = lowest vocab index
        self.out_ids[:n].copy_(idxs.gather(0, best)[0])

    # ------------------------------------------------------------------------------ decode
    def _decode_forward(self, B: int) -> None:
        sd = self.stage_dev
        ids = sd[self.d_ids:self.d_ids + B]
        self._pos = sd[self.d_pos:self.d_pos + B]
        self._slots = sd[self.d_slots:self.d_slots + B]
        lens = sd[self.d_lens:self.d_lens + B]
        bt = sd[self.d_bt:self.d_bt + B * self.max_blocks].view(B, self.max_blocks)
        self._embed(ids, B)
        splits, ws = self._decode_attn_plan(B)

        def attn(l):
            ops.paged_attn_decode(self.qkv[:B], self.qkv_width, self.k_cache[l], self.v_cache[l], bt, lens,
                                  self.attn_out[:B], self.Hq, self.Hkv, self.D, self.scale,
                                  num_splits=splits, workspace=ws)

        self._layers(B, attn)
        self._sample(B, None, sd[self.d_temp:self.d_temp + B].view(torch.float32), sd[self.d_seed:self.d_seed + B])

    def capture_graphs(self) -> None:
        if self.ecfg.enforce_eager:
            return
        self.stage_dev.zero_()
        s = torch.cuda.Stream()
        s.wait_stream(torch.cuda.current_stream())
        with torch.cuda.stream(s):
            for B in sorted(self.ecfg.graph_buckets(), reverse=True):
                k0 = self.stats["kernel_launches"]
                self._decode_forward(B)          # warm-up (allocates workspaces, sets func attrs)
                self.launches_per_decode = self.stats["kernel_launches"] - k0
                torch.cuda.synchronize()
                g = torch.cuda.CUDAGraph()
                with torch.cuda.graph(g, pool=self.graph_pool, stream=s):
                    self._decode_forward(B)
                if self.graph_pool is None:
                    self.graph_pool = g.pool()
                self.graphs[B] = g
        torch.cuda.current_stream().wait_stream(s)
        torch.cuda.synchronize()
        self._buckets = sorted(self.graphs)
        self.stats["kernel_launches"] = 0

    def decode(self, B: int, last_token: np.ndarray, num_computed: np.ndarray, block_table: np.ndarray,
               temps: Optional[np.ndarray] = None, seeds: Optional[np.ndarray] = None) -> np.ndarray:
        """One decode step for rows 0..B-1 of the scheduler's persistent batch.  temps (fp32) / seeds
        (uint32, already mixed with the position) per row drive the in-kernel sampler; None = greedy."""
        S, bs = self.max_seqs, self.block_size
        Bp = B
        if self.graphs:
            Bp = next((b for b in self._buckets if b >= B), None)
            if Bp is None:
                raise B200Error(f"decode batch {B} exceeds the largest captured graph")
        st = self.stage_np
        pos = num_computed[:B]
        st[self.d_ids:self.d_ids + B] = last_token[:B]
        st[self.d_pos:self.d_pos + B] = pos
        st[self.d_slots:self.d_slots + B] = block_table[np.arange(B), pos // bs] * bs + pos % bs
        st[self.d_lens:self.d_lens + B] = pos + 1
        if temps is None:
            st[self.d_temp:self.d_temp + Bp] = 0
        else:
            st[self.d_temp:self.d_temp + B] = np.asarray(temps[:B], dtype=np.float32).view(np.int32)
            st[self.d_seed:self.d_seed + B] = np.asarray(seeds[:B], dtype=np.uint32).view(np.int32)
            st[self.d_temp + B:self.d_temp + Bp] = 0
        if Bp > B:   # padding rows: no cache write, zero-length attention
            st[self.d_ids + B:self.d_ids + Bp] = 0
            st[self.d_pos + B:self.d_pos + Bp] = 0
            st[self.d_slots + B:self.d_slots + Bp] = -1
            st[self.d_lens + B:self.d_lens + Bp] = 0
        st[self.d_bt:self.d_bt + B * self.max_blocks] = block_table[:B].reshape(-1)
        n = self.d_bt + B * self.max_blocks
        t_gpu = time.perf_counter()
        self.stage_dev[:n].copy_(self.stage_host[:n], non_blocking=True)
        if self.graphs:
            self.graphs[Bp].replay()
            self.stats["graph_replays"] += 1
            self.stats["kernel_launches"] += self.launches_per_decode
        else:
            self._decode_forward(Bp)
        self.out_host[:B].copy_(self.out_ids[:B], non_blocking=True)
        torch.cuda.current_stream().synchronize()
        self.stats["decode_gpu_s"] = self.stats.get("decode_gpu_s", 0.0) + time.perf_counter() - t_gpu
        self.stats["decode_steps"] += 1
        self.stats["h2d_bytes"] += n * 4
        self.stats["d2h_bytes"] += B * 4
        self.stats["decode_kv_tokens"] += int(pos.sum()) + B
        return self.out_host[:B].numpy()

    # ------------------------------------------------------------------------------ prefill
    def prefill(self, token_chunks: Sequence[Sequence[int]], starts: Sequence[int], block_tables: Sequence[Sequence[int]],
                sample: Sequence[bool], decode=None, sample_temps: Optional[Sequence[float]] = None,
                sample_seeds: Optional[Sequence[int]] = None) -> np.ndarray:
        """Runs one (chunked) prefill step.  token_chunks[s] are the ids of chunk s, which occupies
        positions [starts[s], starts[s]+len).  sample[s] is the number of TRAILING rows of chunk s whose
        next token is sampled (False/0: none, True/1: the last row = a completed prompt, 1 + k: every row
        of a speculative verification chunk).  Returns the sampled tokens, chunk after chunk.
        decode = (B, last_token, num_computed, block_table) makes it a MIXED step: rows 0..B-1 of the
        scheduler's persistent batch ride along with one decode token each (appended after the
        prefill tokens: same GEMMs / norms / RoPE launches, the weights are streamed once for
        both); their B tokens follow the prefill tokens in the returned array."""
        S = len(token_chunks)
        lens = [len(c) for c in token_chunks]
        Tp = sum(lens)
        Bd = 0 if decode is None else int(decode[0])
        T = Tp + Bd
        if T > self.max_tokens or S + Bd > self.max_seqs:
            raise B200Error(f"prefill step too large: {T} tokens / {S + Bd} sequences")
        bs = self.block_size
        st = self.stage_np
        o_ids, o_pos, o_slots = 0, _align(T), 2 * _align(T)
        counts = [min(int(c), n) for c, n in zip(sample, lens)]     # bool -> 0 / 1
        NS = sum(counts) + Bd
        if NS > self.max_sample:
            raise B200Error(f"prefill step samples {NS} rows, the logits buffer holds {self.max_sample}")
        o_cu = 3 * _align(T)
        o_ctx = o_cu + _align(S + 1)
        o_rows = o_ctx + _align(S)
        tile_seq, tile_q0 = ops.build_prefill_tiles(lens, self.prefill_tile, starts)
        nt = len(tile_seq)
        o_ts = o_rows + _align(NS)
        o_tq = o_ts + _align(nt)
        o_dl = o_tq + _align(nt)              # decode rows: sequence lengths
        o_st = o_dl + _align(Bd)              # sampler inputs of the sampled rows: temperatures, seeds
        o_ss = o_st + _align(NS)
        o_bt = o_ss + _align(NS)
        end = o_bt + (S + Bd) * self.max_blocks
        if end > self.stage_words:
            raise B200Error("step inputs exceed the staging buffer")
        cu = np.zeros(S + 1, dtype=np.int32)
        np.cumsum(lens, out=cu[1:])
        bt = np.zeros((S + Bd, self.max_blocks), dtype=np.int32)
        for s in range(S):
            a, b = cu[s], cu[s + 1]
            tbl = np.asarray(block_tables[s], dtype=np.int32)
            bt[s, :len(tbl)] = tbl
            p = np.arange(starts[s], starts[s] + lens[s], dtype=np.int32)
            st[o_ids + a:o_ids + b] = token_chunks[s]
            st[o_pos + a:o_pos + b] = p
            st[o_slots + a:o_slots + b] = tbl[p // bs] * bs + p % bs
        st[o_cu:o_cu + S + 1] = cu
        st[o_ctx:o_ctx + S] = starts
        sample_rows = [r for s in range(S) for r in range(int(cu[s + 1]) - counts[s], int(cu[s + 1]))]
        if Bd:
            _, last_token, num_computed, dbt = decode
            dpos = np.asarray(num_computed[:Bd], dtype=np.int32)
            st[o_ids + Tp:o_ids + T] = last_token[:Bd]
            st[o_pos + Tp:o_pos + T] = dpos
            st[o_slots + Tp:o_slots + T] = dbt[np.arange(Bd), dpos // bs] * bs + dpos % bs
            st[o_dl:o_dl + Bd] = dpos + 1
            bt[S:] = dbt[:Bd]
            sample_rows += list(range(Tp, T))
            self.stats["decode_kv_tokens"] += int(dpos.sum()) + Bd
        ns = len(sample_rows)
        st[o_rows:o_rows + ns] = sample_rows
        st[o_st:o_st + ns] = 0
        if sample_temps is not None and ns:      # one entry per sampled row (prefill rows first, then the riders)
            st[o_st:o_st + ns] = np.asarray(sample_temps, dtype=np.float32).view(np.int32)
            st[o_ss:o_ss + ns] = np.asarray(sample_seeds, dtype=np.uint32).view(np.int32)
        st[o_ts:o_ts + nt] = tile_seq
        st[o_tq:o_tq + nt] = tile_q0
        st[o_bt:end] = bt.reshape(-1)
        self.stage_dev[:end].copy_(self.stage_host[:end], non_blocking=True)
        sd = self.stage_dev
        ids = sd[o_ids:o_ids + T]
        self._pos, self._slots = sd[o_pos:o_pos + T], sd[o_slots:o_slots + T]
        cu_d, ctx_d = sd[o_cu:o_cu + S + 1], sd[o_ctx:o_ctx + S]
        ts_d, tq_d = sd[o_ts:o_ts + nt], sd[o_tq:o_tq + nt]
        bt_all = sd[o_bt:end].view(S + Bd, self.max_blocks)
        bt_d = bt_all[:S]
        self._embed(ids, T)
        if Bd:
            dl_d, dbt_d = sd[o_dl:o_dl + Bd], bt_all[S:]
            splits, ws = self._decode_attn_plan(Bd)

        def attn(l):
            # the prefill kernel only ever touches the first Tp rows (q is read through cu_q / tiles)
            if nt:
                ops.paged_attn_prefill(self.qkv[:Tp], self.qkv_width, self.k_cache[l], self.v_cache[l], bt_d, cu_d,
                                       ctx_d, ts_d, tq_d, self.attn_out[:Tp], self.Hq, self.Hkv, self.D, self.scale)
            if Bd:   # the riding decode rows: one query token each, the split-KV decode kernel
                ops.paged_attn_decode(self.qkv[Tp:T], self.qkv_width, self.k_cache[l], self.v_cache[l], dbt_d, dl_d,
                                      self.attn_out[Tp:T], self.Hq, self.Hkv, self.D, self.scale,
                                      num_splits=splits, workspace=ws)
                self.stats["kernel_launches"] += 1

        self._layers(T, attn)
        self.stats["prefill_steps"] += 1
        self.stats["mixed_steps"] = self.stats.get("mixed_steps", 0) + (1 if Bd else 0)
        self.stats["h2d_bytes"] += end * 4
        self.stats["d2h_bytes"] += ns * 4
        self.stats["prefill_tokens"] += Tp
        if ns == 0:
            torch.cuda.current_stream().synchronize()
            return np.zeros(0, dtype=np.int32)
        self._sample(ns, sd[o_rows:o_rows + ns], sd[o_st:o_st + ns].view(torch.float32), sd[o_ss:o_ss + ns])
        self.out_host[:ns].copy_(self.out_ids[:ns], non_blocking=True)
        torch.cuda.current_stream().synchronize()
        return self.out_host[:ns].numpy().copy()

    # ------------------------------------------------------------------------------ sampling
    def gather_logits_rows(self, rows: Sequence[int]) -> torch.Tensor:
        """Full-vocabulary fp32 logits [len(rows), V] of the given rows of the last step.  TP>1: the
        vocab shards are all-gathered over NCCL (every rank must call this with the same rows; the
        reference engine gathers logits the same way, $VLLM/model_executor/layers/logits_processor.py)."""
        idx = torch.tensor(list(rows), dtype=torch.long, device=self.device)
        local = self.logits.index_select(0, idx)
        if self.tp == 1:
            return local
        V = self.mcfg.vocab_size
        per = (V + self.tp - 1) // self.tp
        if local.shape[1] < per:                     # last rank of an uneven split
            local = torch.nn.functional.pad(local, (0, per - local.shape[1]), value=float("-inf"))
        out = torch.empty(self.tp, len(rows), per, dtype=torch.float32, device=self.device)
        torch.distributed.all_gather_into_tensor(out, local.contiguous(), group=self.tp_group)
        return out.permute(1, 0, 2).reshape(len(rows), self.tp * per)[:, :V]

    def sample_rows(self, rows: Sequence[int], temps: Sequence[float], top_ps: Sequence[float],
                    generators: Sequence[Optional[torch.Generator]]) -> List[int]:
        """Temperature / nucleus sampling for the given rows of the last step's logits (the greedy
        path stays inside the captured graph; sampled rows are overridden afterwards, as
        $VLLM/v1/sample/sampler.py does with its greedy/random split).  TP>1: rank 0 samples from the
        gathered full-vocabulary logits (the other ranks only take part in the gather)."""
        lg = self.gather_logits_rows(rows) / torch.tensor(list(temps), dtype=torch.float32,
                                                          device=self.device).clamp_min(1e-5).unsqueeze(1)
        probs = torch.softmax(lg, dim=-1)
        tp = torch.tensor(list(top_ps), dtype=torch.float32, device=self.device).unsqueeze(1)
        if bool((tp < 1.0).any()):
            sp, si = torch.sort(probs, dim=-1, descending=True)
            keep = (torch.cumsum(sp, dim=-1) - sp) < tp          # always keeps the top token
            sp = torch.where(keep, sp, torch.zeros_like(sp))
            probs = torch.zeros_like(probs).scatter_(1, si, sp)
        probs = probs / probs.sum(dim=-1, keepdim=True)
        out = [0] * len(generators)
        shared = [i for i, g in enumerate(generators) if g is None]
        if shared:   # one batched draw for every row without its own seed
            draw = torch.multinomial(probs[shared], 1, generator=self._gen).flatten().tolist()
            for i, t in zip(shared, draw):
                out[i] = int(t)
        for i, g in enumerate(generators):
            if g is not None:
                out[i] = int(torch.multinomial(probs[i], 1, generator=g))
        return out

    # -------------------------------------------------------------------- test / debug hooks
    def last_logits(self, n: int) -> torch.Tensor:
        """fp32 logits of the last _sample() call (local vocab shard), for parity tests."""
        return self.logits[:n].clone()
