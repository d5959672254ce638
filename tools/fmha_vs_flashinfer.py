#!/usr/bin/env python
"""Head-to-head of the prefill attention kernel (csrc/attention_prefill_tc.cu) against the kernel
the reference's vLLM backend uses on sm_100: FlashInfer's TRT-LLM-gen context FMHA
(`flashinfer.prefill.trtllm_batch_context_with_kv_cache`; vLLM logs "TRTLLM-gen (FlashInfer)" as its
prefill attention on B200, profiles/r2_bench_tp8.json gpu_baseline).  Same paged K/V cache
(`[pages, Hkv, 16, D]` = FlashInfer's HND tuple layout), same block tables, same Q; both outputs are
checked against each other, both are timed with CUDA events after warm-up, L2 flushed between
iterations.  Measurement infrastructure: FlashInfer is LIBRARY code and is never on the product path.

  python tools/fmha_vs_flashinfer.py            # one JSON line per shape
"""
import json
import sys
from pathlib import Path

import torch

ROOT = Path(__file__).resolve().parents[1]
sys.path.insert(0, str(ROOT))
from gpustack_b200 import ops  # noqa: E402

BF = torch.bfloat16
HQ, HKV, D, PAGE = 32, 8, 128, 16


def build(batch, seqlen, dev):
    g = torch.Generator(device=dev).manual_seed(batch * 7919 + seqlen)
    pages_per = (seqlen + PAGE - 1) // PAGE
    total = batch * pages_per + 3
    kc = torch.randn(total, HKV, PAGE, D, generator=g, device=dev, dtype=torch.float32).to(BF)
    vc = torch.randn(total, HKV, PAGE, D, generator=g, device=dev, dtype=torch.float32).to(BF)
    perm = torch.randperm(total, generator=g, device=dev)[: batch * pages_per].to(torch.int32)
    bt = perm.reshape(batch, pages_per).contiguous()
    stride = (HQ + 2 * HKV) * D
    qkv = torch.randn(batch * seqlen, stride, generator=g, device=dev, dtype=torch.float32).to(BF)
    return kc, vc, bt, qkv, stride


def time_it(fn, flush, iters=20, warmup=5):
    for _ in range(warmup):
        fn()
    torch.cuda.synchronize()
    tot = 0.0
    for _ in range(iters):
        flush.add_(1)                                   # > L2: the next launch starts cold like in a real step
        a, b = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
        a.record()
        fn()
        b.record()
        torch.cuda.synchronize()
        tot += a.elapsed_time(b)
    return tot / iters * 1e3


def main():
    dev = torch.device("cuda:0")
    torch.cuda.set_device(dev)
    flush = torch.zeros(256 << 20, dtype=torch.uint8, device=dev)
    try:
        from flashinfer.prefill import trtllm_batch_context_with_kv_cache as fi_ctx
    except Exception as e:  # noqa: BLE001
        fi_ctx, fi_err = None, repr(e)
    ws = torch.zeros(256 << 20, dtype=torch.uint8, device=dev)
    scale = D ** -0.5
    for batch, seqlen in ((16, 512), (4, 2048), (1, 8192)):
        kc, vc, bt, qkv, stride = build(batch, seqlen, dev)
        T = batch * seqlen
        cu = torch.arange(0, T + 1, seqlen, dtype=torch.int32, device=dev)
        ctx = torch.zeros(batch, dtype=torch.int32, device=dev)
        tile = ops.attn_prefill_tile(HQ, HKV)
        ts, tq = ops.build_prefill_tiles([seqlen] * batch, tile)
        ts = torch.tensor(ts, dtype=torch.int32, device=dev)
        tq = torch.tensor(tq, dtype=torch.int32, device=dev)
        out = torch.zeros(T, HQ * D, dtype=BF, device=dev)

        def ours():
            ops.paged_attn_prefill(qkv, stride, kc, vc, bt, cu, ctx, ts, tq, out, HQ, HKV, D, scale)

        flops = 4.0 * HQ * D * batch * (seqlen * (seqlen + 1) / 2)          # causal, useful
        us = time_it(ours, flush)
        # spot check of our output against plain fp32 attention: last sequence, first kv-head group
        sq = batch - 1
        pages = bt[sq].long()
        kk = kc[pages, 0].reshape(-1, D)[:seqlen].float()
        vv = vc[pages, 0].reshape(-1, D)[:seqlen].float()
        qq = qkv[sq * seqlen:(sq + 1) * seqlen, :D].float()                 # head 0
        att = (qq @ kk.T) * scale
        att = att.masked_fill(torch.triu(torch.ones(seqlen, seqlen, dtype=torch.bool, device=dev), 1), float("-inf"))
        ref = torch.softmax(att, -1) @ vv
        ref_err = (out[sq * seqlen:(sq + 1) * seqlen, :D].float() - ref).abs().max().item()
        bad = torch.isnan(out.float()).reshape(T, HQ, D).any(-1).nonzero()[:12].tolist()      # (token, head) pairs
        line = {"shape": f"{batch} x {seqlen} causal, Hq {HQ} / Hkv {HKV} x {D}, page {PAGE}", "flops": flops,
                "nan_rows_ours_after_timing": bad, "page0_owner": (bt == 0).nonzero()[:2].tolist(),
                "ours_us": round(us, 2), "ours_tflops": round(flops / us / 1e6, 1),
                "ours_max_abs_err_vs_fp32_reference": ref_err}
        if fi_ctx is None:
            line["flashinfer"] = {"unavailable": fi_err}
        else:
            q = qkv[:, : HQ * D].reshape(T, HQ, D).contiguous()
            seq_lens = torch.full((batch,), seqlen, dtype=torch.int32, device=dev)
            fo = torch.zeros(T, HQ, D, dtype=BF, device=dev)

            def theirs():
                fi_ctx(q, (kc, vc), ws, bt, seq_lens, seqlen, seqlen, scale, 1.0, batch, cu, cu, out=fo, kv_layout="HND")

            try:
                fus = time_it(theirs, flush)
                diff = (fo.reshape(T, -1).float() - out.float()).abs().max().item()
                line.update({"nan_ours": int(torch.isnan(out.float()).sum()),
                             "nan_flashinfer": int(torch.isnan(fo.float()).sum()),
                             "flashinfer_trtllm_gen_us": round(fus, 2), "flashinfer_tflops": round(flops / fus / 1e6, 1),
                             "ours_over_flashinfer": round(fus / us, 3), "max_abs_diff_between_the_two": diff,
                             "note": "FlashInfer is given a contiguous Q copy (made outside its timed region); ours reads Q "
                                     "strided out of the fused QKV buffer as the engine does"})
            except Exception as e:  # noqa: BLE001
                line["flashinfer"] = {"failed": repr(e)[:300]}
        print(json.dumps(line), flush=True)


if __name__ == "__main__":
    main()
