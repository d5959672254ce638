#!/usr/bin/env python
"""Latency effect of n-gram speculative decoding at low concurrency (where decode is weight-bandwidth bound and a
k+1-token verification forward costs about what a 1-token step costs).

Random-init weights have no language to speculate on, but greedy decoding of a random transformer falls into
repetition loops quickly, and prompts with repeated spans give the prompt-lookup index material from the first
step -- enough to exercise acceptance.  The acceptance rate is therefore a property of this synthetic setup, not of
the method; what the run measures is the COST side (a verification step vs a decode step) and that accepted tokens
turn into wall-clock savings.  One JSON line per (concurrency, mode).

  python tools/spec_decode_bench.py --model llama3-8b --concurrency 1,4 --output-len 256
"""
import argparse
import json
import sys
import time
from pathlib import Path

import torch

ROOT = Path(__file__).resolve().parents[1]
sys.path.insert(0, str(ROOT))
from gpustack_b200.cli import make_engine  # noqa: E402
from gpustack_b200.config import EngineConfig  # noqa: E402
from gpustack_b200.scheduler import SamplingParams  # noqa: E402


def main():
    ap = argparse.ArgumentParser()
    ap.add_argument("--model", default="llama3-8b")
    ap.add_argument("--concurrency", default="1,4")
    ap.add_argument("--input-len", type=int, default=256)
    ap.add_argument("--output-len", type=int, default=256)
    ap.add_argument("--k", type=int, default=4)
    a = ap.parse_args()
    g = torch.Generator().manual_seed(11)
    cs = [int(x) for x in a.concurrency.split(",")]
    prompts = []
    for i in range(max(cs)):
        span = torch.randint(300, 30000, (32,), generator=g).tolist()
        prompts.append((span * (a.input_len // 32 + 1))[: a.input_len])      # a repeated span: drafts from step one
    sp = SamplingParams(max_tokens=a.output_len, ignore_eos=True)
    cfg = {"method": "ngram", "num_speculative_tokens": a.k, "prompt_lookup_max": 4, "prompt_lookup_min": 2}
    truth = {}
    # plain decoding; the real proposer; then the two bounds of the method on this hardware: every draft wrong
    # (pure cost of k+1-token verification steps) and every draft right (drafts = the plain run's own tokens)
    for mode, spec in (("plain", None), ("ngram", cfg), ("all-drafts-wrong", cfg), ("all-drafts-right", cfg)):
        ecfg = EngineConfig(model=f"random:{a.model}", max_model_len=2048, max_num_seqs=16, max_num_batched_tokens=2048,
                            num_gpu_blocks=4096, enable_prefix_caching=False, speculative_config=spec)
        eng = make_engine(ecfg)
        eng.generate(prompts[:1], SamplingParams(max_tokens=8, ignore_eos=True))          # warm-up
        if mode == "all-drafts-wrong":
            eng.sched.proposer = lambda r, k: [7] * k
        elif mode == "all-drafts-right":
            eng.sched.proposer = lambda r, k: truth[tuple(r.prompt_ids)][len(r.output_ids):len(r.output_ids) + k]
        for c in cs:
            torch.cuda.synchronize()
            st0 = dict(eng.sched.spec_stats)
            steps0 = eng.counters["steps"]
            t0 = time.perf_counter()
            outs = eng.generate(prompts[:c], sp)
            torch.cuda.synchronize()
            dt = time.perf_counter() - t0
            st = eng.sched.spec_stats
            if mode == "plain":
                truth.update({tuple(p): o for p, o in zip(prompts[:c], outs)})
            print(json.dumps({"model": a.model, "concurrency": c, "mode": mode,
                              "input_len": a.input_len, "output_len": a.output_len, "seconds": round(dt, 4),
                              "output_tok_s": round(c * a.output_len / dt, 1), "engine_steps": eng.counters["steps"] - steps0,
                              "tokens_per_step_per_seq": round(a.output_len / max(1, eng.counters["steps"] - steps0), 3),
                              "drafted": st["drafted"] - st0["drafted"], "accepted": st["accepted"] - st0["accepted"],
                              "distinct_tokens_in_first_output": len(set(outs[0]))}), flush=True)
        eng.shutdown()
        del eng
        torch.cuda.empty_cache()


if __name__ == "__main__":
    main()
