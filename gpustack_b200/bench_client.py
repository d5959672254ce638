"""Streaming load generator with the metric definitions of the reference's own harness
(benchmarks/benchmark_serving.py): concurrency-bounded streaming requests, TTFT = wall time from
request send to the FIRST SSE chunk (:74-81,120), usage read from the chunk that carries it,
output tok/s = sum(completion_tokens) / total wall time and percentiles via numpy (:321-425).

Differences, on purpose: token-id prompts through /v1/completions (token-exact lengths, seeded --
the reference script's prompts are unseeded and tokenizer-dependent) and raw aiohttp
instead of the openai SDK (httpx_aiohttp is not installed offline).
Run: python -m gpustack_b200.bench_client --url http://127.0.0.1:8000 --model m -n 256 -c 256 \
         --input-len 512 --output-len 128 --seed 0
"""
from __future__ import annotations

import argparse
import asyncio
import json
import random
import sys
import time

import aiohttp
import numpy as np


async def one_request(session, url, model, prompt_ids, max_tokens, chat: bool, results, sem, delay: float = 0.0):
    if delay > 0:
        await asyncio.sleep(delay)
    async with sem:
        if chat:
            payload = {"model": model, "messages": [{"role": "user", "content": prompt_ids}],
                       "max_completion_tokens": max_tokens, "stream": True,
                       "stream_options": {"include_usage": True}, "ignore_eos": True}
            path = "/v1/chat/completions"
        else:
            payload = {"model": model, "prompt": prompt_ids, "max_tokens": max_tokens, "stream": True,
                       "stream_options": {"include_usage": True}, "ignore_eos": True, "temperature": 0}
            path = "/v1/completions"
        t0 = time.time()
        first, usage, last, n_chunks = None, None, None, 0
        try:
            async with session.post(url + path, json=payload) as resp:
                if resp.status != 200:
                    results.append(None)
                    return
                async for raw in resp.content:
                    if not raw.startswith(b"data:"):
                        continue
                    now = time.time()
                    if first is None:
                        first = now
                    last = now
                    n_chunks += 1
                    data = raw[5:].strip()
                    if data == b"[DONE]":
                        break
                    if b'"usage"' in data:
                        u = json.loads(data).get("usage")
                        if u:
                            usage = u
        except Exception as e:  # noqa: BLE001
            print(f"request failed: {e!r}", file=sys.stderr)
            results.append(None)
            return
        t1 = time.time()
        results.append((usage, t1 - t0, (first - t0) * 1000.0 if first else None))
        if usage and first and usage.get("completion_tokens", 0) > 1:
            results[-1] = results[-1] + ((t1 - first) * 1000.0 / (usage["completion_tokens"] - 1),)


async def run(url, model, n, concurrency, input_len, output_len, seed, vocab, chat=False, text_prompts=None,
              rate: float = 0.0):
    """rate > 0: constant-rate arrivals (request i starts at i / rate seconds, the guidellm
    `constant` profile GPUStack's in-product benchmark uses, gpustack/worker/benchmark/runner.py:208-276);
    rate <= 0: all requests offered at once, bounded by `concurrency`."""
    rng = random.Random(seed)
    sem = asyncio.Semaphore(concurrency)
    results = []
    conn = aiohttp.TCPConnector(limit=0, force_close=False)
    timeout = aiohttp.ClientTimeout(total=1800)
    async with aiohttp.ClientSession(connector=conn, timeout=timeout) as session:
        prompts = []
        for i in range(n):
            if text_prompts is not None:
                prompts.append(text_prompts[i % len(text_prompts)])
            else:
                lo = 256 if vocab > 1024 else 3
                prompts.append([rng.randrange(lo, vocab - lo) for _ in range(input_len)])
        t0 = time.time()
        await asyncio.gather(*(one_request(session, url, model, p, output_len, chat, results, sem,
                                           i / rate if rate > 0 else 0.0) for i, p in enumerate(prompts)))
        total = time.time() - t0
    ok = [r for r in results if r is not None and r[0] is not None]
    comp = sum(r[0]["completion_tokens"] for r in ok)
    prm = sum(r[0]["prompt_tokens"] for r in ok)
    ttft = [r[2] for r in ok if r[2] is not None]
    lat = [r[1] for r in ok]
    tpot = [r[3] for r in ok if len(r) > 3]
    pct = lambda a, q: float(np.percentile(a, q)) if a else None
    return {
        "requests": n, "successful": len(ok), "concurrency": concurrency, "total_time_s": total,
        "output_tok_s": comp / total if total > 0 else 0.0, "total_tok_s": (comp + prm) / total if total > 0 else 0.0,
        "completion_tokens": comp, "prompt_tokens": prm, "mean_prompt_tokens": prm / max(1, len(ok)),
        "ttft_ms_p50": pct(ttft, 50), "ttft_ms_p95": pct(ttft, 95), "ttft_ms_p99": pct(ttft, 99),
        "latency_s_p50": pct(lat, 50), "latency_s_p99": pct(lat, 99),
        "tpot_ms_p50": pct(tpot, 50), "tpot_ms_p99": pct(tpot, 99), "request_rate": rate,
    }


def main():
    ap = argparse.ArgumentParser()
    ap.add_argument("--url", required=True)
    ap.add_argument("--model", "-m", required=True)
    ap.add_argument("-n", "--num-requests", type=int, default=256)
    ap.add_argument("-c", "--concurrency", type=int, default=256)
    ap.add_argument("--input-len", type=int, default=512)
    ap.add_argument("--output-len", type=int, default=128)
    ap.add_argument("--seed", type=int, default=0)
    ap.add_argument("--vocab", type=int, default=128256)
    ap.add_argument("--repeat", type=int, default=1, help="run the workload this many times; one JSON line each")
    ap.add_argument("--rate", type=float, default=0.0, help="constant request rate (req/s); <= 0 = all at once")
    a = ap.parse_args()
    for k in range(a.repeat):
        r = asyncio.run(run(a.url, a.model, a.num_requests, a.concurrency, a.input_len, a.output_len,
                            a.seed + k, a.vocab, rate=a.rate))
        print(json.dumps(r), flush=True)


if __name__ == "__main__":
    main()
