#!/usr/bin/env python
"""Stand-in for the container GPUStack's in-product benchmark launches, speaking ITS command line.

gpustack/worker/benchmark/runner.py:208-276 (`_build_command_args`) composes

    benchmark run --target <endpoint> --profile constant --rate <request_rate> --sample-requests 0
        --processor <model_path> --output-dir <dir> --outputs <id>.dual_json --progress-url .. --progress-auth ..
        --backend-kwargs '{"timeout": ..}' --backend openai_http_error_detail
        [--data prompt_tokens=<N>,output_tokens=<M>] [--random-seed <S>] [--max-requests <K>]

for the profiles of gpustack/assets/profiles_config/profiles_config.yaml (Throughput 1024/128 unlimited rate,
Latency 128/128 @ 1 req/s, Long Context 32000/100, Generation Heavy 1000/2000; seed 42).  The real runner image
(a guidellm build) does not exist offline; this script accepts the same argv and drives the target with
gpustack_b200/bench_client.py: token-exact synthetic prompts (the Random dataset: `prompt_tokens` ids drawn with
`--random-seed`), `ignore_eos` + `max_tokens = output_tokens`, constant request rate (`--rate`, <= 0 = unlimited),
at most 256 requests in flight.  It writes `<output-dir>/<outputs stem>.json` with the request-level aggregates
the product's benchmark table shows (requests/s, output tok/s, total tok/s, TTFT / TPOT / latency percentiles).
The ShareGPT dataset (a file inside the runner image) is not available: `--data <path>` is refused.

  python tools/benchmark_runner.py benchmark run --target http://127.0.0.1:8000 --profile constant --rate -1 \\
      --data prompt_tokens=1024,output_tokens=128 --random-seed 42 --max-requests 1000 --output-dir /tmp/b --outputs 7.dual_json
"""
import argparse
import asyncio
import json
import sys
import urllib.request
from pathlib import Path

ROOT = Path(__file__).resolve().parents[1]
sys.path.insert(0, str(ROOT))
from gpustack_b200.bench_client import run  # noqa: E402


def parse(argv):
    ap = argparse.ArgumentParser(prog="benchmark_runner")
    ap.add_argument("group", choices=["benchmark"])
    ap.add_argument("verb", choices=["run"])
    ap.add_argument("--target", required=True)
    ap.add_argument("--profile", default="constant")
    ap.add_argument("--rate", type=float, default=-1.0)
    ap.add_argument("--sample-requests", type=int, default=0)
    ap.add_argument("--processor", default=None, help="tokenizer path (unused: prompts are token ids)")
    ap.add_argument("--output-dir", default=".")
    ap.add_argument("--outputs", default="benchmark.dual_json")
    ap.add_argument("--progress-url", default=None)
    ap.add_argument("--progress-auth", default=None)
    ap.add_argument("--backend-kwargs", default="{}")
    ap.add_argument("--backend", default="openai_http_error_detail")
    ap.add_argument("--data", default="prompt_tokens=1024,output_tokens=128")
    ap.add_argument("--random-seed", type=int, default=42)
    ap.add_argument("--max-requests", type=int, default=1000)
    ap.add_argument("--max-concurrency", type=int, default=256)
    return ap.parse_args(argv)


def parse_data(spec: str):
    if "=" not in spec:
        raise SystemExit(f"dataset file {spec!r} is not available offline; only the Random dataset "
                         "(prompt_tokens=N,output_tokens=M) is supported")
    kv = dict(item.split("=", 1) for item in spec.split(","))
    return int(kv["prompt_tokens"]), int(kv["output_tokens"])


def main(argv=None):
    a = parse(argv if argv is not None else sys.argv[1:])
    if a.profile != "constant":
        raise SystemExit(f"profile {a.profile!r} is not implemented (GPUStack only ever passes 'constant')")
    n_in, n_out = parse_data(a.data)
    target = a.target.rstrip("/")
    if target.endswith("/v1"):
        target = target[:-3]
    models = json.loads(urllib.request.urlopen(target + "/v1/models", timeout=30).read())["data"]
    model = models[0]["id"]
    max_len = int(models[0].get("max_model_len") or 0)
    note = None
    if max_len and n_in + n_out > max_len:
        note = f"prompt_tokens clipped from {n_in} to fit max_model_len={max_len}"
        n_in = max_len - n_out - 8
    vocab = json.loads(urllib.request.urlopen(urllib.request.Request(
        target + "/tokenize", data=b'{"prompt": "a"}', headers={"content-type": "application/json"}), timeout=30).read())
    n = max(1, a.max_requests)
    r = asyncio.run(run(target, model, n, min(n, a.max_concurrency), n_in, n_out, a.random_seed,
                        int(vocab.get("vocab_size") or 32000), rate=a.rate))
    out = {"runner": "tools/benchmark_runner.py (stand-in for GPUStack's benchmark-runner container)",
           "args": {"target": a.target, "profile": a.profile, "rate": a.rate, "data": a.data, "random_seed": a.random_seed,
                    "max_requests": a.max_requests}, "model": model, "input_tokens": n_in, "output_tokens": n_out,
           "note": note, "metrics": r}
    d = Path(a.output_dir)
    d.mkdir(parents=True, exist_ok=True)
    path = d / (Path(a.outputs).stem + ".json")
    path.write_text(json.dumps(out, indent=1))
    print(json.dumps({"output": str(path), "successful": r["successful"], "requests": r["requests"],
                      "output_tok_s": r["output_tok_s"], "ttft_ms_p50": r["ttft_ms_p50"]}))
    return 0 if r["successful"] == r["requests"] else 1


if __name__ == "__main__":
    sys.exit(main())
