#!/usr/bin/env python
"""BASELINE.json configs[2] / configs[4] topology as a declared stand-in: R independent
`b200-engine serve` replicas (TP = --tp each, on disjoint GPUs) behind tools/rr_relay.py (the
reference's RoundRobinStrategy + worker-proxy relay semantics), loaded by the reference's own
benchmarks/benchmark_serving.py (unmodified) at a sweep of concurrencies -- the "QPS sweep".

  python tools/replica_bench.py --model llama3-8b --replicas 2 --tp 4 --concurrency 8,32,128,512

One JSON line per concurrency on stdout.  GPUStack's gateway (Higress) is not runnable offline; the
relay reproduces only instance selection and the 1 KiB chunk relay hop (see tools/rr_relay.py)."""
import argparse
import json
import os
import socket
import subprocess
import sys
import time
import urllib.request
from pathlib import Path

ROOT = Path(__file__).resolve().parents[1]


def free_port():
    with socket.socket() as s:
        s.bind(("127.0.0.1", 0))
        return s.getsockname()[1]


def wait_http(url, timeout_s):
    t0 = time.time()
    while time.time() - t0 < timeout_s:
        try:
            if urllib.request.urlopen(url, timeout=2).status == 200:
                return True
        except Exception:  # noqa: BLE001
            time.sleep(0.5)
    return False


def main():
    ap = argparse.ArgumentParser()
    ap.add_argument("--model", default="llama3-8b")
    ap.add_argument("--replicas", type=int, default=2)
    ap.add_argument("--tp", type=int, default=1)
    ap.add_argument("--concurrency", default="8,32,128,512")
    ap.add_argument("--requests-per-level", type=int, default=0, help="0 = 4 x concurrency (min 64)")
    ap.add_argument("--max-completion-tokens", type=int, default=128)
    ap.add_argument("--prompt-multiplier", type=int, default=4)
    ap.add_argument("--max-num-seqs", type=int, default=256)
    a = ap.parse_args()
    procs, ports = [], []
    env0 = {k: v for k, v in os.environ.items() if k not in ("RANK", "LOCAL_RANK", "WORLD_SIZE", "MASTER_ADDR", "MASTER_PORT")}
    try:
        for r in range(a.replicas):
            port = free_port()
            ports.append(port)
            gpus = ",".join(str(r * a.tp + i) for i in range(a.tp))
            env = dict(env0, CUDA_VISIBLE_DEVICES=gpus, PYTHONPATH=str(ROOT))
            procs.append(subprocess.Popen([sys.executable, "-m", "gpustack_b200.cli", "serve", f"random:{a.model}", "--host",
                                           "127.0.0.1", "--port", str(port), "--served-model-name", a.model,
                                           "--tensor-parallel-size", str(a.tp), "--max-model-len", "2048", "--max-num-seqs",
                                           str(a.max_num_seqs), "--no-enable-prefix-caching", "--log-level", "warning"],
                                          env=env, cwd=str(ROOT)))
        for port in ports:
            if not wait_http(f"http://127.0.0.1:{port}/health", 900):
                raise SystemExit(f"replica on port {port} did not become healthy")
        rport = free_port()
        relay = subprocess.Popen([sys.executable, str(ROOT / "tools" / "rr_relay.py"), "--port", str(rport), "--model", a.model] +
                                 [f"http://127.0.0.1:{p}" for p in ports], env=env0, cwd=str(ROOT))
        procs.append(relay)
        if not wait_http(f"http://127.0.0.1:{rport}/health", 60):
            raise SystemExit("relay did not come up")
        url = f"http://127.0.0.1:{rport}"
        harness = [sys.executable, str(ROOT / "tools" / "run_reference_harness.py"), "-m", a.model, "--max-completion-tokens",
                   str(a.max_completion_tokens), "--ignore-eos", "--prompt-multiplier", str(a.prompt_multiplier),
                   "--request-timeout", "900", "--server-url", url, "--json"]
        subprocess.run(harness + ["-n", "32", "-c", "16"], capture_output=True, text=True, timeout=600)   # warm-up
        for c in [int(x) for x in a.concurrency.split(",")]:
            n = a.requests_per_level or max(64, 4 * c)
            r = subprocess.run(harness + ["-n", str(n), "-c", str(c)], capture_output=True, text=True, timeout=1800)
            if r.returncode != 0 or "{" not in r.stdout:
                print(json.dumps({"concurrency": c, "error": (r.stderr or r.stdout)[-300:]}), flush=True)
                continue
            d = json.loads(r.stdout[r.stdout.index("{"):])
            st = json.loads(urllib.request.urlopen(url + "/relay/stats", timeout=5).read())
            ok = max(1, d["successful_requests"])
            print(json.dumps({"topology": f"{a.replicas} replicas x TP={a.tp} behind the round-robin relay (stand-in)",
                              "model": a.model, "concurrency": c, "requests": d["total_requests"],
                              "successful": d["successful_requests"], "output_tok_s": d["total_completion_tokens_per_second"],
                              "requests_per_s": d["requests_per_second"], "p50_ttft_ms": d["time_to_first_token"]["p50"],
                              "p99_ttft_ms": d["time_to_first_token"]["p99"], "p50_latency_s": d["latency"]["p50"],
                              "mean_prompt_tokens": d["total_prompt_tokens"] / ok,
                              "requests_per_instance": st["requests_per_instance"],
                              "harness": "benchmarks/benchmark_serving.py (reference, unmodified)"}), flush=True)
    finally:
        for p in procs[::-1]:
            p.terminate()
        for p in procs:
            try:
                p.wait(timeout=20)
            except Exception:  # noqa: BLE001
                p.kill()


if __name__ == "__main__":
    main()
