"""GPUStack's in-product benchmark profiles (gpustack/assets/profiles_config/profiles_config.yaml:1-56:
Random dataset, fixed prompt/output token counts, seed 42, constant request rate) driven against
`b200-engine serve`, i.e. the whole boundary: CLI -> engine core process -> frontend process ->
HTTP/SSE -> client process.  One JSON line per profile.

  python tools/profile_bench.py [--scale 0.1] [--model random:llama3-8b]
"""
import argparse
import asyncio
import json
import os
import socket
import subprocess
import sys
import time
import urllib.request
from pathlib import Path

ROOT = Path(__file__).resolve().parents[1]
sys.path.insert(0, str(ROOT))
from gpustack_b200.bench_client import run  # noqa: E402

# name, input tokens, output tokens, request rate (-1 = unlimited), total requests
PROFILES = [("Throughput", 1024, 128, -1, 1000), ("Latency", 128, 128, 1, 100),
            ("Long Context", 32000, 100, 1, 100), ("Generation Heavy", 1000, 2000, 1, 200)]


def main():
    ap = argparse.ArgumentParser()
    ap.add_argument("--model", default="random:llama3-8b")
    ap.add_argument("--scale", type=float, default=1.0, help="fraction of each profile's total_requests to run")
    ap.add_argument("--max-model-len", type=int, default=8192)
    ap.add_argument("--vocab", type=int, default=128256)
    a = ap.parse_args()
    with socket.socket() as s:
        s.bind(("127.0.0.1", 0))
        port = s.getsockname()[1]
    srv = subprocess.Popen([sys.executable, "-m", "gpustack_b200.cli", "serve", a.model, "--host", "127.0.0.1",
                            "--port", str(port), "--served-model-name", "profile-model", "--max-model-len",
                            str(a.max_model_len), "--max-num-seqs", "256"], cwd=str(ROOT),
                           env=dict(os.environ, PYTHONPATH=str(ROOT)))
    url = f"http://127.0.0.1:{port}"
    try:
        for _ in range(1200):
            try:
                if urllib.request.urlopen(url + "/health", timeout=1).status == 200:
                    break
            except Exception:
                if srv.poll() is not None:
                    raise SystemExit("server exited during start-up")
                time.sleep(0.25)
        asyncio.run(run(url, "profile-model", 8, 8, 64, 8, 1, a.vocab))     # warm-up
        for name, n_in, n_out, rate, total in PROFILES:
            note = None
            if n_in + n_out > a.max_model_len:     # Llama-3-8B: 8192 positions
                note = f"prompt clipped from {n_in} to fit max_model_len={a.max_model_len}"
                n_in = a.max_model_len - n_out - 8
            n = max(4, int(total * a.scale))
            r = asyncio.run(run(url, "profile-model", n, min(n, 256), n_in, n_out, 42, a.vocab, rate=float(rate)))
            r.update({"profile": name, "input_tokens": n_in, "output_tokens": n_out, "note": note})
            print(json.dumps(r), flush=True)
    finally:
        srv.terminate()
        try:
            srv.wait(timeout=20)
        except Exception:
            srv.kill()


if __name__ == "__main__":
    main()
