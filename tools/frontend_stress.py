#!/usr/bin/env python
"""How many streamed tokens per second can the HTTP side carry?  (no GPU)

The real serving topology -- frontend process (uvicorn + OpenAI routes + detokeniser) <-pipe-> engine-core
loop -- with a STUB engine whose "GPU step" is a sleep: every running sequence gets one token per
`--step-ms`, prompts are "prefilled" in `--prefill-ms` per 8192 tokens.  The load is the same client the
bench's `e2e` leg uses (gpustack_b200/bench_client.py, token-id prompts, /v1/completions SSE) or, with
`--harness`, the reference's benchmark_serving.py.  If the achieved rate is below the stub's ideal rate the
frontend / pipe / client is the bottleneck, not the engine.  Measurement infrastructure.

  python tools/frontend_stress.py --step-ms 8 -c 256 -n 1024
"""
import argparse
import json
import os
import socket
import subprocess
import sys
import threading
import time
from pathlib import Path
from types import SimpleNamespace

import numpy as np

ROOT = Path(__file__).resolve().parents[1]
sys.path.insert(0, str(ROOT))
from gpustack_b200.scheduler import MixedPlan, PrefillPlan, Request, Scheduler  # noqa: E402


class StubEngine:
    def __init__(self, max_seqs, step_ms, prefill_ms, vocab=128256, max_len=2048):
        self.sched = Scheduler(max_seqs * (max_len // 16 + 2) + 64, 16, max_seqs, 8192, max_len, eos_token_id=None)
        self.requests, self.counters = {}, {"generation_tokens": 0}
        self.step_s, self.prefill_s = step_ms / 1e3, prefill_ms / 1e3
        self.mcfg = SimpleNamespace(vocab_size=vocab, bos_token_id=1, eos_token_id=2)
        self.ecfg = SimpleNamespace(max_model_len=max_len)

    def add_request(self, rid, prompt_ids, params):
        r = Request(rid, list(prompt_ids), params)
        self.sched.add(r)
        self.requests[rid] = r

    def abort_request(self, rid):
        self.sched.abort(rid)

    def has_work(self):
        return self.sched.has_work()

    def step(self):
        plan = self.sched.schedule()
        if plan is None:
            return None
        s = self.sched
        if isinstance(plan, MixedPlan):
            time.sleep(self.prefill_s * plan.num_tokens / 8192)
            out = s.update_mixed(plan, [65] * sum(plan.prefill.sample), np.full(plan.batch, 66, dtype=np.int32))
        elif isinstance(plan, PrefillPlan):
            time.sleep(self.prefill_s * plan.num_tokens / 8192)
            out = s.update_prefill(plan, [65] * sum(plan.sample))
        else:
            time.sleep(self.step_s)
            out = s.update_decode(plan, np.full(plan.batch, 66, dtype=np.int32))
        self.counters["generation_tokens"] += len(out.token_ids)
        for rid, fin in zip(out.req_ids, out.finished):
            if fin is not None:
                self.requests.pop(rid, None)
        return out

    def shutdown(self):
        pass


def main():
    ap = argparse.ArgumentParser()
    ap.add_argument("--step-ms", type=float, default=8.0)
    ap.add_argument("--prefill-ms", type=float, default=93.0, help="per 8192 prompt tokens")
    ap.add_argument("-c", "--concurrency", type=int, default=256)
    ap.add_argument("-n", "--num-requests", type=int, default=1024)
    ap.add_argument("--output-len", type=int, default=128)
    ap.add_argument("--harness", action="store_true", help="drive with the reference's benchmark_serving.py instead")
    ap.add_argument("--clients", type=int, default=1, help="split the load over this many client PROCESSES")
    a = ap.parse_args()
    from gpustack_b200.frontend import spawn_frontend

    eng = StubEngine(a.concurrency, a.step_ms, a.prefill_ms)
    with socket.socket() as s:
        s.bind(("127.0.0.1", 0))
        port = s.getsockname()[1]
    front, core = spawn_frontend(eng, "127.0.0.1", port, "stub", None, "warning")
    t = threading.Thread(target=core.run, daemon=True)
    t.start()
    import urllib.request

    for _ in range(200):
        try:
            if urllib.request.urlopen(f"http://127.0.0.1:{port}/health", timeout=1).status == 200:
                break
        except Exception:  # noqa: BLE001
            time.sleep(0.1)
    env = dict(os.environ, PYTHONPATH=str(ROOT))
    if a.harness:
        cmd = [sys.executable, str(ROOT / "tools" / "run_reference_harness.py"), "-m", "stub", "-n", str(a.num_requests),
               "-c", str(a.concurrency), "--max-completion-tokens", str(a.output_len), "--ignore-eos",
               "--prompt-multiplier", "4", "--server-url", f"http://127.0.0.1:{port}", "--json"]
    else:
        cmd = [sys.executable, "-m", "gpustack_b200.bench_client", "--url", f"http://127.0.0.1:{port}", "-m", "stub",
               "-n", str(a.num_requests), "-c", str(a.concurrency), "--input-len", "512", "--output-len", str(a.output_len)]
    if a.clients > 1 and not a.harness:
        k = a.clients
        cmd[cmd.index("-n") + 1] = str(a.num_requests // k)
        cmd[cmd.index("-c") + 1] = str(a.concurrency // k)
        t0 = time.time()
        ps = [subprocess.Popen(cmd, stdout=subprocess.PIPE, stderr=subprocess.PIPE, text=True, env=env) for _ in range(k)]
        outs = [p.communicate(timeout=900)[0] for p in ps]
        rs = [json.loads(o[o.index("{"):]) for o in outs]
        wall = max(x["total_time_s"] for x in rs)
        line = json.dumps({"clients": k, "successful": sum(x["successful"] for x in rs),
                           "output_tok_s": sum(x["completion_tokens"] for x in rs) / wall, "total_time_s": wall,
                           "tpot_ms_p50": float(np.median([x["tpot_ms_p50"] for x in rs])),
                           "ttft_ms_p50": float(np.median([x["ttft_ms_p50"] for x in rs]))})
    else:
        r = subprocess.run(cmd, capture_output=True, text=True, env=env, timeout=900)
        line = r.stdout[r.stdout.index("{"):] if "{" in r.stdout else r.stdout + r.stderr[-500:]
    waves = -(-a.num_requests // a.concurrency)
    ideal_s = waves * (a.prefill_ms / 1e3 * a.concurrency * 512 / 8192 + (a.output_len - 1) * a.step_ms / 1e3)
    print(json.dumps({"ideal_output_tok_s_of_the_stub": round(a.num_requests * a.output_len / ideal_s, 1),
                      "client": "benchmark_serving.py" if a.harness else "bench_client", "result": json.loads(line)
                      if line.strip().startswith("{") else line}))
    front.terminate()


if __name__ == "__main__":
    main()
