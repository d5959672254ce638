"""HTTP surface conformance without a GPU: a stub engine core produces tokens; every streamed
chunk must validate as openai.types.chat.ChatCompletionChunk exactly as GPUStack's
ModelUsageMiddleware parses them (gpustack/api/middlewares.py:241-250 in the reference), usage must
arrive in the final chunk, and the plumbing endpoints must satisfy GPUStack's probes."""
import json
import time

import numpy as np
import pytest
from fastapi.testclient import TestClient
from openai.types.chat import ChatCompletion, ChatCompletionChunk
from openai.types.completion import Completion

from gpustack_b200.async_engine import AsyncEngine
from gpustack_b200.scheduler import DecodePlan, MixedPlan, PrefillPlan, Request, SamplingParams, Scheduler
from gpustack_b200.server import create_app
from gpustack_b200.tokenizer import ByteTokenizer


class StubEngine:
    """Scheduler-accurate engine core whose 'GPU' emits the bytes of a fixed sentence."""

    TEXT = "Hello from the stub engine. "

    def __init__(self):
        self.sched = Scheduler(256, 16, 8, 128, 512, eos_token_id=2)
        self.requests, self.counters = {}, {"generation_tokens": 0}
        self.tok = ByteTokenizer(600)
        self.ids = self.tok.encode(self.TEXT * 20, add_bos=False)

    def add_request(self, rid, prompt_ids, params):
        r = Request(rid, list(prompt_ids), params)
        self.sched.add(r)
        self.requests[rid] = r

    def abort_request(self, rid):
        self.sched.abort(rid)

    def has_work(self):
        return self.sched.has_work()

    def _next(self, r):
        return self.ids[len(r.output_ids) % len(self.ids)]

    def step(self):
        plan = self.sched.schedule()
        if plan is None:
            return None
        time.sleep(0.001)
        if isinstance(plan, MixedPlan):
            pre = [self._next(r) for r, s in zip(plan.prefill.reqs, plan.prefill.sample) if s]
            dec = np.array([self._next(r) for r in self.sched.rows[:plan.batch]], dtype=np.int32)
            out = self.sched.update_mixed(plan, pre, dec)
        elif isinstance(plan, PrefillPlan):
            out = self.sched.update_prefill(plan, [self._next(r) for r, s in zip(plan.reqs, plan.sample) if s])
        else:
            toks = np.array([self._next(r) for r in self.sched.rows], dtype=np.int32)
            out = self.sched.update_decode(plan, toks)
        self.counters["generation_tokens"] += len(out.token_ids)
        return out

    def shutdown(self):
        pass


@pytest.fixture(scope="module")
def client():
    eng = StubEngine()
    aeng = AsyncEngine(eng)
    app = create_app(aeng, eng.tok, "stub-model", 512)
    with TestClient(app) as c:
        yield c
    aeng._stop = True


def _sse(resp):
    events = []
    for line in resp.iter_lines():
        if line.startswith("data: "):
            events.append(line[6:])
    return events


def test_plumbing_endpoints(client):
    assert client.get("/health").status_code == 200
    assert client.get("/version").json()["version"]
    m = client.get("/v1/models").json()
    assert m["data"][0]["id"] == "stub-model" and m["data"][0]["max_model_len"] == 512
    text = client.get("/metrics").text
    for name in ("vllm:num_requests_running", "vllm:num_requests_waiting", "vllm:kv_cache_usage_perc",
                 "vllm:prompt_tokens_total", "vllm:generation_tokens_total", "vllm:request_success_total",
                 "vllm:time_to_first_token_seconds_bucket", "vllm:time_per_output_token_seconds_bucket",
                 "vllm:e2e_request_latency_seconds_count"):
        assert name in text, name


def test_chat_stream_chunks_validate_like_gpustack_parses_them(client):
    body = {"model": "stub-model", "messages": [{"role": "user", "content": "hi"}], "max_completion_tokens": 24,
            "stream": True, "stream_options": {"include_usage": True}, "ignore_eos": True}
    with client.stream("POST", "/v1/chat/completions", json=body) as resp:
        assert resp.status_code == 200 and resp.headers["content-type"].startswith("text/event-stream")
        events = _sse(resp)
    assert events[-1] == "[DONE]"
    chunks = [ChatCompletionChunk.model_validate(json.loads(e)) for e in events[:-1]]
    assert chunks[0].choices[0].delta.role == "assistant"
    text = "".join(c.choices[0].delta.content or "" for c in chunks if c.choices)
    assert text == (StubEngine.TEXT * 20)[:24]
    finals = [c for c in chunks if c.choices and c.choices[0].finish_reason]
    assert len(finals) == 1 and finals[0].choices[0].finish_reason == "length"
    usage = chunks[-1].usage
    assert chunks[-1].choices == [] and usage.completion_tokens == 24
    assert usage.total_tokens == usage.prompt_tokens + 24


def test_non_stream_chat_and_completions(client):
    r = client.post("/v1/chat/completions", json={"model": "stub-model", "max_tokens": 10, "ignore_eos": True,
                                                  "messages": [{"role": "user", "content": "x"}]})
    c = ChatCompletion.model_validate(r.json())
    assert c.choices[0].message.content == StubEngine.TEXT[:10] and c.usage.completion_tokens == 10
    r = client.post("/v1/completions", json={"model": "stub-model", "prompt": [5, 6, 7, 8], "max_tokens": 7,
                                             "ignore_eos": True})
    c = Completion.model_validate(r.json())
    assert c.usage.prompt_tokens == 4 and c.choices[0].text == StubEngine.TEXT[:7]
    assert c.choices[0].finish_reason == "length"


def test_stop_string_truncates_and_reports_stop(client):
    r = client.post("/v1/completions", json={"model": "stub-model", "prompt": "abc", "max_tokens": 60,
                                             "ignore_eos": True, "stop": ["stub"]})
    j = r.json()
    assert j["choices"][0]["text"] == "Hello from the " and j["choices"][0]["finish_reason"] == "stop"


def test_stop_string_is_counted_as_stop_not_abort(client):
    def counts():
        out = {}
        for line in client.get("/metrics").text.splitlines():
            if line.startswith("vllm:request_success_total"):
                out[line.split('finished_reason="')[1].split('"')[0]] = float(line.rsplit(" ", 1)[1])
        return out

    before = counts()
    client.post("/v1/completions", json={"model": "stub-model", "prompt": "abc", "max_tokens": 60,
                                         "ignore_eos": True, "stop": ["stub"]})
    after = counts()
    assert after["stop"] == before["stop"] + 1 and after["abort"] == before["abort"]


def test_error_envelopes(client):
    r = client.post("/v1/chat/completions", json={"model": "other", "messages": [{"role": "user", "content": "x"}]})
    assert r.status_code == 404 and r.json()["error"]["code"] == "model_not_found"
    r = client.post("/v1/completions", json={"model": "stub-model", "prompt": "x" * 600, "max_tokens": 5})
    assert r.status_code == 400 and "maximum context length" in r.json()["error"]["message"]
    r = client.post("/v1/chat/completions", json={"model": "stub-model", "messages": []})
    assert r.status_code == 400 and r.json()["error"]["type"] == "invalid_request_error"
    r = client.post("/v1/completions", content=b"{not json", headers={"content-type": "application/json"})
    assert r.status_code == 400


def test_many_concurrent_streams(client):
    import concurrent.futures as cf

    def one(i):
        with client.stream("POST", "/v1/completions", json={"model": "stub-model", "prompt": [3 + i, 9], "max_tokens": 16,
                                                            "stream": True, "ignore_eos": True,
                                                            "stream_options": {"include_usage": True}}) as resp:
            ev = _sse(resp)
        u = json.loads(ev[-2])["usage"]
        return u["completion_tokens"]

    with cf.ThreadPoolExecutor(16) as ex:
        assert list(ex.map(one, range(32))) == [16] * 32


def test_frontend_process_split_streams_over_the_pipe():
    """`b200-engine serve` topology: engine-core busy loop here (stub GPU), OpenAI frontend in a
    child process, requests over real HTTP."""
    import socket
    import threading
    import types
    import urllib.request

    from gpustack_b200.frontend import spawn_frontend

    eng = StubEngine()
    eng.mcfg = types.SimpleNamespace(vocab_size=600, bos_token_id=1, eos_token_id=2)
    eng.ecfg = types.SimpleNamespace(max_model_len=512)
    with socket.socket() as s:
        s.bind(("127.0.0.1", 0))
        port = s.getsockname()[1]
    front, core = spawn_frontend(eng, "127.0.0.1", port, "stub-model", None, "error")
    th = threading.Thread(target=core.run, daemon=True)
    th.start()
    try:
        url = f"http://127.0.0.1:{port}"
        for _ in range(300):
            try:
                if urllib.request.urlopen(url + "/health", timeout=1).status == 200:
                    break
            except Exception:
                time.sleep(0.1)
        body = json.dumps({"model": "stub-model", "prompt": [5, 6, 7], "max_tokens": 12, "stream": True,
                           "ignore_eos": True, "temperature": 0, "stream_options": {"include_usage": True}}).encode()
        req = urllib.request.Request(url + "/v1/completions", data=body, headers={"content-type": "application/json"})
        lines = [l.decode().strip() for l in urllib.request.urlopen(req, timeout=30) if l.startswith(b"data: ")]
        assert lines[-1] == "data: [DONE]"
        chunks = [json.loads(l[6:]) for l in lines[:-1]]
        assert "".join(c["choices"][0]["text"] for c in chunks if c["choices"]) == StubEngine.TEXT[:12]
        assert chunks[-1]["usage"] == {"prompt_tokens": 3, "completion_tokens": 12, "total_tokens": 15}
        metrics = urllib.request.urlopen(url + "/metrics", timeout=5).read().decode()
        assert 'vllm:request_success_total{model_name="stub-model",engine="0",finished_reason="length"} 1' in metrics
        bad = urllib.request.Request(url + "/v1/completions", data=json.dumps(
            {"model": "stub-model", "prompt": [1] * 600, "max_tokens": 5}).encode(), headers={"content-type": "application/json"})
        try:
            urllib.request.urlopen(bad, timeout=10)
            assert False, "expected HTTP 400"
        except urllib.error.HTTPError as e:
            assert e.code == 400
    finally:
        front.terminate()
        front.join(timeout=10)
        th.join(timeout=10)
    assert not th.is_alive()


def test_bench_client_against_the_frontend_process():
    """The load generator (reference metric definitions: TTFT = first SSE chunk, tok/s = completion
    tokens / wall) against the real frontend process, offered-at-once and constant-rate modes."""
    import asyncio
    import socket
    import threading
    import types
    import urllib.request

    from gpustack_b200.bench_client import run
    from gpustack_b200.frontend import spawn_frontend

    eng = StubEngine()
    eng.mcfg = types.SimpleNamespace(vocab_size=600, bos_token_id=1, eos_token_id=2)
    eng.ecfg = types.SimpleNamespace(max_model_len=512)
    with socket.socket() as s:
        s.bind(("127.0.0.1", 0))
        port = s.getsockname()[1]
    front, core = spawn_frontend(eng, "127.0.0.1", port, "stub-model", None, "error")
    th = threading.Thread(target=core.run, daemon=True)
    th.start()
    try:
        url = f"http://127.0.0.1:{port}"
        for _ in range(300):
            try:
                if urllib.request.urlopen(url + "/health", timeout=1).status == 200:
                    break
            except Exception:
                time.sleep(0.1)
        r = asyncio.run(run(url, "stub-model", 12, 6, 20, 16, 0, 600))
        assert r["successful"] == 12 and r["completion_tokens"] == 12 * 16 and r["prompt_tokens"] == 12 * 20
        assert r["ttft_ms_p50"] > 0 and r["output_tok_s"] > 0 and r["tpot_ms_p50"] is not None
        t0 = time.time()
        r = asyncio.run(run(url, "stub-model", 6, 6, 10, 8, 0, 600, rate=10.0))
        assert r["successful"] == 6 and time.time() - t0 >= 0.5      # 6 requests at 10 req/s
    finally:
        front.terminate()
        front.join(timeout=10)
        th.join(timeout=10)


def test_reference_benchmark_serving_runs_unmodified_against_the_frontend():
    """BASELINE.json's harness, benchmarks/benchmark_serving.py of the reference, run UNMODIFIED
    (tools/run_reference_harness.py: runpy + a stand-in for the missing httpx_aiohttp + a seed)
    against the real frontend process: chat completions, streaming, usage from the last chunk,
    TTFT on the first chunk -- every request must succeed and the token accounting must add up."""
    import socket
    import subprocess
    import sys
    import threading
    import types
    import urllib.request
    from pathlib import Path

    root = Path(__file__).resolve().parents[1]
    sys.path.insert(0, str(root / "tools"))
    from run_reference_harness import harness_path

    if harness_path() is None:
        pytest.skip("the reference tree is not mounted here (GPU box): no benchmark_serving.py")
    from gpustack_b200.frontend import spawn_frontend

    eng = StubEngine()
    eng.sched = Scheduler(2048, 16, 8, 512, 2048, eos_token_id=2)
    eng.mcfg = types.SimpleNamespace(vocab_size=600, bos_token_id=1, eos_token_id=2)
    eng.ecfg = types.SimpleNamespace(max_model_len=2048)
    with socket.socket() as s:
        s.bind(("127.0.0.1", 0))
        port = s.getsockname()[1]
    front, core = spawn_frontend(eng, "127.0.0.1", port, "stub-model", None, "error")
    th = threading.Thread(target=core.run, daemon=True)
    th.start()
    try:
        url = f"http://127.0.0.1:{port}"
        for _ in range(300):
            try:
                if urllib.request.urlopen(url + "/health", timeout=1).status == 200:
                    break
            except Exception:
                time.sleep(0.1)
        r = subprocess.run([sys.executable, str(root / "tools" / "run_reference_harness.py"), "-m", "stub-model", "-n", "12",
                            "-c", "4", "--max-completion-tokens", "16", "--ignore-eos", "--prompt-multiplier", "2",
                            "--server-url", url, "--json"], capture_output=True, text=True, timeout=120)
        assert r.returncode == 0, r.stdout[-1500:] + r.stderr[-1500:]
        res = json.loads(r.stdout[r.stdout.index("{"):])
        assert res["successful_requests"] == 12 and res["total_requests"] == 12
        assert res["total_completion_tokens"] == 12 * 16
        assert res["total_prompt_tokens"] > 12 * 200          # two sample prompts of > 100 bytes each
        assert res["time_to_first_token"]["p50"] > 0 and res["total_completion_tokens_per_second"] > 0
    finally:
        front.terminate()
        front.join(timeout=10)
        th.join(timeout=10)


def test_round_robin_relay_alternates_between_two_replicas_and_streams():
    """tools/rr_relay.py (stand-in for GPUStack's server-side RoundRobinStrategy + worker proxy,
    gpustack/http_proxy/strategies.py:18-35, routes/worker/proxy.py:26-98): requests alternate over
    the running instances, SSE streams pass through intact, an unknown model is a 404 envelope."""
    import socket
    import subprocess
    import sys
    import threading
    import types
    import urllib.request
    from pathlib import Path

    from gpustack_b200.frontend import spawn_frontend

    root = Path(__file__).resolve().parents[1]
    sys.path.insert(0, str(root / "tools"))
    from rr_relay import RoundRobin

    rr = RoundRobin()
    assert [rr.select("m", ["a", "b", "c"]) for _ in range(5)] == ["a", "b", "c", "a", "b"]
    assert rr.select("m", ["a", "c"]) == "a"             # the instance list changed: a new cycle

    def port():
        with socket.socket() as s:
            s.bind(("127.0.0.1", 0))
            return s.getsockname()[1]

    fronts, threads, ports = [], [], []
    relay = None
    try:
        for _ in range(2):
            eng = StubEngine()
            eng.mcfg = types.SimpleNamespace(vocab_size=600, bos_token_id=1, eos_token_id=2)
            eng.ecfg = types.SimpleNamespace(max_model_len=512)
            p = port()
            front, core = spawn_frontend(eng, "127.0.0.1", p, "stub-model", None, "error")
            th = threading.Thread(target=core.run, daemon=True)
            th.start()
            fronts.append(front); threads.append(th); ports.append(p)
        for p in ports:      # both replicas RUNNING before they are handed to the relay (GPUStack routes to RUNNING instances only)
            for _ in range(300):
                try:
                    if urllib.request.urlopen(f"http://127.0.0.1:{p}/health", timeout=1).status == 200:
                        break
                except Exception:
                    time.sleep(0.1)
        rp = port()
        relay = subprocess.Popen([sys.executable, str(root / "tools" / "rr_relay.py"), "--port", str(rp), "--model", "stub-model"] +
                                 [f"http://127.0.0.1:{p}" for p in ports])
        url = f"http://127.0.0.1:{rp}"
        for _ in range(300):
            try:
                if urllib.request.urlopen(url + "/health", timeout=1).status == 200:
                    break
            except Exception:
                time.sleep(0.1)
        for i in range(6):
            body = json.dumps({"model": "stub-model", "prompt": [5, 6, 7], "max_tokens": 10, "stream": True,
                               "ignore_eos": True, "stream_options": {"include_usage": True}}).encode()
            req = urllib.request.Request(url + "/v1/completions", data=body, headers={"content-type": "application/json"})
            lines = [l.decode().strip() for l in urllib.request.urlopen(req, timeout=30) if l.startswith(b"data: ")]
            assert lines[-1] == "data: [DONE]"
            assert json.loads(lines[-2][6:])["usage"]["completion_tokens"] == 10
        st = json.loads(urllib.request.urlopen(url + "/relay/stats", timeout=5).read())["requests_per_instance"]
        # 6 generations + the health probes: strictly alternating, so the two counts differ by at most one
        counts = sorted(st.values())
        assert counts[1] - counts[0] <= 1 and sum(counts) >= 7
        bad = urllib.request.Request(url + "/v1/completions", data=json.dumps({"model": "other", "prompt": "x"}).encode(),
                                     headers={"content-type": "application/json"})
        try:
            urllib.request.urlopen(bad, timeout=10)
            assert False, "expected HTTP 404"
        except urllib.error.HTTPError as e:
            assert e.code == 404 and json.loads(e.read())["error"]["code"] == "model_not_found"
    finally:
        if relay is not None:
            relay.terminate()
            relay.wait(timeout=10)
        for f in fronts:
            f.terminate()
            f.join(timeout=10)
        for th in threads:
            th.join(timeout=10)


def test_benchmark_runner_accepts_gpustacks_command_line(tmp_path):
    """tools/benchmark_runner.py is launched with the argv gpustack/worker/benchmark/runner.py:208-276 composes for
    its runner container (Random dataset, constant rate, seed) and writes the result file into --output-dir."""
    import socket
    import subprocess
    import sys
    import threading
    import types
    import urllib.request
    from pathlib import Path

    from gpustack_b200.frontend import spawn_frontend

    eng = StubEngine()
    eng.mcfg = types.SimpleNamespace(vocab_size=600, bos_token_id=1, eos_token_id=2)
    eng.ecfg = types.SimpleNamespace(max_model_len=512)
    with socket.socket() as s:
        s.bind(("127.0.0.1", 0))
        port = s.getsockname()[1]
    front, core = spawn_frontend(eng, "127.0.0.1", port, "stub-model", None, "error")
    th = threading.Thread(target=core.run, daemon=True)
    th.start()
    try:
        url = f"http://127.0.0.1:{port}"
        for _ in range(300):
            try:
                if urllib.request.urlopen(url + "/health", timeout=1).status == 200:
                    break
            except Exception:
                time.sleep(0.1)
        root = Path(__file__).resolve().parents[1]
        argv = ["benchmark", "run", "--target", url + "/v1", "--profile", "constant", "--rate", "20", "--sample-requests", "0",
                "--processor", "/models/none", "--output-dir", str(tmp_path), "--outputs", "17.dual_json",
                "--progress-url", "http://127.0.0.1:1/unused", "--progress-auth", "key",
                "--backend-kwargs", json.dumps({"timeout": 600, "response_handlers": {"chat_completions": "x"}}),
                "--backend", "openai_http_error_detail", "--data", "prompt_tokens=40,output_tokens=9",
                "--random-seed", "42", "--max-requests", "10"]
        r = subprocess.run([sys.executable, str(root / "tools" / "benchmark_runner.py")] + argv, capture_output=True,
                           text=True, timeout=120)
        assert r.returncode == 0, r.stdout + r.stderr
        out = json.loads((tmp_path / "17.json").read_text())
        m = out["metrics"]
        assert m["successful"] == 10 and m["completion_tokens"] == 90 and m["prompt_tokens"] == 400
        assert out["args"]["rate"] == 20.0 and out["model"] == "stub-model" and m["total_time_s"] >= 0.4   # 10 requests at 20 req/s
    finally:
        front.terminate()
        front.join(timeout=10)
        th.join(timeout=10)


def test_streams_stay_ordered_when_a_step_emits_several_tokens_per_request():
    """Speculative decoding makes one engine step emit up to k+1 tokens for the same request: the step output then
    carries the request id several times.  Text, order, finish reason and usage must come out as with one token
    per step."""
    from gpustack_b200.scheduler import SpecPlan
    from gpustack_b200.spec_decode import SpeculativeConfig

    class SpecStub(StubEngine):
        def __init__(self):
            super().__init__()
            self.sched = Scheduler(256, 16, 8, 128, 512, eos_token_id=2,
                                   speculative=SpeculativeConfig(num_speculative_tokens=3), max_sample_rows=32)
            # drafts = the sentence's true continuation: always accepted
            self.sched.proposer = lambda r, k: [self.ids[(len(r.output_ids) + j) % len(self.ids)] for j in range(k)]
            self.multi = 0

        def step(self):
            plan = self.sched.schedule()
            if isinstance(plan, SpecPlan):
                sampled = []
                for i in range(plan.batch):
                    r = self.sched.rows[i]
                    n0 = len(r.output_ids)
                    sampled += [self.ids[(n0 + j) % len(self.ids)] for j in range(1 + len(plan.drafts[i]))]
                out = self.sched.update_spec(plan, sampled)
                self.multi += int(len(out.req_ids) > len(set(out.req_ids)))
                self.counters["generation_tokens"] += len(out.token_ids)
                return out
            if plan is None:
                return None
            # not a spec plan: replay StubEngine.step on the plan that was already scheduled
            if isinstance(plan, MixedPlan):
                pre = [self._next(r) for r, s in zip(plan.prefill.reqs, plan.prefill.sample) if s]
                dec = np.array([self._next(r) for r in self.sched.rows[:plan.batch]], dtype=np.int32)
                out = self.sched.update_mixed(plan, pre, dec)
            elif isinstance(plan, PrefillPlan):
                out = self.sched.update_prefill(plan, [self._next(r) for r, s in zip(plan.reqs, plan.sample) if s])
            else:
                out = self.sched.update_decode(plan, np.array([self._next(r) for r in self.sched.rows], dtype=np.int32))
            self.counters["generation_tokens"] += len(out.token_ids)
            return out

    eng = SpecStub()
    aeng = AsyncEngine(eng)
    app = create_app(aeng, eng.tok, "stub-model", 512)
    try:
        with TestClient(app) as c:
            with c.stream("POST", "/v1/chat/completions", json={
                    "model": "stub-model", "messages": [{"role": "user", "content": "hi"}], "max_tokens": 41,
                    "stream": True, "temperature": 0, "ignore_eos": True, "stream_options": {"include_usage": True}}) as resp:
                events = _sse(resp)
            chunks = [ChatCompletionChunk.model_validate_json(e) for e in events[:-1]]
            text = "".join(ch.choices[0].delta.content or "" for ch in chunks if ch.choices)
            assert text == (StubEngine.TEXT * 3)[:41]
            assert [ch.choices[0].finish_reason for ch in chunks if ch.choices][-1] == "length"
            assert chunks[-1].usage.completion_tokens == 41
            # non-streaming, same engine: several requests at once
            r = c.post("/v1/completions", json={"model": "stub-model", "prompt": "abc", "max_tokens": 30, "temperature": 0,
                                                "ignore_eos": True})
            assert r.json()["choices"][0]["text"] == (StubEngine.TEXT * 3)[:30] and r.json()["usage"]["completion_tokens"] == 30
        assert eng.multi > 0 and eng.sched.spec_stats["accepted"] > 0
    finally:
        aeng._stop = True


def test_stream_interval_folds_outputs_but_never_the_first_or_the_last():
    eng = StubEngine()
    aeng = AsyncEngine(eng)
    app = create_app(aeng, eng.tok, "stub-model", 512, stream_interval=5)
    try:
        with TestClient(app) as c:
            with c.stream("POST", "/v1/chat/completions", json={
                    "model": "stub-model", "messages": [{"role": "user", "content": "hi"}], "max_tokens": 23,
                    "stream": True, "temperature": 0, "ignore_eos": True, "stream_options": {"include_usage": True}}) as resp:
                events = _sse(resp)
            chunks = [ChatCompletionChunk.model_validate_json(e) for e in events[:-1]]
            body = [ch for ch in chunks if ch.choices]
            assert "".join(ch.choices[0].delta.content or "" for ch in body) == StubEngine.TEXT[:23]
            assert body[0].choices[0].delta.role == "assistant" and len(body[0].choices[0].delta.content) <= 2   # first output alone
            assert body[-1].choices[0].finish_reason == "length" and all(ch.choices[0].finish_reason is None for ch in body[:-1])
            assert len(body) <= 2 + (23 + 4) // 5 and max(len(ch.choices[0].delta.content or "") for ch in body[1:]) >= 4
            assert chunks[-1].usage.completion_tokens == 23
    finally:
        aeng._stop = True
