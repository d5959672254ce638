"""Async front of the engine core: a background thread runs the busy loop
(schedule -> GPU step -> outputs), asyncio handlers submit requests and await token streams.

Counterpart of the reference engine's AsyncLLM ($VLLM/v1/engine/async_llm.py).  One engine step
produces ONE cross-thread hand-off (loop.call_soon_threadsafe) carrying every request's new token,
so per-token host work stays off the GPU critical path."""
from __future__ import annotations

import asyncio
import itertools
import queue
import threading
import time
from dataclasses import dataclass
from typing import AsyncIterator, Dict, List, Optional, Sequence, Tuple

from .engine import LLMEngine
from .scheduler import SamplingParams, StepOutput


@dataclass
class RequestStats:
    arrival: float
    prompt_tokens: int
    first_token: Optional[float] = None
    last_token: Optional[float] = None
    completion_tokens: int = 0


class Metrics:
    """Counters/histograms exported on /metrics with the vllm:* names GPUStack's runtime mapping
    expects (gpustack/assets/metrics_config/metrics_config.yaml:74-98)."""

    TTFT_BUCKETS = [0.001, 0.005, 0.01, 0.02, 0.04, 0.06, 0.08, 0.1, 0.25, 0.5, 0.75, 1.0, 2.5, 5.0, 7.5,
                    10.0, 20.0, 40.0, 80.0, 160.0, 640.0, 2560.0]
    TPOT_BUCKETS = [0.001, 0.0025, 0.005, 0.01, 0.015, 0.02, 0.025, 0.03, 0.04, 0.05, 0.075, 0.1, 0.15, 0.2,
                    0.3, 0.4, 0.5, 0.75, 1.0, 2.5, 5.0]
    E2E_BUCKETS = [0.3, 0.5, 0.8, 1.0, 1.5, 2.0, 2.5, 5.0, 10.0, 15.0, 20.0, 30.0, 40.0, 50.0, 60.0, 120.0,
                   240.0, 480.0, 960.0, 1920.0, 7680.0]

    def __init__(self):
        self.prompt_tokens = 0
        self.generation_tokens = 0
        self.success: Dict[str, int] = {"stop": 0, "length": 0, "abort": 0}
        self.hist = {"ttft": ([0] * (len(self.TTFT_BUCKETS) + 1), 0.0),
                     "tpot": ([0] * (len(self.TPOT_BUCKETS) + 1), 0.0),
                     "e2e": ([0] * (len(self.E2E_BUCKETS) + 1), 0.0)}
        self._buckets = {"ttft": self.TTFT_BUCKETS, "tpot": self.TPOT_BUCKETS, "e2e": self.E2E_BUCKETS}

    def observe(self, name: str, v: float) -> None:
        counts, total = self.hist[name]
        b = self._buckets[name]
        i = 0
        while i < len(b) and v > b[i]:
            i += 1
        counts[i] += 1
        self.hist[name] = (counts, total + v)

    def finish(self, st: RequestStats, reason: str) -> None:
        self.success[reason] = self.success.get(reason, 0) + 1
        self.generation_tokens += st.completion_tokens
        if st.first_token is not None:
            self.observe("ttft", st.first_token - st.arrival)
            if st.completion_tokens > 1 and st.last_token is not None:
                self.observe("tpot", (st.last_token - st.first_token) / (st.completion_tokens - 1))
            self.observe("e2e", (st.last_token or st.first_token) - st.arrival)


class AsyncEngine:
    def __init__(self, engine: LLMEngine):
        self.engine = engine
        self.metrics = Metrics()
        self._inbox: "queue.SimpleQueue" = queue.SimpleQueue()
        self._streams: Dict[str, Tuple[asyncio.AbstractEventLoop, asyncio.Queue]] = {}
        self._stats: Dict[str, RequestStats] = {}
        self._ids = itertools.count()
        self._stop = False
        self._wake = threading.Event()
        self.errored: Optional[BaseException] = None
        self._thread = threading.Thread(target=self._loop, name="b200-engine-core", daemon=True)
        self._thread.start()

    # ------------------------------------------------------------------ engine thread
    def _loop(self) -> None:
        eng = self.engine
        try:
            while not self._stop:
                self._drain_inbox()
                if not eng.has_work():
                    self._wake.wait(timeout=0.05)
                    self._wake.clear()
                    continue
                out = eng.step()
                if out is not None and out.req_ids:
                    self._dispatch(out)
        except BaseException as e:  # the server turns unhealthy; GPUStack restarts the instance
            self.errored = e
            for rid, (loop, q) in list(self._streams.items()):
                loop.call_soon_threadsafe(q.put_nowait, ([], "error"))
            raise

    def _drain_inbox(self) -> None:
        while True:
            try:
                kind, rid, payload = self._inbox.get_nowait()
            except queue.Empty:
                return
            if kind == "add":
                prompt_ids, params = payload
                try:
                    self.engine.add_request(rid, prompt_ids, params)
                except Exception as e:  # invalid request: report to the waiting handler
                    ent = self._streams.get(rid)
                    if ent:
                        ent[0].call_soon_threadsafe(ent[1].put_nowait, (e, "invalid"))
            elif kind == "abort":
                self.engine.abort_request(rid)

    def _dispatch(self, out: StepOutput) -> None:
        by_loop: Dict[asyncio.AbstractEventLoop, List] = {}
        for rid, tok, fin in zip(out.req_ids, out.token_ids, out.finished):
            ent = self._streams.get(rid)
            if ent is not None:
                by_loop.setdefault(ent[0], []).append((ent[1], tok, fin))
        now = time.monotonic()
        for loop, items in by_loop.items():
            loop.call_soon_threadsafe(self._deliver, items, now)

    @staticmethod
    def _deliver(items, now) -> None:
        for q, tok, fin in items:
            q.put_nowait(([tok], fin, now))

    # ------------------------------------------------------------------ asyncio side
    def new_request_id(self, prefix: str = "cmpl") -> str:
        return f"{prefix}-{next(self._ids):08x}{int(time.time() * 1000) & 0xffffff:06x}"

    async def generate(self, req_id: str, prompt_ids: Sequence[int], params: SamplingParams
                       ) -> AsyncIterator[Tuple[List[int], Optional[str]]]:
        """Yields (new_token_ids, finish_reason) -- tokens produced while the consumer was busy
        are coalesced into one item."""
        if self.errored is not None:
            raise RuntimeError(f"engine core died: {self.errored!r}")
        loop = asyncio.get_running_loop()
        q: asyncio.Queue = asyncio.Queue()
        self._streams[req_id] = (loop, q)
        st = self._stats[req_id] = RequestStats(arrival=time.monotonic(), prompt_tokens=len(prompt_ids))
        self.metrics.prompt_tokens += len(prompt_ids)
        self._inbox.put(("add", req_id, (list(prompt_ids), params)))
        self._wake.set()
        finished = None
        try:
            while finished is None:
                item = await q.get()
                if item[1] == "invalid":
                    raise ValueError(str(item[0]))
                if item[1] == "error":
                    raise RuntimeError(f"engine core died: {self.errored!r}")
                toks, finished, ts = list(item[0]), item[1], item[2]
                while finished is None and not q.empty():     # coalesce
                    more = q.get_nowait()
                    if more[1] in ("invalid", "error"):
                        raise RuntimeError("engine core died")
                    toks.extend(more[0]); finished = more[1]; ts = more[2]
                if st.first_token is None:
                    st.first_token = ts
                st.last_token = ts
                st.completion_tokens += len(toks)
                yield toks, finished
        finally:
            self._streams.pop(req_id, None)
            self._stats.pop(req_id, None)
            if finished is None:            # client went away
                self._inbox.put(("abort", req_id, None))
                self._wake.set()
                self.metrics.finish(st, "abort")
            else:
                self.metrics.finish(st, finished)

    def is_healthy(self) -> bool:
        return self.errored is None and self._thread.is_alive()

    def shutdown(self) -> None:
        self._stop = True
        self._wake.set()
        self._thread.join(timeout=5)
        self.engine.shutdown()
