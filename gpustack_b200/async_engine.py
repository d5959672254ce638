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


class _StreamHub:
    """Request/stream bookkeeping shared by the in-process (thread) and out-of-process (pipe) engine
    fronts: per-request asyncio queues, token coalescing, latency metrics."""

    def __init__(self):
        self.metrics = Metrics()
        self._streams: Dict[str, Tuple[asyncio.AbstractEventLoop, asyncio.Queue]] = {}
        self._stats: Dict[str, RequestStats] = {}
        self._ids = itertools.count()
        self._forced: Dict[str, str] = {}      # finish reason decided by the HTTP layer (stop strings)
        self.errored: Optional[BaseException] = None

    # -- to be provided by the transport
    def _submit(self, rid: str, prompt_ids: List[int], params: SamplingParams) -> None:
        raise NotImplementedError

    def _abort(self, rid: str) -> None:
        raise NotImplementedError

    def snapshot(self) -> dict:
        raise NotImplementedError

    def is_healthy(self) -> bool:
        raise NotImplementedError

    # -- delivery (called from the transport's thread)
    def _dispatch(self, req_ids, token_ids, finished) -> None:
        by_loop: Dict[asyncio.AbstractEventLoop, List] = {}
        for rid, tok, fin in zip(req_ids, token_ids, finished):
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

    def _fail_request(self, rid: str, exc: Exception) -> None:
        ent = self._streams.get(rid)
        if ent:
            ent[0].call_soon_threadsafe(ent[1].put_nowait, (exc, "invalid"))

    def _fail_all(self) -> None:
        for rid, (loop, q) in list(self._streams.items()):
            loop.call_soon_threadsafe(q.put_nowait, ([], "error"))

    def finish_early(self, req_id: str, reason: str) -> None:
        """The consumer ends the stream itself (e.g. a stop STRING matched): the engine-side request
        is still aborted when the generator closes, but it is counted with `reason`, not as an abort."""
        self._forced[req_id] = reason

    # -- asyncio side
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
        self._submit(req_id, list(prompt_ids), params)
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
            forced = self._forced.pop(req_id, None)
            if finished is None:            # consumer stopped reading: client went away, or a stop string
                self._abort(req_id)
                self.metrics.finish(st, forced or "abort")
            else:
                self.metrics.finish(st, finished)


def engine_snapshot(engine) -> dict:
    s = engine.sched
    return {"running": s.num_running, "waiting": s.num_waiting, "kv_usage": s.kv_usage(),
            "preemptions": s.num_preemptions, "generation_tokens": engine.counters["generation_tokens"],
            "prefix_queries": s.pool.queries, "prefix_hits": s.pool.hits}


class AsyncEngine(_StreamHub):
    """Engine core on a background THREAD of the serving process (tests, small deployments)."""

    def __init__(self, engine: LLMEngine):
        super().__init__()
        self.engine = engine
        self._inbox: "queue.SimpleQueue" = queue.SimpleQueue()
        self._stop = False
        self._wake = threading.Event()
        self._thread = threading.Thread(target=self._loop, name="b200-engine-core", daemon=True)
        self._thread.start()

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
                    self._dispatch(out.req_ids, out.token_ids, out.finished)
        except BaseException as e:  # the server turns unhealthy; GPUStack restarts the instance
            self.errored = e
            self._fail_all()
            raise

    def _drain_inbox(self) -> None:
        while True:
            try:
                kind, rid, payload = self._inbox.get_nowait()
            except queue.Empty:
                return
            if kind == "add":
                try:
                    self.engine.add_request(rid, payload[0], payload[1])
                except Exception as e:  # invalid request: report to the waiting handler
                    self._fail_request(rid, e)
            elif kind == "abort":
                self.engine.abort_request(rid)

    def _submit(self, rid, prompt_ids, params) -> None:
        self._inbox.put(("add", rid, (prompt_ids, params)))
        self._wake.set()

    def _abort(self, rid) -> None:
        self._inbox.put(("abort", rid, None))
        self._wake.set()

    def snapshot(self) -> dict:
        return engine_snapshot(self.engine)

    def is_healthy(self) -> bool:
        return self.errored is None and self._thread.is_alive()

    def shutdown(self) -> None:
        self._stop = True
        self._wake.set()
        self._thread.join(timeout=5)
        self.engine.shutdown()


# --------------------------------------------------------------------------- process split
class EngineCoreServer:
    """Runs in the ENGINE process: the busy loop (schedule -> GPU step -> outputs) fed by a pipe.
    Counterpart of the reference engine's EngineCoreProc ($VLLM/v1/engine/core.py): the HTTP front
    lives in another process, so JSON/SSE work never competes with the step loop for the GIL."""

    def __init__(self, engine: LLMEngine, conn):
        self.engine, self.conn = engine, conn

    def run(self) -> None:
        eng, conn = self.engine, self.conn
        last_beat = 0.0
        try:
            while True:
                while conn.poll(0.0 if eng.has_work() else 0.02):
                    msg = conn.recv()
                    kind = msg[0]
                    if kind == "add":
                        _, rid, prompt_ids, params = msg
                        try:
                            eng.add_request(rid, prompt_ids, params)
                        except Exception as e:  # noqa: BLE001
                            conn.send(("invalid", rid, str(e)))
                    elif kind == "abort":
                        eng.abort_request(msg[1])
                    elif kind == "shutdown":
                        return
                if eng.has_work():
                    out = eng.step()
                    if out is not None and out.req_ids:
                        conn.send(("out", out.req_ids, out.token_ids, out.finished, engine_snapshot(eng)))
                now = time.monotonic()
                if now - last_beat > 1.0:
                    conn.send(("beat", engine_snapshot(eng)))
                    last_beat = now
        except (EOFError, OSError):      # the frontend went away (BrokenPipe / ConnectionReset): normal shutdown
            return
        except BaseException as e:  # noqa: BLE001
            try:
                conn.send(("died", repr(e)))
            except (OSError, ValueError):
                pass
            raise


class EngineClient(_StreamHub):
    """Runs in the FRONTEND process: same interface as AsyncEngine, transport = a pipe to an
    EngineCoreServer."""

    def __init__(self, conn):
        super().__init__()
        self.conn = conn
        self._send_lock = threading.Lock()
        self._snap = {"running": 0, "waiting": 0, "kv_usage": 0.0, "preemptions": 0, "generation_tokens": 0,
                      "prefix_queries": 0, "prefix_hits": 0}
        self._last_msg = time.monotonic()
        self._thread = threading.Thread(target=self._reader, name="b200-engine-client", daemon=True)
        self._thread.start()

    def _reader(self) -> None:
        try:
            while True:
                msg = self.conn.recv()
                self._last_msg = time.monotonic()
                kind = msg[0]
                if kind == "out":
                    self._snap = msg[4]
                    self._dispatch(msg[1], msg[2], msg[3])
                elif kind == "beat":
                    self._snap = msg[1]
                elif kind == "invalid":
                    self._fail_request(msg[1], ValueError(msg[2]))
                elif kind == "died":
                    raise RuntimeError(msg[1])
        except BaseException as e:  # noqa: BLE001
            self.errored = e
            self._fail_all()

    def _send(self, msg) -> None:
        with self._send_lock:
            self.conn.send(msg)

    def _submit(self, rid, prompt_ids, params) -> None:
        self._send(("add", rid, prompt_ids, params))

    def _abort(self, rid) -> None:
        try:
            self._send(("abort", rid))
        except Exception:  # noqa: BLE001
            pass

    def snapshot(self) -> dict:
        return dict(self._snap)

    def is_healthy(self) -> bool:
        return self.errored is None and self._thread.is_alive() and time.monotonic() - self._last_msg < 30.0

    def shutdown(self) -> None:
        try:
            self._send(("shutdown",))
        except Exception:  # noqa: BLE001
            pass
