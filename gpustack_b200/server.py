"""OpenAI-compatible HTTP surface of the engine -- the drop-in boundary GPUStack talks to.

Contract:
  * POST /v1/chat/completions, /v1/completions: stream = SSE `data: {json}\\n\\n` ... `data: [DONE]`;
    chunks validate as openai.types.chat.ChatCompletionChunk (parsed that way by
    gpustack/api/middlewares.py:219-262); final usage chunk when stream_options.include_usage
    (GPUStack forces it on, gpustack/routes/openai.py:284-287); accepts max_completion_tokens,
    max_tokens, ignore_eos, temperature, stop, seed (benchmarks/benchmark_serving.py:106-113);
  * GET /v1/models -> {data:[{id, max_model_len}]} (gpustack/worker/model_meta.py:47-66);
  * GET /health (readiness probe, gpustack/worker/serve_manager.py:1000-1048), GET /version
    (gpustack/worker/runtime_metrics_client.py:12-16), GET /metrics (Prometheus text with the
    vllm:* names of gpustack/assets/metrics_config/metrics_config.yaml:74-98);
  * errors are OpenAI error envelopes {error:{message,type,code}} (gpustack/routes/openai.py:305-328).
"""
from __future__ import annotations

import json
import time
from typing import Any, Dict, List, Optional

from fastapi import FastAPI, Request
from fastapi.responses import JSONResponse, PlainTextResponse, StreamingResponse

from . import __version__
from .async_engine import AsyncEngine
from .scheduler import SamplingParams
from .tokenizer import IncrementalDetokenizer

_dumps = lambda o: json.dumps(o, separators=(",", ":"), ensure_ascii=False)


def _error(status: int, message: str, etype: str = "invalid_request_error", code: Optional[str] = None):
    return JSONResponse(status_code=status,
                        content={"error": {"message": message, "type": etype, "param": None, "code": code or status}})


def _sampling(body: Dict[str, Any], max_model_len: int, prompt_len: int, tokenizer) -> SamplingParams:
    mt = body.get("max_completion_tokens")
    if mt is None:
        mt = body.get("max_tokens")
    if mt is None:
        mt = max(1, max_model_len - prompt_len)
    mt = int(mt)
    if mt < 1:
        raise ValueError("max_tokens must be at least 1")
    stop_ids = tuple(int(t) for t in (body.get("stop_token_ids") or ()))
    temp = body.get("temperature")
    # OpenAI / vLLM default: temperature 1.0 when the request does not set it
    return SamplingParams(max_tokens=mt, temperature=1.0 if temp is None else float(temp),
                          top_p=float(body.get("top_p") or 1.0), ignore_eos=bool(body.get("ignore_eos", False)),
                          stop_token_ids=stop_ids, seed=body.get("seed"))


def create_app(aeng: AsyncEngine, tokenizer, served_model_name: str, max_model_len: int,
               stream_interval: int = 1) -> FastAPI:
    """stream_interval: engine outputs folded into one SSE chunk (the reference engine's `--stream-interval`): 1 =
    a chunk per output, N = a chunk every N outputs; the first output (TTFT) and the last (finish_reason) are
    always sent at once."""
    app = FastAPI(title="b200-native engine", version=__version__)
    created = int(time.time())

    def model_ok(body) -> bool:
        m = body.get("model")
        return m in (None, "", served_model_name)

    # ------------------------------------------------------------------ plumbing endpoints
    @app.get("/health")
    async def health():
        if not aeng.is_healthy():
            return PlainTextResponse("engine core is not running", status_code=503)
        return PlainTextResponse("ok")

    @app.get("/version")
    async def version():
        return {"version": __version__}

    @app.get("/v1/models")
    async def models():
        return {"object": "list", "data": [{
            "id": served_model_name, "object": "model", "created": created, "owned_by": "b200-native",
            "root": served_model_name, "parent": None, "max_model_len": max_model_len,
            "meta": {"engine": "b200-native", "n_ctx_train": max_model_len}}]}

    @app.get("/metrics")
    async def metrics():
        return PlainTextResponse(render_metrics(aeng, served_model_name),
                                 media_type="text/plain; version=0.0.4; charset=utf-8")

    @app.post("/tokenize")
    async def tokenize(req: Request):
        body = await req.json()
        ids = tokenizer.encode(body.get("prompt", ""))
        return {"tokens": ids, "count": len(ids), "max_model_len": max_model_len,
                "vocab_size": getattr(tokenizer, "vocab_size", None)}

    # ------------------------------------------------------------------ generation
    async def _parse(req: Request):
        try:
            body = await req.json()
        except Exception:
            return None, _error(400, "request body is not valid JSON")
        if not isinstance(body, dict):
            return None, _error(400, "request body must be a JSON object")
        if not model_ok(body):
            return None, _error(404, f"The model `{body.get('model')}` does not exist.", "NotFoundError", "model_not_found")
        if int(body.get("n") or 1) != 1:
            return None, _error(400, "n > 1 is not supported")
        return body, None

    def _stops(body) -> List[str]:
        s = body.get("stop")
        if s is None:
            return []
        return [s] if isinstance(s, str) else [x for x in s if isinstance(x, str)]

    async def _run(body, prompt_ids: List[int], chat: bool):
        try:
            if len(prompt_ids) == 0:
                raise ValueError("prompt is empty")
            sp = _sampling(body, max_model_len, len(prompt_ids), tokenizer)
            if len(prompt_ids) + sp.max_tokens > max_model_len:
                if body.get("max_tokens") is None and body.get("max_completion_tokens") is None:
                    sp.max_tokens = max_model_len - len(prompt_ids)
                if sp.max_tokens < 1 or len(prompt_ids) + sp.max_tokens > max_model_len:
                    raise ValueError(f"This model's maximum context length is {max_model_len} tokens. However, you "
                                     f"requested {len(prompt_ids) + sp.max_tokens} tokens ({len(prompt_ids)} in the "
                                     f"prompt, {sp.max_tokens} in the completion).")
        except ValueError as e:
            return _error(400, str(e))
        rid = aeng.new_request_id("chatcmpl" if chat else "cmpl")
        stops = _stops(body)
        stream = bool(body.get("stream", False))
        include_usage = bool((body.get("stream_options") or {}).get("include_usage", False))
        obj = "chat.completion" if chat else "text_completion"
        n_prompt = len(prompt_ids)

        def usage(n_out):
            return {"prompt_tokens": n_prompt, "completion_tokens": n_out, "total_tokens": n_prompt + n_out}

        def chunk(text, finish, first):
            if chat:
                delta = {"content": text}
                if first:
                    delta = {"role": "assistant", "content": text}
                ch = {"index": 0, "delta": delta, "logprobs": None, "finish_reason": finish}
                o = "chat.completion.chunk"
            else:
                ch = {"index": 0, "text": text, "logprobs": None, "finish_reason": finish}
                o = "text_completion"
            return {"id": rid, "object": o, "created": int(time.time()), "model": served_model_name, "choices": [ch]}

        async def events():
            """yields (text_delta, finish_reason, is_first, n_out).  With stop strings the last
            len(longest stop)-1 characters are held back until they can no longer start a match."""
            detok = IncrementalDetokenizer(tokenizer)
            n_out, first, pending = 0, True, ""
            hold = max((len(s) for s in stops), default=1) - 1
            gen = aeng.generate(rid, prompt_ids, sp)
            try:
                async for toks, fin in gen:
                    n_out += len(toks)
                    delta = detok.push(toks, final=fin is not None)
                    if stops:
                        pending += delta
                        cut = min((pending.find(s) for s in stops if s in pending), default=-1)
                        if cut >= 0:
                            aeng.finish_early(rid, "stop")      # counted as a stop, not as an abort
                            yield pending[:cut], "stop", first, n_out
                            return
                        if fin is None and hold > 0:
                            delta, pending = pending[:-hold] if len(pending) > hold else "", pending[-hold:]
                        else:
                            delta, pending = pending, ""
                    yield delta, fin, first, n_out
                    first = False
            finally:
                await gen.aclose()

        if stream:
            async def sse():
                n_out = 0
                try:
                    held, n_held = "", 0
                    async for delta, fin, first, n_out in events():
                        if stream_interval > 1 and not first:
                            held += delta
                            n_held += 1
                            if fin is None and n_held < stream_interval:
                                continue
                            delta, held, n_held = held, "", 0
                        if delta or fin is not None or first:
                            yield f"data: {_dumps(chunk(delta, fin, first))}\n\n"
                    if include_usage:
                        u = chunk("", None, False)
                        u["choices"] = []
                        u["usage"] = usage(n_out)
                        yield f"data: {_dumps(u)}\n\n"
                    yield "data: [DONE]\n\n"
                except (ValueError, RuntimeError) as e:
                    err = {"error": {"message": str(e), "type": "invalid_request_error", "code": 400}}
                    yield f"data: {_dumps(err)}\n\n"
            return StreamingResponse(sse(), media_type="text/event-stream",
                                     headers={"Cache-Control": "no-cache", "X-Accel-Buffering": "no"})
        text, fin, n_out = "", None, 0
        try:
            async for delta, fin, _first, n_out in events():
                text += delta
        except ValueError as e:
            return _error(400, str(e))
        except RuntimeError as e:
            return _error(500, str(e), "internal_error")
        if chat:
            choice = {"index": 0, "message": {"role": "assistant", "content": text}, "logprobs": None, "finish_reason": fin}
        else:
            choice = {"index": 0, "text": text, "logprobs": None, "finish_reason": fin}
        return {"id": rid, "object": obj, "created": int(time.time()), "model": served_model_name,
                "choices": [choice], "usage": usage(n_out)}

    @app.post("/v1/chat/completions")
    async def chat_completions(req: Request):
        body, err = await _parse(req)
        if err is not None:
            return err
        msgs = body.get("messages")
        if not isinstance(msgs, list) or not msgs:
            return _error(400, "`messages` must be a non-empty list")
        try:
            ids = tokenizer.apply_chat_template(msgs)
        except Exception as e:
            return _error(400, f"could not render the chat template: {e}")
        return await _run(body, ids, chat=True)

    @app.post("/v1/completions")
    async def completions(req: Request):
        body, err = await _parse(req)
        if err is not None:
            return err
        prompt = body.get("prompt")
        if isinstance(prompt, str):
            ids = tokenizer.encode(prompt)
        elif isinstance(prompt, list) and prompt and all(isinstance(t, int) for t in prompt):
            ids = list(prompt)                       # token-id prompt (token-exact benchmarks)
        elif isinstance(prompt, list) and len(prompt) == 1 and isinstance(prompt[0], str):
            ids = tokenizer.encode(prompt[0])
        elif isinstance(prompt, list) and len(prompt) == 1 and isinstance(prompt[0], list):
            ids = list(prompt[0])
        else:
            return _error(400, "`prompt` must be a string or a list of token ids")
        return await _run(body, ids, chat=False)

    return app


def render_metrics(aeng: AsyncEngine, model: str) -> str:
    """Prometheus exposition with vLLM's metric names so that GPUStack's `vLLM` runtime mapping
    (metrics_config.yaml:74-98) scrapes this engine unchanged."""
    m, snap = aeng.metrics, aeng.snapshot()
    lab = f'{{model_name="{model}",engine="0"}}'
    L: List[str] = []

    def gauge(name, help_, v):
        L.append(f"# HELP {name} {help_}\n# TYPE {name} gauge\n{name}{lab} {v}")

    def counter(name, help_, v, extra=""):
        L.append(f"# HELP {name} {help_}\n# TYPE {name} counter")
        L.append(f"{name}_total{lab if not extra else extra} {v}")

    gauge("vllm:num_requests_running", "Number of requests in model execution batches.", snap["running"])
    gauge("vllm:num_requests_waiting", "Number of requests waiting to be processed.", snap["waiting"])
    gauge("vllm:kv_cache_usage_perc", "KV-cache usage. 1 means 100 percent usage.", round(snap["kv_usage"], 6))
    gauge("vllm:gpu_cache_usage_perc", "KV-cache usage (legacy name).", round(snap["kv_usage"], 6))
    counter("vllm:num_preemptions", "Cumulative number of preemptions from the engine.", snap["preemptions"])
    counter("vllm:prompt_tokens", "Number of prefill tokens processed.", m.prompt_tokens)
    counter("vllm:prefix_cache_queries", "Prefix cache queries, in terms of number of queried tokens.",
            snap.get("prefix_queries", 0))
    counter("vllm:prefix_cache_hits", "Prefix cache hits, in terms of number of cached tokens.",
            snap.get("prefix_hits", 0))
    counter("vllm:generation_tokens", "Number of generation tokens processed.", snap["generation_tokens"])
    L.append("# HELP vllm:request_success Count of successfully processed requests.\n# TYPE vllm:request_success counter")
    for reason, n in m.success.items():
        L.append(f'vllm:request_success_total{{model_name="{model}",engine="0",finished_reason="{reason}"}} {n}')
    for key, name, help_ in (("ttft", "vllm:time_to_first_token_seconds", "Histogram of time to first token in seconds."),
                             ("tpot", "vllm:time_per_output_token_seconds", "Histogram of time per output token in seconds."),
                             ("tpot", "vllm:inter_token_latency_seconds", "Histogram of inter-token latency in seconds."),
                             ("e2e", "vllm:e2e_request_latency_seconds", "Histogram of e2e request latency in seconds.")):
        counts, total = m.hist[key]
        buckets = m._buckets[key]
        L.append(f"# HELP {name} {help_}\n# TYPE {name} histogram")
        acc = 0
        for b, c in zip(buckets, counts):
            acc += c
            L.append(f'{name}_bucket{{model_name="{model}",engine="0",le="{b}"}} {acc}')
        acc += counts[-1]
        L.append(f'{name}_bucket{{model_name="{model}",engine="0",le="+Inf"}} {acc}')
        L.append(f"{name}_sum{lab} {total}\n{name}_count{lab} {acc}")
    return "\n".join(L) + "\n"
