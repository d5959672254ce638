#!/usr/bin/env python
"""Declared STAND-IN for the part of GPUStack that sits in front of two engine replicas
(BASELINE.json configs[2] / configs[4]: "through gpustack/gateway", "TP=4 x 2 replicas, gateway
round-robin").  GPUStack itself (server + worker + Higress) cannot run offline (SURVEY.md section 0.2,
Appendix C.1); this relay reproduces only the two behaviours that touch the request path:

  * instance choice: `RoundRobinStrategy.select_instance` -- one `itertools.cycle` per model id,
    rebuilt when the list of running instances changes (gpustack/http_proxy/strategies.py:18-35,
    caller gpustack/routes/openai.py:381-390);
  * the worker proxy hop: the request body is forwarded as is and the response is streamed back in
    1 KiB chunks over a fresh upstream connection per request
    (gpustack/routes/worker/proxy.py:26-98, `iter_chunked(1024)`, connector `force_close=True`
    gpustack/worker/worker.py:318-330).

Measurement infrastructure (tools/), not part of the engine.

  python tools/rr_relay.py --port 9000 --model llama3-8b http://127.0.0.1:8001 http://127.0.0.1:8002
"""
from __future__ import annotations

import argparse
import itertools
import json
from typing import Dict, List

import aiohttp
from aiohttp import web


class RoundRobin:
    """Same state machine as the reference's RoundRobinStrategy, keyed by model id."""

    def __init__(self):
        self._iterators: Dict[str, itertools.cycle] = {}
        self._lists: Dict[str, List[str]] = {}

    def select(self, model_id: str, instances: List[str]) -> str:
        if not instances:
            raise RuntimeError("No instances available")
        if model_id not in self._iterators or self._lists[model_id] != instances:
            self._iterators[model_id] = itertools.cycle(instances)
            self._lists[model_id] = list(instances)
        return next(self._iterators[model_id])


def make_app(model: str, upstreams: List[str]) -> web.Application:
    lb = RoundRobin()
    counts = {u: 0 for u in upstreams}

    async def on_startup(app):
        app["session"] = aiohttp.ClientSession(connector=aiohttp.TCPConnector(limit=2000, force_close=True),
                                               timeout=aiohttp.ClientTimeout(total=1800))

    async def on_cleanup(app):
        await app["session"].close()

    async def stats(_req):
        return web.json_response({"model": model, "requests_per_instance": counts})

    async def proxy(req: web.Request):
        body = await req.read()
        if req.method == "POST" and body:
            try:
                m = json.loads(body).get("model")
            except ValueError:
                m = None
            if m not in (None, "", model):
                return web.json_response({"error": {"message": f"The model `{m}` does not exist.", "type": "NotFoundError",
                                                    "param": None, "code": "model_not_found"}}, status=404)
        target = lb.select(model, upstreams)
        counts[target] += 1
        headers = {k: v for k, v in req.headers.items() if k.lower() not in ("host", "transfer-encoding", "content-length")}
        try:
            async with req.app["session"].request(req.method, target + req.path_qs, headers=headers, data=body or None) as up:
                resp = web.StreamResponse(status=up.status, headers={k: v for k, v in up.headers.items()
                                                                     if k.lower() not in ("transfer-encoding", "content-length",
                                                                                          "connection")})
                await resp.prepare(req)
                async for chunk in up.content.iter_chunked(1024):
                    await resp.write(chunk)
                await resp.write_eof()
                return resp
        except aiohttp.ClientError as e:
            return web.json_response({"error": {"message": f"An unexpected error occurred: {e}", "type": "ServiceUnavailable",
                                                "param": None, "code": 503}}, status=503)

    app = web.Application(client_max_size=64 << 20)
    app.on_startup.append(on_startup)
    app.on_cleanup.append(on_cleanup)
    app.router.add_get("/relay/stats", stats)
    app.router.add_route("*", "/{tail:.*}", proxy)
    return app


def main():
    ap = argparse.ArgumentParser()
    ap.add_argument("--host", default="127.0.0.1")
    ap.add_argument("--port", type=int, default=9000)
    ap.add_argument("--model", required=True, help="served model name (= GPUStack model id of the replicas)")
    ap.add_argument("upstreams", nargs="+", help="base URLs of the RUNNING instances, e.g. http://127.0.0.1:8001")
    a = ap.parse_args()
    web.run_app(make_app(a.model, [u.rstrip("/") for u in a.upstreams]), host=a.host, port=a.port, print=None,
                access_log=None)


if __name__ == "__main__":
    main()
