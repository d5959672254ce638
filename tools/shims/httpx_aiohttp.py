"""Minimal stand-in for the third-party `httpx_aiohttp` package (not installed offline), so that the
reference's benchmarks/benchmark_serving.py runs UNMODIFIED: it only needs `AiohttpTransport`, an
httpx async transport that sends requests through an aiohttp ClientSession
(/root/reference/benchmarks/benchmark_serving.py:13,241-256).  Test / measurement infrastructure --
not imported by the engine."""
from __future__ import annotations

import typing

import aiohttp
import httpx


class _ResponseStream(httpx.AsyncByteStream):
    def __init__(self, resp: aiohttp.ClientResponse):
        self._resp = resp

    async def __aiter__(self) -> typing.AsyncIterator[bytes]:
        async for chunk in self._resp.content.iter_any():
            if chunk:
                yield chunk

    async def aclose(self) -> None:
        self._resp.release()


class AiohttpTransport(httpx.AsyncBaseTransport):
    def __init__(self, client: aiohttp.ClientSession):
        self.client = client

    async def handle_async_request(self, request: httpx.Request) -> httpx.Response:
        timeout = request.extensions.get("timeout", {}) or {}
        body = request.content if hasattr(request, "_content") else b"".join([c async for c in request.stream])
        headers = [(k, v) for k, v in request.headers.items() if k.lower() not in ("content-length", "host")]
        try:
            resp = await self.client.request(
                request.method, str(request.url), headers=headers, data=body or None, allow_redirects=False,
                timeout=aiohttp.ClientTimeout(total=None, connect=timeout.get("connect"), sock_read=timeout.get("read")),
                skip_auto_headers=("User-Agent",))
        except aiohttp.ClientConnectorError as e:
            raise httpx.ConnectError(str(e), request=request) from e
        except aiohttp.ServerTimeoutError as e:
            raise httpx.ReadTimeout(str(e), request=request) from e
        except aiohttp.ClientError as e:
            raise httpx.TransportError(str(e)) from e
        return httpx.Response(status_code=resp.status, headers=list(resp.headers.items()), stream=_ResponseStream(resp),
                              request=request)

    async def aclose(self) -> None:   # the session belongs to the caller
        return None


class HttpxAiohttpClient(httpx.AsyncClient):
    """Only referenced by `openai` at import time (its optional DefaultAioHttpClient subclasses it when
    the package is importable); the harness itself passes an AiohttpTransport to a plain httpx client."""
