"""HTTP frontend process: uvicorn + OpenAI routes + tokenizer, talking to the engine core process
over a pipe (EngineClient <-> EngineCoreServer).  Keeping JSON/SSE work in its own process is
what the reference engine does too (API server <-> EngineCore over ZMQ, $VLLM/v1/engine/)."""
from __future__ import annotations

import logging
from typing import Optional


def run_frontend(conn, host: str, port: int, served_model_name: str, max_model_len: int, vocab_size: int,
                 bos: Optional[int], eos: Optional[int], model_path: Optional[str], log_level: str = "warning",
                 stream_interval: int = 1) -> None:
    import uvicorn

    from .async_engine import EngineClient
    from .server import create_app
    from .tokenizer import load_tokenizer

    logging.basicConfig(level=getattr(logging, log_level.upper(), logging.WARNING))
    tok = load_tokenizer(model_path, vocab_size, bos, eos)
    client = EngineClient(conn)
    app = create_app(client, tok, served_model_name, max_model_len, stream_interval=stream_interval)
    try:
        uvicorn.run(app, host=host, port=port, log_level=log_level, access_log=False, timeout_keep_alive=5,
                    backlog=4096)
    finally:
        client.shutdown()


def spawn_frontend(engine, host: str, port: int, served_model_name: str, model_path: Optional[str] = None,
                   log_level: str = "warning", stream_interval: int = 1):
    """Start the frontend process for `engine` (running in THIS process); returns
    (process, core_server).  The caller runs core_server.run() (blocking busy loop)."""
    import multiprocessing as mp

    from .async_engine import EngineCoreServer

    ctx = mp.get_context("spawn")
    parent, child = ctx.Pipe(duplex=True)
    m = engine.mcfg
    proc = ctx.Process(target=run_frontend, name="b200-frontend", daemon=True,
                       args=(child, host, port, served_model_name, engine.ecfg.max_model_len, m.vocab_size,
                             m.bos_token_id, m.eos_token_id, model_path, log_level, stream_interval))
    proc.start()
    child.close()
    return proc, EngineCoreServer(engine, parent)
