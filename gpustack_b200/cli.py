"""`b200-engine serve <model_path> ...` -- the process GPUStack's worker launches.

The argv contract is what the reference launcher composes:
  * custom backend (gpustack/worker/backends/custom.py:141-159): run_command with {{model_path}},
    {{port}}, {{worker_ip}}, {{model_name}} substituted + flattened backend_parameters;
  * custom *version* of the built-in vLLM backend (gpustack/worker/backends/vllm.py:395-507,
    602-639): `<model_path> [--max-model-len N] [--tensor-parallel-size N] ... --host H --port P
    --served-model-name NAME`.
So the CLI accepts vLLM's flag names for everything it implements and tolerates (warns about and
ignores) vLLM flags it does not.  With a custom backend nothing injects --tensor-parallel-size
(custom.py:141-159): TP then defaults to the number of visible GPUs when --tensor-parallel-size is
`auto`, else 1.
"""
from __future__ import annotations

import argparse
import logging
import os
import socket
import subprocess
import sys
from typing import List, Optional

log = logging.getLogger("b200-engine")

RANDOM_MODELS = ("random:llama3-8b", "random:tiny", "random:tiny128", "random:mixtral-8x7b", "random:tiny-mixtral")


def build_parser() -> argparse.ArgumentParser:
    p = argparse.ArgumentParser(prog="b200-engine", description="B200-native OpenAI-compatible LLM engine")
    sub = p.add_subparsers(dest="cmd", required=True)
    s = sub.add_parser("serve", help="serve a local HF-format model directory")
    s.add_argument("model", help="model directory (config.json + safetensors) or random:<llama3-8b|tiny>")
    s.add_argument("--host", default="0.0.0.0")
    s.add_argument("--port", type=int, default=8000)
    s.add_argument("--served-model-name", nargs="+", default=None)
    s.add_argument("--tensor-parallel-size", "-tp", default="1")
    s.add_argument("--max-model-len", type=int, default=None)
    s.add_argument("--gpu-memory-utilization", type=float, default=0.90)
    s.add_argument("--max-num-seqs", type=int, default=256)
    s.add_argument("--max-num-batched-tokens", type=int, default=8192)
    s.add_argument("--num-gpu-blocks-override", type=int, default=None)
    s.add_argument("--enforce-eager", action="store_true")
    s.add_argument("--enable-prefix-caching", action=argparse.BooleanOptionalAction, default=True)
    s.add_argument("--mixed-steps", action=argparse.BooleanOptionalAction, default=True,
                   help="prefill steps also carry one decode token per running sequence (unified token budget)")
    s.add_argument("--seed", type=int, default=0)
    s.add_argument("--quantization", "-q", default=None, choices=[None, "fp8"])
    s.add_argument("--speculative-config", default=None,
                   help='JSON as the GPUStack launcher composes it (gpustack/worker/backends/vllm.py:333-367), e.g. '
                        '\'{"method": "ngram", "num_speculative_tokens": 3, "prompt_lookup_max": 4, '
                        '"prompt_lookup_min": 2}\'; draft-model methods (eagle3, mtp) run without speculation')
    s.add_argument("--kv-transfer-config", default=None,
                   help='what GPUStack\'s launcher passes when the model has extended_kv_cache enabled '
                        '(\'{"kv_connector":"LMCacheConnectorV1","kv_role":"kv_both"}\', worker/backends/vllm.py:468-483): '
                        'turns on the host-memory tier of the prefix cache, sized by LMCACHE_MAX_LOCAL_CPU_SIZE (GiB, '
                        'the env the launcher sets; default 5)')
    s.add_argument("--kv-host-cache-gib", type=float, default=None,
                   help="host-memory tier of the prefix cache in GiB (native spelling of the above)")
    s.add_argument("--stream-interval", type=int, default=1,
                   help="engine outputs per SSE chunk (first and last outputs of a request are never held back)")
    s.add_argument("--dtype", default="bfloat16")
    s.add_argument("--log-level", default="info")
    est = sub.add_parser("estimate", help="print the engine's exact VRAM requirement (GPUSTACK_MODEL_VRAM_CLAIM)")
    est.add_argument("model", help="model directory (config.json) or random:<llama3-8b|tiny>")
    est.add_argument("--tensor-parallel-size", "-tp", type=int, default=1)
    est.add_argument("--max-model-len", type=int, default=None)
    est.add_argument("--max-num-seqs", type=int, default=256)
    est.add_argument("--max-num-batched-tokens", type=int, default=8192)
    est.add_argument("--gpu-memory-utilization", type=float, default=0.90)
    est.add_argument("--quantization", "-q", default=None, choices=[None, "fp8"])
    est.add_argument("--kv-tokens", type=int, default=None, help="KV-cache capacity to size the claim for")
    w = sub.add_parser("_worker", help=argparse.SUPPRESS)
    w.add_argument("--engine-json", required=True)
    return p


def _free_port() -> int:
    with socket.socket() as s:
        s.bind(("127.0.0.1", 0))
        return s.getsockname()[1]


def _engine_config(args, tp: int):
    from .config import EngineConfig
    from .spec_decode import SpeculativeConfig

    spec = SpeculativeConfig.parse(args.speculative_config)          # ValueError on malformed JSON
    if spec is not None and not spec.supported:
        log.warning("speculative method %r needs a draft model and is not implemented: serving WITHOUT speculation "
                    "(results are unaffected, only latency)", spec.method)
        spec = None
    host_gib = args.kv_host_cache_gib
    if host_gib is None and args.kv_transfer_config:
        import json

        kt = json.loads(args.kv_transfer_config)
        if "LMCache" in str(kt.get("kv_connector", "")):
            # LMCACHE_CHUNK_SIZE is accepted and unused: this tier works at KV-block (16-token) granularity
            host_gib = float(os.environ.get("LMCACHE_MAX_LOCAL_CPU_SIZE", "5"))
        else:
            log.warning("kv connector %r is not implemented; serving without an extended KV cache", kt.get("kv_connector"))
    return EngineConfig(speculative_config=spec.to_dict() if spec is not None else None,
                        kv_host_cache_gib=float(host_gib or 0.0), model=args.model, served_model_name=(args.served_model_name or [None])[0],
                        max_model_len=args.max_model_len, tensor_parallel_size=tp,
                        gpu_memory_utilization=args.gpu_memory_utilization, max_num_seqs=args.max_num_seqs,
                        max_num_batched_tokens=args.max_num_batched_tokens,
                        num_gpu_blocks=args.num_gpu_blocks_override, enforce_eager=args.enforce_eager,
                        enable_prefix_caching=args.enable_prefix_caching, quantization=args.quantization,
                        mixed_steps=args.mixed_steps, seed=args.seed)


def make_engine(ecfg):
    from .config import llama3_8b, mixtral_8x7b, tiny_llama, tiny_mixtral
    from .engine import LLMEngine

    if ecfg.model.startswith("random:"):
        name = ecfg.model.split(":", 1)[1]
        mcfg = {"llama3-8b": llama3_8b, "tiny": tiny_llama, "tiny128": lambda: tiny_llama(128),
                "mixtral-8x7b": mixtral_8x7b, "tiny-mixtral": tiny_mixtral}[name]()
        return LLMEngine.from_engine_config(ecfg, mcfg=mcfg)
    return LLMEngine.from_engine_config(ecfg)


def _spawn_workers(ecfg, tp: int) -> List[subprocess.Popen]:
    import dataclasses
    import json

    port = _free_port()
    os.environ.update(RANK="0", LOCAL_RANK="0", WORLD_SIZE=str(tp), MASTER_ADDR="127.0.0.1", MASTER_PORT=str(port))
    procs = []
    for r in range(1, tp):
        env = dict(os.environ, RANK=str(r), LOCAL_RANK=str(r))
        procs.append(subprocess.Popen([sys.executable, "-m", "gpustack_b200.cli", "_worker", "--engine-json",
                                       json.dumps(dataclasses.asdict(ecfg))], env=env))
    return procs


def main(argv: Optional[List[str]] = None) -> int:
    args, unknown = build_parser().parse_known_args(argv)
    logging.basicConfig(level=getattr(logging, getattr(args, "log_level", "info").upper(), logging.INFO),
                        format="%(asctime)s %(levelname)s %(name)s: %(message)s")
    if args.cmd == "estimate":
        import json

        from .config import EngineConfig, ModelConfig, llama3_8b, tiny_llama
        from .resource_fit import estimate

        if args.model.startswith("random:"):
            mcfg, disk = (llama3_8b() if "8b" in args.model else tiny_llama()), None
        else:
            mcfg = ModelConfig.from_pretrained(args.model)
            disk = sum(f.stat().st_size for f in __import__("pathlib").Path(args.model).glob("*.safetensors")) or None
        ecfg = EngineConfig(model=args.model, tensor_parallel_size=args.tensor_parallel_size,
                            max_model_len=args.max_model_len, max_num_seqs=args.max_num_seqs,
                            max_num_batched_tokens=args.max_num_batched_tokens,
                            gpu_memory_utilization=args.gpu_memory_utilization, quantization=args.quantization)
        print(json.dumps(estimate(mcfg, ecfg, weights_on_disk=disk, kv_tokens=args.kv_tokens).to_dict(), indent=1))
        return 0
    if args.cmd == "_worker":
        import json

        from .config import EngineConfig
        from .engine import TPWorker

        eng = make_engine(EngineConfig(**json.loads(args.engine_json)))
        TPWorker(eng).serve_forever()
        return 0

    if unknown:
        log.warning("ignoring unsupported engine flags: %s", " ".join(unknown))
    if args.dtype not in ("auto", "bfloat16", "bf16"):
        log.error("only bfloat16 is supported (got --dtype %s)", args.dtype)
        return 2
    import torch

    if not torch.cuda.is_available():
        log.error("no CUDA device visible: this engine has no CPU path")
        return 2
    tp = torch.cuda.device_count() if str(args.tensor_parallel_size) == "auto" else int(args.tensor_parallel_size)
    ecfg = _engine_config(args, tp)
    workers = _spawn_workers(ecfg, tp) if tp > 1 and "RANK" not in os.environ else []
    try:
        from .frontend import spawn_frontend

        eng = make_engine(ecfg)
        name = ecfg.served_model_name or os.path.basename(os.path.normpath(args.model))
        model_path = None if args.model.startswith("random:") else args.model
        # this (GPU-owning) process runs the engine-core busy loop; HTTP lives in a child process
        front, core = spawn_frontend(eng, args.host, args.port, name, model_path, args.log_level,
                                     stream_interval=max(1, args.stream_interval))
        log.info("serving %s on %s:%d (tp=%d, %d KV blocks, max_model_len=%d)", name, args.host, args.port, tp,
                 eng.num_blocks, ecfg.max_model_len)
        try:
            core.run()                      # returns when the frontend goes away / asks to stop
        finally:
            eng.shutdown()
            front.terminate()
    finally:
        for p in workers:
            p.terminate()
    return 0


if __name__ == "__main__":
    sys.exit(main())
