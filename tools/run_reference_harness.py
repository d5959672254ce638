#!/usr/bin/env python
"""Runs the reference's own load generator, benchmarks/benchmark_serving.py, UNMODIFIED (runpy, its
own __main__ and CLI) against any OpenAI-compatible server -- `b200-engine serve` or `vllm serve`.

  python tools/run_reference_harness.py -m <model> -n 512 -c 256 --max-completion-tokens 128 \
         --ignore-eos --prompt-multiplier 4 --server-url http://127.0.0.1:8000 --json

The script is looked up at /root/reference/benchmarks/benchmark_serving.py (build container) or
baseline/_ref/benchmarks/benchmark_serving.py (an untracked, unmodified copy that
`__graft_entry__.build()` places there so that it travels to the GPU box; reference sources are
never committed).  Two things it needs are supplied from outside, without touching it: the missing
third-party `httpx_aiohttp` module (tools/shims/) and a seed for its unseeded `random` prompts
(`random.seed(0)` via a wrapper).
"""
import os
import random
import runpy
import sys
from pathlib import Path

ROOT = Path(__file__).resolve().parents[1]
CANDIDATES = [Path("/root/reference/benchmarks/benchmark_serving.py"),
              ROOT / "baseline" / "_ref" / "benchmarks" / "benchmark_serving.py"]


def harness_path():
    for p in CANDIDATES:
        if p.exists():
            return p
    return None


def main():
    path = harness_path()
    if path is None:
        sys.stderr.write("benchmark_serving.py not found (looked in: %s)\n" % ", ".join(map(str, CANDIDATES)))
        return 3
    try:
        import httpx_aiohttp  # noqa: F401  (the real package, if it is ever installed)
    except ImportError:
        sys.path.insert(0, str(ROOT / "tools" / "shims"))
    random.seed(int(os.environ.get("HARNESS_SEED", "0")))
    sys.argv = [str(path)] + sys.argv[1:]
    runpy.run_path(str(path), run_name="__main__")
    return 0


if __name__ == "__main__":
    sys.exit(main())
