"""Prefix-caching benefit on the Llama-3-8B shape: R requests that share a long system prompt
(P tokens) and differ in a short suffix (S tokens), O output tokens each, all submitted at once.
The same engine runs the workload with the block pool's prefix cache disabled and enabled
(cold: the cache is reset first, so the shared prefix is computed once and re-attached by every
later admission).  One JSON line."""
import json
import sys
import time
from pathlib import Path

sys.path.insert(0, str(Path(__file__).resolve().parents[1]))
import numpy as np
import torch

from gpustack_b200.config import EngineConfig, llama3_8b
from gpustack_b200.engine import LLMEngine
from gpustack_b200.scheduler import SamplingParams


def main():
    R, P, S, Out = 128, 1536, 256, 64
    mcfg = llama3_8b()
    eng = LLMEngine.from_engine_config(EngineConfig(model="random:llama3-8b", max_num_seqs=R, max_model_len=4096,
                                                    enable_prefix_caching=True, seed=0), mcfg=mcfg)
    rs = np.random.RandomState(5)
    sp = SamplingParams(max_tokens=Out, ignore_eos=True)

    def workload(k):
        head = rs.randint(256, mcfg.vocab_size - 256, size=P).tolist()
        return [head + rs.randint(256, mcfg.vocab_size - 256, size=S).tolist() for _ in range(R)]

    res = {}
    eng.generate(workload(0)[:16], sp)                       # warm-up (autotuned paths, allocator)
    for mode in ("off", "on"):
        pool = eng.sched.pool
        pool.caching = mode == "on"
        pool.reset_prefix_cache()
        pool.queries = pool.hits = 0
        st0 = dict(eng.runner.stats)
        torch.cuda.synchronize()
        t0 = time.perf_counter()
        outs = eng.generate(workload(1), sp)
        torch.cuda.synchronize()
        dt = time.perf_counter() - t0
        st1 = eng.runner.stats
        res[mode] = {"seconds": round(dt, 3), "output_tok_s": round(sum(len(o) for o in outs) / dt, 1),
                     "prefill_tokens_computed": st1["prefill_tokens"] - st0["prefill_tokens"],
                     "prefill_steps": st1["prefill_steps"] - st0["prefill_steps"],
                     "prefix_hit_tokens": pool.hits, "prefix_query_tokens": pool.queries}
    print(json.dumps({"workload": f"{R} requests = shared {P}-token prefix + {S}-token suffix, {Out} out, Llama-3-8B TP=1",
                      "prefix_caching": res,
                      "speedup": round(res["off"]["seconds"] / res["on"]["seconds"], 2)}))


if __name__ == "__main__":
    main()
