"""Reference GPU baseline: the engine GPUStack's vLLM backend launches (vLLM, here 0.22.0) on the
same workload as bench.py -- Llama-3-8B shape, dummy weights (throughput only), token-id prompts,
256 requests x 512-in / 128-out greedy, continuous batching.  Offline LLM API (engine core, no
HTTP), to compare with bench.py's `value`.  Output: one JSON line."""
import json
import sys
import tempfile
import time
from pathlib import Path

sys.path.insert(0, str(Path(__file__).resolve().parents[1]))
import numpy as np

from gpustack_b200.config import llama3_8b


def main():
    R = int(sys.argv[1]) if len(sys.argv) > 1 else 256
    steps = int(sys.argv[2]) if len(sys.argv) > 2 else 2
    in_len, out_len = 512, 128
    mcfg = llama3_8b()
    d = tempfile.mkdtemp(prefix="llama3-8b-dummy-")
    Path(d, "config.json").write_text(json.dumps(mcfg.to_hf_dict()))
    import vllm
    from vllm import LLM, SamplingParams

    t0 = time.time()
    llm = LLM(model=d, skip_tokenizer_init=True, load_format="dummy", dtype="bfloat16", max_model_len=1024,
              max_num_seqs=R, max_num_batched_tokens=8192, gpu_memory_utilization=0.90,
              enable_prefix_caching=False, seed=0)
    init_s = time.time() - t0
    sp = SamplingParams(max_tokens=out_len, ignore_eos=True, temperature=0.0, detokenize=False)
    res = []
    for k in range(1 + steps):
        rs = np.random.RandomState(1000 + k)
        prompts = [{"prompt_token_ids": p} for p in rs.randint(256, mcfg.vocab_size - 256, size=(R, in_len)).tolist()]
        t = time.time()
        outs = llm.generate(prompts, sp, use_tqdm=False)
        dt = time.time() - t
        n = sum(len(o.outputs[0].token_ids) for o in outs)
        res.append({"step": k, "seconds": dt, "output_tok_s": n / dt})
        print(f"[vllm] step {k}: {dt:.2f}s {n / dt:.0f} tok/s", file=sys.stderr, flush=True)
    timed = res[1:]
    print(json.dumps({"engine": f"vllm {vllm.__version__}", "requests": R, "input_len": in_len, "output_len": out_len,
                      "init_s": init_s, "output_tok_s": sum(r["output_tok_s"] for r in timed) / len(timed),
                      "ms_per_step": 1e3 * sum(r["seconds"] for r in timed) / len(timed), "steps": res,
                      "weights": "dummy", "path": "offline LLM.generate (engine core, no HTTP)"}))


if __name__ == "__main__":
    main()
