"""Reference GPU baseline: the engine GPUStack's vLLM backend launches (vLLM, here 0.22.0) on the
same workload as bench.py -- Llama-3-8B shape, dummy weights (throughput only), token-id prompts,
R requests x 512-in / 128-out greedy, continuous batching, tensor parallel over --tp GPUs (the
`--tensor-parallel-size` GPUStack's launcher injects, gpustack/worker/backends/vllm.py:632-637).
Offline LLM API (engine core, no HTTP), to compare with bench.py's `value`.  Output: one JSON line
(last line of stdout).

  python tools/vllm_baseline.py [requests] [timed_steps] [--tp N] [--warmup W]
"""
import argparse
import json
import os
import re
import sys
import tempfile
import time
from pathlib import Path

sys.path.insert(0, str(Path(__file__).resolve().parents[1]))
import numpy as np

from gpustack_b200.config import llama3_8b


def main():
    ap = argparse.ArgumentParser()
    ap.add_argument("requests", nargs="?", type=int, default=256)
    ap.add_argument("steps", nargs="?", type=int, default=2)
    ap.add_argument("--tp", type=int, default=1)
    ap.add_argument("--warmup", type=int, default=1)
    ap.add_argument("--input-len", type=int, default=512)
    ap.add_argument("--output-len", type=int, default=128)
    a = ap.parse_args()
    R, steps, in_len, out_len = a.requests, a.steps, a.input_len, a.output_len
    mcfg = llama3_8b()
    d = tempfile.mkdtemp(prefix="llama3-8b-dummy-")
    Path(d, "config.json").write_text(json.dumps(mcfg.to_hf_dict()))
    import vllm
    from vllm import LLM, SamplingParams

    t0 = time.time()
    llm = LLM(model=d, skip_tokenizer_init=True, load_format="dummy", dtype="bfloat16", max_model_len=1024,
              max_num_seqs=R, max_num_batched_tokens=8192, gpu_memory_utilization=0.90,
              enable_prefix_caching=False, seed=0, tensor_parallel_size=a.tp)
    init_s = time.time() - t0
    sp = SamplingParams(max_tokens=out_len, ignore_eos=True, temperature=0.0, detokenize=False)
    res = []
    for k in range(a.warmup + steps):
        rs = np.random.RandomState(1000 + k)
        prompts = [{"prompt_token_ids": p} for p in rs.randint(256, mcfg.vocab_size - 256, size=(R, in_len)).tolist()]
        t = time.time()
        outs = llm.generate(prompts, sp, use_tqdm=False)
        dt = time.time() - t
        n = sum(len(o.outputs[0].token_ids) for o in outs)
        res.append({"step": k, "seconds": dt, "output_tok_s": n / dt})
        print(f"[vllm] step {k}: {dt:.2f}s {n / dt:.0f} tok/s", file=sys.stderr, flush=True)
    timed = res[a.warmup:]
    tot_s = sum(r["seconds"] for r in timed)
    info = {"attention_backend": os.environ.get("VLLM_ATTENTION_BACKEND"), "allreduce_path": None}
    try:   # what the engine actually selected (private attributes: best effort, the log has it too)
        cc = llm.llm_engine.vllm_config
        info["cudagraph_mode"] = str(getattr(cc.compilation_config, "cudagraph_mode", None))
    except Exception:  # noqa: BLE001
        pass
    print(json.dumps({"engine": "vllm", "version": vllm.__version__, "tensor_parallel_size": a.tp, "requests": R,
                      "input_len": in_len, "output_len": out_len, "init_s": init_s,
                      "output_tok_s": R * out_len * len(timed) / tot_s if tot_s else None,
                      "ms_per_step": 1e3 * tot_s / max(1, len(timed)), "timed_steps": len(timed), "warmup": a.warmup,
                      "steps": res, "weights": "dummy", "path": "offline LLM.generate (engine core, no HTTP)", **info}),
          flush=True)


def parse_backend_log(text: str) -> dict:
    """Attention backend / all-reduce path lines of a vLLM log (used by bench.py's gpu_baseline)."""
    out = {}
    m = re.search(r"Using (\w+) attention backend", text)
    if m:
        out["attention_backend"] = m.group(1)
    if "TRTLLM prefill attention" in text:
        out["prefill_attention"] = "TRTLLM-gen (FlashInfer)"
    ar = []
    for pat, name in ((r"[Cc]ustom all-?reduce", "custom_all_reduce"), (r"symm_mem|SymmMem", "torch_symm_mem"),
                      (r"[Ff]lash[Ii]nfer all-?reduce|flashinfer.*allreduce", "flashinfer_allreduce"),
                      (r"pynccl|PyNccl", "pynccl")):
        if re.search(pat, text):
            ar.append(name)
    if ar:
        out["allreduce_path"] = "+".join(ar)
    return out


if __name__ == "__main__":
    main()
