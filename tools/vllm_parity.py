#!/usr/bin/env python
"""Live-engine parity, reference side (SURVEY.md section 8c protocol iv): the engine GPUStack's vLLM
backend launches -- vLLM, whatever version is installed -- scores given token sequences from the SAME
safetensors directory this repo's engine served.  For every sequence (prompt + the tokens OUR engine
generated) vLLM returns, for each generated position, its top-k next-token log-probabilities given
the true prefix (`prompt_logprobs`): a teacher-forced comparison, so one near-tie cannot derail the
rest of the sequence.

  python tools/vllm_parity.py <model_dir> <in.json> <out.json> [gpu_memory_utilization]
in.json : {"sequences": [[token ids...], ...], "n_prompt": [int, ...]}
out.json: {"vllm": version, "top": [[[[token, logprob], ...k] per generated position] per sequence]}
"""
import json
import sys


def main():
    model_dir, fin, fout = sys.argv[1:4]
    util = float(sys.argv[4]) if len(sys.argv) > 4 else 0.30
    job = json.load(open(fin))
    import vllm
    from vllm import LLM, SamplingParams

    llm = LLM(model=model_dir, skip_tokenizer_init=True, dtype="bfloat16", max_model_len=2048, enforce_eager=True,
              gpu_memory_utilization=util, enable_prefix_caching=False, seed=0)
    sp = SamplingParams(max_tokens=1, temperature=0.0, prompt_logprobs=5, detokenize=False)
    outs = llm.generate([{"prompt_token_ids": s} for s in job["sequences"]], sp, use_tqdm=False)
    top = []
    for o, n_prompt in zip(outs, job["n_prompt"]):
        rows = []
        for pos in range(n_prompt, len(o.prompt_token_ids)):
            lp = o.prompt_logprobs[pos]      # distribution of token `pos` given tokens [0, pos)
            rows.append(sorted(([int(t), float(v.logprob)] for t, v in lp.items()), key=lambda x: -x[1]))
        top.append(rows)
    json.dump({"vllm": vllm.__version__, "top": top}, open(fout, "w"))


if __name__ == "__main__":
    main()
