"""Tensor-parallel parity check, launched with torch.distributed.run (one rank per GPU) by
tests/test_tp_gpu.py:  the TP engine must reproduce the CPU oracle's logits (stated tolerance
2e-2 * max|logit|) and the single-GPU engine's greedy tokens.  TP_FAMILY selects the model family:
dense (default), fp8 (W8A8; every rank quantises its own shard, tolerance 0.12 as in
tests/test_engine_gpu.py) or mixtral (experts tensor-parallel over I, tolerance 5e-2)."""
import json
import os
import sys
from pathlib import Path

sys.path.insert(0, str(Path(__file__).resolve().parents[1]))
import torch

from gpustack_b200.config import EngineConfig, ModelConfig
from gpustack_b200.engine import LLMEngine, TPWorker
from gpustack_b200.scheduler import SamplingParams
from oracle import llama_oracle as O


def main():
    world = int(os.environ["WORLD_SIZE"])
    rank = int(os.environ["RANK"])
    torch.cuda.set_device(int(os.environ["LOCAL_RANK"]))
    family = os.environ.get("TP_FAMILY", "dense")
    tol = {"dense": 2e-2, "fp8": 0.12, "mixtral": 5e-2}[family]
    moe = dict(num_local_experts=8, num_experts_per_tok=2) if family == "mixtral" else {}
    c = O.LlamaConfig(vocab_size=1000, hidden_size=512, intermediate_size=256 * world if family == "mixtral" else 1024,
                      num_hidden_layers=2, num_attention_heads=8, num_key_value_heads=world if world <= 8 else 8,
                      head_dim=64, max_position_embeddings=512, **moe)
    if c.num_attention_heads % c.num_key_value_heads:
        c.num_attention_heads = c.num_key_value_heads * 2
    mcfg = ModelConfig(vocab_size=c.vocab_size, hidden_size=c.hidden_size, intermediate_size=c.intermediate_size,
                       num_hidden_layers=c.num_hidden_layers, num_attention_heads=c.num_attention_heads,
                       num_key_value_heads=c.num_key_value_heads, head_dim=c.head_dim, rope_theta=c.rope_theta,
                       max_position_embeddings=c.max_position_embeddings, eos_token_id=2,
                       num_local_experts=c.num_local_experts, num_experts_per_tok=c.num_experts_per_tok,
                       architecture="MixtralForCausalLM" if c.num_local_experts else "LlamaForCausalLM")
    weights = O.init_weights(c, seed=21)
    features = os.environ.get("TP_FEATURES", "0") == "1"     # speculative decoding + host KV tier under TP
    extra = {}
    if features:
        extra = dict(speculative_config={"method": "ngram", "num_speculative_tokens": 3, "prompt_lookup_max": 3,
                                         "prompt_lookup_min": 1}, kv_host_cache_gib=1.0 / 64, enable_prefix_caching=True)
    ecfg = EngineConfig(tensor_parallel_size=world, max_num_seqs=8, max_num_batched_tokens=128,
                        num_gpu_blocks=40 if features else 128,
                        cuda_graph_sizes=[1, 2, 4, 8], enforce_eager=os.environ.get("TP_EAGER", "0") == "1",
                        quantization="fp8" if family == "fp8" else None, **extra)
    eng = LLMEngine.from_engine_config(ecfg, mcfg=mcfg, state_dict={k: v.cuda() for k, v in weights.items()})
    if rank != 0:
        TPWorker(eng).serve_forever()
        return
    g = torch.Generator().manual_seed(4)
    prompts = [torch.randint(3, c.vocab_size, (n,), generator=g).tolist() for n in (9, 40, 77, 130)]
    if features:
        prompts[1] = (prompts[1][:5] * 8)[:40]          # repetitive: the n-gram proposer drafts from the first step on
    oracle = O.OracleLlama(c, weights, quantization=ecfg.quantization)
    res = {"world": world, "family": family, "ok": True, "checks": []}
    outs = eng.generate(prompts, SamplingParams(max_tokens=12, ignore_eos=True))
    for i, (p, o) in enumerate(zip(prompts, outs)):
        oracle.reset()
        logits = oracle.forward(i, p)[-1]
        n_ok = n_chk = 0
        for t in range(len(o)):
            top2 = torch.topk(logits, 2).values
            if float(top2[0] - top2[1]) > tol * float(logits.abs().max()):
                n_chk += 1
                n_ok += int(int(torch.argmax(logits)) == o[t])
            logits = oracle.forward(i, [o[t]])[-1]
        res["checks"].append({"seq": i, "checked": n_chk, "agree": n_ok})
        if n_ok != n_chk or n_chk < (3 if family == "dense" else 1):
            res["ok"] = False
    if features:
        # a second batch evicts the first batch's cached blocks into the host tier; the first batch, asked again,
        # is served from host memory (swap lists replayed by every rank) and must produce oracle-consistent tokens
        others = [torch.randint(3, c.vocab_size, (n,), generator=g).tolist() for n in (150, 160, 140)]
        eng.generate(others, SamplingParams(max_tokens=4, ignore_eos=True))
        again = eng.generate(prompts, SamplingParams(max_tokens=12, ignore_eos=True))
        n_ok = n_chk = 0
        for i, (p, o) in enumerate(zip(prompts, again)):
            oracle.reset()
            logits = oracle.forward(i, p)[-1]
            for t in range(len(o)):
                top2 = torch.topk(logits, 2).values
                if float(top2[0] - top2[1]) > tol * float(logits.abs().max()):
                    n_chk += 1
                    n_ok += int(int(torch.argmax(logits)) == o[t])
                logits = oracle.forward(i, [o[t]])[-1]
        st = eng.sched.spec_stats
        res["features"] = {"spec_steps": st["steps"], "drafted": st["drafted"], "accepted": st["accepted"],
                           "host_stored": eng.sched.pool.host_stored, "host_hit_tokens": eng.sched.pool.host_hits,
                           "swap_in_blocks": eng.runner.stats.get("kv_swap_in_blocks", 0),
                           "again_checked": n_chk, "again_agree": n_ok}
        if not (st["steps"] > 0 and st["drafted"] > 0 and eng.sched.pool.host_hits > 0 and n_ok == n_chk and n_chk >= 8):
            res["ok"] = False
    # sampled requests across vocab shards: the in-kernel Gumbel-max draw is keyed by (seed, position,
    # GLOBAL vocab index), so a seed reproduces its tokens, another seed does not, and nucleus rows
    # (top_p < 1) go through the all-rank logits gather; top_p -> 0 must collapse onto greedy
    sp = dict(max_tokens=12, ignore_eos=True)
    a = eng.generate(prompts[:2], SamplingParams(temperature=1.3, seed=11, **sp))
    b = eng.generate(prompts[:2], SamplingParams(temperature=1.3, seed=11, **sp))
    d = eng.generate(prompts[:2], SamplingParams(temperature=1.3, seed=12, **sp))
    n = eng.generate(prompts[:2], SamplingParams(temperature=1.0, top_p=1e-6, seed=3, **sp))
    nucleus_ok = n == [o[:12] for o in outs[:2]]
    if features and not nucleus_ok:
        # `outs` came out of k+1-token verification forwards, `n` out of 1-token decode steps: a near-tie may
        # resolve differently; the statement that holds is the margin rule against the oracle
        n_ok = n_chk = 0
        for i, (p, o) in enumerate(zip(prompts[:2], n)):
            oracle.reset()
            logits = oracle.forward(i, p)[-1]
            for t in range(len(o)):
                top2 = torch.topk(logits, 2).values
                if float(top2[0] - top2[1]) > tol * float(logits.abs().max()):
                    n_chk += 1
                    n_ok += int(int(torch.argmax(logits)) == o[t])
                logits = oracle.forward(i, [o[t]])[-1]
        nucleus_ok = n_ok == n_chk and n_chk >= 4
    res["sampled"] = {"reproducible": a == b, "seed_matters": a != d, "differs_from_greedy": a != [o[:12] for o in outs[:2]],
                      "nucleus_top_p_0_is_greedy": nucleus_ok}
    if not all(res["sampled"].values()):
        res["ok"] = False
    eng.shutdown()
    print("TP_PARITY " + json.dumps(res), flush=True)
    if not res["ok"]:
        sys.exit(1)


if __name__ == "__main__":
    main()
    import torch.distributed as dist

    if dist.is_initialized():
        dist.destroy_process_group()
