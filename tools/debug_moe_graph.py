"""Debug helper: tiny Mixtral engine, decode under several launch modes, compare against the eager
run (tokens + first-decode-step logits).  Usage: python tools/debug_moe_graph.py [mode]
modes run as sub-processes (B200_PDL is read once per process)."""
import json
import os
import subprocess
import sys
from pathlib import Path

ROOT = Path(__file__).resolve().parents[1]
sys.path.insert(0, str(ROOT))


def run(mode: str):
    import torch

    from gpustack_b200.config import EngineConfig, tiny_llama, tiny_mixtral
    from gpustack_b200.engine import LLMEngine
    from gpustack_b200.scheduler import SamplingParams
    from gpustack_b200.weights import random_state_dict

    dense = mode.startswith("dense-")
    mcfg = tiny_llama(64) if dense else tiny_mixtral()
    mode = mode[6:] if dense else mode
    sd = random_state_dict(mcfg, seed=7, device="cuda")
    kw = dict(max_num_seqs=16, max_num_batched_tokens=256, num_gpu_blocks=128)
    if mode.startswith("eager"):
        kw["enforce_eager"] = True
    elif mode.startswith("graph5"):
        kw["cuda_graph_sizes"] = [5]
    else:
        kw["cuda_graph_sizes"] = [1, 2, 4, 8]
    if mode.startswith("eager-pad8"):
        kw["enforce_eager"] = True
    eng = LLMEngine.from_engine_config(EngineConfig(**kw), mcfg=mcfg, state_dict=sd)
    if mode.startswith("eager-pad8"):
        runner = eng.runner

        class FakeGraph:          # eager execution of the PADDED batch (bucket 8), no CUDA graph
            def replay(self):
                runner._decode_forward(8)

        runner.graphs = {8: FakeGraph()}
        runner._buckets = [8]
    g = torch.Generator().manual_seed(11)
    prompts = [torch.randint(3, mcfg.vocab_size, (n,), generator=g).tolist() for n in (5, 23, 64, 100, 17)]
    sp = SamplingParams(max_tokens=24 if dense else 6, ignore_eos=True)
    for i, p in enumerate(prompts):
        eng.add_request(f"r{i}", p, sp)
    toks = {f"r{i}": [] for i in range(5)}
    logits1 = None
    diag = []
    step = 0
    while eng.sched.has_work():
        out = eng.step()
        for rid, t in zip(out.req_ids, out.token_ids):
            toks[rid].append(int(t))
        if step >= 1 and step <= 3:
            torch.cuda.synchronize()
            n = len(out.req_ids)
            dev_ids = eng.runner.out_ids[:n].cpu().tolist()            # device copy of the sampler output, after a full sync
            host_arg = eng.runner.last_logits(n).float().cpu().argmax(dim=-1).tolist()
            diag.append({"step": step, "returned": [int(t) for t in out.token_ids], "out_ids_after_sync": dev_ids,
                         "argmax_of_logits": host_arg})
        if step == 1:
            lg = eng.runner.last_logits(5).float().cpu()
            logits1 = [[round(float(x), 4) for x in row[:6]] + [float(row.abs().max()), bool(torch.isnan(row).any())]
                       for row in lg]
        step += 1
    print("RESULT " + json.dumps({"mode": ("dense-" if dense else "") + mode, "tokens": toks, "logits_step1": logits1, "diag": diag}))


if __name__ == "__main__":
    if len(sys.argv) > 1:
        run(sys.argv[1])
        sys.exit(0)
    res = {}
    # the default build triggers dependents late (after griddepcontrol.wait); the early-trigger
    # experiment is a separate library: python -m gpustack_b200.build --early
    early = str(ROOT / "gpustack_b200" / "csrc" / "libb200_early.so")
    for mode, env in (("eager", {}), ("graph8-all-pdl", {"B200_PDL_OFF_CLASSES": "0"}),
                      ("graph8-all-pdl-early-trigger", {"B200_PDL_OFF_CLASSES": "0", "B200_LIB": early}),
                      ("dense-eager", {}), ("dense-graph8-early-trigger", {"B200_LIB": early})):
        e = dict(os.environ, **env)
        r = subprocess.run([sys.executable, __file__, mode], capture_output=True, text=True, env=e, timeout=300)
        line = [l for l in r.stdout.splitlines() if l.startswith("RESULT ")]
        if not line:
            print(mode, "FAILED", r.stderr[-800:])
            continue
        res[mode] = json.loads(line[0][7:])
    for mode, r in res.items():
        ref = res.get("dense-eager" if mode.startswith("dense-") else "eager")
        same = r["tokens"] == ref["tokens"] if ref else None
        print(mode, "tokens_equal_eager=", same)
        if not same:
            print("   tokens", r["tokens"])
        for d in r.get("diag", []):
            print("   ", d)
