"""Decode-step time of the Llama-3-8B engine vs the HBM roofline at several batch sizes / contexts
(CUDA-graph replays, device-timed).  One JSON line per point."""
import json
import sys
from pathlib import Path

sys.path.insert(0, str(Path(__file__).resolve().parents[1]))
import numpy as np
import torch

from gpustack_b200.config import EngineConfig, llama3_8b
from gpustack_b200.engine import LLMEngine
from gpustack_b200.scheduler import SamplingParams

peak = 6572.9
p = Path(__file__).resolve().parents[1] / "MEASURED_PEAKS.json"
if p.exists():
    peak = json.loads(p.read_text())["hbm_gbs"]
mcfg = llama3_8b()
eng = LLMEngine.from_engine_config(EngineConfig(max_num_seqs=256, max_num_batched_tokens=8192, max_model_len=4096),
                                   mcfg=mcfg)
rs = np.random.RandomState(0)
for B, ctx in [(1, 512), (8, 512), (32, 512), (64, 512), (128, 512), (256, 512), (64, 3072), (256, 2048)]:
    prompts = rs.randint(256, mcfg.vocab_size - 256, size=(B, ctx)).tolist()
    n_dec = 40
    for i, pr in enumerate(prompts):
        eng.add_request(f"b{B}c{ctx}-{i}", pr, SamplingParams(max_tokens=n_dec + 1, ignore_eos=True))
    while eng.sched.prefilling or eng.sched.waiting:
        eng.step()
    for _ in range(5):
        eng.step()
    torch.cuda.synchronize()
    e0, e1 = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
    e0.record()
    n = 0
    while eng.has_work() and n < 30:
        eng.step(); n += 1
    e1.record()
    torch.cuda.synchronize()
    ms = e0.elapsed_time(e1) / n
    while eng.has_work():
        eng.step()
    mean_ctx = ctx + 5 + n / 2
    nbytes = mcfg.weight_bytes(1) + B * mean_ctx * mcfg.kv_bytes_per_token(1)
    print(json.dumps({"batch": B, "ctx": ctx, "ms_per_step": round(ms, 3), "tok_s": round(B / ms * 1e3),
                      "algorithmic_GB": round(nbytes / 1e9, 2), "achieved_GBs": round(nbytes / ms / 1e6),
                      "hbm_frac": round(nbytes / ms / 1e6 / peak, 3)}), flush=True)
