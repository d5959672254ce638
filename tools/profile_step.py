"""Profile one prefill step (16 x 512 tokens) and one decode step (B=256, ctx~514) of the
Llama-3-8B engine.  Meant to run under ncu with --profile-from-start off: only the region between
cudaProfilerStart/Stop is captured.

  ncu --profile-from-start off --metrics gpu__time_duration.sum --clock-control none --csv \
      --log-file gpurun_out/launches.csv python tools/profile_step.py
"""
import sys
from pathlib import Path

sys.path.insert(0, str(Path(__file__).resolve().parents[1]))
import numpy as np
import torch

from gpustack_b200.config import EngineConfig, llama3_8b
from gpustack_b200.engine import LLMEngine
from gpustack_b200.scheduler import SamplingParams

R = int(sys.argv[1]) if len(sys.argv) > 1 else 256
mcfg = llama3_8b()
eng = LLMEngine.from_engine_config(
    EngineConfig(max_num_seqs=R, max_num_batched_tokens=8192, max_model_len=1024, enforce_eager=True,
                 num_gpu_blocks=R * 48 + 64, mixed_steps=False), mcfg=mcfg)   # pure prefill / pure decode steps
rs = np.random.RandomState(0)
prompts = rs.randint(256, mcfg.vocab_size - 256, size=(R, 512)).tolist()
sp = SamplingParams(max_tokens=6, ignore_eos=True)
for i, p in enumerate(prompts):
    eng.add_request(f"r{i}", p, sp)
n_prefill = (R * 512 + 8191) // 8192
for i in range(n_prefill - 1):
    eng.step()
torch.cuda.synchronize()
torch.cuda.profiler.start()
eng.step()                      # last prefill step: 8192 tokens
torch.cuda.synchronize()
torch.cuda.profiler.stop()
eng.step(); eng.step()          # decode warm-up
torch.cuda.synchronize()
torch.cuda.profiler.start()
eng.step()                      # one decode step at B=R
torch.cuda.synchronize()
torch.cuda.profiler.stop()
while eng.has_work():
    eng.step()
print("profiled: 1 prefill step + 1 decode step; stats", eng.runner.stats)
