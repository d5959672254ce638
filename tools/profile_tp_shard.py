"""Single-GPU stand-in for ONE rank of a TP=t Llama-3-8B decode step (same per-rank shapes, no
communication): shows where a tensor-parallel decode step spends its GPU time.
  ncu --profile-from-start off --metrics gpu__time_duration.sum --clock-control none --csv \
      --log-file gpurun_out/launches_tp8shard.csv python tools/profile_tp_shard.py 8"""
import os
import sys
import time
from pathlib import Path

sys.path.insert(0, str(Path(__file__).resolve().parents[1]))
import numpy as np
import torch

from gpustack_b200.config import EngineConfig, ModelConfig
from gpustack_b200.engine import LLMEngine
from gpustack_b200.scheduler import SamplingParams

t = int(sys.argv[1]) if len(sys.argv) > 1 else 8
R = 256
mcfg = ModelConfig(vocab_size=(128256 // t) // 8 * 8, hidden_size=4096, intermediate_size=14336 // t,
                   num_hidden_layers=32, num_attention_heads=32 // t, num_key_value_heads=8 // t, head_dim=128,
                   max_position_embeddings=8192)
GRAPH = os.environ.get("GRAPH", "0") == "1"      # GRAPH=1: CUDA-graph decode steps, wall-clock ms/step printed
eng = LLMEngine.from_engine_config(EngineConfig(max_num_seqs=R, max_num_batched_tokens=8192, max_model_len=1024,
                                                enforce_eager=not GRAPH, num_gpu_blocks=R * 48 + 64, mixed_steps=False,
                                                cuda_graph_sizes=[R]), mcfg=mcfg)
rs = np.random.RandomState(0)
for i, p in enumerate(rs.randint(256, mcfg.vocab_size - 256, size=(R, 512)).tolist()):
    eng.add_request(f"r{i}", p, SamplingParams(max_tokens=40, ignore_eos=True))
for _ in range((R * 512 + 8191) // 8192):
    eng.step()
eng.step(); eng.step()
torch.cuda.synchronize()
torch.cuda.profiler.start()
eng.step()
torch.cuda.synchronize()
torch.cuda.profiler.stop()
t0 = time.perf_counter()
n = 0
while eng.has_work() and n < 30:
    eng.step(); n += 1
torch.cuda.synchronize()
print(f"TP={t} shard (no communication), B={R}: {1e3 * (time.perf_counter() - t0) / max(1, n):.3f} ms per decode step "
      f"({'graph' if GRAPH else 'eager'}, {n} steps)")
print("done", eng.runner.stats)
