"""All-reduce micro-benchmark under torch.distributed.run: latency of the fused one-shot NVLink
all-reduce+add+RMSNorm kernel vs NCCL all_reduce + our add+RMSNorm kernel for a decode-sized
[T,4096] bf16 message, 64 calls captured in one CUDA graph (as in a decode step)."""
import json
import os
import sys
from pathlib import Path

sys.path.insert(0, str(Path(__file__).resolve().parents[1]))
import torch
import torch.distributed as dist

from gpustack_b200 import ops
from gpustack_b200.engine import TPContext
from gpustack_b200.tp_comm import OneShotAllReduce

tp = TPContext.from_env()
dev = tp.device
H = 4096
ar = OneShotAllReduce(tp, 512, H, dev)
res = {}
for T in (16, 64, 256):
    x = torch.randn(T, H, device=dev).bfloat16()
    resid = torch.randn(T, H, device=dev).bfloat16()
    w = torch.ones(H, device=dev).bfloat16()
    y = torch.empty_like(x); ro = torch.empty_like(x)

    def oneshot():
        for i in range(64):
            ar.stage(i & 1, T).copy_(x)
            ar.allreduce_add_rms_norm(i & 1, resid, w, y, ro, T, 1e-5)

    def nccl():
        for i in range(64):
            h = x.clone()
            dist.all_reduce(h, group=tp.nccl)
            ops.fused_add_rms_norm(h, resid, w, 1e-5, out=y, residual_out=ro)

    for name, fn in (("oneshot", oneshot), ("nccl", nccl)):
        s = torch.cuda.Stream()
        with torch.cuda.stream(s):
            fn(); torch.cuda.synchronize()
            g = torch.cuda.CUDAGraph()
            with torch.cuda.graph(g, stream=s):
                fn()
        torch.cuda.synchronize()
        dist.barrier(group=tp.cpu)
        for _ in range(3):
            g.replay()
        torch.cuda.synchronize()
        e0, e1 = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
        e0.record()
        for _ in range(10):
            g.replay()
        e1.record()
        torch.cuda.synchronize()
        res[f"{name}_T{T}_us_per_call"] = round(e0.elapsed_time(e1) / 10 / 64 * 1e3, 2)
if tp.rank == 0:
    print("AR_BENCH " + json.dumps({"world": tp.world, **res}), flush=True)
torch.cuda.synchronize()
os._exit(0)
