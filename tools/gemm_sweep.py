"""Sweep tile width / split-K of the tcgen05 GEMM over the Llama-3-8B projection shapes.
Weights rotate through enough copies to exceed L2 (126 MB) between timed launches."""
import json
import sys
from pathlib import Path

sys.path.insert(0, str(Path(__file__).resolve().parents[1]))
import torch

from gpustack_b200 import ops

SHAPES = {"qkv": (6144, 4096), "o": (4096, 4096), "gate_up": (28672, 4096), "down": (4096, 14336)}
Ms = [int(x) for x in (sys.argv[1].split(",") if len(sys.argv) > 1 else "16,64,128,256".split(","))]
res = []
for name, (N, K) in SHAPES.items():
    ncopy = max(2, int(400e6 // (N * K * 2)) + 1)
    Ws = [(torch.randn(N, K, device="cuda") * 0.02).bfloat16() for _ in range(ncopy)]
    for M in Ms:
        x = torch.randn(M, K, device="cuda").bfloat16()
        ref_t = None
        for bx in (16, 32, 64, 128, 256):
            if bx * 8 < M or (bx > 16 and bx >= 2 * M and bx > 16 * ((M + 15) // 16)):
                continue
            for splits in (1, 2, 4, 8):
                if (K // 64) % splits or (K // 64) // splits < 4:
                    continue
                out = torch.empty((splits, M, N) if splits > 1 else (M, N), device="cuda",
                                  dtype=torch.float32 if splits > 1 else torch.bfloat16)
                for i in range(3):
                    ops.gemm(x, Ws[i % ncopy], out=out, splits=splits, block_x=bx)
                torch.cuda.synchronize()
                iters = 20
                e0, e1 = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
                e0.record()
                for i in range(iters):
                    ops.gemm(x, Ws[i % ncopy], out=out, splits=splits, block_x=bx)
                e1.record()
                torch.cuda.synchronize()
                us = e0.elapsed_time(e1) / iters * 1e3
                gbs = (N * K * 2) / (us * 1e-6) / 1e9
                tf = 2.0 * M * N * K / (us * 1e-6) / 1e12
                res.append({"shape": name, "M": M, "block_x": bx, "splits": splits, "us": round(us, 2),
                            "w_gbs": round(gbs), "tflops": round(tf, 1)})
        # cuBLAS reference (library bar, not used by the engine)
        for i in range(3):
            torch.matmul(x, Ws[i % ncopy].t())
        torch.cuda.synchronize()
        e0, e1 = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
        e0.record()
        for i in range(20):
            torch.matmul(x, Ws[i % ncopy].t())
        e1.record()
        torch.cuda.synchronize()
        us = e0.elapsed_time(e1) / 20 * 1e3
        res.append({"shape": name, "M": M, "block_x": "cublas", "splits": 0, "us": round(us, 2),
                    "w_gbs": round(N * K * 2 / (us * 1e-6) / 1e9), "tflops": round(2.0 * M * N * K / (us * 1e-6) / 1e12, 1)})
    del Ws
for r in res:
    print(json.dumps(r))
