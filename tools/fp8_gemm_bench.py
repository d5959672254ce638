"""fp8 (e4m3 W8A8) vs bf16 tcgen05 GEMM on the Llama-3-8B / 70B-TP8 projection shapes: CUDA-event
time per launch, TFLOP/s and weight GB/s, plus the cost of the per-token quantiser.  Weights rotate
through enough copies to exceed L2 (126 MB) between timed launches.  One JSON line per case."""
import json
import sys
from pathlib import Path

sys.path.insert(0, str(Path(__file__).resolve().parents[1]))
import torch

from gpustack_b200 import ops

SHAPES = {"8b.qkv": (6144, 4096), "8b.o": (4096, 4096), "8b.gate_up": (28672, 4096), "8b.down": (4096, 14336),
          "70b-tp8.gate_up": (7168, 8192), "70b-tp8.down": (8192, 3584)}
Ms = [int(x) for x in (sys.argv[1].split(",") if len(sys.argv) > 1 else "64,256,8192".split(","))]
ITERS = 20


def timed(fn):
    for i in range(3):
        fn(i)
    torch.cuda.synchronize()
    e0, e1 = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
    e0.record()
    for i in range(ITERS):
        fn(i)
    e1.record()
    torch.cuda.synchronize()
    return e0.elapsed_time(e1) / ITERS * 1e3


for name, (N, K) in SHAPES.items():
    silu = name.endswith("gate_up")
    ncopy = max(2, int(400e6 // (N * K)) + 1)
    Wb = [(torch.randn(N, K, device="cuda") * 0.02).bfloat16() for _ in range(min(ncopy, 4))]
    W8 = [(torch.randn(N, K, device="cuda") * 100).clamp(-448, 448).to(torch.float8_e4m3fn) for _ in range(ncopy)]
    ws = torch.full((1,), 2e-4, device="cuda")
    for M in Ms:
        x = torch.randn(M, K, device="cuda").bfloat16()
        xq, xs = ops.quant_fp8_per_token(x)
        outw = N // 2 if silu else N
        out = torch.empty(M, outw, device="cuda", dtype=torch.bfloat16)
        if silu:
            t_bf = timed(lambda i: ops.gemm_silu_mul(x, Wb[i % len(Wb)], out=out))
        else:
            t_bf = timed(lambda i: ops.gemm(x, Wb[i % len(Wb)], out=out))
        t_f8 = timed(lambda i: ops.gemm_fp8(xq, xs, W8[i % ncopy], ws, out=out, silu_mul=silu))
        t_q = timed(lambda i: ops.quant_fp8_per_token(x, out=xq, scales=xs))
        flops = 2.0 * M * N * K
        print(json.dumps({"shape": name, "M": M, "N": N, "K": K, "silu_epilogue": silu,
                          "bf16_us": round(t_bf, 1), "fp8_us": round(t_f8, 1), "quant_us": round(t_q, 1),
                          "bf16_tflops": round(flops / t_bf / 1e6, 1), "fp8_tflops": round(flops / t_f8 / 1e6, 1),
                          "fp8_w_gbs": round(N * K / t_f8 / 1e3), "bf16_w_gbs": round(2 * N * K / t_bf / 1e3),
                          "quant_gbs": round(3.0 * M * K / t_q / 1e3)}), flush=True)
    del Wb, W8
