#!/usr/bin/env python
"""bench.py -- the reference's headline metric on the hot path: output tok/s (+ p50 TTFT) of
continuous-batched Llama-3-8B bf16 serving, 512-token prompts / 128 greedy output tokens
(BASELINE.json configs[1]; TP = --gpus), random-init weights, synthetic token-id prompts.

  python bench.py --gpus 1 --steps K --warmup W                      (this engine)
  python -m torch.distributed.run --nnodes=1 --nproc-per-node N --master-addr 127.0.0.1 \
         --master-port P bench.py --gpus N --steps K --warmup W     (tensor parallel, 1 rank/GPU)
  python bench.py --impl reference ...                               (CPU restatement on host cores)

A "step" = one pass of the workload: R requests (default 256) admitted at once, run to completion
by the continuous-batching engine.  `value` is measured on the engine core (device-timed, max over
ranks); `e2e` is the same metric through the reference-facing boundary: the OpenAI HTTP server
(/v1/completions, SSE streaming) driven by a separate client process with the TTFT / output-tok/s
definitions of the reference's benchmarks/benchmark_serving.py.
"""
from __future__ import annotations

import argparse
import json
import os
import socket
import statistics
import subprocess
import sys
import threading
import time
from pathlib import Path

ROOT = Path(__file__).resolve().parent
sys.path.insert(0, str(ROOT))

METRIC = "output_tok_per_s"
UNIT = "tok/s"
_T0 = time.time()


def log(msg):
    print(f"[bench +{time.time() - _T0:7.1f}s] {msg}", file=sys.stderr, flush=True)


def parse_args():
    ap = argparse.ArgumentParser()
    ap.add_argument("--gpus", type=int, default=1)
    ap.add_argument("--steps", type=int, default=3)
    ap.add_argument("--warmup", type=int, default=3)
    ap.add_argument("--impl", default="ours", choices=["ours", "reference"])
    ap.add_argument("--requests", type=int, default=256, help="requests per step (= concurrency)")
    ap.add_argument("--input-len", type=int, default=512)
    ap.add_argument("--output-len", type=int, default=128)
    ap.add_argument("--model", default="llama3-8b",
                    choices=["llama3-8b", "llama3-70b", "qwen2-0.5b", "tiny", "tiny128", "mixtral-8x7b", "tiny-mixtral"],
                    help="llama3-8b is the BASELINE config; mixtral-8x7b is a side measurement of the MoE family")
    ap.add_argument("--max-num-batched-tokens", type=int, default=8192, help="prefill token budget per engine step")
    ap.add_argument("--quantization", default=None, choices=[None, "fp8"],
                    help="fp8 = W8A8 side measurement (NOT the BASELINE config, which is bf16)")
    ap.add_argument("--no-e2e", action="store_true", help="skip the HTTP end-to-end leg")
    ap.add_argument("--no-cpu-baseline", action="store_true")
    ap.add_argument("--no-parity-probe", action="store_true",
                    help="skip the correctness probe (2-layer Llama-3-8B-width model vs the CPU oracle) run before timing")
    ap.add_argument("--gpu-baseline", default="auto", choices=["auto", "on", "off"],
                    help="time the reference GPU engine (vLLM, tools/vllm_baseline.py) on the same workload in the "
                         "same lease; auto = at TP 1 and 8 when the time budget allows")
    ap.add_argument("--time-budget-s", type=float, default=float(os.environ.get("B200_BENCH_BUDGET_S", "780")),
                    help="wall-clock budget of the whole run; optional legs are skipped / bounded to stay inside it")
    ap.add_argument("--cpu-sample-requests", type=int, default=2)
    ap.add_argument("--cpu-sample-output", type=int, default=16)
    ap.add_argument("--no-cpu-configs0", action="store_true",
                    help="skip the BASELINE configs[0] CPU leg (Qwen2-0.5B shape, 32 requests on the host cores)")
    return ap.parse_args()


def peaks():
    p = ROOT / "MEASURED_PEAKS.json"
    if p.exists():
        d = json.loads(p.read_text())
        return {"hbm_gbs": d["hbm_gbs"], "bf16_tflops": d["bf16_tflops"],
                "bf16_tflops_sustained": d.get("bf16_tflops_sustained", d["bf16_tflops"]), "source": "measured"}
    return {"hbm_gbs": 6650.0, "bf16_tflops": 1590.0, "bf16_tflops_sustained": 1400.0, "source": "fallback"}


def host_threads() -> int:
    """Usable host cores: min(affinity, cgroup CPU quota).  (The GPU boxes expose 128 logical CPUs
    but cap the container at 16 via cpu.max; asking torch for 128 threads there is pathological.)"""
    n = len(os.sched_getaffinity(0)) if hasattr(os, "sched_getaffinity") else (os.cpu_count() or 1)
    try:
        quota, period = open("/sys/fs/cgroup/cpu.max").read().split()[:2]
        if quota != "max":
            n = min(n, max(1, int(int(quota) / int(period))))
    except Exception:
        pass
    return max(1, n)


def model_config(name):
    from gpustack_b200.config import llama3_8b, llama3_70b, mixtral_8x7b, qwen2_05b, tiny_llama, tiny_mixtral

    return {"llama3-8b": llama3_8b, "llama3-70b": llama3_70b, "tiny": tiny_llama, "tiny128": lambda: tiny_llama(128),
            "qwen2-0.5b": qwen2_05b,
            "mixtral-8x7b": mixtral_8x7b, "tiny-mixtral": tiny_mixtral}[name]()


def workload(step, R, in_len, vocab):
    import numpy as np

    rs = np.random.RandomState(1000 + step)
    lo = 256 if vocab > 1024 else 3
    return rs.randint(lo, vocab - lo, size=(R, in_len)).tolist()


# ------------------------------------------------------------------------------------ clocks
class ClockSampler:
    Q = ("clocks.sm,clocks.max.sm,power.draw,clocks_event_reasons.hw_slowdown,"
         "clocks_event_reasons.hw_thermal_slowdown,clocks_event_reasons.sw_thermal_slowdown,"
         "clocks_event_reasons.sw_power_cap")

    def __init__(self, gpu_index=0):
        self.proc, self.lines = None, []
        try:
            self.proc = subprocess.Popen(["nvidia-smi", f"--id={gpu_index}", f"--query-gpu={self.Q}",
                                          "--format=csv,noheader,nounits", "-lms", "200"],
                                         stdout=subprocess.PIPE, stderr=subprocess.DEVNULL, text=True)
            self.t = threading.Thread(target=self._read, daemon=True)
            self.t.start()
        except Exception:
            self.proc = None

    def _read(self):
        for line in self.proc.stdout:
            self.lines.append(line.strip())

    def stop(self):
        if self.proc is None:
            return {"sm_mhz": None, "sm_max_mhz": None, "reasons": ["nvidia-smi unavailable"]}
        self.proc.terminate()
        try:
            self.proc.wait(timeout=2)
        except Exception:
            self.proc.kill()
        sm, mx, reasons = [], [], set()
        names = ["hw_slowdown", "hw_thermal_slowdown", "sw_thermal_slowdown", "sw_power_cap"]
        for ln in self.lines:
            f = [x.strip() for x in ln.split(",")]
            if len(f) < 7:
                continue
            try:
                sm.append(float(f[0])); mx.append(float(f[1]))
            except ValueError:
                continue
            for n, v in zip(names, f[3:7]):
                if v.lower().startswith("active"):
                    reasons.add(n)
        busy = [x for x in sm if x > 0]
        return {"sm_mhz": statistics.median(busy) if busy else None, "sm_max_mhz": max(mx) if mx else None,
                "reasons": sorted(reasons), "samples": len(sm)}


# ------------------------------------------------------------------------- CPU baseline / reference
def cpu_reference(mname, n_req, in_len, out_len, state_cpu=None):
    """Times the CPU restatement of the path (oracle/llama_oracle.py: PyTorch fp32 over the same
    bf16 weights, all host cores) on a bounded sample: n_req requests, run one after another,
    each a full prefill + greedy decode with KV history.  Returns (tok/s, p50 TTFT ms, seconds)."""
    import torch

    from oracle import llama_oracle as O

    cfg = {"llama3-8b": O.LLAMA3_8B, "tiny": O.TINY, "tiny128": O.TINY128, "qwen2-0.5b": O.QWEN2_05B_SHAPE,
           "tiny-mixtral": O.TINY_MIXTRAL}[mname]
    torch.set_num_threads(host_threads())
    if state_cpu is None:
        # values are irrelevant for timing; uniform fill is much faster than randn for 8B params
        state_cpu = {}
        g = torch.Generator().manual_seed(0)
        base = O.init_weights(O.LlamaConfig(**{**cfg.__dict__, "num_hidden_layers": 1}), seed=0)
        for l in range(cfg.num_hidden_layers):
            for k, v in base.items():
                if k.startswith("model.layers.0."):
                    state_cpu[k.replace("layers.0.", f"layers.{l}.")] = v.clone()
        for k in ("model.embed_tokens.weight", "model.norm.weight", "lm_head.weight"):
            if k in base:
                state_cpu[k] = base[k]                     # tied models (Qwen2-0.5B) have no lm_head entry
    model = O.OracleLlama(cfg, state_cpu, fp32_cache=True)
    prompts = workload(9999, n_req, in_len, cfg.vocab_size)
    ttfts, toks = [], 0
    t0 = time.perf_counter()
    for i, p in enumerate(prompts):
        ts = time.perf_counter()
        logits = model.forward(i, p, last_only=True)[-1]
        nxt = int(torch.argmax(logits))
        ttfts.append((time.perf_counter() - ts) * 1e3)
        toks += 1
        for _ in range(out_len - 1):
            logits = model.forward(i, [nxt], last_only=True)[-1]
            nxt = int(torch.argmax(logits))
            toks += 1
        model.kv.pop(i, None)
    dt = time.perf_counter() - t0
    return toks / dt, statistics.median(ttfts), dt


def run_reference(a):
    rank = int(os.environ.get("RANK", "0"))
    if rank != 0:
        return
    import torch

    vals, ttf, secs = [], [], 0.0
    n_req, out_len = max(1, a.cpu_sample_requests), a.cpu_sample_output
    for k in range(a.warmup + a.steps):
        v, t, dt = cpu_reference(a.model, n_req if k >= a.warmup else 1, a.input_len, out_len if k >= a.warmup else 2)
        if k >= a.warmup:
            vals.append(v); ttf.append(t); secs += dt
    value = statistics.mean(vals)
    cores = torch.get_num_threads()
    sample = f"{n_req} requests x {a.input_len} in / {out_len} out per step, sequential, KV history on host"
    print(json.dumps({
        "impl": "reference", "metric": METRIC, "value": value, "unit": UNIT, "n_gpus": a.gpus, "steps": a.steps,
        "warmup": a.warmup, "ms_per_step": secs / a.steps * 1e3, "higher_is_better": True, "scaling": "strong",
        "vs_baseline": None, "dtype": "bf16 weights / fp32 math", "data": "synthetic",
        "config": {"workload": f"{a.model} {a.input_len}-in/{a.output_len}-out greedy (bounded CPU sample)",
                   "requests_per_step": n_req},
        "p50_ttft_ms": statistics.median(ttf),
        "cpu_baseline": {"value": value, "unit": UNIT, "cores": cores, "kind": "port", "sample": sample},
        "e2e": {"value": value, "unit": UNIT, "h2d_bytes_per_step": 0, "d2h_bytes_per_step": 0},
        "note": "llama-box / llama.cpp are absent offline; this is oracle/llama_oracle.py on host cores",
    }))


# ------------------------------------------------------------------------------------ ours
def free_port():
    with socket.socket() as s:
        s.bind(("127.0.0.1", 0))
        return s.getsockname()[1]


# ------------------------------------------------------------------------------------ parity probe
PROBE_TOL = 2e-2   # relative to max|logit|: the engine-level tolerance of tests/test_engine_gpu.py


def parity_probe(a, rank, world):
    """Correctness of the EXACT code path that is about to be timed, at BASELINE width: a 2-layer
    Llama-3-8B-shaped model (H 4096, I 14336, 32/8 heads of 128, vocab 32768; the same tile shapes,
    split plans and -- under TP -- all-reduce paths as the benchmark model) generates 16 greedy
    tokens for 8 seeded prompts; the CPU oracle is teacher-forced on the engine's tokens and every
    step whose oracle top-2 margin exceeds the tolerance must agree.  Run twice: 8 sequences, and
    256 concurrent sequences (decode batch 256 = the benchmark's batch, which under TP takes the
    large-message all-reduce path) of which the same 8 are checked.  Every rank takes part."""
    import torch

    from gpustack_b200.config import EngineConfig, ModelConfig
    from gpustack_b200.engine import LLMEngine, TPWorker
    from gpustack_b200.scheduler import SamplingParams
    from gpustack_b200.weights import random_state_dict

    mcfg = ModelConfig(vocab_size=32768, hidden_size=4096, intermediate_size=14336, num_hidden_layers=2,
                       num_attention_heads=32, num_key_value_heads=8, head_dim=128, rms_norm_eps=1e-5,
                       rope_theta=500000.0, max_position_embeddings=1024, eos_token_id=2, bos_token_id=1)
    device = f"cuda:{torch.cuda.current_device()}"
    t0 = time.perf_counter()
    runs, oracle = [], None
    import numpy as np

    rs = np.random.RandomState(4242)
    checked_prompts = [rs.randint(3, mcfg.vocab_size, size=n).tolist() for n in (33, 64, 100, 128, 200, 17, 256, 77)]
    filler = [rs.randint(3, mcfg.vocab_size, size=48).tolist() for _ in range(248)]
    for max_seqs in (8, 256):
        sd = random_state_dict(mcfg, seed=7, device=device)
        if rank == 0 and oracle is None:
            from oracle import llama_oracle as O   # the checker; never on the timed path

            oc = O.LlamaConfig(vocab_size=mcfg.vocab_size, hidden_size=mcfg.hidden_size,
                               intermediate_size=mcfg.intermediate_size, num_hidden_layers=mcfg.num_hidden_layers,
                               num_attention_heads=mcfg.num_attention_heads,
                               num_key_value_heads=mcfg.num_key_value_heads, head_dim=mcfg.head_dim,
                               rms_norm_eps=mcfg.rms_norm_eps, rope_theta=mcfg.rope_theta,
                               max_position_embeddings=mcfg.max_position_embeddings)
            torch.set_num_threads(host_threads())
            oracle = O.OracleLlama(oc, {k: v.cpu() for k, v in sd.items()}, fp32_cache=True)
        ecfg = EngineConfig(model="random:probe", tensor_parallel_size=world, max_num_seqs=max_seqs,
                            max_num_batched_tokens=4096, max_model_len=512, num_gpu_blocks=8192,
                            enable_prefix_caching=False, seed=7,
                            cuda_graph_sizes=[1, 2, 4, 8] + ([128, 256] if max_seqs == 256 else []))
        eng = LLMEngine.from_engine_config(ecfg, mcfg=mcfg, state_dict=sd)
        del sd
        if rank != 0:
            TPWorker(eng).serve_forever()
            del eng
            torch.cuda.empty_cache()
            continue
        prompts = checked_prompts + (filler if max_seqs == 256 else [])
        outs = eng.generate(prompts, SamplingParams(max_tokens=16, ignore_eos=True))
        n_chk = n_ok = 0
        for i, (p, o) in enumerate(zip(checked_prompts, outs[:len(checked_prompts)])):
            oracle.reset()
            logits = oracle.forward(i, p, last_only=True)[-1]
            for t in range(len(o)):
                top2 = torch.topk(logits, 2).values
                if float(top2[0] - top2[1]) > PROBE_TOL * float(logits.abs().max()):
                    n_chk += 1
                    n_ok += int(int(torch.argmax(logits)) == int(o[t]))
                logits = oracle.forward(i, [int(o[t])], last_only=True)[-1]
        runs.append({"max_num_seqs": max_seqs, "concurrent": len(prompts), "checked": n_chk, "agree": n_ok,
                     "decode_steps": eng.runner.stats["decode_steps"]})
        eng.shutdown()
        del eng
        torch.cuda.empty_cache()
    if rank != 0:
        return None
    del oracle
    return {"tp": world, "model": "2-layer Llama-3-8B width (H4096 I14336 32/8x128 V32768), seeded N(0,0.02)",
            "protocol": "8 prompts x 16 greedy tokens, CPU oracle teacher-forced on the engine's tokens, steps with "
                        f"top-2 margin > {PROBE_TOL} x max|logit| must agree",
            "checked": sum(r["checked"] for r in runs), "agree": sum(r["agree"] for r in runs), "runs": runs,
            "seconds": round(time.perf_counter() - t0, 1)}


# ------------------------------------------------------------------------------------ reference GPU engine
def gpu_baseline(a, budget_s):
    """vLLM (the engine GPUStack's own backend launches) on the same workload, same lease, same
    step counts: tools/vllm_baseline.py in a fresh process (no torchrun environment), bounded by
    what is left of the time budget.  Returns a dict (with "error" when it could not be measured)."""
    env = {k: v for k, v in os.environ.items()
           if not (k in ("RANK", "LOCAL_RANK", "WORLD_SIZE", "LOCAL_WORLD_SIZE", "GROUP_RANK", "GROUP_WORLD_SIZE",
                         "ROLE_RANK", "ROLE_WORLD_SIZE", "ROLE_NAME", "MASTER_ADDR", "MASTER_PORT", "OMP_NUM_THREADS",
                         "NCCL_DEBUG", "NCCL_DEBUG_FILE", "NCCL_ASYNC_ERROR_HANDLING")
                   or k.startswith("TORCHELASTIC") or k.startswith("TORCH_NCCL"))}
    env["VLLM_LOGGING_LEVEL"] = "INFO"
    cmd = [sys.executable, str(ROOT / "tools" / "vllm_baseline.py"), str(a.requests), str(a.steps), "--tp", str(a.gpus),
           "--warmup", str(min(a.warmup, 2)), "--input-len", str(a.input_len), "--output-len", str(a.output_len)]
    t0 = time.perf_counter()
    try:
        r = subprocess.run(cmd, capture_output=True, text=True, timeout=max(60.0, budget_s), env=env, cwd=str(ROOT))
    except subprocess.TimeoutExpired:
        return {"engine": "vllm", "error": f"timed out after {budget_s:.0f}s (start-up + {a.steps} steps did not fit the budget)"}
    except Exception as e:  # noqa: BLE001
        return {"engine": "vllm", "error": repr(e)[:300]}
    out = None
    for line in reversed(r.stdout.splitlines()):
        if line.startswith("{"):
            try:
                out = json.loads(line)
                break
            except ValueError:
                continue
    if r.returncode != 0 or out is None:
        return {"engine": "vllm", "error": f"rc={r.returncode}: " + (r.stderr or r.stdout)[-400:].replace("\n", " | ")}
    try:
        sys.path.insert(0, str(ROOT / "tools"))
        from vllm_baseline import parse_backend_log

        out.update({k: v for k, v in parse_backend_log(r.stdout + r.stderr).items() if v})
    except Exception:  # noqa: BLE001
        pass
    out.pop("steps", None)
    out["wall_s"] = round(time.perf_counter() - t0, 1)
    out["same_lease"] = True
    return out


def gpus_idle(n, limit_mib=4000, wait_s=30.0):
    """True once the first n GPUs hold less than limit_mib each (the other ranks have exited)."""
    t0 = time.time()
    while time.time() - t0 < wait_s:
        try:
            r = subprocess.run(["nvidia-smi", "--query-gpu=memory.used", "--format=csv,noheader,nounits"],
                               capture_output=True, text=True, timeout=10)
            used = [int(x) for x in r.stdout.split()[:n]]
            if used and max(used) < limit_mib:
                return True
        except Exception:  # noqa: BLE001
            return True
        time.sleep(1.0)
    return False


def micro_rooflines(engine, a, mcfg, pk):
    """Device-timed (CUDA events, eager launches, weights of all layers rotate through HBM) roofline
    entries for the kernels that the dominant-kernel `roofline` would otherwise hide: the
    decode-shaped projections at M = batch, and prefill attention on a 16 x 512-token batch."""
    import torch

    from gpustack_b200 import ops

    run, w = engine.runner, engine.runner.w
    if run.fp8 or mcfg.num_local_experts:
        return {}
    M = min(a.requests, run.max_seqs)
    x = run.xn[:M]
    act = run.act[:M]
    res = {}

    def layer_pass():
        for lw in w.layers:
            ops.gemm(x, lw.qkv, out=run.qkv[:M])
            ops.gemm(run.attn_out[:M], lw.o, out=run.hbuf[:M])
            ops.gemm_silu_mul(x, lw.gate_up, out=act)
            ops.gemm(act, lw.down, out=run.hbuf[:M])

    layer_pass()
    torch.cuda.synchronize()
    e0, e1 = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
    reps = 3
    e0.record()
    for _ in range(reps):
        layer_pass()
    e1.record()
    torch.cuda.synchronize()
    L = len(w.layers)
    us_layer = e0.elapsed_time(e1) * 1e3 / (reps * L)
    wbytes = sum(t.numel() * 2 for t in (w.layers[0].qkv, w.layers[0].o, w.layers[0].gate_up, w.layers[0].down))
    flops = 2.0 * M * wbytes / 2
    res["roofline_decode_gemm"] = {
        "kernel": f"QKV + O + gate_up(SiLU) + down projections at M={M} (one decoder layer, 4 launches)",
        "us_per_layer": us_layer, "bound": "hbm", "achieved": wbytes / (us_layer * 1e-6) / 1e9, "peak": pk["hbm_gbs"],
        "unit": "GB/s", "frac": wbytes / (us_layer * 1e-6) / 1e9 / pk["hbm_gbs"],
        "tensor_tflops": flops / (us_layer * 1e-6) / 1e12, "tensor_frac": flops / (us_layer * 1e-6) / 1e12 / pk["bf16_tflops"],
        "algorithmic_bytes_per_layer": wbytes, "timing": "CUDA events around 3 x 32-layer eager passes (PDL launches)"}
    # prefill attention: 16 sequences x 512 tokens, causal, through the engine's own prefill path inputs
    S, n = max(1, min(16, run.max_tokens // a.input_len)), a.input_len
    if run.num_blocks > S * (n // run.block_size + 1) + 1:
        T = S * n
        i32 = lambda v: torch.tensor(v, dtype=torch.int32, device=run.device)   # noqa: E731
        bpp = n // run.block_size + 1
        bt = torch.zeros(S, run.max_blocks, dtype=torch.int32, device=run.device)
        for s_ in range(S):
            bt[s_, :bpp] = torch.arange(1 + s_ * bpp, 1 + (s_ + 1) * bpp, dtype=torch.int32)
        ts, tq = ops.build_prefill_tiles([n] * S, run.prefill_tile)
        cu, ctx, ts_d, tq_d = i32([i * n for i in range(S + 1)]), i32([0] * S), i32(ts), i32(tq)
        run.qkv[:T].normal_()

        def attn(l):
            ops.paged_attn_prefill(run.qkv[:T], run.qkv_width, run.k_cache[l], run.v_cache[l], bt, cu, ctx, ts_d, tq_d,
                                   run.attn_out[:T], run.Hq, run.Hkv, run.D, run.scale)

        for l in range(min(4, L)):
            attn(l)
        torch.cuda.synchronize()
        e0.record()
        for l in range(L):
            attn(l)
        e1.record()
        torch.cuda.synchronize()
        us = e0.elapsed_time(e1) * 1e3 / L
        fl = 4.0 * S * (n * (n + 1) / 2) * run.Hq * run.D
        res["roofline_prefill_attn"] = {
            "kernel": "paged_attn_prefill (tcgen05/TMEM FMHA for head_dim 128, GQA 4|8)", "us_per_launch": us,
            "bound": "tensor", "achieved": fl / (us * 1e-6) / 1e12, "peak": pk["bf16_tflops"], "unit": "TFLOP/s",
            "frac": fl / (us * 1e-6) / 1e12 / pk["bf16_tflops"], "flops_per_launch": fl,
            "shape": f"{S} x {n} tokens causal, Hq {run.Hq} / Hkv {run.Hkv} x {run.D}",
            "peak_source": f"{pk['source']} (burst cuBLAS bf16: kernel timed alone)"}
    return res


def run_reference_harness(url, a):
    """One pass of the reference's own load generator (benchmarks/benchmark_serving.py, unmodified:
    /v1/chat/completions, text prompts, TTFT = first SSE chunk, tok/s = completion tokens / wall)
    against the server that is up for the e2e leg.  Prompt length is tokenizer dependent (synthetic
    byte tokenizer here): the achieved mean prompt_tokens is reported."""
    sys.path.insert(0, str(ROOT / "tools"))
    try:
        from run_reference_harness import harness_path

        if harness_path() is None:
            return {"unavailable": "benchmark_serving.py not present (baseline/_ref/ is filled by __graft_entry__.build())"}
        cmd = [sys.executable, str(ROOT / "tools" / "run_reference_harness.py"), "-m", "bench-model", "-n", str(a.requests),
               "-c", str(a.requests), "--max-completion-tokens", str(a.output_len), "--ignore-eos",
               "--prompt-multiplier", "4", "--request-timeout", "600", "--server-url", url, "--json"]
        r = subprocess.run(cmd, capture_output=True, text=True, cwd=str(ROOT), timeout=600)
        if r.returncode != 0 or "{" not in r.stdout:
            return {"error": (r.stderr or r.stdout)[-300:]}
        d = json.loads(r.stdout[r.stdout.index("{"):])
        ok = max(1, d.get("successful_requests", 0))
        return {"script": "benchmarks/benchmark_serving.py (reference, unmodified; tools/run_reference_harness.py)",
                "endpoint": "/v1/chat/completions", "requests": d.get("total_requests"), "successful": d.get("successful_requests"),
                "concurrency": d.get("concurrency"), "output_tok_s": d.get("total_completion_tokens_per_second"),
                "p50_ttft_ms": (d.get("time_to_first_token") or {}).get("p50"),
                "mean_prompt_tokens": d.get("total_prompt_tokens", 0) / ok,
                "mean_completion_tokens": d.get("total_completion_tokens", 0) / ok, "seed": 0, "prompt_multiplier": 4}
    except Exception as e:  # noqa: BLE001
        return {"error": repr(e)[:300]}


def run_http_leg(engine, a, mcfg, n_warm, n_timed):
    """Serve the SAME engine over HTTP exactly as `b200-engine serve` does -- this process keeps
    running the engine-core busy loop (on a thread here, bench.py's main thread drives the client),
    the OpenAI frontend runs in a child process -- and load it from a separate client process."""
    from gpustack_b200.frontend import spawn_frontend

    port = free_port()
    front, core = spawn_frontend(engine, "127.0.0.1", port, "bench-model", None, "error")
    th = threading.Thread(target=core.run, name="engine-core", daemon=True)
    th.start()
    import urllib.request

    for _ in range(600):
        try:
            if urllib.request.urlopen(f"http://127.0.0.1:{port}/health", timeout=1).status == 200:
                break
        except Exception:
            time.sleep(0.1)
    st0 = None
    out = []
    try:
        cmd = [sys.executable, "-m", "gpustack_b200.bench_client", "--url", f"http://127.0.0.1:{port}", "--model",
               "bench-model", "-n", str(a.requests), "-c", str(a.requests), "--input-len", str(a.input_len),
               "--output-len", str(a.output_len), "--vocab", str(mcfg.vocab_size)]
        if n_warm:
            subprocess.run(cmd + ["--seed", "5000", "--repeat", str(n_warm)], check=True, capture_output=True,
                           cwd=str(ROOT), timeout=600)
            log("http warm-up done")
        st0 = dict(engine.runner.stats)
        r = subprocess.run(cmd + ["--seed", "6000", "--repeat", str(n_timed)], check=True, capture_output=True,
                           text=True, cwd=str(ROOT), timeout=900)
        out = [json.loads(l) for l in r.stdout.splitlines() if l.startswith("{")]
        harness = run_reference_harness(f"http://127.0.0.1:{port}", a)
    finally:
        front.terminate()
        front.join(timeout=10)
        th.join(timeout=10)          # the core loop ends when its pipe closes
    st1 = engine.runner.stats
    comp = sum(o["completion_tokens"] for o in out)
    tt = sum(o["total_time_s"] for o in out)
    ttft = [o["ttft_ms_p50"] for o in out if o["ttft_ms_p50"] is not None]
    return {
        "value": comp / tt if tt else 0.0, "unit": UNIT,
        "h2d_bytes_per_step": (st1["h2d_bytes"] - st0["h2d_bytes"]) // max(1, n_timed),
        "d2h_bytes_per_step": (st1["d2h_bytes"] - st0["d2h_bytes"]) // max(1, n_timed),
        "p50_ttft_ms": statistics.median(ttft) if ttft else None,
        "p99_ttft_ms": max((o["ttft_ms_p99"] or 0) for o in out) if out else None,
        "successful_requests": sum(o["successful"] for o in out), "requests": sum(o["requests"] for o in out),
        "path": "HTTP /v1/completions SSE: frontend process <-pipe-> engine core; client in a third process "
                "(host JSON in, host text out)",
        "harness": harness,
    }


def run_ours(a):
    import torch

    from gpustack_b200 import ops
    from gpustack_b200.config import EngineConfig
    from gpustack_b200.engine import LLMEngine, TPWorker
    from gpustack_b200.scheduler import SamplingParams

    # NCCL_DEBUG is left as the launcher set it (the communicator log is the driver's evidence of the
    # rank count); the JSON line is printed LAST, after the process group is gone.
    world = int(os.environ.get("WORLD_SIZE", "1"))
    rank = int(os.environ.get("RANK", "0"))
    if world != a.gpus:
        raise SystemExit(f"--gpus {a.gpus} needs one process per GPU (WORLD_SIZE={world}); launch with "
                         "python -m torch.distributed.run --nproc-per-node N bench.py --gpus N ...")
    torch.cuda.set_device(int(os.environ.get("LOCAL_RANK", "0")))
    parity = None
    if not a.no_parity_probe:
        tpp = time.perf_counter()
        parity = parity_probe(a, rank, world)
        if rank == 0:
            log(f"parity probe ({time.perf_counter() - tpp:.1f}s): {parity}")
    mcfg = model_config(a.model)
    ecfg = EngineConfig(model=f"random:{a.model}", tensor_parallel_size=a.gpus, max_num_seqs=a.requests,
                        max_num_batched_tokens=a.max_num_batched_tokens,
                        max_model_len=min(mcfg.max_position_embeddings, max(1024, a.input_len + a.output_len + 16)),
                        enable_prefix_caching=False,    # every timed token is computed (same as the vLLM baseline run)
                        quantization=a.quantization,
                        mixed_steps=os.environ.get("B200_MIXED_STEPS", "1") != "0",
                        seed=0)
    t_init = time.perf_counter()
    engine = LLMEngine.from_engine_config(ecfg, mcfg=mcfg)
    init_s = time.perf_counter() - t_init
    log(f"rank {rank}: engine ready in {init_s:.1f}s, {engine.num_blocks} KV blocks, graphs {len(engine.runner.graphs)}")
    if rank != 0:
        TPWorker(engine).serve_forever()
        return None
    sp = SamplingParams(max_tokens=a.output_len, ignore_eos=True)
    R, K, W = a.requests, a.steps, a.warmup

    for k in range(W):
        tw = time.perf_counter()
        engine.generate(workload(k, R, a.input_len, mcfg.vocab_size), sp)
        log(f"warm-up step {k}: {time.perf_counter() - tw:.2f}s  step_time={engine.step_time}")
    clocks = ClockSampler(int(os.environ.get("LOCAL_RANK", "0")))
    st0 = dict(engine.runner.stats)
    stime0 = dict(engine.step_time)
    ops.TIMER = ops.KernelTimer()
    e0, e1 = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
    engine.sync_mark()
    e0.record()
    n_out = 0
    for k in range(K):
        outs = engine.generate(workload(W + k, R, a.input_len, mcfg.vocab_size), sp)
        n_out += sum(len(o) for o in outs)
    e1.record()
    engine.sync_mark()
    elapsed = engine.max_elapsed()
    ev_ms = e0.elapsed_time(e1)
    log(f"timed region: {K} steps in {elapsed:.2f}s -> {n_out / elapsed:.0f} tok/s")
    clk = clocks.stop()
    timer = ops.TIMER.summary()
    ops.TIMER = None
    st1, stime1 = engine.runner.stats, engine.step_time
    value = n_out / elapsed
    pk = peaks()

    # ---- roofline of the dominant kernel (prefill-shaped tcgen05 GEMM), timed live with CUDA events
    gp = timer.get("gemm_prefill", {"launches": 0, "ms": 0.0, "flops": 0.0, "bytes": 0.0})
    gd = timer.get("gemm_decode", {"launches": 0, "ms": 0.0, "flops": 0.0, "bytes": 0.0})
    ach = gp["flops"] / (gp["ms"] * 1e-3) / 1e12 if gp["ms"] else 0.0
    traffic, traffic_src = None, None
    tp_ = ROOT / "profiles" / "r2_gemm_prefill_traffic.json"
    if tp_.exists() and a.model == "llama3-8b" and a.gpus == 1:
        tj = json.loads(tp_.read_text())     # dram__bytes_read+write per launch from the committed ncu capture
        traffic, traffic_src = tj["avg_dram_bytes_per_launch"], tj["source"]
    roofline = {"kernel": "gemm_bf16_tcgen05_2sm_kernel<256,*> (prefill projections: QKV, O, gate_up+SiLU, down)",
                "bound": "tensor", "achieved": ach, "peak": pk["bf16_tflops_sustained"], "unit": "TFLOP/s",
                "frac": ach / pk["bf16_tflops_sustained"], "traffic": traffic, "traffic_source": traffic_src,
                "algorithmic_bytes_per_launch": gp["bytes"] / max(1, gp["launches"]), "launches": gp["launches"],
                "avg_launch_ms": gp["ms"] / max(1, gp["launches"]), "share_of_step": gp["ms"] / ev_ms,
                "peak_source": f"{pk['source']} (sustained cuBLAS bf16)"}
    # ---- decode phase as a whole against the HBM roofline (weights + KV streamed per step)
    dsteps = st1["decode_steps"] - st0["decode_steps"]
    dtime = stime1["decode"] - stime0["decode"]
    kv_tok = st1["decode_kv_tokens"] - st0["decode_kv_tokens"]
    dbytes = dsteps * mcfg.weight_bytes(a.gpus, a.quantization) + kv_tok * mcfg.kv_bytes_per_token(a.gpus)
    dgbs = dbytes / dtime / 1e9 if dtime else 0.0
    decode_roof = {"bound": "hbm", "achieved": dgbs, "peak": pk["hbm_gbs"], "unit": "GB/s",
                   "frac": dgbs / pk["hbm_gbs"], "steps": dsteps, "ms_per_decode_step": dtime / max(1, dsteps) * 1e3,
                   "note": "whole decode steps (host-timed incl. launch+sync), algorithmic bytes = weights + KV read"}
    ptime = stime1["prefill"] - stime0["prefill"]
    ptok = st1["prefill_tokens"] - st0["prefill_tokens"]

    res = {
        "metric": METRIC, "value": value, "unit": UNIT, "n_gpus": a.gpus, "steps": K, "warmup": W,
        "ms_per_step": elapsed / K * 1e3, "higher_is_better": True, "scaling": "strong", "vs_baseline": None,
        "dtype": "bf16" if a.quantization is None else "fp8 e4m3 W8A8 linears, bf16 elsewhere", "data": "synthetic",
        "config": {"workload": f"Llama-3-8B {'bf16' if a.quantization is None else 'fp8-W8A8'} TP={a.gpus}, continuous batching, {a.input_len}-in/{a.output_len}-out greedy"
                   if a.model == "llama3-8b" else
                   f"{a.model} {'bf16' if a.quantization is None else 'fp8-W8A8'} TP={a.gpus} {a.input_len}-in/{a.output_len}-out",
                   "requests_per_step": R, "concurrency": R, "weights": "seeded N(0,0.02) random init",
                   "prompts": "seeded uniform token ids, fresh per step (working set 16 GB weights >> 126 MB L2)",
                   "parallelism": f"tp{a.gpus}", "max_num_batched_tokens": a.max_num_batched_tokens,
                   "prefix_caching": False,
                   "inputs": "token ids staged through one pinned buffer per engine step"},
        "gpu_launches": st1["kernel_launches"] - st0["kernel_launches"],
        "clocks": clk, "roofline": roofline, "decode_roofline": decode_roof,
        "phase_ms_per_step": {"prefill": ptime / K * 1e3, "decode": dtime / K * 1e3,
                              "prefill_tok_s": ptok / ptime if ptime else None},
        "gemm_decode_live": {"launches": gd["launches"], "ms": gd["ms"],
                             "gbs": gd["bytes"] / (gd["ms"] * 1e-3) / 1e9 if gd["ms"] else None},
        "event_ms_total": ev_ms, "init_s": init_s, "kv_blocks": engine.num_blocks,
        "decode_step_breakdown_ms": {
            "wall": dtime / max(1, dsteps) * 1e3,
            "h2d+graph+sync": (st1.get("decode_gpu_s", 0.0) - st0.get("decode_gpu_s", 0.0)) / max(1, dsteps) * 1e3,
            "tp_step_broadcast": (stime1.get("decode_bcast", 0.0) - stime0.get("decode_bcast", 0.0)) / max(1, dsteps) * 1e3},
    }
    if not a.no_e2e:
        try:
            log("HTTP end-to-end leg ...")
            res["e2e"] = run_http_leg(engine, a, mcfg, n_warm=min(W, 2), n_timed=K)
            res["p50_ttft_ms"] = res["e2e"]["p50_ttft_ms"]
            log(f"e2e: {res['e2e']}")
        except Exception as e:  # noqa: BLE001
            res["e2e"] = {"value": None, "unit": UNIT, "error": repr(e)[:300]}
    if parity is not None:
        res["tp_parity"] = parity
    engine.shutdown()                      # the other ranks leave serve_forever and exit
    try:
        res.update(micro_rooflines(engine, a, mcfg, pk))
    except Exception as e:  # noqa: BLE001
        res["roofline_decode_gemm"] = {"error": repr(e)[:200]}
    del engine
    torch.cuda.empty_cache()
    if a.gpus == 1 and not a.no_cpu_baseline:
        log("CPU baseline (oracle port) ...")
        v, t, dt = cpu_reference(a.model, a.cpu_sample_requests, a.input_len, a.cpu_sample_output)
        log(f"cpu baseline {v:.2f} tok/s in {dt:.1f}s")
        res["cpu_baseline"] = {"value": v, "unit": UNIT, "cores": torch.get_num_threads(), "kind": "port",
                               "p50_ttft_ms": t,
                               "sample": f"{a.cpu_sample_requests} requests x {a.input_len} in / {a.cpu_sample_output} out, "
                                         f"sequential, {dt:.1f} s of CPU work (oracle/llama_oracle.py, fp32 math)"}
        if not a.no_cpu_configs0 and a.time_budget_s - (time.time() - _T0) > 420:
            # BASELINE.json configs[0]: the reference's CPU-runnable case -- Qwen2-0.5B shape,
            # the full 32 requests of the plumbing run, on this box's host cores, in this same run
            log("CPU baseline, configs[0]: Qwen2-0.5B shape x 32 requests ...")
            v0, t0_, dt0 = cpu_reference("qwen2-0.5b", 32, 64, 8)
            res["cpu_baseline"]["configs0"] = {
                "model": "Qwen2-0.5B shape (random init; llama-box / GGUF unavailable offline: oracle port)", "value": v0,
                "unit": UNIT, "p50_ttft_ms": t0_, "cores": torch.get_num_threads(),
                "sample": f"32 requests x 64 in / 8 out, sequential, {dt0:.1f} s of CPU work"}
            log(f"cpu configs[0]: {v0:.1f} tok/s in {dt0:.1f}s")
    return res


def finish_ours(a, res):
    """Runs on rank 0 after the process group is destroyed: the same-lease reference-GPU-engine leg,
    then the ONE JSON line, last on stdout."""
    want = a.gpu_baseline == "on" or (a.gpu_baseline == "auto" and a.gpus in (1, 8) and a.model == "llama3-8b"
                                      and a.quantization is None)
    if want:
        left = a.time_budget_s - (time.time() - _T0)
        if left < 240 and a.gpu_baseline != "on":
            res["gpu_baseline"] = {"engine": "vllm", "error": f"skipped: {left:.0f}s of the {a.time_budget_s:.0f}s budget left"}
        elif not gpus_idle(a.gpus):
            res["gpu_baseline"] = {"engine": "vllm", "error": "skipped: the GPUs were still in use by exiting ranks"}
        else:
            log(f"reference GPU engine (vLLM, TP={a.gpus}), {left:.0f}s of budget left ...")
            res["gpu_baseline"] = gpu_baseline(a, left - 20.0)
            log(f"gpu_baseline: {res['gpu_baseline']}")
        gb = res["gpu_baseline"]
        if gb.get("output_tok_s"):
            gb["ours_over_baseline"] = res["value"] / gb["output_tok_s"]
    sys.stdout.flush()
    print(json.dumps(res), flush=True)
    bad = res.get("tp_parity") and res["tp_parity"]["agree"] != res["tp_parity"]["checked"]
    if bad:
        log("PARITY PROBE FAILED: " + json.dumps(res["tp_parity"]))
    return 1 if bad else 0


def main():
    a = parse_args()
    if a.impl == "reference":
        run_reference(a)
        return 0
    res = run_ours(a)
    try:
        import torch.distributed as dist

        if dist.is_initialized():
            dist.destroy_process_group()
    except Exception:
        pass
    if res is None:          # ranks > 0
        return 0
    if a.gpus > 1:
        time.sleep(2.0)      # let the other ranks' last NCCL / torchrun lines drain before the JSON line
    return finish_ours(a, res)


if __name__ == "__main__":
    sys.exit(main())
