"""Engine core: scheduler + model runner + (optional) tensor-parallel step broadcast.

Counterpart of the reference engine's EngineCore busy loop ($VLLM/v1/engine/core.py): one
iteration = schedule -> execute on the GPU -> update request state -> emit new tokens.
With tensor parallelism there is one process per GPU (torch.distributed, NCCL for tensors);
rank 0 owns the scheduler and broadcasts each step's inputs, the other ranks replay the same
kernel sequence on their weight / KV shards (TPWorker.serve_forever).
"""
from __future__ import annotations

import logging
import os
import time
from typing import Dict, Iterable, List, Optional, Sequence

import numpy as np
import torch

from .config import EngineConfig, ModelConfig
from .lib import B200Error
from .runner import ModelRunner
from .scheduler import (DecodePlan, MixedPlan, PrefillPlan, Request, SamplingParams, Scheduler, SpecPlan, StepOutput,
                        kernel_temperature, step_seeds)
from .weights import (ModelWeights, lazy_random_source, load_safetensors_dir, random_state_dict,
                      shard_state_dict)

log = logging.getLogger("b200.engine")

_KIND_STOP, _KIND_DECODE, _KIND_PREFILL, _KIND_SYNC, _KIND_MAXELAPSED, _KIND_GATHER, _KIND_MIXED = 0, 1, 2, 3, 4, 5, 6
_KIND_SWAP = 7          # extended KV cache: ordered device<->host block copies, executed before the step that follows


class ShmStepChannel:
    """Single-writer (rank 0) / multi-reader step channel in POSIX shared memory: the per-step
    scheduler output (a few KB to ~2 MB of int32) reaches the other ranks of the node in ~10 us
    instead of two gloo broadcasts (~250 us).  Counterpart of the reference engine's shared-memory
    MessageQueue ($VLLM/distributed/device_communicators/shm_broadcast.py).
    Layout: [0:8) sequence number (written LAST), [8:16) payload words, [16:80) 8 x int64 header,
    [128 + 8 r) ack of rank r, [4096: ...) int32 payload.  x86 store ordering + the copy-out before
    the ack make a fence-free protocol sufficient here."""

    PAYLOAD_OFF = 4096

    def __init__(self, rank: int, world: int, dist, cpu_group, capacity_words: int = 1 << 20):
        from multiprocessing import shared_memory

        self.rank, self.world = rank, world
        names = [None]
        if rank == 0:
            self.shm = shared_memory.SharedMemory(create=True, size=self.PAYLOAD_OFF + capacity_words * 4)
            self.shm.buf[: self.PAYLOAD_OFF] = bytes(self.PAYLOAD_OFF)
            names[0] = self.shm.name
        dist.broadcast_object_list(names, src=0, group=cpu_group)
        if rank != 0:
            self.shm = shared_memory.SharedMemory(name=names[0])
            try:   # only the creator may unlink; keep this process's resource tracker out of it
                from multiprocessing import resource_tracker

                resource_tracker.unregister(self.shm._name, "shared_memory")
            except Exception:  # noqa: BLE001
                pass
        self.cap = capacity_words
        self.seq = np.ndarray((1,), dtype=np.uint64, buffer=self.shm.buf, offset=0)
        self.nwords = np.ndarray((1,), dtype=np.int64, buffer=self.shm.buf, offset=8)
        self.header = np.ndarray((8,), dtype=np.int64, buffer=self.shm.buf, offset=16)
        self.acks = np.ndarray((world,), dtype=np.uint64, buffer=self.shm.buf, offset=128)
        self.payload = np.ndarray((capacity_words,), dtype=np.int32, buffer=self.shm.buf, offset=self.PAYLOAD_OFF)
        self.last = 0
        dist.barrier(group=cpu_group)

    @staticmethod
    def _spin(cond) -> None:
        n = 0
        while not cond():
            n += 1
            if n > 20000:
                time.sleep(0.0002 if n < 200000 else 0.002)   # idle engine: stop burning a core
            elif (n & 63) == 0:
                time.sleep(0)

    def send(self, header: List[int], payload: Optional[np.ndarray]) -> None:
        cur = self.last
        self._spin(lambda: all(int(self.acks[r]) == cur for r in range(1, self.world)))
        n = 0 if payload is None else len(payload)
        if n > self.cap:
            raise B200Error(f"step payload of {n} words exceeds the shared-memory channel ({self.cap})")
        if n:
            self.payload[:n] = payload
        self.header[:] = 0
        self.header[: len(header)] = header
        self.nwords[0] = n
        self.last = cur + 1
        self.seq[0] = self.last            # publish

    def recv(self):
        nxt = self.last + 1
        self._spin(lambda: int(self.seq[0]) >= nxt)
        n = int(self.nwords[0])
        header = self.header.tolist()
        payload = self.payload[:n].copy() if n else None
        self.last = nxt
        self.acks[self.rank] = nxt         # rank 0 may now overwrite the buffer
        return header, payload

    def close(self) -> None:
        try:
            for a in ("seq", "nwords", "header", "acks", "payload"):
                setattr(self, a, None)
            self.shm.close()
            if self.rank == 0:
                self.shm.unlink()
        except Exception:  # noqa: BLE001
            pass


class TPContext:
    """torch.distributed plumbing for one tensor-parallel group (one process per GPU)."""

    def __init__(self, rank: int, world: int, device: torch.device):
        import torch.distributed as dist

        self.rank, self.world, self.device = rank, world, device
        self.dist = dist
        self.nccl = dist.group.WORLD
        self.cpu = dist.new_group(backend="gloo")   # bootstrap / barriers; fallback step channel
        self.shm: Optional[ShmStepChannel] = None

    def open_step_channel(self, capacity_words: int) -> None:
        """Shared-memory step channel sized for the largest step payload of this engine (so an
        oversized step can never surface in the middle of serving).  The fence-free protocol relies
        on x86 store ordering: on other hosts (aarch64 Grace) the gloo broadcast is used instead."""
        import platform

        dist = self.dist
        want = os.environ.get("B200_TP_STEP_CHANNEL", "shm") == "shm" and platform.machine() in ("x86_64", "AMD64")
        if want:
            try:
                self.shm = ShmStepChannel(self.rank, self.world, dist, self.cpu, capacity_words=capacity_words)
            except Exception:  # noqa: BLE001  (no /dev/shm, ...): gloo still works
                self.shm = None
        ok = torch.tensor([1 if self.shm is not None else 0])
        dist.all_reduce(ok, op=dist.ReduceOp.MIN, group=self.cpu)
        if int(ok[0]) == 0 and self.shm is not None:
            self.shm.close()
            self.shm = None

    @staticmethod
    def from_env() -> Optional["TPContext"]:
        import torch.distributed as dist

        world = int(os.environ.get("WORLD_SIZE", "1"))
        if world <= 1:
            return None
        rank = int(os.environ["RANK"])
        local = int(os.environ.get("LOCAL_RANK", rank))
        torch.cuda.set_device(local)
        if not dist.is_initialized():
            dist.init_process_group(backend="nccl", device_id=torch.device("cuda", local))
        return TPContext(rank, world, torch.device("cuda", local))

    def bcast_step(self, header: List[int], payload: Optional[np.ndarray]):
        """rank 0 sends (header, int32 payload); the other ranks receive them."""
        if self.shm is not None:
            if self.rank == 0:
                self.shm.send(header, None if payload is None else np.ascontiguousarray(payload, dtype=np.int32))
                return header + [0] * (8 - len(header)), payload
            return self.shm.recv()
        h = torch.zeros(8, dtype=torch.int64)
        if self.rank == 0:
            h[:len(header)] = torch.tensor(header, dtype=torch.int64)
            h[7] = 0 if payload is None else len(payload)
        self.dist.broadcast(h, src=0, group=self.cpu)
        n = int(h[7])
        if n:
            t = torch.from_numpy(np.ascontiguousarray(payload)) if self.rank == 0 else torch.empty(n, dtype=torch.int32)
            self.dist.broadcast(t, src=0, group=self.cpu)
            payload = t.numpy()
        return h.tolist(), payload


class LLMEngine:
    def __init__(self, mcfg: ModelConfig, ecfg: EngineConfig, weights: ModelWeights, tp: Optional[TPContext] = None,
                 device: str = "cuda"):
        self.mcfg, self.ecfg, self.tp = mcfg, ecfg, tp
        if ecfg.max_model_len is None:
            ecfg.max_model_len = mcfg.max_position_embeddings
        ecfg.max_model_len = min(ecfg.max_model_len, mcfg.max_position_embeddings)
        if mcfg.sliding_window and ecfg.max_model_len > mcfg.sliding_window:
            log.warning("sliding-window attention is not implemented: max_model_len capped at the window (%d)",
                        mcfg.sliding_window)
            ecfg.max_model_len = mcfg.sliding_window
        self.runner = ModelRunner(mcfg, ecfg, weights, device=device, tp_group=tp.nccl if tp else None, tp_ctx=tp)
        if tp and tp.shm is None:
            r = self.runner
            # largest payloads: decode = B*(2+max_blocks); prefill = 4*S + S*max_blocks + T
            tp.open_step_channel(r.max_seqs * (r.max_blocks + 4) + r.max_tokens + 4096)
        n = self.runner.profile_num_blocks()
        if tp:   # every rank must use the same block count
            t = torch.tensor([n], dtype=torch.int64)
            tp.dist.all_reduce(t, op=tp.dist.ReduceOp.MIN, group=tp.cpu)
            n = int(t[0])
        self.num_blocks = n
        self.runner.init_kv_cache(n)
        self.runner.capture_graphs()
        # extended KV cache: evicted prefix blocks go to pinned host memory (same slot count on every rank)
        host_slots = 0
        if ecfg.kv_host_cache_gib > 0 and ecfg.enable_prefix_caching:
            host_slots = int(ecfg.kv_host_cache_gib * (1 << 30)) // self.runner.kv_block_bytes()
        self.runner.init_host_kv(host_slots)
        self.sched = Scheduler(n, ecfg.block_size, ecfg.max_num_seqs, ecfg.max_num_batched_tokens,
                               ecfg.max_model_len, mcfg.eos_token_ids or mcfg.eos_token_id, ecfg.enable_prefix_caching,
                               mixed_steps=ecfg.mixed_steps, speculative=ecfg.speculative(),
                               max_sample_rows=self.runner.max_sample, kv_host_slots=host_slots)
        self.requests: Dict[str, Request] = {}
        self.counters = {"prompt_tokens": 0, "generation_tokens": 0, "finished": 0, "steps": 0}
        self.step_time = {"prefill": 0.0, "decode": 0.0}

    # ------------------------------------------------------------------ construction
    @classmethod
    def from_engine_config(cls, ecfg: EngineConfig, mcfg: Optional[ModelConfig] = None,
                           state_dict: Optional[dict] = None) -> "LLMEngine":
        """Model from `ecfg.model` (an HF-format directory), or from an explicit state dict, or --
        when ecfg.model is 'random:<name>' / mcfg is given without weights -- seeded random init."""
        tp = TPContext.from_env() if ecfg.tensor_parallel_size > 1 else None
        if ecfg.tensor_parallel_size > 1 and (tp is None or tp.world != ecfg.tensor_parallel_size):
            raise B200Error(f"tensor_parallel_size={ecfg.tensor_parallel_size} needs WORLD_SIZE to match "
                            "(launch one process per GPU, e.g. with torch.distributed.run)")
        rank = tp.rank if tp else 0
        device = f"cuda:{torch.cuda.current_device()}"
        if state_dict is None and mcfg is None:
            mcfg = ModelConfig.from_pretrained(ecfg.model)
            quant = ecfg.quantization or mcfg.quantization      # CLI flag, else the checkpoint's own config
            weights = load_safetensors_dir(mcfg, ecfg.model, rank, ecfg.tensor_parallel_size, device,
                                           quantization=quant)
        elif state_dict is None and (mcfg.num_local_experts or mcfg.weight_bytes(1) > 48e9):
            # mixture-of-experts / 70B shapes are too large to hold the synthetic checkpoint and its
            # fused shards at the same time: synthesise tensor by tensor
            get, has = lazy_random_source(mcfg, seed=ecfg.seed, device=device)
            weights = shard_state_dict(mcfg, get, rank, ecfg.tensor_parallel_size, device,
                                       quantization=ecfg.quantization or mcfg.quantization, has=has)
            torch.cuda.empty_cache()
        else:
            if state_dict is None:
                state_dict = random_state_dict(mcfg, seed=ecfg.seed, device=device)
            weights = shard_state_dict(mcfg, lambda k: state_dict[k], rank, ecfg.tensor_parallel_size, device,
                                       quantization=ecfg.quantization or mcfg.quantization,
                                       has=lambda k: k in state_dict)
            del state_dict
            torch.cuda.empty_cache()
        return cls(mcfg, ecfg, weights, tp=tp, device=device)

    # ------------------------------------------------------------------ request side
    def add_request(self, req_id: str, prompt_ids: Sequence[int], params: SamplingParams) -> None:
        if req_id in self.requests:
            raise ValueError(f"duplicate request id {req_id}")
        V = self.mcfg.vocab_size
        if any((t < 0 or t >= V) for t in prompt_ids):
            raise ValueError("prompt token id out of range")
        need = (len(prompt_ids) + params.max_tokens + self.ecfg.block_size) // self.ecfg.block_size
        if need > self.sched.pool.num_total:
            raise ValueError("request does not fit in the KV cache")
        r = Request(req_id=req_id, prompt_ids=list(prompt_ids), params=params)
        self.sched.add(r)
        self.requests[req_id] = r
        self.counters["prompt_tokens"] += len(prompt_ids)

    def abort_request(self, req_id: str) -> None:
        if self.sched.abort(req_id):
            self.requests.pop(req_id, None)

    def has_work(self) -> bool:
        return self.sched.has_work()

    # ------------------------------------------------------------------ one iteration
    def step(self) -> Optional[StepOutput]:
        plan = self.sched.schedule()
        swaps = self.sched.pool.take_swap_ops() if self.sched.pool.host_slots else None
        if swaps:       # host-tier copies decided while scheduling: on the stream BEFORE this step's kernels
            if self.tp:
                self.tp.bcast_step([_KIND_SWAP, len(swaps)], np.asarray(swaps, dtype=np.int32).reshape(-1))
            self.runner.kv_swap(swaps)
        if plan is None:
            return None
        t0 = time.perf_counter()
        if isinstance(plan, DecodePlan):
            B = plan.batch
            s = self.sched
            temps = s.temperature[:B]
            seeds = step_seeds(s.sample_seed[:B], s.num_computed[:B])
            if self.tp:
                tb = time.perf_counter()
                payload = np.concatenate([s.last_token[:B], s.num_computed[:B], temps.view(np.int32), seeds.view(np.int32),
                                          s.block_table[:B].reshape(-1)])
                self.tp.bcast_step([_KIND_DECODE, B], payload.astype(np.int32))
                self.step_time["decode_bcast"] = self.step_time.get("decode_bcast", 0.0) + time.perf_counter() - tb
            toks = self.runner.decode(B, s.last_token, s.num_computed, s.block_table, temps, seeds)
            toks = self._apply_sampling(list(s.rows[:B]), toks)
            out = s.update_decode(plan, toks)
            self.step_time["decode"] += time.perf_counter() - t0
        elif isinstance(plan, SpecPlan):
            # speculative verification: row i runs [last_token, drafts...] at positions num_computed[i]..
            # through the chunked-prefill path, every row of every chunk is sampled (spec_decode.py)
            s = self.sched
            B = plan.batch
            chunks = [[int(s.last_token[i])] + plan.drafts[i] for i in range(B)]
            starts = [int(x) for x in s.num_computed[:B]]
            tables = [s.rows[i].blocks for i in range(B)]
            counts = [len(c) for c in chunks]
            stemps = np.repeat(s.temperature[:B], counts).astype(np.float32)
            pos = np.concatenate([np.arange(st, st + n) for st, n in zip(starts, counts)])
            sseeds = step_seeds(np.repeat(s.sample_seed[:B], counts), pos)
            if self.tp:
                self._bcast_prefill(chunks, starts, tables, counts, None, stemps, sseeds)
            toks = self.runner.prefill(chunks, starts, tables, counts, sample_temps=stemps, sample_seeds=sseeds)
            # nucleus rows never carry drafts (one sampled row each): re-sampled on the host like in a decode step
            toks = self._apply_sampling([s.rows[i] for i in range(B) for _ in range(counts[i])], toks)
            out = s.update_spec(plan, toks)
            self.step_time["decode"] += time.perf_counter() - t0
        else:
            mixed = isinstance(plan, MixedPlan)
            pre = plan.prefill if mixed else plan
            s = self.sched
            B = plan.batch if mixed else 0
            dec = (B, s.last_token, s.num_computed, s.block_table) if mixed else None
            chunks = [r.all_ids[st:st + n] for r, st, n in zip(pre.reqs, pre.starts, pre.lens)]
            tables = [r.blocks for r in pre.reqs]
            sampled = [r for r, smp in zip(pre.reqs, pre.sample) if smp]
            # in-kernel sampler inputs, one per sampled row: completed prompts first, then the riding decode rows
            stemps = np.asarray([kernel_temperature(r.params) for r in sampled] + s.temperature[:B].tolist(), dtype=np.float32)
            sseeds = np.concatenate([step_seeds([r.sample_seed for r in sampled], [r.prefill_target for r in sampled]),
                                     step_seeds(s.sample_seed[:B], s.num_computed[:B])]).astype(np.uint32)
            if self.tp:
                self._bcast_prefill(chunks, pre.starts, tables, pre.sample, dec, stemps, sseeds)
            toks = self.runner.prefill(chunks, pre.starts, tables, pre.sample, decode=dec, sample_temps=stemps,
                                       sample_seeds=sseeds)
            toks = self._apply_sampling(sampled + list(s.rows[:B]), toks)
            ns = len(sampled)
            if mixed:
                out = s.update_mixed(plan, toks[:ns], np.asarray(toks[ns:ns + B], dtype=np.int32))
            else:
                out = s.update_prefill(pre, toks)
            self.step_time["prefill"] += time.perf_counter() - t0
        self.counters["steps"] += 1
        self.counters["generation_tokens"] += len(out.token_ids)
        for rid, fin in zip(out.req_ids, out.finished):
            if fin is not None:
                self.counters["finished"] += 1
                self.requests.pop(rid, None)
        return out

    def _apply_sampling(self, reqs, toks):
        """Plain temperature sampling happens in the sampler kernel (Gumbel-max, runner._sample); only
        NUCLEUS requests (temperature > 0 and top_p < 1) are re-sampled here from the step's logits."""
        rows = [i for i, r in enumerate(reqs) if r.params.temperature > 0.0 and r.params.top_p < 1.0]
        if not rows:
            return toks
        if self.tp is not None:      # every rank joins the all-gather of the vocab-sharded logits rows
            self.tp.bcast_step([_KIND_GATHER, len(rows)], np.asarray(rows, dtype=np.int32))
        gens = []
        for i in rows:
            r = reqs[i]
            if r.params.seed is not None and not hasattr(r, "_gen"):
                r._gen = torch.Generator(device=self.runner.device)
                r._gen.manual_seed(int(r.params.seed))
            gens.append(getattr(r, "_gen", None))
        sampled = self.runner.sample_rows(rows, [reqs[i].params.temperature for i in rows],
                                          [reqs[i].params.top_p for i in rows], gens)
        toks = np.array(toks, dtype=np.int32, copy=True)
        toks[rows] = sampled
        return toks

    def _bcast_prefill(self, chunks, starts, tables, sample, dec=None, stemps=None, sseeds=None) -> None:
        S = len(chunks)
        lens = [len(c) for c in chunks]
        nb = [len(t) for t in tables]
        parts = [np.asarray(lens), np.asarray(starts), np.asarray(sample, dtype=np.int32), np.asarray(nb)] + \
            [np.asarray(t) for t in tables] + [np.asarray(c) for c in chunks]
        B = 0
        if dec is not None:     # mixed step: the decode rows' (last token, position, block table) follow
            B, last, comp, bt = dec
            parts += [last[:B], comp[:B], bt[:B].reshape(-1)]
        ns = 0 if stemps is None else len(stemps)
        if ns:
            parts += [np.asarray(stemps, dtype=np.float32).view(np.int32), np.asarray(sseeds, dtype=np.uint32).view(np.int32)]
        flat = np.concatenate([np.asarray(p).reshape(-1) for p in parts]).astype(np.int32)
        self.tp.bcast_step([_KIND_MIXED if dec is not None else _KIND_PREFILL, S, B, ns], flat)

    def shutdown(self) -> None:
        if self.tp and self.tp.rank == 0:
            self.tp.bcast_step([_KIND_STOP], None)
            if self.tp.shm is not None:
                time.sleep(0.05)            # let the workers read the STOP before the segment goes
                self.tp.shm.close()
                self.tp.shm = None

    # ------------------------------------------------------------------ timing helpers (bench.py)
    def sync_mark(self) -> float:
        """Device-synchronise every rank, barrier, and remember the time (rank 0 drives)."""
        if self.tp and self.tp.rank == 0:
            self.tp.bcast_step([_KIND_SYNC], None)
        return self._do_sync_mark()

    def _do_sync_mark(self) -> float:
        torch.cuda.synchronize()
        if self.tp:
            self.tp.dist.barrier(group=self.tp.cpu)
        t = time.perf_counter()
        self._marks = getattr(self, "_marks", []) + [t]
        return t

    def max_elapsed(self) -> float:
        """max over ranks of the time between the last two sync marks (seconds)."""
        if self.tp and self.tp.rank == 0:
            self.tp.bcast_step([_KIND_MAXELAPSED], None)
        return self._do_max_elapsed()

    def _do_max_elapsed(self) -> float:
        e = torch.tensor([self._marks[-1] - self._marks[-2]], dtype=torch.float64)
        if self.tp:
            self.tp.dist.all_reduce(e, op=self.tp.dist.ReduceOp.MAX, group=self.tp.cpu)
        return float(e[0])

    # ------------------------------------------------------------------ offline API
    def generate(self, prompts: Sequence[Sequence[int]], params: SamplingParams | Sequence[SamplingParams]) -> List[List[int]]:
        """Run the given token-id prompts to completion (continuous batching); returns output ids."""
        reqs = []
        for i, p in enumerate(prompts):
            sp = params if isinstance(params, SamplingParams) else params[i]
            rid = f"gen-{self.counters['finished']}-{i}-{time.monotonic_ns()}"
            self.add_request(rid, p, sp)
            reqs.append(self.requests[rid])
        while self.has_work():
            self.step()
        return [r.output_ids for r in reqs]


class TPWorker:
    """Ranks > 0 of a tensor-parallel group: replay rank 0's steps on the local shard."""

    def __init__(self, engine: LLMEngine):
        self.engine = engine

    def serve_forever(self) -> None:
        tp, run = self.engine.tp, self.engine.runner
        mb = run.max_blocks
        while True:
            h, p = tp.bcast_step([0], None)
            kind = h[0]
            if kind == _KIND_STOP:
                if tp.shm is not None:
                    tp.shm.close()
                    tp.shm = None
                return
            if kind == _KIND_SYNC:
                self.engine._do_sync_mark()
                continue
            if kind == _KIND_MAXELAPSED:
                self.engine._do_max_elapsed()
                continue
            if kind == _KIND_GATHER:
                run.gather_logits_rows(p[:h[1]].tolist())     # rank 0 samples; this rank only contributes
                continue
            if kind == _KIND_SWAP:
                run.kv_swap(p[:3 * h[1]].reshape(-1, 3).tolist())
                continue
            if kind == _KIND_DECODE:
                B = h[1]
                run.decode(B, p[:B], p[B:2 * B], p[4 * B:].reshape(B, mb), p[2 * B:3 * B].view(np.float32),
                           p[3 * B:4 * B].view(np.uint32))
            else:
                S = h[1]
                lens, starts, sample, nb = p[:S], p[S:2 * S], p[2 * S:3 * S], p[3 * S:4 * S]
                off = 4 * S
                tables = []
                for s in range(S):
                    tables.append(p[off:off + nb[s]]); off += nb[s]
                chunks = []
                for s in range(S):
                    chunks.append(p[off:off + lens[s]]); off += lens[s]
                dec = None
                if kind == _KIND_MIXED:
                    B = h[2]
                    dec = (B, p[off:off + B], p[off + B:off + 2 * B], p[off + 2 * B:off + 2 * B + B * mb].reshape(B, mb))
                    off += 2 * B + B * mb
                ns = h[3]
                st_ = p[off:off + ns].view(np.float32) if ns else None
                ss_ = p[off + ns:off + 2 * ns].view(np.uint32) if ns else None
                run.prefill(chunks, starts.tolist(), tables, [int(x) for x in sample], decode=dec, sample_temps=st_,
                            sample_seeds=ss_)
