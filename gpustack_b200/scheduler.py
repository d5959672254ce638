"""Continuous-batching scheduler (host side).

Counterpart of the reference engine's Scheduler.schedule ($VLLM/v1/core/sched/scheduler.py:329):
token-budgeted steps: the running sequences' decode tokens first, the rest of the budget for prefill
chunks of admitted / waiting prompts (with chunking of long prompts) -- ONE mixed step serves both, so
running requests keep producing a token per step while new prompts are prefilled (the reference
scheduler's unified token budget).  Steps without prompts to prefill are pure decode steps (CUDA-graph
path); block allocation / prefix-cache reuse / preemption go through the BlockPool.
`mixed_steps=False` restores the round-1 policy (prefill steps never carry decode tokens).

Running sequences live in a compact "persistent batch" of numpy arrays (rows 0..B-1) so that a
decode step's inputs are produced by a handful of vectorised numpy ops instead of per-request
Python.
"""
from __future__ import annotations

import time
from collections import deque
from dataclasses import dataclass, field
from typing import Deque, Dict, List, Optional, Sequence, Tuple

import numpy as np

from .kv_cache import BlockPool, block_digests
from .spec_decode import NgramIndex, SpeculativeConfig, verify


@dataclass
class SamplingParams:
    max_tokens: int = 16
    temperature: float = 0.0
    top_p: float = 1.0
    ignore_eos: bool = False
    stop_token_ids: Tuple[int, ...] = ()
    seed: Optional[int] = None


def kernel_temperature(p: "SamplingParams") -> float:
    """Temperature handed to the in-kernel (Gumbel-max) sampler: plain temperature sampling is done
    there; greedy rows and nucleus rows (top_p < 1: re-sampled on the host from the logits) get 0."""
    return float(p.temperature) if (p.temperature > 0.0 and p.top_p >= 1.0) else 0.0


def step_seeds(seeds, positions) -> np.ndarray:
    """Noise stream of one sampling step: the request's 32-bit seed mixed with the position of the token
    being sampled (same request + same position = same draw, whatever the batch composition)."""
    s = np.asarray(seeds, dtype=np.uint64) * np.uint64(0x9E3779B1) + np.asarray(positions, dtype=np.uint64) * np.uint64(0x85EBCA77)
    return (s & np.uint64(0xFFFFFFFF)).astype(np.uint32)


@dataclass
class Request:
    req_id: str
    prompt_ids: List[int]
    params: SamplingParams
    arrival: float = field(default_factory=time.monotonic)
    output_ids: List[int] = field(default_factory=list)
    num_computed: int = 0          # prompt tokens whose K/V are in the cache (prefill progress)
    row: int = -1                  # row in the persistent batch, -1 = not resident
    blocks: List[int] = field(default_factory=list)
    first_token_time: Optional[float] = None
    finished: bool = False
    finish_reason: Optional[str] = None
    num_preemptions: int = 0
    sample_seed: int = 0           # 32-bit stream id of the in-kernel sampler (params.seed, else drawn at admission)
    digests: List[bytes] = field(default_factory=list)   # chained keys of the full KV blocks
    num_cached_tokens: int = 0     # prompt tokens served from the prefix cache at admission

    @property
    def all_ids(self) -> List[int]:
        return self.prompt_ids + self.output_ids

    @property
    def prefill_target(self) -> int:
        # after a preemption the generated tokens are recomputed as part of the prompt
        return len(self.prompt_ids) + len(self.output_ids)


@dataclass
class PrefillPlan:
    reqs: List[Request]
    starts: List[int]              # first position of each chunk
    lens: List[int]                # tokens in each chunk
    sample: List[bool]             # chunk completes the prompt -> sample a token

    @property
    def num_tokens(self) -> int:
        return sum(self.lens)


@dataclass
class DecodePlan:
    batch: int                     # rows 0..batch-1 of the persistent batch


@dataclass
class MixedPlan:
    """One step = prefill chunks + one decode token for rows 0..batch-1 of the persistent batch."""
    prefill: PrefillPlan
    batch: int

    @property
    def num_tokens(self) -> int:
        return self.prefill.num_tokens + self.batch


@dataclass
class SpecPlan:
    """Speculative decode step: every resident row runs `[last_token, draft_1 .. draft_k]` (k may be 0)
    through one forward; all rows of every chunk are sampled (spec_decode.py)."""
    batch: int
    drafts: List[List[int]]        # per row 0..batch-1

    @property
    def num_tokens(self) -> int:
        return self.batch + sum(len(d) for d in self.drafts)


@dataclass
class StepOutput:
    req_ids: List[str]
    token_ids: List[int]
    finished: List[Optional[str]]  # finish reason or None


class Scheduler:
    def __init__(self, num_blocks: int, block_size: int, max_num_seqs: int, max_num_batched_tokens: int,
                 max_model_len: int, eos_token_id, enable_prefix_caching: bool = False, mixed_steps: bool = True,
                 speculative: Optional[SpeculativeConfig] = None, max_sample_rows: Optional[int] = None,
                 kv_host_slots: int = 0):
        """eos_token_id: one id, an iterable of ids (every turn terminator of the checkpoint), or None.
        speculative: n-gram speculative decoding on pure decode steps (spec_decode.py); max_sample_rows
        bounds the rows one step may sample (the runner's logits buffer)."""
        self.pool = BlockPool(num_blocks, block_size, enable_prefix_caching, host_slots=kv_host_slots)
        self.block_size = block_size
        self.max_num_seqs = max_num_seqs
        self.max_tokens_per_step = max_num_batched_tokens
        self.mixed_steps = mixed_steps
        self.max_model_len = max_model_len
        if eos_token_id is None:
            eos_ids = []
        elif isinstance(eos_token_id, (int, np.integer)):
            eos_ids = [int(eos_token_id)]
        else:
            eos_ids = [int(t) for t in eos_token_id]
        self.eos_ids = frozenset(eos_ids)
        self.eos = eos_ids[0] if eos_ids else -1
        self._eos_arr = np.asarray(sorted(self.eos_ids), dtype=np.int32)
        self.max_blocks_per_seq = (max_model_len + block_size - 1) // block_size
        self.waiting: Deque[Request] = deque()
        self.prefilling: List[Request] = []     # admitted, prompt not fully computed yet
        self.rows: List[Optional[Request]] = []  # decode-resident requests, compact
        S = max_num_seqs
        self.last_token = np.zeros(S, dtype=np.int32)
        self.num_computed = np.zeros(S, dtype=np.int32)
        self.remaining = np.zeros(S, dtype=np.int32)
        self.check_eos = np.zeros(S, dtype=bool)
        self.n_blocks = np.zeros(S, dtype=np.int32)
        self.block_table = np.zeros((S, self.max_blocks_per_seq), dtype=np.int32)
        # in-kernel sampler inputs of the resident rows: temperature (0 = greedy; nucleus rows are left
        # greedy here and re-sampled on the host) and the per-request noise stream
        self.temperature = np.zeros(S, dtype=np.float32)
        self.sample_seed = np.zeros(S, dtype=np.uint32)
        self._seed_rng = np.random.RandomState(0x5EED)
        self.num_preemptions = 0
        self.spec = speculative if (speculative is not None and speculative.supported) else None
        self.max_sample_rows = max_sample_rows if max_sample_rows is not None else max_num_seqs
        self.proposer = self._ngram_propose      # (request, k) -> draft token ids; tests substitute their own
        self.spec_stats = {"steps": 0, "drafted": 0, "accepted": 0}

    # ------------------------------------------------------------------ queue side
    def add(self, req: Request) -> None:
        total = len(req.prompt_ids) + req.params.max_tokens
        if len(req.prompt_ids) == 0:
            raise ValueError("empty prompt")
        if total > self.max_model_len:
            raise ValueError(f"prompt ({len(req.prompt_ids)}) + max_tokens ({req.params.max_tokens}) "
                             f"exceeds max_model_len ({self.max_model_len})")
        req.sample_seed = (int(req.params.seed) if req.params.seed is not None
                           else int(self._seed_rng.randint(0, 2 ** 31))) & 0xFFFFFFFF
        self.waiting.append(req)

    def abort(self, req_id: str) -> bool:
        for q in (self.waiting,):
            for r in list(q):
                if r.req_id == req_id:
                    q.remove(r)
                    r.finished, r.finish_reason = True, "abort"
                    return True
        for r in list(self.prefilling):
            if r.req_id == req_id:
                self.prefilling.remove(r)
                self._release(r)
                r.finished, r.finish_reason = True, "abort"
                return True
        for r in self.rows:
            if r is not None and r.req_id == req_id:
                self._evict_row(r.row)
                self._release(r)
                r.finished, r.finish_reason = True, "abort"
                return True
        return False

    @property
    def num_running(self) -> int:
        return len(self.rows) + len(self.prefilling)

    @property
    def num_waiting(self) -> int:
        return len(self.waiting)

    def has_work(self) -> bool:
        return bool(self.waiting or self.prefilling or self.rows)

    def kv_usage(self) -> float:
        return self.pool.usage()

    # ------------------------------------------------------------------ scheduling
    def _blocks_for(self, ntok: int) -> int:
        return (ntok + self.block_size - 1) // self.block_size

    def _publish(self, r: Request) -> None:
        """Key the request's full, already-written blocks in the prefix cache."""
        if not self.pool.caching or not r.blocks:
            return
        n_full = min(r.num_computed // self.block_size, len(r.blocks))
        if n_full > len(r.digests):
            block_digests(r.all_ids, self.block_size, r.digests, upto_blocks=n_full)
        self.pool.register(r.blocks[:n_full], r.digests[:n_full])

    def _release(self, r: Request) -> None:
        if r.blocks:
            self._publish(r)
            self.pool.free(r.blocks)
            r.blocks = []

    def schedule(self):
        """Returns a PrefillPlan, a MixedPlan, a DecodePlan, or None when idle."""
        ride = len(self.rows) if self.mixed_steps else 0     # decode tokens come out of the same budget
        plan = self._schedule_prefill(self.max_tokens_per_step - ride) if ride < self.max_tokens_per_step else None
        if plan is not None:
            if ride:
                self._ensure_decode_blocks()                 # may preempt rows (never the ones being prefilled)
                if self.rows:
                    return MixedPlan(plan, len(self.rows))
            return plan
        if self.rows:
            self._ensure_decode_blocks()
            if self.rows:
                if self.spec is not None and len(self.rows) < self.spec.disable_by_batch_size:
                    sp = self._schedule_spec()
                    if sp is not None:
                        return sp
                return DecodePlan(batch=len(self.rows))
        return None

    # ------------------------------------------------------------------ speculative decoding
    def _ngram_propose(self, r: Request, k: int) -> List[int]:
        idx = getattr(r, "_ngram", None)
        n_ids = len(r.prompt_ids) + len(r.output_ids)
        if idx is None:
            idx = r._ngram = NgramIndex(self.spec.prompt_lookup_min, self.spec.prompt_lookup_max)
        if len(idx.ids) > n_ids:
            idx.truncate(0)
        if len(idx.ids) < n_ids:
            idx.extend(r.all_ids[len(idx.ids):])
        return idx.propose(k)

    def _schedule_spec(self) -> Optional[SpecPlan]:
        """Drafts for the resident rows (rows already hold a block for position num_computed).  A draft
        is cut to what the request may still emit, to the model length, to the step's token / sample-row
        budgets and to the KV blocks that are free right now -- speculation never preempts."""
        B = len(self.rows)
        k_max = self.spec.num_speculative_tokens
        tok_budget = min(self.max_tokens_per_step, self.max_sample_rows) - B
        drafts: List[List[int]] = []
        any_draft = False
        for i, r in enumerate(self.rows):
            d: List[int] = []
            p = r.params
            room = min(k_max, int(self.remaining[i]) - 1, tok_budget,
                       self.max_model_len - 1 - int(self.num_computed[i]) - 1)
            if room > 0 and not (p.temperature > 0.0 and p.top_p < 1.0):     # nucleus rows are sampled on the host
                d = [int(t) for t in self.proposer(r, room)[:room]]
            if d:
                last_pos = int(self.num_computed[i]) + len(d)
                need = last_pos // self.block_size + 1 - int(self.n_blocks[i])
                if need > 0:
                    if need > self.pool.num_free:
                        d = d[:max(0, (int(self.n_blocks[i]) * self.block_size - 1) - int(self.num_computed[i]))]
                    else:
                        for b in self.pool.allocate(need):
                            r.blocks.append(b)
                            self.block_table[i, self.n_blocks[i]] = b
                            self.n_blocks[i] += 1
            tok_budget -= len(d)
            any_draft |= bool(d)
            drafts.append(d)
        if not any_draft:
            return None
        return SpecPlan(B, drafts)

    def update_spec(self, plan: SpecPlan, sampled: Sequence[int]) -> StepOutput:
        """`sampled` holds, row after row, the tokens sampled at every position of the row's chunk
        (1 + len(draft) each).  Emits 1 + accepted tokens per row, stopping at a finish condition."""
        B = plan.batch
        out = StepOutput([], [], [])
        rows = list(self.rows[:B])
        now_done: List[Tuple[Request, str]] = []
        off = 0
        self.spec_stats["steps"] += 1
        for i, r in enumerate(rows):
            d = plan.drafts[i]
            got = [int(t) for t in sampled[off:off + 1 + len(d)]]
            off += 1 + len(d)
            a = verify(d, got)
            self.spec_stats["drafted"] += len(d)
            emit = got[:a + 1]
            reason = None
            n_emit = 0
            for tok in emit:
                r.output_ids.append(tok)
                n_emit += 1
                out.req_ids.append(r.req_id); out.token_ids.append(tok); out.finished.append(None)
                reason = self._finish_reason(r, tok)
                if reason is not None:
                    out.finished[-1] = reason
                    break
            self.spec_stats["accepted"] += max(0, n_emit - 1)
            self.last_token[i] = emit[n_emit - 1]
            self.num_computed[i] += n_emit       # K/V of last_token and of the accepted drafts are in the cache
            self.remaining[i] -= n_emit
            if reason is not None:
                now_done.append((r, reason))
        for r, reason in now_done[::-1]:          # rows only move down: evict from the highest row first
            r.finished, r.finish_reason = True, reason
            self._evict_row(r.row)
            self._release(r)
        return out

    def _schedule_prefill(self, budget: Optional[int] = None) -> Optional[PrefillPlan]:
        if budget is None:
            budget = self.max_tokens_per_step
        reqs, starts, lens, sample = [], [], [], []
        # 1. continue partially prefilled prompts
        for r in self.prefilling:
            if budget <= 0:
                break
            n = min(budget, r.prefill_target - r.num_computed)
            reqs.append(r); starts.append(r.num_computed); lens.append(n)
            sample.append(r.num_computed + n == r.prefill_target)
            budget -= n
        # 2. admit new requests while sequence slots, tokens and KV blocks last
        while self.waiting and budget > 0:
            r = self.waiting[0]
            if len(self.rows) + len(self.prefilling) >= self.max_num_seqs:
                break
            target = r.prefill_target
            entries: List[tuple] = []        # cached prefix, block by block: (True, device block) | (False, digest in the host tier)
            if self.pool.caching:
                # the last token is always recomputed: its logits are what gets sampled
                block_digests(r.all_ids, self.block_size, r.digests, upto_blocks=(target - 1) // self.block_size)
                entries = self.pool.match_tiers(r.digests[: (target - 1) // self.block_size])
            hit = [b for on_dev, b in entries if on_dev]
            need = self._blocks_for(target + 1) - len(hit)      # host-tier hits need a device block each
            # keep one spare block per running sequence so that decode does not preempt at once
            if need + self.pool.count_free(hit) + len(self.rows) > self.pool.num_free:
                break
            self.waiting.popleft()
            self.pool.attach(hit)
            self.pool.pin_host([d for on_dev, d in entries if not on_dev])
            fresh = iter(self.pool.allocate(need))
            r.blocks = []
            for on_dev, x in entries:
                if on_dev:
                    r.blocks.append(x)
                else:                          # copied back from host memory before this step's kernels run
                    b = next(fresh)
                    self.pool.load_from_host(x, b)
                    r.blocks.append(b)
            r.blocks.extend(fresh)
            r.num_computed = len(entries) * self.block_size
            if self.pool.caching:
                self.pool.queries += target
                self.pool.hits += r.num_computed
                if r.num_preemptions == 0:
                    r.num_cached_tokens = r.num_computed
            self.prefilling.append(r)
            n = min(budget, target - r.num_computed)
            reqs.append(r); starts.append(r.num_computed); lens.append(n)
            sample.append(r.num_computed + n == target)
            budget -= n
        if not reqs:
            return None
        return PrefillPlan(reqs, starts, lens, sample)

    def _ensure_decode_blocks(self) -> None:
        """Every resident row needs a slot for position num_computed; allocate, preempting the
        most recently admitted rows when the pool runs dry."""
        while True:
            B = len(self.rows)
            pos = self.num_computed[:B]
            need = (pos // self.block_size) >= self.n_blocks[:B]
            idx = np.nonzero(need)[0]
            if len(idx) <= self.pool.num_free:
                for i in idx:
                    r = self.rows[i]
                    b = self.pool.allocate(1)[0]
                    r.blocks.append(b)
                    self.block_table[i, self.n_blocks[i]] = b
                    self.n_blocks[i] += 1
                return
            self._preempt_last()
            if not self.rows:
                return

    def _preempt_last(self) -> None:
        r = self.rows[-1]
        self._evict_row(r.row)
        self._release(r)
        r.num_computed = 0
        r.num_preemptions += 1
        self.num_preemptions += 1
        self.waiting.appendleft(r)

    # ------------------------------------------------------------------ persistent batch
    def _install_row(self, r: Request, first_token: int) -> None:
        i = len(self.rows)
        self.rows.append(r)
        r.row = i
        self.last_token[i] = first_token
        self.num_computed[i] = r.num_computed   # K/V tokens in the cache; `first_token` sits at this position
        self.remaining[i] = r.params.max_tokens - len(r.output_ids)
        self.check_eos[i] = not r.params.ignore_eos
        self.temperature[i] = kernel_temperature(r.params)
        self.sample_seed[i] = r.sample_seed
        nb = len(r.blocks)
        self.n_blocks[i] = nb
        self.block_table[i, :nb] = r.blocks
        self.block_table[i, nb:] = 0

    def _evict_row(self, i: int) -> None:
        last = len(self.rows) - 1
        r = self.rows[i]
        r.num_computed = int(self.num_computed[i])      # tokens whose K/V are in the cache
        if i != last:
            m = self.rows[last]
            self.rows[i] = m
            m.row = i
            for a in (self.last_token, self.num_computed, self.remaining, self.check_eos, self.n_blocks,
                      self.temperature, self.sample_seed):
                a[i] = a[last]
            self.block_table[i] = self.block_table[last]
        self.rows.pop()
        r.row = -1

    # ------------------------------------------------------------------ results
    def _finish_reason(self, r: Request, tok: int) -> Optional[str]:
        p = r.params
        # ignore_eos suppresses only EOS; stop_token_ids always terminate (reference engine semantics,
        # $VLLM/v1/core/sched/utils.py check_stop)
        if (not p.ignore_eos and tok in self.eos_ids) or tok in p.stop_token_ids:
            return "stop"
        if len(r.output_ids) >= p.max_tokens:
            return "length"
        return None

    def update_prefill(self, plan: PrefillPlan, sampled: Sequence[int]) -> StepOutput:
        """`sampled[j]` is the token of the j-th plan entry with sample=True."""
        out = StepOutput([], [], [])
        now = time.monotonic()
        j = 0
        for r, n, smp in zip(plan.reqs, plan.lens, plan.sample):
            r.num_computed += n
            self._publish(r)
            if not smp:
                continue
            tok = int(sampled[j]); j += 1
            self.prefilling.remove(r)
            r.output_ids.append(tok)
            if r.first_token_time is None:
                r.first_token_time = now
            reason = self._finish_reason(r, tok)
            out.req_ids.append(r.req_id); out.token_ids.append(tok); out.finished.append(reason)
            if reason is not None:
                r.finished, r.finish_reason = True, reason
                self._release(r)
            else:
                self._install_row(r, tok)
        return out

    def update_mixed(self, plan: MixedPlan, sampled_prefill: Sequence[int], sampled_decode: np.ndarray) -> StepOutput:
        """Decode rows first (their eviction swaps only touch rows < batch), then the prefill chunks
        (completed prompts are installed as new rows at the end)."""
        out = self.update_decode(DecodePlan(plan.batch), sampled_decode)
        pre = self.update_prefill(plan.prefill, sampled_prefill)
        return StepOutput(out.req_ids + pre.req_ids, out.token_ids + pre.token_ids, out.finished + pre.finished)

    def update_decode(self, plan: DecodePlan, sampled: np.ndarray) -> StepOutput:
        B = plan.batch
        sampled = np.asarray(sampled[:B], dtype=np.int32)
        self.last_token[:B] = sampled
        self.num_computed[:B] += 1
        self.remaining[:B] -= 1
        done = self.remaining[:B] <= 0
        if len(self._eos_arr) == 1:
            stop = self.check_eos[:B] & (sampled == self.eos)
        else:
            stop = self.check_eos[:B] & np.isin(sampled, self._eos_arr)
        out = StepOutput([], sampled.tolist(), [None] * B)
        rows = list(self.rows)
        out.req_ids = [r.req_id for r in rows]
        toks = out.token_ids
        for i, r in enumerate(rows):
            r.output_ids.append(toks[i])
        fin = done | stop
        # per-request stop_token_ids are rare; only requests that set them pay for the check
        for i, r in enumerate(rows):
            p = r.params
            if p.stop_token_ids and toks[i] in p.stop_token_ids:
                fin[i] = True
                stop[i] = True
        for i in np.nonzero(fin)[0][::-1].tolist():   # descending: swaps never touch pending rows
            r = rows[i]
            reason = "stop" if stop[i] else "length"
            out.finished[i] = reason
            r.finished, r.finish_reason = True, reason
            self._evict_row(r.row)
            self._release(r)
        return out
