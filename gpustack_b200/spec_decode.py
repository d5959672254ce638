"""Speculative decoding, n-gram ("prompt lookup") method.

The reference launcher composes `--speculative-config '{"method": "ngram", "num_speculative_tokens": k,
"prompt_lookup_max": N, "prompt_lookup_min": n}'` for its vLLM backend
(gpustack/worker/backends/vllm.py:333-367, `_get_speculative_arguments`); the reference engine's
proposer is $VLLM/v1/spec_decode/ngram_proposer.py (longest suffix n-gram of the sequence that occurred
before -> the tokens that followed it are the draft).  Here:

  * drafts come from an incremental per-request index (n-gram -> position after its latest occurrence),
    O(N) dictionary operations per generated token instead of a scan of the sequence per step;
  * verification is ONE forward over `[last_token, d_1 .. d_k]` per sequence through the chunked-prefill
    path (the K/V of all k+1 positions is written, every row is sampled);
  * acceptance is exact for greedy AND for temperature sampling: the sampler kernel's Gumbel noise is a
    function of (request seed, position, vocabulary index) only (csrc/elementwise.cu), so "the token the
    target model samples at position p" is well defined before the step runs; draft d_j is kept iff it
    equals the token sampled from the logits of the row before it.  The output sequence is therefore
    the one non-speculative decoding would have produced (up to the bf16 numerics of a k+1-token
    forward vs a 1-token forward).  Rejected drafts leave garbage K/V beyond `num_computed`, which the
    next step overwrites.

Methods that need a draft model (eagle3, mtp) are not implemented: the engine then runs without
speculation and says so (speculation never changes results, only latency).
"""
from __future__ import annotations

import json
from dataclasses import asdict, dataclass
from typing import Dict, List, Optional, Sequence, Tuple


@dataclass
class SpeculativeConfig:
    method: str = "ngram"
    num_speculative_tokens: int = 3
    prompt_lookup_max: int = 4
    prompt_lookup_min: int = 2
    # like the reference engine's `disable_by_batch_size`: with this many running sequences or more the
    # step is compute-bound and verification tokens only cost throughput
    disable_by_batch_size: int = 32

    @classmethod
    def parse(cls, text) -> Optional["SpeculativeConfig"]:
        """`--speculative-config` value (JSON text or dict) -> config, None when absent / disabled.
        Raises ValueError on malformed input; unsupported methods return a config whose `supported` is False."""
        if text in (None, "", {}):
            return None
        d = json.loads(text) if isinstance(text, str) else dict(text)
        if not isinstance(d, dict):
            raise ValueError("--speculative-config must be a JSON object")
        c = cls(method=str(d.get("method", "ngram")))
        for k in ("num_speculative_tokens", "prompt_lookup_max", "prompt_lookup_min", "disable_by_batch_size"):
            if d.get(k) is not None:
                setattr(c, k, int(d[k]))
        if c.num_speculative_tokens < 1 or c.prompt_lookup_min < 1 or c.prompt_lookup_max < c.prompt_lookup_min:
            raise ValueError(f"invalid speculative config: {d}")
        return c

    @property
    def supported(self) -> bool:
        return self.method in ("ngram", "[ngram]")

    def to_dict(self) -> dict:
        return asdict(self)


class NgramIndex:
    """Incremental prompt-lookup index of ONE sequence.

    For every n in [n_min, n_max] it remembers, per n-gram, the position just after its most recent
    occurrence.  `propose(k)` looks the current suffix up (longest n first) and returns up to k tokens
    that followed that earlier occurrence -- the same draft the reference proposer finds by scanning.
    The suffix itself is only indexed once a further token arrives, so a match is always an EARLIER
    occurrence."""

    def __init__(self, n_min: int, n_max: int):
        self.n_min, self.n_max = n_min, n_max
        self.ids: List[int] = []
        self.maps: List[Dict[Tuple[int, ...], int]] = [dict() for _ in range(n_max - n_min + 1)]
        self._indexed = 0          # n-grams ending before this position are in the maps

    def extend(self, tokens: Sequence[int]) -> None:
        self.ids.extend(int(t) for t in tokens)

    def truncate(self, length: int) -> None:
        """Forget everything (a preempted request is re-indexed from its ids)."""
        if length != len(self.ids):
            self.ids = self.ids[:length]
            self.maps = [dict() for _ in self.maps]
            self._indexed = 0

    def _index_upto(self, end: int) -> None:
        ids = self.ids
        for e in range(max(self._indexed, self.n_min), end):       # n-grams ending at position e (exclusive)
            for i, m in enumerate(self.maps):
                n = self.n_min + i
                if e >= n:
                    m[tuple(ids[e - n:e])] = e
        self._indexed = max(self._indexed, end)

    def propose(self, k: int) -> List[int]:
        L = len(self.ids)
        if k <= 0 or L < self.n_min + 1:
            return []
        self._index_upto(L)                  # everything ending strictly before L: the suffix (ending AT L) is not in
        ids = self.ids
        for i in range(len(self.maps) - 1, -1, -1):
            n = self.n_min + i
            if L < n + 1:
                continue
            pos = self.maps[i].get(tuple(ids[L - n:L]))
            if pos is not None and pos < L:
                return ids[pos:pos + k]
        return []


def verify(drafts: Sequence[int], sampled: Sequence[int]) -> int:
    """sampled[j] is the target model's token after chunk row j (row 0 = the last accepted token, row j
    = draft j).  Returns the number of accepted drafts a: tokens sampled[0..a] are emitted."""
    a = 0
    while a < len(drafts) and int(drafts[a]) == int(sampled[a]):
        a += 1
    return a
