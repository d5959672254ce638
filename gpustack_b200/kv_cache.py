"""Paged KV-cache bookkeeping (host side): a block pool with optional prefix caching.

Counterpart of the reference engine's BlockPool / KVCacheManager
($VLLM/v1/core/block_pool.py, $VLLM/v1/core/kv_cache_manager.py, hashing in
$VLLM/v1/core/kv_cache_utils.py `hash_block_tokens`).  Blocks hold 16 tokens; the device tensors
are [num_blocks][Hkv][16][D] per layer for K and for V (see csrc/attention.cu).  Block 0 is
reserved as the "null" block that padded block-table entries point to.

Prefix caching (GPUStack passes `--enable-prefix-caching` straight through,
gpustack/worker/backends/vllm.py:602-639; on by default in the reference engine): every FULL block
is keyed by a chained digest of (parent digest, its 16 token ids).  A finished or preempted
sequence leaves its blocks in the free queue *with* their keys, so a later request whose prompt
starts with the same tokens re-attaches them (ref-count + 1) and starts its prefill behind them.
The free queue is LRU: a block loses its key only when it is handed out again.

Extended KV cache (GPUStack's `extended_kv_cache`, which its vLLM launcher turns into
`--kv-transfer-config {"kv_connector":"LMCacheConnectorV1",...}` + `LMCACHE_MAX_LOCAL_CPU_SIZE`,
gpustack/worker/backends/vllm.py:256-275,468-483): an optional SECOND TIER in pinned host memory.  A keyed
block that is evicted from the device pool is copied to a host slot first; a later prompt whose chain of
digests continues into the host tier gets those blocks copied back instead of recomputed.  This module
only does the bookkeeping and records the copies as an ordered list of swap operations; the engine executes
them on the step's stream before the step's kernels (runner.kv_swap).
"""
from __future__ import annotations

import hashlib
from collections import OrderedDict
from typing import Dict, List, Optional, Sequence

import numpy as np

_ROOT = b"b200-kv-root"


def block_digests(token_ids: Sequence[int], block_size: int, have: Optional[List[bytes]] = None,
                  upto_blocks: Optional[int] = None) -> List[bytes]:
    """Chained 16-byte digests of the full blocks of `token_ids`; extends `have` in place."""
    out = have if have is not None else []
    n_full = len(token_ids) // block_size
    if upto_blocks is not None:
        n_full = min(n_full, upto_blocks)
    if len(out) >= n_full:
        return out
    arr = np.asarray(token_ids[len(out) * block_size: n_full * block_size], dtype=np.int32)
    parent = out[-1] if out else _ROOT
    for i in range(n_full - len(out)):
        h = hashlib.blake2b(parent, digest_size=16)
        h.update(arr[i * block_size:(i + 1) * block_size].tobytes())
        parent = h.digest()
        out.append(parent)
    return out


class BlockPool:
    def __init__(self, num_blocks: int, block_size: int = 16, enable_prefix_caching: bool = False,
                 host_slots: int = 0):
        if num_blocks < 2:
            raise ValueError("need at least 2 KV blocks")
        self.num_blocks = num_blocks
        self.block_size = block_size
        self.caching = enable_prefix_caching
        # host tier: digest -> slot (LRU, oldest first), free slots, ordered swap operations of this step
        self.host_slots = host_slots if enable_prefix_caching else 0
        self._host: "OrderedDict[bytes, int]" = OrderedDict()
        self._host_free: List[int] = list(range(self.host_slots - 1, -1, -1))
        self._host_pinned: Dict[bytes, int] = {}
        self.swap_ops: List[tuple] = []          # (SWAP_OUT | SWAP_IN, device block, host slot), in execution order
        self.host_hits = 0                       # tokens served from the host tier
        self.host_stored = 0                     # blocks ever copied out
        # free queue, head = next block handed out; block 0 = null block, never in the queue
        self._free: "OrderedDict[int, None]" = OrderedDict((b, None) for b in range(1, num_blocks))
        self._ref = [0] * num_blocks
        self._key: List[Optional[bytes]] = [None] * num_blocks     # digest a block is cached under
        self._cached: Dict[bytes, int] = {}
        self.queries = 0          # tokens looked up / found, for vllm:prefix_cache_{queries,hits}
        self.hits = 0

    @property
    def num_free(self) -> int:
        return len(self._free)

    @property
    def num_total(self) -> int:
        return self.num_blocks - 1

    @property
    def num_cached(self) -> int:
        return len(self._cached)

    def usage(self) -> float:
        return 1.0 - self.num_free / max(1, self.num_total)

    def can_allocate(self, n: int) -> bool:
        return n <= len(self._free)

    def allocate(self, n: int) -> List[int]:
        if n > len(self._free):
            raise MemoryError(f"KV cache exhausted: want {n} blocks, {len(self._free)} free")
        out = []
        for _ in range(n):
            b, _ = self._free.popitem(last=False)
            k = self._key[b]
            if k is not None:                      # evicted from the prefix cache
                del self._cached[k]
                self._key[b] = None
                if self.host_slots:
                    self._offload(b, k)
            self._ref[b] = 1
            out.append(b)
        return out

    # ------------------------------------------------------------------ host tier
    SWAP_OUT, SWAP_IN = 0, 1

    def _offload(self, block: int, key: bytes) -> None:
        """The device copy of `key` is about to be overwritten: keep it in a host slot."""
        if key in self._host:
            self._host.move_to_end(key)          # an older host copy is still valid
            return
        if key in self._host_pinned:
            return
        if self._host_free:
            slot = self._host_free.pop()
        elif self._host:
            _, slot = self._host.popitem(last=False)     # host LRU victim
        else:
            return                               # every slot is pinned by loads of this step: drop the block
        self._host[key] = slot
        self.swap_ops.append((self.SWAP_OUT, block, slot))
        self.host_stored += 1

    def match_tiers(self, digests: Sequence[bytes]) -> List[tuple]:
        """Longest run of cached blocks over BOTH tiers: [(True, device block) | (False, digest)] (no state change)."""
        out: List[tuple] = []
        if not self.caching:
            return out
        for d in digests:
            b = self._cached.get(d)
            if b is not None:
                out.append((True, b))
            elif d in self._host:
                out.append((False, d))
            else:
                break
        return out

    def pin_host(self, digests: Sequence[bytes]) -> None:
        """Host entries about to be loaded must survive the evictions of the allocation that follows."""
        for d in digests:
            self._host_pinned[d] = self._host.pop(d)

    def load_from_host(self, digest: bytes, block: int) -> None:
        """Queue the copy host slot -> device `block` and put the (pinned) entry back as most recently used."""
        slot = self._host_pinned.pop(digest)
        self._host[digest] = slot
        self.swap_ops.append((self.SWAP_IN, block, slot))
        self.host_hits += self.block_size

    def take_swap_ops(self) -> List[tuple]:
        ops, self.swap_ops = self.swap_ops, []
        return ops

    def free(self, blocks: List[int]) -> None:
        if self.caching:
            # tail of the sequence first: it is evicted before the (more widely shared) head
            for b in reversed(blocks):
                self._ref[b] -= 1
                if self._ref[b] == 0:
                    self._free[b] = None
            return
        # no caching: most recently used blocks go back to the head so that hot L2 lines are reused
        for b in reversed(blocks):
            self._ref[b] -= 1
            if self._ref[b] == 0:
                self._free[b] = None
                self._free.move_to_end(b, last=False)

    # ------------------------------------------------------------------ prefix cache
    def match_prefix(self, digests: Sequence[bytes]) -> List[int]:
        """Longest run of cached blocks for the chained `digests` (no state change)."""
        hit: List[int] = []
        if not self.caching:
            return hit
        for d in digests:
            b = self._cached.get(d)
            if b is None:
                break
            hit.append(b)
        return hit

    def count_free(self, blocks: Sequence[int]) -> int:
        """How many of `blocks` currently sit in the free queue (attaching them shrinks it)."""
        return sum(1 for b in blocks if self._ref[b] == 0)

    def attach(self, blocks: Sequence[int]) -> None:
        """Take a reference on cached blocks returned by match_prefix."""
        for b in blocks:
            if self._ref[b] == 0:
                del self._free[b]
            self._ref[b] += 1

    def register(self, blocks: Sequence[int], digests: Sequence[bytes]) -> None:
        """Publish full, already-written blocks under their digests (first writer wins)."""
        if not self.caching:
            return
        for b, d in zip(blocks, digests):
            if self._key[b] is not None or d in self._cached:
                continue
            self._key[b] = d
            self._cached[d] = b

    def reset_prefix_cache(self) -> None:
        for d, b in self._cached.items():
            self._key[b] = None
        self._cached.clear()
        self._host_free.extend(self._host.values())
        self._host.clear()


class SequenceBlocks:
    """Block table of one sequence."""

    __slots__ = ("blocks", "num_tokens")

    def __init__(self):
        self.blocks: List[int] = []
        self.num_tokens = 0          # tokens whose K/V are (or are being) written

    def blocks_needed(self, new_total_tokens: int, block_size: int) -> int:
        need = (new_total_tokens + block_size - 1) // block_size
        return max(0, need - len(self.blocks))


class KVCacheManager:
    """Stand-alone block-table manager over a BlockPool (used by kernel tests and tools; the
    scheduler keeps its tables in the persistent batch instead)."""

    def __init__(self, num_blocks: int, block_size: int = 16):
        self.pool = BlockPool(num_blocks, block_size)
        self.block_size = block_size
        self.seqs: Dict[int, SequenceBlocks] = {}

    def can_append(self, seq_id: int, new_total_tokens: int, reserve: int = 0) -> bool:
        sb = self.seqs.get(seq_id)
        have = len(sb.blocks) if sb else 0
        need = (new_total_tokens + self.block_size - 1) // self.block_size - have
        return max(0, need) + reserve <= self.pool.num_free

    def append(self, seq_id: int, new_total_tokens: int) -> SequenceBlocks:
        """Grow the sequence's table so that it covers `new_total_tokens` tokens."""
        sb = self.seqs.setdefault(seq_id, SequenceBlocks())
        n = sb.blocks_needed(new_total_tokens, self.block_size)
        if n:
            sb.blocks.extend(self.pool.allocate(n))
        sb.num_tokens = new_total_tokens
        return sb

    def slot(self, seq_id: int, pos: int) -> int:
        sb = self.seqs[seq_id]
        return sb.blocks[pos // self.block_size] * self.block_size + pos % self.block_size

    def free(self, seq_id: int) -> None:
        sb = self.seqs.pop(seq_id, None)
        if sb is not None:
            self.pool.free(sb.blocks)

    def usage(self) -> float:
        return self.pool.usage()
