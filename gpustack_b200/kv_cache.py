"""Paged KV-cache bookkeeping (host side): a free-list block pool and per-sequence block tables.

Counterpart of the reference engine's BlockPool / KVCacheManager
($VLLM/v1/core/block_pool.py, $VLLM/v1/core/kv_cache_manager.py).  Blocks hold 16 tokens; the
device tensors are [num_blocks][Hkv][16][D] per layer for K and for V (see csrc/attention.cu).
Block 0 is reserved as the "null" block that padded block-table entries point to.
"""
from __future__ import annotations

from collections import deque
from typing import Deque, Dict, List


class BlockPool:
    def __init__(self, num_blocks: int, block_size: int = 16):
        if num_blocks < 2:
            raise ValueError("need at least 2 KV blocks")
        self.num_blocks = num_blocks
        self.block_size = block_size
        self._free: Deque[int] = deque(range(1, num_blocks))  # block 0 = null block

    @property
    def num_free(self) -> int:
        return len(self._free)

    @property
    def num_total(self) -> int:
        return self.num_blocks - 1

    def usage(self) -> float:
        return 1.0 - self.num_free / max(1, self.num_total)

    def can_allocate(self, n: int) -> bool:
        return n <= len(self._free)

    def allocate(self, n: int) -> List[int]:
        if n > len(self._free):
            raise MemoryError(f"KV cache exhausted: want {n} blocks, {len(self._free)} free")
        return [self._free.popleft() for _ in range(n)]

    def free(self, blocks: List[int]) -> None:
        # most recently used blocks go back first so that hot L2 lines are reused
        self._free.extendleft(reversed(blocks))


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
