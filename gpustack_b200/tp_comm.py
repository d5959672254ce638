"""Peer-memory plumbing for the fused one-shot all-reduce (csrc/allreduce.cu).

Each rank owns two bf16 staging buffers (the row-parallel GEMMs write their partial [T,H] result
straight into them; O-proj and down-proj alternate) and one flag buffer.  The buffers are shared
with the other ranks of the node through CUDA IPC (torch's own tensor-sharing reduction, which
opens the handles with peer access enabled), so a kernel on rank r can load every peer's partials
over NVLink.  Counterpart of the reference engine's CustomAllreduce buffer registration
($VLLM/distributed/device_communicators/custom_all_reduce.py:50-58).
"""
from __future__ import annotations

import ctypes as C
from typing import List

import torch

from .lib import check, load

MAX_WORLD = 8


class OneShotAllReduce:
    def __init__(self, tp, max_tokens: int, hidden: int, device: torch.device):
        from torch.multiprocessing.reductions import reduce_tensor

        self.tp, self.world, self.rank = tp, tp.world, tp.rank
        self.max_tokens, self.H = max_tokens, hidden
        lib = load()
        self.bufs = [torch.zeros(max_tokens * hidden, dtype=torch.bfloat16, device=device) for _ in range(2)]
        self.flags = torch.zeros(lib.b200_allreduce_flag_words(self.world), dtype=torch.int32, device=device)
        torch.cuda.synchronize()
        mine = [reduce_tensor(t) for t in (*self.bufs, self.flags)]
        gathered: List = [None] * self.world
        tp.dist.all_gather_object(gathered, mine, group=tp.cpu)
        self._peer_tensors = []          # keep the rebuilt tensors alive (they own the IPC mappings)
        self._data_ptrs = [(C.c_void_p * MAX_WORLD)() for _ in range(2)]
        self._flag_ptrs = (C.c_void_p * MAX_WORLD)()
        for r in range(self.world):
            if r == self.rank:
                tensors = [*self.bufs, self.flags]
            else:
                tensors = [fn(*args) for fn, args in gathered[r]]
                self._peer_tensors.append(tensors)
            self._data_ptrs[0][r] = tensors[0].data_ptr()
            self._data_ptrs[1][r] = tensors[1].data_ptr()
            self._flag_ptrs[r] = tensors[2].data_ptr()
        tp.dist.barrier(group=tp.cpu)

    def stage(self, which: int, T: int) -> torch.Tensor:
        """[T,H] view of staging buffer `which` -- the GEMM writes its partial result here."""
        return self.bufs[which][: T * self.H].view(T, self.H)

    def allreduce_add_rms_norm(self, which: int, residual, w, y, residual_out, T: int, eps: float) -> None:
        """y = rmsnorm(sum_r stage_r + residual) * w, residual_out = sum + residual (w=None: plain sum)."""
        p = lambda t: None if t is None else t.data_ptr()
        check(load().b200_allreduce_add_rms_norm(self._data_ptrs[which], self._flag_ptrs, self.world, self.rank, 0,
                                                 p(residual), p(w), p(y), p(residual_out), T, self.H, eps,
                                                 torch.cuda.current_stream().cuda_stream), "allreduce_add_rms_norm")
