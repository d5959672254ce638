"""Peer-memory plumbing for the fused one-shot all-reduce (csrc/allreduce.cu).

Each rank owns two bf16 staging buffers (the row-parallel GEMMs write their partial [T,H] result
straight into them; O-proj and down-proj alternate) and one flag buffer, allocated with cudaMalloc
through the C ABI and exported with CUDA IPC; every rank opens the peers' handles in ITS OWN device
context (cudaIpcMemLazyEnablePeerAccess), so its kernels can load the peers' partials over NVLink.
Counterpart of the reference engine's CustomAllreduce buffer registration
($VLLM/distributed/device_communicators/custom_all_reduce.py:50-58).
"""
from __future__ import annotations

import ctypes as C
from typing import List

import torch

from .lib import check, load

MAX_WORLD = 8


class _RawCudaBuffer:
    """Exposes a raw device allocation to torch through __cuda_array_interface__."""

    def __init__(self, ptr: int, nbytes: int, typestr: str, itemsize: int):
        self.ptr, self.nbytes = ptr, nbytes
        self.__cuda_array_interface__ = {"shape": (nbytes // itemsize,), "typestr": typestr,
                                         "data": (ptr, False), "version": 3, "strides": None}


class OneShotAllReduce:
    def __init__(self, tp, max_tokens: int, hidden: int, device: torch.device):
        self.tp, self.world, self.rank = tp, tp.world, tp.rank
        self.max_tokens, self.H = max_tokens, hidden
        self.lib = lib = load()
        nbytes = max_tokens * hidden * 2
        fbytes = lib.b200_allreduce_flag_words(self.world) * 4
        self._own: List[int] = []
        handles = []
        with torch.cuda.device(device):
            for size in (nbytes, nbytes, fbytes):
                p = C.c_void_p()
                check(lib.b200_dev_malloc(size, C.byref(p)), "dev_malloc")
                h = C.create_string_buffer(64)
                check(lib.b200_ipc_get_handle(p, h), "ipc_get_handle")
                self._own.append(p.value)
                handles.append(h.raw)
            torch.cuda.synchronize()
            gathered: List = [None] * self.world
            tp.dist.all_gather_object(gathered, handles, group=tp.cpu)
            self._opened: List[int] = []
            self._data_ptrs = [(C.c_void_p * MAX_WORLD)() for _ in range(2)]
            self._flag_ptrs = (C.c_void_p * MAX_WORLD)()
            for r in range(self.world):
                if r == self.rank:
                    ptrs = self._own
                else:
                    ptrs = []
                    for raw in gathered[r]:
                        q = C.c_void_p()
                        check(lib.b200_ipc_open_handle(C.create_string_buffer(raw, 64), C.byref(q)), "ipc_open_handle")
                        ptrs.append(q.value)
                    self._opened.extend(ptrs)
                self._data_ptrs[0][r], self._data_ptrs[1][r], self._flag_ptrs[r] = ptrs[0], ptrs[1], ptrs[2]
        # torch views of the local staging buffers (GEMM output targets)
        self.bufs = [torch.as_tensor(_RawCudaBuffer(self._own[k], nbytes, "<u2", 2), device=device).view(torch.bfloat16)
                     for k in range(2)]
        tp.dist.barrier(group=tp.cpu)

    def stage(self, which: int, T: int) -> torch.Tensor:
        """[T,H] view of staging buffer `which` -- the GEMM writes its partial result here."""
        return self.bufs[which][: T * self.H].view(T, self.H)

    def allreduce_add_rms_norm(self, which: int, residual, w, y, residual_out, T: int, eps: float) -> None:
        """y = rmsnorm(sum_r stage_r + residual) * w, residual_out = sum + residual (w=None: plain sum)."""
        p = lambda t: None if t is None else t.data_ptr()
        check(self.lib.b200_allreduce_add_rms_norm(self._data_ptrs[which], self._flag_ptrs, self.world, self.rank, 0,
                                                   p(residual), p(w), p(y), p(residual_out), T, self.H, eps,
                                                   torch.cuda.current_stream().cuda_stream), "allreduce_add_rms_norm")

    def close(self) -> None:
        for q in self._opened:
            self.lib.b200_ipc_close_handle(q)
        self._opened = []
