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


def _peer_buffers(tp, sizes, device):
    """Allocates one device buffer per entry of `sizes` on this rank (cudaMalloc through the C ABI,
    zero-filled), exchanges their CUDA-IPC handles and opens the peers' buffers in this process.
    Returns (own_ptrs, opened_ptrs, tables) with tables[k][r] = rank r's buffer k as mapped here."""
    lib = load()
    own: List[int] = []
    handles = []
    with torch.cuda.device(device):
        for size in sizes:
            p = C.c_void_p()
            check(lib.b200_dev_malloc(size, C.byref(p)), "dev_malloc")
            h = C.create_string_buffer(64)
            check(lib.b200_ipc_get_handle(p, h), "ipc_get_handle")
            own.append(p.value)
            handles.append(h.raw)
        torch.cuda.synchronize()
        gathered: List = [None] * tp.world
        tp.dist.all_gather_object(gathered, handles, group=tp.cpu)
        opened: List[int] = []
        tables = [(C.c_void_p * MAX_WORLD)() for _ in sizes]
        for r in range(tp.world):
            if r == tp.rank:
                ptrs = own
            else:
                ptrs = []
                for raw in gathered[r]:
                    q = C.c_void_p()
                    check(lib.b200_ipc_open_handle(C.create_string_buffer(raw, 64), C.byref(q)), "ipc_open_handle")
                    ptrs.append(q.value)
                opened.extend(ptrs)
            for k in range(len(sizes)):
                tables[k][r] = ptrs[k]
    return own, opened, tables


class TwoShotAllReduce:
    """Buffers of the fused two-shot kernel (csrc/allreduce.cu: barrier -> reduce-scatter by token
    rows -> residual add + RMSNorm on this rank's rows -> all-gather by peer stores -> barrier).
    `stage(T)` is where the row-parallel GEMM writes its partial result; `xn` (the gathered,
    normalised activations every rank reads next) is itself peer-mapped, because the peers write it."""

    def __init__(self, tp, max_tokens: int, hidden: int, device: torch.device):
        self.tp, self.world, self.rank = tp, tp.world, tp.rank
        self.max_tokens, self.H = max_tokens, hidden
        self.lib = lib = load()
        nbytes = max_tokens * hidden * 2
        fbytes = lib.b200_allreduce_twoshot_flag_words() * 4
        self._own, self._opened, (self._stage, self._xn, self._flags) = _peer_buffers(tp, (nbytes, nbytes, fbytes), device)
        self.stage_buf = torch.as_tensor(_RawCudaBuffer(self._own[0], nbytes, "<u2", 2), device=device).view(torch.bfloat16)
        self.xn = torch.as_tensor(_RawCudaBuffer(self._own[1], nbytes, "<u2", 2), device=device) \
            .view(torch.bfloat16).view(max_tokens, hidden)
        tp.dist.barrier(group=tp.cpu)

    def stage(self, T: int) -> torch.Tensor:
        return self.stage_buf[: T * self.H].view(T, self.H)

    def rs_norm_ag(self, residual, w, residual_out, T: int, eps: float) -> None:
        """xn[:T] (on every rank) = rmsnorm(sum_r stage_r + residual) * w; residual rows of this rank updated."""
        check(self.lib.b200_allreduce_rs_norm_ag(self._stage, self._xn, self._flags, self.world, self.rank,
                                                 None if residual is None else residual.data_ptr(), w.data_ptr(),
                                                 residual_out.data_ptr(), T, self.H,
                                                 eps, torch.cuda.current_stream().cuda_stream), "allreduce_rs_norm_ag")

    def close(self) -> None:
        for q in self._opened:
            self.lib.b200_ipc_close_handle(q)
        self._opened = []


class PeerArgmax:
    """Peer-mapped (max, argmax) arrays for greedy sampling over vocabulary-parallel logits
    (csrc/allreduce.cu: tp_argmax_select_kernel)."""

    def __init__(self, tp, max_seqs: int, vocab_per_rank: int, device: torch.device):
        self.tp, self.world, self.rank, self.per = tp, tp.world, tp.rank, vocab_per_rank
        self.lib = load()
        n = max(32, max_seqs)
        self._own, self._opened, (self._val, self._idx, self._flags) = _peer_buffers(tp, (n * 4, n * 4, 64 * 4), device)
        self.vals = torch.as_tensor(_RawCudaBuffer(self._own[0], n * 4, "<f4", 4), device=device)
        self.idxs = torch.as_tensor(_RawCudaBuffer(self._own[1], n * 4, "<i4", 4), device=device)
        tp.dist.barrier(group=tp.cpu)

    def select(self, out_idx: torch.Tensor, n: int) -> None:
        """out_idx[:n] = global argmax, from every rank's (self.vals[:n], self.idxs[:n])."""
        check(self.lib.b200_tp_argmax_select(self._val, self._idx, self._flags, self.world, self.rank, self.per,
                                             out_idx.data_ptr(), n, torch.cuda.current_stream().cuda_stream),
              "tp_argmax_select")

    def close(self) -> None:
        for q in self._opened:
            self.lib.b200_ipc_close_handle(q)
        self._opened = []


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
