"""ctypes binding of libb200.so (include/b200.h).

There is no CPU fallback: if the shared library is missing it is built with nvcc; if that fails,
or a call returns an error, a B200Error is raised.
"""
from __future__ import annotations

import ctypes as C
import os
import threading
from pathlib import Path

from . import build as _build

_LIB = None
_LOCK = threading.Lock()


class B200Error(RuntimeError):
    pass


_vp, _i32, _i64, _f32, _u32, _sz = C.c_void_p, C.c_int, C.c_int64, C.c_float, C.c_uint32, C.c_size_t

# name -> (restype, argtypes); mirrors include/b200.h one to one
SIGNATURES = {
    "b200_abi_version": (_i32, []),
    "b200_last_error": (C.c_char_p, []),
    "b200_device_info": (_i32, [C.POINTER(_i32)] * 4),
    "b200_rms_norm": (_i32, [_vp, _vp, _vp, _i32, _i32, _f32, _vp]),
    "b200_fused_add_rms_norm": (_i32, [_vp, _vp, _vp, _vp, _vp, _i32, _i32, _f32, _vp]),
    "b200_fused_add_rms_norm_partials": (_i32, [_vp, _i32, _vp, _vp, _vp, _vp, _i32, _i32, _f32, _vp]),
    "b200_rope_kv_write": (_i32, [_vp, _vp, _vp, _vp, _vp, _vp, _i32, _i32, _i32, _i32, _i32, _vp]),
    "b200_rope_kv_write_partials": (_i32, [_vp, _i32, _vp, _vp, _vp, _vp, _vp, _vp, _vp, _i32, _i32, _i32, _i32, _i32, _vp]),
    "b200_silu_and_mul": (_i32, [_vp, _vp, _i32, _i32, _vp]),
    "b200_embedding": (_i32, [_vp, _vp, _vp, _i32, _i32, _i32, _i32, _vp]),
    "b200_argmax_f32": (_i32, [_vp, _i64, _vp, _vp, _i32, _i32, _vp]),
    "b200_sample_f32": (_i32, [_vp, _i64, _vp, _vp, _i32, _vp, _vp, _i32, _i32, _vp]),
    "b200_gather_rows": (_i32, [_vp, _vp, _vp, _i32, _i32, _vp]),
    "b200_gemm_bf16": (_i32, [_vp, _vp, _vp, _i32, _i32, _i32, _i64, _i64, _i64, _i32, _i32, _vp]),
    "b200_gemm_bf16_ex": (_i32, [_vp, _vp, _vp, _i32, _i32, _i32, _i64, _i64, _i64, _i32, _i32, _i32, _vp]),
    "b200_gemm_bf16_silu_mul": (_i32, [_vp, _vp, _vp, _i32, _i32, _i32, _i64, _i64, _i64, _vp]),
    "b200_gemm_fp8": (_i32, [_vp, _vp, _vp, _i32, _i32, _i32, _i64, _i64, _i64, _vp, _vp, _i32, _i32, _i32, _i32, _vp]),
    "b200_quant_fp8_per_token": (_i32, [_vp, _i64, _vp, _i64, _vp, _i32, _i32, _vp]),
    "b200_moe_route": (_i32, [_vp, _i64, _vp, _vp, _i32, _i32, _i32, _i32, _vp]),
    "b200_moe_capacity_rows": (_i32, [_i32, _i32, _i32]),
    "b200_moe_align": (_i32, [_vp, _vp, _i32, _i32, _i32, _i32, _vp, _vp, _vp, _vp, _vp, _vp]),
    "b200_moe_gather": (_i32, [_vp, _vp, _vp, _vp, _i32, _i32, _i32, _vp]),
    "b200_moe_grouped_gemm": (_i32, [_vp, _vp, _vp, _i32, _i32, _i32, _i32, _i64, _i64, _i64, _vp, _vp, _vp,
                                     _i32, _i32, _vp]),
    "b200_moe_combine": (_i32, [_vp, _vp, _vp, _i32, _i32, _i32, _vp]),
    "b200_gemm_pick_block_x": (_i32, [_i32, _i32, _i32]),
    "b200_splitk_reduce": (_i32, [_vp, _i32, _vp, _i32, _i32, _i64, _vp]),
    "b200_gemm_suggest_splits": (_i32, [_i32, _i32, _i32]),
    "b200_attn_decode_splits": (_i32, [_i32, _i32, _i32, _i32]),
    "b200_attn_decode_workspace_bytes": (_sz, [_i32, _i32, _i32, _i32]),
    "b200_paged_attn_decode": (_i32, [_vp, _i64, _vp, _vp, _vp, _i32, _vp, _vp, _vp, _i32, _i32, _i32,
                                      _i32, _i32, _i32, _f32, _vp]),
    "b200_attn_prefill_tile": (_i32, [_i32, _i32]),
    "b200_paged_attn_prefill": (_i32, [_vp, _i64, _i32, _vp, _vp, _i32, _vp, _i32, _vp, _vp, _vp, _vp, _i32, _vp,
                                       _i32, _i32, _i32, _i32, _f32, _vp]),
    "b200_dev_malloc": (_i32, [_sz, C.POINTER(_vp)]),
    "b200_dev_free": (_i32, [_vp]),
    "b200_ipc_get_handle": (_i32, [_vp, _vp]),
    "b200_ipc_open_handle": (_i32, [_vp, C.POINTER(_vp)]),
    "b200_ipc_close_handle": (_i32, [_vp]),
    "b200_allreduce_flag_words": (_i32, [_i32]),
    "b200_allreduce_add_rms_norm": (_i32, [C.POINTER(_vp), C.POINTER(_vp), _i32, _i32, _u32, _vp, _vp,
                                           _vp, _vp, _i32, _i32, _f32, _vp]),
    "b200_allreduce_twoshot_flag_words": (_i32, []),
    "b200_tp_argmax_select": (_i32, [C.POINTER(_vp), C.POINTER(_vp), C.POINTER(_vp), _i32, _i32, _i32, _vp, _i32, _vp]),
    "b200_allreduce_rs_norm_ag": (_i32, [C.POINTER(_vp), C.POINTER(_vp), C.POINTER(_vp), _i32, _i32, _vp, _vp, _vp,
                                         _i32, _i32, _f32, _vp]),
}


def lib_path() -> Path:
    # B200_LIB selects an experimental build of the same ABI (e.g. csrc/libb200_late.so)
    override = os.environ.get("B200_LIB")
    return Path(override) if override else _build.LIB


def load(build_if_missing: bool = True):
    """Load (building first if needed) libb200.so and attach prototypes."""
    global _LIB
    with _LOCK:
        if _LIB is not None:
            return _LIB
        path = lib_path()
        if not path.exists():
            if not build_if_missing:
                raise B200Error(f"{path} is missing; run `python -m gpustack_b200.build`")
            _build.build()
        lib = C.CDLL(str(path))
        for name, (res, args) in SIGNATURES.items():
            try:
                fn = getattr(lib, name)
            except AttributeError as e:  # pragma: no cover - a stale .so
                raise B200Error(f"{path} does not export {name}; rebuild it") from e
            fn.restype = res
            fn.argtypes = args
        if lib.b200_abi_version() != 1:
            raise B200Error("libb200.so ABI version mismatch; rebuild it")
        _LIB = lib
        return lib


def check(rc: int, what: str = "") -> None:
    if rc != 0:
        msg = load().b200_last_error().decode(errors="replace")
        raise B200Error(f"{what or 'libb200'} failed (rc={rc}): {msg}")
