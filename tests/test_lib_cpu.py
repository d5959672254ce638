"""The C-ABI library loads on a CPU-only box and exports every symbol include/b200.h declares
(no compute calls without a GPU)."""
import ctypes
import re
from pathlib import Path

import pytest

ROOT = Path(__file__).resolve().parents[1]


def _declared():
    text = (ROOT / "include" / "b200.h").read_text()
    text = re.sub(r"/\*.*?\*/", "", text, flags=re.S)
    return sorted(set(re.findall(r"\b(b200_[a-z0-9_]+)\s*\(", text)))


def test_header_symbols_are_exported_and_bound():
    from gpustack_b200 import lib

    handle = lib.load()
    names = _declared()
    assert len(names) >= 20
    for n in names:
        assert hasattr(handle, n), f"libb200.so does not export {n}"
        assert n in lib.SIGNATURES, f"{n} has no ctypes prototype in gpustack_b200/lib.py"
    for n in lib.SIGNATURES:
        assert n in names, f"{n} is bound in lib.py but not declared in include/b200.h"
    assert handle.b200_abi_version() == 1


def test_argument_validation_without_gpu():
    """Argument checks run before any CUDA call, so they are testable here."""
    from gpustack_b200 import lib

    h = lib.load()
    assert h.b200_rms_norm(None, None, None, 4, 4097, 1e-5, None) < 0      # H not a multiple of 8
    assert b"multiple of 8" in h.b200_last_error()
    assert h.b200_gemm_bf16(None, None, None, 4, 128, 100, 100, 100, 128, 0, 1, None) < 0   # K % 64
    assert h.b200_paged_attn_decode(None, 8, None, None, None, 1, None, None, None, 1, 1, 32, 8, 128, 32, 1.0, None) < 0
    assert b"block_size" in h.b200_last_error()
    assert h.b200_gemm_suggest_splits(64, 4096, 14336) >= 1
    assert h.b200_gemm_pick_block_x(1, 1, 64) == 16 and h.b200_gemm_pick_block_x(300, 1, 64) == 256
    assert h.b200_attn_prefill_tile(32, 8) == 32 and h.b200_attn_prefill_tile(64, 8) == 16
    # fp32 partial O + (m, l) per (row, split), preceded by the 256-byte-aligned arrival counters of the in-kernel merge
    assert h.b200_attn_decode_workspace_bytes(4, 32, 128, 8) == 4 * 32 * 8 * 130 * 4 + 512


def test_ops_refuse_cpu_tensors():
    import torch

    from gpustack_b200 import ops
    from gpustack_b200.lib import B200Error

    with pytest.raises(B200Error):
        ops.rms_norm(torch.zeros(2, 64, dtype=torch.bfloat16), torch.ones(64, dtype=torch.bfloat16), 1e-5)
    with pytest.raises(B200Error):
        ops.gemm(torch.zeros(2, 64, dtype=torch.bfloat16), torch.zeros(128, 64, dtype=torch.bfloat16))


def test_product_code_never_imports_the_oracle():
    for f in (ROOT / "gpustack_b200").rglob("*.py"):
        src = f.read_text()
        assert "import oracle" not in src and "from oracle" not in src, f"{f} references oracle/"
