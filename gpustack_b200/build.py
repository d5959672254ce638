"""Build libb200.so (the sm_100a kernel library) in-tree with nvcc.

The shared object lands next to the sources (gpustack_b200/csrc/libb200.so) so that it travels
with the repository snapshot to the GPU box; it is git-ignored.  nvcc cross-compiles without a GPU.
"""
from __future__ import annotations

import hashlib
import os
import shutil
import subprocess
import sys
from concurrent.futures import ThreadPoolExecutor
from pathlib import Path

CSRC = Path(__file__).resolve().parent / "csrc"
LIB = CSRC / "libb200.so"
SOURCES = [
    "runtime.cu",
    "norm.cu",
    "rope_kv.cu",
    "elementwise.cu",
    "gemm_tcgen05.cu",
    "quant_fp8.cu",
    "moe.cu",
    "attention.cu",
    "attention_prefill_tc.cu",
    "allreduce.cu",
]
NVCC_FLAGS = [
    "-gencode", "arch=compute_100a,code=sm_100a",
    "-O3", "-std=c++17", "-lineinfo",
    "-Xcompiler", "-fPIC",
    "--expt-relaxed-constexpr",
    "-Xptxas", "-v",
]


def _nvcc() -> str:
    for cand in (os.environ.get("NVCC"), shutil.which("nvcc"), "/usr/local/cuda/bin/nvcc"):
        if cand and Path(cand).exists():
            return cand
    raise RuntimeError("nvcc not found (set NVCC=/path/to/nvcc)")


def _digest(paths) -> str:
    h = hashlib.sha256()
    for p in sorted(paths):
        h.update(p.name.encode())
        h.update(p.read_bytes())
    h.update(" ".join(NVCC_FLAGS).encode())
    return h.hexdigest()


def build(force: bool = False, verbose: bool = False, variant: str = "") -> Path:
    """variant "" = the product library (dependents released after the kernel's own dependency wait);
    "early" = libb200_early.so, the experimental build with the round-1 entry-time trigger
    (-DB200_PDL_EARLY_TRIGGER), kept only to reproduce the hazard documented in DESIGN.md section 9."""
    srcs = [CSRC / s for s in SOURCES]
    deps = srcs + sorted(CSRC.glob("*.cuh")) + sorted(CSRC.glob("*.h")) + [
        CSRC.parent.parent / "include" / "b200.h"
    ]
    lib = LIB if not variant else CSRC / f"libb200_{variant}.so"
    extra = {"": [], "early": ["-DB200_PDL_EARLY_TRIGGER=1"]}[variant]
    objdir = CSRC / ("build" if not variant else f"build_{variant}")
    stamp = objdir / "stamp"
    digest = _digest(deps) + variant
    if not force and lib.exists() and stamp.exists() and stamp.read_text() == digest:
        return lib
    nvcc = _nvcc()
    objdir.mkdir(exist_ok=True)

    def compile_one(src: Path):
        obj = objdir / (src.stem + ".o")
        cmd = [nvcc, *NVCC_FLAGS, *extra, "-c", str(src), "-o", str(obj)]
        r = subprocess.run(cmd, capture_output=True, text=True)
        log = objdir / (src.stem + ".log")
        log.write_text(r.stdout + r.stderr)
        if r.returncode != 0:
            raise RuntimeError(f"nvcc failed for {src.name}:\n{r.stdout}\n{r.stderr}")
        if verbose:
            sys.stderr.write(r.stderr)
        return obj

    with ThreadPoolExecutor(max_workers=min(8, len(srcs))) as ex:
        objs = list(ex.map(compile_one, srcs))
    cmd = [nvcc, "-shared", "-o", str(lib), *map(str, objs)]
    r = subprocess.run(cmd, capture_output=True, text=True)
    if r.returncode != 0:
        raise RuntimeError(f"link failed:\n{r.stdout}\n{r.stderr}")
    stamp.write_text(digest)
    return lib


if __name__ == "__main__":
    print(build(force="--force" in sys.argv, verbose=True, variant="early" if "--early" in sys.argv else ""))
