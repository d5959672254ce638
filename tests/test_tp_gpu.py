"""Tensor-parallel engine parity (needs >= 2 GPUs; skipped otherwise)."""
import os
import socket
import subprocess
import sys
from pathlib import Path

import pytest
import torch

pytestmark = pytest.mark.gpu
ROOT = Path(__file__).resolve().parents[1]


def _port():
    with socket.socket() as s:
        s.bind(("127.0.0.1", 0))
        return s.getsockname()[1]


@pytest.mark.parametrize("world", [2, 4, 8])
@pytest.mark.parametrize("allreduce", ["nccl", "oneshot", "twoshot"])
@pytest.mark.parametrize("eager", ["1", "0"])
def test_tp_engine_matches_oracle(world, eager, allreduce):
    if torch.cuda.device_count() < world:
        pytest.skip(f"needs {world} GPUs")
    cmd = [sys.executable, "-m", "torch.distributed.run", "--nnodes=1", f"--nproc-per-node={world}",
           "--master-addr", "127.0.0.1", "--master-port", str(_port()), str(ROOT / "tests" / "tp_parity_main.py")]
    env = dict(os.environ, TP_EAGER=eager, B200_TP_ALLREDUCE=allreduce)
    if allreduce == "twoshot":   # every message takes the reduce-scatter / norm-own-rows / all-gather kernel
        env.update(B200_TP_ALLREDUCE="fused", B200_TP_ONESHOT_MAX_BYTES="0")
    r = subprocess.run(cmd, capture_output=True, text=True, timeout=600, env=env)
    assert r.returncode == 0, r.stdout[-2000:] + r.stderr[-3000:]
    assert "TP_PARITY" in r.stdout and '"ok": true' in r.stdout


@pytest.mark.parametrize("family", ["fp8", "mixtral"])
@pytest.mark.parametrize("eager", ["1", "0"])
def test_tp2_fp8_and_mixtral_families(family, eager):
    """W8A8 shards quantised per rank, and Mixtral experts tensor-parallel over I, at TP=2."""
    if torch.cuda.device_count() < 2:
        pytest.skip("needs 2 GPUs")
    cmd = [sys.executable, "-m", "torch.distributed.run", "--nnodes=1", "--nproc-per-node=2",
           "--master-addr", "127.0.0.1", "--master-port", str(_port()), str(ROOT / "tests" / "tp_parity_main.py")]
    r = subprocess.run(cmd, capture_output=True, text=True, timeout=600,
                       env=dict(os.environ, TP_EAGER=eager, TP_FAMILY=family,
                                B200_TP_ONESHOT_MAX_BYTES=os.environ.get("TP_TEST_ONESHOT_MAX_BYTES", str(8 << 20))))
    assert r.returncode == 0, r.stdout[-2000:] + r.stderr[-3000:]
    assert "TP_PARITY" in r.stdout and '"ok": true' in r.stdout


def test_tp2_speculative_decoding_and_host_kv_tier():
    """The speculative verification step (per-chunk sample COUNTS in the prefill broadcast) and the extended
    KV cache's swap list (`_KIND_SWAP`) replayed by a second rank: outputs still follow the oracle."""
    if torch.cuda.device_count() < 2:
        pytest.skip("needs 2 GPUs")
    cmd = [sys.executable, "-m", "torch.distributed.run", "--nnodes=1", "--nproc-per-node=2",
           "--master-addr", "127.0.0.1", "--master-port", str(_port()), str(ROOT / "tests" / "tp_parity_main.py")]
    r = subprocess.run(cmd, capture_output=True, text=True, timeout=600, env=dict(os.environ, TP_EAGER="0", TP_FEATURES="1"))
    assert r.returncode == 0, r.stdout[-2000:] + r.stderr[-3000:]
    assert "TP_PARITY" in r.stdout and '"ok": true' in r.stdout and '"features"' in r.stdout
