"""World-size-2 gloo test (CPU) of the tensor-parallel step protocol: rank 0's scheduler output
must reach rank 1's runner unchanged for decode and (chunked) prefill steps, and the sync/timing
helpers must not deadlock."""
import multiprocessing as mp
import os
import socket

import numpy as np
import pytest


def _port():
    with socket.socket() as s:
        s.bind(("127.0.0.1", 0))
        return s.getsockname()[1]


class FakeRunner:
    max_blocks = 6

    def __init__(self):
        self.calls = []

    def decode(self, B, last, computed, table, temps=None, seeds=None):
        self.calls.append(("decode", int(B), np.asarray(last[:B]).tolist(), np.asarray(computed[:B]).tolist(),
                           np.asarray(table[:B]).tolist()))
        self.sampler = (np.asarray(temps[:B], dtype=np.float32).tolist(), np.asarray(seeds[:B], dtype=np.uint32).tolist())
        return np.arange(B, dtype=np.int32)

    def prefill(self, chunks, starts, tables, sample, decode=None, sample_temps=None, sample_seeds=None):
        call = ("prefill", [list(map(int, c)) for c in chunks], list(map(int, starts)),
                [list(map(int, t)) for t in tables], [int(s) for s in sample])
        if decode is not None:     # mixed step: decode rows ride along
            B, last, comp, tab = decode
            call += ((int(B), np.asarray(last[:B]).tolist(), np.asarray(comp[:B]).tolist(), np.asarray(tab[:B]).tolist()),)
        self.calls.append(call)
        if sample_temps is not None:
            self.sampler = (np.asarray(sample_temps, dtype=np.float32).tolist(), np.asarray(sample_seeds, dtype=np.uint32).tolist())
        return np.zeros(sum(int(s) for s in sample), dtype=np.int32)


def _worker(rank, port, q):
    os.environ.update(MASTER_ADDR="127.0.0.1", MASTER_PORT=str(port), RANK=str(rank), WORLD_SIZE="2")
    import torch
    import torch.distributed as dist

    from gpustack_b200 import engine as E

    dist.init_process_group(backend="gloo", rank=rank, world_size=2)
    tp = E.TPContext(rank, 2, torch.device("cpu"))
    tp.open_step_channel(1 << 16)
    eng = E.LLMEngine.__new__(E.LLMEngine)      # protocol only: no GPU objects
    eng.tp, eng.runner = tp, FakeRunner()
    torch.cuda.synchronize = lambda *a, **k: None
    steps = [
        ("decode", 3, [7, 8, 9], [20, 33, 5], [[1, 2, 0, 0, 0, 0], [3, 4, 5, 0, 0, 0], [6, 0, 0, 0, 0, 0]]),
        ("prefill", [[5, 6, 7, 8], [9, 10]], [0, 16], [[2, 3], [4]], [True, False]),
        ("mixed", [[11, 12, 13]], [32], [[1, 5, 6]], [True], 2, [40, 41], [7, 19], [[2, 0, 0, 0, 0, 0], [3, 4, 0, 0, 0, 0]]),
        # speculative verification: [last_token, drafts...] per sequence, EVERY row sampled (counts, not flags)
        ("spec", [[21, 22, 23, 24], [30]], [50, 9], [[1, 2, 3, 4], [5]], [4, 1]),
    ]
    if rank == 0:
        for s in steps:
            if s[0] == "decode":
                _, B, last, comp, tab = s
                temps = np.array([0.0, 0.7, 1.5], dtype=np.float32)
                seeds = np.array([1, 2 ** 31 + 5, 2 ** 32 - 1], dtype=np.uint32)
                payload = np.concatenate([np.array(last), np.array(comp), temps.view(np.int32), seeds.view(np.int32),
                                          np.array(tab).reshape(-1)]).astype(np.int32)
                tp.bcast_step([E._KIND_DECODE, B], payload)
            elif s[0] == "mixed":
                eng._bcast_prefill(s[1], s[2], s[3], s[4], (s[5], np.array(s[6]), np.array(s[7]), np.array(s[8])))
            elif s[0] == "spec":
                eng._bcast_prefill(s[1], s[2], s[3], s[4], None, np.array([0.0, 0.0, 0.0, 0.0, 0.7], dtype=np.float32),
                                   np.array([1, 2, 3, 4, 2 ** 32 - 1], dtype=np.uint32))
            else:
                eng._bcast_prefill(s[1], s[2], s[3], s[4])
        eng.sync_mark(); eng.sync_mark()
        dt = eng.max_elapsed()
        eng.shutdown()
        q.put(("rank0", dt >= 0.0))
    else:
        E.TPWorker(eng).serve_forever()
        # sampler inputs of the last step (the speculative one), bit-exact round trip
        assert eng.runner.sampler == ([0.0, 0.0, 0.0, 0.0, 0.699999988079071], [1, 2, 3, 4, 2 ** 32 - 1])
        q.put(("rank1", eng.runner.calls))
    dist.destroy_process_group()


def test_step_broadcast_roundtrip():
    ctx = mp.get_context("spawn")
    q = ctx.Queue()
    port = _port()
    ps = [ctx.Process(target=_worker, args=(r, port, q)) for r in range(2)]
    for p in ps:
        p.start()
    got = dict(q.get(timeout=120) for _ in range(2))
    for p in ps:
        p.join(timeout=60)
        assert p.exitcode == 0
    assert got["rank0"] is True
    calls = got["rank1"]
    assert calls[0] == ("decode", 3, [7, 8, 9], [20, 33, 5],
                        [[1, 2, 0, 0, 0, 0], [3, 4, 5, 0, 0, 0], [6, 0, 0, 0, 0, 0]])
    assert calls[1] == ("prefill", [[5, 6, 7, 8], [9, 10]], [0, 16], [[2, 3], [4]], [1, 0])
    assert calls[2] == ("prefill", [[11, 12, 13]], [32], [[1, 5, 6]], [1],
                        (2, [40, 41], [7, 19], [[2, 0, 0, 0, 0, 0], [3, 4, 0, 0, 0, 0]]))
    assert calls[3] == ("prefill", [[21, 22, 23, 24], [30]], [50, 9], [[1, 2, 3, 4], [5]], [4, 1])


def _stress(rank, port, q, channel):
    os.environ.update(MASTER_ADDR="127.0.0.1", MASTER_PORT=str(port), RANK=str(rank), WORLD_SIZE="2",
                      B200_TP_STEP_CHANNEL=channel)
    import torch
    import torch.distributed as dist

    from gpustack_b200 import engine as E

    dist.init_process_group(backend="gloo", rank=rank, world_size=2)
    tp = E.TPContext(rank, 2, torch.device("cpu"))
    tp.open_step_channel(8192)                  # sized by the engine from max_num_seqs / max_model_len
    assert (tp.shm is not None) == (channel == "shm")
    if tp.shm is not None and rank == 0:        # an oversized payload is refused, never truncated
        try:
            tp.shm.send([0], np.zeros(8193, dtype=np.int32)); ok_big = False
        except E.B200Error:
            ok_big = True
        assert ok_big
    rs = np.random.RandomState(0)
    ok, n_steps = True, 1500
    for i in range(n_steps):
        n = int(rs.randint(1, 5000))
        payload = rs.randint(0, 1 << 30, size=n).astype(np.int32)
        if rank == 0:
            tp.bcast_step([E._KIND_DECODE, i, n], payload)
        else:
            h, p = tp.bcast_step([0], None)
            ok = ok and h[0] == E._KIND_DECODE and h[1] == i and h[2] == n and np.array_equal(p, payload)
    if tp.shm is not None:
        dist.barrier(group=tp.cpu)
        tp.shm.close()
    q.put((rank, ok))
    dist.destroy_process_group()


@pytest.mark.parametrize("channel", ["shm", "gloo"])
def test_step_channel_stress(channel):
    ctx = mp.get_context("spawn")
    q = ctx.Queue()
    port = _port()
    ps = [ctx.Process(target=_stress, args=(r, port, q, channel)) for r in range(2)]
    for p in ps:
        p.start()
    got = dict(q.get(timeout=180) for _ in range(2))
    for p in ps:
        p.join(timeout=60)
        assert p.exitcode == 0
    assert got == {0: True, 1: True}
