"""Host logic of the continuous-batching scheduler (no GPU): admission, chunked prefill,
decode-slot allocation, preemption + recompute, finish handling, block accounting."""
import numpy as np
import pytest

from gpustack_b200.kv_cache import BlockPool, KVCacheManager
from gpustack_b200.scheduler import DecodePlan, PrefillPlan, Request, SamplingParams, Scheduler


def mk(num_blocks=64, max_seqs=8, budget=64, max_len=256, eos=2):
    return Scheduler(num_blocks, 16, max_seqs, budget, max_len, eos)


def req(i, n, **kw):
    return Request(f"r{i}", list(range(3, 3 + n)), SamplingParams(**kw))


def drive(s, plan, tok=7):
    if isinstance(plan, PrefillPlan):
        return s.update_prefill(plan, [tok] * sum(plan.sample))
    return s.update_decode(plan, np.full(plan.batch, tok, dtype=np.int32))


def test_block_pool_accounting():
    p = BlockPool(10)
    assert p.num_total == 9 and p.num_free == 9       # block 0 is the null block
    a = p.allocate(4)
    assert 0 not in a and len(set(a)) == 4 and p.num_free == 5
    with pytest.raises(MemoryError):
        p.allocate(6)
    p.free(a)
    assert p.num_free == 9 and abs(p.usage()) < 1e-9
    m = KVCacheManager(10)
    m.append(1, 17)
    assert len(m.seqs[1].blocks) == 2 and m.slot(1, 16) == m.seqs[1].blocks[1] * 16
    m.free(1)
    assert m.pool.num_free == 9


def test_prefill_first_then_decode_until_length():
    s = mk()
    s.add(req(0, 10, max_tokens=3, ignore_eos=True))
    s.add(req(1, 20, max_tokens=2, ignore_eos=True))
    p = s.schedule()
    assert isinstance(p, PrefillPlan) and p.lens == [10, 20] and p.sample == [True, True]
    out = drive(s, p)
    assert out.req_ids == ["r0", "r1"] and out.finished == [None, None]
    assert s.num_computed[:2].tolist() == [10, 20]
    d = s.schedule()
    assert isinstance(d, DecodePlan) and d.batch == 2
    out = drive(s, d)
    assert out.finished == [None, "length"]            # r1 wanted 2 tokens
    out = drive(s, s.schedule())
    assert out.req_ids == ["r0"] and out.finished == ["length"]
    assert s.schedule() is None and not s.has_work()
    assert s.pool.num_free == s.pool.num_total


def test_chunked_prefill_respects_token_budget():
    s = mk(budget=32)
    s.add(req(0, 70, max_tokens=1, ignore_eos=True))
    s.add(req(1, 10, max_tokens=1, ignore_eos=True))
    lens, samples = [], []
    while s.has_work():
        p = s.schedule()
        assert isinstance(p, PrefillPlan) and p.num_tokens <= 32
        lens.append(list(p.lens)); samples.append(list(p.sample))
        drive(s, p)
    assert lens == [[32], [32], [6, 10]] and samples == [[False], [False], [True, True]]


def test_eos_and_stop_ids_finish_with_stop():
    s = mk()
    s.add(req(0, 5, max_tokens=10))
    s.add(req(1, 5, max_tokens=10, stop_token_ids=(9,)))
    s.add(req(2, 5, max_tokens=10, ignore_eos=True))
    drive(s, s.schedule(), tok=7)
    out = s.update_decode(s.schedule(), np.array([2, 9, 2], dtype=np.int32))
    assert out.finished == ["stop", "stop", None]
    assert [r.req_id for r in s.rows] == ["r2"]


def test_preemption_recomputes_and_never_loses_tokens():
    s = mk(num_blocks=8, max_seqs=4, budget=64, max_len=128)   # 7 usable blocks = 112 token slots
    for i in range(3):
        s.add(req(i, 30, max_tokens=40, ignore_eos=True))
    produced = {f"r{i}": 0 for i in range(3)}
    steps = 0
    while s.has_work():
        plan = s.schedule()
        assert plan is not None, "scheduler stalled"
        out = drive(s, plan)
        for rid in out.req_ids:
            produced[rid] += 1
        steps += 1
        assert steps < 2000
    assert produced == {"r0": 40, "r1": 40, "r2": 40}
    assert s.num_preemptions > 0
    assert s.pool.num_free == s.pool.num_total


def test_rejects_oversized_and_empty_requests():
    s = mk(max_len=64)
    with pytest.raises(ValueError):
        s.add(req(0, 60, max_tokens=10))
    with pytest.raises(ValueError):
        s.add(Request("e", [], SamplingParams()))


def test_persistent_batch_stays_compact_after_eviction():
    s = mk()
    for i in range(4):
        s.add(req(i, 8 + i, max_tokens=2 + i, ignore_eos=True))
    drive(s, s.schedule())
    for _ in range(6):
        p = s.schedule()
        if p is None:
            break
        drive(s, p)
        B = len(s.rows)
        assert [r.row for r in s.rows] == list(range(B))
        for r in s.rows:
            assert s.block_table[r.row, : len(r.blocks)].tolist() == r.blocks
            assert s.num_computed[r.row] == len(r.prompt_ids) + len(r.output_ids) - 1
