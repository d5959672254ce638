"""Host logic of the continuous-batching scheduler (no GPU): admission, chunked prefill,
decode-slot allocation, preemption + recompute, finish handling, block accounting."""
import numpy as np
import pytest

from gpustack_b200.kv_cache import BlockPool, KVCacheManager
from gpustack_b200.scheduler import DecodePlan, MixedPlan, PrefillPlan, Request, SamplingParams, Scheduler, SpecPlan
from gpustack_b200.spec_decode import SpeculativeConfig


def mk(num_blocks=64, max_seqs=8, budget=64, max_len=256, eos=2):
    return Scheduler(num_blocks, 16, max_seqs, budget, max_len, eos)


def req(i, n, **kw):
    return Request(f"r{i}", list(range(3, 3 + n)), SamplingParams(**kw))


def drive(s, plan, tok=7):
    if isinstance(plan, MixedPlan):
        return s.update_mixed(plan, [tok] * sum(plan.prefill.sample), np.full(plan.batch, tok, dtype=np.int32))
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


def test_mixed_steps_carry_decode_tokens_with_the_prefill_chunks():
    """A prompt that arrives while another request is decoding is prefilled in a MIXED step: the
    running row produces its token in the same step, the decode token is charged to the budget."""
    s = mk(budget=32)
    s.add(req(0, 10, max_tokens=5, ignore_eos=True))
    out = drive(s, s.schedule())
    assert out.req_ids == ["r0"]
    s.add(req(1, 40, max_tokens=3, ignore_eos=True))
    p = s.schedule()
    assert isinstance(p, MixedPlan) and p.batch == 1
    assert p.prefill.lens == [31] and p.prefill.sample == [False] and p.num_tokens == 32   # 32 - 1 decode token
    out = drive(s, p, tok=9)
    assert out.req_ids == ["r0"] and out.token_ids == [9] and s.num_computed[0] == 11
    p = s.schedule()
    assert isinstance(p, MixedPlan) and p.prefill.lens == [9] and p.prefill.sample == [True]
    out = drive(s, p, tok=8)
    assert out.req_ids == ["r0", "r1"] and out.finished == [None, None]
    assert [r.req_id for r in s.rows] == ["r0", "r1"] and s.num_computed[:2].tolist() == [12, 40]
    d = s.schedule()
    assert isinstance(d, DecodePlan) and d.batch == 2            # nothing left to prefill: pure decode step
    # the round-1 policy is still available
    s2 = Scheduler(64, 16, 8, 32, 256, 2, mixed_steps=False)
    s2.add(req(0, 10, max_tokens=5, ignore_eos=True))
    drive(s2, s2.schedule())
    s2.add(req(1, 40, max_tokens=3, ignore_eos=True))
    p2 = s2.schedule()
    assert isinstance(p2, PrefillPlan) and p2.lens == [32]


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


def test_ignore_eos_suppresses_only_eos_and_every_eos_id_counts():
    """Reference-engine semantics: ignore_eos never disables stop_token_ids; a checkpoint may list
    several turn terminators (config.json list + generation_config.json)."""
    s = Scheduler(64, 16, 8, 64, 256, (2, 5))
    s.add(req(0, 5, max_tokens=10, ignore_eos=True, stop_token_ids=(9,)))
    s.add(req(1, 5, max_tokens=10))
    s.add(req(2, 5, max_tokens=10, ignore_eos=True))
    drive(s, s.schedule(), tok=7)
    out = s.update_decode(s.schedule(), np.array([9, 5, 5], dtype=np.int32))
    assert out.finished == ["stop", "stop", None]
    # the same rules on the prefill path (first sampled token)
    s2 = Scheduler(64, 16, 8, 64, 256, (2, 5))
    s2.add(req(0, 5, max_tokens=10, ignore_eos=True, stop_token_ids=(9,)))
    s2.add(req(1, 5, max_tokens=10))
    out = s2.update_prefill(s2.schedule(), [9, 5])
    assert out.finished == ["stop", "stop"]


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


# ---------------------------------------------------------------------------- prefix caching
def mkc(num_blocks=64, max_seqs=8, budget=256, max_len=512, eos=2):
    return Scheduler(num_blocks, 16, max_seqs, budget, max_len, eos, enable_prefix_caching=True)


def run_to_completion(s, tok=7):
    n = 0
    while s.has_work():
        drive(s, s.schedule(), tok)
        n += 1
        assert n < 10_000
    return n


def test_block_digests_chain_and_extend():
    from gpustack_b200.kv_cache import block_digests

    a = block_digests(list(range(40)), 16)
    assert len(a) == 2 and len(a[0]) == 16 and a[0] != a[1]
    b = block_digests(list(range(48)), 16, have=list(a))
    assert b[:2] == a and len(b) == 3
    c = block_digests([9] + list(range(1, 48)), 16)
    assert c[0] != b[0] and c[1] != b[1] and c[2] != b[2]        # a different head changes every key
    assert block_digests(list(range(48)), 16, upto_blocks=1) == a[:1]


def test_prefix_cache_hit_skips_cached_blocks_and_shares_them():
    s = mkc()
    prompt = list(range(3, 3 + 50))                               # 3 full blocks + 2 tokens
    s.add(Request("a", list(prompt), SamplingParams(max_tokens=2, ignore_eos=True)))
    p = s.schedule()
    assert p.starts == [0] and p.lens == [50]
    a_blocks = list(p.reqs[0].blocks)
    drive(s, p)
    run_to_completion(s)
    assert s.pool.num_free == s.pool.num_total and s.pool.num_cached == 3
    # same prompt again + a request sharing only the first 2 blocks
    s.add(Request("b", list(prompt), SamplingParams(max_tokens=2, ignore_eos=True)))
    s.add(Request("c", prompt[:32] + [400] * 30, SamplingParams(max_tokens=2, ignore_eos=True)))
    p = s.schedule()
    assert p.starts == [48, 32] and p.lens == [2, 30] and p.sample == [True, True]
    rb, rc = p.reqs
    assert rb.blocks[:3] == a_blocks[:3] and rc.blocks[:2] == a_blocks[:2] and rc.blocks[2] != rb.blocks[2]
    assert rb.num_cached_tokens == 48 and rc.num_cached_tokens == 32
    assert s.pool.hits == 80 and s.pool.queries == 50 + 50 + 62
    run_to_completion(s)
    assert s.pool.num_free == s.pool.num_total


def test_prefix_cache_never_covers_the_last_prompt_token():
    s = mkc()
    prompt = list(range(3, 3 + 32))                               # exactly 2 full blocks
    s.add(Request("a", list(prompt), SamplingParams(max_tokens=1, ignore_eos=True)))
    run_to_completion(s)
    s.add(Request("b", list(prompt), SamplingParams(max_tokens=1, ignore_eos=True)))
    p = s.schedule()
    assert p.starts == [16] and p.lens == [16] and p.sample == [True]   # last block recomputed for its logits


def test_generated_tokens_are_cached_for_the_next_turn():
    s = mkc()
    prompt = list(range(3, 3 + 20))
    s.add(Request("a", list(prompt), SamplingParams(max_tokens=30, ignore_eos=True)))
    run_to_completion(s, tok=7)
    # K/V in the cache at the end: 20 prompt + 29 generated tokens = 49 -> 3 full blocks
    assert s.pool.num_cached == 3
    turn2 = prompt + [7] * 30 + [11, 12, 13]
    s.add(Request("b", turn2, SamplingParams(max_tokens=1, ignore_eos=True)))
    p = s.schedule()
    assert p.starts == [48] and p.lens == [5]


def test_cached_blocks_are_evicted_lru_and_only_when_needed():
    s = mkc(num_blocks=9, budget=512)                             # 8 usable blocks
    s.add(Request("a", list(range(100, 148)), SamplingParams(max_tokens=1, ignore_eos=True)))   # 3 full + spare
    run_to_completion(s)
    assert s.pool.num_cached == 3 and s.pool.num_free == 8
    s.add(Request("b", list(range(200, 264)), SamplingParams(max_tokens=1, ignore_eos=True)))   # needs 5 blocks
    run_to_completion(s)
    # 8 blocks, b took 5: the never-keyed spare block of `a` + fresh ones go first; a's head survives
    s.add(Request("c", list(range(100, 148)), SamplingParams(max_tokens=1, ignore_eos=True)))
    p = s.schedule()
    assert p.starts[0] >= 16                                      # at least the first block of `a` is still cached
    run_to_completion(s)
    # now force eviction of everything: a prompt that needs all 8 blocks
    s.add(Request("d", list(range(300, 300 + 120)), SamplingParams(max_tokens=1, ignore_eos=True)))
    p = s.schedule()
    assert p.starts == [0] and p.lens == [120]
    assert s.pool.num_cached == 0
    run_to_completion(s)
    assert s.pool.num_free == s.pool.num_total and s.pool.num_cached == 7


def test_preempted_request_reattaches_its_own_blocks():
    s = mkc(num_blocks=8, max_seqs=4, budget=256, max_len=256)     # 7 usable blocks
    s.add(Request("a", list(range(10, 42)), SamplingParams(max_tokens=64, ignore_eos=True)))
    s.add(Request("b", list(range(50, 82)), SamplingParams(max_tokens=64, ignore_eos=True)))
    pre = 0
    steps = 0
    seen_restart = None
    while s.has_work():
        plan = s.schedule()
        if isinstance(plan, PrefillPlan) and any(r.num_preemptions for r in plan.reqs):
            seen_restart = (plan.starts[0], plan.lens[0])
        drive(s, plan)
        steps += 1
        assert steps < 2000
    assert s.num_preemptions >= 1
    assert seen_restart is not None and seen_restart[0] >= 16       # recompute resumes behind its surviving blocks
    assert s.pool.num_free == s.pool.num_total


def test_shared_block_refcounts_survive_one_owner_finishing():
    s = mkc()
    prompt = list(range(3, 3 + 40))
    s.add(Request("a", list(prompt), SamplingParams(max_tokens=1, ignore_eos=True)))
    run_to_completion(s)
    s.add(Request("b", list(prompt), SamplingParams(max_tokens=3, ignore_eos=True)))
    s.add(Request("c", list(prompt), SamplingParams(max_tokens=9, ignore_eos=True)))
    p = s.schedule()
    rb, rc = p.reqs
    shared = rb.blocks[:2]
    assert rc.blocks[:2] == shared and all(s.pool._ref[b] == 2 for b in shared)
    drive(s, p)
    while not rb.finished:
        drive(s, s.schedule())
    assert all(s.pool._ref[b] == 1 for b in shared) and not rc.finished
    assert all(b not in s.pool._free for b in shared)
    run_to_completion(s)
    assert all(s.pool._ref[b] == 0 for b in shared) and s.pool.num_free == s.pool.num_total


def test_disabled_prefix_caching_never_reuses():
    s = mk()
    prompt = list(range(3, 3 + 50))
    s.add(Request("a", list(prompt), SamplingParams(max_tokens=1, ignore_eos=True)))
    while s.has_work():
        drive(s, s.schedule())
    s.add(Request("b", list(prompt), SamplingParams(max_tokens=1, ignore_eos=True)))
    p = s.schedule()
    assert p.starts == [0] and p.lens == [50] and s.pool.num_cached == 0


def test_prefix_cache_randomised_content_model():
    """Model the device cache as token ids per slot: whatever the scheduler plans, the context a
    chunk attends to must hold exactly the request's own tokens, shared blocks must never be
    written, and the ref counts must equal the number of live owners."""
    rng = np.random.RandomState(0)
    BS = 16
    s = Scheduler(40, BS, 6, 96, 400, 2, enable_prefix_caching=True)
    store = np.full((40, BS), -1, dtype=np.int64)
    heads = [rng.randint(3, 500, size=n).tolist() for n in (16, 48, 70, 33)]
    live, nreq, steps = {}, 0, 0
    while steps < 1500:
        steps += 1
        if nreq < 120 and rng.rand() < 0.35:
            h = heads[rng.randint(len(heads))]
            tail = rng.randint(3, 500, size=rng.randint(1, 40)).tolist()
            r = Request(f"q{nreq}", h[: rng.randint(1, len(h) + 1)] + tail,
                        SamplingParams(max_tokens=int(rng.randint(1, 40)), ignore_eos=True))
            s.add(r); live[r.req_id] = r; nreq += 1
        plan = s.schedule()
        if plan is None:
            if nreq >= 120:
                break
            continue
        if isinstance(plan, PrefillPlan):
            for r, st, n in zip(plan.reqs, plan.starts, plan.lens):
                ids = r.all_ids
                for pos in range(st):                            # context the chunk will read
                    assert store[r.blocks[pos // BS], pos % BS] == ids[pos], (r.req_id, pos)
                for pos in range(st, st + n):                    # K/V the chunk writes
                    b = r.blocks[pos // BS]
                    assert s.pool._ref[b] == 1 or store[b, pos % BS] == ids[pos], "write into a shared block"
                    store[b, pos % BS] = ids[pos]
            toks = rng.randint(3, 500, size=sum(plan.sample)).tolist()
            s.update_prefill(plan, toks)
        else:
            B = plan.batch
            for i in range(B):
                r = s.rows[i]
                pos = int(s.num_computed[i])
                b = int(s.block_table[i, pos // BS])
                assert b == r.blocks[pos // BS] and s.pool._ref[b] == 1
                for q in range(pos):
                    assert store[int(s.block_table[i, q // BS]), q % BS] == r.all_ids[q]
                store[b, pos % BS] = int(s.last_token[i])
                assert int(s.last_token[i]) == r.all_ids[pos]
            s.update_decode(plan, rng.randint(3, 500, size=B).astype(np.int32))
        # ref counts == number of live owners; free queue == exactly the unreferenced blocks
        owners = np.zeros(40, dtype=int)
        for r in list(s.prefilling) + [x for x in s.rows if x is not None]:
            for b in r.blocks:
                owners[b] += 1
        assert owners[1:].tolist() == s.pool._ref[1:]
        assert set(s.pool._free) == {b for b in range(1, 40) if owners[b] == 0}
        for d, b in s.pool._cached.items():
            assert s.pool._key[b] == d
    assert nreq == 120 and not s.has_work() and s.pool.hits > 0
    assert all(r.finished for r in live.values())


# ------------------------------------------------------------------ property test (hypothesis)
from hypothesis import HealthCheck, given, settings, strategies as st  # noqa: E402


@settings(max_examples=60, deadline=None, suppress_health_check=[HealthCheck.too_slow])
@given(seed=st.integers(0, 10_000), num_blocks=st.integers(12, 40), max_seqs=st.integers(1, 6),
       budget=st.sampled_from([16, 48, 64, 200]), caching=st.booleans(), abort_p=st.floats(0.0, 0.2),
       spec_k=st.sampled_from([0, 0, 2, 4]), host_slots=st.sampled_from([0, 0, 3, 40]))
def test_scheduler_invariants_under_random_load(seed, num_blocks, max_seqs, budget, caching, abort_p, spec_k, host_slots):
    """Random arrivals / aborts / EOS under KV pressure (forces chunking and preemption): every step
    respects the token and sequence budgets, chunks continue exactly where the request stopped, block
    tables of resident rows mirror the requests' block lists, no block is owned twice unless it is a
    shared (cached, full, read-only) prefix block, and everything is returned at the end."""
    rng = np.random.RandomState(seed)
    BS, max_len = 16, 160
    # speculative decoding (random drafts, randomly right) and the host KV tier ride along in some examples: the
    # same budgets / ownership / leak invariants must hold with verification steps and swap lists in the mix
    spec = SpeculativeConfig(num_speculative_tokens=spec_k) if spec_k else None
    s = Scheduler(num_blocks, BS, max_seqs, budget, max_len, 2, enable_prefix_caching=caching, speculative=spec,
                  max_sample_rows=max_seqs * (1 + spec_k), kv_host_slots=host_slots)
    if spec:
        s.proposer = lambda r, k: rng.choice([7, 9], size=int(rng.randint(0, k + 1))).tolist()
    heads = [rng.randint(3, 300, size=n).tolist() for n in (16, 40)]
    live, nreq, steps, finished, mixed_seen = {}, 0, 0, 0, 0

    def check_ownership():
        # ownership: a block has several owners only if it is a cached full block inside every owner's computed prefix
        owners = {}
        for r in list(s.prefilling) + [x for x in s.rows if x is not None]:
            assert not r.finished
            for j, b in enumerate(r.blocks):
                assert 0 < b < num_blocks
                owners.setdefault(b, []).append((r, j))
        for b, lst in owners.items():
            assert s.pool._ref[b] == len(lst) and b not in s.pool._free
            if len(lst) > 1:
                assert caching and s.pool._key[b] is not None
                assert len({j for _, j in lst}) == 1                        # same logical position in every owner
        assert len(s.pool._free) + len(owners) == num_blocks - 1

    while steps < 600:
        steps += 1
        if nreq < 25 and rng.rand() < 0.4:
            n_prompt = int(rng.randint(1, 90))
            head = heads[rng.randint(2)][: rng.randint(0, 41)] if rng.rand() < 0.5 else []
            prompt = (head + rng.randint(3, 300, size=n_prompt).tolist())[: max_len - 2]
            max_new = int(rng.randint(1, max(2, min(40, max_len - len(prompt)))))
            need_blocks = (len(prompt) + max_new + BS) // BS
            if need_blocks >= num_blocks - 1:
                continue                                       # the engine rejects these up front
            r = Request(f"h{nreq}", prompt, SamplingParams(max_tokens=max_new, ignore_eos=bool(rng.rand() < 0.7)))
            s.add(r); live[r.req_id] = r; nreq += 1
        if live and rng.rand() < abort_p:
            rid = list(live)[rng.randint(len(live))]
            if not live[rid].finished:
                assert s.abort(rid)
                assert live[rid].finished and live[rid].finish_reason == "abort" and not live[rid].blocks
        plan = s.schedule()
        for kind, blk, slot in s.pool.take_swap_ops():          # host tier: well-formed operations only
            assert kind in (0, 1) and 0 < blk < num_blocks and 0 <= slot < host_slots
        if plan is None:
            if nreq >= 25 and not s.has_work():
                break
            continue
        if isinstance(plan, SpecPlan):
            B = plan.batch
            assert B == len(s.rows) and len(plan.drafts) == B and any(plan.drafts)
            assert plan.num_tokens <= budget and plan.num_tokens <= s.max_sample_rows
            sampled = []
            for i, r in enumerate(s.rows):
                d = plan.drafts[i]
                assert len(d) <= spec_k and len(d) <= r.params.max_tokens - len(r.output_ids) - 1
                assert (int(s.num_computed[i]) + len(d)) // BS < len(r.blocks)          # K/V slots for every draft
                assert int(s.n_blocks[i]) == len(r.blocks) and s.block_table[i, : len(r.blocks)].tolist() == r.blocks
                # the "model": agrees with a draft with probability 1/2, otherwise says 2 (EOS), 7 or 9
                sampled += [int(t) if rng.rand() < 0.5 else int(rng.choice([2, 7, 9])) for t in d] + [int(rng.choice([7, 9]))]
            before = {r.req_id: len(r.output_ids) for r in s.rows}
            out = s.update_spec(plan, sampled)
            finished += sum(f is not None for f in out.finished)
            for rid in set(out.req_ids):
                emitted = out.req_ids.count(rid)
                assert 1 <= emitted <= spec_k + 1 and len(live[rid].output_ids) == before[rid] + emitted
            check_ownership()
            continue
        pre = plan.prefill if isinstance(plan, MixedPlan) else plan if isinstance(plan, PrefillPlan) else None
        B = 0 if isinstance(plan, PrefillPlan) else plan.batch
        if pre is not None:
            assert 0 < pre.num_tokens and pre.num_tokens + B <= budget      # decode tokens share the budget
            assert len(pre.reqs) == len(set(id(r) for r in pre.reqs))
            assert len(s.rows) + len(s.prefilling) <= max_seqs
            assert not any(r in s.rows for r in pre.reqs)
            for r, st_, n, smp in zip(pre.reqs, pre.starts, pre.lens, pre.sample):
                assert n > 0 and st_ == r.num_computed and st_ + n <= r.prefill_target
                assert smp == (st_ + n == r.prefill_target)
                assert len(r.blocks) * BS >= r.prefill_target + 1            # room for the first decode slot
                if caching and st_ > 0 and st_ % BS == 0 and r.num_cached_tokens:
                    assert r.num_cached_tokens <= r.prefill_target - 1
        if B:
            assert 0 < B <= max_seqs and B == len(s.rows)
            for i, r in enumerate(s.rows):
                assert r.row == i and int(s.n_blocks[i]) == len(r.blocks)
                assert s.block_table[i, : len(r.blocks)].tolist() == r.blocks
                pos = int(s.num_computed[i])
                assert pos // BS < len(r.blocks) and pos == len(r.prompt_ids) + len(r.output_ids) - 1
                assert int(s.last_token[i]) == r.all_ids[-1]
        ptoks = rng.choice([2, 7, 9], size=sum(pre.sample), p=[0.1, 0.45, 0.45]).tolist() if pre is not None else []
        dtoks = rng.choice([2, 7, 9], size=B, p=[0.1, 0.45, 0.45]).astype(np.int32)
        if isinstance(plan, MixedPlan):
            mixed_seen += 1
            out = s.update_mixed(plan, ptoks, dtoks)
        elif pre is not None:
            out = s.update_prefill(plan, ptoks)
        else:
            out = s.update_decode(plan, dtoks)
        finished += sum(f is not None for f in out.finished)
        check_ownership()
    assert not s.has_work() or steps >= 600
    if not s.has_work():
        assert s.pool.num_free == s.pool.num_total
        for r in live.values():
            assert r.finished and (r.finish_reason in ("stop", "length", "abort"))
            if r.finish_reason == "length":
                assert len(r.output_ids) == r.params.max_tokens
            if r.finish_reason == "stop":
                assert r.output_ids[-1] == 2 and not r.params.ignore_eos
