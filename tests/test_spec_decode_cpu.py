"""Speculative decoding, host side (no GPU): the n-gram proposer against a brute-force scan, the
launcher's --speculative-config contract (gpustack/worker/backends/vllm.py:333-367), and the scheduler's
draft / verify cycle against a deterministic stand-in model -- speculation must never change what a
request emits, only in how many steps."""
import json

import numpy as np
import pytest

from gpustack_b200.scheduler import DecodePlan, MixedPlan, PrefillPlan, Request, SamplingParams, Scheduler, SpecPlan
from gpustack_b200.spec_decode import NgramIndex, SpeculativeConfig, verify


# ------------------------------------------------------------------ proposer
def brute_force_draft(ids, n_min, n_max, k):
    """The reference proposer's definition: the longest suffix n-gram (n_max .. n_min) that occurred
    earlier; the draft is what followed its most recent earlier occurrence."""
    L = len(ids)
    for n in range(n_max, n_min - 1, -1):
        if L < n + 1:
            continue
        suf = ids[L - n:]
        for e in range(L - 1, n - 1, -1):            # e = end (exclusive) of the earlier occurrence
            if ids[e - n:e] == suf:
                return ids[e:e + k]
    return []


@pytest.mark.parametrize("seed", range(6))
def test_ngram_index_equals_brute_force_scan(seed):
    rng = np.random.RandomState(seed)
    n_min, n_max, k = int(rng.randint(1, 3)), int(rng.randint(3, 6)), int(rng.randint(1, 6))
    ids = rng.randint(0, 4, size=200).tolist()       # tiny vocabulary: many repeats
    idx = NgramIndex(n_min, n_max)
    pos = 0
    while pos < len(ids):
        step = int(rng.randint(1, 5))                # tokens arrive one or several at a time (accepted drafts)
        idx.extend(ids[pos:pos + step])
        pos = min(len(ids), pos + step)
        assert idx.propose(k) == brute_force_draft(ids[:pos], n_min, n_max, k), pos
    assert idx.propose(0) == []
    idx.truncate(50)                                 # preemption: rebuilt from scratch
    assert idx.propose(k) == brute_force_draft(ids[:50], n_min, n_max, k)


def test_launcher_json_contract():
    # exactly what _get_speculative_arguments composes for SpeculativeAlgorithmEnum.NGRAM
    text = json.dumps({"method": "ngram", "num_speculative_tokens": 5, "prompt_lookup_max": 6, "prompt_lookup_min": 3})
    c = SpeculativeConfig.parse(text)
    assert c.supported and (c.num_speculative_tokens, c.prompt_lookup_max, c.prompt_lookup_min) == (5, 6, 3)
    assert SpeculativeConfig.parse(None) is None and SpeculativeConfig.parse("") is None
    # draft-model methods parse but are flagged: the CLI then serves without speculation
    e = SpeculativeConfig.parse('{"method": "eagle3", "num_speculative_tokens": 2, "model": "/models/draft"}')
    assert e is not None and not e.supported
    assert not SpeculativeConfig.parse('{"method": "mtp"}').supported
    for bad in ('{"method": "ngram", "num_speculative_tokens": 0}', '{"method": "ngram", "prompt_lookup_max": 1, '
                '"prompt_lookup_min": 2}', "[1, 2]"):
        with pytest.raises(ValueError):
            SpeculativeConfig.parse(bad)
    with pytest.raises(ValueError):
        SpeculativeConfig.parse("{not json")


def test_cli_accepts_the_launcher_flag():
    from gpustack_b200.cli import _engine_config, build_parser

    a = build_parser().parse_args(["serve", "random:tiny", "--speculative-config",
                                   '{"method": "ngram", "num_speculative_tokens": 4, "prompt_lookup_max": 5}'])
    e = _engine_config(a, 1)
    assert e.num_speculative_tokens == 4 and e.speculative().prompt_lookup_max == 5
    a = build_parser().parse_args(["serve", "random:tiny", "--speculative-config", '{"method": "eagle3", "model": "/d"}'])
    assert _engine_config(a, 1).num_speculative_tokens == 0
    a = build_parser().parse_args(["serve", "random:tiny"])
    assert _engine_config(a, 1).speculative() is None


def test_verify_counts_the_matching_prefix():
    assert verify([], [9]) == 0
    assert verify([5, 6, 7], [5, 6, 7, 8]) == 3
    assert verify([5, 6, 7], [5, 9, 7, 8]) == 1       # a later coincidence does not count
    assert verify([5, 6, 7], [4, 6, 7, 8]) == 0


# ------------------------------------------------------------------ scheduler: draft / verify cycle
PERIOD = 11


def model_next(prev_tok, pos):
    """Stand-in target model: greedy next token as a function of the token it is fed (a cycle of
    PERIOD tokens starting at 100) -- a wrong draft therefore also yields a wrong token after it."""
    return 100 + ((prev_tok - 100 + 1) % PERIOD) if prev_tok >= 100 else 100 + pos % PERIOD


def run(spec, reqs, proposer=None, max_seqs=8, num_blocks=64, budget=64):
    s = Scheduler(num_blocks, 16, max_seqs, budget, 256, 104, speculative=spec,
                  max_sample_rows=max_seqs * (1 + (spec.num_speculative_tokens if spec else 0)))
    if proposer is not None:
        s.proposer = proposer
    for r in reqs:
        s.add(r)
    steps = {"prefill": 0, "decode": 0, "spec": 0}
    emitted = {r.req_id: [] for r in reqs}
    reasons = {}
    for _ in range(2000):
        plan = s.schedule()
        if plan is None:
            break
        if isinstance(plan, SpecPlan):
            steps["spec"] += 1
            B = plan.batch
            assert B == len(s.rows) and plan.num_tokens <= budget
            sampled = []
            for i in range(B):
                chunk = [int(s.last_token[i])] + plan.drafts[i]
                r = s.rows[i]
                # the drafts must fit in the row's blocks and in what the request may still emit
                assert (int(s.num_computed[i]) + len(chunk) - 1) // 16 < len(r.blocks)
                assert len(plan.drafts[i]) <= r.params.max_tokens - len(r.output_ids) - 1
                assert int(s.n_blocks[i]) == len(r.blocks) and s.block_table[i, :len(r.blocks)].tolist() == r.blocks
                sampled += [model_next(t, int(s.num_computed[i]) + j) for j, t in enumerate(chunk)]
            assert len(sampled) <= s.max_sample_rows
            out = s.update_spec(plan, sampled)
        elif isinstance(plan, DecodePlan):
            steps["decode"] += 1
            B = plan.batch
            toks = np.asarray([model_next(int(s.last_token[i]), int(s.num_computed[i])) for i in range(B)], dtype=np.int32)
            out = s.update_decode(plan, toks)
        else:
            steps["prefill"] += 1
            pre = plan.prefill if isinstance(plan, MixedPlan) else plan
            ptoks = [model_next(r.all_ids[st + n - 1], st + n - 1) for r, st, n, smp in
                     zip(pre.reqs, pre.starts, pre.lens, pre.sample) if smp]
            if isinstance(plan, MixedPlan):
                B = plan.batch
                dt = np.asarray([model_next(int(s.last_token[i]), int(s.num_computed[i])) for i in range(B)], dtype=np.int32)
                out = s.update_mixed(plan, ptoks, dt)
            else:
                out = s.update_prefill(plan, ptoks)
        for rid, tok, fin in zip(out.req_ids, out.token_ids, out.finished):
            emitted[rid].append(tok)
            if fin is not None:
                reasons[rid] = fin
        for i, r in enumerate(s.rows):           # resident rows: the cache holds everything but the last token
            assert int(s.num_computed[i]) == len(r.prompt_ids) + len(r.output_ids) - 1
            assert int(s.last_token[i]) == r.all_ids[-1] and r.row == i
    assert not s.has_work() and s.pool.num_free == s.pool.num_total
    for r in reqs:
        assert emitted[r.req_id] == r.output_ids
    return emitted, reasons, steps, s


def mk_reqs():
    cyc = [100 + i % PERIOD for i in range(40)]
    return [
        Request("a", cyc[:30], SamplingParams(max_tokens=40, ignore_eos=True)),            # periodic prompt: drafts hit
        Request("b", list(range(3, 20)), SamplingParams(max_tokens=25, ignore_eos=True)),  # enters the cycle after 1 token
        Request("c", cyc[3:25], SamplingParams(max_tokens=30)),                            # stops at EOS (104) mid-draft
        Request("d", cyc[:14], SamplingParams(max_tokens=9, ignore_eos=True, stop_token_ids=(109,))),
        Request("e", cyc[:20], SamplingParams(max_tokens=1, ignore_eos=True)),             # finishes in the prefill step
        Request("f", cyc[5:31], SamplingParams(max_tokens=12, ignore_eos=True, temperature=0.8, top_p=0.9)),  # nucleus: no drafts
    ]


def test_speculation_changes_step_count_not_output():
    base, base_reasons, base_steps, _ = run(None, mk_reqs())
    spec = SpeculativeConfig(num_speculative_tokens=4, prompt_lookup_max=4, prompt_lookup_min=2)
    got, reasons, steps, s = run(spec, mk_reqs())
    assert got == base and reasons == base_reasons
    assert base_steps["spec"] == 0 and steps["spec"] > 0
    assert steps["decode"] + steps["spec"] < base_steps["decode"]          # fewer forward passes
    assert s.spec_stats["accepted"] > 0 and s.spec_stats["accepted"] <= s.spec_stats["drafted"]
    assert reasons == {"a": "length", "b": "length", "c": "stop", "d": "stop", "e": "length", "f": "length"}


def test_wrong_and_partially_wrong_drafts_are_rejected():
    """An adversarial proposer: always k drafts, right for the first `good` tokens then wrong."""
    calls = {"n": 0}

    def proposer(r, k):
        calls["n"] += 1
        good = calls["n"] % 4                       # 0 .. 3 correct tokens, then garbage
        t, out = r.all_ids[-1], []
        for j in range(k):
            t = model_next(t, 0)
            out.append(t if j < good else 7)
        return out

    base, base_reasons, _, _ = run(None, mk_reqs())
    spec = SpeculativeConfig(num_speculative_tokens=3)
    got, reasons, steps, s = run(spec, mk_reqs(), proposer=proposer)
    assert got == base and reasons == base_reasons
    assert 0 < s.spec_stats["accepted"] < s.spec_stats["drafted"]


def test_drafts_respect_kv_pressure_and_batch_threshold():
    spec = SpeculativeConfig(num_speculative_tokens=4, disable_by_batch_size=3)
    # 1) batch threshold: with 3 running sequences speculation is off
    reqs = [Request(f"r{i}", [100 + j % PERIOD for j in range(20)], SamplingParams(max_tokens=20, ignore_eos=True))
            for i in range(3)]
    _, _, steps, _ = run(spec, reqs)
    assert steps["spec"] == 0 and steps["decode"] > 0
    # 2) a pool with no spare block: drafts are cut to the row's own blocks, nothing is preempted
    one = [Request("x", [100 + j % PERIOD for j in range(30)], SamplingParams(max_tokens=30, ignore_eos=True))]
    base, _, _, _ = run(None, [Request("x", list(one[0].prompt_ids), SamplingParams(max_tokens=30, ignore_eos=True))],
                        num_blocks=5)
    got, _, steps, s = run(spec, one, num_blocks=5)          # 4 usable blocks = 64 positions for 30 + 30 tokens
    assert got == base and s.num_preemptions == 0 and steps["spec"] > 0
