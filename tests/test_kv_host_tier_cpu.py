"""Extended KV cache (host-memory tier of the prefix cache), host logic only: the BlockPool decides which
evicted blocks are copied out, which prompt prefixes are copied back, and in which ORDER -- the engine merely
executes `swap_ops` in order before each step.  A content model (what digest every device block / host slot
holds) replays those operations and checks that every block a request is told it may skip really holds the K/V
of exactly those tokens.  Launcher contract: gpustack/worker/backends/vllm.py:256-275,468-483."""
import numpy as np
import pytest

from gpustack_b200.kv_cache import BlockPool, block_digests
from gpustack_b200.scheduler import DecodePlan, MixedPlan, PrefillPlan, Request, SamplingParams, Scheduler

BS = 16


class ContentModel:
    """What is stored where, replayed from the pool's swap operations and the scheduler's plans."""

    def __init__(self, sched):
        self.s = sched
        self.dev, self.host = {}, {}
        self.loaded = self.stored = 0

    def run_swaps(self):
        for kind, blk, slot in self.s.pool.take_swap_ops():
            if kind == BlockPool.SWAP_OUT:
                assert blk in self.dev, "a block that was never written is being saved"
                self.host[slot] = self.dev[blk]
                self.stored += 1
            else:
                self.dev[blk] = self.host[slot]
                self.loaded += 1
        for d, slot in self.s.pool._host.items():      # what the pool believes each slot holds is what it holds
            assert self.host.get(slot) == d

    def check_plan(self, pre):
        for r, st in zip(pre.reqs, pre.starts):
            digs = block_digests(r.all_ids, BS, [], upto_blocks=st // BS)
            for j in range(st // BS):              # every skipped block must hold exactly these tokens' K/V
                assert self.dev.get(r.blocks[j]) == digs[j], (r.req_id, j)

    def snapshot(self, extra=()):
        live = {id(r): (r, list(r.blocks)) for r in list(self.s.rows) + list(self.s.prefilling) + list(extra)}
        return list(live.values())

    def after_step(self, snap):
        """Every full block below num_computed of a request that took part in the step now holds that
        request's tokens (prefill chunks and decode tokens alike; finished requests keep the count they
        were released with)."""
        for r, blocks in snap:
            digs = block_digests(r.all_ids, BS, [], upto_blocks=r.num_computed // BS)
            for j, d in enumerate(digs):
                if j < len(blocks):
                    self.dev[blocks[j]] = d


def drive(s, model, max_steps=4000):
    steps = 0
    while s.has_work() and steps < max_steps:
        steps += 1
        plan = s.schedule()
        model.run_swaps()                          # the engine runs them before the step's kernels
        if plan is None:
            continue
        if isinstance(plan, DecodePlan):
            snap = model.snapshot()
            s.update_decode(plan, np.full(plan.batch, 7, dtype=np.int32))
            model.after_step(snap)
            continue
        pre = plan.prefill if isinstance(plan, MixedPlan) else plan
        model.check_plan(pre)
        snap = model.snapshot(pre.reqs)
        if isinstance(plan, MixedPlan):
            s.update_mixed(plan, [7] * sum(pre.sample), np.full(plan.batch, 7, dtype=np.int32))
        else:
            s.update_prefill(plan, [7] * sum(pre.sample))
        model.after_step(snap)
    assert not s.has_work()


def test_evicted_prefix_comes_back_from_the_host_tier():
    s = Scheduler(14, BS, 2, 256, 512, None, enable_prefix_caching=True, kv_host_slots=32)
    m = ContentModel(s)
    a = list(range(100, 100 + 96))                 # 6 full blocks
    b = list(range(500, 500 + 160))                # 10 blocks: pushes A's blocks out of the 13-block device pool
    s.add(Request("a1", a + [1, 2, 3], SamplingParams(max_tokens=2, ignore_eos=True)))
    drive(s, m)
    s.add(Request("b1", b + [4], SamplingParams(max_tokens=2, ignore_eos=True)))
    drive(s, m)
    assert m.stored > 0 and s.pool.host_stored == m.stored          # A's keyed blocks were saved, not dropped
    before = m.loaded
    r = Request("a2", a + [9, 9, 9, 9], SamplingParams(max_tokens=2, ignore_eos=True))
    s.add(r)
    drive(s, m)
    assert m.loaded > before and s.pool.host_hits >= (m.loaded - before) * BS
    assert r.num_cached_tokens == 96               # the whole shared prefix was served without recomputation
    assert s.pool.num_free == s.pool.num_total


def test_without_a_host_tier_the_prefix_is_recomputed():
    s = Scheduler(14, BS, 2, 256, 512, None, enable_prefix_caching=True)
    m = ContentModel(s)
    a = list(range(100, 196))
    s.add(Request("a1", a + [1, 2, 3], SamplingParams(max_tokens=2, ignore_eos=True))); drive(s, m)
    s.add(Request("b1", list(range(500, 660)) + [4], SamplingParams(max_tokens=2, ignore_eos=True))); drive(s, m)
    r = Request("a2", a + [9, 9, 9, 9], SamplingParams(max_tokens=2, ignore_eos=True))
    s.add(r); drive(s, m)
    assert m.stored == 0 and m.loaded == 0 and r.num_cached_tokens < 96


@pytest.mark.parametrize("seed,host_slots,num_blocks", [(0, 4, 12), (1, 64, 12), (2, 2, 20), (3, 16, 30), (4, 1, 12)])
def test_random_shared_prefix_load_keeps_contents_right(seed, host_slots, num_blocks):
    """Random arrivals over a few shared prefixes with a device pool that keeps evicting and a host tier
    from tiny (every load competes with the evictions of its own admission: pinning) to roomy."""
    rng = np.random.RandomState(seed)
    s = Scheduler(num_blocks, BS, 3, 96, 160, None, enable_prefix_caching=True, kv_host_slots=host_slots)
    m = ContentModel(s)
    heads = [rng.randint(3, 400, size=n).tolist() for n in (48, 80, 64)]
    n_req = 0
    for _ in range(60):
        for _ in range(int(rng.randint(1, 4))):
            h = heads[rng.randint(3)][: int(rng.choice([16, 32, 48, 64, 80]))]
            tail = rng.randint(3, 400, size=int(rng.randint(1, 20))).tolist()
            p = (h + tail)[:140]
            if (len(p) + 4 + BS) // BS >= num_blocks - 1:
                continue
            s.add(Request(f"q{n_req}", p, SamplingParams(max_tokens=int(rng.randint(1, 4)), ignore_eos=True)))
            n_req += 1
        drive(s, m)
    assert s.pool.num_free == s.pool.num_total
    assert len(s.pool._host) + len(s.pool._host_free) == host_slots and not s.pool._host_pinned
    assert len(set(s.pool._host.values())) == len(s.pool._host)      # no slot owned twice
    if host_slots >= 16:
        assert m.loaded > 0


def test_cli_maps_the_launcher_arguments(monkeypatch):
    from gpustack_b200.cli import _engine_config, build_parser

    monkeypatch.setenv("LMCACHE_MAX_LOCAL_CPU_SIZE", "24")          # GiB, as _set_lmcache_env exports it
    a = build_parser().parse_args(["serve", "random:tiny", "--kv-transfer-config",
                                   '{"kv_connector":"LMCacheConnectorV1","kv_role":"kv_both"}'])
    assert _engine_config(a, 1).kv_host_cache_gib == 24.0
    a = build_parser().parse_args(["serve", "random:tiny", "--kv-host-cache-gib", "3.5"])
    assert _engine_config(a, 1).kv_host_cache_gib == 3.5
    a = build_parser().parse_args(["serve", "random:tiny", "--kv-transfer-config", '{"kv_connector":"NixlConnector"}'])
    assert _engine_config(a, 1).kv_host_cache_gib == 0.0
    assert _engine_config(build_parser().parse_args(["serve", "random:tiny"]), 1).kv_host_cache_gib == 0.0
