"""CPU unit tests: id allocator, paged KV + prefix cache invariants, scheduler policies, batch builder."""
import random

import numpy as np
import pytest

from gllm_b200.id_allocator import IDAllocator
from gllm_b200.input_data import BatchArrays, build_batch
from gllm_b200.memory_manager import MemoryManager, PrefixMemoryManager
from gllm_b200.scheduler import (ScheduledSeq, Scheduler, balanced_decode_budget, kv_headroom_tokens,
                                 throttled_prefill_budget)
from gllm_b200.sequence import Sequence


def mkseq(i, n, out=4, **kw):
    return Sequence(i, list(range(1000 * i, 1000 * i + n)), [2], output_len=out, ignore_eos=True, **kw)


# ------------------------------------------------------------------------------------------------
def test_id_allocator_fifo_tail_and_specific():
    a = IDAllocator(0, 7)
    assert [a.allocate() for _ in range(3)] == [0, 1, 2]
    a.free(1)
    # freed id goes to the TAIL: 3..7 are handed out before 1 comes back
    assert [a.allocate() for _ in range(5)] == [3, 4, 5, 6, 7]
    assert a.allocate() == 1
    assert a.get_num_free_ids() == 0
    with pytest.raises(RuntimeError):
        a.allocate()
    a.free(5); a.free(2); a.free(7)
    assert a.allocate(2) == 2 and not a.is_free(2)      # O(1) specific re-acquire from the middle
    assert a.allocate(2) == 2                            # already taken: no-op (shared page)
    assert [a.allocate(), a.allocate()] == [5, 7]
    with pytest.raises(RuntimeError):
        a.free(3); a.free(3)


def test_id_allocator_random_against_model():
    rnd = random.Random(0)
    a = IDAllocator(10, 59)
    model_free = list(range(10, 60))
    used = set()
    for _ in range(5000):
        r = rnd.random()
        if r < 0.45 and model_free:
            x = a.allocate()
            assert x == model_free.pop(0)
            used.add(x)
        elif r < 0.6 and model_free:
            x = rnd.choice(model_free)
            assert a.allocate(x) == x
            model_free.remove(x)
            used.add(x)
        elif used:
            x = rnd.choice(sorted(used))
            used.remove(x)
            a.free(x)
            model_free.append(x)
        assert a.get_num_free_ids() == len(model_free)


# ------------------------------------------------------------------------------------------------
def test_memory_manager_alloc_free():
    mm = MemoryManager(10, 4, reserve_dummy_page=True)
    assert mm.dummy_page == 9 and mm.get_num_free_pages() == 9
    s = mkseq(1, 10)
    s.scheduled_token_num = 10
    mm.pre_allocate_page([s])
    assert len(s.page_table) == 3 and mm.get_num_free_pages() == 6
    s.scheduled_token_num = 13
    mm.pre_allocate_page([s])
    assert len(s.page_table) == 4
    mm.free(s)
    assert mm.get_num_free_pages() == 9 and s.page_table == []


def test_prefix_cache_hit_refcount_eviction():
    ps = 4
    mm = PrefixMemoryManager(8, ps)
    a = Sequence(1, list(range(10)), [2])
    mm.pre_allocate_computed_page([a])
    assert a.computed_token_num == 0
    a.scheduled_token_num = 10
    mm.pre_allocate_page([a])
    assert len(a.page_table) == 3
    # nothing is cacheable before the chunk has run: an identical prompt arriving now shares nothing
    early = Sequence(9, list(range(10)), [2])
    mm.pre_allocate_computed_page([early])
    assert early.page_table == [] and early.computed_token_num == 0
    a.computed_token_num = 10
    mm.publish_computed(a)
    # identical prefix: two full pages are shared, refcount 2
    b = Sequence(2, list(range(10)), [2])
    mm.pre_allocate_computed_page([b])
    assert b.page_table == a.page_table[:2] and b.computed_token_num == 8 == b.scheduled_token_num
    assert mm.page_ref[a.page_table[0]] == 2
    assert mm.get_cache_hit_rate() > 0
    # a prompt that is an exact multiple of the page size never gets its LAST page from the cache
    c = Sequence(3, list(range(8)), [2])
    mm.pre_allocate_computed_page([c])
    assert c.computed_token_num == 4
    mm.free(c)
    free_before = mm.get_num_free_pages()
    mm.free(a)
    # shared pages stay allocated (b still holds them); a's private third page is released
    assert mm.get_num_free_pages() == free_before + 1
    mm.free(b)
    assert mm.get_num_free_pages() == 8
    # freed pages keep their hash until they are handed out again ...
    d = Sequence(4, list(range(10)), [2])
    mm.pre_allocate_computed_page([d])
    assert d.computed_token_num == 8
    mm.free(d)
    # ... and lose it on re-allocation
    for i in range(8):
        mm.allocate_page()
    assert not mm.hash2page
    e = Sequence(5, list(range(10)), [2])
    for p in range(8):
        mm.free_page(p)
    mm.pre_allocate_computed_page([e])
    assert e.computed_token_num == 0


def test_prefix_cache_decode_page_registered():
    ps = 4
    mm = PrefixMemoryManager(16, ps)
    a = Sequence(1, [1, 2, 3, 4, 5, 6], [2])
    a.scheduled_token_num = 6
    mm.pre_allocate_page([a])
    a.computed_token_num = 6
    a.append(7)  # decode
    a.append(8)  # -> 8 tokens: second page full
    a.scheduled_token_num = 8
    mm.pre_allocate_page([a])
    a.computed_token_num = 8      # the decode steps that wrote tokens 6 and 7 have returned
    mm.publish_computed(a)
    b = Sequence(2, [1, 2, 3, 4, 5, 6, 7, 8, 9], [2])
    mm.pre_allocate_computed_page([b])
    assert b.computed_token_num == 8


def test_prefix_cache_never_serves_pages_that_were_not_computed():
    """a sequence aborted (or stall-broken) between two prefill chunks left the hashes of pages it had only
    allocated; a retry of the same prompt then skipped tokens whose KV was never written."""
    ps = 16
    mm = PrefixMemoryManager(64, ps, reserve_dummy_page=True)
    sch = Scheduler(mm, maxp=40, maxd=8, page_size=ps, log=False, kvthresh=0.0)
    prompt = [(7 * i) % 101 + 3 for i in range(100)]
    a = Sequence(1, prompt, [2], output_len=4)
    sch.add_new_requests([a])
    batch = sch.schedule_once()
    assert batch[0].n == 40 and not batch[0].emits
    sch.add_next_tokens([])
    sch.process_output()
    assert a.computed_token_num == 40 and a.published == 2        # tokens 32..39 sit in a page that is not full
    sch.add_abort_ids([1])
    out = sch.check_abort_seqs()
    assert out.free_ids == [1] and not sch.seqs_to_prefill
    b = Sequence(2, prompt, [2], output_len=4)
    sch.add_new_requests([b])
    batch = sch.schedule_once()
    assert batch[0].start == 32, batch      # only the two pages that were really written are reused


def test_abort_between_prefill_chunks_reports_once_and_leaves_no_stale_id():
    """(a) a sequence flagged while its chunk is in flight is freed and reported exactly once and never
    scheduled again; (b) an abort id that matches nothing alive is dropped instead of lingering (it disabled
    lookahead scheduling and hit the next request re-using the id)."""
    ps = 16
    mm = PrefixMemoryManager(64, ps, reserve_dummy_page=True)
    sch = Scheduler(mm, maxp=16, maxd=8, page_size=ps, log=False, kvthresh=0.0)
    a = Sequence(0, list(range(3, 63)), [2], output_len=4)
    sch.add_new_requests([a])
    assert sch.schedule_once()
    a.is_abort = True                       # worst case: flagged from outside, still queued, chunk in flight
    sch.add_abort_ids([0])
    assert sch.check_abort_seqs() is None   # in flight: nothing to report yet
    sch.add_next_tokens([])
    out = sch.process_output()
    assert out.free_ids == [0] and not a.page_table
    assert not sch.seqs_to_prefill and not sch.has_work() and not sch.abort_ids
    assert mm.get_num_free_pages() == 63
    assert not sch.schedule_once()
    sch.add_abort_ids([0, 77])              # late duplicates / unknown ids
    assert sch.check_abort_seqs() is None and not sch.abort_ids
    b = Sequence(0, list(range(3, 20)), [2], output_len=2)     # the id is re-used by a new request
    sch.add_new_requests([b])
    run_engine(sch)
    assert b.num_output_tokens == 2 and not b.is_abort


# ------------------------------------------------------------------------------------------------
def test_budget_functions():
    assert balanced_decode_budget(3, 4, 100) == 1
    assert balanced_decode_budget(5, 4, 100, rnd=0) == 1 and balanced_decode_budget(5, 4, 100, rnd=3) == 2
    assert balanced_decode_budget(1000, 2, 64, rnd=1) == 64
    assert kv_headroom_tokens(100, 5, 16) == 95 * 16 and kv_headroom_tokens(3, 5, 16) == 0
    # single GPU: plain cap
    assert throttled_prefill_budget(10 ** 6, 1, 0.5, 0.05, 2048, 32, 8, 10, 10 ** 5) == 2048
    # UT: free ratio 0.525 -> ratio 0.5 -> 1024
    assert throttled_prefill_budget(10 ** 6, 4, 0.525, 0.05, 2048, 32, 8, 1, 10 ** 5) == 1024
    # WT: 800 waiting tokens over 8 iterations -> 100; floor minp
    assert throttled_prefill_budget(10 ** 6, 4, 1.0, 0.05, 2048, 32, 8, 3, 800) == 100
    assert throttled_prefill_budget(10 ** 6, 4, 1.0, 0.05, 2048, 32, 8, 3, 80) == 32
    assert throttled_prefill_budget(0, 4, 1.0, 0.05, 2048, 32, 8, 3, 800) == 0


def run_engine(sch: Scheduler, n_steps=10000, token=lambda e: 7):
    """Drive the scheduler with an instant 'model' honouring the <= pp_size in-flight invariant."""
    inflight = []
    produced = {}
    batches = []
    for _ in range(n_steps):
        if not sch.has_work():
            break
        b = sch.schedule_once()
        if b:
            assert len(sch.batch_running) <= sch.pp_size
            # decode-first ordering
            kinds = [e.is_decode for e in b]
            assert kinds == sorted(kinds, reverse=True), kinds
            inflight.append(b)
            batches.append(b)
        if inflight and (len(inflight) == sch.pp_size or not b):
            done = inflight.pop(0)
            sch.add_next_tokens([token(e) for e in done if e.emits])
            out = sch.process_output()
            for sid, tok in zip(out.act_schedule_ids, out.next_tokens):
                produced.setdefault(sid, []).append(tok)
    return produced, batches


@pytest.mark.parametrize("method", ["chunked_prefill", "split_pd", "token_throttling"])
@pytest.mark.parametrize("pp", [1, 2, 4])
def test_scheduler_completes_all(method, pp):
    mm = PrefixMemoryManager(64, 16)
    sch = Scheduler(mm, pp_size=pp, world_size=pp, schedule_method=method, maxd=8, maxp=64, minp=8, iterp=4,
                    kvthresh=0.05, page_size=16, log=False)
    seqs = [mkseq(i, n, out=5) for i, n in enumerate([10, 100, 33, 64, 7, 150])]
    sch.add_new_requests(seqs)
    produced, batches = run_engine(sch)
    assert all(len(produced[s.seq_id]) == 5 for s in seqs)
    assert mm.get_num_free_pages() == 64
    for b in batches:
        assert sum(e.n for e in b) <= (64 if method != "token_throttling" else 64 + 8)


def test_chunk_continuation_and_pipelined_chunks():
    mm = MemoryManager(64, 16)
    sch = Scheduler(mm, pp_size=2, world_size=2, schedule_method="chunked_prefill", maxd=8, maxp=32, page_size=16,
                    log=False)
    s = mkseq(1, 100, out=2)
    sch.add_new_requests([s])
    b1 = sch.schedule_once()
    b2 = sch.schedule_once()  # second chunk scheduled while the first is still in flight (pp=2)
    assert [(e.start, e.n) for e in b1] == [(0, 32)] and [(e.start, e.n) for e in b2] == [(32, 32)]
    assert sch.schedule_once() == []  # pp_size batches in flight
    sch.add_next_tokens([]); out = sch.process_output()
    assert out.act_schedule_ids == [] and s.computed_token_num == 32
    b3 = sch.schedule_once()
    assert [(e.start, e.n) for e in b3] == [(64, 32)]


def test_preemption_recompute():
    mm = MemoryManager(6, 4)  # tiny KV: forces preemption
    sch = Scheduler(mm, pp_size=1, schedule_method="chunked_prefill", maxd=8, maxp=64, kvthresh=0.0, page_size=4,
                    log=False)
    seqs = [mkseq(i, 7, out=10) for i in range(3)]
    sch.add_new_requests(seqs)
    produced, _ = run_engine(sch)
    assert all(len(produced[s.seq_id]) >= 10 for s in seqs) or sch.num_preempt_seqs > 0
    assert sch.num_preempt_seqs > 0
    assert mm.get_num_free_pages() == 6
    for s in seqs:
        assert len(s.token_ids) == 7 + 10


def test_abort_frees_pages():
    mm = MemoryManager(32, 16)
    sch = Scheduler(mm, pp_size=1, maxd=8, maxp=64, page_size=16, log=False)
    a, b = mkseq(1, 20, out=50), mkseq(2, 20, out=50)
    sch.add_new_requests([a, b])
    bt = sch.schedule_once()
    sch.add_next_tokens([5, 5]); sch.process_output()
    sch.add_abort_ids([1])
    out = sch.check_abort_seqs()
    assert out.free_ids == [1]
    produced, _ = run_engine(sch)
    assert len(produced[2]) == 49 and 1 not in produced
    assert mm.get_num_free_pages() == 32


def test_split_pd_prioritises_prefill():
    mm = MemoryManager(64, 16)
    sch = Scheduler(mm, pp_size=1, schedule_method="split_pd", maxd=8, maxp=32, kvthresh=0.05, page_size=16, log=False)
    sch.add_new_requests([mkseq(1, 10, out=5)])
    b = sch.schedule_once(); sch.add_next_tokens([1]); sch.process_output()
    sch.add_new_requests([mkseq(2, 10, out=5)])
    b = sch.schedule_once()
    assert [e.seq.seq_id for e in b] == [2]  # decode of seq 1 is held back while a prefill waits


# ------------------------------------------------------------------------------------------------
def test_build_batch_layout_and_wire():
    ps = 4
    a = mkseq(1, 6, out=5)   # will be a decode entry
    a.page_table = [3, 5]
    a.computed_token_num = 6; a.append(99)
    b = mkseq(2, 10, out=5)  # prefill chunk [4, 10) with context
    b.page_table = [7, 1, 2]
    ents = [ScheduledSeq(a, 6, 1), ScheduledSeq(b, 4, 6)]
    bt = build_batch(ents, ps, vocab_size=50000, batch_id=3)
    assert bt.num_decode_seqs == 1 and bt.num_seqs == 2 and bt.num_tokens == 7
    assert bt.tokens.tolist() == [99] + list(range(2004, 2010))
    assert bt.positions.tolist() == [6, 4, 5, 6, 7, 8, 9]
    assert bt.slot_mapping.tolist() == [5 * 4 + 2, 1 * 4 + 0, 1 * 4 + 1, 1 * 4 + 2, 1 * 4 + 3, 2 * 4 + 0, 2 * 4 + 1]
    assert bt.seq_lens.tolist() == [7, 10] and bt.query_start_loc.tolist() == [0, 1, 7]
    assert bt.logits_idx.tolist() == [0, 6] and bt.block_table[1].tolist() == [7, 1, 2]
    assert bt.max_q_len == 6 and bt.max_seq_len == 10 and bt.all_greedy is False  # default top_k=10
    hdr, bufs = bt.to_wire()
    back = BatchArrays.from_wire(hdr, [memoryview(np.ascontiguousarray(x)) for x in bufs])
    for f in ("tokens", "positions", "slot_mapping", "block_table", "seq_lens", "query_start_loc", "logits_idx",
              "temperature", "top_k", "top_p"):
        assert np.array_equal(getattr(bt, f), getattr(back, f)), f
    assert back.batch_id == 3 and back.num_decode_seqs == 1


def test_incremental_decode_batch_equals_full_rebuild():
    """The vectorised decode fast path (batch N+1 derived from batch N, rows possibly permuted, pages crossing a
    boundary) must produce exactly the arrays of a from-scratch build."""
    import numpy as np
    from gllm_b200.input_data import build_batch
    from gllm_b200.scheduler import ScheduledSeq
    from gllm_b200.sequence import Sequence
    rng = np.random.default_rng(0)
    page = 16
    seqs = []
    for i in range(13):
        n = int(rng.integers(20, 90))
        s = Sequence(i, rng.integers(3, 200, size=n).tolist(), [1], 500, True, 0.7 if i % 3 else 0.0, 0.9, 5 if i % 3 else 1,
                     1.0)
        s.prompt_len = n - 3
        s.page_table = [int(x) for x in rng.permutation(4000)[: (n + page) // page + 1]]
        s.slot = i + 1
        seqs.append(s)
    order = list(range(len(seqs)))
    prev = None
    for step in range(40):
        order.reverse()                      # the scheduler re-queues head-first: row order flips every step
        ents = [ScheduledSeq(seqs[j], len(seqs[j].token_ids) - 1, 1) for j in order]
        for e in ents:                       # make sure the page for the new token exists
            need = e.start // page + 1
            while len(e.seq.page_table) < need:
                e.seq.page_table.append(int(rng.integers(4000, 8000)))
        fast = build_batch(ents, page, 1000, step, prev=prev)
        full = build_batch(ents, page, 1000, step)
        for name in ("tokens", "positions", "slot_mapping", "seq_lens", "query_start_loc", "logits_idx", "emit_seq",
                     "temperature", "top_k", "top_p", "rep_penalty", "state_slot"):
            assert np.array_equal(getattr(fast, name), getattr(full, name)), (step, name)
        w = full.block_table.shape[1]
        assert np.array_equal(fast.block_table[:, :w], full.block_table)
        assert (fast.num_seqs, fast.num_decode_seqs, fast.num_tokens, fast.max_seq_len, fast.all_greedy) == \
               (full.num_seqs, full.num_decode_seqs, full.num_tokens, full.max_seq_len, full.all_greedy)
        prev = fast
        for s in seqs:
            s.token_ids.append(int(rng.integers(3, 200)))


def test_oversized_requests_do_not_hang_the_scheduler():
    """A prompt that can never be resident is rejected (finished with no output) and an output budget that the
    pool cannot hold is capped — neither may block the queue forever."""
    from gllm_b200.memory_manager import MemoryManager
    from gllm_b200.scheduler import Scheduler
    from gllm_b200.sequence import Sequence
    mm = MemoryManager(8, 16, reserve_dummy_page=True)          # 7 usable pages = 112 tokens
    sch = Scheduler(mm, maxp=32, maxd=8, page_size=16, log=False)
    big = Sequence(1, list(range(200)), [1], 4, True)
    greedy = Sequence(2, list(range(40)), [1], 500, True)
    sch.add_new_requests([big, greedy])
    assert greedy.output_len == 112 - 40
    out = sch.check_abort_seqs()
    assert out is not None and out.free_ids == [1] and big not in sch.seqs_to_prefill
    ents = sch.schedule_once()
    assert ents and ents[0].seq is greedy
    # the prefill budget excludes the kvthresh reserve: a prompt that only fits into the reserve is rejected too
    mm = MemoryManager(20, 16)                                  # 20 pages, reserve = int(0.25 * 20) = 5 pages
    sch = Scheduler(mm, maxp=64, maxd=8, page_size=16, kvthresh=0.25, log=False)
    fits, too_big = Sequence(3, list(range(15 * 16 - 1)), [1], 4, True), Sequence(4, list(range(15 * 16)), [1], 4, True)
    sch.add_new_requests([too_big, fits])
    out = sch.check_abort_seqs()
    assert out is not None and out.free_ids == [4]
    for _ in range(40):                                         # the admitted one really gets through its prefill
        ents = sch.schedule_once()
        assert ents, "admitted prompt stalled"
        sch.add_next_tokens([5] * sum(e.emits for e in ents))
        sch.process_output()
        if fits.computed_token_num >= fits.prompt_len:
            break
    assert fits.computed_token_num >= fits.prompt_len


@pytest.mark.parametrize("tp", [1, 2, 4, 8])
def test_weight_sharding_round_trip(tp):
    """Random HF-layout tensors -> per-rank shards (reference: gllm/models/weight_utils.py:6-84) -> reassembled ==
    original, for QKV (incl. KV-head replication when tp > kv heads), gate/up, row / column splits, the padded
    vocabulary, contiguous expert blocks and the PP layer partition."""
    import torch
    from gllm_b200.models import weight_utils as wu
    from gllm_b200.parallel.state import partition_layers
    torch.manual_seed(tp)
    heads, kv_heads, d, h, inter, vocab = 8, 2, 4, 16, 24 * 8, 1000
    q, k, v = torch.randn(heads * d, h), torch.randn(kv_heads * d, h), torch.randn(kv_heads * d, h)
    hq = heads // tp
    qs, ks, vs = [], {}, {}
    for r in range(tp):
        w = wu.shard_qkv(q, k, v, heads, kv_heads, d, r, tp)
        kv0, nkv = wu.kv_head_range(kv_heads, r, tp)
        assert w.shape[0] == (hq + 2 * nkv) * d
        qs.append(w[: hq * d])
        for j in range(nkv):
            ks[kv0 + j] = w[hq * d + j * d: hq * d + (j + 1) * d]
            vs[kv0 + j] = w[(hq + nkv) * d + j * d: (hq + nkv) * d + (j + 1) * d]
    assert torch.equal(torch.cat(qs), q)
    assert torch.equal(torch.cat([ks[i] for i in range(kv_heads)]), k)
    assert torch.equal(torch.cat([vs[i] for i in range(kv_heads)]), v)
    gate, up, down = torch.randn(inter, h), torch.randn(inter, h), torch.randn(h, inter)
    gu = [wu.shard_gate_up(gate, up, r, tp) for r in range(tp)]
    assert torch.equal(torch.cat([g[: inter // tp] for g in gu]), gate)
    assert torch.equal(torch.cat([g[inter // tp:] for g in gu]), up)
    assert torch.equal(torch.cat([wu.shard_cols(down, r, tp) for r in range(tp)], dim=1), down)
    emb = torch.randn(vocab, h)
    sh = [wu.shard_vocab(emb, r, tp) for r in range(tp)]
    full = torch.cat(sh)
    assert full.shape[0] == wu.pad_vocab(vocab, tp) and torch.equal(full[:vocab], emb) and not full[vocab:].any()
    for e_total in (8, 60, 257):
        covered = []
        for r in range(tp):
            s0, n = wu.expert_range(e_total, r, tp)
            covered += list(range(s0, s0 + n))
        assert covered == list(range(e_total))
    for layers in (36, 61, 7):
        for pp in (1, 2, 4):
            parts = partition_layers(layers, pp)
            assert [i for p in parts for i in p] == list(range(layers)) and all(len(p) > 0 for p in parts)
    assert [len(p) for p in partition_layers(64, 4, [16, 16, 17, 15])] == [16, 16, 17, 15]


# ---------------------------------------------------------------------------------------------------------
# utils: chat post-processing, model path resolution, helpers
# ---------------------------------------------------------------------------------------------------------
def test_glm_process_response_plain_and_tool_call():
    from gllm_b200.utils.chat import process_response
    hist = [{"role": "user", "content": "hi"}]
    content, h2 = process_response("ChatGLMModel", "\n hello there ", hist)
    assert content == "hello there" and h2[-1] == {"role": "assistant", "metadata": "", "content": "hello there"}
    assert len(hist) == 1                                           # input history is not mutated
    tools = [{"role": "system", "content": "tools", "tools": [{"name": "get_weather"}]},
             {"role": "user", "content": "weather?"}]
    out = "get_weather\n```python\ntool_call(city='Paris', days=3)\n```"
    content, h3 = process_response("ChatGLMModel", out, tools)
    assert content == {"name": "get_weather", "parameters": {"city": "Paris", "days": 3}}
    assert h3[-1]["metadata"] == "get_weather"
    # model output is never evaluated: a call expression inside an argument stays unparsed text
    content, _ = process_response("ChatGLMModel", "t\n```python\ntool_call(x=__import__('os').getcwd())\n```", tools)
    assert isinstance(content["parameters"], str)
    # without advertised tools the payload is passed through; other architectures append one assistant turn
    content, _ = process_response("ChatGLMModel", "interpreter\nprint(1)", hist)
    assert content == {"name": "interpreter", "content": "print(1)"}
    content, h4 = process_response("Qwen3ForCausalLM", "plain", hist)
    assert content == "plain" and h4[-1] == {"role": "assistant", "content": "plain"}


def test_resolve_model_path(tmp_path):
    from gllm_b200.utils import resolve_model_path
    assert resolve_model_path("preset:tiny") == "preset:tiny"
    assert resolve_model_path({"a": 1}) == {"a": 1}
    assert resolve_model_path(str(tmp_path)) == str(tmp_path)
    calls = []

    def fake_download(repo, cache_dir=None, allow_patterns=None):
        calls.append((repo, tuple(allow_patterns)))
        return str(tmp_path / "snap")

    assert resolve_model_path("Qwen/Qwen3-8B", cache_dir=str(tmp_path), _download=fake_download) == str(tmp_path / "snap")
    assert calls and calls[0][0] == "Qwen/Qwen3-8B" and "*.safetensors" in calls[0][1]
    with pytest.raises(FileNotFoundError):
        resolve_model_path("/no/such/dir/model")


def test_util_helpers():
    import asyncio
    import torch
    from gllm_b200 import utils
    assert utils.round_up(17, 16) == 32 and utils.round_down(17, 16) == 16 and utils.cdiv(17, 16) == 2
    assert utils.dtype_bytes(torch.bfloat16) == 2
    x = torch.tensor([1.0, float("inf"), float("-inf"), float("nan")], dtype=torch.float16)
    y = utils.clamp_overflow(x)
    assert torch.isfinite(y).all() and y[0] == 1 and y[1] > 6e4 and y[2] < -6e4
    assert utils.async_tensor_h2d([1, 2, 3], torch.int32, "cpu").tolist() == [1, 2, 3]

    async def run():
        return await utils.make_async(lambda a, b=0: a + b)(2, b=3)
    assert asyncio.run(run()) == 5


def test_shm_ring_broadcast_wrap_and_flow_control():
    """engine/shm_ring.py: every consumer sees every record in order, records never wrap (skip marker), the
    producer blocks instead of overwriting unread bytes, late attachers start from the beginning."""
    import os
    import uuid
    from gllm_b200.engine.shm_ring import RingReader, RingWriter
    name = f"gllm_b200_test_{uuid.uuid4().hex[:10]}"
    r_early = RingReader(name, 0)
    assert r_early.recv() is None                          # producer not there yet: lazy attach
    w = RingWriter(name, 2, capacity=32768)
    r_late = RingReader(name, 1)
    try:
        rng = random.Random(0)
        sent = []
        got = {0: [], 1: []}
        for i in range(600):
            big = i % 50 == 49
            payload = bytes([i % 251]) * (9000 if big else rng.randrange(1, 1500))    # 9000 > ring / 4: chained records
            kind = i % 2
            if big:       # a chained send must not time out half way (single-threaded test): make room first
                for k, r in ((0, r_early), (1, r_late)):
                    while (m := r.recv()) is not None:
                        got[k].append(m)
            while True:
                try:
                    w.send(payload, kind, timeout_s=0.0)
                    break
                except TimeoutError:                       # ring full: consumers must release first
                    for k, r in ((0, r_early), (1, r_late)):
                        m = r.recv()
                        if m is not None:
                            got[k].append(m)
            sent.append((kind, payload))
            if rng.random() < 0.5:
                m = r_early.recv()
                if m is not None:
                    got[0].append(m)
        for k, r in ((0, r_early), (1, r_late)):
            while True:
                m = r.recv()
                if m is None:
                    break
                got[k].append(m)
        assert got[0] == sent and got[1] == sent
        assert w.write > 5 * 32768                         # wrapped several times
    finally:
        r_early.close(); r_late.close(); w.close()
    assert not os.path.exists(f"/dev/shm/{name}")


def test_incremental_detokenizer_streams_exactly_the_final_text():
    """`Sequence.detokenize_inc` with a byte-level BPE (multi-byte characters split across tokens, random id
    sequences with invalid UTF-8): the concatenated deltas equal the one-shot decode; only a trailing partial
    character may still be held back."""
    pytest.importorskip("tokenizers")
    from tokenizers import Tokenizer, decoders, models, pre_tokenizers, trainers
    from transformers import PreTrainedTokenizerFast
    from gllm_b200.sequence import Sequence
    corpus = ["hello world, how are you today?", "naïve café — déjà vu ☕ 你好世界 こんにちは",
              "The quick brown fox jumps over the lazy dog 12345", "emoji 😀😃😄 test",
              "tabs\tand\nnewlines  double  spaces", "Ünïcödé strings ßtraße"] * 20
    tok = Tokenizer(models.BPE())
    tok.pre_tokenizer = pre_tokenizers.ByteLevel(add_prefix_space=False)
    tok.decoder = decoders.ByteLevel()
    tok.train_from_iterator(corpus, trainers.BpeTrainer(vocab_size=400,
                                                        initial_alphabet=pre_tokenizers.ByteLevel.alphabet()))
    fast = PreTrainedTokenizerFast(tokenizer_object=tok)
    rng = random.Random(0)
    for trial in range(120):
        text = " ".join(rng.choice(corpus) for _ in range(rng.randrange(1, 4)))
        ids = fast.encode(text)
        if rng.random() < 0.3:   # random ids too (invalid utf-8 byte sequences)
            ids = [rng.randrange(0, fast.vocab_size) for _ in range(rng.randrange(1, 40))]
        seq = Sequence(0, [1], [0], output_len=len(ids), ignore_eos=True)
        out, i = "", 0
        while i < len(ids):
            k = rng.randrange(1, 4)
            for t in ids[i:i + k]:
                seq.append(t)
            i += k
            out += seq.detokenize_inc(fast)
        full = fast.decode(ids, skip_special_tokens=True)
        # text still held back (trailing partial character) is allowed to be missing, nothing else
        assert full == out or (full.startswith(out) and "\ufffd" in fast.decode(ids[-4:])), (trial, full, out)


def test_vocab_parallel_sampling_model_matches_the_full_vocab_filter():
    """per-shard candidates + (max, sum-exp) statistics are enough to finish top-k / top-p / penalty
    sampling exactly — every draw lies in the support of the full-vocabulary filter, greedy rows give the argmax, and
    the empirical distribution of a row matches the filtered distribution."""
    import torch
    from gllm_b200.ops import ref
    torch.manual_seed(3)
    tp, per, b, c = 4, 96, 6, 16
    v_full = tp * per - 5                               # the last shard ends with padding columns
    logits = torch.randn(b, tp * per) * 2.5
    logits[:, v_full:] = 0.0
    temperature = torch.tensor([0.7, 1.0, 1.3, 0.0, 0.9, 1.0])
    top_k = torch.tensor([8, v_full, 3, 1, 16, v_full], dtype=torch.int32)
    top_p = torch.tensor([0.9, 0.6, 1.0, 1.0, 0.5, 1.0])
    pen = torch.tensor([1.0, 1.3, 1.0, 1.2, 1.0, 1.0])
    seen = torch.rand(b, tp * per) < 0.3
    probs = ref.sample_filter(logits[:, :v_full], temperature, top_k, top_p, pen, seen[:, :v_full])
    hist = torch.zeros(b, v_full)
    n_draw = 300
    for it in range(n_draw):
        g = torch.Generator().manual_seed(1000 + it)
        race = torch.empty(b, tp * per).exponential_(1.0, generator=g)
        recs = []
        for r in range(tp):
            lo = r * per
            valid = max(0, min(per, v_full - lo))
            recs.append(ref.vp_candidates(logits[:, lo:lo + per], valid, v_full, c, temperature, top_k, top_p, pen,
                                          seen[:, lo:lo + valid], race[:, lo:lo + valid], vocab_offset=lo))
        toks = ref.vp_final(torch.stack(recs), c, v_full, top_k, top_p, generator=g)
        assert int(toks.min()) >= 0 and int(toks.max()) < v_full
        hist[torch.arange(b), toks.long()] += 1
    assert bool((probs[hist > 0] > 0).all()), "drew a token outside the filtered support"
    x = ref.apply_penalty_temperature(logits[:, :v_full], temperature, pen, seen[:, :v_full])
    assert int(hist[3].argmax()) == int(x[3].argmax()) and hist[3].max() == n_draw          # greedy row
    for r in (0, 2, 4):                                                                       # small supports
        emp = hist[r] / n_draw
        assert float((emp - probs[r]).abs().max()) < 0.12, (r, emp[probs[r] > 0], probs[r][probs[r] > 0])
    # unfiltered rows (pure temperature sampling over the whole vocabulary): the shard races reproduce a full-vocab
    # Gumbel-max draw exactly
    g = torch.Generator().manual_seed(7)
    race = torch.empty(b, tp * per).exponential_(1.0, generator=g)
    recs = [ref.vp_candidates(logits[:, r * per:(r + 1) * per], max(0, min(per, v_full - r * per)), v_full, c,
                              temperature, top_k, top_p, pen, seen[:, r * per:r * per + max(0, min(per, v_full - r * per))],
                              race[:, r * per:r * per + max(0, min(per, v_full - r * per))], vocab_offset=r * per)
            for r in range(tp)]
    toks = ref.vp_final(torch.stack(recs), c, v_full, top_k, top_p, generator=g)
    full = (torch.log_softmax(x[5], -1) - torch.log(race[5, :v_full])).argmax()
    assert int(toks[5]) == int(full)


def test_small_t_threshold_rules(monkeypatch):
    """Fused TP: the replicated-rows form is used up to a threshold (default 64; `auto` scales it with the LL
    all-reduce's incoming traffic (tp-1)*T), and always when some rank would own no rows of the sharded layout."""
    import importlib
    import gllm_b200.parallel.fused as fused
    assert fused.small_threshold(2) == fused.small_threshold(8) == fused.SMALL_T == 64      # default: one value
    monkeypatch.setenv("GLLM_TP_SMALL_T", "auto")
    f2 = importlib.reload(fused)
    try:
        assert [f2.small_threshold(tp) for tp in (2, 4, 8)] == [64, 37, 16] and f2.SMALL_T == 64

        def small(t, tp):
            return t <= f2.small_threshold(tp) or (t <= f2.SMALL_T and (tp - 1) * ((t + tp - 1) // tp) >= t)
        for tp in (2, 4, 8):
            for t in range(1, 200):
                rpr = (t + tp - 1) // tp
                if not small(t, tp):
                    assert t > f2.small_threshold(tp)
                    assert (tp - 1) * rpr < t, (t, tp)      # every rank owns at least one row on the sharded path
        assert small(17, 8) and not small(32, 8) and not small(64, 8) and small(64, 2)
    finally:
        monkeypatch.delenv("GLLM_TP_SMALL_T")
        importlib.reload(fused)


def test_generate_takes_per_request_sampling_parameters():
    from gllm_b200 import LLM
    from gllm_b200.models.presets import tiny
    llm = LLM(tiny("Qwen3ForCausalLM", num_hidden_layers=1), load_format="dummy", device="cpu", maxp=32, maxd=8,
              num_cpu_pages=32, model_max_length=64, log_stats=False)
    outs = llm.generate(tokens=[[3, 4, 5], [6, 7]], output_lens=[3, 2], ignore_eos=True, temperature=[0.0, 0.9],
                        top_k=[1, 5], top_p=[1.0, 0.8], repetition_penalty=[1.0, 1.3])
    assert [len(s.token_ids) for s in outs] == [6, 4]
    assert (outs[0].top_k, outs[1].top_k, outs[1].repetition_penalty, outs[1].top_p) == (1, 5, 1.3, 0.8)
    assert not llm.worker._seq_slots                       # the penalised request gave its state row back
    llm.shutdown()
