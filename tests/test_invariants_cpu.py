"""Property tests (hypothesis): random request streams through the real Scheduler + PrefixMemoryManager must
keep the paged-KV bookkeeping consistent at every step (allocator / prefix-cache refcount and
eviction invariants; scheduler budgets; abort and preemption paths)."""
import random

import numpy as np
import pytest

hypothesis = pytest.importorskip("hypothesis")
from hypothesis import HealthCheck, example, given, settings, strategies as st  # noqa: E402

from gllm_b200.input_data import build_batch  # noqa: E402
from gllm_b200.memory_manager import PrefixMemoryManager  # noqa: E402
from gllm_b200.scheduler import Scheduler  # noqa: E402
from gllm_b200.sequence import Sequence  # noqa: E402

PAGE = int(__import__("os").environ.get("GLLM_HYP_PAGE", "4"))


def check_invariants(mm: PrefixMemoryManager, live):
    """live: sequences that may hold pages (running, waiting with cached prefix, ...)."""
    holders = {}
    for s in live:
        assert len(set(s.page_table)) == len(s.page_table), "a sequence holds a page twice"
        for p in s.page_table:
            holders[p] = holders.get(p, 0) + 1
    for p in range(mm.num_pages):
        ref = mm.page_ref[p]
        assert ref >= 0
        assert ref == holders.get(p, 0), (p, ref, holders.get(p, 0))        # refcount == number of holders
        assert mm.id_allocator.is_free(p) == (ref == 0), p                  # free list <=> nobody holds it
    assert mm.get_num_free_pages() == mm.num_pages - len(holders)
    for h, p in mm.hash2page.items():                                       # hash maps are mutually consistent
        assert mm.page2hash[p] == h
    # two sequences share a page only if they agree on every token up to the end of that page
    owner = {}
    for s in live:
        for i, p in enumerate(s.page_table):
            if p in owner and owner[p][0] is not s:
                o, j = owner[p]
                assert i == j and o.token_ids[:(i + 1) * PAGE] == s.token_ids[:(i + 1) * PAGE]
            owner.setdefault(p, (s, i))
    # a cache entry is never ahead of the computation: a page that only this sequence holds and that did not come
    # from a prefix-cache hit is published only once its whole token range has been computed (hashes were
    # registered at allocation; an abort / stall-break between two chunks then left entries for unwritten KV)
    for s in live:
        hit_pages = s.num_cached_tokens // mm.page_size
        for i, p in enumerate(s.page_table):
            h = mm.page2hash[p]
            if h is not None and mm.hash2page.get(h) == p and holders[p] == 1 and i >= hit_pages:
                assert (i + 1) * mm.page_size <= s.computed_token_num, (s.seq_id, i, s.computed_token_num)


@settings(max_examples=int(__import__("os").environ.get("GLLM_HYP_EXAMPLES", "60")), deadline=None, derandomize=not __import__("os").environ.get("GLLM_HYP_RANDOM"), suppress_health_check=[HealthCheck.too_slow])
@given(seed=st.integers(0, 10 ** 6), n_req=st.integers(1, 14), pages=st.integers(10, 40),
       method=st.sampled_from(["chunked_prefill", "token_throttling", "split_pd"]),
       maxp=st.sampled_from([8, 16, 64]), abort_rate=st.sampled_from([0.0, 0.1]), pp=st.sampled_from([1, 2]),
       kvthresh=st.sampled_from([0.0, 0.1, 0.3]), page=st.sampled_from([2, 4]))
@example(seed=185, n_req=10, pages=10, method="split_pd", maxp=8, abort_rate=0.1, pp=1, kvthresh=0.0, page=2)
def test_random_streams_keep_kv_bookkeeping_consistent(seed, n_req, pages, method, maxp, abort_rate, pp, kvthresh,
                                                       page):
    global PAGE
    PAGE = page       # small pages: preemption + prefix-cache hits + page-boundary cases every few tokens
    rng = random.Random(seed)
    mm = PrefixMemoryManager(pages, PAGE)
    cap = 6 if method == "token_throttling" else maxp      # config.max_running_seqs of the engine
    sch = Scheduler(mm, pp_size=pp, world_size=pp, schedule_method=method, maxd=6, maxp=maxp, minp=4, iterp=2,
                    kvthresh=kvthresh, page_size=PAGE, log=False, max_seqs=cap)
    # prompts drawn from a few shared prefixes so that the prefix cache really gets hits and shared pages
    stems = [[rng.randrange(50) for _ in range(rng.randrange(2, 14))] for _ in range(3)]
    reqs = []
    for i in range(n_req):
        stem = rng.choice(stems)
        toks = stem[:rng.randrange(1, len(stem) + 1)] + [rng.randrange(50) for _ in range(rng.randrange(0, 6))]
        out = rng.randrange(1, 9)
        if (len(toks) + out + PAGE - 1) // PAGE > pages - 1 - int(kvthresh * pages):
            continue
        reqs.append(Sequence(i, toks, [2], output_len=out, ignore_eos=True))
    pending = list(reqs)
    everyone = list(reqs)
    produced = {}
    inflight = []
    prev_arrays, bid = None, 0
    for step in range(1200):
        if pending and rng.random() < 0.5:
            k = rng.randrange(1, len(pending) + 1)
            sch.add_new_requests(pending[:k])
            del pending[:k]
        if abort_rate and rng.random() < abort_rate and everyone:
            sch.add_abort_ids([rng.choice(everyone).seq_id])
            out = sch.check_abort_seqs()
        batch = sch.schedule_once()
        check_invariants(mm, everyone)
        if batch:
            assert len(sch.batch_running) <= pp                              # <= pp micro-batches in flight
            assert len(batch) <= cap                                         # fits the runner's per-sequence buffers
            n_dec = next((i for i, e in enumerate(batch) if not e.is_decode), len(batch))
            assert all(e.n == 1 and e.emits for e in batch[:n_dec])          # leading decode rows: one token each
            assert len({e.seq.seq_id for e in batch}) == len(batch)          # a sequence appears once per batch
            for e in batch:
                need = (e.start + e.n + PAGE - 1) // PAGE
                assert len(e.seq.page_table) >= need                         # KV room exists for what will run
            inflight.append(batch)
            # the incremental decode path of build_batch must equal a from-scratch build of the same entries
            bid += 1
            inc = build_batch(batch, PAGE, 64, bid, prev=prev_arrays if pp == 1 else None)
            full = build_batch(batch, PAGE, 64, bid, prev=None)
            for name in ("tokens", "positions", "slot_mapping", "seq_lens", "query_start_loc", "logits_idx",
                         "temperature", "top_k", "top_p", "rep_penalty", "emit_seq", "state_slot"):
                assert np.array_equal(getattr(inc, name), getattr(full, name)), name
            w = full.block_table.shape[1]
            for r in range(len(batch)):   # compare the pages each row really uses
                nblk = (int(full.seq_lens[r]) + PAGE - 1) // PAGE
                assert np.array_equal(inc.block_table[r, :nblk], full.block_table[r, :nblk])
            assert (inc.num_decode_seqs, inc.num_seqs, inc.num_tokens, inc.max_q_len, inc.max_seq_len) == \
                (full.num_decode_seqs, full.num_seqs, full.num_tokens, full.max_q_len, full.max_seq_len)
            prev_arrays = inc
        if inflight and (len(inflight) == pp or not batch):
            done = inflight.pop(0)
            sch.add_next_tokens([rng.randrange(3, 50) for e in done if e.emits])
            out = sch.process_output()
            for sid, tok in zip(out.act_schedule_ids, out.next_tokens):
                produced.setdefault(sid, []).append(tok)
            check_invariants(mm, everyone)
        if not pending and not sch.has_work():
            break
    assert not sch.has_work() and not pending, "engine did not drain"
    for s in reqs:
        assert not s.page_table                                              # everything was given back
        if not s.is_abort:
            assert len(produced.get(s.seq_id, [])) == s.output_len
    assert mm.get_num_free_pages() == pages
    check_invariants(mm, [])


def _tok(seq_id: int, pos: int) -> int:
    """Deterministic stand-in for the model: the token at position `pos` of a sequence depends on nothing else,
    so synchronous and lookahead scheduling must produce identical streams."""
    return 3 + (seq_id * 7919 + pos * 104729) % 45


def _check_incremental(entries, prev_arrays):
    """The worker builds every batch with `prev=` the previous one: the incremental result must equal a full build."""
    inc = build_batch(entries, PAGE, 64, 0, prev=prev_arrays)
    full = build_batch(entries, PAGE, 64, 0, prev=None)
    for name in ("tokens", "positions", "slot_mapping", "seq_lens", "query_start_loc", "logits_idx"):
        assert np.array_equal(getattr(inc, name), getattr(full, name)), name
    for r in range(len(entries)):
        nblk = (int(full.seq_lens[r]) + PAGE - 1) // PAGE
        assert np.array_equal(inc.block_table[r, :nblk], full.block_table[r, :nblk])
    return inc


def _drive(reqs_spec, pages, method, maxp, lookahead, arrivals, abort_plan):
    """Emulates the driver loop of engine/worker.py (run_driver) around the real Scheduler."""
    mm = PrefixMemoryManager(pages, PAGE)
    sch = Scheduler(mm, pp_size=1, world_size=1, schedule_method=method, maxd=6, maxp=maxp, minp=4, iterp=2,
                    kvthresh=0.0, page_size=PAGE, log=False)
    reqs = [Sequence(i, list(t), [2], output_len=o, ignore_eos=True) for i, (t, o) in enumerate(reqs_spec)]
    pending_reqs = list(reqs)
    inflight = []            # launched batches whose tokens are not back yet (<= 2 with lookahead)
    produced = {}
    n_look = 0
    prev_arrays = None
    for step in range(3000):
        if pending_reqs and arrivals[step % len(arrivals)]:
            sch.add_new_requests([pending_reqs.pop(0)])
        if step in abort_plan and abort_plan[step] < len(reqs):
            sch.add_abort_ids([abort_plan[step]])
        sch.check_abort_seqs()
        if lookahead and len(inflight) == 1:
            look = sch.schedule_lookahead()
            if look:
                n_look += 1
                inflight.append(look)
                prev_arrays = _check_incremental(look, prev_arrays)
        keep = 1 if (lookahead and len(inflight) == 2) else 0
        while len(inflight) > keep:
            done = inflight.pop(0)
            sch.add_next_tokens([_tok(e.seq.seq_id, e.start + e.n) for e in done if e.emits])
        while True:
            out = sch.process_output()
            if out is None:
                break
            for sid, tok in zip(out.act_schedule_ids, out.next_tokens):
                produced.setdefault(sid, []).append(tok)
        check_invariants(mm, reqs)
        entries = sch.schedule_once()
        if entries:
            inflight.append(entries)
            prev_arrays = _check_incremental(entries, prev_arrays)
        if not pending_reqs and not sch.has_work() and not inflight:
            break
    assert not pending_reqs and not sch.has_work() and not inflight, "engine did not drain"
    assert mm.get_num_free_pages() == pages
    for s in reqs:
        assert not s.page_table and s.pending < 0 and not s.zombie
    return reqs, produced, n_look


@settings(max_examples=int(__import__("os").environ.get("GLLM_HYP_EXAMPLES", "60")), deadline=None,
          derandomize=not __import__("os").environ.get("GLLM_HYP_RANDOM"), suppress_health_check=[HealthCheck.too_slow])
@given(seed=st.integers(0, 10 ** 6), n_req=st.integers(1, 10), pages=st.integers(12, 48),
       method=st.sampled_from(["chunked_prefill", "token_throttling"]), maxp=st.sampled_from([8, 16, 64]),
       aborts=st.booleans(), page=st.sampled_from([2, 4]))
def test_lookahead_scheduling_equals_synchronous(seed, n_req, pages, method, maxp, aborts, page):
    global PAGE
    PAGE = page
    """Async (lookahead) scheduling — placeholder tokens, zombies, pages reserved one step ahead — must give every
    request exactly the tokens the synchronous loop gives it and leave no page behind."""
    rng = random.Random(seed)
    spec = []
    for _ in range(n_req):
        toks = [rng.randrange(3, 50) for _ in range(rng.randrange(1, 18))]
        out = rng.randrange(1, 12)
        if (len(toks) + out + PAGE - 1) // PAGE <= pages - 2:
            spec.append((toks, out))
    arrivals = [rng.random() < 0.6 for _ in range(17)] + [True]
    abort_plan = {rng.randrange(0, 40): rng.randrange(0, max(len(spec), 1)) for _ in range(2)} if aborts else {}
    reqs_s, prod_s, _ = _drive(spec, pages, method, maxp, False, arrivals, abort_plan)
    reqs_a, prod_a, n_look = _drive(spec, pages, method, maxp, True, arrivals, abort_plan)
    for s, a in zip(reqs_s, reqs_a):
        want = [_tok(s.seq_id, s.prompt_len + i) for i in range(s.output_len)]
        if not s.is_abort:
            assert prod_s.get(s.seq_id, []) == want
        if not a.is_abort:
            assert prod_a.get(a.seq_id, []) == want, (a.seq_id, prod_a.get(a.seq_id), want)
        else:   # an aborted request may have streamed a prefix of its tokens, never anything else
            got = prod_a.get(a.seq_id, [])
            assert got == want[:len(got)]


@settings(max_examples=200, deadline=None, derandomize=True)
@given(ops=st.lists(st.tuples(st.sampled_from(["alloc", "alloc_id", "free"]), st.integers(0, 15)), max_size=120))
def test_id_allocator_matches_a_simple_model(ops):
    """The O(1) linked-list allocator behaves like the obvious model: allocate from the head of the free order,
    freed ids go to the tail, specific ids can be taken from anywhere."""
    from gllm_b200.id_allocator import IDAllocator
    a = IDAllocator(0, 15)
    order = list(range(16))            # model of the free list, head first
    for op, x in ops:
        if op == "alloc":
            if order:
                assert a.allocate() == order.pop(0)
            else:
                with pytest.raises(RuntimeError):
                    a.allocate()
        elif op == "alloc_id":
            assert a.allocate(x) == x
            if x in order:
                order.remove(x)
        else:
            if x in order:
                with pytest.raises(RuntimeError):
                    a.free(x)
            else:
                a.free(x)
                order.append(x)
        assert a.get_num_free_ids() == len(order)
        assert all(a.is_free(i) == (i in order) for i in range(16))


@settings(max_examples=300, deadline=None, derandomize=True)
@given(layers=st.integers(1, 130), pp=st.integers(1, 16))
def test_layer_partition_covers_every_layer_once(layers, pp):
    from gllm_b200.parallel.state import partition_layers
    if pp > layers:
        return
    parts = partition_layers(layers, pp)
    assert len(parts) == pp and all(len(r) > 0 for r in parts)
    flat = [i for r in parts for i in r]
    assert flat == list(range(layers))                       # contiguous, ordered, complete
    assert max(len(r) for r in parts) - min(len(r) for r in parts) <= max(1, (layers + pp - 1) // pp - 1)


@settings(max_examples=80, deadline=None, derandomize=True)
@given(seed=st.integers(0, 10 ** 6), b=st.integers(1, 6), v=st.integers(2, 300), k=st.integers(-1, 40),
       p=st.floats(0.05, 1.0), temp=st.floats(0.1, 2.0))
def test_reference_sampler_draws_inside_the_filtered_support(seed, b, v, k, p, temp):
    """ops/ref.py is the oracle of the sm_100a sampler kernel: a drawn token must survive the top-k and nucleus
    filters, greedy rows (top_k == 1) must be the argmax, the filtered distribution is normalised."""
    import torch
    from gllm_b200.ops import ref
    g = torch.Generator().manual_seed(seed)
    logits = torch.randn(b, v, generator=g) * 3
    tk = torch.full((b,), k, dtype=torch.int32)
    tk[0] = 1                                                   # one greedy row in every batch
    tp = torch.full((b,), p)
    t = torch.full((b,), temp)
    probs = ref.sample_filter(logits, t, tk, tp)
    assert torch.allclose(probs.sum(-1), torch.ones(b), atol=1e-5)
    kk = tk.clone().long()
    kk[(kk <= 0) | (kk > v)] = v
    assert ((probs > 0).sum(-1) <= kk).all() or bool((logits.sort(-1).values.diff(dim=-1) == 0).any())  # ties may widen top-k
    for s in range(3):
        tok = ref.sample(logits, t, tk, tp, generator=torch.Generator().manual_seed(seed + s)).long()
        assert (probs.gather(1, tok.view(-1, 1)) > 0).all()
        assert int(tok[0]) == int(logits[0].argmax())
    # nucleus: the kept set is the smallest prefix (by probability) whose mass reaches top_p
    base = ref.sample_filter(logits, t, tk, None)
    for r in range(b):
        kept = probs[r] > 0
        mass_kept = float(base[r][kept].sum())
        assert mass_kept >= min(p, 1.0) - 1e-4
        if kept.sum() > 1:
            smallest = float(base[r][kept].min())
            assert mass_kept - smallest < p + 1e-4              # dropping the weakest kept token would fall short


@settings(max_examples=100, deadline=None, derandomize=True)
@given(seed=st.integers(0, 10 ** 6), b=st.integers(1, 9), mrope=st.booleans(), penalty=st.booleans(),
       feed=st.booleans(), mm=st.booleans())
def test_batch_wire_round_trip(seed, b, mrope, penalty, feed, mm):
    """Every field a peer needs survives Comm's single-buffer encoding (optional arrays, 2-D positions, the
    multimodal payload dict), bit for bit, for arbitrary shapes."""
    from gllm_b200.engine.comm import Comm
    from gllm_b200.input_data import BatchArrays
    rng = np.random.default_rng(seed)
    q = rng.integers(1, 7, b).astype(np.int32)
    t = int(q.sum())
    i32 = lambda *shape: rng.integers(0, 1000, shape).astype(np.int32)    # noqa: E731
    f32 = lambda *shape: rng.random(shape).astype(np.float32)             # noqa: E731
    batch = BatchArrays(
        tokens=i32(t), positions=i32(3, t) if mrope else i32(t), slot_mapping=i32(t),
        block_table=i32(b, int(rng.integers(1, 9))), seq_lens=i32(b), query_start_loc=np.concatenate([[0], q.cumsum()]).astype(np.int32),
        logits_idx=i32(b), emit_seq=np.arange(b, dtype=np.int32), temperature=f32(b), top_k=i32(b), top_p=f32(b),
        rep_penalty=f32(b), state_slot=i32(b), num_decode_seqs=int(rng.integers(0, b + 1)), num_seqs=b, num_tokens=t,
        max_q_len=int(q.max()), max_seq_len=77, all_greedy=not penalty, need_penalty=penalty, batch_id=seed % 1000,
        seen_rows=i32(5) if penalty else None, seen_tokens=i32(5) if penalty else None,
        clear_slots=i32(2) if penalty else None, feed_src=i32(b) if feed else None,
        mm={"rows": [1, 2], "pixel_values": f32(6, 10), "grid": [[1, 2, 3]]} if mm else None)
    got = Comm._decode_batch(memoryview(Comm._encode_batch(batch)))
    for name in ("tokens", "positions", "slot_mapping", "block_table", "seq_lens", "query_start_loc", "logits_idx",
                 "emit_seq", "temperature", "top_k", "top_p", "rep_penalty", "state_slot", "seen_rows", "seen_tokens",
                 "clear_slots", "feed_src"):
        a, g = getattr(batch, name), getattr(got, name)
        assert (a is None and g is None) or (a.dtype == g.dtype and a.shape == g.shape and np.array_equal(a, g)), name
    for name in ("num_decode_seqs", "num_seqs", "num_tokens", "max_q_len", "max_seq_len", "all_greedy", "need_penalty",
                 "batch_id"):
        assert getattr(batch, name) == getattr(got, name), name
    if mm:
        assert got.mm["rows"] == [1, 2] and np.array_equal(got.mm["pixel_values"], batch.mm["pixel_values"])
    else:
        assert got.mm is None
