"""Greedy-decode token-exact match against HuggingFace transformers on tiny random models (CPU, fp32),
through the full engine: scheduler, chunked prefill, paged KV, prefix cache, loader."""
from conftest import scratch_dir
import os

import pytest
import torch

transformers = pytest.importorskip("transformers")


def _save(model):
    d = scratch_dir("gllm_b200_test_")
    model.save_pretrained(d, safe_serialization=True)
    return d


def _hf_greedy(model, prompt, n):
    with torch.no_grad():
        out = model.generate(torch.tensor([prompt]), max_new_tokens=n, do_sample=False, eos_token_id=None,
                             pad_token_id=0)
    return out[0, len(prompt):].tolist()


def _engine(path, **kw):
    from gllm_b200 import LLM
    args = dict(maxp=64, maxd=64, page_size=16, num_cpu_pages=96, model_max_length=320, log_stats=False)
    args.update(kw)
    return LLM(path, **args)


PROMPTS = [[5, 17, 99, 200, 3, 45, 7], [9] * 40, list(range(20, 150)), [300, 301]]


def _check(model, **kw):
    d = _save(model)
    llm = _engine(d, **kw)
    outs = llm.generate(tokens=PROMPTS, output_lens=[10] * len(PROMPTS), ignore_eos=True)
    for p, s in zip(PROMPTS, outs):
        assert s.token_ids[len(p):] == _hf_greedy(model, p, 10), (len(p),)
    llm.shutdown()
    return d


def test_llama_matches_hf():
    from transformers import LlamaConfig, LlamaForCausalLM
    torch.manual_seed(0)
    cfg = LlamaConfig(hidden_size=128, intermediate_size=256, num_hidden_layers=2, num_attention_heads=4,
                      num_key_value_heads=2, vocab_size=512, max_position_embeddings=512, eos_token_id=1,
                      bos_token_id=0)
    _check(LlamaForCausalLM(cfg).eval().float())


def test_llama3_rope_scaling_matches_hf():
    from transformers import LlamaConfig, LlamaForCausalLM
    torch.manual_seed(1)
    cfg = LlamaConfig(hidden_size=128, intermediate_size=192, num_hidden_layers=2, num_attention_heads=4,
                      num_key_value_heads=4, vocab_size=512, max_position_embeddings=512, eos_token_id=1,
                      rope_parameters={"rope_type": "llama3", "factor": 8.0, "low_freq_factor": 1.0,
                                       "high_freq_factor": 4.0, "original_max_position_embeddings": 64,
                                       "rope_theta": 10000.0})
    _check(LlamaForCausalLM(cfg).eval().float())


def test_qwen2_bias_tied_matches_hf():
    from transformers import Qwen2Config, Qwen2ForCausalLM
    torch.manual_seed(2)
    cfg = Qwen2Config(hidden_size=128, intermediate_size=256, num_hidden_layers=3, num_attention_heads=4,
                      num_key_value_heads=2, vocab_size=512, max_position_embeddings=512, tie_word_embeddings=True,
                      eos_token_id=1)
    m = Qwen2ForCausalLM(cfg).eval().float()
    for n, p in m.named_parameters():
        if n.endswith("bias"):
            torch.nn.init.normal_(p, std=0.5)
    _check(m)


def test_qwen3_qk_norm_matches_hf_with_prefix_cache_and_tiny_chunks():
    from transformers import Qwen3Config, Qwen3ForCausalLM
    torch.manual_seed(3)
    cfg = Qwen3Config(hidden_size=128, intermediate_size=256, num_hidden_layers=2, num_attention_heads=4,
                      num_key_value_heads=2, head_dim=32, vocab_size=512, max_position_embeddings=512,
                      eos_token_id=1, tie_word_embeddings=False)
    m = Qwen3ForCausalLM(cfg).eval().float()
    for n, p in m.named_parameters():
        if "q_norm" in n or "k_norm" in n:
            torch.nn.init.normal_(p, mean=1.0, std=0.2)
    d = _check(m, maxp=24, enable_prefix_caching=True)
    # second engine: same prompts twice -> second round hits the prefix cache and must not change tokens
    llm = _engine(d, maxp=48, enable_prefix_caching=True)
    r1 = llm.generate(tokens=PROMPTS, output_lens=[6] * 4, ignore_eos=True)
    r2 = llm.generate(tokens=PROMPTS, output_lens=[6] * 4, ignore_eos=True)
    assert [s.token_ids for s in r1] == [s.token_ids for s in r2]
    assert llm.worker.mm.get_cache_hit_rate() > 0
    assert any(s.num_cached_tokens > 0 for s in r2)
    llm.shutdown()


def test_token_throttling_and_preemption_keep_tokens_exact():
    from transformers import LlamaConfig, LlamaForCausalLM
    torch.manual_seed(4)
    cfg = LlamaConfig(hidden_size=64, intermediate_size=128, num_hidden_layers=2, num_attention_heads=2,
                      num_key_value_heads=2, vocab_size=256, max_position_embeddings=512, eos_token_id=1)
    m = LlamaForCausalLM(cfg).eval().float()
    d = _save(m)
    prompts = [[3 + i, 9, 27, 81, 5] * 4 for i in range(6)]
    # KV for ~3 sequences only -> decode preemption + recompute must still give HF's tokens
    llm = _engine(d, schedule_method="token_throttling", num_cpu_pages=10, kvthresh=0.0, maxp=32, maxd=8,
                  enable_prefix_caching=False)
    outs = llm.generate(tokens=prompts, output_lens=[24] * 6, ignore_eos=True)
    assert llm.worker.scheduler.num_preempt_seqs > 0
    for p, s in zip(prompts, outs):
        assert s.token_ids[len(p):] == _hf_greedy(m, p, 24)
    llm.shutdown()


def test_eos_and_output_len_finish():
    from transformers import LlamaConfig, LlamaForCausalLM
    torch.manual_seed(5)
    cfg = LlamaConfig(hidden_size=64, intermediate_size=128, num_hidden_layers=1, num_attention_heads=2,
                      num_key_value_heads=2, vocab_size=64, max_position_embeddings=256, eos_token_id=1)
    m = LlamaForCausalLM(cfg).eval().float()
    d = _save(m)
    llm = _engine(d)
    p = [4, 8, 15, 16, 23, 42]
    ref = _hf_greedy(m, p, 40)
    llm.finish_tokens = [ref[5]]  # pretend the 6th generated token is EOS
    s = llm.generate(tokens=[p], output_lens=[40])[0]
    first = ref.index(ref[5])
    assert s.token_ids[len(p):] == ref[: first + 1]
    with pytest.raises(ValueError):
        llm.generate(tokens=[[1] * 400], output_lens=[4])
    llm.shutdown()


def test_mixtral_matches_hf():
    from transformers import MixtralConfig, MixtralForCausalLM
    torch.manual_seed(6)
    cfg = MixtralConfig(hidden_size=64, intermediate_size=128, num_hidden_layers=2, num_attention_heads=4,
                        num_key_value_heads=2, vocab_size=512, max_position_embeddings=512, num_local_experts=4,
                        num_experts_per_tok=2, eos_token_id=1, sliding_window=None)
    _check(MixtralForCausalLM(cfg).eval().float())


def test_qwen3_moe_matches_hf():
    from transformers import Qwen3MoeConfig, Qwen3MoeForCausalLM
    torch.manual_seed(7)
    cfg = Qwen3MoeConfig(hidden_size=64, intermediate_size=128, moe_intermediate_size=64, num_hidden_layers=2,
                         num_attention_heads=4, num_key_value_heads=2, head_dim=16, vocab_size=512,
                         max_position_embeddings=512, num_experts=8, num_experts_per_tok=2, norm_topk_prob=True,
                         decoder_sparse_step=1, mlp_only_layers=[], eos_token_id=1)
    _check(Qwen3MoeForCausalLM(cfg).eval().float())


def test_qwen2_moe_shared_expert_matches_hf():
    from transformers import Qwen2MoeConfig, Qwen2MoeForCausalLM
    torch.manual_seed(8)
    cfg = Qwen2MoeConfig(hidden_size=64, intermediate_size=128, moe_intermediate_size=64,
                         shared_expert_intermediate_size=128, num_hidden_layers=2, num_attention_heads=4,
                         num_key_value_heads=2, vocab_size=512, max_position_embeddings=512, num_experts=4,
                         num_experts_per_tok=2, norm_topk_prob=False, decoder_sparse_step=1, mlp_only_layers=[],
                         eos_token_id=1)
    _check(Qwen2MoeForCausalLM(cfg).eval().float())


def test_deepseek_v3_mla_moe_matches_hf():
    from transformers import DeepseekV3Config, DeepseekV3ForCausalLM
    torch.manual_seed(9)
    cfg = DeepseekV3Config(hidden_size=64, intermediate_size=128, moe_intermediate_size=32, num_hidden_layers=3,
                           num_attention_heads=4, num_key_value_heads=4, n_routed_experts=8, n_shared_experts=1,
                           num_experts_per_tok=2, n_group=2, topk_group=1, first_k_dense_replace=1,
                           routed_scaling_factor=2.5, norm_topk_prob=True, q_lora_rank=24, kv_lora_rank=32,
                           qk_nope_head_dim=16, qk_rope_head_dim=8, v_head_dim=16, vocab_size=512,
                           max_position_embeddings=512, eos_token_id=1, rope_scaling=None, rope_interleave=True)
    m = DeepseekV3ForCausalLM(cfg).eval().float()
    for n, p in m.named_parameters():
        if "e_score_correction_bias" in n:
            torch.nn.init.normal_(p, std=0.05)
    for n, b in m.named_buffers():
        if "e_score_correction_bias" in n:
            b.normal_(std=0.05)
    _check(m)


def test_fp8_block_quantised_model_close_to_hf():
    """`quantization_config: fp8 / weight_block_size [128,128]` -> linears run the block-scaled fp8 path
    (activations quantised per token per 128 group). Quantisation noise may flip late tokens; the first
    tokens must agree with the unquantised HF model."""
    import json
    from transformers import Qwen3Config, Qwen3ForCausalLM
    torch.manual_seed(11)
    cfg = Qwen3Config(hidden_size=256, intermediate_size=512, num_hidden_layers=2, num_attention_heads=4,
                      num_key_value_heads=2, head_dim=64, vocab_size=512, max_position_embeddings=512,
                      eos_token_id=1, tie_word_embeddings=False)
    m = Qwen3ForCausalLM(cfg).eval().float()
    d = _save(m)
    c = json.load(open(os.path.join(d, "config.json")))
    c["quantization_config"] = {"quant_method": "fp8", "activation_scheme": "dynamic", "fmt": "e4m3",
                                "weight_block_size": [128, 128]}
    json.dump(c, open(os.path.join(d, "config.json"), "w"))
    llm = _engine(d)
    model = llm.worker.runner.model
    assert model.spec.quant == "fp8" and model.layers[0].attn.qkv_w.dtype == torch.float8_e4m3fn
    outs = llm.generate(tokens=PROMPTS, output_lens=[4] * len(PROMPTS), ignore_eos=True)
    same_first = sum(s.token_ids[len(p)] == _hf_greedy(m, p, 1)[0] for p, s in zip(PROMPTS, outs))
    assert same_first >= len(PROMPTS) - 1
    llm.shutdown()


def test_fp8_block_quantised_moe_experts_close_to_hf():
    """fp8 checkpoints also keep the MoE EXPERT weights block-quantised (e4m3 + a scale row per 64 weight rows,
    reference Fp8MoEMethod): first tokens must agree with the unquantised HF model."""
    import json
    from transformers import Qwen3MoeConfig, Qwen3MoeForCausalLM
    torch.manual_seed(12)
    cfg = Qwen3MoeConfig(hidden_size=256, intermediate_size=512, moe_intermediate_size=128, num_hidden_layers=2,
                         num_attention_heads=4, num_key_value_heads=2, head_dim=64, vocab_size=512,
                         max_position_embeddings=512, eos_token_id=1, tie_word_embeddings=False, num_experts=4,
                         num_experts_per_tok=2, decoder_sparse_step=1, mlp_only_layers=[], norm_topk_prob=True)
    m = Qwen3MoeForCausalLM(cfg).eval().float()
    d = _save(m)
    c = json.load(open(os.path.join(d, "config.json")))
    c["quantization_config"] = {"quant_method": "fp8", "activation_scheme": "dynamic", "fmt": "e4m3",
                                "weight_block_size": [128, 128]}
    json.dump(c, open(os.path.join(d, "config.json"), "w"))
    llm = _engine(d)
    ex = llm.worker.runner.model.layers[0].mlp.experts
    assert ex.quant == "fp8" and ex.w13.dtype == torch.float8_e4m3fn and ex.w13_ws.shape == (4, 4, 2)
    outs = llm.generate(tokens=PROMPTS, output_lens=[4] * len(PROMPTS), ignore_eos=True)
    same_first = sum(s.token_ids[len(p)] == _hf_greedy(m, p, 1)[0] for p, s in zip(PROMPTS, outs))
    assert same_first >= len(PROMPTS) - 1
    llm.shutdown()


def test_async_lookahead_scheduling_is_token_exact():
    """`async_schedule=True`: the next decode step is queued before the previous step's tokens reach the host
    (placeholder + device-side token feed). Must reproduce HF greedy tokens, and the synchronous engine's tokens
    with EOS finishing mid-flight (zombie sequences) and under KV pressure (preemption)."""
    from transformers import Qwen3Config, Qwen3ForCausalLM
    from gllm_b200 import LLM
    from gllm_b200 import scheduler as S
    torch.manual_seed(5)
    cfg = Qwen3Config(hidden_size=64, intermediate_size=128, num_hidden_layers=2, num_attention_heads=4,
                      num_key_value_heads=2, head_dim=16, vocab_size=512, max_position_embeddings=512,
                      eos_token_id=1, tie_word_embeddings=False)
    m = Qwen3ForCausalLM(cfg).eval().float()
    d = _save(m)
    looks = {"n": 0}
    orig = S.Scheduler.schedule_lookahead

    def counted(self):
        r = orig(self)
        looks["n"] += bool(r)
        return r

    S.Scheduler.schedule_lookahead = counted
    try:
        for prefix in (False, True):      # with the prefix cache, pages completed by decode are hashed one step later
            looks["n"] = 0
            llm = _engine(d, async_schedule=True, enable_prefix_caching=prefix)
            outs = llm.generate(tokens=PROMPTS, output_lens=[10] * len(PROMPTS), ignore_eos=True)
            for p, s in zip(PROMPTS, outs):
                assert s.token_ids[len(p):] == _hf_greedy(m, p, 10), (len(p),)
            assert looks["n"] >= 5, "lookahead scheduling never engaged"
            llm.shutdown()
        # EOS finishing while a lookahead step is in flight + KV pressure: async == sync
        gen = [t for p, s in zip(PROMPTS, outs) for t in s.token_ids[len(p):]]
        eos = max(set(gen), key=gen.count)
        res = {}
        for mode in (False, True):
            llm = _engine(d, async_schedule=mode, enable_prefix_caching=False, num_cpu_pages=40, kvthresh=0.0)
            llm.finish_tokens = [eos]
            o = llm.generate(tokens=PROMPTS, output_lens=[40] * len(PROMPTS), ignore_eos=False)
            res[mode] = [s.token_ids for s in o]
            mm = llm.worker.mm
            assert mm.get_num_free_pages() == mm.num_pages - 1, "pages leaked"    # all but the dummy page
            llm.shutdown()
        assert res[False] == res[True]
        assert any(len(x) - len(p) < 40 for x, p in zip(res[True], PROMPTS)), "EOS never hit: test is vacuous"
        # requests arriving while others decode (mixed lookahead batches: decode rows fed from the device + new
        # prefill chunks with known tokens)
        staged = {}
        for mode in (False, True):
            llm = _engine(d, async_schedule=mode, maxp=32)
            seqs = [llm.allocate_seq(p, n, True, top_k=1) for p, n in zip(PROMPTS, (30, 12, 20, 25))]
            llm.add_requests(seqs[:1])
            n = 0
            while len(llm.finished) < len(seqs):
                llm.schedule()
                n += 1
                if n in (6, 11, 15):
                    llm.add_requests([seqs[(6, 11, 15).index(n) + 1]])
                assert n < 5000
            staged[mode] = [s.token_ids for s in seqs]
            llm.shutdown()
        assert staged[False] == staged[True]
    finally:
        S.Scheduler.schedule_lookahead = orig


def test_deepseek_fp8_checkpoint_keeps_experts_quantised():
    """DeepSeek-V3-style fp8 checkpoints: the routed experts stay block-scaled e4m3 (what lets V3 fit on 8 GPUs),
    MLA / dense / shared weights are de-quantised; first tokens agree with the unquantised HF model."""
    import json
    from transformers import DeepseekV3Config, DeepseekV3ForCausalLM
    torch.manual_seed(13)
    cfg = DeepseekV3Config(hidden_size=256, intermediate_size=256, moe_intermediate_size=128, num_hidden_layers=2,
                           num_attention_heads=4, num_key_value_heads=4, n_routed_experts=4, n_shared_experts=1,
                           num_experts_per_tok=2, n_group=2, topk_group=1, first_k_dense_replace=1,
                           routed_scaling_factor=1.0, norm_topk_prob=True, q_lora_rank=32, kv_lora_rank=32,
                           qk_nope_head_dim=16, qk_rope_head_dim=8, v_head_dim=16, vocab_size=512,
                           max_position_embeddings=512, eos_token_id=1, rope_scaling=None, rope_interleave=True)
    m = DeepseekV3ForCausalLM(cfg).eval().float()
    d = _save(m)
    c = json.load(open(os.path.join(d, "config.json")))
    c["quantization_config"] = {"quant_method": "fp8", "activation_scheme": "dynamic", "fmt": "e4m3",
                                "weight_block_size": [128, 128]}
    json.dump(c, open(os.path.join(d, "config.json"), "w"))
    llm = _engine(d)
    ex = llm.worker.runner.model.layers[1].mlp.experts
    assert ex.quant == "fp8" and ex.w13.dtype == torch.float8_e4m3fn
    outs = llm.generate(tokens=PROMPTS, output_lens=[4] * len(PROMPTS), ignore_eos=True)
    same_first = sum(s.token_ids[len(p)] == _hf_greedy(m, p, 1)[0] for p, s in zip(PROMPTS, outs))
    assert same_first >= len(PROMPTS) - 1
    llm.shutdown()


@pytest.mark.parametrize("async_schedule", [False, True])
def test_random_arrivals_and_aborts_through_the_engine(async_schedule):
    """Front-end + worker + scheduler + runner under churn: requests arrive at random ticks, some are aborted at
    random ticks (possibly before they were ever scheduled, mid-prefill, mid-decode). The engine must drain, give
    every page back, free every sequence id, and the surviving requests must get exactly the tokens of a quiet run."""
    import random
    from gllm_b200 import LLM
    from gllm_b200.models.presets import tiny
    cfg = tiny("Qwen3ForCausalLM", num_hidden_layers=1, max_position_embeddings=512)

    def engine():
        return LLM(cfg, load_format="dummy", maxp=24, maxd=6, num_cpu_pages=pool, page_size=4, log_stats=False,
                   device="cpu", kvthresh=0.0, async_schedule=async_schedule, enable_prefix_caching=True, seed=0)

    for seed in range(int(os.environ.get("GLLM_CHURN_SEEDS", "4"))):
        rng = random.Random(seed)
        pool = (96, 40, 28)[seed % 3]        # roomy / tight / very tight KV pool (preemption + recompute)
        stems = [[rng.randrange(5, 300) for _ in range(rng.randrange(4, 24))] for _ in range(3)]   # shared prefixes
        prompts = []
        for _ in range(14):
            st_ = rng.choice(stems)
            prompts.append(st_[:rng.randrange(1, len(st_) + 1)] + [rng.randrange(5, 300) for _ in range(rng.randrange(0, 8))])
        outs = [rng.randrange(1, 14) for _ in prompts]
        llm = engine()
        quiet = [s.token_ids for s in llm.generate(tokens=prompts, output_lens=outs, ignore_eos=True)]
        llm.shutdown()

        llm = engine()
        seqs = [llm.allocate_seq(p, o, True, top_k=1) for p, o in zip(prompts, outs)]
        arrive = sorted((rng.randrange(0, 25), i) for i in range(len(seqs)))
        aborts = {i: rng.randrange(0, 40) for i in rng.sample(range(len(seqs)), 5)}
        aborted = set()
        tick = 0
        while len(llm.finished) < len(seqs):
            while arrive and arrive[0][0] <= tick:
                llm.add_requests([seqs[arrive.pop(0)[1]]])
            for i, t in aborts.items():
                if t == tick and seqs[i].seq_id in llm.running_maps:
                    llm.abort([seqs[i].seq_id])
                    aborted.add(i)
            llm.schedule()
            tick += 1
            if tick > 40 and not arrive and not llm.running_maps and not llm.wait_lists:
                break
            assert tick < 20000, "engine did not drain"
        for i, s in enumerate(seqs):
            if i in aborted:
                assert s.token_ids == quiet[i][:len(s.token_ids)]      # a prefix of the quiet stream at most
            else:
                assert s.token_ids == quiet[i], (seed, i)
        mm = llm.worker.mm
        assert mm.get_num_free_pages() == mm.num_pages - 1, "pages leaked"          # all but the dummy page
        assert not llm.running_maps
        assert llm.id_allocator.get_num_free_ids() == 100000 - 0                  # every sequence id returned
        llm.shutdown()


def test_requests_submitted_from_another_thread_are_never_lost():
    """The API server adds requests from the event-loop thread while the engine tick runs in a worker thread: a
    request arriving in the middle of `_send` must neither vanish nor lose its tokens."""
    import threading
    import random
    from gllm_b200 import LLM
    from gllm_b200.models.presets import tiny
    cfg = tiny("Qwen3ForCausalLM", num_hidden_layers=1, max_position_embeddings=256)
    llm = LLM(cfg, load_format="dummy", maxp=64, maxd=64, num_cpu_pages=512, page_size=4, log_stats=False, device="cpu")
    n = 400
    seqs = []
    rng = random.Random(0)

    def feeder():
        for i in range(n):
            s = llm.allocate_seq([5 + i % 50, 7, 9], 1 + i % 3, True, top_k=1)
            seqs.append(s)
            llm.add_requests([s])
            if i % 7 == 0:
                time.sleep(rng.random() * 0.0005)

    import time
    th = threading.Thread(target=feeder)
    th.start()
    t0 = time.time()
    while len(llm.finished) < n:
        llm.schedule()
        assert time.time() - t0 < 120, f"only {len(llm.finished)} of {n} requests finished: some were lost"
    th.join()
    assert all(s.num_output_tokens == s.output_len for s in seqs)
    assert not llm.running_maps and llm.id_allocator.get_num_free_ids() == 100000
    llm.shutdown()


def test_random_dense_configs_match_hf():
    """Loader / spec robustness: randomly drawn dense configurations (GQA ratios incl. MQA, head_dim != hidden /
    heads, tied embeddings, attention biases, odd intermediate sizes, page sizes, chunk sizes) stay token-exact
    against HuggingFace."""
    import random
    from transformers import LlamaConfig, LlamaForCausalLM, Qwen2Config, Qwen2ForCausalLM, Qwen3Config, Qwen3ForCausalLM
    n = int(os.environ.get("GLLM_CFG_SEEDS", "5"))
    for seed in range(n):
        rng = random.Random(seed)
        torch.manual_seed(100 + seed)
        heads = rng.choice([2, 4, 8])
        kv = rng.choice([h for h in (1, 2, 4, 8) if h <= heads and heads % h == 0])
        hd = rng.choice([16, 32])
        fam = rng.choice(["llama", "qwen2", "qwen3"])
        common = dict(hidden_size=rng.choice([64, 96, 128]), intermediate_size=rng.choice([80, 128, 176]),
                      num_hidden_layers=rng.choice([1, 2, 3]), num_attention_heads=heads, num_key_value_heads=kv,
                      vocab_size=rng.choice([320, 512, 777]), max_position_embeddings=512, eos_token_id=1,
                      tie_word_embeddings=rng.random() < 0.5, rms_norm_eps=rng.choice([1e-5, 1e-6]))
        if fam == "llama":
            cfg = LlamaConfig(**common, head_dim=hd, attention_bias=rng.random() < 0.3)
            model = LlamaForCausalLM(cfg)
        elif fam == "qwen2":
            common["hidden_size"] = heads * hd        # Qwen2 has no explicit head_dim
            cfg = Qwen2Config(**common)
            model = Qwen2ForCausalLM(cfg)
        else:
            cfg = Qwen3Config(**common, head_dim=hd)
            model = Qwen3ForCausalLM(cfg)
        model = model.eval().float()
        try:
            _check(model, page_size=rng.choice([4, 8, 16]), maxp=rng.choice([16, 48, 64]),
                   enable_prefix_caching=rng.random() < 0.5)
        except AssertionError as e:
            raise AssertionError(f"seed {seed}: {fam} {cfg.to_dict()}") from e


def test_random_moe_configs_match_hf():
    """Same idea for the MoE families: expert counts that are not multiples of 8 (router padding), top-k, shared
    experts, renormalisation, sparse-step / mlp-only layer patterns."""
    import random
    from transformers import (MixtralConfig, MixtralForCausalLM, Qwen2MoeConfig, Qwen2MoeForCausalLM, Qwen3MoeConfig,
                              Qwen3MoeForCausalLM)
    n = int(os.environ.get("GLLM_CFG_SEEDS", "4"))
    for seed in range(n):
        rng = random.Random(1000 + seed)
        torch.manual_seed(200 + seed)
        heads = rng.choice([2, 4])
        kv = rng.choice([h for h in (1, 2, 4) if h <= heads and heads % h == 0])
        experts = rng.choice([3, 4, 6, 8])
        topk = rng.randrange(1, min(experts, 3) + 1)
        layers = rng.choice([2, 3])
        fam = rng.choice(["mixtral", "qwen2_moe", "qwen3_moe"])
        base = dict(hidden_size=64, num_hidden_layers=layers, num_attention_heads=heads, num_key_value_heads=kv,
                    vocab_size=512, max_position_embeddings=512, eos_token_id=1, tie_word_embeddings=False)
        if fam == "mixtral":
            cfg = MixtralConfig(**base, intermediate_size=rng.choice([64, 96]), num_local_experts=experts,
                                num_experts_per_tok=topk)
            model = MixtralForCausalLM(cfg)
        elif fam == "qwen2_moe":
            cfg = Qwen2MoeConfig(**base, intermediate_size=96, moe_intermediate_size=rng.choice([32, 48]),
                                 shared_expert_intermediate_size=rng.choice([32, 64]), num_experts=experts,
                                 num_experts_per_tok=topk, norm_topk_prob=rng.random() < 0.5,
                                 decoder_sparse_step=rng.choice([1, 2]),
                                 mlp_only_layers=[0] if rng.random() < 0.3 else [])
            model = Qwen2MoeForCausalLM(cfg)
        else:
            cfg = Qwen3MoeConfig(**base, head_dim=16, intermediate_size=96, moe_intermediate_size=rng.choice([32, 48]),
                                 num_experts=experts, num_experts_per_tok=topk, norm_topk_prob=rng.random() < 0.5,
                                 decoder_sparse_step=rng.choice([1, 2]), mlp_only_layers=[])
            model = Qwen3MoeForCausalLM(cfg)
        model = model.eval().float()
        try:
            _check(model, page_size=rng.choice([8, 16]), maxp=rng.choice([32, 64]))
        except AssertionError as e:
            raise AssertionError(f"seed {seed}: {fam} {cfg.to_dict()}") from e


def test_random_deepseek_v3_configs_match_hf():
    """MLA + grouped routing under randomly drawn shapes: with / without the q LoRA, different nope / rope / v head
    sizes, group counts, dense-prefix depth, shared-expert counts, routing scale."""
    import random
    from transformers import DeepseekV3Config, DeepseekV3ForCausalLM
    n = int(os.environ.get("GLLM_CFG_SEEDS", "3"))
    for seed in range(n):
        rng = random.Random(2000 + seed)
        torch.manual_seed(300 + seed)
        groups = rng.choice([1, 2, 4])
        experts = groups * rng.choice([2, 3])
        topk_group = rng.randrange(1, groups + 1)
        topk = rng.randrange(1, min(4, topk_group * (experts // groups)) + 1)
        cfg = DeepseekV3Config(
            hidden_size=64, intermediate_size=96, moe_intermediate_size=rng.choice([32, 48]),
            num_hidden_layers=rng.choice([2, 3]), num_attention_heads=rng.choice([2, 4]), num_key_value_heads=4,
            n_routed_experts=experts, n_shared_experts=rng.choice([1, 2]), num_experts_per_tok=topk, n_group=groups,
            topk_group=topk_group, first_k_dense_replace=rng.choice([0, 1]), routed_scaling_factor=rng.choice([1.0, 2.5]),
            norm_topk_prob=rng.random() < 0.7, q_lora_rank=rng.choice([None, 24]), kv_lora_rank=rng.choice([16, 32]),
            qk_nope_head_dim=rng.choice([16, 32]), qk_rope_head_dim=rng.choice([8, 16]), v_head_dim=rng.choice([16, 32]),
            vocab_size=512, max_position_embeddings=512, eos_token_id=1, rope_scaling=None, rope_interleave=True)
        cfg.num_key_value_heads = cfg.num_attention_heads
        m = DeepseekV3ForCausalLM(cfg).eval().float()
        for name, b in list(m.named_buffers()) + list(m.named_parameters()):
            if "e_score_correction_bias" in name:
                b.data.normal_(std=0.05)
        try:
            _check(m, page_size=rng.choice([8, 16]))
        except AssertionError as e:
            raise AssertionError(f"seed {seed}: {cfg.to_dict()}") from e


def test_all_schedule_methods_generate_the_same_tokens():
    """Greedy decoding does not depend on the batching policy. (token_throttling used to overflow the runner's
    per-sequence buffers — capacity maxd — as soon as a full decode batch was joined by a prompt.)"""
    import random
    from gllm_b200 import LLM
    from gllm_b200.models.presets import tiny
    cfg = tiny("Qwen3ForCausalLM", num_hidden_layers=1)
    rng = random.Random(3)
    prompts = [[rng.randrange(5, 300) for _ in range(rng.randrange(1, 50))] for _ in range(16)]
    outs = [rng.randrange(1, 20) for _ in prompts]
    res = {}
    for m in ("chunked_prefill", "split_pd", "token_throttling"):
        llm = LLM(cfg, load_format="dummy", maxp=24, maxd=6, num_cpu_pages=64, page_size=4, log_stats=False,
                  device="cpu", schedule_method=m, seed=0)
        res[m] = [s.token_ids for s in llm.generate(tokens=prompts, output_lens=outs, ignore_eos=True)]
        llm.shutdown()
    assert res["split_pd"] == res["chunked_prefill"] and res["token_throttling"] == res["chunked_prefill"]
