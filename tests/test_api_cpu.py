"""OpenAI-API contract tests on CPU with fastapi.testclient: routes, SSE framing, usage, errors, metrics."""
from conftest import scratch_dir
import json
import os

import pytest
import torch

pytest.importorskip("fastapi")
transformers = pytest.importorskip("transformers")


def _make_model_dir():
    from tokenizers import Tokenizer, models, pre_tokenizers
    from transformers import LlamaConfig, LlamaForCausalLM, PreTrainedTokenizerFast
    torch.manual_seed(0)
    words = ["<unk>", "<s>", "</s>", "<|user|>", "<|assistant|>"] + [f"w{i}" for i in range(200)] + \
        ["hello", "world", "how", "are", "you", "?", "the", "a"]
    vocab = {w: i for i, w in enumerate(words)}
    tok = Tokenizer(models.WordLevel(vocab, unk_token="<unk>"))
    tok.pre_tokenizer = pre_tokenizers.Whitespace()
    fast = PreTrainedTokenizerFast(tokenizer_object=tok, unk_token="<unk>", bos_token="<s>", eos_token="</s>")
    fast.chat_template = "{% for m in messages %}<|{{ m['role'] }}|> {{ m['content'] }} {% endfor %}" \
                         "{% if add_generation_prompt %}<|assistant|> {% endif %}"
    cfg = LlamaConfig(hidden_size=64, intermediate_size=128, num_hidden_layers=2, num_attention_heads=2,
                      num_key_value_heads=2, vocab_size=len(words), max_position_embeddings=256, eos_token_id=2,
                      bos_token_id=1)
    d = scratch_dir("gllm_b200_api_")
    LlamaForCausalLM(cfg).eval().float().save_pretrained(d, safe_serialization=True)
    fast.save_pretrained(d)
    return d


@pytest.fixture(scope="module")
def client():
    from fastapi.testclient import TestClient
    from gllm_b200.engine.async_llm_engine import AsyncLLM
    from gllm_b200.entrypoints.api_server import build_app
    d = _make_model_dir()
    engine = AsyncLLM(d, maxp=64, maxd=16, num_cpu_pages=64, model_max_length=128, log_stats=False)
    assert engine.tokenizer is not None
    with TestClient(build_app(engine)) as c:
        yield c, engine
    engine.shutdown()


def test_models_health_metrics(client):
    c, engine = client
    r = c.get("/v1/models")
    assert r.status_code == 200 and r.json()["object"] == "list" and r.json()["data"][0]["id"]
    assert c.get("/health").json() == {"status": "ok"}
    assert "gllm_requests_total" in c.get("/metrics").text


def test_completion_non_stream_and_usage(client):
    c, engine = client
    r = c.post("/v1/completions", json={"model": "m", "prompt": "hello world how are you", "max_tokens": 6,
                                         "temperature": 0, "top_k": 1, "ignore_eos": True})
    assert r.status_code == 200, r.text
    body = r.json()
    assert body["object"] == "text_completion" and len(body["choices"]) == 1
    assert body["usage"]["prompt_tokens"] == 5 and body["usage"]["completion_tokens"] == 6
    assert body["usage"]["total_tokens"] == 11 and body["choices"][0]["finish_reason"] == "length"
    # token-id prompts are accepted too
    r2 = c.post("/v1/completions", json={"prompt": [5, 6, 7], "max_tokens": 3, "ignore_eos": True})
    assert r2.status_code == 200 and r2.json()["usage"]["completion_tokens"] == 3


def test_completion_stream_sse_framing(client):
    c, engine = client
    with c.stream("POST", "/v1/completions", json={"prompt": "hello world", "max_tokens": 5, "stream": True,
                                                    "ignore_eos": True, "top_k": 1}) as r:
        assert r.status_code == 200 and r.headers["content-type"].startswith("text/event-stream")
        raw = "".join(r.iter_text())
    events = [e for e in raw.split("\n\n") if e]
    assert all(e.startswith("data: ") for e in events) and events[-1] == "data: [DONE]"
    chunks = [json.loads(e[6:]) for e in events[:-1]]
    assert chunks[-1]["usage"]["completion_tokens"] == 5 and chunks[-1]["choices"][0]["finish_reason"] == "length"
    assert len({ch["id"] for ch in chunks}) == 1


def test_chat_completion_stream_and_non_stream(client):
    c, engine = client
    msgs = [{"role": "user", "content": "hello how are you ?"}]
    r = c.post("/v1/chat/completions", json={"model": "m", "messages": msgs, "max_completion_tokens": 4,
                                              "ignore_eos": True, "top_k": 1})
    assert r.status_code == 200, r.text
    body = r.json()
    assert body["object"] == "chat.completion" and body["choices"][0]["message"]["role"] == "assistant"
    assert body["usage"]["completion_tokens"] == 4
    with c.stream("POST", "/v1/chat/completions", json={"messages": msgs, "max_tokens": 4, "stream": True,
                                                         "ignore_eos": True}) as r:
        raw = "".join(r.iter_text())
    events = [e for e in raw.split("\n\n") if e]
    assert events[-1] == "data: [DONE]"
    first = json.loads(events[0][6:])
    assert first["object"] == "chat.completion.chunk" and first["choices"][0]["delta"].get("role") == "assistant"


def test_over_length_is_400(client):
    c, engine = client
    r = c.post("/v1/completions", json={"prompt": "hello " * 300, "max_tokens": 4})
    assert r.status_code == 400 and r.json()["type"] == "BadRequestError"
    r = c.post("/v1/completions", json={"prompt": "hello", "max_tokens": 100000})
    assert r.status_code == 400


def test_tokenize_roundtrip_and_profile_endpoints(client):
    c, engine = client
    t = c.post("/tokenize", json={"prompt": "hello world"}).json()
    assert t["count"] == 2
    assert "hello" in c.post("/detokenize", json={"tokens": t["tokens"]}).json()["prompt"]


def test_engine_metrics_after_requests(client):
    c, engine = client
    assert engine.metrics["requests_finished"] >= 4
    assert engine.metrics["generation_tokens_total"] > 0
    assert len(engine.running_maps) == 0
    text = c.get("/metrics").text
    for name in ("time_to_first_token", "time_per_output_token", "e2e_request_latency"):
        assert f'gllm_{name}_seconds_bucket{{le="+Inf"}}' in text
    assert 'gllm_step_seconds_total{phase="decode"}' in text and engine.last_stats["decode_step_count"] > 0
    assert engine.last_stats["prefill_step_tokens"] > 0 and engine.last_stats["decode_step_seconds"] > 0
    h = engine.hist["ttft"]
    assert h.count == sum(h.counts) >= 4 and engine.hist["tpot"].count > 0


def test_stop_strings(client):
    """OpenAI `stop`: generation ends at the first stop string, which is not part of the returned text; text that
    might still become a stop string is held back while streaming (the reference ignores `stop`)."""
    from gllm_b200.engine.async_llm_engine import AsyncStream
    import asyncio

    async def drain(parts, stop):
        st = AsyncStream(None, stop)
        for p in parts:
            st.put(p)
        st.finish("length")
        return "".join([x async for x in st]), st.finish_reason, st.stop_hit

    assert asyncio.run(drain(["ab", "c<", "/s", ">zz"], ["</s>"])) == ("abc", "stop", True)      # split across deltas
    assert asyncio.run(drain(["ab", "c<", "/x", "y"], ["</s>"])) == ("abc</xy", "length", False)   # false alarm: flushed
    assert asyncio.run(drain(["abc<"], ["</s>", "q"])) == ("abc<", "length", False)               # held tail released
    assert asyncio.run(drain(["hello world"], "o w")) == ("hell", "stop", True)

    c, engine = client
    body = {"prompt": "hello world how are you", "max_tokens": 12, "temperature": 0.0, "top_k": 1, "ignore_eos": True}
    full = c.post("/v1/completions", json=body).json()["choices"][0]["text"]
    words = full.split()
    assert len(words) >= 6
    stop = " " + words[3] + " "
    r = c.post("/v1/completions", json=dict(body, stop=[stop])).json()
    assert r["choices"][0]["finish_reason"] == "stop"
    assert r["choices"][0]["text"] == full[:full.find(stop)] and stop not in r["choices"][0]["text"]
    assert r["usage"]["completion_tokens"] < 12                       # the request was cut short in the engine too
    # streaming: same text, [DONE] framing intact
    chunks = []
    with c.stream("POST", "/v1/completions", json=dict(body, stop=stop, stream=True)) as resp:
        for line in resp.iter_lines():
            if line.startswith("data: ") and line != "data: [DONE]":
                chunks.append(json.loads(line[6:]))
    assert "".join(ch["choices"][0]["text"] for ch in chunks) == full[:full.find(stop)]
    assert chunks[-1]["choices"][0]["finish_reason"] == "stop"


def test_malformed_requests_are_refused_and_do_not_hurt_the_engine(client):
    """Empty prompts, non-positive budgets, out-of-vocabulary ids and nonsensical sampling parameters get a 400;
    none of them may reach the engine (an empty prompt used to take the whole engine loop down)."""
    c, engine = client
    bad = [
        ("/v1/completions", {"prompt": "", "max_tokens": 4}),
        ("/v1/completions", {"prompt": [[]], "max_tokens": 4}),
        ("/v1/completions", {"prompt": [], "max_tokens": 4}),
        ("/v1/completions", {"prompt": "hello", "max_tokens": 0}),
        ("/v1/completions", {"prompt": "hello", "max_tokens": -3}),
        ("/v1/completions", {"prompt": [9999999], "max_tokens": 4}),
        ("/v1/completions", {"prompt": [-1], "max_tokens": 4}),
        ("/v1/completions", {"prompt": "hello", "max_tokens": 100000}),
        ("/v1/completions", {"prompt": "hello", "max_tokens": 4, "temperature": -1.0}),
        ("/v1/completions", {"prompt": "hello", "max_tokens": 4, "top_p": 0.0}),
        ("/v1/completions", {"prompt": "hello", "max_tokens": 4, "repetition_penalty": 0.0}),
        ("/v1/chat/completions", {"messages": [], "max_tokens": 4}),
        ("/v1/chat/completions", {"messages": [{"role": "user", "content": "hello"}], "max_completion_tokens": 0}),
    ]
    before = engine.metrics["requests_total"]
    for url, body in bad:
        r = c.post(url, json=body)
        assert r.status_code == 400 and r.json()["object"] == "error", (url, body, r.status_code, r.text[:200])
    assert engine.metrics["requests_total"] == before and engine.failed is None
    ok = c.post("/v1/completions", json={"prompt": "hello world", "max_tokens": 3, "stop": ["", "zzz"]})
    assert ok.status_code == 200 and ok.json()["usage"]["completion_tokens"] == 3
    assert c.get("/health").status_code == 200
    with pytest.raises(ValueError):
        engine.allocate_seq([], 4)


def test_disable_thinking_reaches_the_chat_template():
    """`--disable-thinking` / `use_thinking=False` -> `enable_thinking=False` in `apply_chat_template`
    (reference: model_runner.py:201-214)."""
    from gllm_b200 import LLM
    from transformers import AutoTokenizer
    d = _make_model_dir()
    tok = AutoTokenizer.from_pretrained(d)
    tok.chat_template = ("{% for m in messages %}<|{{ m['role'] }}|> {{ m['content'] }} {% endfor %}"
                         "{% if add_generation_prompt %}<|assistant|> {% if enable_thinking is defined and not "
                         "enable_thinking %}w1 w2 {% endif %}{% endif %}")
    tok.save_pretrained(d)
    msgs = [{"role": "user", "content": "hello world"}]
    ids = {}
    for flag in (True, False):
        llm = LLM(d, maxp=32, maxd=8, num_cpu_pages=32, model_max_length=128, log_stats=False, use_thinking=flag)
        ids[flag] = llm.encode(None, True, msgs)
        llm.shutdown()
    extra = tok.encode("w1 w2", add_special_tokens=False)
    assert ids[False][-len(extra):] == extra and ids[True] == ids[False][:-len(extra)]


def test_disconnect_mid_chunked_prefill_and_penalty_flood_do_not_kill_the_engine():
    """ADVICE r1 (high + medium): (1) a client that goes away while its prompt is between two prefill chunks used to
    leave a freed sequence at the head of the prefill queue -> double free of its id -> engine loop dead for good;
    (2) more queued repetition-penalty requests than max_running_seqs exhausted the penalty-state pool on arrival."""
    import asyncio
    from gllm_b200.engine.async_llm_engine import AsyncLLM
    d = _make_model_dir()
    engine = AsyncLLM(d, maxp=16, maxd=16, num_cpu_pages=96, model_max_length=256, log_stats=False)

    async def scenario():
        prompt = [5 + (i * 7) % 190 for i in range(60)]
        st = await engine.add_requests_async(None, prompt, output_len=8, ignore_eos=True, temperature=0.0, top_k=1)
        for _ in range(2000):                      # let the first chunk(s) run, not the whole prompt
            await asyncio.sleep(0.001)
            seq = engine.running_maps.get(st.seq_id)
            if seq is not None and 0 < seq.computed_token_num < 60:
                break
        engine.abort_stream(st)
        engine.abort_stream(st)                    # idempotent
        # the same prompt again plus a flood of penalty requests (24 > maxd = 16 = max_running_seqs)
        streams = [await engine.add_requests_async(None, prompt, output_len=6, ignore_eos=True, temperature=0.0,
                                                   top_k=1)]
        for i in range(24):
            streams.append(await engine.add_requests_async(None, [9 + i, 17, 23 + i, 31], output_len=5,
                                                           ignore_eos=True, temperature=0.0, top_k=1,
                                                           repetition_penalty=1.1))
        texts = [await asyncio.wait_for(engine.collect(s), timeout=120) for s in streams]
        return st, streams, texts

    st, streams, texts = asyncio.run(scenario())
    assert engine.failed is None
    assert all(s.finish_reason == "length" for s in streams), [s.finish_reason for s in streams]
    assert streams[0].completion_tokens == 6 and all(s.completion_tokens == 5 for s in streams[1:])
    assert engine.metrics["requests_aborted"] == 1
    sch = engine.worker.scheduler
    assert not sch.abort_ids and not sch.has_work()
    assert engine.worker.mm.get_num_free_pages() == engine.worker.mm.usable_pages      # nothing leaked
    assert not engine.worker._seq_slots and not engine.running_maps
    engine.shutdown()
