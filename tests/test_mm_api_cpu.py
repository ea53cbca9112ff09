"""Multimodal request path through the OpenAI API on CPU: image_url (data URL) parsing, processor call (a small
stand-in for the HF `AutoProcessor`, which needs real model files), M-RoPE position computation, pixel payload to
the engine, vision tower + DeepStack in the forward. Text-only requests on the same server still work."""
from conftest import scratch_dir
import base64
import io
import json

import numpy as np
import pytest
import torch

pytest.importorskip("fastapi")
transformers = pytest.importorskip("transformers")
PIL = pytest.importorskip("PIL")

IMG, VID, VSTART = 290, 291, 292


class _FakeProcessor:
    """Implements the two calls `encode_mm` makes; every image becomes a 1 x 4 x 6 patch grid (6 merged tokens)."""

    def __init__(self, tok):
        self.tok = tok

    def apply_chat_template(self, msgs, tokenize=False, add_generation_prompt=True):
        out = []
        for m in msgs:
            c = m["content"]
            if isinstance(c, str):
                out.append(f"<|{m['role']}|> {c}")
            else:
                parts = ["<|vision_start|> <|image_pad|>" if p["type"] == "image" else p["text"] for p in c]
                out.append(f"<|{m['role']}|> " + " ".join(parts))
        return " ".join(out) + (" <|assistant|>" if add_generation_prompt else "")

    def __call__(self, text, images=None, videos=None, return_tensors="np"):
        ids = []
        for t in self.tok.encode(text[0]):
            ids += [IMG] * 6 if t == IMG else [t]
        pix = []
        for im in images:
            a = np.asarray(im.resize((96, 64)), dtype=np.float32) / 255.0       # [64, 96, 3]
            g = np.random.default_rng(int(a.sum() * 1000) % (2 ** 31))
            pix.append(g.standard_normal((24, 3 * 2 * 16 * 16)).astype(np.float32))
        return {"input_ids": np.asarray([ids]), "pixel_values": np.concatenate(pix, 0),
                "image_grid_thw": np.asarray([[1, 4, 6]] * len(images))}


def _make_vl_dir():
    from tokenizers import Tokenizer, models, pre_tokenizers
    from transformers import PreTrainedTokenizerFast, Qwen3VLConfig, Qwen3VLForConditionalGeneration
    torch.manual_seed(0)
    words = ["<unk>", "<s>", "</s>", "<|user|>", "<|assistant|>"] + [f"w{i}" for i in range(277)] + \
        ["hello", "what", "is", "this", "?", "a", "picture", "of"]
    assert len(words) == IMG
    words += ["<|image_pad|>", "<|video_pad|>", "<|vision_start|>"] + [f"x{i}" for i in range(7)]
    vocab = {w: i for i, w in enumerate(words)}
    tok = Tokenizer(models.WordLevel(vocab, unk_token="<unk>"))
    tok.pre_tokenizer = pre_tokenizers.WhitespaceSplit()
    fast = PreTrainedTokenizerFast(tokenizer_object=tok, unk_token="<unk>", bos_token="<s>", eos_token="</s>")
    fast.chat_template = "{% for m in messages %}<|{{ m['role'] }}|> {{ m['content'] }} {% endfor %}" \
                         "{% if add_generation_prompt %}<|assistant|> {% endif %}"
    text = dict(hidden_size=64, num_hidden_layers=4, num_attention_heads=4, num_key_value_heads=2, head_dim=16,
                intermediate_size=128, vocab_size=len(words), max_position_embeddings=512,
                rope_parameters={"rope_type": "default", "mrope_section": [2, 3, 3], "mrope_interleaved": True,
                                 "rope_theta": 10000.0})
    vision = dict(depth=3, hidden_size=32, num_heads=2, intermediate_size=64, out_hidden_size=64, patch_size=16,
                  spatial_merge_size=2, temporal_patch_size=2, num_position_embeddings=16,
                  deepstack_visual_indexes=[0, 1], in_channels=3)
    cfg = Qwen3VLConfig(text_config=text, vision_config=vision, image_token_id=IMG, video_token_id=VID,
                        vision_start_token_id=VSTART, tie_word_embeddings=False)
    d = scratch_dir("gllm_b200_vlapi_")
    Qwen3VLForConditionalGeneration(cfg).eval().float().save_pretrained(d, safe_serialization=True)
    fast.save_pretrained(d)
    return d


def _data_url(color):
    from PIL import Image
    buf = io.BytesIO()
    Image.new("RGB", (40, 30), color).save(buf, format="PNG")
    return "data:image/png;base64," + base64.b64encode(buf.getvalue()).decode()


def test_chat_completion_with_images_and_text_only(monkeypatch):
    from fastapi.testclient import TestClient
    from gllm_b200.engine.async_llm_engine import AsyncLLM
    from gllm_b200.entrypoints.api_server import build_app
    from gllm_b200.models import multimodal as mmod
    d = _make_vl_dir()
    engine = AsyncLLM(d, maxp=64, maxd=16, num_cpu_pages=64, model_max_length=256, log_stats=False)
    assert engine.loader.use_mm and engine.tokenizer is not None
    monkeypatch.setattr(mmod, "get_processor", lambda llm: _FakeProcessor(llm.tokenizer))
    try:
        with TestClient(build_app(engine)) as c:
            content = [{"type": "image_url", "image_url": {"url": _data_url((255, 0, 0))}},
                       {"type": "text", "text": "what is this ?"},
                       {"type": "image_url", "image_url": {"url": _data_url((0, 0, 255))}}]
            r = c.post("/v1/chat/completions", json={"messages": [{"role": "user", "content": content}],
                                                      "max_completion_tokens": 5, "ignore_eos": True, "top_k": 1})
            assert r.status_code == 200, r.text
            body = r.json()
            assert body["usage"]["completion_tokens"] == 5
            # 2 images x (vision_start + 6 pads) + role tokens + 4 words + assistant tag
            assert body["usage"]["prompt_tokens"] == 1 + 2 * 7 + 4 + 1
            # the same conversation again must give the same greedy answer (pixels -> same embeddings); a different
            # picture is a different prompt for the prefix cache (hash salt) even though the token ids are equal
            r2 = c.post("/v1/chat/completions", json={"messages": [{"role": "user", "content": content}],
                                                       "max_completion_tokens": 5, "ignore_eos": True, "top_k": 1})
            assert r2.json()["choices"][0]["message"]["content"] == body["choices"][0]["message"]["content"]
            # text only on the multimodal server
            r3 = c.post("/v1/chat/completions", json={"messages": [{"role": "user", "content": "hello what is this"}],
                                                       "max_completion_tokens": 3, "ignore_eos": True, "top_k": 1})
            assert r3.status_code == 200 and r3.json()["usage"]["completion_tokens"] == 3
    finally:
        engine.shutdown()
