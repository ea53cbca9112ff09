"""Multimodal plumbing shared by the Qwen-VL families.

Front-end side (driver process): OpenAI-style message parsing, HF processor call, 3-axis M-RoPE positions
computed on the CPU once per request (reference: gllm/model_runner.py:219-246, gllm/layers/rotary_embedding.py:627-740).
Worker side: `VLCausalLM` — vision tower on the first pipeline stage only, per-sequence embedding cache so a
chunked prefill only runs the tower once (reference: gllm/model_runner.py:248-340), placeholder merge and
Qwen3-VL DeepStack injection (reference: gllm/models/qwen3_vl.py:398-450).

Wire format: `BatchArrays.mm = {"new": {seq_id: {"pixel_values", "grid_thw"}}, "chunks": [(seq_id, tok_off,
n_tok, vis_before, vis_total)]}` — pixel data travels once per (re)computation of a sequence.
"""
from __future__ import annotations

import base64
import hashlib
import io
from dataclasses import dataclass
from typing import List, Optional, Tuple

import numpy as np
import torch

from gllm_b200.models.decoder import CausalLM


@dataclass
class MMInfo:
    image_token_id: int
    video_token_id: int
    spatial_merge_size: int
    tokens_per_second: float = 1.0
    per_frame_video: bool = False   # Qwen3-VL: every video frame is its own (t=1) vision segment

    @staticmethod
    def from_config(cfg) -> "MMInfo":
        vc = cfg.get("vision_config", {}) or {}
        arch = cfg["architectures"][0]
        return MMInfo(image_token_id=cfg.get("image_token_id", 151655), video_token_id=cfg.get("video_token_id", 151656),
                      spatial_merge_size=vc.get("spatial_merge_size", 2),
                      tokens_per_second=vc.get("tokens_per_second", 1.0),
                      per_frame_video=arch.startswith("Qwen3VL"))


# ------------------------------------------------------------------------------------------------
# M-RoPE positions
# ------------------------------------------------------------------------------------------------
def compute_mrope_positions(token_ids, image_grid_thw, video_grid_thw, info: MMInfo,
                            second_per_grid_ts=None) -> Tuple[np.ndarray, int]:
    """-> (positions int32 [3, L], delta) where position of generated token i is i + delta.

    Text runs advance all three axes together; a vision segment of grid (t, h, w) starting at p gets
    t-axis p + frame*interval, h-axis p + row, w-axis p + col (rows/cols after the spatial merge) and the next
    text token continues at max + 1."""
    toks = np.asarray(token_ids, dtype=np.int64)
    L = toks.shape[0]
    pos = np.zeros((3, L), dtype=np.int64)
    m = info.spatial_merge_size
    img = [tuple(int(x) for x in g) for g in (image_grid_thw if image_grid_thw is not None else [])]
    vid = [tuple(int(x) for x in g) for g in (video_grid_thw if video_grid_thw is not None else [])]
    if info.per_frame_video:  # one placeholder run per frame
        vid = [(1, h, w) for (t, h, w) in vid for _ in range(t)]
        spg = [1.0] * len(vid)
    else:
        spg = list(second_per_grid_ts) if second_per_grid_ts is not None else [1.0] * len(vid)
    ii = vi = 0
    i, cur = 0, 0
    order = []   # vision segments in prompt order, indexed [images..., videos...] (per-frame runs collapse)
    is_vis = (toks == info.image_token_id) | (toks == info.video_token_id)
    while i < L:
        if not is_vis[i]:
            j = i
            while j < L and not is_vis[j]:
                j += 1
            pos[:, i:j] = cur + np.arange(j - i)
            cur += j - i
            i = j
            continue
        if toks[i] == info.image_token_id:
            t, h, w = img[ii]
            order.append(("i", ii))
            ii += 1
            interval = 1.0
        else:
            t, h, w = vid[vi]
            order.append(("v", vi))
            interval = info.tokens_per_second * float(spg[vi]) if not info.per_frame_video else 1.0
            vi += 1
        gh, gw = h // m, w // m
        n = t * gh * gw
        assert i + n <= L and bool(is_vis[i:i + n].all()), "placeholder run does not match the vision grid"
        tt = (np.arange(t) * interval).astype(np.int64).repeat(gh * gw)
        hh = np.tile(np.arange(gh).repeat(gw), t)
        ww = np.tile(np.arange(gw), t * gh)
        pos[0, i:i + n] = cur + tt
        pos[1, i:i + n] = cur + hh
        pos[2, i:i + n] = cur + ww
        cur = int(pos[:, i:i + n].max()) + 1
        i += n
    delta = int(cur - L)
    compute_mrope_positions.last_order = order
    return pos.astype(np.int32), delta


def prepare_mm_sequence(seq, info: MMInfo):
    """Fill `seq.mm_state` (positions, delta, vision-token count, prefix-hash salt) from `seq.mm_contents`."""
    mc = seq.mm_contents
    if not mc:
        return
    pos, delta = compute_mrope_positions(seq.token_ids[:seq.prompt_len], mc.get("image_grid_thw"),
                                         mc.get("video_grid_thw"), info, mc.get("second_per_grid_ts"))
    toks = np.asarray(seq.token_ids[:seq.prompt_len])
    vis = (toks == info.image_token_id) | (toks == info.video_token_id)
    h = hashlib.sha1()
    for key in ("pixel_values", "pixel_values_videos"):
        if mc.get(key) is not None:
            h.update(np.ascontiguousarray(mc[key]).tobytes())
    n_img = 0 if mc.get("image_grid_thw") is None else len(mc["image_grid_thw"])
    order, seen = [], set()
    vg = mc.get("video_grid_thw")
    frame_owner = []  # per-frame video runs -> index of the video they belong to
    if info.per_frame_video and vg is not None:
        for k, g in enumerate(vg):
            frame_owner += [k] * int(g[0])
    for kind, k in compute_mrope_positions.last_order:
        seg = k if kind == "i" else n_img + (frame_owner[k] if frame_owner else k)
        if seg not in seen:
            seen.add(seg)
            order.append(seg)
    seq.mm_state = {"positions": pos, "order": order, "vis_cum": np.concatenate([[0], np.cumsum(vis)]).astype(np.int32),
                    "salt": int.from_bytes(h.digest()[:8], "little"), "sent": False}
    seq.mrope_delta = delta


def seq_positions(seq, start: int, n: int) -> np.ndarray:
    """[3, n] positions of tokens [start, start+n) of a sequence (text-only sequences: plain arange)."""
    st = seq.mm_state
    ar = np.arange(start, start + n, dtype=np.int32)
    if not st:
        return np.broadcast_to(ar, (3, n))
    p = st["positions"]
    pl = p.shape[1]
    if start + n <= pl:
        return p[:, start:start + n]
    out = np.empty((3, n), dtype=np.int32)
    k = max(0, pl - start)
    if k:
        out[:, :k] = p[:, start:]
    out[:, k:] = ar[k:] + seq.mrope_delta
    return out


def batch_mm_payload(entries, qsl) -> Optional[dict]:
    """Driver side: the `mm` dict for a batch (None when no scheduled chunk touches vision tokens)."""
    new, chunks = {}, []
    for i, e in enumerate(entries):
        seq = e.seq
        st = seq.mm_state
        if not st or e.start >= seq.prompt_len:
            continue
        cum = st["vis_cum"]
        end = min(e.start + e.n, seq.prompt_len)
        before, upto = int(cum[e.start]), int(cum[end])
        if upto == before:
            continue
        if not st["sent"] or e.start == 0:
            mc = seq.mm_contents
            pv = [mc[k] for k in ("pixel_values", "pixel_values_videos") if mc.get(k) is not None]
            gr = [mc[k] for k in ("image_grid_thw", "video_grid_thw") if mc.get(k) is not None]
            # image segments precede video segments in the cache only when the prompt orders them so; mixed
            # prompts keep per-modality order, which is what the placeholder scan in the worker relies on
            new[seq.seq_id] = {"pixel_values": np.concatenate(pv, 0), "grid_thw": np.concatenate(gr, 0),
                               "order": st["order"]}
            st["sent"] = True
        chunks.append((seq.seq_id, int(qsl[i]), int(e.n), before, int(cum[-1])))
    if not chunks:
        return None
    return {"new": new, "chunks": chunks}


# ------------------------------------------------------------------------------------------------
# front end: messages -> (token ids, mm_contents)
# ------------------------------------------------------------------------------------------------
def _load_image(url: str):
    from PIL import Image
    if url.startswith("data:"):
        raw = base64.b64decode(url.split(",", 1)[1])
        return Image.open(io.BytesIO(raw)).convert("RGB")
    if url.startswith("file://"):
        url = url[len("file://"):]
    if url.startswith(("http://", "https://")):
        import requests
        return Image.open(io.BytesIO(requests.get(url, timeout=30).content)).convert("RGB")
    return Image.open(url).convert("RGB")


def split_messages(messages):
    """OpenAI-style content parts -> (HF-style messages with bare image/video placeholders, images, videos)."""
    out, images, videos = [], [], []
    for m in messages:
        m = m if isinstance(m, dict) else m.model_dump()
        c = m.get("content")
        if isinstance(c, str) or c is None:
            out.append({"role": m["role"], "content": c or ""})
            continue
        parts = []
        for p in c:
            t = p.get("type")
            if t == "text":
                parts.append({"type": "text", "text": p["text"]})
            elif t in ("image_url", "image"):
                u = p.get("image_url", p.get("image"))
                u = u["url"] if isinstance(u, dict) else u
                images.append(_load_image(u))
                parts.append({"type": "image"})
            elif t in ("video_url", "video"):
                u = p.get("video_url", p.get("video"))
                u = u["url"] if isinstance(u, dict) else u
                videos.append(u[len("file://"):] if u.startswith("file://") else u)
                parts.append({"type": "video"})
        out.append({"role": m["role"], "content": parts})
    return out, images, videos


def get_processor(llm):
    proc = getattr(llm, "_mm_processor", None)
    if proc is None:
        from transformers import AutoProcessor
        kw = {}
        if llm.cfg.mm_processor_min_pixels:
            kw["min_pixels"] = llm.cfg.mm_processor_min_pixels
        if llm.cfg.mm_processor_max_pixels:
            kw["max_pixels"] = llm.cfg.mm_processor_max_pixels
        proc = AutoProcessor.from_pretrained(llm.cfg.tokenizer_path or llm.cfg.model_path, **kw)
        llm._mm_processor = proc
    return proc


def encode_mm(llm, messages) -> Tuple[List[int], Optional[dict]]:
    """Chat messages (possibly with images / videos) -> prompt token ids with expanded placeholders + the
    pixel payload. Text-only conversations return (ids, None)."""
    msgs, images, videos = split_messages(messages)
    if not images and not videos:
        return llm.encode(None, True, msgs), None
    proc = get_processor(llm)
    text = proc.apply_chat_template(msgs, tokenize=False, add_generation_prompt=True)
    inputs = proc(text=[text], images=images or None, videos=videos or None, return_tensors="np")
    mm = {}
    for k in ("pixel_values", "image_grid_thw", "pixel_values_videos", "video_grid_thw", "second_per_grid_ts"):
        if k in inputs and inputs[k] is not None:
            mm[k] = np.asarray(inputs[k])
    return [int(t) for t in inputs["input_ids"][0]], mm


def extract_mm_contents(llm, messages) -> Optional[dict]:
    return encode_mm(llm, messages)[1]


# ------------------------------------------------------------------------------------------------
# worker side
# ------------------------------------------------------------------------------------------------
class VLCausalLM(CausalLM):
    """Text decoder + vision tower. `self.visual(pixel_values, grid_thw)` returns (embeds [N, H], deepstack
    list of [N, H]); only the first pipeline stage owns it (reference: gllm/models/qwen2_5_vl.py:686-690)."""

    def __init__(self, spec, cfg, device, moe_factory=None, vision_factory=None):
        super().__init__(spec, device, moe_factory)
        self.mm_info = MMInfo.from_config(cfg)
        self.visual = vision_factory(cfg["vision_config"], spec.dtype, device) if self.is_first else None
        self.num_deepstack = len((cfg.get("vision_config") or {}).get("deepstack_visual_indexes", []) or [])
        if self.num_deepstack:
            assert not self.is_first or self.num_deepstack < len(self.layers), \
                "DeepStack layers must live on the first pipeline stage"
        self._mm_cache = {}

    def init_dummy(self, seed: int = 0):
        super().init_dummy(seed)
        if self.visual is not None:
            g = torch.Generator(device="cpu").manual_seed(seed + 77)   # replicated: same on every rank
            for name, p in self.visual.named_parameters():
                if "norm" in name or name.endswith("ln_q.weight"):
                    p.data.fill_(0.0 if name.endswith("bias") else 1.0)
                elif p.dim() == 1:
                    p.data.zero_()
                else:
                    p.data.copy_((torch.randn(p.shape, generator=g) * 0.02).to(p.dtype))

    def load_weights(self, reader, progress=None):
        super().load_weights(reader, progress)
        if self.visual is not None:
            from gllm_b200.models.vision import load_vision_weights
            load_vision_weights(self.visual, reader)

    @torch.no_grad()
    def _vision(self, item):
        pv = torch.from_numpy(np.ascontiguousarray(item["pixel_values"])).to(self.device)
        grid = torch.from_numpy(np.asarray(item["grid_thw"], dtype=np.int64))
        emb, deep = self.visual(pv.to(self.spec.dtype), grid)
        order = item.get("order")
        if order is not None and list(order) != sorted(order):
            # the tower ran over [images..., videos...]; the prompt interleaves them differently
            m2 = self.mm_info.spatial_merge_size ** 2
            lens = [int(t * h * w) // m2 for t, h, w in grid.tolist()]
            perm = torch.cat([torch.arange(sum(lens[:k]), sum(lens[:k + 1])) for k in order]).to(emb.device)
            emb, deep = emb[perm], [d[perm] for d in deep]
        return emb, deep

    def forward(self, inp, kv_cache, tpc, hidden=None, residual=None, inputs_embeds=None, recv_tiles=None):
        mm = getattr(inp.batch, "mm", None) if inp.batch is not None else None
        deep = None
        if self.is_first and mm and inputs_embeds is None:
            x = self.embed(inp, tpc)
            toks = inp.batch.tokens
            idx_all, emb_all, deep_all = [], [], [[] for _ in range(self.num_deepstack)]
            for sid, item in mm["new"].items():
                self._mm_cache[sid] = self._vision(item)
            for sid, off, n, before, total in mm["chunks"]:
                emb, ds = self._mm_cache[sid]
                t = toks[off:off + n]
                loc = np.nonzero((t == self.mm_info.image_token_id) | (t == self.mm_info.video_token_id))[0]
                idx_all.append(loc + off)
                emb_all.append(emb[before:before + len(loc)])
                for k in range(self.num_deepstack):
                    deep_all[k].append(ds[k][before:before + len(loc)])
                if before + len(loc) >= total:
                    self._mm_cache.pop(sid, None)
            idx = torch.from_numpy(np.concatenate(idx_all)).to(self.device)
            x = x.clone() if x.data_ptr() == self.embed_w.data_ptr() else x
            x.index_copy_(0, idx, torch.cat(emb_all, 0).to(x.dtype))
            inputs_embeds = x
            if self.num_deepstack:
                deep = (idx, [torch.cat(d, 0).to(x.dtype) for d in deep_all])
        return super().forward(inp, kv_cache, tpc, hidden, residual, inputs_embeds=inputs_embeds,
                               recv_tiles=recv_tiles, deepstack=deep)
