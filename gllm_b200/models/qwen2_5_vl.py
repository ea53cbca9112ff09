"""Qwen2.5-VL: Qwen2 text decoder with 3-axis M-RoPE + windowed-attention vision tower.

Reference: gllm/models/qwen2_5_vl.py:686-989 (model wiring, weight names), HF config layout of
transformers >= 4.52 (`text_config` / `vision_config`) and the older flat layout.
"""
from __future__ import annotations

from gllm_b200.models.multimodal import VLCausalLM
from gllm_b200.models.vision import Qwen2_5_VisionTower


def text_config(cfg):
    """The decoder's config: nested `text_config` (new layout) overlaid on the top-level keys (old layout)."""
    from gllm_b200.models.registry import HFConfig
    tc = dict(cfg)
    tc.update(cfg.get("text_config") or {})
    tc.pop("text_config", None)
    tc["architectures"] = cfg["architectures"]
    if "torch_dtype" not in tc and "dtype" not in tc:
        tc["torch_dtype"] = cfg.get("torch_dtype", cfg.get("dtype", "bfloat16"))
    return HFConfig(tc)


def build_qwen2_5_vl(cfg, device):
    from gllm_b200.models.registry import spec_qwen2
    spec = spec_qwen2(text_config(cfg))
    spec.arch = "qwen2_5_vl"
    assert spec.rope_scaling and "mrope_section" in spec.rope_scaling, "Qwen2.5-VL needs rope mrope_section"
    return VLCausalLM(spec, cfg, device, vision_factory=Qwen2_5_VisionTower)
