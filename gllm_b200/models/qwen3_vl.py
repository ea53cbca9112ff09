"""Qwen3-VL / Qwen3-VL-MoE: Qwen3 (dense or MoE) decoder with interleaved M-RoPE, ViT with learned
position-table interpolation and DeepStack feature injection into the first decoder layers.

Reference: gllm/models/qwen3_vl.py:310-568, gllm/models/qwen3_vl_moe.py:85-129 (fused expert tensors
`experts.gate_up_proj [E, H, 2I]` / `experts.down_proj [E, I, H]`, handled by SparseMoeBlock.load_weights).
"""
from __future__ import annotations

from gllm_b200.models.multimodal import VLCausalLM
from gllm_b200.models.qwen2_5_vl import text_config
from gllm_b200.models.vision import Qwen3VisionTower


def build_qwen3_vl(cfg, device):
    from gllm_b200.layers.moe import make_moe_block
    from gllm_b200.models.registry import spec_qwen3, spec_qwen3_moe
    tc = text_config(cfg)
    moe = cfg["architectures"][0] == "Qwen3VLMoeForConditionalGeneration"
    spec = spec_qwen3_moe(tc) if moe else spec_qwen3(tc)
    spec.arch = "qwen3_vl_moe" if moe else "qwen3_vl"
    rs = dict(spec.rope_scaling or {})
    assert "mrope_section" in rs, "Qwen3-VL needs rope mrope_section"
    rs.setdefault("mrope_interleaved", True)
    spec.rope_scaling = rs
    return VLCausalLM(spec, cfg, device, moe_factory=make_moe_block if moe else None,
                      vision_factory=Qwen3VisionTower)
