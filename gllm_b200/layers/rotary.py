"""RoPE frequency tables: default / linear / llama3 / YaRN (+ M-RoPE sections).

One fp32 `[max_pos, rot_dim]` table (cos | sin) per model; the fused rope kernel gathers rows by
position (csrc/elemwise/rope_kv.cu). Reference: gllm/layers/rotary_embedding.py:27-336.
"""
from __future__ import annotations

import math
from dataclasses import dataclass
from typing import List, Optional

import torch

from gllm_b200.ops.ref import build_cos_sin_cache


def _default_inv_freq(rot_dim: int, base: float) -> torch.Tensor:
    return 1.0 / (base ** (torch.arange(0, rot_dim, 2, dtype=torch.float32) / rot_dim))


def _llama3_inv_freq(rot_dim, base, factor, low_freq_factor, high_freq_factor, orig_max_pos):
    inv = _default_inv_freq(rot_dim, base)
    low_wavelen = orig_max_pos / low_freq_factor
    high_wavelen = orig_max_pos / high_freq_factor
    wavelen = 2 * math.pi / inv
    if low_freq_factor != high_freq_factor:
        smooth = (orig_max_pos / wavelen - low_freq_factor) / (high_freq_factor - low_freq_factor)
    else:
        smooth = torch.zeros_like(inv)
    return torch.where(wavelen < high_wavelen, inv,
                       torch.where(wavelen > low_wavelen, inv / factor,
                                   (1 - smooth) * inv / factor + smooth * inv))


def yarn_get_mscale(scale: float = 1.0, mscale: float = 1.0) -> float:
    if scale <= 1:
        return 1.0
    return 0.1 * mscale * math.log(scale) + 1.0


def _yarn_find_correction_dim(num_rot, dim, base, max_pos):
    return (dim * math.log(max_pos / (num_rot * 2 * math.pi))) / (2 * math.log(base))


def _yarn_inv_freq(rot_dim, base, factor, orig_max_pos, beta_fast=32, beta_slow=1):
    pos_freqs = base ** (torch.arange(0, rot_dim, 2, dtype=torch.float32) / rot_dim)
    extrap = 1.0 / pos_freqs
    interp = 1.0 / (factor * pos_freqs)
    low = max(math.floor(_yarn_find_correction_dim(beta_fast, rot_dim, base, orig_max_pos)), 0)
    high = min(math.ceil(_yarn_find_correction_dim(beta_slow, rot_dim, base, orig_max_pos)), rot_dim - 1)
    if low == high:
        high += 0.001
    ramp = ((torch.arange(rot_dim // 2, dtype=torch.float32) - low) / (high - low)).clamp(0, 1)
    mask = 1 - ramp
    return interp * (1 - mask) + extrap * mask


@dataclass
class RopeSpec:
    rot_dim: int
    neox: bool = True
    mrope_section: Optional[List[int]] = None
    cos_sin: Optional[torch.Tensor] = None  # fp32 [max_pos, rot_dim]
    attn_mscale: float = 1.0  # extra softmax-scale factor (YaRN with mscale_all_dim)


def build_rope(head_dim: int, max_position: int, base: float, rope_scaling: Optional[dict] = None,
               rot_dim: Optional[int] = None, neox: bool = True, device="cpu") -> RopeSpec:
    rot = head_dim if rot_dim is None else rot_dim
    spec = RopeSpec(rot_dim=rot, neox=neox)
    max_pos = max_position
    inv_freq = None
    mscale = 1.0
    if rope_scaling:
        kind = rope_scaling.get("rope_type", rope_scaling.get("type", "default"))
        if "mrope_section" in rope_scaling:
            spec.mrope_section = list(rope_scaling["mrope_section"])[:3]
            if rope_scaling.get("mrope_interleaved"):
                spec.mrope_section.append(1)  # 4th entry flags the interleaved layout
        if kind == "linear":
            factor = rope_scaling["factor"]
            inv_freq = _default_inv_freq(rot, base) / factor
            max_pos = int(max_position * factor)
        elif kind == "llama3":
            inv_freq = _llama3_inv_freq(rot, base, rope_scaling["factor"], rope_scaling["low_freq_factor"],
                                        rope_scaling["high_freq_factor"],
                                        rope_scaling["original_max_position_embeddings"])
        elif kind in ("yarn", "deepseek_yarn"):
            factor = rope_scaling["factor"]
            orig = rope_scaling["original_max_position_embeddings"]
            inv_freq = _yarn_inv_freq(rot, base, factor, orig, rope_scaling.get("beta_fast", 32),
                                      rope_scaling.get("beta_slow", 1))
            m = rope_scaling.get("mscale", 1.0)
            m_all = rope_scaling.get("mscale_all_dim", 0.0)
            if kind == "deepseek_yarn" or "mscale_all_dim" in rope_scaling:
                mscale = yarn_get_mscale(factor, m) / yarn_get_mscale(factor, m_all) if m_all else yarn_get_mscale(factor, m)
                if m_all:
                    am = yarn_get_mscale(factor, m_all)
                    spec.attn_mscale = am * am
            else:
                mscale = yarn_get_mscale(factor)
            max_pos = int(orig * factor)
        elif kind in ("default", "mrope", None):
            pass
        elif kind == "dynamic":
            pass  # dynamic NTK only matters beyond the trained length; table is the default one
        else:
            raise NotImplementedError(f"rope scaling type {kind}")
    spec.cos_sin = build_cos_sin_cache(rot, max_pos, base, inv_freq, mscale).to(device)
    return spec
