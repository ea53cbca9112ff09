"""Vision towers of the Qwen-VL families (Qwen2.5-VL windowed ViT, Qwen3-VL ViT with DeepStack).

Runs once per image on the first pipeline stage; not on the serving hot path, so it uses PyTorch
library ops (SDPA per attention segment) — SURVEY §7 P8. Parameter names mirror the HF checkpoints
(`visual.*`), so loading is a name walk. Every TP rank holds the full tower (the reference shards its QKV
and re-interleaves with an all-gather, gllm/models/qwen2_5_vl.py:160-239; at <1 B parameters replication is
cheaper than the extra collectives).

Reference: gllm/models/qwen2_5_vl.py:177-381, gllm/models/qwen3_vl.py:310-450.
"""
from __future__ import annotations

from typing import List

import torch
import torch.nn.functional as F
from torch import nn
from gllm_b200.utils import clamp_overflow


def _rotate_half(x):
    h = x.shape[-1] // 2
    return torch.cat((-x[..., h:], x[..., :h]), dim=-1)


def _apply_rope(q, k, cos, sin):
    dt = q.dtype
    q, k = q.float(), k.float()
    cos, sin = cos.unsqueeze(-2).float(), sin.unsqueeze(-2).float()
    return (q * cos + _rotate_half(q) * sin).to(dt), (k * cos + _rotate_half(k) * sin).to(dt)


# ------------------------------------------------------------------------------------------------
# tensor parallelism inside the towers (reference: QKVParallel + all_gather_interleave in
# gllm/models/qwen2_5_vl.py:160-239). Here: attention heads and the MLP intermediate dimension
# are split over the TP group, the two row-parallel projections of a block end in an all-reduce; patch embedding,
# position tables and the mergers stay replicated (small). `GLLM_VISION_TP=0` keeps the whole tower replicated.
# ------------------------------------------------------------------------------------------------
def vision_tp(heads: int, inter: int):
    """-> (rank, size) the tower is sharded over; (0, 1) when not distributed or the shapes do not divide."""
    import os
    from gllm_b200.parallel import state as ps
    st = ps.get_state()
    if st.tp_size > 1 and os.environ.get("GLLM_VISION_TP", "1") == "1" and heads % st.tp_size == 0 \
            and inter % st.tp_size == 0:
        return st.tp_rank, st.tp_size
    return 0, 1


class _ColLinear(nn.Linear):
    """Output features sharded over the TP group (`tp_kind` tells the weight loader how to slice)."""
    tp_kind = "col"

    def __init__(self, in_f, out_f, tp, bias=True):
        super().__init__(in_f, out_f // tp[1], bias=bias)
        self.tp = tp


class _RowLinear(nn.Linear):
    """Input features sharded; partial products are summed over the TP group, the (replicated) bias is added once."""
    tp_kind = "row"

    def __init__(self, in_f, out_f, tp, bias=True):
        super().__init__(in_f // tp[1], out_f, bias=bias)
        self.tp = tp

    def forward(self, x):
        if self.tp[1] == 1:
            return F.linear(x, self.weight, self.bias)     # unsharded: exactly nn.Linear (bias fused in the GEMM)
        from gllm_b200.parallel import state as ps
        y = ps.tp_all_reduce(F.linear(x, self.weight))
        return y if self.bias is None else y + self.bias


class VisionAttention(nn.Module):
    def __init__(self, dim: int, num_heads: int, tp=(0, 1)):
        super().__init__()
        self.num_heads = num_heads // tp[1]            # heads of this rank
        self.qkv = _ColLinear(dim, dim * 3, tp)
        self.qkv.tp_kind = "qkv"                       # rows are [3][heads][head_dim]: slice the heads axis
        self.qkv.total_heads = num_heads
        self.proj = _RowLinear(dim, dim, tp)

    def forward(self, x, cu_seqlens: List[int], cos, sin):
        s = x.shape[0]
        q, k, v = self.qkv(x).reshape(s, 3, self.num_heads, -1).permute(1, 0, 2, 3).unbind(0)
        q, k = _apply_rope(q, k, cos, sin)
        outs = []
        for a, b in zip(cu_seqlens[:-1], cu_seqlens[1:]):
            if b <= a:
                continue
            o = F.scaled_dot_product_attention(q[a:b].transpose(0, 1).unsqueeze(0), k[a:b].transpose(0, 1).unsqueeze(0),
                                               v[a:b].transpose(0, 1).unsqueeze(0))
            outs.append(o.squeeze(0).transpose(0, 1))
        return self.proj(torch.cat(outs, 0).reshape(s, -1))


class RMSNorm(nn.Module):
    def __init__(self, dim, eps=1e-6):
        super().__init__()
        self.weight = nn.Parameter(torch.ones(dim))
        self.eps = eps

    def forward(self, x):
        xf = x.float()
        return self.weight * (xf * torch.rsqrt(xf.pow(2).mean(-1, keepdim=True) + self.eps)).to(x.dtype)


def _rot_table(dim: int, n: int, device, theta: float = 10000.0):
    inv = 1.0 / (theta ** (torch.arange(0, dim, 2, dtype=torch.float, device=device) / dim))
    return torch.outer(torch.arange(n, dtype=torch.float, device=device), inv)


def _hw_pos_ids(grid, merge):
    """(row, col) of every patch in merge-block-major order, repeated over frames."""
    out = []
    for t, h, w in grid:
        r = torch.arange(h).unsqueeze(1).expand(-1, w).reshape(h // merge, merge, w // merge, merge)
        c = torch.arange(w).unsqueeze(0).expand(h, -1).reshape(h // merge, merge, w // merge, merge)
        rc = torch.stack([r.permute(0, 2, 1, 3).flatten(), c.permute(0, 2, 1, 3).flatten()], dim=-1)
        out.append(rc.repeat(t, 1))
    return torch.cat(out, 0)


def _frame_cu_seqlens(grid):
    cu = [0]
    for t, h, w in grid:
        for _ in range(t):
            cu.append(cu[-1] + h * w)
    return cu


# ------------------------------------------------------------------------------------------------
# Qwen2.5-VL
# ------------------------------------------------------------------------------------------------
class _PatchEmbed(nn.Module):
    def __init__(self, in_ch, tps, ps, dim, bias):
        super().__init__()
        self.shape = (in_ch, tps, ps, ps)
        self.proj = nn.Conv3d(in_ch, dim, kernel_size=(tps, ps, ps), stride=(tps, ps, ps), bias=bias)

    def forward(self, x):
        # a stride==kernel Conv3d over pre-cut patches is one GEMM
        w = self.proj.weight.view(self.proj.weight.shape[0], -1)
        return F.linear(x.view(x.shape[0], -1).to(w.dtype), w, self.proj.bias)


class _SwiGLU(nn.Module):
    def __init__(self, dim, inter, tp=(0, 1)):
        super().__init__()
        self.gate_proj = _ColLinear(dim, inter, tp)
        self.up_proj = _ColLinear(dim, inter, tp)
        self.down_proj = _RowLinear(inter, dim, tp)

    def forward(self, x):
        return self.down_proj(F.silu(self.gate_proj(x)) * self.up_proj(x))


class _Block25(nn.Module):
    def __init__(self, dim, heads, inter, tp=(0, 1)):
        super().__init__()
        self.norm1, self.norm2 = RMSNorm(dim), RMSNorm(dim)
        self.attn = VisionAttention(dim, heads, tp)
        self.mlp = _SwiGLU(dim, inter, tp)

    def forward(self, x, cu, cos, sin):
        x = x + self.attn(self.norm1(x), cu, cos, sin)
        return x + self.mlp(self.norm2(x))


class _Merger25(nn.Module):
    def __init__(self, out_dim, ctx_dim, merge):
        super().__init__()
        self.hidden = ctx_dim * merge * merge
        self.ln_q = RMSNorm(ctx_dim)
        self.mlp = nn.Sequential(nn.Linear(self.hidden, self.hidden), nn.GELU(), nn.Linear(self.hidden, out_dim))

    def forward(self, x):
        return self.mlp(self.ln_q(x).view(-1, self.hidden))


class Qwen2_5_VisionTower(nn.Module):
    def __init__(self, vc: dict, dtype, device):
        super().__init__()
        dim, heads = vc["hidden_size"], vc["num_heads"]
        self.merge = vc.get("spatial_merge_size", 2)
        self.patch_size = vc.get("patch_size", 14)
        self.window_size = vc.get("window_size", 112)
        self.fullatt = set(vc.get("fullatt_block_indexes", [7, 15, 23, 31]))
        self.head_dim = dim // heads
        self.patch_embed = _PatchEmbed(vc.get("in_channels", vc.get("in_chans", 3)), vc.get("temporal_patch_size", 2),
                                       self.patch_size, dim, bias=False)
        self.tp = vision_tp(heads, vc["intermediate_size"])
        self.blocks = nn.ModuleList([_Block25(dim, heads, vc["intermediate_size"], self.tp)
                                     for _ in range(vc["depth"])])
        self.merger = _Merger25(vc["out_hidden_size"], dim, self.merge)
        self.to(device=device, dtype=dtype)

    def _window_index(self, grid):
        """Permutation of merged-token indices into window-major order + cumulative window lengths."""
        idx, cu, base = [], [0], 0
        vw = self.window_size // self.merge // self.patch_size
        unit = self.merge * self.merge
        for t, h, w in grid:
            gh, gw = h // self.merge, w // self.merge
            index = torch.arange(t * gh * gw).reshape(t, gh, gw)
            ph, pw = vw - gh % vw, vw - gw % vw
            nh, nw = (gh + ph) // vw, (gw + pw) // vw
            pad = F.pad(index, (0, pw, 0, ph), "constant", -100)
            pad = pad.reshape(t, nh, vw, nw, vw).permute(0, 1, 3, 2, 4).reshape(t, nh * nw, vw, vw)
            lens = (pad != -100).sum([2, 3]).reshape(-1)
            flat = pad.reshape(-1)
            idx.append(flat[flat != -100] + base)
            for n in (lens.cumsum(0) * unit + cu[-1]).tolist():
                cu.append(int(n))
            base += t * gh * gw
        cu = [c for i, c in enumerate(cu) if i == 0 or c != cu[i - 1]]
        return torch.cat(idx, 0), cu

    def forward(self, pixel_values, grid_thw):
        grid = [tuple(int(v) for v in g) for g in grid_thw.tolist()]
        dev = pixel_values.device
        x = self.patch_embed(pixel_values)
        s = x.shape[0]
        unit = self.merge * self.merge
        pos = _hw_pos_ids(grid, self.merge).to(dev)
        table = _rot_table(self.head_dim // 2, max(max(h, w) for _, h, w in grid), dev)
        rot = table[pos].flatten(1)
        widx, cu_win = self._window_index(grid)
        widx = widx.to(dev)
        x = x.reshape(s // unit, unit, -1)[widx].reshape(s, -1)
        rot = rot.reshape(s // unit, unit, -1)[widx].reshape(s, -1)
        emb = torch.cat((rot, rot), dim=-1)
        cos, sin = emb.cos(), emb.sin()
        cu_full = _frame_cu_seqlens(grid)
        for i, blk in enumerate(self.blocks):
            x = blk(x, cu_full if i in self.fullatt else cu_win, cos, sin)
            if x.dtype == torch.float16:     # fp16 checkpoints can overflow in the ViT residual stream
                x = clamp_overflow(x)
        out = self.merger(x)
        return out[torch.argsort(widx)], []


# ------------------------------------------------------------------------------------------------
# Qwen3-VL
# ------------------------------------------------------------------------------------------------
class _MLP3(nn.Module):
    def __init__(self, dim, inter, act, tp=(0, 1)):
        super().__init__()
        self.linear_fc1 = _ColLinear(dim, inter, tp)
        self.linear_fc2 = _RowLinear(inter, dim, tp)
        self.tanh = "tanh" in act

    def forward(self, x):
        return self.linear_fc2(F.gelu(self.linear_fc1(x), approximate="tanh" if self.tanh else "none"))


class _Block3(nn.Module):
    def __init__(self, dim, heads, inter, act, tp=(0, 1)):
        super().__init__()
        self.norm1, self.norm2 = nn.LayerNorm(dim, eps=1e-6), nn.LayerNorm(dim, eps=1e-6)
        self.attn = VisionAttention(dim, heads, tp)
        self.mlp = _MLP3(dim, inter, act, tp)

    def forward(self, x, cu, cos, sin):
        x = x + self.attn(self.norm1(x), cu, cos, sin)
        return x + self.mlp(self.norm2(x))


class _Merger3(nn.Module):
    def __init__(self, dim, out_dim, merge, postshuffle):
        super().__init__()
        self.hidden = dim * merge * merge
        self.postshuffle = postshuffle
        self.norm = nn.LayerNorm(self.hidden if postshuffle else dim, eps=1e-6)
        self.linear_fc1 = nn.Linear(self.hidden, self.hidden)
        self.linear_fc2 = nn.Linear(self.hidden, out_dim)

    def forward(self, x):
        x = self.norm(x.view(-1, self.hidden) if self.postshuffle else x).view(-1, self.hidden)
        return self.linear_fc2(F.gelu(self.linear_fc1(x)))


class Qwen3VisionTower(nn.Module):
    def __init__(self, vc: dict, dtype, device):
        super().__init__()
        dim, heads = vc["hidden_size"], vc["num_heads"]
        self.merge = vc.get("spatial_merge_size", 2)
        self.head_dim = dim // heads
        self.patch_embed = _PatchEmbed(vc.get("in_channels", 3), vc.get("temporal_patch_size", 2),
                                       vc.get("patch_size", 16), dim, bias=True)
        self.pos_embed = nn.Embedding(vc["num_position_embeddings"], dim)
        self.side = int(vc["num_position_embeddings"] ** 0.5)
        act = vc.get("hidden_act", "gelu_pytorch_tanh")
        self.tp = vision_tp(heads, vc["intermediate_size"])
        self.blocks = nn.ModuleList([_Block3(dim, heads, vc["intermediate_size"], act, self.tp)
                                     for _ in range(vc["depth"])])
        self.merger = _Merger3(dim, vc["out_hidden_size"], self.merge, False)
        self.deepstack_idx = list(vc.get("deepstack_visual_indexes", []) or [])
        self.deepstack_merger_list = nn.ModuleList(
            [_Merger3(dim, vc["out_hidden_size"], self.merge, True) for _ in self.deepstack_idx])
        self.to(device=device, dtype=dtype)

    def _pos_embed(self, grid, dev):
        """Bilinear interpolation of the learned side x side position table to every (h, w) grid."""
        outs = []
        n, m = self.side, self.merge
        wt = self.pos_embed.weight
        for t, h, w in grid:
            hi = torch.linspace(0, n - 1, h, device=dev)
            wi = torch.linspace(0, n - 1, w, device=dev)
            h0, w0 = hi.int(), wi.int()
            h1, w1 = (h0 + 1).clip(max=n - 1), (w0 + 1).clip(max=n - 1)
            dh, dw = (hi - h0).unsqueeze(1), (wi - w0).unsqueeze(0)
            acc = 0
            for hh, ww, wgt in ((h0, w0, (1 - dh) * (1 - dw)), (h0, w1, (1 - dh) * dw),
                                (h1, w0, dh * (1 - dw)), (h1, w1, dh * dw)):
                idx = (hh.long().unsqueeze(1) * n + ww.long().unsqueeze(0)).flatten()
                acc = acc + wt[idx] * wgt.flatten().unsqueeze(1).to(wt.dtype)
            pe = acc.repeat(t, 1).view(t, h // m, m, w // m, m, -1).permute(0, 1, 3, 2, 4, 5).flatten(0, 4)
            outs.append(pe)
        return torch.cat(outs, 0)

    def forward(self, pixel_values, grid_thw):
        grid = [tuple(int(v) for v in g) for g in grid_thw.tolist()]
        dev = pixel_values.device
        x = self.patch_embed(pixel_values)
        x = x + self._pos_embed(grid, dev).to(x.dtype)
        pos = _hw_pos_ids(grid, self.merge).to(dev)
        table = _rot_table(self.head_dim // 2, max(max(h, w) for _, h, w in grid), dev)
        rot = table[pos].flatten(1)
        emb = torch.cat((rot, rot), dim=-1)
        cos, sin = emb.cos(), emb.sin()
        cu = _frame_cu_seqlens(grid)
        deep = []
        for i, blk in enumerate(self.blocks):
            x = blk(x, cu, cos, sin)
            if i in self.deepstack_idx:
                deep.append(self.deepstack_merger_list[self.deepstack_idx.index(i)](x))
        return self.merger(x), deep


def shard_vision_param(mod: nn.Module, pname: str, full: torch.Tensor) -> torch.Tensor:
    """This rank's slice of a full checkpoint tensor for parameter `pname` ("weight" / "bias") of `mod`."""
    kind = getattr(mod, "tp_kind", None)
    rank, size = getattr(mod, "tp", (0, 1))
    if kind is None or size == 1:
        return full
    if kind == "col":
        n = full.shape[0] // size
        return full[rank * n:(rank + 1) * n]
    if kind == "row":
        if pname == "bias":
            return full
        n = full.shape[1] // size
        return full[:, rank * n:(rank + 1) * n]
    if kind == "qkv":      # rows are [3][heads][head_dim]
        heads = mod.total_heads
        hl = heads // size
        v = full.view(3, heads, full.shape[0] // (3 * heads), *full.shape[1:])
        return v[:, rank * hl:(rank + 1) * hl].reshape(-1, *full.shape[1:])
    raise ValueError(kind)


def load_vision_weights(tower: nn.Module, reader, prefix: str = "visual."):
    for mname, mod in tower.named_modules():
        for pname, p in mod.named_parameters(recurse=False):
            name = f"{mname}.{pname}" if mname else pname
            p.data.copy_(shard_vision_param(mod, pname, reader.get(prefix + name)).to(p.dtype))
