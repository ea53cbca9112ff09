"""One decoder layer's kernels at a decode batch size, launched once each after a warm-up pass — the target of
`ncu --metrics gpu__time_duration.sum,dram__bytes_read.sum ...` (per-kernel device time without launch gaps).

    ncu --clock-control none --metrics gpu__time_duration.sum,dram__bytes_read.sum,dram__bytes_write.sum \
        --csv --log-file gpurun_out/layer.csv python benchmarks/ncu_decode_layer.py --batch 64 --ctx 512
"""
import argparse
import os
import sys

import torch

sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
from gllm_b200.ops import ref, sm100  # noqa: E402


def main():
    ap = argparse.ArgumentParser()
    ap.add_argument("--batch", type=int, default=64)
    ap.add_argument("--ctx", type=int, default=512)
    ap.add_argument("--passes", type=int, default=2)
    a = ap.parse_args()
    dev = "cuda"
    H, I, hq, hkv, d, page = 4096, 12288, 32, 8, 128, 16
    b = a.batch
    torch.manual_seed(0)
    x = (torch.randn(b, H, device=dev) * 0.1).bfloat16()
    res = (torch.randn(b, H, device=dev) * 0.1).bfloat16()
    w_qkv = (torch.randn((hq + 2 * hkv) * d, H, device=dev) * 0.02).bfloat16()
    w_o = (torch.randn(H, hq * d, device=dev) * 0.02).bfloat16()
    w_gu = (torch.randn(2 * I, H, device=dev) * 0.02).bfloat16()
    w_dn = (torch.randn(H, I, device=dev) * 0.02).bfloat16()
    nw = torch.ones(H, device=dev).bfloat16()
    qn = torch.ones(d, device=dev).bfloat16()
    pages_per_seq = (a.ctx + page - 1) // page
    n_pages = b * pages_per_seq + 1
    shape = ref.kv_cache_shape(n_pages, hkv, d, page)
    kc = (torch.randn(shape, device=dev) * 0.1).bfloat16()
    vc = (torch.randn(shape, device=dev) * 0.1).bfloat16()
    bt = torch.arange(b * pages_per_seq, device=dev, dtype=torch.int32).view(b, pages_per_seq)
    seq_lens = torch.full((b,), a.ctx, device=dev, dtype=torch.int32)
    qsl = torch.arange(b + 1, device=dev, dtype=torch.int32)
    pos = torch.full((b,), a.ctx - 1, device=dev, dtype=torch.int32)
    slots = (bt[:, -1] * page + (a.ctx - 1) % page).to(torch.int32)
    cs = ref.build_cos_sin_cache(d, 4096, 1e6).to(dev)
    for _ in range(a.passes):
        h, r = sm100.rmsnorm(x, nw, 1e-6, res)
        qkv = sm100.linear(h, w_qkv)
        q = qkv[:, : hq * d].view(b, hq, d)
        k = qkv[:, hq * d: (hq + hkv) * d].view(b, hkv, d)
        v = qkv[:, (hq + hkv) * d:].view(b, hkv, d)
        sm100.rope_kv_write(q, k, v, pos, cs, d, True, qn, qn, 1e-6, kc, vc, slots)
        o = sm100.paged_attention(qkv[:, : hq * d], kc, vc, bt, seq_lens, qsl, d ** -0.5, hq, d, b, b, 1, a.ctx)
        y = sm100.linear(o, w_o)
        h2, r = sm100.rmsnorm(y, nw, 1e-6, r)
        act = sm100.linear_silu_mul(h2, w_gu)
        y2 = sm100.linear(act, w_dn)
    torch.cuda.synchronize()


if __name__ == "__main__":
    main()
