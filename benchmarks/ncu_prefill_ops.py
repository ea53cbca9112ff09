"""Prefill-shaped hot kernels, launched a few times each — target of `ncu --set full` (Qwen3-8B shapes, 4096 tokens).

    ncu --set full --clock-control none --import-source on -k regex:'gemm_bf16_kernel|attn_prefill' -s 3 -c 3 \
        -o gpurun_out/prefill python benchmarks/ncu_prefill_ops.py
"""
import os
import sys

import torch

sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
from gllm_b200.ops import ref, sm100  # noqa: E402


def main():
    dev = "cuda"
    H, I, hq, hkv, d, page = 4096, 12288, 32, 8, 128, 16
    nseq, qlen = 4, 1024
    t = nseq * qlen
    torch.manual_seed(0)
    x = (torch.randn(t, H, device=dev) * 0.1).bfloat16()
    w_gu = (torch.randn(2 * I, H, device=dev) * 0.02).bfloat16()
    w_o = (torch.randn(H, hq * d, device=dev) * 0.02).bfloat16()
    q = (torch.randn(t, hq * d, device=dev) * 0.1).bfloat16()
    pages_per_seq = qlen // page
    shape = ref.kv_cache_shape(nseq * pages_per_seq + 1, hkv, d, page)
    kc = (torch.randn(shape, device=dev) * 0.1).bfloat16()
    vc = (torch.randn(shape, device=dev) * 0.1).bfloat16()
    bt = torch.arange(nseq * pages_per_seq, device=dev, dtype=torch.int32).view(nseq, pages_per_seq)
    seq_lens = torch.full((nseq,), qlen, device=dev, dtype=torch.int32)
    qsl = (torch.arange(nseq + 1, device=dev, dtype=torch.int32) * qlen).to(torch.int32)
    for _ in range(2):
        act = sm100.linear_silu_mul(x, w_gu)                      # gemm_bf16_kernel<256, 1>
        y = sm100.linear(q, w_o)                                  # gemm_bf16_kernel<*, 0>
        o = sm100.paged_attention(q, kc, vc, bt, seq_lens, qsl, d ** -0.5, hq, d, 0, nseq, qlen, qlen)
    torch.cuda.synchronize()


if __name__ == "__main__":
    main()
