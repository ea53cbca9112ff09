"""Run a few GEMM shapes once each (for ncu captures). usage: prof_gemm.py M N K [reps]"""
import os
import sys

import torch

sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
from gllm_b200.ops import sm100  # noqa: E402

m, n, k = (int(x) for x in sys.argv[1:4])
reps = int(sys.argv[4]) if len(sys.argv) > 4 else 3
x = (torch.randn(m, k, device="cuda") * 0.1).bfloat16()
w = (torch.randn(n, k, device="cuda") * 0.1).bfloat16()
out = torch.empty(m, n, device="cuda", dtype=torch.bfloat16)
flush = torch.empty(256 << 20, dtype=torch.uint8, device="cuda")
for _ in range(reps):
    flush.zero_()
    sm100.linear(x, w, out=out)
torch.cuda.synchronize()
