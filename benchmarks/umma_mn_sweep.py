"""Sweep the shared-memory descriptor encodings for an MN-major (row-major [K, N]) B operand of tcgen05.mma and
report which ones reproduce torch (csrc/bench/umma_mn_probe.cu). One GPU, a few hundred tiny launches; every
launch is followed by a synchronize so an illegal encoding surfaces as a CUDA error for that combination only
(run under `timeout`: a bad descriptor can also hang the tensor pipe).

    timeout 120 python benchmarks/umma_mn_sweep.py
"""
import itertools
import os
import sys

import torch

sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
from gllm_b200.ops import lib as _lib  # noqa: E402


def main():
    L = _lib.load()
    P, I = _lib.c_void_p, _lib.c_int
    L.gllm_umma_mn_probe.argtypes = [P, P, P, I, I, I, I, I, I, P]
    L.gllm_umma_mn_probe.restype = I
    torch.manual_seed(0)
    a = (torch.randn(128, 64, device="cuda") * 0.5).bfloat16()
    bt = (torch.randn(64, 128, device="cuda") * 0.5).bfloat16()
    want = a.float() @ bt.float()
    d = torch.empty(128, 128, device="cuda", dtype=torch.float32)
    st = torch.cuda.current_stream().cuda_stream
    hits = []
    lbos = [0, 1, 8, 64, 128, 512, 1024]          # x16 B: 0, 16, 128, 1 KB, 2 KB, 8 KB, 16 KB
    sbos = [8, 64, 128, 512]                       # 128 B, 1 KB, 2 KB, 8 KB
    kadvs = [2, 8, 64, 128]                        # 32 B, 128 B, 1 KB (8 rows), 2 KB (16 rows of 128 B)
    for b_major, split_n, lbo, sbo, kadv in itertools.product((1, 0), (1, 0), lbos, sbos, kadvs):
        d.zero_()
        rc = L.gllm_umma_mn_probe(a.data_ptr(), bt.data_ptr(), d.data_ptr(), lbo, sbo, kadv, 8192, b_major, split_n, st)
        try:
            torch.cuda.synchronize()
        except RuntimeError as e:  # noqa: BLE001
            print("CUDA error at", dict(b_major=b_major, split_n=split_n, lbo16=lbo, sbo16=sbo, k_adv16=kadv), e)
            return
        if rc == 0:
            err = ((d - want).norm() / want.norm()).item()
            if err < 2e-2:
                hits.append((b_major, split_n, lbo, sbo, kadv, err))
                print("MATCH b_major=%d split_n=%d lbo16=%d sbo16=%d k_adv16=%d rel_err=%.4f" % hits[-1], flush=True)
    print(f"{len(hits)} matching encodings")


if __name__ == "__main__":
    main()
