"""Per-op device-time breakdown of one engine forward (CUDA events around every kernel front-end).
    python benchmarks/step_breakdown.py --model preset:qwen3-8b --batch 256 --ctx 512 [--prefill 0]
"""
import argparse
import collections
import os
import sys

import numpy as np
import torch

sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
from gllm_b200.config import EngineConfig  # noqa: E402
from gllm_b200.input_data import BatchArrays  # noqa: E402
from gllm_b200.model_runner import ModelRunner  # noqa: E402
from gllm_b200.ops import sm100  # noqa: E402
from gllm_b200.parallel import state as ps  # noqa: E402

REC = []


def wrap(name):
    fn = getattr(sm100, name)

    def inner(*a, **kw):
        e0, e1 = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
        e0.record()
        out = fn(*a, **kw)
        e1.record()
        shape = tuple(a[0].shape) if len(a) and hasattr(a[0], "shape") else ()
        extra = tuple(a[1].shape) if name.startswith("linear") else ()
        REC.append((name, shape, extra, e0, e1))
        return out
    setattr(sm100, name, inner)


def make_batch(b, ctx, prefill, page):
    if prefill:
        q = np.full(b, prefill, np.int32)
        sl = q.copy()
    else:
        q = np.ones(b, np.int32)
        sl = np.full(b, ctx, np.int32)
    qsl = np.zeros(b + 1, np.int32)
    np.cumsum(q, out=qsl[1:])
    t = int(qsl[-1])
    nb = (int(sl.max()) + page - 1) // page
    bt = (np.arange(b * nb, dtype=np.int32).reshape(b, nb)) % 60000
    pos = np.concatenate([np.arange(s - n, s, dtype=np.int32) for s, n in zip(sl, q)])
    slots = np.concatenate([bt[i, (np.arange(s - n, s) // page)] * page + np.arange(s - n, s) % page
                            for i, (s, n) in enumerate(zip(sl, q))]).astype(np.int32)
    e = b
    return BatchArrays(tokens=np.random.randint(0, 1000, t).astype(np.int32), positions=pos, slot_mapping=slots,
                       block_table=bt, seq_lens=sl, query_start_loc=qsl, logits_idx=(qsl[1:] - 1).astype(np.int32),
                       emit_seq=np.arange(e, dtype=np.int32), temperature=np.ones(e, np.float32),
                       top_k=np.ones(e, np.int32), top_p=np.ones(e, np.float32), rep_penalty=np.ones(e, np.float32),
                       state_slot=np.zeros(e, np.int32), num_decode_seqs=0 if prefill else b, num_seqs=b,
                       num_tokens=t, max_q_len=int(q.max()), max_seq_len=int(sl.max()))


def main():
    ap = argparse.ArgumentParser()
    ap.add_argument("--model", default="preset:qwen3-8b")
    ap.add_argument("--batch", type=int, default=256)
    ap.add_argument("--ctx", type=int, default=512)
    ap.add_argument("--prefill", type=int, default=0, help="tokens per sequence (0 = decode)")
    args = ap.parse_args()
    ps.init_dist(1, 1, 0, 0)
    cfg = EngineConfig(model_path=args.model, load_format="dummy", maxp=8192, maxd=1024, disable_cuda_graph=True,
                       num_gpu_pages=65536, model_max_length=4096)
    r = ModelRunner(cfg)
    r.init("cuda:0")
    batch = make_batch(args.batch, args.ctx, args.prefill, cfg.page_size)
    for _ in range(3):
        r.step(batch)
    torch.cuda.synchronize()
    for n in ("linear", "linear_silu_mul", "rmsnorm", "rope_kv_write", "paged_attention", "embedding", "gather_rows",
              "sample"):
        wrap(n)
    e0, e1 = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
    e0.record()
    r.step(batch)
    e1.record()
    torch.cuda.synchronize()
    agg = collections.OrderedDict()
    for name, shape, extra, a, b in REC:
        k = (name, shape, extra)
        agg.setdefault(k, [0, 0.0])
        agg[k][0] += 1
        agg[k][1] += a.elapsed_time(b)
    tot = e0.elapsed_time(e1)
    print(f"forward total {tot:.3f} ms (batch={args.batch} ctx={args.ctx} prefill={args.prefill})")
    s = 0.0
    for (name, shape, extra), (cnt, ms) in sorted(agg.items(), key=lambda kv: -kv[1][1]):
        s += ms
        print(f"{ms:8.3f} ms {100 * ms / tot:5.1f}%  x{cnt:<4d} {ms / cnt * 1e3:8.1f} us/call  {name} {shape} {extra}")
    print(f"sum of ops {s:.3f} ms; gaps/launch overhead {tot - s:.3f} ms")


if __name__ == "__main__":
    main()
