"""Host-side cost of one engine iteration (no GPU, no model): scheduler + batch assembly + wire packing.

The driver's Python work per step is serial with the GPU unless async scheduling is on, so every 0.1 ms here is
~1 % of a decode step. Drives the real Scheduler / MemoryManager / build_batch with an instant "model".

    python benchmarks/host_loop_bench.py [--seqs 256] [--prompt 200] [--out 200] [--profile]
"""
import argparse
import cProfile
import os
import pstats
import sys
import time

import numpy as np

sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
from gllm_b200.input_data import BatchArrays, build_batch  # noqa: E402
from gllm_b200.memory_manager import PrefixMemoryManager  # noqa: E402
from gllm_b200.scheduler import Scheduler  # noqa: E402
from gllm_b200.sequence import Sequence  # noqa: E402


def run(args):
    page = 16
    pages = args.seqs * ((args.prompt + args.out) // page + 2) + 64
    mm = PrefixMemoryManager(pages, page)
    sch = Scheduler(mm, pp_size=1, world_size=1, schedule_method=args.method, maxd=args.seqs, maxp=4096,
                    page_size=page, log=False)
    rng = np.random.default_rng(0)
    seqs = [Sequence(i, rng.integers(0, 30000, args.prompt).tolist(), [2], output_len=args.out, ignore_eos=True)
            for i in range(args.seqs)]
    sch.add_new_requests(seqs)
    t = {"schedule": 0.0, "build": 0.0, "wire": 0.0, "output": 0.0}
    steps = decode_steps = 0
    prev = None
    bid = 0
    while sch.has_work():
        t0 = time.perf_counter()
        entries = sch.schedule_once()
        t1 = time.perf_counter()
        if not entries:
            break
        bid += 1
        batch = build_batch(entries, page, 32000, bid, prev=prev)
        prev = batch
        t2 = time.perf_counter()
        if args.wire:
            hdr, bufs = batch.to_wire()
            BatchArrays.from_wire(hdr, [memoryview(b) for b in bufs])
        t3 = time.perf_counter()
        sch.add_next_tokens([7] * sum(1 for e in entries if e.emits))
        while sch.process_output() is not None:
            pass
        t4 = time.perf_counter()
        steps += 1
        if batch.is_decode_only():
            decode_steps += 1
            t["schedule"] += t1 - t0
            t["build"] += t2 - t1
            t["wire"] += t3 - t2
            t["output"] += t4 - t3
    d = max(decode_steps, 1)
    per = {k: round(v / d * 1e3, 4) for k, v in t.items()}
    per["total_ms_per_decode_step"] = round(sum(t.values()) / d * 1e3, 4)
    print({"seqs": args.seqs, "steps": steps, "decode_steps": decode_steps, **per})


if __name__ == "__main__":
    ap = argparse.ArgumentParser()
    ap.add_argument("--seqs", type=int, default=256)
    ap.add_argument("--prompt", type=int, default=200)
    ap.add_argument("--out", type=int, default=200)
    ap.add_argument("--method", default="chunked_prefill")
    ap.add_argument("--wire", action="store_true", help="include the TP fan-out packing / unpacking")
    ap.add_argument("--profile", action="store_true")
    a = ap.parse_args()
    if a.profile:
        pr = cProfile.Profile()
        pr.enable()
        run(a)
        pr.disable()
        pstats.Stats(pr).sort_stats("tottime").print_stats(18)
    else:
        run(a)
