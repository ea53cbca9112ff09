"""Driver -> peers batch fan-out latency (host only): how long from `send_batch` on the driver until EVERY peer has
the decoded BatchArrays in hand. This sits on the critical path of every TP/PP step (the peers cannot launch
their forward before it). Compares the ZeroMQ ipc transport with the shared-memory ring (engine/shm_ring.py).

    python benchmarks/fanout_bench.py [--peers 7] [--seqs 64] [--iters 2000] [--transport zmq|shm]
"""
import argparse
import multiprocessing as mp
import os
import sys
import time

import numpy as np

sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))


def make_batch(b, ctx=512, page=16):
    from gllm_b200.input_data import BatchArrays
    nb = ctx // page
    i32 = np.int32
    return BatchArrays(tokens=np.arange(b, dtype=i32), positions=np.full(b, ctx - 1, i32),
                       slot_mapping=np.arange(b, dtype=i32), block_table=np.arange(b * nb, dtype=i32).reshape(b, nb),
                       seq_lens=np.full(b, ctx, i32), query_start_loc=np.arange(b + 1, dtype=i32),
                       logits_idx=np.arange(b, dtype=i32), emit_seq=np.arange(b, dtype=i32),
                       temperature=np.ones(b, np.float32), top_k=np.ones(b, i32), top_p=np.ones(b, np.float32),
                       rep_penalty=np.ones(b, np.float32), state_slot=np.zeros(b, i32), num_decode_seqs=b, num_seqs=b,
                       num_tokens=b, max_q_len=1, max_seq_len=ctx)


def peer(rank, world, base, transport, iters, stamps, ready):
    from gllm_b200.engine.comm import Comm
    os.environ["GLLM_BATCH_TRANSPORT"] = transport
    c = Comm(base, rank, world, 0).init()
    ready[rank - 1] = 1
    n = 0
    while n < iters:
        msg = c.recv_batch(0)
        if msg is None:
            continue
        stamps[(rank - 1) * iters + n] = time.perf_counter_ns()
        n += 1
    c.close()


def main():
    ap = argparse.ArgumentParser()
    ap.add_argument("--peers", type=int, default=7)
    ap.add_argument("--seqs", type=int, default=64)
    ap.add_argument("--iters", type=int, default=2000)
    ap.add_argument("--transport", default="zmq")
    a = ap.parse_args()
    from gllm_b200.engine.comm import Comm, ipc_base
    os.environ["GLLM_BATCH_TRANSPORT"] = a.transport
    base = ipc_base()
    world = a.peers + 1
    ctx = mp.get_context("spawn")
    stamps = ctx.Array("q", a.peers * a.iters, lock=False)
    ready = ctx.Array("i", a.peers, lock=False)
    procs = [ctx.Process(target=peer, args=(r, world, base, a.transport, a.iters, stamps, ready), daemon=True)
             for r in range(1, world)]
    for p in procs:
        p.start()
    while not all(ready):                 # peers have bound their inboxes
        time.sleep(0.05)
    drv = Comm(base, 0, world, 0).init()
    time.sleep(0.5)
    batch = make_batch(a.seqs)
    sent = np.zeros(a.iters, dtype=np.int64)
    send_cost = np.zeros(a.iters, dtype=np.int64)
    for i in range(a.iters):
        t0 = time.perf_counter_ns()
        sent[i] = t0
        drv.send_batch(batch)
        send_cost[i] = time.perf_counter_ns() - t0
        time.sleep(0.0005)                # one decode step apart
    for p in procs:
        p.join(timeout=30)
    arr = np.frombuffer(stamps, dtype=np.int64).reshape(a.peers, a.iters)
    last = arr.max(axis=0) - sent         # until the slowest peer has it
    w = a.iters // 10
    print({"transport": a.transport, "peers": a.peers, "seqs": a.seqs,
           "driver_send_us_p50": round(float(np.median(send_cost[w:])) / 1e3, 1),
           "all_peers_have_it_us_p50": round(float(np.median(last[w:])) / 1e3, 1),
           "all_peers_have_it_us_p99": round(float(np.percentile(last[w:], 99)) / 1e3, 1)})
    drv.close(unlink_all=True)


if __name__ == "__main__":
    main()
