"""Prefix-cache serving benchmark (reference: benchmarks/benchmark_prefix_serving.py): every user holds a
multi-round conversation (rounds are sequential per user, users are concurrent); all users share a system
prompt, and each round re-sends the whole history — so a prefix cache turns most prefill into cache hits.

    python benchmarks/benchmark_prefix_serving.py --num-users 32 --rounds 6
Run the server with and without --enable-prefix-caching and compare TTFT / cache hit rate (/metrics).
"""
import argparse
import asyncio
import os
import sys
import time

import numpy as np

sys.path.insert(0, os.path.dirname(os.path.abspath(__file__)))
from backend_request_func import RequestFuncInput, async_request_openai_completions  # noqa: E402
from workloads import multi_round_conversations  # noqa: E402


async def user_session(uid, system, turns, answer_len, api_url, vocab, rng, results):
    history = list(system)
    for turn in turns:
        history = history + turn
        inp = RequestFuncInput(prompt=history, api_url=api_url, prompt_len=len(history), output_len=answer_len)
        out = await async_request_openai_completions(inp)
        results.append(out)
        # the synthetic "assistant answer" joins the history (ids only; the text is irrelevant here)
        history = history + rng.integers(10, vocab - 10, size=answer_len).tolist()


async def run(args):
    import aiohttp
    base = f"http://{args.host}:{args.port}"
    system, users, answer_len = multi_round_conversations(args.num_users, args.rounds, args.vocab_size, args.seed,
                                                          args.system_len, args.turn_len, args.answer_len)
    rng = np.random.default_rng(args.seed + 1)
    results = []
    t0 = time.perf_counter()
    await asyncio.gather(*[user_session(i, system, u, answer_len, base + "/v1/completions", args.vocab_size, rng,
                                        results) for i, u in enumerate(users)])
    dur = time.perf_counter() - t0
    ok = [r for r in results if r.success]
    ttft = np.array([r.ttft for r in ok]) * 1e3
    print(f"requests {len(ok)}/{len(results)}  duration {dur:.2f}s  mean TTFT {ttft.mean():.1f} ms  "
          f"median TTFT {np.median(ttft):.1f} ms  p99 TTFT {np.percentile(ttft, 99):.1f} ms  "
          f"output tok/s {sum(r.output_tokens for r in ok) / dur:.1f}")
    async with aiohttp.ClientSession() as s:
        async with s.get(base + "/metrics") as r:
            for line in (await r.text()).splitlines():
                if "cache_hit_rate" in line and not line.startswith("#"):
                    print(line)


def main():
    ap = argparse.ArgumentParser()
    ap.add_argument("--host", default="127.0.0.1")
    ap.add_argument("--port", type=int, default=8000)
    ap.add_argument("--num-users", type=int, default=32)
    ap.add_argument("--rounds", type=int, default=6)
    ap.add_argument("--system-len", type=int, default=256)
    ap.add_argument("--turn-len", type=int, default=64)
    ap.add_argument("--answer-len", type=int, default=64)
    ap.add_argument("--vocab-size", type=int, default=150000)
    ap.add_argument("--seed", type=int, default=0)
    asyncio.run(run(ap.parse_args()))


if __name__ == "__main__":
    main()
