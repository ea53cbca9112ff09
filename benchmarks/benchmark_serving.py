"""Online serving benchmark (reference: benchmarks/benchmark_serving.py): Poisson / gamma arrivals against a
running OpenAI-compatible server; reports TTFT / TPOT / ITL / E2EL mean, median, std, percentiles, output
throughput and goodput.

    python -m gllm_b200.entrypoints.api_server --model-path preset:qwen3-8b --load-format dummy --tp 8 &
    python benchmarks/benchmark_serving.py --num-prompts 1000 --request-rate 32
Prompts are sent as token-id arrays (synthetic ShareGPT-shaped), so no tokenizer/dataset is needed.
"""
import argparse
import asyncio
import json
import os
import sys
import time

import numpy as np

sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
sys.path.insert(0, os.path.dirname(os.path.abspath(__file__)))
from backend_request_func import ASYNC_REQUEST_FUNCS, RequestFuncInput  # noqa: E402
from workloads import sharegpt_shaped  # noqa: E402


async def arrival_times(n, rate, burstiness, rng):
    """Inter-arrival gaps: gamma(shape=burstiness, scale=1/(rate*burstiness)); burstiness=1 is Poisson."""
    for i in range(n):
        yield i
        if rate == float("inf"):
            continue
        await asyncio.sleep(float(rng.gamma(shape=burstiness, scale=1.0 / (rate * burstiness))))


def summarise(outs, duration, percentiles, goodput):
    ok = [o for o in outs if o.success]
    ttft = np.array([o.ttft for o in ok]) * 1e3
    e2el = np.array([o.latency for o in ok]) * 1e3
    tpot = np.array([(o.latency - o.ttft) / (o.output_tokens - 1) for o in ok if o.output_tokens > 1]) * 1e3
    itl = np.array([x for o in ok for x in o.itl]) * 1e3
    res = {"completed": len(ok), "failed": len(outs) - len(ok), "duration_s": round(duration, 3),
           "total_input_tokens": int(sum(o.prompt_len for o in ok)),
           "total_output_tokens": int(sum(o.output_tokens for o in ok)),
           "request_throughput": round(len(ok) / duration, 3),
           "output_throughput": round(sum(o.output_tokens for o in ok) / duration, 2),
           "total_token_throughput": round(sum(o.output_tokens + o.prompt_len for o in ok) / duration, 2)}
    for name, arr in (("ttft", ttft), ("tpot", tpot), ("itl", itl), ("e2el", e2el)):
        if arr.size == 0:
            continue
        res[f"mean_{name}_ms"] = round(float(arr.mean()), 3)
        res[f"median_{name}_ms"] = round(float(np.median(arr)), 3)
        res[f"std_{name}_ms"] = round(float(arr.std()), 3)
        for p in percentiles:
            res[f"p{int(p)}_{name}_ms"] = round(float(np.percentile(arr, p)), 3)
    if goodput:
        good = 0
        for o in ok:
            vals = {"ttft": o.ttft * 1e3, "e2el": o.latency * 1e3,
                    "tpot": (o.latency - o.ttft) / max(o.output_tokens - 1, 1) * 1e3}
            if all(vals[k] <= v for k, v in goodput.items()):
                good += 1
        res["request_goodput"] = round(good / duration, 3)
    return res


async def run(args):
    import aiohttp
    from tqdm import tqdm
    base = f"http://{args.host}:{args.port}"
    api_url = base + args.endpoint
    async with aiohttp.ClientSession() as s:
        async with s.get(base + "/v1/models") as r:
            info = await r.json()
    model = args.model or info["data"][0]["id"]
    vocab = args.vocab_size
    prompts, outs = sharegpt_shaped(args.num_prompts, vocab, args.seed, max_output=args.max_output_len)
    fn = ASYNC_REQUEST_FUNCS[args.backend]
    rng = np.random.default_rng(args.seed)
    if args.profile:
        async with aiohttp.ClientSession() as s:
            await s.post(base + "/start_profile")
    pbar = tqdm(total=len(prompts))
    sem = asyncio.Semaphore(args.max_concurrency) if args.max_concurrency else None

    async def one(i):
        inp = RequestFuncInput(prompt=prompts[i], api_url=api_url, prompt_len=len(prompts[i]), output_len=outs[i],
                               model=model, ignore_eos=True)
        if sem is None:
            return await fn(inp, pbar)
        async with sem:
            return await fn(inp, pbar)

    if args.prompt_format == "words":
        # servers that only take text (the reference's /v1/completions tokenises `prompt`): the benchmark model
        # directories carry a one-word-per-id vocabulary ("t<id>"), so the same token ids go over the wire as words
        prompts = [" ".join(f"t{t}" for t in p) for p in prompts]
    t0 = time.perf_counter()
    tasks = []
    # staged arrivals (reference: benchmarks/benchmark_serving.py:686-718): the request set is cut into
    # `--arrival-stage` equal parts, each sent with its own Poisson process, `--stage-interval` seconds apart
    n_stage = max(1, args.arrival_stage)
    per = len(prompts) // n_stage
    for st in range(n_stage):
        lo, hi = st * per, (st + 1) * per if st != n_stage - 1 else len(prompts)
        async for i in arrival_times(hi - lo, args.request_rate, args.burstiness, rng):
            tasks.append(asyncio.create_task(one(lo + i)))
        if n_stage != 1:
            await asyncio.sleep(args.stage_interval)
    results = await asyncio.gather(*tasks)
    dur = time.perf_counter() - t0
    pbar.close()
    if args.profile:
        async with aiohttp.ClientSession() as s:
            await s.post(base + "/stop_profile")
    goodput = None
    if args.goodput:
        goodput = {kv.split(":")[0]: float(kv.split(":")[1]) for kv in args.goodput}
    res = summarise(results, dur, [float(p) for p in args.metric_percentiles.split(",")], goodput)
    res.update({"backend": args.backend, "request_rate": args.request_rate, "num_prompts": args.num_prompts,
                "arrival_stage": args.arrival_stage, "stage_interval": args.stage_interval})
    print("{s:=^50}".format(s=" Serving Benchmark Result "))
    for k, v in res.items():
        print(f"{k:<32}{v}")
    errs = [o.error for o in results if not o.success][:3]
    if errs:
        print("sample errors:", errs)
    if args.save_result:
        with open(args.save_result, "w") as f:
            json.dump(res, f, indent=2)
    return res


def main():
    ap = argparse.ArgumentParser()
    ap.add_argument("--backend", default="gllm_b200", choices=list(ASYNC_REQUEST_FUNCS))
    ap.add_argument("--host", default="127.0.0.1")
    ap.add_argument("--port", type=int, default=8000)
    ap.add_argument("--endpoint", default="/v1/completions")
    ap.add_argument("--model", default=None)
    ap.add_argument("--num-prompts", type=int, default=1000)
    ap.add_argument("--request-rate", type=float, default=float("inf"))
    ap.add_argument("--burstiness", type=float, default=1.0)
    ap.add_argument("--max-concurrency", type=int, default=None)
    ap.add_argument("--max-output-len", type=int, default=512)
    ap.add_argument("--vocab-size", type=int, default=150000)
    ap.add_argument("--seed", type=int, default=0)
    ap.add_argument("--prompt-format", default="ids", choices=["ids", "words"],
                    help="send prompts as token-id lists, or as 't<id>' words for text-only servers")
    ap.add_argument("--arrival-stage", type=int, default=1, help="number of stages the requests are sent in")
    ap.add_argument("--stage-interval", type=float, default=10.0, help="seconds between stages")
    ap.add_argument("--metric-percentiles", default="50,90,99")
    ap.add_argument("--goodput", nargs="*", default=None, help="SLOs like ttft:500 tpot:50 e2el:10000 (ms)")
    ap.add_argument("--profile", action="store_true")
    ap.add_argument("--save-result", default=None)
    asyncio.run(run(ap.parse_args()))


if __name__ == "__main__":
    main()
