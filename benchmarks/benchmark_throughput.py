"""Offline throughput benchmark (reference: benchmarks/benchmark_throughput.py): all requests are handed
to the engine at t = 0; reports requests/s, total tokens/s and output tokens/s.

    python benchmarks/benchmark_throughput.py --model-path preset:qwen3-8b --load-format dummy \
        --num-prompts 1000 [--tp 8] [--dataset ShareGPT_V3.json]
Under torchrun (one rank per GPU) the engine runs in-process on every rank; otherwise it spawns workers.
"""
import argparse
import json
import os
import sys
import time

sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
sys.path.insert(0, os.path.dirname(os.path.abspath(__file__)))


def main():
    ap = argparse.ArgumentParser()
    ap.add_argument("--backend", default="gllm_b200", choices=["gllm_b200"])
    ap.add_argument("--model-path", "--model", dest="model_path", required=True)
    ap.add_argument("--load-format", default="auto", choices=["auto", "dummy"])
    ap.add_argument("--dataset", default=None, help="local ShareGPT json; default: synthetic ShareGPT-shaped ids")
    ap.add_argument("--num-prompts", type=int, default=1000)
    ap.add_argument("--seed", type=int, default=0)
    ap.add_argument("--tp", type=int, default=1)
    ap.add_argument("--pp", type=int, default=1)
    ap.add_argument("--maxp", type=int, default=4096)
    ap.add_argument("--maxd", type=int, default=1024)
    ap.add_argument("--gpu-memory-util", "--gpu-memory-utilization", dest="gpu_memory_util", type=float, default=0.9)
    ap.add_argument("--schedule-method", default="chunked_prefill")
    ap.add_argument("--enable-prefix-caching", action="store_true")
    ap.add_argument("--max-cuda-graph-bs", type=int, default=512)
    ap.add_argument("--tp-mode", default="fused", choices=["fused", "nccl"])
    ap.add_argument("--output-json", default=None)
    args = ap.parse_args()
    from gllm_b200 import LLM
    from workloads import from_sharegpt_file, sharegpt_shaped
    llm = LLM(args.model_path, load_format=args.load_format, tp_size=args.tp, pp_size=args.pp, maxp=args.maxp,
              maxd=args.maxd, gpu_memory_util=args.gpu_memory_util, schedule_method=args.schedule_method,
              enable_prefix_caching=args.enable_prefix_caching, max_cuda_graph_bs=args.max_cuda_graph_bs,
              model_max_length=2048 + 16, tp_mode=args.tp_mode, log_stats=False)
    vocab = llm.loader.config["vocab_size"]
    if args.dataset:
        prompts, outs = from_sharegpt_file(args.dataset, llm.tokenizer, args.num_prompts, args.seed)
    else:
        prompts, outs = sharegpt_shaped(args.num_prompts, vocab, args.seed)
    t0 = time.perf_counter()
    llm.generate(tokens=prompts, output_lens=outs, ignore_eos=True, top_k=1, temperature=0.0, progress=True)
    dt = time.perf_counter() - t0
    if llm.is_driver_process:
        n_in, n_out = sum(len(p) for p in prompts), sum(outs)
        res = {"backend": args.backend, "num_prompts": len(prompts), "elapsed_s": round(dt, 3),
               "requests_per_s": round(len(prompts) / dt, 2), "total_tokens_per_s": round((n_in + n_out) / dt, 1),
               "output_tokens_per_s": round(n_out / dt, 1), "input_tokens": n_in, "output_tokens": n_out}
        print(f"Throughput: {res['requests_per_s']:.2f} requests/s, {res['total_tokens_per_s']:.2f} total tokens/s, "
              f"{res['output_tokens_per_s']:.2f} output tokens/s")
        if args.output_json:
            with open(args.output_json, "w") as f:
                json.dump(res, f, indent=2)
    llm.shutdown()
    os._exit(0)


if __name__ == "__main__":
    main()
