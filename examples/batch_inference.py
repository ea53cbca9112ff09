"""Offline batch inference (reference: examples/batch_inference.py): throughput in requests/s and
input / output tokens/s; prompts are synthetic ShareGPT-shaped ids unless --prompts-file is given."""
import argparse
import os
import sys
import time

sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
sys.path.insert(0, os.path.join(os.path.dirname(os.path.dirname(os.path.abspath(__file__))), "benchmarks"))

if __name__ == "__main__":
    ap = argparse.ArgumentParser()
    ap.add_argument("--model-path", required=True)
    ap.add_argument("--load-format", default="auto")
    ap.add_argument("--num-prompt", "--num-prompts", dest="num_prompts", type=int, default=64)
    ap.add_argument("--prompts-file", default=None, help="one text prompt per line (needs a tokenizer)")
    ap.add_argument("--output-len", type=int, default=128)
    ap.add_argument("--tp", type=int, default=1)
    ap.add_argument("--pp", type=int, default=1)
    ap.add_argument("--print-output", action="store_true")
    args = ap.parse_args()
    from gllm_b200 import LLM
    from workloads import sharegpt_shaped
    llm = LLM(args.model_path, load_format=args.load_format, tp_size=args.tp, pp_size=args.pp)
    t0 = time.time()
    if args.prompts_file:
        prompts = [l.strip() for l in open(args.prompts_file) if l.strip()]
        seqs = llm.generate(prompts=prompts, output_lens=[args.output_len] * len(prompts))
    else:
        toks, outs = sharegpt_shaped(args.num_prompts, llm.loader.config["vocab_size"])
        seqs = llm.generate(tokens=toks, output_lens=outs, ignore_eos=True)
    dt = time.time() - t0
    if llm.is_driver_process:
        n_in = sum(s.prompt_len for s in seqs)
        n_out = sum(len(s.token_ids) - s.prompt_len for s in seqs)
        print(f"{len(seqs) / dt:.2f} reqs/s  {n_in / dt:.1f} input tok/s  {n_out / dt:.1f} output tok/s")
        if args.print_output:
            for s in seqs:
                print("-" * 40, "\n", s.prompt, "\n>>>", s.output or s.token_ids[s.prompt_len:])
    llm.shutdown()
