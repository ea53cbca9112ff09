"""Reference arm of bench.py: the UNMODIFIED reference (installed by baseline/install_reference.sh into
baseline/_ref) driven through its own public API — `from gllm import LLM; LLM(...).generate(tokens=..)` — on the
headline workload (Qwen3-8B shape, random-init weights via its `load_format="dummy"`, the same synthetic
ShareGPT-shaped requests as our arm). Prints ONE JSON line on stdout; any failure is reported by the caller
(bench.py) as {"impl": "reference", "unavailable": ...}.

The reference spawns its own worker processes (one per GPU), so this runs in ONE process whatever N is; timing is
wall clock around `generate` (the front-end process owns no CUDA context), K timed passes after W warm-up passes.
"""
import argparse
import json
import os
import sys
import tempfile
import time

HERE = os.path.dirname(os.path.abspath(__file__))
ROOT = os.path.dirname(HERE)


def qwen3_8b_dir(path):
    """config.json + a stand-in tokenizer (one token per id; the benchmark feeds token ids) for Qwen3-8B."""
    from tokenizers import Tokenizer, models, pre_tokenizers
    from transformers import PreTrainedTokenizerFast
    vocab = 151936
    cfg = {"architectures": ["Qwen3ForCausalLM"], "model_type": "qwen3", "hidden_size": 4096,
           "intermediate_size": 12288, "num_hidden_layers": 36, "num_attention_heads": 32, "num_key_value_heads": 8,
           "head_dim": 128, "vocab_size": vocab, "max_position_embeddings": 40960, "rms_norm_eps": 1e-6,
           "rope_theta": 1000000.0, "rope_scaling": None, "hidden_act": "silu", "tie_word_embeddings": False,
           "attention_bias": False, "torch_dtype": "bfloat16", "bos_token_id": 151643, "eos_token_id": 151645}
    os.makedirs(path, exist_ok=True)
    with open(os.path.join(path, "config.json"), "w") as f:
        json.dump(cfg, f)
    with open(os.path.join(path, "generation_config.json"), "w") as f:
        # repetition_penalty: transformers 4 defaulted it to 1.0, transformers 5 leaves it None and the reference
        # feeds generation_config.repetition_penalty straight into a float tensor (llm_engine.py:320-324)
        json.dump({"eos_token_id": [151645, 151643], "temperature": 0.6, "top_p": 0.95, "top_k": 20,
                   "repetition_penalty": 1.0}, f)
    tok = Tokenizer(models.WordLevel({f"t{i}": i for i in range(vocab)}, unk_token="t0"))
    tok.pre_tokenizer = pre_tokenizers.Whitespace()
    # model_max_length as in the real checkpoint's tokenizer_config.json: the reference derives its admission
    # limit from it (model_runner.py:135-143) and would otherwise fall back to generation_config.max_length
    PreTrainedTokenizerFast(tokenizer_object=tok, unk_token="t0", eos_token="t151645", pad_token="t151643",
                            model_max_length=40960).save_pretrained(path)
    return path


def main():
    ap = argparse.ArgumentParser()
    ap.add_argument("--gpus", type=int, default=1)
    ap.add_argument("--steps", type=int, default=3)
    ap.add_argument("--warmup", type=int, default=3)
    ap.add_argument("--num-prompts", type=int, default=1000)
    ap.add_argument("--maxp", type=int, default=4096)
    ap.add_argument("--maxd", type=int, default=1024)
    ap.add_argument("--max-cuda-graph-bs", type=int, default=512)
    ap.add_argument("--seed", type=int, default=0)
    args = ap.parse_args()
    sys.path.insert(0, ROOT)
    from bench import synth_requests          # the same request generator as our arm
    prompts, outs = synth_requests(args.num_prompts, 151936, args.seed)
    from gllm import LLM                      # baseline/_ref on PYTHONPATH (set by bench.py)
    model_dir = qwen3_8b_dir(tempfile.mkdtemp(prefix="gllm_ref_qwen3_8b_"))
    llm = LLM(model_dir, load_format="dummy", tp_size=args.gpus, pp_size=1, maxp=args.maxp, maxd=args.maxd,
              max_cuda_graph_bs=args.max_cuda_graph_bs, enable_prefix_caching=True,
              schedule_method="chunked_prefill", master_port=str(int(os.environ.get("GLLM_REF_PORT", "18000"))),
              zmq_port_base=int(os.environ.get("GLLM_REF_PORT", "18000")) + 1)

    n_pass = args.warmup + args.steps
    pass_prompts = [prompts if i == 0 else synth_requests(args.num_prompts, 151936, args.seed, i)[0]
                    for i in range(n_pass)]     # fresh token ids per pass, as in our arm (no cross-pass cache hits)
    pass_no = [0]

    def one_pass():
        prompts = pass_prompts[min(pass_no[0], n_pass - 1)]
        pass_no[0] += 1
        t0 = time.perf_counter()
        # sampling arguments left at the reference's defaults: top_k defaults to 1, i.e. greedy (llm_engine.py:316-325)
        seqs = llm.generate(tokens=[list(p) for p in prompts], output_lens=list(outs))
        dt = time.perf_counter() - t0
        n_out = sum(len(s.token_ids) - s.prompt_len for s in seqs)
        return dt, n_out

    # The reference's generate() has no ignore_eos switch: a sequence may stop early on an EOS id; the value counts
    # the tokens it really produced (and `output_tokens_expected` says how many the workload asks for).
    t_start = time.perf_counter()
    budget = float(os.environ.get("GLLM_REF_BUDGET_S", "1e9"))      # seconds this script may spend in passes
    warm_done = 0
    warm_allowed = args.warmup
    for _ in range(args.warmup):
        if warm_done >= warm_allowed:
            break
        dt, _n = one_pass()
        warm_done += 1
        # The reference needs ~1 minute per pass on this workload: the K timed passes have priority over the later
        # warm-up passes inside the caller's time limit (the first warm-up pass always runs: it compiles / caches).
        left = budget - (time.perf_counter() - t_start)
        warm_allowed = min(warm_allowed, warm_done + max(0, int((left - args.steps * dt) // max(dt, 1e-3))))
    tot_t = tot_tok = 0
    steps_done = 0
    for _ in range(args.steps):
        dt, n = one_pass()
        tot_t += dt
        tot_tok += n
        steps_done += 1
        if time.perf_counter() - t_start + dt > budget:
            break
    value = tot_tok / tot_t
    print(json.dumps({
        "impl": "reference", "metric": "output_tokens_per_s", "value": round(value, 2), "unit": "tokens/s",
        "n_gpus": args.gpus, "steps": steps_done, "warmup": warm_done,
        "steps_requested": args.steps, "warmup_requested": args.warmup,
        "ms_per_step": round(tot_t / steps_done * 1e3, 2),
        "output_tokens_per_step": tot_tok // steps_done, "output_tokens_expected": sum(outs), "higher_is_better": True, "scaling": "strong",
        "vs_baseline": None, "dtype": "bf16",
        "data": "synthetic ShareGPT-shaped token ids (same generator and seed as our arm); random-init weights "
                "(reference load_format=dummy)",
        "timing": "wall clock around LLM.generate in the front-end process (workers are separate processes)",
        "config": {"model": "Qwen3-8B", "num_prompts": args.num_prompts, "parallelism": f"tp{args.gpus}",
                   "maxp": args.maxp, "maxd": args.maxd, "max_cuda_graph_bs": args.max_cuda_graph_bs,
                   "enable_prefix_caching": True, "schedule_method": "chunked_prefill", "native_kernels": "vLLM 0.22 libraries of this image "
                   "(the reference pins vLLM 0.11)"},
        "e2e": {"value": round(value, 2), "unit": "tokens/s"}}), flush=True)


def _leave():
    """The reference's worker processes hold our stdout / stderr pipes: bench.py would wait for EOF until they are
    gone (also after a failure, if a worker hangs in a collective). bench.py starts this script as the leader of its
    own process group, so take the whole group down."""
    sys.stdout.flush()
    sys.stderr.flush()
    if os.getpgid(0) == os.getpid():
        import signal
        os.killpg(os.getpid(), signal.SIGKILL)
    os._exit(0)


if __name__ == "__main__":
    try:
        main()
    except BaseException:  # noqa: BLE001
        import traceback
        traceback.print_exc()
    finally:
        _leave()
