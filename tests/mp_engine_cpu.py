"""CPU multi-process engine check (gloo): run under torchrun with WORLD_SIZE = pp*tp ranks.
usage: mp_engine_cpu.py <pp> <tp> <out_json>   — rank 0 writes the generated tokens."""
import json
import os
import sys

import torch

sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))


def main():
    pp, tp, out = int(sys.argv[1]), int(sys.argv[2]), sys.argv[3]
    arch = sys.argv[4] if len(sys.argv) > 4 else "Qwen3ForCausalLM"
    method = sys.argv[5] if len(sys.argv) > 5 else "chunked_prefill"
    from gllm_b200 import LLM
    from gllm_b200.models.presets import tiny
    over = {}
    if arch == "MixtralForCausalLM":
        over = dict(num_local_experts=4, num_experts_per_tok=2)
    over.update(json.loads(os.environ.get("GLLM_TEST_CFG", "{}")))     # shape overrides for sharding edge cases
    cfg = tiny(arch, **{"num_hidden_layers": 4, **over})
    torch.manual_seed(0)
    llm = LLM(cfg, load_format="dummy", pp_size=pp, tp_size=tp, maxp=48, maxd=16, num_cpu_pages=128,
              model_max_length=256, log_stats=False, device="cpu", launch_mode="inproc", schedule_method=method,
              assigned_layers=os.environ.get("GLLM_TEST_ASSIGNED") or None,
              use_ep=os.environ.get("GLLM_TEST_NO_EP") != "1",
              seed=0, async_schedule=os.environ.get("GLLM_TEST_ASYNC") == "1",
              enable_prefix_caching=os.environ.get("GLLM_TEST_ASYNC") != "1")
    # identical weights on every layout: re-initialise from one global state dict
    sys.path.insert(0, os.path.dirname(os.path.abspath(__file__)))
    from shard_util import load_global_weights
    load_global_weights(llm.worker.runner.model, cfg, seed=123)
    if os.environ.get("GLLM_TEST_TWO_ENGINES") == "1":
        # a second engine in the same processes, right after the first one was torn down: its ipc endpoints must not
        # be the files the first engine's (asynchronously closing) sockets still own
        warm = llm.generate(tokens=[[5, 6, 7]], output_lens=[3], ignore_eos=True)
        assert int(os.environ.get("RANK", "0")) != 0 or len(warm[0].token_ids) == 6
        llm.close()
        torch.manual_seed(0)
        llm = LLM(cfg, load_format="dummy", pp_size=pp, tp_size=tp, maxp=48, maxd=16, num_cpu_pages=128,
                  model_max_length=256, log_stats=False, device="cpu", launch_mode="inproc", schedule_method=method,
                  seed=0)
        load_global_weights(llm.worker.runner.model, cfg, seed=123)
    prompts = [[5, 17, 99, 200, 3, 45, 7], [9] * 40, list(range(20, 120)), [300, 301]]
    if os.environ.get("GLLM_TEST_SAMPLED") == "1":
        # sampled + penalty requests next to greedy ones (vocab-parallel sampling under TP): rows 0 and 2 are greedy
        # and must stay token-identical on every layout; rows 1 and 3 must come back complete and in vocabulary
        seqs = llm.generate(tokens=prompts, output_lens=[8] * 4, ignore_eos=True, temperature=[0.0, 0.8, 0.0, 1.0],
                            top_p=[1.0, 0.9, 1.0, 1.0], top_k=[1, 8, 1, 0], repetition_penalty=[1.0, 1.2, 1.3, 1.0])
        outs = seqs
        if int(os.environ.get("RANK", "0")) == 0:
            vocab = cfg["vocab_size"]
            assert all(len(s.token_ids) == len(p) + 8 and 0 <= min(s.token_ids) and max(s.token_ids) < vocab
                       for s, p in zip(seqs, prompts)), [s.token_ids for s in seqs]
            if tp > 1:
                assert llm.worker.runner.stats.get("vp_sample_steps", 0) > 0, "vocab-parallel sampling did not run"
            with open(out, "w") as f:
                json.dump([seqs[0].token_ids, seqs[2].token_ids], f)
        out = os.devnull
    else:
        outs = llm.generate(tokens=prompts, output_lens=[8] * 4, ignore_eos=True)
    if int(os.environ.get("RANK", "0")) == 0:
        with open(out, "w") as f:
            json.dump([s.token_ids for s in outs], f)
    if os.environ.get("GLLM_TEST_PP_STATS"):      # per-rank evidence that the sharded stage transfer really ran
        from gllm_b200.parallel import state as ps
        with open(f"{out}.pp{os.environ.get('RANK', '0')}", "w") as f:
            json.dump(ps.PP_STATS, f)
    llm.shutdown()
    if torch.distributed.is_initialized():
        torch.distributed.barrier()
        torch.distributed.destroy_process_group()


if __name__ == "__main__":
    main()
