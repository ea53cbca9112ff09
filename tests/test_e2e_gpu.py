"""End-to-end on the GPU: the engine (sm_100a kernels, CUDA graphs) must reproduce the CPU oracle
path's greedy tokens on a small random model."""
import pytest
import torch

pytestmark = pytest.mark.gpu


def _cfg():
    from gllm_b200.models.presets import tiny
    return tiny("Qwen3ForCausalLM", hidden_size=256, num_hidden_layers=3, num_attention_heads=4,
                num_key_value_heads=2, head_dim=64, intermediate_size=512, vocab_size=1024, torch_dtype="bfloat16")


def _gen(device, graphs, prompts, out_len, state=None):
    from gllm_b200 import LLM
    llm = LLM(_cfg(), load_format="dummy", maxp=128, maxd=64, max_cuda_graph_bs=8, num_gpu_pages=256,
              num_cpu_pages=256, model_max_length=512, log_stats=False, device=device,
              disable_cuda_graph=not graphs)
    model = llm.worker.runner.model
    if state is not None:
        for (n, p), (_, q) in zip(model.named_parameters(), state):
            p.data.copy_(q.to(p.device))
    outs = llm.generate(tokens=prompts, output_lens=[out_len] * len(prompts), ignore_eos=True)
    toks = [s.token_ids[len(p):] for s, p in zip(outs, prompts)]
    params = [(n, p.detach().cpu().clone()) for n, p in model.named_parameters()]
    logits_stats = llm.worker.runner.stats.copy()
    llm.shutdown()
    return toks, params, logits_stats


def test_gpu_matches_cpu_oracle():
    torch.manual_seed(0)
    prompts = [[5, 9, 100, 7], list(range(20, 190)), [77] * 33, [3, 1, 4, 1, 5, 9, 2, 6]]
    cpu_toks, params, _ = _gen("cpu", False, prompts, 6)
    gpu_toks, _, st = _gen("cuda", True, prompts, 6, state=params)
    assert st["graph_steps"] > 0, "CUDA graphs were not used for decode"
    # bf16 kernels vs fp32-accumulated oracle: allow rare argmax flips after the first tokens
    agree = sum(a == b for x, y in zip(cpu_toks, gpu_toks) for a, b in zip(x, y))
    total = sum(len(x) for x in cpu_toks)
    assert [x[0] for x in cpu_toks] == [y[0] for y in gpu_toks], (cpu_toks, gpu_toks)
    # once a near-tie flips, the rest of that sequence differs: require most sequences to match fully
    same_seqs = sum(x == y for x, y in zip(cpu_toks, gpu_toks))
    assert same_seqs >= len(prompts) // 2 and agree / total >= 0.6, (cpu_toks, gpu_toks)


def test_graph_and_eager_agree():
    prompts = [[5, 9, 100, 7], list(range(20, 150)), [77] * 33]
    t1, params, _ = _gen("cuda", False, prompts, 8)
    t2, _, st = _gen("cuda", True, prompts, 8, state=params)
    assert st["graph_steps"] > 0
    assert t1 == t2
