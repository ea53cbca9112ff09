"""Soak test of the hand-rolled NVLink synchronisation (epochs, parity double buffers, arrival counters) in
csrc/comm/tp_fused.cu: thousands of CUDA-graph replays of fused-TP decode steps at alternating batch sizes — LL /
NVLS one-shot all-reduce for the small ones, GEMM⊕reduce-scatter / all-gather⊕GEMM for the large — with the logits
of every 500th replay compared against the NCCL strategy on the same inputs. Run under torchrun (one rank per GPU):

    torchrun --nproc-per-node 2 tests/mp_tp_soak.py [replays]

Prints TP_SOAK_OK on rank 0. A lost flag or a parity slip shows up as a bounded-spin trap, a stall (the watchdog
dumps stacks) or a logits mismatch."""
import faulthandler
import os
import sys

import torch
import torch.distributed as dist

sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
sys.path.insert(0, os.path.join(os.path.dirname(os.path.dirname(os.path.abspath(__file__))), "benchmarks"))


def main():
    n_replays = int(sys.argv[1]) if len(sys.argv) > 1 else 10000
    faulthandler.dump_traceback_later(int(os.environ.get("GLLM_TP_SOAK_TIMEOUT", "240")), exit=True)
    rank, local, world = int(os.environ["RANK"]), int(os.environ["LOCAL_RANK"]), int(os.environ["WORLD_SIZE"])
    torch.cuda.set_device(local)
    from gllm_b200.config import EngineConfig
    from gllm_b200.model_runner import ModelRunner
    from gllm_b200.models.presets import tiny
    from gllm_b200.parallel import state as ps
    from step_breakdown import make_batch
    ps.init_dist(1, world, rank, local)
    cfg_model = tiny("Qwen3ForCausalLM", hidden_size=1024, num_hidden_layers=4, num_attention_heads=16,
                     num_key_value_heads=8, head_dim=64, intermediate_size=2048, vocab_size=4096,
                     torch_dtype="bfloat16")
    runners = {}
    for mode in ("nccl", "fused"):
        torch.manual_seed(1234 + rank)
        cfg = EngineConfig(model_path=cfg_model, load_format="dummy", maxp=512, maxd=256, tp_size=world, tp_mode=mode,
                           max_cuda_graph_bs=256, num_gpu_pages=8192, model_max_length=1024, seed=5)
        r = ModelRunner(cfg)
        r.init(f"cuda:{local}")
        r.keep_logits = True
        runners[mode] = r
    # identical weights in both runners
    for (na, pa), (nb, pb) in zip(runners["nccl"].model.named_parameters(), runners["fused"].model.named_parameters()):
        pb.data.copy_(pa.data)
    import numpy as np
    np.random.seed(0)       # the same synthetic tokens on every rank
    sizes = [1, 8, 33, 64, 65, 128, 200, 256]
    batches = {b: make_batch(b, 300, 0, 16) for b in sizes}
    worst = 0.0
    checked = 0
    for it in range(n_replays):
        b = sizes[it % len(sizes)]
        check = it % 500 < len(sizes)
        f = runners["fused"]
        f.logit_log.clear()
        f.keep_logits = check
        f.step(batches[b])
        if check:
            lf = f.logit_log[-1][1]
            n = runners["nccl"]
            n.logit_log.clear()
            n.step(batches[b])
            ln = n.logit_log[-1][1]
            err = float((lf - ln).abs().max()) / (float(ln.abs().max()) + 1e-6)
            worst = max(worst, err)
            checked += 1
    torch.cuda.synchronize()
    ok = torch.tensor([1 if worst < 4e-2 else 0], device="cuda")
    dist.all_reduce(ok, op=dist.ReduceOp.MIN)
    if rank == 0:
        st = runners["fused"].stats
        print(f"{n_replays} fused-TP decode steps ({st['graph_steps']} graph replays, NVLS calls "
              f"{getattr(runners['fused'].tpc, 'nvls_calls', 0)}), {checked} logits comparisons vs NCCL, worst max-abs err "
              f"{worst:.4f} of the row scale", flush=True)
        print("TP_SOAK_OK" if ok.item() == 1 else "TP_SOAK_FAILED", flush=True)
    for r in runners.values():
        r.close()
    dist.barrier()
    dist.destroy_process_group()


if __name__ == "__main__":
    main()
