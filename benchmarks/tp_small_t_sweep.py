"""Where should the fused TP path switch from the latency-optimised form (replicated rows, one-kernel LL
all-reduce + add + norm) to the bandwidth-optimised one (GEMM⊕reduce-scatter / all-gather⊕GEMM over token shards)?
Engine-level decode step time (CUDA graphs) per batch size for two thresholds, under torchrun:

    GLLM_TP_SMALL_T=64  torchrun ... benchmarks/tp_small_t_sweep.py      # small form up to 64 tokens
    GLLM_TP_SMALL_T=512 torchrun ... benchmarks/tp_small_t_sweep.py      # small form everywhere it fits

Prints one JSON line per batch size (device time, max over ranks)."""
import json
import os
import sys

import torch
import torch.distributed as dist

sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
sys.path.insert(0, os.path.dirname(os.path.abspath(__file__)))


def main():
    rank, local, world = int(os.environ["RANK"]), int(os.environ["LOCAL_RANK"]), int(os.environ["WORLD_SIZE"])
    torch.cuda.set_device(local)
    from gllm_b200.config import EngineConfig
    from gllm_b200.model_runner import ModelRunner
    from gllm_b200.parallel import fused, state as ps
    from step_breakdown import make_batch
    ps.init_dist(1, world, rank, local)
    cfg = EngineConfig(model_path=os.environ.get("MODEL", "preset:qwen3-8b"), load_format="dummy", maxp=4096,
                       maxd=1024, tp_size=world, tp_mode="fused", max_cuda_graph_bs=512, num_gpu_pages=40000,
                       model_max_length=2064)
    r = ModelRunner(cfg)
    r.init(f"cuda:{local}")
    for b in (1, 8, 32, 64, 96, 128, 192, 256, 384, 512):
        batch = make_batch(b, 512, 0, cfg.page_size)
        for _ in range(3):
            r.step(batch)
        torch.cuda.synchronize()
        dist.barrier()
        e0, e1 = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
        n = 20
        e0.record()
        for _ in range(n):
            r.step(batch)
        e1.record()
        torch.cuda.synchronize()
        t = torch.tensor([e0.elapsed_time(e1) / n], device="cuda")
        dist.all_reduce(t, op=dist.ReduceOp.MAX)
        if rank == 0:
            print(json.dumps({"tp": world, "small_t": fused.SMALL_T, "batch": b, "ctx": 512,
                              "ms_per_step": round(t.item(), 3)}), flush=True)
        dist.barrier()
    r.close()
    dist.barrier()
    dist.destroy_process_group()


if __name__ == "__main__":
    main()
