"""Engine-level decode/prefill step time under TP (torchrun): fused vs nccl, eager vs CUDA graph."""
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
    from gllm_b200.parallel import state as ps
    from step_breakdown import make_batch
    ps.init_dist(1, world, rank, local)
    res = {}
    for mode in ("nccl", "fused"):
        cfg = EngineConfig(model_path=os.environ.get("MODEL", "preset:qwen3-8b"), load_format="dummy", maxp=4096,
                           maxd=1024, tp_size=world, tp_mode=mode, max_cuda_graph_bs=256, num_gpu_pages=40000,
                           model_max_length=2064)
        r = ModelRunner(cfg)
        r.init(f"cuda:{local}")
        for b, ctx, pre in ((16, 512, 0), (64, 512, 0), (128, 512, 0), (256, 512, 0), (4, 0, 1024)):
            batch = make_batch(b, ctx, pre, cfg.page_size)
            for _ in range(3):
                r.step(batch)
            torch.cuda.synchronize()
            dist.barrier()
            e0, e1 = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
            n = 10
            e0.record()
            for _ in range(n):
                r.step(batch)
            e1.record()
            torch.cuda.synchronize()
            t = torch.tensor([e0.elapsed_time(e1) / n], device="cuda")
            dist.all_reduce(t, op=dist.ReduceOp.MAX)
            res[(mode, b, pre)] = round(t.item(), 3)
            dist.barrier()
        del r
        torch.cuda.empty_cache()
    if rank == 0:
        for (b, pre) in ((16, 0), (128, 0), (256, 0), (4, 1024)):
            print(json.dumps({"tp": world, "batch": b, "prefill": pre, "ms_nccl": res[("nccl", b, pre)],
                              "ms_fused": res[("fused", b, pre)]}), flush=True)
    dist.barrier()
    sys.stdout.flush()
    os._exit(0)


if __name__ == "__main__":
    main()
