"""Run under torchrun (gloo, 2 ranks): a TP-sharded vision tower fed with slices of ONE global state dict must
reproduce the replicated tower's output on every rank. argv: <out file> <qwen2_5|qwen3>"""
import json
import os
import sys

import torch

sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))


class _DictReader:
    def __init__(self, sd):
        self.sd = sd

    def get(self, name):
        return self.sd[name]


def main():
    out, which = sys.argv[1], sys.argv[2]
    rank, world = int(os.environ["RANK"]), int(os.environ["WORLD_SIZE"])
    from gllm_b200.models import vision
    from gllm_b200.parallel import state as ps
    if which == "qwen2_5":
        vc = dict(hidden_size=64, num_heads=4, intermediate_size=96, depth=3, out_hidden_size=48, patch_size=14,
                  spatial_merge_size=2, temporal_patch_size=2, in_channels=3, window_size=56,
                  fullatt_block_indexes=[1])
        cls, ps_ = vision.Qwen2_5_VisionTower, 14
    else:
        vc = dict(hidden_size=64, num_heads=4, intermediate_size=96, depth=3, out_hidden_size=48, patch_size=16,
                  spatial_merge_size=2, temporal_patch_size=2, in_channels=3, num_position_embeddings=16,
                  deepstack_visual_indexes=[0, 1])
        cls, ps_ = vision.Qwen3VisionTower, 16
    # 1. replicated tower (no process group yet) with seeded weights -> the global state dict and the oracle
    torch.manual_seed(1234)
    full = cls(vc, torch.float32, torch.device("cpu"))
    for p in full.parameters():
        p.data.copy_(torch.randn(p.shape) * 0.05)
    sd = {"visual." + k: v.clone() for k, v in full.state_dict().items()}
    grid = torch.tensor([[1, 4, 6], [2, 4, 4]])
    n = int((grid[:, 0] * grid[:, 1] * grid[:, 2]).sum())
    pix = torch.randn(n, 3 * 2 * ps_ * ps_)
    with torch.no_grad():
        ref, ref_deep = full(pix, grid)
    # 2. sharded tower
    ps.init_dist(1, world, rank, rank, os.environ.get("MASTER_ADDR", "127.0.0.1"),
                 int(os.environ.get("MASTER_PORT", "29500")) + 1)
    tower = cls(vc, torch.float32, torch.device("cpu"))
    assert tower.tp == (rank, world), tower.tp
    assert tower.blocks[0].attn.qkv.weight.shape[0] == 3 * 64 // world
    vision.load_vision_weights(tower, _DictReader(sd))
    with torch.no_grad():
        got, got_deep = tower(pix, grid)
    err = float((got - ref).abs().max())
    errs = [float((a - b).abs().max()) for a, b in zip(got_deep, ref_deep)]
    ok = err < 1e-4 and all(e < 1e-4 for e in errs) and len(got_deep) == len(ref_deep)
    flag = torch.tensor([1 if ok else 0])
    torch.distributed.all_reduce(flag, op=torch.distributed.ReduceOp.MIN)
    if rank == 0:
        with open(out, "w") as f:
            json.dump({"ok": bool(flag.item()), "err": err, "deep": errs}, f)
    torch.distributed.barrier()
    torch.distributed.destroy_process_group()


if __name__ == "__main__":
    main()
