import os, sys, time
t0 = time.time()
def log(*a):
    print(f"[probe r{os.environ.get('RANK')} +{time.time()-t0:.1f}s]", *a, flush=True)
log("start")
import torch
import torch.distributed as dist
log("torch imported")
rank, local, world = int(os.environ["RANK"]), int(os.environ["LOCAL_RANK"]), int(os.environ["WORLD_SIZE"])
torch.cuda.set_device(local)
x = torch.ones(4, device="cuda")
log("cuda ready", torch.cuda.get_device_name(local))
dist.init_process_group("nccl", device_id=torch.device("cuda", local))
log("pg ready")
dist.all_reduce(x)
torch.cuda.synchronize()
log("allreduce ok", x[0].item())
g = dist.new_group(list(range(world)))
dist.barrier(group=g)
log("subgroup ok")
import torch.distributed._symmetric_memory as symm
t = symm.empty(1 << 20, dtype=torch.uint8, device=torch.device("cuda", local))
log("symm.empty ok")
h = symm.rendezvous(t, g.group_name)
log("rendezvous ok", [hex(p) for p in h.buffer_ptrs][:2], "mc", getattr(h, "multicast_ptr", None))
dist.barrier()
dist.destroy_process_group()
log("done")
