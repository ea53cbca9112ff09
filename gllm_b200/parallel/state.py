"""Distributed state: ranks, groups, PP layer partition, PP p2p, baseline collectives.

Reference: gllm/dist_utils.py. Same rank mapping (`rank = pp_rank * tp_size + tp_rank`), same EP
convention (EP group == TP group), same layer partition rule (ceil(L / pp) per stage with the
remainder on the last stage, or an explicit `assigned_layers` list).

`torch.distributed` (NCCL on GPUs, gloo on CPU for the plumbing tests) provides rendezvous, PP
p2p, control-plane collectives and the *baseline* TP collectives. The product TP/EP hot paths use
the fused NVLink kernels in `gllm_b200.parallel.fused` instead.
"""
from __future__ import annotations

import datetime
from dataclasses import dataclass, field
from typing import List, Optional

import torch
import torch.distributed as dist

from gllm_b200.utils.logging import logger


@dataclass
class ParallelState:
    rank: int = 0
    world_size: int = 1
    pp_rank: int = 0
    pp_size: int = 1
    tp_rank: int = 0
    tp_size: int = 1
    ep_rank: int = 0
    ep_size: int = 1
    local_rank: int = 0
    tp_group: Optional[object] = None
    tp_ranks: List[int] = field(default_factory=lambda: [0])
    assigned_layers: Optional[List[int]] = None
    initialized: bool = False
    backend: str = "none"


_STATE = ParallelState()


def get_state() -> ParallelState:
    return _STATE


def reset_state():
    global _STATE
    _STATE = ParallelState()


def get_rank(): return _STATE.rank
def get_world_size(): return _STATE.world_size
def get_pp_rank(): return _STATE.pp_rank
def get_pp_size(): return _STATE.pp_size
def get_tp_rank(): return _STATE.tp_rank
def get_tp_size(): return _STATE.tp_size
def get_ep_rank(): return _STATE.ep_rank
def get_ep_size(): return _STATE.ep_size
def get_local_rank(): return _STATE.local_rank
def get_tp_group(): return _STATE.tp_group
def is_first_pp_rank(): return _STATE.pp_rank == 0
def is_last_pp_rank(): return _STATE.pp_rank == _STATE.pp_size - 1
def is_driver(): return _STATE.rank == 0


def get_output_rank() -> int:
    """First TP rank of the last stage samples and reports tokens (gllm/dist_utils.py:71-76)."""
    return (_STATE.pp_size - 1) * _STATE.tp_size


def is_output_rank() -> bool:
    return _STATE.rank == get_output_rank()


def get_next_pp_rank() -> int:
    return _STATE.rank + _STATE.tp_size


def get_prev_pp_rank() -> int:
    return _STATE.rank - _STATE.tp_size


def init_dist(pp_size: int, tp_size: int, rank: int, local_rank: int, master_addr: str = "127.0.0.1",
              master_port: int = 8001, use_ep: bool = True, assigned_layers: Optional[List[int]] = None,
              backend: Optional[str] = None, init_process_group: bool = True, timeout_s: int = 600):
    """Create the world group (+ one TP group per stage)."""
    global _STATE
    world = pp_size * tp_size
    st = ParallelState(rank=rank, world_size=world, pp_rank=rank // tp_size, pp_size=pp_size,
                       tp_rank=rank % tp_size, tp_size=tp_size, local_rank=local_rank,
                       assigned_layers=assigned_layers)
    if use_ep:
        st.ep_rank, st.ep_size = st.tp_rank, st.tp_size
    if backend is None:
        backend = "nccl" if torch.cuda.is_available() else "gloo"
    st.backend = backend
    if world > 1:
        if init_process_group and not dist.is_initialized():
            kw = {}
            if backend == "nccl":
                kw["device_id"] = torch.device("cuda", local_rank)
            import os
            init_method = f"tcp://{master_addr}:{master_port}"
            if os.environ.get("TORCHELASTIC_USE_AGENT_STORE") == "True" and \
                    int(os.environ.get("WORLD_SIZE", "0")) == world:
                # launched by torchrun: the elastic agent hosts the store (a tcp:// init on another port
                # would wait forever for a server nobody starts)
                init_method = "env://"
            dist.init_process_group(backend=backend, init_method=init_method,
                                    world_size=world, rank=rank,
                                    timeout=datetime.timedelta(seconds=timeout_s), **kw)
            logger.info("dist init: rank %d / %d (pp %d/%d, tp %d/%d) backend=%s", rank, world, st.pp_rank,
                        pp_size, st.tp_rank, tp_size, backend)
        # every rank must create every group, in the same order
        for stage in range(pp_size):
            ranks = list(range(stage * tp_size, (stage + 1) * tp_size))
            grp = dist.new_group(ranks) if tp_size > 1 else None
            if stage == st.pp_rank:
                st.tp_group, st.tp_ranks = grp, ranks
    st.initialized = True
    _STATE = st
    return st


def destroy():
    if dist.is_initialized():
        try:
            dist.destroy_process_group()
        except Exception:  # noqa: BLE001
            pass
    reset_state()


# ------------------------------------------------------------------------------------------------
# PP layer partition
# ------------------------------------------------------------------------------------------------
def partition_layers(num_layers: int, pp_size: int, assigned_layers: Optional[List[int]] = None) -> List[range]:
    """-> layer range per stage. Default: ceil(L/pp) per stage, remainder on the last stage
    (gllm/dist_utils.py:176-187)."""
    if assigned_layers is not None:
        assert len(assigned_layers) == pp_size and sum(assigned_layers) == num_layers, \
            f"assigned_layers {assigned_layers} must have {pp_size} entries summing to {num_layers}"
        out, s = [], 0
        for n in assigned_layers:
            out.append(range(s, s + n))
            s += n
        return out
    per = (num_layers + pp_size - 1) // pp_size
    out = []
    for i in range(pp_size):
        a = min(i * per, num_layers)
        b = num_layers if i == pp_size - 1 else min((i + 1) * per, num_layers)
        out.append(range(a, b))
    # stages before the last take `per`; make sure the last one is not empty when L < pp*per
    if pp_size > 1 and len(out[-1]) == 0:
        # fall back to an even floor split with the remainder spread from the front
        base, rem = divmod(num_layers, pp_size)
        out, s = [], 0
        for i in range(pp_size):
            n = base + (1 if i < rem else 0)
            out.append(range(s, s + n))
            s += n
    return out


def get_pp_layers(num_layers: int) -> range:
    return partition_layers(num_layers, _STATE.pp_size, _STATE.assigned_layers)[_STATE.pp_rank]


# ------------------------------------------------------------------------------------------------
# baseline collectives (NCCL / gloo)
# ------------------------------------------------------------------------------------------------
def tp_all_reduce(x: torch.Tensor) -> torch.Tensor:
    if _STATE.tp_size == 1:
        return x
    dist.all_reduce(x, group=_STATE.tp_group)
    return x


def tp_all_gather_last_dim(x: torch.Tensor) -> torch.Tensor:
    """[.., n] per rank -> [.., n * tp] (gllm/dist_utils.py:230-250)."""
    if _STATE.tp_size == 1:
        return x
    tp = _STATE.tp_size
    x2 = x.contiguous().view(-1, x.shape[-1])
    out = torch.empty((tp * x2.shape[0], x2.shape[1]), dtype=x.dtype, device=x.device)
    dist.all_gather_into_tensor(out, x2, group=_STATE.tp_group)
    out = out.view((tp,) + tuple(x.shape))
    return out.movedim(0, -2).reshape(*x.shape[:-1], tp * x.shape[-1])


def tp_all_gather_first_dim(x: torch.Tensor) -> torch.Tensor:
    if _STATE.tp_size == 1:
        return x
    out = torch.empty((_STATE.tp_size * x.shape[0],) + tuple(x.shape[1:]), dtype=x.dtype, device=x.device)
    dist.all_gather_into_tensor(out, x.contiguous(), group=_STATE.tp_group)
    return out


def pp_send(tensors: List[torch.Tensor], dst: Optional[int] = None):
    dst = get_next_pp_rank() if dst is None else dst
    return [dist.isend(t, dst) for t in tensors]


def pp_recv(tensors: List[torch.Tensor], src: Optional[int] = None):
    src = get_prev_pp_rank() if src is None else src
    for t in tensors:
        dist.recv(t, src)


def divide(a: int, b: int) -> int:
    assert a % b == 0, f"{a} is not divisible by {b}"
    return a // b


# ------------------------------------------------------------------------------------------------
# tile-streamed PP p2p: the activation travels as row tiles over NCCL p2p; the receiver
# posts every tile's irecv up front and lets its compute stream wait per tile, so the first layer's
# add+RMSNorm and QKV GEMM of tile i overlap the transfer of tiles i+1.. (the reference blocks on the
# whole tensor, gllm/worker.py:141-157).
# ------------------------------------------------------------------------------------------------
def pp_row_tiles(num_rows: int, min_rows: Optional[int] = None, max_tiles: int = 8):
    if min_rows is None:
        import os
        min_rows = int(os.environ.get("GLLM_PP_TILE_ROWS", "512"))
    n = max(1, min(max_tiles, num_rows // min_rows))
    per = (num_rows + n - 1) // n
    return [(i * per, min((i + 1) * per, num_rows)) for i in range(n) if i * per < num_rows]


def _pp_shard_rows(r0: int, r1: int):
    """TP>1: instead of every TP rank of a stage shipping the same replicated [rows, H] tile to its peer in the
    next stage (what the reference does: tp_size redundant copies), rank k ships only the k-th
    1/tp slice of the tile and the receiving stage re-assembles it with one all-gather over its own TP group
    (NVLink) — tp x fewer bytes on the inter-stage link, which is the slow one when stages sit on different nodes.
    Small tiles (decode) keep the replicated send: one more collective would cost more than the bytes saved.
    -> (a, b, per) = this rank's row range and the padded slice length, or None for a replicated tile."""
    import os
    tp = get_tp_size()
    if tp == 1 or os.environ.get("GLLM_PP_SHARD", "1") != "1":
        return None
    if r1 - r0 < int(os.environ.get("GLLM_PP_SHARD_MIN_ROWS", "256")):
        return None
    per = -(-(r1 - r0) // tp)
    a = min(r0 + get_tp_rank() * per, r1)
    return a, min(a + per, r1), per


PP_STATS = {"sharded_tiles": 0, "replicated_tiles": 0}   # stage-boundary tiles received by this rank


def pp_send_tiled(tensors: List[torch.Tensor], dst: Optional[int] = None):
    dst = get_next_pp_rank() if dst is None else dst
    handles = []
    for r0, r1 in pp_row_tiles(tensors[0].shape[0]):
        sh = _pp_shard_rows(r0, r1)
        a, b = (r0, r1) if sh is None else sh[:2]
        if b > a:
            for t in tensors:
                handles.append(dist.isend(t[a:b], dst))
    return handles


class _ShardedTileWork:
    """`wait()`-compatible handle of one sharded tile: p2p receive of this rank's slice, then the all-gather that
    rebuilds rows r0:r1 of `tensor` on every TP rank of the stage."""

    def __init__(self, works, tensor: torch.Tensor, r0: int, r1: int, a: int, b: int, per: int):
        self.works, self.tensor, self.r0, self.r1, self.a, self.b, self.per = works, tensor, r0, r1, a, b, per

    def wait(self):
        for w in self.works:
            w.wait()
        t = self.tensor
        mine = torch.zeros(self.per, *t.shape[1:], dtype=t.dtype, device=t.device)
        mine[: self.b - self.a].copy_(t[self.a:self.b])
        full = torch.empty(get_tp_size() * self.per, *t.shape[1:], dtype=t.dtype, device=t.device)
        dist.all_gather_into_tensor(full, mine, group=get_tp_group())
        t[self.r0:self.r1].copy_(full[: self.r1 - self.r0])   # slice k of the tile sits at rows k*per.. of `full`


def pp_recv_tiled(tensors: List[torch.Tensor], src: Optional[int] = None):
    """-> [(r0, r1, [work per tensor])]; call `w.wait()` (stream-level) before touching rows r0:r1."""
    src = get_prev_pp_rank() if src is None else src
    out = []
    for r0, r1 in pp_row_tiles(tensors[0].shape[0]):
        sh = _pp_shard_rows(r0, r1)
        if sh is None:
            PP_STATS["replicated_tiles"] += 1
            out.append((r0, r1, [dist.irecv(t[r0:r1], src) for t in tensors]))
            continue
        PP_STATS["sharded_tiles"] += 1
        a, b, per = sh
        out.append((r0, r1, [_ShardedTileWork([dist.irecv(t[a:b], src)] if b > a else [], t, r0, r1, a, b, per)
                             for t in tensors]))
    return out
