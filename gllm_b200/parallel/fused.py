"""Fused tensor-parallel strategy (the product path for tp > 1, pp == 1): token-sharded activations,
row-parallel GEMM ⊕ reduce-scatter and all-gather ⊕ column-parallel GEMM over NVLink peer memory.

Per block (kernels: csrc/gemm/gemm_bf16.cu hooks + csrc/comm/tp_fused.cu):

    attn out ──► O-proj GEMM, epilogue stores partial tiles into the OWNER rank's staging slot (P2P st)
             ──► rs_reduce_norm on the owner: wait tile counters, sum tp partials (fixed order), + residual
                 shard, RMSNorm, push normed rows into EVERY rank's gather buffer, raise epoch flags
             ──► gate/up GEMM whose TMA producer warp waits (ld.acquire.sys) for the row shards covering
                 its M tile ... (same again for down-proj / next layer's QKV)

Symmetric buffers come from `torch.distributed._symmetric_memory` (CUDA VMM + handle exchange); all
synchronisation state (expected counters, epochs) lives in device memory so CUDA graphs can replay it.
Reference behaviour replaced: GEMM -> NCCL all_reduce -> fused_add_rms_norm on replicated activations
(gllm/layers/linear.py:247-250, gllm/dist_utils.py:253-256).
"""
from __future__ import annotations

import ctypes
from ctypes import c_float, c_int, c_int64, c_uint32, c_void_p
from typing import Optional

import torch
import torch.distributed as dist

from gllm_b200.layers import functional as Fn
from gllm_b200.ops import lib as _lib
from gllm_b200.ops.lib import MAX_PEERS, GemmComm, check, stream_ptr
from gllm_b200.parallel import state as ps
from gllm_b200.parallel.tp import TPComm
from gllm_b200.utils.logging import logger


class ReduceNormArgs(ctypes.Structure):
    _fields_ = [
        ("stage", c_void_p), ("cnt", c_void_p), ("n_tiles", c_uint32 * MAX_PEERS), ("local_x", c_void_p),
        ("local_ld", c_int64), ("residual", c_void_p), ("residual_in", c_int), ("norm_w", c_void_p),
        ("ag_peers", c_void_p * MAX_PEERS), ("flag_peers", c_void_p * MAX_PEERS), ("unnormed_out", c_void_p),
        ("st", c_void_p), ("parity", c_int), ("ag_idx", c_int), ("tp", c_int), ("rank", c_int),
        ("rows_per_rank", c_int), ("rows_valid", c_int), ("H", c_int), ("eps", c_float), ("T", c_int),
    ]


MAX_BLOCKS = 256   # kMaxBlocks in csrc/comm/tp_fused.cu (128-row blocks per gather buffer)
SMALL_T = 64       # forwards with <= this many tokens use the NCCL strategy (see begin_forward)


def _declare(L):
    L.gllm_rs_reduce_norm.argtypes = [ctypes.POINTER(ReduceNormArgs), c_void_p]
    L.gllm_rs_reduce_norm.restype = c_int
    L.gllm_push_partial_rows.argtypes = [c_void_p, c_int64, c_int, c_int, c_int, c_int, c_void_p, c_void_p, c_void_p]
    L.gllm_push_partial_rows.restype = c_int
    L.gllm_wait_ag_flags.argtypes = [c_void_p, c_void_p, c_int, c_int, c_void_p]
    L.gllm_wait_ag_flags.restype = c_int
    L.gllm_tp_state_bytes.argtypes = []
    L.gllm_tp_state_bytes.restype = c_int


class FusedTPComm(TPComm):
    fused = True

    def __init__(self, max_tokens: int, hidden_size: int, dtype=torch.bfloat16, device=None, **_):
        super().__init__()
        assert dtype == torch.bfloat16, "fused TP path is bf16"
        import torch.distributed._symmetric_memory as symm
        st = ps.get_state()
        self.group = st.tp_group
        self.device = torch.device(device)
        tp = self.tp_size
        self.H = hidden_size
        self.max_tokens = max_tokens
        self.rpr_max = (max_tokens + tp - 1) // tp
        t_pad = self.rpr_max * tp
        L = _lib.load()
        _declare(L)
        self.L = L
        # ---- one symmetric blob: [stage x2][ag x3][cnt x2][flags x3] ----
        stage_bytes = tp * self.rpr_max * hidden_size * 2
        ag_bytes = t_pad * hidden_size * 2
        self.off_stage = [0, stage_bytes]
        self.off_ag = [2 * stage_bytes + i * ag_bytes for i in range(3)]
        base_sync = 2 * stage_bytes + 3 * ag_bytes
        self.off_cnt = [base_sync, base_sync + 64]
        self.off_flag = [base_sync + 128 + 4 * MAX_BLOCKS * i for i in range(3)]
        total = base_sync + 128 + 4 * MAX_BLOCKS * 3
        assert t_pad <= 128 * MAX_BLOCKS
        total = (total + 255) // 256 * 256
        self.blob = symm.empty(total, dtype=torch.uint8, device=self.device)
        self.blob.zero_()
        self.hdl = symm.rendezvous(self.blob, self.group.group_name)
        self.peer_base = [int(p) for p in self.hdl.buffer_ptrs]
        assert len(self.peer_base) == tp
        self.local_base = self.peer_base[self.tp_rank]
        assert self.local_base == self.blob.data_ptr()
        n_state = L.gllm_tp_state_bytes()
        self.state = torch.zeros(n_state, dtype=torch.uint8, device=self.device)
        # TpState layout: rs_expected[2][8] | ag_epoch[3] | ticket[4] | pad[9] | ag_expected[3][MAX_BLOCKS] (u32)
        assert n_state == 128 + 3 * MAX_BLOCKS * 4, n_state
        self.ag_expected_ptr = [self.state.data_ptr() + 128 + 4 * MAX_BLOCKS * i for i in range(3)]
        # device tables of peer pointers for the row-push kernel
        self.stage_tbl = [torch.tensor([b + self.off_stage[p] for b in self.peer_base], dtype=torch.int64,
                                       device=self.device) for p in range(2)]
        self.cnt_tbl = [torch.tensor([b + self.off_cnt[p] for b in self.peer_base], dtype=torch.int64,
                                     device=self.device) for p in range(2)]
        self.residual_buf = torch.zeros(self.rpr_max, hidden_size, dtype=dtype, device=self.device)
        torch.cuda.synchronize()
        dist.barrier(group=self.group)
        self.T = 0
        self.rpr = 0
        self.rs_call = 0
        self.ag_call = 0
        self.small = False
        self.cur_ag = None  # (ag_idx, tensor view) produced by the last reduce_norm
        logger.info("fused TP: %d MB symmetric buffer per rank, peers mapped over NVLink", total >> 20)

    # -------------------------------------------------------------------------------------------
    def begin_forward(self, num_tokens: int):
        assert num_tokens <= self.max_tokens
        # Tiny decode batches are latency bound: the swap-AB weight-streaming GEMMs + one NCCL
        # all-reduce beat the sharded dataflow there (measured on 2xB200: 4.6 ms vs 7.0 ms per step at
        # batch 16), so such forwards run the baseline strategy; everything else runs fused.
        self.small = num_tokens <= SMALL_T
        self.T = num_tokens
        self.rpr = (num_tokens + self.tp_size - 1) // self.tp_size
        self.rs_call = 0
        self.ag_call = 0
        self.cur_ag = None

    def _rows_valid(self) -> int:
        return max(0, min(self.rpr, self.T - self.tp_rank * self.rpr))

    def _next_ag_idx(self) -> int:
        i = 2 if self.ag_call == 0 else (self.ag_call - 1) % 2
        self.ag_call += 1
        return i

    def _ag_view(self, idx: int) -> torch.Tensor:
        off = self.off_ag[idx]
        return self.blob[off: off + self.T * self.H * 2].view(torch.bfloat16).view(self.T, self.H)

    def _reduce_norm(self, parity: int, n_tiles, local_x, residual_in: bool, norm_w, eps: float):
        ag_idx = self._next_ag_idx()
        a = ReduceNormArgs()
        a.stage = self.local_base + self.off_stage[parity]
        a.cnt = self.local_base + self.off_cnt[parity]
        for s in range(self.tp_size):
            a.n_tiles[s] = n_tiles
        a.local_x = local_x.data_ptr() if local_x is not None else None
        a.local_ld = local_x.stride(0) if local_x is not None else 0
        a.residual = self.residual_buf.data_ptr()
        a.residual_in = 1 if residual_in else 0
        a.norm_w = norm_w.data_ptr()
        for p in range(self.tp_size):
            a.ag_peers[p] = self.peer_base[p] + self.off_ag[ag_idx]
            a.flag_peers[p] = self.peer_base[p] + self.off_flag[ag_idx]
        a.unnormed_out = None
        a.st = self.state.data_ptr()
        a.parity, a.ag_idx, a.tp, a.rank = parity, ag_idx, self.tp_size, self.tp_rank
        a.rows_per_rank, a.rows_valid, a.H, a.eps = self.rpr, self._rows_valid(), self.H, float(eps)
        a.T = self.T
        check(self.L.gllm_rs_reduce_norm(ctypes.byref(a), stream_ptr()), "rs_reduce_norm")
        from gllm_b200.ops import sm100
        sm100._count()
        h = self._ag_view(ag_idx)
        self.cur_ag = (ag_idx, h)
        return h, self.residual_buf[: self.rpr]

    def _ag_comm(self, x: torch.Tensor) -> Optional[GemmComm]:
        """GemmComm that gates the A operand on the gather flags, if x is the live gather buffer."""
        if self.cur_ag is None or x.data_ptr() != self.cur_ag[1].data_ptr():
            return None
        idx = self.cur_ag[0]
        c = GemmComm()
        c.a_ready = self.local_base + self.off_flag[idx]
        c.a_expected = self.ag_expected_ptr[idx]
        num_m = (self.T + 127) // 128
        c.m_rot = ((self.tp_rank * self.rpr) // 128) % max(num_m, 1)  # start with this rank's own rows
        c.rs_world = 0
        return c

    # -------------------------------------------------------------------------------------------
    def first_norm(self, x: torch.Tensor, norm_w: torch.Tensor, eps: float):
        """Embedding output (replicated) -> (normed gather buffer, residual shard)."""
        if self.small:
            return super().first_norm(x, norm_w, eps)
        return self._reduce_norm(0, 0, x, False, norm_w, eps)

    def materialize(self, h: torch.Tensor) -> torch.Tensor:
        """Make the gather buffer safe to read by a kernel that does not understand the flags."""
        if not self.small and self.cur_ag is not None and h.data_ptr() == self.cur_ag[1].data_ptr():
            idx = self.cur_ag[0]
            check(self.L.gllm_wait_ag_flags(self.local_base + self.off_flag[idx], self.state.data_ptr(), idx,
                                            self.T, stream_ptr()), "wait_ag_flags")
            from gllm_b200.ops import sm100
            sm100._count()
        return h

    def col_linear(self, x, w, bias=None):
        from gllm_b200.ops import sm100
        comm = None if self.small else self._ag_comm(x)
        if comm is None:
            return sm100.linear(x, w, bias)
        return sm100.linear(x, w, bias, comm=comm)

    def col_linear_silu_mul(self, x, w_interleaved):
        from gllm_b200.ops import sm100
        comm = None if self.small else self._ag_comm(x)
        if comm is None:
            return sm100.linear_silu_mul(x, w_interleaved)
        return sm100.linear_silu_mul(x, w_interleaved, comm=comm)

    def row_linear_add_norm(self, x, w, residual, norm_w, eps, bias=None):
        from gllm_b200.ops import sm100
        if self.small:
            return super().row_linear_add_norm(x, w, residual, norm_w, eps, bias)
        parity = self.rs_call % 2
        self.rs_call += 1
        t, n = x.shape[0], w.shape[0]
        assert t == self.T and n == self.H
        c = GemmComm()
        c.a_ready = None
        c.rs_world, c.rs_rank, c.rows_per_rank, c.rs_inc = self.tp_size, self.tp_rank, self.rpr, 1
        for p in range(self.tp_size):
            c.peer_out[p] = self.peer_base[p] + self.off_stage[parity]
            c.peer_cnt[p] = self.peer_base[p] + self.off_cnt[parity]
        # the epilogue writes into the peers' staging slots; `out` is only a shape carrier
        dummy = self.blob[self.off_stage[parity]: self.off_stage[parity] + 16].view(torch.bfloat16)
        sm100.linear(x, w, bias if self.tp_rank == 0 else None, out=_FakeOut(t, n, dummy), comm=c)
        r0 = self.tp_rank * self.rpr
        n_tiles = self.L.gllm_gemm_bf16_tiles_covering(t, n, 0, sm100._FORCE_BN, r0, min(r0 + self.rpr, t))
        return self._reduce_norm(parity, n_tiles, None, residual is not None, norm_w, eps)

    def reduce_add_norm(self, partial, residual, norm_w, eps):
        from gllm_b200.ops import sm100
        if self.small:
            return super().reduce_add_norm(partial, residual, norm_w, eps)
        parity = self.rs_call % 2
        self.rs_call += 1
        assert partial.shape[0] == self.T and partial.shape[1] == self.H and partial.stride(1) == 1
        check(self.L.gllm_push_partial_rows(partial.data_ptr(), partial.stride(0), self.T, self.H, self.tp_rank,
                                            self.rpr, self.stage_tbl[parity].data_ptr(),
                                            self.cnt_tbl[parity].data_ptr(), stream_ptr()), "push_partial_rows")
        sm100._count()
        return self._reduce_norm(parity, self._rows_valid(), None, residual is not None, norm_w, eps)

    def row_linear(self, x, w, bias=None):
        if self.small:
            return super().row_linear(x, w, bias)
        raise NotImplementedError("fused TP is used with pp_size == 1 (no un-normalised stage boundary)")


class _FakeOut:
    """Duck-typed `out=` for sm100.linear when the epilogue writes to peer memory instead."""

    def __init__(self, m, n, t):
        self.shape = (m, n)
        self._t = t

    def stride(self, i):
        return self.shape[1] if i == 0 else 1

    def data_ptr(self):
        return self._t.data_ptr()
