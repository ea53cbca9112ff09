"""Single-producer / multi-consumer broadcast ring in POSIX shared memory: the driver's per-step batch fan-out to
the other ranks of a node without sockets.

Why: the ZeroMQ ipc path costs one send per peer on the driver plus an I/O-thread hop and a poll wake-up on every
peer — measured ~300 us from `send_batch` until the slowest of 3 peers holds the batch
(benchmarks/fanout_bench.py), on the critical path of every TP / PP step because the peers cannot launch their
forward before it. Here the driver writes the message ONCE and bumps a cursor; the peers, which busy-poll anyway,
see it within a cache-line transfer.

Layout (little endian):
    [0:8)     magic | [8:16) capacity | [64:72) write cursor (bytes ever written, monotonic)
    [128 + 64*i : +8)  read cursor of consumer i (one cache line each)
    [4096 : 4096 + capacity)  data: records  u32 length | u32 kind | payload (padded to 16); a length of
                              0xFFFFFFFF means "skip to the start of the ring" (records never wrap)
Flow control: the producer never overwrites bytes some consumer has not released (`write - min(read) <= capacity`).

Memory ordering: payload bytes are stored before the cursor, cursors are single aligned 8-byte stores. That is
sufficient on x86-64 (total store order); CPython offers no fences, so on weakly ordered hosts this transport is not
selected (engine/comm.py keeps ZeroMQ unless GLLM_BATCH_TRANSPORT=shm is set on x86-64).
"""
from __future__ import annotations

import struct
import time
from multiprocessing import shared_memory
from typing import Optional, Tuple

import numpy as np

MAGIC = 0x676C6C6D5F623230     # "gllm_b20"
HDR_BYTES = 4096
SKIP = 0xFFFFFFFF
KIND_BATCH, KIND_CONTROL = 0, 1
MORE = 0x100                   # kind flag: the message continues in the next record


def _attach_untracked(name: str) -> shared_memory.SharedMemory:
    """Attach to an existing segment WITHOUT registering it with this process's resource tracker: the tracker
    would unlink the producer's segment when a consumer exits (Python < 3.13 has no `track=False`)."""
    from multiprocessing import resource_tracker
    orig = resource_tracker.register
    resource_tracker.register = lambda *a, **k: None
    try:
        return shared_memory.SharedMemory(name=name)
    finally:
        resource_tracker.register = orig


class RingWriter:
    def __init__(self, name: str, num_consumers: int, capacity: int = 32 << 20):
        assert capacity % 4096 == 0 and num_consumers >= 1
        try:    # a crashed run with the same name may have left a segment behind
            old = _attach_untracked(name)
            old.close()
            shared_memory._posixshmem.shm_unlink("/" + name)   # noqa: SLF001 (unlink without tracker traffic)
        except FileNotFoundError:
            pass
        self.shm = shared_memory.SharedMemory(name=name, create=True, size=HDR_BYTES + capacity)
        self.capacity = capacity
        self.n = num_consumers
        self.u64 = np.ndarray((HDR_BYTES // 8,), dtype=np.uint64, buffer=self.shm.buf)
        self.u64[:] = 0
        self.u64[1] = capacity
        self.data = np.ndarray((capacity,), dtype=np.uint8, buffer=self.shm.buf, offset=HDR_BYTES)
        self.write = 0
        self.min_read = 0              # cached lower bound of the consumers' cursors
        self.u64[0] = MAGIC            # last: consumers attach only to an initialised ring

    def _min_read(self) -> int:
        self.min_read = int(self.u64[16:16 + 8 * self.n:8].min())
        return self.min_read

    def _reserve(self, need: int, timeout_s: float):
        if self.write + need - self.min_read <= self.capacity:
            return          # enough room even by the last cursor values we saw: no shared-memory reads at all
        t0 = None
        while self.write + need - self._min_read() > self.capacity:
            if t0 is None:
                t0 = time.monotonic()
            elif time.monotonic() - t0 > timeout_s:
                raise TimeoutError("shm ring full: a consumer stopped reading")
            time.sleep(0)

    def send(self, payload, kind: int = KIND_BATCH, timeout_s: float = 60.0):
        """Messages larger than a quarter of the ring (pixel payloads of multimodal prefills) travel as a chain of
        records flagged MORE; the consumers re-assemble them."""
        mv = memoryview(payload).cast("B")
        limit = self.capacity // 4 - 64
        while len(mv) > limit:
            self._send_one(mv[:limit], kind | MORE, timeout_s)
            mv = mv[limit:]
        self._send_one(mv, kind, timeout_s)

    def _send_one(self, mv, kind: int, timeout_s: float):
        n = len(mv)
        rec = 8 + (n + 15) // 16 * 16
        off = self.write % self.capacity
        if off + rec > self.capacity:              # does not fit before the end: skip marker, start over
            pad = self.capacity - off
            self._reserve(pad + rec, timeout_s)
            self.data[off:off + 4] = np.frombuffer(struct.pack("<I", SKIP), dtype=np.uint8)
            self.write += pad
            off = 0
        else:
            self._reserve(rec, timeout_s)
        self.data[off + 8:off + 8 + n] = np.frombuffer(mv, dtype=np.uint8)
        self.data[off:off + 8] = np.frombuffer(struct.pack("<II", n, kind), dtype=np.uint8)
        self.write += rec
        self.u64[8] = self.write                   # publish

    def close(self, unlink: bool = True):
        self.u64 = self.data = None
        try:
            self.shm.close()
            if unlink:
                self.shm.unlink()
        except Exception:  # noqa: BLE001
            pass


class RingReader:
    """Consumer `index` (0-based). Attaches lazily: `recv()` returns None until the producer has created the ring."""

    def __init__(self, name: str, index: int):
        self.name, self.index = name, index
        self.shm: Optional[shared_memory.SharedMemory] = None
        self.read = 0
        self.parts = []                # records of a chained message received so far

    def _attach(self) -> bool:
        try:
            shm = _attach_untracked(self.name)
        except FileNotFoundError:
            return False
        u64 = np.ndarray((HDR_BYTES // 8,), dtype=np.uint64, buffer=shm.buf)
        if int(u64[0]) != MAGIC:
            del u64
            shm.close()
            return False
        self.shm, self.u64 = shm, u64
        self.capacity = int(u64[1])
        self.data = np.ndarray((self.capacity,), dtype=np.uint8, buffer=shm.buf, offset=HDR_BYTES)
        return True

    def recv(self) -> Optional[Tuple[int, bytes]]:
        """-> (kind, payload copy) or None when nothing is pending."""
        if self.shm is None and not self._attach():
            return None
        while True:
            if int(self.u64[8]) == self.read:
                return None
            off = self.read % self.capacity
            n, kind = struct.unpack_from("<II", self.data[off:off + 8].tobytes())
            if n == SKIP:
                self.read += self.capacity - off
                continue
            payload = self.data[off + 8:off + 8 + n].tobytes()     # copy out: the slot is recycled after release
            self.read += 8 + (n + 15) // 16 * 16
            self.u64[16 + 8 * self.index] = self.read              # release
            if kind & MORE:                                        # chained message: keep collecting
                self.parts.append(payload)
                continue
            if self.parts:
                payload = b"".join(self.parts) + payload
                self.parts = []
            return kind, payload

    def close(self):
        self.u64 = self.data = None
        if self.shm is not None:
            try:
                self.shm.close()
            except Exception:  # noqa: BLE001
                pass
            self.shm = None
