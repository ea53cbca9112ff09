"""Free-list of integer ids with O(1) *specific-id* acquisition.

The reference's deque-based allocator does an O(n) `in` + `remove` whenever the prefix cache
re-acquires a particular page (gllm/id_allocator.py:9-16). Here the free list is an intrusive
doubly-linked list over two int arrays, so `allocate()`, `allocate(id)` and `free(id)` are all
O(1) while keeping the same policy: allocate from the head, free to the **tail** (freed pages
— and their cached prefix hashes — survive as long as possible).
"""
from __future__ import annotations


class IDAllocator:
    __slots__ = ("start", "size", "_prev", "_next", "_free", "_head", "_tail", "_num_free")

    def __init__(self, start_num: int = 0, end_num: int = 999):
        self.start = start_num
        self.size = end_num - start_num + 1
        n = self.size
        self._prev = list(range(-1, n - 1))
        self._next = list(range(1, n + 1))
        if n:
            self._next[-1] = -1
        self._free = [True] * n
        self._head = 0 if n else -1
        self._tail = n - 1
        self._num_free = n

    # -- internal ---------------------------------------------------------------------------
    def _unlink(self, i: int):
        p, nx = self._prev[i], self._next[i]
        if p >= 0:
            self._next[p] = nx
        else:
            self._head = nx
        if nx >= 0:
            self._prev[nx] = p
        else:
            self._tail = p
        self._free[i] = False
        self._num_free -= 1

    # -- public -----------------------------------------------------------------------------
    def allocate(self, id: int | None = None) -> int:
        if id is None:
            if self._head < 0:
                raise RuntimeError("IDAllocator exhausted")
            i = self._head
            self._unlink(i)
            return i + self.start
        i = id - self.start
        if self._free[i]:
            self._unlink(i)
        return id

    def free(self, id: int):
        i = id - self.start
        if self._free[i]:
            raise RuntimeError(f"double free of id {id}")
        self._free[i] = True
        self._prev[i] = self._tail
        self._next[i] = -1
        if self._tail >= 0:
            self._next[self._tail] = i
        else:
            self._head = i
        self._tail = i
        self._num_free += 1

    def is_free(self, id: int) -> bool:
        return self._free[id - self.start]

    def get_num_used_ids(self) -> int:
        return self.size - self._num_free

    def get_num_free_ids(self) -> int:
        return self._num_free
