"""Single-producer single-consumer sample FIFO (mirror of radiocore/tools/ringbuffer.py:10-160).

Host plumbing between the SDR thread and the DSP thread; no arithmetic.  The
reference's ``atomics`` counter is replaced by a lock-protected integer.
"""
import threading
from typing import Union

import numpy as np


class RingBuffer:
    def __init__(self, capacity: Union[int, float], dtype: str = "complex64", cuda: bool = False,
                 print_overflow: bool = True, allow_overflow: bool = True):
        self._capacity = int(capacity)
        self._cuda = cuda
        self._print_overflow = print_overflow
        self._allow_overflow = allow_overflow
        self._buffer = np.zeros(self._capacity, dtype=dtype)
        self._occupancy = 0
        self._head = 0
        self._tail = 0
        self._guard = threading.Lock()
        self._event = threading.Event()

    def __str__(self) -> str:
        return str(self._buffer)

    @property
    def data(self):
        """The backing array (oldest sample at the read position, not at index 0)."""
        return self._buffer

    @property
    def dtype(self):
        return self._buffer.dtype

    @property
    def is_cuda(self) -> bool:
        return self._cuda

    @property
    def capacity(self) -> int:
        return self._capacity

    @property
    def occupancy(self) -> int:
        with self._guard:
            return self._occupancy

    @property
    def vacancy(self) -> int:
        return self._capacity - self.occupancy

    @property
    def is_empty(self) -> bool:
        return self.occupancy == 0

    @property
    def is_full(self) -> bool:
        return self.occupancy == self._capacity

    def reset(self):
        with self._guard:
            self._occupancy = self._head = self._tail = 0

    def put(self, arr) -> bool:
        """Append samples; on overflow reset the ring (and report it) like the reference."""
        n = len(arr)
        if n > self._capacity:
            raise ValueError("input buffer is larger than the ring capacity")
        if self.vacancy < n:
            if not self._allow_overflow:
                raise ValueError("Overflow happened.")
            if self._print_overflow:
                print("overflow")
            self.reset()
        first = min(n, self._capacity - self._tail)
        self._buffer[self._tail:self._tail + first] = arr[:first]
        if first < n:
            self._buffer[:n - first] = arr[first:]
        self._tail = (self._tail + n) % self._capacity
        with self._guard:
            self._occupancy += n
        self._event.set()
        return True

    def get(self, arr, timeout: float = 3.0) -> bool:
        """Fill ``arr`` with the oldest samples; False (None-like) on timeout."""
        n = len(arr)
        if n > self._capacity:
            raise ValueError("output buffer is larger than the ring capacity")
        while self.occupancy < n:
            self._event.clear()
            if self.occupancy >= n:
                break
            if not self._event.wait(timeout):
                return False
        first = min(n, self._capacity - self._head)
        arr[:first] = self._buffer[self._head:self._head + first]
        if first < n:
            arr[first:] = self._buffer[:n - first]
        self._head = (self._head + n) % self._capacity
        with self._guard:
            self._occupancy -= n
        return True
