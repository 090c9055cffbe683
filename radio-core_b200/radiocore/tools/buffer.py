"""Fixed-size sample buffer (mirror of radiocore/tools/buffer.py:10-93).

Host plumbing only: a pinned NumPy-visible array (so ``Tuner.load`` can DMA it
straight to the GPU) with the reference's optional mutex and ``consume()``.
"""
import threading
from contextlib import contextmanager
from typing import Union

import numpy as np


class Buffer:
    def __init__(self, size: Union[int, float], dtype: str = "complex64", lock: bool = False,
                 cuda: bool = False):
        self._lock = lock
        self._cuda = cuda
        self._size = int(size)
        if self._lock:
            self._mtx = threading.Lock()
        self._pinned = None
        self._buffer = np.zeros(self._size, dtype=dtype)
        if cuda:
            try:
                import torch
                if torch.cuda.is_available():
                    t = torch.from_numpy(self._buffer).pin_memory()
                    self._pinned, self._buffer = t, t.numpy()
            except Exception:
                pass

    @property
    def dtype(self):
        return self._buffer.dtype

    @property
    def is_cuda(self) -> bool:
        return self._cuda

    @property
    def size(self) -> int:
        return self._size

    def __len__(self) -> int:
        return self.size

    @property
    def is_locked(self) -> bool:
        if not self._lock:
            raise ValueError("locking is not enabled in this instance")
        return self._mtx.locked()

    @property
    def data(self):
        return self._buffer

    @contextmanager
    def consume(self):
        try:
            if self._lock:
                self._mtx.acquire()
            yield self._buffer
        finally:
            if self._lock:
                self._mtx.release()
