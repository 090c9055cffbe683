"""Fixed-size sample buffer with the interface of the reference's ``Buffer``
(radiocore/tools/buffer.py:10-93): ``data``, ``size``, ``dtype``, ``is_cuda``, ``is_locked``,
``consume()``.

Host plumbing only.  Where the reference asks cuSignal for CUDA managed memory, this keeps a
page-locked host array (when a GPU is present and ``cuda=True``), which is what ``Tuner.load`` /
``Tuner.submit`` can DMA from at full PCIe rate; the NumPy view handed out is the same memory.
"""
import threading
from typing import Union

import numpy as np


def _page_locked(array):
    """(owner, view): a pinned copy of ``array`` if a CUDA device is usable, else the array itself."""
    try:
        import torch
        if torch.cuda.is_available():
            pinned = torch.from_numpy(array).pin_memory()
            return pinned, pinned.numpy()
    except Exception:                      # no torch / no driver: plain pageable memory still works
        pass
    return None, array


class _Lease:
    """Context manager handing out the array, holding the buffer's mutex if it has one."""

    def __init__(self, array, guard):
        self._array, self._guard = array, guard

    def __enter__(self):
        if self._guard is not None:
            self._guard.acquire()
        return self._array

    def __exit__(self, *exc):
        if self._guard is not None:
            self._guard.release()
        return False


class Buffer:
    def __init__(self, size: Union[int, float], dtype: str = "complex64", lock: bool = False,
                 cuda: bool = False):
        self._cuda = bool(cuda)
        self._guard = threading.Lock() if lock else None
        storage = np.zeros(int(size), dtype=dtype)
        self._owner, self._array = _page_locked(storage) if self._cuda else (None, storage)

    def __len__(self) -> int:
        return self._array.shape[0]

    @property
    def size(self) -> int:
        return len(self)

    @property
    def dtype(self):
        return self._array.dtype

    @property
    def is_cuda(self) -> bool:
        return self._cuda

    @property
    def is_locked(self) -> bool:
        if self._guard is None:
            raise ValueError("locking is not enabled in this instance")
        return self._guard.locked()

    @property
    def data(self):
        """The underlying array (no locking)."""
        return self._array

    def consume(self):
        """``with buffer.consume() as array:`` -- exclusive while inside when ``lock=True``."""
        return _Lease(self._array, self._guard)
