"""Ring of reusable items with the interface of the reference's ``Carrousel``
(radiocore/tools/carrousel.py:8-118): a queue that never frees what it hands out, so device or
pinned buffers are allocated once.  Single producer, single consumer; host plumbing.

State is the index of the oldest filled item and the number of filled items; the write position
is derived.  Writing into a full ring overwrites the oldest item and counts an overflow
(the reference's tests/test_carrousel.py pins that behaviour)."""
from typing import List

from radiocore.tools.buffer import Buffer


class _Slot:
    """Context manager around one item; ``done`` runs when the block exits, even on error."""

    def __init__(self, item, done):
        self._item, self._done, self._inner = item, done, None

    def __enter__(self):
        if isinstance(self._item, Buffer):           # Buffers are entered through their own lease
            self._inner = self._item.consume()
            return self._inner.__enter__()
        return self._item

    def __exit__(self, *exc):
        try:
            if self._inner is not None:
                self._inner.__exit__(*exc)
        finally:
            self._done()
        return False


class Carrousel:
    def __init__(self, items: List, print_overflow: bool = True):
        self._items = items
        self._oldest = 0
        self._filled = 0
        self._overflows = 0
        self._report = bool(print_overflow)

    def __str__(self):
        return str(self._items)

    @property
    def capacity(self) -> int:
        return len(self._items)

    @property
    def occupancy(self) -> int:
        return self._filled

    @property
    def overflow(self) -> int:
        """Number of items overwritten before they were read."""
        return self._overflows

    @property
    def is_empty(self) -> bool:
        return self._filled == 0

    @property
    def is_full(self) -> bool:
        return self._filled >= self.capacity

    @property
    def is_healthy(self) -> bool:
        """At least one item is ready to be read."""
        return self._filled >= 1

    def reset(self):
        self._oldest = self._filled = 0

    def enqueue(self):
        """``with ring.enqueue() as item:`` -- the next item to fill."""
        if self.is_full:
            self._overflows += 1
            self._drop_oldest()
            if self._report:
                print("overflow")
        return _Slot(self._items[(self._oldest + self._filled) % self.capacity], self._filled_one)

    def dequeue(self):
        """``with ring.dequeue() as item:`` -- the oldest filled item."""
        if self.is_empty:
            raise ValueError("carrousel is empty")
        return _Slot(self._items[self._oldest], self._drop_oldest)

    def _filled_one(self):
        self._filled += 1

    def _drop_oldest(self):
        self._oldest = (self._oldest + 1) % self.capacity
        self._filled -= 1
