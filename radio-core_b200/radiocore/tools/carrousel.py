"""Ring of reusable items (mirror of radiocore/tools/carrousel.py:8-118); host plumbing."""
from contextlib import contextmanager
from typing import List


class Carrousel:
    """Single-producer ring: ``enqueue()`` hands out the next free item, ``dequeue()``
    the oldest filled one; when full the oldest item is dropped and reused."""

    def __init__(self, items: List):
        self._items = list(items)
        self._capacity = len(self._items)
        self._head = 0
        self._tail = 0
        self._occupancy = 0

    @property
    def capacity(self) -> int:
        return self._capacity

    @property
    def occupancy(self) -> int:
        return self._occupancy

    @property
    def is_full(self) -> bool:
        return self._occupancy == self._capacity

    @property
    def is_empty(self) -> bool:
        return self._occupancy == 0

    def reset(self):
        self._head = self._tail = self._occupancy = 0

    @contextmanager
    def enqueue(self):
        if self.is_full:                     # overflow: drop the oldest
            self._head = (self._head + 1) % self._capacity
            self._occupancy -= 1
        try:
            yield self._items[self._tail]
        finally:
            self._tail = (self._tail + 1) % self._capacity
            self._occupancy += 1

    @contextmanager
    def dequeue(self):
        if self.is_empty:
            raise ValueError("carrousel is empty")
        try:
            yield self._items[self._head]
        finally:
            self._head = (self._head + 1) % self._capacity
            self._occupancy -= 1
