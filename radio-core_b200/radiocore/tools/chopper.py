"""Fixed-size slicer (mirror of radiocore/tools/chopper.py:4-55); host plumbing."""


class Chopper:
    """Iterate over consecutive ``chunk_size`` slices of a ``size``-long buffer."""

    def __init__(self, size, chunk_size):
        self._size = int(size)
        self._chunk = int(chunk_size)
        if self._chunk <= 0 or self._size % self._chunk:
            raise ValueError("size must be a positive multiple of chunk_size")

    def chop(self, buffer):
        if len(buffer) != self._size:
            raise ValueError("buffer size mismatch")
        for start in range(0, self._size, self._chunk):
            yield buffer[start:start + self._chunk]
