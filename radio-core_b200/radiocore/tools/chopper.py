"""Fixed-size slicer with the interface of the reference's ``Chopper``
(radiocore/tools/chopper.py:4-55): views of consecutive chunks of a larger array; host plumbing."""
from typing import Union


class Chopper:
    def __init__(self, size: Union[int, float], chunk_size: Union[int, float]):
        self._total, self._step = int(size), int(chunk_size)
        if self._step <= 0 or self._total % self._step:
            raise ValueError(f"cannot evenly divide array by chunk size ({self._total}, {self._step})")

    @property
    def size(self) -> int:
        return self._total

    @property
    def chunk_size(self) -> int:
        return self._step

    def chop(self, input_arr):
        """Yield ``size // chunk_size`` consecutive views (no copies) of ``input_arr``."""
        for begin in range(0, self._total, self._step):
            yield input_arr[begin:begin + self._step]
