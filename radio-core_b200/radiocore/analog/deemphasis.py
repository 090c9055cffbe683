"""FM broadcast de-emphasis (mirror of radiocore/analog/deemphasis.py:7-66)."""
import ctypes as C
from typing import Union

import numpy as np
import torch

from radiocore import _device, _native


class Deemphasis:
    """51-tap FIR image of the one-pole de-emphasis, filter state carried across calls."""

    def __init__(self, input_size: Union[int, float], rate: float = 75e-6, dtype: str = "float32",
                 cuda: bool = False):
        self._cuda = cuda
        self._dtype = dtype
        self._rate = rate
        self._input_size = int(input_size)
        self._handle = None

    @property
    def taps(self):
        """(taps[51], zi[50]) as float32 NumPy arrays (host-side design, no GPU needed)."""
        b = (C.c_float * 51)()
        z = (C.c_float * 50)()
        _native.check(_native.lib().rc_deemph_taps(self._rate, self._input_size, b, z))
        return np.array(b, dtype=np.float32), np.array(z, dtype=np.float32)

    def _native_handle(self):
        if self._handle is None:
            h = C.c_void_p()
            _native.check(_native.lib().rc_deemph_create(_device.device_index(), self._input_size,
                                                         self._rate, C.byref(h)))
            self._handle = h
        return self._handle

    def __del__(self):
        h, self._handle = getattr(self, "_handle", None), None
        if h is not None:
            try:
                _native.lib().rc_deemph_destroy(h)
            except Exception:
                pass

    def run(self, input_sig, numpy_output: bool = False):
        if len(input_sig) != self._input_size:
            raise ValueError("input_sig size and input_size mismatch")
        x = _device.to_device(input_sig, torch.float32)
        out = torch.empty_like(x)
        _native.check(_native.lib().rc_deemph_run(self._native_handle(), x.data_ptr(), out.data_ptr(),
                                                  _device.stream_ptr()))
        return _device.to_host(out) if numpy_output else out

    __call__ = run
