"""Zero-phase band-pass filter (mirror of radiocore/analog/bandpass.py:7-74)."""
import ctypes as C
from typing import Union

import numpy as np
import torch

from radiocore import _device, _native


class Bandpass:
    """Windowed-sinc band-pass taps applied forward and backward (filtfilt)."""

    def __init__(self, input_size: Union[int, float], start_freq: Union[int, float],
                 stop_freq: Union[int, float], dtype: str = "float32", num_taps: int = 61,
                 window: str = "hamm", cuda: bool = False):
        self._cuda = cuda
        self._dtype = dtype
        self._window = window
        self._num_taps = int(num_taps)
        self._input_size = int(input_size)
        self._stop_freq = float(stop_freq)
        self._start_freq = float(start_freq)
        self._handle = None

    def _native_handle(self):
        if self._handle is None:
            h = C.c_void_p()
            _native.check(_native.lib().rc_bandpass_create(
                _device.device_index(), self._input_size, self._start_freq, self._stop_freq,
                self._num_taps, self._window.encode(), C.byref(h)))
            self._handle = h
        return self._handle

    @property
    def taps(self):
        buf = (C.c_float * self._num_taps)()
        _native.check(_native.lib().rc_bandpass_taps(self._native_handle(), buf, self._num_taps))
        return np.array(buf, dtype=np.float32)

    def __del__(self):
        h, self._handle = getattr(self, "_handle", None), None
        if h is not None:
            try:
                _native.lib().rc_bandpass_destroy(h)
            except Exception:
                pass

    def run(self, input_sig, numpy_output: bool = False):
        if len(input_sig) != self._input_size:
            raise ValueError("input_sig size and input_size mismatch")
        x = _device.to_device(input_sig, torch.float32)
        out = torch.empty_like(x)
        _native.check(_native.lib().rc_bandpass_run(self._native_handle(), x.data_ptr(), out.data_ptr(),
                                                    _device.stream_ptr()))
        return _device.to_host(out) if numpy_output else out

    __call__ = run
