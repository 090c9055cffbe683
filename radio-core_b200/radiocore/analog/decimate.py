"""Fourier-method decimator (mirror of radiocore/analog/decimate.py:7-50)."""
import ctypes as C
from typing import Union

import torch

from radiocore import _device, _native


class Decimate:
    """Resample a block to ``output_size`` samples with a Hamming spectral taper."""

    def __init__(self, input_size: Union[int, float], output_size: Union[int, float], cuda: bool = False):
        self._cuda = cuda
        self._input_size = int(input_size)
        self._output_size = int(output_size)
        self._handle = None

    def _native_handle(self):
        if self._handle is None:
            h = C.c_void_p()
            _native.check(_native.lib().rc_decimate_create(_device.device_index(), self._input_size,
                                                           self._output_size, C.byref(h)))
            self._handle = h
        return self._handle

    def __del__(self):
        h, self._handle = getattr(self, "_handle", None), None
        if h is not None:
            try:
                _native.lib().rc_decimate_destroy(h)
            except Exception:
                pass

    def run(self, input_sig, numpy_output: bool = False):
        """Complex input -> complex64 output, real input -> float32 output.

        Returns a CUDA tensor (the reference returns the backend array,
        decimate.py:48-50); it converts with ``numpy.asarray(t.cpu())``.
        """
        if len(input_sig) != self._input_size:
            raise ValueError("input_sig size and input_size mismatch")
        if hasattr(input_sig, "tensor"):
            input_sig = input_sig.tensor()
        lib = _native.lib()
        if _device.is_complex_input(input_sig):
            x = _device.to_device(input_sig, torch.complex64)
            out = torch.empty(self._output_size, dtype=torch.complex64, device=x.device)
            _native.check(lib.rc_decimate_run_complex(self._native_handle(), x.data_ptr(), out.data_ptr(),
                                                      _device.stream_ptr()))
        else:
            x = _device.to_device(input_sig, torch.float32)
            out = torch.empty(self._output_size, dtype=torch.float32, device=x.device)
            _native.check(lib.rc_decimate_run_real(self._native_handle(), x.data_ptr(), out.data_ptr(),
                                                   _device.stream_ptr()))
        return _device.to_host(out) if numpy_output else out

    __call__ = run
