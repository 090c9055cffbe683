"""Hilbert-transform 'PLL' (mirror of radiocore/analog/pll.py:6-58)."""
import ctypes as C

import torch

from radiocore import _device, _native


class PLL:
    """Phase reference from the analytic signal of the input; ``real``/``image``
    return cos / sin of ``mult`` times its phase."""

    def __init__(self, cuda: bool = False):
        self._cuda = cuda
        self._handle = None
        self._size = 0

    def _drop(self):
        h, self._handle = getattr(self, "_handle", None), None
        if h is not None:
            try:
                _native.lib().rc_pll_destroy(h)
            except Exception:
                pass

    __del__ = _drop

    def step(self, input_sig):
        x = _device.to_device(input_sig, torch.float32)
        if self._handle is None or self._size != x.numel():
            self._drop()
            h = C.c_void_p()
            _native.check(_native.lib().rc_pll_create(_device.device_index(), x.numel(), C.byref(h)))
            self._handle, self._size = h, x.numel()
        _native.check(_native.lib().rc_pll_step(self._handle, x.data_ptr(), _device.stream_ptr()))
        torch.cuda.current_stream().synchronize()   # x may be a temporary; the kernels read it

    def _eval(self, mult, imag, numpy_output):
        if self._handle is None:
            raise RuntimeError("PLL.step must be called first")
        out = torch.empty(self._size, dtype=torch.float32, device="cuda")
        _native.check(_native.lib().rc_pll_eval(self._handle, float(mult), imag, out.data_ptr(),
                                                _device.stream_ptr()))
        return _device.to_host(out) if numpy_output else out

    def real(self, mult: float = 1.0, numpy_output: bool = False):
        return self._eval(mult, 0, numpy_output)

    def image(self, mult: float = 1.0, numpy_output: bool = False):
        return self._eval(mult, 1, numpy_output)
