"""Shared host logic of FM / MFM / WBFM: one native demodulator handle per instance."""
import ctypes as C
from typing import Union

import torch

from radiocore import _device, _native

MODE_FM, MODE_MFM, MODE_WBFM = 0, 1, 2


class ChannelView:
    """What ``Tuner.run(i)`` returns: a lazy handle on channel ``i`` of the loaded block.

    ``Tuner.load`` already knows every registered demodulator, so the engine
    runs channeliser + demodulators for all channels in batched kernels; a
    demodulator that is handed the view of its own channel just picks up its
    audio.  Used as an array (``numpy.asarray``, ``len``, ``.tensor()``) the view
    materialises the channel IQ (complex64) like the reference's Tuner.run.
    """

    def __init__(self, tuner, index, serial, size):
        self.tuner, self.index, self.serial, self.size = tuner, index, serial, size

    def __len__(self):
        return self.size

    @property
    def shape(self):
        return (self.size,)

    @property
    def dtype(self):
        import numpy as np
        return np.dtype("complex64")

    def tensor(self):
        return self.tuner._channel_iq(self.index, self.serial)

    def __array__(self, dtype=None, copy=None):
        a = _device.to_host(self.tensor())
        return a.astype(dtype) if dtype is not None else a

    def tobytes(self):
        return self.__array__().tobytes()


class DemodBase:
    _mode = MODE_FM
    _channels = 1

    def __init__(self, input_size: Union[int, float], output_size: Union[int, float],
                 deemphasis: float = 75e-6, cuda: bool = False):
        self._cuda = cuda            # accepted for signature compatibility; always GPU
        self._input_size = int(input_size)
        self._output_size = int(output_size)
        self._deemphasis_rate = float(deemphasis)
        self._handle = None

    @property
    def channels(self):
        """Return the number of audio channels of the output."""
        return self._channels

    def _native_handle(self):
        if self._handle is None:
            h = C.c_void_p()
            _native.check(_native.lib().rc_demod_create(
                _device.device_index(), self._mode, self._input_size, self._output_size,
                self._deemphasis_rate, 1, C.byref(h)))
            self._handle = h
        return self._handle

    def __del__(self):
        h, self._handle = getattr(self, "_handle", None), None
        if h is not None:
            try:
                _native.lib().rc_demod_destroy(h)
            except Exception:
                pass

    def _shape(self, flat):
        a = flat.view(self._output_size, self._channels)
        return a.unsqueeze(0) if self._channels == 2 else a

    def run(self, input_sig, numpy_output: bool = True):
        """Demodulate one block; returns float32 audio shaped like the reference's
        ((A, 1) for FM/MFM, (1, A, 2) for WBFM)."""
        if len(input_sig) != self._input_size:
            raise ValueError("input_sig size and input_size mismatch")
        if isinstance(input_sig, ChannelView):
            got = input_sig.tuner._audio_for(input_sig, self, numpy_output)
            if got is not None:
                return got
            input_sig = input_sig.tensor()
        x = _device.to_device(input_sig, torch.complex64)
        out = torch.empty(self._output_size * self._channels, dtype=torch.float32, device=x.device)
        _native.check(_native.lib().rc_demod_run(self._native_handle(), x.data_ptr(), out.data_ptr(),
                                                 _device.stream_ptr()))
        out = self._shape(out)
        return _device.to_host(out) if numpy_output else out

    __call__ = run
