"""ctypes + NumPy binding of the CPU *replay* build of the kernels
(tests/native/_build/librc_emulate.so, compiled with -DRC_EMULATE).

TEST INFRASTRUCTURE ONLY: it lets the GPU-less build container run the exact
per-thread code of every kernel (serially, on host memory) against the oracle.
The product package never loads this library.
"""
import ctypes as C
import os
import subprocess

import numpy as np

ROOT = os.path.dirname(os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
BUILD = os.path.join(ROOT, "tests", "native", "_build")
LIB = os.path.join(BUILD, "librc_emulate.so")
SRC = os.path.join(ROOT, "radio-core_b200", "csrc")


def build(force=False):
    srcs = [os.path.join(SRC, f) for f in os.listdir(SRC)] + [os.path.join(ROOT, "include", "radiocore_b200.h")]
    import hashlib
    dg = hashlib.sha256()
    for s_ in sorted(srcs):
        dg.update(open(s_, "rb").read())
    stamp = LIB + ".sha256"
    if not force and os.path.exists(LIB) and os.path.exists(stamp) and open(stamp).read() == dg.hexdigest():
        return LIB
    os.makedirs(BUILD, exist_ok=True)
    import sys
    if ROOT not in sys.path:
        sys.path.insert(0, ROOT)
    import __graft_entry__ as g
    flags = ["-DRC_EMULATE", "-O2", "-std=c++17", "-Xcompiler", "-fPIC", "-Wno-deprecated-gpu-targets",
             "-I", os.path.join(ROOT, "include")]
    g.compile_units([s_ for s_ in sorted(srcs) if s_.endswith(".cu")], flags, os.path.join(BUILD, "obj"), LIB)
    open(stamp, "w").write(dg.hexdigest())
    return LIB


_lib = None


def lib():
    global _lib
    if _lib is None:
        _lib = C.CDLL(build())
        _lib.rc_last_error.restype = C.c_char_p
    return _lib


def _check(rc):
    if rc < 0:
        raise ValueError(lib().rc_last_error().decode())
    return rc


def _p(a):
    return C.c_void_p(a.ctypes.data)


def _c64(x):
    return np.ascontiguousarray(np.asarray(x), dtype=np.complex64)


def _f32(x):
    return np.ascontiguousarray(np.asarray(x), dtype=np.float32)


class _Demod:
    mode = 0
    channels = 1

    def __init__(self, input_size, output_size, deemphasis=75e-6, cuda=False):
        self._input_size, self._output_size = int(input_size), int(output_size)
        self._tau = deemphasis
        self._h = C.c_void_p()
        _check(lib().rc_demod_create(0, self.mode, C.c_int64(self._input_size), C.c_int64(self._output_size),
                                     C.c_double(deemphasis), 1, C.byref(self._h)))

    def run(self, x, numpy_output=True):
        if len(x) != self._input_size:
            raise ValueError("input_sig size and input_size mismatch")
        x = _c64(x)
        out = np.zeros((self._output_size, self.channels), dtype=np.float32)
        _check(lib().rc_demod_run(self._h, _p(x), _p(out), None))
        return out[None] if self.channels == 2 else out


class FM(_Demod):
    mode = 0


class MFM(_Demod):
    mode = 1


class WBFM(_Demod):
    mode = 2
    channels = 2


class Decimate:
    def __init__(self, input_size, output_size, cuda=False):
        self._in, self._out = int(input_size), int(output_size)
        self._h = C.c_void_p()
        _check(lib().rc_decimate_create(0, C.c_int64(self._in), C.c_int64(self._out), C.byref(self._h)))

    def run(self, x):
        if len(x) != self._in:
            raise ValueError("input_sig size and input_size mismatch")
        x = np.asarray(x)
        if np.iscomplexobj(x):
            x = _c64(x)
            out = np.zeros(self._out, dtype=np.complex64)
            _check(lib().rc_decimate_run_complex(self._h, _p(x), _p(out), None))
        else:
            x = _f32(x)
            out = np.zeros(self._out, dtype=np.float32)
            _check(lib().rc_decimate_run_real(self._h, _p(x), _p(out), None))
        return out


class Deemphasis:
    def __init__(self, input_size, rate=75e-6, dtype="float32", cuda=False):
        self._n = int(input_size)
        self._h = C.c_void_p()
        _check(lib().rc_deemph_create(0, C.c_int64(self._n), C.c_double(rate), C.byref(self._h)))

    def run(self, x):
        if len(x) != self._n:
            raise ValueError("input_sig size and input_size mismatch")
        x = _f32(x)
        out = np.zeros(self._n, dtype=np.float32)
        _check(lib().rc_deemph_run(self._h, _p(x), _p(out), None))
        return out


class Bandpass:
    def __init__(self, input_size, start_freq, stop_freq, dtype="float32", num_taps=61, window="hamm", cuda=False):
        self._n = int(input_size)
        self._h = C.c_void_p()
        _check(lib().rc_bandpass_create(0, C.c_int64(self._n), C.c_double(start_freq), C.c_double(stop_freq),
                                        int(num_taps), window.encode(), C.byref(self._h)))

    def run(self, x):
        if len(x) != self._n:
            raise ValueError("input_sig size and input_size mismatch")
        x = _f32(x)
        out = np.zeros(self._n, dtype=np.float32)
        _check(lib().rc_bandpass_run(self._h, _p(x), _p(out), None))
        return out


class PLL:
    def __init__(self, cuda=False):
        self._h = None
        self._n = 0

    def step(self, sig):
        sig = _f32(sig)
        if self._h is None or self._n != len(sig):
            self._h = C.c_void_p()
            self._n = len(sig)
            _check(lib().rc_pll_create(0, C.c_int64(self._n), C.byref(self._h)))
        _check(lib().rc_pll_step(self._h, _p(sig), None))

    def _eval(self, mult, imag):
        out = np.zeros(self._n, dtype=np.float32)
        _check(lib().rc_pll_eval(self._h, C.c_double(mult), imag, _p(out), None))
        return out

    def real(self, mult=1.0):
        return self._eval(mult, 0)

    def image(self, mult=1.0):
        return self._eval(mult, 1)


class _Channel:
    def __init__(self, index, bandwidth, demodulator, center):
        self.index, self.bandwidth, self.demodulator, self.center_frequency = index, bandwidth, demodulator, center
        self.lower_frequency, self.higher_frequency = center - bandwidth / 2, center + bandwidth / 2


class Tuner:
    """Minimal host mirror: band plan as tuner.py:163-174, arithmetic in the replay lib."""

    def __init__(self, cuda=False):
        self._bounds = []
        self.input_frequency = 0.0
        self.input_bandwidth = 0.0
        self._eng = None
        self._audio = None
        self._subband = None

    def needed_bins(self):
        n = int(self.input_bandwidth)
        return [((-(int(c.bandwidth) // 2) - int(self.input_frequency - c.center_frequency)) % n, int(c.bandwidth) + 1)
                for c in self._bounds]

    def set_subband(self, x_lo, x_len):
        self._subband = (int(x_lo), int(x_len))

    def load_subband(self, spectrum):
        """rc_engine_load_subband: the engine reads `spectrum` (kept alive here) in place."""
        if self._eng is None:
            self._commit()
        self._sub = _c64(spectrum)
        _check(lib().rc_engine_load_subband(self._eng, _p(self._sub)))

    def channels(self):
        return self._bounds

    def add_channel(self, frequency, bandwidth, demodulator):
        self._bounds.append(_Channel(len(self._bounds), bandwidth, demodulator, frequency))
        lo = min(c.lower_frequency for c in self._bounds)
        hi = max(c.higher_frequency for c in self._bounds)
        self.input_frequency = (lo + hi) / 2
        self.input_bandwidth = hi - lo
        mean_bw = sum(c.bandwidth for c in self._bounds) // len(self._bounds)
        self.input_bandwidth += (self.input_bandwidth * -1) % mean_bw

    def request_bandwidth(self, bw):
        if bw < self.input_bandwidth:
            raise ValueError("requested bandwidth is too low")
        self.input_bandwidth = bw

    def _commit(self):
        N = int(self.input_bandwidth)
        self._eng = C.c_void_p()
        _check(lib().rc_engine_create(0, C.c_int64(N), C.byref(self._eng)))
        for ch in self._bounds:
            d = ch.demodulator
            roll = int(self.input_frequency - ch.center_frequency) % N
            mode = d.mode if d is not None else 3          # RC_MODE_NONE: IQ only
            A = d._output_size if d is not None else 2
            tau = d._tau if d is not None else 75e-6
            idx = C.c_int()
            _check(lib().rc_engine_add_channel(self._eng, C.c_int64(roll), C.c_int64(int(ch.bandwidth)),
                                               C.c_int64(A), mode, C.c_double(tau), C.byref(idx)))
        if self._subband is not None:
            _check(lib().rc_engine_set_subband(self._eng, C.c_int64(self._subband[0]), C.c_int64(self._subband[1])))
        _check(lib().rc_engine_commit(self._eng))
        tot = C.c_int64()
        _check(lib().rc_engine_audio_floats(self._eng, C.byref(tot)))
        self._audio = np.zeros(tot.value, dtype=np.float32)

    def load(self, x):
        if self._eng is None:
            self._commit()
        x = _c64(x)
        if len(x) != int(self.input_bandwidth):
            raise ValueError("input size mismatch")
        _check(lib().rc_engine_load(self._eng, _p(x), None))
        self._fresh = False

    def run(self, index):
        B = int(self._bounds[index].bandwidth)
        out = np.zeros(B, dtype=np.complex64)
        _check(lib().rc_engine_channel_iq(self._eng, int(index), _p(out), None))
        return out

    def run_all(self):
        """Fused path: every channel's audio in one call (state carried in the engine)."""
        _check(lib().rc_engine_run(self._eng, _p(self._audio), None))
        res = []
        for i in range(len(self._bounds)):
            off, A, nch = C.c_int64(), C.c_int64(), C.c_int()
            _check(lib().rc_engine_channel_layout(self._eng, i, C.byref(off), C.byref(A), C.byref(nch)))
            a = self._audio[off.value: off.value + A.value * nch.value].reshape(A.value, nch.value).copy()
            res.append(a[None] if nch.value == 2 else a)
        return res


def fft(x, sign=-1):
    """rc_fft_c2c on the replay build: batched unnormalised FFT of a (batch, n) complex64 array."""
    x = _c64(np.atleast_2d(x))
    out = np.empty_like(x)
    _check(lib().rc_fft_c2c(0, C.c_int64(x.shape[1]), x.shape[0], sign, _p(x), _p(out), None))
    return out


class Fft:
    """rc_fft_* (persistent plan) on the replay build."""

    def __init__(self, n, batch=1):
        self.n, self.batch = int(n), int(batch)
        self._h = C.c_void_p()
        _check(lib().rc_fft_create(0, C.c_int64(self.n), self.batch, C.byref(self._h)))

    def __call__(self, x, sign=-1):
        x = _c64(x)
        out = np.empty_like(x)
        _check(lib().rc_fft_exec(self._h, sign, _p(x), _p(out), None))
        return out


def subband_combine(pieces, n_input, k0_base):
    """rc_subband_combine: pieces (G, P) complex64 -> bins (G, P)."""
    pieces = _c64(pieces)
    g, p = pieces.shape
    out = np.empty_like(pieces)
    _check(lib().rc_subband_combine(0, g, C.c_int64(p), C.c_int64(n_input), C.c_int64(k0_base), _p(pieces), _p(out), None))
    return out


class ScatterSeg(C.Structure):
    _fields_ = [("k1", C.c_int32), ("reserved", C.c_int32), ("j_lo", C.c_int64), ("j_hi", C.c_int64), ("dst", C.c_void_p)]


def fft_scatter(fft_plan, x, pieces):
    """rc_fft_exec_scatter: output piece p of the transform lands in the complex64 array pieces[p]."""
    x = _c64(x)
    arr = (C.c_void_p * len(pieces))(*[p.ctypes.data for p in pieces])
    _check(lib().rc_fft_exec_scatter(fft_plan._h, -1, _p(x), arr, len(pieces), C.c_int64(len(pieces[0])), None))


def subband_combine_scatter(pieces, n_input, k0_base, segs):
    """rc_subband_combine_scatter; segs: [(k1, j_lo, j_hi, destination array, first index)]."""
    pieces = _c64(pieces)
    g, p = pieces.shape
    arr = (ScatterSeg * len(segs))()
    for a, (k1, j0, j1, dst, pos) in zip(arr, segs):
        a.k1, a.reserved, a.j_lo, a.j_hi, a.dst = k1, 0, j0, j1, dst.ctypes.data + 8 * pos
    _check(lib().rc_subband_combine_scatter(0, g, C.c_int64(p), C.c_int64(n_input), C.c_int64(k0_base), _p(pieces),
                                            arr, len(segs), None))
