"""CPU oracle for the Tuner -> {FM | MFM | WBFM} receive chain.

TEST INFRASTRUCTURE ONLY.  Nothing under ``radio-core_b200/`` may import this
module; it is used by ``tests/``, by ``__graft_entry__.smoke()`` and by the
``cpu_baseline`` / ``--impl reference`` legs of ``bench.py``, always as the
checker or the timed CPU baseline, never as the product.

What it is: a float64 NumPy restatement, in closed form, of the arithmetic the
reference (luigifcruz/radio-core @ 209dc88) performs on its NumPy/SciPy path.
The reference itself only *sequences* SciPy calls (``scipy.signal.resample``,
``lfilter``, ``filtfilt``, ``hilbert``, ``firwin``, ``dimpulse`` ...), so the
algorithm restated here is that of SciPy 1.18.1 / NumPy 2.3.5 (the versions the
reference runs against in this image; the reference pins only ``scipy ^1.5``,
``numpy ^1.21`` in pyproject.toml:25-26 and has no lock file).  The only
library primitives used are FFTs (``scipy.fft``) and ``numpy`` elementwise ops.

Parity pin: the reference's own tests hold NO golden vector for this path
(SURVEY.md section 4 / 8c), so the oracle is pinned against outputs of the
reference itself run in the build container:
``tests/golden/make_golden.py`` imports ``/root/reference`` (with an in-memory
``atomics`` stub), runs it on seeded inputs and commits the outputs under
``tests/golden/``; ``tests/test_oracle_golden.py`` checks this file against
them to <= 1e-9, and ``tests/test_oracle_vs_reference.py`` does the same live
whenever ``/root/reference`` is present.

Each function cites the reference file:line it follows.
"""

from __future__ import annotations

import math
from dataclasses import dataclass
from typing import List, Optional

import numpy as np
from scipy import fft as _fft

__all__ = [
    "shifted_window", "resample_freq_two_sided", "resample_real",
    "fm_discriminator", "deemphasis_taps", "fir_zi", "fir_stateful",
    "firwin_bandpass", "filtfilt_fir", "analytic_signal",
    "Decimate", "Deemphasis", "Bandpass", "PLL", "FM", "MFM", "WBFM",
    "Channel", "Tuner",
]


# --------------------------------------------------------------------------
# windows
# --------------------------------------------------------------------------

def shifted_window(kind: str, n: int) -> np.ndarray:
    """fftshift(get_window(kind, n)) for the periodic ('fftbins') cosine windows.

    Reference: tuner.py:155-157 ('hann'), decimate.py:32-33 ('hamm').
    Periodic window w[i] = a0 - a1*cos(2*pi*i/n); fftshift moves index
    ceil(n/2).. to the front, so result[k] = w[(k + n//2) % n] for even n and
    w[(k + (n+1)//2) % n] in general (numpy.fft.fftshift shifts by n//2 to the
    right: out[(i + n//2) % n] = w[i]).
    """
    n = int(n)
    if kind in ("hann", "hanning"):
        a0, a1 = 0.5, 0.5
    elif kind in ("hamm", "hamming"):
        a0, a1 = 0.54, 0.46
    else:
        raise ValueError(f"unsupported window {kind!r}")
    if n == 1:
        return np.ones(1)
    k = np.arange(n)
    i = (k - n // 2) % n          # out[k] = w[(k - n//2) mod n]
    return a0 - a1 * np.cos(2.0 * np.pi * i / n)


# --------------------------------------------------------------------------
# Fourier-method resampling (SciPy 1.18.1 scipy.signal.resample)
# --------------------------------------------------------------------------

def resample_freq_two_sided(X: np.ndarray, num: int, W: Optional[np.ndarray]) -> np.ndarray:
    """Two-sided branch of scipy.signal.resample given the spectrum X.

    Used by Tuner.run (tuner.py:159-161, domain='freq') and by Decimate.run on
    complex input (decimate.py:48).  SciPy _signaltools.py 'else' branch.
    """
    n_x = X.shape[-1]
    num = int(num)
    m = min(num, n_x)
    m2 = m // 2 + 1
    Z = X * W if W is not None else X
    Y = np.zeros(num, dtype=Z.dtype)
    Y[:m2] = Z[:m2]
    if m2 < m:
        Y[m2 - m:] = Z[m2 - m:]
    if m % 2 == 0:
        if num < n_x:
            Y[-(m // 2)] += Z[-(m // 2)]
        elif n_x < num:
            Y[m // 2] /= 2
            Y[num - m // 2] = Y[m // 2]
    return _fft.ifft(Y * (num / n_x), n=num)


def folded_window(W: np.ndarray) -> np.ndarray:
    """One-sided window of the rfft branch: Wf[0]=W[0], Wf[l]=(W[l]+W[n-l])/2."""
    n = W.shape[0]
    n_X = n // 2 + 1
    Wf = W[:n_X].copy()
    Wf[1:n_X] = 0.5 * (W[1:n_X] + W[-1:-n_X:-1])
    return Wf


def resample_real(x: np.ndarray, num: int, W: Optional[np.ndarray]) -> np.ndarray:
    """rfft branch of scipy.signal.resample (real input, domain='time').

    Used by Decimate.run on the FM discriminator output (fm.py:66,
    wbfm.py:86-87).  Note num == n_x is NOT the identity when a window is
    given: the folded taper is still applied (wbfm.py:42-43).
    """
    n_x = x.shape[-1]
    num = int(num)
    m = min(num, n_x)
    m2 = m // 2 + 1
    X = _fft.rfft(x)
    if W is not None:
        X = X * folded_window(W)
    X = X[:m2].copy()
    if m % 2 == 0 and num != n_x:
        X[m // 2] *= 2 if num < n_x else 0.5
    return _fft.irfft(X * (num / n_x), n=num)


# --------------------------------------------------------------------------
# FM discriminator
# --------------------------------------------------------------------------

def fm_discriminator(x: np.ndarray) -> np.ndarray:
    """angle -> unwrap -> diff -> pad(1,0) -> /pi   (fm.py:60-65).

    numpy.unwrap (period 2*pi) adds to each raw difference dd the correction
    (mod(dd+pi, 2*pi)-pi) - dd when |dd| >= pi, keeping +pi when dd > 0; the
    following diff() undoes the cumsum, so d[n] is the wrapped phase step.
    """
    p = np.angle(x)
    dd = np.diff(p)
    ddmod = np.mod(dd + np.pi, 2.0 * np.pi) - np.pi
    ddmod[(ddmod == -np.pi) & (dd > 0)] = np.pi
    step = np.where(np.abs(dd) < np.pi, dd, ddmod)
    out = np.empty(p.shape[0], dtype=step.dtype)
    out[0] = 0.0
    out[1:] = step
    return out / np.pi


# --------------------------------------------------------------------------
# FIR helpers (lfilter with a = [1])
# --------------------------------------------------------------------------

def deemphasis_taps(size: int, rate: float, dtype="float32") -> np.ndarray:
    """51-tap FIR version of the one-pole de-emphasis (deemphasis.py:37-46).

    dimpulse of H(z) = (1-x)/(z-x):  h[0] = 0, h[n] = (1-x) x^(n-1), n=1..50,
    with x^(n-1) built by repeated multiplication exactly as scipy's dlsim
    state recurrence does (matters only for the float32 rounding of the taps).
    """
    x = float(np.exp(-1.0 / (int(size) * rate)))
    b = np.zeros(51)
    p = 1.0                      # dlsim state recurrence: p[n+1] = x * p[n]
    for n in range(1, 51):
        b[n] = (1.0 - x) * p
        p = x * p
    return b.astype(dtype)


def fir_zi(b: np.ndarray) -> np.ndarray:
    """scipy.signal.lfilter_zi(b, 1) (deemphasis.py:48; also inside filtfilt).

    SciPy 1.18.1 evaluates flip(cumsum(flip(b - y_inf*a)))[1:] with
    y_inf = sum(b)/sum(a) IN THE DTYPE OF THE TAPS (float32 here), i.e.
    zi[i] = sum_{k>i} b[k] accumulated from the last tap backwards.
    """
    b = np.asarray(b)
    if not np.issubdtype(b.dtype, np.floating):
        b = b.astype(np.float64)
    a = np.zeros_like(b)
    a[0] = 1
    y_inf = np.sum(b) / np.sum(a)
    return np.flip(np.cumsum(np.flip(b - y_inf * a)))[1:].copy()


def fir_stateful(b: np.ndarray, x: np.ndarray, zi: np.ndarray):
    """y, zf = lfilter(b, 1, x, zi=zi) for an FIR filter, float64 arithmetic.

    Transposed direct form II: y[n] = sum_k b[k] x[n-k] + zi[n] (n < K);
    zf[i] = sum_{k>i} b[k] x[L-(k-i)]  (+ the not-yet-flushed part of zi).
    """
    b = np.asarray(b, dtype=np.float64)
    x = np.asarray(x, dtype=np.float64)
    zi = np.asarray(zi, dtype=np.float64)
    K = b.shape[0] - 1
    L = x.shape[0]
    full = np.convolve(x, b)              # length L + K
    full[:K] += zi
    y = full[:L]
    zf = full[L:L + K].copy()
    return y, zf


def firwin_bandpass(num_taps: int, lo: float, hi: float, window: str = "hamm") -> np.ndarray:
    """scipy.signal.firwin(num_taps, [lo, hi], pass_zero=False, window=...)

    bandpass.py:50-54.  Windowed-sinc difference, symmetric (non-periodic)
    window, scaled to unit gain at the band centre.
    """
    n = int(num_taps)
    alpha = 0.5 * (n - 1)
    m = np.arange(n) - alpha
    h = hi * np.sinc(hi * m) - lo * np.sinc(lo * m)
    i = np.arange(n)
    if window in ("hamm", "hamming"):
        w = 0.54 - 0.46 * np.cos(2.0 * np.pi * i / (n - 1))
    elif window in ("hann", "hanning"):
        w = 0.5 - 0.5 * np.cos(2.0 * np.pi * i / (n - 1))
    elif window in ("boxcar", "rect"):
        w = np.ones(n)
    elif window == "blackman":
        w = (0.42 - 0.5 * np.cos(2.0 * np.pi * i / (n - 1))
             + 0.08 * np.cos(4.0 * np.pi * i / (n - 1)))
    else:
        raise ValueError(f"unsupported window {window!r}")
    h = h * w
    fc = 0.5 * (lo + hi)
    s = np.sum(h * np.cos(np.pi * m * fc))
    return h / s


def filtfilt_fir(b: np.ndarray, x: np.ndarray) -> np.ndarray:
    """scipy.signal.filtfilt(b, 1, x) with defaults (bandpass.py:72).

    padtype='odd', padlen = 3*len(b); forward pass started at zi*ext[0],
    backward pass started at zi*y[-1].
    """
    b = np.asarray(b, dtype=np.float64)
    x = np.asarray(x, dtype=np.float64)
    ntaps = b.shape[0]
    edge = 3 * ntaps
    if x.shape[0] <= edge:
        raise ValueError("The length of the input vector x must be greater than padlen")
    left = 2.0 * x[0] - x[edge:0:-1]
    right = 2.0 * x[-1] - x[-2:-(edge + 2):-1]
    ext = np.concatenate((left, x, right))
    zi = fir_zi(b)
    y, _ = fir_stateful(b, ext, zi * ext[0])
    y, _ = fir_stateful(b, y[::-1], zi * y[-1])
    return y[::-1][edge:-edge]


def analytic_signal(p: np.ndarray) -> np.ndarray:
    """scipy.signal.hilbert(p) (pll.py:34): one-sided spectrum doubling."""
    n = p.shape[0]
    P = _fft.fft(p)
    h = np.zeros(n)
    if n % 2 == 0:
        h[0] = h[n // 2] = 1.0
        h[1:n // 2] = 2.0
    else:
        h[0] = 1.0
        h[1:(n + 1) // 2] = 2.0
    return _fft.ifft(P * h)


# --------------------------------------------------------------------------
# classes mirroring the reference surface
# --------------------------------------------------------------------------

class Decimate:
    """decimate.py:21-50."""

    def __init__(self, input_size, output_size):
        self.input_size = int(input_size)
        self.output_size = int(output_size)
        self.win = shifted_window("hamm", self.input_size)

    def run(self, x):
        if len(x) != self.input_size:
            raise ValueError("input_sig size and input_size mismatch")
        x = np.asarray(x)
        if np.iscomplexobj(x):
            return resample_freq_two_sided(_fft.fft(x), self.output_size, self.win)
        return resample_real(x, self.output_size, self.win)


class Deemphasis:
    """deemphasis.py:26-66 (stateful 51-tap FIR)."""

    def __init__(self, input_size, rate=75e-6, dtype="float32"):
        self.input_size = int(input_size)
        self.taps = deemphasis_taps(self.input_size, rate, dtype)
        self.state = fir_zi(self.taps).astype(dtype)   # already `dtype`

    def run(self, x):
        if len(x) != self.input_size:
            raise ValueError("input_sig size and input_size mismatch")
        y, self.state = fir_stateful(self.taps, x, self.state)
        return y


class Bandpass:
    """bandpass.py:29-74 (firwin + zero-phase filtfilt)."""

    def __init__(self, input_size, start_freq, stop_freq, dtype="float32",
                 num_taps=61, window="hamm"):
        self.input_size = int(input_size)
        nyq = 0.5 * self.input_size
        self.taps = firwin_bandpass(int(num_taps), float(start_freq) / nyq,
                                    float(stop_freq) / nyq, window).astype(dtype)

    def run(self, x):
        if len(x) != self.input_size:
            raise ValueError("input_sig size and input_size mismatch")
        return filtfilt_fir(self.taps, x)


class PLL:
    """pll.py:19-58 (Hilbert-transform 'PLL')."""

    def __init__(self):
        self.baseline = None

    def step(self, sig):
        self.baseline = analytic_signal(np.asarray(sig, dtype=np.float64))

    def real(self, mult=1.0):
        z = self.baseline ** mult
        return np.real(z) / np.abs(z)

    def image(self, mult=1.0):
        z = self.baseline ** mult
        return np.imag(z) / np.abs(z)


class FM:
    """fm.py:26-72."""

    channels = 1

    def __init__(self, input_size, output_size, deemphasis=75e-6):
        self.input_size = int(input_size)
        self.output_size = int(output_size)
        self.decimate = Decimate(self.input_size, self.output_size)

    def run(self, x):
        if len(x) != self.input_size:
            raise ValueError("input_sig size and input_size mismatch")
        d = fm_discriminator(np.asarray(x))
        return self.decimate.run(d)[:, None]


class MFM:
    """mfm.py:29-71."""

    channels = 1

    def __init__(self, input_size, output_size, deemphasis=75e-6):
        self.fm = FM(input_size, output_size)
        self.deemph = Deemphasis(int(output_size), deemphasis)

    def run(self, x):
        a = self.fm.run(x)[:, 0]
        a = self.deemph.run(a)
        a = a - np.mean(a)
        a = np.clip(a, -0.999, 0.999)
        return a[:, None]


class WBFM:
    """wbfm.py:32-105."""

    channels = 2

    def __init__(self, input_size, output_size, deemphasis=75e-6):
        self.input_size = int(input_size)
        self.output_size = int(output_size)
        self.fm = FM(self.input_size, self.input_size)
        self.pilot = Bandpass(self.input_size, 19e3 - 50, 19e3 + 50, num_taps=41)
        self.pll = PLL()
        self.decimate = Decimate(self.input_size, self.output_size)
        self.left = Deemphasis(self.output_size, deemphasis)
        self.right = Deemphasis(self.output_size, deemphasis)

    def run(self, x):
        mpx = self.fm.run(x)[:, 0]
        self.pll.step(self.pilot.run(mpx))
        lmr = (self.pll.image(2) * mpx) * 1.0175
        l = self.decimate.run(mpx + lmr)
        r = self.decimate.run(mpx - lmr)
        l = self.left.run(l)
        r = self.right.run(r)
        lr = np.dstack((l, r))               # (1, A, 2)
        lr = lr - np.mean(lr)
        return np.clip(lr, -0.999, 0.999)


@dataclass
class Channel:
    """tuner.py:9-35."""
    index: int
    bandwidth: float
    demodulator: object
    lower_frequency: float
    center_frequency: float
    higher_frequency: float

    @property
    def address_bytes(self) -> bytes:
        return int(self.center_frequency).to_bytes(4, byteorder="little")


class Tuner:
    """tuner.py:38-174.

    ``literal=True`` performs the reference's O(N)-per-channel roll and
    full-length window multiply (what the CPU baseline times); the default
    gathers only the B+1 bins a channel keeps (same numbers, O(B)).
    """

    def __init__(self, literal: bool = False, fft_workers: Optional[int] = None):
        self.literal = literal
        self.fft_workers = fft_workers
        self._win = None
        self._buffer = None
        self.input_frequency = 0.0
        self.input_bandwidth = 0.0
        self._bounds: List[Channel] = []

    def channels(self):
        return self._bounds

    def request_bandwidth(self, bandwidth):
        if bandwidth < self.input_bandwidth:
            raise ValueError(f"requested bandwidth ({bandwidth}) is too low, "
                             f"minimum is {self.input_bandwidth}")
        self.input_bandwidth = bandwidth

    def add_channel(self, frequency, bandwidth, demodulator):
        self._bounds.append(Channel(len(self._bounds), bandwidth, demodulator,
                                    frequency - bandwidth / 2, frequency,
                                    frequency + bandwidth / 2))
        lo = min(c.lower_frequency for c in self._bounds)
        hi = max(c.higher_frequency for c in self._bounds)
        self.input_frequency = (lo + hi) / 2
        self.input_bandwidth = hi - lo
        mean_bw = sum(c.bandwidth for c in self._bounds) // len(self._bounds)
        self.input_bandwidth += (self.input_bandwidth * -1) % mean_bw

    def load(self, x):
        x = np.asarray(x)
        self._buffer = _fft.fft(x, workers=self.fft_workers)   # c64 in -> c64 FFT

    def roll_of(self, index):
        ch = self._bounds[int(index)]
        return int(self.input_frequency - ch.center_frequency)

    def run(self, index):
        ch = self._bounds[int(index)]
        r = self.roll_of(index)
        B = int(ch.bandwidth)
        X = self._buffer
        N = X.shape[0]
        if int(self.input_bandwidth) != N:
            raise ValueError("window length is not equal to number of frequency bins")
        if self.literal:
            if self._win is None:
                self._win = shifted_window("hann", N)
            return resample_freq_two_sided(np.roll(X, r), B, self._win)
        return _tuner_gather(X, r, B)


def _tuner_gather(X: np.ndarray, r: int, B: int) -> np.ndarray:
    """O(B) evaluation of roll + Hann + two-sided truncation (B < N)."""
    N = X.shape[0]
    if not B < N:
        return resample_freq_two_sided(np.roll(X, r), B, shifted_window("hann", N))
    m2 = B // 2 + 1
    k = np.concatenate((np.arange(m2), np.arange(N - (B - m2), N)))
    w = 0.5 + 0.5 * np.cos(2.0 * np.pi * k / N) if N % 2 == 0 else shifted_window("hann", N)[k]
    Y = X[(k - r) % N].astype(np.complex128) * w
    if B % 2 == 0:
        kk = N - B // 2
        wk = 0.5 + 0.5 * np.cos(2.0 * np.pi * kk / N) if N % 2 == 0 else shifted_window("hann", N)[kk]
        Y[B // 2] += complex(X[(kk - r) % N]) * wk
    return _fft.ifft(Y * (B / N), n=B)
