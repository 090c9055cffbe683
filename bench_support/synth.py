"""Deterministic synthetic wideband IQ for the Tuner -> FM/MFM/WBFM chain.

SURVEY.md section 8(d): every channel carries a unit-amplitude FM station
(mono tone or stereo multiplex), the stations are summed, complex AWGN is
added, the sum is scaled by 1/sqrt(C) and cast to complex64.  All phases are
closed-form integrals of the modulating tones, so consecutive one-second
blocks are phase continuous (``block`` selects the second) and the generator
can be evaluated in chunks.

Frequencies are in "bin units": the block of N samples spans one second, so a
channel of B bins is B Hz wide.  ``deviation`` defaults to 0.3*B (75 kHz in a
250 kHz channel).
"""
from __future__ import annotations

import numpy as np

TWO_PI = 2.0 * np.pi


def mono_phase(t, chan_index, deviation):
    """Phase (radians) 2*pi*dev*integral of 0.5*sin(2*pi*fm*t), fm = 300 + 50*c Hz."""
    fm = 300.0 + 50.0 * (chan_index % 64)
    return (deviation * 0.5 / fm) * (1.0 - np.cos(TWO_PI * fm * t))


def stereo_phase(t, chan_index, deviation):
    """Integrated stereo multiplex 0.45(L+R) + 0.1 pilot + 0.45(L-R) sin(2*pi*38k t)."""
    fl, fr, fp = 1000.0 + 10.0 * (chan_index % 16), 2500.0, 19000.0

    def isin(f):                       # integral of sin(2 pi f t)
        return (1.0 - np.cos(TWO_PI * f * t)) / (TWO_PI * f)

    def isinsin(f, g):                 # integral of sin(2 pi f t) sin(2 pi g t)
        return 0.5 * (np.sin(TWO_PI * (f - g) * t) / (TWO_PI * (f - g))
                      - np.sin(TWO_PI * (f + g) * t) / (TWO_PI * (f + g)))

    lpr = 0.8 * isin(fl) + 0.8 * isin(fr)
    lmr = 0.8 * isinsin(fl, 2 * fp) - 0.8 * isinsin(fr, 2 * fp)
    return TWO_PI * deviation * (0.45 * lpr + 0.10 * isin(fp) + 0.45 * lmr)


def station(n_samples, rate, chan_index, offset_hz=0.0, deviation=None, stereo=False,
            block=0, phase0=None, start=0, count=None):
    """Complex128 FM station sampled at ``rate`` Hz, samples [start, start+count)."""
    count = n_samples if count is None else count
    n = np.arange(start, start + count, dtype=np.float64)
    t = (block * float(n_samples) + n) / float(rate)
    if stereo:
        ph = stereo_phase(t, chan_index, deviation)
    else:
        ph = mono_phase(t, chan_index, deviation)
    if phase0 is None:
        phase0 = 0.61803398875 * chan_index
    # carrier offset: exact modular phase to stay accurate at large offsets
    cyc = np.mod(offset_hz * t, 1.0)
    return np.exp(1j * (ph + TWO_PI * cyc + phase0))


def wideband(N, centers_hz, bandwidth, seed, stereo=False, noise_sigma=0.05,
             block=0, deviation=None):
    """Sum of stations at offsets ``centers_hz`` (relative to the tuner centre).

    Returns complex64[N].  O(C*N): meant for the small/medium parity sizes.
    """
    rng = np.random.default_rng([seed, block])
    C = len(centers_hz)
    dev = 0.3 * bandwidth if deviation is None else deviation
    x = np.zeros(N, dtype=np.complex128)
    for c, off in enumerate(centers_hz):
        x += station(N, N, c, offset_hz=off, deviation=dev, stereo=stereo, block=block)
    x += noise_sigma * (rng.standard_normal(N) + 1j * rng.standard_normal(N))
    x /= np.sqrt(C)
    return x.astype(np.complex64)


def tiling_centers(N, C, B):
    """Channel-centre offsets (Hz, relative to band centre) tiling the band."""
    return [-(C * B) / 2.0 + B / 2.0 + c * B for c in range(C)]
