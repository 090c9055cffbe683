"""Seeded sweep of odd geometries through the CPU replay of the kernels against the oracle:
block / channel / audio sizes 2^a 3^b 5^c that no BASELINE configuration uses (ragged tiles,
single-pass and generic-kernel plans, channels off the bin grid, partial band occupancy), two
blocks each so the carried de-emphasis state is exercised.

Signals are kept well conditioned for a *relative* bound: the audio rate
stays above twice the test tone, so the reference's output is not near-silent (the float32
pipeline has an absolute error floor of about 1e-7 of full scale), and every channel holds a
whole station (on a channel of noise or of half a station the envelope crosses zero, the phase
steps by almost exactly pi, and the sign of that step is decided by rounding -- in the reference
too, whose channel IQ comes from a float32 FFT)."""
import numpy as np
import pytest

import radiocore_oracle as oracle
from bench_support import synth
from tests import parity


@pytest.fixture(scope="module")
def emu():
    from tests.native import emu as m
    m.lib()
    return m


def _smooth(lo, hi):
    out = set()
    a = 1
    while a <= hi:
        b = a
        while b <= hi:
            c = b
            while c <= hi:
                if c >= lo and c % 2 == 0:
                    out.add(c)
                c *= 5
            b *= 3
        a *= 2
    return sorted(out)


def _cases(count, seed):
    rng = np.random.default_rng(seed)
    Bs, As = _smooth(6000, 60000), _smooth(2400, 16000)
    cases = []
    while len(cases) < count:
        B = int(rng.choice(Bs))
        A = int(rng.choice([a for a in As if a <= B]))
        D = int(rng.choice([3, 4, 5, 6, 8]))
        C_ = int(rng.integers(1, min(D - 1, 4) + 1))
        kind = str(rng.choice(["FM", "MFM"]))
        # channels on a subset of the D - 1 tile positions, each nudged off the grid by up to B/8
        slots = sorted(rng.choice(D - 1, size=C_, replace=False).tolist())
        centers = tuple(int((k + 0.5) * B + rng.integers(-B // 8, B // 8 + 1)) for k in slots)
        cases.append((B * D, B, A, kind, centers, int(rng.integers(1 << 30))))
    return cases


@pytest.mark.parametrize("N,B,A,kind,centers,seed", _cases(10, 2026))
def test_random_geometry_replay(emu, N, B, A, kind, centers, seed):
    g, o = emu.Tuner(), oracle.Tuner()
    for c in centers:
        g.add_channel(100e6 + c, B, getattr(emu, kind)(B, A))
        o.add_channel(100e6 + c, B, getattr(oracle, kind)(B, A))
    g.request_bandwidth(N)
    o.request_bandwidth(N)
    assert g.input_frequency == o.input_frequency
    # the band plan centres the block on the registered channels (tools/tuner.py:163-174)
    offs = [100e6 + c - o.input_frequency for c in centers]
    for blk in range(2):
        x = synth.wideband(N, offs, B, seed=seed, block=blk)
        g.load(x)
        o.load(x)
        audio = g.run_all()                                    # engine path (all channels batched)
        for ch in g.channels():
            ref = o.channels()[ch.index].demodulator.run(o.run(ch.index))
            assert np.max(np.abs(ref)) > 1e-2                  # conditioning of the relative bound
            parity.assert_parity(audio[ch.index], ref, f"N={N} B={B} A={A} {kind} ch{ch.index} blk{blk}")
