"""Live check of the oracle against the real reference; build container only
(/root/reference is absent on the GPU box, where this file skips)."""
import numpy as np
import pytest

import ref_shim
import radiocore_oracle as oracle
from tests.golden import cases

pytestmark = pytest.mark.skipif(not ref_shim.available(), reason="reference checkout not present")


@pytest.mark.parametrize("name", list(cases.CASES))
def test_case_live(name):
    ref = ref_shim.load_reference()
    a = cases.CASES[name](ref, np)
    b = cases.CASES[name](oracle, np)
    assert a.keys() == b.keys()
    for k in a:
        scale = max(float(np.max(np.abs(a[k]))), 1e-300)
        assert float(np.max(np.abs(a[k] - b[k]))) / scale <= 1e-9, (name, k)


@pytest.mark.parametrize("size,rate", [(48000, 75e-6), (48000, 50e-6), (32000, 75e-6), (1500, 2.4e-3)])
def test_deemphasis_taps_bit_exact(size, rate):
    ref = ref_shim.load_reference()
    a, b = oracle.Deemphasis(size, rate), ref.Deemphasis(size, rate)
    assert np.array_equal(a.taps, b._taps[0])
    assert np.array_equal(a.state, b._state)


def test_bandpass_taps_bit_exact():
    ref = ref_shim.load_reference()
    for n, lo, hi, k in ((250000, 18950, 19050, 41), (240000, 18950, 19050, 41), (5000, 300, 900, 61)):
        a = oracle.Bandpass(n, lo, hi, num_taps=k)
        b = ref.Bandpass(n, lo, hi, num_taps=k)
        assert np.allclose(a.taps, b._taps[0], rtol=0, atol=1e-9)


def test_band_plan_matches():
    ref = ref_shim.load_reference()
    for freqs, bw in (((96.9e6, 94.5e6, 97.5e6), 240e3), ((100e6,), 250e3), ((1e6, 1.3e6), 200e3)):
        a, b = oracle.Tuner(), ref.Tuner()
        for f in freqs:
            a.add_channel(f, bw, None)
            b.add_channel(f, bw, None)
        assert a.input_frequency == b.input_frequency
        assert a.input_bandwidth == b.input_bandwidth
