"""Seeded sweep of random band plans on the GPU kernels against the oracle (SURVEY.md section 4:
"hypothesis-style random band plans"): block / channel / audio sizes 2^a 3^b 5^c the BASELINE
configurations do not use (ragged tiles, TMA and per-thread gathers, generic-kernel plans),
channels off the bin grid and on a subset of the band, demodulator kinds MIXED inside one tuner
(several banks per engine), two blocks each (carried de-emphasis state).  Conditioning rules as in
tests/test_replay_fuzz.py."""
import numpy as np
import pytest

import radiocore_oracle as oracle
from bench_support import synth
from tests import parity
from tests.test_replay_fuzz import _smooth

pytestmark = pytest.mark.gpu


def _cases(count, seed):
    rng = np.random.default_rng(seed)
    Bs, As = _smooth(40_000, 400_000), _smooth(8_000, 48_000)
    cases = []
    while len(cases) < count:
        B = int(rng.choice(Bs))
        A = int(rng.choice([a for a in As if a <= B // 2]))
        D = int(rng.choice([4, 5, 6, 8, 10, 12]))
        C_ = int(rng.integers(2, min(D - 1, 6) + 1))
        slots = sorted(rng.choice(D - 1, size=C_, replace=False).tolist())
        centers = tuple(int((k + 0.5) * B + rng.integers(-B // 8, B // 8 + 1)) for k in slots)
        kinds = tuple(str(rng.choice(["FM", "MFM", "WBFM"] if B >= 200_000 else ["FM", "MFM"])) for _ in slots)
        cases.append((B * D, B, A, kinds, centers, int(rng.integers(1 << 30))))
    return cases


@pytest.mark.parametrize("N,B,A,kinds,centers,seed", _cases(12, 4242))
def test_random_band_plan_gpu(N, B, A, kinds, centers, seed):
    import radiocore as rc
    g, o = rc.Tuner(cuda=True), oracle.Tuner()
    for c, kind in zip(centers, kinds):
        g.add_channel(100e6 + c, B, getattr(rc, kind)(B, A, cuda=True))
        o.add_channel(100e6 + c, B, getattr(oracle, kind)(B, A))
    g.request_bandwidth(N)
    o.request_bandwidth(N)
    assert g.input_frequency == o.input_frequency and g.input_bandwidth == o.input_bandwidth
    offs = [100e6 + c - o.input_frequency for c in centers]
    stereo = "WBFM" in kinds
    for blk in range(2):
        x = synth.wideband(N, offs, B, seed=seed, block=blk, stereo=stereo, deviation=min(0.3 * B, 75e3))
        g.load(x)
        o.load(x)
        for ch in g.channels():
            got = ch.demodulator.run(g.run(ch.index))
            ref = o.channels()[ch.index].demodulator.run(o.run(ch.index))
            assert got.shape == ref.shape
            assert np.max(np.abs(ref)) > 1e-2                  # conditioning of the relative bound
            parity.assert_parity(got, ref, f"N={N} B={B} A={A} {kinds[ch.index]} ch{ch.index} blk{blk}")
