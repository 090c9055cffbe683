"""Property tests (hypothesis) of the pure host arithmetic of the sharded Tuner.load:
`covering_arc` really covers, is minimal up to the even-start rule, and `SubbandPlan` delivers every
bin of every rank's arc exactly once, from the rank that combines it, for arbitrary arcs."""
import numpy as np
import pytest

hypothesis = pytest.importorskip("hypothesis")
from hypothesis import given, settings, strategies as st  # noqa: E402

from radiocore.tools import sharding  # noqa: E402


@settings(max_examples=200, deadline=None)
@given(n=st.integers(64, 5000), data=st.data())
def test_covering_arc_covers_every_interval(n, data):
    k = data.draw(st.integers(1, 6))
    intervals = [(data.draw(st.integers(0, n - 1)), data.draw(st.integers(1, max(1, n // 8)))) for _ in range(k)]
    lo, length = sharding.covering_arc(intervals, n)
    assert lo % 2 == 0 and 0 <= lo < n and 1 <= length <= n
    covered = np.zeros(n, dtype=bool)
    covered[(lo + np.arange(length)) % n] = True
    for first, count in intervals:
        assert covered[(first + np.arange(count)) % n].all()
    # minimal: no arc starting at an interval's first bin is more than one bin shorter (even-start rule)
    best = min(max(((s - s0) % n) + c for s, c in intervals) for s0, _ in intervals)
    assert length <= min(best + 1, n)


@settings(max_examples=60, deadline=None)
@given(world=st.sampled_from([1, 2, 4, 8]), scale=st.integers(1, 40), data=st.data())
def test_subband_plan_delivers_every_bin_once(world, scale, data):
    n = world * world * 2 * scale
    arcs = []
    for _ in range(world):
        lo = data.draw(st.integers(0, n // 2 - 1)) * 2
        arcs.append((lo, data.draw(st.integers(1, n))))
    plan = sharding.SubbandPlan(n, world, arcs)
    assert plan.m * world == n and plan.p * world == plan.m
    for d, (lo, length) in enumerate(arcs):
        seen = np.full(length, -1, dtype=np.int64)
        for src in range(world):
            runs = plan.runs(src, d)
            assert [r[0] for r in runs] == sorted(r[0] for r in runs)                  # enumerated by k1: both ends agree on the order
            for k1, j0, j1, pos in runs:
                assert 0 <= j0 < j1 <= plan.p and 0 <= k1 < world
                bins = k1 * plan.m + src * plan.p + np.arange(j0, j1)
                assert (seen[pos: pos + (j1 - j0)] == -1).all()            # never written twice
                seen[pos: pos + (j1 - j0)] = bins
        assert np.array_equal(seen, (lo + np.arange(length)) % n)           # every bin, in arc order


def test_channel_slice_owner_roundtrip():
    for n in range(1, 40):
        for w in range(1, 10):
            owners = [sharding.owner_of(c, n, w) for c in range(n)]
            for r in range(w):
                assert [c for c in range(n) if owners[c] == r] == list(sharding.channel_slice(n, w, r))
