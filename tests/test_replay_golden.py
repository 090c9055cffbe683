"""CPU replay of the CUDA kernels' per-thread code (built with -DRC_EMULATE)
against the golden outputs of the real reference.  Validates the index
arithmetic and numerics of every kernel in the GPU-less build container; the
GPU tests (-m gpu) repeat the same scenarios on the real kernels."""
import os
import shutil

import numpy as np
import pytest

from tests import parity
from tests.golden import cases

pytestmark = pytest.mark.skipif(shutil.which("nvcc") is None, reason="nvcc needed to build the replay library")

GOLDEN = np.load(os.path.join(os.path.dirname(__file__), "golden", "golden.npz"))



@pytest.fixture(scope="module")
def emu():
    from tests.native import emu as m
    m.build()
    return m


@pytest.mark.parametrize("name", list(cases.CASES))
def test_replay_matches_reference(emu, name):
    res = cases.CASES[name](emu, np)
    for key, val in res.items():
        ref = GOLDEN[f"{name}/{key}"]
        if key == "f_in":
            assert np.array_equal(ref, val)
            continue
        parity.assert_parity(val, ref, f"{name}/{key}", tol_scale=parity.LOOSE.get(f"{name}/{key}", 1.0))


def test_fused_engine_equals_per_channel(emu):
    """rc_engine_run (batched, fused) == Tuner.run + demod.run per channel."""
    from bench_support import synth
    N, B, A, C = 80000, 20000, 4000, 4
    offs = synth.tiling_centers(N, C, B)
    t = emu.Tuner()
    solo = [emu.MFM(B, A) for _ in offs]
    for off in offs:
        t.add_channel(1e8 + off, B, emu.MFM(B, A))
    t.request_bandwidth(N)
    for blk in range(2):
        t.load(synth.wideband(N, offs, B, seed=42, block=blk))
        fused = t.run_all()
        for i in range(C):
            ref = GOLDEN[f"tuner_mfm/b{blk}c{i}"]
            parity.assert_parity(cases.pin(fused[i].astype(np.float64)), ref, f"fused b{blk}c{i}")
            one = solo[i].run(t.run(i))
            assert np.max(np.abs(one - fused[i])) <= 2e-6


def test_mixed_banks_and_iq_only(emu):
    """Channels of different kinds share one engine; a channel without demodulator yields IQ only."""
    import radiocore_oracle as oracle
    from bench_support import synth
    N, B = 256000, 64000
    offs = [-96000.0, -32000.0, 32000.0, 96000.0]
    kinds = [("WBFM", 16000), ("MFM", 16000), ("FM", 8000), (None, 0)]
    te, to = emu.Tuner(), oracle.Tuner()
    for off, (kind, A) in zip(offs, kinds):
        te.add_channel(1e8 + off, B, getattr(emu, kind)(B, A) if kind else None)
        to.add_channel(1e8 + off, B, getattr(oracle, kind)(B, A) if kind else None)
    te.request_bandwidth(N)
    to.request_bandwidth(N)
    x = sum(synth.station(N, N, c, offset_hz=off, deviation=15000.0, stereo=(c == 0)) for c, off in enumerate(offs))
    x = (x / 2).astype(np.complex64)
    te.load(x)
    to.load(x)
    fused = te.run_all()
    for i, (kind, A) in enumerate(kinds):
        iq_ref = to.run(i)
        parity.assert_parity(te.run(i), iq_ref, f"iq{i}")
        if kind:
            parity.assert_parity(fused[i], to.channels()[i].demodulator.run(iq_ref), f"audio{i} {kind}")
        else:
            assert fused[i].size == 0


def test_error_codes(emu):
    with pytest.raises(ValueError):
        emu.FM(1000, 100).run(np.zeros(999, dtype=np.complex64))
    with pytest.raises(ValueError):
        emu.FM(2 * 7 * 11, 14)                       # 7 and 11 are not supported factors
    with pytest.raises(ValueError):
        emu.WBFM(30000, 6000)                        # 19 kHz pilot above Nyquist
    with pytest.raises(ValueError):
        emu.Bandpass(100, 10, 20, num_taps=61).run(np.zeros(100))   # shorter than padlen


def test_wide_tile_schedules_replay(emu, monkeypatch):
    """The 64-column schedule of the R = 50 later pass (500 000 = 200 x 50 x 50, ragged last tile:
    2500 columns are not a multiple of 64) against the 32-column one and against numpy."""
    n = 500_000
    rng = np.random.default_rng(64)
    x = (rng.standard_normal((2, n)) + 1j * rng.standard_normal((2, n))).astype(np.complex64)
    wide = emu.fft(x, +1)
    monkeypatch.setenv("RC_NO_WIDE64", "1")
    narrow = emu.fft(x, +1)
    assert np.array_equal(wide, narrow)               # same arithmetic, different tiling
    ref = np.fft.ifft(x.astype(np.complex128), axis=1) * n
    assert np.max(np.abs(wide - ref)) / np.sqrt(np.mean(np.abs(ref) ** 2)) < 3e-6


@pytest.mark.parametrize("world", [2, 4, 8])
def test_sharded_load_replay(emu, world):
    """The native pieces of the sharded Tuner.load (rc_fft_exec on the commutator branches,
    rc_subband_combine, an engine in sub-band mode) with the exchanges done by array slicing in one
    process: every virtual rank's audio against the oracle, and its sub-band against a float64 FFT."""
    import radiocore_oracle as oracle
    from bench_support import synth
    from radiocore.tools import sharding
    N, B, A, C_ = 128_000, 16_000, 3_200, 8
    offs = synth.tiling_centers(N, C_, B)
    x = synth.wideband(N, offs, B, seed=77)
    o = oracle.Tuner()
    for off in offs:
        o.add_channel(1e8 + off, B, oracle.MFM(B, A))
    o.request_bandwidth(N)
    o.load(x)
    X64 = np.fft.fft(x.astype(np.complex128))
    tuners, arcs = [], []
    for r in range(world):
        t = emu.Tuner()
        sharding.shard_tuner(t, [1e8 + f for f in offs], B, lambda c: emu.MFM(B, A), 1e8, N, world, r)
        arcs.append(sharding.covering_arc(t.needed_bins(), N))
        t.set_subband(*arcs[-1])
        tuners.append(t)
    plan = sharding.SubbandPlan(N, world, arcs)
    fft = emu.Fft(plan.m)
    F = [fft(x[g::world]) for g in range(world)]                                   # local transforms
    Y = [emu.subband_combine(np.stack([F[g][p * plan.p:(p + 1) * plan.p] for g in range(world)]), N, p * plan.p)
         for p in range(world)]                                                    # exchange 1 + combine
    for d in range(world):
        sub = np.zeros(arcs[d][1] + 64, dtype=np.complex64)
        for src in range(world):                                                   # exchange 2
            for k1, j0, j1, pos in plan.runs(src, d):
                sub[pos: pos + (j1 - j0)] = Y[src][k1, j0:j1]
        want = X64[(arcs[d][0] + np.arange(arcs[d][1])) % N]
        assert np.max(np.abs(sub[:arcs[d][1]] - want)) <= 3e-6 * np.sqrt(np.mean(np.abs(X64) ** 2)) * 10
        tuners[d].load_subband(sub)
        audio = tuners[d].run_all()
        mine = list(sharding.channel_slice(C_, world, d))
        for i, c in enumerate(mine):
            ref = o.channels()[c].demodulator.run(o.run(c))
            parity.assert_parity(audio[i], ref, f"world {world} rank {d} ch {c}")


@pytest.mark.parametrize("world", [2, 8])
def test_sharded_load_fused_stores_replay(emu, world):
    """The fused variants -- rc_fft_exec_scatter (last pass stores each piece into its combiner's
    buffer) and rc_subband_combine_scatter (the combine stores every bin into the sub-bands that read
    it) -- against the plain kernels + array-slicing exchanges: bit-identical sub-bands."""
    from bench_support import synth
    from radiocore.tools import sharding
    N, B, C_ = 128_000, 16_000, 8
    offs = synth.tiling_centers(N, C_, B)
    x = synth.wideband(N, offs, B, seed=78)
    arcs = []
    for r in range(world):
        t = emu.Tuner()
        sharding.shard_tuner(t, [1e8 + f for f in offs], B, lambda c: None, 1e8, N, world, r)
        arcs.append(sharding.covering_arc(t.needed_bins(), N))
    plan = sharding.SubbandPlan(N, world, arcs)
    fft = emu.Fft(plan.m)
    # plain
    F = [fft(x[g::world]) for g in range(world)]
    Y = [emu.subband_combine(np.stack([F[g][p * plan.p:(p + 1) * plan.p] for g in range(world)]), N, p * plan.p)
         for p in range(world)]
    want = []
    for d in range(world):
        sub = np.zeros(arcs[d][1], dtype=np.complex64)
        for src in range(world):
            for k1, j0, j1, pos in plan.runs(src, d):
                sub[pos: pos + (j1 - j0)] = Y[src][k1, j0:j1]
        want.append(sub)
    # fused: R[p] is rank p's [G][P] receive buffer, subs[d] rank d's sub-band
    R = [np.zeros((world, plan.p), dtype=np.complex64) for _ in range(world)]
    subs = [np.zeros(arcs[d][1], dtype=np.complex64) for d in range(world)]
    for g in range(world):
        emu.fft_scatter(fft, x[g::world], [R[p][g] for p in range(world)])
    for p in range(world):
        segs = [(k1, j0, j1, subs[d], pos) for d in range(world) for k1, j0, j1, pos in plan.runs(p, d)]
        emu.subband_combine_scatter(R[p], N, p * plan.p, segs)
    for d in range(world):
        assert np.array_equal(subs[d], want[d]), d


@pytest.mark.parametrize("world,seed", [(2, 1), (4, 2), (4, 3), (8, 4)])
def test_sharded_load_offgrid_replay(emu, world, seed):
    """Sharded load with channels OFF the bin grid, unevenly spread over the ranks, overlapping
    between neighbours and reaching across the ends of the spectrum (arcs that wrap from bin N-1 to
    bin 0, halos on both sides): every virtual rank's sub-band against a float64 FFT and its audio
    against the oracle."""
    import radiocore_oracle as oracle
    from bench_support import synth
    from radiocore.tools import sharding
    rng = np.random.default_rng(seed)
    N, B, A = 192_000, 12_000, 2_400
    C_ = int(rng.integers(world, 14))
    # centres anywhere in the band (the edge channels' bins wrap around the spectrum), sorted: contiguous slices
    centers = np.sort(rng.integers(-N // 2 + B // 2 + 1, N // 2 - B // 2 - 1, size=C_)).astype(float)
    centers[0] = -N / 2 + B / 2 + 3                     # lowest channel: its lowest bins sit just above bin N/2 (cyclic)
    centers[-1] = N / 2 - B / 2 - 5
    x = synth.wideband(N, list(centers), B, seed=seed)
    o = oracle.Tuner()
    for f in centers:
        o.add_channel(1e8 + f, B, oracle.FM(B, A))
    o.input_frequency = 1e8
    o.request_bandwidth(N)
    o.load(x)
    X64 = np.fft.fft(x.astype(np.complex128))
    tuners, arcs = [], []
    for r in range(world):
        t = emu.Tuner()
        mine = sharding.shard_tuner(t, [1e8 + f for f in centers], B, lambda c: emu.FM(B, A), 1e8, N, world, r)
        arcs.append(sharding.covering_arc(t.needed_bins(), N) if mine else (0, 2))
        if mine:
            t.set_subband(*arcs[-1])
        tuners.append((t, mine))
    plan = sharding.SubbandPlan(N, world, arcs)
    fft = emu.Fft(plan.m)
    R = [np.zeros((world, plan.p), dtype=np.complex64) for _ in range(world)]
    subs = [np.zeros(arcs[d][1] + 64, dtype=np.complex64) for d in range(world)]
    for g in range(world):
        emu.fft_scatter(fft, x[g::world], [R[p][g] for p in range(world)])
    for p in range(world):
        segs = [(k1, j0, j1, subs[d], pos) for d in range(world) for k1, j0, j1, pos in plan.runs(p, d)]
        emu.subband_combine_scatter(R[p], N, p * plan.p, segs)
    rms = np.sqrt(np.mean(np.abs(X64) ** 2))
    for d, (t, mine) in enumerate(tuners):
        if not mine:
            continue
        lo, length = arcs[d]
        want = X64[(lo + np.arange(length)) % N]
        err = np.abs(subs[d][:length] - want)
        assert np.max(err) <= 2e-6 * np.max(np.abs(X64)) + 5e-6 * rms, (d, np.max(err))
        t.load_subband(subs[d])
        audio = t.run_all()
        for i, c in enumerate(mine):
            ref = o.channels()[c].demodulator.run(o.run(c))
            if np.max(np.abs(ref)) > 1e-2:          # conditioning of the relative bound (overlapping stations fade)
                parity.assert_parity(audio[i], ref, f"world {world} rank {d} ch {c}", tol_scale=2.0)
