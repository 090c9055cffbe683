"""GPU parity tests proper: the CUDA kernels, called through the C ABI by the
drop-in ``radiocore`` classes, against (a) the committed outputs of the real
reference and (b) the oracle on seeded inputs at larger sizes."""
import ctypes as C
import os

import numpy as np
import pytest

import radiocore_oracle as oracle
from bench_support import synth
from tests import parity
from tests.golden import cases

pytestmark = pytest.mark.gpu

GOLDEN = np.load(os.path.join(os.path.dirname(__file__), "golden", "golden.npz"))


@pytest.fixture(scope="module")
def rc():
    import torch
    assert torch.cuda.is_available(), "GPU tests need a CUDA device"
    import radiocore
    assert radiocore.HasCuda()
    return radiocore


@pytest.mark.parametrize("name", list(cases.CASES))
def test_cuda_matches_reference_golden(rc, name):
    res = cases.CASES[name](rc, np)
    for key, val in res.items():
        ref = GOLDEN[f"{name}/{key}"]
        if key == "f_in":
            assert np.array_equal(ref, val)
            continue
        parity.assert_parity(val, ref, f"{name}/{key}", tol_scale=parity.LOOSE.get(f"{name}/{key}", 1.0))


@pytest.mark.parametrize("n,batch", [(1, 2), (2, 3), (30, 4), (625, 3), (1000, 2), (4800, 2), (24000, 2),
                                     (25000, 2), (250000, 1), (1000000, 1), (2500000, 1), (10000000, 1)])
def test_fft_engine_vs_numpy(rc, n, batch):
    import torch
    from radiocore import _native
    rng = np.random.default_rng(n)
    x = (rng.standard_normal((batch, n)) + 1j * rng.standard_normal((batch, n))).astype(np.complex64)
    xd = torch.from_numpy(x).cuda()
    for sign in (-1, 1):
        out = torch.empty_like(xd)
        _native.check(_native.lib().rc_fft_c2c(0, n, batch, sign, xd.data_ptr(), out.data_ptr(), None))
        torch.cuda.synchronize()
        ref = np.fft.fft(x.astype(np.complex128), axis=1) if sign < 0 else np.fft.ifft(x.astype(np.complex128), axis=1) * n
        err = np.max(np.abs(out.cpu().numpy() - ref)) / np.sqrt(np.mean(np.abs(ref) ** 2))
        assert err < 3e-6, (n, sign, err)


def _pair(rc, N, B, A, C_, kind, offs=None, f0=100e6):
    offs = synth.tiling_centers(N, C_, B) if offs is None else offs
    g, o = rc.Tuner(cuda=True), oracle.Tuner()
    for off in offs:
        g.add_channel(f0 + off, B, getattr(rc, kind)(B, A, cuda=True))
        o.add_channel(f0 + off, B, getattr(oracle, kind)(B, A))
    g.request_bandwidth(N)
    o.request_bandwidth(N)
    return g, o, offs


def test_config2_tuner_32_mfm(rc):
    """BASELINE config 2: 10 MHz -> 32 x 250 kHz -> MFM 48 kHz, two blocks, every channel."""
    N, B, A, C_ = 10_000_000, 250_000, 48_000, 32
    offs = [-4e6 + 125e3 + c * 250e3 for c in range(C_)]
    g, o, _ = _pair(rc, N, B, A, C_, "MFM", offs)
    worst = 0.0
    for blk in range(2):
        x = synth.wideband(N, offs, B, seed=42, block=blk)
        g.load(x)
        o.load(x)
        for ch in g.channels():
            got = ch.demodulator.run(g.run(ch.index))
            ref = o.channels()[ch.index].demodulator.run(o.run(ch.index))
            assert got.shape == (A, 1) and got.dtype == np.float32
            worst = max(worst, parity.assert_parity(got, ref, f"cfg2 b{blk} ch{ch.index}"))
    print("config2 worst rel err", worst)


def test_config4_wbfm_stereo(rc):
    """BASELINE config 4 as specified: N = 16e6, all 64 x 250 kHz stereo channels,
    WBFM(250e3 -> 48e3, 75 us), two blocks (carried de-emphasis state), every channel vs the oracle."""
    import bench
    N, B, A, C_ = 16_000_000, 250_000, 48_000, 64
    g, o, offs = _pair(rc, N, B, A, C_, "WBFM")
    worst = 0.0
    for blk in range(2):
        x = bench.make_wideband_gpu(N, C_, B, 4 + 100 * blk, True, "cuda")[0]
        g.load(x)
        audio = g.run_all(numpy_output=True).copy()
        o.load(x.cpu().numpy())
        del x
        for c, (off_c, size, nch) in enumerate(g.audio_slices()):
            got = audio[off_c: off_c + size * nch].reshape(1, size, nch)
            ref = o.channels()[c].demodulator.run(o.run(c))
            assert ref.shape == (1, A, 2)
            worst = max(worst, parity.assert_parity(got, ref, f"cfg4 b{blk} ch{c}"))
    print("config4 (64 ch, N=16e6) worst rel err", worst)


def test_config1_decimate_wbfm(rc):
    """BASELINE config 1 plumbing (examples/receive_fm.py:76-82,100-101): Decimate 2.5M->250k, WBFM ->48k."""
    n_in, B, A = 2_500_000, 250_000, 48_000
    x = synth.station(n_in, n_in, 0, offset_hz=12_345.0, deviation=75e3, stereo=True)
    rng = np.random.default_rng(1234)
    x = (x + 0.02 * (rng.standard_normal(n_in) + 1j * rng.standard_normal(n_in))).astype(np.complex64)
    gd, gw = rc.Decimate(n_in, B, cuda=True), rc.WBFM(B, A, cuda=True)
    od, ow = oracle.Decimate(n_in, B), oracle.WBFM(B, A)
    iq_g, iq_o = gd.run(x), od.run(x)
    parity.assert_parity(iq_g.cpu().numpy(), iq_o, "decimate 2.5M->250k")
    parity.assert_parity(gw.run(iq_g), ow.run(iq_o), "wbfm after decimate")


def test_config3_shape_fm_1m_channels(rc):
    """BASELINE config 3 geometry scaled to 16 channels: N=16e6, B=1e6, FM(1e6 -> 48e3), full oracle."""
    N, B, A, C_ = 16_000_000, 1_000_000, 48_000, 16
    g, o, offs = _pair(rc, N, B, A, C_, "FM")
    x = synth.wideband(N, offs, B, seed=3, deviation=75e3)
    g.load(x)
    o.load(x)
    for ch in g.channels():
        got = ch.demodulator.run(g.run(ch.index))
        ref = o.channels()[ch.index].demodulator.run(o.run(ch.index))
        parity.assert_parity(got, ref, f"cfg3 ch{ch.index}")


def test_channel_iq_and_view(rc):
    N, B, A, C_ = 400_000, 50_000, 12_000, 8
    g, o, offs = _pair(rc, N, B, A, C_, "FM")
    x = synth.wideband(N, offs, B, seed=9)
    g.load(x)
    o.load(x)
    v = g.run(3)
    assert len(v) == B and v.shape == (B,)
    parity.assert_parity(np.asarray(v), o.run(3), "channel IQ")
    dev = g.channels()[3].demodulator.run(v, numpy_output=False)
    assert dev.is_cuda and tuple(dev.shape) == (A, 1)
    # a foreign demodulator handed the view falls back to the standalone kernels: same numbers
    other = rc.FM(B, A, cuda=True).run(v)
    assert np.max(np.abs(other - dev.cpu().numpy())) <= 2e-6


def test_errors(rc):
    with pytest.raises(ValueError):
        rc.FM(1000, 100).run(np.zeros(999, dtype=np.complex64))
    with pytest.raises(ValueError):
        rc.Decimate(1000, 100).run(np.zeros(1001))
    with pytest.raises(ValueError):
        rc.Deemphasis(1000).run(np.zeros(1001))
    with pytest.raises(ValueError):
        rc.FM(2 * 7 * 11, 14).run(np.zeros(154, dtype=np.complex64))
    t = rc.Tuner()
    t.add_channel(1e6, 1000, None)
    with pytest.raises(ValueError):
        t.request_bandwidth(10)


def test_block_pipeline_matches_synchronous_path(rc):
    """Tuner.submit()/collect() (copy / kernels / read-back overlapped, 2 blocks in flight) gives
    the same audio, block for block, as load() + run_all(), including the carried de-emphasis state."""
    import torch
    N, B, A, C_ = 400_000, 50_000, 12_000, 8
    offs = synth.tiling_centers(N, C_, B)
    blocks = [synth.wideband(N, offs, B, seed=21, block=b) for b in range(4)]
    ref_t, pipe_t = rc.Tuner(cuda=True), rc.Tuner(cuda=True)
    for off in offs:
        ref_t.add_channel(100e6 + off, B, rc.MFM(B, A, cuda=True))
        pipe_t.add_channel(100e6 + off, B, rc.MFM(B, A, cuda=True))
    ref_t.request_bandwidth(N)
    pipe_t.request_bandwidth(N)
    want = []
    for x in blocks:
        ref_t.load(x)
        want.append(ref_t.run_all(numpy_output=True).copy())
    pinned = [torch.from_numpy(x).pin_memory() for x in blocks]
    got, prev = [], None
    for x in pinned:
        t = pipe_t.submit(x)
        if prev is not None:
            got.append(pipe_t.collect(prev).copy())
        prev = t
    got.append(pipe_t.collect(prev).copy())
    for a, b in zip(got, want):
        assert np.array_equal(a, b)
    with pytest.raises(RuntimeError):
        pipe_t.collect(0)                       # expired ticket


def test_fft_one_billion_points(rc):
    """BASELINE config 5 size (N = 1e9 = 2^9 5^9, four passes): index arithmetic at the top of the
    32-bit range.  Two complex exponentials -> two spectral lines of known height, everything else
    at the fp32 noise floor."""
    import torch
    from radiocore import _native
    if torch.cuda.get_device_properties(0).total_memory < 60e9:
        pytest.skip("needs ~40 GB of device memory")
    n, k0, k1 = 1_000_000_000, 123_456_789, 999_999_937
    t = torch.arange(n, device="cuda", dtype=torch.int64)
    x = torch.empty(n, dtype=torch.complex64, device="cuda")
    step = 50_000_000
    for s in range(0, n, step):                      # build in slices: exact integer phase index, fp64 angle
        tt = t[s:s + step]
        ph0 = torch.remainder(tt * k0, n).to(torch.float64) * (2 * np.pi / n)
        ph1 = torch.remainder(tt * k1, n).to(torch.float64) * (2 * np.pi / n)
        x[s:s + step] = (torch.polar(torch.ones_like(ph0), ph0) + 0.5 * torch.polar(torch.ones_like(ph1), ph1)).to(torch.complex64)
        del tt, ph0, ph1
    del t
    out = torch.empty_like(x)
    _native.check(_native.lib().rc_fft_c2c(0, n, 1, -1, x.data_ptr(), out.data_ptr(), None))
    torch.cuda.synchronize()
    a0, a1 = out[k0].item(), out[k1].item()
    assert abs(a0 - n) / n < 2e-5, a0
    assert abs(a1 - 0.5 * n) / n < 2e-5, a1
    out[k0] = 0
    out[k1] = 0
    mag = out.abs()
    assert float(mag.max()) / n < 2e-5                # nothing leaked into other bins
    assert float(mag.pow(2).sum().sqrt()) / n < 2e-5 * 3


def test_example_server_loop(rc):
    """examples/multi_fm_synthetic.py: the reference's multi_fm_server loop (RingBuffer -> Buffer ->
    Tuner.load -> per-channel run -> tobytes) with stand-ins for the radio and the socket."""
    import importlib.util
    path = os.path.join(os.path.dirname(os.path.dirname(__file__)), "examples", "multi_fm_synthetic.py")
    spec = importlib.util.spec_from_file_location("multi_fm_synthetic", path)
    mod = importlib.util.module_from_spec(spec)
    spec.loader.exec_module(mod)
    frames = mod.main(blocks=2)
    assert len(frames) == 6
    topics = [int.from_bytes(a, "little") for a, _ in frames[:3]]
    assert topics == [int(f) for f, _, _ in mod.Config.channels]
    assert [n for _, n in frames[:3]] == [48000 * 2 * 4, 48000 * 4, 48000 * 4]
    if mod.zmq is not None:
        # every subscriber got both blocks of its own station, decoded as the reference's receiver does
        shapes = [[a.shape for a in rx] for rx in mod.main.received]
        assert shapes == [[(48000, 2)] * 2, [(48000, 1)] * 2, [(48000, 1)] * 2]
        for rx in mod.main.received:
            assert all(np.all(np.isfinite(a)) and np.max(np.abs(a)) <= 0.999 + 1e-6 for a in rx)


def test_config3_literal_block(rc):
    """BASELINE config 3 as benchmarked: N = 256e6, 256 x 1 MHz channels, with FM(1e6 -> 48e3)
    (`cfg3`) AND with WBFM(1e6 -> 48e3) stereo demodulators (`cfg3-wbfm`, the north star's headline
    chain) -- both engines on the full block, two blocks (carried de-emphasis state), against the
    oracle on a subset of channels (ch 0, 1, 127, 255)."""
    import torch
    import bench
    if torch.cuda.get_device_properties(0).total_memory < 100e9:
        pytest.skip("needs ~60 GB of device memory")
    N, C_, B, A = 256_000_000, 256, 1_000_000, 48_000
    subset = (0, 1, 127, 255)
    offs = bench.tiling_offsets(N, C_, B)
    tuners, oracles = {}, {}
    for kind in ("FM", "WBFM"):
        g = rc.Tuner(cuda=True)
        o = oracle.Tuner(fft_workers=os.cpu_count())
        for c, off in enumerate(offs):
            g.add_channel(100e6 + off, B, getattr(rc, kind)(B, A, cuda=True))
            o.add_channel(100e6 + off, B, getattr(oracle, kind)(B, A) if c in subset else None)
        g.request_bandwidth(N)
        o.request_bandwidth(N)
        tuners[kind], oracles[kind] = g, o
    worst = {"FM": 0.0, "WBFM": 0.0}
    for blk in range(2):
        x_dev, _ = bench.make_wideband_gpu(N, C_, B, 3 + 100 * blk, True, "cuda")
        got = {}
        for kind, g in tuners.items():
            g.load(x_dev)
            got[kind] = g.run_all(numpy_output=True).copy()
        x = x_dev.cpu().numpy()
        del x_dev
        torch.cuda.empty_cache()
        oracles["FM"].load(x)
        oracles["WBFM"]._buffer = oracles["FM"]._buffer          # same block: one 256 M-point CPU FFT
        del x
        for kind, g in tuners.items():
            slices = g.audio_slices()
            for c in subset:
                off_c, size, nch = slices[c]
                a = got[kind][off_c: off_c + size * nch].reshape(size, nch)
                ref = oracles[kind].channels()[c].demodulator.run(oracles[kind].run(c))
                ref = ref.reshape(size, nch)
                worst[kind] = max(worst[kind], parity.assert_parity(a, ref, f"cfg3 literal {kind} b{blk} ch{c}"))
    print("config3 literal worst rel err", worst)


def test_config3_short_block(rc):
    """Short-block variant of config 3: T = 1/32 s blocks, N = 8e6, 256 channels of
    B*T = 31250 bins, FM to A*T = 1500 samples.  The reference is size-driven (tuner.py:155-157,
    decimate.py:32-33), so the oracle instantiated in BIN UNITS -- Tuner.add_channel(f*T, bw*T),
    FM(bw*T, 48e3*T) -- is the reference for this mode; every channel, two blocks, eager and
    through the captured CUDA graph (Tuner.step)."""
    import bench
    N, C_, B, A = bench.WORKLOADS["cfg3-short"][:4]
    g, o, offs = _pair(rc, N, B, A, C_, "FM")
    g2, _, _ = _pair(rc, N, B, A, C_, "FM")
    worst = 0.0
    for blk in range(3):
        x = bench.make_wideband_gpu(N, C_, B, 3 + 100 * blk, False, "cuda")[0]
        g.load(x)
        audio = g.run_all(numpy_output=True).copy()
        stepped = g2.step(x, numpy_output=True).copy()          # block 0 eager, block 1 captured + replayed, block 2 replayed
        assert np.array_equal(audio, stepped), f"graph replay differs, block {blk}"
        if blk == 2:
            continue
        o.load(x.cpu().numpy())
        for c, (off_c, size, nch) in enumerate(g.audio_slices()):
            ref = o.channels()[c].demodulator.run(o.run(c))
            worst = max(worst, parity.assert_parity(audio[off_c: off_c + size].reshape(size, 1), ref, f"short b{blk} ch{c}"))
    print("config3 short-block worst rel err", worst)


def test_graph_step_carries_state(rc):
    """Tuner.step (one CUDA graph per block) == load + run_all for demodulators with carried
    de-emphasis state (MFM, WBFM), four blocks."""
    N, B, A, C_ = 2_000_000, 250_000, 48_000, 8
    offs = synth.tiling_centers(N, C_, B)
    for kind in ("MFM", "WBFM"):
        a, _, _ = _pair(rc, N, B, A, C_, kind)
        b, _, _ = _pair(rc, N, B, A, C_, kind)
        for blk in range(4):
            x = synth.wideband(N, offs, B, seed=17, stereo=kind == "WBFM", block=blk)
            a.load(x)
            want = a.run_all(numpy_output=True).copy()
            got = b.step(x, numpy_output=True).copy()
            assert np.array_equal(want, got), (kind, blk)


def test_config5_one_gpu_slice(rc):
    """BASELINE configs[4]: N = 1e9, 2048 x 250 kHz FM channels, 256 per GPU.  One GPU's slice
    (rank 3 of 8: channels 768..1023, band plan of the full list) on the full 1 G-sample block;
    channels at both ends and in the middle of the slice against the oracle's O(B) gather
    (oracle/radiocore_oracle.py `_tuner_gather`) from a CPU FFT of the same block."""
    import torch
    import bench
    from radiocore.tools import sharding
    if torch.cuda.get_device_properties(0).total_memory < 100e9:
        pytest.skip("needs ~60 GB of device memory")
    N, C_, B, A = 1_000_000_000, 2048, 250_000, 48_000
    offs = bench.tiling_offsets(N, C_, B)
    x_dev, _ = bench.make_wideband_gpu(N, C_, B, 5, False, "cuda")
    g = rc.Tuner(cuda=True)
    mine = sharding.shard_tuner(g, [100e6 + f for f in offs], B, lambda c: rc.FM(B, A, cuda=True), 100e6, N, 8, 3)
    assert mine[0] == 768 and len(mine) == 256
    g.load(x_dev)
    audio = g.run_all(numpy_output=True).copy()
    slices = g.audio_slices()
    x = x_dev.cpu().numpy()
    del x_dev
    del g
    torch.cuda.empty_cache()
    o = oracle.Tuner()
    check = (768, 769, 900, 1023)
    for c, off in enumerate(offs):
        o.add_channel(100e6 + off, B, oracle.FM(B, A) if c in check else None)
    o.request_bandwidth(N)
    o.load(x)                                   # one 1e9-point complex64 CPU FFT (minutes)
    del x
    worst = 0.0
    for c in check:
        off_c, size, nch = slices[c - mine[0]]
        got = audio[off_c: off_c + size * nch].reshape(size, nch)
        ref = o.channels()[c].demodulator.run(o.run(c))
        worst = max(worst, parity.assert_parity(got, ref, f"cfg5 ch{c}"))
    print("config5 slice worst rel err", worst)


def test_unaligned_device_inputs(rc):
    """CUDA tensors whose data pointer is only 8-byte aligned (a view starting at an odd complex
    sample) cannot be described to the TMA unit or read with 16-byte loads: the kernels must fall
    back to per-thread loads and give the same numbers."""
    import torch
    N, B, A, C_ = 400_000, 50_000, 12_000, 8
    offs = synth.tiling_centers(N, C_, B)
    x = torch.from_numpy(synth.wideband(N, offs, B, seed=33)).cuda()
    big = torch.empty(N + 1, dtype=torch.complex64, device="cuda")
    big[1:] = x
    view = big[1:]
    assert view.data_ptr() % 16 == 8
    outs = []
    for inp in (x, view):
        t = rc.Tuner(cuda=True)
        for off in offs:
            t.add_channel(100e6 + off, B, rc.MFM(B, A, cuda=True))
        t.request_bandwidth(N)
        t.load(inp)
        outs.append(t.run_all(numpy_output=True).copy())
    assert np.array_equal(outs[0], outs[1])
    # stand-alone demodulator on an unaligned channel block
    iq = torch.from_numpy(synth.station(B, B, 1, offset_hz=321.0, deviation=0.3 * B).astype(np.complex64)).cuda()
    big2 = torch.empty(B + 1, dtype=torch.complex64, device="cuda")
    big2[1:] = iq
    a = rc.FM(B, A, cuda=True).run(iq)
    b = rc.FM(B, A, cuda=True).run(big2[1:])
    assert np.array_equal(a, b)


def test_reference_benchmark_call_pattern(rc):
    """tests/benchmark.py of the reference, call for call (one iteration each): demodulators fed a
    zero-filled ``Buffer.data``, ``Decimate`` on 10 M and 2.5 M complex samples, a ``Tuner`` whose
    channels carry the demodulator *class* (never invoked) -- sizes and argument types as there."""
    buff = rc.Buffer(int(256e3), dtype=np.complex64, cuda=True)
    assert buff.data.dtype == np.complex64 and len(buff) == 256000
    for cls, shape in ((rc.WBFM, (1, 32000, 2)), (rc.MFM, (32000, 1)), (rc.FM, (32000, 1))):
        out = cls(256e3, 32e3, cuda=True).run(buff.data)
        assert out.shape == shape and out.dtype == np.float32
        if cls is not rc.WBFM:                      # WBFM on silence is 0/0 in the reference as well
            assert np.all(np.isfinite(out))
    for n_in in (10e6, 2.5e6):
        big = rc.Buffer(n_in, dtype=np.complex64, cuda=True)
        y = rc.Decimate(n_in, 250e3, cuda=True).run(big.data)
        assert len(y) == 250000 and y.is_cuda and not bool(y.any())      # backend array, as the reference's cuda path
    tuner = rc.Tuner(cuda=True)
    for f in (94.5e6, 97.5e6, 96.9e6):
        tuner.add_channel(f, int(250e3), rc.FM)
    tuner.request_bandwidth(int(10e6))
    block = rc.Buffer(10e6, dtype=np.complex64, cuda=True)
    block.data[:] = synth.wideband(10_000_000, [f - tuner.input_frequency for f in (94.5e6, 97.5e6, 96.9e6)],
                                   250_000, seed=8)
    o = oracle.Tuner()
    for f in (94.5e6, 97.5e6, 96.9e6):
        o.add_channel(f, int(250e3), None)
    o.request_bandwidth(int(10e6))
    o.load(block.data)
    tuner.load(block.data)
    parity.assert_parity(np.asarray(tuner.run(0)), o.run(0), "benchmark tuner channel 0")
