#!/usr/bin/env python
"""Small pass over every kernel family for compute-sanitizer (memcheck / racecheck / initcheck):

    compute-sanitizer --tool memcheck python tools/sanitize_run.py

Tuner + MFM and WBFM channels (TMA and per-thread loaders, aligned and 8-byte-aligned input),
the block pipeline, and the stand-alone Decimate / Bandpass / PLL / Deemphasis operators.
Sizes are a few hundred thousand samples so the instrumented run takes seconds."""
import os
import sys

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
for p in (ROOT, os.path.join(ROOT, "radio-core_b200")):
    if p not in sys.path:
        sys.path.insert(0, p)

import numpy as np  # noqa: E402
import torch  # noqa: E402

import radiocore as rc  # noqa: E402
from bench_support import synth  # noqa: E402


def main():
    torch.cuda.set_device(0)
    N, B, A, C_ = 400_000, 50_000, 12_000, 8
    offs = synth.tiling_centers(N, C_, B)
    x = synth.wideband(N, offs, B, seed=2, stereo=True)
    for kind in ("MFM", "WBFM", "FM"):
        t = rc.Tuner(cuda=True)
        for off in offs:
            t.add_channel(100e6 + off, B, getattr(rc, kind)(B, A, cuda=True))
        t.request_bandwidth(N)
        t.load(x)
        a = t.run_all(numpy_output=True).copy()
        big = torch.empty(N + 1, dtype=torch.complex64, device="cuda")
        big[1:] = torch.from_numpy(x).cuda()
        t.load(big[1:])                                    # 8-byte aligned: per-thread loaders
        t.run_all(numpy_output=True)
        tk = t.submit(torch.from_numpy(x).pin_memory())    # block pipeline
        t.collect(tk)
        assert np.all(np.isfinite(a))
        print(kind, "ok", float(np.abs(a).max()))
    # literal config-2 geometry, one block (R = 500 passes, 250 kHz channels)
    N2, B2, A2 = 2_000_000, 250_000, 48_000
    offs2 = synth.tiling_centers(N2, 8, B2)
    t = rc.Tuner(cuda=True)
    for off in offs2:
        t.add_channel(100e6 + off, B2, rc.WBFM(B2, A2, cuda=True))
    t.request_bandwidth(N2)
    t.load(synth.wideband(N2, offs2, B2, seed=4, stereo=True))
    t.run_all(numpy_output=True)
    print("cfg4-geometry ok")
    iq = synth.station(250_000, 250_000, 1, offset_hz=1234.0, deviation=75e3).astype(np.complex64)
    rc.Decimate(250_000, 48_000, cuda=True).run(iq)
    mpx = np.real(iq).astype(np.float64)
    rc.Decimate(250_000, 48_000, cuda=True).run(mpx)
    rc.Bandpass(250_000, 19e3 - 50, 19e3 + 50, num_taps=41, cuda=True).run(mpx)
    pll = rc.PLL(cuda=True)
    pll.step(mpx)
    pll.image(2.0)
    rc.Deemphasis(48_000, cuda=True).run(mpx[:48_000])
    torch.cuda.synchronize()
    print("stand-alone operators ok")
    sharded_load_single_gpu()
    # one block as a CUDA graph (Tuner.step): eager warm-up, capture, replay
    t = rc.Tuner(cuda=True)
    for off in offs:
        t.add_channel(100e6 + off, B, rc.MFM(B, A, cuda=True))
    t.request_bandwidth(N)
    for _ in range(3):
        t.step(x, numpy_output=True)
    print("graph step ok")


def sharded_load_single_gpu(world=4):
    """The kernels of the sharded Tuner.load with every 'rank' on this GPU: local transforms whose
    last pass scatters pieces (rc_fft_exec_scatter), the scattering combine
    (rc_subband_combine_scatter) into each rank's sub-band, and engines in sub-band mode."""
    import ctypes as C
    from radiocore import _native
    from radiocore.tools import sharding
    lib = _native.lib()
    N, B, A, C_ = 1_600_000, 100_000, 20_000, 16
    offs = synth.tiling_centers(N, C_, B)
    x = torch.from_numpy(synth.wideband(N, offs, B, seed=3)).cuda()
    tuners, arcs = [], []
    for r in range(world):
        t = rc.Tuner(cuda=True)
        sharding.shard_tuner(t, [100e6 + f for f in offs], B, lambda c: rc.MFM(B, A, cuda=True), 100e6, N, world, r)
        arcs.append(sharding.covering_arc(t.needed_bins(), N))
        t.set_subband(*arcs[-1])
        tuners.append(t)
    plan = sharding.SubbandPlan(N, world, arcs)
    fft = C.c_void_p()
    _native.check(lib.rc_fft_create(0, plan.m, 1, C.byref(fft)))
    R = [torch.zeros(world * plan.p, dtype=torch.complex64, device="cuda") for _ in range(world)]
    subs = [torch.zeros(arcs[d][1] + (1 << 16), dtype=torch.complex64, device="cuda") for d in range(world)]
    for g in range(world):
        bases = (C.c_void_p * world)(*[R[p].data_ptr() + 8 * g * plan.p for p in range(world)])
        branch = x[g::world].contiguous()
        _native.check(lib.rc_fft_exec_scatter(fft, -1, branch.data_ptr(), bases, world, plan.p, None))
    for p in range(world):
        segs = [(k1, j0, j1, subs[d].data_ptr() + 8 * pos) for d in range(world) for k1, j0, j1, pos in plan.runs(p, d)]
        arr = (_native.ScatterSeg * len(segs))()
        for a, (k1, j0, j1, dst) in zip(arr, segs):
            a.k1, a.reserved, a.j_lo, a.j_hi, a.dst = k1, 0, j0, j1, dst
        _native.check(lib.rc_subband_combine_scatter(0, world, plan.p, N, p * plan.p, R[p].data_ptr(), arr, len(segs), None))
    full = rc.Tuner(cuda=True)
    for off in offs:
        full.add_channel(100e6 + off, B, rc.MFM(B, A, cuda=True))
    full.request_bandwidth(N)
    full.load(x)
    want = full.run_all(numpy_output=True).copy()
    got = []
    for d in range(world):
        tuners[d].load_subband(subs[d])
        got.append(tuners[d].run_all(numpy_output=True).copy())
    err = float(np.max(np.abs(np.concatenate(got) - want)))
    assert err < 5e-6, err
    lib.rc_fft_destroy(fft)
    print("sharded load (single GPU, %d virtual ranks) ok, max |diff| vs one FFT %.2e" % (world, err))


if __name__ == "__main__":
    main()
