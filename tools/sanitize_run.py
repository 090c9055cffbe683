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


if __name__ == "__main__":
    main()
