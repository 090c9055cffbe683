#!/usr/bin/env python
"""GPU check of the FFT engine (rc_fft_c2c) against numpy for a list of sizes."""
import os, sys
ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
sys.path.insert(0, os.path.join(ROOT, "radio-core_b200"))
import numpy as np, torch
from radiocore import _native
lib = _native.lib()
sizes = [int(a) for a in sys.argv[1:]] or [10000, 16000, 20000, 80000, 250000, 400000, 1000000]
for n in sizes:
    rng = np.random.default_rng(n)
    batch = 3 if n <= 100000 else 1
    x = (rng.standard_normal((batch, n)) + 1j * rng.standard_normal((batch, n))).astype(np.complex64)
    xd = torch.from_numpy(x).cuda()
    for sign in (-1, 1):
        out = torch.empty_like(xd)
        _native.check(lib.rc_fft_c2c(0, n, batch, sign, xd.data_ptr(), out.data_ptr(), None))
        torch.cuda.synchronize()
        ref = np.fft.fft(x.astype(np.complex128), axis=1) if sign < 0 else np.fft.ifft(x.astype(np.complex128), axis=1) * n
        d = np.abs(out.cpu().numpy() - ref)
        err = d.max() / np.sqrt(np.mean(np.abs(ref) ** 2))
        bad = np.argwhere(d > 1e-3 * np.sqrt(n))
        print(n, sign, "err %.3g" % err, "bad", len(bad), bad[:8].tolist() if len(bad) else "", flush=True)
