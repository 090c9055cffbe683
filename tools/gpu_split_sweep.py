#!/usr/bin/env python
"""Time alternative pass splits of the FFT plans (RC_FFT_SPLIT) on the GPU with the library's
per-kernel event profiler:  python tools/gpu_split_sweep.py
Each line: n, batch, split, per-pass ms, total ms.  Plain complex64 in/out (no fused loaders)."""
import ctypes as C
import json
import os
import sys

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
sys.path.insert(0, os.path.join(ROOT, "radio-core_b200"))
import torch  # noqa: E402
from radiocore import _native  # noqa: E402

lib = _native.lib()
CASES = {
    (256_000_000, 1): ["640x800x500", "800x500x640", "800x640x500", "640x800x500", "800x500x640", "800x800x400"],
}
if len(sys.argv) > 1 and sys.argv[1] == "all":
    CASES.update({
        (256_000_000, 1): ["640x800x500", "640x500x800", "800x500x640", "400x800x800", "640x625x640", "512x625x800"],
        (1_000_000, 256): ["200x50x100", "200x100x50", "100x100x100", "100x50x200"],
        (500_000, 256): ["200x50x50", "100x100x50", "100x50x100"],
        (250_000, 32): ["500x500", "100x50x50", "50x100x50", "50x50x100", "625x400", "125x40x50", "250x1000"],
        (125_000, 32): ["250x500", "50x50x50", "200x625", "625x200", "125x1000"],
        (10_000_000, 1): ["200x100x500", "200x500x100", "250x200x200", "200x250x200", "160x250x250"],
        (16_000_000, 1): ["200x100x800", "200x200x400", "250x256x250", "256x250x250", "320x500x100"],
    })


if len(sys.argv) > 1 and sys.argv[1] == "small":
    # the transforms of configs 2 / 4 and of the short-block mode, at their batch sizes
    CASES = {
        (250_000, 32): ["500x500", "50x50x100", "100x50x50", "50x100x50", "400x625", "625x400", "250x1000", "40x125x50", "125x40x50", "125x50x40"],
        (250_000, 64): ["500x500", "50x50x100", "100x50x50", "400x625", "125x40x50"],
        (125_000, 32): ["250x500", "500x250", "200x625", "625x200", "50x50x50", "125x1000"],
        (125_000, 64): ["250x500", "500x250", "200x625", "50x50x50"],
        (125_000, 192): ["250x500", "500x250", "200x625", "50x50x50"],
        (24_000, 32): ["160x150", "150x160", "300x80", "80x300", "600x40", "40x600"],
        (24_000, 128): ["160x150", "150x160", "300x80", "80x300", "600x40"],
        (24_000, 256): ["160x150", "150x160", "300x80", "80x300"],
        (10_000_000, 1): ["200x100x500", "200x250x200", "250x200x200", "160x250x250", "125x400x200", "400x250x100", "320x125x250", "500x200x100", "100x100x1000"],
        (16_000_000, 1): ["200x100x800", "200x200x400", "256x250x250", "250x256x250", "320x500x100", "160x400x250", "400x400x100", "640x250x100", "500x320x100"],
        (8_000_000, 1): ["200x200x200", "160x250x200", "320x250x100", "500x160x100", "128x250x250", "250x256x125", "400x200x100", "250x160x200"],
        (31_250, 256): ["125x250", "250x125", "50x625", "625x50"],
        (15_625, 256): ["125x125"],
        (32_000_000, 1): ["160x800x250", "320x400x250", "400x320x250", "200x400x400", "250x320x400", "500x256x250", "320x250x400"],
        (128_000_000, 1): ["200x800x800", "800x640x250", "640x800x250", "500x512x500", "400x640x500", "640x400x500", "800x400x400"],
    }


def run(n, batch, split, reps=5):
    os.environ["RC_FFT_SPLIT"] = f"{n}:{split}"
    x = torch.randn(batch * n, 2, device="cuda")                 # interleaved complex64
    out = torch.empty_like(x)
    for _ in range(2):
        _native.check(lib.rc_fft_c2c(0, n, batch, -1, x.data_ptr(), out.data_ptr(), None))
    lib.rc_profile_reset()
    lib.rc_profile_enable(1)
    for _ in range(reps):
        _native.check(lib.rc_fft_c2c(0, n, batch, -1, x.data_ptr(), out.data_ptr(), None))
    need = lib.rc_profile_report(None, 0)
    buf = C.create_string_buffer(need + 16)
    lib.rc_profile_report(buf, need + 16)
    lib.rc_profile_enable(0)
    lib.rc_profile_reset()
    k = json.loads(buf.value.decode())
    per = {t: v["total_ms"] / v["count"] for t, v in k.items()}
    return per


for (n, batch), splits in CASES.items():
    for s in splits:
        try:
            per = run(n, batch, s)
            print(n, batch, s, " ".join(f"{t.split('/')[-1]}={v:.4f}" for t, v in per.items()), "total=%.4f" % sum(per.values()), flush=True)
        except Exception as exc:
            print(n, batch, s, "failed:", exc, flush=True)
