"""radiocore -- B200-native drop-in for the FM receive path of luigifcruz/radio-core.

Same import name and class surface as the reference package
(``radiocore/__init__.py:3-28``): ``from radiocore import Tuner, WBFM, MFM, FM,
Decimate, Deemphasis, Bandpass, PLL, Buffer, RingBuffer`` keeps working, but
every ``run``/``load`` executes hand-written sm_100a kernels through the C ABI
in ``include/radiocore_b200.h``.  The ``cuda=`` keyword of the reference is
accepted and ignored: there is exactly one backend, and no CPU fallback.
"""
from radiocore.analog import *   # noqa: F401,F403
from radiocore.tools import *    # noqa: F401,F403


def HasCuda():
    """True when the native library loads and a CUDA device is visible
    (reference radiocore/__init__.py:6-26 probes cupy + cusignal instead)."""
    try:
        import torch
        from radiocore import _native
        _native.load_library()
        return bool(torch.cuda.is_available())
    except Exception:
        return False


__version__ = "1.0.0"
