"""Broadcast mono FM demodulator (mirror of radiocore/analog/mfm.py:8-71)."""
from radiocore.analog._demod import DemodBase, MODE_MFM


class MFM(DemodBase):
    """FM -> stateful de-emphasis -> block mean removal -> clip(+-0.999)."""
    _mode = MODE_MFM
    _channels = 1
