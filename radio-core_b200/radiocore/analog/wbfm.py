"""Broadcast stereo FM demodulator (mirror of radiocore/analog/wbfm.py:11-105)."""
from radiocore.analog._demod import DemodBase, MODE_WBFM


class WBFM(DemodBase):
    """FM(same size) -> 19 kHz pilot filtfilt -> Hilbert 'PLL' -> L-R recovery ->
    two decimations -> two stateful de-emphases -> joint mean removal -> clip."""
    _mode = MODE_WBFM
    _channels = 2
