"""Generic FM demodulator (mirror of radiocore/analog/fm.py:8-72)."""
from radiocore.analog._demod import DemodBase, MODE_FM


class FM(DemodBase):
    """Phase-difference discriminator followed by Fourier decimation to ``output_size``.

    ``deemphasis`` is unused in this mode, as in the reference.
    """
    _mode = MODE_FM
    _channels = 1
