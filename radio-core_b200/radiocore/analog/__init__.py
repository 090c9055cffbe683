"""FM demodulators and their building blocks (mirror of radiocore/analog/__init__.py:3-9)."""
from radiocore.analog.pll import PLL
from radiocore.analog.wbfm import WBFM
from radiocore.analog.mfm import MFM
from radiocore.analog.fm import FM
from radiocore.analog.deemphasis import Deemphasis
from radiocore.analog.decimate import Decimate
from radiocore.analog.bandpass import Bandpass

__all__ = ["PLL", "WBFM", "MFM", "FM", "Deemphasis", "Decimate", "Bandpass"]
