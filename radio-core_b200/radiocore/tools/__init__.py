"""Channeliser and host-side plumbing (mirror of radiocore/tools/__init__.py:3-7)."""
from radiocore.tools.tuner import Tuner, Channel
from radiocore.tools.buffer import Buffer
from radiocore.tools.ringbuffer import RingBuffer
from radiocore.tools.carrousel import Carrousel
from radiocore.tools.chopper import Chopper
from radiocore.tools import sharding

__all__ = ["Tuner", "Channel", "Buffer", "RingBuffer", "Carrousel", "Chopper", "sharding"]
