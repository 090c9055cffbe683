#!/usr/bin/env python
"""The per-block loop of the reference's examples/multi_fm_server.py:86-106,123-136 against this
package, with the radio and the ZeroMQ socket replaced by stand-ins (no SoapySDR / pyzmq here):

    SDR thread  -> RingBuffer -> DSP thread: Tuner.load, per channel Tuner.run + demodulator.run
                                            -> "socket.send_multipart([address_bytes, audio.tobytes()])"

Run:  python examples/multi_fm_synthetic.py [blocks]
"""
import os
import sys
import threading
import time

import numpy as np

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
for p in (ROOT, os.path.join(ROOT, "radio-core_b200")):
    if p not in sys.path:
        sys.path.insert(0, p)

from radiocore import Buffer, MFM, RingBuffer, Tuner, WBFM   # noqa: E402  (same import as the reference)
from bench_support import synth                              # noqa: E402


class Config:
    input_rate = 2_000_000            # one-second blocks: samples == Hz, as in the reference
    channels = [(100.0e6 - 500e3, 250e3, "wbfm"), (100.0e6, 250e3, "mfm"), (100.0e6 + 500e3, 250e3, "mfm")]
    audio_rate = 48_000
    deemphasis = 75e-6


class FakeSocket:
    """Stands in for the ZeroMQ PUB socket: collects the multipart frames."""

    def __init__(self):
        self.frames = []

    def send_multipart(self, parts):
        self.frames.append((bytes(parts[0]), len(parts[1])))


def sdr_thread(ring, blocks, offsets):
    """Stands in for SoapySDR.readStream: pushes phase-continuous synthetic IQ in chunks."""
    chunk = Config.input_rate // 8
    for blk in range(blocks):
        x = synth.wideband(Config.input_rate, offsets, 250_000, seed=11, block=blk)
        for s in range(0, len(x), chunk):
            while ring.vacancy < chunk:               # a real radio would overflow; the stand-in waits
                time.sleep(0.001)
            ring.put(x[s:s + chunk])


def main(blocks=2):
    cfg = Config
    tuner = Tuner(cuda=True)
    for freq, bw, kind in cfg.channels:
        demod = (WBFM if kind == "wbfm" else MFM)(bw, cfg.audio_rate, deemphasis=cfg.deemphasis, cuda=True)
        tuner.add_channel(freq, bw, demod)
    tuner.request_bandwidth(cfg.input_rate)
    offsets = [f - tuner.input_frequency for f, _, _ in cfg.channels]

    ring = RingBuffer(cfg.input_rate * 3, cuda=True)
    producer = threading.Thread(target=sdr_thread, args=(ring, blocks, offsets), daemon=True)
    producer.start()

    socket = FakeSocket()
    tmp_buffer = Buffer(cfg.input_rate, cuda=True)
    done = 0
    while done < blocks:
        if not ring.get(tmp_buffer.data):            # blocks up to 3 s, like the reference
            continue
        tuner.load(tmp_buffer.data)
        for channel in tuner.channels():
            tmp = tuner.run(channel.index)
            tmp = channel.demodulator.run(tmp)
            socket.send_multipart([channel.address_bytes, tmp.tobytes()])
        done += 1
    producer.join()
    return socket.frames


if __name__ == "__main__":
    frames = main(int(sys.argv[1]) if len(sys.argv) > 1 else 2)
    for addr, nbytes in frames:
        print("topic", int.from_bytes(addr, "little"), "Hz ->", nbytes, "bytes of float32 audio")
