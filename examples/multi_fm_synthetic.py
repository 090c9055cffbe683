#!/usr/bin/env python
"""The per-block loop of the reference's examples/multi_fm_server.py:86-106,123-136 against this
package.  The radio is replaced by a stand-in (no SoapySDR here); the egress is the reference's:
a ZeroMQ PUB socket, one multipart message per channel and block,
``[int32-LE centre frequency, float32 audio bytes]``, decoded by subscribers exactly as
examples/multi_fm_receiver.py:23-24,47-49 does (falls back to a frame recorder without pyzmq).

    SDR thread  -> RingBuffer -> DSP thread: Tuner.load, per channel Tuner.run + demodulator.run
                                            -> socket.send_multipart([address_bytes, audio.tobytes()])

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


try:
    import zmq
except ImportError:                                           # pragma: no cover
    zmq = None


class RecordingSocket:
    """Wraps the PUB socket (or nothing, without pyzmq) and keeps (topic, payload size) of every frame."""

    def __init__(self, socket=None):
        self.frames, self.socket = [], socket

    def send_multipart(self, parts):
        self.frames.append((bytes(parts[0]), len(parts[1])))
        if self.socket is not None:
            self.socket.send_multipart(parts)


class Receiver(threading.Thread):
    """examples/multi_fm_receiver.py without the sound card: subscribe to one station by its
    address bytes, decode each payload as float32 and reshape to (samples, channels)."""

    def __init__(self, context, endpoint, frequency, channels, blocks):
        super().__init__(daemon=True)
        self.socket = context.socket(zmq.SUB)
        self.socket.connect(endpoint)
        self.socket.setsockopt(zmq.SUBSCRIBE, int(frequency).to_bytes(4, byteorder="little"))
        self.socket.setsockopt(zmq.RCVTIMEO, 60_000)
        self.channels, self.blocks, self.audio = channels, blocks, []

    def run(self):
        try:
            for _ in range(self.blocks):
                _, payload = self.socket.recv_multipart()
                audio = np.frombuffer(payload, dtype=np.float32)
                self.audio.append(audio.reshape((len(audio) // self.channels, self.channels)))
        except zmq.Again:                                     # publisher died: leave what arrived
            pass
        finally:
            self.socket.close(0)


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

    context = pub = None
    receivers = []
    if zmq is not None:
        context = zmq.Context()
        pub = context.socket(zmq.PUB)
        port = pub.bind_to_random_port("tcp://127.0.0.1")
        for channel in tuner.channels():
            receivers.append(Receiver(context, f"tcp://127.0.0.1:{port}", channel.center_frequency,
                                      channel.demodulator.channels, blocks))
            receivers[-1].start()
        time.sleep(0.3)                               # PUB/SUB: let the subscriptions reach the publisher
    socket = RecordingSocket(pub)
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
    main.received = []
    if zmq is not None:
        for r in receivers:
            r.join(30)
            main.received.append(r.audio)
        pub.close(0)
        context.term()
    return socket.frames


if __name__ == "__main__":
    frames = main(int(sys.argv[1]) if len(sys.argv) > 1 else 2)
    for addr, nbytes in frames:
        print("topic", int.from_bytes(addr, "little"), "Hz ->", nbytes, "bytes of float32 audio")
    for (freq, _, kind), blocks_rx in zip(Config.channels, main.received):
        print("subscriber", int(freq), kind, "received", [a.shape for a in blocks_rx])
