"""The reference's examples/multi_fm_server.py, UNMODIFIED, on top of this package.

Runs in a CHILD process (`python tests/test_reference_example.py <out.npy>`): the script never closes
its ZeroMQ socket, so the interpreter that ran it is left with a context that blocks at exit; the
child saves what the subscriber received and leaves with os._exit.

The script is the vendored byte-for-byte copy (oracle/_ref/examples, made by oracle/make_ref.py;
/root/reference/examples in the build container).  It is executed as ``__main__`` with two
stand-ins registered before it starts: a ``SoapySDR`` module whose device plays a synthetic
wideband block in a loop (the image has no radio and no SoapySDR), and -- through ``sys.path`` --
this repository's ``radiocore`` package in place of the reference's.  A ZeroMQ subscriber decodes
what the server publishes exactly like examples/multi_fm_receiver.py:23-24,46-50; once it has two
blocks of every station the main thread is interrupted, which the script handles itself
(``KeyboardInterrupt`` -> ``dsp.stop(); rx.stop(); sys.exit``).  The audio received is compared with
the oracle fed the same stream.
"""
import _thread
import os
import runpy
import sys
import threading
import time
import types

import numpy as np
import pytest

import radiocore_oracle as oracle
from bench_support import synth
from tests import parity

pytestmark = pytest.mark.gpu

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))


def _script():
    for base in ("/root/reference", os.path.join(ROOT, "oracle", "_ref")):
        p = os.path.join(base, "examples", "multi_fm_server.py")
        if os.path.exists(p):
            return p
    return None


class _Result:
    def __init__(self, ret):
        self.ret = ret


def _soapy_stub(block, rate_limit):
    """A SoapySDR module whose Device streams `block` (one second of IQ) cyclically."""
    mod = types.ModuleType("SoapySDR")
    mod.SOAPY_SDR_CF32, mod.SOAPY_SDR_RX = "CF32", 0

    class Device:
        calls = []

        def __init__(self, args):
            self.args, self.pos, self.open, self.t0, self.sent = args, 0, False, None, 0

        def setSampleRate(self, direction, channel, rate):
            Device.calls.append(("rate", rate))

        def setFrequency(self, direction, channel, freq):
            Device.calls.append(("freq", freq))

        def setGainMode(self, direction, channel, automatic):
            pass

        def setupStream(self, direction, fmt):
            return object()

        def activateStream(self, stream):
            self.open, self.t0 = True, time.perf_counter()

        def readStream(self, stream, buffs, count, timeoutUs=0):
            if not self.open:
                time.sleep(0.01)
                return _Result(0)
            while self.sent > rate_limit * (time.perf_counter() - self.t0):     # pace like a radio would
                time.sleep(0.0005)
            n = min(count, len(block) - self.pos)
            buffs[0][:n] = block[self.pos:self.pos + n]
            self.pos = (self.pos + n) % len(block)
            self.sent += n
            return _Result(n)

        def deactivateStream(self, stream):
            self.open = False

        def closeStream(self, stream):
            pass

    mod.Device = Device
    return mod


def _serve(out_path):
    """Child process: run the unmodified script until the subscriber has two blocks per station."""
    import zmq
    script = _script()
    for p in (ROOT, os.path.join(ROOT, "radio-core_b200"), os.path.join(ROOT, "oracle")):
        if p not in sys.path:
            sys.path.insert(0, p)
    import radiocore
    assert "radio-core_b200" in radiocore.__file__            # the drop-in, not the reference package
    block, f_in, freqs = _stream()
    stub = _soapy_stub(block, rate_limit=40e6)
    sys.modules["SoapySDR"] = stub
    got = {f: [] for f in freqs}

    def subscriber():
        ctx = zmq.Context.instance()
        sock = ctx.socket(zmq.SUB)
        sock.connect("tcp://127.0.0.1:5555")
        sock.setsockopt(zmq.SUBSCRIBE, b"")
        sock.setsockopt(zmq.RCVTIMEO, 200)
        deadline = time.time() + 120
        while time.time() < deadline:
            try:
                topic, payload = sock.recv_multipart()
            except zmq.Again:
                continue
            f = float(int.from_bytes(topic, "little"))
            if f in got:
                got[f].append(np.frombuffer(payload, dtype=np.float32).copy())
            if all(len(v) >= 2 for v in got.values()):
                break
        sock.close(0)
        _thread.interrupt_main()                # what Ctrl-C does to the script

    threading.Thread(target=subscriber, daemon=True).start()
    sys.argv = [script]
    code = 1
    try:
        runpy.run_path(script, run_name="__main__")
    except SystemExit:                          # the script's own KeyboardInterrupt handler: stop threads, sys.exit
        code = 0
    np.save(out_path, {"got": got, "calls": stub.Device.calls}, allow_pickle=True)
    sys.stdout.flush()
    os._exit(code)                              # skip interpreter teardown: the script's PUB socket is still open


N, B, A = 10_000_000, 240_000, 48_000           # the script's own Config: 10 Msps, three 240 kHz stations, 48 kHz audio
FREQS = (96.9e6, 94.5e6, 97.5e6)
KINDS = ("WBFM", "MFM", "FM")


def _oracle_tuner():
    o = oracle.Tuner()
    for f, k in zip(FREQS, KINDS):
        o.add_channel(f, B, getattr(oracle, k)(B, A, 75e-6))
    o.request_bandwidth(N)
    return o


def _stream():
    f_in = _oracle_tuner().input_frequency
    block = synth.wideband(N, [f - f_in for f in FREQS], B, seed=19, stereo=True, deviation=60e3)
    return block, f_in, FREQS


def test_multi_fm_server_unmodified(tmp_path):
    import subprocess
    pytest.importorskip("zmq")
    if _script() is None:
        pytest.skip("no copy of the reference's examples (run oracle/make_ref.py in the build container)")
    out = str(tmp_path / "served.npy")
    env = dict(os.environ, PYTHONPATH=os.pathsep.join([ROOT, os.path.join(ROOT, "radio-core_b200"), os.path.join(ROOT, "oracle")]))
    proc = subprocess.run([sys.executable, os.path.abspath(__file__), out], env=env, timeout=300,
                          stdout=subprocess.PIPE, stderr=subprocess.STDOUT)
    assert proc.returncode == 0, proc.stdout.decode(errors="replace")[-3000:]
    res = np.load(out, allow_pickle=True).item()
    got, calls = res["got"], res["calls"]

    block, f_in, _ = _stream()
    o = _oracle_tuner()
    want = []                                   # oracle audio of the first three blocks of the looped stream
    for _ in range(3):
        o.load(block)
        want.append([ch.demodulator.run(o.run(ch.index)) for ch in o.channels()])
    assert ("rate", float(N)) in calls and ("freq", f_in) in calls
    for i, (f, k) in enumerate(zip(FREQS, KINDS)):
        assert len(got[f]) >= 2, f"no audio received for {f}"
        nch = 2 if k == "WBFM" else 1
        for payload in got[f][:2]:
            assert payload.size == A * nch
            a = payload.reshape(A, nch)
            # a late subscriber may miss the first block; the stream repeats, so blocks >= 1 coincide
            errs = [parity.errors(a, np.asarray(w[i]).reshape(A, nch)) for w in want]
            rel_peak, margin = min(errs, key=lambda e: e[1])
            assert rel_peak <= parity.TOL and margin <= 1.0, (f, k, errs)


if __name__ == "__main__":
    _serve(sys.argv[1])
