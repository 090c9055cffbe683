"""CPU-side checks of the product package: the C-ABI library loads and exports
every symbol include/radiocore_b200.h declares (no compute calls), the band
plan and plumbing behave like the reference's, and nothing in the product
imports the oracle."""
import ctypes as C
import os
import re

import numpy as np
import pytest

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
LIB = os.path.join(ROOT, "radio-core_b200", "radiocore", "_native", "libradiocore_b200.so")


def _declared_symbols():
    text = open(os.path.join(ROOT, "include", "radiocore_b200.h")).read()
    return sorted(set(re.findall(r"\b(rc_[a-z0-9_]+)\s*\(", text)))


@pytest.fixture(scope="module")
def built():
    if not os.path.exists(LIB):
        import __graft_entry__ as g
        g.build()
    return LIB


def test_library_exports_every_declared_symbol(built):
    from radiocore import _native
    lib = _native.load_library()
    declared = _declared_symbols()
    assert len(declared) >= 30
    for name in declared:
        assert hasattr(lib, name), name
    assert sorted(_native.EXPORTED_SYMBOLS) == declared
    assert lib.rc_version() >= 100
    assert lib.rc_size_supported(256_000_000) == 1 and lib.rc_size_supported(14) == 0


def test_host_side_tap_design_matches_oracle(built):
    import radiocore
    import radiocore_oracle as oracle
    for size, tau in ((48000, 75e-6), (48000, 50e-6), (32000, 75e-6)):
        taps, zi = radiocore.Deemphasis(size, tau).taps
        ref = oracle.Deemphasis(size, tau)
        assert np.array_equal(taps, ref.taps)
        assert np.array_equal(zi, ref.state)


def test_band_plan_matches_oracle():
    import radiocore
    import radiocore_oracle as oracle
    for freqs, bw in (((96.9e6, 94.5e6, 97.5e6), 240e3), ((100e6,), 250e3), ((1e6, 1.3e6, 0.9e6), 200e3)):
        a, b = radiocore.Tuner(), oracle.Tuner()
        for f in freqs:
            a.add_channel(f, bw, None)
            b.add_channel(f, bw, None)
        assert a.input_frequency == b.input_frequency
        assert a.input_bandwidth == b.input_bandwidth
        assert [c.address_bytes for c in a.channels()] == [c.address_bytes for c in b.channels()]
    a.request_bandwidth(10e6)
    assert a.input_bandwidth == 10e6
    with pytest.raises(ValueError):
        a.request_bandwidth(1e3)


def test_no_gpu_means_loud_failure():
    import torch
    if torch.cuda.is_available():
        pytest.skip("GPU present")
    import radiocore
    assert radiocore.HasCuda() is False
    with pytest.raises(RuntimeError):
        radiocore.FM(1000, 100).run(np.zeros(1000, dtype=np.complex64))


def test_product_never_imports_the_oracle():
    pkg = os.path.join(ROOT, "radio-core_b200")
    for dirpath, _, files in os.walk(pkg):
        for f in files:
            if f.endswith((".py", ".cu", ".cuh", ".h")):
                text = open(os.path.join(dirpath, f)).read()
                for bad in ("radiocore_oracle", "import scipy", "from scipy", "ref_shim", "librc_emulate"):
                    assert bad not in text, (bad, os.path.join(dirpath, f))


def test_ringbuffer_and_buffer_plumbing():
    import radiocore
    rb = radiocore.RingBuffer(8, dtype="float32", print_overflow=False)
    assert rb.capacity == 8 and rb.occupancy == 0
    rb.put(np.arange(5, dtype=np.float32))
    out = np.zeros(3, dtype=np.float32)
    assert rb.get(out) and list(out) == [0, 1, 2]
    rb.put(np.arange(5, 10, dtype=np.float32))        # wraps
    out = np.zeros(7, dtype=np.float32)
    assert rb.get(out) and list(out) == [3, 4, 5, 6, 7, 8, 9]
    assert rb.get(np.zeros(1, dtype=np.float32), timeout=0.01) is False
    b = radiocore.Buffer(8, "float32")
    with b.consume() as a:
        a[:] = 1
    assert b.data.sum() == 8 and len(b) == 8


def test_wire_format_round_trip():
    """Egress of the server loop (multi_fm_server.py:103-106): topic = Channel.address_bytes
    (int32-LE centre frequency), payload = float32 audio bytes; the receiver
    (multi_fm_receiver.py:23-24,47-49) subscribes by that prefix and reshapes to (samples, channels)."""
    zmq = pytest.importorskip("zmq")
    import time
    import radiocore
    tuner = radiocore.Tuner()
    for f in (96.9e6, 94.5e6):
        tuner.add_channel(f, 240e3, None)
    want = tuner.channels()[1]
    assert want.address_bytes == int(94.5e6).to_bytes(4, byteorder="little")
    ctx = zmq.Context()
    pub, sub = ctx.socket(zmq.PUB), ctx.socket(zmq.SUB)
    try:
        port = pub.bind_to_random_port("tcp://127.0.0.1")
        sub.connect(f"tcp://127.0.0.1:{port}")
        sub.setsockopt(zmq.SUBSCRIBE, int(94.5e6).to_bytes(4, byteorder="little"))
        sub.setsockopt(zmq.RCVTIMEO, 5000)
        stereo = np.random.default_rng(0).uniform(-1, 1, (1, 480, 2)).astype(np.float32)   # WBFM.run's shape
        for _ in range(50):                      # slow joiner: publish until the subscription is live
            for ch in tuner.channels():
                pub.send_multipart([ch.address_bytes, stereo.tobytes()])
            try:
                topic, payload = sub.recv_multipart(flags=zmq.NOBLOCK)
                break
            except zmq.Again:
                time.sleep(0.05)
        else:
            pytest.fail("no frame received")
        assert topic == want.address_bytes      # only the subscribed station arrives
        audio = np.frombuffer(payload, dtype=np.float32)
        audio = audio.reshape((len(audio) // 2, 2))
        assert np.array_equal(audio, stereo[0])
    finally:
        pub.close(0)
        sub.close(0)
        ctx.term()


def test_carrousel_chopper_semantics():
    """Behaviour the reference's tests/test_carrousel.py pins: a full ring overwrites its oldest item
    and counts the overflow; Buffer items are handed out through their own consume()."""
    import radiocore
    ring = radiocore.Carrousel([[0], [0], [0]], print_overflow=False)
    assert ring.is_empty and not ring.is_healthy and ring.capacity == 3
    for v in (1, 2, 3, 4):
        with ring.enqueue() as item:
            item[0] = v
    assert ring.is_full and ring.occupancy == 3 and ring.overflow == 1
    got = []
    while not ring.is_empty:
        with ring.dequeue() as item:
            got.append(item[0])
    assert got == [2, 3, 4]
    with pytest.raises(ValueError):
        ring.dequeue()
    bufs = radiocore.Carrousel([radiocore.Buffer(4, dtype="float32", lock=True) for _ in range(2)])
    with bufs.enqueue() as arr:
        assert bufs._items[0].is_locked
        arr[:] = 7
    assert not bufs._items[0].is_locked
    with bufs.dequeue() as arr:
        assert np.all(arr == 7)
    chop = radiocore.Chopper(12, 4)
    data = np.arange(12)
    parts = list(chop.chop(data))
    assert [p.tolist() for p in parts] == [[0, 1, 2, 3], [4, 5, 6, 7], [8, 9, 10, 11]] and parts[0].base is data
    assert (chop.size, chop.chunk_size) == (12, 4)
    with pytest.raises(ValueError):
        radiocore.Chopper(10, 4)
    ringbuf = radiocore.RingBuffer(8, dtype=np.float32, allow_overflow=False)
    ringbuf.put([1, 2, 3, 4, 5, 6])
    with pytest.raises(ValueError):
        ringbuf.put([7, 8, 9])
    assert np.allclose(ringbuf.data[:6], [1, 2, 3, 4, 5, 6])
