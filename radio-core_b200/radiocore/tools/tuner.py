"""Multi-channel tuner (mirror of radiocore/tools/tuner.py:9-174).

Host side: the channel registry and the band plan.  Device side (C ABI,
``rc_engine_*``): ``load`` is one N-point forward FFT of the one-second block;
each channel is then the inverse B-point FFT of B+1 gathered, Hann-weighted
bins.  Because ``add_channel`` receives the demodulator instance, ``load``
knows every channel's chain, and the first ``demodulator.run(tuner.run(i))``
after a ``load`` computes *all* channels' audio with batched kernels; the
per-channel calls of ``examples/multi_fm_server.py:100-106`` then just pick up
their slice.
"""
import ctypes as C
from dataclasses import dataclass
from typing import List

import numpy as np
import torch

from radiocore import _device, _native
from radiocore.analog._demod import ChannelView, DemodBase

MODE_NONE = 3


@dataclass
class Channel:
    """Frequency boundaries and demodulator of one registered channel."""
    index: int
    bandwidth: float
    demodulator: None
    lower_frequency: float
    center_frequency: float
    higher_frequency: float

    @property
    def address_bytes(self) -> bytes:
        """Little-endian int32 centre frequency: the ZeroMQ topic of the channel."""
        return int(self.center_frequency).to_bytes(4, byteorder="little")


class Tuner:
    """Channelise one-second blocks of wideband IQ into the registered channels."""

    def __init__(self, cuda: bool = False):
        self._cuda = cuda
        self._input_frequency = 0.0
        self._input_bandwidth = 0.0
        self._bounds: List[Channel] = []
        self._engine = None
        self._engine_key = None
        self._layout = []
        self._audio_dev = None
        self._audio_host = None
        self._serial = 0
        self._audio_serial = -1
        self._host_serial = -1
        self._input_ref = None
        self._graph = None         # captured load + run_all of one block (step)
        self._subband = None       # (x_lo, x_len): the engine is handed a sub-band of the spectrum (sharding.ShardedLoad)
        self._stage_dev = None     # persistent device copy of a pinned host block (load)
        self._stage_ev = None
        self._pipe = None          # block pipeline state of submit()/collect()

    # ------------------------------------------------------------ band plan
    @property
    def input_frequency(self) -> float:
        """Centre frequency the wideband input must be tuned to."""
        return self._input_frequency

    @property
    def input_bandwidth(self) -> float:
        """Bandwidth (= samples per one-second block) of the wideband input."""
        return self._input_bandwidth

    def channels(self) -> List[Channel]:
        return self._bounds

    def request_bandwidth(self, bandwidth: float):
        """Widen the input bandwidth (e.g. to the SDR's fixed rate); never narrows."""
        if bandwidth < self._input_bandwidth:
            raise ValueError(f"requested bandwidth ({bandwidth}) is too low, "
                             f"minimum is {self._input_bandwidth}")
        self._input_bandwidth = bandwidth

    def add_channel(self, frequency: float, bandwidth: float, demodulator):
        half = bandwidth / 2
        self._bounds.append(Channel(len(self._bounds), bandwidth, demodulator,
                                    frequency - half, frequency, frequency + half))
        self._replan()

    def reset(self):
        """Forget all channels.  As in the reference (tuner.py:121-124,164) the re-plan of the now
        empty channel list then raises ``ValueError`` (``min()`` of an empty sequence) -- the
        channels are gone when it does, and ``add_channel`` can be called again."""
        self._bounds = []
        self._input_frequency = 0.0
        self._input_bandwidth = 0.0
        self._drop_engine()
        self._replan()

    def _replan(self):
        edges_lo = [c.lower_frequency for c in self._bounds]
        edges_hi = [c.higher_frequency for c in self._bounds]
        lo, hi = min(edges_lo), max(edges_hi)
        self._input_frequency = (lo + hi) / 2
        span = hi - lo
        # pad the span up to a multiple of the (floored) mean channel bandwidth
        mean_bw = sum(c.bandwidth for c in self._bounds) // len(self._bounds)
        self._input_bandwidth = span + (-span) % mean_bw

    # --------------------------------------------------------------- engine
    def _drop_engine(self):
        h, self._engine = self._engine, None
        self._engine_key = None
        self._graph = None
        if self._pipe is not None:
            self._pipe = None
            try:
                torch.cuda.synchronize()
            except Exception:
                pass
        if h is not None:
            try:
                _native.lib().rc_engine_destroy(h)
            except Exception:
                pass

    def __del__(self):
        self._drop_engine()

    def _spec(self):
        n = int(self._input_bandwidth)
        spec = []
        for ch in self._bounds:
            d = ch.demodulator
            roll = int(self._input_frequency - ch.center_frequency) % n
            if isinstance(d, DemodBase) and d._input_size == int(ch.bandwidth):
                spec.append((roll, int(ch.bandwidth), d._output_size, d._mode, d._deemphasis_rate))
            else:
                spec.append((roll, int(ch.bandwidth), 2, MODE_NONE, 75e-6))
        return n, tuple(spec), self._subband

    def _ensure_engine(self):
        if not self._bounds:
            raise ValueError("no channels registered")
        key = self._spec()
        if self._engine is not None and key == self._engine_key:
            return
        self._drop_engine()
        lib = _native.lib()
        n, spec, subband = key
        h = C.c_void_p()
        _native.check(lib.rc_engine_create(_device.device_index(), n, C.byref(h)))
        try:
            for roll, bw, audio, mode, tau in spec:
                _native.check(lib.rc_engine_add_channel(h, roll, bw, audio, mode, tau, None))
            if subband is not None:
                _native.check(lib.rc_engine_set_subband(h, int(subband[0]), int(subband[1])))
            _native.check(lib.rc_engine_commit(h))
        except Exception:
            lib.rc_engine_destroy(h)
            raise
        self._engine, self._engine_key = h, key
        total = C.c_int64()
        _native.check(lib.rc_engine_audio_floats(h, C.byref(total)))
        self._audio_dev = torch.empty(max(total.value, 1), dtype=torch.float32, device="cuda")
        self._audio_host = torch.empty(max(total.value, 1), dtype=torch.float32).pin_memory()
        self._layout = []
        for i in range(len(spec)):
            off, size, nch = C.c_int64(), C.c_int64(), C.c_int()
            _native.check(lib.rc_engine_channel_layout(h, i, C.byref(off), C.byref(size), C.byref(nch)))
            self._layout.append((off.value, size.value, nch.value))

    # ------------------------------------------------------------- hot path
    def load(self, input_signal):
        """Forward FFT of one second of wideband IQ (complex64, len == input_bandwidth)."""
        self._ensure_engine()
        if len(input_signal) != int(self._input_bandwidth):
            raise ValueError("input_signal size and input_bandwidth mismatch")
        if self._pipe is not None:                   # blocks queued by submit() share the engine's scratch
            self._pipe["compute"].synchronize()
        x = self._stage_input(input_signal)
        self._input_ref = x
        _native.check(_native.lib().rc_engine_load(self._engine, x.data_ptr(), _device.stream_ptr()))
        self._serial += 1

    # ------------------------------------------------ sharded load (sharding.ShardedLoad)
    def needed_bins(self):
        """(first_bin, count) of every channel's gathered bins, cyclic indices of the N-bin
        spectrum: channel i reads bins (k - roll) mod N for k in [-B/2, B/2] (tuner.py:151-161)."""
        n = int(self._input_bandwidth)
        out = []
        for ch in self._bounds:
            roll = int(self._input_frequency - ch.center_frequency) % n
            bw = int(ch.bandwidth)
            out.append(((-(bw // 2) - roll) % n, bw + 1))
        return out

    def set_subband(self, x_lo, x_len):
        """From now on ``load_subband`` hands the engine bins [x_lo, x_lo + x_len) (cyclic) of the
        block's spectrum instead of ``load`` handing it the block (rebuilds the engine)."""
        self._subband = None if x_lo is None else (int(x_lo), int(x_len))

    def load_subband(self, spectrum):
        """``Tuner.load`` for a sub-band computed elsewhere: a CUDA complex64 tensor holding bins
        [x_lo, x_lo + x_len) of fft(block); it is read (not copied) by the following run calls."""
        self._ensure_engine()
        if self._subband is None:
            raise RuntimeError("Tuner.set_subband must be called first")
        if not (isinstance(spectrum, torch.Tensor) and spectrum.is_cuda and spectrum.dtype == torch.complex64
                and spectrum.is_contiguous() and spectrum.numel() >= self._subband[1]):
            raise ValueError("spectrum must be a contiguous CUDA complex64 tensor of at least x_len bins")
        self._input_ref = spectrum
        _native.check(_native.lib().rc_engine_load_subband(self._engine, spectrum.data_ptr()))
        self._serial += 1

    def _stage_input(self, input_signal):
        """Device copy of one block.  A page-locked host array (``Buffer(cuda=True).data``, as
        examples/multi_fm_server.py:87 hands over) is DMA-ed asynchronously into a persistent
        device buffer on the current stream: the kernels queue behind the copy and ``load``
        returns without waiting for either.  The caller may refill the host buffer once the copy
        has finished; ``load`` therefore waits for the PREVIOUS block's copy only."""
        if isinstance(input_signal, torch.Tensor) and input_signal.is_cuda:
            return _device.to_device(input_signal, torch.complex64)
        host = input_signal if isinstance(input_signal, torch.Tensor) else None
        if host is None:
            a = np.asarray(input_signal)
            if a.dtype == np.complex64 and a.flags.c_contiguous:
                host = torch.from_numpy(a)
        if host is None or host.dtype != torch.complex64 or not host.is_pinned():
            return _device.to_device(input_signal, torch.complex64)
        n = host.numel()
        if self._stage_dev is None or self._stage_dev.numel() != n:
            self._stage_dev = torch.empty(n, dtype=torch.complex64, device="cuda")
            self._stage_ev = torch.cuda.Event()
        self._stage_dev.copy_(host, non_blocking=True)
        self._stage_ev.record()
        # the host array is the caller's to overwrite as soon as load() returns (the reference's
        # load is synchronous): wait for the DMA, not for the kernels queued behind it
        self._stage_ev.synchronize()
        return self._stage_dev

    def run(self, channel_index: int):
        """Channel ``channel_index`` of the loaded block, as a lazy ChannelView."""
        ch = self._bounds[int(channel_index)]
        if self._engine is None or self._serial == 0:
            raise RuntimeError("Tuner.load must be called before Tuner.run")
        return ChannelView(self, ch.index, self._serial, int(ch.bandwidth))

    def run_all(self, numpy_output: bool = False):
        """All channels' audio of the loaded block in one batched pass.

        Returns the packed float32 buffer (CUDA tensor, or pinned host tensor's
        NumPy view); ``audio_slices()`` gives each channel's (offset, size, nch).
        """
        self._compute_audio()
        if not numpy_output:
            return self._audio_dev
        self._fetch_host()
        return self._audio_host.numpy()

    # ------------------------------------------------- one block as one CUDA graph
    def _graph_state(self):
        self._ensure_engine()
        g = self._graph
        if g is None or g["engine"] is not self._engine:
            n = int(self._input_bandwidth)
            g = self._graph = {"engine": self._engine, "x": torch.empty(n, dtype=torch.complex64, device="cuda"),
                               "graph": None, "warm": 0}
        return g

    def graph_input(self):
        """The persistent device buffer the captured graph of ``step`` reads: a producer (the
        host-to-device copy of the ingest side, another kernel) that writes the block straight into
        it and then calls ``step(tuner.graph_input())`` saves the extra device copy."""
        return self._graph_state()["x"]

    def step(self, input_signal, numpy_output: bool = False):
        """``load`` + ``run_all`` of one block replayed as ONE CUDA graph.

        The kernel sequence of a block is fixed once the channels are registered, so it is
        captured on first use (after one eager warm-up block) and replayed afterwards: the
        block is copied into a persistent device buffer the captured kernels read, and one
        graph launch replaces the ~10-25 kernel launches of the block -- what matters for
        short blocks, whose kernels run for microseconds.  Same arithmetic and
        the same carried de-emphasis state as ``load`` + ``run_all``; returns like ``run_all``."""
        n = int(self._input_bandwidth)
        if len(input_signal) != n:
            raise ValueError("input_signal size and input_bandwidth mismatch")
        lib = _native.lib()
        g = self._graph_state()
        src = input_signal if isinstance(input_signal, torch.Tensor) else torch.from_numpy(
            np.ascontiguousarray(np.asarray(input_signal), dtype=np.complex64))
        if not (src.is_cuda and src.data_ptr() == g["x"].data_ptr()):     # a block written straight into graph_input() needs no copy
            g["x"].copy_(src, non_blocking=True)
        if g["graph"] is None and g["warm"] >= 1:
            torch.cuda.synchronize()
            graph = torch.cuda.CUDAGraph()
            with torch.cuda.graph(graph):
                _native.check(lib.rc_engine_load(self._engine, g["x"].data_ptr(), _device.stream_ptr()))
                _native.check(lib.rc_engine_run(self._engine, self._audio_dev.data_ptr(), _device.stream_ptr()))
            g["graph"] = graph
            # capture records the kernels without running them: this block still has to run
        if g["graph"] is not None:
            g["graph"].replay()
        else:
            _native.check(lib.rc_engine_load(self._engine, g["x"].data_ptr(), _device.stream_ptr()))
            _native.check(lib.rc_engine_run(self._engine, self._audio_dev.data_ptr(), _device.stream_ptr()))
            g["warm"] += 1
        self._serial += 1
        self._audio_serial = self._serial
        self._input_ref = g["x"]
        if not numpy_output:
            return self._audio_dev
        self._fetch_host()
        return self._audio_host.numpy()

    # ------------------------------------------------- pipelined block interface
    # The reference's loop (examples/multi_fm_server.py:95-106) is synchronous: get a block,
    # load, demodulate, send.  A one-second block is 8 bytes per sample of host->device copy,
    # which at PCIe rates takes longer than the kernels; submit()/collect() keep `depth` blocks
    # in flight so the copy of block k+1 overlaps the kernels of block k and the read-back of
    # block k-1.  Same arithmetic, same carried de-emphasis state, blocks processed in order.
    def submit(self, input_signal, depth: int = 2):
        """Queue one block (pinned host array or CUDA tensor); returns a ticket for collect()."""
        self._ensure_engine()
        n = int(self._input_bandwidth)
        if len(input_signal) != n:
            raise ValueError("input_signal size and input_bandwidth mismatch")
        if self._pipe is None or self._pipe["depth"] != depth or self._pipe["engine"] is not self._engine:
            total = max(self._audio_dev.numel(), 1)
            self._pipe = {
                "depth": depth, "engine": self._engine, "next": 0,
                "copy": torch.cuda.Stream(), "compute": torch.cuda.Stream(), "out": torch.cuda.Stream(),
                "x": [torch.empty(n, dtype=torch.complex64, device="cuda") for _ in range(depth)],
                "audio": [torch.empty(total, dtype=torch.float32, device="cuda") for _ in range(depth)],
                "host": [torch.empty(total, dtype=torch.float32).pin_memory() for _ in range(depth)],
                "ev_in": [torch.cuda.Event() for _ in range(depth)],
                "ev_loaded": [None] * depth,
                "ev_done": [torch.cuda.Event() for _ in range(depth)],
                "ev_out": [torch.cuda.Event() for _ in range(depth)],
            }
            # order the pipeline after whatever the caller queued on the current stream
            for st in ("copy", "compute", "out"):
                self._pipe[st].wait_stream(torch.cuda.current_stream())
        p = self._pipe
        ticket = p["next"]
        slot = ticket % depth
        p["next"] += 1
        lib = _native.lib()
        src = input_signal if isinstance(input_signal, torch.Tensor) else torch.from_numpy(
            np.ascontiguousarray(np.asarray(input_signal), dtype=np.complex64))
        # everything the caller queued on its stream (the producer of a CUDA `src`, a load() or
        # run_all() on the engine's scratch) is ordered before this block's copy and kernels
        cur = torch.cuda.current_stream()
        p["copy"].wait_stream(cur)
        p["compute"].wait_stream(cur)
        if src.is_cuda:
            src.record_stream(p["copy"])                  # the allocator must not recycle it under the copy
        with torch.cuda.stream(p["copy"]):
            if p["ev_loaded"][slot] is not None:          # the slot's previous block has been transformed
                p["copy"].wait_event(p["ev_loaded"][slot])
            p["x"][slot].copy_(src, non_blocking=True)
            p["ev_in"][slot].record(p["copy"])
        with torch.cuda.stream(p["compute"]):
            p["compute"].wait_event(p["ev_in"][slot])
            p["compute"].wait_event(p["ev_out"][slot])    # the slot's previous audio has been read back
            _native.check(lib.rc_engine_load(self._engine, p["x"][slot].data_ptr(), p["compute"].cuda_stream))
            ev = torch.cuda.Event()
            ev.record(p["compute"])
            p["ev_loaded"][slot] = ev
            _native.check(lib.rc_engine_run(self._engine, p["audio"][slot].data_ptr(), p["compute"].cuda_stream))
            p["ev_done"][slot].record(p["compute"])
        with torch.cuda.stream(p["out"]):
            p["out"].wait_event(p["ev_done"][slot])
            p["host"][slot].copy_(p["audio"][slot], non_blocking=True)
            p["ev_out"][slot].record(p["out"])
        self._serial += 1
        self._audio_serial = -1
        return ticket

    def collect(self, ticket: int):
        """Packed float32 audio of a submitted block (NumPy view of pinned memory, valid until
        `depth` more blocks are submitted); ``audio_slices()`` locates each channel."""
        p = self._pipe
        if p is None or ticket >= p["next"] or ticket < p["next"] - p["depth"]:
            raise RuntimeError("unknown or expired ticket")
        slot = ticket % p["depth"]
        p["ev_out"][slot].synchronize()
        return p["host"][slot].numpy()

    def audio_slices(self):
        self._ensure_engine()
        return list(self._layout)

    def _compute_audio(self):
        if self._audio_serial == self._serial:
            return
        if self._engine is None or self._serial == 0:
            raise RuntimeError("Tuner.load must be called before the channels are demodulated")
        _native.check(_native.lib().rc_engine_run(self._engine, self._audio_dev.data_ptr(),
                                                  _device.stream_ptr()))
        self._audio_serial = self._serial

    def _fetch_host(self):
        if self._host_serial == self._serial:
            return
        self._audio_host.copy_(self._audio_dev, non_blocking=True)
        torch.cuda.current_stream().synchronize()
        self._host_serial = self._serial

    def _channel_iq(self, index, serial):
        if serial != self._serial:
            raise RuntimeError("stale channel view: Tuner.load was called again")
        bw = int(self._bounds[index].bandwidth)
        out = torch.empty(bw, dtype=torch.complex64, device="cuda")
        _native.check(_native.lib().rc_engine_channel_iq(self._engine, index, out.data_ptr(),
                                                         _device.stream_ptr()))
        return out

    def _audio_for(self, view, demod, numpy_output):
        """Audio of ``view``'s channel if ``demod`` is the instance registered for it."""
        if view.tuner is not self or view.serial != self._serial:
            return None
        ch = self._bounds[view.index]
        off, size, nch = self._layout[view.index]
        if ch.demodulator is not demod or nch == 0:
            return None
        self._compute_audio()
        if numpy_output:
            self._fetch_host()
            flat = self._audio_host.numpy()[off: off + size * nch]
            a = flat.reshape(size, nch).copy()
            return a[None] if nch == 2 else a
        # a fresh tensor, like the reference's return value: the packed buffer is rewritten per block
        a = self._audio_dev[off: off + size * nch].clone().view(size, nch)
        return a.unsqueeze(0) if nch == 2 else a
