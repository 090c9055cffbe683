"""Multi-GPU plumbing of the receive path: one process per GPU, channels sharded.

The reference is single-process (examples/multi_fm_server.py:113-140); scaling it across the
GPUs of one box is new here.  Channels are independent after ``Tuner.load``, so each rank keeps
a contiguous slice of the channel list and its carried de-emphasis state; the only exchange is
one broadcast of the wideband block (complex64, 8*N bytes) when all ranks listen to the same
stream.  There is no reduction and no gather of audio: every rank emits its own channels.

Works on any ``torch.distributed`` backend (NCCL on the GPUs, gloo in the CPU tests).
"""
import collections
import os
from typing import List, Optional, Sequence

import torch
import torch.distributed as dist


def channel_slice(n_channels: int, world_size: int, rank: int) -> range:
    """Contiguous, balanced slice of ``range(n_channels)`` owned by ``rank``
    (the first ``n_channels % world_size`` ranks get one channel more)."""
    if not (0 <= rank < world_size):
        raise ValueError("rank out of range")
    base, extra = divmod(n_channels, world_size)
    start = rank * base + min(rank, extra)
    return range(start, start + base + (1 if rank < extra else 0))


def owner_of(channel: int, n_channels: int, world_size: int) -> int:
    """Rank that owns ``channel`` under ``channel_slice``."""
    base, extra = divmod(n_channels, world_size)
    edge = extra * (base + 1)
    if channel < edge:
        return channel // (base + 1)
    return extra + (channel - edge) // max(base, 1)


def shard_tuner(tuner, centers: Sequence[float], bandwidth: float, make_demodulator, input_frequency: float,
                input_bandwidth: float, world_size: int, rank: int) -> List[int]:
    """Register this rank's slice of the channel list on ``tuner`` while keeping the band plan of
    the FULL list: the reference derives ``input_frequency`` from the registered channels
    (tools/tuner.py:163-174), so a slice alone would re-centre the plan.  Returns the global
    indices of the registered channels."""
    mine = list(channel_slice(len(centers), world_size, rank))
    for c in mine:
        tuner.add_channel(centers[c], bandwidth, make_demodulator(c))
    # plan of the whole band, not of the slice (the attribute is private in the package's Tuner,
    # public in the oracle's)
    if hasattr(tuner, "_input_frequency"):
        tuner._input_frequency = float(input_frequency)
    else:
        tuner.input_frequency = float(input_frequency)
    tuner.request_bandwidth(input_bandwidth)
    return mine


def broadcast_block(block: torch.Tensor, src: int = 0) -> torch.Tensor:
    """Broadcast one wideband block (complex64 tensor, in place) from ``src`` to every rank."""
    if not dist.is_initialized() or dist.get_world_size() == 1:
        return block
    dist.broadcast(torch.view_as_real(block) if block.is_complex() else block, src=src)
    return block


class BlockBroadcaster:
    """Double-buffered broadcast of the wideband block: block k+1 travels while block k is in
    the kernels.

    ``post()`` starts the (asynchronous) broadcast of the next block, ``take()`` returns the
    oldest posted block once it has arrived -- on NCCL "arrived" is a stream dependency, the host
    does not block.  The source rank sends straight from the tensor it passes to ``post`` (no
    staging copy; it must leave that tensor alone until the block has been taken); the other
    ranks receive into ``depth`` rotating device buffers, so a buffer is only overwritten after
    the work queued on it ``depth`` posts earlier, which the collective is ordered behind.
    Every rank must call post/take in the same order.  With one rank it degenerates to a queue.
    """

    def __init__(self, n_samples: int, device, src: int = 0, depth: int = 2):
        if depth < 2:
            raise ValueError("depth must be at least 2")
        self._src = int(src)
        self._world = dist.get_world_size() if dist.is_initialized() else 1
        self._rank = dist.get_rank() if dist.is_initialized() else 0
        self._is_src = self._world == 1 or self._rank == self._src
        self._n = int(n_samples)
        self._slots = [] if self._is_src else [torch.empty(self._n, dtype=torch.complex64, device=device)
                                                for _ in range(depth)]
        self._depth = depth
        self._turn = 0
        self._pending = collections.deque()

    def post(self, block: Optional[torch.Tensor] = None) -> None:
        """Start sending (source rank: ``block`` required) / receiving the next block."""
        if len(self._pending) >= self._depth:
            raise RuntimeError("too many blocks in flight: take() before the next post()")
        if self._is_src:
            if block is None or block.numel() != self._n or block.dtype != torch.complex64:
                raise ValueError("the source rank posts a complex64 block of the configured size")
            buf = block
        else:
            buf = self._slots[self._turn]
            self._turn = (self._turn + 1) % self._depth
        work = None
        if self._world > 1:
            work = dist.broadcast(torch.view_as_real(buf), src=self._src, async_op=True)
        self._pending.append((buf, work))

    def take(self) -> torch.Tensor:
        """Oldest posted block, ordered after its arrival."""
        if not self._pending:
            raise RuntimeError("take() without a posted block")
        buf, work = self._pending.popleft()
        if work is not None:
            work.wait()
        return buf

    def in_flight(self) -> int:
        return len(self._pending)


# =====================================================================================
# Sharded Tuner.load: the N-point FFT itself divided over the ranks
# =====================================================================================
# BlockBroadcaster replicates the block, so every GPU repeats the whole N-point FFT
# (tools/tuner.py:137-138) and receives 8*N bytes per block over NVLink.  The sharded load
# divides both: with G ranks, M = N/G and P = M/G,
#
#   rank g holds the commutator branch x_g[m] = x[G*m + g]                (polyphase input commutator)
#   F_g = FFT_M(x_g)                                                      local, 1/G of the FFT work
#   exchange 1 (all-to-all, equal pieces): rank p collects F_g[k0], k0 in [p*P, (p+1)*P), of every g
#   X[k0 + M*k1] = sum_g W_G^{g k1} W_N^{g k0} F_g[k0]                    the remaining radix-G step, local
#   exchange 2: rank p sends each rank d the bins of d's sub-band that p now holds
#
# after which rank d owns the contiguous (cyclic) sub-band [x_lo_d, x_lo_d + x_len_d) its channels
# gather from (Tuner.needed_bins) -- a halo of neighbouring bins included, so no channel needs a
# second rank.  Per rank and block 2 * 8*N/G * (G-1)/G bytes cross NVLink instead of 8*N*(G-1)/G,
# there is no reduction, and the result is the same N-point DFT (SURVEY.md 7.3-1 with D = G).


def covering_arc(intervals: Sequence, n: int):
    """Smallest cyclic arc (start, length) of Z_n containing every (first, count) interval;
    start is even (the channel gather stages even-aligned bins by TMA)."""
    if not intervals:
        raise ValueError("no intervals")
    best = None
    starts = sorted({int(s) % n for s, _ in intervals})
    for s0 in starts:
        length = max(((int(s) - s0) % n) + int(c) for s, c in intervals)
        if best is None or length < best[1]:
            best = (s0, length)
    lo, length = best
    if lo % 2:
        lo, length = lo - 1, length + 1
    return lo % n, min(length, n)


def _arc_parts(lo: int, length: int, n: int):
    """The cyclic arc as linear (first_bin, count, position_in_arc) parts of [0, n)."""
    if lo + length <= n:
        return [(lo, length, 0)]
    return [(lo, n - lo, 0), (0, lo + length - n, n - lo)]


class SubbandPlan:
    """Who sends which bins to whom in exchange 2 (pure host arithmetic; every rank builds the
    same plan from the list of all ranks' arcs)."""

    def __init__(self, n: int, world: int, arcs: Sequence):
        if n % (world * world):
            raise ValueError("the block length must be a multiple of world_size**2")
        self.n, self.world = int(n), int(world)
        self.m = self.n // self.world            # local FFT length, and the stride between a piece's bin blocks
        self.p = self.m // self.world            # piece length
        self.arcs = [(int(a), int(b)) for a, b in arcs]

    def runs(self, src: int, dst: int):
        """[(k1, j_lo, j_hi, pos)]: rank ``src`` holds bins k1*M + src*P + j as Y[k1][j]; those with
        j in [j_lo, j_hi) go to position ``pos`` onward of ``dst``'s sub-band.  Ordered by (k1, part)
        -- the order both ends enumerate them in."""
        out = []
        lo, length = self.arcs[dst]
        for k1 in range(self.world):
            a = k1 * self.m + src * self.p               # first bin of this block of the piece
            for first, count, pos0 in _arc_parts(lo, length, self.n):
                b0, b1 = max(a, first), min(a + self.p, first + count)
                if b0 < b1:
                    out.append((k1, b0 - a, b1 - a, pos0 + (b0 - first)))
        return out


class ShardedLoad:
    """Distributed ``Tuner.load``: every rank posts its commutator branch of the block and takes
    back the sub-band of the spectrum its own channels need.

        load = ShardedLoad(tuner)          # after the channels are registered (collective: arcs are exchanged)
        load.post(x_branch)                # x[rank::world] of block k+1: local FFT, exchanges, combine -- on a side stream
        tuner.load_subband(load.take())    # block k: ordered after its arrival
        audio = tuner.run_all()

    ``depth`` sub-band buffers rotate, so block k+1 travels while block k is in the channel kernels.
    ``kernels`` supplies the local arithmetic (the CUDA library by default; the gloo CPU tests pass
    a NumPy stand-in to check the plan and the exchanges).
    """

    def __init__(self, tuner, depth: int = 0, kernels=None, group=None, lanes: int = 0):
        self._world = dist.get_world_size(group) if dist.is_initialized() else 1
        self._rank = dist.get_rank(group) if dist.is_initialized() else 0
        self._group = group
        n = int(tuner.input_bandwidth)
        arc = covering_arc(tuner.needed_bins(), n)
        arcs = [None] * self._world
        if self._world > 1:
            dist.all_gather_object(arcs, arc, group=group)
        else:
            arcs = [arc]
        self.plan = SubbandPlan(n, self._world, arcs)
        self.x_lo, self.x_len = arcs[self._rank]
        tuner.set_subband(self.x_lo, self.x_len)
        # Pipeline lanes: consecutive blocks alternate between `lanes` independent side streams (own FFT
        # plan, own piece buffer R, own barrier channels), so the HBM-bound local FFT of block k+2 runs
        # beside the NVLink-bound exchange of block k+1.  `lanes` blocks can be posted ahead of take().
        if lanes <= 0:
            lanes = int(os.environ.get("RC_SHARD_LANES", "1")) if kernels is None else 1   # measured: 2 lanes gain nothing
        self.lanes = max(1, lanes)
        plan = self.plan
        self._ks = [kernels] if kernels is not None else [_NativeKernels(plan, self._rank) for _ in range(self.lanes)]
        self.lanes = len(self._ks)
        self._k = k = self._ks[0]
        self._send = [plan.runs(self._rank, d) for d in range(self._world)]
        self._recv = [plan.runs(p, self._rank) for p in range(self._world)]
        self._depth = max(self.lanes + 1, depth)
        pad = 1 << 16                                               # the gather's tensor map describes whole rows past x_len
        slot_len = max(b for _, b in arcs) + pad
        self._Fs = self._Ys = None                                  # staging arrays of the unfused paths (allocated below)
        # Transport of the two exchanges.  "peer": the receive buffers (R and the sub-band slots) live in
        # symmetric memory mapped into every rank (torch.distributed._symmetric_memory); a rank PUSHES its
        # pieces / runs straight into the peers' buffers with device-to-device copies over NVLink and a
        # stream-ordered barrier closes each exchange -- no NCCL send/recv kernels competing for SMs, and
        # the copies run at link rate.  "collective": torch.distributed all-to-all / batched send-recv
        # (NCCL or gloo) into private buffers -- the portable path the CPU tests exercise.
        self._peer = None
        want = os.environ.get("RC_SHARD_TRANSPORT", "peer" if kernels is None else "collective")
        if want == "peer" and self._world > 1:
            try:
                self._peer = _PeerBuffers(plan.m, slot_len, self._depth, self._world, self._rank, group, self.lanes)
            except Exception as exc:                                # pragma: no cover - depends on the box
                import warnings
                warnings.warn(f"radiocore: symmetric-memory transport unavailable ({exc!r}); using NCCL send/recv")
        self.transport = "peer" if self._peer is not None else "collective"
        if self._peer is not None:
            self._Rs, self._slots = self._peer.R, self._peer.slots
        else:
            self._Rs = [k.empty(plan.m) for _ in range(self.lanes)]  # [G][P]: piece `rank` of every F_g
            self._slots = [k.empty(slot_len) for _ in range(self._depth)]
        self.phase_events = None                                    # set to [] to record (name, event) pairs per post()
        # With peer-mapped buffers the two exchanges can be the STORES of the kernels themselves: the
        # last pass of the local FFT scatters each piece into the rank that combines it, and the combine
        # writes every bin into the sub-band of the rank(s) that read it -- compute and NVLink transfer in
        # one kernel each, a barrier after each, no staging arrays and no copies.  RC_SHARD_FUSED=0 keeps
        # the kernels local and pushes with device-to-device copies instead.
        mode = os.environ.get("RC_SHARD_FUSED", "1")
        ok = self._peer is not None and hasattr(k, "fft_scatter") and plan.p % 2 == 0
        self.fused_fft = ok and mode in ("1", "fft")            # exchange 1 = stores of the FFT's last pass
        self.fused_combine = ok and mode in ("1", "combine")     # exchange 2 = stores of the combine
        self.fused = self.fused_fft and self.fused_combine
        if not self.fused_fft:
            self._Fs = [k.empty(plan.m) for _ in range(self.lanes)]     # F_g, natural order: piece p = [p*P, (p+1)*P)
        if not self.fused_combine:
            self._Ys = [k.empty(plan.m) for _ in range(self.lanes)]     # [G][P]: bins k1*M + rank*P + j
        if ok:
            g, p8 = self._rank, 8 * plan.p
            self._piece_bases = [[self._peer.ptrs[d] + self._peer.r_offset_bytes(ln) + g * p8 for d in range(self._world)]
                                 for ln in range(self.lanes)]
            self._segs = []
            for idx in range(self._depth):
                segs = []
                for d in range(self._world):
                    base = self._peer.ptrs[d] + self._peer.slot_offset_bytes(idx)
                    segs += [(k1, j0, j1, base + 8 * pos) for k1, j0, j1, pos in self._send[d]]
                self._segs.append(k.make_segments(segs))
        self._turn = 0
        self._block = 0
        self._last_fft_done = None
        self._pending = collections.deque()
        self.bytes_exchanged = 8 * (2 * plan.m - 2 * plan.p) if self._world > 1 else 0   # sent per block, both exchanges (halo aside)

    # ---- the two exchanges
    def _exchange_pieces(self, lane):
        """R[g] = F_g[rank*P : (rank+1)*P] from every rank g (equal-split all-to-all)."""
        F, R = self._Fs[lane], self._Rs[lane]
        if self._world == 1:
            R.copy_(F)
            return
        if self._peer is not None:
            p, g = self.plan.p, self._rank
            for step in range(self._world):                          # start with the neighbour: spread the traffic
                d = (g + step) % self._world
                self._peer.R_of(d, lane)[g * p:(g + 1) * p].copy_(F[d * p:(d + 1) * p], non_blocking=True)
            self._peer.barrier(2 * lane)
            return
        out, inp = _as_real(R), _as_real(F)
        if dist.get_backend(self._group) == "nccl":
            dist.all_to_all_single(out, inp, group=self._group)
            return
        p, ops = self.plan.p, []
        for r in range(self._world):
            if r == self._rank:
                out[r * p:(r + 1) * p].copy_(inp[r * p:(r + 1) * p])
            else:
                ops.append(dist.P2POp(dist.isend, inp[r * p:(r + 1) * p], r, group=self._group))
                ops.append(dist.P2POp(dist.irecv, out[r * p:(r + 1) * p], r, group=self._group))
        for w in dist.batch_isend_irecv(ops):
            w.wait()

    def _exchange_bins(self, slot, lane):
        """Exchange 2: the runs of SubbandPlan straight from Y into the destination sub-bands."""
        p = self.plan.p
        if self._peer is not None:
            idx = next(i for i, t in enumerate(self._slots) if t is slot)
            for step in range(self._world):
                d = (self._rank + step) % self._world
                dst = self._peer.slot_of(d, idx)
                for k1, j0, j1, pos in self._send[d]:
                    dst[pos: pos + (j1 - j0)].copy_(self._Ys[lane][k1 * p + j0: k1 * p + j1], non_blocking=True)
            self._peer.barrier(2 * lane + 1)
            return
        Y, X = _as_real(self._Ys[lane]), _as_real(slot)
        ops = []
        for d in range(self._world):
            for k1, j0, j1, pos in self._send[d]:
                src = Y[k1 * p + j0: k1 * p + j1]
                if d == self._rank:
                    X[pos: pos + (j1 - j0)].copy_(src)
                else:
                    ops.append(dist.P2POp(dist.isend, src, d, group=self._group))
        for q in range(self._world):
            if q == self._rank:
                continue
            for k1, j0, j1, pos in self._recv[q]:
                ops.append(dist.P2POp(dist.irecv, X[pos: pos + (j1 - j0)], q, group=self._group))
        if ops:
            for w in dist.batch_isend_irecv(ops):
                w.wait()

    # ---- block interface
    def post(self, x_branch, ready=None) -> None:
        """Start the load of the next block from this rank's branch x[rank::world] (M samples).
        ``ready``: optional CUDA event after which ``x_branch`` is valid (e.g. its host-to-device
        copy on another stream); without it the branch is ordered after the caller's stream."""
        if len(self._pending) >= self._depth:
            raise RuntimeError("too many blocks in flight: take() before the next post()")
        if x_branch.numel() != self.plan.m:
            raise ValueError("the branch holds N / world_size samples")
        slot = self._slots[self._turn]
        self._turn = (self._turn + 1) % self._depth
        lane = self._block % self.lanes
        self._block += 1
        k = self._k = self._ks[lane]
        done = k.begin(x_branch, slot, ready)
        try:
            self._post_on_lane(k, lane, x_branch, slot)
        finally:
            ev = k.end(done)                        # always leave the side-stream context
        self._pending.append((slot, ev, k))

    def _post_on_lane(self, k, lane, x_branch, slot):
        self._mark("begin")
        if self.fused_fft:
            k.fft_scatter(x_branch, self._piece_bases[lane])
            self._mark("local_fft+scatter")
            self._last_fft_done = getattr(k, "_fft_done", None)
            self._peer.barrier(2 * lane)
            self._mark("barrier_pieces")
        else:
            k.fft(x_branch, self._Fs[lane])
            self._mark("local_fft")
            self._last_fft_done = getattr(k, "_fft_done", None)
            self._exchange_pieces(lane)
            self._mark("exchange_pieces")
        if self.fused_combine:
            idx = next(i for i, t in enumerate(self._slots) if t is slot)
            k.combine_scatter(self._Rs[lane], self._rank * self.plan.p, self._segs[idx])
            self._mark("combine+scatter")
            self._peer.barrier(2 * lane + 1)
            self._mark("barrier_bins")
        else:
            k.combine(self._Rs[lane], self._Ys[lane], self._rank * self.plan.p)
            self._mark("combine")
            self._exchange_bins(slot, lane)
            self._mark("exchange_bins")

    def _mark(self, name):
        if self.phase_events is not None and hasattr(self._k, "mark"):
            self.phase_events.append((name, self._k.mark()))

    def phase_ms(self):
        """Mean device milliseconds of each phase of post() over the recorded blocks (side stream)."""
        ev, out, count = self.phase_events or [], {}, {}
        for (n0, e0), (n1, e1) in zip(ev, ev[1:]):
            if n1 == "begin":
                continue
            out[n1] = out.get(n1, 0.0) + e0.elapsed_time(e1)
            count[n1] = count.get(n1, 0) + 1
        return {k: v / count[k] for k, v in out.items()}

    def take(self):
        """Sub-band of the oldest posted block (bins [x_lo, x_lo + x_len)), ordered after its arrival."""
        if not self._pending:
            raise RuntimeError("take() without a posted block")
        slot, ev, k = self._pending.popleft()
        k.wait(ev, slot, self._last_fft_done)
        return slot

    def in_flight(self) -> int:
        return len(self._pending)


def _as_real(t):
    return torch.view_as_real(t) if t.is_complex() else t


class _NativeKernels:
    """Local arithmetic of ShardedLoad on the GPU (C ABI: rc_fft_*, rc_subband_combine); the whole
    pipeline of a block runs on a side stream so that it overlaps the channel kernels of the
    previous block on the caller's stream."""

    def __init__(self, plan, rank):
        import ctypes as C
        from radiocore import _device, _native
        self._native, self._C = _native, C
        self._plan = plan
        self._dev = _device.device_index()
        # RC_SHARD_PRIORITY=1 runs the pipeline on a high-priority stream (measured: no effect, 4.00 vs
        # 3.91 ms per block at 2 GPUs -- the phases of the two streams do not overlap usefully either way)
        prio = -1 if os.environ.get("RC_SHARD_PRIORITY", "0") == "1" else 0
        self._stream = torch.cuda.Stream(priority=prio)
        self._fft = C.c_void_p()
        _native.check(_native.lib().rc_fft_create(self._dev, plan.m, 1, C.byref(self._fft)))
        self._ctx = None
        self._fft_done = None

    def __del__(self):
        h, self._fft = getattr(self, "_fft", None), None
        if h:
            try:
                self._native.lib().rc_fft_destroy(h)
            except Exception:
                pass

    def empty(self, n):
        return torch.empty(int(n), dtype=torch.complex64, device="cuda")

    def begin(self, x_branch, slot, ready=None):
        cur = torch.cuda.current_stream()
        self._stream.wait_stream(cur)                     # the producer of x_branch, and the previous readers of `slot`
        if ready is not None:
            self._stream.wait_event(ready)
        x_branch.record_stream(self._stream)
        self._ctx = torch.cuda.stream(self._stream)
        self._ctx.__enter__()
        return None

    def fft(self, x, out):
        x = x.contiguous()
        self._native.check(self._native.lib().rc_fft_exec(self._fft, -1, x.data_ptr(), out.data_ptr(),
                                                          self._stream.cuda_stream))
        self._fft_done = torch.cuda.Event()
        self._fft_done.record(self._stream)

    def fft_scatter(self, x, piece_bases):
        C = self._C
        x = x.contiguous()
        arr = (C.c_void_p * len(piece_bases))(*piece_bases)
        self._native.check(self._native.lib().rc_fft_exec_scatter(
            self._fft, -1, x.data_ptr(), arr, len(piece_bases), self._plan.p, self._stream.cuda_stream))
        self._fft_done = torch.cuda.Event()
        self._fft_done.record(self._stream)

    def make_segments(self, segs):
        arr = (self._native.ScatterSeg * len(segs))()
        for a, (k1, j0, j1, dst) in zip(arr, segs):
            a.k1, a.reserved, a.j_lo, a.j_hi, a.dst = k1, 0, j0, j1, dst
        return arr

    def combine_scatter(self, pieces, k0_base, segs):
        p = self._plan
        self._native.check(self._native.lib().rc_subband_combine_scatter(
            self._dev, p.world, p.p, p.n, int(k0_base), pieces.data_ptr(), segs, len(segs), self._stream.cuda_stream))

    def combine(self, pieces, bins, k0_base):
        p = self._plan
        self._native.check(self._native.lib().rc_subband_combine(
            self._dev, p.world, p.p, p.n, int(k0_base), pieces.data_ptr(), bins.data_ptr(), self._stream.cuda_stream))

    def mark(self):
        ev = torch.cuda.Event(enable_timing=True)
        ev.record(self._stream)
        return ev

    def end(self, _):
        ev = torch.cuda.Event()
        ev.record(self._stream)
        self._ctx.__exit__(None, None, None)
        self._ctx = None
        return ev

    def wait(self, ev, slot, newest_fft_done=None):
        cur = torch.cuda.current_stream()
        cur.wait_event(ev)
        # Take turns on the SMs: the caller's channel kernels for this block start once the local FFT
        # of the block posted last has finished, so they run beside that block's NVLink-bound exchange
        # instead of sharing HBM bandwidth with its FFT passes.
        if newest_fft_done is not None and os.environ.get("RC_SHARD_TURNS", "1") != "0":
            cur.wait_event(newest_fft_done)


class _PeerBuffers:
    """Receive buffers of ShardedLoad in symmetric memory: one allocation per rank holding R ([G][P]
    pieces) and the sub-band slots, mapped into every rank of the group, plus the stream-ordered
    barrier of the mapping (signal pads in the same symmetric allocation)."""

    def __init__(self, m, slot_len, depth, world, rank, group, lanes=1):
        import torch.distributed._symmetric_memory as symm
        self._m, self._slot_len, self._depth, self._lanes = int(m), int(slot_len), int(depth), int(lanes)
        total = 2 * (self._lanes * self._m + self._depth * self._slot_len)   # float32 words (complex64 = 2)
        self._buf = symm.empty(total, dtype=torch.float32, device=torch.device("cuda", torch.cuda.current_device()))
        self._hdl = symm.rendezvous(self._buf, group if group is not None else dist.group.WORLD)
        self._views = [self._buf if r == rank else self._hdl.get_buffer(r, (total,), torch.float32) for r in range(world)]
        self.R = [self._complex(self._buf, ln * self._m, self._m) for ln in range(self._lanes)]
        self.slots = [self._complex(self._buf, self._slot_first(i), self._slot_len) for i in range(depth)]
        self.ptrs = list(self._hdl.buffer_ptrs)                       # peer base addresses (for kernels that store remotely)

    @staticmethod
    def _complex(buf, first, count):
        return torch.view_as_complex(buf[2 * first: 2 * (first + count)].view(count, 2))

    def _slot_first(self, i):
        return self._lanes * self._m + i * self._slot_len

    def r_offset_bytes(self, lane):
        return 8 * lane * self._m

    def slot_offset_bytes(self, i):
        return 8 * self._slot_first(i)

    def R_of(self, r, lane):
        return self._complex(self._views[r], lane * self._m, self._m)

    def slot_of(self, r, i):
        return self._complex(self._views[r], self._slot_first(i), self._slot_len)

    def barrier(self, channel):
        self._hdl.barrier(channel=channel)
