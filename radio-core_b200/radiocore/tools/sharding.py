"""Multi-GPU plumbing of the receive path: one process per GPU, channels sharded.

The reference is single-process (examples/multi_fm_server.py:113-140); scaling it across the
GPUs of one box is new here.  Channels are independent after ``Tuner.load``, so each rank keeps
a contiguous slice of the channel list and its carried de-emphasis state; the only exchange is
one broadcast of the wideband block (complex64, 8*N bytes) when all ranks listen to the same
stream.  There is no reduction and no gather of audio: every rank emits its own channels.

Works on any ``torch.distributed`` backend (NCCL on the GPUs, gloo in the CPU tests).
"""
import collections
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
