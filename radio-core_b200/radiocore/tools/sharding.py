"""Multi-GPU plumbing of the receive path: one process per GPU, channels sharded.

The reference is single-process (examples/multi_fm_server.py:113-140); scaling it across the
GPUs of one box is new here.  Channels are independent after ``Tuner.load``, so each rank keeps
a contiguous slice of the channel list and its carried de-emphasis state; the only exchange is
one broadcast of the wideband block (complex64, 8*N bytes) when all ranks listen to the same
stream.  There is no reduction and no gather of audio: every rank emits its own channels.

Works on any ``torch.distributed`` backend (NCCL on the GPUs, gloo in the CPU tests).
"""
from typing import List, Sequence

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
