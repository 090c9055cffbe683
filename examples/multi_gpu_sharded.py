#!/usr/bin/env python
"""One wideband stream on G GPUs: channels AND Tuner.load sharded (radiocore.tools.sharding).

    python -m torch.distributed.run --nproc-per-node 2 --master-addr 127.0.0.1 examples/multi_gpu_sharded.py [blocks]

Every rank registers its contiguous slice of the stations (band plan of the full list), receives
its commutator branch of each block -- samples rank, rank+G, rank+2G, ... as the ingest side would
deal them out of the RingBuffer -- and publishes its own stations' audio exactly as the
single-process loop of the reference's examples/multi_fm_server.py:98-106 does; there is no
gather of audio and no reduction.  The radio is a synthetic stand-in (every rank can compute its
own branch of the same stream); the "socket" counts the payload bytes.
"""
import os
import sys

import numpy as np
import torch
import torch.distributed as dist

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
for p in (ROOT, os.path.join(ROOT, "radio-core_b200")):
    if p not in sys.path:
        sys.path.insert(0, p)

from radiocore import MFM, Tuner              # noqa: E402
from radiocore.tools import sharding          # noqa: E402
from bench_support import synth               # noqa: E402

INPUT_RATE, BANDWIDTH, AUDIO_RATE, STATIONS = 8_000_000, 250_000, 48_000, 24
F0 = 100.0e6


def main(blocks=3):
    rank, world = int(os.environ.get("RANK", "0")), int(os.environ.get("WORLD_SIZE", "1"))
    torch.cuda.set_device(int(os.environ.get("LOCAL_RANK", "0")))
    if world > 1:
        os.environ.setdefault("MASTER_ADDR", "127.0.0.1")
        dist.init_process_group("nccl", device_id=torch.device("cuda", torch.cuda.current_device()))
    offsets = synth.tiling_centers(INPUT_RATE, STATIONS, BANDWIDTH)
    tuner = Tuner(cuda=True)
    mine = sharding.shard_tuner(tuner, [F0 + o for o in offsets], BANDWIDTH,
                                lambda c: MFM(BANDWIDTH, AUDIO_RATE, cuda=True), F0, INPUT_RATE, world, rank)
    sent = 0
    if world == 1:
        for blk in range(blocks):
            tuner.load(synth.wideband(INPUT_RATE, offsets, BANDWIDTH, seed=1, block=blk))
            for ch in tuner.channels():
                sent += len(ch.demodulator.run(tuner.run(ch.index)).tobytes())
    else:
        load = sharding.ShardedLoad(tuner)                       # collective: the ranks exchange the arcs they need
        branch = lambda blk: torch.from_numpy(                   # noqa: E731  the ingest side's strided copy
            synth.wideband(INPUT_RATE, offsets, BANDWIDTH, seed=1, block=blk)[rank::world].copy()).cuda()
        load.post(branch(0))
        for blk in range(blocks):
            if blk + 1 < blocks:
                load.post(branch(blk + 1))                       # block k+1 travels while block k is demodulated
            tuner.load_subband(load.take())
            for ch in tuner.channels():                          # the reference's per-channel loop, unchanged
                audio = ch.demodulator.run(tuner.run(ch.index))
                sent += len(ch.address_bytes) + len(audio.tobytes())
        dist.barrier()
    print(f"rank {rank}/{world}: stations {mine[0]}..{mine[-1]}, {blocks} blocks, {sent} payload bytes")
    if world > 1:
        dist.destroy_process_group()


if __name__ == "__main__":
    main(int(sys.argv[1]) if len(sys.argv) > 1 else 3)
