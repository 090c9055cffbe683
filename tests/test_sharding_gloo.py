"""World-size-2 check (gloo, CPU) of the multi-GPU host logic: channel slices, the band plan of
a sharded Tuner and the single broadcast of the wideband block.  The arithmetic on each rank is
the oracle's (the GPU kernels are covered by the -m gpu tests); what is tested here is that two
ranks together produce exactly the audio one process produces for all channels."""
import os
import socket

import numpy as np
import pytest
import torch
import torch.distributed as dist
import torch.multiprocessing as mp

import radiocore_oracle as oracle
from bench_support import synth
from radiocore.tools import sharding

N, B, A, C_ = 80_000, 10_000, 2_000, 6
F0 = 100e6


def _free_port():
    with socket.socket() as s:
        s.bind(("127.0.0.1", 0))
        return s.getsockname()[1]


def _worker(rank, world, port, out_dir):
    os.environ["MASTER_ADDR"] = "127.0.0.1"
    os.environ["MASTER_PORT"] = str(port)
    dist.init_process_group("gloo", rank=rank, world_size=world)
    try:
        offs = synth.tiling_centers(N, C_, B)
        centers = [F0 + o for o in offs]
        tuner = oracle.Tuner()
        mine = sharding.shard_tuner(tuner, centers, B, lambda c: oracle.MFM(B, A), F0, N, world, rank)
        assert mine == list(sharding.channel_slice(C_, world, rank))
        audio = {}
        # block 0 through the plain broadcast, block 1 through the double-buffered feed
        # (posted one block ahead, as bench.py --mode bcast does on the GPUs)
        block = torch.zeros(N, dtype=torch.complex64)
        if rank == 0:
            block = torch.from_numpy(synth.wideband(N, offs, B, seed=5, block=0))
        sharding.broadcast_block(block, src=0)
        feed = sharding.BlockBroadcaster(N, "cpu", src=0)
        nxt = torch.from_numpy(synth.wideband(N, offs, B, seed=5, block=1)) if rank == 0 else None
        feed.post(nxt)
        for blk in range(2):
            if blk == 1:
                block = feed.take()
                assert feed.in_flight() == 0
            tuner.load(block.numpy())
            for ch in tuner.channels():
                audio[(blk, mine[ch.index])] = ch.demodulator.run(tuner.run(ch.index))
        with pytest.raises(RuntimeError):
            feed.take()
        np.save(os.path.join(out_dir, f"rank{rank}.npy"), audio, allow_pickle=True)
        # no reduction, no gather on the data path: only a barrier to end together
        dist.barrier()
    finally:
        dist.destroy_process_group()


def test_channel_slices_partition():
    for n, w in ((256, 8), (2048, 8), (7, 2), (5, 8), (32, 3)):
        seen = []
        for r in range(w):
            sl = sharding.channel_slice(n, w, r)
            seen.extend(sl)
            for c in sl:
                assert sharding.owner_of(c, n, w) == r
        assert seen == list(range(n))
    with pytest.raises(ValueError):
        sharding.channel_slice(4, 2, 2)


def test_two_ranks_equal_one_process(tmp_path):
    world = 2
    mp.start_processes(_worker, args=(world, _free_port(), str(tmp_path)), nprocs=world, join=True,
                       start_method="fork")
    offs = synth.tiling_centers(N, C_, B)
    ref = oracle.Tuner()
    for o in offs:
        ref.add_channel(F0 + o, B, oracle.MFM(B, A))
    ref.request_bandwidth(N)
    got = {}
    for r in range(world):
        got.update(np.load(os.path.join(str(tmp_path), f"rank{r}.npy"), allow_pickle=True).item())
    assert sorted(got) == [(b, c) for b in range(2) for c in range(C_)]
    for blk in range(2):
        ref.load(synth.wideband(N, offs, B, seed=5, block=blk))
        for ch in ref.channels():
            want = ch.demodulator.run(ref.run(ch.index))
            assert np.array_equal(got[(blk, ch.index)], want), (blk, ch.index)
