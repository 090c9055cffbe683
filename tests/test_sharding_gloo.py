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


# ------------------------------------------------------------------ sharded Tuner.load
class _NumpyKernels:
    """Stand-in for the CUDA kernels of sharding.ShardedLoad (complex128 on the CPU): what is under
    test is the plan and the two exchanges, not the arithmetic."""

    def __init__(self, n, world):
        self.n, self.world = n, world

    def empty(self, count):
        return torch.zeros(int(count), dtype=torch.complex128)

    def begin(self, x, slot, ready=None):
        return None

    def end(self, token):
        return None

    def wait(self, ev, slot, newest_fft_done=None):
        pass

    def fft(self, x, out):
        out.copy_(torch.fft.fft(x.to(torch.complex128)))

    def combine(self, pieces, bins, k0_base):
        g_, p = self.world, pieces.numel() // self.world
        f = pieces.view(g_, p).numpy()
        k0 = k0_base + np.arange(p)
        g = np.arange(g_)[:, None]
        v = f * np.exp(-2j * np.pi * g * k0[None, :] / self.n)
        dft = np.exp(-2j * np.pi * np.outer(np.arange(g_), np.arange(g_)) / g_)
        bins.view(g_, p).copy_(torch.from_numpy(dft @ v))


class _OracleSubbandTuner(oracle.Tuner):
    """The oracle's Tuner with the three hooks ShardedLoad uses on the product's Tuner."""

    def needed_bins(self):
        n = int(self.input_bandwidth)
        return [((-(int(c.bandwidth) // 2) - self.roll_of(c.index)) % n, int(c.bandwidth) + 1) for c in self._bounds]

    def set_subband(self, x_lo, x_len):
        self.subband = (x_lo, x_len)

    def load_subband(self, spectrum):
        n = int(self.input_bandwidth)
        x_lo, x_len = self.subband
        full = np.zeros(n, dtype=np.complex128)
        full[(x_lo + np.arange(x_len)) % n] = spectrum.numpy()[:x_len]
        self._buffer = full


def _sharded_worker(rank, world, port, out_dir):
    os.environ["MASTER_ADDR"] = "127.0.0.1"
    os.environ["MASTER_PORT"] = str(port)
    dist.init_process_group("gloo", rank=rank, world_size=world)
    try:
        offs = synth.tiling_centers(N, C_, B)
        centers = [F0 + o for o in offs]
        tuner = _OracleSubbandTuner()
        mine = sharding.shard_tuner(tuner, centers, B, lambda c: oracle.MFM(B, A), F0, N, world, rank)
        load = sharding.ShardedLoad(tuner, kernels=_NumpyKernels(N, world))
        assert load.plan.m == N // world and load.x_len <= N
        blocks = [synth.wideband(N, offs, B, seed=5, block=b) for b in range(3)]       # every rank can make its branch
        out = {"arc": (load.x_lo, load.x_len), "mine": mine}
        load.post(torch.from_numpy(blocks[0][rank::world].copy()))
        for blk in range(3):
            if blk + 1 < 3:
                load.post(torch.from_numpy(blocks[blk + 1][rank::world].copy()))     # one block ahead
            sub = load.take()
            out[("X", blk)] = sub.numpy()[:load.x_len].copy()
            tuner.load_subband(sub)
            for ch in tuner.channels():
                out[(blk, mine[ch.index])] = ch.demodulator.run(tuner.run(ch.index))
        assert load.in_flight() == 0
        np.save(os.path.join(out_dir, f"rank{rank}.npy"), out, allow_pickle=True)
        dist.barrier()
    finally:
        dist.destroy_process_group()


@pytest.mark.parametrize("world", [2, 4])
def test_sharded_load_equals_one_fft(tmp_path, world):
    """Commutated branches -> local FFTs -> two exchanges: every rank ends up with exactly the bins
    of fft(block) its channels gather from, and the audio of the plain single-process chain."""
    mp.start_processes(_sharded_worker, args=(world, _free_port(), str(tmp_path)), nprocs=world, join=True,
                       start_method="fork")
    offs = synth.tiling_centers(N, C_, B)
    ref = oracle.Tuner()
    for o in offs:
        ref.add_channel(F0 + o, B, oracle.MFM(B, A))
    ref.request_bandwidth(N)
    got = [np.load(os.path.join(str(tmp_path), f"rank{r}.npy"), allow_pickle=True).item() for r in range(world)]
    assert sorted(c for g in got for c in g["mine"]) == list(range(C_))
    for blk in range(3):
        x = synth.wideband(N, offs, B, seed=5, block=blk)
        X = np.fft.fft(x.astype(np.complex128))
        ref.load(x)
        for g in got:
            lo, length = g["arc"]
            assert lo % 2 == 0 and length < N
            want = X[(lo + np.arange(length)) % N]
            assert np.max(np.abs(g[("X", blk)] - want)) <= 1e-9 * np.max(np.abs(X))
            for c in g["mine"]:
                a = ref.channels()[c].demodulator.run(ref.run(c))
                assert np.max(np.abs(g[(blk, c)] - a)) <= 2e-6 * np.max(np.abs(a)), (blk, c)


def test_subband_plan_covers_every_arc():
    """Every bin of every rank's arc is sent exactly once, by the rank that combines it."""
    rng = np.random.default_rng(0)
    for world, n in ((2, 64), (4, 1600), (8, 64 * 50)):
        m = n // world
        arcs = []
        for r in range(world):
            lo = int(rng.integers(0, n // 2)) * 2
            arcs.append((lo, int(rng.integers(m // 2, min(n, m + m // 3)))))
        plan = sharding.SubbandPlan(n, world, arcs)
        for d, (lo, length) in enumerate(arcs):
            seen = np.full(length, -1)
            for src in range(world):
                for k1, j0, j1, pos in plan.runs(src, d):
                    bins = k1 * plan.m + src * plan.p + np.arange(j0, j1)
                    assert np.all(seen[pos: pos + (j1 - j0)] == -1)
                    seen[pos: pos + (j1 - j0)] = bins
            assert np.array_equal(seen, (lo + np.arange(length)) % n)
    assert sharding.covering_arc([(990, 20), (10, 5)], 1000) == (990, 25)
    with pytest.raises(ValueError):
        sharding.SubbandPlan(100, 8, [(0, 10)] * 8)
