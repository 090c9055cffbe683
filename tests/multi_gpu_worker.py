"""Worker of tests/test_gpu_multi.py: one process per GPU (NCCL), importable for spawn.

Each rank runs the SAME four blocks three ways on its own GPU and writes what it got:
  full     every channel on one GPU (plain Tuner.load + run_all) -- the single-GPU answer
  bcast    one stream, NCCL broadcast of the block (sharding.BlockBroadcaster), this rank's slice
  sharded  commutated branches, sharded Tuner.load (sharding.ShardedLoad), this rank's slice
"""
import os
import sys

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
for p in (ROOT, os.path.join(ROOT, "radio-core_b200")):
    if p not in sys.path:
        sys.path.insert(0, p)

N, B, A, C_ = 1_600_000, 100_000, 20_000, 16
F0 = 100e6
BLOCKS = 4


def run(rank, world, port, out_dir, kind="MFM"):
    import numpy as np
    import torch
    import torch.distributed as dist
    import radiocore
    from bench_support import synth
    from radiocore.tools import sharding

    os.environ["MASTER_ADDR"] = "127.0.0.1"
    os.environ["MASTER_PORT"] = str(port)
    torch.cuda.set_device(rank)
    dist.init_process_group("nccl", rank=rank, world_size=world, device_id=torch.device("cuda", rank))
    try:
        offs = synth.tiling_centers(N, C_, B)
        centers = [F0 + o for o in offs]
        blocks = [torch.from_numpy(synth.wideband(N, offs, B, seed=31, block=b)).cuda() for b in range(BLOCKS)]
        make = lambda c: getattr(radiocore, kind)(B, A, cuda=True)                     # noqa: E731
        mine = list(sharding.channel_slice(C_, world, rank))

        def slice_of(tuner, packed, channels):
            return np.concatenate([packed[o: o + s * n] for (o, s, n) in (tuner.audio_slices()[c] for c in channels)])

        out = {"mine": mine}
        full = radiocore.Tuner(cuda=True)
        for c in range(C_):
            full.add_channel(centers[c], B, make(c))
        full.request_bandwidth(N)
        for b, x in enumerate(blocks):
            full.load(x)
            out[("full", b)] = slice_of(full, full.run_all(numpy_output=True), mine).copy()

        tb = radiocore.Tuner(cuda=True)
        sharding.shard_tuner(tb, centers, B, make, F0, N, world, rank)
        feed = sharding.BlockBroadcaster(N, "cuda", src=0)
        feed.post(blocks[0] if rank == 0 else None)
        for b in range(BLOCKS):
            if b + 1 < BLOCKS:
                feed.post(blocks[b + 1] if rank == 0 else None)                       # one block ahead
            tb.load(feed.take())
            out[("bcast", b)] = slice_of(tb, tb.run_all(numpy_output=True), range(len(mine))).copy()

        branches = [x[rank::world].contiguous() for x in blocks]
        # default: peer-mapped buffers, exchanges fused into the kernels' stores; then the same with
        # device-to-device copies (RC_SHARD_FUSED=0) and over NCCL (RC_SHARD_TRANSPORT=collective)
        for tag, env in (("sharded", {}), ("sharded_copy", {"RC_SHARD_FUSED": "0"}),
                         ("sharded_nccl", {"RC_SHARD_TRANSPORT": "collective"})):
            os.environ.update(env)
            ts = radiocore.Tuner(cuda=True)
            sharding.shard_tuner(ts, centers, B, make, F0, N, world, rank)
            load = sharding.ShardedLoad(ts)
            for k_ in env:
                del os.environ[k_]
            out[tag + "_mode"] = (load.transport, load.fused)
            out["arc"] = (load.x_lo, load.x_len)
            ahead = load.lanes                                      # blocks posted ahead: one per pipeline lane
            for b in range(min(ahead, BLOCKS)):
                load.post(branches[b])
            for b in range(BLOCKS):
                if b + ahead < BLOCKS:
                    load.post(branches[b + ahead])
                sub = load.take()
                ts.load_subband(sub)
                out[(tag, b)] = slice_of(ts, ts.run_all(numpy_output=True), range(len(mine))).copy()
                if b == 0 and tag == "sharded":
                    out["subband"] = sub[:load.x_len].cpu().numpy()
            torch.cuda.synchronize()
            dist.barrier()
            del load, ts
        torch.cuda.synchronize()
        np.save(os.path.join(out_dir, f"rank{rank}.npy"), out, allow_pickle=True)
        dist.barrier()
    finally:
        dist.destroy_process_group()
