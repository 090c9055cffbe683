"""Multi-GPU data checks (need >= 2 GPUs; run with `gpurun --gpus 2 -- python -m pytest tests/test_gpu_multi.py -m gpu`).

One stream, channels sharded over the ranks.  `bcast` (NCCL broadcast of the block, every rank
repeats the N-point FFT) must give BIT-IDENTICAL audio to one GPU running every channel: same
kernels, same inputs.  `sharded` (commutated branches, local N/G-point FFTs, two NVLink exchanges,
radix-G combine) is a different factorisation of the same DFT: its sub-band is checked against a
float64 FFT and its audio against the oracle at the 1e-5 bar."""
import os
import socket

import numpy as np
import pytest

import radiocore_oracle as oracle
from bench_support import synth
from tests import multi_gpu_worker as w
from tests import parity

pytestmark = pytest.mark.gpu


def _free_port():
    with socket.socket() as s:
        s.bind(("127.0.0.1", 0))
        return s.getsockname()[1]


def _worlds():
    import torch
    n = torch.cuda.device_count() if torch.cuda.is_available() else 0
    return [g for g in (2, 4, 8) if g <= n]


@pytest.mark.parametrize("world", [2, 4, 8])
def test_one_stream_channel_shards(tmp_path, world):
    import torch.multiprocessing as mp
    if world not in _worlds():
        pytest.skip(f"needs {world} GPUs")
    mp.start_processes(w.run, args=(world, _free_port(), str(tmp_path)), nprocs=world, join=True, start_method="spawn")
    N, B, A, C_ = w.N, w.B, w.A, w.C_
    offs = synth.tiling_centers(N, C_, B)
    o = oracle.Tuner()
    for off in offs:
        o.add_channel(w.F0 + off, B, oracle.MFM(B, A))
    o.request_bandwidth(N)
    got = [np.load(os.path.join(str(tmp_path), f"rank{r}.npy"), allow_pickle=True).item() for r in range(world)]
    assert sorted(c for g in got for c in g["mine"]) == list(range(C_))
    worst = 0.0
    for b in range(w.BLOCKS):
        x = synth.wideband(N, offs, B, seed=31, block=b)
        o.load(x)
        ref = {c: o.channels()[c].demodulator.run(o.run(c)).reshape(-1) for c in range(C_)}
        if b == 0:
            X = np.fft.fft(x.astype(np.complex128))
        for g in got:
            assert np.array_equal(g[("bcast", b)], g[("full", b)]), f"bcast differs from one GPU, block {b}"
            want = np.concatenate([ref[c] for c in g["mine"]])
            worst = max(worst, parity.assert_parity(g[("sharded", b)], want, f"sharded b{b}"))
            # the three transports move the same numbers
            assert np.array_equal(g[("sharded_copy", b)], g[("sharded", b)]), f"copy transport differs, block {b}"
            assert np.array_equal(g[("sharded_nccl", b)], g[("sharded", b)]), f"NCCL transport differs, block {b}"
            parity.assert_parity(g[("full", b)], want, f"full b{b}")
            if b == 0:
                lo, length = g["arc"]
                sub = X[(lo + np.arange(length)) % N]
                # fp32 transform: rounding noise ~ eps * rms over all bins, plus eps * |X_k| on the strong
                # lines of the FM stations (twiddle rounding scales with the bin itself)
                err = np.abs(g["subband"] - sub)
                assert np.sqrt(np.mean(err ** 2)) <= 2e-6 * np.sqrt(np.mean(np.abs(X) ** 2))
                assert np.max(err) <= 2e-6 * np.max(np.abs(X)) + 5e-6 * np.sqrt(np.mean(np.abs(X) ** 2))
    print(f"world {world}: sharded-load audio worst rel err {worst:.2e}; modes",
          got[0]["sharded_mode"], got[0]["sharded_copy_mode"], got[0]["sharded_nccl_mode"])
