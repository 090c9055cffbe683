"""CPU arm of the benchmark: the reference's own NumPy/SciPy path on the host cores.

BASELINE INFRASTRUCTURE (only bench.py's `--impl reference` / `cpu_baseline` legs import this).
What is timed is the UNMODIFIED reference package (`oracle/_ref`, vendored by `oracle/make_ref.py`;
`kind: "reference"`) called exactly as examples/multi_fm_server.py:98-106 does --
`tuner.load(block)`, then per channel `tuner.run(i)` + `channel.demodulator.run(...)` -- or, when
the copy is absent, the oracle port of the same algorithm (`kind: "port"`).

Two figures:
  * single core, as written: `scipy.fft` with its default `workers=1`, channels one after another;
  * all cores: the load FFT as written (single thread: the reference passes no `workers`), the
    channel loop spread over a pool of forked processes that share the loaded spectrum.
A one-second block at the large configurations costs minutes of CPU, so a step times a bounded
sample -- the load FFT (warm-up steps only) and k of the C channels (every step) -- and the block
time is `t_load + t_channels * C / k`; the line says so.  The pool is created and warmed outside
the timed region.
"""
import os
import sys
import time

import numpy as np

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))

_STATE = {}


def load_impl():
    """(namespace with Tuner / FM / MFM / WBFM, kind)."""
    p = os.path.join(ROOT, "oracle")
    if p not in sys.path:
        sys.path.insert(0, p)
    try:
        import ref_shim
        if ref_shim.available():
            return ref_shim.load_reference(), "reference"
    except Exception:
        pass
    import radiocore_oracle
    return radiocore_oracle, "port"


def host_workers(N, limit=None):
    cores = os.cpu_count() or 1
    try:
        import psutil
        avail = psutil.virtual_memory().available
    except Exception:
        avail = 64 << 30
    per_worker = 8 * N * 5            # roll (c64) + window product (c128) + temporaries
    w = max(1, min(cores, int((avail * 0.6 - 8 * N * 6) // max(per_worker, 1))))
    return min(w, limit) if limit else w


def _one_channel(i):
    st = _STATE
    t0 = time.perf_counter()
    iq = st["tuner"].run(i)
    st["channels"][i].demodulator.run(iq)
    return time.perf_counter() - t0


class CpuArm:
    def __init__(self, N, Cn, B, A, kind, offsets, x_host, f0=100e6):
        self.impl, self.impl_kind = load_impl()
        self.N, self.Cn, self.B, self.A, self.kind = N, Cn, B, A, kind
        self.x = x_host
        ctor = getattr(self.impl, kind)
        self.tuner = self.impl.Tuner()
        for off in offsets:
            self.tuner.add_channel(f0 + off, B, ctor(B, A))
        self.tuner.request_bandwidth(N)
        self.pool = None
        self.workers = 1
        self.t_load = []
        self.t_single = None

    # ---- single core, as written
    def load(self):
        t0 = time.perf_counter()
        self.tuner.load(self.x)
        dt = time.perf_counter() - t0
        self.t_load.append(dt)
        return dt

    def single_channel(self, index=0):
        """One channel in this process (also builds the reference's cached Hann window before the
        workers are forked, tuner.py:155-157)."""
        _STATE.update(tuner=self.tuner, channels=self.tuner.channels())
        first = _one_channel(index)           # includes get_window(N) on the first call
        self.t_single = _one_channel(index)
        return first, self.t_single

    # ---- all cores: forked workers sharing the loaded spectrum
    def start_pool(self, workers):
        import multiprocessing as mp
        self.workers = max(1, int(workers))
        if self.workers > 1:
            _STATE.update(tuner=self.tuner, channels=self.tuner.channels())
            self.pool = mp.get_context("fork").Pool(self.workers)
            self.pool.map(_one_channel, list(range(self.workers)), chunksize=1)     # warm: page in, first-touch

    def sample_channels(self, k):
        k = max(1, min(self.Cn, k))
        step = max(1, self.Cn // k)
        return list(range(0, self.Cn, step))[:k]

    def run_channels(self, sample):
        """Wall seconds for `sample` over the pool, and the per-channel seconds seen by the workers."""
        t0 = time.perf_counter()
        if self.pool is not None:
            per = self.pool.map(_one_channel, sample, chunksize=1)
        else:
            per = [_one_channel(i) for i in sample]
        return time.perf_counter() - t0, per

    def close(self):
        if self.pool is not None:
            self.pool.close()
            self.pool.join()
            self.pool = None
        _STATE.clear()

    # ---- figures
    def block_seconds(self, t_channels, k):
        return float(np.mean(self.t_load)) + t_channels * (self.Cn / k)

    def single_core_value(self):
        return self.N / (float(np.mean(self.t_load)) + self.t_single * self.Cn) / 1e6


def measure(N, Cn, B, A, kind, offsets, x_host, steps, warmup, step_seconds=6.0, limit_workers=None):
    """Run `warmup` + `steps` bounded steps; returns the dict bench.py turns into its JSON line."""
    arm = CpuArm(N, Cn, B, A, kind, offsets, x_host)
    try:
        arm.load()
        first, t1 = arm.single_channel(0)
        workers = host_workers(N, limit_workers)
        arm.start_pool(workers)
        # channels per step: what the pool gets through in about `step_seconds`
        k = int(step_seconds / max(t1, 1e-6)) * workers
        k = max(workers, min(Cn, k // workers * workers if k >= workers else workers))
        if k >= Cn:
            k = Cn
        sample = arm.sample_channels(k)
        walls = []
        for s in range(warmup + steps):
            if s < warmup and s < 2 and 2 * np.mean(arm.t_load) < step_seconds * 4:
                arm.load()                      # a second (and third) sample of the FFT, outside the timed steps
            wall, _ = arm.run_channels(sample)
            walls.append(wall)
        timed = walls[warmup:] or walls
        t_ch = float(np.mean(timed))
        block = arm.block_seconds(t_ch, len(sample))
        what = (f"{arm.impl_kind} package on {workers} host processes: Tuner.load of the {N}-sample block timed "
                f"{len(arm.t_load)}x (mean {np.mean(arm.t_load):.2f} s, scipy.fft workers=1 as written); every step = "
                f"{len(sample)} of {Cn} channels Tuner.run+{kind}.run (mean {t_ch:.2f} s wall, {len(timed)} steps measured); "
                f"block time = t_load + t_channels*{Cn}/{len(sample)}")
        return {"value": N / block / 1e6, "block_s": block, "step_wall_s": t_ch, "steps_measured": len(timed),
                "cores": workers, "kind": arm.impl_kind, "sample": what,
                "single_core": {"value": arm.single_core_value(), "cores": 1,
                                "sample": f"t_load {np.mean(arm.t_load):.2f} s + {Cn} x one channel {t1:.3f} s, one process, as written"},
                "t_load_s": float(np.mean(arm.t_load)), "channels_per_step": len(sample)}
    finally:
        arm.close()
