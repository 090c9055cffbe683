#!/usr/bin/env python
"""Headline benchmark: complex Msamples/s of wideband IQ through the
Tuner -> {FM | MFM | WBFM} chain (BASELINE.json `metric`).

    python bench.py [--gpus N] [--steps K] [--warmup W] [--workload cfg3] [--impl reference]

One "step" = one one-second block of synthetic wideband IQ through Tuner.load
and every channel's Tuner.run + demodulator.run.

  value     inputs already resident in HBM, CUDA-event timed, whole job
  e2e       the same through the public classes with HOST buffers: pinned
            host IQ -> H2D -> kernels -> D2H of every channel's audio, per step
  roofline  dominant kernel: its compulsory bytes / its event-timed duration
            against the measured HBM peak (MEASURED_PEAKS.json)
  cpu_baseline  the oracle port (reference algorithm, NumPy/SciPy) timed on the
            host cores on a bounded sample of the same block (N=1 only)

--impl reference times the reference's own CPU algorithm (oracle port: the
reference is pure Python whose arithmetic lives in SciPy) with a process pool
over channels.  N>1 (torchrun): every rank owns its own sub-band stream
(256 channels each, no data-path collective) -> weak scaling; `--mode bcast`
instead replicates one stream with an NCCL broadcast per block.
"""
import argparse
import ctypes as C
import json
import math
import os
import sys
import threading
import time

ROOT = os.path.dirname(os.path.abspath(__file__))
for _p in (ROOT, os.path.join(ROOT, "radio-core_b200")):
    if _p not in sys.path:
        sys.path.insert(0, _p)

import numpy as np  # noqa: E402

METRIC = "complex Msamples/s through multi-WBFM chain at 1/2/4/8 GPU; HBM GB/s vs peak"
UNIT = "Msamples/s"

WORKLOADS = {
    # name: (N, C, B, A, demod, description)
    "cfg3": (256_000_000, 256, 1_000_000, 48_000, "FM",
             "configs[2]: 256-channel channelizer + FM demod at 256 Msps complex, literal one-second block "
             "(N=256e6, C=256, B=1e6, A=48e3)"),
    "cfg3-wbfm": (256_000_000, 256, 1_000_000, 48_000, "WBFM",
                  "configs[2] geometry with WBFM stereo demodulators (N=256e6, C=256, B=1e6, A=48e3)"),
    "cfg2": (10_000_000, 32, 250_000, 48_000, "MFM",
             "configs[1]: Tuner 10 MHz -> 32 x 250 kHz + MFM (N=10e6, C=32, B=250e3, A=48e3)"),
    "cfg4": (16_000_000, 64, 250_000, 48_000, "WBFM",
             "configs[3]: 64-channel WBFM stereo with de-emphasis (N=16e6, C=64, B=250e3, A=48e3)"),
    "cfg5": (1_000_000_000, 2048, 250_000, 48_000, "FM",
             "configs[4]: 1 Gsps wideband block, 2048 x 250 kHz FM channels (256 per GPU at 8 GPUs), "
             "use with --gpus 8 --mode bcast (N=1e9, C=2048, B=250e3, A=48e3)"),
    "small": (4_000_000, 16, 250_000, 48_000, "MFM", "smoke-sized: N=4e6, C=16, B=250e3, MFM"),
}


def parse():
    ap = argparse.ArgumentParser()
    ap.add_argument("--gpus", type=int, default=1)
    ap.add_argument("--steps", type=int, default=10)
    ap.add_argument("--warmup", type=int, default=3)
    ap.add_argument("--impl", default="b200", choices=["b200", "reference"])
    ap.add_argument("--workload", default="cfg3", choices=sorted(WORKLOADS))
    ap.add_argument("--mode", default="independent", choices=["independent", "bcast"])
    ap.add_argument("--no-cpu-baseline", action="store_true")
    ap.add_argument("--no-e2e", action="store_true")
    return ap.parse_args()


# ------------------------------------------------------------------ synthetic
def tiling_offsets(N, Cn, B):
    return [-(Cn * B) / 2.0 + B / 2.0 + c * B for c in range(Cn)]


def make_wideband_gpu(N, Cn, B, seed, stereo, device):
    """Band-limited sum of FM stations synthesised per channel and placed in the
    wideband spectrum (polyphase synthesis with torch.fft -- setup, not the timed path).
    SURVEY.md 8(d): unit-amplitude stations, + complex AWGN sigma 0.05, scaled 1/sqrt(C)."""
    import torch
    from bench_support import synth
    g = torch.Generator(device=device)
    g.manual_seed(seed)
    X = torch.zeros(N, dtype=torch.complex64, device=device)
    t = torch.arange(B, dtype=torch.float64, device=device) / B
    offs = tiling_offsets(N, Cn, B)
    jj = torch.arange(-B // 2, B // 2, device=device)
    dev = 75e3 if B >= 200_000 else 0.3 * B
    for c, off in enumerate(offs):
        if stereo:
            fl, fr, fp = 1000.0 + 10.0 * (c % 16), 2500.0, 19000.0
            tp = 2 * math.pi

            def isin(f):
                return (1.0 - torch.cos(tp * f * t)) / (tp * f)

            def isinsin(f, gq):
                return 0.5 * (torch.sin(tp * (f - gq) * t) / (tp * (f - gq)) - torch.sin(tp * (f + gq) * t) / (tp * (f + gq)))

            ph = tp * dev * (0.45 * (0.8 * isin(fl) + 0.8 * isin(fr)) + 0.10 * isin(fp)
                             + 0.45 * (0.8 * isinsin(fl, 2 * fp) - 0.8 * isinsin(fr, 2 * fp)))
        else:
            fm = 300.0 + 50.0 * (c % 64)
            ph = (dev * 0.5 / fm) * (1.0 - torch.cos(2 * math.pi * fm * t))
        ph = ph + 0.61803398875 * c
        s = torch.polar(torch.ones_like(ph), ph).to(torch.complex64)
        S = torch.fft.fftshift(torch.fft.fft(s))
        idx = (int(off) + jj) % N
        X[idx] = S
        del s, S, ph
    x = torch.fft.ifft(X)
    del X
    x *= (N / B) / math.sqrt(Cn)
    nz = torch.randn(N, 2, generator=g, device=device, dtype=torch.float32)
    x += torch.view_as_complex(nz) * (0.05 / math.sqrt(Cn))
    del nz
    return x, offs


# --------------------------------------------------------------------- clocks
class ClockSampler(threading.Thread):
    REASONS = {0x2: "applications_clocks_setting", 0x4: "sw_power_cap", 0x8: "hw_slowdown",
               0x20: "sw_thermal_slowdown", 0x40: "hw_thermal_slowdown", 0x80: "hw_power_brake_slowdown"}

    def __init__(self, index):
        super().__init__(daemon=True)
        self.index, self.samples, self.mask, self.stop_flag, self.max_mhz = index, [], 0, False, None
        try:
            import pynvml
            pynvml.nvmlInit()
            self.nv = pynvml
            self.h = pynvml.nvmlDeviceGetHandleByIndex(index)
            self.max_mhz = pynvml.nvmlDeviceGetMaxClockInfo(self.h, pynvml.NVML_CLOCK_SM)
        except Exception:
            self.nv = None

    def run(self):
        if self.nv is None:
            return
        n = 0
        while not self.stop_flag:
            self._sample()
            n += 1
            time.sleep(0.002 if n < 50 else 0.05)      # short timed regions (config 2: a few ms) still get samples

    def _sample(self):
        try:
            self.samples.append(self.nv.nvmlDeviceGetClockInfo(self.h, self.nv.NVML_CLOCK_SM))
            self.mask |= int(self.nv.nvmlDeviceGetCurrentClocksThrottleReasons(self.h))
        except Exception:
            pass

    def summary(self):
        self.stop_flag = True
        if self.nv is not None and not self.samples:
            self._sample()                              # region shorter than the thread's start-up
        if self.nv is None or not self.samples:
            return {"sm_mhz": None, "sm_max_mhz": self.max_mhz, "reasons": ["unavailable"]}
        return {"sm_mhz": float(np.median(self.samples)), "sm_max_mhz": self.max_mhz,
                "reasons": [n for bit, n in self.REASONS.items() if self.mask & bit], "samples": len(self.samples)}


# ------------------------------------------------------------ CPU (reference)
def _oracle():
    p = os.path.join(ROOT, "oracle")
    if p not in sys.path:
        sys.path.insert(0, p)
    import radiocore_oracle
    return radiocore_oracle


_POOL_STATE = {}


def _pool_channel(i):
    """Literal reference arithmetic for channel i: roll + full-length Hann multiply +
    truncation + inverse FFT (tuner.py:151-161), then demodulator.run."""
    st = _POOL_STATE
    t0 = time.perf_counter()
    iq = st["tuner"].run(i)
    st["demods"][i].run(iq)
    return time.perf_counter() - t0


def cpu_reference_block(x_host, wl, n_sample, workers):
    """Time the reference algorithm on one block: full Tuner.load FFT + `n_sample`
    channels (literal O(N)-per-channel path), channels spread over `workers`
    processes.  Returns (seconds per full block extrapolated to all C channels,
    description, detail)."""
    import multiprocessing as mp
    oracle = _oracle()
    N, Cn, B, A, kind, _ = wl
    offs = tiling_offsets(N, Cn, B)
    tuner = oracle.Tuner(literal=True, fft_workers=workers)
    demods = []
    for off in offs:
        d = getattr(oracle, kind)(B, A)
        demods.append(d)
        tuner.add_channel(100e6 + off, B, d)
    tuner.request_bandwidth(N)
    t0 = time.perf_counter()
    tuner.load(x_host)
    t_load = time.perf_counter() - t0
    tuner._win = oracle.shifted_window("hann", N)      # cached by the reference after the first run
    sample = list(range(0, Cn, max(1, Cn // n_sample)))[:n_sample]
    _POOL_STATE.update(tuner=tuner, demods=demods)
    t0 = time.perf_counter()
    if workers > 1 and len(sample) > 1:
        with mp.get_context("fork").Pool(min(workers, len(sample))) as pool:
            pool.map(_pool_channel, sample)
    else:
        for i in sample:
            _pool_channel(i)
    t_ch = time.perf_counter() - t0
    _POOL_STATE.clear()
    per_block = t_load + t_ch * (Cn / len(sample))
    what = (f"one {N}-sample block: full Tuner.load FFT ({t_load:.2f} s, scipy.fft workers={workers}) + "
            f"{len(sample)} of {Cn} channels Tuner.run+{kind}.run literal ({t_ch:.2f} s on "
            f"{min(workers, len(sample))} processes), channel time extrapolated x{Cn / len(sample):.0f}")
    return per_block, what, {"t_load_s": t_load, "t_channels_s": t_ch, "channels_timed": len(sample)}


def host_workers(N, limit=None):
    cores = os.cpu_count() or 1
    try:
        import psutil
        avail = psutil.virtual_memory().available
    except Exception:
        avail = 64 << 30
    per_worker = 8 * N * 5            # roll (c64) + window product (c128) + temporaries
    w = max(1, min(cores, int((avail * 0.6 - 8 * N * 4) // max(per_worker, 1))))
    return min(w, limit) if limit else w


def run_reference(args, rank, world):
    if rank != 0:
        return
    wl = WORKLOADS[args.workload]
    N, Cn, B, A, kind, desc = wl
    import torch
    if torch.cuda.is_available():
        x, _ = make_wideband_gpu(N, Cn, B, 3, kind == "WBFM", "cuda")
        x_host = x.cpu().numpy()
        del x
        torch.cuda.empty_cache()
    else:
        rng = np.random.default_rng(3)
        x_host = (rng.standard_normal(N, dtype=np.float32) + 1j * rng.standard_normal(N, dtype=np.float32)).astype(np.complex64)
    workers = host_workers(N)
    n_sample = max(1, min(Cn, workers))
    budget = 150.0                            # seconds of wall clock for the whole warmup + steps loop
    times, what, wall, measured = [], "", 0.0, 0
    t_start = time.perf_counter()
    for step in range(args.warmup + args.steps):
        if times and (time.perf_counter() - t_start) + wall > budget:
            times.append(times[-1])          # bounded run: re-use the last measured sample
            continue
        t_step = time.perf_counter()
        per_block, what, _ = cpu_reference_block(x_host, wl, n_sample, workers)
        wall = time.perf_counter() - t_step  # what one more sample would cost (per_block is the extrapolated block time)
        times.append(per_block)
        measured += 1
    if measured < args.warmup + args.steps:
        what += f"; {measured} of {args.warmup + args.steps} steps measured within the {budget:.0f} s bound, the rest repeat the last sample"
    timed = times[args.warmup:] or times
    sec = float(np.mean(timed))
    value = N / sec / 1e6
    line = {"impl": "reference", "metric": METRIC, "value": value, "unit": UNIT, "n_gpus": args.gpus,
            "steps": args.steps, "warmup": args.warmup, "ms_per_step": sec * 1e3, "higher_is_better": True,
            "scaling": "weak", "vs_baseline": None, "dtype": "f64", "data": "synthetic",
            "config": {"workload": desc, "demodulator": kind},
            "cpu_baseline": {"value": value, "unit": UNIT, "cores": workers, "kind": "port", "sample": what},
            "e2e": {"value": value, "unit": UNIT, "h2d_bytes_per_step": 0, "d2h_bytes_per_step": 0},
            "gpu_launches": 0}
    emit(line)


# ------------------------------------------------------------------ B200 arm
def run_b200(args, rank, world, local_rank):
    import torch
    import torch.distributed as dist
    import radiocore
    from radiocore import _native

    torch.cuda.set_device(local_rank)
    device = torch.device("cuda", local_rank)
    lib = _native.lib()
    wl = WORKLOADS[args.workload]
    N, Cn, B, A, kind, desc = wl
    nch = 2 if kind == "WBFM" else 1
    bcast = args.mode == "bcast" and world > 1

    from radiocore.tools import sharding
    tuner = radiocore.Tuner(cuda=True)
    feed = None
    if bcast:
        # one wideband stream, replicated: rank 0 generates, one NCCL broadcast per block
        # (posted one block ahead, so it overlaps the kernels of the previous block),
        # every rank demodulates its contiguous slice of the channels
        offs = tiling_offsets(N, Cn, B)
        x_dev = make_wideband_gpu(N, Cn, B, 3, kind == "WBFM", device)[0] if rank == 0 else None
        feed = sharding.BlockBroadcaster(N, device, src=0)
        my = sharding.shard_tuner(tuner, [100e6 + o for o in offs], B, lambda c: getattr(radiocore, kind)(B, A, cuda=True),
                                  100e6, N, world, rank)
    else:
        my = list(range(Cn))
        x_dev, offs = make_wideband_gpu(N, Cn, B, 3 + rank, kind == "WBFM", device)
        for c in my:
            tuner.add_channel(100e6 + offs[c], B, getattr(radiocore, kind)(B, A, cuda=True))
        tuner.request_bandwidth(N)

    def barrier():
        if world > 1:
            dist.barrier()
        torch.cuda.synchronize()

    def step_device():
        if bcast:
            feed.post(x_dev)                       # block k+1 starts travelling ...
            tuner.load(feed.take())                # ... while block k is demodulated
        else:
            tuner.load(x_dev)
        tuner.run_all()

    if bcast:
        feed.post(x_dev)                           # prime: one block ahead
    for _ in range(max(args.warmup, 3)):
        step_device()
    barrier()

    # ---- timed region: device-resident input, per-kernel events on the same stream
    lib.rc_profile_reset()
    lib.rc_profile_enable(1)
    sampler = ClockSampler(local_rank)
    sampler.start()
    ev0, ev1 = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
    barrier()
    ev0.record()
    for _ in range(args.steps):
        step_device()
    ev1.record()
    barrier()
    clocks = sampler.summary()
    ms_total = ev0.elapsed_time(ev1)
    launches = int(lib.rc_profile_launches())
    need = lib.rc_profile_report(None, 0)
    buf = C.create_string_buffer(need + 16)
    lib.rc_profile_report(buf, need + 16)
    kernels = json.loads(buf.value.decode())
    lib.rc_profile_enable(0)
    lib.rc_profile_reset()

    # ---- e2e: host IQ (pinned) -> public API -> host audio of every channel, every step.
    # Tuner.submit()/collect() is the package's block pipeline: the H2D copy of block k+1 runs
    # while block k is in the kernels and block k-1's audio is read back (depth 2).
    e2e = None
    if not args.no_e2e and bcast:
        # one host stream: rank 0 copies each pinned block to the device on a side stream and
        # broadcasts it from there; every rank demodulates its channel slice and reads its audio
        # back.  Block k+1's copy and broadcast are queued before block k's kernels.
        copy_stream = torch.cuda.Stream()
        stage = [torch.empty(N, dtype=torch.complex64, device=device) for _ in range(3)] if rank == 0 else None
        x_host = [torch.empty(N, dtype=torch.complex64).pin_memory() for _ in range(2)] if rank == 0 else None
        if rank == 0:
            for xh in x_host:
                xh.copy_(x_dev)
        torch.cuda.synchronize()
        slices = tuner.audio_slices()
        d2h = 4 * sum(size * nchn for _, size, nchn in slices)
        state = {"i": 0}

        def send_next():
            i = state["i"]
            state["i"] += 1
            if rank != 0:
                feed.post(None)
                return
            with torch.cuda.stream(copy_stream):
                copy_stream.wait_stream(torch.cuda.current_stream())    # the slot's previous reader is queued there
                stage[i % 3].copy_(x_host[i % 2], non_blocking=True)
                feed.post(stage[i % 3])                                  # ordered behind the copy, not behind the kernels

        def run_e2e(steps):
            checksum = 0.0
            for _ in range(steps):
                send_next()
                tuner.load(feed.take())
                checksum += float(tuner.run_all(numpy_output=True)[0])
            return checksum

        while feed.in_flight():                     # the device-resident leg leaves one block posted
            feed.take()
        send_next()
        run_e2e(3)
        barrier()
        t0 = time.perf_counter()
        run_e2e(args.steps)
        torch.cuda.synchronize()
        t_e2e = time.perf_counter() - t0
        feed.take()
        e2e = {"t": t_e2e, "h2d": 8 * N, "d2h": d2h}
    elif not args.no_e2e:
        x_host = [torch.empty(N, dtype=torch.complex64).pin_memory() for _ in range(2)]
        for xh in x_host:
            xh.copy_(x_dev)
        torch.cuda.synchronize()
        slices = tuner.audio_slices()
        d2h = 4 * sum(size * nchn for _, size, nchn in slices)

        def run_e2e(steps):
            prev, checksum = None, 0.0
            for i in range(steps):
                t = tuner.submit(x_host[i % 2])
                if prev is not None:
                    audio = tuner.collect(prev)
                    checksum += float(audio[0])
                prev = t
            audio = tuner.collect(prev)
            return checksum + float(audio[0])

        run_e2e(3)
        barrier()
        t0 = time.perf_counter()
        run_e2e(args.steps)
        torch.cuda.synchronize()
        t_e2e = time.perf_counter() - t0
        e2e = {"t": t_e2e, "h2d": 8 * N, "d2h": d2h}

    # ---- reduce over ranks (max time)
    if bcast:
        while feed.in_flight():                     # nothing left travelling when the ranks part
            feed.take()
    t_dev = torch.tensor([ms_total, (e2e["t"] if e2e else 0.0)], dtype=torch.float64, device=device)
    if world > 1:
        dist.all_reduce(t_dev, op=dist.ReduceOp.MAX)
    ms_total, t_e2e = float(t_dev[0]), float(t_dev[1])
    streams = 1 if bcast else world
    samples_per_step = N * streams
    ms_per_step = ms_total / args.steps
    value = samples_per_step / (ms_per_step * 1e-3) / 1e6

    if rank != 0:
        return

    peaks = {}
    try:
        peaks = json.load(open(os.path.join(ROOT, "MEASURED_PEAKS.json")))
    except Exception:
        pass
    peak_gbs, peak_src = (peaks.get("hbm_gbs"), "measured (MEASURED_PEAKS.json hbm_gbs)") if peaks.get("hbm_gbs") \
        else (6650.0, "fallback (B200_PROFILING.md)")

    table = {}
    for tag, k in kernels.items():
        avg_ms = k["total_ms"] / max(k["count"], 1)
        gbs = k["bytes_per_launch"] / (avg_ms * 1e-3) / 1e9 if avg_ms > 0 else 0.0
        table[tag] = {"launches_per_step": k["count"] / args.steps, "avg_ms": round(avg_ms, 4),
                      "ms_per_step": round(k["total_ms"] / args.steps, 4),
                      "bytes_per_launch": k["bytes_per_launch"], "GBps": round(gbs, 1),
                      "frac_of_hbm_peak": round(gbs / peak_gbs, 4)}
    top = max(table, key=lambda t: table[t]["ms_per_step"]) if table else None
    traffic = None
    try:
        traffic = json.load(open(os.path.join(ROOT, "profiles", "dram_traffic.json"))).get(args.workload, {}).get(top)
    except Exception:
        pass
    roofline = None
    if top:
        kt = table[top]
        roofline = {"kernel": top, "bound": "hbm", "achieved": kt["GBps"], "peak": peak_gbs, "unit": "GB/s",
                    "frac": round(kt["GBps"] / peak_gbs, 4), "traffic": traffic, "peak_source": peak_src,
                    "avg_launch_ms": kt["avg_ms"], "share_of_step": round(kt["ms_per_step"] / ms_per_step, 4)}
    algo_bytes = 8 * N + 4 * A * nch * len(my)
    path_gbs = algo_bytes / (ms_per_step * 1e-3) / 1e9
    line = {"metric": METRIC, "value": value, "unit": UNIT, "n_gpus": world, "steps": args.steps,
            "warmup": max(args.warmup, 3), "ms_per_step": ms_per_step, "higher_is_better": True,
            "scaling": "strong" if bcast else "weak", "vs_baseline": None, "dtype": "f32", "data": "synthetic",
            "config": {"workload": desc, "demodulator": kind, "channels_per_gpu": len(my),
                       "multi_gpu": ("one stream, NCCL broadcast per block, channel slices" if bcast else
                                     "independent sub-band stream per GPU, no data-path collective") if world > 1 else "n/a",
                       "l2": "input block (%.0f MB) larger than the 126 MB L2, no flush needed" % (8 * N / 1e6)
                             if 8 * N > 130e6 else "input smaller than L2: steady-state L2-resident run"},
            "gpu_launches": launches,
            "roofline": roofline,
            "roofline_path": {"algorithmic_bytes_per_step": algo_bytes, "achieved": round(path_gbs, 1), "peak": peak_gbs,
                              "unit": "GB/s", "frac": round(path_gbs / peak_gbs, 4),
                              "note": "SURVEY 8(d) bytes (read IQ once + write audio) over the whole step"},
            "kernels": table, "clocks": clocks}
    if e2e:
        api = ("rank 0: pinned host IQ -> H2D -> BlockBroadcaster (NCCL, one block ahead); every rank: Tuner.load + run_all "
               "-> D2H of its channels' audio") if bcast else \
            "Tuner.submit()/collect(): pinned host IQ -> H2D -> kernels -> D2H of all channels' audio, 2 blocks in flight"
        line["e2e"] = {"value": samples_per_step * args.steps / t_e2e / 1e6, "unit": UNIT,
                       "h2d_bytes_per_step": e2e["h2d"] * streams, "d2h_bytes_per_step": e2e["d2h"] * world,
                       "ms_per_step": t_e2e / args.steps * 1e3, "timing": "wall clock between device synchronisations, max over ranks",
                       "api": api}
    if world == 1 and not args.no_cpu_baseline:
        try:
            workers = host_workers(N, limit=None)
            x_host_np = x_dev.cpu().numpy()
            sec, what, _ = cpu_reference_block(x_host_np, wl, max(1, min(Cn, 2 if workers < 4 else workers)), workers)
            line["cpu_baseline"] = {"value": N / sec / 1e6, "unit": UNIT, "cores": workers, "kind": "port", "sample": what}
        except Exception as exc:       # never lose the GPU numbers to a host-side problem
            line["cpu_baseline"] = {"value": None, "unit": UNIT, "cores": os.cpu_count(), "kind": "port",
                                    "sample": f"failed: {exc!r}"}
    emit(line)


_JSON_FD = None


def emit(line: dict):
    """The ONE JSON line of the contract, on the real stdout (see main: everything else a library
    prints on fd 1 -- e.g. NCCL's version banner -- is diverted to stderr)."""
    data = (json.dumps(line) + "\n").encode()
    if _JSON_FD is None:
        sys.stdout.write(data.decode())
        sys.stdout.flush()
    else:
        os.write(_JSON_FD, data)


def main():
    global _JSON_FD
    args = parse()
    sys.stdout.flush()
    _JSON_FD = os.dup(1)
    os.dup2(2, 1)
    rank = int(os.environ.get("RANK", "0"))
    world = int(os.environ.get("WORLD_SIZE", "1"))
    local_rank = int(os.environ.get("LOCAL_RANK", "0"))
    if args.impl == "reference":
        run_reference(args, rank, world)
        return
    import torch
    if not torch.cuda.is_available():
        raise SystemExit("bench.py: no CUDA device (the B200 arm has no CPU fallback); use --impl reference for the CPU arm")
    if world > 1:
        import torch.distributed as dist
        os.environ.setdefault("MASTER_ADDR", "127.0.0.1")
        os.environ.setdefault("NCCL_DEBUG", "WARN")      # keep NCCL's version banner off stdout: one JSON line only
        torch.cuda.set_device(local_rank)
        dist.init_process_group("nccl", device_id=torch.device("cuda", local_rank))
    try:
        run_b200(args, rank, world, local_rank)
    finally:
        if world > 1:
            import torch.distributed as dist
            dist.destroy_process_group()


if __name__ == "__main__":
    main()
