#!/usr/bin/env python
"""Headline benchmark: complex Msamples/s of wideband IQ through the
Tuner -> {FM | MFM | WBFM} chain (BASELINE.json `metric`).

    python bench.py [--gpus N] [--steps K] [--warmup W] [--workload cfg3] [--impl reference]

One "step" = one block of synthetic wideband IQ through Tuner.load and every channel's
Tuner.run + demodulator.run.

  value     inputs already resident in HBM, CUDA-event timed, whole job
  e2e       the same through the public classes with HOST buffers: pinned host IQ -> H2D ->
            kernels -> D2H of every channel's audio, every step
  roofline  dominant kernel: its compulsory bytes / its event-timed duration against the
            measured HBM peak (MEASURED_PEAKS.json); roofline_path: SURVEY 8(d) bytes of the
            whole step over the step time
  cpu_baseline  the reference package itself (oracle/_ref; `kind: "reference"`) timed on the
            host cores on a bounded sample of the same block (N=1 only)

N = 1: configs[2] (256 x 1 MHz FM, 256 Msps, literal one-second block) is the line; the north
star's WBFM chain on the same geometry (`wbfm_chain`), the short-block variant (`short_block`,
SURVEY 7.3-3) and the literal drop-in loop (`e2e_dropin`) ride along as extra keys.
N > 1 (torchrun): ONE wideband stream; every rank owns a contiguous slice of the channels and the
commutator branch x[rank::N] of the block; Tuner.load itself is sharded (local N/G-point FFTs, two
NVLink exchanges, radix-G combine: radiocore/tools/sharding.py) -> "strong" scaling.  Extra keys:
`bcast` (NCCL broadcast of the block, every rank repeats the FFT), `replicas` (independent
sub-band stream per GPU, no collective) and, at 8 GPUs, `cfg5` (configs[4]: N = 1e9, 2048 channels).

--impl reference times the reference's own CPU implementation (bench_support/cpu_arm.py).
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
NVLINK_PEAK_GBS = 770.0          # measured peer copy, per direction per GPU (B200_PROFILING.md)

WORKLOADS = {
    # name: (N, C, B, A, demod, description)
    "cfg3": (256_000_000, 256, 1_000_000, 48_000, "FM",
             "configs[2]: 256-channel channelizer + FM demod at 256 Msps complex, literal one-second block "
             "(N=256e6, C=256, B=1e6, A=48e3)"),
    "cfg3-wbfm": (256_000_000, 256, 1_000_000, 48_000, "WBFM",
                  "configs[2] geometry with WBFM stereo demodulators (N=256e6, C=256, B=1e6, A=48e3)"),
    "cfg3-short": (8_000_000, 256, 31_250, 1_500, "FM",
                   "configs[2] in short blocks of T = 1/32 s (SURVEY 7.3-3): N=8e6, C=256, B*T=31250, A*T=1500, "
                   "32 blocks per second of signal; reference instantiated in bin units"),
    "cfg2": (10_000_000, 32, 250_000, 48_000, "MFM",
             "configs[1]: Tuner 10 MHz -> 32 x 250 kHz + MFM (N=10e6, C=32, B=250e3, A=48e3)"),
    "cfg4": (16_000_000, 64, 250_000, 48_000, "WBFM",
             "configs[3]: 64-channel WBFM stereo with de-emphasis (N=16e6, C=64, B=250e3, A=48e3)"),
    "cfg5": (1_000_000_000, 2048, 250_000, 48_000, "FM",
             "configs[4]: 1 Gsps wideband block, 2048 x 250 kHz FM channels (256 per GPU at 8 GPUs) "
             "(N=1e9, C=2048, B=250e3, A=48e3)"),
    "small": (4_000_000, 16, 250_000, 48_000, "MFM", "smoke-sized: N=4e6, C=16, B=250e3, MFM"),
}


def parse():
    ap = argparse.ArgumentParser()
    ap.add_argument("--gpus", type=int, default=1)
    ap.add_argument("--steps", type=int, default=10)
    ap.add_argument("--warmup", type=int, default=3)
    ap.add_argument("--impl", default="b200", choices=["b200", "reference"])
    ap.add_argument("--workload", default="cfg3", choices=sorted(WORKLOADS))
    ap.add_argument("--mode", default="auto", choices=["auto", "sharded", "bcast", "independent"],
                    help="N>1: auto = sharded (one stream, sharded Tuner.load)")
    ap.add_argument("--no-cpu-baseline", action="store_true")
    ap.add_argument("--no-e2e", action="store_true")
    ap.add_argument("--no-extras", action="store_true")
    return ap.parse_args()


# ------------------------------------------------------------------ synthetic
def tiling_offsets(N, Cn, B):
    return [-(Cn * B) / 2.0 + B / 2.0 + c * B for c in range(Cn)]


def make_wideband_gpu(N, Cn, B, seed, stereo, device):
    """Band-limited sum of FM stations synthesised per channel and placed in the
    wideband spectrum (polyphase synthesis with torch.fft -- setup, not the timed path).
    SURVEY.md 8(d): unit-amplitude stations, + complex AWGN sigma 0.05, scaled 1/sqrt(C)."""
    import torch
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
            fm = (300.0 + 50.0 * (c % 64)) if B >= 200_000 else (3.0 + (c % 16))     # short blocks: a few cycles per block
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
def config_of(wl):
    """The `config` object: identical in the B200 arm and the reference arm."""
    return {"workload": wl[5], "demodulator": wl[4]}


def host_block(wl, seed=3):
    """The block of the workload on the host (synthesised on the GPU when there is one)."""
    N, Cn, B, A, kind, _ = wl
    import torch
    if torch.cuda.is_available():
        x, _ = make_wideband_gpu(N, Cn, B, seed, kind == "WBFM", "cuda")
        x_host = x.cpu().numpy()
        del x
        torch.cuda.empty_cache()
        return x_host
    rng = np.random.default_rng(seed)
    return (rng.standard_normal(N, dtype=np.float32) + 1j * rng.standard_normal(N, dtype=np.float32)).astype(np.complex64)


def run_reference(args, rank, world):
    if rank != 0:
        return
    from bench_support import cpu_arm
    wl = WORKLOADS[args.workload]
    N, Cn, B, A, kind, desc = wl
    x_host = host_block(wl)
    r = cpu_arm.measure(N, Cn, B, A, kind, tiling_offsets(N, Cn, B), x_host, steps=args.steps, warmup=args.warmup,
                        step_seconds=6.0)
    line = {"impl": "reference", "metric": METRIC, "value": r["value"], "unit": UNIT, "n_gpus": args.gpus,
            "steps": args.steps, "warmup": args.warmup, "ms_per_step": r["step_wall_s"] * 1e3, "higher_is_better": True,
            "scaling": "strong" if args.gpus > 1 else "weak", "vs_baseline": None, "dtype": "f64", "data": "synthetic",
            "config": config_of(wl),
            "step": f"bounded sample: {r['channels_per_step']} of {Cn} channels per step (wall ms_per_step); value = N / "
                    f"(t_load + t_step * {Cn}/{r['channels_per_step']}) = one block in {r['block_s']:.2f} s",
            "cpu_baseline": {"value": r["value"], "unit": UNIT, "cores": r["cores"], "kind": r["kind"], "sample": r["sample"],
                             "single_core": r["single_core"]},
            "e2e": {"value": r["value"], "unit": UNIT, "h2d_bytes_per_step": 0, "d2h_bytes_per_step": 0},
            "gpu_launches": 0}
    emit(line)


# ------------------------------------------------------------------ B200 arm
class Ctx:
    """What the measurement helpers share."""

    def __init__(self, args, rank, world, local_rank):
        import torch
        import torch.distributed as dist
        import radiocore
        from radiocore import _native
        from radiocore.tools import sharding
        self.torch, self.dist, self.rc, self.sharding = torch, dist, radiocore, sharding
        self.args, self.rank, self.world, self.local_rank = args, rank, world, local_rank
        self.device = torch.device("cuda", local_rank)
        self.lib = _native.lib()

    def barrier(self):
        if self.world > 1:
            self.dist.barrier()
        self.torch.cuda.synchronize()

    def max_over_ranks(self, *vals):
        t = self.torch.tensor(list(vals), dtype=self.torch.float64, device=self.device)
        if self.world > 1:
            self.dist.all_reduce(t, op=self.dist.ReduceOp.MAX)
        return [float(v) for v in t]


def build_tuner(ctx, wl, channels=None):
    """Tuner with the workload's channels (all of them, or the given global indices with the band
    plan of the full list)."""
    N, Cn, B, A, kind, _ = wl
    rc = ctx.rc
    offs = tiling_offsets(N, Cn, B)
    tuner = rc.Tuner(cuda=True)
    for c in (range(Cn) if channels is None else channels):
        tuner.add_channel(100e6 + offs[c], B, getattr(rc, kind)(B, A, cuda=True))
    if channels is not None:
        tuner._input_frequency = 100e6
    tuner.request_bandwidth(N)
    return tuner


def time_steps(ctx, step, steps, warmup, profile=False):
    """Warm up, then time exactly `steps` calls of step() between barriers with CUDA events on the
    current stream; returns (ms_total (max over ranks), launches, per-kernel table or None, clocks)."""
    torch, lib = ctx.torch, ctx.lib
    for _ in range(max(warmup, 3)):
        step()
    ctx.barrier()
    lib.rc_profile_reset()
    lib.rc_profile_enable(1 if profile else 0)
    sampler = ClockSampler(ctx.local_rank)
    sampler.start()
    ev0, ev1 = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
    ctx.barrier()
    ev0.record()
    for _ in range(steps):
        step()
    ev1.record()
    ctx.barrier()
    clocks = sampler.summary()
    ms_total = ev0.elapsed_time(ev1)
    launches = int(lib.rc_profile_launches())
    kernels = None
    if profile:
        need = lib.rc_profile_report(None, 0)
        buf = C.create_string_buffer(need + 16)
        lib.rc_profile_report(buf, need + 16)
        kernels = json.loads(buf.value.decode())
    lib.rc_profile_enable(0)
    lib.rc_profile_reset()
    return ctx.max_over_ranks(ms_total)[0], launches, kernels, clocks


def e2e_pipeline(ctx, tuner, x_dev, steps):
    """Host IQ (pinned) -> Tuner.submit()/collect() -> host audio of every channel, every step: the H2D
    copy of block k+1 runs while block k is in the kernels and block k-1's audio is read back."""
    torch = ctx.torch
    n = x_dev.numel()
    x_host = [torch.empty(n, dtype=torch.complex64).pin_memory() for _ in range(2)]
    for xh in x_host:
        xh.copy_(x_dev)
    torch.cuda.synchronize()
    d2h = 4 * sum(size * nchn for _, size, nchn in tuner.audio_slices())

    def run(k):
        prev, checksum = None, 0.0
        for i in range(k):
            t = tuner.submit(x_host[i % 2])
            if prev is not None:
                checksum += float(tuner.collect(prev)[0])
            prev = t
        return checksum + float(tuner.collect(prev)[0])

    run(3)
    ctx.barrier()
    t0 = time.perf_counter()
    run(steps)
    torch.cuda.synchronize()
    t = time.perf_counter() - t0
    del x_host
    return {"t": ctx.max_over_ranks(t)[0], "h2d": 8 * n, "d2h": d2h,
            "api": "Tuner.submit()/collect(): pinned host IQ -> H2D -> kernels -> D2H of all channels' audio, 2 blocks in flight"}


def e2e_dropin(ctx, wl, x_dev, steps):
    """The reference's loop, literally (examples/multi_fm_server.py:87,95-106): a Buffer the radio
    thread would fill, `tuner.load(buffer.data)`, then per channel `tuner.run(i)`,
    `channel.demodulator.run(...)`, `.tobytes()`.  Synchronous, one block at a time."""
    torch, rc = ctx.torch, ctx.rc
    N, Cn, B, A, kind, _ = wl
    tuner = build_tuner(ctx, wl)
    buf = rc.Buffer(N, dtype="complex64", cuda=True)
    buf.data[:] = x_dev.cpu().numpy()
    sent = 0

    def block():
        nonlocal sent
        tuner.load(buf.data)
        for channel in tuner.channels():
            tmp = tuner.run(channel.index)
            tmp = channel.demodulator.run(tmp)
            sent += len(tmp.tobytes())

    for _ in range(2):
        block()
    torch.cuda.synchronize()
    sent = 0
    t0 = time.perf_counter()
    for _ in range(steps):
        block()
    torch.cuda.synchronize()
    t = time.perf_counter() - t0
    del tuner
    return {"value": N * steps / t / 1e6, "unit": UNIT, "ms_per_step": t / steps * 1e3, "h2d_bytes_per_step": 8 * N,
            "d2h_bytes_per_step": sent // steps,
            "api": "Buffer(cuda=True).data -> Tuner.load -> per channel Tuner.run + demodulator.run + tobytes "
                   "(examples/multi_fm_server.py:95-106 unmodified, synchronous)"}


def kernel_table(kernels, steps, peak_gbs):
    table = {}
    for tag, k in (kernels or {}).items():
        avg_ms = k["total_ms"] / max(k["count"], 1)
        gbs = k["bytes_per_launch"] / (avg_ms * 1e-3) / 1e9 if avg_ms > 0 else 0.0
        table[tag] = {"launches_per_step": k["count"] / steps, "avg_ms": round(avg_ms, 4),
                      "ms_per_step": round(k["total_ms"] / steps, 4),
                      "bytes_per_launch": k["bytes_per_launch"], "GBps": round(gbs, 1),
                      "frac_of_hbm_peak": round(gbs / peak_gbs, 4)}
    return table


def peaks():
    try:
        p = json.load(open(os.path.join(ROOT, "MEASURED_PEAKS.json")))
        if p.get("hbm_gbs"):
            return float(p["hbm_gbs"]), "measured (MEASURED_PEAKS.json hbm_gbs)"
    except Exception:
        pass
    return 6650.0, "fallback (B200_PROFILING.md)"


def path_roofline(wl, n_channels, ms_per_step, peak_gbs, streams=1, gpus=1):
    N, Cn, B, A, kind, _ = wl
    nch = 2 if kind == "WBFM" else 1
    algo = (8 * N + 4 * A * nch * n_channels) * streams
    gbs = algo / (ms_per_step * 1e-3) / 1e9
    return {"algorithmic_bytes_per_step": algo, "achieved": round(gbs, 1), "peak": peak_gbs * gpus, "unit": "GB/s",
            "frac": round(gbs / (peak_gbs * gpus), 4),
            "note": "SURVEY 8(d) bytes (read IQ once + write audio) over the whole step; peak = measured HBM peak x GPUs"}


# ---- one GPU: a workload on device-resident input
def measure_single(ctx, wl, x_dev, steps, warmup, profile, e2e, graph=False):
    tuner = build_tuner(ctx, wl)
    if graph:
        x_graph = tuner.graph_input()           # the block is resident where the captured graph reads it
        x_graph.copy_(x_dev)

    def step():
        if graph:
            tuner.step(x_graph)                 # the block's kernels as one captured CUDA graph
        else:
            tuner.load(x_dev)
            tuner.run_all()

    ms_total, launches, kernels, clocks = time_steps(ctx, step, steps, warmup, profile)
    res = {"ms_per_step": ms_total / steps, "launches": launches, "kernels": kernels, "clocks": clocks,
           "channels": len(tuner.channels())}
    if e2e:
        res["e2e"] = e2e_pipeline(ctx, tuner, x_dev, steps)
    del tuner
    ctx.torch.cuda.empty_cache()
    return res


# ---- N GPUs, one stream
def measure_sharded(ctx, wl, x_dev, steps, warmup, profile, e2e):
    """Channels sliced over the ranks, Tuner.load sharded (sharding.ShardedLoad): every rank holds
    its commutator branch x[rank::world] of the block."""
    torch, sharding = ctx.torch, ctx.sharding
    N, Cn, B, A, kind, _ = wl
    mine = list(sharding.channel_slice(Cn, ctx.world, ctx.rank))
    tuner = build_tuner(ctx, wl, mine)
    load = sharding.ShardedLoad(tuner)
    branch = x_dev[ctx.rank::ctx.world].contiguous()

    def step():
        load.post(branch)                           # block k+lanes: local FFT + exchanges on a side stream ...
        tuner.load_subband(load.take())             # ... while block k is demodulated
        tuner.run_all()

    for _ in range(load.lanes):                     # prime: one block ahead per pipeline lane
        load.post(branch)
    ms_total, launches, kernels, clocks = time_steps(ctx, step, steps, warmup, profile)
    res = {"ms_per_step": ms_total / steps, "launches": launches, "kernels": kernels, "clocks": clocks,
           "channels": len(mine), "nvlink_bytes_sent_per_rank_per_step": load.bytes_exchanged}
    if e2e:
        # every rank copies ITS branch from pinned host memory over its own PCIe link (the ingest side
        # deals samples round-robin to the ranks' buffers: the polyphase input commutator), then the
        # same sharded load; audio of its channels is read back every step
        m = branch.numel()
        copy_stream = torch.cuda.Stream()
        host = [torch.empty(m, dtype=torch.complex64).pin_memory() for _ in range(2)]
        for h in host:
            h.copy_(branch)
        nstage = load.lanes + 2
        stage = [torch.empty(m, dtype=torch.complex64, device=ctx.device) for _ in range(nstage)]
        d2h = 4 * sum(size * nchn for _, size, nchn in tuner.audio_slices())
        state = {"i": 0}

        def send_next():
            i = state["i"]
            state["i"] += 1
            ev = torch.cuda.Event()
            with torch.cuda.stream(copy_stream):
                stage[i % nstage].copy_(host[i % 2], non_blocking=True)
                ev.record(copy_stream)
            load.post(stage[i % nstage], ready=ev)

        def run(k):
            checksum = 0.0
            for _ in range(k):
                send_next()
                tuner.load_subband(load.take())
                checksum += float(tuner.run_all(numpy_output=True)[0])
            return checksum

        while load.in_flight():
            load.take()
        torch.cuda.synchronize()
        for _ in range(load.lanes):
            send_next()
        run(3)
        ctx.barrier()
        t0 = time.perf_counter()
        run(steps)
        torch.cuda.synchronize()
        t = time.perf_counter() - t0
        h2d, d2h_all = ctx.max_over_ranks(8.0 * m, float(d2h))
        res["e2e"] = {"t": ctx.max_over_ranks(t)[0], "h2d": int(h2d) * ctx.world, "d2h": int(d2h_all) * ctx.world,
                      "api": "every rank: pinned host branch x[rank::G] (samples dealt round-robin to the ranks' buffers by the "
                             "ingest side as they arrive; that host-side dealing is not timed) -> H2D over its own PCIe link -> "
                             "sharding.ShardedLoad (local FFT, 2 NVLink exchanges, combine) -> Tuner.load_subband + run_all "
                             "-> D2H of its channels' audio"}
    while load.in_flight():
        load.take()
    torch.cuda.synchronize()
    # decomposition, outside the timed value: the load pipeline on its own (phases on the side stream,
    # nothing else running), then the channel stage on its own
    load.phase_events = []
    sub = None
    for _ in range(3):
        load.post(branch)
        sub = load.take()
        torch.cuda.synchronize()
    phases = load.phase_ms()
    load.phase_events = None
    tuner.load_subband(sub)
    tuner.run_all()
    ev0, ev1 = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
    ev0.record()
    for _ in range(3):
        tuner.load_subband(sub)
        tuner.run_all()
    ev1.record()
    torch.cuda.synchronize()
    res["decomposition"] = {"transport": load.transport, "lanes": load.lanes, "fused": bool(getattr(load, "fused", False)), "load_pipeline_ms": {k: round(v, 4) for k, v in phases.items()},
                            "load_pipeline_total_ms": round(sum(phases.values()), 4),
                            "channel_stage_ms": round(ev0.elapsed_time(ev1) / 3, 4),
                            "note": "each measured alone on rank 0; in the timed step block k+1's load pipeline overlaps block k's channel stage"}
    del load, tuner
    torch.cuda.empty_cache()
    return res


def measure_bcast(ctx, wl, x_dev, steps, warmup):
    """One stream replicated by an NCCL broadcast per block (posted one block ahead); every rank
    repeats the N-point FFT and demodulates its channel slice."""
    sharding = ctx.sharding
    N, Cn = wl[0], wl[1]
    mine = list(sharding.channel_slice(Cn, ctx.world, ctx.rank))
    tuner = build_tuner(ctx, wl, mine)
    feed = sharding.BlockBroadcaster(N, ctx.device, src=0)
    src = x_dev if ctx.rank == 0 else None

    def step():
        feed.post(src)
        tuner.load(feed.take())
        tuner.run_all()

    feed.post(src)
    ms_total, _, _, _ = time_steps(ctx, step, steps, warmup)
    while feed.in_flight():
        feed.take()
    ctx.torch.cuda.synchronize()
    del tuner, feed
    ctx.torch.cuda.empty_cache()
    return {"ms_per_step": ms_total / steps, "value": N / (ms_total / steps * 1e-3) / 1e6, "unit": UNIT, "scaling": "strong",
            "what": "one stream, NCCL broadcast of the block per step, every rank repeats the N-point FFT"}


def run_b200(args, rank, world, local_rank):
    import torch
    torch.cuda.set_device(local_rank)
    ctx = Ctx(args, rank, world, local_rank)
    wl = WORKLOADS[args.workload]
    N, Cn, B, A, kind, desc = wl
    mode = args.mode
    if mode == "auto":
        mode = "sharded" if world > 1 else "single"
    if world == 1:
        mode = "single"
    peak_gbs, peak_src = peaks()
    steps, warmup = args.steps, max(args.warmup, 3)

    # every rank synthesises the same block (same seed, same arithmetic); `independent` gives each its own
    x_dev, _ = make_wideband_gpu(N, Cn, B, 3 + (rank if mode == "independent" else 0), kind == "WBFM", ctx.device)

    extras = {}
    if mode == "single" or mode == "independent":
        res = measure_single(ctx, wl, x_dev, steps, warmup, True, not args.no_e2e)
        streams = world if mode == "independent" else 1
        scaling = "weak"
    elif mode == "sharded":
        try:
            res = measure_sharded(ctx, wl, x_dev, steps, warmup, True, not args.no_e2e)
        except ValueError as exc:                    # block length not a multiple of world_size**2: broadcast instead
            r = measure_bcast(ctx, wl, x_dev, steps, warmup)
            res = {"ms_per_step": r["ms_per_step"], "launches": 0, "kernels": None, "clocks": None, "channels": Cn // world}
            mode = "bcast"
            sys.stderr.write(f"bench.py: sharded load unavailable ({exc}); one stream by NCCL broadcast\n")
        streams, scaling = 1, "strong"
    else:
        r = measure_bcast(ctx, wl, x_dev, steps, warmup)
        res = {"ms_per_step": r["ms_per_step"], "launches": 0, "kernels": None, "clocks": None, "channels": Cn // world}
        streams, scaling = 1, "strong"
    ms_per_step = res["ms_per_step"]
    value = N * streams / (ms_per_step * 1e-3) / 1e6

    # ---- extra keys (never change the line's value)
    if not args.no_extras and args.workload == "cfg3":
        short_steps = max(3, min(steps, 10))
        if world == 1:
            try:
                w2 = WORKLOADS["cfg3-wbfm"]
                r = measure_single(ctx, w2, x_dev, short_steps, 3, False, not args.no_e2e)
                extras["wbfm_chain"] = {
                    "workload": w2[5], "value": N / (r["ms_per_step"] * 1e-3) / 1e6, "unit": UNIT, "ms_per_step": r["ms_per_step"],
                    "roofline_path": path_roofline(w2, Cn, r["ms_per_step"], peak_gbs),
                    "e2e": ({"value": N * short_steps / r["e2e"]["t"] / 1e6, "unit": UNIT, "h2d_bytes_per_step": r["e2e"]["h2d"],
                             "d2h_bytes_per_step": r["e2e"]["d2h"]} if "e2e" in r else None),
                    "parity": "tests/test_gpu_parity.py::test_config3_literal_block (WBFM, B=1e6, 2 blocks, ch 0/1/127/255 vs oracle)"}
            except Exception as exc:
                extras["wbfm_chain"] = {"failed": repr(exc)}
            try:
                w3 = WORKLOADS["cfg3-short"]
                xs, _ = make_wideband_gpu(w3[0], w3[1], w3[2], 3, False, ctx.device)
                r_eager = measure_single(ctx, w3, xs, max(32, short_steps), 3, False, False)
                r = measure_single(ctx, w3, xs, max(32, short_steps), 3, False, False, graph=True)
                del xs
                extras["short_block"] = {
                    "workload": w3[5], "value": w3[0] / (r["ms_per_step"] * 1e-3) / 1e6, "unit": UNIT, "ms_per_block": r["ms_per_step"],
                    "ms_per_second_of_signal": r["ms_per_step"] * 32,
                    "launch": "one CUDA graph per block (Tuner.step on Tuner.graph_input())",
                    "eager_ms_per_block": r_eager["ms_per_step"], "launches_per_block": r_eager["launches"] / max(32, short_steps),
                    "roofline_path": path_roofline(w3, w3[1], r["ms_per_step"], peak_gbs),
                    "l2": "block (64 MB) and every intermediate fit the 126 MB L2: steady-state L2-resident run",
                    "parity": "tests/test_gpu_parity.py::test_config3_short_block (bin-unit oracle, every channel)"}
            except Exception as exc:
                extras["short_block"] = {"failed": repr(exc)}
            if not args.no_e2e:
                try:
                    extras["e2e_dropin"] = e2e_dropin(ctx, wl, x_dev, max(3, min(steps, 5)))
                except Exception as exc:
                    extras["e2e_dropin"] = {"failed": repr(exc)}
        else:
            try:
                extras["bcast"] = measure_bcast(ctx, wl, x_dev, short_steps, 3)
            except Exception as exc:
                extras["bcast"] = {"failed": repr(exc)}
            try:
                r = measure_single(ctx, wl, x_dev, short_steps, 3, False, False)
                extras["replicas"] = {"value": N * world / (r["ms_per_step"] * 1e-3) / 1e6, "unit": UNIT, "ms_per_step": r["ms_per_step"],
                                      "scaling": "weak", "what": "independent sub-band stream per GPU (256 channels each), no data-path collective"}
            except Exception as exc:
                extras["replicas"] = {"failed": repr(exc)}
            if world == 8:
                try:
                    del x_dev
                    torch.cuda.empty_cache()
                    w5 = WORKLOADS["cfg5"]
                    x5, _ = make_wideband_gpu(w5[0], w5[1], w5[2], 5, False, ctx.device)
                    r = measure_sharded(ctx, w5, x5, max(3, min(steps, 5)), 3, False, False)
                    del x5
                    extras["cfg5"] = {"workload": w5[5], "value": w5[0] / (r["ms_per_step"] * 1e-3) / 1e6, "unit": UNIT,
                                      "ms_per_step": r["ms_per_step"], "channels_per_gpu": r["channels"], "scaling": "strong",
                                      "nvlink_bytes_sent_per_rank_per_step": r["nvlink_bytes_sent_per_rank_per_step"],
                                      "parity": "tests/test_gpu_parity.py::test_config5_one_gpu_slice; tests/test_gpu_multi.py (sharded load)"}
                    x_dev = None
                except Exception as exc:
                    extras["cfg5"] = {"failed": repr(exc)}

    if rank != 0:
        return

    table = kernel_table(res.get("kernels"), steps, peak_gbs)
    top = max(table, key=lambda t: table[t]["ms_per_step"]) if table else None
    traffic = None
    try:
        traffic = json.load(open(os.path.join(ROOT, "profiles", "dram_traffic.json"))).get(args.workload if world == 1 else "", {}).get(top)
    except Exception:
        pass
    # kernels whose stores ARE an NVLink exchange (sharded load, fused transport): bound by the link, not HBM
    nvl = res.get("nvlink_bytes_sent_per_rank_per_step")
    fused_tags = []
    if nvl and res.get("decomposition", {}).get("fused"):
        last_fft = [t for t in table if t.startswith("tuner.local_fft/")]
        fused_tags = [t for t in (last_fft[-1:] + ["tuner.subband_combine_scatter"]) if t in table]
        for t in fused_tags:
            g = (nvl / 2) / (table[t]["avg_ms"] * 1e-3) / 1e9
            table[t]["nvlink_bytes_per_launch"] = nvl / 2
            table[t]["nvlink_GBps"] = round(g, 1)
            table[t]["frac_of_nvlink_peak"] = round(g / NVLINK_PEAK_GBS, 4)
    roofline = None
    if top:
        kt = table[top]
        roofline = {"kernel": top, "bound": "hbm", "achieved": kt["GBps"], "peak": peak_gbs, "unit": "GB/s",
                    "frac": round(kt["GBps"] / peak_gbs, 4), "traffic": traffic, "peak_source": peak_src,
                    "avg_launch_ms": kt["avg_ms"], "share_of_step": round(kt["ms_per_step"] / ms_per_step, 4)}
        if top in fused_tags:
            roofline.update({"bound": "nvlink", "achieved": kt["nvlink_GBps"], "peak": NVLINK_PEAK_GBS,
                             "frac": kt["frac_of_nvlink_peak"], "traffic": None,
                             "peak_source": "measured peer-copy bandwidth per direction per GPU (B200_PROFILING.md: 770 GB/s)",
                             "note": "fused compute + exchange kernel: its remote stores cross NVLink; HBM side: %.1f GB/s" % kt["GBps"]})
    multi = {"single": "n/a", "independent": "independent sub-band stream per GPU, no data-path collective",
             "sharded": "one stream: channel slices + sharded Tuner.load (commutator branches, local N/G-point FFT, "
                        "two NVLink all-to-all exchanges, radix-G combine), no reduction",
             "bcast": "one stream, NCCL broadcast per block, channel slices"}[mode]
    line = {"metric": METRIC, "value": value, "unit": UNIT, "n_gpus": world, "steps": steps,
            "warmup": warmup, "ms_per_step": ms_per_step, "higher_is_better": True,
            "scaling": scaling, "vs_baseline": None, "dtype": "f32", "data": "synthetic",
            "config": config_of(wl),
            "run": {"channels_per_gpu": res["channels"], "multi_gpu": multi,
                    "l2": "input block (%.0f MB) larger than the 126 MB L2, no flush needed" % (8 * N / 1e6)
                          if 8 * N > 130e6 else "input smaller than L2: steady-state L2-resident run"},
            "gpu_launches": res["launches"],
            "roofline": roofline,
            "roofline_path": path_roofline(wl, Cn, ms_per_step, peak_gbs, streams, world),
            "kernels": table, "clocks": res["clocks"]}
    if "nvlink_bytes_sent_per_rank_per_step" in res:
        line["run"]["nvlink_bytes_sent_per_rank_per_step"] = res["nvlink_bytes_sent_per_rank_per_step"]
    if "decomposition" in res:
        line["run"]["decomposition"] = res["decomposition"]
    if "e2e" in res:
        e = res["e2e"]
        line["e2e"] = {"value": N * streams * steps / e["t"] / 1e6, "unit": UNIT,
                       "h2d_bytes_per_step": e["h2d"] * (streams if mode != "sharded" else 1),
                       "d2h_bytes_per_step": e["d2h"] * (world if mode == "independent" else 1),
                       "ms_per_step": e["t"] / steps * 1e3, "timing": "wall clock between device synchronisations, max over ranks",
                       "api": e["api"]}
    line.update(extras)
    if world == 1 and not args.no_cpu_baseline:
        try:
            from bench_support import cpu_arm
            x_host_np = x_dev.cpu().numpy()
            r = cpu_arm.measure(N, Cn, B, A, kind, tiling_offsets(N, Cn, B), x_host_np, steps=1, warmup=0, step_seconds=6.0)
            line["cpu_baseline"] = {"value": r["value"], "unit": UNIT, "cores": r["cores"], "kind": r["kind"], "sample": r["sample"],
                                    "single_core": r["single_core"]}
        except Exception as exc:       # never lose the GPU numbers to a host-side problem
            line["cpu_baseline"] = {"value": None, "unit": UNIT, "cores": os.cpu_count(), "kind": "reference",
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
