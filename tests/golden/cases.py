"""Seeded scenarios exercised identically against the reference (to make the
golden vectors), the oracle and the CUDA drop-in.

Every function takes ``impl`` -- a namespace exposing the reference's class
surface (Tuner, FM, MFM, WBFM, Decimate, Deemphasis, Bandpass, PLL) -- and
returns a dict of float64/complex128 arrays.  Demodulators called directly get
complex128 input (the reference's float32 ``unwrap`` on complex64
input is wrong by up to 1e-2; behind Tuner.run / Decimate.run the input is
complex128 anyway).
"""
import numpy as np

from bench_support import synth


def pin(a, limit=4096):
    """Reduce a long 1-D signal to head + tail + strided interior samples.

    Applied identically to every implementation, so the stored golden vectors
    stay small while still pinning both block edges and the interior.
    """
    a = np.asarray(a)
    flat = a.reshape(-1) if a.ndim <= 1 else a.reshape(-1, a.shape[-1]) if a.shape[-1] <= 2 else a.reshape(-1)
    n = flat.shape[0]
    if n <= limit:
        return flat
    step = max(1, n // 1500) | 1
    return np.concatenate((flat[:640], flat[-640:], flat[640:-640:step]))


def _f64(a):
    if hasattr(a, "detach"):                 # torch tensor (possibly on the GPU)
        a = a.detach().cpu().numpy()
    a = np.asarray(a)
    return pin(a.astype(np.complex128 if np.iscomplexobj(a) else np.float64))


def _tuner_setup(impl, N, B, A, demod_cls, offsets, f0=100e6, **kw):
    tuner = impl.Tuner()
    for off in offsets:
        demod = demod_cls(B, A, **kw) if demod_cls is not None else None
        tuner.add_channel(f0 + off, B, demod)
    tuner.request_bandwidth(N)
    return tuner


def tuner_mfm(impl, np_=np):
    """N=80000 -> 4 x 20000 -> MFM 4000, two phase-continuous blocks."""
    N, B, A, C = 80000, 20000, 4000, 4
    offs = synth.tiling_centers(N, C, B)
    tuner = _tuner_setup(impl, N, B, A, impl.MFM, offs)
    out = {}
    for blk in range(2):
        x = synth.wideband(N, offs, B, seed=42, block=blk)
        tuner.load(x)
        for ch in tuner.channels():
            iq = tuner.run(ch.index)
            if blk == 0 and ch.index in (0, 3):
                out[f"iq{ch.index}"] = _f64(iq)
            out[f"b{blk}c{ch.index}"] = _f64(ch.demodulator.run(iq))
    return out


def tuner_offgrid_fm(impl, np_=np):
    """Channels that do not tile the band (overlap, odd offsets), generic FM."""
    N, B, A = 96000, 24000, 4800
    offs = [-30001.0, -2500.0, 17.0, 33000.0]
    tuner = _tuner_setup(impl, N, B, A, impl.FM, offs)
    x = synth.wideband(N, offs, B, seed=7)
    tuner.load(x)
    out = {"f_in": np.array([tuner.input_frequency, tuner.input_bandwidth])}
    for ch in tuner.channels():
        out[f"c{ch.index}"] = _f64(ch.demodulator.run(tuner.run(ch.index)))
    return out


def tuner_single_channel(impl, np_=np):
    """One channel and no request_bandwidth: the plan's input bandwidth equals the channel's
    (tuner.py:163-174), so Tuner.run is window multiply + inverse FFT of every bin (num == Nx)."""
    B, A = 24000, 4800
    tuner = impl.Tuner()
    tuner.add_channel(98.7e6, B, impl.MFM(B, A))
    x = synth.wideband(B, [0.0], B, seed=12)
    tuner.load(x)
    iq = tuner.run(0)
    return {"f_in": np.array([tuner.input_frequency, tuner.input_bandwidth]), "iq": _f64(iq),
            "audio": _f64(tuner.channels()[0].demodulator.run(iq))}


def fm_direct(impl, np_=np):
    B, A = 25000, 4800
    x = synth.station(B, B, 3, offset_hz=1234.0, deviation=0.3 * B).astype(np.complex64)
    fm = impl.FM(B, A)
    return {"audio": _f64(fm.run(x.astype(np.complex128)))}


def wbfm_direct(impl, np_=np):
    """WBFM 64000 -> 16000, stereo multiplex, three blocks (state carry)."""
    B, A = 64000, 16000
    rng = np.random.default_rng(11)
    demod = impl.WBFM(B, A)
    out = {}
    for blk in range(3):
        x = synth.station(B, B, 1, offset_hz=-700.0, deviation=15000.0, stereo=True, block=blk)
        x = x + 0.01 * (rng.standard_normal(B) + 1j * rng.standard_normal(B))
        x = x.astype(np.complex64).astype(np.complex128)
        out[f"b{blk}"] = _f64(demod.run(x))
    return out


def wbfm_50us(impl, np_=np):
    B, A = 60000, 12000
    demod = impl.WBFM(B, A, deemphasis=50e-6)
    x = synth.station(B, B, 2, offset_hz=311.0, deviation=12000.0, stereo=True)
    x = x.astype(np.complex64).astype(np.complex128)
    return {"b0": _f64(demod.run(x))}


def decimate_cases(impl, np_=np):
    rng = np.random.default_rng(5)
    out = {}
    for n_in, n_out in ((25000, 2500), (6000, 6000), (4000, 6000), (3125, 625), (2187, 729)):
        xr = rng.standard_normal(n_in)
        xc = (rng.standard_normal(n_in) + 1j * rng.standard_normal(n_in)).astype(np.complex64)
        out[f"r{n_in}_{n_out}"] = _f64(impl.Decimate(n_in, n_out).run(xr))
        out[f"c{n_in}_{n_out}"] = _f64(impl.Decimate(n_in, n_out).run(xc))
    return out


def deemphasis_cases(impl, np_=np):
    rng = np.random.default_rng(6)
    de = impl.Deemphasis(4800)
    de50 = impl.Deemphasis(4800, 50e-6)
    out = {}
    for blk in range(3):
        x = rng.standard_normal(4800)
        out[f"b{blk}"] = _f64(de.run(x))
        out[f"e{blk}"] = _f64(de50.run(x))
    return out


def bandpass_pll_cases(impl, np_=np):
    rng = np.random.default_rng(8)
    n = 64000
    t = np.arange(n) / n
    x = 0.1 * np.sin(2 * np.pi * 19e3 * t + 0.3) + 0.3 * rng.standard_normal(n)
    bp = impl.Bandpass(n, 19e3 - 50, 19e3 + 50, num_taps=41)
    p = bp.run(x)
    pll = impl.PLL()
    pll.step(p)
    out = {"pilot": _f64(p), "image2": _f64(pll.image(2)), "real1": _f64(pll.real(1.0)),
           "image3": _f64(pll.image(3.0))}
    bp2 = impl.Bandpass(5000, 300, 900)          # default 61 taps
    out["bp61"] = _f64(bp2.run(rng.standard_normal(5000)))
    return out


CASES = {
    "tuner_mfm": tuner_mfm,
    "tuner_offgrid_fm": tuner_offgrid_fm,
    "tuner_single_channel": tuner_single_channel,
    "fm_direct": fm_direct,
    "wbfm_direct": wbfm_direct,
    "wbfm_50us": wbfm_50us,
    "decimate": decimate_cases,
    "deemphasis": deemphasis_cases,
    "bandpass_pll": bandpass_pll_cases,
}
