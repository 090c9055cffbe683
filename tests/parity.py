"""The parity metric (BASELINE.md section 4 / SURVEY.md 7.3-7).

north_star: "audio matches the reference's NumPy/SciPy path to within 1e-5
relative per sample".  Per-sample relative error is undefined at zero
crossings, so it is made precise as: for each channel-block

    max|a - b| <= 1e-5 * max|b|       and
    |a - b|    <= 1e-5 * rms(b) + 1e-5 * |b|   for every sample.
"""
import numpy as np

TOL = 1e-5


def errors(got, ref):
    got = np.asarray(got, dtype=np.float64) if not np.iscomplexobj(got) else np.asarray(got, dtype=np.complex128)
    ref = np.asarray(ref)
    assert got.shape == ref.shape, (got.shape, ref.shape)
    diff = np.abs(got - ref)
    peak = float(np.max(np.abs(ref)))
    rms = float(np.sqrt(np.mean(np.abs(ref) ** 2)))
    rel_peak = float(np.max(diff)) / max(peak, 1e-300)
    margin = float(np.max(diff / (TOL * rms + TOL * np.abs(ref) + 1e-300)))
    return rel_peak, margin


def assert_parity(got, ref, what="", tol_scale=1.0):
    rel_peak, margin = errors(got, ref)
    assert rel_peak <= TOL * tol_scale, f"{what}: max|a-b|/max|b| = {rel_peak:.3e}"
    assert margin <= tol_scale, f"{what}: per-sample criterion exceeded by x{margin:.2f}"
    return rel_peak
