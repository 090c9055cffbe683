"""Generate tests/golden/*.npz by running the REAL reference (build container only).

    python tests/golden/make_golden.py

Imports /root/reference through oracle/ref_shim.py (atomics stub), feeds it
seeded inputs from bench_support/synth.py and stores the float64 outputs.
Inputs are regenerated from the seeds at test time (numpy's PCG64 streams are
stable), only outputs are stored.  The NumPy/SciPy versions are recorded so
SciPy drift (scipy.signal.resample changed behaviour across versions) is
detectable.
"""
import os
import sys

import numpy as np
import scipy

HERE = os.path.dirname(os.path.abspath(__file__))
ROOT = os.path.dirname(os.path.dirname(HERE))
sys.path.insert(0, ROOT)
sys.path.insert(0, os.path.join(ROOT, "oracle"))

import ref_shim                                    # noqa: E402
from bench_support import synth                    # noqa: E402
from tests.golden import cases                     # noqa: E402


def main():
    ref = ref_shim.load_reference()
    out = {"versions": np.array([np.__version__, scipy.__version__, "209dc88"])}
    for name, fn in cases.CASES.items():
        res = fn(ref, np)
        for k, v in res.items():
            out[f"{name}/{k}"] = np.asarray(v)
        print(name, {k: np.asarray(v).shape for k, v in res.items()})
    path = os.path.join(HERE, "golden.npz")
    np.savez_compressed(path, **out)
    print("wrote", path, os.path.getsize(path), "bytes")


if __name__ == "__main__":
    main()
