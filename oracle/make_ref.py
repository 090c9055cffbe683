"""Recipe: vendor the UNMODIFIED reference package into ``oracle/_ref/`` (git-ignored).

TEST / BASELINE INFRASTRUCTURE ONLY.  ``/root/reference`` exists in the build container but not on
the GPU box; ``gpurun`` ships the repository tree including git-ignored files, so a byte-for-byte
copy of the reference's pure-Python package under ``oracle/_ref/radiocore`` travels with it and
``bench.py --impl reference`` / ``cpu_baseline`` can time the reference ITSELF on the box's host
cores (``kind: "reference"``).  Nothing is edited: the only addition is ``oracle/_ref/atomics.py``,
a stand-in for the PyPI module ``atomics`` that the reference's ``RingBuffer`` imports
(radiocore/tools/ringbuffer.py:3) and that is absent from this image.  ``PROVENANCE.json`` records
the source path, the sha256 of every copied file and the NumPy / SciPy versions.

    python oracle/make_ref.py            # no-op when /root/reference is absent (GPU box)

Called by ``__graft_entry__.build()``.  The product package never imports anything from here.
"""
import hashlib
import json
import os
import shutil
import sys

HERE = os.path.dirname(os.path.abspath(__file__))
SRC = os.environ.get("RADIOCORE_REFERENCE", "/root/reference")
DST = os.path.join(HERE, "_ref")

ATOMICS_STUB = '''"""Stand-in for the PyPI package `atomics` (absent here): the reference's RingBuffer keeps its
occupancy counter in atomics.atomic(width=4, atype=atomics.INT) (radiocore/tools/ringbuffer.py:46)."""
import threading

INT = "INT"


class _Atomic:
    def __init__(self):
        self._v, self._m = 0, threading.Lock()

    def load(self):
        return self._v

    def store(self, v):
        with self._m:
            self._v = v

    def add(self, v):
        with self._m:
            self._v += v

    def sub(self, v):
        with self._m:
            self._v -= v


def atomic(width=4, atype=None):
    return _Atomic()
'''


def make(force: bool = False):
    """Copy the reference package; returns the destination or None when there is no source."""
    src_pkg = os.path.join(SRC, "radiocore")
    if not os.path.isdir(src_pkg):
        return DST if os.path.isdir(os.path.join(DST, "radiocore")) else None
    if os.path.isdir(DST) and not force and os.path.exists(os.path.join(DST, "PROVENANCE.json")):
        return DST
    if os.path.isdir(DST):
        shutil.rmtree(DST)
    os.makedirs(DST)
    shutil.copytree(src_pkg, os.path.join(DST, "radiocore"), ignore=shutil.ignore_patterns("__pycache__"))
    for extra in ("examples", "tests"):             # the example scripts / host tests run unmodified in tests/
        if os.path.isdir(os.path.join(SRC, extra)):
            shutil.copytree(os.path.join(SRC, extra), os.path.join(DST, extra), ignore=shutil.ignore_patterns("__pycache__"))
    open(os.path.join(DST, "atomics.py"), "w").write(ATOMICS_STUB)
    files = {}
    for root, _, names in os.walk(DST):
        for n in sorted(names):
            p = os.path.join(root, n)
            files[os.path.relpath(p, DST)] = hashlib.sha256(open(p, "rb").read()).hexdigest()
    import numpy
    import scipy
    json.dump({"source": SRC, "commit": "209dc88", "numpy": numpy.__version__, "scipy": scipy.__version__,
               "files": files}, open(os.path.join(DST, "PROVENANCE.json"), "w"), indent=1)
    return DST


if __name__ == "__main__":
    print(make(force="--force" in sys.argv))
