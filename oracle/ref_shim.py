"""Import the real, unmodified reference package under the name ``_ref_radiocore``.

TEST / BASELINE INFRASTRUCTURE ONLY.  Source, in order: ``$RADIOCORE_REFERENCE``,
``/root/reference`` (build container, read-only), ``oracle/_ref`` (the byte-for-byte copy that
``oracle/make_ref.py`` makes so that the reference travels to the GPU box).  Used to (a) validate
``radiocore_oracle.py`` live, (b) generate the golden vectors under ``tests/golden/``
(``tests/golden/make_golden.py``) and (c) time the reference itself in ``bench.py``'s CPU legs.

The reference's package ``__init__`` imports ``atomics`` (PyPI, absent here)
through ``radiocore/tools/ringbuffer.py:3``; only RingBuffer uses it, so an
in-memory stand-in is registered before the import.
"""
import importlib
import os
import sys
import types

def _find_root():
    here = os.path.dirname(os.path.abspath(__file__))
    for cand in (os.environ.get("RADIOCORE_REFERENCE"), "/root/reference", os.path.join(here, "_ref")):
        if cand and os.path.isdir(os.path.join(cand, "radiocore")):
            return cand
    return os.environ.get("RADIOCORE_REFERENCE", "/root/reference")


REFERENCE_ROOT = _find_root()


def available() -> bool:
    return os.path.isdir(os.path.join(REFERENCE_ROOT, "radiocore"))


def _install_atomics_stub():
    if "atomics" in sys.modules:
        return
    mod = types.ModuleType("atomics")

    class _Atomic:
        def __init__(self):
            self._v = 0

        def load(self):
            return self._v

        def store(self, v):
            self._v = v

        def add(self, v):
            self._v += v

        def sub(self, v):
            self._v -= v

    mod.INT = "INT"
    mod.atomic = lambda width=4, atype=None: _Atomic()
    sys.modules["atomics"] = mod


def load_reference():
    """Return the reference ``radiocore`` module under the name ``_ref_radiocore``.

    The product drop-in is also called ``radiocore``; to keep both importable
    in one process the reference is imported first-come under its own name and
    then moved aside in ``sys.modules``.
    """
    if "_ref_radiocore" in sys.modules:
        return sys.modules["_ref_radiocore"]
    if not available():
        raise ImportError(f"reference not found at {REFERENCE_ROOT}")
    _install_atomics_stub()
    saved = {k: v for k, v in sys.modules.items()
             if k == "radiocore" or k.startswith("radiocore.")}
    for k in saved:
        del sys.modules[k]
    sys.path.insert(0, REFERENCE_ROOT)
    try:
        ref = importlib.import_module("radiocore")
    finally:
        sys.path.remove(REFERENCE_ROOT)
    for k in [k for k in sys.modules if k == "radiocore" or k.startswith("radiocore.")]:
        sys.modules["_ref_" + k] = sys.modules.pop(k)
    sys.modules.update(saved)
    return ref
