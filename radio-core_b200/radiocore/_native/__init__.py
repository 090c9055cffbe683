"""ctypes binding of libradiocore_b200.so (the C ABI in include/radiocore_b200.h).

There is no CPU fallback: if the shared library has not been built
(``python -c "import __graft_entry__ as g; g.build()"``) or no CUDA device is
visible, the first use raises ``RuntimeError``.
"""
import ctypes as C
import os

_HERE = os.path.dirname(os.path.abspath(__file__))
# RADIOCORE_B200_LIB: point at an alternative build of the same C ABI (kernel experiments)
LIB_PATH = os.environ.get("RADIOCORE_B200_LIB") or os.path.join(_HERE, "libradiocore_b200.so")

_lib = None

_i64, _int, _dbl, _vp, _fp = C.c_int64, C.c_int, C.c_double, C.c_void_p, C.c_void_p
_pp = C.POINTER(C.c_void_p)

_SIGNATURES = {
    "rc_version": ([], _int),
    "rc_size_supported": ([_i64], _int),
    "rc_engine_create": ([_int, _i64, _pp], _int),
    "rc_engine_destroy": ([_vp], _int),
    "rc_engine_add_channel": ([_vp, _i64, _i64, _i64, _int, _dbl, C.POINTER(_int)], _int),
    "rc_engine_commit": ([_vp], _int),
    "rc_engine_audio_floats": ([_vp, C.POINTER(_i64)], _int),
    "rc_engine_channel_layout": ([_vp, _int, C.POINTER(_i64), C.POINTER(_i64), C.POINTER(_int)], _int),
    "rc_engine_load": ([_vp, _vp, _vp], _int),
    "rc_engine_run": ([_vp, _fp, _vp], _int),
    "rc_engine_channel_iq": ([_vp, _int, _vp, _vp], _int),
    "rc_engine_spectrum": ([_vp, _vp, _vp], _int),
    "rc_engine_reset_state": ([_vp], _int),
    "rc_engine_workspace_bytes": ([_vp, C.POINTER(_i64)], _int),
    "rc_engine_set_subband": ([_vp, _i64, _i64], _int),
    "rc_engine_load_subband": ([_vp, _vp], _int),
    "rc_subband_combine": ([_int, _int, _i64, _i64, _i64, _vp, _vp, _vp], _int),
    "rc_fft_create": ([_int, _i64, _int, _pp], _int),
    "rc_fft_destroy": ([_vp], _int),
    "rc_fft_exec": ([_vp, _int, _vp, _vp, _vp], _int),
    "rc_fft_exec_scatter": ([_vp, _int, _vp, C.POINTER(C.c_void_p), _int, _i64, _vp], _int),
    "rc_subband_combine_scatter": ([_int, _int, _i64, _i64, _i64, _vp, _vp, _int, _vp], _int),
    "rc_demod_create": ([_int, _int, _i64, _i64, _dbl, _int, _pp], _int),
    "rc_demod_destroy": ([_vp], _int),
    "rc_demod_run": ([_vp, _vp, _fp, _vp], _int),
    "rc_demod_reset_state": ([_vp], _int),
    "rc_decimate_create": ([_int, _i64, _i64, _pp], _int),
    "rc_decimate_destroy": ([_vp], _int),
    "rc_decimate_run_real": ([_vp, _fp, _fp, _vp], _int),
    "rc_decimate_run_complex": ([_vp, _vp, _vp, _vp], _int),
    "rc_deemph_create": ([_int, _i64, _dbl, _pp], _int),
    "rc_deemph_destroy": ([_vp], _int),
    "rc_deemph_run": ([_vp, _fp, _fp, _vp], _int),
    "rc_deemph_reset_state": ([_vp], _int),
    "rc_deemph_taps": ([_dbl, _i64, C.POINTER(C.c_float), C.POINTER(C.c_float)], _int),
    "rc_bandpass_create": ([_int, _i64, _dbl, _dbl, _int, C.c_char_p, _pp], _int),
    "rc_bandpass_destroy": ([_vp], _int),
    "rc_bandpass_run": ([_vp, _fp, _fp, _vp], _int),
    "rc_bandpass_taps": ([_vp, C.POINTER(C.c_float), _int], _int),
    "rc_pll_create": ([_int, _i64, _pp], _int),
    "rc_pll_destroy": ([_vp], _int),
    "rc_pll_step": ([_vp, _fp, _vp], _int),
    "rc_pll_eval": ([_vp, _dbl, _int, _fp, _vp], _int),
    "rc_profile_enable": ([_int], _int),
    "rc_profile_reset": ([], _int),
    "rc_profile_launches": ([], _i64),
    "rc_profile_report": ([C.c_char_p, _int], _int),
    "rc_fft_c2c": ([_int, _i64, _int, _int, _vp, _vp, _vp], _int),
}



class ScatterSeg(C.Structure):
    """rc_scatter_seg of include/radiocore_b200.h."""
    _fields_ = [("k1", C.c_int32), ("reserved", C.c_int32), ("j_lo", C.c_int64), ("j_hi", C.c_int64), ("dst", C.c_void_p)]


EXPORTED_SYMBOLS = ["rc_last_error"] + sorted(_SIGNATURES)

RC_ERR_INVALID, RC_ERR_UNSUPPORTED, RC_ERR_CUDA, RC_ERR_STATE = -1, -2, -3, -4


def load_library(path=None):
    """dlopen the shared library and attach argtypes; no CUDA call is made."""
    global _lib
    if _lib is not None and path is None:
        return _lib
    p = path or LIB_PATH
    if not os.path.exists(p):
        raise RuntimeError(
            f"radiocore (B200): native library not found at {p}; build it with "
            "`python -c 'import __graft_entry__ as g; g.build()'`. There is no CPU fallback.")
    lib = C.CDLL(p)
    lib.rc_last_error.restype = C.c_char_p
    lib.rc_last_error.argtypes = []
    for name, (args, res) in _SIGNATURES.items():
        fn = getattr(lib, name)
        fn.argtypes = args
        fn.restype = res
    if path is None:
        _lib = lib
    return lib


def lib():
    return load_library()


def check(rc):
    """Map a negative rc_status to the exception the reference would raise."""
    if rc >= 0:
        return rc
    msg = lib().rc_last_error().decode(errors="replace")
    if rc in (RC_ERR_INVALID, RC_ERR_UNSUPPORTED):
        raise ValueError(msg)
    raise RuntimeError(msg)
