"""ctypes binding of ``libspeechdsp.so`` (C ABI declared in ``include/speechdsp.h``).

There is no CPU fallback: if the shared library has not been built, or there is no CUDA
device, every compute call raises.  Build with ``python -c "import __graft_entry__ as g; g.build()"``
or ``python -m speech_cloner_b200.build``.
"""
from __future__ import annotations

import ctypes as C
import os

_HERE = os.path.dirname(os.path.abspath(__file__))
LIB_PATH = os.environ.get("SPEECHDSP_LIB") or os.path.join(_HERE, "libspeechdsp.so")   # override: A/B-testing builds

SC_OK = 0
SC_ERR_INVALID = -1
SC_ERR_CUDA = -2
SC_ERR_UNSUPPORTED = -3
SC_ERR_NO_DEVICE = -4


class ScParams(C.Structure):
    """Mirror of ``struct sc_params``."""
    _fields_ = [
        ("sample_rate", C.c_int32), ("n_fft", C.c_int32), ("win_length", C.c_int32),
        ("hop_length", C.c_int32), ("n_mels", C.c_int32), ("n_mfcc", C.c_int32),
        ("mfcc_normalize_first", C.c_int32), ("calc_mfcc_derivative", C.c_int32),
        ("clip_output", C.c_int32), ("fft_precision", C.c_int32),
        ("pre_emphasis", C.c_double), ("mfcc_norm_factor", C.c_double),
        ("m_db_norm_factor", C.c_double), ("p_db_norm_factor", C.c_double),
        ("mean_abs_amp_norm", C.c_double),
        ("window_host", C.POINTER(C.c_double)),
    ]


class SpeechDspError(RuntimeError):
    """A CUDA / library failure that is not a bad argument."""


_P = C.c_void_p
_I64P = C.POINTER(C.c_int64)

# name -> (restype, argtypes); kept in sync with include/speechdsp.h (tests/test_abi.py checks it)
PROTOTYPES = {
    "sc_plan_create": (C.c_int, [C.POINTER(ScParams), C.POINTER(_P)]),
    "sc_plan_destroy": (None, [_P]),
    "sc_plan_reserve": (C.c_int, [_P, C.c_int64, C.c_int64, C.c_int32]),
    "sc_alloc_count": (C.c_int64, []),
    "sc_plan_poll_status": (C.c_int, [_P, C.POINTER(C.c_int32), _P]),
    "sc_plan_is_fast_path": (C.c_int, [_P]),
    "sc_num_frames": (C.c_int64, [_P, C.c_int64]),
    "sc_frontend_batch": (C.c_int, [_P, _P, _I64P, _I64P, C.c_int32, _P, _P, _P, _I64P, _P]),
    "sc_mean_abs_batch": (C.c_int, [_P, _P, _I64P, _I64P, C.c_int32, _P, _P]),
    "sc_phn_target_batch": (C.c_int, [_P, _P, _P, _I64P, _I64P, C.c_int32, C.c_int32, C.c_int32, _P, _I64P, _P]),
    "sc_window_gather": (C.c_int, [C.POINTER(_P), C.POINTER(_P), _I64P, C.c_int32, C.c_int64, _P, _P, C.c_int32,
                                   C.c_int32, _P]),
    "sc_preemphasis": (C.c_int, [_P, C.c_int64, C.c_double, _P, _P]),
    "sc_inv_preemphasis": (C.c_int, [_P, C.c_int64, C.c_double, _P, _P]),
    "sc_preemphasis_f64": (C.c_int, [_P, C.c_int64, C.c_double, _P, _P]),
    "sc_inv_preemphasis_f64": (C.c_int, [_P, C.c_int64, C.c_double, _P, _P]),
    "sc_power_to_amp_batch": (C.c_int, [_P, _P, _I64P, _I64P, C.c_int32, C.c_double, C.c_double, _P, _P]),
    "sc_griffinlim_batch": (C.c_int, [_P, _P, _P, _I64P, _I64P, C.c_int32, C.c_int32, _P, _I64P, _P, _P]),
    "sc_deemph_renorm_batch": (C.c_int, [_P, _P, _I64P, _I64P, C.c_int32, C.c_double, C.c_double, _P, _P]),
    "sc_transpose_to_f32": (C.c_int, [_P, C.c_int32, C.c_int64, C.c_int64, _P, _P]),
    "sc_griffinlim_chunk_step": (C.c_int, [_P, _P, _P, C.c_int64, C.c_int64, C.c_int64, _P, C.c_int64,
                                           C.c_int64, _P, C.c_int64, C.c_int64, _P]),
    "sc_chunk_geometry": (C.c_int, [_P, _I64P, _I64P, _I64P, _I64P]),
    "sc_griffinlim_chunk_run": (C.c_int, [_P, _P, _P, C.c_int64, C.c_int64, C.c_int64, _P, _P, C.c_int64, C.c_int64,
                                          C.c_int64, C.c_int64, C.c_int32, C.c_int64, _P]),
    "sc_p2a_chunk_partial": (C.c_int, [_P, _P, C.c_int64, C.c_double, _P, _P]),
    "sc_p2a_chunk_apply": (C.c_int, [_P, _P, C.c_int64, C.c_double, C.c_double, _P, C.c_int64, _P, _P]),
    "sc_deemph_chunk_window": (C.c_int, [C.c_double]),
    "sc_deemph_chunk_local": (C.c_int, [_P, _P, C.c_int64, C.c_int64, C.c_int64, C.c_double, _P, _P]),
    "sc_deemph_chunk_apply": (C.c_int, [_P, _P, C.c_int64, C.c_int64, C.c_int64, C.c_double, _P, C.c_int32, _P, _P, _P]),
    "sc_renorm_chunk": (C.c_int, [_P, _P, C.c_int64, _P, C.c_int64, C.c_int64, C.c_double, _P]),
    "sc_profile_enable": (C.c_int, [_P, C.c_int32]),
    "sc_profile_read": (C.c_int, [_P, C.POINTER(C.c_double)]),
    "sc_launch_count": (C.c_int64, []),
    "sc_launch_count_reset": (None, []),
    "sc_last_error": (C.c_char_p, []),
    "sc_version": (C.c_char_p, []),
}

_lib = None


def load():
    """Load the shared library (once).  Raises ImportError with build instructions if absent."""
    global _lib
    if _lib is not None:
        return _lib
    if not os.path.exists(LIB_PATH):
        raise ImportError(
            f"{LIB_PATH} is missing: the CUDA library has not been built and there is no CPU "
            "fallback.  Run `python -m speech_cloner_b200.build` (needs nvcc).")
    lib = C.CDLL(LIB_PATH)
    for name, (res, args) in PROTOTYPES.items():
        fn = getattr(lib, name)
        fn.restype = res
        fn.argtypes = args
    _lib = lib
    return lib


def check(rc: int, what: str = ""):
    """Translate a status code: bad arguments -> ValueError, the rest -> SpeechDspError."""
    if rc == SC_OK:
        return
    msg = load().sc_last_error().decode("utf-8", "replace")
    if rc == SC_ERR_INVALID:
        raise ValueError(msg or what)
    if rc == SC_ERR_UNSUPPORTED:
        raise NotImplementedError(msg or what)
    raise SpeechDspError(f"{what}: {msg} (status {rc})")


def i64_array(values):
    arr = (C.c_int64 * len(values))(*[int(v) for v in values])
    return arr
