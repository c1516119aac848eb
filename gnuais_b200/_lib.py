"""ctypes binding of ``libgaisb200.so`` (the C-ABI declared in ``include/gais_b200.h``).

The library is built in-tree by ``gnuais_b200/csrc/Makefile`` (``__graft_entry__.build()``).
There is no fallback of any kind: if the shared object is missing, or no sm_100 GPU is
visible when a context is created, the caller gets an exception.
"""
from __future__ import annotations

import ctypes as C
import os
from pathlib import Path

_HERE = Path(__file__).resolve().parent
LIB_PATH = _HERE / "lib" / "libgaisb200.so"

ABI_VERSION = 1
LAYOUT_PLANAR, LAYOUT_INTERLEAVED = 0, 1
FIR_GUARD, FIR_EXACT = 0, 1
KEEP_BITS, KEEP_SIGNS, KEEP_PEAK = 1, 2, 4
NMEA_STRIDE = 176
E_OVERFLOW = -5


class Config(C.Structure):
    _fields_ = [
        ("abi_version", C.c_int32),
        ("device", C.c_int32),
        ("n_channels", C.c_int32),
        ("layout", C.c_int32),
        ("max_frames_per_run", C.c_int64),
        ("fir_mode", C.c_int32),
        ("flags", C.c_uint32),
        ("reserved", C.c_int32 * 8),
    ]


class Msg(C.Structure):
    _fields_ = [
        ("payload", C.c_uint8 * 53),
        ("flags", C.c_uint8),
        ("nbits", C.c_uint16),
        ("channel", C.c_uint32),
        ("end_bit", C.c_uint32),
    ]


class Counters(C.Structure):
    _fields_ = [("ok", C.c_int32), ("crcfail", C.c_int32), ("sizefail", C.c_int32)]


class ChanState(C.Structure):
    _fields_ = [
        ("pll", C.c_uint32),
        ("prev", C.c_int32),
        ("lastbit", C.c_int32),
        ("fsm_state", C.c_int32),
        ("seqnr", C.c_int32),
        ("n_bits", C.c_uint32),
    ]


class NmeaRec(C.Structure):
    _fields_ = [("len", C.c_uint8), ("text", C.c_char * (NMEA_STRIDE - 1))]


class Timing(C.Structure):
    _fields_ = [
        ("total_ms", C.c_float),
        ("fir_ms", C.c_float),
        ("track_ms", C.c_float),
        ("post_ms", C.c_float),
        ("launches", C.c_int32),
        ("nmea_ms", C.c_float),
    ]


class Synth(C.Structure):
    _fields_ = [
        ("seed", C.c_uint64),
        ("amplitude", C.c_int32),
        ("noise_q16", C.c_int32),
        ("rho_q16", C.c_int32),
        ("jitter", C.c_int32),
    ]


assert C.sizeof(Msg) == 64 and C.sizeof(NmeaRec) == NMEA_STRIDE

# every symbol include/gais_b200.h declares: name -> (restype, argtypes)
_P = C.c_void_p
SYMBOLS = {
    "gais_last_error": (C.c_char_p, []),
    "gais_abi_version": (C.c_int, []),
    "gais_device_count": (C.c_int, []),
    "gais_create": (C.c_int, [C.POINTER(Config), C.POINTER(_P)]),
    "gais_destroy": (None, [_P]),
    "gais_reset": (C.c_int, [_P]),
    "gais_reset_fsm": (C.c_int, [_P]),
    "gais_run_device": (C.c_int, [_P, _P, C.c_int64, C.c_int64, _P]),
    "gais_run_host": (C.c_int, [_P, _P, C.c_int64, C.c_int64]),
    "gais_sync": (C.c_int, [_P]),
    "gais_run_bits_device": (C.c_int, [_P, _P, C.c_int64, C.c_int64, _P]),
    "gais_run_bits_host": (C.c_int, [_P, _P, C.c_int64, C.c_int64]),
    "gais_get_peaks": (C.c_int, [_P, _P]),
    "gais_message_count": (C.c_int, [_P, C.POINTER(C.c_int64)]),
    "gais_get_messages": (C.c_int, [_P, _P, C.c_int64, C.POINTER(C.c_int64)]),
    "gais_host_alloc": (C.c_int, [C.POINTER(_P), C.c_size_t]),
    "gais_host_free": (None, [_P]),
    "gais_device_messages": (C.c_int, [_P, C.POINTER(_P), C.POINTER(C.c_int64)]),
    "gais_get_nmea": (C.c_int, [_P, _P, C.c_int64, C.POINTER(C.c_int64)]),
    "gais_device_nmea": (C.c_int, [_P, C.POINTER(_P), C.POINTER(_P), C.POINTER(C.c_int64), C.POINTER(C.c_int64)]),
    "gais_get_nmea_text": (C.c_int, [_P, _P, C.c_int64, C.POINTER(C.c_int64)]),
    "gais_get_counters": (C.c_int, [_P, _P]),
    "gais_get_state": (C.c_int, [_P, _P]),
    "gais_get_totals": (C.c_int, [_P, C.POINTER(C.c_int64 * 3)]),
    "gais_bits_row_words": (C.c_int, [_P, C.POINTER(C.c_int64)]),
    "gais_get_bits": (C.c_int, [_P, _P, _P]),
    "gais_get_signs": (C.c_int, [_P, _P, C.c_int64]),
    "gais_get_timing": (C.c_int, [_P, C.POINTER(Timing)]),
    "gais_nmea_format": (C.c_int, [C.POINTER(Msg), C.c_char_p]),
    "gais_text_format": (C.c_int, [C.POINTER(Msg), C.c_char, C.c_char_p, C.c_int]),
    "gais_synth_host": (C.c_int, [C.POINTER(Synth), C.c_uint32, C.c_int32, C.c_int64, _P, C.c_int32, C.c_int64]),
    "gais_synth_device": (C.c_int, [C.POINTER(Synth), C.c_uint32, C.c_int32, C.c_int64, _P, C.c_int32, C.c_int64, _P]),
}

_lib = None


class GaisError(RuntimeError):
    def __init__(self, code: int, msg: str):
        super().__init__(f"libgaisb200 error {code}: {msg}")
        self.code = code


def load() -> C.CDLL:
    """Load the in-tree shared object and bind every declared symbol (raises if absent)."""
    global _lib
    if _lib is not None:
        return _lib
    path = Path(os.environ.get("GAIS_B200_LIB", LIB_PATH))
    if not path.exists():
        raise ImportError(
            f"{path} is missing: build it with `python -c 'import __graft_entry__ as g; g.build()'` "
            "(nvcc, sm_100a). gnuais_b200 has no CPU or PyTorch fallback."
        )
    lib = C.CDLL(str(path))
    for name, (res, args) in SYMBOLS.items():
        fn = getattr(lib, name)  # AttributeError if the .so does not export it
        fn.restype = res
        fn.argtypes = args
    if lib.gais_abi_version() != ABI_VERSION:
        raise ImportError(f"{path}: ABI version {lib.gais_abi_version()} != {ABI_VERSION}")
    _lib = lib
    return lib


def check(rc: int, allow=()) -> int:
    if rc != 0 and rc not in allow:
        raise GaisError(rc, load().gais_last_error().decode(errors="replace"))
    return rc
