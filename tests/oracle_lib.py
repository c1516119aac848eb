"""ctypes access to the CHECKERS under oracle/ (test infrastructure only).

* ``port``  -- oracle/_build/libgais_oracle.so, this repo's C restatement (always buildable)
* ``ref``   -- oracle/_ref/libgnuais_ref{,_tap}.so, the unmodified reference objects + harness
               (built from /root/reference when present; prebuilt files travel to the GPU box)
"""
from __future__ import annotations

import ctypes as C
import subprocess
from dataclasses import dataclass
from pathlib import Path

import numpy as np

ROOT = Path(__file__).resolve().parent.parent
ORACLE = ROOT / "oracle"

FRAME_DTYPE = np.dtype([("end_bit", "<u4"), ("nbits", "<i2"), ("status", "u1"), ("nbytes", "u1"), ("payload", "u1", 56)])
assert FRAME_DTYPE.itemsize == 64


def build_oracle(ref: bool = True) -> None:
    subprocess.run(["make", "-C", str(ORACLE), "port"] + (["ref"] if ref else []), check=True,
                   stdout=subprocess.DEVNULL)


@dataclass
class OracleResult:
    bits: np.ndarray | None
    signs: np.ndarray | None
    nmea: bytes
    ok: int
    crcfail: int
    sizefail: int
    pll: int
    prev: int
    lastbit: int
    fsm_state: int
    seqnr: int
    frames: np.ndarray | None = None

    def counters(self):
        return (self.ok, self.crcfail, self.sizefail)


_P = C.c_void_p
_RUN_COMMON = [_P, C.c_int64, C.c_int, C.c_int, C.c_int, _P, C.c_int64, C.POINTER(C.c_int64), _P,
               _P, C.c_int64, C.POINTER(C.c_int64), _P]


class _Checker:
    def __init__(self, path: Path, prefix: str, has_frames: bool):
        self.lib = C.CDLL(str(path))
        self.run_fn = getattr(self.lib, prefix + "_run")
        self.run_fn.restype = C.c_int
        self.run_fn.argtypes = _RUN_COMMON + ([_P, C.c_int64, C.POINTER(C.c_int64)] if has_frames else [])
        self.bench_fn = getattr(self.lib, prefix + "_bench")
        self.bench_fn.restype = C.c_double
        self.bench_fn.argtypes = [_P, C.c_int64, C.c_int64, C.c_int, C.c_int, C.POINTER(C.c_int64)]
        self.has_frames = has_frames

    def run(self, buf: np.ndarray, num_ch: int = 1, ch_ofs: int = 0, chunk: int = 1020, want_bits: bool = True,
            want_signs: bool = False, want_frames: bool = False) -> OracleResult:
        """buf: int16, frame-interleaved [n_frames, num_ch] (or 1-D for num_ch == 1)."""
        a = np.ascontiguousarray(buf, dtype=np.int16).reshape(-1)
        n_frames = a.size // num_ch
        bits_cap = n_frames // 4 + 16
        bits = np.zeros(bits_cap, dtype=np.uint8) if want_bits else None
        signs = np.zeros(n_frames, dtype=np.uint8) if want_signs else None
        nmea_cap = (n_frames // 250 + 4) * 100
        nmea = np.zeros(nmea_cap, dtype=np.uint8)
        n_bits, nmea_len, nfr = C.c_int64(), C.c_int64(), C.c_int64()
        stats = np.zeros(8, dtype=np.int32)
        args = [a.ctypes.data_as(_P), n_frames, num_ch, ch_ofs, chunk,
                bits.ctypes.data_as(_P) if want_bits else None, bits_cap, C.byref(n_bits),
                signs.ctypes.data_as(_P) if want_signs else None,
                nmea.ctypes.data_as(_P), nmea_cap, C.byref(nmea_len), stats.ctypes.data_as(_P)]
        frames = None
        if self.has_frames:
            fcap = n_frames // 200 + 16 if want_frames else 0
            frames = np.zeros(fcap, dtype=FRAME_DTYPE) if want_frames else None
            args += [frames.ctypes.data_as(_P) if want_frames else None, fcap, C.byref(nfr)]
        rc = self.run_fn(*args)
        if rc != 0:
            raise RuntimeError(f"oracle run failed: {rc}")
        assert nmea_len.value <= nmea_cap and (not want_bits or n_bits.value <= bits_cap)
        if frames is not None:
            assert nfr.value <= len(frames)
            frames = frames[: nfr.value]
        return OracleResult(
            bits=bits[: n_bits.value] if want_bits else None, signs=signs, nmea=nmea[: nmea_len.value].tobytes(),
            ok=int(stats[0]), crcfail=int(stats[1]), sizefail=int(stats[2]), pll=int(stats[3]) & 0xFFFFFFFF,
            prev=int(stats[4]), lastbit=int(stats[5]), fsm_state=int(stats[6]), seqnr=int(stats[7]), frames=frames)

    def bench(self, planar: np.ndarray, n_threads: int, chunk: int = 1020):
        """planar [n_channels, n_samples] int16 -> (seconds, ok_total)."""
        a = np.ascontiguousarray(planar, dtype=np.int16)
        ok = C.c_int64()
        secs = self.bench_fn(a.ctypes.data_as(_P), a.shape[0], a.shape[1], n_threads, chunk, C.byref(ok))
        return secs, ok.value


_cache: dict = {}


def port() -> _Checker:
    if "port" not in _cache:
        p = ORACLE / "_build" / "libgais_oracle.so"
        if not p.exists():
            build_oracle(ref=False)
        _cache["port"] = _Checker(p, "goracle", True)
    return _cache["port"]


def ref_available() -> bool:
    return (ORACLE / "_ref" / "libgnuais_ref_tap.so").exists() or Path("/root/reference/src/receiver.c").exists()


def ref(tap: bool = True, quiet: bool = True) -> _Checker:
    """quiet=True (tests): skip_type[] gates the per-message printf + field decoders and hlog is
    raised to LOG_ERR; quiet=False (CPU baseline): the reference's default behaviour."""
    key = "ref_tap" if tap else "ref"
    if key not in _cache:
        p = ORACLE / "_ref" / ("libgnuais_ref_tap.so" if tap else "libgnuais_ref.so")
        if not p.exists():
            build_oracle(ref=True)
        chk = _Checker(p, "gref", False)
        chk.lib.gref_set_quiet.argtypes = [C.c_int]
        _cache[key] = chk
    _cache[key].lib.gref_set_quiet(1 if quiet else 0)
    return _cache[key]
