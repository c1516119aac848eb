"""CPU check of the FIR kernel's tier-1 guard band (gnuais_b200/csrc/gais_fir.cuh): for int16 windows drawn
to stress it, the 10-tap FMA sum a10 must stay within E1 = 0.36 of the reference's float32 sequential
sum R (src/filter.h:40-49: multiply, then add, tap order), so that |a10| > E1 implies sign(R) = sign(a10).
R is evaluated with numpy float32 operations (one rounding per operation, like the reference's SSE code);
the FMA chain with float64 products and sums rounded to float32 after every step.  No GPU involved."""
import struct

import numpy as np

HALF = [0x00000000, 0x00000000, 0x00000069, 0x0130bd6d, 0x0982c347, 0x112a6907, 0x18439833, 0x1ec5b74e, 0x24b00698,
        0x2a0a0629, 0x2ebea222, 0x32e7e4d5, 0x36786fe0, 0x396a68bf, 0x3bc2cc99, 0x3d8e92d5, 0x3eb7cd8a, 0x3f50b242]
TAPS = np.array([struct.unpack("<f", struct.pack("<I", h))[0] for h in HALF + HALF[::-1]], dtype=np.float32)
E1 = np.float32(0.36)


def reference_sum(win: np.ndarray) -> np.ndarray:
    """win [N, 36] int16 -> the reference's float32 result for each window"""
    s = np.zeros(len(win), dtype=np.float32)
    x = win.astype(np.float32)
    for i in range(36):
        s = (s + x[:, i] * TAPS[i]).astype(np.float32)       # float32 multiply, float32 add
    return s


def tier1_sum(win: np.ndarray) -> np.ndarray:
    """the kernel's chain: a = t13*x13, then a = fma(t_k, x_k, a) for k = 14..22 (float32 FMA)"""
    x = win.astype(np.float64)
    t = TAPS.astype(np.float64)
    a = (t[13] * x[:, 13]).astype(np.float32)
    for k in range(14, 23):
        a = (t[k] * x[:, k] + a.astype(np.float64)).astype(np.float32)
    return a


def windows(rng, n):
    full = rng.choice(np.array([-32768, 32767], dtype=np.int16), size=(n, 36))                 # full-scale square noise
    same = np.where(rng.random((n, 1)) < 0.5, np.int16(-32768), np.int16(32767)) * np.ones((1, 36), np.int16)
    noise = rng.integers(-32768, 32768, size=(n, 36)).astype(np.int16)                         # full-scale white noise
    quiet = np.clip(np.rint(rng.normal(0, 300, size=(n, 36))), -32768, 32767).astype(np.int16) # the bench's idle channel
    # centre taps cancelling (a10 near zero) under full-scale outer samples: the hardest case for the bound
    cancel = full.copy()
    cancel[:, 13:23] = rng.integers(-3, 4, size=(n, 10))
    return np.concatenate([full, same.astype(np.int16), noise, quiet, cancel])


def test_tier1_within_guard_band_of_reference():
    rng = np.random.default_rng(20261017)
    win = windows(rng, 40000)
    r, a = reference_sum(win), tier1_sum(win)
    err = np.abs(a.astype(np.float64) - r.astype(np.float64))
    assert err.max() < 0.3567, err.max()              # the bound derived in gais_fir.cuh; E1 = 0.36 sits above it
    sure = np.abs(a) > E1
    assert sure.any() and (~sure).any()               # both sides of the guard are exercised
    assert np.array_equal(a[sure] > 0, r[sure] > 0)   # outside the band the cheap sign IS the reference's sign
    assert not np.any(r[sure] == 0)
