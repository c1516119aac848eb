"""CPU model of the tensor-core FIR arithmetic (gnuais_b200/csrc/gais_fir_umma.cuh, shared by gais_fir_tc.cuh and the
fused kernel gais_fused.cuh): the exact integer Toeplitz contraction that tcgen05.mma.kind::i8 evaluates -- taps 12..23
quantised to 24 bits and split into three bytes, every sample with bit 7 flipped read as two signed bytes, three int32
accumulators D24 / D16 / D8 -- replayed in numpy against the reference's float32 sequential sum (src/filter.h:40-49).
Checks what the kernels rely on: the byte split is exact, the accumulators fit int32, |V / 65536 - R| stays below the
0.103 the header derives, and the decision rule (Q = round(V / 65536) != 0 decides; Q == 0 with |V| >= 8192 / 65536
decides; the rest stays open for tiers 2 / 3) never contradicts the reference's sign.  No GPU involved."""
import numpy as np

from test_fir_guard_bound import TAPS, reference_sum, windows

TAP_LO, TAP_HI, VGUARD = 12, 23, 8192          # U_TAP_LO, U_TAP_HI, U_VGUARD


def quantised_taps():
    T = np.zeros(36, dtype=np.int64)
    for i in range(TAP_LO, TAP_HI + 1):
        T[i] = int(float(TAPS[i]) * 16777216.0 + 0.5)          # umma_build_taps()
    return T


def contraction(win: np.ndarray):
    """win [N, 36] int16 -> (D24, D16, D8) exactly as the three accumulator slices of one MMA row hold them"""
    T = quantised_taps()
    T1, T2, T3 = (T >> 16) & 255, (T >> 8) & 255, T & 255
    xf = win.astype(np.uint16) ^ np.uint16(0x0080)
    hi = (xf >> 8).astype(np.uint8).view(np.int8).astype(np.int64)
    lo = (xf & 0xFF).astype(np.uint8).view(np.int8).astype(np.int64)
    assert np.array_equal(256 * hi + lo + 128, win.astype(np.int64))          # the flip makes both bytes signed, exactly
    d24 = (hi * T1).sum(axis=1)
    d16 = (hi * T2 + lo * T1).sum(axis=1)
    d8 = (hi * T3 + lo * T2).sum(axis=1)
    return d24, d16, d8, T


def test_integer_contraction_decides_like_the_reference():
    rng = np.random.default_rng(20261018)
    win = windows(rng, 40000)
    # more stress: the tap window alone at full scale with everything else cancelling, and ramps through zero
    extra = rng.integers(-2, 3, size=(40000, 36)).astype(np.int16)
    extra[:, 24:] = rng.choice(np.array([-32768, 32767], dtype=np.int16), size=(40000, 12))
    extra[:, :12] = rng.choice(np.array([-32768, 32767], dtype=np.int16), size=(40000, 12))
    win = np.concatenate([win, extra])
    r = reference_sum(win).astype(np.float64)
    d24, d16, d8, T = contraction(win)
    for d in (d24, d16, d8):
        assert np.abs(d).max() < 2 ** 31                                       # int32 accumulators never wrap
    sum_t = int(T.sum())
    # V in units of 2^-16: what the contraction knows of sum T_i x_i / 2^24 (the D0 = sum T3 lo' term is dropped)
    v = 65536 * d24 + 256 * d16 + d8 + sum_t // 2
    err = np.abs(v / 65536.0 - r)
    assert err.max() < 0.103, err.max()
    # the kernels' decision (umma_half_word / x_half_word): Q and, when Q == 0, the exact low half of V + 32768
    kc = sum_t // 2 + 32768
    p = 256 * d16 + d8 + kc
    q = d24 + (p >> 16)
    low = (p & 0xFFFF) - 32768
    pos = (q >= 1) | ((q == 0) & (low >= VGUARD))
    neg = (q <= -1) | ((q == 0) & (low <= -VGUARD))
    open_ = ~(pos | neg)
    assert pos.any() and neg.any() and open_.any()
    assert np.all(r[pos] > 0) and np.all(~(r[neg] > 0))
    # the open outputs are few on ordinary audio (1e-3 of the bench's idle-channel noise before the refinement)
    quiet = np.clip(np.rint(rng.normal(0, 300, size=(200000, 36))), -32768, 32767).astype(np.int16)
    d24, d16, d8, _ = contraction(quiet)
    p = 256 * d16 + d8 + kc
    q = d24 + (p >> 16)
    low = (p & 0xFFFF) - 32768
    frac_open = np.mean((q == 0) & (np.abs(low) < VGUARD))
    assert frac_open < 1e-3, frac_open


def test_no_flip_contraction_decides_like_the_reference():
    """the fused kernel's form (gais_fused.cuh, X_NOFLIP): the raw bytes multiplied twice -- high bytes as s8, low bytes as
    u8 -- so x = 256 hi + lo exactly, no 128-offset; the dropped D0 = sum T3 lo is one-sided (0 .. 780300) and its midpoint
    goes into the rounding constant"""
    rng = np.random.default_rng(20261019)
    win = windows(rng, 40000)
    T = quantised_taps()
    T1, T2, T3 = (T >> 16) & 255, (T >> 8) & 255, T & 255
    xu = win.astype(np.uint16)
    hi = (xu >> 8).astype(np.uint8).view(np.int8).astype(np.int64)
    lo = (xu & 0xFF).astype(np.int64)
    assert np.array_equal(256 * hi + lo, win.astype(np.int64))
    d24 = (hi * T1).sum(axis=1)
    d16 = (hi * T2).sum(axis=1) + (lo * T1).sum(axis=1)
    d8 = (hi * T3).sum(axis=1) + (lo * T2).sum(axis=1)
    d0 = (lo * T3).sum(axis=1)
    for d in (d24, d16, d8):
        assert np.abs(d).max() < 2 ** 31
    assert d0.min() >= 0 and d0.max() <= 780300
    r = reference_sum(win).astype(np.float64)
    mid = 780300 // 512                                                       # X_KC_NOFLIP - 32768
    v = 65536 * d24 + 256 * d16 + d8 + mid
    err = np.abs(v / 65536.0 - r)
    assert err.max() < 0.1151, err.max()
    kc = 32768 + mid
    p = 256 * d16 + d8 + kc
    q = d24 + (p >> 16)
    low = (p & 0xFFFF) - 32768
    pos = (q >= 1) | ((q == 0) & (low >= VGUARD))
    neg = (q <= -1) | ((q == 0) & (low <= -VGUARD))
    assert pos.any() and neg.any() and (~(pos | neg)).any()
    assert np.all(r[pos] > 0) and np.all(~(r[neg] > 0))
