"""CPU check of the arithmetic the tracking kernel is built on (gnuais_b200/csrc/gais_track.cuh), against the
oracle: a line-by-line Python mirror of the kernel's word loop -- 32-bit word-base phase register (phase in
the lower half, slices since the hand-over in the upper half, nudges added to the base), NRZI bits as toggles
of a difference accumulator, hand-over of 24+ bits at a time -- must give the reference's NRZI bit stream
and final DPLL phase for the FIR signs the oracle reports.  No GPU involved."""
import numpy as np
import pytest

import cases
import oracle_lib as O

INC, NUDGE, M32 = 13107, 819, 0xFFFFFFFF


def track_words(signs: np.ndarray):
    """signs: uint8 0/1 per sample (filtered > 0).  Returns (NRZI bits, final pll, final prev)."""
    n = len(signs)
    pad = (-n) % 32
    words = np.packbits(np.concatenate([signs, np.zeros(pad, np.uint8)]), bitorder="little").view("<u4")
    zb, dlo, hb, prevword = 0, 0, 0, 0          # pll = 0, prev = 0, nothing pending (src/receiver.c:52-74)
    bits = []

    def hand_over(k):
        nonlocal dlo, hb, zb
        w = ~dlo & M32
        bits.extend((w >> i) & 1 for i in range(k))
        dlo >>= k
        hb += k
        zb = (zb - (k << 16)) & M32

    for wi, sw in enumerate(int(w) for w in words):
        nb = min(32, n - 32 * wi)
        x = (sw ^ (((sw << 1) | (prevword >> 31)) & M32)) & ((1 << nb) - 1)
        prevword = (sw << (32 - nb)) & M32
        while x:
            iso = x & -x
            j = iso.bit_length() - 1
            x ^= iso
            zj = (j * INC + zb) & M32                  # the register at sample j of this word
            dlo ^= 1 << (zj >> 16)                     # the bit of the NEXT slice flips
            zb = (zb + (-NUDGE if zj & 0x8000 else NUDGE)) & M32
        zb = (zb + nb * INC) & M32                     # base of the next word
        nd = zb >> 16
        assert nd <= 30 and dlo < (1 << 31)
        if nd >= 24:
            hand_over(nd & ~3)                         # whole nibbles, as hdlc_chunk() consumes them
    hand_over(zb >> 16)                                # end of the run: everything that was sliced
    return np.array(bits, dtype=np.uint8), zb & 0xFFFF, prevword >> 31


@pytest.mark.parametrize("seed,n_frames,sigma", [(11, 20000, 300.0), (12, 33333, 1500.0), (13, 4097, 3000.0), (14, 31, 300.0)])
def test_word_base_dpll_equals_oracle(seed, n_frames, sigma):
    x = cases.synth_case(seed, 1, n_frames, sigma=sigma)[:, 0]
    want = O.port().run(np.ascontiguousarray(x), want_signs=True)
    bits, pll, _ = track_words(want.signs)
    assert np.array_equal(bits, want.bits)
    assert pll == want.pll


def test_word_base_dpll_on_edge_inputs():
    for name, x in cases.edge_cases().items():
        want = O.port().run(np.ascontiguousarray(x), want_signs=True)
        bits, pll, _ = track_words(want.signs)
        assert np.array_equal(bits, want.bits), name
        assert pll == want.pll, name
