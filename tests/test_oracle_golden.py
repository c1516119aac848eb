"""The oracle restatement against the committed reference-generated golden vectors
(tests/golden/reference_outputs.npz, made by tests/golden/make_golden.py from the unmodified
reference).  Runs everywhere, including the GPU box where /root/reference does not exist."""
import hashlib
from pathlib import Path

import numpy as np
import pytest

import cases
import oracle_lib as O

GOLD = np.load(Path(__file__).parent / "golden" / "reference_outputs.npz")
INPUTS = cases.golden_inputs()


@pytest.mark.parametrize("name", sorted(INPUTS))
def test_port_matches_golden(name):
    x, num_ch = INPUTS[name]
    # the generator is integer-only: the inputs must be reproduced exactly on any host
    assert hashlib.sha256(np.ascontiguousarray(x).tobytes()).digest() == GOLD[f"{name}/sha256"].tobytes()
    for ch in range(num_ch):
        p = O.port().run(x, num_ch=num_ch, ch_ofs=ch, want_signs=True, want_frames=True)
        k = f"{name}/ch{ch}"
        assert len(p.bits) == int(GOLD[k + "/n_bits"])
        assert np.array_equal(np.packbits(p.bits, bitorder="little"), GOLD[k + "/bits"])
        assert np.array_equal(np.packbits(p.signs, bitorder="little"), GOLD[k + "/signs"])
        assert p.nmea == GOLD[k + "/nmea"].tobytes()
        assert [p.ok, p.crcfail, p.sizefail, p.pll, p.prev, p.lastbit, p.fsm_state, p.seqnr] == GOLD[k + "/stats"].tolist()
        assert len(p.frames) == p.ok + p.crcfail + p.sizefail


def test_golden_covers_all_frame_outcomes():
    tot = np.zeros(3, np.int64)
    multi = 0
    for k in GOLD.files:
        if k.endswith("/stats"):
            tot += GOLD[k][:3]
        if k.endswith("/nmea"):
            multi += GOLD[k].tobytes().count(b"!AIVDM,2,")
    assert (tot > 0).all() and multi > 0     # ok, CRC-fail, size-fail and multi-sentence all exercised


def test_port_chunk_and_stereo_invariance():
    x = cases.synth_case(31, 2, 60000)
    a = O.port().run(x, num_ch=2, ch_ofs=1, chunk=1020)
    b = O.port().run(x, num_ch=2, ch_ofs=1, chunk=7)
    c = O.port().run(np.ascontiguousarray(x[:, 1]), chunk=4096)
    assert a.nmea == b.nmea == c.nmea and np.array_equal(a.bits, b.bits) and np.array_equal(a.bits, c.bits)


def test_crc_known_answer():
    """CRC-16/X.25 check value: "123456789" -> 0x906E; a frame followed by its FCS leaves 0x0f47."""
    import ctypes as C
    lib = O.port().lib
    lib.goracle_crc16.restype = C.c_uint16
    lib.goracle_crc16.argtypes = [C.c_char_p, C.c_uint]
    assert lib.goracle_crc16(b"123456789", 9) == 0x906E
    fcs = lib.goracle_crc16(b"123456789", 9)
    framed = b"123456789" + bytes([fcs & 0xFF, fcs >> 8])
    assert lib.goracle_crc16(framed, 11) == 0x0F47


def test_tap_bit_patterns():
    import ctypes as C
    lib = O.port().lib
    lib.goracle_tap_bits.restype = C.POINTER(C.c_uint32)
    t = [lib.goracle_tap_bits()[i] for i in range(36)]
    assert t[0] == t[1] == t[34] == t[35] == 0 and t[2] == t[33] == 0x69 and t[17] == t[18] == 0x3F50B242
    assert t == t[::-1]
