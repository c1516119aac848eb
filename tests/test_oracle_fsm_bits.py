"""The oracle's HDLC bit machine against the reference's protodec_decode(), fed with BITS instead of audio
(oracle/gais_oracle.c goracle_fsm_bits vs oracle/ref_harness.c gref_fsm_bits -> src/protodec.c:988-1122):
well-formed frames with a CRC, stuffed random payloads up to and beyond the 449-bit buffer, broken stuffing,
missing closing flags, long alternations and noise reach corners of the state machine that demodulated
audio rarely visits.  Pins the FSM the tracking kernel's tables are checked against (tests/test_hdlc_table.py)."""
import ctypes as C
import os

import numpy as np
import pytest

import oracle_lib as O

FLAG = [0, 1, 1, 1, 1, 1, 1, 0]


def crc16_x25(data: bytes) -> int:
    crc = 0xFFFF
    for byte in data:
        crc ^= byte
        for _ in range(8):
            crc = (crc >> 1) ^ 0x8408 if crc & 1 else crc >> 1
    return ~crc & 0xFFFF


def make_bits(seed: int, n_bursts: int) -> np.ndarray:
    rng = np.random.default_rng(seed)
    out = []
    for burst in range(n_bursts):
        out += list(rng.integers(0, 2, rng.integers(0, 40)))
        out += [i & 1 for i in range(int(rng.integers(10, 40)))]       # sometimes too short for > 14 alternations
        out += FLAG
        ones = 0
        kind = burst % 4
        if kind == 0:                                                   # well-formed, types 1..27 in the first 6 bits
            nb = 53 if burst % 8 == 0 else 21
            body = bytearray(rng.integers(0, 256, nb, dtype=np.uint8).tobytes())
            body[0] = (int(rng.integers(0, 28)) << 2) | (body[0] & 3)
            fcs = crc16_x25(bytes(body))
            body += bytes([fcs & 0xFF, fcs >> 8])
            for byte in body:
                for k in range(8):
                    b = (byte >> k) & 1
                    out.append(b)
                    ones = ones + 1 if b else 0
                    if ones == 5:
                        out.append(0)
                        ones = 0
            out += FLAG
        else:
            for _ in range(int(rng.integers(0, 520))):
                b = int(rng.integers(0, 3) != 0)
                out.append(b)
                ones = ones + 1 if b else 0
                if ones == 5 and rng.integers(0, 8):
                    out.append(0)
                    ones = 0
            if kind != 3:
                out += FLAG
    return np.array(out, dtype=np.uint8)


@pytest.mark.skipif(not O.ref_available(), reason="reference objects not available")
@pytest.mark.parametrize("seed", [1, 2, 3])
def test_oracle_fsm_equals_reference_on_bit_streams(seed):
    bits = make_bits(seed, 1500)
    port, ref = O.port(), O.ref(tap=False, quiet=True)
    ps, rs = (C.c_int32 * 3)(), (C.c_int32 * 3)()
    nfr = C.c_int64()
    port.lib.goracle_fsm_bits.argtypes = [C.c_void_p, C.c_int64, C.c_void_p, C.c_void_p, C.c_int64, C.c_void_p]
    frames = np.zeros(20000, dtype=O.FRAME_DTYPE)
    assert port.lib.goracle_fsm_bits(bits.ctypes.data, len(bits), ps, frames.ctypes.data, len(frames), C.byref(nfr)) == 0
    ref.lib.gref_fsm_bits.argtypes = [C.c_void_p, C.c_int64, C.c_void_p, C.c_int]
    fd = os.memfd_create("nmea")
    try:
        assert ref.lib.gref_fsm_bits(bits.ctypes.data, len(bits), rs, fd) == 0
        size = os.lseek(fd, 0, os.SEEK_END)
        os.lseek(fd, 0, os.SEEK_SET)
        ref_nmea = os.read(fd, size)
    finally:
        os.close(fd)
    assert list(ps) == list(rs), (list(ps), list(rs))
    assert ps[0] > 200 and ps[1] > 200 and ps[2] > 100                  # every verdict is exercised
    assert nfr.value == sum(ps)
    # the NMEA the reference wrote for the CRC-ok frames == the oracle's armouring of the oracle's frames
    port.lib.goracle_nmea.argtypes = [C.c_void_p, C.c_int, C.c_void_p, C.c_char_p]
    seq = C.c_uint8(0)
    got = b""
    buf = C.create_string_buffer(512)
    for fr in frames[: nfr.value]:
        if fr["status"] == 0:
            pl = np.ascontiguousarray(fr["payload"])
            n = port.lib.goracle_nmea(pl.ctypes.data, int(fr["nbits"]), C.byref(seq), buf)
            got += buf.raw[:n]
    assert got == ref_nmea
