"""Generates tests/golden/text_lines.npz: the text line the UNMODIFIED reference prints
(protodec_getdata(), src/protodec.c:896-986, via oracle/_ref) for random CRC-ok frames of every
message type 0..27 and many lengths.  Build-container only (needs /root/reference)."""
import ctypes as C
import sys
from pathlib import Path

import numpy as np

HERE = Path(__file__).resolve().parent
sys.path.insert(0, str(HERE.parent.parent))
sys.path.insert(0, str(HERE.parent))
import oracle_lib as O  # noqa: E402


def records(n=3000, seed=2718):
    rng = np.random.default_rng(seed)
    pay = rng.integers(0, 256, size=(n, 53), dtype=np.uint8)
    nbits = rng.integers(1, 427, size=n).astype(np.int32)
    nbits[: n // 2] = rng.choice([168, 424, 256, 312, 96, 160, 360, 368, 72, 200], size=n // 2)
    typ = rng.integers(0, 28, size=n)
    pay[:, 0] = (typ << 2) | rng.integers(0, 4, size=n)
    # every fourth binary message (types 6/8) gets DAC 1 and FI 11 or 40 so the IFM decoders run
    for i in range(n):
        if typ[i] == 6 and i % 2 == 0:      # appid at bit 72: dac 10 bits, fi 6 bits
            v = (1 << 6) | (11 if i % 4 == 0 else 40)
            pay[i, 9] = v >> 8; pay[i, 10] = v & 0xFF
        if typ[i] == 8 and i % 2 == 0:      # appid at bit 40
            v = (1 << 6) | (11 if i % 4 == 0 else 40)
            pay[i, 5] = v >> 8; pay[i, 6] = v & 0xFF
        nb = nbits[i] // 8
        pay[i, nb:] = 0
    seq = rng.integers(0, 10, size=n).astype(np.int32)
    return pay, nbits, seq


def reference_lines(pay, nbits, seq):
    lib = O.ref(tap=True, quiet=False).lib
    lib.gref_getdata_text.restype = C.c_int64
    lib.gref_getdata_text.argtypes = [C.c_void_p, C.c_int, C.c_int, C.c_char, C.c_char_p, C.c_int64]
    buf = C.create_string_buffer(4096)
    out = []
    for i in range(len(nbits)):
        row = np.ascontiguousarray(pay[i])
        n = lib.gref_getdata_text(row.ctypes.data_as(C.c_void_p), int(nbits[i]), int(seq[i]), b"AB"[i % 2:i % 2 + 1], buf, 4096)
        out.append(buf.raw[:n])
    O.ref(tap=True, quiet=True)
    return out


if __name__ == "__main__":
    pay, nbits, seq = records()
    lines = reference_lines(pay, nbits, seq)
    blob = b"".join(lines)
    lens = np.array([len(x) for x in lines], np.int32)
    np.savez_compressed(HERE / "text_lines.npz", payload=pay, nbits=nbits, seqnr=seq, lens=lens, text=np.frombuffer(blob, np.uint8))
    print("wrote", len(lines), "records,", (lens > 0).sum(), "with text,", len(blob), "bytes")
