"""Generates tests/golden/reference_outputs.npz by running the UNMODIFIED reference objects
(oracle/_ref, built by oracle/Makefile from /root/reference) on the inputs of tests/cases.py.

Run in the build container only (needs /root/reference):  python tests/golden/make_golden.py
The GPU box has no /root/reference; there the tests compare against this file."""
import hashlib
import sys
from pathlib import Path

import numpy as np

HERE = Path(__file__).resolve().parent
sys.path.insert(0, str(HERE.parent.parent))
sys.path.insert(0, str(HERE.parent))

import cases  # noqa: E402
import oracle_lib as O  # noqa: E402


def main():
    ref = O.ref(tap=True)
    out = {}
    for name, (x, num_ch) in cases.golden_inputs().items():
        out[f"{name}/sha256"] = np.frombuffer(hashlib.sha256(np.ascontiguousarray(x).tobytes()).digest(), np.uint8)
        for ch in range(num_ch):
            r = ref.run(x, num_ch=num_ch, ch_ofs=ch, chunk=1020, want_bits=True, want_signs=True)
            k = f"{name}/ch{ch}"
            out[k + "/bits"] = np.packbits(r.bits, bitorder="little")
            out[k + "/n_bits"] = np.array(len(r.bits), np.int64)
            out[k + "/signs"] = np.packbits(r.signs, bitorder="little")
            out[k + "/nmea"] = np.frombuffer(r.nmea, np.uint8)
            out[k + "/stats"] = np.array([r.ok, r.crcfail, r.sizefail, r.pll, r.prev, r.lastbit, r.fsm_state, r.seqnr], np.int64)
            print(f"{k}: ok={r.ok} crcfail={r.crcfail} sizefail={r.sizefail} bits={len(r.bits)} nmea={len(r.nmea)}B pll={r.pll}")
    np.savez_compressed(HERE / "reference_outputs.npz", **out)
    print("wrote", HERE / "reference_outputs.npz")


if __name__ == "__main__":
    main()
