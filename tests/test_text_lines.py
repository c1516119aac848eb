"""the per-message stdout line (field decoders) restated in
gnuais_b200/csrc/gais_text.cpp, against reference-generated golden lines for every message type
(tests/golden/text_lines.npz) and, where the reference objects exist, against the reference live."""
from pathlib import Path

import numpy as np
import pytest

import oracle_lib as O
from gnuais_b200 import MSG_DTYPE, text_format

GOLD = np.load(Path(__file__).parent / "golden" / "text_lines.npz")


def _msg(pay, nbits, seq):
    m = np.zeros((), dtype=MSG_DTYPE)
    m["payload"] = pay
    m["nbits"] = nbits
    typ = int(pay[0]) >> 2
    m["flags"] = int(seq) | (16 if 1 <= typ <= 24 else 0)
    return m


def test_text_lines_match_golden():
    off = np.concatenate([[0], np.cumsum(GOLD["lens"])])
    text = GOLD["text"].tobytes()
    types = set()
    for i in range(len(GOLD["nbits"])):
        want = text[off[i]:off[i + 1]]
        got = text_format(_msg(GOLD["payload"][i], GOLD["nbits"][i], GOLD["seqnr"][i]), "AB"[i % 2])
        assert got == want, (i, got, want)
        if want:
            types.add(int(want.split()[3]))
    assert types == set(range(1, 25))          # every gated type 1..24 produced a line


@pytest.mark.skipif(not O.ref_available(), reason="reference objects not available")
def test_text_lines_match_reference_live():
    import sys
    sys.path.insert(0, str(Path(__file__).parent / "golden"))
    import make_text_golden as G
    pay, nbits, seq = G.records(n=600, seed=99)
    want = G.reference_lines(pay, nbits, seq)
    for i in range(len(nbits)):
        assert text_format(_msg(pay[i], nbits[i], seq[i]), "AB"[i % 2]) == want[i], i
