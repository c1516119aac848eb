"""GPU tests (-m gpu) of the round-2 additions to the batched C-ABI:
  * gais_run_bits_*  = protodec_decode() for a batch of channels, fed NRZI bits (src/protodec.c:988-1122)
  * size-fails are counted by the tracker and never take a slot: a bit stream that closes a frame every ~32
    bits must give the reference's lostframes2 count, not GAIS_EOVERFLOW
  * a run that does overflow its slots still reports the reference's counters
  * gais_get_peaks   = filter_run_buf()'s return value (src/filter.c:112-119): positive samples only
  * the three FIR implementations (tensor-core with TMA ring, first tensor-core kernel, FFMA2 guard band) give the
    same sign words as the exact chain
Nothing here reads /root/reference; the reference objects used are the prebuilt oracle/_ref ones."""
import ctypes as C
import os
import subprocess
import sys
from pathlib import Path

import numpy as np
import pytest

import cases
import oracle_lib as O
from gnuais_b200 import BatchReceiver, GaisError, SynthParams, synth_host
from test_oracle_fsm_bits import FLAG, make_bits

pytestmark = pytest.mark.gpu
ROOT = Path(__file__).resolve().parent.parent


def oracle_bits(bits: np.ndarray):
    """(ok, crcfail, sizefail), NMEA bytes of the oracle's bit machine (pinned to protodec_decode() by
    tests/test_oracle_fsm_bits.py)."""
    port = O.port()
    ps, nfr = (C.c_int32 * 3)(), C.c_int64()
    port.lib.goracle_fsm_bits.argtypes = [C.c_void_p, C.c_int64, C.c_void_p, C.c_void_p, C.c_int64, C.c_void_p]
    frames = np.zeros(len(bits) // 20 + 64, dtype=O.FRAME_DTYPE)
    assert port.lib.goracle_fsm_bits(bits.ctypes.data, len(bits), ps, frames.ctypes.data, len(frames), C.byref(nfr)) == 0
    assert nfr.value <= len(frames)
    port.lib.goracle_nmea.argtypes = [C.c_void_p, C.c_int, C.c_void_p, C.c_char_p]
    seq, text, buf = C.c_uint8(0), b"", C.create_string_buffer(512)
    for fr in frames[: nfr.value]:
        if fr["status"] == 0:
            pl = np.ascontiguousarray(fr["payload"])
            n = port.lib.goracle_nmea(pl.ctypes.data, int(fr["nbits"]), C.byref(seq), buf)
            text += buf.raw[:n]
    return tuple(ps), text


@pytest.mark.parametrize("device_path", [False, True])
def test_run_bits_equals_protodec_decode(device_path):
    """four channels with different streams, fed in two calls of different lengths (FSM state carries over)"""
    streams = [make_bits(seed, 400) for seed in (11, 12, 13, 14)]
    n = min(len(s) for s in streams)
    bits = np.stack([s[:n] for s in streams])
    cut = n // 3 + 5
    with BatchReceiver(4, n, slot_cap=n // 54 + 2) as rx:
        text = [b""] * 4
        for part in (bits[:, :cut], bits[:, cut:]):
            part = np.ascontiguousarray(part)
            if device_path:
                import torch
                rx.run_bits(torch.from_numpy(part).cuda())
            else:
                rx.run_bits(part)
            msgs, recs = rx.messages(), rx.nmea_records()
            raw = recs.view(np.uint8).reshape(-1, 176)
            for i, m in enumerate(msgs):
                text[int(m["channel"])] += raw[i, 1:1 + raw[i, 0]].tobytes()
        cnt = rx.counters()
        st = rx.state()
    for c in range(4):
        want_cnt, want_text = oracle_bits(np.ascontiguousarray(bits[c]))
        assert (int(cnt[c]["ok"]), int(cnt[c]["crcfail"]), int(cnt[c]["sizefail"])) == want_cnt, c
        assert text[c] == want_text, c
        assert int(st[c]["n_bits"]) == n
        assert want_cnt[0] > 20 and want_cnt[1] > 20 and want_cnt[2] > 5


def test_dense_size_fail_stream_counts_instead_of_overflowing():
    """preamble + flag + a few data bits + closing flag, over and over: one lostframes2 every ~50 bits, none of
    which may take a message slot (default capacity: one per 1024 samples)"""
    unit = [i & 1 for i in range(16)] + FLAG + [1, 0, 1, 1, 0, 0, 1, 0] + FLAG
    bits = np.array(unit * 4000, dtype=np.uint8)
    want_cnt, want_text = oracle_bits(bits)
    assert want_cnt[2] > 3500 and want_cnt[0] == 0
    with BatchReceiver(2, len(bits)) as rx:
        rx.run_bits(np.stack([bits, bits[::-1].copy()]))
        cnt = rx.counters()
        assert rx.message_count() == 0
    assert (int(cnt[0]["ok"]), int(cnt[0]["crcfail"]), int(cnt[0]["sizefail"])) == want_cnt
    want_rev, _ = oracle_bits(np.ascontiguousarray(bits[::-1]))
    assert (int(cnt[1]["ok"]), int(cnt[1]["crcfail"]), int(cnt[1]["sizefail"])) == want_rev


def test_counters_survive_slot_overflow():
    """with far too few slots the run reports GAIS_EOVERFLOW -- and still the reference's counters"""
    x = cases.synth_case(109, 1, 120000, rho=1.0)
    want = O.port().run(np.ascontiguousarray(x[:, 0]))
    assert want.ok > 20
    rx = BatchReceiver(1, 120000, slot_cap=3)
    rx.run(np.ascontiguousarray(x.T))
    with pytest.raises(GaisError) as e:
        rx.sync()
    assert e.value.code == -5
    cnt = rx.counters()[0]
    assert (int(cnt["ok"]), int(cnt["crcfail"]), int(cnt["sizefail"])) == want.counters()
    assert rx.message_count() == 3
    rx.close()


def test_peaks_are_filter_run_buf_return_values():
    import torch
    p = SynthParams(seed=31, sigma=300.0, rho=0.6)
    n = 50000
    x = synth_host(p, 6, n)
    x[1] = -np.abs(x[1]) - 1            # nothing positive: the reference reports 0
    x[2, :] = 0
    x[3, 777] = 32767
    x[4] = np.minimum(x[4], 5)
    want = np.maximum(x.max(axis=1), 0).astype(np.int16)
    for layout in ("planar", "interleaved"):
        with BatchReceiver(6, n, layout=layout, keep_peak=True, tile_frames=16384) as rx:
            buf = x if layout == "planar" else np.ascontiguousarray(x.T)
            rx.run(torch.from_numpy(buf).cuda())
            assert np.array_equal(rx.peaks(), want), layout
            rx.run(torch.from_numpy(np.ascontiguousarray(buf[:, :1000] if layout == "planar" else buf[:1000])).cuda())
            first = x[:, :1000]
            assert np.array_equal(rx.peaks(), np.maximum(first.max(axis=1), 0).astype(np.int16))     # per run, not cumulative
    if O.ref_available():
        ref = O.ref(tap=True, quiet=True)
        ref.lib.gref_peak.argtypes = [C.c_void_p, C.c_int64, C.c_int, C.c_int, C.c_int]
        for c in range(6):
            row = np.ascontiguousarray(x[c])
            assert ref.lib.gref_peak(row.ctypes.data, n, 1, 0, 1020) == int(want[c]), c


def test_every_fir_implementation_gives_the_exact_signs():
    """GAIS_FIR_IMPL selects the fast kernel when the library is loaded, so each runs in its own process"""
    code = r'''
import sys, numpy as np, torch
sys.path.insert(0, %r); sys.path.insert(0, %r)
from gnuais_b200 import BatchReceiver, SynthParams, synth_device
C_, N = 1024, 70000 + 512
d = torch.empty((C_, N), dtype=torch.int16, device="cuda")
synth_device(SynthParams(seed=77, sigma=300.0), d[:512], 512, N)
synth_device(SynthParams(seed=78, sigma=1500.0, rho=0.9), d[512:960], 448, N, first_channel=512)
d[960:976] = 0
d[976:992] = torch.randint(-2, 3, (16, N), device="cuda", dtype=torch.int16)
d[992:] = torch.randint(-32768, 32768, (32, N), device="cuda", dtype=torch.int32).to(torch.int16)
out = []
for mode, tile in (("exact", 0), ("guard", 0), ("guard", 16384)):
    rx = BatchReceiver(C_, N, fir_mode=mode, keep_signs=True, tile_frames=tile)
    rx.run(d[:, :30000]); a = rx.signs(30000); m1 = rx.messages()
    rx.run(d[:, 30000:]); b = rx.signs(N - 30000); m2 = rx.messages()
    out.append((a, b, m1, m2, rx.counters(), rx.state()))
    rx.close()
for o in out[1:]:
    for x, y in zip(out[0], o):
        assert x.tobytes() == y.tobytes()
assert out[0][4]["ok"].sum() > 1000
print("same")
''' % (str(ROOT), str(ROOT / "tests"))
    for impl in ("tc", "umma", "ffma2"):
        env = dict(os.environ, GAIS_FIR_IMPL=impl)
        r = subprocess.run([sys.executable, "-c", code], env=env, capture_output=True, text=True, timeout=600)
        assert r.returncode == 0 and "same" in r.stdout, (impl, r.stdout[-2000:], r.stderr[-2000:])


def test_packed_nmea_equals_fixed_records_and_host_formatter():
    """gais_device_nmea / gais_get_nmea_text (one warp per message, packed) against the per-message formatter the host
    shim uses and the fixed-stride records -- incl. two-sentence type 5, ungated types 25-27 and odd lengths"""
    import torch
    from gnuais_b200 import nmea_format
    C_, N = 512, 96000
    d = torch.empty((C_, N), dtype=torch.int16, device="cuda")
    from gnuais_b200 import synth_device
    synth_device(SynthParams(seed=55, sigma=300.0, rho=0.9), d, C_, N)
    with BatchReceiver(C_, N) as rx:
        rx.run(d)
        msgs, recs = rx.messages(), rx.nmea_records()
        text = rx.nmea()
        tptr, optr, n, nbytes = rx.device_nmea()
        tm = rx.timing()
    raw = recs.view(np.uint8).reshape(-1, 176)
    want = b"".join(raw[i, 1:1 + raw[i, 0]].tobytes() for i in range(len(recs)))
    assert n == len(msgs) > 20000 and nbytes == len(text)
    assert text == want
    two = sum(1 for m in msgs[:4000] if nmea_format(m).count(b"\r\n") == 2)
    none = sum(1 for m in msgs[:4000] if not (m["flags"] & 16))
    assert two > 50 and none > 10
    assert b"".join(nmea_format(m) for m in msgs[:4000]) == text[: sum(int(x) for x in raw[:4000, 0])]
    assert tm["nmea_ms"] > 0
    # empty run: no text
    with BatchReceiver(16, 4096) as rx:
        rx.run(torch.zeros((16, 4096), dtype=torch.int16, device="cuda"))
        assert rx.nmea() == b""
