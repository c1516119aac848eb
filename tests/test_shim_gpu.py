"""The link-compatible legacy entry points (include/gais_compat.h, libgnuais_rx_b200.so) driven
by a C program that mimics gnuais' main() loop (tests/c/shim_main.c), against the oracle."""
import subprocess
from pathlib import Path

import numpy as np
import pytest

import cases
import oracle_lib as O
from gnuais_b200 import MSG_DTYPE, text_format

ROOT = Path(__file__).resolve().parent.parent
LIBDIR = ROOT / "gnuais_b200" / "lib"


def _build(tmp_path: Path) -> Path:
    exe = tmp_path / "shim_main"
    subprocess.run(["gcc", "-O2", "-Wall", "-I", str(ROOT / "include"), str(ROOT / "tests" / "c" / "shim_main.c"), "-o", str(exe),
                    "-L", str(LIBDIR), "-lgnuais_rx_b200", "-lgaisb200", f"-Wl,-rpath,{LIBDIR}"], check=True)
    return exe


def _build_protodec(tmp_path: Path) -> Path:
    exe = tmp_path / "protodec_main"
    subprocess.run(["gcc", "-O2", "-Wall", "-I", str(ROOT / "include"), str(ROOT / "tests" / "c" / "protodec_main.c"), "-o", str(exe),
                    "-L", str(LIBDIR), "-lgnuais_rx_b200", "-lgaisb200", f"-Wl,-rpath,{LIBDIR}"], check=True)
    return exe


def test_shim_links_without_gpu(tmp_path):
    """CPU: the shim library resolves against the batched library and exports the three entry points"""
    _build(tmp_path)
    out = subprocess.run(["nm", "-D", "--defined-only", str(LIBDIR / "libgnuais_rx_b200.so")], capture_output=True, text=True).stdout
    for sym in ("init_receiver", "receiver_run", "free_receiver", "gais_compat_flush", "protodec_initialize", "protodec_reset",
                "protodec_decode", "protodec_getdata", "gais_compat_flush_decoder", "gais_compat_free_decoder"):
        assert f" T {sym}" in out
    _build_protodec(tmp_path)


@pytest.mark.gpu
@pytest.mark.parametrize("channels", [1, 2])
def test_shim_matches_reference_loop(tmp_path, channels):
    n = 240000
    x = cases.synth_case(301, channels, n, rho=0.8)            # [n, channels] interleaved = the raw file gnuais reads
    raw = tmp_path / "capture.raw"
    x.tofile(raw)
    exe = _build(tmp_path)
    r = subprocess.run([str(exe), str(raw), str(channels), str(tmp_path / "out")], capture_output=True, text=True, timeout=300,
                       env={"GAIS_SHIM_BATCH_FRAMES": "48000", "PATH": "/usr/bin:/bin"})
    assert r.returncode == 0, r.stderr
    lines = r.stdout.strip().splitlines()
    stats = [ln for ln in lines if ": Received correctly:" in ln]
    for c in range(channels):
        want = O.port().run(x, num_ch=channels, ch_ofs=c, want_frames=True)
        got = (tmp_path / f"out.{'AB'[c]}.nmea").read_bytes()
        assert got == want.nmea
        assert stats[c] == (f"{'AB'[c]}: Received correctly: {want.ok} packets, wrong CRC: {want.crcfail} packets, "
                            f"wrong size: {want.sizefail} packets")
        assert want.ok > 50
        # the per-message stdout lines (src/protodec.c:934-985), in order, for this channel
        text = [ln for ln in lines if ln.startswith(f"ch {'AB'[c]} type ")]
        seq, expect = 0, []
        for f in want.frames[want.frames["status"] == 0]:
            m = np.zeros((), dtype=MSG_DTYPE)
            nb = int(f["nbytes"])
            m["payload"][:nb] = f["payload"][:nb]
            m["nbits"] = f["nbits"]
            gate = 1 <= (int(f["payload"][0]) >> 2) <= 24
            m["flags"] = seq | (16 if gate else 0)
            if gate:
                seq = (seq + 1) % 10
                expect.append(text_format(m, "AB"[c]).decode().rstrip("\n"))
        assert text == expect


@pytest.mark.gpu
@pytest.mark.parametrize("reset_at", [-1, 20011])
def test_protodec_entry_points_match_reference(tmp_path, reset_at):
    """protodec_initialize / protodec_decode / protodec_reset of the shim (src/protodec.h:73-76) on raw bit streams: the
    counters, the bytes handed to serial_write() and to ipc_write() and the stdout lines (with the host's skip_type[5]
    set) against the oracle's bit machine, which tests/test_oracle_fsm_bits.py pins to the reference's protodec_decode()"""
    import ctypes as C
    from test_oracle_fsm_bits import make_bits
    bits = make_bits(21, 700)
    (tmp_path / "bits.raw").write_bytes(bits.tobytes())
    exe = _build_protodec(tmp_path)
    r = subprocess.run([str(exe), str(tmp_path / "bits.raw"), str(tmp_path / "out"), str(reset_at)], capture_output=True, text=True,
                       timeout=300, env={"GAIS_SHIM_BATCH_BITS": "9600", "PATH": "/usr/bin:/bin"})
    assert r.returncode == 0, r.stderr
    port = O.port()
    port.lib.goracle_fsm_bits.argtypes = [C.c_void_p, C.c_int64, C.c_void_p, C.c_void_p, C.c_int64, C.c_void_p]
    port.lib.goracle_nmea.argtypes = [C.c_void_p, C.c_int, C.c_void_p, C.c_char_p]
    parts = [bits] if reset_at < 0 else [bits[:reset_at], bits[reset_at:]]      # protodec_reset() = a fresh bit machine, same counters
    cnt, text, seq, buf, lines = [0, 0, 0], b"", C.c_uint8(0), C.create_string_buffer(512), []
    for part in parts:
        part = np.ascontiguousarray(part)
        ps, nfr = (C.c_int32 * 3)(), C.c_int64()
        frames = np.zeros(len(part) // 20 + 64, dtype=O.FRAME_DTYPE)
        assert port.lib.goracle_fsm_bits(part.ctypes.data, len(part), ps, frames.ctypes.data, len(frames), C.byref(nfr)) == 0
        cnt = [a + b for a, b in zip(cnt, ps)]
        for fr in frames[: nfr.value]:
            if fr["status"] != 0:
                continue
            m = np.zeros((), dtype=MSG_DTYPE)
            nb = int(fr["nbytes"])
            m["payload"][:nb] = fr["payload"][:nb]
            m["nbits"] = fr["nbits"]
            kind = int(fr["payload"][0]) >> 2
            m["flags"] = seq.value | (16 if 1 <= kind <= 24 else 0)
            pl = np.ascontiguousarray(fr["payload"])
            n = port.lib.goracle_nmea(pl.ctypes.data, int(fr["nbits"]), C.byref(seq), buf)
            text += buf.raw[:n]
            if 1 <= kind <= 24 and kind != 5:
                lines.append(text_format(m, "B").decode().rstrip("\n"))
    out = r.stdout.strip().splitlines()
    assert out[-1].startswith(f"Received correctly: {cnt[0]} packets, wrong CRC: {cnt[1]} packets, wrong size: {cnt[2]} packets, "
                              f"seqnr {seq.value},")
    assert cnt[0] > 50 and cnt[1] > 50 and cnt[2] > 20
    assert (tmp_path / "out.serial").read_bytes() == text
    assert (tmp_path / "out.ipc").read_bytes() == text.replace(b"\r\n", b"\n")       # "!%s" without the CR LF, src/protodec.c:886-888
    assert [ln for ln in out if ln.startswith("ch B type ")] == lines and not any(" type 5 " in ln for ln in out)
