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


def test_shim_links_without_gpu(tmp_path):
    """CPU: the shim library resolves against the batched library and exports the three entry points"""
    _build(tmp_path)
    out = subprocess.run(["nm", "-D", "--defined-only", str(LIBDIR / "libgnuais_rx_b200.so")], capture_output=True, text=True).stdout
    for sym in ("init_receiver", "receiver_run", "free_receiver", "gais_compat_flush"):
        assert f" T {sym}" in out


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
