"""The HDLC tables of the tracking kernel, checked on the host (no GPU): tests/c/hdlc_table_check.cu includes
gnuais_b200/csrc/gais_track.cuh and, on 1.3 Mbit of flags, well-formed frames (CRC-16), stuffed random payloads
and noise,
  * steps the FSM four bits at a time through hdlc_nibble_entry() -- reading the fields where the kernel reads
    them -- against bit-by-bit hdlc_transition(), from every one of the 80 states: same states, same stored
    bits, same frames;
  * runs the table FSM with the kernel's buffer rules (reset at 449 stored bits, frame judged by stop bit, length
    and CRC) against the oracle's bit machine (oracle/gais_oracle.c fsm_bit(), itself pinned to the reference):
    same frame events, same ok / crcfail / sizefail counts."""
import shutil
import subprocess
from pathlib import Path

import pytest

import oracle_lib as O

ROOT = Path(__file__).resolve().parent.parent


def test_hdlc_tables_equal_bit_fsm_and_oracle(tmp_path):
    nvcc = shutil.which("nvcc") or "/usr/local/cuda/bin/nvcc"
    if not Path(nvcc).exists():
        pytest.skip("nvcc not available")
    O.port()                                   # builds oracle/_build/libgais_oracle.so if it is not there yet
    lib_dir = ROOT / "oracle" / "_build"
    exe = tmp_path / "hdlc_table_check"
    subprocess.run([nvcc, "-O1", "-std=c++17", f"-I{ROOT / 'include'}", f"-I{ROOT / 'gnuais_b200' / 'csrc'}", f"-I{ROOT / 'oracle'}",
                    "-o", str(exe), str(ROOT / "tests" / "c" / "hdlc_table_check.cu"), f"-L{lib_dir}", "-lgais_oracle",
                    "-Xlinker", "-rpath", "-Xlinker", str(lib_dir)], check=True, capture_output=True)
    out = subprocess.run([str(exe)], capture_output=True, text=True)
    assert out.returncode == 0, out.stdout + out.stderr
    lines = out.stdout.strip().splitlines()
    assert len(lines) == 2 and all(l.startswith("ok:") for l in lines), out.stdout
