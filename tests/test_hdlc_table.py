"""The HDLC tables of the tracking kernel, checked on the host (no GPU): tests/c/hdlc_table_check.cu includes
gnuais_b200/csrc/gais_track.cuh and steps the FSM four bits at a time through hdlc_nibble_entry() -- reading
the fields where the kernel reads them -- against bit-by-bit hdlc_transition() on 1.2 Mbit of flags, stuffed
payloads and noise, from every one of the 80 states: same states, same stored bits, same frames."""
import shutil
import subprocess
from pathlib import Path

import pytest

ROOT = Path(__file__).resolve().parent.parent


def test_nibble_table_equals_bit_fsm(tmp_path):
    nvcc = shutil.which("nvcc") or "/usr/local/cuda/bin/nvcc"
    if not Path(nvcc).exists():
        pytest.skip("nvcc not available")
    exe = tmp_path / "hdlc_table_check"
    subprocess.run([nvcc, "-O1", "-std=c++17", f"-I{ROOT / 'include'}", f"-I{ROOT / 'gnuais_b200' / 'csrc'}", "-o", str(exe),
                    str(ROOT / "tests" / "c" / "hdlc_table_check.cu")], check=True, capture_output=True)
    out = subprocess.run([str(exe)], capture_output=True, text=True)
    assert out.returncode == 0, out.stdout + out.stderr
    assert out.stdout.startswith("ok:")
