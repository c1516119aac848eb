import subprocess
import sys
from pathlib import Path

import pytest

ROOT = Path(__file__).resolve().parent.parent
sys.path.insert(0, str(ROOT))
sys.path.insert(0, str(ROOT / "tests"))


def pytest_configure(config):
    config.addinivalue_line("markers", "gpu: needs a real B200 (run with -m gpu on the GPU box)")
    # the product library and the checkers are built in-tree; build them if a fresh clone lacks them
    lib = ROOT / "gnuais_b200" / "lib" / "libgaisb200.so"
    if not lib.exists():
        subprocess.run(["make", "-C", str(ROOT / "gnuais_b200" / "csrc")], check=True, stdout=subprocess.DEVNULL)
    if not (ROOT / "oracle" / "_build" / "libgais_oracle.so").exists():
        subprocess.run(["make", "-C", str(ROOT / "oracle"), "port"], check=True, stdout=subprocess.DEVNULL)
    if Path("/root/reference/src/receiver.c").exists() and not (ROOT / "oracle" / "_ref" / "libgnuais_ref_tap.so").exists():
        subprocess.run(["make", "-C", str(ROOT / "oracle"), "ref"], check=True, stdout=subprocess.DEVNULL)


def has_gpu() -> bool:
    try:
        import torch
        return torch.cuda.is_available()
    except Exception:
        return False
