"""CPU-side checks of the C-ABI boundary: the shared object loads, exports every symbol
include/gais_b200.h declares, struct layouts agree, the host NMEA formatter equals the oracle's,
and -- with no GPU -- creating a context FAILS LOUDLY (there is no CPU fallback)."""
import ctypes as C
import re
from pathlib import Path

import numpy as np
import pytest

import oracle_lib as O
from conftest import has_gpu
from gnuais_b200 import _lib as L
from gnuais_b200 import MSG_DTYPE, nmea_format

ROOT = Path(__file__).resolve().parent.parent


def _declared(header: str):
    text = (ROOT / "include" / header).read_text()
    text = re.sub(r"/\*.*?\*/", "", text, flags=re.S)
    return sorted(set(re.findall(r"\b(gais_[a-z_0-9]+)\s*\(", text)))


def test_every_declared_symbol_is_exported_and_bound():
    lib = L.load()
    names = _declared("gais_b200.h")
    assert len(names) >= 20
    for n in names:
        assert hasattr(lib, n), f"{n} declared in include/gais_b200.h but not exported"
    assert set(names) == set(L.SYMBOLS), set(names) ^ set(L.SYMBOLS)
    assert lib.gais_abi_version() == L.ABI_VERSION


def test_struct_layouts():
    assert C.sizeof(L.Msg) == 64 and L.Msg.flags.offset == 53 and L.Msg.nbits.offset == 54
    assert L.Msg.channel.offset == 56 and L.Msg.end_bit.offset == 60
    assert C.sizeof(L.Config) == 64 and C.sizeof(L.Counters) == 12 and C.sizeof(L.ChanState) == 24
    assert C.sizeof(L.NmeaRec) == 176 and C.sizeof(L.Synth) == 24 and C.sizeof(L.Timing) == 24


@pytest.mark.skipif(has_gpu(), reason="checks the no-GPU failure mode")
def test_create_fails_loudly_without_gpu():
    lib = L.load()
    assert lib.gais_device_count() == 0
    cfg = L.Config(abi_version=L.ABI_VERSION, device=0, n_channels=4, layout=0, max_frames_per_run=1024)
    ctx = C.c_void_p()
    rc = lib.gais_create(C.byref(cfg), C.byref(ctx))
    assert rc == -2 and not ctx.value                      # GAIS_ENODEV
    assert b"no CPU fallback" in lib.gais_last_error()
    from gnuais_b200 import BatchReceiver, GaisError
    with pytest.raises(GaisError):
        BatchReceiver(4, 1024)


def test_bad_arguments_rejected():
    lib = L.load()
    ctx = C.c_void_p()
    cfg = L.Config(abi_version=99, device=0, n_channels=4, layout=0, max_frames_per_run=1024)
    assert lib.gais_create(C.byref(cfg), C.byref(ctx)) == -1
    cfg = L.Config(abi_version=L.ABI_VERSION, device=0, n_channels=0, layout=0, max_frames_per_run=1024)
    assert lib.gais_create(C.byref(cfg), C.byref(ctx)) == -1
    cfg = L.Config(abi_version=L.ABI_VERSION, device=0, n_channels=1, layout=7, max_frames_per_run=1024)
    assert lib.gais_create(C.byref(cfg), C.byref(ctx)) == -1
    assert lib.gais_sync(None) == -1


@pytest.mark.skipif(not Path("/root/reference/src/receiver.h").exists(), reason="needs the reference's headers")
def test_compat_structs_and_prototypes_match_the_reference_headers(tmp_path):
    """tests/c/layout_check.c includes the reference's own receiver.h / protodec.h (struct tags renamed) next to
    include/gais_compat.h: sizeof / offsetof of every field of struct receiver and struct demod_state_t, and the
    prototypes of the seven entry points, are compile-time assertions"""
    import subprocess
    (tmp_path / "config.h").write_text('#define HAVE_ALSA 1\n#define PACKAGE "gnuais"\n#define VERSION "0.3.3"\n')
    r = subprocess.run(["gcc", "-fsyntax-only", "-Wall", "-Werror", "-fcommon", "-I", str(tmp_path), "-I", "/root/reference/src",
                        "-I", str(ROOT / "include"), str(ROOT / "tests" / "c" / "layout_check.c")], capture_output=True, text=True)
    assert r.returncode == 0, r.stderr


def _oracle_nmea(payload: bytes, nbits: int, seqnr: int):
    lib = O.port().lib
    lib.goracle_nmea.restype = C.c_int
    lib.goracle_nmea.argtypes = [C.c_char_p, C.c_int, C.POINTER(C.c_uint8), C.c_char_p]
    s = C.c_uint8(seqnr)
    out = C.create_string_buffer(512)
    n = lib.goracle_nmea(payload, nbits, C.byref(s), out)
    return out.raw[:n], s.value


def test_host_nmea_formatter_matches_oracle():
    rng = np.random.default_rng(5)
    for i in range(3000):
        nbits = int(rng.integers(1, 427))
        nb = nbits // 8
        pay = bytearray(rng.integers(0, 256, size=53, dtype=np.uint8).tobytes())
        if i % 3 == 0:
            pay[0] = (int(rng.integers(0, 28)) << 2) | int(rng.integers(0, 4))   # bias towards gated/ungated types
        for j in range(nb, 53):
            pay[j] = 0
        seq = int(rng.integers(0, 10))
        m = np.zeros((), dtype=MSG_DTYPE)
        m["payload"] = np.frombuffer(bytes(pay), np.uint8)
        m["nbits"], m["flags"] = nbits, seq
        want, _ = _oracle_nmea(bytes(pay[:max(nb, 1)]) + b"\0" * 60, nbits, seq)
        assert nmea_format(m) == want, (nbits, seq)


def test_host_synth_is_deterministic_and_layout_consistent():
    from gnuais_b200 import SynthParams, synth_host
    p = SynthParams(seed=99, sigma=300)
    a = synth_host(p, 3, 6000)
    b = synth_host(p, 3, 6000, layout="interleaved")
    assert np.array_equal(a, b.T)
    # channels are keyed by absolute index; prefixes of longer runs agree
    c = synth_host(p, 1, 3000, first_channel=2)
    assert np.array_equal(c[0], a[2, :3000])
    assert a.std() > 1000 and abs(int(a.max())) < 20000


def test_host_alloc_fails_cleanly_without_a_gpu():
    """gais_host_alloc() is page-locked memory from the CUDA runtime: on a box without a device it must
    return an error code (and a NULL pointer), never crash; gais_host_free(NULL) is a no-op"""
    import torch
    if torch.cuda.is_available():
        import pytest
        pytest.skip("a GPU is visible: the allocation succeeds (covered by the -m gpu tests)")
    lib = L.load()
    p = C.c_void_p(123)
    assert lib.gais_host_alloc(None, 64) != 0
    rc = lib.gais_host_alloc(C.byref(p), 1 << 20)
    assert rc != 0 and not p.value
    assert b"cudaHostAlloc" in lib.gais_last_error()
    lib.gais_host_free(None)
