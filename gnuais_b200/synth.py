"""Synthetic GMSK/AIS discriminator audio: ctypes front-end of the integer-only
generator in csrc/synth_core.h.  Host and device versions produce identical int16 samples."""
from __future__ import annotations

import ctypes as C
from dataclasses import dataclass

import numpy as np

from . import _lib as L


def sigma_to_q16(sigma: float) -> int:
    """noise_q16 for a target noise sigma (Irwin-Hall of four u16: sigma_IH = 65536/sqrt(3))."""
    return int(round(sigma * 65536.0 / 37837.227))


@dataclass
class SynthParams:
    seed: int = 1
    amplitude: int = 12000
    sigma: float = 300.0
    rho: float = 0.5
    jitter: bool = True

    def c_struct(self) -> L.Synth:
        s = L.Synth()
        s.seed = self.seed & 0xFFFFFFFFFFFFFFFF
        s.amplitude = self.amplitude
        s.noise_q16 = sigma_to_q16(self.sigma)
        s.rho_q16 = int(round(self.rho * 65536))
        s.jitter = 1 if self.jitter else 0
        return s


_synth_lib = None


def _host_lib():
    """libgais_synth.so: the generator alone (plain C).  Loading it maps nothing of the product, so the reference arm
    of bench.py and the CPU tests can synthesise audio without touching libgaisb200.so."""
    global _synth_lib
    if _synth_lib is None:
        path = L.LIB_PATH.parent / "libgais_synth.so"
        if not path.exists():
            raise ImportError(f"{path} is missing: run __graft_entry__.build()")
        lib = C.CDLL(str(path))
        lib.gais_synth_host.restype = C.c_int
        lib.gais_synth_host.argtypes = [C.POINTER(L.Synth), C.c_uint32, C.c_int32, C.c_int64, C.c_void_p, C.c_int32, C.c_int64]
        _synth_lib = lib
    return _synth_lib


def synth_host(p: SynthParams, n_channels: int, n_frames: int, first_channel: int = 0, layout: str = "planar",
               stride: int | None = None) -> np.ndarray:
    """int16 array, planar [n_channels, n_frames] or interleaved [n_frames, n_channels]; no GPU needed."""
    lib = _host_lib()
    planar = layout == "planar"
    if stride is None:
        stride = n_frames if planar else n_channels
    out = np.zeros((n_channels, stride) if planar else (n_frames, stride), dtype=np.int16)
    s = p.c_struct()
    rc = lib.gais_synth_host(C.byref(s), first_channel, n_channels, n_frames, out.ctypes.data_as(C.c_void_p),
                             L.LAYOUT_PLANAR if planar else L.LAYOUT_INTERLEAVED, stride)
    if rc != 0:
        raise ValueError(f"gais_synth_host failed: {rc}")
    return out


def synth_device(p: SynthParams, out, n_channels: int, n_frames: int, first_channel: int = 0, layout: str = "planar",
                 stride: int | None = None, stream=None) -> None:
    """Fill a CUDA int16 torch tensor in place (async on the current torch stream)."""
    import torch

    lib = L.load()
    planar = layout == "planar"
    if stride is None:
        stride = out.stride(0)
    if stream is None:
        stream = torch.cuda.current_stream(out.device).cuda_stream
    s = p.c_struct()
    L.check(lib.gais_synth_device(C.byref(s), first_channel, n_channels, n_frames, C.c_void_p(out.data_ptr()),
                                  L.LAYOUT_PLANAR if planar else L.LAYOUT_INTERLEAVED, stride, C.c_void_p(stream)))
