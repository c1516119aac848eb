"""gnuais_b200 -- B200-native batched AIS receive path (gnuais receiver_run()/protodec_decode()).

Only what the hot path needs lives here: ``csrc/`` (CUDA kernels + the C-ABI), ``lib/`` (the
built shared objects, git-ignored) and the ctypes / host mirror of the reference interface.
"""
from ._lib import GaisError, load  # noqa: F401
from .receiver import (  # noqa: F401
    BatchReceiver, MSG_DTYPE, free_receiver, init_receiver, nmea_format, receiver_run, text_format,
)
from .synth import SynthParams, synth_device, synth_host  # noqa: F401

__all__ = [
    "BatchReceiver", "GaisError", "MSG_DTYPE", "SynthParams", "free_receiver", "init_receiver", "load",
    "nmea_format", "receiver_run", "synth_device", "synth_host", "text_format",
]
