"""Deterministic test inputs shared by the golden-fixture generator and the parity tests.

Every case is (samples, num_ch) with samples int16 [n_frames, num_ch] frame-interleaved,
exactly the buffer shape receiver_run() takes (src/receiver.c:87-102)."""
from __future__ import annotations

import numpy as np

from gnuais_b200 import SynthParams, synth_host


def synth_case(seed: int, n_channels: int, n_frames: int, sigma: float = 300.0, rho: float = 0.5, jitter: bool = True,
               first_channel: int = 0) -> np.ndarray:
    p = SynthParams(seed=seed, sigma=sigma, rho=rho, jitter=jitter)
    return synth_host(p, n_channels, n_frames, first_channel=first_channel, layout="interleaved")


def _rng(seed):
    return np.random.default_rng(seed)


def edge_cases() -> dict:
    """name -> int16 [n_frames, 1]; the corner inputs of SURVEY.md H1/H6."""
    out = {}
    out["silence"] = np.zeros((5000, 1), np.int16)
    x = np.zeros((400, 1), np.int16); x[100] = 1000; x[200] = -1; x[300] = 32767
    out["impulses_denormal_taps"] = x                       # single samples under the denormal taps 2/33
    t = np.arange(20000)
    out["square_fullscale"] = np.where((t // 5) % 2 == 0, 32767, -32768).astype(np.int16)[:, None]
    out["alternating_1"] = np.where(t % 2 == 0, 1, -1).astype(np.int16)[:, None]
    out["dc_positive"] = np.full((3000, 1), 12345, np.int16)
    out["dc_negative"] = np.full((3000, 1), -32768, np.int16)
    out["noise_small"] = _rng(1).integers(-3, 4, size=(30000, 1)).astype(np.int16)
    out["noise_full"] = _rng(2).integers(-32768, 32768, size=(30000, 1)).astype(np.int16)
    # exact cancellation candidates: antisymmetric pairs around the two centre taps
    x = np.zeros((600, 1), np.int16)
    for k in range(10):
        x[50 + 50 * k] = 1000 + k; x[51 + 50 * k] = -(1000 + k)
    out["cancel_pairs"] = x
    out["one_sample"] = np.array([[1234]], np.int16)
    out["len_31"] = _rng(3).integers(-2000, 2000, size=(31, 1)).astype(np.int16)
    out["len_33"] = _rng(4).integers(-2000, 2000, size=(33, 1)).astype(np.int16)
    out["len_1025"] = _rng(5).integers(-2000, 2000, size=(1025, 1)).astype(np.int16)
    return out


GOLDEN_SYNTH = {
    # name: (seed, n_channels, n_frames, sigma, rho, jitter)
    "synth_s300": (11, 2, 96000, 300.0, 0.5, True),
    "synth_s1500": (12, 2, 96000, 1500.0, 0.7, True),
    "synth_dense_nojit": (13, 1, 96000, 300.0, 1.0, False),
}


def golden_inputs() -> dict:
    """name -> (samples [n_frames, num_ch], num_ch)"""
    d = {}
    for name, (seed, nch, nfr, sigma, rho, jit) in GOLDEN_SYNTH.items():
        d[name] = (synth_case(seed, nch, nfr, sigma, rho, jit), nch)
    for name, x in edge_cases().items():
        d["edge_" + name] = (x, 1)
    return d
