"""Prints the Q14 Gaussian BT=0.4 symbol pulse used by gnuais_b200/csrc/synth_core.h (gs_pulse)."""
import numpy as np

sps, BT = 5, 0.4
t = (np.arange(31) - 15) / sps
h = np.exp(-2 * np.pi ** 2 * BT ** 2 * t ** 2 / np.log(2))
h /= h.sum()
p = np.convolve(h, np.ones(sps))
q = np.round(p * (1 << 14)).astype(int)
print(q[8:27].tolist())
