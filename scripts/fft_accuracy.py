#!/usr/bin/env python
"""rms error of the device FFT vs float64, next to NumPy's own complex64 FFT on the same input."""
import sys
from pathlib import Path

import numpy as np
import torch

sys.path.insert(0, str(Path(__file__).resolve().parents[1]))
from psa_b200.engine import Engine  # noqa: E402

eng = Engine()
for n_t in (250, 512, 2048, 8192, 16384, 32768):
    rng = np.random.default_rng(n_t)
    n_k = 4
    P = rng.standard_normal((2 * n_k, 3, n_t)).astype(np.float32)
    # a strong line on top of the noise, like a phonon peak
    t = np.arange(n_t)
    P[0::2] += 30 * np.cos(2 * np.pi * 37 * t / n_t).astype(np.float32)
    P[1::2] += 30 * np.sin(2 * np.pi * 37 * t / n_t).astype(np.float32)
    ldp = -(-n_t // 4) * 4
    Pp = np.zeros((2 * n_k, 3, ldp), np.float32)
    Pp[:, :, :n_t] = P
    out = torch.zeros((n_t, n_k, 3), dtype=torch.complex64, device=eng.device)
    eng.fft_sed(torch.from_numpy(Pp).to(eng.device), 1, Pp.size, n_k, n_t, ldp, 0, out, n_k, 0)
    z = P[0::2].astype(np.float64) + 1j * P[1::2].astype(np.float64)
    want = (np.fft.fft(z, axis=-1) / n_t).transpose(2, 0, 1)
    z32 = (P[0::2] + 1j * P[1::2]).astype(np.complex64)
    np32 = (np.fft.fft(z32, axis=-1) / n_t).astype(np.complex64).transpose(2, 0, 1)
    rms = np.sqrt(np.mean(np.abs(want) ** 2))
    e_dev = np.sqrt(np.mean(np.abs(out.cpu().numpy() - want) ** 2)) / rms
    e_np = np.sqrt(np.mean(np.abs(np32 - want) ** 2)) / rms
    print(f"n_t={n_t:6d}  device rms err {e_dev:.2e}   numpy complex64 rms err {e_np:.2e}   ratio {e_dev / e_np:.2f}")
