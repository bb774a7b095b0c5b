#!/usr/bin/env python
"""Time psa_fft_sed alone (CUDA events) for one shape.
    python scripts/fft_tune.py 16384 1024 [mode]        PSA_FFT4=0 selects the one-CTA-per-column kernel
"""
import os
import sys
from pathlib import Path

import torch

sys.path.insert(0, str(Path(__file__).resolve().parents[1]))
from psa_b200.engine import Engine  # noqa: E402

n_t, n_k = int(sys.argv[1]), int(sys.argv[2])
mode = int(sys.argv[3]) if len(sys.argv) > 3 else 0
eng = Engine()
P = torch.randn((2 * n_k, 3, n_t), device=eng.device)
out = torch.empty((n_t, n_k, 3), dtype=torch.complex64, device=eng.device) if mode == 0 else \
    torch.empty((n_t, n_k), dtype=torch.float32, device=eng.device)
for _ in range(3):
    eng.fft_sed(P, 1, P.numel(), n_k, n_t, n_t, mode, out, n_k, 0)
torch.cuda.synchronize()
a, b = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
reps = 10
a.record()
for _ in range(reps):
    eng.fft_sed(P, 1, P.numel(), n_k, n_t, n_t, mode, out, n_k, 0)
b.record()
torch.cuda.synchronize()
ms = a.elapsed_time(b) / reps
gbs = (48.0 if mode == 0 else 28.0) * n_k * n_t / ms / 1e6
print(f"fft4={os.environ.get('PSA_FFT4', '1')} n_t={n_t} n_k={n_k} mode={mode}: {ms:.3f} ms  {gbs:.0f} GB/s algorithmic "
      f"({gbs / 6548.8:.3f} of the measured HBM peak)")
