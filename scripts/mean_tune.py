#!/usr/bin/env python
"""Time psa_mean_positions alone: scripts/mean_tune.py n_t n_a"""
import os
import sys
from pathlib import Path

import torch

sys.path.insert(0, str(Path(__file__).resolve().parents[1]))
from psa_b200.engine import Engine  # noqa: E402

n_t, n_a = int(sys.argv[1]), int(sys.argv[2])
eng = Engine()
pos = torch.randn((n_t, n_a, 3), device=eng.device)
for _ in range(3):
    eng.mean_positions(pos)
torch.cuda.synchronize()
a, b = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
a.record()
for _ in range(10):
    eng.mean_positions(pos)
b.record()
torch.cuda.synchronize()
ms = a.elapsed_time(b) / 10
print(f"lib={os.path.basename(os.environ.get('PSA_B200_LIB', 'default'))} n_t={n_t} n_a={n_a}: {ms:.3f} ms  "
      f"{12.0 * n_t * n_a / ms / 1e6:.0f} GB/s")
